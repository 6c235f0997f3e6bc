// TEST INFRASTRUCTURE ONLY -- entry point of oracle/_ref/reference_tests: the reference's own tests/*.cc (compiled where they lie)
// over oracle/eigen_shim and oracle/gtest_shim.
//   reference_tests [--skip=Suite.Name,Suite.Name,...] [--list] [substring filter on "Suite.Name"]
#include <gtest/gtest.h>

#include <cstring>
#include <set>

int main(int argc, char** argv)
{
    std::set<std::string> skip;
    const char* filter = nullptr;
    bool list = false;
    for (int i = 1; i < argc; ++i)
    {
        if (std::strncmp(argv[i], "--skip=", 7) == 0)
        {
            std::stringstream ss(argv[i] + 7);
            std::string name;
            while (std::getline(ss, name, ',')) skip.insert(name);
        }
        else if (std::strcmp(argv[i], "--list") == 0)
            list = true;
        else
            filter = argv[i];
    }
    auto& reg = ::testing::registry();
    if (list)
    {
        for (auto& t : reg) std::cout << t.suite << "." << t.name << "\n";
        return 0;
    }
    for (auto it = reg.begin(); it != reg.end();)
    {
        const std::string full = std::string(it->suite) + "." + it->name;
        if (skip.count(full))
        {
            std::cout << "[ SKIPPED  ] " << full << std::endl;
            it = reg.erase(it);
        }
        else
            ++it;
    }
    return ::testing::run_all(filter);
}
