// TEST INFRASTRUCTURE ONLY -- the handful of googletest macros the reference's tests/*.cc use (TEST, ASSERT_/EXPECT_ EQ, NEAR,
// TRUE, FALSE, GT, GE, LT, LE, NE, THROW, NO_THROW, with `<< message` streaming), so that those files compile where they lie
// under /root/reference/tests and run as oracle/_ref/reference_tests.  googletest is not in the image.  This is NOT googletest.
#pragma once

#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing
{

struct TestInfo
{
    const char* suite;
    const char* name;
    std::function<void()> body;
};

inline std::vector<TestInfo>& registry()
{
    static std::vector<TestInfo> r;
    return r;
}

inline int& current_failures()
{
    static int n = 0;
    return n;
}

struct Registrar
{
    Registrar(const char* suite, const char* name, std::function<void()> body) { registry().push_back({suite, name, std::move(body)}); }
};

// collects the streamed message and reports when it dies
class Failure
{
    std::ostringstream msg_;
    bool active_;

public:
    Failure(bool failed, const char* file, int line, const std::string& what) : active_(failed)
    {
        if (active_) msg_ << file << ":" << line << ": Failure\n  " << what << "\n  ";
    }
    Failure(Failure&& o) : msg_(std::move(o.msg_)), active_(o.active_) { o.active_ = false; }
    ~Failure()
    {
        if (active_)
        {
            ++current_failures();
            std::cout << msg_.str() << std::endl;
        }
    }
    template <typename T>
    Failure& operator<<(const T& t)
    {
        if (active_) msg_ << t;
        return *this;
    }
    Failure& operator<<(std::ostream& (*f)(std::ostream&))
    {
        if (active_) msg_ << f;
        return *this;
    }
    explicit operator bool() const { return active_; }
};

// `return Voidify() = failure << "msg"` gives ASSERT_* its early return while keeping the stream syntax
struct Voidify
{
    void operator=(const Failure&) const {}
};

template <typename A, typename B>
std::string describe(const char* op, const char* ea, const char* eb, const A& a, const B& b)
{
    std::ostringstream s;
    s << "Expected: (" << ea << ") " << op << " (" << eb << "), actual: " << a << " vs " << b;
    return s.str();
}

inline void InitGoogleTest(int*, char**) {}

inline int run_all(const char* filter = nullptr)
{
    int failed_tests = 0, ran = 0;
    std::vector<std::string> failed_names;
    for (auto& t : registry())
    {
        const std::string full = std::string(t.suite) + "." + t.name;
        if (filter && full.find(filter) == std::string::npos) continue;
        std::cout << "[ RUN      ] " << full << std::endl;
        const int before = current_failures();
        try
        {
            t.body();
        }
        catch (const std::exception& e)
        {
            ++current_failures();
            std::cout << "  uncaught exception: " << e.what() << std::endl;
        }
        catch (...)
        {
            ++current_failures();
            std::cout << "  uncaught exception" << std::endl;
        }
        ++ran;
        if (current_failures() != before)
        {
            ++failed_tests;
            failed_names.push_back(full);
            std::cout << "[  FAILED  ] " << full << std::endl;
        }
        else
            std::cout << "[       OK ] " << full << std::endl;
    }
    std::cout << "[==========] " << ran << " tests ran, " << (ran - failed_tests) << " passed, " << failed_tests << " failed" << std::endl;
    for (auto& n : failed_names) std::cout << "[  FAILED  ] " << n << std::endl;
    return failed_tests == 0 ? 0 : 1;
}

}  // namespace testing

#define RUN_ALL_TESTS() ::testing::run_all()

#define GTEST_SHIM_CAT_(a, b) a##b
#define GTEST_SHIM_CAT(a, b) GTEST_SHIM_CAT_(a, b)

#define TEST(suite, name)                                                                                            \
    static void GTEST_SHIM_CAT(suite##_##name##_body_, __LINE__)();                                                  \
    static ::testing::Registrar GTEST_SHIM_CAT(suite##_##name##_reg_, __LINE__)(#suite, #name,                       \
                                                                               &GTEST_SHIM_CAT(suite##_##name##_body_, __LINE__)); \
    static void GTEST_SHIM_CAT(suite##_##name##_body_, __LINE__)()

#define GTEST_SHIM_CHECK_(fatal, cond, what)                                         \
    if (::testing::Failure gtest_shim_f_{!(cond), __FILE__, __LINE__, what}; !gtest_shim_f_) \
        ;                                                                            \
    else                                                                             \
        GTEST_SHIM_##fatal ::testing::Voidify() = gtest_shim_f_

#define GTEST_SHIM_FATAL return
#define GTEST_SHIM_NONFATAL

#define GTEST_SHIM_CMP_(fatal, op, a, b)                                                                  \
    if (::testing::Failure gtest_shim_f_{!((a)op(b)), __FILE__, __LINE__, std::string(#a " " #op " " #b)}; !gtest_shim_f_) \
        ;                                                                                                 \
    else                                                                                                  \
        GTEST_SHIM_##fatal ::testing::Voidify() = gtest_shim_f_

#define ASSERT_TRUE(c) GTEST_SHIM_CHECK_(FATAL, (c), "Value of: " #c " expected true")
#define ASSERT_FALSE(c) GTEST_SHIM_CHECK_(FATAL, !(c), "Value of: " #c " expected false")
#define EXPECT_TRUE(c) GTEST_SHIM_CHECK_(NONFATAL, (c), "Value of: " #c " expected true")
#define EXPECT_FALSE(c) GTEST_SHIM_CHECK_(NONFATAL, !(c), "Value of: " #c " expected false")
#define ASSERT_EQ(a, b) GTEST_SHIM_CMP_(FATAL, ==, a, b)
#define ASSERT_NE(a, b) GTEST_SHIM_CMP_(FATAL, !=, a, b)
#define ASSERT_GT(a, b) GTEST_SHIM_CMP_(FATAL, >, a, b)
#define ASSERT_GE(a, b) GTEST_SHIM_CMP_(FATAL, >=, a, b)
#define ASSERT_LT(a, b) GTEST_SHIM_CMP_(FATAL, <, a, b)
#define ASSERT_LE(a, b) GTEST_SHIM_CMP_(FATAL, <=, a, b)
#define EXPECT_EQ(a, b) GTEST_SHIM_CMP_(NONFATAL, ==, a, b)
#define EXPECT_NE(a, b) GTEST_SHIM_CMP_(NONFATAL, !=, a, b)
#define EXPECT_GT(a, b) GTEST_SHIM_CMP_(NONFATAL, >, a, b)
#define EXPECT_LT(a, b) GTEST_SHIM_CMP_(NONFATAL, <, a, b)
#define ASSERT_NEAR(a, b, tol) GTEST_SHIM_CHECK_(FATAL, std::abs((double)(a) - (double)(b)) <= (double)(tol), "|" #a " - " #b "| <= " #tol)
#define EXPECT_NEAR(a, b, tol) GTEST_SHIM_CHECK_(NONFATAL, std::abs((double)(a) - (double)(b)) <= (double)(tol), "|" #a " - " #b "| <= " #tol)

#define SUCCEED() GTEST_SHIM_CHECK_(NONFATAL, true, "")
#define FAIL() GTEST_SHIM_CHECK_(FATAL, false, "Failed")
#define ADD_FAILURE() GTEST_SHIM_CHECK_(NONFATAL, false, "Failed")

#define GTEST_SHIM_THROW_(fatal, stmt, extype)                                                       \
    if (::testing::Failure gtest_shim_f_{[&]() {                                                     \
            try { stmt; } catch (const extype&) { return false; } catch (...) { return true; }       \
            return true; }(), __FILE__, __LINE__, "Expected: " #stmt " throws " #extype}; !gtest_shim_f_) \
        ;                                                                                            \
    else                                                                                             \
        GTEST_SHIM_##fatal ::testing::Voidify() = gtest_shim_f_
#define ASSERT_THROW(stmt, extype) GTEST_SHIM_THROW_(FATAL, stmt, extype)
#define EXPECT_THROW(stmt, extype) GTEST_SHIM_THROW_(NONFATAL, stmt, extype)
#define GTEST_SHIM_NOTHROW_(fatal, stmt)                                                             \
    if (::testing::Failure gtest_shim_f_{[&]() {                                                     \
            try { stmt; } catch (...) { return true; }                                               \
            return false; }(), __FILE__, __LINE__, "Expected: " #stmt " does not throw"}; !gtest_shim_f_) \
        ;                                                                                            \
    else                                                                                             \
        GTEST_SHIM_##fatal ::testing::Voidify() = gtest_shim_f_
#define ASSERT_NO_THROW(stmt) GTEST_SHIM_NOTHROW_(FATAL, stmt)
#define EXPECT_NO_THROW(stmt) GTEST_SHIM_NOTHROW_(NONFATAL, stmt)
