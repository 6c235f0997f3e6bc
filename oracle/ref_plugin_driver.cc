// TEST INFRASTRUCTURE ONLY -- the drop-in, exercised from the reference's side.
//
// This translation unit is reference code plus the binding of include/reference_binding/B200ObjectiveTerm.hh: the UNMODIFIED
// TinyAD::ScalarFunction of /root/reference (compiled in place over oracle/eigen_shim, like ref_driver.cc) gets a
// B200ScalarObjectiveTerm pushed onto its own `objective_terms` (ScalarFunction.hh:235) and is then evaluated through its own
// eval / eval_with_gradient / eval_with_derivatives / eval_with_hessian_proj (Detail/ScalarFunctionImpl.hh:256-416).  The device
// terms come from the product's energies library (tadx_* of tinyad_b200/csrc/energies.cu, the "user .cu translation unit"),
// reached through the C ABI only.  tests/test_reference_plugin_gpu.py compares the result with the same ScalarFunction holding the
// reference's own CPU terms (ref_driver.cc).  Built by `make -C oracle _ref/libtinyad_plugin.so` into oracle/_ref/.
#include <TinyAD/ScalarFunction.hh>
#include <TinyAD/Utils/Helpers.hh>

#include <reference_binding/B200ObjectiveTerm.hh>

#include <cstdint>
#include <memory>
#include <string>

extern "C" {
// tinyad_b200/csrc/energies.cu
const char* tadx_last_error();
int tadx_create(int d, int64_t n_vertices, int is_vector, int device, int assembly, void** out);
void tadx_destroy(void* h);
void* tadx_handle(void* h);
int tadx_add_term(void* h, int kind, int64_t n_elements, const int32_t* conn, int valence, const double* data, int n_data);

struct plugin_term
{
    int kind;
    std::int64_t n_elements;
    const std::int32_t* conn;
    const double* data;
    int n_data;
    int valence;
};
}

namespace
{

thread_local std::string g_last_error;

struct Result
{
    double f = 0.0;
    Eigen::VectorXd g;
    Eigen::SparseMatrix<double> H;
};

struct DeviceProblem  // owns the product-side function for the lifetime of one evaluation
{
    void* h = nullptr;
    ~DeviceProblem()
    {
        if (h) tadx_destroy(h);
    }
};

template <int d>
void eval_impl(std::int64_t n_vertices, int n_terms, const plugin_term* terms, int assembly, int mode, const double* x_in, double eps, Result& out)
{
    // the product side: element functors on the device
    DeviceProblem dev;
    if (tadx_create(d, n_vertices, 0, 0, assembly, &dev.h) != 0) throw std::runtime_error(std::string("tadx_create: ") + tadx_last_error());
    Eigen::Index n_elements = 0;
    for (int i = 0; i < n_terms; ++i)
    {
        if (tadx_add_term(dev.h, terms[i].kind, terms[i].n_elements, terms[i].conn, terms[i].valence, terms[i].data, terms[i].n_data) != 0)
            throw std::runtime_error(std::string("tadx_add_term: ") + tadx_last_error());
    }
    const tad_function fn = static_cast<tad_function>(tadx_handle(dev.h));
    n_elements = (Eigen::Index)tad_function_n_elements(fn);

    // the reference side: its own facade with the B200 term plugged in
    auto func = TinyAD::scalar_function<d>(TinyAD::range(n_vertices));
    func.objective_terms.push_back(std::make_unique<TinyAD::B200ScalarObjectiveTerm>(fn, n_elements));
    func.n_elements += n_elements;

    Eigen::VectorXd x = Eigen::Map<const Eigen::VectorXd>(x_in, func.n_vars);
    switch (mode)
    {
    case 0: out.f = func.eval(x); break;
    case 1: func.eval_with_gradient(x, out.f, out.g); break;
    case 2: func.eval_with_derivatives(x, out.f, out.g, out.H); break;
    case 3: func.eval_with_hessian_proj(x, out.f, out.g, out.H, eps); break;
    default: throw std::runtime_error("plugin: bad mode");
    }
}

}  // namespace

// Only the plugin_* entry points are exported (the library is built with -fvisibility=hidden): this library holds the REFERENCE's
// TinyAD:: symbols, the product's energies library holds its own facade under the same names, and neither may interpose the other.
#define PLUGIN_API __attribute__((visibility("default")))

extern "C" {

PLUGIN_API const char* plugin_last_error() { return g_last_error.c_str(); }

PLUGIN_API void* plugin_scalar_eval(int d, std::int64_t n_vertices, int n_terms, const plugin_term* terms, int assembly, int mode, const double* x, double eps)
{
    auto res = std::make_unique<Result>();
    try
    {
        if (d == 1) eval_impl<1>(n_vertices, n_terms, terms, assembly, mode, x, eps, *res);
        else if (d == 2) eval_impl<2>(n_vertices, n_terms, terms, assembly, mode, x, eps, *res);
        else if (d == 3) eval_impl<3>(n_vertices, n_terms, terms, assembly, mode, x, eps, *res);
        else throw std::runtime_error("plugin: unsupported variable dimension");
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return nullptr;
    }
    return res.release();
}

PLUGIN_API double plugin_result_f(void* r) { return ((Result*)r)->f; }
PLUGIN_API std::int64_t plugin_result_nnz(void* r) { return ((Result*)r)->H.nonZeros(); }
PLUGIN_API std::int64_t plugin_result_cols(void* r) { return ((Result*)r)->H.cols(); }
PLUGIN_API std::int64_t plugin_result_g_size(void* r) { return ((Result*)r)->g.size(); }
PLUGIN_API void plugin_result_copy(void* rp, double* g, std::int32_t* outer, std::int32_t* inner, double* values)
{
    Result* r = (Result*)rp;
    if (g) std::copy(r->g.data(), r->g.data() + r->g.size(), g);
    if (outer) std::copy(r->H.outerIndexPtr(), r->H.outerIndexPtr() + r->H.cols() + 1, outer);
    if (inner) std::copy(r->H.innerIndexPtr(), r->H.innerIndexPtr() + r->H.nonZeros(), inner);
    if (values) std::copy(r->H.valuePtr(), r->H.valuePtr() + r->H.nonZeros(), values);
}
PLUGIN_API void plugin_result_free(void* r) { delete (Result*)r; }

}  // extern "C"
