// tinyad_b200 -- "user translation unit" of the tests and the benchmark.
//
// The element functors below restate the element lambdas of the reference's own tests and of the
// configurations in BASELINE.json, written against the public surface of this framework
// (TinyAD::scalar_function<d>, add_elements<N>, TINYAD_SCALAR_TYPE, element.variables(...)):
//   SymDirichlet2D   tests/NewtonTest.cc:28-44      (Double<6>,  config C1)
//   Penalty<d>       tests/NewtonTest.cc:48-55
//   SymDirichlet3D   3-D analogue, SURVEY 8(d) C2 / C5 (Double<12>) -- the headline workload
//   Sos*             tests/GaussNewtonTest.cc:34-72  (VectorFunction twins), polycurl stand-in (C4)
//   misc             tests/ScalarFunctionTest.cc, tests/DynamicElementsTest.cc:60-91
// nvcc instantiates the element kernels (TinyAD/Kernels.cuh) for them here; everything else happens
// in the functor-independent runtime behind the C ABI (include/tinyad_b200.h).
// The extern "C" tadx_* functions only exist so that pytest / bench.py can drive the C++ facade via ctypes.
#include <memory>
#include <string>
#include <vector>

#include "energies.cuh"

#include <TinyAD/Utils/LineSearch.hh>
#include <TinyAD/Utils/NewtonDecrement.hh>
#include <TinyAD/Utils/NewtonDirection.hh>

using namespace TinyAD;
using namespace tadx;

// tests/HandleTypeTest.cc:8-42: a user-defined handle type becomes usable by overloading idx_from_handle (found by
// argument-dependent lookup in the handle's namespace, or declared in namespace TinyAD before the facade is included)
namespace tadx
{
struct CustomVariableHandle { int idx = -1; };
struct CustomElementHandle { int idx = -1; };
inline int64_t idx_from_handle(const CustomVariableHandle& vh) { return vh.idx; }
inline int64_t idx_from_handle(const CustomElementHandle& eh) { return eh.idx; }
}  // namespace tadx

// The Double<12> tet kernels are compiled in parallel, one Hessian part per translation unit (energies_tet_part.cu).
namespace TinyAD { namespace detail {
#define TADX_EXTERN_PART(P) extern template int launch_second_order_part<tadx::SymDirichlet3D, 3, 4, TADX_TET_PARTS, P, false>(const tadx::SymDirichlet3D&, const tad_launch_args&);
TADX_EXTERN_PART(0)
#if TADX_TET_PARTS > 1
TADX_EXTERN_PART(1)
#endif
#if TADX_TET_PARTS > 2
TADX_EXTERN_PART(2)
#endif
#if TADX_TET_PARTS > 3
TADX_EXTERN_PART(3)
#endif
#if TADX_TET_PARTS > 4
TADX_EXTERN_PART(4)
#endif
#if TADX_TET_PARTS > 5
TADX_EXTERN_PART(5)
#endif
#if TADX_TET_PARTS > 6
#error "add more TADX_EXTERN_PART lines"
#endif
} }

namespace
{

thread_local std::string g_err;

enum
{
    K_SYMDIRICHLET2D = 1, K_PENALTY2D = 2, K_SYMDIRICHLET3D = 3, K_PENALTY3D = 4, K_EDGE_DIRICHLET1D = 5, K_ARAP2D = 12, K_DYN_SUM_SQR2D = 10, K_DYN_ONERING1D = 11,
    K_QUADRATIC2D = 6, K_REPEATED_HANDLE = 7, K_TRIG_MIX2D = 8, K_SQRT1D = 9, K_BRANCH_ON_X1D = 13,
    K_SOS_SYMDIRICHLET2D = 101, K_SOS_PENALTY2D = 102, K_SOS_POLYCURL2D = 103, K_SOS_TEST1D_A = 104, K_SOS_TEST1D_B = 105,
};

struct DeviceArray
{
    void* p = nullptr;
    ~DeviceArray() { if (p) cudaFree(p); }
};

// row-major (n x c) host array -> SoA [c][stride] on the device
template <class T>
bool upload_soa(const T* host, int64_t n, int c, int64_t stride, DeviceArray& out)
{
    std::vector<T> tmp((size_t)std::max<int64_t>(1, c * stride), T(0));
    for (int64_t e = 0; e < n; ++e)
        for (int j = 0; j < c; ++j) tmp[(size_t)(j * stride + e)] = host[e * c + j];
    if (cudaMalloc(&out.p, tmp.size() * sizeof(T)) != cudaSuccess) return false;
    return cudaMemcpy(out.p, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

struct Problem
{
    int d = 0;
    bool is_vector = false;
    std::unique_ptr<ScalarFunction<1>> s1;
    std::unique_ptr<ScalarFunction<2>> s2;
    std::unique_ptr<ScalarFunction<3>> s3;
    std::unique_ptr<VectorFunction<2>> v2;
    std::unique_ptr<VectorFunction<1>> v1;
    std::vector<std::unique_ptr<DeviceArray>> arrays;
    tad_function handle() const
    {
        if (s1) return s1->handle();
        if (s2) return s2->handle();
        if (s3) return s3->handle();
        if (v2) return v2->handle();
        if (v1) return v1->handle();
        return nullptr;
    }
};

template <class Fn>
int guarded(Fn&& fn)
{
    try { fn(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}

}  // namespace

extern "C" {

const char* tadx_last_error() { return g_err.c_str(); }

int tadx_create(int d, int64_t n_vertices, int is_vector, int device, int assembly, void** out)
{
    *out = nullptr;
    return guarded([&] {
        auto P = std::make_unique<Problem>();
        P->d = d;
        P->is_vector = is_vector != 0;
        EvalSettings s;
        s.device = device;
        s.assembly = assembly;
        if (is_vector)
        {
            if (d == 2) P->v2 = std::make_unique<VectorFunction<2>>(vector_function<2>(range(n_vertices), s));
            else if (d == 1) P->v1 = std::make_unique<VectorFunction<1>>(vector_function<1>(range(n_vertices), s));
            else throw std::runtime_error("vector functions are instantiated for d = 1, 2 only");
        }
        else if (d == 1) P->s1 = std::make_unique<ScalarFunction<1>>(scalar_function<1>(range(n_vertices), s));
        else if (d == 2) P->s2 = std::make_unique<ScalarFunction<2>>(scalar_function<2>(range(n_vertices), s));
        else if (d == 3) P->s3 = std::make_unique<ScalarFunction<3>>(scalar_function<3>(range(n_vertices), s));
        else throw std::runtime_error("unsupported variable dimension");
        *out = P.release();
    });
}

void tadx_destroy(void* h) { delete static_cast<Problem*>(h); }

void* tadx_handle(void* h) { return h ? static_cast<Problem*>(h)->handle() : nullptr; }

// conn: n_elements x valence (row-major int32), data: n_elements x n_data (row-major double)
int tadx_add_term(void* h, int kind, int64_t n_elements, const int32_t* conn, int valence, const double* data, int n_data)
{
    Problem* P = static_cast<Problem*>(h);
    return guarded([&] {
        const int64_t stride = ((n_elements + 31) / 32) * 32;
        auto dc = std::make_unique<DeviceArray>();
        auto dd = std::make_unique<DeviceArray>();
        if (!upload_soa(conn, n_elements, valence, stride, *dc) || !upload_soa(data, n_elements, n_data, stride, *dd))
            throw std::runtime_error("device upload of element data failed");
        ConnView C{static_cast<const int32_t*>(dc->p), stride};
        DataView D{static_cast<const double*>(dd->p), stride};
        P->arrays.push_back(std::move(dc));
        P->arrays.push_back(std::move(dd));
        const auto els = range(n_elements);
        auto need = [&](bool ok) { if (!ok) throw std::runtime_error("term kind does not match the function's variable dimension / type"); };
        switch (kind)
        {
        case K_SYMDIRICHLET2D: need(P->s2 && valence == 3 && n_data == 5); P->s2->add_elements<3>(els, SymDirichlet2D{C, D}); break;
        case K_PENALTY2D: need(P->s2 && valence == 1 && n_data == 2); P->s2->add_elements<1>(els, Penalty<2>{C, D}); break;
        case K_SYMDIRICHLET3D: need(P->s3 && valence == 4 && n_data == 10); P->s3->add_elements<4>(els, SymDirichlet3D{C, D}); break;
        case K_PENALTY3D: need(P->s3 && valence == 1 && n_data == 3); P->s3->add_elements<1>(els, Penalty<3>{C, D}); break;
        case K_EDGE_DIRICHLET1D: need(P->s1 && valence == 2 && n_data == 1); P->s1->add_elements<2>(els, EdgeDirichlet1D{C, D}); break;
        case K_SQRT1D: need(P->s1 && valence == 1 && n_data == 1); P->s1->add_elements<1>(els, Sqrt1D{C, D}); break;
        case K_BRANCH_ON_X1D: need(P->s1 && valence == 2 && n_data == 1); P->s1->add_elements<2>(els, BranchOnX1D{C, D}); break;
        case K_QUADRATIC2D: need(P->s2 && valence == 1 && n_data == 1); P->s2->add_elements<1>(els, Quadratic2D{C, D}); break;
        case K_REPEATED_HANDLE: need(P->s2 && valence == 2 && n_data == 1); P->s2->add_elements<2>(els, RepeatedHandle{C, D}); break;
        case K_TRIG_MIX2D: need(P->s2 && valence == 2 && n_data == 1); P->s2->add_elements<2>(els, TrigMix2D{C, D}); break;
        case K_ARAP2D: need(P->s2 && valence == 3 && n_data == 5); P->s2->add_elements<3>(els, Arap2D{C, D}); break;
        case K_DYN_SUM_SQR2D: need(P->s2 != nullptr); P->s2->add_elements_dynamic<3, 1>(els, DynSumSqr2D{C, D}); break;
        case K_DYN_ONERING1D: need(P->s1 != nullptr); P->s1->add_elements_dynamic<4, 6, 7, 10>(els, OneRingDirichlet1D{C, D, valence}); break;
        case K_SOS_SYMDIRICHLET2D: need(P->v2 && valence == 3 && n_data == 5); P->v2->add_elements<3, 8>(els, SosSymDirichlet2D{C, D}); break;
        case K_SOS_PENALTY2D: need(P->v2 && valence == 1 && n_data == 2); P->v2->add_elements<1, 2>(els, SosPenalty2D{C, D}); break;
        case K_SOS_POLYCURL2D: need(P->v2 && valence == 2 && n_data == 3); P->v2->add_elements<2, 2>(els, SosPolycurl2D{C, D}); break;
        case K_SOS_TEST1D_A: need(P->v1 && valence == 1); P->v1->add_elements<1, 2>(els, SosTest1DA{C, D}); break;
        case K_SOS_TEST1D_B: need(P->v1 && valence == 1); P->v1->add_elements<1, 1>(els, SosTest1DB{C, D}); break;
        default: throw std::runtime_error("unknown term kind");
        }
    });
}

// ---- known-answer scalar cases on the device (one thread) or on the host ----
int tadx_scalar_case(int id, const double* params16, double* out, int on_device);

// ---- facade semantics of the reference's tests, exercised in C++ (returns 0 on success, else a failing step) ----
int tadx_selftest(int device)
{
    g_err.clear();
    try
    {
        EvalSettings es;
        es.device = device;
        // default-constructed functions evaluate to 0 / empty (tests/ScalarFunctionTest.cc:183-189)
        ScalarFunction<2> empty;
        if (empty.eval(std::vector<double>()) != 0.0) return 1;
        auto [f0, g0, H0] = empty.eval_with_hessian_proj(std::vector<double>());
        if (f0 != 0.0 || !g0.empty() || H0.nonZeros() != 0) return 2;
        // a function built in place, then moved: the moved-to object evaluates, the moved-from one is empty (:190-252)
        DeviceArray dc, dd;
        const int32_t conn[1] = {0};
        const double data[1] = {1.0};
        if (!upload_soa(conn, 1, 1, 32, dc) || !upload_soa(data, 1, 1, 32, dd)) return 3;
        ConnView C{static_cast<const int32_t*>(dc.p), 32};
        DataView D{static_cast<const double*>(dd.p), 32};
        auto func = scalar_function<2>(range(1), es);
        func.add_elements<1>(range(1), Quadratic2D{C, D});
        ScalarFunction<2> moved(std::move(func));
        const std::vector<double> x = {1.0, 2.0};
        if (moved.eval(x) != 12.0) return 4;                                   // tests/ScalarFunctionTest.cc:72-110
        auto [f, g, H] = moved.eval_with_derivatives(x);
        if (f != 12.0 || g[0] != 9.0 || g[1] != 6.0) return 5;
        if (H.coeff(0, 0) != 4.0 || H.coeff(0, 1) != 2.0 || H.coeff(1, 0) != 2.0 || H.coeff(1, 1) != 2.0) return 6;
        if (func.n_vars != 0 || func.eval(std::vector<double>()) != 0.0) return 7;
        ScalarFunction<2> assigned;
        assigned = std::move(moved);
        if (assigned(x) != 12.0) return 8;
        // size mismatch throws (ScalarFunctionImpl.hh:262) and the object stays usable
        bool threw = false;
        try { assigned.eval(std::vector<double>(3, 0.0)); } catch (const std::runtime_error&) { threw = true; }
        if (!threw || assigned.eval(x) != 12.0) return 9;
        // an error inside the element evaluation is reported to the caller and the function stays usable
        // (tests/ExceptionTest.cc:8-54; here: non-finite derivative, ScalarObjectiveTerm.hh:210,252-253)
        auto fs = scalar_function<1>(range(1), es);
        fs.add_elements<1>(range(1), Sqrt1D{C, D});
        threw = false;
        try { fs.eval_with_gradient(std::vector<double>{-1.0}); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) return 10;
        auto [fv, gv] = fs.eval_with_gradient(std::vector<double>{4.0});
        if (fv != 2.0 || gv[0] != 0.25) return 11;
        // x_from_data / x_to_data (ScalarFunctionImpl.hh:216-254)
        auto xs = assigned.x_from_data([](int64_t) { return std::vector<double>{3.0, 5.0}; });
        if (xs.size() != 2 || xs[0] != 3.0 || xs[1] != 5.0) return 12;
        double seen = 0.0;
        assigned.x_to_data(xs, [&](int64_t, const Vec<double, 2>& pv) { seen = pv[0] + pv[1]; });
        if (seen != 8.0) return 13;
        // non-compact variable indices are rejected (ScalarFunctionImpl.hh:59)
        threw = false;
        try { ScalarFunction<2> bad(std::vector<int64_t>{0, 2}, es); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) return 14;
        // one projected-Newton step with the solver utilities (loop of tests/NewtonTest.cc:68-75) on the convex quadratic:
        // H = [[4,2],[2,2]], g = (9,6) at x = (1,2)  ->  d = -H^-1 g = (-1.5,-1.5), decrement 11.25, full step accepted, g(x+d) = 0
        {
            auto [fq, gq, Hq] = assigned.eval_with_hessian_proj(x);
            LinearSolver<double> solver;
            solver.block_dim = 2;
            solver.rel_tol = 1e-14;
            const std::vector<double> dq = newton_direction(gq, Hq, solver);
            if (std::fabs(dq[0] + 1.5) > 1e-12 || std::fabs(dq[1] + 1.5) > 1e-12) return 15;
            if (std::fabs(newton_decrement(dq, gq) - 11.25) > 1e-11) return 16;
            const std::vector<double> xn = line_search(x, dq, fq, gq, assigned);
            if (std::fabs(xn[0] + 0.5) > 1e-12 || std::fabs(xn[1] - 0.5) > 1e-12) return 17;
            auto [fn2, gn2] = assigned.eval_with_gradient(xn);
            if (std::fabs(gn2[0]) > 1e-11 || std::fabs(gn2[1]) > 1e-11 || !(fn2 < fq)) return 18;
            // an indefinite matrix is reported like the reference's "Linear solve failed." (NewtonDirection.hh:43-44)
            SparseMatrix Hneg = Hq;
            for (double& v : Hneg.values) v = -v;
            threw = false;
            try { newton_direction(gq, Hneg, solver); } catch (const std::runtime_error&) { threw = true; }
            if (!threw) return 19;
        }
        // tests/VectorFunctionTest.cc:73-172 (test_eval) through the facade, incl. eval_with_derivatives (per-residual Hessians)
        {
            DeviceArray ca, cb;
            const int32_t conn_a[1] = {0}, conn_b[1] = {1};
            if (!upload_soa(conn_a, 1, 1, 32, ca) || !upload_soa(conn_b, 1, 1, 32, cb)) return 30;
            auto vf = vector_function<1>(range(2), es);
            vf.add_elements<1, 2>(range(1), SosTest1DA{ConnView{static_cast<const int32_t*>(ca.p), 32}, D});
            vf.add_elements<1, 1>(range(1), SosTest1DB{ConnView{static_cast<const int32_t*>(cb.p), 32}, D});
            const std::vector<double> xv = {3.0, 4.0};
            const double r_exp[3] = {6.0, 9.0, 16.0};
            const std::vector<double> rv = vf.eval(xv);
            for (int i = 0; i < 3; ++i) if (rv[(size_t)i] != r_exp[i]) return 31;
            if (vf.eval_sum_of_squares(xv) != 373.0) return 32;
            auto [r1, J1] = vf.eval_with_jacobian(xv);
            if (J1.coeff(0, 0) != 2.0 || J1.coeff(1, 0) != 6.0 || J1.coeff(2, 1) != 8.0 || J1.coeff(2, 0) != 0.0 || J1.nonZeros() != 3) return 33;
            auto [r2, J2, H2] = vf.eval_with_derivatives(xv);
            for (int i = 0; i < 3; ++i) if (r2[(size_t)i] != r_exp[i]) return 34;
            if (J2.coeff(0, 0) != 2.0 || J2.coeff(1, 0) != 6.0 || J2.coeff(2, 1) != 8.0) return 35;
            if (H2.size() != 3) return 36;
            // H_expected[1](0,0) = 2, H_expected[2](1,1) = 2, everything else 0 (:101-103)
            for (int i = 0; i < 3; ++i)
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b)
                    {
                        const double want = ((i == 1 && a == 0 && b == 0) || (i == 2 && a == 1 && b == 1)) ? 2.0 : 0.0;
                        if (H2[(size_t)i].coeff(a, b) != want) return 37;
                    }
            auto [f3, g3, r3, J3] = vf.eval_sum_of_squares_with_derivatives(xv);
            if (f3 != 373.0 || g3[0] != 8.0 * 3.0 + 4.0 * 27.0 || g3[1] != 4.0 * 64.0) return 38;
        }
        // mesh-library handle types (Support/Common.hh; reference Support/{OpenMesh,PMP,Polymesh,GeometryCentral}.hh): handles with
        // .idx(), .idx.value or .getIndex() enumerate variables and elements on the host; the device functor sees the integer
        {
            struct IdxHandle { int i; int idx() const { return i; } };                       // OpenMesh::BaseHandle / pmp::Handle
            struct PmIndex { int value; };
            struct PmHandle { PmIndex idx; };                                               // pm::primitive_handle<tag>
            struct GcHandle { size_t i; size_t getIndex() const { return i; } };             // geometrycentral::Element<T, M>
            if (idx_from_handle(IdxHandle{3}) != 3 || idx_from_handle(PmHandle{{4}}) != 4 || idx_from_handle(GcHandle{5}) != 5 ||
                idx_from_handle(7) != 7 || idx_from_handle((int64_t)8) != 8) return 20;
            threw = false;
            try { struct Opaque {}; idx_from_handle(Opaque{}); } catch (const std::runtime_error&) { threw = true; }
            if (!threw) return 21;
            std::vector<IdxHandle> vertices = {{0}};
            std::vector<GcHandle> faces = {{0}};
            auto fm = scalar_function<2>(vertices, es);
            fm.add_elements<1>(faces, Quadratic2D{C, D});
            if (fm.eval(x) != 12.0) return 22;
            auto xm = fm.x_from_data([](IdxHandle v) { return std::vector<double>{1.0 + v.idx(), 2.0}; });
            if (xm.size() != 2 || xm[0] != 1.0 || xm[1] != 2.0) return 23;
            // tests/HandleTypeTest.cc:44-66: custom variable / element handle types with user overloads of idx_from_handle
            std::vector<tadx::CustomVariableHandle> cv = {{0}};
            std::vector<tadx::CustomElementHandle> ce = {{0}};
            auto fc = scalar_function<2>(cv, es);
            fc.add_elements<1>(ce, Quadratic2D{C, D});
            if (fc.eval(x) != 12.0) return 25;
            auto [fcv, fcg, fcH] = fc.eval_with_hessian_proj(x);
            if (fcv != 12.0 || fcH.nonZeros() != 4) return 26;
            std::vector<IdxHandle> sparse_ids = {{0}, {2}};                                // not compact -> rejected like integer handles
            threw = false;
            try { auto bad = scalar_function<2>(sparse_ids, es); } catch (const std::runtime_error&) { threw = true; }
            if (!threw) return 24;
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"

namespace
{
__global__ void scalar_case_kernel(int id, const double* p, double* out, int* n) { *n = tadx::scalar_case_run(id, p, out); }
}  // namespace

extern "C" int tadx_scalar_case(int id, const double* params16, double* out, int on_device)
{
    if (!on_device) return tadx::scalar_case_run(id, params16, out);
    double *dp = nullptr, *dout = nullptr;
    int* dn = nullptr;
    int n = -100;
    if (cudaMalloc(&dp, 16 * sizeof(double)) != cudaSuccess || cudaMalloc(&dout, 256 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&dn, sizeof(int)) != cudaSuccess)
        return -100;
    cudaMemcpy(dp, params16, 16 * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 256 * sizeof(double));
    scalar_case_kernel<<<1, 1>>>(id, dp, dout, dn);
    if (cudaDeviceSynchronize() == cudaSuccess)
    {
        cudaMemcpy(&n, dn, sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(out, dout, 256 * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cudaFree(dp); cudaFree(dout); cudaFree(dn);
    return n;
}
