// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (see tinyad_oracle.hh).
//
// The energies below restate the element lambdas the reference's own tests use,
// written against the oracle's Eigen-free types so that they read like the originals:
//   symdirichlet2d      tests/NewtonTest.cc:28-44        (2-D symmetric Dirichlet, Double<6>)
//   penalty<d>          tests/NewtonTest.cc:48-55        (positional penalty, Double<d>)
//   symdirichlet3d      3-D analogue (vol*(|J|^2+|J^-1|^2), Double<12>)
//   sos terms           tests/GaussNewtonTest.cc:34-72   (VectorFunction twins)
//   edge dirichlet      tests/DynamicElementsTest.cc:60-91 (Hessian == graph Laplacian)
//   quadratic tests     tests/ScalarFunctionTest.cc:10-179
//   polycurl stand-in   complex-valued per-edge residuals (Scalar.hh:1151-1320), SURVEY 8(d) C4
#include "tinyad_oracle.hh"

#include <cstring>
#include <map>

using namespace oracle;

extern "C" {

// One objective term.  `conn` is n_elements x valence (row-major int32 vertex handles),
// `data` is n_elements x n_data (row-major doubles).
struct oracle_term
{
    int kind;
    std::int64_t n_elements;
    const std::int32_t* conn;
    const double* data;
    int n_data;
};

enum
{
    ORC_SYMDIRICHLET2D = 1,      // d=2 N=3 data: Mr(0,0) Mr(0,1) Mr(1,0) Mr(1,1) w
    ORC_PENALTY2D = 2,           // d=2 N=1 data: tx ty
    ORC_SYMDIRICHLET3D = 3,      // d=3 N=4 data: Mr^-1 row-major (9) vol
    ORC_PENALTY3D = 4,           // d=3 N=1 data: tx ty tz
    ORC_EDGE_DIRICHLET1D = 5,    // d=1 N=2 data: w        w*(x_a-x_b)^2
    ORC_QUADRATIC2D = 6,         // d=2 N=1 data: sign     sign*(2x0^2+2x0x1+x1^2+x0+1)   (ScalarFunctionTest.cc:72-146)
    ORC_REPEATED_HANDLE = 7,     // d=2 N=2 accesses conn[0] twice then conn[1]             (ScalarFunctionTest.cc:153-179)
    ORC_TRIG_MIX2D = 8,          // d=2 N=2 exercises sin/cos/exp/log/sqrt/atan2/pow/hypot/tanh on 4 variables
    ORC_ARAP2D = 12,             // d=2 N=3 data: Mr (4) w      w * |J - closest_orthogonal(J)|^2, J = M Mr^-1   (Operations/SVD.hh)
    ORC_DYN_SUM_SQR2D = 10,      // d=2 dynamic <3,1>: element e accesses handles 0..e-1, |sum|^2   (DynamicElementsTest.cc:9-33)
    ORC_DYN_ONERING1D = 11,      // d=1 dynamic <4,6,7,10>: conn = padded neighbour table (n_data columns, -1 = none)
                                 //     0.25 * sum_n (x_v - x_n)^2                                (DynamicElementsTest.cc:92-141)
    ORC_SOS_SYMDIRICHLET2D = 101,  // d=2 N=3 M=8 data: Mr (4) scale
    ORC_SOS_PENALTY2D = 102,       // d=2 N=1 M=2 data: tx ty
    ORC_SOS_POLYCURL2D = 103,      // d=2 N=2 M=2 data: ex ey w   complex residual, see below
};

}  // extern "C"

namespace
{

std::vector<Index> handles(Index n) { return range(n); }

template <typename FuncT>
void add_scalar_term(FuncT& func, const oracle_term& t)
{
    const std::int32_t* conn = t.conn;
    const double* data = t.data;
    const int nd = t.n_data;
    switch (t.kind)
    {
    case ORC_SYMDIRICHLET2D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<2>>)
            func.template add_elements<3>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index e = element.handle;
                const double* dd = data + e * nd;
                Mat<double, 2, 2> Mr;
                Mr(0, 0) = dd[0]; Mr(0, 1) = dd[1]; Mr(1, 0) = dd[2]; Mr(1, 1) = dd[3];
                Vec<T, 2> a = element.variables(conn[3 * e + 0]);
                Vec<T, 2> b = element.variables(conn[3 * e + 1]);
                Vec<T, 2> c = element.variables(conn[3 * e + 2]);
                Mat<T, 2, 2> M = col_mat(b - a, c - a);
                if (M.determinant() <= 0.0) return (T)INFINITY;
                return ((M * Mr.inverse()).squaredNorm() + (Mr * M.inverse()).squaredNorm()) * dd[4];
            });
        break;
    case ORC_PENALTY2D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<2>>)
            func.template add_elements<1>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index e = element.handle;
                Vec<double, 2> p_target;
                p_target[0] = data[e * nd + 0]; p_target[1] = data[e * nd + 1];
                Vec<T, 2> p = element.variables(conn[e]);
                return (p_target - p).squaredNorm();
            });
        break;
    case ORC_SYMDIRICHLET3D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<3>>)
            func.template add_elements<4>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index e = element.handle;
                const double* dd = data + e * nd;
                Mat<double, 3, 3> Mr_inv;
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) Mr_inv(i, j) = dd[3 * i + j];
                Vec<T, 3> a = element.variables(conn[4 * e + 0]);
                Vec<T, 3> b = element.variables(conn[4 * e + 1]);
                Vec<T, 3> c = element.variables(conn[4 * e + 2]);
                Vec<T, 3> d = element.variables(conn[4 * e + 3]);
                Mat<T, 3, 3> M = col_mat(b - a, c - a, d - a);
                if (M.determinant() <= 0.0) return (T)INFINITY;
                Mat<T, 3, 3> J = M * Mr_inv;
                return (J.squaredNorm() + J.inverse().squaredNorm()) * dd[9];
            });
        break;
    case ORC_PENALTY3D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<3>>)
            func.template add_elements<1>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index e = element.handle;
                Vec<double, 3> p_target;
                for (int i = 0; i < 3; ++i) p_target[i] = data[e * nd + i];
                Vec<T, 3> p = element.variables(conn[e]);
                return (p_target - p).squaredNorm();
            });
        break;
    case ORC_EDGE_DIRICHLET1D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<1>>)
            func.template add_elements<2>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index e = element.handle;
                T xa = element.variable(conn[2 * e + 0]);
                T xb = element.variable(conn[2 * e + 1]);
                return data[e * nd] * sqr(xa - xb);
            });
        break;
    case ORC_QUADRATIC2D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<2>>)
            func.template add_elements<1>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index e = element.handle;
                Vec<T, 2> x = element.variables(conn[e]);
                return data[e * nd] * (2.0 * sqr(x[0]) + 2.0 * x[0] * x[1] + sqr(x[1]) + x[0] + 1.0);
            });
        break;
    case ORC_REPEATED_HANDLE:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<2>>)
            func.template add_elements<2>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index e = element.handle;
                Vec<T, 2> v = element.variables(conn[2 * e + 0]);
                Vec<T, 2> v2 = element.variables(conn[2 * e + 0]);  // same handle again -> same local slots
                Vec<T, 2> w = element.variables(conn[2 * e + 1]);
                return v[0] * v2[1] + sqr(w[0]) * v2[0] + w[1] * v[1] * 3.0;
            });
        break;
    case ORC_TRIG_MIX2D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<2>>)
            func.template add_elements<2>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index e = element.handle;
                Vec<T, 2> p = element.variables(conn[2 * e + 0]);
                Vec<T, 2> q = element.variables(conn[2 * e + 1]);
                T r = hypot(p[0] - q[0], p[1] - q[1]) + 0.5;
                T s = sin(p[0]) * cos(q[1]) + exp(0.25 * p[1]) / (1.0 + sqr(q[0]));
                T u = log(r) + sqrt(r + sqr(s)) + atan2(p[1] + 2.0, q[0] + 3.0);
                T v = pow(r, 3) - pow(r, 1.5) + tanh(s) * data[e * nd];
                return u * v + fabs(s - 0.1) + 2.0 / r - (1.0 - s) / 3.0;
            });
        break;
    case ORC_ARAP2D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<2>>)
            func.template add_elements<3>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index e = element.handle;
                const double* dd = data + e * nd;
                Mat<double, 2, 2> Mr;
                Mr(0, 0) = dd[0]; Mr(0, 1) = dd[1]; Mr(1, 0) = dd[2]; Mr(1, 1) = dd[3];
                Vec<T, 2> a = element.variables(conn[3 * e + 0]);
                Vec<T, 2> b = element.variables(conn[3 * e + 1]);
                Vec<T, 2> c = element.variables(conn[3 * e + 2]);
                Mat<T, 2, 2> J = col_mat(b - a, c - a) * Mr.inverse();
                Mat<T, 2, 2> R = closest_orthogonal(J);
                return (J - R).squaredNorm() * dd[4];
            });
        break;
    case ORC_DYN_SUM_SQR2D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<2>>)
            func.template add_elements_dynamic<3, 1>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const int e = (int)element.handle;
                Vec<T, 2> sum;
                for (int i = 0; i < 2; ++i) sum[i] = T(0.0);
                for (int v = 0; v < e; ++v) sum = sum + element.variables(v);
                return sum.squaredNorm();
            });
        break;
    case ORC_DYN_ONERING1D:
        if constexpr (std::is_same_v<FuncT, ScalarFunction<1>>)
            func.template add_elements_dynamic<4, 6, 7, 10>(handles(t.n_elements), [=](auto& element) -> ORACLE_SCALAR_TYPE(element) {
                using T = ORACLE_SCALAR_TYPE(element);
                const Index v = element.handle;
                T v_val = element.variable(v);
                std::vector<T> neigh_vals;
                for (int i = 0; i < nd && conn[v * nd + i] >= 0; ++i) neigh_vals.push_back(element.variable(conn[v * nd + i]));
                T dirichlet = 0.0;
                for (size_t i = 0; i < neigh_vals.size(); ++i) dirichlet = dirichlet + 0.25 * sqr(v_val - neigh_vals[i]);
                return dirichlet;
            });
        break;
    default:
        error_throw("oracle: unknown scalar term kind");
    }
}

template <int d>
std::unique_ptr<ScalarFunction<d>> build_scalar(std::int64_t n_vertices, int n_terms, const oracle_term* terms, int n_threads)
{
    EvalSettings s;
    s.n_threads = n_threads;
    auto func = std::make_unique<ScalarFunction<d>>(n_vertices, s);
    for (int i = 0; i < n_terms; ++i) add_scalar_term(*func, terms[i]);
    return func;
}

template <typename FuncT>
void add_vector_term(FuncT& func, const oracle_term& t)
{
    const std::int32_t* conn = t.conn;
    const double* data = t.data;
    const int nd = t.n_data;
    switch (t.kind)
    {
    case ORC_SOS_SYMDIRICHLET2D:
        func.template add_elements<3, 8>(handles(t.n_elements), [=](auto& element) -> ORACLE_VECTOR_TYPE(element) {
            using T = ORACLE_SCALAR_TYPE(element);
            const Index e = element.handle;
            const double* dd = data + e * nd;
            Mat<double, 2, 2> Mr;
            Mr(0, 0) = dd[0]; Mr(0, 1) = dd[1]; Mr(1, 0) = dd[2]; Mr(1, 1) = dd[3];
            Vec<T, 2> a = element.variables(conn[3 * e + 0]);
            Vec<T, 2> b = element.variables(conn[3 * e + 1]);
            Vec<T, 2> c = element.variables(conn[3 * e + 2]);
            Mat<T, 2, 2> M = col_mat(b - a, c - a);
            if (M.determinant() <= 0.0) return Vec<T, 8>::Constant((T)INFINITY);
            Mat<T, 2, 2> J = M * Mr.inverse();
            Mat<T, 2, 2> J_inv = Mr * M.inverse();
            Vec<T, 8> E;
            E[0] = J(0, 0); E[1] = J(0, 1); E[2] = J(1, 0); E[3] = J(1, 1);
            E[4] = J_inv(0, 0); E[5] = J_inv(0, 1); E[6] = J_inv(1, 0); E[7] = J_inv(1, 1);
            return dd[4] * E;
        });
        break;
    case ORC_SOS_PENALTY2D:
        func.template add_elements<1, 2>(handles(t.n_elements), [=](auto& element) -> ORACLE_VECTOR_TYPE(element) {
            using T = ORACLE_SCALAR_TYPE(element);
            const Index e = element.handle;
            Vec<double, 2> p_target;
            p_target[0] = data[e * nd + 0]; p_target[1] = data[e * nd + 1];
            Vec<T, 2> p = element.variables(conn[e]);
            return p_target - p;
        });
        break;
    case ORC_SOS_POLYCURL2D:
        // Synthetic stand-in for a polycurl-style frame-field residual (the real lambda lives in
        // TinyAD-Examples, not in the reference tree): per interior edge with unit direction e=(ex,ey)
        // and per-face complex frame variables z_f = (u,v), z_g: residual = w * (z_f^4 - z_g^4) * conj(e)^4 ... simplified to
        //   c = (sqr(sqr(z_f)) - sqr(sqr(z_g))) * conj(edge)   with edge = ex + i ey,  r = w * (Re c, Im c)
        func.template add_elements<2, 2>(handles(t.n_elements), [=](auto& element) -> ORACLE_VECTOR_TYPE(element) {
            using T = ORACLE_SCALAR_TYPE(element);
            const Index e = element.handle;
            const double* dd = data + e * nd;
            Vec<T, 2> pf = element.variables(conn[2 * e + 0]);
            Vec<T, 2> pg = element.variables(conn[2 * e + 1]);
            Vec<T, 2> r;
            if constexpr (ORACLE_ACTIVE_MODE(element))
            {
                std::complex<T> zf(pf[0], pf[1]), zg(pg[0], pg[1]);
                std::complex<double> edge(dd[0], -dd[1]);
                std::complex<T> c = (sqr(sqr(zf)) - sqr(sqr(zg))) * edge;
                r[0] = dd[2] * c.real(); r[1] = dd[2] * c.imag();
            }
            else
            {
                std::complex<double> zf(pf[0], pf[1]), zg(pg[0], pg[1]);
                std::complex<double> edge(dd[0], -dd[1]);
                auto csqr = [](const std::complex<double>& a) { return std::complex<double>(a.real() * a.real() - a.imag() * a.imag(), 2.0 * a.real() * a.imag()); };
                std::complex<double> df = csqr(csqr(zf)), dg = csqr(csqr(zg));
                std::complex<double> dz(df.real() - dg.real(), df.imag() - dg.imag());
                std::complex<double> c(dz.real() * edge.real() - dz.imag() * edge.imag(), dz.real() * edge.imag() + dz.imag() * edge.real());
                r[0] = dd[2] * c.real(); r[1] = dd[2] * c.imag();
            }
            return r;
        });
        break;
    default:
        error_throw("oracle: unknown vector term kind");
    }
}

thread_local std::string g_last_error;

struct Result
{
    double f = 0.0;
    std::vector<double> g, r;
    SparseMatrix H;  // Hessian or Jacobian
    PhaseTimes pt;
};

template <int d>
int scalar_eval_impl(std::int64_t n_vertices, int n_terms, const oracle_term* terms, int mode, const double* x_in,
                     double eps, int n_threads, Result& out)
{
    auto func = build_scalar<d>(n_vertices, n_terms, terms, n_threads);
    std::vector<double> x(x_in, x_in + func->n_vars);
    out.pt.abs_sum = (mode & 8) != 0;    // modes 2|8, 3|8: g / H values = sum of |contributions| (scale of the per-entry parity bound)
    out.pt.norm_sum = (mode & 16) != 0;  // modes 2|16, 3|16: sum over the contributing elements of the element's max |entry|
    mode &= 7;
    switch (mode)
    {
    case 0: out.f = func->eval(x); break;
    case 1: func->eval_with_gradient(x, out.f, out.g); break;
    case 2: func->eval_with_derivatives(x, out.f, out.g, out.H, &out.pt); break;
    case 3: func->eval_with_hessian_proj(x, out.f, out.g, out.H, eps, &out.pt); break;
    default: error_throw("oracle: bad mode");
    }
    return 0;
}

}  // namespace

extern "C" {

const char* oracle_last_error() { return g_last_error.c_str(); }

// mode: 0 eval, 1 eval_with_gradient, 2 eval_with_derivatives, 3 eval_with_hessian_proj
// Returns an opaque result (free with oracle_result_free) or nullptr on error (see oracle_last_error()).
void* oracle_scalar_eval(int d, std::int64_t n_vertices, int n_terms, const oracle_term* terms, int mode,
                         const double* x, double eps, int n_threads)
{
    auto res = std::make_unique<Result>();
    try
    {
        if (d == 1) scalar_eval_impl<1>(n_vertices, n_terms, terms, mode, x, eps, n_threads, *res);
        else if (d == 2) scalar_eval_impl<2>(n_vertices, n_terms, terms, mode, x, eps, n_threads, *res);
        else if (d == 3) scalar_eval_impl<3>(n_vertices, n_terms, terms, mode, x, eps, n_threads, *res);
        else error_throw("oracle: unsupported variable dimension");
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return nullptr;
    }
    return res.release();
}

// mode: 0 eval (r), 1 eval_with_jacobian (r, J), 2 eval_sum_of_squares (f), 3 eval_sum_of_squares_with_derivatives (f,g,r,J)
void* oracle_vector_eval(int d, std::int64_t n_vertices, int n_terms, const oracle_term* terms, int mode,
                         const double* x_in, int n_threads)
{
    auto res = std::make_unique<Result>();
    try
    {
        if (d != 2) error_throw("oracle: vector functions are instantiated for d=2 only");
        EvalSettings s;
        s.n_threads = n_threads;
        VectorFunction<2> func(n_vertices, s);
        for (int i = 0; i < n_terms; ++i) add_vector_term(func, terms[i]);
        std::vector<double> x(x_in, x_in + func.n_vars);
        switch (mode)
        {
        case 0: res->r = func.eval(x); break;
        case 1: func.eval_with_jacobian(x, res->r, res->H); break;
        case 2: res->f = func.eval_sum_of_squares(x); break;
        case 3: func.eval_sum_of_squares_with_derivatives(x, res->f, res->g, res->r, res->H); break;
        default: error_throw("oracle: bad mode");
        }
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return nullptr;
    }
    return res.release();
}

double oracle_result_f(void* r) { return ((Result*)r)->f; }
std::int64_t oracle_result_nnz(void* r) { return ((Result*)r)->H.nonZeros(); }
std::int64_t oracle_result_rows(void* r) { return ((Result*)r)->H.rows; }
std::int64_t oracle_result_cols(void* r) { return ((Result*)r)->H.cols; }
std::int64_t oracle_result_g_size(void* r) { return (std::int64_t)((Result*)r)->g.size(); }
std::int64_t oracle_result_r_size(void* r) { return (std::int64_t)((Result*)r)->r.size(); }
void oracle_result_copy(void* rp, double* g, double* rvec, std::int32_t* outer, std::int32_t* inner, double* values)
{
    Result* r = (Result*)rp;
    if (g) std::copy(r->g.begin(), r->g.end(), g);
    if (rvec) std::copy(r->r.begin(), r->r.end(), rvec);
    if (outer) std::copy(r->H.outer.begin(), r->H.outer.end(), outer);
    if (inner) std::copy(r->H.inner.begin(), r->H.inner.end(), inner);
    if (values) std::copy(r->H.values.begin(), r->H.values.end(), values);
}
// t[0..2] = phase seconds (element eval+projection, serial accumulate, COO->CSC); n[0] = #decomposed, n[1] = #rebuilt
void oracle_result_phases(void* rp, double* t, std::int64_t* n)
{
    Result* r = (Result*)rp;
    t[0] = r->pt.eval; t[1] = r->pt.accumulate; t[2] = r->pt.compress;
    n[0] = r->pt.n_decomposed; n[1] = r->pt.n_projected;
}
void oracle_result_free(void* r) { delete (Result*)r; }

int oracle_default_threads() { return get_n_threads(EvalSettings()); }
int oracle_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Dense projection of one k x k matrix (row-major, in place). Returns the code of project_positive_definite().
int oracle_project(int k, double* H, double eps)
{
    try { return project_positive_definite(k, H, eps); }
    catch (const std::exception& e) { g_last_error = e.what(); return -1; }
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Scalar known-answer cases (mirrors tests/ScalarTest*.cc, ComplexTest.cc of the reference).
// Shared case vocabulary with the product's device test kernel: see tests/golden/scalar_cases.json.
// out = [val, grad(k), Hess(k*k)] for each returned scalar (n_out scalars), k in {1,2,6}.
// ---------------------------------------------------------------------------
namespace
{
template <int k>
void put(const Scalar<k, true>& s, double*& out)
{
    *out++ = s.val;
    for (int i = 0; i < k; ++i) *out++ = s.grad[i];
    for (int i = 0; i < k * k; ++i) *out++ = s.Hess[i];
}

int scalar_case_impl(const std::string& name, const double* p, double* out)
{
    using A1 = Scalar<1, true>;
    using A2 = Scalar<2, true>;
    auto kd = [&](int o) { return A1::known_derivatives(p[o], p[o + 1], p[o + 2]); };
    // --- unary on known_derivatives(p0,p1,p2) ---
    static const std::map<std::string, std::function<A1(const A1&)>> unary = {
        {"neg", [](const A1& a) { return -a; }}, {"sqrt", [](const A1& a) { return sqrt(a); }},
        {"sqr", [](const A1& a) { return sqr(a); }}, {"fabs", [](const A1& a) { return fabs(a); }},
        {"abs", [](const A1& a) { return abs(a); }}, {"exp", [](const A1& a) { return exp(a); }},
        {"log", [](const A1& a) { return log(a); }}, {"log2", [](const A1& a) { return log2(a); }},
        {"log10", [](const A1& a) { return log10(a); }}, {"sin", [](const A1& a) { return sin(a); }},
        {"cos", [](const A1& a) { return cos(a); }}, {"tan", [](const A1& a) { return tan(a); }},
        {"asin", [](const A1& a) { return asin(a); }}, {"acos", [](const A1& a) { return acos(a); }},
        {"atan", [](const A1& a) { return atan(a); }}, {"sinh", [](const A1& a) { return sinh(a); }},
        {"cosh", [](const A1& a) { return cosh(a); }}, {"tanh", [](const A1& a) { return tanh(a); }},
        {"asinh", [](const A1& a) { return asinh(a); }}, {"acosh", [](const A1& a) { return acosh(a); }},
        {"atanh", [](const A1& a) { return atanh(a); }},
    };
    if (auto it = unary.find(name); it != unary.end()) { put(it->second(kd(0)), out); return 1; }
    if (name == "pow_int") { put(pow(kd(0), (int)p[3]), out); return 1; }
    if (name == "pow_real") { put(pow(kd(0), p[3]), out); return 1; }
    // --- binary on a=kd(0), b=kd(3); scalar operand p[6] ---
    if (name == "add") { put(kd(0) + kd(3), out); return 1; }
    if (name == "sub") { put(kd(0) - kd(3), out); return 1; }
    if (name == "mul") { put(kd(0) * kd(3), out); return 1; }
    if (name == "div") { put(kd(0) / kd(3), out); return 1; }
    if (name == "add_s") { put(kd(0) + p[6], out); return 1; }
    if (name == "s_add") { put(p[6] + kd(0), out); return 1; }
    if (name == "sub_s") { put(kd(0) - p[6], out); return 1; }
    if (name == "s_sub") { put(p[6] - kd(0), out); return 1; }
    if (name == "mul_s") { put(kd(0) * p[6], out); return 1; }
    if (name == "s_mul") { put(p[6] * kd(0), out); return 1; }
    if (name == "div_s") { put(kd(0) / p[6], out); return 1; }
    if (name == "s_div") { put(p[6] / kd(0), out); return 1; }
    if (name == "iadd") { A1 a = kd(0); a += kd(3); put(a, out); return 1; }
    if (name == "isub") { A1 a = kd(0); a -= kd(3); put(a, out); return 1; }
    if (name == "imul") { A1 a = kd(0); a *= kd(3); put(a, out); return 1; }
    if (name == "idiv") { A1 a = kd(0); a /= kd(3); put(a, out); return 1; }
    if (name == "iadd_s") { A1 a = kd(0); a += p[6]; put(a, out); return 1; }
    if (name == "isub_s") { A1 a = kd(0); a -= p[6]; put(a, out); return 1; }
    if (name == "imul_s") { A1 a = kd(0); a *= p[6]; put(a, out); return 1; }
    if (name == "idiv_s") { A1 a = kd(0); a /= p[6]; put(a, out); return 1; }
    if (name == "min") { put(min(kd(0), kd(3)), out); return 1; }
    if (name == "max") { put(max(kd(0), kd(3)), out); return 1; }
    if (name == "clamp") { put(clamp(kd(0), kd(3), kd(6)), out); return 1; }
    if (name == "fmin") { put(fmin(kd(0), kd(3)), out); return 1; }
    if (name == "fmax") { put(fmax(kd(0), kd(3)), out); return 1; }
    if (name == "clamp_d") { put(clamp(kd(0), p[3], p[4]), out); return 1; }   // ScalarTestComparison.cc:141-160 (double bounds)
    if (name == "cmp")   // ScalarTestComparison.cc:41-108: one bit per comparison operator
    {
        const A1 a = kd(0), b = kd(3);
        const double sc = p[6];
        unsigned m = 0;
        int bit = 0;
        auto put_bit = [&](bool v) { if (v) m |= 1u << bit; ++bit; };
        put_bit(a == b); put_bit(a != b); put_bit(a < b); put_bit(a <= b); put_bit(a > b); put_bit(a >= b);
        put_bit(a == sc); put_bit(a != sc); put_bit(a < sc); put_bit(a <= sc); put_bit(a > sc); put_bit(a >= sc);
        put_bit(sc == a); put_bit(sc != a); put_bit(sc < a); put_bit(sc <= a); put_bit(sc > a); put_bit(sc >= a);
        put(A1((double)m), out);
        return 1;
    }
    if (name == "isnan_isinf")   // ScalarTestComparison.cc:12-32
    {
        const A1 v(p[0]);
        put(A1((double)((isnan(v) ? 1 : 0) | (isinf(v) ? 2 : 0) | (isfinite(v) ? 4 : 0))), out);
        return 1;
    }
    if (name == "quadratic") { A1 a(p[0], 0); put(sqr(a) + a + 2.0, out); return 1; }  // ScalarTestMisc.cc:11-26
    if (name == "atan2_1")   // ScalarTestBinaryOperators.cc:468-510
    {
        A1 x(p[0], 0);
        auto y = sqr(x) - x - 1.0;
        put(atan2(y, x), out);
        return 1;
    }
    // --- k = 2, x = (p0, idx0), y = (p1, idx1) ---
    A2 x(p[0], 0), y(p[1], 1);
    if (name == "sqr_pow_mul")  // ScalarTestUnaryOperators.cc:66-101
    {
        A2 a = x * x + 7.0 * y * y - 3.0 * x * 3.0 + x + 2 * y;
        put(sqr(a), out); put(pow(a, 2), out); put(a * a, out);
        return 3;
    }
    if (name == "atan2_const") { put(atan2(y, x), out); return 1; }
    if (name == "atan2_2")
    {
        auto a = 0.5 * sqr(x) - sqr(y) - y;
        auto b = -sqr(x - 2) - sqr(y - 3) + 1;
        put(atan2(b, a), out); put(atan(b / a), out);
        return 2;
    }
    if (name == "hypot") { put(hypot(x, y), out); return 1; }
    if (name == "div2d") { put(sqr(x) / y, out); return 1; }
    if (name == "div2d_2")
    {
        auto a = 0.5 * sqr(x) - sqr(y) + 2.0 * x - y;
        auto b = -sqr(x - 2.0) - sqr(y - 3.0) + 1.0;
        put(a / b, out);
        return 1;
    }
    if (name == "plus_minus_mult_div_2d") { put((sqr(x) + x) * (sqr(y) - y) / (y - 1.0), out); return 1; }
    if (name == "sphere")  // ScalarTestMisc.cc:38-86
    {
        put(sin(x) * cos(y), out); put(sin(x) * sin(y), out); put(cos(x), out);
        return 3;
    }
    // complex (ComplexTest.cc): a = x + i y, b = (p2 + i p3) passive or active copies
    {
        using C = std::complex<A2>;
        C a(x, y);
        std::complex<double> bd(p[2], p[3]);
        C b(A2(p[2]) + 0.5 * x, A2(p[3]) - 0.25 * y);
        auto putc = [&](const C& c) { put(c.real(), out); put(c.imag(), out); };
        if (name == "c_mul") { putc(a * b); return 2; }
        if (name == "c_mul_d") { putc(a * bd); return 2; }
        if (name == "c_d_mul") { putc(bd * a); return 2; }
        if (name == "c_div") { putc(a / b); return 2; }
        if (name == "c_div_d") { putc(a / bd); return 2; }
        if (name == "c_add") { putc(a + b); return 2; }
        if (name == "c_sub") { putc(a - b); return 2; }
        if (name == "c_sqr") { putc(sqr(a)); return 2; }
        if (name == "c_conj") { putc(conj(a)); return 2; }
        if (name == "c_abs") { put(abs(a), out); return 1; }
        if (name == "c_arg") { put(arg(a), out); return 1; }
    }
    // k = 6: symmetric Dirichlet of one triangle (ScalarTestHessianBlock.cc:50-90, ScalarTestMisc.cc:121-149)
    if (name == "symm_dirich6")
    {
        using A6 = Scalar<6, true>;
        Vec<double, 2> ar, br, cr;
        ar[0] = p[6]; ar[1] = p[7]; br[0] = p[8]; br[1] = p[9]; cr[0] = p[10]; cr[1] = p[11];
        Mat<double, 2, 2> Mr = col_mat(br - ar, cr - ar);
        Vec<A6, 2> a, b, c;
        a[0] = A6(p[0], 0); a[1] = A6(p[1], 1); b[0] = A6(p[2], 2); b[1] = A6(p[3], 3); c[0] = A6(p[4], 4); c[1] = A6(p[5], 5);
        Mat<A6, 2, 2> M = col_mat(b - a, c - a);
        Mat<A6, 2, 2> J = M * Mr.inverse();
        A6 E = J.squaredNorm() + J.inverse().squaredNorm();
        put(E, out);
        return 1;
    }
    return -1;
}
}  // namespace

extern "C" int oracle_scalar_case(const char* name, const double* params, double* out)
{
    try { return scalar_case_impl(name, params, out); }
    catch (const std::exception& e) { g_last_error = e.what(); return -2; }
}
