// TEST INFRASTRUCTURE ONLY -- C entry points over the UNMODIFIED reference headers.
//
// Compiled by `make -C oracle _ref` as
//     g++ -I oracle/eigen_shim -I /root/reference/include ref_driver.cc -> oracle/_ref/libtinyad_ref.so
// i.e. TinyAD/ScalarFunction.hh, VectorFunction.hh, Scalar.hh, Detail/*.hh, Utils/HessianProjection.hh and Operations/SVD.hh
// are taken where they lie under /root/reference; nothing of them is copied into this repository.  Eigen, which those
// headers include, is not in the image: oracle/eigen_shim/ provides the subset of its API they need (see the header of
// eigen_shim/Eigen/src/Shim.h -- eager evaluation, textbook eigen-solver: the arithmetic of TinyAD::Scalar, Element,
// the objective terms and ScalarFunction / VectorFunction is the reference's own, the dense/sparse primitives are not
// Eigen's).
//
// The element lambdas below are the ones the reference's tests use (tests/NewtonTest.cc:28-55, GaussNewtonTest.cc:34-72,
// DynamicElementsTest.cc:9-33 and 92-141, ScalarFunctionTest.cc:72-179), written the way those tests write them, on the
// term vocabulary of oracle_capi.cc so that tests/test_oracle_vs_reference.py can feed both libraries the same inputs.
#include <TinyAD/ScalarFunction.hh>
#include <TinyAD/VectorFunction.hh>
#include <TinyAD/Operations/SVD.hh>
#include <TinyAD/Utils/Helpers.hh>
#include <TinyAD/Utils/HessianProjection.hh>

#include <chrono>
#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>

extern "C" {

// same layout as oracle_capi.cc: conn is n_elements x valence (int32 handles), data is n_elements x n_data doubles
struct ref_term
{
    int kind;
    std::int64_t n_elements;
    const std::int32_t* conn;
    const double* data;
    int n_data;
};

enum
{
    REF_SYMDIRICHLET2D = 1,
    REF_PENALTY2D = 2,
    REF_SYMDIRICHLET3D = 3,
    REF_PENALTY3D = 4,
    REF_EDGE_DIRICHLET1D = 5,
    REF_QUADRATIC2D = 6,
    REF_REPEATED_HANDLE = 7,
    REF_TRIG_MIX2D = 8,
    REF_ARAP2D = 12,
    REF_DYN_SUM_SQR2D = 10,
    REF_DYN_ONERING1D = 11,
    REF_SOS_SYMDIRICHLET2D = 101,
    REF_SOS_PENALTY2D = 102,
    REF_SOS_POLYCURL2D = 103,
};

}  // extern "C"

namespace
{

template <typename FuncT>
void add_scalar_term(FuncT& func, const ref_term& t, int d)
{
    const std::int32_t* conn = t.conn;
    const double* data = t.data;
    const int nd = t.n_data;
    const auto elements = TinyAD::range(t.n_elements);
    switch (t.kind)
    {
    case REF_SYMDIRICHLET2D:  // tests/NewtonTest.cc:28-44
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 2)
            func.template add_elements<3>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const Eigen::Index e = element.handle;
                const double* dd = data + e * nd;
                Eigen::Matrix2d Mr;
                Mr << dd[0], dd[1], dd[2], dd[3];
                Eigen::Vector2<T> a = element.variables(conn[3 * e + 0]);
                Eigen::Vector2<T> b = element.variables(conn[3 * e + 1]);
                Eigen::Vector2<T> c = element.variables(conn[3 * e + 2]);
                Eigen::Matrix2<T> M = TinyAD::col_mat(b - a, c - a);
                if (M.determinant() <= 0.0) return (T)INFINITY;
                return ((M * Mr.inverse()).squaredNorm() + (Mr * M.inverse()).squaredNorm()) * dd[4];
            });
        break;
    case REF_PENALTY2D:  // tests/NewtonTest.cc:48-55
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 2)
            func.template add_elements<1>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const Eigen::Index e = element.handle;
                Eigen::Vector2d p_target(data[e * nd + 0], data[e * nd + 1]);
                Eigen::Vector2<T> p = element.variables(conn[e]);
                return (p_target - p).squaredNorm();
            });
        break;
    case REF_SYMDIRICHLET3D:
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 3)
            func.template add_elements<4>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const Eigen::Index e = element.handle;
                const double* dd = data + e * nd;
                Eigen::Matrix3d Mr_inv;
                Mr_inv << dd[0], dd[1], dd[2], dd[3], dd[4], dd[5], dd[6], dd[7], dd[8];
                Eigen::Vector3<T> a = element.variables(conn[4 * e + 0]);
                Eigen::Vector3<T> b = element.variables(conn[4 * e + 1]);
                Eigen::Vector3<T> c = element.variables(conn[4 * e + 2]);
                Eigen::Vector3<T> dv = element.variables(conn[4 * e + 3]);
                Eigen::Matrix3<T> M = TinyAD::col_mat(b - a, c - a, dv - a);
                if (M.determinant() <= 0.0) return (T)INFINITY;
                Eigen::Matrix3<T> J = M * Mr_inv;
                return (J.squaredNorm() + J.inverse().squaredNorm()) * dd[9];
            });
        break;
    case REF_PENALTY3D:
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 3)
            func.template add_elements<1>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const Eigen::Index e = element.handle;
                Eigen::Vector3d p_target(data[e * nd + 0], data[e * nd + 1], data[e * nd + 2]);
                Eigen::Vector3<T> p = element.variables(conn[e]);
                return (p_target - p).squaredNorm();
            });
        break;
    case REF_EDGE_DIRICHLET1D:  // tests/DynamicElementsTest.cc:60-91 (per-edge form)
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 1)
            func.template add_elements<2>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const Eigen::Index e = element.handle;
                T xa = element.variable(conn[2 * e + 0]);
                T xb = element.variable(conn[2 * e + 1]);
                return data[e * nd] * sqr(xa - xb);
            });
        break;
    case REF_QUADRATIC2D:  // tests/ScalarFunctionTest.cc:72-146
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 2)
            func.template add_elements<1>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const Eigen::Index e = element.handle;
                Eigen::Vector2<T> x = element.variables(conn[e]);
                return data[e * nd] * (2.0 * sqr(x[0]) + 2.0 * x[0] * x[1] + sqr(x[1]) + x[0] + 1.0);
            });
        break;
    case REF_REPEATED_HANDLE:  // tests/ScalarFunctionTest.cc:153-179
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 2)
            func.template add_elements<2>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const Eigen::Index e = element.handle;
                Eigen::Vector2<T> v = element.variables(conn[2 * e + 0]);
                Eigen::Vector2<T> v2 = element.variables(conn[2 * e + 0]);
                Eigen::Vector2<T> w = element.variables(conn[2 * e + 1]);
                return v[0] * v2[1] + sqr(w[0]) * v2[0] + w[1] * v[1] * 3.0;
            });
        break;
    case REF_TRIG_MIX2D:
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 2)
            func.template add_elements<2>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const Eigen::Index e = element.handle;
                Eigen::Vector2<T> p = element.variables(conn[2 * e + 0]);
                Eigen::Vector2<T> q = element.variables(conn[2 * e + 1]);
                T r = hypot(p[0] - q[0], p[1] - q[1]) + 0.5;
                T s = sin(p[0]) * cos(q[1]) + exp(0.25 * p[1]) / (1.0 + sqr(q[0]));
                T u = log(r) + sqrt(r + sqr(s)) + atan2(p[1] + 2.0, q[0] + 3.0);
                T v = pow(r, 3) - pow(r, 1.5) + tanh(s) * data[e * nd];
                return u * v + fabs(s - 0.1) + 2.0 / r - (1.0 - s) / 3.0;
            });
        break;
    case REF_ARAP2D:  // Operations/SVD.hh:69-99
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 2)
            func.template add_elements<3>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const Eigen::Index e = element.handle;
                const double* dd = data + e * nd;
                Eigen::Matrix2d Mr;
                Mr << dd[0], dd[1], dd[2], dd[3];
                Eigen::Vector2<T> a = element.variables(conn[3 * e + 0]);
                Eigen::Vector2<T> b = element.variables(conn[3 * e + 1]);
                Eigen::Vector2<T> c = element.variables(conn[3 * e + 2]);
                Eigen::Matrix2<T> J = TinyAD::col_mat(b - a, c - a) * Mr.inverse();
                Eigen::Matrix2<T> R = TinyAD::closest_orthogonal(J);
                return (J - R).squaredNorm() * dd[4];
            });
        break;
    case REF_DYN_SUM_SQR2D:  // tests/DynamicElementsTest.cc:9-33
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 2)
            func.template add_elements_dynamic<3, 1>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const int e = (int)element.handle;
                Eigen::Vector2<T> sum = Eigen::Vector2<T>::Zero();
                for (int v = 0; v < e; ++v) sum += element.variables(v);
                return sum.squaredNorm();
            });
        break;
    case REF_DYN_ONERING1D:  // tests/DynamicElementsTest.cc:92-141
        if constexpr ((int)FuncT::PassiveVariableVectorType::RowsAtCompileTime == 1)
            func.template add_elements_dynamic<4, 6, 7, 10>(elements, [=](auto& element) -> TINYAD_SCALAR_TYPE(element) {
                using T = TINYAD_SCALAR_TYPE(element);
                const int v = (int)element.handle;
                T v_val = element.variable(v);
                std::vector<T> neigh_vals;
                for (int i = 0; i < nd && conn[v * nd + i] >= 0; ++i) neigh_vals.push_back(element.variable(conn[v * nd + i]));
                T dirichlet = 0.0;
                for (size_t i = 0; i < neigh_vals.size(); ++i) dirichlet += 0.25 * sqr(v_val - neigh_vals[i]);
                return dirichlet;
            });
        break;
    default:
        throw std::runtime_error("ref: unknown scalar term kind");
    }
    (void)d;
}

template <typename FuncT>
void add_vector_term(FuncT& func, const ref_term& t)
{
    const std::int32_t* conn = t.conn;
    const double* data = t.data;
    const int nd = t.n_data;
    const auto elements = TinyAD::range(t.n_elements);
    switch (t.kind)
    {
    case REF_SOS_SYMDIRICHLET2D:  // tests/GaussNewtonTest.cc:34-60
        func.template add_elements<3, 8>(elements, [=](auto& element) -> TINYAD_VECTOR_TYPE(element) {
            using T = TINYAD_SCALAR_TYPE(element);
            const Eigen::Index e = element.handle;
            const double* dd = data + e * nd;
            Eigen::Matrix2d Mr;
            Mr << dd[0], dd[1], dd[2], dd[3];
            Eigen::Vector2<T> a = element.variables(conn[3 * e + 0]);
            Eigen::Vector2<T> b = element.variables(conn[3 * e + 1]);
            Eigen::Vector2<T> c = element.variables(conn[3 * e + 2]);
            Eigen::Matrix2<T> M = TinyAD::col_mat(b - a, c - a);
            if (M.determinant() <= 0.0) return Eigen::Vector<T, 8>::Constant((T)INFINITY);
            Eigen::Matrix2<T> J = M * Mr.inverse();
            Eigen::Matrix2<T> J_inv = Mr * M.inverse();
            Eigen::Vector<T, 8> E;
            E << J(0, 0), J(0, 1), J(1, 0), J(1, 1), J_inv(0, 0), J_inv(0, 1), J_inv(1, 0), J_inv(1, 1);
            return (dd[4] * E).eval();
        });
        break;
    case REF_SOS_PENALTY2D:  // tests/GaussNewtonTest.cc:64-72
        func.template add_elements<1, 2>(elements, [=](auto& element) -> TINYAD_VECTOR_TYPE(element) {
            using T = TINYAD_SCALAR_TYPE(element);
            const Eigen::Index e = element.handle;
            Eigen::Vector2d p_target(data[e * nd + 0], data[e * nd + 1]);
            Eigen::Vector2<T> p = element.variables(conn[e]);
            return (p_target - p).eval();
        });
        break;
    case REF_SOS_POLYCURL2D:  // complex residual on std::complex<TinyAD::Scalar> (Scalar.hh:1151-1320)
        func.template add_elements<2, 2>(elements, [=](auto& element) -> TINYAD_VECTOR_TYPE(element) {
            using T = TINYAD_SCALAR_TYPE(element);
            const Eigen::Index e = element.handle;
            const double* dd = data + e * nd;
            Eigen::Vector2<T> pf = element.variables(conn[2 * e + 0]);
            Eigen::Vector2<T> pg = element.variables(conn[2 * e + 1]);
            Eigen::Vector2<T> r;
            if constexpr (TINYAD_ACTIVE_MODE(element))
            {
                std::complex<T> zf(pf[0], pf[1]), zg(pg[0], pg[1]);
                std::complex<double> edge(dd[0], -dd[1]);
                std::complex<T> c = (sqr(sqr(zf)) - sqr(sqr(zg))) * edge;
                r[0] = dd[2] * c.real();
                r[1] = dd[2] * c.imag();
            }
            else
            {
                std::complex<double> zf(pf[0], pf[1]), zg(pg[0], pg[1]);
                std::complex<double> edge(dd[0], -dd[1]);
                auto csqr = [](const std::complex<double>& a) { return std::complex<double>(a.real() * a.real() - a.imag() * a.imag(), 2.0 * a.real() * a.imag()); };
                std::complex<double> df = csqr(csqr(zf)), dg = csqr(csqr(zg));
                std::complex<double> dz(df.real() - dg.real(), df.imag() - dg.imag());
                std::complex<double> c(dz.real() * edge.real() - dz.imag() * edge.imag(), dz.real() * edge.imag() + dz.imag() * edge.real());
                r[0] = dd[2] * c.real();
                r[1] = dd[2] * c.imag();
            }
            return r;
        });
        break;
    default:
        throw std::runtime_error("ref: unknown vector term kind");
    }
}

thread_local std::string g_last_error;

struct Result
{
    double f = 0.0;
    Eigen::VectorXd g, r;
    Eigen::SparseMatrix<double> H;  // Hessian or Jacobian
    double seconds = 0.0;           // the eval* call alone
    // VectorFunction::eval_with_derivatives: the Hessian of every residual as (residual, row, col, value) entries
    std::vector<std::int32_t> hess_res, hess_row, hess_col;
    std::vector<double> hess_val;
};

template <int d>
void scalar_eval_impl(std::int64_t n_vertices, int n_terms, const ref_term* terms, int mode, const double* x_in, double eps,
                      int n_threads, Result& out)
{
    TinyAD::EvalSettings settings;
    settings.n_threads = n_threads;
    auto func = TinyAD::scalar_function<d>(TinyAD::range(n_vertices), settings);
    for (int i = 0; i < n_terms; ++i) add_scalar_term(func, terms[i], d);
    Eigen::VectorXd x = Eigen::Map<const Eigen::VectorXd>(x_in, func.n_vars);
    const auto t0 = std::chrono::steady_clock::now();
    switch (mode)
    {
    case 0: out.f = func.eval(x); break;
    case 1: func.eval_with_gradient(x, out.f, out.g); break;
    case 2: func.eval_with_derivatives(x, out.f, out.g, out.H); break;
    case 3: func.eval_with_hessian_proj(x, out.f, out.g, out.H, eps); break;
    default: throw std::runtime_error("ref: bad mode");
    }
    out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

template <int k>
int project_impl(double* H, double eps)
{
    Eigen::Matrix<double, k, k> M;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) M(i, j) = H[i * k + j];
    TinyAD::project_positive_definite<k, double>(M, eps);
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) H[i * k + j] = M(i, j);
    return 0;
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_last_error.c_str(); }

// which reference tree and which linear-algebra layer this library was built from
const char* ref_description()
{
#ifdef TINYAD_EIGEN_SHIM
    return "TinyAD headers (unmodified, compiled in place) over oracle/eigen_shim (not Eigen)";
#else
    return "TinyAD headers (unmodified, compiled in place) over Eigen";
#endif
}

// mode: 0 eval, 1 eval_with_gradient, 2 eval_with_derivatives, 3 eval_with_hessian_proj (Detail/ScalarFunctionImpl.hh:256-416)
void* ref_scalar_eval(int d, std::int64_t n_vertices, int n_terms, const ref_term* terms, int mode, const double* x, double eps,
                      int n_threads)
{
    auto res = std::make_unique<Result>();
    try
    {
        if (d == 1) scalar_eval_impl<1>(n_vertices, n_terms, terms, mode, x, eps, n_threads, *res);
        else if (d == 2) scalar_eval_impl<2>(n_vertices, n_terms, terms, mode, x, eps, n_threads, *res);
        else if (d == 3) scalar_eval_impl<3>(n_vertices, n_terms, terms, mode, x, eps, n_threads, *res);
        else throw std::runtime_error("ref: unsupported variable dimension");
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return nullptr;
    }
    return res.release();
}

// mode: 0 eval, 1 eval_with_jacobian, 2 eval_sum_of_squares, 3 eval_sum_of_squares_with_derivatives, 4 eval_with_derivatives
// (Detail/VectorFunctionImpl.hh:143-301)
void* ref_vector_eval(int d, std::int64_t n_vertices, int n_terms, const ref_term* terms, int mode, const double* x_in, int n_threads)
{
    auto res = std::make_unique<Result>();
    try
    {
        if (d != 2) throw std::runtime_error("ref: vector functions are instantiated for d=2 only");
        TinyAD::EvalSettings settings;
        settings.n_threads = n_threads;
        auto func = TinyAD::vector_function<2>(TinyAD::range(n_vertices), settings);
        for (int i = 0; i < n_terms; ++i) add_vector_term(func, terms[i]);
        Eigen::VectorXd x = Eigen::Map<const Eigen::VectorXd>(x_in, func.n_vars);
        const auto t0 = std::chrono::steady_clock::now();
        switch (mode)
        {
        case 0: res->r = func.eval(x); break;
        case 1: func.eval_with_jacobian(x, res->r, res->H); break;
        case 2: res->f = func.eval_sum_of_squares(x); break;
        case 3: func.eval_sum_of_squares_with_derivatives(x, res->f, res->g, res->r, res->H); break;
        case 4:  // Detail/VectorFunctionImpl.hh:207-235: r, J and one n x n sparse Hessian per residual
        {
            std::vector<Eigen::SparseMatrix<double>> H;
            func.eval_with_derivatives(x, res->r, res->H, H);
            for (std::size_t i = 0; i < H.size(); ++i)
                for (Eigen::Index j = 0; j < H[i].cols(); ++j)
                    for (Eigen::SparseMatrix<double>::InnerIterator it(H[i], j); it; ++it)
                    {
                        res->hess_res.push_back((std::int32_t)i);
                        res->hess_row.push_back((std::int32_t)it.row());
                        res->hess_col.push_back((std::int32_t)it.col());
                        res->hess_val.push_back(it.value());
                    }
            break;
        }
        default: throw std::runtime_error("ref: bad mode");
        }
        res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return nullptr;
    }
    return res.release();
}

double ref_result_f(void* r) { return ((Result*)r)->f; }
double ref_result_seconds(void* r) { return ((Result*)r)->seconds; }
std::int64_t ref_result_nnz(void* r) { return ((Result*)r)->H.nonZeros(); }
std::int64_t ref_result_rows(void* r) { return ((Result*)r)->H.rows(); }
std::int64_t ref_result_cols(void* r) { return ((Result*)r)->H.cols(); }
std::int64_t ref_result_g_size(void* r) { return ((Result*)r)->g.size(); }
std::int64_t ref_result_r_size(void* r) { return ((Result*)r)->r.size(); }
void ref_result_copy(void* rp, double* g, double* rvec, std::int32_t* outer, std::int32_t* inner, double* values)
{
    Result* r = (Result*)rp;
    if (g) std::copy(r->g.data(), r->g.data() + r->g.size(), g);
    if (rvec) std::copy(r->r.data(), r->r.data() + r->r.size(), rvec);
    if (outer) std::copy(r->H.outerIndexPtr(), r->H.outerIndexPtr() + r->H.cols() + 1, outer);
    if (inner) std::copy(r->H.innerIndexPtr(), r->H.innerIndexPtr() + r->H.nonZeros(), inner);
    if (values) std::copy(r->H.valuePtr(), r->H.valuePtr() + r->H.nonZeros(), values);
}
std::int64_t ref_result_hess_count(void* r) { return (std::int64_t)((Result*)r)->hess_val.size(); }
void ref_result_hess_copy(void* rp, std::int32_t* res, std::int32_t* row, std::int32_t* col, double* val)
{
    Result* r = (Result*)rp;
    std::copy(r->hess_res.begin(), r->hess_res.end(), res);
    std::copy(r->hess_row.begin(), r->hess_row.end(), row);
    std::copy(r->hess_col.begin(), r->hess_col.end(), col);
    std::copy(r->hess_val.begin(), r->hess_val.end(), val);
}
void ref_result_free(void* r) { delete (Result*)r; }

int ref_default_threads() { return TinyAD::get_n_threads(TinyAD::EvalSettings()); }

// TinyAD::project_positive_definite (Utils/HessianProjection.hh:48-101) on one k x k matrix (row-major, in place)
int ref_project(int k, double* H, double eps)
{
    try
    {
        switch (k)
        {
        case 2: return project_impl<2>(H, eps);
        case 3: return project_impl<3>(H, eps);
        case 4: return project_impl<4>(H, eps);
        case 6: return project_impl<6>(H, eps);
        case 8: return project_impl<8>(H, eps);
        case 9: return project_impl<9>(H, eps);
        case 12: return project_impl<12>(H, eps);
        case 16: return project_impl<16>(H, eps);
        default: throw std::runtime_error("ref: project is instantiated for k in {2,3,4,6,8,9,12,16}");
        }
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return -1;
    }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Scalar cases on the reference's TinyAD::Scalar: the case vocabulary of oracle_capi.cc / tests/golden/scalar_cases.json
// (tests/ScalarTest*.cc, ComplexTest.cc of the reference), so that the oracle's and the product's Scalar can be compared
// with the reference's on arbitrary parameters, not only on the transcribed golden ones.
// out = [val, grad(k), Hess(k*k) row-major] per returned scalar.
// ---------------------------------------------------------------------------------------------------------------------
namespace
{

template <int k>
void put(const TinyAD::Double<k>& s, double*& out)
{
    *out++ = s.val;
    for (int i = 0; i < k; ++i) *out++ = s.grad(i);
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) *out++ = s.Hess(i, j);
}

int scalar_case_impl(const std::string& name, const double* p, double* out)
{
    using A1 = TinyAD::Double<1>;
    using A2 = TinyAD::Double<2>;
    auto kd = [&](int o) { return A1::known_derivatives(p[o], p[o + 1], p[o + 2]); };
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
    if (name == "clamp_d") { put(clamp(kd(0), p[3], p[4]), out); return 1; }
    if (name == "cmp")
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
    if (name == "isnan_isinf")
    {
        const A1 v(p[0]);
        put(A1((double)((isnan(v) ? 1 : 0) | (isinf(v) ? 2 : 0) | (isfinite(v) ? 4 : 0))), out);
        return 1;
    }
    if (name == "quadratic") { A1 a(p[0], 0); put(sqr(a) + a + 2.0, out); return 1; }
    if (name == "atan2_1")
    {
        A1 x(p[0], 0);
        A1 y = sqr(x) - x - 1.0;
        put(atan2(y, x), out);
        return 1;
    }
    A2 x(p[0], 0), y(p[1], 1);
    if (name == "sqr_pow_mul")
    {
        A2 a = x * x + 7.0 * y * y - 3.0 * x * 3.0 + x + 2 * y;
        put(sqr(a), out); put(pow(a, 2), out); put(a * a, out);
        return 3;
    }
    if (name == "atan2_const") { put(atan2(y, x), out); return 1; }
    if (name == "atan2_2")
    {
        A2 a = 0.5 * sqr(x) - sqr(y) - y;
        A2 b = -sqr(x - 2) - sqr(y - 3) + 1;
        put(atan2(b, a), out); put(atan(b / a), out);
        return 2;
    }
    if (name == "hypot") { put(hypot(x, y), out); return 1; }
    if (name == "div2d") { put(sqr(x) / y, out); return 1; }
    if (name == "div2d_2")
    {
        A2 a = 0.5 * sqr(x) - sqr(y) + 2.0 * x - y;
        A2 b = -sqr(x - 2.0) - sqr(y - 3.0) + 1.0;
        put(a / b, out);
        return 1;
    }
    if (name == "plus_minus_mult_div_2d") { put((sqr(x) + x) * (sqr(y) - y) / (y - 1.0), out); return 1; }
    if (name == "sphere")
    {
        put(sin(x) * cos(y), out); put(sin(x) * sin(y), out); put(cos(x), out);
        return 3;
    }
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
    if (name == "symm_dirich6")  // tests/ScalarTestHessianBlock.cc:50-90
    {
        using A6 = TinyAD::Double<6>;
        const Eigen::Vector2d ar(p[6], p[7]), br(p[8], p[9]), cr(p[10], p[11]);
        const Eigen::Matrix2d Mr = TinyAD::col_mat(br - ar, cr - ar);
        Eigen::Matrix<double, Eigen::Dynamic, 1> xv(6);
        for (int i = 0; i < 6; ++i) xv[i] = p[i];
        const Eigen::Matrix<A6, 6, 1> xa = A6::make_active(xv);
        const Eigen::Vector2<A6> a(xa[0], xa[1]), b(xa[2], xa[3]), c(xa[4], xa[5]);
        const Eigen::Matrix2<A6> M = TinyAD::col_mat(b - a, c - a);
        const Eigen::Matrix2<A6> J = M * Mr.inverse();
        A6 E = J.squaredNorm() + J.inverse().squaredNorm();
        put(E, out);
        return 1;
    }
    return -1;
}

}  // namespace

extern "C" int ref_scalar_case(const char* name, const double* params, double* out)
{
    try { return scalar_case_impl(name, params, out); }
    catch (const std::exception& e) { g_last_error = e.what(); return -2; }
}
