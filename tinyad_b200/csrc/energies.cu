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

#include <TinyAD/ScalarFunction.hh>
#include <TinyAD/VectorFunction.hh>

using namespace TinyAD;

namespace
{

thread_local std::string g_err;

// per-element data, structure of arrays: column j of element e at p[j * stride + e]
struct ConnView { const int32_t* p; int64_t stride; TINYAD_HD int32_t operator()(int64_t e, int j) const { return p[j * stride + e]; } };
struct DataView { const double* p; int64_t stride; TINYAD_HD double operator()(int64_t e, int j) const { return p[j * stride + e]; } };

struct SymDirichlet2D  // data: Mr(0,0) Mr(0,1) Mr(1,0) Mr(1,1) w
{
    ConnView F; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        Mat<double, 2, 2> Mr;
        Mr(0, 0) = D(e, 0); Mr(0, 1) = D(e, 1); Mr(1, 0) = D(e, 2); Mr(1, 1) = D(e, 3);
        Vec<T, 2> a = element.variables(F(e, 0));
        Vec<T, 2> b = element.variables(F(e, 1));
        Vec<T, 2> c = element.variables(F(e, 2));
        Mat<T, 2, 2> M = col_mat(b - a, c - a);
        if (M.determinant() <= 0.0) return (T)INFINITY;
        return ((M * Mr.inverse()).squaredNorm() + (Mr * M.inverse()).squaredNorm()) * D(e, 4);
    }
};

template <int d>
struct Penalty  // data: target (d)
{
    ConnView B; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        Vec<double, d> p_target;
        for (int i = 0; i < d; ++i) p_target[i] = D(e, i);
        Vec<T, d> p = element.variables(B(e, 0));
        return (p_target - p).squaredNorm();
    }
};

struct SymDirichlet3D  // data: Mr^-1 row-major (9), vol
{
    ConnView Tt; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        Mat<double, 3, 3> Mr_inv;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Mr_inv(i, j) = D(e, 3 * i + j);
        Vec<T, 3> a = element.variables(Tt(e, 0));
        Vec<T, 3> b = element.variables(Tt(e, 1));
        Vec<T, 3> c = element.variables(Tt(e, 2));
        Vec<T, 3> dd = element.variables(Tt(e, 3));
        Mat<T, 3, 3> M = col_mat(b - a, c - a, dd - a);
        if (M.determinant() <= 0.0) return (T)INFINITY;
        Mat<T, 3, 3> J = M * Mr_inv;
        return (J.squaredNorm() + J.inverse().squaredNorm()) * D(e, 9);
    }
};

struct EdgeDirichlet1D  // data: w ; w * (x_a - x_b)^2
{
    ConnView C; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        T xa = element.variable(C(e, 0));
        T xb = element.variable(C(e, 1));
        return D(e, 0) * sqr(xa - xb);
    }
};

struct Quadratic2D  // ScalarFunctionTest.cc:72-146, data: sign
{
    ConnView C; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        Vec<T, 2> x = element.variables(C(e, 0));
        return D(e, 0) * (2.0 * sqr(x[0]) + 2.0 * x[0] * x[1] + sqr(x[1]) + x[0] + 1.0);
    }
};

struct RepeatedHandle  // ScalarFunctionTest.cc:153-179: same handle requested twice
{
    ConnView C; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        Vec<T, 2> v = element.variables(C(e, 0));
        Vec<T, 2> v2 = element.variables(C(e, 0));
        Vec<T, 2> w = element.variables(C(e, 1));
        return v[0] * v2[1] + sqr(w[0]) * v2[0] + w[1] * v[1] * 3.0;
    }
};

struct TrigMix2D
{
    ConnView C; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        Vec<T, 2> p = element.variables(C(e, 0));
        Vec<T, 2> q = element.variables(C(e, 1));
        T r = hypot(p[0] - q[0], p[1] - q[1]) + 0.5;
        T s = sin(p[0]) * cos(q[1]) + exp(0.25 * p[1]) / (1.0 + sqr(q[0]));
        T u = log(r) + sqrt(r + sqr(s)) + atan2(p[1] + 2.0, q[0] + 3.0);
        T v = pow(r, 3) - pow(r, 1.5) + tanh(s) * D(e, 0);
        return u * v + fabs(s - 0.1) + 2.0 / r - (1.0 - s) / 3.0;
    }
};

struct SosSymDirichlet2D  // GaussNewtonTest.cc:34-62, data: Mr (4), scale
{
    ConnView F; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_VECTOR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        Mat<double, 2, 2> Mr;
        Mr(0, 0) = D(e, 0); Mr(0, 1) = D(e, 1); Mr(1, 0) = D(e, 2); Mr(1, 1) = D(e, 3);
        Vec<T, 2> a = element.variables(F(e, 0));
        Vec<T, 2> b = element.variables(F(e, 1));
        Vec<T, 2> c = element.variables(F(e, 2));
        Mat<T, 2, 2> M = col_mat(b - a, c - a);
        if (M.determinant() <= 0.0) return Vec<T, 8>::Constant((T)INFINITY);
        Mat<T, 2, 2> J = M * Mr.inverse();
        Mat<T, 2, 2> J_inv = Mr * M.inverse();
        Vec<T, 8> Ev;
        Ev[0] = J(0, 0); Ev[1] = J(0, 1); Ev[2] = J(1, 0); Ev[3] = J(1, 1);
        Ev[4] = J_inv(0, 0); Ev[5] = J_inv(0, 1); Ev[6] = J_inv(1, 0); Ev[7] = J_inv(1, 1);
        return D(e, 4) * Ev;
    }
};

struct SosPenalty2D  // GaussNewtonTest.cc:63-70
{
    ConnView B; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_VECTOR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        Vec<double, 2> p_target(D(e, 0), D(e, 1));
        Vec<T, 2> p = element.variables(B(e, 0));
        return p_target - p;
    }
};

struct SosPolycurl2D  // synthetic polycurl-style complex residual (config C4 stand-in), data: ex ey w
{
    ConnView C; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_VECTOR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        Vec<T, 2> pf = element.variables(C(e, 0));
        Vec<T, 2> pg = element.variables(C(e, 1));
        Complex<T> zf(pf[0], pf[1]), zg(pg[0], pg[1]);
        Complex<double> edge(D(e, 0), -D(e, 1));
        Complex<T> c = (sqr(sqr(zf)) - sqr(sqr(zg))) * edge;
        Vec<T, 2> r;
        r[0] = D(e, 2) * c.real();
        r[1] = D(e, 2) * c.imag();
        return r;
    }
};

enum
{
    K_SYMDIRICHLET2D = 1, K_PENALTY2D = 2, K_SYMDIRICHLET3D = 3, K_PENALTY3D = 4, K_EDGE_DIRICHLET1D = 5,
    K_QUADRATIC2D = 6, K_REPEATED_HANDLE = 7, K_TRIG_MIX2D = 8,
    K_SOS_SYMDIRICHLET2D = 101, K_SOS_PENALTY2D = 102, K_SOS_POLYCURL2D = 103,
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
    std::vector<std::unique_ptr<DeviceArray>> arrays;
    tad_function handle() const
    {
        if (s1) return s1->handle();
        if (s2) return s2->handle();
        if (s3) return s3->handle();
        if (v2) return v2->handle();
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
            if (d != 2) throw std::runtime_error("vector functions are instantiated for d = 2 only");
            P->v2 = std::make_unique<VectorFunction<2>>(vector_function<2>(range(n_vertices), s));
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
        case K_QUADRATIC2D: need(P->s2 && valence == 1 && n_data == 1); P->s2->add_elements<1>(els, Quadratic2D{C, D}); break;
        case K_REPEATED_HANDLE: need(P->s2 && valence == 2 && n_data == 1); P->s2->add_elements<2>(els, RepeatedHandle{C, D}); break;
        case K_TRIG_MIX2D: need(P->s2 && valence == 2 && n_data == 1); P->s2->add_elements<2>(els, TrigMix2D{C, D}); break;
        case K_SOS_SYMDIRICHLET2D: need(P->v2 && valence == 3 && n_data == 5); P->v2->add_elements<3, 8>(els, SosSymDirichlet2D{C, D}); break;
        case K_SOS_PENALTY2D: need(P->v2 && valence == 1 && n_data == 2); P->v2->add_elements<1, 2>(els, SosPenalty2D{C, D}); break;
        case K_SOS_POLYCURL2D: need(P->v2 && valence == 2 && n_data == 3); P->v2->add_elements<2, 2>(els, SosPolycurl2D{C, D}); break;
        default: throw std::runtime_error("unknown term kind");
        }
    });
}

}  // extern "C"
