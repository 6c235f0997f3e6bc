// tinyad_b200 -- element functors of the tests and the benchmark (shared by the translation units that
// instantiate their kernels).  See energies.cu for what each one restates.
#pragma once

#include <TinyAD/ScalarFunction.hh>
#include <TinyAD/VectorFunction.hh>

#ifndef TADX_TET_PARTS
#define TADX_TET_PARTS 4
#endif

namespace tadx
{
using namespace TinyAD;

// per-element data, structure of arrays: column j of element e at p[j * stride + e]
struct ConnView { const int32_t* p; int64_t stride; TINYAD_HD int32_t operator()(int64_t e, int j) const { return p[j * stride + e]; } };
struct DataView { const double* p; int64_t stride; TINYAD_HD double operator()(int64_t e, int j) const { return p[j * stride + e]; } };

struct SymDirichlet2D  // data: Mr(0,0) Mr(0,1) Mr(1,0) Mr(1,1) w
{
    static constexpr bool tinyad_unique_handles = true;
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
    static constexpr int tinyad_parts = TADX_TET_PARTS;      // Hessian parts (one kernel each)
    static constexpr bool tinyad_unique_handles = true;      // a tet never names a vertex twice
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


}  // namespace tadx
