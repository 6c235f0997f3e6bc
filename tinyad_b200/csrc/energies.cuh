// tinyad_b200 -- element functors of the tests and the benchmark (shared by the translation units that
// instantiate their kernels).  See energies.cu for what each one restates.
#pragma once

#include <TinyAD/ScalarFunction.hh>
#include <TinyAD/VectorFunction.hh>
#include <TinyAD/Operations/SVD.hh>

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

struct Arap2D  // data: Mr(0,0) Mr(0,1) Mr(1,0) Mr(1,1) w;  w |J - closest_orthogonal(J)|^2 with Operations/SVD.hh on active scalars
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
        Mat<T, 2, 2> J = col_mat(b - a, c - a) * Mr.inverse();
        Mat<T, 2, 2> R = closest_orthogonal(J);
        return (J - R).squaredNorm() * D(e, 4);
    }
};

// ---- dynamic-valence elements (add_elements_dynamic, tests/DynamicElementsTest.cc) ----
struct DynSumSqr2D  // DynamicElementsTest.cc:9-33: element e accesses the handles 0 .. e-1
{
    ConnView C; DataView D;  // unused
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int e = (int)element.handle;
        Vec<T, 2> sum;
        sum[0] = T(0.0);
        sum[1] = T(0.0);
        for (int v = 0; v < e; ++v) sum = sum + element.variables(v);
        return sum.squaredNorm();
    }
};

struct OneRingDirichlet1D  // DynamicElementsTest.cc:112-131: conn = padded neighbour table (W columns, -1 = none)
{
    ConnView Nb; DataView D;  // D unused
    int W;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t v = element.handle;
        T v_val = element.variable(v);
        T dirichlet = 0.0;
        for (int i = 0; i < W; ++i)
        {
            const int32_t nb = Nb(v, i);
            if (nb < 0) break;
            dirichlet = dirichlet + 0.25 * sqr(v_val - element.variable(nb));
        }
        return dirichlet;
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

// A functor whose variables() calls depend on x (it returns before touching its second handle when x_a > data): the recorded
// pattern (x = 0 at add_elements time) does not cover such an evaluation -> TAD_PATTERN_MISMATCH.
struct BranchOnX1D  // data: threshold
{
    ConnView C; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        const int64_t e = element.handle;
        T a = element.variable(C(e, 0));
        if (a > D(e, 0)) return sqr(a);
        T b = element.variable(C(e, 1));
        return sqr(a - b);
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

struct Sqrt1D  // ExceptionTest analogue: sqrt(x) has a NaN derivative for x < 0 -> TAD_NONFINITE_DERIVATIVE; data unused
{
    ConnView C; DataView D;
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_SCALAR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        T x = element.variable(C(element.handle, 0));
        return sqrt(x) * D(element.handle, 0);
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

// tests/VectorFunctionTest.cc:73-93 (test_eval): R^2 -> R^3, element A returns (2 x0, x0^2), element B returns (x1^2)
struct SosTest1DA
{
    ConnView C; DataView D;  // D unused
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_VECTOR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        T x0 = element.variables(C(element.handle, 0))[0];
        Vec<T, 2> r;
        r[0] = 2.0 * x0;
        r[1] = sqr(x0);
        return r;
    }
};
struct SosTest1DB
{
    ConnView C; DataView D;  // D unused
    template <class E>
    TINYAD_HD auto operator()(E& element) const -> TINYAD_VECTOR_TYPE(element)
    {
        using T = TINYAD_SCALAR_TYPE(element);
        T x1 = element.variables(C(element.handle, 0))[0];
        Vec<T, 1> r;
        r[0] = sqr(x1);
        return r;
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


// ---------------------------------------------------------------------------------------------
// Known-answer cases of the reference's Scalar tests (tests/ScalarTest*.cc, ComplexTest.cc) on the product's
// Scalar.  Same case vocabulary as oracle_capi.cc / tests/golden/scalar_cases.json; runs on host and device.
// out = [val, grad(k), Hess(k*k)] per returned scalar.
// ---------------------------------------------------------------------------------------------
enum ScalarCase
{
    SC_NEG, SC_SQRT, SC_SQR, SC_FABS, SC_ABS, SC_EXP, SC_LOG, SC_LOG2, SC_LOG10, SC_SIN, SC_COS, SC_TAN, SC_ASIN, SC_ACOS, SC_ATAN,
    SC_SINH, SC_COSH, SC_TANH, SC_ASINH, SC_ACOSH, SC_ATANH, SC_POW_INT, SC_POW_REAL,
    SC_ADD, SC_SUB, SC_MUL, SC_DIV, SC_ADD_S, SC_S_ADD, SC_SUB_S, SC_S_SUB, SC_MUL_S, SC_S_MUL, SC_DIV_S, SC_S_DIV,
    SC_IADD, SC_ISUB, SC_IMUL, SC_IDIV, SC_IADD_S, SC_ISUB_S, SC_IMUL_S, SC_IDIV_S, SC_MIN, SC_MAX, SC_CLAMP, SC_QUADRATIC, SC_ATAN2_1,
    SC_SQR_POW_MUL, SC_ATAN2_CONST, SC_ATAN2_2, SC_HYPOT, SC_DIV2D, SC_DIV2D_2, SC_PMMD_2D, SC_SPHERE,
    SC_C_MUL, SC_C_MUL_D, SC_C_D_MUL, SC_C_DIV, SC_C_DIV_D, SC_C_ADD, SC_C_SUB, SC_C_SQR, SC_C_CONJ, SC_C_ABS, SC_C_ARG, SC_SYMM_DIRICH6,
    SC_SVD2, SC_CLOSEST_ORTHOGONAL2,
    SC_FMIN, SC_FMAX, SC_CLAMP_D, SC_CMP, SC_ISNAN_ISINF,   // tests/ScalarTestComparison.cc
    SC_HESS_BLOCK_ISSUE13, SC_HESS_BLOCK_SYMDIR,             // tests/ScalarTestHessianBlock.cc
    SC_COUNT
};

template <int k, int NP, int P>
TINYAD_HD inline void sc_put(const Scalar<k, true, NP, P>& s, double*& out)
{
    *out++ = s.val;
    for (int i = 0; i < k; ++i) *out++ = s.grad[i];
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) *out++ = s.Hess(i, j);
}

TINYAD_HD inline int scalar_case_run(int id, const double* p, double* out)
{
    using A1 = Scalar<1, true>;
    using A2 = Scalar<2, true>;
    const A1 a = A1::known_derivatives(p[0], p[1], p[2]);
    const A1 b = A1::known_derivatives(p[3], p[4], p[5]);
    const double sc = p[6];
    switch (id)
    {
    case SC_NEG: sc_put(-a, out); return 1;
    case SC_SQRT: sc_put(sqrt(a), out); return 1;
    case SC_SQR: sc_put(sqr(a), out); return 1;
    case SC_FABS: sc_put(fabs(a), out); return 1;
    case SC_ABS: sc_put(abs(a), out); return 1;
    case SC_EXP: sc_put(exp(a), out); return 1;
    case SC_LOG: sc_put(log(a), out); return 1;
    case SC_LOG2: sc_put(log2(a), out); return 1;
    case SC_LOG10: sc_put(log10(a), out); return 1;
    case SC_SIN: sc_put(sin(a), out); return 1;
    case SC_COS: sc_put(cos(a), out); return 1;
    case SC_TAN: sc_put(tan(a), out); return 1;
    case SC_ASIN: sc_put(asin(a), out); return 1;
    case SC_ACOS: sc_put(acos(a), out); return 1;
    case SC_ATAN: sc_put(atan(a), out); return 1;
    case SC_SINH: sc_put(sinh(a), out); return 1;
    case SC_COSH: sc_put(cosh(a), out); return 1;
    case SC_TANH: sc_put(tanh(a), out); return 1;
    case SC_ASINH: sc_put(asinh(a), out); return 1;
    case SC_ACOSH: sc_put(acosh(a), out); return 1;
    case SC_ATANH: sc_put(atanh(a), out); return 1;
    case SC_POW_INT: sc_put(pow(a, (int)p[3]), out); return 1;
    case SC_POW_REAL: sc_put(pow(a, p[3]), out); return 1;
    case SC_ADD: sc_put(a + b, out); return 1;
    case SC_SUB: sc_put(a - b, out); return 1;
    case SC_MUL: sc_put(a * b, out); return 1;
    case SC_DIV: sc_put(a / b, out); return 1;
    case SC_ADD_S: sc_put(a + sc, out); return 1;
    case SC_S_ADD: sc_put(sc + a, out); return 1;
    case SC_SUB_S: sc_put(a - sc, out); return 1;
    case SC_S_SUB: sc_put(sc - a, out); return 1;
    case SC_MUL_S: sc_put(a * sc, out); return 1;
    case SC_S_MUL: sc_put(sc * a, out); return 1;
    case SC_DIV_S: sc_put(a / sc, out); return 1;
    case SC_S_DIV: sc_put(sc / a, out); return 1;
    case SC_IADD: { A1 t = a; t += b; sc_put(t, out); return 1; }
    case SC_ISUB: { A1 t = a; t -= b; sc_put(t, out); return 1; }
    case SC_IMUL: { A1 t = a; t *= b; sc_put(t, out); return 1; }
    case SC_IDIV: { A1 t = a; t /= b; sc_put(t, out); return 1; }
    case SC_IADD_S: { A1 t = a; t += sc; sc_put(t, out); return 1; }
    case SC_ISUB_S: { A1 t = a; t -= sc; sc_put(t, out); return 1; }
    case SC_IMUL_S: { A1 t = a; t *= sc; sc_put(t, out); return 1; }
    case SC_IDIV_S: { A1 t = a; t /= sc; sc_put(t, out); return 1; }
    case SC_MIN: sc_put(min(a, b), out); return 1;
    case SC_MAX: sc_put(max(a, b), out); return 1;
    case SC_CLAMP: sc_put(clamp(a, b, A1::known_derivatives(p[6], p[7], p[8])), out); return 1;
    case SC_FMIN: sc_put(fmin(a, b), out); return 1;
    case SC_FMAX: sc_put(fmax(a, b), out); return 1;
    case SC_CLAMP_D: sc_put(clamp(a, p[3], p[4]), out); return 1;   // double bounds convert to passive scalars (ScalarTestComparison.cc:141-160)
    case SC_CMP:  // every comparison operator of ScalarTestComparison.cc:41-108 as one bit of the returned value
    {
        unsigned m = 0;
        int bit = 0;
        auto put_bit = [&](bool v) { if (v) m |= 1u << bit; ++bit; };
        put_bit(a == b); put_bit(a != b); put_bit(a < b); put_bit(a <= b); put_bit(a > b); put_bit(a >= b);
        put_bit(a == sc); put_bit(a != sc); put_bit(a < sc); put_bit(a <= sc); put_bit(a > sc); put_bit(a >= sc);
        put_bit(sc == a); put_bit(sc != a); put_bit(sc < a); put_bit(sc <= a); put_bit(sc > a); put_bit(sc >= a);
        sc_put(A1((double)m), out);
        return 1;
    }
    case SC_ISNAN_ISINF:  // ScalarTestComparison.cc:12-32: bit 0 isnan, bit 1 isinf, bit 2 isfinite of a passive scalar
    {
        const A1 v(p[0]);
        sc_put(A1((double)((isnan(v) ? 1 : 0) | (isinf(v) ? 2 : 0) | (isfinite(v) ? 4 : 0))), out);
        return 1;
    }
    case SC_QUADRATIC: { A1 x(p[0], 0); sc_put(sqr(x) + x + 2.0, out); return 1; }
    case SC_ATAN2_1: { A1 x(p[0], 0); A1 y = sqr(x) - x - 1.0; sc_put(atan2(y, x), out); return 1; }
    default: break;
    }
    const A2 x(p[0], 0), y(p[1], 1);
    switch (id)
    {
    case SC_SQR_POW_MUL:
    {
        A2 q = x * x + 7.0 * y * y - 3.0 * x * 3.0 + x + 2.0 * y;
        sc_put(sqr(q), out); sc_put(pow(q, 2), out); sc_put(q * q, out);
        return 3;
    }
    case SC_ATAN2_CONST: sc_put(atan2(y, x), out); return 1;
    case SC_ATAN2_2:
    {
        A2 u = 0.5 * sqr(x) - sqr(y) - y;
        A2 v = -sqr(x - 2.0) - sqr(y - 3.0) + 1.0;
        sc_put(atan2(v, u), out); sc_put(atan(v / u), out);
        return 2;
    }
    case SC_HYPOT: sc_put(hypot(x, y), out); return 1;
    case SC_DIV2D: sc_put(sqr(x) / y, out); return 1;
    case SC_DIV2D_2:
    {
        A2 u = 0.5 * sqr(x) - sqr(y) + 2.0 * x - y;
        A2 v = -sqr(x - 2.0) - sqr(y - 3.0) + 1.0;
        sc_put(u / v, out);
        return 1;
    }
    case SC_PMMD_2D: sc_put((sqr(x) + x) * (sqr(y) - y) / (y - 1.0), out); return 1;
    case SC_SPHERE: sc_put(sin(x) * cos(y), out); sc_put(sin(x) * sin(y), out); sc_put(cos(x), out); return 3;
    default: break;
    }
    {
        using C = Complex<A2>;
        const C ca(x, y);
        const Complex<double> cd(p[2], p[3]);
        const C cb(A2(p[2]) + 0.5 * x, A2(p[3]) - 0.25 * y);
        switch (id)
        {
        case SC_C_MUL: { auto r = ca * cb; sc_put(r.re, out); sc_put(r.im, out); return 2; }
        case SC_C_MUL_D: { auto r = ca * cd; sc_put(r.re, out); sc_put(r.im, out); return 2; }
        case SC_C_D_MUL: { auto r = cd * ca; sc_put(r.re, out); sc_put(r.im, out); return 2; }
        case SC_C_DIV: { auto r = ca / cb; sc_put(r.re, out); sc_put(r.im, out); return 2; }
        case SC_C_DIV_D: { auto r = ca / cd; sc_put(r.re, out); sc_put(r.im, out); return 2; }
        case SC_C_ADD: { auto r = ca + cb; sc_put(r.re, out); sc_put(r.im, out); return 2; }
        case SC_C_SUB: { auto r = ca - cb; sc_put(r.re, out); sc_put(r.im, out); return 2; }
        case SC_C_SQR: { auto r = sqr(ca); sc_put(r.re, out); sc_put(r.im, out); return 2; }
        case SC_C_CONJ: { auto r = conj(ca); sc_put(r.re, out); sc_put(r.im, out); return 2; }
        case SC_C_ABS: sc_put(abs(ca), out); return 1;
        case SC_C_ARG: sc_put(arg(ca), out); return 1;
        default: break;
        }
    }
    if (id == SC_SYMM_DIRICH6)
    {
        using A6 = Scalar<6, true>;
        Vec<double, 2> ar(p[6], p[7]), br(p[8], p[9]), cr(p[10], p[11]);
        Mat<double, 2, 2> Mr = col_mat(br - ar, cr - ar);
        Vec<A6, 2> va(A6(p[0], 0), A6(p[1], 1)), vb(A6(p[2], 2), A6(p[3], 3)), vc(A6(p[4], 4), A6(p[5], 5));
        Mat<A6, 2, 2> M = col_mat(vb - va, vc - va);
        Mat<A6, 2, 2> J = M * Mr.inverse();
        A6 E = J.squaredNorm() + J.inverse().squaredNorm();
        sc_put(E, out);
        return 1;
    }
    if (id == SC_HESS_BLOCK_ISSUE13)  // tests/ScalarTestHessianBlock.cc:12-45: f = y3 ((x1 - y1)^2 + (x2 - y2)^2), block d2f / dx dy = (0, 2, 2, 3)
    {
        auto fn = [&](auto tag) {
            using T = typename decltype(tag)::type;
            const T x1(p[0], 0), x2(p[1], 1), y1(p[2], 2), y2(p[3], 3), y3(p[4], 4);
            const T r1 = x1 - y1, r2 = x2 - y2;
            return y3 * (r1 * r1 + r2 * r2);
        };
        struct FullTag { using type = Scalar<5, true>; };
        struct BlockTag { using type = ScalarHessianBlock<5, 0, 2, 2, 3>; };
        sc_put(fn(FullTag{}), out);
        sc_put(fn(BlockTag{}), out);
        return 2;
    }
    if (id == SC_HESS_BLOCK_SYMDIR)   // tests/ScalarTestHessianBlock.cc:50-100: symmetric Dirichlet, k = 6; p[12] selects the block
    {
        Vec<double, 2> ar(p[6], p[7]), br(p[8], p[9]), cr(p[10], p[11]);
        Mat<double, 2, 2> Mr = col_mat(br - ar, cr - ar);
        auto fn = [&](auto tag) {
            using T = typename decltype(tag)::type;
            Vec<T, 2> va(T(p[0], 0), T(p[1], 1)), vb(T(p[2], 2), T(p[3], 3)), vc(T(p[4], 4), T(p[5], 5));
            Mat<T, 2, 2> M = col_mat(vb - va, vc - va);
            Mat<T, 2, 2> J = M * Mr.inverse();
            T E = J.squaredNorm() + J.inverse().squaredNorm();
            return E;
        };
        struct FullTag { using type = Scalar<6, true>; };
        struct B0 { using type = ScalarHessianBlock<6, 0, 0, 0, 0>; };
        struct B1 { using type = ScalarHessianBlock<6, 0, 0, 6, 6>; };
        struct B2 { using type = ScalarHessianBlock<6, 0, 0, 3, 1>; };
        struct B3 { using type = ScalarHessianBlock<6, 0, 0, 1, 3>; };
        struct B4 { using type = ScalarHessianBlock<6, 2, 2, 2, 2>; };
        struct B5 { using type = ScalarHessianBlock<6, 1, 4, 5, 2>; };
        sc_put(fn(FullTag{}), out);
        switch ((int)p[12])
        {
        case 0: sc_put(fn(B0{}), out); break;
        case 1: sc_put(fn(B1{}), out); break;
        case 2: sc_put(fn(B2{}), out); break;
        case 3: sc_put(fn(B3{}), out); break;
        case 4: sc_put(fn(B4{}), out); break;
        default: sc_put(fn(B5{}), out); break;
        }
        return 2;
    }
    if (id == SC_SVD2 || id == SC_CLOSEST_ORTHOGONAL2)  // tests/SVDTest.cc:9-97, Scalar<4>: A = [[p0, p1], [p2, p3]]
    {
        using A4 = Scalar<4, true>;
        Mat<A4, 2, 2> A;
        A(0, 0) = A4(p[0], 0); A(0, 1) = A4(p[1], 1); A(1, 0) = A4(p[2], 2); A(1, 1) = A4(p[3], 3);
        Mat<A4, 2, 2> R;
        if (id == SC_SVD2)
        {
            Mat<A4, 2, 2> U, V;
            Vec<A4, 2> S;
            svd(A, U, S, V);
            Mat<A4, 2, 2> US;
            US(0, 0) = U(0, 0) * S[0]; US(0, 1) = U(0, 1) * S[1]; US(1, 0) = U(1, 0) * S[0]; US(1, 1) = U(1, 1) * S[1];
            R = US * V.transpose();  // U * S.asDiagonal() * V^T
            sc_put(R(0, 0), out); sc_put(R(0, 1), out); sc_put(R(1, 0), out); sc_put(R(1, 1), out);
            sc_put(S[0], out); sc_put(S[1], out);
            return 6;
        }
        R = closest_orthogonal(A);
        sc_put(R(0, 0), out); sc_put(R(0, 1), out); sc_put(R(1, 0), out); sc_put(R(1, 1), out);
        return 4;
    }
    return -1;
}

}  // namespace tadx
