// TEST INFRASTRUCTURE ONLY -- CPU oracle for tinyad_b200.
//
// A plain C++17 (no Eigen, no CUDA) restatement of the reference's per-element
// sparse-derivative path.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, link or call this; the
// product (tinyad_b200/) never includes it.
//
// What is restated (reference file:line, relative to /root/reference):
//   Scalar<k>            include/TinyAD/Scalar.hh:24-1347   (full k x k Hessian, same formulas
//                                                            and per-entry evaluation order)
//   Element              include/TinyAD/Detail/Element.hh:159-170,198-295
//   objective terms      include/TinyAD/Detail/ScalarObjectiveTerm.hh:162-278,
//                        include/TinyAD/Detail/VectorObjectiveTerm.hh:158-243,326-351
//   facades              include/TinyAD/Detail/ScalarFunctionImpl.hh:256-416,
//                        include/TinyAD/Detail/VectorFunctionImpl.hh:143-188,254-283
//   projection           include/TinyAD/Utils/HessianProjection.hh:16-101
//   parallel_for         include/TinyAD/Detail/Parallel.hh:20-69
//   helpers              include/TinyAD/Utils/Helpers.hh:17-76, Utils/ToPassive.hh:15-30
//
// Third-party dependency of the reference that is NOT in /root/reference: Eigen3
// (unpinned, CMakeLists.txt:14-17; 3.4.0 on the reference's CI).  The pieces of
// Eigen the path relies on are restated from their published behaviour:
//   * fixed-size determinant()/inverse() (2x2, 3x3 cofactor form), squaredNorm(),
//     matrix products  -> struct Mat below;
//   * SelfAdjointEigenSolver (reads the lower triangle, eigenvalues ascending,
//     orthonormal eigenvectors) -> sym_eig_jacobi(), a cyclic Jacobi solver: any
//     backward-stable solver gives the same projected matrix to O(eps*|H|);
//   * SparseMatrix::setFromTriplets (compressed column storage, ascending inner
//     indices, duplicates summed in input order, zeros kept) -> set_from_triplets().
//
// Pinning: tests/test_oracle_golden.py checks this file against the reference's
// own known-answer tests (tests/ScalarTest*.cc, ComplexTest.cc,
// ScalarFunctionTest.cc, VectorFunctionTest.cc, NewtonTest.cc, GaussNewtonTest.cc,
// DynamicElementsTest.cc) -- see oracle/README.md for the list -- and
// tests/test_oracle_vs_reference.py checks it against the reference itself: oracle/_ref
// (ref_driver.cc = the unmodified TinyAD headers compiled in place over oracle/eigen_shim)
// on the same inputs: patterns bit-exact, f / g / H 1e-13, projected H 1e-10.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle
{

using Index = std::int64_t;

[[noreturn]] inline void error_throw(const std::string& msg)
{
    // Utils/Out.hh:73-79: print + throw std::runtime_error (printing omitted here)
    throw std::runtime_error(msg);
}

// ---------------------------------------------------------------------------
// Scalar<k, with_hessian>  (Scalar.hh:24-1347)
// ---------------------------------------------------------------------------
template <int k, bool with_hessian = true>
struct Scalar
{
    static constexpr int k_ = k;
    static constexpr bool with_hessian_ = with_hessian;
    static constexpr int hs = with_hessian ? k * k : 0;

    double val = 0.0;                 // Scalar.hh:1341
    std::array<double, k> grad{};     // Scalar.hh:1342 (zero-initialised)
    std::array<double, hs> Hess{};    // Scalar.hh:1344 full k x k, (i,j) at i*k+j

    Scalar() = default;
    Scalar(double _val) : val(_val) {}                          // passive, Scalar.hh:62-66
    Scalar(double _val, Index _idx) : val(_val)                 // active,  Scalar.hh:70-78
    {
        if (_idx < 0 || _idx >= k) error_throw("Scalar: index out of range");
        grad[_idx] = 1.0;
    }

    double& H(int i, int j) { return Hess[i * k + j]; }
    const double& H(int i, int j) const { return Hess[i * k + j]; }

    // Scalar.hh:81-106
    static Scalar known_derivatives(double _val, const std::array<double, k>& _grad, const std::array<double, hs>& _Hess)
    {
        Scalar res;
        res.val = _val;
        res.grad = _grad;
        res.Hess = _Hess;
        return res;
    }
    static Scalar known_derivatives(double _val, double _grad, double _Hess)
    {
        static_assert(k == 1, "univariate only");
        Scalar res;
        res.val = _val;
        res.grad[0] = _grad;
        if constexpr (with_hessian) res.Hess[0] = _Hess;
        return res;
    }

    // Scalar.hh:199-214
    static Scalar chain(double val, double grad, double Hess, const Scalar& a)
    {
        Scalar res;
        res.val = val;
        for (int i = 0; i < k; ++i) res.grad[i] = grad * a.grad[i];
        if constexpr (with_hessian)
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                    res.H(i, j) = Hess * (a.grad[i] * a.grad[j]) + grad * a.H(i, j);
        return res;
    }

    // ---- unary (Scalar.hh:220-552) ----
    friend Scalar operator-(const Scalar& a)
    {
        Scalar res;
        res.val = -a.val;
        for (int i = 0; i < k; ++i) res.grad[i] = -a.grad[i];
        for (int i = 0; i < hs; ++i) res.Hess[i] = -a.Hess[i];
        return res;
    }
    friend Scalar sqrt(const Scalar& a)
    {
        const double f = std::sqrt(a.val);
        return chain(f, 0.5 / f, -0.25 / (f * a.val), a);
    }
    friend Scalar sqr(const Scalar& a)
    {
        Scalar res;
        res.val = a.val * a.val;
        for (int i = 0; i < k; ++i) res.grad[i] = 2.0 * a.val * a.grad[i];
        if constexpr (with_hessian)
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                    res.H(i, j) = 2.0 * (a.val * a.H(i, j) + a.grad[i] * a.grad[j]);
        return res;
    }
    friend Scalar pow(const Scalar& a, const int& e)
    {
        if (e == 0) return chain(1.0, 0.0, 0.0, a);
        else if (e == 1) return chain(a.val, 1.0, 0.0, a);
        else
        {
            const double f2 = std::pow(a.val, e - 2);
            const double f1 = f2 * a.val;
            const double f = f1 * a.val;
            return chain(f, e * f1, e * (e - 1) * f2, a);
        }
    }
    friend Scalar pow(const Scalar& a, const double& e)
    {
        const double f2 = std::pow(a.val, e - 2.0);
        const double f1 = f2 * a.val;
        const double f = f1 * a.val;
        return chain(f, e * f1, e * (e - 1.0) * f2, a);
    }
    friend Scalar fabs(const Scalar& a)
    {
        if (a.val >= 0.0) return chain(a.val, 1.0, 0.0, a);
        else return chain(-a.val, -1.0, 0.0, a);
    }
    friend Scalar abs(const Scalar& a) { return fabs(a); }
    friend Scalar exp(const Scalar& a)
    {
        const double e = std::exp(a.val);
        return chain(e, e, e, a);
    }
    friend Scalar log(const Scalar& a)
    {
        const double a_inv = 1.0 / a.val;
        return chain(std::log(a.val), a_inv, -a_inv / a.val, a);
    }
    friend Scalar log2(const Scalar& a)
    {
        const double a_inv = 1.0 / a.val / std::log(2.0);
        return chain(std::log2(a.val), a_inv, -a_inv / a.val, a);
    }
    friend Scalar log10(const Scalar& a)
    {
        const double a_inv = 1.0 / a.val / std::log(10.0);
        return chain(std::log10(a.val), a_inv, -a_inv / a.val, a);
    }
    friend Scalar sin(const Scalar& a)
    {
        const double s = std::sin(a.val);
        return chain(s, std::cos(a.val), -s, a);
    }
    friend Scalar cos(const Scalar& a)
    {
        const double c = std::cos(a.val);
        return chain(c, -std::sin(a.val), -c, a);
    }
    friend Scalar tan(const Scalar& a)
    {
        const double c = std::cos(a.val);
        const double c2 = c * c;
        const double c3 = c2 * c;
        return chain(std::tan(a.val), 1.0 / c2, 2.0 * std::sin(a.val) / c3, a);
    }
    friend Scalar asin(const Scalar& a)
    {
        const double s = 1.0 - a.val * a.val;
        const double s_sqrt = std::sqrt(s);
        return chain(std::asin(a.val), 1.0 / s_sqrt, a.val / s_sqrt / s, a);
    }
    friend Scalar acos(const Scalar& a)
    {
        if (!(a.val > -1.0) || !(a.val < 1.0)) error_throw("acos: argument out of (-1,1)");  // Scalar.hh:433-434
        const double s = 1.0 - a.val * a.val;
        const double s_sqrt = std::sqrt(s);
        return chain(std::acos(a.val), -1.0 / s_sqrt, -a.val / s_sqrt / s, a);
    }
    friend Scalar atan(const Scalar& a)
    {
        const double s = a.val * a.val + 1.0;
        return chain(std::atan(a.val), 1.0 / s, -2.0 * a.val / s / s, a);
    }
    friend Scalar sinh(const Scalar& a)
    {
        const double s = std::sinh(a.val);
        return chain(s, std::cosh(a.val), s, a);
    }
    friend Scalar cosh(const Scalar& a)
    {
        const double c = std::cosh(a.val);
        return chain(c, std::sinh(a.val), c, a);
    }
    friend Scalar tanh(const Scalar& a)
    {
        const double c = std::cosh(a.val);
        const double c2 = c * c;
        const double c3 = c2 * c;
        return chain(std::tanh(a.val), 1.0 / c2, -2.0 * std::sinh(a.val) / c3, a);
    }
    friend Scalar asinh(const Scalar& a)
    {
        const double s = a.val * a.val + 1.0;
        const double s_sqrt = std::sqrt(s);
        return chain(std::asinh(a.val), 1.0 / s_sqrt, -a.val / s_sqrt / s, a);
    }
    friend Scalar acosh(const Scalar& a)
    {
        const double sm = a.val - 1.0;
        const double sp = a.val + 1.0;
        const double prod = std::sqrt(sm) * std::sqrt(sp);
        return chain(std::acosh(a.val), 1.0 / prod, -a.val / prod / sm / sp, a);
    }
    friend Scalar atanh(const Scalar& a)
    {
        const double s = 1.0 - a.val * a.val;
        return chain(std::atanh(a.val), 1.0 / s, 2.0 * a.val / s / s, a);
    }
    friend bool isnan(const Scalar& a) { return std::isnan(a.val); }
    friend bool isinf(const Scalar& a) { return std::isinf(a.val); }
    friend bool isfinite(const Scalar& a) { return std::isfinite(a.val); }

    // ---- binary (Scalar.hh:579-931) ----
    friend Scalar operator+(const Scalar& a, const Scalar& b)
    {
        Scalar res;
        res.val = a.val + b.val;
        for (int i = 0; i < k; ++i) res.grad[i] = a.grad[i] + b.grad[i];
        for (int i = 0; i < hs; ++i) res.Hess[i] = a.Hess[i] + b.Hess[i];
        return res;
    }
    friend Scalar operator+(const Scalar& a, const double& b) { Scalar res = a; res.val += b; return res; }
    friend Scalar operator+(const double& a, const Scalar& b) { Scalar res = b; res.val += a; return res; }
    Scalar& operator+=(const Scalar& b)
    {
        val += b.val;
        for (int i = 0; i < k; ++i) grad[i] += b.grad[i];
        for (int i = 0; i < hs; ++i) Hess[i] += b.Hess[i];
        return *this;
    }
    Scalar& operator+=(const double& b) { val += b; return *this; }
    friend Scalar operator-(const Scalar& a, const Scalar& b)
    {
        Scalar res;
        res.val = a.val - b.val;
        for (int i = 0; i < k; ++i) res.grad[i] = a.grad[i] - b.grad[i];
        for (int i = 0; i < hs; ++i) res.Hess[i] = a.Hess[i] - b.Hess[i];
        return res;
    }
    friend Scalar operator-(const Scalar& a, const double& b) { Scalar res = a; res.val -= b; return res; }
    friend Scalar operator-(const double& a, const Scalar& b)
    {
        Scalar res;
        res.val = a - b.val;
        for (int i = 0; i < k; ++i) res.grad[i] = -b.grad[i];
        for (int i = 0; i < hs; ++i) res.Hess[i] = -b.Hess[i];
        return res;
    }
    Scalar& operator-=(const Scalar& b)
    {
        val -= b.val;
        for (int i = 0; i < k; ++i) grad[i] -= b.grad[i];
        for (int i = 0; i < hs; ++i) Hess[i] -= b.Hess[i];
        return *this;
    }
    Scalar& operator-=(const double& b) { val -= b; return *this; }
    friend Scalar operator*(const Scalar& a, const Scalar& b)
    {
        Scalar res;
        res.val = a.val * b.val;
        for (int i = 0; i < k; ++i) res.grad[i] = b.val * a.grad[i] + a.val * b.grad[i];
        if constexpr (with_hessian)  // Scalar.hh:765, left-to-right per entry
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                    res.H(i, j) = b.val * a.H(i, j) + a.grad[i] * b.grad[j] + b.grad[i] * a.grad[j] + a.val * b.H(i, j);
        return res;
    }
    friend Scalar operator*(const Scalar& a, const double& b)
    {
        Scalar res = a;
        res.val *= b;
        for (int i = 0; i < k; ++i) res.grad[i] *= b;
        for (int i = 0; i < hs; ++i) res.Hess[i] *= b;
        return res;
    }
    friend Scalar operator*(const double& a, const Scalar& b)
    {
        Scalar res = b;
        res.val *= a;
        for (int i = 0; i < k; ++i) res.grad[i] *= a;
        for (int i = 0; i < hs; ++i) res.Hess[i] *= a;
        return res;
    }
    Scalar& operator*=(const Scalar& b) { *this = *this * b; return *this; }
    Scalar& operator*=(const double& b) { *this = *this * b; return *this; }
    friend Scalar operator/(const Scalar& a, const Scalar& b)
    {
        Scalar res;
        res.val = a.val / b.val;
        for (int i = 0; i < k; ++i) res.grad[i] = (b.val * a.grad[i] - a.val * b.grad[i]) / (b.val * b.val);
        if constexpr (with_hessian)  // Scalar.hh:837
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                    res.H(i, j) = (a.H(i, j) - res.grad[i] * b.grad[j] - b.grad[i] * res.grad[j] - res.val * b.H(i, j)) / b.val;
        return res;
    }
    friend Scalar operator/(const Scalar& a, const double& b)
    {
        Scalar res = a;
        res.val /= b;
        for (int i = 0; i < k; ++i) res.grad[i] /= b;
        for (int i = 0; i < hs; ++i) res.Hess[i] /= b;
        return res;
    }
    friend Scalar operator/(const double& a, const Scalar& b)
    {
        Scalar res;
        res.val = a / b.val;
        const double c = -a / (b.val * b.val);
        for (int i = 0; i < k; ++i) res.grad[i] = c * b.grad[i];
        if constexpr (with_hessian)  // Scalar.hh:873
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                    res.H(i, j) = ((-res.grad[i]) * b.grad[j] - b.grad[i] * res.grad[j] - res.val * b.H(i, j)) / b.val;
        return res;
    }
    Scalar& operator/=(const Scalar& b) { *this = *this / b; return *this; }
    Scalar& operator/=(const double& b) { *this = *this / b; return *this; }
    friend Scalar atan2(const Scalar& y, const Scalar& x)
    {
        Scalar res;
        res.val = std::atan2(y.val, x.val);
        std::array<double, k> u;
        for (int i = 0; i < k; ++i) u[i] = x.val * y.grad[i] - y.val * x.grad[i];
        const double v = x.val * x.val + y.val * y.val;
        for (int i = 0; i < k; ++i) res.grad[i] = u[i] / v;
        if constexpr (with_hessian)  // Scalar.hh:909-915
        {
            std::array<double, k> dv;
            for (int i = 0; i < k; ++i) dv[i] = 2.0 * (x.val * x.grad[i] + y.val * y.grad[i]);
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                {
                    const double du = x.val * y.H(i, j) - y.val * x.H(i, j) + y.grad[i] * x.grad[j] - x.grad[i] * y.grad[j];
                    res.H(i, j) = (du - res.grad[i] * dv[j]) / v;
                }
        }
        return res;
    }
    friend Scalar hypot(const Scalar& a, const Scalar& b) { return sqrt(a * a + b * b); }

    // ---- comparisons on val only (Scalar.hh:933-1095) ----
    friend bool operator==(const Scalar& a, const Scalar& b) { return a.val == b.val; }
    friend bool operator==(const Scalar& a, const double& b) { return a.val == b; }
    friend bool operator==(const double& a, const Scalar& b) { return a == b.val; }
    friend bool operator!=(const Scalar& a, const Scalar& b) { return a.val != b.val; }
    friend bool operator!=(const Scalar& a, const double& b) { return a.val != b; }
    friend bool operator!=(const double& a, const Scalar& b) { return a != b.val; }
    friend bool operator<(const Scalar& a, const Scalar& b) { return a.val < b.val; }
    friend bool operator<(const Scalar& a, const double& b) { return a.val < b; }
    friend bool operator<(const double& a, const Scalar& b) { return a < b.val; }
    friend bool operator<=(const Scalar& a, const Scalar& b) { return a.val <= b.val; }
    friend bool operator<=(const Scalar& a, const double& b) { return a.val <= b; }
    friend bool operator<=(const double& a, const Scalar& b) { return a <= b.val; }
    friend bool operator>(const Scalar& a, const Scalar& b) { return a.val > b.val; }
    friend bool operator>(const Scalar& a, const double& b) { return a.val > b; }
    friend bool operator>(const double& a, const Scalar& b) { return a > b.val; }
    friend bool operator>=(const Scalar& a, const Scalar& b) { return a.val >= b.val; }
    friend bool operator>=(const Scalar& a, const double& b) { return a.val >= b; }
    friend bool operator>=(const double& a, const Scalar& b) { return a >= b.val; }
    // Scalar.hh:1097-1145
    friend Scalar min(const Scalar& a, const Scalar& b) { return (b < a) ? b : a; }
    friend Scalar fmin(const Scalar& a, const Scalar& b) { return min(a, b); }
    friend Scalar max(const Scalar& a, const Scalar& b) { return (a < b) ? b : a; }
    friend Scalar fmax(const Scalar& a, const Scalar& b) { return max(a, b); }
    friend Scalar clamp(const Scalar& x, const Scalar& a, const Scalar& b)
    {
        if (x < a) return a;
        else if (x > b) return b;
        else return x;
    }

    // ---- std::complex overloads (Scalar.hh:1151-1320) ----
    using C = std::complex<Scalar>;
    using Cd = std::complex<double>;
    friend C operator+(const C& a, const C& b) { return C(a.real() + b.real(), a.imag() + b.imag()); }
    friend C operator+(const Cd& a, const C& b) { return C(a.real() + b.real(), a.imag() + b.imag()); }
    friend C operator+(const C& a, const Cd& b) { return C(a.real() + b.real(), a.imag() + b.imag()); }
    friend C operator-(const C& a, const C& b) { return C(a.real() - b.real(), a.imag() - b.imag()); }
    friend C operator-(const Cd& a, const C& b) { return C(a.real() - b.real(), a.imag() - b.imag()); }
    friend C operator-(const C& a, const Cd& b) { return C(a.real() - b.real(), a.imag() - b.imag()); }
    friend C operator*(const C& a, const C& b)
    {
        return C(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
    }
    friend C operator*(const Cd& a, const C& b)
    {
        return C(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
    }
    friend C operator*(const C& a, const Cd& b)
    {
        return C(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
    }
    friend C sqr(const C& a) { return C(sqr(a.real()) - sqr(a.imag()), 2.0 * a.real() * a.imag()); }
    friend C operator/(const C& a, const C& b)
    {
        const Scalar denom = b.real() * b.real() + b.imag() * b.imag();
        return C((a.real() * b.real() + a.imag() * b.imag()) / denom, (a.imag() * b.real() - a.real() * b.imag()) / denom);
    }
    friend C operator/(const C& a, const Cd& b)
    {
        const double denom = b.real() * b.real() + b.imag() * b.imag();
        return C((a.real() * b.real() + a.imag() * b.imag()) / denom, (a.imag() * b.real() - a.real() * b.imag()) / denom);
    }
    friend C conj(const C& a) { return C(a.real(), -a.imag()); }
    friend Scalar abs(const C& a) { return hypot(a.real(), a.imag()); }
    friend Scalar arg(const C& a) { return atan2(a.imag(), a.real()); }
};

template <int k>
using Double = Scalar<k, true>;

// Utils/ToPassive.hh:15-30, Scalar.hh:1368-1369
inline double to_passive(const double& a) { return a; }
template <int k, bool wh>
double to_passive(const Scalar<k, wh>& a) { return a.val; }
inline double sqr(const double& x) { return x * x; }

// ---------------------------------------------------------------------------
// Tiny fixed-size matrix with the Eigen semantics the path uses
// ---------------------------------------------------------------------------
template <typename T, int R, int C>
struct Mat
{
    T a[R * C] = {};  // column-major like Eigen
    T& operator()(int i, int j) { return a[j * R + i]; }
    const T& operator()(int i, int j) const { return a[j * R + i]; }
    T& operator[](int i) { static_assert(C == 1 || R == 1); return a[i]; }
    const T& operator[](int i) const { static_assert(C == 1 || R == 1); return a[i]; }
    T& operator()(int i) { return (*this)[i]; }
    const T& operator()(int i) const { return (*this)[i]; }

    static Mat Constant(const T& v) { Mat m; for (auto& e : m.a) e = v; return m; }
    static Mat Zero() { return Constant(T(0.0)); }

    T squaredNorm() const { T s = a[0] * a[0]; for (int i = 1; i < R * C; ++i) s = s + a[i] * a[i]; return s; }
    T sum() const { T s = a[0]; for (int i = 1; i < R * C; ++i) s = s + a[i]; return s; }
    template <typename U>
    auto dot(const Mat<U, R, C>& o) const { auto s = a[0] * o.a[0]; for (int i = 1; i < R * C; ++i) s = s + a[i] * o.a[i]; return s; }
    Mat<T, C, R> transpose() const { Mat<T, C, R> t; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) t(j, i) = (*this)(i, j); return t; }

    T determinant() const
    {
        static_assert(R == C && (R == 2 || R == 3));
        const Mat& m = *this;
        if constexpr (R == 2)
            return m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1);
        else
        {
            auto h = [&](int x, int y, int z) { return m(0, x) * (m(1, y) * m(2, z) - m(1, z) * m(2, y)); };
            return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
        }
    }
    Mat inverse() const
    {
        static_assert(R == C && (R == 2 || R == 3));
        const Mat& m = *this;
        Mat r;
        if constexpr (R == 2)
        {
            const T invdet = T(1.0) / m.determinant();
            r(0, 0) = m(1, 1) * invdet;
            r(1, 0) = -m(1, 0) * invdet;
            r(0, 1) = -m(0, 1) * invdet;
            r(1, 1) = m(0, 0) * invdet;
        }
        else
        {
            auto cof = [&](int i, int j) {
                const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
            };
            const T c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
            const T det = c00 * m(0, 0) + c10 * m(1, 0) + c20 * m(2, 0);
            const T invdet = T(1.0) / det;
            r(0, 0) = c00 * invdet; r(0, 1) = c10 * invdet; r(0, 2) = c20 * invdet;
            r(1, 0) = cof(0, 1) * invdet; r(1, 1) = cof(1, 1) * invdet; r(1, 2) = cof(2, 1) * invdet;
            r(2, 0) = cof(0, 2) * invdet; r(2, 1) = cof(1, 2) * invdet; r(2, 2) = cof(2, 2) * invdet;
        }
        return r;
    }
};
template <typename T, int N>
using Vec = Mat<T, N, 1>;

template <typename T, typename U, int R, int C>
auto operator+(const Mat<T, R, C>& x, const Mat<U, R, C>& y)
{
    Mat<decltype(x.a[0] + y.a[0]), R, C> r;
    for (int i = 0; i < R * C; ++i) r.a[i] = x.a[i] + y.a[i];
    return r;
}
template <typename T, typename U, int R, int C>
auto operator-(const Mat<T, R, C>& x, const Mat<U, R, C>& y)
{
    Mat<decltype(x.a[0] - y.a[0]), R, C> r;
    for (int i = 0; i < R * C; ++i) r.a[i] = x.a[i] - y.a[i];
    return r;
}
template <typename T, typename U, int R, int K, int C>
auto operator*(const Mat<T, R, K>& x, const Mat<U, K, C>& y)
{
    Mat<decltype(x.a[0] * y.a[0]), R, C> r;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j)
        {
            auto s = x(i, 0) * y(0, j);
            for (int l = 1; l < K; ++l) s = s + x(i, l) * y(l, j);
            r(i, j) = s;
        }
    return r;
}
template <typename T, int R, int C>
Mat<T, R, C> operator*(const double& s, const Mat<T, R, C>& x)
{
    Mat<T, R, C> r;
    for (int i = 0; i < R * C; ++i) r.a[i] = s * x.a[i];
    return r;
}
template <typename T, int R, int C>
Mat<T, R, C> operator*(const Mat<T, R, C>& x, const double& s)
{
    Mat<T, R, C> r;
    for (int i = 0; i < R * C; ++i) r.a[i] = x.a[i] * s;
    return r;
}
// Utils/Helpers.hh:48-76
template <typename T, int R>
Mat<T, R, 2> col_mat(const Vec<T, R>& v0, const Vec<T, R>& v1)
{
    Mat<T, R, 2> M;
    for (int i = 0; i < R; ++i) { M(i, 0) = v0[i]; M(i, 1) = v1[i]; }
    return M;
}
template <typename T, int R>
Mat<T, R, 3> col_mat(const Vec<T, R>& v0, const Vec<T, R>& v1, const Vec<T, R>& v2)
{
    Mat<T, R, 3> M;
    for (int i = 0; i < R; ++i) { M(i, 0) = v0[i]; M(i, 1) = v1[i]; M(i, 2) = v2[i]; }
    return M;
}

// ---------------------------------------------------------------------------
// Hessian projection (Utils/HessianProjection.hh:16-101)
// ---------------------------------------------------------------------------
constexpr double default_hessian_projection_eps = 1e-9;  // :16

// Cyclic Jacobi on the LOWER triangle of a (k x k, (i,j) at i*k+j).
// Stands in for Eigen::SelfAdjointEigenSolver (:66): eigenvalues ascending in w,
// eigenvectors as columns of V ((i,j) at i*k+j).
inline void sym_eig_jacobi(int k, const double* a_in, double* w, double* V)
{
    std::vector<double> A(k * k);
    for (int i = 0; i < k; ++i)
        for (int j = 0; j <= i; ++j)
            A[i * k + j] = A[j * k + i] = a_in[i * k + j];
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) V[i * k + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; ++sweep)
    {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j)
                (i == j ? diag : off) += A[i * k + j] * A[i * k + j];
        // converged when the off-diagonal Frobenius norm is below machine epsilon times the norm of the matrix (eigenvalue error
        // <= off^2 / gap).  A stricter test never fires for the element Hessians of the benchmark: their three-dimensional null
        // space leaves rounding noise of ~1e-17 |A| off the diagonal, and the solver would burn all 100 sweeps on every element
        // (that made the CPU baseline ~6x slower than the algorithm it stands for).
        if (off == 0.0 || off <= 4.9e-32 * (diag + off)) break;
        for (int p = 0; p < k - 1; ++p)
            for (int q = p + 1; q < k; ++q)
            {
                const double apq = A[p * k + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * k + q] - A[p * k + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0);
                const double s = t * c;
                for (int r = 0; r < k; ++r)
                {
                    const double arp = A[r * k + p], arq = A[r * k + q];
                    A[r * k + p] = c * arp - s * arq;
                    A[r * k + q] = s * arp + c * arq;
                }
                for (int r = 0; r < k; ++r)
                {
                    const double apr = A[p * k + r], aqr = A[q * k + r];
                    A[p * k + r] = c * apr - s * aqr;
                    A[q * k + r] = s * apr + c * aqr;
                }
                for (int r = 0; r < k; ++r)
                {
                    const double vrp = V[r * k + p], vrq = V[r * k + q];
                    V[r * k + p] = c * vrp - s * vrq;
                    V[r * k + q] = s * vrp + c * vrq;
                }
            }
    }
    std::vector<int> order(k);
    for (int i = 0; i < k; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return A[x * k + x] < A[y * k + y]; });
    std::vector<double> Vs(k * k);
    for (int c = 0; c < k; ++c)
    {
        w[c] = A[order[c] * k + order[c]];
        for (int r = 0; r < k; ++r) Vs[r * k + c] = V[r * k + order[c]];
    }
    std::copy(Vs.begin(), Vs.end(), V);
}

// :23-42
inline bool positive_diagonally_dominant(int k, const double* H, double eps)
{
    for (int i = 0; i < k; ++i)
    {
        double off_diag_abs_sum = 0.0;
        for (int j = 0; j < k; ++j)
            if (i != j) off_diag_abs_sum += std::abs(H[i * k + j]);
        if (H[i * k + i] < off_diag_abs_sum + eps) return false;
    }
    return true;
}

// :48-101.  Returns 0 = early-out (diag. dominant), 1 = decomposed but untouched, 2 = rebuilt.
inline int project_positive_definite(int k, double* H, double eps)
{
    if (k == 0) return 0;
    if (positive_diagonally_dominant(k, H, eps)) return 0;
    std::vector<double> w(k), V(k * k);
    sym_eig_jacobi(k, H, w.data(), V.data());
    bool all_positive = true;
    for (int i = 0; i < k; ++i)
    {
        if (eps < 0)
        {
            if (w[i] < 0) { w[i] = -w[i]; all_positive = false; }
        }
        else
        {
            if (w[i] < eps) { w[i] = eps; all_positive = false; }
        }
    }
    if (all_positive) return 1;
    // H = (V * D) * V^T
    std::vector<double> VD(k * k);
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) VD[i * k + j] = V[i * k + j] * w[j];
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j)
        {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += VD[i * k + l] * V[j * k + l];
            H[i * k + j] = s;
            if (!std::isfinite(s)) error_throw("project_positive_definite: non-finite result");
        }
    return 2;
}

// ---------------------------------------------------------------------------
// parallel_for (Detail/Parallel.hh:20-69), EvalSettings (Detail/EvalSettings.hh:14-21)
// ---------------------------------------------------------------------------
struct EvalSettings { int n_threads = -1; };

inline int get_n_threads(const EvalSettings& s)
{
#ifdef _OPENMP
    if (s.n_threads > 0) return s.n_threads;
    return std::max(1, omp_get_max_threads() - 1);  // Parallel.hh:27
#else
    (void)s;
    return 1;
#endif
}

template <typename F>
void parallel_for(Index n, const EvalSettings& settings, F&& body)
{
    std::exception_ptr first;
    bool cancel = false;
    const int nt = get_n_threads(settings);
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (Index i = 0; i < n; ++i)
    {
        if (cancel) continue;
        try { body(i); }
        catch (...)
        {
#pragma omp critical(oracle_exc)
            { if (!first) first = std::current_exception(); cancel = true; }
        }
    }
    if (first) std::rethrow_exception(first);
}

// ---------------------------------------------------------------------------
// Sparse matrix, setFromTriplets semantics
// ---------------------------------------------------------------------------
struct Triplet { std::int32_t row, col; double value; };

struct SparseMatrix  // compressed column storage, StorageIndex = int32
{
    Index rows = 0, cols = 0;
    std::vector<std::int32_t> outer{0};  // cols+1
    std::vector<std::int32_t> inner;     // row index per nz, ascending per column
    std::vector<double> values;
    Index nonZeros() const { return (Index)values.size(); }
    double coeff(Index r, Index c) const
    {
        for (Index p = outer[c]; p < outer[c + 1]; ++p)
            if (inner[p] == r) return values[p];
        return 0.0;
    }
};

inline SparseMatrix set_from_triplets(Index rows, Index cols, const std::vector<Triplet>& T)
{
    // pass 1: bucket by row in input order (row-major temporary)
    std::vector<Index> rptr(rows + 1, 0);
    for (const auto& t : T) ++rptr[t.row + 1];
    for (Index r = 0; r < rows; ++r) rptr[r + 1] += rptr[r];
    std::vector<std::int32_t> rcol(T.size());
    std::vector<double> rval(T.size());
    {
        std::vector<Index> pos(rptr.begin(), rptr.end() - 1);
        for (const auto& t : T) { const Index p = pos[t.row]++; rcol[p] = t.col; rval[p] = t.value; }
    }
    // collapse duplicates per row: first occurrence keeps the slot, later ones add in input order
    std::vector<Index> wi(cols, -1);
    std::vector<Index> rend(rows);
    Index count = 0;
    std::vector<Index> rstart(rows + 1, 0);
    for (Index r = 0; r < rows; ++r)
    {
        const Index start = count;
        rstart[r] = start;
        for (Index p = rptr[r]; p < rptr[r + 1]; ++p)
        {
            const std::int32_t c = rcol[p];
            if (wi[c] >= start) rval[wi[c]] += rval[p];
            else { rval[count] = rval[p]; rcol[count] = c; wi[c] = count; ++count; }
        }
        rend[r] = count;
    }
    rstart[rows] = count;
    // pass 2: transpose to column-major; rows visited ascending => inner indices ascending
    SparseMatrix M;
    M.rows = rows; M.cols = cols;
    M.outer.assign(cols + 1, 0);
    for (Index p = 0; p < count; ++p) ++M.outer[rcol[p] + 1];
    for (Index c = 0; c < cols; ++c) M.outer[c + 1] += M.outer[c];
    M.inner.resize(count);
    M.values.resize(count);
    std::vector<Index> pos(M.outer.begin(), M.outer.end() - 1);
    for (Index r = 0; r < rows; ++r)
        for (Index p = rstart[r]; p < rend[r]; ++p)
        {
            const Index q = pos[rcol[p]]++;
            M.inner[q] = (std::int32_t)r;
            M.values[q] = rval[p];
        }
    return M;
}

// ---------------------------------------------------------------------------
// Element (Detail/Element.hh:16-295)
// ---------------------------------------------------------------------------
template <int d, int N, int M, typename ScalarT, bool active_mode_>
struct Element
{
    static constexpr int n_element = d * N;
    static constexpr bool active_mode = active_mode_;
    using ScalarType = ScalarT;
    using VariableVectorType = Vec<ScalarT, d>;
    using PassiveVectorType = Vec<double, d>;
    using OutputVectorType = Vec<ScalarT, M>;

    Element() = default;
    Element(const Element&) = delete;             // Element.hh:78
    Element(Element&&) = default;
    Element& operator=(Element&&) = default;
    Element(Index _handle, const std::vector<double>& _x) : handle(_handle), x(&_x)
    {
        if constexpr (active_mode) idx_local_to_global.reserve(n_element);  // :205
    }

    Index global_idx(Index vh, Index off) const  // :159-170
    {
        const Index g = d * vh + off;
        if (g < 0 || g >= (Index)x->size()) error_throw("Element: variable index out of range");
        return g;
    }

    VariableVectorType variables(Index vh)  // :208-260
    {
        const Index start = global_idx(vh, 0);
        if (start + d > (Index)x->size()) error_throw("Element: variable segment out of range");
        Index local = -1;
        for (Index i = 0; i < (Index)idx_local_to_global.size(); ++i)
            if (idx_local_to_global[i] == start) { local = i; break; }
        if (local == -1)
        {
            if ((Index)idx_local_to_global.size() >= n_element)
                error_throw("Too many variables requested via element.variables(...).");
            local = (Index)idx_local_to_global.size();
            for (Index i = 0; i < d; ++i) idx_local_to_global.push_back(start + i);
        }
        VariableVectorType v;
        for (Index i = 0; i < d; ++i)
        {
            if constexpr (active_mode) v[i] = ScalarT((*x)[start + i], local + i);
            else v[i] = (*x)[start + i];
        }
        return v;
    }
    ScalarT variable(Index vh) { static_assert(d == 1); return variables(vh)[0]; }
    PassiveVectorType variables_passive(Index vh) const  // :272-285
    {
        const Index start = global_idx(vh, 0);
        PassiveVectorType v;
        for (Index i = 0; i < d; ++i) v[i] = (*x)[start + i];
        return v;
    }
    double variable_passive(Index vh) const { static_assert(d == 1); return variables_passive(vh)[0]; }

    Index handle = 0;
    std::vector<Index> idx_local_to_global;
    const std::vector<double>* x = nullptr;
};

#define ORACLE_SCALAR_TYPE(element) typename std::decay_t<decltype(element)>::ScalarType
#define ORACLE_VECTOR_TYPE(element) typename std::decay_t<decltype(element)>::OutputVectorType
#define ORACLE_ACTIVE_MODE(element) std::decay_t<decltype(element)>::active_mode

template <int k, bool wh>
bool all_finite(const Scalar<k, wh>& s, bool check_hess)
{
    for (double g : s.grad) if (!std::isfinite(g)) return false;
    if (check_hess) for (double h : s.Hess) if (!std::isfinite(h)) return false;
    return true;
}

// Phase timers filled by the second-order path:
// [0] parallel element evaluation + projection, [1] serial accumulation + triplet push, [2] COO->CSC
// norm_sum: like abs_sum, but every contribution of an element counts with the element's largest |grad| / |H| entry (the scale of the
// rounding error of an element-level computation: cancellation inside an element leaves errors relative to its largest intermediates).
// abs_sum (test metric only, SURVEY 8(c) "per-entry |delta| <= tol * sum |contributions|"): accumulate |grad_i| and |H_ij| instead of the
// signed values, so that g / H come back as the per-entry scale of the parity bound.
struct PhaseTimes { double eval = 0, accumulate = 0, compress = 0; Index n_projected = 0, n_decomposed = 0; bool abs_sum = false, norm_sum = false; };

// ---------------------------------------------------------------------------
// Scalar objective terms (Detail/ScalarObjectiveTerm.hh:22-289)
// ---------------------------------------------------------------------------
struct ScalarObjectiveTermBase
{
    virtual ~ScalarObjectiveTermBase() = default;
    virtual Index n_elements() const = 0;
    virtual double eval(const std::vector<double>& x) const = 0;
    virtual void eval_with_gradient_add(const std::vector<double>& x, double& f, std::vector<double>& g) const = 0;
    virtual void eval_with_derivatives_add(const std::vector<double>& x, double& f, std::vector<double>& g,
                                           std::vector<Triplet>& T, bool project, double eps, PhaseTimes* pt) const = 0;
};

inline double now_s()
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

template <int d, int N, typename F>
struct ScalarObjectiveTerm : ScalarObjectiveTermBase
{
    static constexpr int k = d * N;
    using PassiveElement = Element<d, N, 1, double, false>;
    using FirstElement = Element<d, N, 1, Scalar<k, false>, true>;
    using SecondElement = Element<d, N, 1, Scalar<k, true>, true>;

    ScalarObjectiveTerm(std::vector<Index> handles, F f, Index n_global, const EvalSettings& s)
        : n_vars_global(n_global), element_handles(std::move(handles)), settings(s), func(std::move(f)) {}

    Index n_elements() const override { return (Index)element_handles.size(); }

    double eval(const std::vector<double>& x) const override  // :162-186
    {
        std::vector<double> res(element_handles.size());
        parallel_for((Index)element_handles.size(), settings, [&](Index i) {
            PassiveElement element(element_handles[i], x);
            res[i] = func(element);
        });
        double f = 0.0;
        for (double r : res) f += r;
        return f;
    }

    void eval_with_gradient_add(const std::vector<double>& x, double& f, std::vector<double>& g) const override  // :188-222
    {
        std::vector<FirstElement> elements(element_handles.size());
        std::vector<Scalar<k, false>> results(element_handles.size());
        parallel_for((Index)element_handles.size(), settings, [&](Index i) {
            elements[i] = FirstElement(element_handles[i], x);
            results[i] = func(elements[i]);
            if (!all_finite(results[i], false)) error_throw("non-finite gradient");
        });
        for (size_t e = 0; e < element_handles.size(); ++e)
        {
            f += results[e].val;
            for (size_t i = 0; i < elements[e].idx_local_to_global.size(); ++i)
                g[elements[e].idx_local_to_global[i]] += results[e].grad[i];
        }
    }

    void eval_with_derivatives_add(const std::vector<double>& x, double& f, std::vector<double>& g,
                                   std::vector<Triplet>& T, bool project, double eps, PhaseTimes* pt) const override  // :224-278
    {
        const double t0 = now_s();
        std::vector<SecondElement> elements(element_handles.size());
        std::vector<Scalar<k, true>> results(element_handles.size());
        std::vector<unsigned char> proj_code(pt ? element_handles.size() : 0);
        parallel_for((Index)element_handles.size(), settings, [&](Index i) {
            elements[i] = SecondElement(element_handles[i], x);
            results[i] = func(elements[i]);
            if (project)
            {
                const int code = project_positive_definite(k, results[i].Hess.data(), eps);
                if (pt) proj_code[i] = (unsigned char)code;
            }
            if (!all_finite(results[i], true)) error_throw("non-finite gradient or Hessian");
        });
        const double t1 = now_s();
        for (size_t e = 0; e < element_handles.size(); ++e)
        {
            f += results[e].val;
            const auto& l2g = elements[e].idx_local_to_global;
            const bool abs_sum = pt && pt->abs_sum, norm_sum = pt && pt->norm_sum;
            double gmax = 0.0, hmax = 0.0;
            if (norm_sum)
            {
                for (double v : results[e].grad) gmax = std::max(gmax, std::fabs(v));
                for (double v : results[e].Hess) hmax = std::max(hmax, std::fabs(v));
            }
            for (size_t i = 0; i < l2g.size(); ++i)
                g[l2g[i]] += norm_sum ? gmax : (abs_sum ? std::fabs(results[e].grad[i]) : results[e].grad[i]);
            for (size_t i = 0; i < l2g.size(); ++i)
                for (size_t j = 0; j < l2g.size(); ++j)
                {
                    const double h = results[e].H((int)i, (int)j);
                    T.push_back(Triplet{(std::int32_t)l2g[i], (std::int32_t)l2g[j], norm_sum ? hmax : (abs_sum ? std::fabs(h) : h)});
                }
        }
        const double t2 = now_s();
        if (pt)
        {
            pt->eval += t1 - t0;
            pt->accumulate += t2 - t1;
            for (unsigned char c : proj_code) { pt->n_decomposed += (c >= 1); pt->n_projected += (c == 2); }
        }
    }

    const Index n_vars_global;
    const std::vector<Index> element_handles;
    const EvalSettings& settings;
    F func;
};

// ---------------------------------------------------------------------------
// ScalarFunction (ScalarFunction.hh:36-236, Detail/ScalarFunctionImpl.hh)
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Operations/SVD.hh:12-21, :70-99 -- sign() and the closed-form closest orthogonal 2x2 matrix U V^T
// ---------------------------------------------------------------------------
template <typename T>
int svd_sign(const T& x)
{
    if (x < T(0.0)) return -1;
    else if (x > T(0.0)) return 1;
    else return 0;
}

template <typename T>
Mat<T, 2, 2> closest_orthogonal(const Mat<T, 2, 2>& A)
{
    Mat<T, 2, 2> Su = A * A.transpose();                                        // :77
    T phi = 0.5 * atan2(Su(0, 1) + Su(1, 0), Su(0, 0) - Su(1, 1));              // :78
    T Cphi = cos(phi), Sphi = sin(phi);
    Mat<T, 2, 2> U;
    U(0, 0) = Cphi; U(0, 1) = -Sphi; U(1, 0) = Sphi; U(1, 1) = Cphi;            // :81-83
    Mat<T, 2, 2> Sw = A.transpose() * A;                                        // :86
    T theta = 0.5 * atan2(Sw(0, 1) + Sw(1, 0), Sw(0, 0) - Sw(1, 1));
    T Ctheta = cos(theta), Stheta = sin(theta);
    Mat<T, 2, 2> W;
    W(0, 0) = Ctheta; W(0, 1) = -Stheta; W(1, 0) = Stheta; W(1, 1) = Ctheta;    // :90-92
    Mat<T, 2, 2> S = U.transpose() * A * W;                                     // :95
    const double c0 = (double)svd_sign(S(0, 0)), c1 = (double)svd_sign(S(1, 1));
    Mat<T, 2, 2> V;                                                             // V = W * C.asDiagonal(), :96-97
    V(0, 0) = W(0, 0) * c0; V(0, 1) = W(0, 1) * c1; V(1, 0) = W(1, 0) * c0; V(1, 1) = W(1, 1) * c1;
    return U * V.transpose();                                                   // :99
}

// RecorderElement (ScalarFunctionImpl.hh:106-129): records which variable handles the functor accesses
template <int d>
struct RecorderElement
{
    using ScalarType = double;
    static constexpr bool active_mode = false;
    Index handle;
    std::vector<Index> accessed;
    explicit RecorderElement(Index h) : handle(h) {}
    Vec<double, d> variables(Index vh)
    {
        if (std::find(accessed.begin(), accessed.end(), vh) == accessed.end()) accessed.push_back(vh);
        return Vec<double, d>();
    }
    double variable(Index vh) { return variables(vh)[0]; }
    Vec<double, d> variables_passive(Index) { return Vec<double, d>(); }
    double variable_passive(Index) { return 0.0; }
};

inline std::vector<Index> range(Index n)  // Utils/Helpers.hh:17-27
{
    std::vector<Index> r(n);
    for (Index i = 0; i < n; ++i) r[i] = i;
    return r;
}

template <int d>
struct ScalarFunction
{
    ScalarFunction() = default;
    ScalarFunction(ScalarFunction&&) = default;
    ScalarFunction(const ScalarFunction&) = delete;
    explicit ScalarFunction(Index n_handles, const EvalSettings& s = EvalSettings())
        : settings(std::make_unique<EvalSettings>(s)), n_vars(d * n_handles) {}

    template <int N, typename F>
    void add_elements(const std::vector<Index>& handles, F f)  // ScalarFunctionImpl.hh:63-100
    {
        objective_terms.push_back(std::make_unique<ScalarObjectiveTerm<d, N, F>>(handles, std::move(f), n_vars, *settings));
        n_elements += (Index)handles.size();
    }

    // add_elements_dynamic (ScalarFunction.hh:80-108, ScalarFunctionImpl.hh:134-214): record how many variable handles each
    // element accesses (RecorderElement, :106-129), group the elements by the exact or next larger static valence, add one
    // term per non-empty group in the ORDER OF THE TEMPLATE ARGUMENTS (:190-213).
    template <int... ElementValences, typename F>
    void add_elements_dynamic(const std::vector<Index>& handles, F f)
    {
        std::vector<int> static_valences_sorted = {ElementValences...};
        std::sort(static_valences_sorted.begin(), static_valences_sorted.end());
        if (std::unique(static_valences_sorted.begin(), static_valences_sorted.end()) != static_valences_sorted.end())
            error_throw("Element valences passed to add_elements<..>(..) are not unique. Please pass unique element valences.");  // :158-160
        std::map<int, std::vector<Index>> groups;
        for (const Index e : handles)
        {
            RecorderElement<d> rec(e);
            f(rec);
            const int element_valence = (int)rec.accessed.size();
            auto it = std::lower_bound(static_valences_sorted.begin(), static_valences_sorted.end(), element_valence);
            if (it == static_valences_sorted.end())
                error_throw("Element valence exceeds maximum static valence passed to add_elements<..>(..).");  // :175-180
            groups[*it].push_back(e);
        }
        (add_dynamic_group<ElementValences>(groups, f), ...);
    }

    double eval(const std::vector<double>& x) const  // :256-273
    {
        check(x);
        double f = 0.0;
        for (auto& o : objective_terms)
        {
            if (f == INFINITY) return INFINITY;
            f += o->eval(x);
        }
        return f;
    }
    void eval_with_gradient(const std::vector<double>& x, double& f, std::vector<double>& g) const  // :284-299
    {
        check(x);
        f = 0.0;
        g.assign(n_vars, 0.0);
        for (auto& o : objective_terms) o->eval_with_gradient_add(x, f, g);
    }
    void eval_with_derivatives(const std::vector<double>& x, double& f, std::vector<double>& g, SparseMatrix& H, PhaseTimes* pt = nullptr) const  // :316-336
    {
        eval_second(x, f, g, H, false, NAN, pt);
    }
    void eval_with_hessian_proj(const std::vector<double>& x, double& f, std::vector<double>& g, SparseMatrix& H,
                                double eps = default_hessian_projection_eps, PhaseTimes* pt = nullptr) const  // :378-399
    {
        eval_second(x, f, g, H, true, eps, pt);
    }

    std::unique_ptr<EvalSettings> settings = std::make_unique<EvalSettings>();
    Index n_vars = 0;
    Index n_elements = 0;
    std::vector<std::unique_ptr<ScalarObjectiveTermBase>> objective_terms;

private:
    template <int N, typename F>
    void add_dynamic_group(const std::map<int, std::vector<Index>>& groups, const F& f)
    {
        const auto it = groups.find(N);
        if (it != groups.end()) add_elements<N>(it->second, f);
    }

    void check(const std::vector<double>& x) const
    {
        if ((Index)x.size() != n_vars) error_throw("x.size() != n_vars");
    }
    void eval_second(const std::vector<double>& x, double& f, std::vector<double>& g, SparseMatrix& H, bool project, double eps, PhaseTimes* pt) const
    {
        check(x);
        f = 0.0;
        g.assign(n_vars, 0.0);
        std::vector<Triplet> T;  // not reserved, like :393
        for (auto& o : objective_terms) o->eval_with_derivatives_add(x, f, g, T, project, eps, pt);
        const double t0 = now_s();
        H = set_from_triplets(n_vars, n_vars, T);
        if (pt) pt->compress += now_s() - t0;
    }
};

// ---------------------------------------------------------------------------
// Vector objective terms / VectorFunction (first-order part only; the
// per-residual Hessian tensor VectorObjectiveTerm.hh:245-324 is out of scope)
// ---------------------------------------------------------------------------
struct VectorObjectiveTermBase
{
    virtual ~VectorObjectiveTermBase() = default;
    virtual Index n_elements() const = 0;
    virtual Index n_outputs() const = 0;
    virtual void eval(const std::vector<double>& x, double* r) const = 0;
    virtual void eval_with_jacobian_add(const std::vector<double>& x, std::vector<double>& r, std::vector<Triplet>& T) const = 0;
    virtual double eval_sum_of_squares(const std::vector<double>& x) const = 0;
};

template <int d, int N, int M, typename F>
struct VectorObjectiveTerm : VectorObjectiveTermBase
{
    static constexpr int k = d * N;
    using PassiveElement = Element<d, N, M, double, false>;
    using FirstElement = Element<d, N, M, Scalar<k, false>, true>;

    VectorObjectiveTerm(std::vector<Index> handles, F f, Index n_global, const EvalSettings& s)
        : n_vars_global(n_global), element_handles(std::move(handles)), settings(s), func(std::move(f)) {}

    Index n_elements() const override { return (Index)element_handles.size(); }
    Index n_outputs() const override { return M * (Index)element_handles.size(); }

    void eval(const std::vector<double>& x, double* r) const override  // VectorObjectiveTerm.hh:158-178
    {
        parallel_for((Index)element_handles.size(), settings, [&](Index i) {
            PassiveElement element(element_handles[i], x);
            const Vec<double, M> res = func(element);
            for (int j = 0; j < M; ++j) r[M * i + j] = res[j];
        });
    }
    void eval_with_jacobian_add(const std::vector<double>& x, std::vector<double>& r, std::vector<Triplet>& T) const override  // :180-243
    {
        std::vector<FirstElement> elements(element_handles.size());
        std::vector<Vec<Scalar<k, false>, M>> results(element_handles.size());
        parallel_for((Index)element_handles.size(), settings, [&](Index i) {
            elements[i] = FirstElement(element_handles[i], x);
            results[i] = func(elements[i]);
            for (int j = 0; j < M; ++j)
                if (!all_finite(results[i][j], false)) error_throw("non-finite Jacobian");
        });
        const Index start = (Index)r.size();
        r.resize(r.size() + M * element_handles.size());
        Index i_res = 0;
        for (size_t e = 0; e < element_handles.size(); ++e)
            for (int j = 0; j < M; ++j)
            {
                const auto& res = results[e][j];
                r[start + i_res] = res.val;
                for (size_t v = 0; v < elements[e].idx_local_to_global.size(); ++v)
                    T.push_back(Triplet{(std::int32_t)(start + i_res), (std::int32_t)elements[e].idx_local_to_global[v], res.grad[v]});
                ++i_res;
            }
    }
    double eval_sum_of_squares(const std::vector<double>& x) const override  // :326-351
    {
        std::vector<double> sq(element_handles.size());
        parallel_for((Index)element_handles.size(), settings, [&](Index i) {
            PassiveElement element(element_handles[i], x);
            const Vec<double, M> res = func(element);
            sq[i] = res.dot(res);
        });
        double result = 0.0;
        for (double s : sq) result += s;
        return result;
    }

    const Index n_vars_global;
    const std::vector<Index> element_handles;
    const EvalSettings& settings;
    F func;
};

template <int d>
struct VectorFunction
{
    VectorFunction() = default;
    VectorFunction(VectorFunction&&) = default;
    explicit VectorFunction(Index n_handles, const EvalSettings& s = EvalSettings())
        : settings(std::make_unique<EvalSettings>(s)), n_vars(d * n_handles) {}

    template <int N, int M, typename F>
    void add_elements(const std::vector<Index>& handles, F f)  // VectorFunctionImpl.hh:64-101
    {
        objective_terms.push_back(std::make_unique<VectorObjectiveTerm<d, N, M, F>>(handles, std::move(f), n_vars, *settings));
        n_elements += (Index)handles.size();
        n_outputs += M * (Index)handles.size();
    }
    std::vector<double> eval(const std::vector<double>& x) const  // :143-159
    {
        std::vector<double> r(n_outputs);
        Index start = 0;
        for (auto& o : objective_terms) { o->eval(x, r.data() + start); start += o->n_outputs(); }
        return r;
    }
    void eval_with_jacobian(const std::vector<double>& x, std::vector<double>& r, SparseMatrix& J) const  // :170-188
    {
        r.clear();
        std::vector<Triplet> T;
        for (auto& o : objective_terms) o->eval_with_jacobian_add(x, r, T);
        J = set_from_triplets((Index)r.size(), n_vars, T);
    }
    double eval_sum_of_squares(const std::vector<double>& x) const  // :254-266
    {
        double result = 0.0;
        for (auto& o : objective_terms) result += o->eval_sum_of_squares(x);
        return result;
    }
    void eval_sum_of_squares_with_derivatives(const std::vector<double>& x, double& f, std::vector<double>& g,
                                              std::vector<double>& r, SparseMatrix& J) const  // :268-283
    {
        eval_with_jacobian(x, r, J);
        f = 0.0;
        for (double v : r) f += v * v;
        g.assign(n_vars, 0.0);  // g = 2 J^T r
        for (Index c = 0; c < J.cols; ++c)
        {
            double s = 0.0;
            for (Index p = J.outer[c]; p < J.outer[c + 1]; ++p) s += J.values[p] * r[J.inner[p]];
            g[c] = 2.0 * s;
        }
    }

    std::unique_ptr<EvalSettings> settings = std::make_unique<EvalSettings>();
    Index n_vars = 0, n_elements = 0, n_outputs = 0;
    std::vector<std::unique_ptr<VectorObjectiveTermBase>> objective_terms;
};

}  // namespace oracle
