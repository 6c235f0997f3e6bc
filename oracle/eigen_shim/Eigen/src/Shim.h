// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the part of the Eigen 3 API that the reference's headers
// (include/TinyAD/{Scalar,ScalarFunction,VectorFunction}.hh, Detail/*.hh, Utils/HessianProjection.hh, Utils/Helpers.hh,
// Operations/SVD.hh) and small element lambdas use.  Eigen itself is not in this image and cannot be fetched; with this
// shim on the include path the UNMODIFIED reference headers compile where they lie (oracle/Makefile, target _ref), which
// gives the oracle a second, independent implementation to be checked against (tests/test_oracle_vs_reference.py).
//
// This is NOT Eigen: every expression is evaluated eagerly into a plain column-major matrix, nothing is vectorised and
// only the members listed below exist.  Algorithms are textbook ones written for this file (cofactor inverse /
// determinant, Householder tridiagonalisation + implicit QL for SelfAdjointEigenSolver, ordered duplicate summation in
// setFromTriplets as documented for Eigen 3.4).  Run time of code built on it says nothing about real Eigen.
#pragma once

// the standard headers Eigen/Core itself pulls in (the reference relies on them transitively)
#include <algorithm>
#include <array>
#include <cassert>
#include <climits>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <numeric>
#include <sstream>
#include <string>
#include <ostream>
#include <stdexcept>
#include <type_traits>
#include <utility>
#include <vector>

#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3
#define EIGEN_MINOR_VERSION 0
#define TINYAD_EIGEN_SHIM 1

namespace Eigen
{

using Index = std::ptrdiff_t;
constexpr int Dynamic = -1;
constexpr int Infinity = -1;
enum { ColMajor = 0, RowMajor = 1 };
enum { Lower = 1, Upper = 2 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

template <typename T>
struct NumTraits
{
    typedef T Real;
    typedef T NonInteger;
    typedef T Nested;
    typedef T Literal;
    enum
    {
        IsComplex = 0,
        IsInteger = std::is_integral<T>::value ? 1 : 0,
        IsSigned = std::is_signed<T>::value ? 1 : 0,
        RequireInitialization = 0,
        ReadCost = 1,
        AddCost = 1,
        MulCost = 1
    };
    static T epsilon() { return std::numeric_limits<T>::epsilon(); }
    static T dummy_precision() { return T(1e-12); }
    static T highest() { return (std::numeric_limits<T>::max)(); }
    static T lowest() { return std::numeric_limits<T>::lowest(); }
    static int digits10() { return std::numeric_limits<T>::digits10; }
};

template <typename A, typename B, typename BinaryOp = void>
struct ScalarBinaryOpTraits
{
};
template <typename T, typename BinaryOp>
struct ScalarBinaryOpTraits<T, T, BinaryOp>
{
    typedef T ReturnType;
};

struct EigenTag
{
};
template <typename X>
struct is_eigen : std::is_base_of<EigenTag, std::decay_t<X>>
{
};

template <typename T, int R, int C, int Options = 0, int MaxR = R, int MaxC = C>
class Matrix;
template <typename M>
class Map;
template <typename X, int BR, int BC>
class Block;
template <typename V>
class DiagonalWrapper;
template <typename Derived>
class MatrixBase;
template <typename Derived>
class CommaInitializer;

namespace internal
{

template <typename D>
struct traits;
template <typename T, int R, int C, int O, int MR, int MC>
struct traits<Matrix<T, R, C, O, MR, MC>>
{
    using Scalar = T;
    enum { Rows = R, Cols = C };
};
template <typename M>
struct traits<Map<M>> : traits<std::remove_const_t<M>>
{
};
template <typename X, int BR, int BC>
struct traits<Block<X, BR, BC>>
{
    using Scalar = typename traits<std::remove_const_t<X>>::Scalar;
    enum { Rows = BR, Cols = BC };
};

constexpr int pick_dim(int a, int b) { return a != Dynamic ? a : b; }
constexpr int mul_dim(int a, int b) { return (a == Dynamic || b == Dynamic) ? Dynamic : a * b; }

template <typename T, int R, int C, bool Fixed = (R >= 0 && C >= 0)>
struct Storage;

template <typename T, int R, int C>
struct Storage<T, R, C, true>
{
    std::array<T, (std::size_t)(R * C)> a{};
    static constexpr Index rows() { return R; }
    static constexpr Index cols() { return C; }
    void resize(Index r, Index c)
    {
        if (r != R || c != C) throw std::logic_error("eigen shim: resize of a fixed-size matrix");
    }
    void conservative_resize(Index r, Index c) { resize(r, c); }
    T* data() { return a.data(); }
    const T* data() const { return a.data(); }
};

template <typename T, int R, int C>
struct Storage<T, R, C, false>
{
    std::vector<T> a;
    Index r = R < 0 ? 0 : R, c = C < 0 ? 0 : C;
    Index rows() const { return r; }
    Index cols() const { return c; }
    void resize(Index r_, Index c_)
    {
        if ((R >= 0 && r_ != R) || (C >= 0 && c_ != C)) throw std::logic_error("eigen shim: resize against a fixed dimension");
        if (r_ == r && c_ == c && (Index)a.size() == r_ * c_) return;
        r = r_;
        c = c_;
        a.assign((std::size_t)(r * c), T());
    }
    void conservative_resize(Index r_, Index c_)
    {
        if (c_ == c || r == 0 || c == 0 || c_ == 1)
        {
            // vectors and column appends keep their linear layout
            if (c_ != c && !(c_ == 1 || c == 0 || r == 0))
                throw std::logic_error("eigen shim: conservativeResize of a matrix is not supported");
            if (c_ == c && r_ != r && c > 1) throw std::logic_error("eigen shim: conservativeResize of a matrix is not supported");
            a.resize((std::size_t)(r_ * c_), T());
            r = r_;
            c = c_;
            return;
        }
        throw std::logic_error("eigen shim: conservativeResize of a matrix is not supported");
    }
    T* data() { return a.data(); }
    const T* data() const { return a.data(); }
};

// 2/3/4 determinant and cofactor inverse of small matrices, generic elimination otherwise
template <typename M>
typename M::Scalar determinant_of(const M& m)
{
    using T = typename M::Scalar;
    const Index n = m.rows();
    if (n != m.cols()) throw std::logic_error("eigen shim: determinant of a non-square matrix");
    if (n == 0) return T(1.0);
    if (n == 1) return m(0, 0);
    if (n == 2) return m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1);
    if (n == 3)
    {
        return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
               m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
    }
    // Laplace expansion along the first row (small n only)
    T det = T(0.0);
    for (Index j = 0; j < n; ++j)
    {
        Matrix<T, Dynamic, Dynamic> sub(n - 1, n - 1);
        for (Index r = 1; r < n; ++r)
            for (Index c = 0, cc = 0; c < n; ++c)
                if (c != j) sub(r - 1, cc++) = m(r, c);
        T term = m(0, j) * determinant_of(sub);
        det = (j % 2 == 0) ? T(det + term) : T(det - term);
    }
    return det;
}

}  // namespace internal

// --------------------------------------------------------------------------------------------------------------------
// MatrixBase: everything that only reads coefficients
// --------------------------------------------------------------------------------------------------------------------
template <typename Derived>
class MatrixBase : public EigenTag
{
public:
    using Scalar = typename internal::traits<Derived>::Scalar;
    enum
    {
        RowsAtCompileTime = internal::traits<Derived>::Rows,
        ColsAtCompileTime = internal::traits<Derived>::Cols,
        SizeAtCompileTime = internal::mul_dim(internal::traits<Derived>::Rows, internal::traits<Derived>::Cols),
        IsVectorAtCompileTime = (internal::traits<Derived>::Rows == 1 || internal::traits<Derived>::Cols == 1) ? 1 : 0
    };
    using PlainObject = Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime>;
    using PlainMatrix = PlainObject;
    using RealScalar = Scalar;
    using TransposeType = Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime>;

    Derived& derived() { return *static_cast<Derived*>(this); }
    const Derived& derived() const { return *static_cast<const Derived*>(this); }

    Index rows() const { return derived().rows_impl(); }
    Index cols() const { return derived().cols_impl(); }
    Index size() const { return rows() * cols(); }

    decltype(auto) operator()(Index i, Index j) const { return derived().coeff(i, j); }
    decltype(auto) operator()(Index i, Index j) { return derived().coeffRef(i, j); }
    decltype(auto) operator()(Index i) const { return lin(i); }
    decltype(auto) operator()(Index i) { return lin(i); }
    decltype(auto) operator[](Index i) const { return lin(i); }
    decltype(auto) operator[](Index i) { return lin(i); }
    decltype(auto) coeff(Index i) const { return lin(i); }
    decltype(auto) x() const { return lin(0); }
    decltype(auto) y() const { return lin(1); }
    decltype(auto) z() const { return lin(2); }
    decltype(auto) w() const { return lin(3); }
    decltype(auto) x() { return lin(0); }
    decltype(auto) y() { return lin(1); }
    decltype(auto) z() { return lin(2); }
    decltype(auto) w() { return lin(3); }

    PlainObject eval() const { return PlainObject(*this); }

    TransposeType transpose() const
    {
        TransposeType t;
        t.resize(cols(), rows());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) t(j, i) = (*this)(i, j);
        return t;
    }
    TransposeType adjoint() const { return transpose(); }

    // ---- blocks (views) ----
    template <int N>
    auto segment(Index start) { return vec_block<Derived, N>(derived(), start, N); }
    template <int N>
    auto segment(Index start) const { return vec_block<const Derived, N>(derived(), start, N); }
    template <int N>
    auto segment(Index start, Index n) { return vec_block<Derived, N>(derived(), start, n); }
    template <int N>
    auto segment(Index start, Index n) const { return vec_block<const Derived, N>(derived(), start, n); }
    auto segment(Index start, Index n) { return vec_block<Derived, Dynamic>(derived(), start, n); }
    auto segment(Index start, Index n) const { return vec_block<const Derived, Dynamic>(derived(), start, n); }
    template <int N>
    auto head() { return segment<N>(0); }
    template <int N>
    auto head() const { return segment<N>(0); }
    auto head(Index n) { return segment(0, n); }
    auto head(Index n) const { return segment(0, n); }
    template <int N>
    auto tail() { return segment<N>(size() - N); }
    template <int N>
    auto tail() const { return segment<N>(size() - N); }
    auto tail(Index n) { return segment(size() - n, n); }
    auto tail(Index n) const { return segment(size() - n, n); }

    template <int BR, int BC>
    Block<Derived, BR, BC> block(Index i, Index j) { return Block<Derived, BR, BC>(derived(), i, j, BR, BC); }
    template <int BR, int BC>
    Block<const Derived, BR, BC> block(Index i, Index j) const { return Block<const Derived, BR, BC>(derived(), i, j, BR, BC); }
    Block<Derived, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) { return Block<Derived, Dynamic, Dynamic>(derived(), i, j, r, c); }
    Block<const Derived, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) const
    {
        return Block<const Derived, Dynamic, Dynamic>(derived(), i, j, r, c);
    }
    template <int BR, int BC>
    auto topLeftCorner() { return block<BR, BC>(0, 0); }
    template <int BR, int BC>
    auto topLeftCorner() const { return block<BR, BC>(0, 0); }
    Block<Derived, 1, ColsAtCompileTime> row(Index i) { return Block<Derived, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols()); }
    Block<const Derived, 1, ColsAtCompileTime> row(Index i) const { return Block<const Derived, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols()); }
    Block<Derived, RowsAtCompileTime, 1> col(Index j) { return Block<Derived, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1); }
    Block<const Derived, RowsAtCompileTime, 1> col(Index j) const { return Block<const Derived, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1); }
    Matrix<Scalar, internal::pick_dim(RowsAtCompileTime, ColsAtCompileTime), 1> diagonal() const
    {
        Matrix<Scalar, internal::pick_dim(RowsAtCompileTime, ColsAtCompileTime), 1> d;
        d.resize((std::min)(rows(), cols()), 1);
        for (Index i = 0; i < d.size(); ++i) d[i] = (*this)(i, i);
        return d;
    }

    // ---- reductions ----
    Scalar sum() const
    {
        if (size() == 0) return Scalar(0);
        Scalar s = lin_rc(0);
        for (Index i = 1; i < size(); ++i) s = s + lin_rc(i);
        return s;
    }
    Scalar prod() const
    {
        Scalar s = Scalar(1);
        for (Index i = 0; i < size(); ++i) s = s * lin_rc(i);
        return s;
    }
    Scalar mean() const { return sum() / Scalar((double)size()); }
    Scalar trace() const
    {
        Scalar s = (*this)(0, 0);
        for (Index i = 1; i < (std::min)(rows(), cols()); ++i) s = s + (*this)(i, i);
        return s;
    }
    Scalar squaredNorm() const
    {
        if (size() == 0) return Scalar(0);
        Scalar s = lin_rc(0) * lin_rc(0);
        for (Index i = 1; i < size(); ++i) s = s + lin_rc(i) * lin_rc(i);
        return s;
    }
    Scalar norm() const
    {
        using std::sqrt;
        return sqrt(squaredNorm());
    }
    PlainObject normalized() const { return PlainObject(*this) / norm(); }
    template <typename O>
    auto dot(const MatrixBase<O>& o) const
    {
        using RT = decltype(std::declval<Scalar>() * std::declval<typename O::Scalar>());
        if (size() != o.size()) throw std::logic_error("eigen shim: dot of different sizes");
        if (size() == 0) return RT(0);
        RT s = lin(0) * o[0];
        for (Index i = 1; i < size(); ++i) s = s + lin(i) * o[i];
        return s;
    }
    template <typename O>
    auto cross(const MatrixBase<O>& o) const
    {
        using RT = decltype(std::declval<Scalar>() * std::declval<typename O::Scalar>());
        if (size() != 3 || o.size() != 3) throw std::logic_error("eigen shim: cross needs 3-vectors");
        Matrix<RT, 3, 1> r;
        r[0] = lin(1) * o[2] - lin(2) * o[1];
        r[1] = lin(2) * o[0] - lin(0) * o[2];
        r[2] = lin(0) * o[1] - lin(1) * o[0];
        return r;
    }
    Scalar maxCoeff() const
    {
        Scalar m = lin_rc(0);
        for (Index i = 1; i < size(); ++i)
            if (m < lin_rc(i)) m = lin_rc(i);
        return m;
    }
    Scalar minCoeff() const
    {
        Scalar m = lin_rc(0);
        for (Index i = 1; i < size(); ++i)
            if (lin_rc(i) < m) m = lin_rc(i);
        return m;
    }
    bool allFinite() const
    {
        using std::isfinite;
        for (Index i = 0; i < size(); ++i)
            if (!isfinite(lin_rc(i))) return false;
        return true;
    }
    bool hasNaN() const
    {
        for (Index i = 0; i < size(); ++i)
            if (!(lin_rc(i) == lin_rc(i))) return true;
        return false;
    }
    bool isZero(double prec = 1e-12) const
    {
        using std::abs;
        for (Index i = 0; i < size(); ++i)
            if ((double)abs(lin_rc(i)) > prec) return false;
        return true;
    }
    Matrix<Scalar, Dynamic, Dynamic> replicate(Index rf, Index cf) const
    {
        Matrix<Scalar, Dynamic, Dynamic> r(rows() * rf, cols() * cf);
        for (Index j = 0; j < r.cols(); ++j)
            for (Index i = 0; i < r.rows(); ++i) r(i, j) = (*this)(i % rows(), j % cols());
        return r;
    }
    template <typename O>
    bool isApprox(const MatrixBase<O>& o, double prec = 1e-12) const
    {
        double d = 0.0, a = 0.0, b = 0.0;
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i)
            {
                const double u = (double)(*this)(i, j), v = (double)o(i, j);
                d += (u - v) * (u - v);
                a += u * u;
                b += v * v;
            }
        return d <= prec * prec * (std::min)(a, b);
    }

    Scalar determinant() const { return internal::determinant_of(derived()); }

    PlainObject inverse() const
    {
        const Index n = rows();
        if (n != cols()) throw std::logic_error("eigen shim: inverse of a non-square matrix");
        PlainObject r;
        r.resize(n, n);
        const Derived& m = derived();
        if (n == 1)
        {
            r(0, 0) = Scalar(1.0) / m(0, 0);
        }
        else if (n == 2)
        {
            const Scalar invdet = Scalar(1.0) / determinant();
            r(0, 0) = m(1, 1) * invdet;
            r(1, 0) = -m(1, 0) * invdet;
            r(0, 1) = -m(0, 1) * invdet;
            r(1, 1) = m(0, 0) * invdet;
        }
        else if (n == 3)
        {
            // cofactors; the determinant is the first column times the first cofactor row
            const Scalar c00 = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1);
            const Scalar c10 = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2);
            const Scalar c20 = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
            const Scalar invdet = Scalar(1.0) / (m(0, 0) * c00 + m(0, 1) * c10 + m(0, 2) * c20);
            r(0, 0) = c00 * invdet;
            r(1, 0) = c10 * invdet;
            r(2, 0) = c20 * invdet;
            r(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) * invdet;
            r(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) * invdet;
            r(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) * invdet;
            r(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) * invdet;
            r(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) * invdet;
            r(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) * invdet;
        }
        else
        {
            // Gauss-Jordan with partial pivoting on the magnitude of the passive value
            Matrix<Scalar, Dynamic, Dynamic> a(n, n);
            for (Index j = 0; j < n; ++j)
                for (Index i = 0; i < n; ++i)
                {
                    a(i, j) = m(i, j);
                    r(i, j) = Scalar(i == j ? 1.0 : 0.0);
                }
            for (Index c = 0; c < n; ++c)
            {
                Index p = c;
                for (Index i = c + 1; i < n; ++i)
                    if (abs_of(a(p, c)) < abs_of(a(i, c))) p = i;
                if (p != c)
                    for (Index j = 0; j < n; ++j)
                    {
                        std::swap(a(p, j), a(c, j));
                        std::swap(r(p, j), r(c, j));
                    }
                const Scalar inv = Scalar(1.0) / a(c, c);
                for (Index j = 0; j < n; ++j)
                {
                    a(c, j) = a(c, j) * inv;
                    r(c, j) = r(c, j) * inv;
                }
                for (Index i = 0; i < n; ++i)
                    if (i != c)
                    {
                        const Scalar fac = a(i, c);
                        for (Index j = 0; j < n; ++j)
                        {
                            a(i, j) = a(i, j) - fac * a(c, j);
                            r(i, j) = r(i, j) - fac * r(c, j);
                        }
                    }
            }
        }
        return r;
    }

    // ---- coefficient-wise ----
    template <typename F>
    auto unaryExpr(F f) const
    {
        using RT = std::decay_t<decltype(f(std::declval<Scalar>()))>;
        Matrix<RT, RowsAtCompileTime, ColsAtCompileTime> r;
        r.resize(rows(), cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r(i, j) = f((*this)(i, j));
        return r;
    }
    template <typename U>
    Matrix<U, RowsAtCompileTime, ColsAtCompileTime> cast() const
    {
        return unaryExpr([](const Scalar& s) { return (U)s; });
    }
    PlainObject cwiseAbs() const
    {
        return unaryExpr([](const Scalar& s) {
            using std::abs;
            return (Scalar)abs(s);
        });
    }
    PlainObject cwiseAbs2() const
    {
        return unaryExpr([](const Scalar& s) { return (Scalar)(s * s); });
    }
    PlainObject cwiseSqrt() const
    {
        return unaryExpr([](const Scalar& s) {
            using std::sqrt;
            return (Scalar)sqrt(s);
        });
    }
    PlainObject cwiseInverse() const
    {
        return unaryExpr([](const Scalar& s) { return (Scalar)(Scalar(1.0) / s); });
    }
    PlainObject abs() const { return cwiseAbs(); }     // reachable through .array()
    PlainObject square() const { return cwiseAbs2(); }
    template <typename O>
    auto cwiseProduct(const MatrixBase<O>& o) const
    {
        using RT = decltype(std::declval<Scalar>() * std::declval<typename O::Scalar>());
        Matrix<RT, RowsAtCompileTime, ColsAtCompileTime> r;
        r.resize(rows(), cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r(i, j) = (*this)(i, j) * o(i, j);
        return r;
    }
    template <typename O>
    auto cwiseQuotient(const MatrixBase<O>& o) const
    {
        using RT = decltype(std::declval<Scalar>() / std::declval<typename O::Scalar>());
        Matrix<RT, RowsAtCompileTime, ColsAtCompileTime> r;
        r.resize(rows(), cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r(i, j) = (*this)(i, j) / o(i, j);
        return r;
    }
    const Derived& array() const { return derived(); }
    const Derived& matrix() const { return derived(); }
    const Derived& noalias() const { return derived(); }
    Derived& noalias() { return derived(); }

    DiagonalWrapper<PlainObject> asDiagonal() const;

    // ---- writers ----
    Derived& setZero()
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) (*this)(i, j) = Scalar(0);
        return derived();
    }
    Derived& setConstant(const Scalar& v)
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) (*this)(i, j) = v;
        return derived();
    }
    Derived& setOnes() { return setConstant(Scalar(1)); }
    Derived& fill(const Scalar& v) { return setConstant(v); }
    Derived& setIdentity()
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) (*this)(i, j) = Scalar(i == j ? 1 : 0);
        return derived();
    }
    void normalize()
    {
        const Scalar n = norm();
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) (*this)(i, j) = (*this)(i, j) / n;
    }
    template <typename O>
    Derived& operator+=(const MatrixBase<O>& o)
    {
        check_same(o);
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) (*this)(i, j) = (*this)(i, j) + o(i, j);
        return derived();
    }
    template <typename O>
    Derived& operator-=(const MatrixBase<O>& o)
    {
        check_same(o);
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) (*this)(i, j) = (*this)(i, j) - o(i, j);
        return derived();
    }
    template <typename S, typename = std::enable_if_t<!is_eigen<S>::value>>
    Derived& operator*=(const S& s)
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) (*this)(i, j) = (*this)(i, j) * s;
        return derived();
    }
    template <typename S, typename = std::enable_if_t<!is_eigen<S>::value>>
    Derived& operator/=(const S& s)
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) (*this)(i, j) = (*this)(i, j) / s;
        return derived();
    }

    // comma initialiser
    template <typename A>
    CommaInitializer<Derived> operator<<(const A& a);

    // element-wise copy used by every operator= ; vectors may be assigned across orientation
    template <typename O>
    void assign_from(const MatrixBase<O>& o)
    {
        if (rows() == o.rows() && cols() == o.cols())
        {
            for (Index j = 0; j < cols(); ++j)
                for (Index i = 0; i < rows(); ++i) (*this)(i, j) = Scalar(o(i, j));
        }
        else if ((rows() == 1 || cols() == 1) && (o.rows() == 1 || o.cols() == 1) && size() == o.size())
        {
            for (Index i = 0; i < size(); ++i) (*this)[i] = Scalar(o[i]);
        }
        else
            throw std::logic_error("eigen shim: assignment of different shapes");
    }

protected:
    template <typename O>
    void check_same(const MatrixBase<O>& o) const
    {
        if (rows() != o.rows() || cols() != o.cols()) throw std::logic_error("eigen shim: operands of different shapes");
    }
    static Scalar abs_of(const Scalar& s)
    {
        using std::abs;
        return abs(s);
    }
    // linear (vector) index
    decltype(auto) lin(Index i) const { return cols() == 1 ? derived().coeff(i, 0) : (rows() == 1 ? derived().coeff(0, i) : derived().coeff(i % rows(), i / rows())); }
    decltype(auto) lin(Index i) { return cols() == 1 ? derived().coeffRef(i, 0) : (rows() == 1 ? derived().coeffRef(0, i) : derived().coeffRef(i % rows(), i / rows())); }
    // column-major traversal index
    decltype(auto) lin_rc(Index i) const { return derived().coeff(i % rows(), i / rows()); }

    template <typename X, int N>
    static auto vec_block(X& x, Index start, Index n)
    {
        if constexpr (ColsAtCompileTime == 1)
            return Block<X, N, 1>(x, start, 0, n, 1);
        else if constexpr (RowsAtCompileTime == 1)
            return Block<X, 1, N>(x, 0, start, 1, n);
        else
        {
            if (x.cols() == 1) return Block<X, N, 1>(x, start, 0, n, 1);
            throw std::logic_error("eigen shim: segment of a matrix");
        }
    }
};

// --------------------------------------------------------------------------------------------------------------------
// Matrix
// --------------------------------------------------------------------------------------------------------------------
template <typename T, int R, int C, int Options, int MaxR, int MaxC>
class Matrix : public MatrixBase<Matrix<T, R, C, Options, MaxR, MaxC>>
{
    static constexpr bool fixed = (R >= 0 && C >= 0);
    internal::Storage<T, R, C> s_;

public:
    using Base = MatrixBase<Matrix>;
    using Scalar = T;
    using StorageIndex = int;

    Matrix() = default;
    Matrix(const Matrix&) = default;
    Matrix(Matrix&&) = default;
    Matrix& operator=(const Matrix&) = default;
    Matrix& operator=(Matrix&&) = default;

    template <typename D>
    Matrix(const MatrixBase<D>& o)
    {
        resize_like(o);
        this->assign_from(o);
    }
    template <typename V>
    Matrix(const DiagonalWrapper<V>& d);

    // one argument that is not a matrix: a size (vectors) or the single coefficient of a 1x1
    template <typename A, typename = std::enable_if_t<!is_eigen<A>::value>>
    explicit Matrix(const A& a)
    {
        if constexpr (fixed && R * C == 1 && std::is_convertible<A, T>::value)
            s_.a[0] = T(a);
        else if constexpr (std::is_convertible<A, Index>::value)
        {
            if constexpr (C == 1 || R != 1)
                resize((Index)a, 1);
            else
                resize(1, (Index)a);
        }
        else
            static_assert(sizeof(A) == 0, "eigen shim: unsupported constructor argument");
    }
    // two arguments: two coefficients for fixed size-2 vectors, (rows, cols) otherwise
    template <typename A, typename B, typename = std::enable_if_t<!is_eigen<A>::value && !is_eigen<B>::value>>
    Matrix(const A& a, const B& b)
    {
        if constexpr (fixed && R * C == 2)
        {
            s_.a[0] = T(a);
            s_.a[1] = T(b);
        }
        else
            resize((Index)a, (Index)b);
    }
    template <typename A, typename B, typename D, typename = std::enable_if_t<!is_eigen<A>::value>>
    Matrix(const A& a, const B& b, const D& c)
    {
        resize(R == 1 ? 1 : 3, R == 1 ? 3 : 1);
        s_.data()[0] = T(a);
        s_.data()[1] = T(b);
        s_.data()[2] = T(c);
    }
    template <typename A, typename B, typename D, typename E, typename = std::enable_if_t<!is_eigen<A>::value>>
    Matrix(const A& a, const B& b, const D& c, const E& d)
    {
        resize(R == 1 ? 1 : 4, R == 1 ? 4 : 1);
        s_.data()[0] = T(a);
        s_.data()[1] = T(b);
        s_.data()[2] = T(c);
        s_.data()[3] = T(d);
    }

    // { a, b, c, ... }: the coefficients of a vector
    Matrix(std::initializer_list<T> l)
    {
        if constexpr (!fixed) resize_vec((Index)l.size());
        if ((Index)l.size() != this->size()) throw std::logic_error("eigen shim: initializer list of the wrong length");
        Index i = 0;
        for (const T& v : l) s_.data()[i++] = v;
    }

    template <typename D>
    Matrix& operator=(const MatrixBase<D>& o)
    {
        resize_like(o);
        this->assign_from(o);
        return *this;
    }
    template <typename V>
    Matrix& operator=(const DiagonalWrapper<V>& d)
    {
        *this = Matrix(d);
        return *this;
    }

    Index rows_impl() const { return s_.rows(); }
    Index cols_impl() const { return s_.cols(); }
    const T& coeff(Index i, Index j) const { return s_.data()[i + j * s_.rows()]; }
    T& coeffRef(Index i, Index j) { return s_.data()[i + j * s_.rows()]; }
    using Base::coeff;
    T* data() { return s_.data(); }
    const T* data() const { return s_.data(); }
    Index innerStride() const { return 1; }
    Index outerStride() const { return s_.rows(); }

    void resize(Index r, Index c) { s_.resize(r, c); }
    void resize(Index n)
    {
        if constexpr (R == 1 && C != 1)
            s_.resize(1, n);
        else
            s_.resize(n, 1);
    }
    void conservativeResize(Index n)
    {
        if constexpr (R == 1 && C != 1)
            s_.conservative_resize(1, n);
        else
            s_.conservative_resize(n, 1);
    }
    void conservativeResize(Index r, Index c) { s_.conservative_resize(r, c); }

    // ---- named constructors ----
    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(Index n)
    {
        Matrix m;
        m.resize_vec(n);
        m.setZero();
        return m;
    }
    static Matrix Zero(Index r, Index c)
    {
        Matrix m;
        m.resize(r, c);
        m.setZero();
        return m;
    }
    static Matrix Constant(const T& v)
    {
        Matrix m;
        m.setConstant(v);
        return m;
    }
    static Matrix Constant(Index n, const T& v)
    {
        Matrix m;
        m.resize_vec(n);
        m.setConstant(v);
        return m;
    }
    static Matrix Constant(Index r, Index c, const T& v)
    {
        Matrix m;
        m.resize(r, c);
        m.setConstant(v);
        return m;
    }
    // uniform in [-1, 1] from a fixed-seed generator (tests only ask for "some" matrix)
    static Matrix Random() { return Matrix().randomize(); }
    static Matrix Random(Index n)
    {
        Matrix m;
        m.resize_vec(n);
        return m.randomize();
    }
    static Matrix Random(Index r, Index c)
    {
        Matrix m;
        m.resize(r, c);
        return m.randomize();
    }
    Matrix& setRandom() { return randomize(); }
    static Matrix Ones() { return Constant(T(1)); }
    static Matrix Ones(Index n) { return Constant(n, T(1)); }
    static Matrix Ones(Index r, Index c) { return Constant(r, c, T(1)); }
    static Matrix Identity()
    {
        Matrix m;
        m.setIdentity();
        return m;
    }
    static Matrix Identity(Index r, Index c)
    {
        Matrix m;
        m.resize(r, c);
        m.setIdentity();
        return m;
    }
    static Matrix Unit(Index i)
    {
        Matrix m;
        m[i] = T(1);
        return m;
    }
    static Matrix UnitX() { return Unit(0); }
    static Matrix UnitY() { return Unit(1); }
    static Matrix UnitZ() { return Unit(2); }

private:
    Matrix& randomize()
    {
        static thread_local unsigned long long state = 0x9E3779B97F4A7C15ull;
        for (Index i = 0; i < this->size(); ++i)
        {
            state = state * 6364136223846793005ull + 1442695040888963407ull;
            s_.data()[i] = T((double)(state >> 11) / 9007199254740992.0 * 2.0 - 1.0);
        }
        return *this;
    }
    void resize_vec(Index n)
    {
        if constexpr (fixed)
            (void)n;  // Zero(k) / Constant(k, v) on a fixed-size vector: the size is already known
        else
            resize(n);
    }
    template <typename D>
    void resize_like(const MatrixBase<D>& o)
    {
        if constexpr (fixed)
            return;
        else if constexpr (R == Dynamic && C == Dynamic)
            resize(o.rows(), o.cols());
        else if constexpr (C == 1)
            resize(o.size(), 1);  // vectors take vectors of either orientation
        else if constexpr (R == 1)
            resize(1, o.size());
        else if constexpr (R == Dynamic)
            resize(o.rows(), C);
        else
            resize(R, o.cols());
    }
};

// --------------------------------------------------------------------------------------------------------------------
// Map and Block (views)
// --------------------------------------------------------------------------------------------------------------------
template <typename M>
class Map : public MatrixBase<Map<M>>
{
    using Plain = std::remove_const_t<M>;
    using T = typename Plain::Scalar;
    using Ptr = std::conditional_t<std::is_const<M>::value, const T*, T*>;
    Ptr p_;
    Index r_, c_;

public:
    using Scalar = T;
    Map(Ptr p) : p_(p), r_(Plain::RowsAtCompileTime), c_(Plain::ColsAtCompileTime) {}
    Map(Ptr p, Index n) : p_(p), r_(Plain::RowsAtCompileTime == 1 ? 1 : n), c_(Plain::RowsAtCompileTime == 1 ? n : 1) {}
    Map(Ptr p, Index r, Index c) : p_(p), r_(r), c_(c) {}
    Map(const Map&) = default;
    Index rows_impl() const { return r_; }
    Index cols_impl() const { return c_; }
    const T& coeff(Index i, Index j) const { return p_[i + j * r_]; }
    decltype(auto) coeffRef(Index i, Index j) { return p_[i + j * r_]; }
    using MatrixBase<Map>::coeff;
    Ptr data() const { return p_; }
    template <typename D>
    Map& operator=(const MatrixBase<D>& o)
    {
        this->assign_from(o);
        return *this;
    }
    Map& operator=(const Map& o)
    {
        this->assign_from(o);
        return *this;
    }
};

template <typename X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC>>
{
    using T = typename internal::traits<std::remove_const_t<X>>::Scalar;
    X& x_;
    Index r0_, c0_, nr_, nc_;

public:
    using Scalar = T;
    Block(X& x, Index r0, Index c0, Index nr, Index nc) : x_(x), r0_(r0), c0_(c0), nr_(nr), nc_(nc)
    {
        if (r0 < 0 || c0 < 0 || nr < 0 || nc < 0 || r0 + nr > x.rows() || c0 + nc > x.cols()) throw std::out_of_range("eigen shim: block out of range");
    }
    Block(const Block&) = default;
    Index rows_impl() const { return nr_; }
    Index cols_impl() const { return nc_; }
    decltype(auto) coeff(Index i, Index j) const { return static_cast<const std::remove_const_t<X>&>(x_).coeff(r0_ + i, c0_ + j); }
    decltype(auto) coeffRef(Index i, Index j) { return x_.coeffRef(r0_ + i, c0_ + j); }
    using MatrixBase<Block>::coeff;
    template <typename D>
    Block& operator=(const MatrixBase<D>& o)
    {
        typename MatrixBase<D>::PlainObject tmp(o);  // the source may alias the target
        this->assign_from(tmp);
        return *this;
    }
    Block& operator=(const Block& o)
    {
        typename MatrixBase<Block>::PlainObject tmp(o);
        this->assign_from(tmp);
        return *this;
    }
};

// --------------------------------------------------------------------------------------------------------------------
// Diagonal wrapper
// --------------------------------------------------------------------------------------------------------------------
template <typename V>
class DiagonalWrapper : public EigenTag
{
    V v_;

public:
    using Scalar = typename V::Scalar;
    explicit DiagonalWrapper(const V& v) : v_(v) {}
    const V& diagonal() const { return v_; }
    Index rows() const { return v_.size(); }
    Index cols() const { return v_.size(); }
};

template <typename Derived>
DiagonalWrapper<typename MatrixBase<Derived>::PlainObject> MatrixBase<Derived>::asDiagonal() const
{
    return DiagonalWrapper<PlainObject>(PlainObject(*this));
}

template <typename T, int R, int C, int O, int MR, int MC>
template <typename V>
Matrix<T, R, C, O, MR, MC>::Matrix(const DiagonalWrapper<V>& d)
{
    const Index n = d.rows();
    if constexpr (!fixed) resize(n, n);
    if (this->rows() != n || this->cols() != n) throw std::logic_error("eigen shim: diagonal of a different size");
    this->setZero();
    for (Index i = 0; i < n; ++i) coeffRef(i, i) = T(d.diagonal()[i]);
}

// --------------------------------------------------------------------------------------------------------------------
// Comma initialiser (row-major filling by scalars or blocks)
// --------------------------------------------------------------------------------------------------------------------
template <typename Derived>
class CommaInitializer
{
    Derived& m_;
    Index row_ = 0, col_ = 0, block_rows_ = 1;

    template <typename A>
    void put(const A& a)
    {
        if constexpr (is_eigen<A>::value)
        {
            if (col_ == m_.cols())
            {
                row_ += block_rows_;
                col_ = 0;
                block_rows_ = a.rows();
            }
            if (row_ + a.rows() > m_.rows() || col_ + a.cols() > m_.cols()) throw std::out_of_range("eigen shim: too many coefficients in comma initialiser");
            for (Index j = 0; j < a.cols(); ++j)
                for (Index i = 0; i < a.rows(); ++i) m_(row_ + i, col_ + j) = typename Derived::Scalar(a(i, j));
            col_ += a.cols();
        }
        else
        {
            if (col_ == m_.cols())
            {
                row_ += block_rows_;
                col_ = 0;
                block_rows_ = 1;
            }
            if (row_ >= m_.rows()) throw std::out_of_range("eigen shim: too many coefficients in comma initialiser");
            m_(row_, col_) = typename Derived::Scalar(a);
            ++col_;
        }
    }

public:
    template <typename A>
    CommaInitializer(Derived& m, const A& a) : m_(m)
    {
        if constexpr (is_eigen<A>::value) block_rows_ = a.rows();
        put(a);
    }
    template <typename A>
    CommaInitializer& operator,(const A& a)
    {
        put(a);
        return *this;
    }
    Derived& finished() { return m_; }
};

template <typename Derived>
template <typename A>
CommaInitializer<Derived> MatrixBase<Derived>::operator<<(const A& a)
{
    return CommaInitializer<Derived>(derived(), a);
}

// --------------------------------------------------------------------------------------------------------------------
// Operators
// --------------------------------------------------------------------------------------------------------------------
template <typename A, typename B>
auto operator+(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    using RT = decltype(std::declval<typename A::Scalar>() + std::declval<typename B::Scalar>());
    Matrix<RT, internal::pick_dim(A::RowsAtCompileTime, B::RowsAtCompileTime), internal::pick_dim(A::ColsAtCompileTime, B::ColsAtCompileTime)> r;
    if (a.rows() != b.rows() || a.cols() != b.cols()) throw std::logic_error("eigen shim: operands of different shapes");
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r(i, j) = a(i, j) + b(i, j);
    return r;
}

template <typename A, typename B>
auto operator-(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    using RT = decltype(std::declval<typename A::Scalar>() - std::declval<typename B::Scalar>());
    Matrix<RT, internal::pick_dim(A::RowsAtCompileTime, B::RowsAtCompileTime), internal::pick_dim(A::ColsAtCompileTime, B::ColsAtCompileTime)> r;
    if (a.rows() != b.rows() || a.cols() != b.cols()) throw std::logic_error("eigen shim: operands of different shapes");
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r(i, j) = a(i, j) - b(i, j);
    return r;
}

template <typename A>
auto operator-(const MatrixBase<A>& a)
{
    typename MatrixBase<A>::PlainObject r;
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r(i, j) = -a(i, j);
    return r;
}

template <typename A, typename B>
auto operator*(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    using RT = decltype(std::declval<typename A::Scalar>() * std::declval<typename B::Scalar>());
    Matrix<RT, A::RowsAtCompileTime, B::ColsAtCompileTime> r;
    if (a.cols() != b.rows()) throw std::logic_error("eigen shim: product of incompatible shapes");
    r.resize(a.rows(), b.cols());
    const Index n = a.cols();
    for (Index j = 0; j < b.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i)
        {
            if (n == 0)
            {
                r(i, j) = RT(0);
                continue;
            }
            RT s = a(i, 0) * b(0, j);
            for (Index k = 1; k < n; ++k) s = s + a(i, k) * b(k, j);
            r(i, j) = s;
        }
    return r;
}

template <typename A, typename S, typename = std::enable_if_t<!is_eigen<S>::value>>
auto operator*(const MatrixBase<A>& a, const S& s)
{
    using RT = decltype(std::declval<typename A::Scalar>() * std::declval<S>());
    Matrix<RT, A::RowsAtCompileTime, A::ColsAtCompileTime> r;
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r(i, j) = a(i, j) * s;
    return r;
}

template <typename S, typename A, typename = std::enable_if_t<!is_eigen<S>::value>>
auto operator*(const S& s, const MatrixBase<A>& a)
{
    using RT = decltype(std::declval<S>() * std::declval<typename A::Scalar>());
    Matrix<RT, A::RowsAtCompileTime, A::ColsAtCompileTime> r;
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r(i, j) = s * a(i, j);
    return r;
}

template <typename A, typename S, typename = std::enable_if_t<!is_eigen<S>::value>>
auto operator/(const MatrixBase<A>& a, const S& s)
{
    using RT = decltype(std::declval<typename A::Scalar>() / std::declval<S>());
    Matrix<RT, A::RowsAtCompileTime, A::ColsAtCompileTime> r;
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r(i, j) = a(i, j) / s;
    return r;
}

template <typename A, typename V>
auto operator*(const MatrixBase<A>& a, const DiagonalWrapper<V>& d)
{
    using RT = decltype(std::declval<typename A::Scalar>() * std::declval<typename V::Scalar>());
    Matrix<RT, A::RowsAtCompileTime, A::ColsAtCompileTime> r;
    if (a.cols() != d.rows()) throw std::logic_error("eigen shim: product of incompatible shapes");
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r(i, j) = a(i, j) * d.diagonal()[j];
    return r;
}

template <typename V, typename A>
auto operator*(const DiagonalWrapper<V>& d, const MatrixBase<A>& a)
{
    using RT = decltype(std::declval<typename V::Scalar>() * std::declval<typename A::Scalar>());
    Matrix<RT, A::RowsAtCompileTime, A::ColsAtCompileTime> r;
    if (a.rows() != d.rows()) throw std::logic_error("eigen shim: product of incompatible shapes");
    r.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r(i, j) = d.diagonal()[i] * a(i, j);
    return r;
}

template <typename A, typename B>
bool operator==(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i)
            if (!(a(i, j) == b(i, j))) return false;
    return true;
}
template <typename A, typename B>
bool operator!=(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    return !(a == b);
}

template <typename A>
std::ostream& operator<<(std::ostream& s, const MatrixBase<A>& a)
{
    for (Index i = 0; i < a.rows(); ++i)
    {
        for (Index j = 0; j < a.cols(); ++j) s << (j ? " " : "") << a(i, j);
        if (i + 1 < a.rows()) s << "\n";
    }
    return s;
}

// --------------------------------------------------------------------------------------------------------------------
// Typedefs of "old" Eigen (the templated Vector<T, n> / Vector2<T> / MatrixX<T> aliases are supplied by the reference's
// Detail/EigenVectorTypedefs.hh and must not be defined here)
// --------------------------------------------------------------------------------------------------------------------
#define TINYAD_SHIM_TYPEDEFS(T, S)                     \
    typedef Matrix<T, 2, 2> Matrix2##S;                \
    typedef Matrix<T, 3, 3> Matrix3##S;                \
    typedef Matrix<T, 4, 4> Matrix4##S;                \
    typedef Matrix<T, Dynamic, Dynamic> MatrixX##S;    \
    typedef Matrix<T, 2, 1> Vector2##S;                \
    typedef Matrix<T, 3, 1> Vector3##S;                \
    typedef Matrix<T, 4, 1> Vector4##S;                \
    typedef Matrix<T, Dynamic, 1> VectorX##S;          \
    typedef Matrix<T, 1, 2> RowVector2##S;             \
    typedef Matrix<T, 1, 3> RowVector3##S;             \
    typedef Matrix<T, 1, 4> RowVector4##S;             \
    typedef Matrix<T, 1, Dynamic> RowVectorX##S;       \
    typedef Matrix<T, 2, Dynamic> Matrix2X##S;         \
    typedef Matrix<T, 3, Dynamic> Matrix3X##S;         \
    typedef Matrix<T, Dynamic, 2> MatrixX2##S;         \
    typedef Matrix<T, Dynamic, 3> MatrixX3##S;
TINYAD_SHIM_TYPEDEFS(double, d)
TINYAD_SHIM_TYPEDEFS(float, f)
TINYAD_SHIM_TYPEDEFS(int, i)
#undef TINYAD_SHIM_TYPEDEFS

// the alias templates of Eigen 3.4 (the reference's tests/Meshes.hh uses them after including only <Eigen/Core>);
// Detail/EigenVectorTypedefs.hh declares the same aliases again, which is a valid redeclaration
template <typename Type> using Matrix2 = Matrix<Type, 2, 2>;
template <typename Type> using Vector2 = Matrix<Type, 2, 1>;
template <typename Type> using RowVector2 = Matrix<Type, 1, 2>;
template <typename Type> using Matrix3 = Matrix<Type, 3, 3>;
template <typename Type> using Vector3 = Matrix<Type, 3, 1>;
template <typename Type> using RowVector3 = Matrix<Type, 1, 3>;
template <typename Type> using Matrix4 = Matrix<Type, 4, 4>;
template <typename Type> using Vector4 = Matrix<Type, 4, 1>;
template <typename Type> using RowVector4 = Matrix<Type, 1, 4>;
template <typename Type> using MatrixX = Matrix<Type, Dynamic, Dynamic>;
template <typename Type> using VectorX = Matrix<Type, Dynamic, 1>;
template <typename Type> using RowVectorX = Matrix<Type, 1, Dynamic>;
template <typename Type> using Matrix2X = Matrix<Type, 2, Dynamic>;
template <typename Type> using MatrixX2 = Matrix<Type, Dynamic, 2>;
template <typename Type> using Matrix3X = Matrix<Type, 3, Dynamic>;
template <typename Type> using MatrixX3 = Matrix<Type, Dynamic, 3>;
template <typename Type> using Matrix4X = Matrix<Type, 4, Dynamic>;
template <typename Type> using MatrixX4 = Matrix<Type, Dynamic, 4>;
template <typename Type, int Size> using Vector = Matrix<Type, Size, 1>;
template <typename Type, int Size> using RowVector = Matrix<Type, 1, Size>;

// --------------------------------------------------------------------------------------------------------------------
// SelfAdjointEigenSolver: Householder tridiagonalisation of the lower triangle, implicit-shift QL on the tridiagonal
// matrix with accumulation of the transformations; eigenvalues ascending, eigenvectors in the columns.
// --------------------------------------------------------------------------------------------------------------------
template <typename MatT>
class SelfAdjointEigenSolver
{
public:
    using Scalar = typename MatT::Scalar;
    using MatrixType = MatT;
    using RealVectorType = Matrix<Scalar, MatT::RowsAtCompileTime, 1>;
    using EigenvectorsType = MatT;

    SelfAdjointEigenSolver() = default;
    template <typename D>
    explicit SelfAdjointEigenSolver(const MatrixBase<D>& m, int /*options*/ = 0)
    {
        compute(m);
    }

    template <typename D>
    SelfAdjointEigenSolver& compute(const MatrixBase<D>& m, int /*options*/ = 0)
    {
        const Index n = m.rows();
        if (n != m.cols()) throw std::logic_error("eigen shim: SelfAdjointEigenSolver needs a square matrix");
        vec_.resize(n, n);
        val_.resize(n, 1);
        std::vector<Scalar> e((std::size_t)n, Scalar(0));
        for (Index j = 0; j < n; ++j)
            for (Index i = 0; i < n; ++i) vec_(i, j) = i >= j ? Scalar(m(i, j)) : Scalar(m(j, i));  // lower triangle only
        info_ = Success;
        if (n == 0) return *this;
        tridiagonalize(n, e);
        if (!ql(n, e)) info_ = NoConvergence;
        sort(n);
        return *this;
    }

    const RealVectorType& eigenvalues() const { return val_; }
    const EigenvectorsType& eigenvectors() const { return vec_; }
    ComputationInfo info() const { return info_; }

private:
    MatT vec_;
    RealVectorType val_;
    ComputationInfo info_ = InvalidInput;

    // reduce vec_ (symmetric) to tridiagonal form, leaving the accumulated orthogonal matrix in vec_,
    // the diagonal in val_ and the sub-diagonal in e[1..n-1]
    void tridiagonalize(Index n, std::vector<Scalar>& e)
    {
        using std::abs;
        using std::sqrt;
        MatT& a = vec_;
        RealVectorType& d = val_;
        for (Index i = n - 1; i > 0; --i)
        {
            const Index l = i - 1;
            Scalar h = 0, scale = 0;
            if (l > 0)
            {
                for (Index k = 0; k <= l; ++k) scale += abs(a(i, k));
                if (scale == Scalar(0))
                    e[i] = a(i, l);
                else
                {
                    for (Index k = 0; k <= l; ++k)
                    {
                        a(i, k) /= scale;
                        h += a(i, k) * a(i, k);
                    }
                    Scalar f = a(i, l);
                    Scalar g = f >= Scalar(0) ? -sqrt(h) : sqrt(h);
                    e[i] = scale * g;
                    h -= f * g;
                    a(i, l) = f - g;
                    f = 0;
                    for (Index j = 0; j <= l; ++j)
                    {
                        a(j, i) = a(i, j) / h;
                        g = 0;
                        for (Index k = 0; k <= j; ++k) g += a(j, k) * a(i, k);
                        for (Index k = j + 1; k <= l; ++k) g += a(k, j) * a(i, k);
                        e[j] = g / h;
                        f += e[j] * a(i, j);
                    }
                    const Scalar hh = f / (h + h);
                    for (Index j = 0; j <= l; ++j)
                    {
                        f = a(i, j);
                        e[j] = g = e[j] - hh * f;
                        for (Index k = 0; k <= j; ++k) a(j, k) -= f * e[k] + g * a(i, k);
                    }
                }
            }
            else
                e[i] = a(i, l);
            d[i] = h;
        }
        d[0] = 0;
        e[0] = 0;
        for (Index i = 0; i < n; ++i)
        {
            const Index l = i - 1;
            if (d[i] != Scalar(0))
            {
                for (Index j = 0; j <= l; ++j)
                {
                    Scalar g = 0;
                    for (Index k = 0; k <= l; ++k) g += a(i, k) * a(k, j);
                    for (Index k = 0; k <= l; ++k) a(k, j) -= g * a(k, i);
                }
            }
            d[i] = a(i, i);
            a(i, i) = 1;
            for (Index j = 0; j <= l; ++j) a(j, i) = a(i, j) = 0;
        }
    }

    bool ql(Index n, std::vector<Scalar>& e)
    {
        using std::abs;
        using std::sqrt;
        MatT& z = vec_;
        RealVectorType& d = val_;
        for (Index i = 1; i < n; ++i) e[i - 1] = e[i];
        e[n - 1] = 0;
        for (Index l = 0; l < n; ++l)
        {
            int iter = 0;
            Index m;
            do
            {
                for (m = l; m < n - 1; ++m)
                {
                    const Scalar dd = abs(d[m]) + abs(d[m + 1]);
                    if (abs(e[m]) <= std::numeric_limits<Scalar>::epsilon() * dd) break;
                }
                if (m != l)
                {
                    if (iter++ == 64) return false;
                    Scalar g = (d[l + 1] - d[l]) / (Scalar(2) * e[l]);
                    Scalar r = hypot_(g, Scalar(1));
                    g = d[m] - d[l] + e[l] / (g + (g >= Scalar(0) ? abs(r) : -abs(r)));
                    Scalar s = 1, c = 1, p = 0;
                    Index i;
                    for (i = m - 1; i >= l; --i)
                    {
                        Scalar f = s * e[i];
                        const Scalar b = c * e[i];
                        e[i + 1] = (r = hypot_(f, g));
                        if (r == Scalar(0))
                        {
                            d[i + 1] -= p;
                            e[m] = 0;
                            break;
                        }
                        s = f / r;
                        c = g / r;
                        g = d[i + 1] - p;
                        r = (d[i] - g) * s + Scalar(2) * c * b;
                        d[i + 1] = g + (p = s * r);
                        g = c * r - b;
                        for (Index k = 0; k < n; ++k)
                        {
                            f = z(k, i + 1);
                            z(k, i + 1) = s * z(k, i) + c * f;
                            z(k, i) = c * z(k, i) - s * f;
                        }
                    }
                    if (r == Scalar(0) && i >= l) continue;
                    d[l] -= p;
                    e[l] = g;
                    e[m] = 0;
                }
            } while (m != l);
        }
        return true;
    }

    static Scalar hypot_(Scalar a, Scalar b)
    {
        using std::abs;
        using std::sqrt;
        const Scalar aa = abs(a), ab = abs(b);
        if (aa > ab) return aa * sqrt(Scalar(1) + (ab / aa) * (ab / aa));
        return ab == Scalar(0) ? Scalar(0) : ab * sqrt(Scalar(1) + (aa / ab) * (aa / ab));
    }

    void sort(Index n)
    {
        for (Index i = 0; i < n - 1; ++i)
        {
            Index k = i;
            for (Index j = i + 1; j < n; ++j)
                if (val_[j] < val_[k]) k = j;
            if (k != i)
            {
                std::swap(val_[i], val_[k]);
                for (Index r = 0; r < n; ++r) std::swap(vec_(r, i), vec_(r, k));
            }
        }
    }
};

// --------------------------------------------------------------------------------------------------------------------
// JacobiSVD: one-sided Jacobi (Hestenes) on the columns, generic in the scalar type (works on TinyAD::Scalar as the
// reference's tests/SVDTest.cc asks); singular values descending, full U and V for square matrices.
// --------------------------------------------------------------------------------------------------------------------
enum { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

template <typename MatT, int QRPreconditioner = 0>
class JacobiSVD
{
public:
    using Scalar = typename MatT::Scalar;
    using SingularValuesType = Matrix<Scalar, MatT::ColsAtCompileTime, 1>;

    JacobiSVD() = default;
    template <typename D>
    explicit JacobiSVD(const MatrixBase<D>& m, unsigned int options = 0)
    {
        compute(m, options);
    }
    template <typename D>
    JacobiSVD& compute(const MatrixBase<D>& m, unsigned int = 0)
    {
        using std::abs;
        using std::sqrt;
        const Index r = m.rows(), c = m.cols();
        if (r != c) throw std::logic_error("eigen shim: JacobiSVD is implemented for square matrices");
        MatT a = m;  // columns converge to U * diag(S)
        v_.resize(c, c);
        v_.setIdentity();
        for (int sweep = 0; sweep < 60; ++sweep)
        {
            bool rotated = false;
            for (Index p = 0; p < c - 1; ++p)
                for (Index q = p + 1; q < c; ++q)
                {
                    Scalar alpha = Scalar(0), beta = Scalar(0), gamma = Scalar(0);
                    for (Index i = 0; i < r; ++i)
                    {
                        alpha = alpha + a(i, p) * a(i, p);
                        beta = beta + a(i, q) * a(i, q);
                        gamma = gamma + a(i, p) * a(i, q);
                    }
                    if (abs(gamma) <= 1e-16 * sqrt(alpha * beta) || gamma == Scalar(0)) continue;
                    rotated = true;
                    const Scalar zeta = (beta - alpha) / (Scalar(2) * gamma);
                    const Scalar t = (zeta >= Scalar(0) ? Scalar(1) : Scalar(-1)) / (abs(zeta) + sqrt(Scalar(1) + zeta * zeta));
                    const Scalar cs = Scalar(1) / sqrt(Scalar(1) + t * t), sn = cs * t;
                    for (Index i = 0; i < r; ++i)
                    {
                        const Scalar x = a(i, p), y = a(i, q);
                        a(i, p) = cs * x - sn * y;
                        a(i, q) = sn * x + cs * y;
                    }
                    for (Index i = 0; i < c; ++i)
                    {
                        const Scalar x = v_(i, p), y = v_(i, q);
                        v_(i, p) = cs * x - sn * y;
                        v_(i, q) = sn * x + cs * y;
                    }
                }
            if (!rotated) break;
        }
        s_.resize(c, 1);
        u_.resize(r, c);
        for (Index j = 0; j < c; ++j)
        {
            Scalar n2 = Scalar(0);
            for (Index i = 0; i < r; ++i) n2 = n2 + a(i, j) * a(i, j);
            s_[j] = sqrt(n2);
            for (Index i = 0; i < r; ++i) u_(i, j) = a(i, j) / s_[j];
        }
        for (Index i = 0; i < c - 1; ++i)  // descending
        {
            Index k = i;
            for (Index j = i + 1; j < c; ++j)
                if (s_[k] < s_[j]) k = j;
            if (k != i)
            {
                std::swap(s_[i], s_[k]);
                for (Index rr = 0; rr < r; ++rr) std::swap(u_(rr, i), u_(rr, k));
                for (Index rr = 0; rr < c; ++rr) std::swap(v_(rr, i), v_(rr, k));
            }
        }
        return *this;
    }
    const MatT& matrixU() const { return u_; }
    const MatT& matrixV() const { return v_; }
    const SingularValuesType& singularValues() const { return s_; }

private:
    MatT u_, v_;
    SingularValuesType s_;
};

// --------------------------------------------------------------------------------------------------------------------
// Sparse: Triplet and a compressed column-major SparseMatrix
// --------------------------------------------------------------------------------------------------------------------
template <typename T, typename StorageIndex_ = int>
class Triplet
{
    StorageIndex_ r_ = 0, c_ = 0;
    T v_ = T(0);

public:
    Triplet() = default;
    Triplet(const StorageIndex_& r, const StorageIndex_& c, const T& v = T(0)) : r_(r), c_(c), v_(v) {}
    const StorageIndex_& row() const { return r_; }
    const StorageIndex_& col() const { return c_; }
    const T& value() const { return v_; }
};

template <typename T, int Options = 0, typename StorageIndex_ = int>
class SparseMatrix : public EigenTag
{
public:
    using Scalar = T;
    using StorageIndex = StorageIndex_;

    SparseMatrix() : outer_(1, 0) {}
    SparseMatrix(Index r, Index c) : rows_(r), cols_(c), outer_((std::size_t)c + 1, 0) {}

    Index rows() const { return rows_; }
    Index cols() const { return cols_; }
    Index outerSize() const { return cols_; }
    Index innerSize() const { return rows_; }
    Index nonZeros() const { return (Index)inner_.size(); }
    bool isCompressed() const { return true; }
    void makeCompressed() {}
    void resize(Index r, Index c)
    {
        rows_ = r;
        cols_ = c;
        outer_.assign((std::size_t)c + 1, 0);
        inner_.clear();
        values_.clear();
    }
    void setZero() { resize(rows_, cols_); }
    void setIdentity()
    {
        const Index n = (std::min)(rows_, cols_);
        inner_.resize((std::size_t)n);
        values_.assign((std::size_t)n, T(1));
        for (Index j = 0; j < n; ++j) inner_[j] = (StorageIndex)j;
        for (Index j = 0; j <= cols_; ++j) outer_[j] = (StorageIndex)(std::min)(j, n);
    }
    const StorageIndex* outerIndexPtr() const { return outer_.data(); }
    const StorageIndex* innerIndexPtr() const { return inner_.data(); }
    const T* valuePtr() const { return values_.data(); }
    StorageIndex* outerIndexPtr() { return outer_.data(); }
    StorageIndex* innerIndexPtr() { return inner_.data(); }
    T* valuePtr() { return values_.data(); }

    // Duplicates are summed in the order of the input range (the first occurrence of an entry receives the later
    // ones one by one), columns are compressed with ascending row indices.
    template <typename It>
    void setFromTriplets(It begin, It end)
    {
        // pass 1: bucket by row, keeping input order (the row-major intermediate of the published algorithm)
        std::vector<std::size_t> row_start((std::size_t)rows_ + 1, 0);
        std::size_t n = 0;
        for (It it = begin; it != end; ++it, ++n)
        {
            if (it->row() < 0 || it->row() >= rows_ || it->col() < 0 || it->col() >= cols_) throw std::out_of_range("eigen shim: triplet out of range");
            ++row_start[(std::size_t)it->row() + 1];
        }
        for (Index i = 0; i < rows_; ++i) row_start[i + 1] += row_start[i];
        std::vector<StorageIndex> tcol(n);
        std::vector<T> tval(n);
        {
            std::vector<std::size_t> pos(row_start.begin(), row_start.end() - 1);
            for (It it = begin; it != end; ++it)
            {
                const std::size_t p = pos[(std::size_t)it->row()]++;
                tcol[p] = (StorageIndex)it->col();
                tval[p] = it->value();
            }
        }
        // pass 2: collapse duplicates within each row, in order
        std::vector<std::ptrdiff_t> where((std::size_t)cols_, -1);
        std::vector<std::size_t> new_start((std::size_t)rows_ + 1, 0);
        std::size_t count = 0;
        for (Index i = 0; i < rows_; ++i)
        {
            const std::size_t first = count;
            new_start[i] = first;
            for (std::size_t k = row_start[i]; k < row_start[i + 1]; ++k)
            {
                const StorageIndex c = tcol[k];
                if (where[c] >= (std::ptrdiff_t)first)
                    tval[(std::size_t)where[c]] += tval[k];
                else
                {
                    tcol[count] = c;
                    tval[count] = tval[k];
                    where[c] = (std::ptrdiff_t)count;
                    ++count;
                }
            }
        }
        new_start[rows_] = count;
        // pass 3: transpose into column-major (rows visited ascending -> sorted inner indices)
        outer_.assign((std::size_t)cols_ + 1, 0);
        for (std::size_t k = 0; k < count; ++k) ++outer_[(std::size_t)tcol[k] + 1];
        for (Index j = 0; j < cols_; ++j) outer_[j + 1] += outer_[j];
        inner_.resize(count);
        values_.resize(count);
        std::vector<StorageIndex> pos(outer_.begin(), outer_.end() - 1);
        for (Index i = 0; i < rows_; ++i)
            for (std::size_t k = new_start[i]; k < new_start[i + 1]; ++k)
            {
                const StorageIndex p = pos[tcol[k]]++;
                inner_[p] = (StorageIndex)i;
                values_[p] = tval[k];
            }
    }

    T coeff(Index i, Index j) const
    {
        for (StorageIndex p = outer_[j]; p < outer_[j + 1]; ++p)
            if (inner_[p] == i) return values_[p];
        return T(0);
    }
    // reference to entry (i, j), inserted as an explicit zero when absent (keeps the columns sorted)
    T& coeffRef(Index i, Index j)
    {
        StorageIndex p = outer_[j];
        while (p < outer_[j + 1] && inner_[p] < i) ++p;
        if (p < outer_[j + 1] && inner_[p] == i) return values_[p];
        inner_.insert(inner_.begin() + p, (StorageIndex)i);
        values_.insert(values_.begin() + p, T(0));
        for (Index c = j + 1; c <= cols_; ++c) ++outer_[c];
        return values_[p];
    }
    T& insert(Index i, Index j) { return coeffRef(i, j); }
    void reserve(Index) {}
    template <typename X>
    void reserve(const X&) {}

    // a*A + b*B on the union pattern
    static SparseMatrix combine(const T& a, const SparseMatrix& A, const T& b, const SparseMatrix& B)
    {
        if (A.rows_ != B.rows_ || A.cols_ != B.cols_) throw std::logic_error("eigen shim: sparse operands of different shapes");
        SparseMatrix r(A.rows_, A.cols_);
        for (Index j = 0; j < A.cols_; ++j)
        {
            StorageIndex p = A.outer_[j], q = B.outer_[j];
            const StorageIndex pe = A.outer_[j + 1], qe = B.outer_[j + 1];
            while (p < pe || q < qe)
            {
                if (q >= qe || (p < pe && A.inner_[p] < B.inner_[q]))
                {
                    r.inner_.push_back(A.inner_[p]);
                    r.values_.push_back(a * A.values_[p]);
                    ++p;
                }
                else if (p >= pe || B.inner_[q] < A.inner_[p])
                {
                    r.inner_.push_back(B.inner_[q]);
                    r.values_.push_back(b * B.values_[q]);
                    ++q;
                }
                else
                {
                    r.inner_.push_back(A.inner_[p]);
                    r.values_.push_back(a * A.values_[p] + b * B.values_[q]);
                    ++p;
                    ++q;
                }
            }
            r.outer_[j + 1] = (StorageIndex)r.inner_.size();
        }
        return r;
    }
    friend SparseMatrix operator+(const SparseMatrix& A, const SparseMatrix& B) { return combine(T(1), A, T(1), B); }
    friend SparseMatrix operator-(const SparseMatrix& A, const SparseMatrix& B) { return combine(T(1), A, T(-1), B); }
    SparseMatrix operator-() const
    {
        SparseMatrix r = *this;
        for (auto& v : r.values_) v = -v;
        return r;
    }
    friend SparseMatrix operator*(const SparseMatrix& A, const SparseMatrix& B)
    {
        if (A.cols_ != B.rows_) throw std::logic_error("eigen shim: sparse product of incompatible shapes");
        SparseMatrix r(A.rows_, B.cols_);
        std::vector<T> acc((std::size_t)A.rows_, T(0));
        std::vector<char> used((std::size_t)A.rows_, 0);
        std::vector<StorageIndex> idx;
        for (Index j = 0; j < B.cols_; ++j)
        {
            idx.clear();
            for (StorageIndex q = B.outer_[j]; q < B.outer_[j + 1]; ++q)
            {
                const StorageIndex k = B.inner_[q];
                for (StorageIndex p = A.outer_[k]; p < A.outer_[k + 1]; ++p)
                {
                    const StorageIndex i = A.inner_[p];
                    if (!used[i])
                    {
                        used[i] = 1;
                        idx.push_back(i);
                    }
                    acc[i] += A.values_[p] * B.values_[q];
                }
            }
            std::sort(idx.begin(), idx.end());
            for (StorageIndex i : idx)
            {
                r.inner_.push_back(i);
                r.values_.push_back(acc[i]);
                acc[i] = T(0);
                used[i] = 0;
            }
            r.outer_[j + 1] = (StorageIndex)r.inner_.size();
        }
        return r;
    }
    T squaredNorm() const
    {
        T s = T(0);
        for (const auto& v : values_) s += v * v;
        return s;
    }
    T norm() const
    {
        using std::sqrt;
        return sqrt(squaredNorm());
    }
    T sum() const
    {
        T s = T(0);
        for (const auto& v : values_) s += v;
        return s;
    }
    template <typename U>
    SparseMatrix<U, Options, StorageIndex_> cast() const
    {
        SparseMatrix<U, Options, StorageIndex_> r(rows_, cols_);
        std::vector<Triplet<U, StorageIndex_>> t;
        for (Index j = 0; j < cols_; ++j)
            for (StorageIndex p = outer_[j]; p < outer_[j + 1]; ++p) t.emplace_back(inner_[p], (StorageIndex_)j, (U)values_[p]);
        r.setFromTriplets(t.begin(), t.end());
        return r;
    }

    SparseMatrix transpose() const
    {
        SparseMatrix t(cols_, rows_);
        for (std::size_t k = 0; k < inner_.size(); ++k) ++t.outer_[(std::size_t)inner_[k] + 1];
        for (Index j = 0; j < rows_; ++j) t.outer_[j + 1] += t.outer_[j];
        t.inner_.resize(inner_.size());
        t.values_.resize(values_.size());
        std::vector<StorageIndex> pos(t.outer_.begin(), t.outer_.end() - 1);
        for (Index j = 0; j < cols_; ++j)
            for (StorageIndex p = outer_[j]; p < outer_[j + 1]; ++p)
            {
                const StorageIndex q = pos[inner_[p]]++;
                t.inner_[q] = (StorageIndex)j;
                t.values_[q] = values_[p];
            }
        return t;
    }

    class InnerIterator
    {
        const SparseMatrix& m_;
        Index j_;
        StorageIndex p_, end_;

    public:
        InnerIterator(const SparseMatrix& m, Index j) : m_(m), j_(j), p_(m.outer_[j]), end_(m.outer_[j + 1]) {}
        InnerIterator& operator++()
        {
            ++p_;
            return *this;
        }
        explicit operator bool() const { return p_ < end_; }
        Index row() const { return m_.inner_[p_]; }
        Index col() const { return j_; }
        Index index() const { return m_.inner_[p_]; }
        const T& value() const { return m_.values_[p_]; }
    };

    template <typename D>
    Matrix<T, Dynamic, 1> operator*(const MatrixBase<D>& v) const
    {
        if (v.size() != cols_) throw std::logic_error("eigen shim: sparse product of incompatible shapes");
        Matrix<T, Dynamic, 1> r = Matrix<T, Dynamic, 1>::Zero(rows_);
        for (Index j = 0; j < cols_; ++j)
            for (StorageIndex p = outer_[j]; p < outer_[j + 1]; ++p) r[inner_[p]] += values_[p] * v[j];
        return r;
    }
    SparseMatrix& operator*=(const T& s)
    {
        for (auto& v : values_) v *= s;
        return *this;
    }
    friend SparseMatrix operator*(const T& s, const SparseMatrix& m)
    {
        SparseMatrix r = m;
        r *= s;
        return r;
    }
    friend SparseMatrix operator*(const SparseMatrix& m, const T& s)
    {
        SparseMatrix r = m;
        r *= s;
        return r;
    }
    Matrix<T, Dynamic, Dynamic> toDense() const
    {
        Matrix<T, Dynamic, Dynamic> d = Matrix<T, Dynamic, Dynamic>::Zero(rows_, cols_);
        for (Index j = 0; j < cols_; ++j)
            for (StorageIndex p = outer_[j]; p < outer_[j + 1]; ++p) d(inner_[p], j) = values_[p];
        return d;
    }

private:
    Index rows_ = 0, cols_ = 0;
    std::vector<StorageIndex> outer_, inner_;
    std::vector<T> values_;
};

// --------------------------------------------------------------------------------------------------------------------
// Linear solvers of Eigen/SparseCholesky, SparseLU, SparseQR: a dense L D L^T for SimplicialLDLT (further down) and ONE dense LU
// with partial pivoting under the other names (only the reference's small test systems are ever solved with them; the solver is
// outside the path under test)
// --------------------------------------------------------------------------------------------------------------------
template <typename MatT>
class DenseFallbackSolver
{
public:
    using Scalar = typename MatT::Scalar;
    DenseFallbackSolver() = default;
    explicit DenseFallbackSolver(const MatT& A) { compute(A); }
    void analyzePattern(const MatT&) {}
    void factorize(const MatT& A)
    {
        using std::abs;
        n_ = A.rows();
        info_ = Success;
        if (A.rows() != A.cols())
        {
            info_ = InvalidInput;
            return;
        }
        lu_ = A.toDense();
        perm_.resize((std::size_t)n_);
        for (Index i = 0; i < n_; ++i) perm_[i] = i;
        for (Index c = 0; c < n_; ++c)
        {
            Index p = c;
            for (Index i = c + 1; i < n_; ++i)
                if (abs(lu_(i, c)) > abs(lu_(p, c))) p = i;
            if (!(abs(lu_(p, c)) > Scalar(0)) || !std::isfinite((double)lu_(p, c)))
            {
                info_ = NumericalIssue;
                return;
            }
            if (p != c)
            {
                std::swap(perm_[p], perm_[c]);
                for (Index j = 0; j < n_; ++j) std::swap(lu_(p, j), lu_(c, j));
            }
            for (Index i = c + 1; i < n_; ++i)
            {
                lu_(i, c) /= lu_(c, c);
                for (Index j = c + 1; j < n_; ++j) lu_(i, j) -= lu_(i, c) * lu_(c, j);
            }
        }
    }
    DenseFallbackSolver& compute(const MatT& A)
    {
        analyzePattern(A);
        factorize(A);
        return *this;
    }
    template <typename D>
    Matrix<Scalar, Dynamic, 1> solve(const MatrixBase<D>& b) const
    {
        Matrix<Scalar, Dynamic, 1> x(n_);
        if (info_ != Success || b.size() != n_)
        {
            x.setConstant(std::numeric_limits<Scalar>::quiet_NaN());
            return x;
        }
        for (Index i = 0; i < n_; ++i)
        {
            Scalar s = b[perm_[i]];
            for (Index j = 0; j < i; ++j) s -= lu_(i, j) * x[j];
            x[i] = s;
        }
        for (Index i = n_ - 1; i >= 0; --i)
        {
            Scalar s = x[i];
            for (Index j = i + 1; j < n_; ++j) s -= lu_(i, j) * x[j];
            x[i] = s / lu_(i, i);
        }
        return x;
    }
    ComputationInfo info() const { return info_; }

private:
    Matrix<Scalar, Dynamic, Dynamic> lu_;
    std::vector<Index> perm_;
    Index n_ = 0;
    ComputationInfo info_ = InvalidInput;
};

template <typename I>
struct COLAMDOrdering
{
};
template <typename I>
struct AMDOrdering
{
};
template <typename I>
struct NaturalOrdering
{
};
// symmetric L D L^T without pivoting on the lower triangle (what a simplicial factorisation computes, here densely)
template <typename MatT, int UpLo = Lower, typename Ordering = AMDOrdering<int>>
class SimplicialLDLT
{
public:
    using Scalar = typename MatT::Scalar;
    SimplicialLDLT() = default;
    explicit SimplicialLDLT(const MatT& A) { compute(A); }
    void analyzePattern(const MatT&) {}
    void factorize(const MatT& A)
    {
        n_ = A.rows();
        info_ = Success;
        if (A.rows() != A.cols())
        {
            info_ = InvalidInput;
            return;
        }
        l_ = A.toDense();
        d_.resize((std::size_t)n_);
        for (Index j = 0; j < n_; ++j)
        {
            Scalar dj = l_(j, j);
            for (Index k = 0; k < j; ++k) dj -= l_(j, k) * l_(j, k) * d_[k];
            d_[j] = dj;
            if (dj == Scalar(0) || !std::isfinite((double)dj))
            {
                info_ = NumericalIssue;
                return;
            }
            for (Index i = j + 1; i < n_; ++i)
            {
                Scalar v = l_(i, j);
                for (Index k = 0; k < j; ++k) v -= l_(i, k) * l_(j, k) * d_[k];
                l_(i, j) = v / dj;
            }
        }
    }
    SimplicialLDLT& compute(const MatT& A)
    {
        factorize(A);
        return *this;
    }
    template <typename D>
    Matrix<Scalar, Dynamic, 1> solve(const MatrixBase<D>& b) const
    {
        Matrix<Scalar, Dynamic, 1> x(n_);
        if (info_ != Success || b.size() != n_)
        {
            x.setConstant(std::numeric_limits<Scalar>::quiet_NaN());
            return x;
        }
        for (Index i = 0; i < n_; ++i)
        {
            Scalar s = b[i];
            for (Index k = 0; k < i; ++k) s -= l_(i, k) * x[k];
            x[i] = s;
        }
        for (Index i = 0; i < n_; ++i) x[i] /= d_[i];
        for (Index i = n_ - 1; i >= 0; --i)
        {
            Scalar s = x[i];
            for (Index k = i + 1; k < n_; ++k) s -= l_(k, i) * x[k];
            x[i] = s;
        }
        return x;
    }
    ComputationInfo info() const { return info_; }

private:
    Matrix<Scalar, Dynamic, Dynamic> l_;
    std::vector<Scalar> d_;
    Index n_ = 0;
    ComputationInfo info_ = InvalidInput;
};
template <typename MatT, int UpLo = Lower, typename Ordering = AMDOrdering<int>>
class SimplicialLLT : public DenseFallbackSolver<MatT>
{
public:
    using DenseFallbackSolver<MatT>::DenseFallbackSolver;
};
template <typename MatT, typename Ordering = COLAMDOrdering<int>>
class SparseLU : public DenseFallbackSolver<MatT>
{
public:
    using DenseFallbackSolver<MatT>::DenseFallbackSolver;
};
template <typename MatT, typename Ordering>
class SparseQR : public DenseFallbackSolver<MatT>
{
public:
    using DenseFallbackSolver<MatT>::DenseFallbackSolver;
};

}  // namespace Eigen
