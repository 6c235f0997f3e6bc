// tinyad_b200 -- tiny fixed-size matrices usable inside __host__ __device__ element functors.
//
// The reference's element lambdas use Eigen fixed-size algebra on active scalars
// (e.g. tests/NewtonTest.cc:28-44: col_mat, determinant, inverse, squaredNorm, products);
// Eigen is neither installed in this image nor device-friendly for custom scalars, so the
// subset of that API the path needs is provided here with Eigen's names and semantics
// (column-major storage, cofactor inverse for 2x2 / 3x3).
#pragma once

#include <TinyAD/Scalar.hh>

namespace TinyAD
{

template <typename T, int R>
struct Inverse;

template <typename T, int R, int C>
struct Mat
{
    static constexpr int RowsAtCompileTime = R;
    static constexpr int ColsAtCompileTime = C;
    using Scalar = T;

    T a[R * C];

    TINYAD_HD TINYAD_INLINE Mat() { detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { a[decltype(ic)::value] = T(0.0); }); }
    // Vec<T,2>(x, y), Vec<T,3>(x, y, z)
    TINYAD_HD TINYAD_INLINE Mat(const T& x, const T& y) { static_assert(R * C == 2, "size"); a[0] = x; a[1] = y; }
    TINYAD_HD TINYAD_INLINE Mat(const T& x, const T& y, const T& z) { static_assert(R * C == 3, "size"); a[0] = x; a[1] = y; a[2] = z; }
    // converting copy (e.g. Mat<double> -> Mat<Scalar>)
    template <typename U, typename = std::enable_if_t<!std::is_same<U, T>::value && std::is_convertible<U, T>::value>>
    TINYAD_HD TINYAD_INLINE Mat(const Mat<U, R, C>& o) { detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { a[decltype(ic)::value] = T(o.a[decltype(ic)::value]); }); }

    TINYAD_HD TINYAD_INLINE T& operator()(int i, int j) { return a[j * R + i]; }
    TINYAD_HD TINYAD_INLINE const T& operator()(int i, int j) const { return a[j * R + i]; }
    TINYAD_HD TINYAD_INLINE T& operator()(int i) { return a[i]; }
    TINYAD_HD TINYAD_INLINE const T& operator()(int i) const { return a[i]; }
    TINYAD_HD TINYAD_INLINE T& operator[](int i) { return a[i]; }
    TINYAD_HD TINYAD_INLINE const T& operator[](int i) const { return a[i]; }
    TINYAD_HD TINYAD_INLINE T& x() { return a[0]; }
    TINYAD_HD TINYAD_INLINE T& y() { return a[1]; }
    TINYAD_HD TINYAD_INLINE T& z() { return a[2]; }
    TINYAD_HD TINYAD_INLINE const T& x() const { return a[0]; }
    TINYAD_HD TINYAD_INLINE const T& y() const { return a[1]; }
    TINYAD_HD TINYAD_INLINE const T& z() const { return a[2]; }
    TINYAD_HD static constexpr int rows() { return R; }
    TINYAD_HD static constexpr int cols() { return C; }
    TINYAD_HD static constexpr int size() { return R * C; }

    TINYAD_HD TINYAD_INLINE static Mat Constant(const T& v) { Mat m; detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { m.a[decltype(ic)::value] = v; }); return m; }
    TINYAD_HD TINYAD_INLINE static Mat Zero() { return Mat(); }
    TINYAD_HD TINYAD_INLINE static Mat Identity()
    {
        Mat m;
        detail::static_for<(R < C ? R : C)>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; m.a[i * R + i] = T(1.0); });
        return m;
    }

    TINYAD_HD TINYAD_INLINE T squaredNorm() const
    {
        T s = a[0] * a[0];
        detail::static_for<R * C - 1>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value + 1; s = s + a[i] * a[i]; });
        return s;
    }
    TINYAD_HD TINYAD_INLINE T norm() const { return sqrt(squaredNorm()); }
    TINYAD_HD TINYAD_INLINE T sum() const
    {
        T s = a[0];
        detail::static_for<R * C - 1>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value + 1; s = s + a[i]; });
        return s;
    }
    TINYAD_HD TINYAD_INLINE T trace() const
    {
        static_assert(R == C, "square");
        T s = a[0];
        detail::static_for<R - 1>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value + 1; s = s + a[i * R + i]; });
        return s;
    }
    template <typename U>
    TINYAD_HD TINYAD_INLINE auto dot(const Mat<U, R, C>& o) const
    {
        auto s = a[0] * o.a[0];
        detail::static_for<R * C - 1>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value + 1; s = s + a[i] * o.a[i]; });
        return s;
    }
    template <typename U>
    TINYAD_HD TINYAD_INLINE auto cwiseProduct(const Mat<U, R, C>& o) const
    {
        Mat<decltype(a[0] * o.a[0]), R, C> r;
        detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = a[i] * o.a[i]; });
        return r;
    }
    template <typename U>
    TINYAD_HD TINYAD_INLINE auto cross(const Mat<U, R, C>& o) const
    {
        static_assert(R * C == 3, "cross needs 3-vectors");
        Mat<decltype(a[0] * o.a[0]), R, C> r;
        r.a[0] = a[1] * o.a[2] - a[2] * o.a[1];
        r.a[1] = a[2] * o.a[0] - a[0] * o.a[2];
        r.a[2] = a[0] * o.a[1] - a[1] * o.a[0];
        return r;
    }
    TINYAD_HD TINYAD_INLINE Mat<T, C, R> transpose() const
    {
        Mat<T, C, R> t;
        detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int l = decltype(ic)::value; constexpr int i = l % R, j = l / R; t.a[i * C + j] = a[l]; });
        return t;
    }
    TINYAD_HD TINYAD_INLINE Mat<T, R, 1> col(int j) const { Mat<T, R, 1> v; for (int i = 0; i < R; ++i) v.a[i] = a[j * R + i]; return v; }
    TINYAD_HD TINYAD_INLINE Mat<T, 1, C> row(int i) const { Mat<T, 1, C> v; for (int j = 0; j < C; ++j) v.a[j] = a[j * R + i]; return v; }

    // Eigen fixed-size determinant (2x2, 3x3 "bruteforce" form)
    TINYAD_HD TINYAD_INLINE T determinant() const
    {
        static_assert(R == C && (R == 1 || R == 2 || R == 3), "determinant: 1x1, 2x2 or 3x3");
        const Mat& m = *this;
        if constexpr (R == 1) return a[0];
        else if constexpr (R == 2) return m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1);
        else
            return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1))
                 - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0))
                 + m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
    }
    // Eigen fixed-size inverse (adjugate times 1/det), returned as a lazy expression like Eigen does:
    // `Mat<T,3,3> Ji = J.inverse();` materialises it, `J.inverse().squaredNorm()` streams entry by entry
    // so that only a handful of active scalars are live at a time (register pressure, Double<12>).
    TINYAD_HD TINYAD_INLINE Inverse<T, R> inverse() const
    {
        static_assert(R == C && (R == 1 || R == 2 || R == 3), "inverse: 1x1, 2x2 or 3x3");
        return Inverse<T, R>(*this);
    }
};

template <typename T, int N>
using Vec = Mat<T, N, 1>;
template <typename T> using Vector2 = Mat<T, 2, 1>;
template <typename T> using Vector3 = Mat<T, 3, 1>;
template <typename T> using Matrix2 = Mat<T, 2, 2>;
template <typename T> using Matrix3 = Mat<T, 3, 3>;

template <typename T, int R>
struct Inverse
{
    Mat<T, R, R> m;
    T invdet;

    TINYAD_HD TINYAD_INLINE explicit Inverse(const Mat<T, R, R>& _m) : m(_m)
    {
        if constexpr (R == 1) invdet = 1.0 / m.a[0];
        else if constexpr (R == 2) invdet = 1.0 / m.determinant();
        else
        {
            // det from the first-column cofactors, like Eigen's compute_inverse_size3_helper
            const T det = cof<0, 0>() * m(0, 0) + cof<1, 0>() * m(1, 0) + cof<2, 0>() * m(2, 0);
            invdet = 1.0 / det;
        }
    }
    // cofactor (i,j) = m(i+1,j+1) m(i+2,j+2) - m(i+1,j+2) m(i+2,j+1) (indices mod 3)
    template <int i, int j>
    TINYAD_HD TINYAD_INLINE T cof() const
    {
        constexpr int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
    }
    // entry (i, j) of the inverse
    template <int i, int j>
    TINYAD_HD TINYAD_INLINE T coeff() const
    {
        if constexpr (R == 1) return invdet;
        else if constexpr (R == 2)
        {
            if constexpr (i == 0 && j == 0) return m(1, 1) * invdet;
            else if constexpr (i == 1 && j == 0) return -m(1, 0) * invdet;
            else if constexpr (i == 0 && j == 1) return -m(0, 1) * invdet;
            else return m(0, 0) * invdet;
        }
        else
            return cof<j, i>() * invdet;
    }
    TINYAD_HD TINYAD_INLINE Mat<T, R, R> eval() const
    {
        Mat<T, R, R> r;
        detail::static_for<R * R>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int l = decltype(ic)::value;
            r.a[l] = this->template coeff<l % R, l / R>();
        });
        return r;
    }
    TINYAD_HD TINYAD_INLINE operator Mat<T, R, R>() const { return eval(); }
    TINYAD_HD TINYAD_INLINE T operator()(int i, int j) const { return eval()(i, j); }
    TINYAD_HD TINYAD_INLINE T squaredNorm() const
    {
        T s = sqr(this->template coeff<0, 0>());
        detail::static_for<R * R - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int l = decltype(ic)::value + 1;
            s = s + sqr(this->template coeff<l % R, l / R>());
        });
        return s;
    }
    TINYAD_HD TINYAD_INLINE T norm() const { return sqrt(squaredNorm()); }
    TINYAD_HD TINYAD_INLINE Mat<T, R, R> transpose() const { return eval().transpose(); }
    TINYAD_HD TINYAD_INLINE T determinant() const { return invdet; }
    TINYAD_HD TINYAD_INLINE T trace() const { return eval().trace(); }
};

template <typename T, typename U, int R, int C>
TINYAD_HD TINYAD_INLINE auto operator*(const Mat<U, C, R>& x, const Inverse<T, R>& y) { return x * y.eval(); }
template <typename T, typename U, int R, int C>
TINYAD_HD TINYAD_INLINE auto operator*(const Inverse<T, R>& x, const Mat<U, R, C>& y) { return x.eval() * y; }
template <typename T, int R>
TINYAD_HD TINYAD_INLINE auto operator*(const double& s, const Inverse<T, R>& y) { return s * y.eval(); }
template <typename T, int R>
TINYAD_HD TINYAD_INLINE auto operator*(const Inverse<T, R>& y, const double& s) { return y.eval() * s; }

template <typename T, typename U, int R, int C>
TINYAD_HD TINYAD_INLINE auto operator+(const Mat<T, R, C>& x, const Mat<U, R, C>& y)
{
    Mat<decltype(x.a[0] + y.a[0]), R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = x.a[i] + y.a[i]; });
    return r;
}
template <typename T, typename U, int R, int C>
TINYAD_HD TINYAD_INLINE auto operator-(const Mat<T, R, C>& x, const Mat<U, R, C>& y)
{
    Mat<decltype(x.a[0] - y.a[0]), R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = x.a[i] - y.a[i]; });
    return r;
}
template <typename T, int R, int C>
TINYAD_HD TINYAD_INLINE Mat<T, R, C> operator-(const Mat<T, R, C>& x)
{
    Mat<T, R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = -x.a[i]; });
    return r;
}
template <typename T, typename U, int R, int K, int C>
TINYAD_HD TINYAD_INLINE auto operator*(const Mat<T, R, K>& x, const Mat<U, K, C>& y)
{
    Mat<decltype(x.a[0] * y.a[0]), R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE {
        constexpr int l = decltype(ic)::value;
        constexpr int i = l % R, j = l / R;
        auto s = x.a[i] * y.a[j * K];
        detail::static_for<K - 1>([&](auto lc) TINYAD_LAMBDA_INLINE { constexpr int q = decltype(lc)::value + 1; s = s + x.a[q * R + i] * y.a[j * K + q]; });
        r.a[l] = s;
    });
    return r;
}
// scalar * matrix, matrix * scalar, matrix / scalar (scalar: double or the active type)
template <typename T, int R, int C>
TINYAD_HD TINYAD_INLINE Mat<T, R, C> operator*(const double& s, const Mat<T, R, C>& x)
{
    Mat<T, R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = s * x.a[i]; });
    return r;
}
template <typename T, int R, int C>
TINYAD_HD TINYAD_INLINE Mat<T, R, C> operator*(const Mat<T, R, C>& x, const double& s)
{
    Mat<T, R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = x.a[i] * s; });
    return r;
}
template <typename T, int R, int C>
TINYAD_HD TINYAD_INLINE Mat<T, R, C> operator/(const Mat<T, R, C>& x, const double& s)
{
    Mat<T, R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = x.a[i] / s; });
    return r;
}
template <int k, bool wh, int NP, int P, typename U, int R, int C>
TINYAD_HD TINYAD_INLINE auto operator*(const Scalar<k, wh, NP, P>& s, const Mat<U, R, C>& x)
{
    Mat<Scalar<k, wh, NP, P>, R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = s * x.a[i]; });
    return r;
}
template <int k, bool wh, int NP, int P, typename U, int R, int C>
TINYAD_HD TINYAD_INLINE auto operator*(const Mat<U, R, C>& x, const Scalar<k, wh, NP, P>& s)
{
    Mat<Scalar<k, wh, NP, P>, R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = x.a[i] * s; });
    return r;
}

// Utils/Helpers.hh:48-76
template <typename T, int R>
TINYAD_HD TINYAD_INLINE Mat<T, R, 2> col_mat(const Mat<T, R, 1>& v0, const Mat<T, R, 1>& v1)
{
    Mat<T, R, 2> M;
    detail::static_for<R>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; M.a[i] = v0.a[i]; M.a[R + i] = v1.a[i]; });
    return M;
}
template <typename T, int R>
TINYAD_HD TINYAD_INLINE Mat<T, R, 3> col_mat(const Mat<T, R, 1>& v0, const Mat<T, R, 1>& v1, const Mat<T, R, 1>& v2)
{
    Mat<T, R, 3> M;
    detail::static_for<R>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; M.a[i] = v0.a[i]; M.a[R + i] = v1.a[i]; M.a[2 * R + i] = v2.a[i]; });
    return M;
}

// Scalar.hh:1371-1383
template <int k, bool wh, int NP, int P, int R, int C>
TINYAD_HD TINYAD_INLINE Mat<double, R, C> to_passive(const Mat<Scalar<k, wh, NP, P>, R, C>& A)
{
    Mat<double, R, C> r;
    detail::static_for<R * C>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; r.a[i] = A.a[i].val; });
    return r;
}
template <int R, int C>
TINYAD_HD TINYAD_INLINE Mat<double, R, C> to_passive(const Mat<double, R, C>& A) { return A; }

}  // namespace TinyAD
