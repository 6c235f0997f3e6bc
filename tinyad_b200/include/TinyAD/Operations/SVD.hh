// tinyad_b200 -- closed-form 2 x 2 SVD and closest orthogonal matrix on passive or active (device) scalars:
// Operations/SVD.hh:12-99 of the reference, with TinyAD::Mat / Vec instead of Eigen fixed-size types.
// Same sequence of scalar operations as the reference (atan2 of the symmetrised off-diagonal, cos / sin of the half angles,
// singular values from the trace and the discriminant of A A^T, sign correction from U^T A W), so derivatives agree with it.
#pragma once

#include <TinyAD/Matrix.hh>
#include <TinyAD/Scalar.hh>

namespace TinyAD
{

template <typename T>
TINYAD_HD TINYAD_INLINE int sign(const T& _x)  // :12-21
{
    if (_x < T(0.0)) return -1;
    else if (_x > T(0.0)) return 1;
    else return 0;
}

namespace detail
{
// U (rotation by phi) and W (rotation by theta) of the reference's construction, :39-53 / :76-90
template <typename T>
TINYAD_HD TINYAD_INLINE void svd_rotations(const Mat<T, 2, 2>& _A, Mat<T, 2, 2>& _Su, Mat<T, 2, 2>& _U, Mat<T, 2, 2>& _W)
{
    _Su = _A * _A.transpose();
    T phi = 0.5 * atan2(_Su(0, 1) + _Su(1, 0), _Su(0, 0) - _Su(1, 1));
    T Cphi = cos(phi);
    T Sphi = sin(phi);
    _U(0, 0) = Cphi; _U(0, 1) = -Sphi;
    _U(1, 0) = Sphi; _U(1, 1) = Cphi;
    Mat<T, 2, 2> Sw = _A.transpose() * _A;
    T theta = 0.5 * atan2(Sw(0, 1) + Sw(1, 0), Sw(0, 0) - Sw(1, 1));
    T Ctheta = cos(theta);
    T Stheta = sin(theta);
    _W(0, 0) = Ctheta; _W(0, 1) = -Stheta;
    _W(1, 0) = Stheta; _W(1, 1) = Ctheta;
}
// V = W * diag(sign(S00), sign(S11)), S = U^T A W   (:60-63 / :92-95)
template <typename T>
TINYAD_HD TINYAD_INLINE Mat<T, 2, 2> svd_right(const Mat<T, 2, 2>& _A, const Mat<T, 2, 2>& _U, const Mat<T, 2, 2>& _W)
{
    Mat<T, 2, 2> S = _U.transpose() * _A * _W;
    const double c0 = (double)sign(S(0, 0)), c1 = (double)sign(S(1, 1));
    Mat<T, 2, 2> V;
    V(0, 0) = _W(0, 0) * c0; V(0, 1) = _W(0, 1) * c1;
    V(1, 0) = _W(1, 0) * c0; V(1, 1) = _W(1, 1) * c1;
    return V;
}
}  // namespace detail

// 2x2 closed-form SVD, A = U * diag(S) * V^T   (:26-64)
template <typename T>
TINYAD_HD TINYAD_INLINE void svd(const Mat<T, 2, 2>& _A, Mat<T, 2, 2>& _U, Vec<T, 2>& _S, Mat<T, 2, 2>& _V)
{
    Mat<T, 2, 2> Su, W;
    detail::svd_rotations(_A, Su, _U, W);
    T SUsum = Su(0, 0) + Su(1, 1);
    T SUdif = sqrt(sqr(Su(0, 0) - Su(1, 1)) + 4.0 * Su(0, 1) * Su(1, 0));
    _S[0] = sqrt((SUsum + SUdif) / 2.0);
    _S[1] = sqrt((SUsum - SUdif) / 2.0);
    _V = detail::svd_right(_A, _U, W);
}

// Closest orthogonal 2x2 matrix, U * V^T for the SVD A = U * S * V^T   (:70-99)
template <typename T>
TINYAD_HD TINYAD_INLINE Mat<T, 2, 2> closest_orthogonal(const Mat<T, 2, 2>& _A)
{
    Mat<T, 2, 2> Su, U, W;
    detail::svd_rotations(_A, Su, U, W);
    Mat<T, 2, 2> V = detail::svd_right(_A, U, W);
    return U * V.transpose();
}

}  // namespace TinyAD
