// tinyad_b200 -- forward-mode second-order AD scalar for sm_100a device code (and host).
//
// Same surface and derivative semantics as the reference's TinyAD::Scalar
// (include/TinyAD/Scalar.hh:24-1347 in /root/reference), redesigned for registers:
//   * the Hessian is PACKED symmetric (h = k(k+1)/2 doubles, tile order -- see
//     Detail/HessLayout.hh) instead of a full k x k Eigen matrix (Scalar.hh:1344);
//   * an element's Hessian can be split over NP cooperating threads: instantiation
//     <k, true, NP, P> carries only part P of the packed entries (plus val and the
//     gradient, which every part recomputes; unused gradient lanes are dead code);
//   * every scalar carries structural sparsity masks of its gradient and Hessian part; after
//     inlining they are compile-time constants, so only structurally non-zero entries cost
//     flops and registers (e.g. the entries of a deformation gradient depend on 4 of 12
//     variables and have no Hessian) -- the reference multiplies the zeros out.
// Everything is fully unrolled (static_for) so all indices are compile-time constants
// and the state lives in registers.
#pragma once

#include <cmath>
#include <cstdint>
#include <type_traits>
#include <utility>

#include <TinyAD/Detail/HessLayout.hh>

namespace TinyAD
{

namespace detail
{
template <class F, int... I>
TINYAD_HD TINYAD_INLINE void static_for_impl(F&& f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
// Calls f(integral_constant<int, 0>) ... f(integral_constant<int, N-1>).
template <int N, class F>
TINYAD_HD TINYAD_INLINE void static_for(F&& f)
{
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// h (+)= t, where `have` says whether h already holds a term (compile-time after inlining)
TINYAD_HD TINYAD_INLINE void acc(double& h, bool& have, double t)
{
    h = have ? h + t : t;
    have = true;
}
// Which packed Hessian entries a Scalar<k, true, NP, P> carries:
//   NP >= 1:  part P of NP, a contiguous cut of the tile order (the cooperating-thread split of the element kernels);
//   NP == -1: the TRUNCATED HESSIAN BLOCK of the reference (Scalar.hh:24,31-38,188-196: hess_row_start, hess_col_start,
//             hess_rows, hess_cols), P = hess_block_code(r0, c0, nr, nc): the packed entries (i, j) ~ (j, i) with one index in the
//             row range and the other in the column range.  Entry (i, j) of a result only depends on entry (i, j) and gradient
//             components i, j of the operands, so every operator works on any such subset unchanged.
TINYAD_HD constexpr int hess_block_code(int r0, int c0, int nr, int nc) { return r0 | (c0 << 6) | (nr << 12) | (nc << 18); }
TINYAD_HD constexpr bool hess_block_has(int code, int i, int j)
{
    const int r0 = code & 63, c0 = (code >> 6) & 63, nr = (code >> 12) & 63, nc = (code >> 18) & 63;
    return (i >= r0 && i < r0 + nr && j >= c0 && j < c0 + nc) || (j >= r0 && j < r0 + nr && i >= c0 && i < c0 + nc);
}
// number of selected packed entries
TINYAD_HD constexpr int hess_sel_count(int k, bool wh, int np, int p)
{
    if (!wh) return 0;
    if (np >= 1) return hess_part_begin(k, np, p + 1) - hess_part_begin(k, np, p);
    int n = 0;
    for (int s = 0; s < hess_size(k); ++s) n += hess_block_has(p, hess_seq_rc(k, s).row, hess_seq_rc(k, s).col) ? 1 : 0;
    return n;
}
// tile-order index of the e-th selected entry
TINYAD_HD constexpr int hess_sel_seq(int k, bool wh, int np, int p, int e)
{
    if (np >= 1) return (wh ? hess_part_begin(k, np, p) : 0) + e;
    int n = 0;
    for (int s = 0; s < hess_size(k); ++s)
        if (hess_block_has(p, hess_seq_rc(k, s).row, hess_seq_rc(k, s).col))
        {
            if (n == e) return s;
            ++n;
        }
    return -1;
}
// position of tile-order index s among the selected entries, -1 if it is not selected
TINYAD_HD constexpr int hess_sel_local(int k, bool wh, int np, int p, int s)
{
    if (!wh) return -1;
    if (np >= 1)
    {
        const int b = hess_part_begin(k, np, p), e = hess_part_begin(k, np, p + 1);
        return (s >= b && s < e) ? s - b : -1;
    }
    int n = 0;
    for (int q = 0; q < hess_size(k); ++q)
        if (hess_block_has(p, hess_seq_rc(k, q).row, hess_seq_rc(k, q).col))
        {
            if (q == s) return n;
            ++n;
        }
    return -1;
}
// (row, col) of the e-th packed entry held by the scalar (free functions: usable inside generic lambdas of friends)
template <int k, bool wh, int NP, int P>
TINYAD_HD constexpr int part_row(int e) { return hess_seq_rc(k, hess_sel_seq(k, wh, NP, P, e)).row; }
template <int k, bool wh, int NP, int P>
TINYAD_HD constexpr int part_col(int e) { return hess_seq_rc(k, hess_sel_seq(k, wh, NP, P, e)).col; }
}  // namespace detail

template <int k, bool with_hessian = true, int NP = 1, int P = 0>
struct Scalar
{
    static_assert(k >= 0 && k <= 32, "static k <= 32 only (dynamic mode of Scalar.hh:30 is out of scope)");
    static_assert((NP >= 1 && P >= 0 && P < NP) || (NP == -1 && with_hessian && P >= 0), "bad Hessian partition / block");
    static constexpr int k_ = k;
    static constexpr bool with_hessian_ = with_hessian;
    static constexpr int n_parts_ = NP;
    static constexpr int part_ = P;
    static constexpr bool truncated_hessian_ = NP == -1;   // Scalar.hh:34
    static constexpr int h_begin = (with_hessian && NP >= 1) ? detail::hess_part_begin(k, NP, P) : 0;   // parts only (staging index)
    static constexpr int nh = detail::hess_sel_count(k, with_hessian, NP, P);  // packed entries held by this scalar
    static constexpr int HW = nh > 0 ? (nh + 63) / 64 : 1;
    static constexpr uint32_t g_all = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);

    // (row, col) of local packed entry e
    TINYAD_HD static constexpr int row(int e) { return detail::hess_seq_rc(k, detail::hess_sel_seq(k, with_hessian, NP, P, e)).row; }
    TINYAD_HD static constexpr int col(int e) { return detail::hess_seq_rc(k, detail::hess_sel_seq(k, with_hessian, NP, P, e)).col; }

    // ---- data (Scalar.hh:1341-1346) ----
    double val;
    double grad[k > 0 ? k : 1];
    double hess[nh > 0 ? nh : 1];  // packed lower triangle, part P
    // Structural sparsity: bit i of gm / bit e of hm is set iff grad[i] / hess[e] MAY be non-zero.
    // Entries whose bit is clear hold an exact 0.0.  After inlining the masks are compile-time
    // constants (seeds are constants, every operator combines masks with | and &), so the
    // `if (bit)` tests below fold away and only structurally non-zero entries cost flops/registers.
    uint32_t gm;
    uint64_t hm[HW];

    TINYAD_HD TINYAD_INLINE bool g(int i) const { return (gm >> i) & 1u; }
    TINYAD_HD TINYAD_INLINE bool h(int e) const { return (hm[e >> 6] >> (e & 63)) & 1ull; }
    TINYAD_HD TINYAD_INLINE void set_h(int e) { hm[e >> 6] |= (1ull << (e & 63)); }
    TINYAD_HD TINYAD_INLINE bool gz() const { return gm == 0u; }   // passive constant
    TINYAD_HD TINYAD_INLINE bool hz() const                        // Hessian identically zero
    {
        uint64_t any = 0;
        for (int w = 0; w < HW; ++w) any |= hm[w];
        return any == 0;
    }

    // ---- constructors (Scalar.hh:54-78) ----
    TINYAD_HD TINYAD_INLINE Scalar() : val(0.0), gm(0u) { zero_derivs(); }
    TINYAD_HD TINYAD_INLINE Scalar(double _val) : val(_val), gm(0u) { zero_derivs(); }
    TINYAD_HD TINYAD_INLINE Scalar(double _val, int _idx) : val(_val), gm(1u << _idx)
    {
        zero_derivs();
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; grad[i] = (i == _idx) ? 1.0 : 0.0; });
    }
    // Active variable whose index is only known at run time: dense masks, no mask-dependent branches.
    TINYAD_HD TINYAD_INLINE static Scalar active_dense(double _val, int _idx)
    {
        Scalar res(_val);
        res.gm = g_all;
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; res.grad[i] = (i == _idx) ? 1.0 : 0.0; });
        return res;
    }

    // Known derivatives; _Hess is a full row-major k x k matrix (Scalar.hh:81-106).
    TINYAD_HD static Scalar known_derivatives(double _val, const double* _grad, const double* _Hess)
    {
        Scalar res;
        res.val = _val;
        res.gm = g_all;
        for (int i = 0; i < k; ++i) res.grad[i] = _grad[i];
        for (int e = 0; e < nh; ++e)
        {
            res.hess[e] = _Hess[row(e) * k + col(e)];
            res.set_h(e);
        }
        return res;
    }
    TINYAD_HD static Scalar known_derivatives(double _val, double _grad, double _Hess)
    {
        static_assert(k == 1, "univariate only");
        return known_derivatives(_val, &_grad, &_Hess);
    }

    // Hessian entry (i, j); only valid for entries owned by this part (all of them for NP == 1).
    TINYAD_HD double Hess(int i, int j) const
    {
        const int s = detail::hess_sel_local(k, with_hessian, NP, P, detail::hess_seq_index(k, i, j));
        return s >= 0 ? hess[s] : 0.0;
    }
    // Truncated block (NP == -1): entry (i, j) of the hess_rows x hess_cols block, i.e. d^2 f / dx_{r0+i} dx_{c0+j} (Scalar.hh:188-196)
    TINYAD_HD double HessBlock(int i, int j) const
    {
        static_assert(NP == -1, "HessBlock() belongs to truncated-Hessian scalars");
        return Hess((P & 63) + i, ((P >> 6) & 63) + j);
    }

    // ---- chain rule (Scalar.hh:199-214) ----
    TINYAD_HD TINYAD_INLINE static Scalar chain(double f, double df, double ddf, const Scalar& a)
    {
        Scalar res;
        res.val = f;
        res.gm = a.gm;
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            if (a.g(i)) res.grad[i] = df * a.grad[i];
        });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int e = decltype(ic)::value;
            constexpr int i = Scalar::row(e), j = Scalar::col(e);
            double hv = 0.0;
            bool have = false;
            if (a.g(i) && a.g(j)) detail::acc(hv, have, ddf * (a.grad[i] * a.grad[j]));
            if (a.h(e)) detail::acc(hv, have, df * a.hess[e]);
            if (have) { res.hess[e] = hv; res.set_h(e); }
        });
        return res;
    }

    TINYAD_HD TINYAD_INLINE static Scalar neg_(const Scalar& a)
    {
        Scalar res;
        res.val = -a.val;
        res.gm = a.gm;
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; if (a.g(i)) res.grad[i] = -a.grad[i]; });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int e = decltype(ic)::value;
            if (a.h(e)) { res.hess[e] = -a.hess[e]; res.set_h(e); }
        });
        return res;
    }
    TINYAD_HD TINYAD_INLINE static Scalar sqr_(const Scalar& a)
    {
        Scalar res;
        res.val = a.val * a.val;
        res.gm = a.gm;
        const double two_a = 2.0 * a.val;
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; if (a.g(i)) res.grad[i] = two_a * a.grad[i]; });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int e = decltype(ic)::value;
            constexpr int i = Scalar::row(e), j = Scalar::col(e);
            double hv = 0.0;
            bool have = false;
            if (a.h(e)) detail::acc(hv, have, a.val * a.hess[e]);
            if (a.g(i) && a.g(j)) detail::acc(hv, have, a.grad[i] * a.grad[j]);
            if (have) { res.hess[e] = 2.0 * hv; res.set_h(e); }
        });
        return res;
    }
    template <bool Minus>
    TINYAD_HD TINYAD_INLINE static Scalar addsub_(const Scalar& a, const Scalar& b)
    {
        Scalar res;
        res.val = Minus ? a.val - b.val : a.val + b.val;
        res.gm = a.gm | b.gm;
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            const bool ca = a.g(i), cb = b.g(i);
            if (ca && cb) res.grad[i] = Minus ? a.grad[i] - b.grad[i] : a.grad[i] + b.grad[i];
            else if (ca) res.grad[i] = a.grad[i];
            else if (cb) res.grad[i] = Minus ? -b.grad[i] : b.grad[i];
        });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int e = decltype(ic)::value;
            const bool ca = a.h(e), cb = b.h(e);
            if (ca && cb) res.hess[e] = Minus ? a.hess[e] - b.hess[e] : a.hess[e] + b.hess[e];
            else if (ca) res.hess[e] = a.hess[e];
            else if (cb) res.hess[e] = Minus ? -b.hess[e] : b.hess[e];
            if (ca || cb) res.set_h(e);
        });
        return res;
    }
    TINYAD_HD TINYAD_INLINE static Scalar add_(const Scalar& a, const Scalar& b) { return addsub_<false>(a, b); }
    TINYAD_HD TINYAD_INLINE static Scalar sub_(const Scalar& a, const Scalar& b) { return addsub_<true>(a, b); }
    TINYAD_HD TINYAD_INLINE static Scalar mul_(const Scalar& a, const Scalar& b)
    {
        Scalar res;
        res.val = a.val * b.val;
        res.gm = a.gm | b.gm;
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            const bool ca = a.g(i), cb = b.g(i);
            if (ca && cb) res.grad[i] = b.val * a.grad[i] + a.val * b.grad[i];
            else if (ca) res.grad[i] = b.val * a.grad[i];
            else if (cb) res.grad[i] = a.val * b.grad[i];
        });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE {  // Scalar.hh:765, same left-to-right order
            constexpr int e = decltype(ic)::value;
            constexpr int i = Scalar::row(e), j = Scalar::col(e);
            double hv = 0.0;
            bool have = false;
            if (a.h(e)) detail::acc(hv, have, b.val * a.hess[e]);
            if (a.g(i) && b.g(j)) detail::acc(hv, have, a.grad[i] * b.grad[j]);
            if (b.g(i) && a.g(j)) detail::acc(hv, have, b.grad[i] * a.grad[j]);
            if (b.h(e)) detail::acc(hv, have, a.val * b.hess[e]);
            if (have) { res.hess[e] = hv; res.set_h(e); }
        });
        return res;
    }
    TINYAD_HD TINYAD_INLINE static Scalar muls_(const Scalar& a, const double& b)
    {
        Scalar res;
        res.val = a.val * b;
        res.gm = a.gm;
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; if (a.g(i)) res.grad[i] = a.grad[i] * b; });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int e = decltype(ic)::value;
            if (a.h(e)) { res.hess[e] = a.hess[e] * b; res.set_h(e); }
        });
        return res;
    }
    TINYAD_HD TINYAD_INLINE static Scalar div_(const Scalar& a, const Scalar& b)
    {
        // Scalar.hh:823-841; divisions by b replaced by one reciprocal (differs by <= 1 ulp per entry)
        const double inv_b = 1.0 / b.val;
        Scalar res;
        res.val = a.val * inv_b;
        res.gm = a.gm | b.gm;
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            const bool ca = a.g(i), cb = b.g(i);
            if (ca && cb) res.grad[i] = (a.grad[i] - res.val * b.grad[i]) * inv_b;
            else if (ca) res.grad[i] = a.grad[i] * inv_b;
            else if (cb) res.grad[i] = -(res.val * b.grad[i]) * inv_b;
        });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int e = decltype(ic)::value;
            constexpr int i = Scalar::row(e), j = Scalar::col(e);
            double hv = 0.0;
            bool have = false;
            if (a.h(e)) detail::acc(hv, have, a.hess[e]);
            if (res.g(i) && b.g(j)) detail::acc(hv, have, -(res.grad[i] * b.grad[j]));
            if (b.g(i) && res.g(j)) detail::acc(hv, have, -(b.grad[i] * res.grad[j]));
            if (b.h(e)) detail::acc(hv, have, -(res.val * b.hess[e]));
            if (have) { res.hess[e] = hv * inv_b; res.set_h(e); }
        });
        return res;
    }
    TINYAD_HD TINYAD_INLINE static Scalar rdiv_(const double& a, const Scalar& b)
    {
        // Scalar.hh:861-877
        const double inv_b = 1.0 / b.val;
        Scalar res;
        res.val = a * inv_b;
        res.gm = b.gm;
        const double c = -res.val * inv_b;
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; if (b.g(i)) res.grad[i] = c * b.grad[i]; });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int e = decltype(ic)::value;
            constexpr int i = Scalar::row(e), j = Scalar::col(e);
            double hv = 0.0;
            bool have = false;
            if (b.g(i) && b.g(j))
            {
                detail::acc(hv, have, -(res.grad[i] * b.grad[j]));
                detail::acc(hv, have, -(b.grad[i] * res.grad[j]));
            }
            if (b.h(e)) detail::acc(hv, have, -(res.val * b.hess[e]));
            if (have) { res.hess[e] = hv * inv_b; res.set_h(e); }
        });
        return res;
    }
    TINYAD_HD TINYAD_INLINE static Scalar atan2_(const Scalar& y, const Scalar& x)
    {
        // Scalar.hh:893-919
        Scalar res;
        res.val = ::atan2(y.val, x.val);
        res.gm = x.gm | y.gm;
        const double inv_v = 1.0 / (x.val * x.val + y.val * y.val);
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            const bool cy = y.g(i), cx = x.g(i);
            if (cy && cx) res.grad[i] = (x.val * y.grad[i] - y.val * x.grad[i]) * inv_v;
            else if (cy) res.grad[i] = (x.val * y.grad[i]) * inv_v;
            else if (cx) res.grad[i] = -(y.val * x.grad[i]) * inv_v;
        });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int e = decltype(ic)::value;
            constexpr int i = Scalar::row(e), j = Scalar::col(e);
            // Entry (i,j), i >= j, of the reference's full matrix (du - grad dv^T)/v.  The antisymmetric
            // parts of du and of grad dv^T cancel exactly, so (j,i) is the same number.
            double hv = 0.0;
            bool have = false;
            if (y.h(e)) detail::acc(hv, have, x.val * y.hess[e]);
            if (x.h(e)) detail::acc(hv, have, -(y.val * x.hess[e]));
            if (y.g(i) && x.g(j)) detail::acc(hv, have, y.grad[i] * x.grad[j]);
            if (x.g(i) && y.g(j)) detail::acc(hv, have, -(x.grad[i] * y.grad[j]));
            if (res.g(i) && res.g(j))
            {
                const double dv_j = 2.0 * (x.val * x.grad[j] + y.val * y.grad[j]);
                detail::acc(hv, have, -(res.grad[i] * dv_j));
            }
            if (have) { res.hess[e] = hv * inv_v; res.set_h(e); }
        });
        return res;
    }

    // nvcc's front end does not see the class scope from generic lambdas inside in-class friend
    // definitions, so the operators with unrolled bodies are the static members above and the friends forward.
    TINYAD_HD TINYAD_INLINE friend Scalar operator-(const Scalar& a) { return neg_(a); }
    TINYAD_HD TINYAD_INLINE friend Scalar sqr(const Scalar& a) { return sqr_(a); }
    TINYAD_HD TINYAD_INLINE friend Scalar operator+(const Scalar& a, const Scalar& b) { return add_(a, b); }
    TINYAD_HD TINYAD_INLINE friend Scalar operator-(const Scalar& a, const Scalar& b) { return sub_(a, b); }
    TINYAD_HD TINYAD_INLINE friend Scalar operator*(const Scalar& a, const Scalar& b) { return mul_(a, b); }
    TINYAD_HD TINYAD_INLINE friend Scalar operator*(const Scalar& a, const double& b) { return muls_(a, b); }
    TINYAD_HD TINYAD_INLINE friend Scalar operator/(const Scalar& a, const Scalar& b) { return div_(a, b); }
    TINYAD_HD TINYAD_INLINE friend Scalar operator/(const double& a, const Scalar& b) { return rdiv_(a, b); }
    TINYAD_HD TINYAD_INLINE friend Scalar atan2(const Scalar& y, const Scalar& x) { return atan2_(y, x); }

    // ---- unary math (Scalar.hh:236-573) ----
    TINYAD_HD TINYAD_INLINE friend Scalar sqrt(const Scalar& a)
    {
        const double f = ::sqrt(a.val);
        return chain(f, 0.5 / f, -0.25 / (f * a.val), a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar pow(const Scalar& a, const int& e)
    {
        if (e == 0) return chain(1.0, 0.0, 0.0, a);
        else if (e == 1) return chain(a.val, 1.0, 0.0, a);
        else
        {
            const double f2 = ::pow(a.val, (double)(e - 2));
            const double f1 = f2 * a.val;
            const double f = f1 * a.val;
            return chain(f, e * f1, e * (e - 1) * f2, a);
        }
    }
    TINYAD_HD TINYAD_INLINE friend Scalar pow(const Scalar& a, const double& e)
    {
        const double f2 = ::pow(a.val, e - 2.0);
        const double f1 = f2 * a.val;
        const double f = f1 * a.val;
        return chain(f, e * f1, e * (e - 1.0) * f2, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar fabs(const Scalar& a)
    {
        if (a.val >= 0.0) return chain(a.val, 1.0, 0.0, a);
        else return chain(-a.val, -1.0, 0.0, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar abs(const Scalar& a) { return fabs(a); }
    TINYAD_HD TINYAD_INLINE friend Scalar exp(const Scalar& a)
    {
        const double e = ::exp(a.val);
        return chain(e, e, e, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar log(const Scalar& a)
    {
        const double a_inv = 1.0 / a.val;
        return chain(::log(a.val), a_inv, -a_inv / a.val, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar log2(const Scalar& a)
    {
        const double a_inv = 1.0 / a.val / ::log(2.0);
        return chain(::log2(a.val), a_inv, -a_inv / a.val, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar log10(const Scalar& a)
    {
        const double a_inv = 1.0 / a.val / ::log(10.0);
        return chain(::log10(a.val), a_inv, -a_inv / a.val, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar sin(const Scalar& a)
    {
        const double s = ::sin(a.val);
        return chain(s, ::cos(a.val), -s, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar cos(const Scalar& a)
    {
        const double c = ::cos(a.val);
        return chain(c, -::sin(a.val), -c, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar tan(const Scalar& a)
    {
        const double c = ::cos(a.val);
        const double c2 = c * c;
        const double c3 = c2 * c;
        return chain(::tan(a.val), 1.0 / c2, 2.0 * ::sin(a.val) / c3, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar asin(const Scalar& a)
    {
        const double s = 1.0 - a.val * a.val;
        const double s_sqrt = ::sqrt(s);
        return chain(::asin(a.val), 1.0 / s_sqrt, a.val / s_sqrt / s, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar acos(const Scalar& a)
    {
        // The reference asserts -1 < a < 1 and throws (Scalar.hh:433-441); device code cannot
        // throw: out-of-range arguments give NaN derivatives, which the finite check reports.
        const double s = 1.0 - a.val * a.val;
        const double s_sqrt = ::sqrt(s);
        return chain(::acos(a.val), -1.0 / s_sqrt, -a.val / s_sqrt / s, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar atan(const Scalar& a)
    {
        const double s = a.val * a.val + 1.0;
        return chain(::atan(a.val), 1.0 / s, -2.0 * a.val / s / s, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar sinh(const Scalar& a)
    {
        const double s = ::sinh(a.val);
        return chain(s, ::cosh(a.val), s, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar cosh(const Scalar& a)
    {
        const double c = ::cosh(a.val);
        return chain(c, ::sinh(a.val), c, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar tanh(const Scalar& a)
    {
        const double c = ::cosh(a.val);
        const double c2 = c * c;
        const double c3 = c2 * c;
        return chain(::tanh(a.val), 1.0 / c2, -2.0 * ::sinh(a.val) / c3, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar asinh(const Scalar& a)
    {
        const double s = a.val * a.val + 1.0;
        const double s_sqrt = ::sqrt(s);
        return chain(::asinh(a.val), 1.0 / s_sqrt, -a.val / s_sqrt / s, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar acosh(const Scalar& a)
    {
        const double sm = a.val - 1.0;
        const double sp = a.val + 1.0;
        const double prod = ::sqrt(sm) * ::sqrt(sp);
        return chain(::acosh(a.val), 1.0 / prod, -a.val / prod / sm / sp, a);
    }
    TINYAD_HD TINYAD_INLINE friend Scalar atanh(const Scalar& a)
    {
        const double s = 1.0 - a.val * a.val;
        return chain(::atanh(a.val), 1.0 / s, 2.0 * a.val / s / s, a);
    }
    TINYAD_HD TINYAD_INLINE friend bool isnan(const Scalar& a) { return a.val != a.val; }
    TINYAD_HD TINYAD_INLINE friend bool isinf(const Scalar& a) { return a.val == INFINITY || a.val == -INFINITY; }
    TINYAD_HD TINYAD_INLINE friend bool isfinite(const Scalar& a) { return !isnan(a) && !isinf(a); }

    // ---- remaining binary operators (Scalar.hh:599-891) ----
    TINYAD_HD TINYAD_INLINE friend Scalar operator+(const Scalar& a, const double& b) { Scalar res = a; res.val += b; return res; }
    TINYAD_HD TINYAD_INLINE friend Scalar operator+(const double& a, const Scalar& b) { Scalar res = b; res.val += a; return res; }
    TINYAD_HD TINYAD_INLINE Scalar& operator+=(const Scalar& b) { *this = *this + b; return *this; }
    TINYAD_HD TINYAD_INLINE Scalar& operator+=(const double& b) { val += b; return *this; }
    TINYAD_HD TINYAD_INLINE friend Scalar operator-(const Scalar& a, const double& b) { Scalar res = a; res.val -= b; return res; }
    TINYAD_HD TINYAD_INLINE friend Scalar operator-(const double& a, const Scalar& b)
    {
        Scalar res = neg_(b);
        res.val = a - b.val;
        return res;
    }
    TINYAD_HD TINYAD_INLINE Scalar& operator-=(const Scalar& b) { *this = *this - b; return *this; }
    TINYAD_HD TINYAD_INLINE Scalar& operator-=(const double& b) { val -= b; return *this; }
    TINYAD_HD TINYAD_INLINE friend Scalar operator*(const double& a, const Scalar& b) { return muls_(b, a); }
    TINYAD_HD TINYAD_INLINE Scalar& operator*=(const Scalar& b) { *this = *this * b; return *this; }
    TINYAD_HD TINYAD_INLINE Scalar& operator*=(const double& b) { *this = *this * b; return *this; }
    TINYAD_HD TINYAD_INLINE friend Scalar operator/(const Scalar& a, const double& b) { return muls_(a, 1.0 / b); }
    TINYAD_HD TINYAD_INLINE Scalar& operator/=(const Scalar& b) { *this = *this / b; return *this; }
    TINYAD_HD TINYAD_INLINE Scalar& operator/=(const double& b) { *this = *this / b; return *this; }
    TINYAD_HD TINYAD_INLINE friend Scalar hypot(const Scalar& a, const Scalar& b) { return sqrt(a * a + b * b); }

    // ---- comparisons on val (Scalar.hh:933-1095) ----
    TINYAD_HD TINYAD_INLINE friend bool operator==(const Scalar& a, const Scalar& b) { return a.val == b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator==(const Scalar& a, const double& b) { return a.val == b; }
    TINYAD_HD TINYAD_INLINE friend bool operator==(const double& a, const Scalar& b) { return a == b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator!=(const Scalar& a, const Scalar& b) { return a.val != b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator!=(const Scalar& a, const double& b) { return a.val != b; }
    TINYAD_HD TINYAD_INLINE friend bool operator!=(const double& a, const Scalar& b) { return a != b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator<(const Scalar& a, const Scalar& b) { return a.val < b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator<(const Scalar& a, const double& b) { return a.val < b; }
    TINYAD_HD TINYAD_INLINE friend bool operator<(const double& a, const Scalar& b) { return a < b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator<=(const Scalar& a, const Scalar& b) { return a.val <= b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator<=(const Scalar& a, const double& b) { return a.val <= b; }
    TINYAD_HD TINYAD_INLINE friend bool operator<=(const double& a, const Scalar& b) { return a <= b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator>(const Scalar& a, const Scalar& b) { return a.val > b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator>(const Scalar& a, const double& b) { return a.val > b; }
    TINYAD_HD TINYAD_INLINE friend bool operator>(const double& a, const Scalar& b) { return a > b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator>=(const Scalar& a, const Scalar& b) { return a.val >= b.val; }
    TINYAD_HD TINYAD_INLINE friend bool operator>=(const Scalar& a, const double& b) { return a.val >= b; }
    TINYAD_HD TINYAD_INLINE friend bool operator>=(const double& a, const Scalar& b) { return a >= b.val; }
    // Scalar.hh:1097-1145
    TINYAD_HD TINYAD_INLINE friend Scalar min(const Scalar& a, const Scalar& b) { return (b < a) ? b : a; }
    TINYAD_HD TINYAD_INLINE friend Scalar fmin(const Scalar& a, const Scalar& b) { return min(a, b); }
    TINYAD_HD TINYAD_INLINE friend Scalar max(const Scalar& a, const Scalar& b) { return (a < b) ? b : a; }
    TINYAD_HD TINYAD_INLINE friend Scalar fmax(const Scalar& a, const Scalar& b) { return max(a, b); }
    TINYAD_HD TINYAD_INLINE friend Scalar clamp(const Scalar& x, const Scalar& a, const Scalar& b)
    {
        if (x < a) return a;
        else if (x > b) return b;
        else return x;
    }

private:
    TINYAD_HD TINYAD_INLINE void zero_derivs()
    {
        detail::static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; grad[i] = 0.0; });
        detail::static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int e = decltype(ic)::value; hess[e] = 0.0; });
        detail::static_for<HW>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int w = decltype(ic)::value; hm[w] = 0ull; });
    }
};

// Scalar.hh:1389-1391 (double only; Float / LongDouble are out of scope)
template <int k, bool with_hessian = true>
using Double = Scalar<k, with_hessian>;

// Utils/ToPassive.hh:15-30, Scalar.hh:1368-1369
TINYAD_HD TINYAD_INLINE double to_passive(const double& a) { return a; }
// Scalar<k, double, true, hess_row_start, hess_col_start, hess_rows, hess_cols> of the reference (Scalar.hh:24-38): only the
// selected block of the Hessian is computed and stored; read it with HessBlock(i, j).
template <int k, int hess_row_start, int hess_col_start, int hess_rows, int hess_cols>
using ScalarHessianBlock = Scalar<k, true, -1, detail::hess_block_code(hess_row_start, hess_col_start, hess_rows, hess_cols)>;

template <int k, bool wh, int NP, int P>
TINYAD_HD TINYAD_INLINE double to_passive(const Scalar<k, wh, NP, P>& a) { return a.val; }
TINYAD_HD TINYAD_INLINE double sqr(const double& x) { return x * x; }

// ---------------------------------------------------------------------------
// Complex numbers over active scalars.  The reference overloads operators on
// std::complex<Scalar> (Scalar.hh:1151-1320); std::complex is not usable in device
// code, so the same formulas live on this small POD.
// ---------------------------------------------------------------------------
template <typename T>
struct Complex
{
    T re, im;
    TINYAD_HD TINYAD_INLINE Complex() : re(0.0), im(0.0) {}
    TINYAD_HD TINYAD_INLINE Complex(const T& r, const T& i) : re(r), im(i) {}
    TINYAD_HD TINYAD_INLINE explicit Complex(const T& r) : re(r), im(0.0) {}
    TINYAD_HD TINYAD_INLINE const T& real() const { return re; }
    TINYAD_HD TINYAD_INLINE const T& imag() const { return im; }
};

template <typename A, typename B>
TINYAD_HD TINYAD_INLINE auto operator+(const Complex<A>& a, const Complex<B>& b)
{
    return Complex<decltype(a.re + b.re)>(a.re + b.re, a.im + b.im);
}
template <typename A, typename B>
TINYAD_HD TINYAD_INLINE auto operator-(const Complex<A>& a, const Complex<B>& b)
{
    return Complex<decltype(a.re - b.re)>(a.re - b.re, a.im - b.im);
}
template <typename A, typename B>
TINYAD_HD TINYAD_INLINE auto operator*(const Complex<A>& a, const Complex<B>& b)
{
    return Complex<decltype(a.re * b.re)>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename A, typename B>
TINYAD_HD TINYAD_INLINE auto operator/(const Complex<A>& a, const Complex<B>& b)
{
    const auto denom = b.re * b.re + b.im * b.im;
    return Complex<decltype(a.re * b.re / denom)>((a.re * b.re + a.im * b.im) / denom, (a.im * b.re - a.re * b.im) / denom);
}
template <typename A>
TINYAD_HD TINYAD_INLINE Complex<A> sqr(const Complex<A>& a)
{
    return Complex<A>(sqr(a.re) - sqr(a.im), 2.0 * a.re * a.im);
}
template <typename A>
TINYAD_HD TINYAD_INLINE Complex<A> conj(const Complex<A>& a) { return Complex<A>(a.re, -a.im); }
template <typename A>
TINYAD_HD TINYAD_INLINE A abs(const Complex<A>& a) { return hypot(a.re, a.im); }
template <typename A>
TINYAD_HD TINYAD_INLINE A arg(const Complex<A>& a) { return atan2(a.im, a.re); }

}  // namespace TinyAD
