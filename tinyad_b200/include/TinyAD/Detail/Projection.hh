// tinyad_b200 -- per-element projection of a packed symmetric K x K Hessian to a positive-definite
// matrix: project_positive_definite of the reference (Utils/HessianProjection.hh:23-101), i.e.
//     H  ->  V max(L, eps) V^T        (eps >= 0)        or        V |L| V^T        (eps < 0)
// with both early-outs (diagonally dominant; nothing clamped) leaving H bit-unchanged.
//
// The reference calls Eigen::SelfAdjointEigenSolver (full eigendecomposition) and rebuilds
// V D V^T.  On the GPU the K x K eigenvector matrix is what hurts: it does not fit in registers
// (K = 12: 288 registers) and the QL/QR rotations index its columns at run time, so it has to
// live in shared memory, where rotating it costs ~25 KB of traffic per element.  This
// implementation never forms it:
//   1. Householder tridiagonalisation A = Q T Q^T with the packed matrix in REGISTERS (fully
//      unrolled, compile-time indices); the reflectors stay in the registers A occupied;
//   2. eigenvalues of T by implicit QL without vectors (EISPACK tql1 scheme) -- O(K^2);
//   3. the projected matrix differs from H by a low-rank term,
//          H + sum_{l_j < eps} (eps - l_j) v_j v_j^T       (or the complementary sum if that is shorter),
//      so only the eigenvectors of the clamped (or the kept) eigenvalues are needed: inverse
//      iteration on T with re-orthogonalisation inside clusters (the LAPACK dstein scheme),
//      back-transformed through the reflectors -- O(K^2) per vector.
// Any backward-stable eigensolver gives the same projected matrix to O(eps_machine |H|), which is
// what the parity tests check (tolerance 1e-10 relative, BASELINE.json north_star).
// The function is __host__ __device__ so that the host unit test exercises exactly the device code.
#pragma once

#include <cmath>
#include <cstdint>

#include <TinyAD/Scalar.hh>

namespace TinyAD
{
namespace detail
{

// 1/sqrt(t) for t in the normal range, without the special-case branches of rsqrt(): hardware approximation plus
// two Newton steps (host: exact).  Only used on data scaled to O(1).
TINYAD_HD TINYAD_INLINE double rsqrt_fast(double t)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
    const double ht = 0.5 * t;
    y = y * fma(-ht * y, y, 1.5);
    y = y * fma(-ht * y, y, 1.5);
    return y;
#else
    return 1.0 / sqrt(t);
#endif
}

enum ProjectCode
{
    PROJ_DOMINANT = 0,   // early-out 1, H untouched
    PROJ_UNCHANGED = 1,  // decomposed, nothing clamped, H untouched
    PROJ_REBUILT = 2,
    PROJ_FALLBACK = 3    // inverse iteration did not converge: caller must use the full eigensolver (H untouched)
};

// Scratch layout between the three phases (structure-of-arrays over elements in the kernels):
//   R[0..K)            d_i = T(i,i)
//   R[K..2K-1)         e_i = T(i+1,i)
//   R[2K-1..3K-3)      tau_k of reflector H_k = I - tau_k v_k v_k^T, k < K-2
//   R[3K-3..)          v_k(2:) for k = 0..K-3, concatenated (v_k(1) = 1 is implicit); then max |H_ij|; then the K eigenvalues
//   W[0] = number of vectors, W[1] = form (0: H + sum, 1: eps I + sum, 2: -H + sum), W[2..2+MAXV) weights,
//   W[2+MAXV + jv*K + i] = component i of eigenvector jv of T
template <int K>
struct ProjLayout
{
    static constexpr int H = K * (K + 1) / 2;
    static constexpr int MAXV = K / 2 + 1;
    static constexpr int n_refl = K > 2 ? K - 2 : 0;
    static constexpr int off_d = 0, off_e = K, off_tau = 2 * K - 1, off_v = off_tau + n_refl;
    static constexpr int n_v = n_refl * (n_refl + 1) / 2;  // sum_{k} (K-k-2)
    static constexpr int off_amax = off_v + n_v;  // max |H_ij| of the element (scale of the accuracy target)
    static constexpr int off_lam = off_amax + 1;  // eigenvalues of T, ascending (written by proj_eigenvalues)
    static constexpr int off_tnull = off_lam + K; // D if the D translations (1,0,..,1,0,..), ... are null vectors of H, else 0 (proj_tridiagonalize<K, D>)
    static constexpr int nR = off_tnull + 1;
    static constexpr int off_wgt = 2, off_vec = 2 + MAXV;
    static constexpr int nW = off_vec + MAXV * K;
    // offset of v_k(2 + i), i < K-k-2
    TINYAD_HD static constexpr int v_index(int k, int i)
    {
        int o = off_v;
        for (int q = 0; q < k; ++q) o += K - q - 2;
        return o + i;
    }
};

// Row p of the 4 x 4 Hadamard matrix (+-1/2 each): row 0 = (1,1,1,1)/2 spans the translation of four handles, rows 1..3 its complement
TINYAD_HD constexpr int hadamard_sign(int p, int s)
{
    // p = 1: + + - -,  p = 2: + - + -,  p = 3: + - - +
    return p == 0 ? 1 : (p == 1 ? (s < 2 ? 1 : -1) : (p == 2 ? ((s % 2 == 0) ? 1 : -1) : ((s == 0 || s == 3) ? 1 : -1)));
}

// Householder tridiagonalisation of the packed symmetric K x K matrix a (tile order, in registers, overwritten) and the stores of
// phase A: d, e, taus, reflectors, max|H| into R (ProjLayout<K>).
template <int K, class StoreRFn>
TINYAD_HD TINYAD_INLINE void proj_tridiag_core(double (&a)[K * (K + 1) / 2], const double amax, StoreRFn&& store_r)
{
    using L = ProjLayout<K>;
#define TAD_A(i, j) a[hess_seq_index(K, (i), (j))]
    double d0[K], e0[K];  // tridiagonal T: d0[i] = T(i,i), e0[i] = T(i+1,i); e0[K-1] = 0
    double tau[K > 2 ? K - 2 : 1];

    if constexpr (K == 1)
    {
        d0[0] = a[0];
        e0[0] = 0.0;
    }
    else
    {
        // ---- 1. Householder tridiagonalisation of the lower triangle, column by column ----
        // After step k: column k of A holds (d_k, e_k, v_k(2:)) with reflector H_k = I - tau_k v_k v_k^T, v_k(1) = 1.
        static_for<(K > 2 ? K - 2 : 0)>([&](auto kc) TINYAD_LAMBDA_INLINE {
            constexpr int k = decltype(kc)::value;
            constexpr int n = K - k - 1;  // order of the trailing block, rows/cols k+1 .. K-1
            const double alpha = TAD_A(k + 1, k);
            double xnorm2 = 0.0;
            static_for<n - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
                constexpr int i = decltype(ic)::value;
                xnorm2 = fma(TAD_A(k + 2 + i, k), TAD_A(k + 2 + i, k), xnorm2);
            });
            double t = 0.0;
            if (xnorm2 > 0.0)
            {
                const double nrm = sqrt(fma(alpha, alpha, xnorm2));
                const double beta = alpha >= 0.0 ? -nrm : nrm;
                t = (beta - alpha) / beta;
                const double sc = 1.0 / (alpha - beta);
                static_for<n - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = decltype(ic)::value;
                    TAD_A(k + 2 + i, k) *= sc;
                });
                TAD_A(k + 1, k) = beta;
                // p = tau * A22 v,  v = (1, A(k+2.., k))
                double p[n];
                static_for<n>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = decltype(ic)::value;
                    double s = TAD_A(k + 1 + i, k + 1);  // v_0 = 1
                    static_for<n - 1>([&](auto jc) TINYAD_LAMBDA_INLINE {
                        constexpr int j = decltype(jc)::value + 1;
                        s = fma(TAD_A(k + 1 + i, k + 1 + j), TAD_A(k + 1 + j, k), s);
                    });
                    p[i] = t * s;
                });
                // w = p - (tau/2) (p^T v) v
                double pv = p[0];
                static_for<n - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = decltype(ic)::value + 1;
                    pv = fma(p[i], TAD_A(k + 1 + i, k), pv);
                });
                const double hk = -0.5 * t * pv;
                p[0] += hk;
                static_for<n - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = decltype(ic)::value + 1;
                    p[i] = fma(hk, TAD_A(k + 1 + i, k), p[i]);
                });
                // A22 -= v w^T + w v^T (lower triangle)
                static_for<n>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = decltype(ic)::value;
                    static_for<i + 1>([&](auto jc) TINYAD_LAMBDA_INLINE {
                        constexpr int j = decltype(jc)::value;
                        double vi = 1.0, vj = 1.0;
                        if constexpr (i > 0) vi = TAD_A(k + 1 + i, k);
                        if constexpr (j > 0) vj = TAD_A(k + 1 + j, k);
                        TAD_A(k + 1 + i, k + 1 + j) -= vi * p[j] + p[i] * vj;
                    });
                });
            }
            tau[k] = t;
        });
        static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            d0[i] = TAD_A(i, i);
            if constexpr (i + 1 < K) e0[i] = TAD_A(i + 1, i);
            else e0[i] = 0.0;
        });
    }

    store_r(L::off_amax, amax);
    static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
        constexpr int i = decltype(ic)::value;
        store_r(L::off_d + i, d0[i]);
        if constexpr (i + 1 < K) store_r(L::off_e + i, e0[i]);
    });
    static_for<L::n_refl>([&](auto kc) TINYAD_LAMBDA_INLINE {
        constexpr int k = decltype(kc)::value;
        store_r(L::off_tau + k, tau[k]);
        static_for<K - k - 2>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            store_r(L::v_index(k, i), TAD_A(k + 2 + i, k));
        });
    });
#undef TAD_A
}

// PROJ_* code bit: the element goes through the REDUCED pipeline (proj_tridiagonalize<K, D, true>): phases B1 / B2 / C on K - D
constexpr int PROJ_REDUCED_BIT = 16;

// Phase A: early-out 1 and Householder tridiagonalisation; the packed matrix lives in registers.
// Returns PROJ_DOMINANT (nothing stored) or PROJ_UNCHANGED (R stored, continue with phase B).
// D > 0 (variable dimension of a term whose K = D * N local variables are ordered handle by handle): also tests whether the D
// TRANSLATIONS t_a (1 on component a of every handle) are null vectors of H, |H t_a|_inf <= 1e-13 max|H|, and stores the answer.
// Every translation-invariant element energy has them (the tet / triangle deformation energies: 3 of the ~4.5 eigenpairs phase B2
// would compute per tet); phase B2 then skips them and phase C adds their exactly known term, eps * (projector onto the translations).
// REDUCE (needs D > 0 and four handles, K = 4 D): when the translations are null vectors, H is mapped to the (K - D) x (K - D)
// matrix of its action on their orthogonal complement -- for four handles a 2-D Hadamard transform over the 4 x 4 grid of D x D
// blocks, additions and one scale -- and THAT matrix is tridiagonalised (stored in ProjLayout<K - D> order; the returned code
// carries PROJ_REDUCED_BIT).  Phases B1 / B2 then run on K - D = 9 instead of 12 (QL ~ K^2) without the null cluster, and phase C
// (proj_apply<K, K - D>) maps the eigenvectors back.  Elements that fail the translation test take the general path.
template <int K, int D = 0, bool REDUCE = false, class LoadFn, class StoreRFn>
TINYAD_HD inline int proj_tridiagonalize(LoadFn&& load, StoreRFn&& store_r, const double eps)
{
    using L = ProjLayout<K>;
    constexpr int H = L::H;
    double a[H];
    double amax = 0.0;
    static_for<H>([&](auto sc) TINYAD_LAMBDA_INLINE {
        constexpr int s = decltype(sc)::value;
        a[s] = load(s);
        amax = fmax(amax, fabs(a[s]));
    });
#define TAD_A(i, j) a[hess_seq_index(K, (i), (j))]

    // ---- early-out 1: positive diagonally dominant (HessianProjection.hh:23-42, :62-63) ----
    {
        double offsum[K];
        static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE { offsum[decltype(ic)::value] = 0.0; });
        static_for<H>([&](auto sc) TINYAD_LAMBDA_INLINE {
            constexpr int s = decltype(sc)::value;
            constexpr int r = hess_seq_rc(K, s).row, c = hess_seq_rc(K, s).col;
            if constexpr (r != c)
            {
                const double v = fabs(a[s]);
                offsum[r] += v;
                offsum[c] += v;
            }
        });
        bool dominant = true;
        static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            if (TAD_A(i, i) < offsum[i] + eps) dominant = false;
        });
        if (dominant) return PROJ_DOMINANT;
    }

    bool tnull = false;
    if constexpr (D > 0 && K % (D > 0 ? D : 1) == 0 && K > D)
    {
        // row sums over the handles, per component: S(i, a) = sum_s H(i, D s + a)
        double worst = 0.0;
        static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            static_for<D>([&](auto ac) TINYAD_LAMBDA_INLINE {
                constexpr int a_ = decltype(ac)::value;
                double sum = 0.0;
                static_for<K / D>([&](auto sc) TINYAD_LAMBDA_INLINE {
                    constexpr int j = D * decltype(sc)::value + a_;
                    sum += a[hess_seq_index(K, i, j)];
                });
                worst = fmax(worst, fabs(sum));
            });
        });
        tnull = worst <= 1e-13 * amax && amax > 0.0;
    }
    if constexpr (REDUCE && D > 0 && K == 4 * (D > 0 ? D : 1))
    {
        if (tnull)
        {
            // H~_pq = 1/4 sum_s sum_t sg(p, s) sg(q, t) B_st for the Hadamard rows p, q = 1..3 (row 0 = the translations): first over
            // s (G_pt = sum_s sg(p, s) B_st), then over t; only the lower block triangle p >= q of the symmetric result is formed
            constexpr int KR = K - D;
            double ar[KR * (KR + 1) / 2];
            static_for<3>([&](auto pc) TINYAD_LAMBDA_INLINE {
                constexpr int p = decltype(pc)::value + 1;
                static_for<D>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = decltype(ic)::value;
                    double g[4][D];   // G_pt(i, j), t = 0..3
                    static_for<4>([&](auto tc) TINYAD_LAMBDA_INLINE {
                        constexpr int t = decltype(tc)::value;
                        static_for<D>([&](auto jc) TINYAD_LAMBDA_INLINE {
                            constexpr int j = decltype(jc)::value;
                            double sum = 0.0;
                            static_for<4>([&](auto sc) TINYAD_LAMBDA_INLINE {
                                constexpr int s_ = decltype(sc)::value;
                                const double v = a[hess_seq_index(K, s_ * D + i, t * D + j)];
                                if constexpr (hadamard_sign(p, s_) > 0) sum += v; else sum -= v;
                            });
                            g[t][j] = sum;
                        });
                    });
                    static_for<p>([&](auto qc) TINYAD_LAMBDA_INLINE {
                        constexpr int q = decltype(qc)::value + 1;   // q <= p
                        static_for<D>([&](auto jc) TINYAD_LAMBDA_INLINE {
                            constexpr int j = decltype(jc)::value;
                            if constexpr (q < p || j <= i)
                            {
                                double sum = 0.0;
                                static_for<4>([&](auto tc) TINYAD_LAMBDA_INLINE {
                                    constexpr int t = decltype(tc)::value;
                                    if constexpr (hadamard_sign(q, t) > 0) sum += g[t][j]; else sum -= g[t][j];
                                });
                                ar[hess_seq_index(KR, (p - 1) * D + i, (q - 1) * D + j)] = 0.25 * sum;
                            }
                        });
                    });
                });
            });
            // max|H| stays the scale of the accuracy targets (the transform is orthogonal)
            proj_tridiag_core<KR>(ar, amax, store_r);
            store_r(ProjLayout<KR>::off_tnull, 0.0);   // nothing left to deflate in the reduced matrix
            return PROJ_UNCHANGED | PROJ_REDUCED_BIT;
        }
        store_r(L::off_tnull, 0.0);
    }
    else
        store_r(L::off_tnull, tnull ? (double)D : 0.0);

#undef TAD_A
    proj_tridiag_core<K>(a, amax, store_r);
    return PROJ_UNCHANGED;
}

// Phase B1: eigenvalues of T by QL with explicit shifts and no vectors (EISPACK tql1 scheme), stored UNSORTED to
// R[off_lam..).  Shaped for the GPU, one thread per matrix:
//   * d / e live in registers and every array index is a compile-time constant: the unreduced block always starts at
//     position 0, because a converged eigenvalue is written out and the arrays are shifted down by one (zeros shift into e,
//     so the block end needs no separate bookkeeping);
//   * ONE run-time loop whose body is a single QL step (one sweep unrolled over its maximal range K-2 .. 0; the rotations at
//     and beyond the first negligible sub-diagonal entry m are predicated off) followed by the deflation test.  The body
//     is ~0.5 k instructions, so it stays in the instruction cache (a version unrolled over the eigenvalue index as well
//     spent 80 % of its time waiting for instruction fetches), and lanes of a warp that need different numbers of steps for
//     one eigenvalue do not wait for each other: each lane simply continues with its next eigenvalue;
//   * no divisions or square roots on the slow paths: reciprocal / reciprocal square root by hardware approximation plus
//     Newton steps (all arguments are in the normal range because T is scaled by 1/|T|_1 first).
// Deflation criterion: |e_i| <= macheps * max_i(|d_i| + |e_i|), i.e. absolute accuracy macheps |T| -- what the projection
// needs (the map l -> max(l, eps) is 1-Lipschitz).
// 32 fixed pseudo-random numbers in (-1, 1) (start vectors of the inverse iteration)
TINYAD_HD TINYAD_INLINE double start_table(unsigned i)
{
    static constexpr double t[32] = {0.4387, -0.7112, 0.2918, 0.9534, -0.1276, 0.6641, -0.8823, 0.3359, -0.5467, 0.7785, 0.1193,
                              -0.9341, 0.5872, -0.2654, 0.8126, -0.4498, 0.0715, 0.6267, -0.7931, 0.3642, -0.1889, 0.9078,
                              -0.6013, 0.2471, 0.7356, -0.3927, 0.5189, -0.8564, 0.1632, -0.6745, 0.8891, -0.0458};
    return t[i & 31u];
}

// 1/x for x in the normal range (callers guarantee it, see the uses): hardware approximation + two Newton steps,
// without the slow-path branches of an IEEE division.  x = 0 gives NaN, which the callers select away.
TINYAD_HD TINYAD_INLINE double rcp_fast(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
#else
    return 1.0 / x;
#endif
}

// (Two multi-matrix drivers of this iteration were measured and dropped, profiles/README.md: lanes that pick up the next matrix of a
// per-warp queue on demand, and static per-lane lists with register prefetch.  Both were slower than one matrix per thread -- the
// kernel is bound by the latency of its dependent FP64 chains, which many short-lived warps hide better than few long-lived ones --
// and the resumable-state form they needed cost 18 registers (80 -> 98) and 13 % of this kernel's time.  A third variant with d / e
// in shared memory and run-time loops (EISPACK tql1 as written: ~32 instead of ~70 instructions per rotation slot, 68 registers) took
// 2.7 ms instead of 0.54 ms: every rotation then waits for a shared-memory store -> load round trip of the previous one.)
template <int K, class LoadRFn, class StoreRFn>
TINYAD_HD inline int proj_eigenvalues(LoadRFn&& load_r, StoreRFn&& store_r)
{
    using L = ProjLayout<K>;
    constexpr double macheps = 2.220446049250313e-16;
    double lam[K], ee[K];
    double onenrm = 0.0;
    static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
        constexpr int i = decltype(ic)::value;
        lam[i] = load_r(L::off_d + i);
        if constexpr (i + 1 < K) ee[i] = load_r(L::off_e + i);
        else ee[i] = 0.0;
    });
    static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
        constexpr int i = decltype(ic)::value;
        double rowsum = fabs(lam[i]) + fabs(ee[i]);
        if constexpr (i > 0) rowsum += fabs(ee[i - 1]);
        onenrm = fmax(onenrm, rowsum);
    });
    if (!(onenrm == onenrm) || onenrm > 1e300) return PROJ_FALLBACK;  // NaN / Inf input: the caller's finite check reports it
    const double inv_nrm = onenrm > 0.0 ? 1.0 / onenrm : 0.0;
    double tn = 0.0;
    static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
        constexpr int i = decltype(ic)::value;
        lam[i] *= inv_nrm;
        ee[i] *= inv_nrm;
        tn = fmax(tn, fabs(lam[i]) + fabs(ee[i]));
    });
    const double thr = macheps * tn;
    double f = 0.0;
    int done = 0, steps = 0;
    while (done < K)
    {
        // m = first index with a negligible sub-diagonal entry (e[K-1] = 0 always)
        int m = K - 1;
        static_for<K - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = K - 2 - decltype(ic)::value;  // K-2 down to 0: the smallest index wins
            if (fabs(ee[i]) <= thr) m = i;
        });
        if constexpr (K > 1)
        {
            if (m > 0)
            {
                // one QL step on the block [0, m], shift from the leading 2 x 2 block
                if (++steps > 40 * K) return PROJ_FALLBACK;
                const double e_l = ee[0];
                double g = lam[0];
                double p = (lam[1] - g) * rcp_fast(2.0 * e_l);
                const double pp1 = fma(p, p, 1.0);
                double r = pp1 * rsqrt_fast(pp1);
                if (p < 0) r = -r;
                const double q = p + r;  // |q| >= 1
                const double qinv = rcp_fast(q);
                const double dl = e_l * qinv;
                lam[0] = dl;
                lam[1] = e_l * q;
                double h = g - dl;
                static_for<(K > 2 ? K - 2 : 0)>([&](auto ic) TINYAD_LAMBDA_INLINE { lam[2 + decltype(ic)::value] -= h; });
                f += h;
                p = lam[K - 1];
                static_for<(K > 2 ? K - 2 : 0)>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = 1 + decltype(ic)::value;  // 1 .. K-2
                    if (m == i) p = lam[i];
                });
                double c = 1.0, c3 = 1.0;  // c3: c after rotation 2 (1 if it does not run)
                const double el1 = ee[1];
                double s = 0.0, s2 = 0.0;  // s2: s after rotation 1 (0 if it does not run)
                static_for<K - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = K - 2 - decltype(ic)::value;  // K-2 down to 0
                    if (i < m)
                    {
                        const double ei = ee[i], di = lam[i];
                        g = c * ei;
                        h = c * p;
                        const double t = fma(p, p, ei * ei);
                        const double rinv = rsqrt_fast(t);
                        ee[i + 1] = s * (t * rinv);
                        s = ei * rinv;
                        c = p * rinv;
                        p = c * di - s * g;
                        lam[i + 1] = h + s * (c * g + s * di);
                        if constexpr (i == 2) c3 = c;
                        if constexpr (i == 1) s2 = s;
                    }
                });
                p = -s * s2 * c3 * el1 * qinv;  // = -s s2 c3 el1 e_l / lam[1]
                ee[0] = s * p;
                lam[0] = c * p;
            }
        }
        if (fabs(ee[0]) <= thr)
        {
            // deflate: eigenvalue found; shift the arrays so that the remaining block starts at 0 again
            store_r(L::off_lam + done, (lam[0] + f) * onenrm);
            ++done;
            static_for<K - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
                constexpr int i = decltype(ic)::value;
                lam[i] = lam[i + 1];
                ee[i] = ee[i + 1];
            });
            ee[K - 1] = 0.0;
        }
    }
    return PROJ_UNCHANGED;
}

// Ascending sort of K doubles held in registers: compare-exchange network with compile-time indices (an optimal
// 39-comparator network for K = 12, odd-even transposition otherwise).  No NaNs (checked by the caller).
TINYAD_HD TINYAD_INLINE void cswap(double& a, double& b)
{
    const bool sw = b < a;
    const double lo = sw ? b : a, hi = sw ? a : b;
    a = lo;
    b = hi;
}
template <int K>
TINYAD_HD TINYAD_INLINE void sort_ascending(double (&v)[K])
{
    if constexpr (K == 12)
    {
        constexpr int net[39][2] = {{0, 8},  {1, 7},  {2, 6},  {3, 11}, {4, 10}, {5, 9},  {0, 1},  {2, 5},  {3, 4},  {6, 9},
                                    {7, 8},  {10, 11}, {0, 2},  {1, 6},  {5, 10}, {9, 11}, {0, 3},  {1, 2},  {4, 6},  {5, 7},
                                    {8, 11}, {9, 10}, {1, 4},  {3, 5},  {6, 8},  {7, 10}, {1, 3},  {2, 5},  {6, 9},  {8, 10},
                                    {2, 3},  {4, 5},  {6, 7},  {8, 9},  {4, 6},  {5, 7},  {3, 4},  {5, 6},  {7, 8}};
        static_for<39>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int c = decltype(ic)::value;
            cswap(v[net[c][0]], v[net[c][1]]);
        });
    }
    else
    {
        static_for<K>([&](auto rc) TINYAD_LAMBDA_INLINE {
            constexpr int round = decltype(rc)::value;
            static_for<(K - (round & 1)) / 2>([&](auto ic) TINYAD_LAMBDA_INLINE {
                constexpr int i = (round & 1) + 2 * decltype(ic)::value;
                cswap(v[i], v[i + 1]);
            });
        });
    }
}

// Dot product / sums of K entries with FOUR independent accumulators: phase B2 runs at two warps per scheduler and is bound by
// the latency of dependent FP64 instructions, so a 12-long serial fma chain costs ~100 cycles where four chains of 3 plus a
// 2-level combine cost ~40 (ncu: the serial dot products of the orthogonalisation passes were 12 % of the kernel's stall samples).
template <int K>
TINYAD_HD TINYAD_INLINE double dot4(const double (&a)[K], const double (&b)[K])
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
        constexpr int i = decltype(ic)::value;
        if constexpr (i % 4 == 0) s0 = fma(a[i], b[i], s0);
        else if constexpr (i % 4 == 1) s1 = fma(a[i], b[i], s1);
        else if constexpr (i % 4 == 2) s2 = fma(a[i], b[i], s2);
        else s3 = fma(a[i], b[i], s3);
    });
    return (s0 + s1) + (s2 + s3);
}
template <int K>
TINYAD_HD TINYAD_INLINE double abssum4(const double (&a)[K])
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
        constexpr int i = decltype(ic)::value;
        if constexpr (i % 4 == 0) s0 += fabs(a[i]);
        else if constexpr (i % 4 == 1) s1 += fabs(a[i]);
        else if constexpr (i % 4 == 2) s2 += fabs(a[i]);
        else s3 += fabs(a[i]);
    });
    return (s0 + s1) + (s2 + s3);
}

// Phase B2: selection of the eigenvalues that move, their eigenvectors (of T) by inverse iteration.
// Scalar recurrences on small arrays, written as plain loops: nvcc unrolls the fixed-trip-count ones (LU, solves)
// so those arrays live in registers.  load_w re-reads vectors already stored through store_w; the eigenvalues are read
// back from R at run-time indices.
// Functors: load_r(i) reads R; load_lam(i) / store_lam(i, v), i < K: the eigenvalues (read unsorted from R[off_lam + i] by the
// first load_lam calls, then rewritten sorted -- the kernel keeps the sorted copy in shared memory, they are read at run-time
// indices); store_w(i, v) writes W; store_vec(jv, v) writes the K components of vector jv (W[off_vec + jv K + q]) and load_vec(jv, v)
// reads vector jv back (the kernel serves the first few vectors, the ones re-read most often, from shared memory).
template <int K, class LoadRFn, class LoadLamFn, class StoreLamFn, class StoreWFn, class StoreVecFn, class LoadVecFn>
TINYAD_HD inline int proj_select_vectors(LoadRFn&& load_r, LoadLamFn&& load_lam, StoreLamFn&& store_lam, StoreWFn&& store_w,
                                         StoreVecFn&& store_vec, LoadVecFn&& load_vec, const double eps)
{
    using L = ProjLayout<K>;
    constexpr double macheps = 2.220446049250313e-16;
    double d0[K], e0[K];
    double onenrm = 0.0;
    for (int i = 0; i < K; ++i)
    {
        d0[i] = load_r(L::off_d + i);
        e0[i] = (i + 1 < K) ? load_r(L::off_e + i) : 0.0;
    }
    for (int i = 0; i < K; ++i)
    {
        const double rowsum = fabs(d0[i]) + fabs(e0[i]) + (i > 0 ? fabs(e0[i - 1]) : 0.0);
        onenrm = fmax(onenrm, rowsum);
    }
    double emax = 0.0, dmax = 0.0;
    for (int i = 0; i < K; ++i)
    {
        emax = fmax(emax, fabs(e0[i]));
        dmax = fmax(dmax, fabs(d0[i]));
    }
    auto lam_at = [&](int i) { return load_lam(i); };
    {
        // B1 leaves the eigenvalues in the order of convergence: sort ascending (odd-even transposition network on
        // registers), check them, write them back for the run-time indexed reads below
        double lam[K];
        double chk = 0.0;
        static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            lam[i] = load_r(L::off_lam + i);  // unsorted, from B1
            chk += fabs(lam[i]);
        });
        if (!(chk <= 1.7e308)) return PROJ_FALLBACK;  // NaN / Inf
        sort_ascending<K>(lam);
        static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; store_lam(i, lam[i]); });
    }

    // ---- 3. which eigenvalues move (HessianProjection.hh:71-91) ----
    const bool abs_mode = eps < 0.0;
    const double thresh = abs_mode ? 0.0 : eps;
    int r = 0;  // the eigenvalues are ascending
    for (int i = 0; i < K; ++i) r += (lam_at(i) < thresh) ? 1 : 0;
    if (r == 0) return PROJ_UNCHANGED;  // early-out 2 (:94-95)
    // form A: H + sum_{j<r} delta_j v_j v_j^T with delta_j = target_j - l_j           (r <= K/2)
    // form B: base + sum_{j>=r} gamma_j v_j v_j^T, base = eps I (clamp) or -H (abs)     (otherwise: fewer vectors)
    const bool form_b = 2 * r > K;
    // Translation null space (proj_tridiagonalize<K, D> found the D translations to be null vectors of H): if the moved eigenvalues
    // are exactly some clearly negative ones plus D eigenvalues that are zero to rounding, the latter ARE the translations -- no
    // inverse iteration for them; phase C adds eps * (projector onto the translations) (the reference adds (eps - l_j) v_j v_j^T
    // over an arbitrary orthonormal basis v_j of that eigenspace, l_j = O(macheps |H|)).
    int tn = (int)load_r(L::off_tnull);
    int n_neg = 0;
    if (tn > 0)
    {
        const double tolz = 1e-11 * load_r(L::off_amax);
        int n_tiny = 0;
        for (int i = 0; i < K; ++i)
        {
            const double li = lam_at(i);
            n_neg += (li < -tolz) ? 1 : 0;
            n_tiny += (fabs(li) <= tolz) ? 1 : 0;
        }
        // exactly D tiny eigenvalues, all moved eigenvalues accounted for, form A
        if (n_tiny != tn || r != n_neg + tn || form_b) tn = 0;
    }
    const int j_begin = form_b ? r : 0, j_end = form_b ? K : (tn > 0 ? n_neg : r);

    // ---- 4. inverse iteration on T (LAPACK dstein scheme: dlagtf / dlagts, re-orthogonalisation in clusters) ----
    const double ortol = 1e-3 * onenrm;
    const double amax = load_r(L::off_amax);
    const double res_limit = 2e-12 * amax;
    double xjm = 0.0;
    int gpind = j_begin;
    double la[K], lb[K], lc[K], ld[K];  // LU factors of T - x_j I (kept across the vectors of a cluster, see below)
    uint32_t pivmask = 0;
    double tol = 0.0, a_last = 0.0;
    for (int j = j_begin; j < j_end; ++j)
    {
        const int jv = j - j_begin;
        const double lj = lam_at(j);
        double xj = lj;
        if (j > j_begin)
        {
            const double pertol = 10.0 * fabs(macheps * xj);
            if (xj - xjm < pertol) xj = xjm + pertol;
        }
        // LU factorisation of T - xj I with partial pivoting (dlagtf); the pivot tests |a_k|/scale1 >= |c_k|/scale2
        // are cross-multiplied, and the pivots are inverted once (with the dlagts perturbation) for all solves.
        // Shifts that coincide to working precision (a multiple eigenvalue, e.g. the null space) share one factorisation.
        const bool reuse_lu = j > j_begin && fabs(xj - xjm) <= 4.0 * macheps * onenrm;
        if (!reuse_lu)
        {
            pivmask = 0;
            for (int i = 0; i < K; ++i)
            {
                la[i] = d0[i] - xj;
                lb[i] = e0[i];
                lc[i] = e0[i];
                ld[i] = 0.0;
            }
            // Branch-free elimination step: "keep" = no row interchange.  One reciprocal per step (hardware approximation
            // + Newton) instead of a division in each of two divergent branches.
            double scale1 = fabs(la[0]) + (K > 1 ? fabs(lb[0]) : 0.0);
            for (int k = 0; k < K - 1; ++k)
            {
                double scale2 = fabs(lc[k]) + fabs(la[k + 1]);
                if (k < K - 2) scale2 += fabs(lb[k + 1]);
                const double ak = la[k], ck = lc[k], bk = lb[k], a1 = la[k + 1];
                const bool czero = ck == 0.0;
                const bool keep = czero || (ak != 0.0 && fabs(ck) * scale1 <= fabs(ak) * scale2);
                const double mult = czero ? 0.0 : (keep ? ck : ak) * rcp_fast(keep ? ak : ck);
                la[k + 1] = fma(-mult, keep ? bk : a1, keep ? a1 : bk);
                la[k] = keep ? ak : ck;
                lb[k] = keep ? bk : a1;
                lc[k] = mult;
                if (k < K - 2)
                {
                    const double b1 = lb[k + 1];
                    ld[k] = keep ? 0.0 : b1;
                    lb[k + 1] = keep ? b1 : -mult * b1;
                }
                if (keep) scale1 = scale2;
                else pivmask |= (1u << k);
            }
            // dlagtf: tol = macheps * max(|a_i|, |b_i|, |d_i|) over the factors.  Their entries are bounded by twice the
            // largest entry of T - xj I (partial pivoting on a tridiagonal matrix), for which max|e| and max|d| + |xj| are
            // upper bounds; a tolerance that is a small factor larger only perturbs a negligible pivot a little more.
            tol = macheps * fmax(emax, dmax + fabs(xj));
            a_last = fabs(la[K - 1]);
            for (int i = 0; i < K; ++i)
            {
                double ak = la[i];
                if (fabs(ak) < tol) ak = (ak < 0.0) ? -tol : tol;
                la[i] = rcp_fast(ak);
            }
        }

        // weight of v v^T in the low-rank term, and distance of l_j to the nearest unselected eigenvalue
        double wj;
        if (!form_b) wj = abs_mode ? -2.0 * lj : eps - lj;
        else wj = abs_mode ? 2.0 * lj : lj - eps;
        double gap = 1e300;
        if (!form_b) { if (j_end < K) gap = lam_at(j_end) - lj; }
        else if (j_begin > 0) gap = lj - lam_at(j_begin - 1);

        // deterministic start vector with entries in (-1, 1): a fixed table of pseudo-random numbers, read at an offset that
        // depends on the vector (dstein draws new random numbers for every vector)
        double x[K];
        for (int i = 0; i < K; ++i) x[i] = start_table((unsigned)(i + 5 * jv));
        if (j > j_begin && fabs(xj - xjm) > ortol) gpind = j;
        // With the eigenvalue known to working precision one solve from a random start already gives a residual of a
        // few eps |T| (more solves do not improve it; for the later members of a cluster they make it worse, because
        // the re-orthogonalised iterate is fed back into a solve that amplifies the removed directions again).  So,
        // instead of dstein's growth test plus two extra iterations: solve, re-orthogonalise inside the cluster, and
        // accept as soon as the residual |(T - l_j) y|_inf of the normalised vector is below res_limit (at most 5 solves,
        // then the element goes to the full eigensolver).  Because l -> max(l, eps) and l -> |l| are 1-Lipschitz, an
        // eigenvector error only enters the projected matrix through this residual (the weight ratio
        // |w_j - w_k| / |l_j - l_k| is <= 2 for every pair), so res_limit = 2e-12 max|H_ij| keeps the projected
        // matrix within ~1e-11 max|H_ij| of the exact one -- an order below the 1e-10 parity tolerance.  (The last member of an
        // exactly degenerate cluster typically stalls at ~1e-13 |T|_1; T is not split into blocks here as dstein does.)
        int its = 0;
        bool converged = false;
        double inv = 0.0;
        while (its < 5)
        {
            ++its;
            // orthogonalise the right-hand side against the vectors of the cluster computed so far before the solve:
            // otherwise the solve amplifies those directions as much as the wanted one and the remainder after removing
            // them is noisy (residual ~100 eps |T| for the third vector of a triple eigenvalue).  (Vectors outside the
            // cluster are amplified >= 1e12 times less than the wanted one; they are removed after the solve only.) ...
            for (int i = gpind; i < j; ++i)
            {
                double v[K];
                load_vec(i - j_begin, v);
                const double dot = dot4<K>(x, v);
                for (int q = 0; q < K; ++q) x[q] = fma(-dot, v[q], x[q]);
            }
            const double xabs = abssum4<K>(x);
            const double scl = (double)K * onenrm * fmax(macheps, fmax(a_last, tol)) / fmax(xabs, 1e-300);
            // solve (dlagts): forward with L and the interchanges (scaling folded in), back substitution
            x[0] *= scl;
            for (int k = 1; k < K; ++k)
            {
                const double t1 = x[k] * scl;
                if (!((pivmask >> (k - 1)) & 1u)) x[k] = t1 - lc[k - 1] * x[k - 1];
                else
                {
                    const double t0 = x[k - 1];
                    x[k - 1] = t1;
                    x[k] = t0 - lc[k - 1] * t1;
                }
            }
            for (int k = K - 1; k >= 0; --k)
            {
                double temp = x[k];
                if (k <= K - 2) temp -= lb[k] * x[k + 1];
                if (k <= K - 3) temp -= ld[k] * x[k + 2];
                x[k] = temp * la[k];
            }
            // ... and against all selected vectors after it (modified Gram-Schmidt; not only the cluster: the other selected
            // eigenvectors are orthogonal anyway, and removing them confines what is left of the error to the
            // unselected subspace, see the acceptance test)
            for (int i = j_begin; i < j; ++i)
            {
                double v[K];
                load_vec(i - j_begin, v);
                const double dot = dot4<K>(x, v);
                for (int q = 0; q < K; ++q) x[q] = fma(-dot, v[q], x[q]);
            }
            const double n2 = dot4<K>(x, x);
            if (!(n2 > 0.0) || !(n2 < 1e300)) break;  // NaN, zero or overflow
            inv = 1.0 / sqrt(n2);
            double res = 0.0;
            for (int i = 0; i < K; ++i)
            {
                double t = (d0[i] - lj) * x[i];
                if (i > 0) t = fma(e0[i - 1], x[i - 1], t);
                if (i + 1 < K) t = fma(e0[i], x[i + 1], t);
                if (fabs(t) > res) res = fabs(t);
            }
#if defined(TAD_PROJ_DEBUG) && !defined(__CUDA_ARCH__)
            printf("  j=%d lj=%.3e xj=%.3e its=%d res=%.3e (limit %.3e) gpind=%d a_last=%.3e tol=%.3e\n", j, lj, xj, its, res * inv, res_limit, gpind, a_last, tol);
#endif
            // After the orthogonalisation the error of y lies in the unselected eigenspace, at distance >= gap from l_j,
            // and enters the projected matrix scaled by |w_j| / gap: a vector with a tiny weight (the null space of an
            // element Hessian has w_j ~ eps) may carry a much larger residual.
            if (res * inv <= res_limit || res * inv * fabs(wj) <= 1e-12 * amax * gap)
            {
                converged = true;
                break;
            }
        }
        if (!converged) return PROJ_FALLBACK;
        {
            double xn[K];
            for (int i = 0; i < K; ++i) xn[i] = x[i] * inv;
            store_vec(jv, xn);   // whole vector at once: the kernel forms the address of component 0 once and steps by the stride
        }
        store_w(L::off_wgt + jv, wj);
        xjm = xj;
    }
    store_w(0, (double)(j_end - j_begin));
    store_w(1, (!form_b ? 0.0 : (abs_mode ? 2.0 : 1.0)) + 8.0 * tn);   // form + 8 * (dimension of the deflated translation null space)
    return PROJ_REBUILT;
}

// Phase C: back-transform the selected eigenvectors through the reflectors and add the low-rank term to H.
// tmp_store(slot, v) / tmp_load(slot): per-thread scratch of MAXV * (K + 1) doubles for the back-transformed vectors and their
// weights between the two halves (the reflectors and the accumulator do not fit in registers together).  The kernels keep it
// in SHARED memory: as a run-time indexed local array it missed L1 92 % of the time and every vector iteration waited for L2.
// The global loads of vector jv + 1 are issued before vector jv is processed (software prefetch), for the same reason.
// preload(nv): optional asynchronous copy of the nv vectors and weights from W straight into the scratch (slot jv (K+1) + i <-
// W[off_vec + jv K + i], slot jv (K+1) + K <- W[off_wgt + jv]); returns true if it did so.  preload_wait() completes it.  The
// fused kernel uses cp.async here, so these loads overlap the reflector loads and cost no registers.
// KR < K (reduced pipeline, proj_tridiagonalize<K, D, true>, D = K - KR, four handles): R and W are in ProjLayout<KR> order, the
// eigenvectors are those of the KR x KR matrix on the complement of the translations; after the reflectors they are mapped back
// with the Hadamard rows 1..3, v[s D + i] = 1/2 sum_p sg(p, s) y[(p - 1) D + i], and the translations' own term eps * projector is
// added in closed form (form 0 only: the bases of the other forms contain it).  preload(nv) then fills slots jv (K + 1) + i, i < KR.
template <int K, int KR = K, class LoadRFn, class LoadWFn, class LoadFn, class StoreFn, class TmpStoreFn, class TmpLoadFn, class PreloadFn, class PreloadWaitFn>
TINYAD_HD inline void proj_apply(LoadRFn&& load_r, LoadWFn&& load_w, LoadFn&& load, StoreFn&& store, const double eps, TmpStoreFn&& tmp_store,
                                 TmpLoadFn&& tmp_load, PreloadFn&& preload, PreloadWaitFn&& preload_wait)
{
    using L = ProjLayout<KR>;          // layout of R / W: the matrix that was decomposed
    constexpr int H = K * (K + 1) / 2;
    constexpr int DR = K - KR;         // dimension of the translation space removed by the reduction (0: none)
    static_assert(KR == K || (DR > 0 && K == 4 * DR), "the reduced pipeline is for four handles");
    const int nv = (int)load_w(0);
    const int form_code = (int)load_w(1);
    const int form = form_code & 7, tn = (KR == K) ? (form_code >> 3) : ((form == 0) ? DR : 0);
    const bool preloaded = preload(nv);
    {
        // reflectors in registers; each vector goes v = H_0 H_1 ... H_{K-3} y
        double refl[L::n_v > 0 ? L::n_v : 1], tau[L::n_refl > 0 ? L::n_refl : 1];
        static_for<L::n_v>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; refl[i] = load_r(L::off_v + i); });
        static_for<L::n_refl>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; tau[i] = load_r(L::off_tau + i); });
        double ynext[KR], wnext = 0.0;
        if (preloaded) preload_wait();
        else if (nv > 0)
        {
            static_for<KR>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; ynext[i] = load_w(L::off_vec + i); });
            wnext = load_w(L::off_wgt);
        }
        for (int jv = 0; jv < nv; ++jv)
        {
            double y[KR];
            double wj = wnext;
            if (preloaded)
                static_for<KR>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; y[i] = tmp_load(jv * (K + 1) + i); });
            else
            {
                static_for<KR>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; y[i] = ynext[i]; });
                if (jv + 1 < nv)
                {
                    static_for<KR>([&](auto ic) TINYAD_LAMBDA_INLINE {
                        constexpr int i = decltype(ic)::value;
                        ynext[i] = load_w(L::off_vec + (jv + 1) * KR + i);
                    });
                    wnext = load_w(L::off_wgt + jv + 1);
                }
            }
            static_for<L::n_refl>([&](auto kc) TINYAD_LAMBDA_INLINE {
                constexpr int k = KR - 3 - decltype(kc)::value;
                constexpr int n = KR - k - 1;
                constexpr int vo = L::v_index(k, 0) - L::off_v;
                double s = y[k + 1];
                static_for<n - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = decltype(ic)::value;
                    s = fma(refl[vo + i], y[k + 2 + i], s);
                });
                s *= tau[k];
                y[k + 1] -= s;
                static_for<n - 1>([&](auto ic) TINYAD_LAMBDA_INLINE {
                    constexpr int i = decltype(ic)::value;
                    y[k + 2 + i] = fma(-s, refl[vo + i], y[k + 2 + i]);
                });
            });
            if constexpr (KR == K)
                static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; tmp_store(jv * (K + 1) + i, y[i]); });
            else
                static_for<4>([&](auto sc) TINYAD_LAMBDA_INLINE {   // back to the K variables of the four handles
                    constexpr int s_ = decltype(sc)::value;
                    static_for<DR>([&](auto ic) TINYAD_LAMBDA_INLINE {
                        constexpr int i = decltype(ic)::value;
                        double v = 0.0;
                        static_for<3>([&](auto pc) TINYAD_LAMBDA_INLINE {
                            constexpr int p = decltype(pc)::value + 1;
                            if constexpr (hadamard_sign(p, s_) > 0) v += y[(p - 1) * DR + i]; else v -= y[(p - 1) * DR + i];
                        });
                        tmp_store(jv * (K + 1) + s_ * DR + i, 0.5 * v);
                    });
                });
            if (!preloaded) tmp_store(jv * (K + 1) + K, wj);
        }
    }
    double acc[H];
    static_for<H>([&](auto sc) TINYAD_LAMBDA_INLINE {
        constexpr int s = decltype(sc)::value;
        constexpr int r_ = hess_seq_rc(K, s).row, c_ = hess_seq_rc(K, s).col;
        if (form == 0) acc[s] = load(s);
        else if (form == 2) acc[s] = -load(s);
        else acc[s] = (r_ == c_) ? eps : 0.0;
    });
    for (int jv = 0; jv < nv; ++jv)
    {
        const double wj = tmp_load(jv * (K + 1) + K);
        double v[K], wv[K];
        static_for<K>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            v[i] = tmp_load(jv * (K + 1) + i);
            wv[i] = wj * v[i];
        });
        static_for<H>([&](auto sc) TINYAD_LAMBDA_INLINE {
            constexpr int s = decltype(sc)::value;
            acc[s] = fma(wv[hess_seq_rc(K, s).row], v[hess_seq_rc(K, s).col], acc[s]);
        });
    }
    if (tn > 0 && eps > 0.0)
    {
        // + eps * (projector onto the tn translations of the K / tn handles): eps / N on the entries (i, j) with equal component
        const double w0 = eps * (double)tn / (double)K;
        static_for<H>([&](auto sc) TINYAD_LAMBDA_INLINE {
            constexpr int s = decltype(sc)::value;
            constexpr int r_ = hess_seq_rc(K, s).row, c_ = hess_seq_rc(K, s).col;
            bool same = false;
            if (tn == 3) same = (K % 3 == 0) && (r_ % 3 == c_ % 3);
            else if (tn == 2) same = (K % 2 == 0) && (r_ % 2 == c_ % 2);
            else same = true;
            if (same) acc[s] += w0;
        });
    }
    static_for<H>([&](auto sc) TINYAD_LAMBDA_INLINE { constexpr int s = decltype(sc)::value; store(s, acc[s]); });
}

// with a thread-local scratch array (host; kernels without shared memory)
template <int K, int KR = K, class LoadRFn, class LoadWFn, class LoadFn, class StoreFn>
TINYAD_HD inline void proj_apply(LoadRFn&& load_r, LoadWFn&& load_w, LoadFn&& load, StoreFn&& store, const double eps)
{
    double tmp[ProjLayout<K>::MAXV * (K + 1)];
    proj_apply<K, KR>(load_r, load_w, load, store, eps, [&](int i, double v) { tmp[i] = v; }, [&](int i) { return tmp[i]; },
                      [](int) { return false; }, [] {});
}

// All three phases on one element through local scratch (host tests; the kernels run the phases separately).
// D: variable dimension of the term (translation null-space deflation, see proj_tridiagonalize); 0 = none.
template <int K, int D = 0, class LoadFn, class StoreFn>
TINYAD_HD inline int project_element(LoadFn&& load, StoreFn&& store, const double eps)
{
    using L = ProjLayout<K>;
    double R[L::nR > 0 ? L::nR : 1], Wb[L::nW];
    int code = proj_tridiagonalize<K, D>(load, [&](int i, double v) { R[i] = v; }, eps);
    if (code == PROJ_DOMINANT) return code;
    code = proj_eigenvalues<K>([&](int i) { return R[i]; }, [&](int i, double v) { R[i] = v; });
    if (code == PROJ_FALLBACK) return code;
    code = proj_select_vectors<K>([&](int i) { return R[i]; }, [&](int i) { return R[L::off_lam + i]; }, [&](int i, double v) { R[L::off_lam + i] = v; },
                                  [&](int i, double v) { Wb[i] = v; }, [&](int jv, const double (&v)[K]) { for (int q = 0; q < K; ++q) Wb[L::off_vec + jv * K + q] = v[q]; },
                                  [&](int jv, double (&v)[K]) { for (int q = 0; q < K; ++q) v[q] = Wb[L::off_vec + jv * K + q]; }, eps);
    if (code != PROJ_REBUILT) return code;
    proj_apply<K>([&](int i) { return R[i]; }, [&](int i) { return Wb[i]; }, load, store, eps);
    return PROJ_REBUILT;
}


// The reduced pipeline on one element (host tests; the kernels run its phases separately): four handles of D variables each.
template <int K, int D, class LoadFn, class StoreFn>
TINYAD_HD inline int project_element_reduced(LoadFn&& load, StoreFn&& store, const double eps)
{
    static_assert(D > 0 && K == 4 * D, "four handles");
    constexpr int KR = K - D;
    using L = ProjLayout<K>;
    using LR = ProjLayout<KR>;
    double R[L::nR], Wb[L::nW];
    int code = proj_tridiagonalize<K, D, true>(load, [&](int i, double v) { R[i] = v; }, eps);
    if (code == PROJ_DOMINANT) return code;
    if (code & PROJ_REDUCED_BIT)
    {
        code = proj_eigenvalues<KR>([&](int i) { return R[i]; }, [&](int i, double v) { R[i] = v; });
        if (code == PROJ_FALLBACK) return code;
        code = proj_select_vectors<KR>([&](int i) { return R[i]; }, [&](int i) { return R[LR::off_lam + i]; }, [&](int i, double v) { R[LR::off_lam + i] = v; },
                                       [&](int i, double v) { Wb[i] = v; }, [&](int jv, const double (&v)[KR]) { for (int q = 0; q < KR; ++q) Wb[LR::off_vec + jv * KR + q] = v[q]; },
                                       [&](int jv, double (&v)[KR]) { for (int q = 0; q < KR; ++q) v[q] = Wb[LR::off_vec + jv * KR + q]; }, eps);
        if (code == PROJ_FALLBACK) return code;
        if (code == PROJ_UNCHANGED)
        {
            // nothing moves in the complement: only the translations' own eigenvalue 0 is below eps > 0
            if (!(eps > 0.0)) return PROJ_UNCHANGED;
            Wb[0] = 0.0;
            Wb[1] = 0.0;
        }
        proj_apply<K, KR>([&](int i) { return R[i]; }, [&](int i) { return Wb[i]; }, load, store, eps);
        return PROJ_REBUILT | PROJ_REDUCED_BIT;
    }
    code = proj_eigenvalues<K>([&](int i) { return R[i]; }, [&](int i, double v) { R[i] = v; });
    if (code == PROJ_FALLBACK) return code;
    code = proj_select_vectors<K>([&](int i) { return R[i]; }, [&](int i) { return R[L::off_lam + i]; }, [&](int i, double v) { R[L::off_lam + i] = v; },
                                  [&](int i, double v) { Wb[i] = v; }, [&](int jv, const double (&v)[K]) { for (int q = 0; q < K; ++q) Wb[L::off_vec + jv * K + q] = v[q]; },
                                  [&](int jv, double (&v)[K]) { for (int q = 0; q < K; ++q) v[q] = Wb[L::off_vec + jv * K + q]; }, eps);
    if (code != PROJ_REBUILT) return code;
    proj_apply<K>([&](int i) { return R[i]; }, [&](int i) { return Wb[i]; }, load, store, eps);
    return PROJ_REBUILT;
}

// Full eigendecomposition of one packed matrix by cyclic Jacobi rotations, all in thread-local arrays: the in-kernel fallback of
// the fused small-k element kernel (TinyAD/Kernels.cuh) for the few elements per million whose inverse iteration does not converge
// (code PROJ_FALLBACK of project_element).  Any backward-stable solver gives the same projected matrix to O(macheps |H|).  Like the
// fast path, H is rebuilt as H + sum_j (clamp(l_j) - l_j) v_j v_j^T over the moved eigenpairs only, which leaves H bit-unchanged
// when nothing moves (HessianProjection.hh:94-95).  Returns PROJ_UNCHANGED or PROJ_REBUILT.
template <int K, class LoadFn, class StoreFn>
TINYAD_HD inline int project_full_jacobi(LoadFn&& load, StoreFn&& store, const double eps)
{
    constexpr int H = K * (K + 1) / 2;
    double A[K][K], V[K][K];
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    double h[H];
    static_for<H>([&](auto sc) TINYAD_LAMBDA_INLINE {
        constexpr int s = decltype(sc)::value;
        constexpr int r = hess_seq_rc(K, s).row, c = hess_seq_rc(K, s).col;
        h[s] = load(s);
        A[r][c] = h[s];
        A[c][r] = h[s];
    });
    double nrm = 0.0;
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) nrm += A[i][j] * A[i][j];
    for (int sweep = 0; sweep < 60; ++sweep)
    {
        double off = 0.0;
        for (int p = 0; p < K; ++p)
            for (int q = p + 1; q < K; ++q) off += A[p][q] * A[p][q];
        if (off <= 1e-32 * nrm) break;
        for (int p = 0; p < K - 1; ++p)
            for (int q = p + 1; q < K; ++q)
            {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int r = 0; r < K; ++r)
                {
                    const double arp = A[r][p], arq = A[r][q];
                    A[r][p] = c * arp - sn * arq;
                    A[r][q] = sn * arp + c * arq;
                }
                for (int r = 0; r < K; ++r)
                {
                    const double apr = A[p][r], aqr = A[q][r];
                    A[p][r] = c * apr - sn * aqr;
                    A[q][r] = sn * apr + c * aqr;
                }
                for (int r = 0; r < K; ++r)
                {
                    const double vrp = V[r][p], vrq = V[r][q];
                    V[r][p] = c * vrp - sn * vrq;
                    V[r][q] = sn * vrp + c * vrq;
                }
            }
    }
    bool moved = false;
    for (int j = 0; j < K; ++j)
    {
        const double lam = A[j][j];
        double w = 0.0;
        if (eps < 0.0) { if (lam < 0.0) w = -2.0 * lam; }      // |l| - l
        else if (lam < eps) w = eps - lam;
        if ((eps < 0.0 && lam < 0.0) || (eps >= 0.0 && lam < eps))
        {
            moved = true;
            static_for<H>([&](auto sc) TINYAD_LAMBDA_INLINE {
                constexpr int s = decltype(sc)::value;
                constexpr int r = hess_seq_rc(K, s).row, c = hess_seq_rc(K, s).col;
                h[s] = fma(w * V[r][j], V[c][j], h[s]);
            });
        }
    }
    if (!moved) return PROJ_UNCHANGED;
    static_for<H>([&](auto sc) TINYAD_LAMBDA_INLINE { constexpr int s = decltype(sc)::value; store(s, h[s]); });
    return PROJ_REBUILT;
}

}  // namespace detail
}  // namespace TinyAD
