// tinyad_b200 -- canonical packed layout of the symmetric k x k element Hessian.
//
// The reference stores a FULL k x k Eigen matrix per scalar (Scalar.hh:1344-1346).  Here
// only the lower triangle (h = k(k+1)/2 entries) is kept, in "tile order": the index range
// is cut into tiles of size t (3 if 3|k, else 2 if 2|k, else 1 -- i.e. one tile per vertex
// for the usual d = 2,3), tiles are visited row-major over the lower block triangle and
// entries row-major inside a tile.  The order depends on k only, so the prebuilt runtime
// (projection / assembly kernels) and the user-TU element kernels agree on it.
//
// A contiguous cut of that sequence into NP parts gives each of the NP cooperating threads
// of an element a near-square set of entries (few distinct row/col gradients needed).
#pragma once

// the unrolled bodies are generic lambdas; they must be inlined for the sparsity masks to fold
#define TINYAD_LAMBDA_INLINE __attribute__((always_inline))
#if defined(__CUDACC__)
#define TINYAD_HD __host__ __device__
#define TINYAD_INLINE __forceinline__
#else
#define TINYAD_HD
#define TINYAD_INLINE inline __attribute__((always_inline))
#endif

namespace TinyAD
{
namespace detail
{

TINYAD_HD constexpr int hess_tile(int k) { return (k % 3 == 0) ? 3 : ((k % 2 == 0) ? 2 : 1); }
TINYAD_HD constexpr int hess_size(int k) { return k * (k + 1) / 2; }

struct HessRC { int row, col; };

// (row, col), row >= col, of the s-th entry in tile order.
TINYAD_HD constexpr HessRC hess_seq_rc(int k, int s)
{
    const int t = hess_tile(k);
    const int nb = k / t;
    int pos = 0;
    for (int bi = 0; bi < nb; ++bi)
        for (int bj = 0; bj <= bi; ++bj)
        {
            const int cnt = (bi == bj) ? t * (t + 1) / 2 : t * t;
            if (s < pos + cnt)
            {
                const int l = s - pos;
                if (bi == bj)
                {
                    int r = 0;
                    while ((r + 1) * (r + 2) / 2 <= l) ++r;
                    return HessRC{bi * t + r, bj * t + (l - r * (r + 1) / 2)};
                }
                return HessRC{bi * t + l / t, bj * t + l % t};
            }
            pos += cnt;
        }
    return HessRC{-1, -1};
}

// Inverse map: sequence index of entry (i, j) (any order of i, j).
TINYAD_HD constexpr int hess_seq_index(int k, int i, int j)
{
    if (i < j) { const int tmp = i; i = j; j = tmp; }
    const int t = hess_tile(k);
    const int bi = i / t, bj = j / t;
    const int ri = i % t, rj = j % t;
    const int diag_cnt = t * (t + 1) / 2;
    // tiles before block row bi: sum_{b<bi} (b*t*t + diag_cnt)
    int pos = (bi * (bi - 1) / 2) * t * t + bi * diag_cnt;
    pos += bj * t * t;  // full tiles before (bi, bj) in this block row (bj <= bi)
    if (bi == bj) return pos + ri * (ri + 1) / 2 + rj;
    return pos + ri * t + rj;
}

// Contiguous cut of [0, h) into NP parts.
TINYAD_HD constexpr int hess_part_begin(int k, int np, int p)
{
    const int h = hess_size(k);
    return (int)(((long long)h * p) / np);
}

}  // namespace detail
}  // namespace TinyAD
