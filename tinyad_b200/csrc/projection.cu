// tinyad_b200 runtime -- batched PSD projection (Utils/HessianProjection.hh:23-101).  Compiled once per group of K
// (-DTAD_PROJ_PART=0..5) so that the heavy instantiations build in parallel; the run-time dispatch lives in runtime.cu.
#include "rt_common.cuh"

#ifndef TAD_PROJ_PART
#define TAD_PROJ_PART 0
#endif

namespace tadrt
{
// ---------------------------------------------------------------------------------------------
// PSD projection (Utils/HessianProjection.hh:23-101), one thread per element.
//
// Symmetric eigensolver: Householder tridiagonalisation + implicit QL with accumulated
// transformations (the classic EISPACK tred2/tql2 scheme; Eigen's SelfAdjointEigenSolver is the QR
// flavour of the same method, any backward-stable variant gives the same projected matrix to
// O(eps |H|)).  Data placement: the K x K work matrix and the tridiagonal live in SHARED memory,
// laid out [entry][lane] so a warp's accesses are conflict-free 256-byte rows; per-thread local
// arrays would spill to L2/DRAM (the first version did: 10 GB of DRAM writes per launch).
//   * rotations are generated with one rsqrt instead of hypot + two divisions;
//   * a QL sweep keeps the running column in registers, so each rotation reads and writes one
//     column of V instead of two;
//   * H is rebuilt as H + sum_j (clamp(l_j) - l_j) v_j v_j^T over the clamped eigenpairs only,
//     which leaves H bit-unchanged when nothing is clamped (HessianProjection.hh:94-95).
// ---------------------------------------------------------------------------------------------
template <int K>
struct ProjSmem
{
    static constexpr int doubles_per_warp = (K * K + 2 * K) * 32;
};

// Full eigendecomposition of one element; S = this thread's column of the [entry][lane] shared-memory block.
template <int K>
__device__ void project_full_one(double* __restrict__ hp, int64_t stride, double eps, double* S, const int SS, unsigned long long* counts,
                                 bool count_decomposed)
{
    constexpr int H = K * (K + 1) / 2;
#define PV(i, j) S[((i) * K + (j)) * SS]
#define PD(i) S[(K * K + (i)) * SS]
#define PE(i) S[(K * K + K + (i)) * SS]
    // ---- load, and early-out 1: positive diagonally dominant (HessianProjection.hh:23-42, :62-63) ----
    {
        double offsum[K], diag[K];
#pragma unroll
        for (int i = 0; i < K; ++i) offsum[i] = 0.0;
#pragma unroll
        for (int s = 0; s < H; ++s)
        {
            constexpr int dummy = 0;
            (void)dummy;
            const int r = hess_seq_rc(K, s).row, c = hess_seq_rc(K, s).col;
            const double v = hp[(int64_t)s * stride];
            PV(r, c) = v;
            if (r != c)
            {
                PV(c, r) = v;
                offsum[r] += fabs(v);
                offsum[c] += fabs(v);
            }
            else
                diag[r] = v;
        }
        bool dominant = true;
#pragma unroll
        for (int i = 0; i < K; ++i)
            if (diag[i] < offsum[i] + eps) dominant = false;
        if (dominant) return;
    }

    // ---- tridiagonalise; V holds the symmetric matrix on entry, the orthogonal transformation on exit ----
    for (int j = 0; j < K; ++j) PD(j) = PV(K - 1, j);
    for (int i = K - 1; i > 0; --i)
    {
        double scale = 0.0, h = 0.0;
        for (int q = 0; q < i; ++q) scale += fabs(PD(q));
        if (scale == 0.0)
        {
            PE(i) = PD(i - 1);
            for (int j = 0; j < i; ++j)
            {
                PD(j) = PV(i - 1, j);
                PV(i, j) = 0.0;
                PV(j, i) = 0.0;
            }
        }
        else
        {
            const double inv_scale = 1.0 / scale;
            for (int q = 0; q < i; ++q)
            {
                const double t = PD(q) * inv_scale;
                PD(q) = t;
                h += t * t;
            }
            double f = PD(i - 1);
            double g = sqrt(h);
            if (f > 0) g = -g;
            PE(i) = scale * g;
            h -= f * g;
            PD(i - 1) = f - g;
            for (int j = 0; j < i; ++j) PE(j) = 0.0;
            for (int j = 0; j < i; ++j)
            {
                f = PD(j);
                PV(j, i) = f;
                g = PE(j) + PV(j, j) * f;
                for (int q = j + 1; q <= i - 1; ++q)
                {
                    const double vqj = PV(q, j);
                    g += vqj * PD(q);
                    PE(q) += vqj * f;
                }
                PE(j) = g;
            }
            f = 0.0;
            const double inv_h = 1.0 / h;
            for (int j = 0; j < i; ++j)
            {
                const double t = PE(j) * inv_h;
                PE(j) = t;
                f += t * PD(j);
            }
            const double hh = f / (h + h);
            for (int j = 0; j < i; ++j) PE(j) -= hh * PD(j);
            for (int j = 0; j < i; ++j)
            {
                f = PD(j);
                g = PE(j);
                for (int q = j; q <= i - 1; ++q) PV(q, j) -= (f * PE(q) + g * PD(q));
                PD(j) = PV(i - 1, j);
                PV(i, j) = 0.0;
            }
        }
        PD(i) = h;
    }
    for (int i = 0; i < K - 1; ++i)
    {
        PV(K - 1, i) = PV(i, i);
        PV(i, i) = 1.0;
        const double h = PD(i + 1);
        if (h != 0.0)
        {
            const double inv_h = 1.0 / h;
            for (int q = 0; q <= i; ++q) PD(q) = PV(q, i + 1) * inv_h;
            for (int j = 0; j <= i; ++j)
            {
                double g = 0.0;
                for (int q = 0; q <= i; ++q) g += PV(q, i + 1) * PV(q, j);
                for (int q = 0; q <= i; ++q) PV(q, j) -= g * PD(q);
            }
        }
        for (int q = 0; q <= i; ++q) PV(q, i + 1) = 0.0;
    }
    for (int j = 0; j < K; ++j)
    {
        PD(j) = PV(K - 1, j);
        PV(K - 1, j) = 0.0;
    }
    PV(K - 1, K - 1) = 1.0;
    PE(0) = 0.0;
    // ---- implicit QL on the tridiagonal (d, e), rotations accumulated into V ----
    for (int i = 1; i < K; ++i) PE(i - 1) = PE(i);
    PE(K - 1) = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double meps = 2.220446049250313e-16;
    for (int l = 0; l < K; ++l)
    {
        tst1 = fmax(tst1, fabs(PD(l)) + fabs(PE(l)));
        int m = l;
        while (m < K - 1)
        {
            if (fabs(PE(m)) <= meps * tst1) break;
            ++m;
        }
        if (m > l)
        {
            int iter = 0;
            double el_abs;
            do
            {
                ++iter;
                const double e_l = PE(l);
                double g = PD(l);
                double p = (PD(l + 1) - g) / (2.0 * e_l);
                double r = (fabs(p) < 1e150) ? sqrt(fma(p, p, 1.0)) : fabs(p);
                if (p < 0) r = -r;
                const double dl = e_l / (p + r);
                const double dl1 = e_l * (p + r);
                PD(l) = dl;
                PD(l + 1) = dl1;
                double h = g - dl;
                for (int i = l + 2; i < K; ++i) PD(i) -= h;
                f += h;
                p = PD(m);
                double c = 1.0, c2 = 1.0, c3 = 1.0;
                const double el1 = PE(l + 1);
                double s = 0.0, s2 = 0.0;
                double x[K];  // running column (column i+1 of V while the sweep moves down)
#pragma unroll
                for (int q = 0; q < K; ++q) x[q] = PV(q, m);
                for (int i = m - 1; i >= l; --i)
                {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    const double ei = PE(i), di = PD(i);
                    g = c * ei;
                    h = c * p;
                    const double t = fma(p, p, ei * ei);
                    double rinv;
                    if (t > 1e-280 && t < 1e280)
                    {
                        rinv = rsqrt(t);
                        r = t * rinv;
                    }
                    else
                    {
                        r = hypot(p, ei);
                        rinv = 1.0 / r;
                    }
                    PE(i + 1) = s * r;
                    s = ei * rinv;
                    c = p * rinv;
                    p = c * di - s * g;
                    PD(i + 1) = h + s * (c * g + s * di);
#pragma unroll
                    for (int q = 0; q < K; ++q)
                    {
                        const double y = PV(q, i);
                        PV(q, i + 1) = s * y + c * x[q];
                        x[q] = c * y - s * x[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < K; ++q) PV(q, l) = x[q];
                p = -s * s2 * c3 * el1 * e_l / dl1;
                PE(l) = s * p;
                PD(l) = c * p;
                el_abs = fabs(s * p);
            } while (el_abs > meps * tst1 && iter < 60);
        }
        PD(l) = PD(l) + f;
        PE(l) = 0.0;
    }

    // ---- clamp (HessianProjection.hh:71-91) and rebuild from the clamped eigenpairs only ----
    if (counts && count_decomposed) atomicAdd(&counts[0], 1ull);
    double acc[H];
#pragma unroll
    for (int s = 0; s < H; ++s) acc[s] = 0.0;
    bool all_positive = true;
    for (int j = 0; j < K; ++j)
    {
        const double lam = PD(j);
        double target = lam;
        if (eps < 0) { if (lam < 0) target = -lam; }
        else if (lam < eps) target = eps;
        if (target != lam)
        {
            all_positive = false;
            const double delta = target - lam;
            double v[K], dv[K];
#pragma unroll
            for (int q = 0; q < K; ++q)
            {
                v[q] = PV(q, j);
                dv[q] = delta * v[q];
            }
#pragma unroll
            for (int s = 0; s < H; ++s) acc[s] = fma(dv[hess_seq_rc(K, s).row], v[hess_seq_rc(K, s).col], acc[s]);
        }
    }
    // early out 2: nothing clamped -> H stays bit-unchanged (:94-95)
    if (all_positive) return;
    if (counts) atomicAdd(&counts[1], 1ull);
#pragma unroll
    for (int s = 0; s < H; ++s) hp[(int64_t)s * stride] += acc[s];
#undef PV
#undef PD
#undef PE
}

template <int K>
__global__ void __launch_bounds__(32) project_kernel_full(double* __restrict__ hess, int64_t n, int64_t stride, double eps,
                                                          unsigned long long* counts)
{
    extern __shared__ double proj_smem[];
    const int64_t el = (int64_t)blockIdx.x * 32 + threadIdx.x;
    if (el >= n) return;
    project_full_one<K>(hess + el, stride, eps, proj_smem + threadIdx.x, 32, counts, true);
}

// The elements the fast path could not finish (counts[2] of them: a few per million on the tet workloads): a small fixed
// grid of single-warp blocks strides over the list.  The cost of this launch is the serial latency of one full solve, so the
// work matrix is kept in STATIC shared memory when it fits the 48 KB static limit (K <= 12; no opt-in / carve-out switch),
// else in a global scratch buffer laid out [entry][thread].
template <int K>
__global__ void __launch_bounds__(kListThreads) project_kernel_list(double* __restrict__ hess, int64_t stride, double eps,
                                                                    unsigned long long* counts, const int64_t* __restrict__ list,
                                                                    double* __restrict__ work)
{
    constexpr bool use_smem = (size_t)(K * K + 2 * K) * kListThreads * sizeof(double) <= 48 * 1024;
    __shared__ double sm[use_smem ? (K * K + 2 * K) * kListThreads : 1];
    const int64_t count = (int64_t)counts[2];
    const int nthreads = gridDim.x * blockDim.x;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = tid; i < count; i += nthreads)
    {
        if (use_smem) project_full_one<K>(hess + list[i], stride, eps, sm + threadIdx.x, kListThreads, counts, false);
        else project_full_one<K>(hess + list[i], stride, eps, work + tid, nthreads, counts, false);
    }
}

// Fast path (Detail/Projection.hh), three kernels with different resource profiles, one thread per element:
//   A  tridiagonalise  -- the packed matrix in registers, fully unrolled (register-heavy, ILP-rich)
//   B  select vectors  -- eigenvalues of T, eigenvectors of the moved eigenvalues by inverse iteration
//                         (scalar recurrences on small arrays: few registers, runs at high occupancy)
//   C  apply           -- back-transform through the reflectors, H += low-rank term (register-heavy)
// Scratch between them is structure-of-arrays over the elements (coalesced 256-byte rows).

// D: variable dimension of the term when its K = D * N local variables are ordered handle by handle (translation null-space test), else 0
// REDUCE: elements whose Hessian has the translation null space continue in the reduced pipeline (K - D; code bit PROJ_REDUCED_BIT)
template <int K, int D, bool REDUCE = false>
__global__ void __launch_bounds__(128) project_kernel_a(const double* __restrict__ hess, int64_t n, int64_t stride, double eps, ProjScratch sc)
{
    const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (el >= n) return;
    const double* hp = hess + el;
    double* rp = sc.R + el;
    sc.codes[el] = TinyAD::detail::proj_tridiagonalize<K, D, REDUCE>([&](int s) { return hp[(int64_t)s * stride]; },
                                                          [&](int i, double v) { rp[(int64_t)i * stride] = v; }, eps);
}

// Which elements a phase-B kernel works on.  filter < 0: all decomposed ones; 0 / 1: those of the general / of the reduced pipeline
// (code bit PROJ_REDUCED_BIT, set by phase A).
__device__ __forceinline__ bool proj_skip(int code, int filter)
{
    if ((code & 15) == TinyAD::detail::PROJ_DOMINANT || (code & 15) == TinyAD::detail::PROJ_FALLBACK) return true;
    if (filter < 0) return false;
    return ((code & TinyAD::detail::PROJ_REDUCED_BIT) != 0) != (filter == 1);
}

// B1: eigenvalues of T (register-resident QL, Detail/Projection.hh proj_eigenvalues); few registers, high occupancy
template <int K>
__global__ void __launch_bounds__(128) project_kernel_b1(int64_t n, int64_t stride, ProjScratch sc, int filter)
{
    const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (el >= n) return;
    if (proj_skip(sc.codes[el], filter)) return;
    double* rp = sc.R + el;
    const int code = TinyAD::detail::proj_eigenvalues<K>([&](int i) { return rp[(int64_t)i * stride]; },
                                                         [&](int i, double v) { rp[(int64_t)i * stride] = v; });
    if (code == TinyAD::detail::PROJ_FALLBACK) sc.codes[el] = code;
}

// B2: selection + inverse iteration.  The sorted eigenvalues and the first B2_SMEM_VECS eigenvectors of every thread live in
// shared memory ([slot][thread], conflict-free): they are re-read at run-time indices / by every later vector's two
// orthogonalisation passes, and as global re-reads (L2 round trips of data the thread has just written) those loads were
// ~25 % of the kernel's stall samples.
constexpr int B2_SMEM_VECS = 4;
template <int K>
constexpr size_t b2_smem_bytes(int threads) { return (size_t)(K + B2_SMEM_VECS * K) * threads * sizeof(double); }

template <int K, int MINB>
__global__ void __launch_bounds__(128, MINB) project_kernel_b(int64_t n, int64_t stride, double eps, unsigned long long* counts, ProjScratch sc,
                                                              int filter)
{
    using L = TinyAD::detail::ProjLayout<K>;
    extern __shared__ double b2_smem[];
    const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (el >= n) return;
    int code = sc.codes[el];
    if ((code & 15) == TinyAD::detail::PROJ_DOMINANT) return;
    if (filter >= 0 && (code & 15) != TinyAD::detail::PROJ_FALLBACK && ((code & TinyAD::detail::PROJ_REDUCED_BIT) != 0) != (filter == 1)) return;
    if (filter == 1 && (code & 15) == TinyAD::detail::PROJ_FALLBACK) return;   // failed in the reduced B1: counted by the general B2 launch
    code &= 15;
    double* rp = sc.R + el;
    double* wp = sc.W + el;
    const int bd = blockDim.x;
    double* sl = b2_smem + threadIdx.x;        // eigenvalues: slot i
    double* sv = sl + (size_t)K * bd;          // vectors: slot jv * K + q, jv < B2_SMEM_VECS
    if (code != TinyAD::detail::PROJ_FALLBACK)
        code = TinyAD::detail::proj_select_vectors<K>(
            [&](int i) { return rp[(int64_t)i * stride]; }, [&](int i) { return sl[i * bd]; }, [&](int i, double v) { sl[i * bd] = v; },
            [&](int i, double v) { wp[(int64_t)i * stride] = v; },
            [&](int jv, const double (&v)[K]) {
                double* p = wp + (int64_t)(L::off_vec + jv * K) * stride;
#pragma unroll
                for (int q = 0; q < K; ++q) { *p = v[q]; p += stride; }
                if (jv < B2_SMEM_VECS)
                {
                    double* q0 = sv + (jv * K) * bd;
#pragma unroll
                    for (int q = 0; q < K; ++q) q0[q * bd] = v[q];
                }
            },
            [&](int jv, double (&v)[K]) {
                if (jv < B2_SMEM_VECS)
                {
                    const double* q0 = sv + (jv * K) * bd;
#pragma unroll
                    for (int q = 0; q < K; ++q) v[q] = q0[q * bd];
                }
                else
                {
                    const double* p = wp + (int64_t)(L::off_vec + jv * K) * stride;
#pragma unroll
                    for (int q = 0; q < K; ++q) { v[q] = *p; p += stride; }
                }
            },
            eps);
    if (filter == 1)
    {
        if (code == TinyAD::detail::PROJ_FALLBACK)
        {
            sc.codes[el] = code;   // counted and listed by the general launch that follows (the full solver works on H itself)
            return;
        }
        // reduced pipeline: the matrix on the complement of the translations.  Nothing to move there still leaves the translations'
        // own eigenvalue 0 < eps: an update with no vectors.  The bit stays on rebuilt elements only (phase C picks the layout by it).
        if (code == TinyAD::detail::PROJ_UNCHANGED && eps > 0.0)
        {
            wp[0] = 0.0;
            wp[stride] = 0.0;
            code = TinyAD::detail::PROJ_REBUILT;
        }
        sc.codes[el] = code == TinyAD::detail::PROJ_REBUILT ? (code | TinyAD::detail::PROJ_REDUCED_BIT) : code;
    }
    else
        sc.codes[el] = code;
    if (counts) atomicAdd(&counts[0], 1ull);
    if (code == TinyAD::detail::PROJ_REBUILT && counts) atomicAdd(&counts[1], 1ull);
    if (code == TinyAD::detail::PROJ_FALLBACK)
    {
        const unsigned long long slot = atomicAdd(&counts[2], 1ull);  // per slab (reset by the caller): index into the list
        sc.list[slot] = el;
        atomicAdd(&counts[3], 1ull);                                  // running total of the evaluation
    }
}

template <int K>
__global__ void __launch_bounds__(128) project_kernel_c(double* __restrict__ hess, int64_t n, int64_t stride, double eps, ProjScratch sc)
{
    const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (el >= n) return;
    if (sc.codes[el] != TinyAD::detail::PROJ_REBUILT) return;   // (the reduced pipeline only exists on the fused path)
    double* hp = hess + el;
    const double* rp = sc.R + el;
    const double* wp = sc.W + el;
    TinyAD::detail::proj_apply<K>([&](int i) { return rp[(int64_t)i * stride]; }, [&](int i) { return wp[(int64_t)i * stride]; },
                                  [&](int s) { return hp[(int64_t)s * stride]; }, [&](int s, double v) { hp[(int64_t)s * stride] = v; }, eps);
}

template <int K>
size_t project_scratch_doubles(int64_t stride)
{
    using L = TinyAD::detail::ProjLayout<K>;
    return (size_t)(L::nR + L::nW) * (size_t)stride + (size_t)(K * K + 2 * K) * kListBlocks * kListThreads;
}

// Phase B2.  2 blocks per SM at 255 registers (no spills) beat 3 blocks at 168 registers with ~400 B of spills by 3-6 %
// (tools/proj_bench.cu); K <= 12: 128-thread blocks (61 KB of shared memory each at K = 12); larger K: smaller blocks keep the
// footprint per SM.
template <int K>
int launch_b2(int64_t n, int64_t stride, double eps, unsigned long long* counts, ProjScratch sc, int filter, cudaStream_t st)
{
    static PerDeviceOnce b2_configured;
    bool config_ok = true;
    b2_configured.run([&] {
        config_ok = cudaFuncSetAttribute(project_kernel_b<K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b2_smem_bytes<K>(128)) == cudaSuccess &&
                    cudaFuncSetAttribute(project_kernel_b<K, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) == cudaSuccess;
    });
    if (!config_ok) return TAD_CUDA_ERROR;
    const int bt = K <= 12 ? 128 : 64;
    project_kernel_b<K, 2><<<(unsigned)((n + bt - 1) / bt), bt, b2_smem_bytes<K>(bt), st>>>(n, stride, eps, counts, sc, filter);
    return TAD_OK;
}

// counts: device uint64[4] = {#decomposed, #rebuilt, #full solver, unused}.  scratch_d: project_scratch_doubles<K>(stride)
// doubles, scratch_i: stride int32 + n int64 (see project_scratch_bytes).
template <int K>
int launch_project(double* hess, int64_t n, int64_t stride, double eps, unsigned long long* counts, double* scratch_d, int32_t* codes,
                   int64_t* list, bool full_only, ProjScratch* fuse_out, const ProjSide* side, cudaStream_t st, int tdim)
{
    using L = TinyAD::detail::ProjLayout<K>;
    constexpr size_t smem = (size_t)ProjSmem<K>::doubles_per_warp * sizeof(double);
    static PerDeviceOnce configured;
    bool config_ok = true;
    configured.run([&] { config_ok = cudaFuncSetAttribute(project_kernel_full<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; });
    if (!config_ok) return fail(TAD_CUDA_ERROR, "cannot configure shared memory of the projection kernel");
    if (full_only)
    {
        count_launch();
        project_kernel_full<K><<<(unsigned)((n + 31) / 32), 32, smem, st>>>(hess, n, stride, eps, counts);
    }
    else
    {
        ProjScratch sc;
        sc.R = scratch_d;
        sc.W = scratch_d + (size_t)L::nR * stride;
        sc.codes = codes;
        sc.list = list;
        const unsigned g = (unsigned)((n + 127) / 128);
        // tdim: variable dimension of the term (translation null-space test of phase A), | 64: reduced pipeline wanted -- for
        // four handles (K = 4 tdim) on the fused path: elements with the translation null space run phases B1 / B2 / C on K - tdim
        const bool want_reduce = (tdim & 64) != 0 && fuse_out != nullptr;
        tdim &= 63;
        bool reduced = false;
        if constexpr (K == 12 || K == 8)
        {
            constexpr int D = K / 4;
            if (want_reduce && tdim == D)
            {
                reduced = true;
                count_launch(6);   // A, B1 and B2 of both pipelines, the full-solver list kernel
                project_kernel_a<K, D, true><<<g, 128, 0, st>>>(hess, n, stride, eps, sc);
                project_kernel_b1<K - D><<<g, 128, 0, st>>>(n, stride, sc, 1);
                project_kernel_b1<K><<<g, 128, 0, st>>>(n, stride, sc, 0);
                if (launch_b2<K - D>(n, stride, eps, counts, sc, 1, st) != TAD_OK || launch_b2<K>(n, stride, eps, counts, sc, 0, st) != TAD_OK)
                    return fail(TAD_CUDA_ERROR, "cannot configure shared memory of the projection kernel B2");
            }
        }
        if (!reduced)
        {
            count_launch(4 + (fuse_out ? 0 : 1));
            // translation null-space test for the element shapes of the fused path (d <= 3 variables per handle, <= 4 handles)
            if constexpr (K % 3 == 0 && K / 3 >= 2 && K / 3 <= 4)
            {
                if (tdim == 3) project_kernel_a<K, 3><<<g, 128, 0, st>>>(hess, n, stride, eps, sc);
            }
            if constexpr (K % 2 == 0 && K / 2 >= 2 && K / 2 <= 4)
            {
                if (tdim == 2) project_kernel_a<K, 2><<<g, 128, 0, st>>>(hess, n, stride, eps, sc);
            }
            if (!((tdim == 3 && K % 3 == 0 && K / 3 >= 2 && K / 3 <= 4) || (tdim == 2 && K % 2 == 0 && K / 2 >= 2 && K / 2 <= 4)))
                project_kernel_a<K, 0><<<g, 128, 0, st>>>(hess, n, stride, eps, sc);
            project_kernel_b1<K><<<g, 128, 0, st>>>(n, stride, sc, -1);
            if (launch_b2<K>(n, stride, eps, counts, sc, -1, st) != TAD_OK)
                return fail(TAD_CUDA_ERROR, "cannot configure shared memory of the projection kernel B2");
        }
        // elements whose inverse iteration did not converge (code PROJ_FALLBACK, listed in `list`): full eigensolver
        double* work = scratch_d + (size_t)(L::nR + L::nW) * stride;
        if (fuse_out && side && side->stream)
        {
            // fused path: on the side stream, next to the phase C / assembly kernel (which skips the listed elements;
            // the caller assembles them after ev_list)
            if (cudaEventRecord(side->ev_b, st) != cudaSuccess || cudaStreamWaitEvent(side->stream, side->ev_b, 0) != cudaSuccess)
                return fail(TAD_CUDA_ERROR, "projection side stream");
            project_kernel_list<K><<<kListBlocks, kListThreads, 0, side->stream>>>(hess, stride, eps, counts, list, work);
            if (cudaEventRecord(side->ev_list, side->stream) != cudaSuccess) return fail(TAD_CUDA_ERROR, "projection side stream");
        }
        else
            project_kernel_list<K><<<kListBlocks, kListThreads, 0, st>>>(hess, stride, eps, counts, list, work);
        sc.reduced = reduced ? 1 : 0;
        if (fuse_out) *fuse_out = sc;  // phase C is fused with the assembly by the caller
        else project_kernel_c<K><<<g, 128, 0, st>>>(hess, n, stride, eps, sc);
    }
    return cudaGetLastError() == cudaSuccess ? TAD_OK : fail(TAD_CUDA_ERROR, "project kernel launch failed");
}


#if TAD_PROJ_PART == 0
// Any k <= 32 without a dedicated instantiation (the Hessian projection is instantiated for k in {1..10, 12, 15, 16, 18}; the
// reference projects any k, Utils/HessianProjection.hh:48-101): cyclic Jacobi with run-time k, the k x k work matrices A and V of a
// thread in a global scratch laid out [entry][thread] (coalesced).  Correct for every k, not tuned -- the sizes the benchmarks and
// the usual elements use (triangles, tets, edges, quads, one-rings) have the register-resident path above.
constexpr int kGenericBlocks = 148 * 2, kGenericThreads = 128;
__global__ void __launch_bounds__(kGenericThreads) project_kernel_generic(int k, double* __restrict__ hess, int64_t n, int64_t stride, double eps,
                                                                          unsigned long long* counts, double* __restrict__ work)
{
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int h = hess_size(k);
    double* A = work + tid;                              // A(i, j) = A[(i * k + j) * nthreads]
    double* V = work + (int64_t)k * k * nthreads + tid;  // V(i, j) likewise
#define GA(i, j) A[(int64_t)((i) * k + (j)) * nthreads]
#define GV(i, j) V[(int64_t)((i) * k + (j)) * nthreads]
    for (int64_t el = tid; el < n; el += nthreads)
    {
        double* hp = hess + el;
        // early-out 1: positive diagonally dominant (HessianProjection.hh:23-42)
        bool dominant = true;
        for (int i = 0; i < k && dominant; ++i)
        {
            double off = 0.0;
            for (int j = 0; j < k; ++j)
                if (j != i) off += fabs(hp[(int64_t)hess_seq_index(k, i, j) * stride]);
            if (hp[(int64_t)hess_seq_index(k, i, i) * stride] < off + eps) dominant = false;
        }
        if (dominant) continue;
        double nrm = 0.0;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j)
            {
                const double v = hp[(int64_t)hess_seq_index(k, i, j) * stride];
                GA(i, j) = v;
                GV(i, j) = i == j ? 1.0 : 0.0;
                nrm += v * v;
            }
        if (!(nrm == nrm) || nrm > 1e300) continue;     // NaN / Inf: the caller's finite check reports it
        for (int sweep = 0; sweep < 60; ++sweep)
        {
            double off = 0.0;
            for (int p = 0; p < k; ++p)
                for (int q = p + 1; q < k; ++q) { const double v = GA(p, q); off += v * v; }
            if (off <= 1e-32 * nrm) break;
            for (int p = 0; p < k - 1; ++p)
                for (int q = p + 1; q < k; ++q)
                {
                    const double apq = GA(p, q);
                    if (apq == 0.0) continue;
                    const double theta = (GA(q, q) - GA(p, p)) / (2.0 * apq);
                    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                    for (int r = 0; r < k; ++r)
                    {
                        const double arp = GA(r, p), arq = GA(r, q);
                        GA(r, p) = c * arp - sn * arq;
                        GA(r, q) = sn * arp + c * arq;
                    }
                    for (int r = 0; r < k; ++r)
                    {
                        const double apr = GA(p, r), aqr = GA(q, r);
                        GA(p, r) = c * apr - sn * aqr;
                        GA(q, r) = sn * apr + c * aqr;
                    }
                    for (int r = 0; r < k; ++r)
                    {
                        const double vrp = GV(r, p), vrq = GV(r, q);
                        GV(r, p) = c * vrp - sn * vrq;
                        GV(r, q) = sn * vrp + c * vrq;
                    }
                }
        }
        if (counts) atomicAdd(&counts[0], 1ull);
        // H += sum over the moved eigenpairs of (target - l) v v^T: leaves H bit-unchanged when nothing moves (:94-95)
        bool moved = false;
        for (int j = 0; j < k; ++j)
        {
            const double lam = GA(j, j);
            const bool mv = eps < 0.0 ? lam < 0.0 : lam < eps;
            if (!mv) continue;
            moved = true;
            const double w = eps < 0.0 ? -2.0 * lam : eps - lam;
            for (int s = 0; s < h; ++s)
            {
                const TinyAD::detail::HessRC rc = hess_seq_rc(k, s);
                hp[(int64_t)s * stride] = fma(w * GV(rc.row, j), GV(rc.col, j), hp[(int64_t)s * stride]);
            }
        }
        if (moved && counts) atomicAdd(&counts[1], 1ull);
    }
#undef GA
#undef GV
}

size_t project_scratch_doubles_generic(int k) { return (size_t)2 * k * k * kGenericBlocks * kGenericThreads; }

int launch_project_generic(int k, double* hess, int64_t n, int64_t stride, double eps, unsigned long long* counts, double* scratch_d, cudaStream_t st)
{
    if (k < 1 || k > 32) return fail(TAD_NOT_SUPPORTED, "Hessian projection supports at most 32 variables per element");
    count_launch();
    project_kernel_generic<<<kGenericBlocks, kGenericThreads, 0, st>>>(k, hess, n, stride, eps, counts, scratch_d);
    return cudaGetLastError() == cudaSuccess ? TAD_OK : fail(TAD_CUDA_ERROR, "project kernel launch failed");
}
#endif  // TAD_PROJ_PART == 0

#define TAD_INST(K)                                                                                                                         \
    template size_t project_scratch_doubles<K>(int64_t);                                                                                    \
    template int launch_project<K>(double*, int64_t, int64_t, double, unsigned long long*, double*, int32_t*, int64_t*, bool, ProjScratch*, \
                                   const ProjSide*, cudaStream_t, int);
#if TAD_PROJ_PART == 0
TAD_INST(1) TAD_INST(2) TAD_INST(3) TAD_INST(4) TAD_INST(5) TAD_INST(6)
#elif TAD_PROJ_PART == 1
TAD_INST(7) TAD_INST(8) TAD_INST(9) TAD_INST(10)
#elif TAD_PROJ_PART == 2
TAD_INST(12)
#elif TAD_PROJ_PART == 3
TAD_INST(15)
#elif TAD_PROJ_PART == 4
TAD_INST(16)
#else
TAD_INST(18)
#endif
#undef TAD_INST

}  // namespace tadrt
