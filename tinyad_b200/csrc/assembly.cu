// tinyad_b200 runtime -- assembly into g and the fixed CSR values (replaces the serial loops of ScalarObjectiveTerm.hh:256-277):
// FP64 atomics (optionally fused with the last projection phase) or the deterministic gather.  Compiled once per variable
// dimension (-DTAD_ASM_PART=0,1,2 -> D = 1,2,3) so that the (D, N) instantiations build in parallel; part 0 also holds the
// run-time dispatch, the generic kernel and the gather kernels.
#include "rt_common.cuh"

#ifndef TAD_ASM_PART
#define TAD_ASM_PART 0
#endif

namespace tadrt
{
// ---------------------------------------------------------------------------------------------
// assembly, atomic mode: one thread per element, FP64 red.global.add on g and the CSR values
// ---------------------------------------------------------------------------------------------

// rec / blockbase / rstride: the term's maps, already offset to the first element of the slab, leading dimension mstride;
// grad / hess: the slab's staging, leading dimension stride; e = position inside the slab.
template <int D, int N>
__global__ void __launch_bounds__(128) assemble_atomic_kernel(const int32_t* __restrict__ rec, const int32_t* __restrict__ blockbase,
                                                              const int32_t* __restrict__ rstride, int64_t mstride,
                                                              const double* __restrict__ grad, const double* __restrict__ hess, int64_t n,
                                                              int64_t stride, double* __restrict__ g, double* __restrict__ Hv, int32_t* err)
{
    constexpr int K = D * N;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    bool finite = true;
#pragma unroll
    for (int bi = 0; bi < N; ++bi)
    {
        const int32_t vi = rec[(int64_t)bi * mstride + e];
        if (vi < 0) continue;
#pragma unroll
        for (int a = 0; a < D; ++a)
        {
            const double v = grad[(int64_t)(D * bi + a) * stride + e];
            finite = finite && isfinite(v);
            atomicAdd(&g[(int64_t)D * vi + a], v);
        }
    }
    if (hess)
    {
#pragma unroll
        for (int bi = 0; bi < N; ++bi)
        {
            const int32_t rs = rstride[(int64_t)bi * mstride + e];
#pragma unroll
            for (int bj = 0; bj < N; ++bj)
            {
                const int32_t base = blockbase[(int64_t)(bi * N + bj) * mstride + e];
                if (base < 0) continue;
#pragma unroll
                for (int a = 0; a < D; ++a)
#pragma unroll
                    for (int b = 0; b < D; ++b)
                    {
                        const int s = hess_seq_index(K, D * bi + a, D * bj + b);
                        const double v = hess[(int64_t)s * stride + e];
                        finite = finite && isfinite(v);
                        atomicAdd(&Hv[(int64_t)base + (int64_t)a * rs + b], v);
                    }
            }
        }
    }
    if (!finite) atomicOr(err, ERR_NONFINITE);
}

// Projection phase C fused with the atomic assembly: the projected Hessian of an element is formed in registers
// (low-rank update of H, Detail/Projection.hh proj_apply) and scattered from there, instead of being written back to
// the staging buffer and read again by the assembly kernel (saves 1.35 KB of HBM traffic per tet and one launch).
// REDUCED: the element went through the reduced pipeline (four handles; R / W in ProjLayout<K - D> order, proj_apply<K, K - D>)
template <int D, int N, bool REDUCED = false>
__device__ __forceinline__ void c_assemble_one(const int64_t e, const double* __restrict__ hess, int64_t stride, double eps, const ProjScratch& sc,
                                               const int32_t* __restrict__ rec, const int32_t* __restrict__ blockbase,
                                               const int32_t* __restrict__ rstride, const int64_t mstride, const double* __restrict__ grad,
                                               double* __restrict__ g, double* __restrict__ Hv, int32_t* err, const int code)
{
    constexpr int K = D * N;
    constexpr int H = K * (K + 1) / 2;
    const double* hp = hess + e;
    // The scatter map of the element (N handles, N row strides, N^2 block bases) is loaded FIRST: in the scatter phase below every
    // block was "load base -> test -> 9 REDs", and with ~210 registers taken by the accumulator the compiler did not hoist those
    // loads, so their latency was exposed 16 times per tet (ncu: 62 % of the stall samples were long-scoreboard waits on the
    // instructions consuming them).  Loaded here, their latency hides behind the projection phase C.
    int32_t m_vi[N], m_rs[N], m_base[N * N];
#pragma unroll
    for (int bi = 0; bi < N; ++bi)
    {
        m_vi[bi] = rec[(int64_t)bi * mstride + e];
        m_rs[bi] = rstride[(int64_t)bi * mstride + e];
#pragma unroll
        for (int bj = 0; bj < N; ++bj) m_base[bi * N + bj] = blockbase[(int64_t)(bi * N + bj) * mstride + e];
    }
    double acc[H];
    if ((code & 15) == TinyAD::detail::PROJ_REBUILT)
    {
        constexpr int KR = REDUCED ? K - D : K;
        const double* rp = sc.R + e;
        const double* wp = sc.W + e;
        // scratch of proj_apply in shared memory: [slot][thread of the block], conflict-free
        extern __shared__ double casm_tmp[];
        double* tp = casm_tmp + threadIdx.x;
        const int bd = blockDim.x;
        TinyAD::detail::proj_apply<K, KR>([&](int i) { return rp[(int64_t)i * stride]; }, [&](int i) { return wp[(int64_t)i * stride]; },
                                      [&](int s) { return hp[(int64_t)s * stride]; }, [&](int s, double v) { acc[s] = v; }, eps,
                                      [&](int i, double v) { tp[i * bd] = v; }, [&](int i) { return tp[i * bd]; },
                                      [&](int nv) {
                                          // asynchronous global -> shared copies of the nv vectors and their weights (cp.async, 8 bytes
                                          // each): no registers, and they overlap the reflector loads that follow
                                          using L = TinyAD::detail::ProjLayout<KR>;
                                          const unsigned dst0 = (unsigned)__cvta_generic_to_shared(tp);
                                          for (int jv = 0; jv < nv; ++jv)
                                          {
#pragma unroll
                                              for (int i = 0; i <= KR; ++i)
                                              {
                                                  // components i < KR to slots jv (K + 1) + i, the weight to slot jv (K + 1) + K
                                                  const double* src = wp + (int64_t)(i < KR ? L::off_vec + jv * KR + i : L::off_wgt + jv) * stride;
                                                  const unsigned dst = dst0 + (unsigned)((jv * (K + 1) + (i < KR ? i : K)) * bd) * 8u;
                                                  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
                                              }
                                          }
                                          return true;
                                      },
                                      [&] { asm volatile("cp.async.wait_all;" ::: "memory"); });
    }
    else
    {
#pragma unroll
        for (int s = 0; s < H; ++s) acc[s] = hp[(int64_t)s * stride];
    }
    bool finite = true;
#pragma unroll
    for (int bi = 0; bi < N; ++bi)
    {
        const int32_t vi = m_vi[bi];
        if (vi < 0) continue;
#pragma unroll
        for (int a = 0; a < D; ++a)
        {
            const double v = grad[(int64_t)(D * bi + a) * stride + e];
            finite = finite && isfinite(v);
            atomicAdd(&g[(int64_t)D * vi + a], v);
        }
    }
#pragma unroll
    for (int bi = 0; bi < N; ++bi)
    {
        const int32_t rs = m_rs[bi];
#pragma unroll
        for (int bj = 0; bj < N; ++bj)
        {
            const int32_t base = m_base[bi * N + bj];
            if (base < 0) continue;
#pragma unroll
            for (int a = 0; a < D; ++a)
#pragma unroll
                for (int b = 0; b < D; ++b)
                {
                    const double v = acc[hess_seq_index(K, D * bi + a, D * bj + b)];
                    finite = finite && isfinite(v);
                    atomicAdd(&Hv[(int64_t)base + (int64_t)a * rs + b], v);
                }
        }
    }
    if (!finite) atomicOr(err, ERR_NONFINITE);
}

// LIST = false: all elements except those handed to the full solver (code PROJ_FALLBACK) when `skip_listed`;
// LIST = true: the listed elements (their staged Hessian was projected in place by project_kernel_list).
// REDUCED (non-LIST only): the launch for the elements of the reduced pipeline (code bit PROJ_REDUCED_BIT); the plain launch skips them.
template <int D, int N, bool LIST, bool REDUCED = false>
__global__ void __launch_bounds__(128) project_c_assemble_kernel(const double* __restrict__ hess, int64_t n, int64_t stride, double eps,
                                                                 ProjScratch sc, const int32_t* __restrict__ rec,
                                                                 const int32_t* __restrict__ blockbase, const int32_t* __restrict__ rstride,
                                                                 int64_t mstride, const double* __restrict__ grad, double* __restrict__ g,
                                                                 double* __restrict__ Hv, int32_t* err, const unsigned long long* counts,
                                                                 bool skip_listed)
{
    if constexpr (LIST)
    {
        const int64_t count = (int64_t)counts[2];
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
            c_assemble_one<D, N>(sc.list[i], hess, stride, eps, sc, rec, blockbase, rstride, mstride, grad, g, Hv, err, TinyAD::detail::PROJ_FALLBACK);
    }
    else
    {
        const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= n) return;
        const int code = sc.codes[e];
        if (skip_listed && code == TinyAD::detail::PROJ_FALLBACK) return;
        if (((code & TinyAD::detail::PROJ_REDUCED_BIT) != 0) != REDUCED) return;
        c_assemble_one<D, N, REDUCED>(e, hess, stride, eps, sc, rec, blockbase, rstride, mstride, grad, g, Hv, err, code);
    }
}


template <int D, int N>
int launch_c_assemble(const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double eps, ProjScratch sc, double* g,
                      double* Hv, int32_t* err, const unsigned long long* counts, const ProjSide* side, cudaStream_t st)
{
    const bool split = side && side->stream;
    // ~200 registers per thread: single-warp blocks fit 10 per SM (10 warps) where 128-thread blocks fit 2 (8 warps); the kernel is
    // bound by the latency of its load phases, so the extra warps pay (C2: 1.18 -> 1.08 ms; capping the registers at 170 for 12 warps: 1.21 ms).
    // TAD_CASM_BLOCK overrides (tuning knob).
    static const int bs = [] { const char* e = getenv("TAD_CASM_BLOCK"); const int v = e ? atoi(e) : 32; return (v == 32 || v == 64 || v == 128) ? v : 32; }();
    constexpr int K = D * N;
    constexpr size_t tmp_thread = (size_t)TinyAD::detail::ProjLayout<K>::MAXV * (K + 1) * sizeof(double);  // 728 B at K = 12
    static PerDeviceOnce configured;
    bool ok = true;
    configured.run([&] {
        ok = cudaFuncSetAttribute(project_c_assemble_kernel<D, N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(128 * tmp_thread)) == cudaSuccess &&
             cudaFuncSetAttribute(project_c_assemble_kernel<D, N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(128 * tmp_thread)) == cudaSuccess &&
             cudaFuncSetAttribute(project_c_assemble_kernel<D, N, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) == cudaSuccess;
    });
    if (!ok) return fail(TAD_CUDA_ERROR, "cannot configure shared memory of the fused projection/assembly kernel");
    count_launch(split ? 2 : 1);
    if constexpr (N == 4)
    {
        if (sc.reduced)
        {
            // the elements of the reduced pipeline (normally all of them) first; the plain launch below skips them
            static PerDeviceOnce configured_r;
            configured_r.run([&] {
                ok = cudaFuncSetAttribute(project_c_assemble_kernel<D, N, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(128 * tmp_thread)) == cudaSuccess &&
                     cudaFuncSetAttribute(project_c_assemble_kernel<D, N, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) == cudaSuccess;
            });
            if (!ok) return fail(TAD_CUDA_ERROR, "cannot configure shared memory of the fused projection/assembly kernel");
            count_launch();
            project_c_assemble_kernel<D, N, false, true><<<(unsigned)((n + bs - 1) / bs), bs, bs * tmp_thread, st>>>(
                hess, n, stride, eps, sc, m.rec, m.blockbase, m.rstride, m.mstride, grad, g, Hv, err, counts, split);
        }
    }
    project_c_assemble_kernel<D, N, false><<<(unsigned)((n + bs - 1) / bs), bs, bs * tmp_thread, st>>>(
        hess, n, stride, eps, sc, m.rec, m.blockbase, m.rstride, m.mstride, grad, g, Hv, err, counts, split);
    if (split)
    {
        cudaStreamWaitEvent(st, side->ev_list, 0);
        project_c_assemble_kernel<D, N, true><<<8, 128, 128 * tmp_thread, st>>>(hess, n, stride, eps, sc, m.rec, m.blockbase, m.rstride, m.mstride,
                                                                                 grad, g, Hv, err, counts, false);
    }
    return TAD_OK;
}

template <int D, int N>
void launch_assemble(const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double* g, double* Hv, int32_t* err,
                     cudaStream_t st)
{
    assemble_atomic_kernel<D, N><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(m.rec, m.blockbase, m.rstride, m.mstride, grad, hess, n, stride, g, Hv, err);
}

#if TAD_ASM_PART == 0
bool fused_c_assemble_supported(int d, int N) { return d >= 1 && d <= 3 && N >= 1 && N <= 4; }

int c_assemble(int d, int N, const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double eps, ProjScratch sc,
               double* g, double* Hv, int32_t* err, const unsigned long long* counts, const ProjSide* side, cudaStream_t st)
{
    if (n <= 0) return TAD_OK;
    if (!fused_c_assemble_supported(d, N)) return fail(TAD_NOT_SUPPORTED, "no fused projection/assembly kernel for this (d, N)");
    int rc = TAD_OK;
    switch (d)
    {
    case 1: rc = c_assemble_d<1>(N, m, grad, hess, n, stride, eps, sc, g, Hv, err, counts, side, st); break;
    case 2: rc = c_assemble_d<2>(N, m, grad, hess, n, stride, eps, sc, g, Hv, err, counts, side, st); break;
    default: rc = c_assemble_d<3>(N, m, grad, hess, n, stride, eps, sc, g, Hv, err, counts, side, st); break;
    }
    if (rc != TAD_OK) return rc;
    return cudaGetLastError() == cudaSuccess ? TAD_OK : fail(TAD_CUDA_ERROR, "fused projection/assembly launch failed");
}
#endif

template <int D>
int c_assemble_d(int N, const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double eps, ProjScratch sc,
                 double* g, double* Hv, int32_t* err, const unsigned long long* counts, const ProjSide* side, cudaStream_t st)
{
    switch (N)
    {
    case 1: return launch_c_assemble<D, 1>(m, grad, hess, n, stride, eps, sc, g, Hv, err, counts, side, st);
    case 2: return launch_c_assemble<D, 2>(m, grad, hess, n, stride, eps, sc, g, Hv, err, counts, side, st);
    case 3: return launch_c_assemble<D, 3>(m, grad, hess, n, stride, eps, sc, g, Hv, err, counts, side, st);
    default: return launch_c_assemble<D, 4>(m, grad, hess, n, stride, eps, sc, g, Hv, err, counts, side, st);
    }
}

template <int D>
bool assemble_atomic_d(int N, const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double* g, double* Hv,
                       int32_t* err, cudaStream_t st)
{
    switch (N)
    {
    case 1: launch_assemble<D, 1>(m, grad, hess, n, stride, g, Hv, err, st); return true;
    case 2: launch_assemble<D, 2>(m, grad, hess, n, stride, g, Hv, err, st); return true;
    case 3: launch_assemble<D, 3>(m, grad, hess, n, stride, g, Hv, err, st); return true;
    case 4: launch_assemble<D, 4>(m, grad, hess, n, stride, g, Hv, err, st); return true;
    default: return false;
    }
}

template int c_assemble_d<TAD_ASM_PART + 1>(int, const SlabMaps&, const double*, const double*, int64_t, int64_t, double, ProjScratch, double*, double*,
                                             int32_t*, const unsigned long long*, const ProjSide*, cudaStream_t);
template bool assemble_atomic_d<TAD_ASM_PART + 1>(int, const SlabMaps&, const double*, const double*, int64_t, int64_t, double*, double*, int32_t*,
                                                   cudaStream_t);

#if TAD_ASM_PART == 0
// generic (runtime d, N) fallback
__global__ void __launch_bounds__(128) assemble_atomic_generic(int D, int N, SeqTable seq, const int32_t* __restrict__ rec,
                                                               const int32_t* __restrict__ blockbase, const int32_t* __restrict__ rstride,
                                                               int64_t mstride, const double* __restrict__ grad, const double* __restrict__ hess,
                                                               int64_t n, int64_t stride, double* __restrict__ g, double* __restrict__ Hv,
                                                               int32_t* err)
{
    const int K = D * N;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    bool finite = true;
    for (int bi = 0; bi < N; ++bi)
    {
        const int32_t vi = rec[(int64_t)bi * mstride + e];
        if (vi < 0) continue;
        for (int a = 0; a < D; ++a)
        {
            const double v = grad[(int64_t)(D * bi + a) * stride + e];
            finite = finite && isfinite(v);
            atomicAdd(&g[(int64_t)D * vi + a], v);
        }
    }
    if (hess)
        for (int bi = 0; bi < N; ++bi)
        {
            const int32_t rs = rstride[(int64_t)bi * mstride + e];
            for (int bj = 0; bj < N; ++bj)
            {
                const int32_t base = blockbase[(int64_t)(bi * N + bj) * mstride + e];
                if (base < 0) continue;
                for (int a = 0; a < D; ++a)
                    for (int b = 0; b < D; ++b)
                    {
                        const int s = seq.idx[(D * bi + a) * K + (D * bj + b)];
                        const double v = hess[(int64_t)s * stride + e];
                        finite = finite && isfinite(v);
                        atomicAdd(&Hv[(int64_t)base + (int64_t)a * rs + b], v);
                    }
            }
        }
    if (!finite) atomicOr(err, ERR_NONFINITE);
}

int assemble_atomic(int d, int N, const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double* g, double* Hv,
                    int32_t* err, cudaStream_t st)
{
    if (n <= 0) return TAD_OK;
    count_launch();
    bool done = false;
    if (d == 1) done = assemble_atomic_d<1>(N, m, grad, hess, n, stride, g, Hv, err, st);
    else if (d == 2) done = assemble_atomic_d<2>(N, m, grad, hess, n, stride, g, Hv, err, st);
    else if (d == 3) done = assemble_atomic_d<3>(N, m, grad, hess, n, stride, g, Hv, err, st);
    if (!done)
    {
        const int K = d * N;
        if (K > 32) return fail(TAD_NOT_SUPPORTED, "assembly supports at most 32 variables per element");
        SeqTable seq;
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) seq.idx[i * K + j] = (int16_t)hess_seq_index(K, i, j);
        assemble_atomic_generic<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d, N, seq, m.rec, m.blockbase, m.rstride, m.mstride, grad, hess, n, stride,
                                                                            g, Hv, err);
    }
    return cudaGetLastError() == cudaSuccess ? TAD_OK : fail(TAD_CUDA_ERROR, "assembly kernel launch failed");
}

// ---------------------------------------------------------------------------------------------
// assembly, gather mode: one thread per CSR entry, contributions summed in (term, element) order --
// the order in which the reference's setFromTriplets adds duplicates -- deterministic, no atomics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gather_hessian(const int64_t* __restrict__ block_ptr, const int32_t* __restrict__ contrib,
                                                      const int64_t* __restrict__ block_key, const int64_t* __restrict__ vrow,
                                                      const TermDev* __restrict__ terms, int n_terms, SeqTable const* __restrict__ seqs,
                                                      int64_t n_blocks, int64_t n_handles, int d, double* __restrict__ Hv, int32_t* err)
{
    const int dd = d * d;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p = tid / dd;
    if (p >= n_blocks) return;
    const int ab = (int)(tid % dd), a = ab / d, b = ab % d;
    double acc = 0.0;
    for (int64_t i = block_ptr[p]; i < block_ptr[p + 1]; ++i)
    {
        const int64_t c = contrib[i];
        if (c < 0) continue;  // structural-only block
        int t = 0;
        while (t + 1 < n_terms && terms[t + 1].off <= c) ++t;
        const TermDev& T = terms[t];
        const int64_t local = c - T.off;
        const int64_t nn = (int64_t)T.N * T.N;
        const int64_t e = local / nn;
        const int bb = (int)(local % nn);
        const int bi = bb / T.N, bj = bb % T.N;
        const int s = seqs[t].idx[(d * bi + a) * T.k + (d * bj + b)];
        acc += T.hess[(int64_t)s * T.stride + e];
    }
    if (!isfinite(acc)) atomicOr(err, ERR_NONFINITE);
    const int64_t vi = block_key[p] / n_handles;
    const int64_t r0 = vrow[vi], deg = vrow[vi + 1] - r0;
    Hv[(int64_t)dd * r0 + (int64_t)a * d * deg + (int64_t)d * (p - r0) + b] = acc;
}

__global__ void __launch_bounds__(128) gather_gradient(const int64_t* __restrict__ block_ptr, const int32_t* __restrict__ contrib,
                                                       const int64_t* __restrict__ block_key, const TermDev* __restrict__ terms,
                                                       int n_terms, int64_t n_blocks, int64_t n_handles, int d, double* __restrict__ g,
                                                       int32_t* err)
{
    // one thread per (vertex, component); contributions of vertex v are those of its diagonal block (v, v)
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t v = tid / d;
    if (v >= n_handles) return;
    const int a = (int)(tid % d);
    const int64_t target = v * n_handles + v;
    int64_t lo = 0, hi = n_blocks;
    while (lo < hi)
    {
        const int64_t mid = (lo + hi) / 2;
        if (block_key[mid] < target) lo = mid + 1; else hi = mid;
    }
    double acc = 0.0;
    if (lo < n_blocks && block_key[lo] == target)
        for (int64_t i = block_ptr[lo]; i < block_ptr[lo + 1]; ++i)
        {
            const int64_t c = contrib[i];
            if (c < 0) continue;  // structural-only block
            int t = 0;
            while (t + 1 < n_terms && terms[t + 1].off <= c) ++t;
            const TermDev& T = terms[t];
            const int64_t local = c - T.off;
            const int64_t nn = (int64_t)T.N * T.N;
            const int64_t e = local / nn;
            const int bi = (int)(local % nn) / T.N;
            acc += T.grad[(int64_t)(d * bi + a) * T.stride + e];
        }
    if (!isfinite(acc)) atomicOr(err, ERR_NONFINITE);
    g[(int64_t)d * v + a] = acc;
}

int gather_assemble(const int64_t* block_ptr, const int32_t* contrib, const int64_t* block_key, const int64_t* vrow, const TermDev* terms,
                    int n_terms, const SeqTable* seqs, int64_t n_blocks, int64_t n_handles, int64_t n_vars, int d, double* g, double* Hv,
                    int32_t* err, cudaStream_t st)
{
    const int64_t nt = n_blocks * d * d;
    count_launch(nt > 0 ? 2 : 1);
    if (nt > 0)
        gather_hessian<<<blocks_for(nt, 128), 128, 0, st>>>(block_ptr, contrib, block_key, vrow, terms, n_terms, seqs, n_blocks, n_handles, d, Hv, err);
    gather_gradient<<<blocks_for(n_vars, 128), 128, 0, st>>>(block_ptr, contrib, block_key, terms, n_terms, n_blocks, n_handles, d, g, err);
    return cudaGetLastError() == cudaSuccess ? TAD_OK : fail(TAD_CUDA_ERROR, "gather assembly launch failed");
}
#endif  // TAD_ASM_PART == 0

}  // namespace tadrt
