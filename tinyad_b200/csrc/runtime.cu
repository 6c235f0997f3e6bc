// tinyad_b200 runtime (libtinyad_b200.so) -- functor-independent half of the B200 path.
// Implements include/tinyad_b200.h.  sm_100a only; no CPU fallback.
//
// What lives here (reference counterpart in /root/reference/include/TinyAD):
//   record pass bookkeeping, fixed CSR pattern + scatter maps   <- rediscovered at every eval by Element.hh:208-260 +
//                                                                  setFromTriplets (ScalarFunctionImpl.hh:398)
//   batched PSD projection kernel                               <- Utils/HessianProjection.hh:23-101
//   assembly kernels (FP64 atomics | deterministic gather)      <- serial loops ScalarObjectiveTerm.hh:256-277
//   f reduction, finite checks, error word                      <- ScalarObjectiveTerm.hh:210,252-253, Utils/Out.hh:73-79
//   VectorFunction r / J (CSC) / sum of squares / g = 2 J^T r   <- VectorObjectiveTerm.hh:158-243, VectorFunctionImpl.hh:143-283
#include "rt_common.cuh"

using namespace tadrt;

namespace tadrt
{
thread_local std::string g_last_error = "";
thread_local int64_t* tl_launch_counter = nullptr;
}  // namespace tadrt

struct tad_function_s
{
    int d = 0;
    int64_t n_handles = 0, n_vars = 0;
    bool is_vector = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<Term> terms;
    int64_t n_elements = 0, n_outputs = 0;
    // options
    int assembly = TAD_ASSEMBLY_ATOMIC;
    int64_t chunk = 0;             // TAD_OPT_CHUNK_ELEMENTS: 0 default, < 0 whole term
    int n_lanes = 2;               // TAD_OPT_LANES
    bool timing = false;
    // pattern
    bool pattern_built = false;
    int64_t nnz = 0, n_outer = 0, n_blocks = 0, n_contrib = 0;
    DevBuf<int32_t> outer, inner;
    DevBuf<int64_t> block_ptr;     // [n_blocks+1] into contrib
    DevBuf<int32_t> contrib;       // sorted contribution ids
    DevBuf<int64_t> block_key;     // [n_blocks]
    DevBuf<int64_t> vrow;          // [n_handles+1] first block of each vertex row
    std::vector<int64_t> vrow_host;   // host copy (row finality of the slab schedule)
    DevBuf<TermDev> terms_dev;
    std::vector<TermDev> terms_dev_host;  // what terms_dev holds
    DevBuf<unsigned char> seqs_dev;   // SeqTable per term (gather assembly), built with the pattern
    std::vector<int64_t> extra_keys;  // vertex pairs injected by tad_function_add_pattern_blocks (halo rows of other ranks)
    // slab schedule of the second-order evaluation (rebuilt when the pattern or the options change)
    struct Schedule
    {
        std::vector<Slab> slabs;
        int64_t chunk = -2;       // slab size the schedule was built for (-1: whole terms)
        bool whole = false;
        bool has_final = false;   // Slab::final_values computed (needs the pattern)
        int n_halo_slabs = 0;     // multi-GPU: the first n_halo_slabs slabs touch vertices owned by other ranks (scheduled first)
        void clear() { slabs.clear(); chunk = -2; has_final = false; n_halo_slabs = 0; }
    };
    Schedule sched[2];            // [0] device-pointer entry points, [1] host-buffer entry points (smaller slabs: finer D2H pipelining)
    // scratch
    DevBuf<double> x_dev, g_dev, H_dev, r_dev;
    DevBuf<double> stage;          // vector functions: staging of the current term
    DevBuf<int32_t> err;           // int32[8]
    DevBuf<double> fpart;          // block partial sums
    DevBuf<double> fterm;          // per-term sums
    int projection_full = 0;       // option: 1 = always use the full eigensolver kernel
    int64_t last_proj[3] = {0, 0, 0};
    float last_ms[4] = {0, 0, 0, 0};
    int64_t n_launches = 0;        // kernels launched so far (tad_function_launch_count)
    cudaEvent_t ev[2] = {nullptr, nullptr};   // begin / end of an evaluation on the main stream
    cudaEvent_t ev_caller = nullptr;
    cudaStream_t caller_stream = nullptr;
    bool wait_caller = false;      // tad_function_set_caller_stream
    std::vector<Lane> lanes;
    std::vector<Lane> prio_lane;           // 0 or 1 lane with high stream priority (partitioned functions: the slabs the exchange waits for)
    std::vector<cudaEvent_t> slab_events;  // one per slab of the schedule: "this slab is assembled"
    cudaStream_t copy_stream = nullptr;    // device -> host copies of finished rows (host-buffer entry points)
    // multi-GPU (tad_function_set_comm): the halo plan is built with the pattern; all NCCL calls of an evaluation go to comm_stream
    tad_comm comm = nullptr;
    HaloPlan halo;
    int replicate_gradient = 0;            // TAD_OPT_REPLICATE_GRADIENT
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_comm[2] = {nullptr, nullptr};   // main stream -> comm stream, comm stream -> main stream
    cudaEvent_t ev_trace[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // TAD_COMM_TRACE=1: timeline of the exchange
    std::recursive_mutex mtx;      // eval* may be called concurrently (ScalarFunctionTest.cc:255-291): calls on one function are serialised
};

namespace
{

// ---------------------------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------------------------
__global__ void fill_i32(int32_t* p, int64_t n, int32_t v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// rec_counts < 0 marks repeated handles; writes flags[0] |= 1 if any
__global__ void scan_counts(const int32_t* counts, int64_t n, int32_t* flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && counts[i] < 0) atomicOr(flags, 1);
}

// Deterministic sum: stage 1, one partial per block (fixed tree), stage 2 one block over partials.
template <bool SQUARE>
__global__ void __launch_bounds__(256) reduce_stage1(const double* v, int64_t n, int64_t stride, int rows, double* partial)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int r = 0; r < rows; ++r)
    {
        const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
        if (i < n)
        {
            const double t = v[(int64_t)r * stride + i];
            s += SQUARE ? t * t : t;
        }
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(1024) reduce_stage2(const double* partial, int64_t nb, double* out)
{
    __shared__ double sh[1024];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < nb; i += 1024) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1)
    {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

}  // namespace

namespace tadrt
{
// k with a dedicated (register-resident) instantiation of the projection kernels; every other k <= 32 runs the run-time k kernel
bool project_has_instance(int k) { return (k >= 1 && k <= 10) || k == 12 || k == 15 || k == 16 || k == 18; }

size_t project_scratch_doubles_rt(int k, int64_t stride)
{
    switch (k)
    {
    case 1: return project_scratch_doubles<1>(stride);
    case 2: return project_scratch_doubles<2>(stride);
    case 3: return project_scratch_doubles<3>(stride);
    case 4: return project_scratch_doubles<4>(stride);
    case 5: return project_scratch_doubles<5>(stride);
    case 6: return project_scratch_doubles<6>(stride);
    case 7: return project_scratch_doubles<7>(stride);
    case 8: return project_scratch_doubles<8>(stride);
    case 9: return project_scratch_doubles<9>(stride);
    case 10: return project_scratch_doubles<10>(stride);
    case 12: return project_scratch_doubles<12>(stride);
    case 15: return project_scratch_doubles<15>(stride);
    case 16: return project_scratch_doubles<16>(stride);
    case 18: return project_scratch_doubles<18>(stride);
    default: return (k >= 1 && k <= 32) ? project_scratch_doubles_generic(k) : 0;
    }
}

int project_dispatch(int k, double* hess, int64_t n, int64_t stride, double eps, unsigned long long* counts, double* scratch_d,
                     int32_t* codes, int64_t* list, bool full_only, ProjScratch* fuse_out, const ProjSide* side, cudaStream_t st, int tdim)
{
    if (n <= 0 || k <= 0) return TAD_OK;
    switch (k)
    {
    case 1: return launch_project<1>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 2: return launch_project<2>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 3: return launch_project<3>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 4: return launch_project<4>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 5: return launch_project<5>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 6: return launch_project<6>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 7: return launch_project<7>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 8: return launch_project<8>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 9: return launch_project<9>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 10: return launch_project<10>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 12: return launch_project<12>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 15: return launch_project<15>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 16: return launch_project<16>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    case 18: return launch_project<18>(hess, n, stride, eps, counts, scratch_d, codes, list, full_only, fuse_out, side, st, tdim);
    default: return launch_project_generic(k, hess, n, stride, eps, counts, scratch_d, st);   // any other k <= 32: run-time k Jacobi
    }
}
}  // namespace tadrt

namespace
{

// ---------------------------------------------------------------------------------------------
// pattern construction (scalar functions): vertex-pair blocks -> CSR + scatter maps
// ---------------------------------------------------------------------------------------------
__global__ void gen_block_keys(TermDev t, int64_t n_handles, int64_t* keys, int32_t* payload)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // local contribution id, element-major
    const int64_t nn = (int64_t)t.N * t.N;
    if (c >= t.n * nn) return;
    const int64_t e = c / nn;
    const int b = (int)(c % nn);
    const int bi = b / t.N, bj = b % t.N;
    const int32_t vi = t.rec[(int64_t)bi * t.stride + e];
    const int32_t vj = t.rec[(int64_t)bj * t.stride + e];
    keys[t.off + c] = (vi < 0 || vj < 0) ? INT64_MAX : (int64_t)vi * n_handles + vj;
    payload[t.off + c] = (int32_t)(t.off + c);
}

__global__ void count_valid(const int64_t* keys, int64_t n, int64_t* out)
{
    // keys sorted ascending; binary search for the first INT64_MAX (single thread)
    int64_t lo = 0, hi = n;
    while (lo < hi)
    {
        const int64_t mid = (lo + hi) / 2;
        if (keys[mid] == INT64_MAX) hi = mid; else lo = mid + 1;
    }
    *out = lo;
}

__global__ void head_flags(const int64_t* keys, int64_t n, int32_t* flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void scatter_heads(const int64_t* keys, const int32_t* flags, const int32_t* pid_incl, int64_t n,
                              int64_t* block_key, int64_t* block_ptr, int64_t n_blocks)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i])
    {
        const int32_t p = pid_incl[i] - 1;
        block_key[p] = keys[i];
        block_ptr[p] = i;
    }
    if (i == 0) block_ptr[n_blocks] = n;
}

__global__ void vertex_rows(const int64_t* block_key, int64_t n_blocks, int64_t n_handles, int64_t* vrow)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > n_handles) return;
    const int64_t target = v * n_handles;
    int64_t lo = 0, hi = n_blocks;
    while (lo < hi)
    {
        const int64_t mid = (lo + hi) / 2;
        if (block_key[mid] < target) lo = mid + 1; else hi = mid;
    }
    vrow[v] = lo;
}

__global__ void fill_csr(const int64_t* block_key, const int64_t* vrow, int64_t n_blocks, int64_t n_handles, int d,
                         int32_t* outer, int32_t* inner)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_blocks)
    {
        const int64_t vi = block_key[p] / n_handles, vj = block_key[p] % n_handles;
        const int64_t r0 = vrow[vi], deg = vrow[vi + 1] - r0;
        for (int a = 0; a < d; ++a)
            for (int b = 0; b < d; ++b)
                inner[(int64_t)d * d * r0 + (int64_t)a * d * deg + (int64_t)d * (p - r0) + b] = (int32_t)(d * vj + b);
    }
    if (p < n_handles)
    {
        const int64_t r0 = vrow[p], deg = vrow[p + 1] - r0;
        for (int a = 0; a < d; ++a) outer[d * p + a] = (int32_t)((int64_t)d * d * r0 + (int64_t)a * d * deg);
    }
    if (p == 0) outer[(int64_t)d * n_handles] = (int32_t)((int64_t)d * d * n_blocks);
}

__global__ void fill_maps(const int32_t* contrib, const int32_t* pid_incl, int64_t n_valid, const TermDev* terms, int n_terms,
                          const int64_t* block_key, const int64_t* vrow, int64_t n_handles, int d)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_valid) return;
    const int64_t c = contrib[i];
    if (c < 0) return;  // structural-only block
    const int64_t p = pid_incl[i] - 1;
    int t = 0;
    while (t + 1 < n_terms && terms[t + 1].off <= c) ++t;
    const TermDev T = terms[t];
    const int64_t local = c - T.off;
    const int64_t nn = (int64_t)T.N * T.N;
    const int64_t e = local / nn;
    const int b = (int)(local % nn);
    const int bi = b / T.N;
    const int64_t vi = block_key[p] / n_handles;
    const int64_t r0 = vrow[vi], deg = vrow[vi + 1] - r0;
    T.blockbase[(int64_t)b * T.stride + e] = (int32_t)((int64_t)d * d * r0 + (int64_t)d * (p - r0));
    T.rstride[(int64_t)bi * T.stride + e] = (int32_t)(d * deg);
}

// ---------------------------------------------------------------------------------------------
// vector functions
// ---------------------------------------------------------------------------------------------
__global__ void gen_jac_keys(TermDev t, int d, int64_t n_outputs, int64_t* keys, int32_t* payload)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // element-major: ((e*M + m)*k + i)
    const int64_t per = (int64_t)t.M * t.k;
    if (c >= t.n * per) return;
    const int64_t e = c / per;
    const int mi = (int)(c % per);
    const int m = mi / t.k, i = mi % t.k;
    const int32_t v = t.rec[(int64_t)(i / d) * t.stride + e];
    const int64_t row = t.out_offset + (int64_t)t.M * e + m;
    keys[t.off + c] = (v < 0) ? INT64_MAX : ((int64_t)d * v + (i % d)) * n_outputs + row;
    payload[t.off + c] = (int32_t)(t.off + c);
}

__global__ void fill_jac(const int64_t* keys, const int32_t* contrib, int64_t n_valid, const TermDev* terms, int n_terms,
                         int64_t n_outputs, int32_t* inner)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_valid) return;
    inner[i] = (int32_t)(keys[i] % n_outputs);
    const int64_t c = contrib[i];
    int t = 0;
    while (t + 1 < n_terms && terms[t + 1].off <= c) ++t;
    const TermDev T = terms[t];
    const int64_t local = c - T.off;
    const int64_t per = (int64_t)T.M * T.k;
    const int64_t e = local / per;
    const int mi = (int)(local % per);
    T.jslot[(int64_t)mi * T.stride + e] = (int32_t)i;
}

__global__ void jac_col_ptr(const int64_t* keys, int64_t n_valid, int64_t n_vars, int64_t n_outputs, int32_t* outer)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_vars) return;
    const int64_t target = c * n_outputs;
    int64_t lo = 0, hi = n_valid;
    while (lo < hi)
    {
        const int64_t mid = (lo + hi) / 2;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    outer[c] = (int32_t)lo;
}

__global__ void scatter_residuals(const double* val, int64_t n, int64_t stride, int M, int64_t out_offset, double* r)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // i = e*M + m
    if (i >= n * M) return;
    const int64_t e = i / M;
    const int m = (int)(i % M);
    r[out_offset + i] = val[(int64_t)m * stride + e];
}

__global__ void scatter_jacobian(const double* grad, const int32_t* jslot, int64_t n, int64_t stride, int rows, double* Jv, int32_t* err)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    bool finite = true;
    for (int r = 0; r < rows; ++r)
    {
        const int32_t s = jslot[(int64_t)r * stride + e];
        if (s < 0) continue;
        const double v = grad[(int64_t)r * stride + e];
        finite = finite && isfinite(v);
        Jv[s] = v;
    }
    if (!finite) atomicOr(err, ERR_NONFINITE);
}

// Per-residual Hessians of a vector term: staged packed entries hess[(m * nh + s) * stride + e] -> dense row-major k x k blocks,
// one per residual, out[((M * e + m) * k + i) * k + j] (element-local variable order).
__global__ void __launch_bounds__(256) scatter_hess_blocks(const double* __restrict__ hess, int64_t n, int64_t stride, int M, int k, int nh, SeqTable seq,
                                                           double* __restrict__ out, int32_t* err)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int kk = k * k;
    if (i >= n * M * kk) return;
    const int64_t em = i / kk;          // M * e + m
    const int ij = (int)(i % kk);
    const int64_t e = em / M;
    const int m = (int)(em % M);
    const double v = hess[(int64_t)(m * nh + seq.idx[ij]) * stride + e];
    if (!isfinite(v)) atomicOr(err, ERR_NONFINITE);
    out[i] = v;
}

__global__ void jt_r(const int32_t* outer, const int32_t* inner, const double* Jv, const double* r, int64_t n_vars, double* g)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_vars) return;
    double s = 0.0;
    for (int32_t p = outer[c]; p < outer[c + 1]; ++p) s += Jv[p] * r[inner[p]];
    g[c] = 2.0 * s;
}


// ---------------------------------------------------------------------------------------------
// FP64 pipe microbenchmark (roofline denominator): 8 independent DFMA chains per thread
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;  // keep the chains alive
}

// ---------------------------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------------------------

int check_error_word(tad_function f, bool sync_already)
{
    int32_t h_err[8] = {0};
    if (!sync_already) TAD_CUDA(cudaStreamSynchronize(f->stream));
    TAD_CUDA(cudaMemcpy(h_err, f->err.p, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (h_err[0] & ERR_RANGE) return fail(TAD_INDEX_OUT_OF_RANGE, "variable handle out of range in element.variables(...)");
    if (h_err[0] & ERR_TOO_MANY) return fail(TAD_TOO_MANY_VARIABLES, "Too many variables requested via element.variables(...).");
    if (h_err[0] & ERR_PATTERN)
        return fail(TAD_PATTERN_MISMATCH, "an element requested different variable handles than when it was added (the sparsity pattern is fixed at add_elements time)");
    if (h_err[0] & ERR_NONFINITE) return fail(TAD_NONFINITE_DERIVATIVE, "non-finite element gradient or Hessian");
    return TAD_OK;
}

int sum_to(tad_function f, const double* v, int64_t n, int64_t stride, int rows, bool square, double* out_dev)
{
    const int64_t nb = std::max<int64_t>(1, (n + 255) / 256);
    TAD_CUDA(f->fpart.ensure((size_t)nb));
    if (n <= 0)
    {
        TAD_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), f->stream));
        return TAD_OK;
    }
    count_launch(2);
    if (square) reduce_stage1<true><<<(unsigned)nb, 256, 0, f->stream>>>(v, n, stride, rows, f->fpart.p);
    else reduce_stage1<false><<<(unsigned)nb, 256, 0, f->stream>>>(v, n, stride, rows, f->fpart.p);
    reduce_stage2<<<1, 1024, 0, f->stream>>>(f->fpart.p, nb, out_dev);
    return cudaGetLastError() == cudaSuccess ? TAD_OK : fail(TAD_CUDA_ERROR, "reduction launch failed");
}

// Launch arguments of the slab [e_begin, e_begin + n) of term t; `stage` holds the slab's outputs with leading dimension sstride.
void fill_launch_args(tad_function f, const Term& t, int mode, const double* x, double* stage, int64_t e_begin, int64_t n, int64_t sstride,
                      cudaStream_t stream, tad_launch_args& a)
{
    std::memset(&a, 0, sizeof(a));
    a.mode = mode;
    a.dedup = t.dedup ? 1 : 0;
    a.n_elements = n;
    a.stride = sstride;
    a.e_begin = e_begin;
    a.rec_stride = t.stride;
    a.elem_handles = t.has_handles ? t.elem_handles.p : nullptr;
    a.x = x;
    a.n_handles = f->n_handles;
    const int rows = std::max(1, t.M);
    a.val = stage;
    a.grad = stage ? stage + (int64_t)rows * sstride : nullptr;
    a.hess = stage ? a.grad + (int64_t)rows * t.k * sstride : nullptr;   // scalar terms: after grad[k]; vector terms: after grad[M * k]
    a.rec_handles = t.rec_handles.p;
    a.rec_counts = t.rec_counts.p;
    a.error_flags = f->err.p;
    a.stream = stream;
    a.launch_counter = &f->n_launches;
}
void fill_launch_args(tad_function f, const Term& t, int mode, const double* x, double* stage, tad_launch_args& a)
{
    fill_launch_args(f, t, mode, x, stage, 0, t.n, t.stride, f->stream, a);
}

size_t stage_doubles(const Term& t, int mode, int64_t sstride)
{
    const int rows = std::max(1, t.M);
    size_t per = rows;
    if (mode >= TAD_MODE_FIRST) per += (size_t)rows * t.k;
    if (mode == TAD_MODE_SECOND) per += (size_t)rows * hess_size(t.k);   // vector terms: one packed Hessian per residual
    return per * (size_t)sstride;
}
size_t stage_doubles(const Term& t, int mode) { return stage_doubles(t, mode, t.stride); }

// Device copy of the term descriptions (pattern construction, gather assembly).  Re-uploaded only when something changed
// (the staging pointers of the gather mode move when a buffer grows).
int upload_terms_dev(tad_function f, bool with_stage_ptrs, int mode)
{
    std::vector<TermDev> td(f->terms.size());
    for (size_t i = 0; i < f->terms.size(); ++i)
    {
        Term& t = f->terms[i];
        std::memset(&td[i], 0, sizeof(TermDev));
        td[i].off = t.contrib_offset;
        td[i].n = t.n;
        td[i].stride = t.stride;
        td[i].N = t.N; td[i].M = t.M; td[i].k = t.k;
        td[i].rec = t.rec_handles.p;
        td[i].blockbase = t.blockbase.p;
        td[i].rstride = t.rstride.p;
        td[i].jslot = t.jslot.p;
        td[i].out_offset = t.out_offset;
        td[i].grad = nullptr; td[i].hess = nullptr;
        if (with_stage_ptrs && t.stage.p)
        {
            td[i].grad = t.stage.p + t.stride;
            td[i].hess = (mode == TAD_MODE_SECOND) ? t.stage.p + (int64_t)(1 + t.k) * t.stride : nullptr;
        }
    }
    if (td.size() == f->terms_dev_host.size() && f->terms_dev.p &&
        (td.empty() || std::memcmp(td.data(), f->terms_dev_host.data(), td.size() * sizeof(TermDev)) == 0))
        return TAD_OK;
    TAD_CUDA(f->terms_dev.ensure(td.size()));
    if (!td.empty())
        TAD_CUDA(cudaMemcpyAsync(f->terms_dev.p, td.data(), td.size() * sizeof(TermDev), cudaMemcpyHostToDevice, f->stream));
    TAD_CUDA(cudaStreamSynchronize(f->stream));  // td is a local
    f->terms_dev_host = td;
    return TAD_OK;
}

int build_pattern_scalar(tad_function f)
{
    cudaStream_t st = f->stream;
    const int d = f->d;
    int64_t nC = 0;
    for (auto& t : f->terms)
    {
        t.contrib_offset = nC;
        nC += (int64_t)t.N * t.N * t.n;
    }
    const int64_t n_term_contrib = nC;
    std::vector<int64_t> structural = f->extra_keys;   // structural-only blocks: injected by the user, or sent here by other ranks
    structural.insert(structural.end(), f->halo.keys_from_peers.begin(), f->halo.keys_from_peers.end());
    nC += (int64_t)structural.size();
    if (nC >= (int64_t)INT32_MAX) return fail(TAD_NOT_SUPPORTED, "more than 2^31 block contributions");
    f->n_contrib = nC;
    for (auto& t : f->terms)
    {
        TAD_CUDA(t.blockbase.ensure((size_t)t.N * t.N * t.stride));
        TAD_CUDA(t.rstride.ensure((size_t)t.N * t.stride));
        const int64_t nb = (int64_t)t.N * t.N * t.stride;
        if (nb) fill_i32<<<blocks_for(nb, 256), 256, 0, st>>>(t.blockbase.p, nb, -1);
        const int64_t nr = (int64_t)t.N * t.stride;
        if (nr) fill_i32<<<blocks_for(nr, 256), 256, 0, st>>>(t.rstride.p, nr, 0);
    }
    TAD_TRY(upload_terms_dev(f, false, 0));

    DevBuf<int64_t> keys_a, keys_b, nvalid_d;
    DevBuf<int32_t> pay_a, pay_b, flags, pid;
    TAD_CUDA(keys_a.ensure((size_t)nC)); TAD_CUDA(keys_b.ensure((size_t)nC));
    TAD_CUDA(pay_a.ensure((size_t)nC)); TAD_CUDA(pay_b.ensure((size_t)nC));
    TAD_CUDA(nvalid_d.ensure(1));
    std::vector<TermDev> td(f->terms.size());
    TAD_CUDA(cudaMemcpy(td.data(), f->terms_dev.p, td.size() * sizeof(TermDev), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < f->terms.size(); ++i)
    {
        const int64_t cnt = (int64_t)td[i].N * td[i].N * td[i].n;
        if (cnt) gen_block_keys<<<blocks_for(cnt, 256), 256, 0, st>>>(td[i], f->n_handles, keys_a.p, pay_a.p);
    }
    if (!structural.empty())
    {
        // structural-only blocks (no local contribution): payload -1
        const size_t ne = structural.size();
        TAD_CUDA(cudaMemcpyAsync(keys_a.p + n_term_contrib, structural.data(), ne * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        fill_i32<<<blocks_for((int64_t)ne, 256), 256, 0, st>>>(pay_a.p + n_term_contrib, (int64_t)ne, -1);
    }
    int64_t n_valid = 0, n_blocks = 0;
    if (nC > 0)
    {
        // number of significant key bits: keys < n_handles^2 (invalid keys are INT64_MAX -> need the full width if present)
        size_t tmp_bytes = 0;
        cub::DoubleBuffer<int64_t> kb(keys_a.p, keys_b.p);
        cub::DoubleBuffer<int32_t> pb(pay_a.p, pay_b.p);
        TAD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, pb, (int)nC, 0, 64, st));
        DevBuf<unsigned char> tmp;
        TAD_CUDA(tmp.ensure(tmp_bytes));
        TAD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kb, pb, (int)nC, 0, 64, st));
        int64_t* keys = kb.Current();
        int32_t* pay = pb.Current();
        count_valid<<<1, 1, 0, st>>>(keys, nC, nvalid_d.p);
        TAD_CUDA(cudaMemcpyAsync(&n_valid, nvalid_d.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        TAD_CUDA(cudaStreamSynchronize(st));
        TAD_CUDA(flags.ensure((size_t)std::max<int64_t>(n_valid, 1)));
        TAD_CUDA(pid.ensure((size_t)std::max<int64_t>(n_valid, 1)));
        if (n_valid > 0)
        {
            head_flags<<<blocks_for(n_valid, 256), 256, 0, st>>>(keys, n_valid, flags.p);
            size_t scan_bytes = 0;
            TAD_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, flags.p, pid.p, (int)n_valid, st));
            DevBuf<unsigned char> tmp2;
            TAD_CUDA(tmp2.ensure(scan_bytes));
            TAD_CUDA(cub::DeviceScan::InclusiveSum(tmp2.p, scan_bytes, flags.p, pid.p, (int)n_valid, st));
            int32_t last = 0;
            TAD_CUDA(cudaMemcpyAsync(&last, pid.p + (n_valid - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            TAD_CUDA(cudaStreamSynchronize(st));
            n_blocks = last;
        }
        const int64_t nnz = (int64_t)d * d * n_blocks;
        if (nnz >= (int64_t)INT32_MAX) return fail(TAD_NOT_SUPPORTED, "Hessian has more than 2^31 non-zeros (int32 StorageIndex)");
        TAD_CUDA(f->block_key.ensure((size_t)std::max<int64_t>(n_blocks, 1)));
        TAD_CUDA(f->block_ptr.ensure((size_t)n_blocks + 1));
        TAD_CUDA(f->vrow.ensure((size_t)f->n_handles + 1));
        TAD_CUDA(f->outer.ensure((size_t)f->n_vars + 1));
        TAD_CUDA(f->inner.ensure((size_t)std::max<int64_t>(nnz, 1)));
        TAD_CUDA(f->contrib.ensure((size_t)std::max<int64_t>(n_valid, 1)));
        if (n_valid > 0)
        {
            scatter_heads<<<blocks_for(n_valid, 256), 256, 0, st>>>(keys, flags.p, pid.p, n_valid, f->block_key.p, f->block_ptr.p, n_blocks);
            TAD_CUDA(cudaMemcpyAsync(f->contrib.p, pay, (size_t)n_valid * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        }
        else
            TAD_CUDA(cudaMemsetAsync(f->block_ptr.p, 0, sizeof(int64_t), st));
        vertex_rows<<<blocks_for(f->n_handles + 1, 256), 256, 0, st>>>(f->block_key.p, n_blocks, f->n_handles, f->vrow.p);
        fill_csr<<<blocks_for(std::max(n_blocks, f->n_handles), 256), 256, 0, st>>>(f->block_key.p, f->vrow.p, n_blocks, f->n_handles, d,
                                                                                  f->outer.p, f->inner.p);
        if (n_valid > 0)
            fill_maps<<<blocks_for(n_valid, 256), 256, 0, st>>>(f->contrib.p, pid.p, n_valid, f->terms_dev.p, (int)f->terms.size(),
                                                                f->block_key.p, f->vrow.p, f->n_handles, d);
        TAD_CUDA(cudaGetLastError());
        TAD_CUDA(cudaStreamSynchronize(st));
        f->nnz = nnz;
    }
    else
    {
        TAD_CUDA(f->outer.ensure((size_t)f->n_vars + 1));
        TAD_CUDA(cudaMemsetAsync(f->outer.p, 0, ((size_t)f->n_vars + 1) * sizeof(int32_t), st));
        TAD_CUDA(f->inner.ensure(1));
        TAD_CUDA(f->block_ptr.ensure(1));
        TAD_CUDA(cudaMemsetAsync(f->block_ptr.p, 0, sizeof(int64_t), st));
        TAD_CUDA(f->block_key.ensure(1));
        TAD_CUDA(f->vrow.ensure((size_t)f->n_handles + 1));
        TAD_CUDA(cudaMemsetAsync(f->vrow.p, 0, ((size_t)f->n_handles + 1) * sizeof(int64_t), st));
        TAD_CUDA(f->contrib.ensure(1));
        TAD_CUDA(cudaStreamSynchronize(st));
        f->nnz = 0;
    }
    f->n_blocks = n_blocks;
    f->n_outer = f->n_vars;
    // host copy of the vertex rows (row finality of the slab schedule) and the per-term packed-index tables of the gather assembly
    f->vrow_host.assign((size_t)f->n_handles + 1, 0);
    TAD_CUDA(cudaMemcpy(f->vrow_host.data(), f->vrow.p, ((size_t)f->n_handles + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    {
        std::vector<SeqTable> seqs(std::max<size_t>(f->terms.size(), 1));
        for (size_t ti = 0; ti < f->terms.size(); ++ti)
        {
            const int K = f->terms[ti].k;
            if (K > 32) continue;  // gather assembly reports TAD_NOT_SUPPORTED for such a term
            for (int i = 0; i < K; ++i)
                for (int j = 0; j < K; ++j) seqs[ti].idx[i * K + j] = (int16_t)hess_seq_index(K, i, j);
        }
        TAD_CUDA(f->seqs_dev.ensure(seqs.size() * sizeof(SeqTable)));
        TAD_CUDA(cudaMemcpy(f->seqs_dev.p, seqs.data(), seqs.size() * sizeof(SeqTable), cudaMemcpyHostToDevice));
    }
    f->sched[0].clear();
    f->sched[1].clear();
    f->pattern_built = true;
    return TAD_OK;
}

int build_pattern_vector(tad_function f)
{
    cudaStream_t st = f->stream;
    int64_t nC = 0, out = 0;
    for (auto& t : f->terms)
    {
        t.contrib_offset = nC;
        t.out_offset = out;
        nC += (int64_t)t.M * t.k * t.n;
        out += (int64_t)t.M * t.n;
    }
    f->n_outputs = out;
    if (nC >= (int64_t)INT32_MAX) return fail(TAD_NOT_SUPPORTED, "Jacobian has more than 2^31 entries");
    for (auto& t : f->terms)
    {
        const int64_t nj = (int64_t)t.M * t.k * t.stride;
        TAD_CUDA(t.jslot.ensure((size_t)nj));
        if (nj) fill_i32<<<blocks_for(nj, 256), 256, 0, st>>>(t.jslot.p, nj, -1);
    }
    TAD_TRY(upload_terms_dev(f, false, 0));
    TAD_CUDA(f->outer.ensure((size_t)f->n_vars + 1));
    int64_t n_valid = 0;
    if (nC > 0)
    {
        DevBuf<int64_t> keys_a, keys_b, nvalid_d;
        DevBuf<int32_t> pay_a, pay_b;
        TAD_CUDA(keys_a.ensure((size_t)nC)); TAD_CUDA(keys_b.ensure((size_t)nC));
        TAD_CUDA(pay_a.ensure((size_t)nC)); TAD_CUDA(pay_b.ensure((size_t)nC));
        TAD_CUDA(nvalid_d.ensure(1));
        std::vector<TermDev> td(f->terms.size());
        TAD_CUDA(cudaMemcpy(td.data(), f->terms_dev.p, td.size() * sizeof(TermDev), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < td.size(); ++i)
        {
            const int64_t cnt = (int64_t)td[i].M * td[i].k * td[i].n;
            if (cnt) gen_jac_keys<<<blocks_for(cnt, 256), 256, 0, st>>>(td[i], f->d, f->n_outputs, keys_a.p, pay_a.p);
        }
        size_t tmp_bytes = 0;
        cub::DoubleBuffer<int64_t> kb(keys_a.p, keys_b.p);
        cub::DoubleBuffer<int32_t> pb(pay_a.p, pay_b.p);
        TAD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, pb, (int)nC, 0, 64, st));
        DevBuf<unsigned char> tmp;
        TAD_CUDA(tmp.ensure(tmp_bytes));
        TAD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kb, pb, (int)nC, 0, 64, st));
        count_valid<<<1, 1, 0, st>>>(kb.Current(), nC, nvalid_d.p);
        TAD_CUDA(cudaMemcpyAsync(&n_valid, nvalid_d.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        TAD_CUDA(cudaStreamSynchronize(st));
        TAD_CUDA(f->inner.ensure((size_t)std::max<int64_t>(n_valid, 1)));
        if (n_valid > 0)
            fill_jac<<<blocks_for(n_valid, 256), 256, 0, st>>>(kb.Current(), pb.Current(), n_valid, f->terms_dev.p, (int)f->terms.size(),
                                                               f->n_outputs, f->inner.p);
        jac_col_ptr<<<blocks_for(f->n_vars + 1, 256), 256, 0, st>>>(kb.Current(), n_valid, f->n_vars, f->n_outputs, f->outer.p);
        TAD_CUDA(cudaGetLastError());
        TAD_CUDA(cudaStreamSynchronize(st));
    }
    else
    {
        TAD_CUDA(cudaMemsetAsync(f->outer.p, 0, ((size_t)f->n_vars + 1) * sizeof(int32_t), st));
        TAD_CUDA(f->inner.ensure(1));
        TAD_CUDA(cudaStreamSynchronize(st));
    }
    f->nnz = n_valid;
    f->n_outer = f->n_vars;
    f->pattern_built = true;
    return TAD_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU: vertex ownership, halo blocks, exchange lists.  Collective: every rank of the communicator runs
// this at the same point (the first pattern query / evaluation after tad_function_set_comm).
// ---------------------------------------------------------------------------------------------
__global__ void mark_touched(const int32_t* __restrict__ rec, int N, int64_t stride, int64_t n, int rank, int32_t* __restrict__ owner)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    for (int j = 0; j < N; ++j)
    {
        const int32_t v = rec[(int64_t)j * stride + e];
        if (v >= 0) atomicMin(&owner[v], rank);
    }
}

// CSR position of the d x d blocks `keys` (vi * n_handles + vj): value index of entry (0,0) and distance between its rows.
__global__ void lookup_blocks(const int64_t* __restrict__ keys, int64_t n, const int64_t* __restrict__ block_key, int64_t n_blocks,
                              const int64_t* __restrict__ vrow, int64_t n_handles, int d, int32_t* __restrict__ base, int32_t* __restrict__ rs,
                              int32_t* __restrict__ missing)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t key = keys[i];
    int64_t lo = 0, hi = n_blocks;
    while (lo < hi)
    {
        const int64_t mid = (lo + hi) / 2;
        if (block_key[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (lo >= n_blocks || block_key[lo] != key) { atomicAdd(missing, 1); base[i] = 0; rs[i] = 0; return; }
    const int64_t vi = key / n_handles;
    const int64_t r0 = vrow[vi], deg = vrow[vi + 1] - r0;
    base[i] = (int32_t)((int64_t)d * d * r0 + (int64_t)d * (lo - r0));
    rs[i] = (int32_t)(d * deg);
}

int build_halo_plan(tad_function f)
{
    HaloPlan& P = f->halo;
    tad_comm c = f->comm;
    const int W = c->world, R = c->rank, d = f->d;
    const int64_t nh = f->n_handles;
    cudaStream_t st = f->stream;
    // 1. ownership: lowest rank that touches the vertex
    TAD_CUDA(P.owner.ensure((size_t)nh));
    fill_i32<<<blocks_for(nh, 256), 256, 0, st>>>(P.owner.p, nh, W);
    for (const Term& t : f->terms)
        if (t.n > 0) mark_touched<<<blocks_for(t.n, 256), 256, 0, st>>>(t.rec_handles.p, t.N, t.stride, t.n, R, P.owner.p);
    TAD_CUDA(cudaGetLastError());
    TAD_TRY(comm_allreduce_min_i32(c, P.owner.p, nh, st));
    P.owner_host.assign((size_t)nh, W);
    TAD_CUDA(cudaMemcpyAsync(P.owner_host.data(), P.owner.p, (size_t)nh * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    // 2. blocks of the local pattern whose row vertex is owned elsewhere, by destination (block keys are sorted)
    std::vector<int64_t> bk((size_t)f->n_blocks);
    if (f->n_blocks) TAD_CUDA(cudaMemcpyAsync(bk.data(), f->block_key.p, (size_t)f->n_blocks * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    TAD_CUDA(cudaStreamSynchronize(st));
    std::vector<std::vector<int64_t>> send_keys((size_t)W);
    for (int64_t key : bk)
    {
        const int32_t o = P.owner_host[(size_t)(key / nh)];
        if (o != R && o < W) send_keys[(size_t)o].push_back(key);
    }
    // 3. who sends how many blocks to whom
    std::vector<int64_t> cnt((size_t)W, 0), all((size_t)W * W, 0);
    for (int p = 0; p < W; ++p) cnt[(size_t)p] = (int64_t)send_keys[(size_t)p].size();
    DevBuf<int64_t> cnt_d, all_d;
    TAD_CUDA(cnt_d.ensure((size_t)W)); TAD_CUDA(all_d.ensure((size_t)W * W));
    TAD_CUDA(cudaMemcpyAsync(cnt_d.p, cnt.data(), (size_t)W * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    TAD_TRY(comm_allgather_i64(c, cnt_d.p, all_d.p, W, st));
    TAD_CUDA(cudaMemcpyAsync(all.data(), all_d.p, (size_t)W * W * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    TAD_CUDA(cudaStreamSynchronize(st));
    P.send_blk_off.assign((size_t)W + 1, 0); P.recv_blk_off.assign((size_t)W + 1, 0);
    for (int p = 0; p < W; ++p)
    {
        P.send_blk_off[(size_t)p + 1] = P.send_blk_off[(size_t)p] + cnt[(size_t)p];
        P.recv_blk_off[(size_t)p + 1] = P.recv_blk_off[(size_t)p] + (p == R ? 0 : all[(size_t)p * W + R]);
    }
    const int64_t ns = P.send_blk_off[(size_t)W], nr = P.recv_blk_off[(size_t)W];
    // 4. the keys themselves
    std::vector<int64_t> sk((size_t)ns), rk((size_t)nr);
    for (int p = 0; p < W; ++p) std::copy(send_keys[(size_t)p].begin(), send_keys[(size_t)p].end(), sk.begin() + P.send_blk_off[(size_t)p]);
    DevBuf<int64_t> sk_d, rk_d;
    TAD_CUDA(sk_d.ensure((size_t)std::max<int64_t>(ns, 1))); TAD_CUDA(rk_d.ensure((size_t)std::max<int64_t>(nr, 1)));
    if (ns) TAD_CUDA(cudaMemcpyAsync(sk_d.p, sk.data(), (size_t)ns * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    TAD_TRY(comm_group_begin());
    TAD_TRY(comm_exchange(c, sk_d.p, P.send_blk_off.data(), rk_d.p, P.recv_blk_off.data(), (int)sizeof(int64_t), st));
    TAD_TRY(comm_group_end());
    if (nr) TAD_CUDA(cudaMemcpyAsync(rk.data(), rk_d.p, (size_t)nr * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    TAD_CUDA(cudaStreamSynchronize(st));
    for (int64_t key : rk)
        if (key < 0 || key / nh >= nh || P.owner_host[(size_t)(key / nh)] != R) return fail(TAD_COMM_ERROR, "halo exchange: received a block of a row this rank does not own");
    // 5. rebuild the pattern with structural slots for the received blocks: owned rows now have the columns of the 1-rank pattern
    P.keys_from_peers = rk;
    TAD_TRY(build_pattern_scalar(f));
    // 6. positions of the exchanged blocks in the local CSR; gradient entries travel with the DIAGONAL blocks' vertices
    TAD_CUDA(P.send_base.ensure((size_t)std::max<int64_t>(ns, 1))); TAD_CUDA(P.send_rs.ensure((size_t)std::max<int64_t>(ns, 1)));
    TAD_CUDA(P.recv_base.ensure((size_t)std::max<int64_t>(nr, 1))); TAD_CUDA(P.recv_rs.ensure((size_t)std::max<int64_t>(nr, 1)));
    DevBuf<int32_t> missing;
    TAD_CUDA(missing.ensure(1));
    TAD_CUDA(cudaMemsetAsync(missing.p, 0, sizeof(int32_t), st));
    if (ns) lookup_blocks<<<blocks_for(ns, 256), 256, 0, st>>>(sk_d.p, ns, f->block_key.p, f->n_blocks, f->vrow.p, nh, d, P.send_base.p, P.send_rs.p, missing.p);
    if (nr) lookup_blocks<<<blocks_for(nr, 256), 256, 0, st>>>(rk_d.p, nr, f->block_key.p, f->n_blocks, f->vrow.p, nh, d, P.recv_base.p, P.recv_rs.p, missing.p);
    int32_t n_missing = 0;
    TAD_CUDA(cudaMemcpyAsync(&n_missing, missing.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    std::vector<int32_t> sv, rv;
    P.send_vtx_off.assign((size_t)W + 1, 0); P.recv_vtx_off.assign((size_t)W + 1, 0);
    int64_t min_recv_vertex = INT64_MAX;
    for (int p = 0; p < W; ++p)
    {
        for (int64_t i = P.send_blk_off[(size_t)p]; i < P.send_blk_off[(size_t)p + 1]; ++i)
            if (sk[(size_t)i] / nh == sk[(size_t)i] % nh) sv.push_back((int32_t)(sk[(size_t)i] / nh));
        for (int64_t i = P.recv_blk_off[(size_t)p]; i < P.recv_blk_off[(size_t)p + 1]; ++i)
        {
            const int64_t vi = rk[(size_t)i] / nh;
            min_recv_vertex = std::min(min_recv_vertex, vi);
            if (vi == rk[(size_t)i] % nh) rv.push_back((int32_t)vi);
        }
        P.send_vtx_off[(size_t)p + 1] = (int64_t)sv.size();
        P.recv_vtx_off[(size_t)p + 1] = (int64_t)rv.size();
    }
    TAD_CUDA(P.send_vtx.ensure(std::max<size_t>(sv.size(), 1))); TAD_CUDA(P.recv_vtx.ensure(std::max<size_t>(rv.size(), 1)));
    if (!sv.empty()) TAD_CUDA(cudaMemcpyAsync(P.send_vtx.p, sv.data(), sv.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (!rv.empty()) TAD_CUDA(cudaMemcpyAsync(P.recv_vtx.p, rv.data(), rv.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    TAD_CUDA(cudaStreamSynchronize(st));
    if (n_missing) return fail(TAD_COMM_ERROR, "halo exchange: a block is missing from the local pattern");
    const int64_t dd = (int64_t)d * d;
    P.send_h_off.resize((size_t)W + 1); P.recv_h_off.resize((size_t)W + 1); P.send_g_off.resize((size_t)W + 1); P.recv_g_off.resize((size_t)W + 1);
    for (int p = 0; p <= W; ++p)
    {
        P.send_h_off[(size_t)p] = dd * P.send_blk_off[(size_t)p]; P.recv_h_off[(size_t)p] = dd * P.recv_blk_off[(size_t)p];
        P.send_g_off[(size_t)p] = d * P.send_vtx_off[(size_t)p]; P.recv_g_off[(size_t)p] = d * P.recv_vtx_off[(size_t)p];
    }
    TAD_CUDA(P.send_h.ensure((size_t)std::max<int64_t>(dd * ns, 1))); TAD_CUDA(P.recv_h.ensure((size_t)std::max<int64_t>(dd * nr, 1)));
    TAD_CUDA(P.send_g.ensure(std::max<size_t>(d * sv.size(), 1))); TAD_CUDA(P.recv_g.ensure(std::max<size_t>(d * rv.size(), 1)));
    P.recv_min_value = (min_recv_vertex == INT64_MAX) ? INT64_MAX : dd * f->vrow_host[(size_t)min_recv_vertex];
    P.ready = true;
    return TAD_OK;
}

int ensure_pattern(tad_function f)
{
    if (f->pattern_built && (!f->comm || f->is_vector || f->halo.ready)) return TAD_OK;
    if (f->is_vector) return build_pattern_vector(f);
    if (f->comm)
    {
        f->halo.clear();                 // first the local pattern alone, then ownership and the blocks the peers send
        TAD_TRY(build_pattern_scalar(f));
        return build_halo_plan(f);
    }
    return build_pattern_scalar(f);
}

struct DeviceGuard
{
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Smallest variable handle touched by the elements [e_begin, e_begin + n) of a term (one block per slab).
// owner / rank (multi-GPU, owner may be null): halo[slab] = 1 if the slab touches a vertex owned by another rank.
__global__ void __launch_bounds__(256) slab_min_vertex(const int32_t* __restrict__ rec, int N, int64_t stride, const int64_t* __restrict__ begin,
                                                       const int64_t* __restrict__ count, int32_t* __restrict__ out,
                                                       const int32_t* __restrict__ owner, int rank, int32_t* __restrict__ halo)
{
    __shared__ int32_t sh[256];
    const int64_t e0 = begin[blockIdx.x], n = count[blockIdx.x];
    int32_t m = INT32_MAX;
    bool foreign = false;
    for (int64_t i = threadIdx.x; i < n; i += 256)
        for (int j = 0; j < N; ++j)
        {
            const int32_t v = rec[(int64_t)j * stride + e0 + i];
            if (v >= 0 && v < m) m = v;
            if (owner && v >= 0 && owner[v] != rank) foreign = true;
        }
    if (owner && __syncthreads_or(foreign ? 1 : 0) && threadIdx.x == 0) halo[blockIdx.x] = 1;
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if ((int)threadIdx.x < o) sh[threadIdx.x] = min(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// Default slab sizes (sweeps in profiles/r02_slab_sweep.txt for C2, profiles/r02c_sweep_c5.txt for C5 on one B200, and
// profiles/r02i_n8_*.json for the 1.27 M tets per rank of C5 on 8 GPUs): a slab costs ~0.1-0.2 ms of launch gaps and partial
// waves, so the device-resident path uses TWO slabs (one per lane; also what lets the halo exchange of a partitioned function
// hide behind the second slab) until a slab would exceed 2 M elements, which bounds the memory: 2.3 KB of staging + scratch per
// tet and lane.  The host-buffer path trades some of that for a finer-grained overlap of the D2H copies with the assembly (about
// a nineteenth of the function, 128 k .. 1 M elements: C2 131072, C5 ~557 k).  Slabs of a term are balanced (build_schedule).
int64_t default_chunk(tad_function f, bool host_path)
{
    const int64_t n = std::max<int64_t>(f->n_elements, 1);
    auto round_up = [](int64_t v, int64_t q) { return (v + q - 1) / q * q; };
    if (host_path) return std::min<int64_t>(1048576, std::max<int64_t>(131072, round_up(n / 19, 32768)));
    return std::min<int64_t>(2097152, std::max<int64_t>(524288, round_up(n / 2, 65536)));
}

int64_t effective_chunk(tad_function f, bool whole_terms, bool host_path)
{
    if (whole_terms || f->chunk < 0) return -1;
    static const int64_t env_chunk = [] { const char* e = getenv("TAD_CHUNK_ELEMENTS"); return e ? (int64_t)atoll(e) : (int64_t)0; }();
    int64_t c = f->chunk > 0 ? f->chunk : (env_chunk != 0 ? env_chunk : default_chunk(f, host_path));
    if (c < 0) return -1;
    return ((c + 255) / 256) * 256;  // multiples of 256: the partial sums of f do not depend on the slab size
}

// The slab schedule of an evaluation.  Atomic assembly: terms in ascending size (a small term -- e.g. a few penalty elements on
// arbitrary vertices -- would otherwise keep rows "open" until the very end), slabs of `chunk` elements; for second-order
// evaluations each slab records how many leading CSR values are final once it is complete (rows of vertices below the smallest
// vertex any LATER slab touches), which is what the host-buffer entry points copy out while later slabs are assembled.
// Gather assembly: terms in their own order, one slab per term (its kernels read every term's complete staging).
int build_schedule(tad_function f, int mode, bool whole_terms, bool host_path, tad_function_s::Schedule** out)
{
    tad_function_s::Schedule& S = f->sched[host_path ? 1 : 0];
    *out = &S;
    const int64_t chunk = effective_chunk(f, whole_terms, host_path);
    const bool cached = S.chunk == chunk && S.whole == whole_terms && (!S.slabs.empty() || f->n_elements == 0);
    if (cached && (S.has_final || (mode != TAD_MODE_SECOND && !(f->comm && f->halo.ready)))) return TAD_OK;
    if (!cached)
    {
        S.clear();
        std::vector<int> order(f->terms.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        if (!whole_terms) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return f->terms[(size_t)a].n < f->terms[(size_t)b].n; });
        for (int ti : order)
        {
            const Term& t = f->terms[(size_t)ti];
            if (t.n <= 0) continue;
            // slabs of (almost) equal size, at most `chunk` elements: no small trailing slab with its partial waves
            int64_t step = t.n;
            if (chunk > 0 && t.n > chunk)
            {
                const int64_t n_slabs = (t.n + chunk - 1) / chunk;
                step = (((t.n + n_slabs - 1) / n_slabs + 255) / 256) * 256;
            }
            for (int64_t e0 = 0; e0 < t.n; e0 += step) S.slabs.push_back(Slab{ti, e0, std::min(step, t.n - e0), 0});
        }
        S.chunk = chunk;
        S.whole = whole_terms;
    }
    const size_t ns = S.slabs.size();
    const bool partitioned = f->comm && f->halo.ready;
    if ((mode == TAD_MODE_SECOND || partitioned) && !f->is_vector && f->pattern_built && ns > 0)
    {
        std::vector<int32_t> vmin(ns, INT32_MAX), halo(ns, 0);
        if (ns > 1 || partitioned)
        {
            std::vector<int64_t> hb(ns), hc(ns);
            DevBuf<int64_t> db, dc;
            DevBuf<int32_t> dout, dhalo;
            TAD_CUDA(db.ensure(ns)); TAD_CUDA(dc.ensure(ns)); TAD_CUDA(dout.ensure(ns)); TAD_CUDA(dhalo.ensure(ns));
            for (size_t ti = 0; ti < f->terms.size(); ++ti)
            {
                // slabs of one term are contiguous in the schedule
                size_t first = ns, cnt = 0;
                for (size_t q = 0; q < ns; ++q)
                    if (S.slabs[q].term == (int)ti) { if (first == ns) first = q; hb[cnt] = S.slabs[q].e_begin; hc[cnt] = S.slabs[q].n; ++cnt; }
                if (!cnt) continue;
                const Term& t = f->terms[ti];
                TAD_CUDA(cudaMemcpyAsync(db.p, hb.data(), cnt * sizeof(int64_t), cudaMemcpyHostToDevice, f->stream));
                TAD_CUDA(cudaMemcpyAsync(dc.p, hc.data(), cnt * sizeof(int64_t), cudaMemcpyHostToDevice, f->stream));
                TAD_CUDA(cudaMemsetAsync(dhalo.p, 0, cnt * sizeof(int32_t), f->stream));
                slab_min_vertex<<<(unsigned)cnt, 256, 0, f->stream>>>(t.rec_handles.p, t.N, t.stride, db.p, dc.p, dout.p,
                                                                      partitioned ? f->halo.owner.p : nullptr, partitioned ? f->comm->rank : 0, dhalo.p);
                TAD_CUDA(cudaMemcpyAsync(vmin.data() + first, dout.p, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, f->stream));
                TAD_CUDA(cudaMemcpyAsync(halo.data() + first, dhalo.p, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, f->stream));
                TAD_CUDA(cudaStreamSynchronize(f->stream));
            }
        }
        if (partitioned && !S.has_final)
        {
            // slabs that touch halo rows go first: their contributions can travel to the owners while the rest is assembled
            std::vector<size_t> order(ns);
            for (size_t q = 0; q < ns; ++q) order[q] = q;
            std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return halo[a] > halo[b]; });
            std::vector<Slab> slabs(ns);
            std::vector<int32_t> vm(ns);
            int nh = 0;
            for (size_t q = 0; q < ns; ++q) { slabs[q] = S.slabs[order[q]]; vm[q] = vmin[order[q]]; nh += halo[order[q]] ? 1 : 0; }
            S.slabs.swap(slabs);
            vmin.swap(vm);
            S.n_halo_slabs = nh;
        }
        const int64_t dd = (int64_t)f->d * f->d;
        int64_t later = INT64_MAX;  // smallest vertex touched by the slabs after q
        for (size_t q = ns; q-- > 0;)
        {
            S.slabs[q].final_values = (later == INT64_MAX || f->vrow_host.empty()) ? f->nnz : dd * f->vrow_host[(size_t)std::min<int64_t>(later, f->n_handles)];
            if (partitioned) S.slabs[q].final_values = std::min(S.slabs[q].final_values, f->halo.recv_min_value);   // rows that receive halo values stay open
            if (vmin[q] != INT32_MAX) later = std::min<int64_t>(later, vmin[q]);
        }
        S.has_final = true;
    }
    return TAD_OK;
}

int create_lane(Lane& L, bool high_priority)
{
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);   // numerically lower = higher priority
    const int prio = high_priority ? greatest : least;
    TAD_CUDA(cudaStreamCreateWithPriority(&L.stream, cudaStreamNonBlocking, prio));
    if (cudaStreamCreateWithPriority(&L.side.stream, cudaStreamNonBlocking, prio) != cudaSuccess ||
        cudaEventCreateWithFlags(&L.side.ev_b, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&L.side.ev_list, cudaEventDisableTiming) != cudaSuccess)
        L.side = ProjSide();  // no side stream: the list kernel stays on the lane's stream
    for (auto& e : L.tev) TAD_CUDA(cudaEventCreate(&e));
    TAD_CUDA(L.counts.ensure(4));
    return TAD_OK;
}

int ensure_lanes(tad_function f, int n)
{
    while ((int)f->lanes.size() < n)
    {
        f->lanes.emplace_back();
        TAD_TRY(create_lane(f->lanes.back(), false));
    }
    return TAD_OK;
}

// Partitioned functions, opt-in (TAD_PRIORITY_LANE=1): the slabs the exchange waits for (the halo slabs, at least the first slab)
// run on a lane of their own with HIGH stream priority, so that they finish after about one slab's time instead of sharing the GPU
// with the next slab (three slabs of 1.7 M tets: 5.3 instead of 9.8 of 15 ms).
int ensure_prio_lane(tad_function f)
{
    if (!f->prio_lane.empty()) return TAD_OK;
    f->prio_lane.emplace_back();
    return create_lane(f->prio_lane.back(), true);
}

int ensure_slab_events(tad_function f, size_t n)
{
    while (f->slab_events.size() < n)
    {
        cudaEvent_t e = nullptr;
        TAD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        f->slab_events.push_back(e);
    }
    return TAD_OK;
}

// Every evaluation first waits for the work queued on the registered caller stream (tad_function_set_caller_stream).
int wait_for_caller(tad_function f)
{
    if (!f->wait_caller) return TAD_OK;
    TAD_CUDA(cudaEventRecord(f->ev_caller, f->caller_stream));
    TAD_CUDA(cudaStreamWaitEvent(f->stream, f->ev_caller, 0));
    return TAD_OK;
}

// Scalar function evaluation core.  mode: PASSIVE / FIRST / SECOND.  x, g, Hv: device pointers.  hc (optional): host
// destinations of g and the CSR values; the values are copied out slab by slab on the copy stream as their rows become final.
int eval_scalar(tad_function f, int mode, const double* x, double* f_host, double* g, double* Hv, bool project, double eps, const HostCopy* hc)
{
    if (f->is_vector) return fail(TAD_INVALID_ARGUMENT, "scalar evaluation called on a vector function");
    if (!x && f->n_vars) return fail(TAD_INVALID_ARGUMENT, "x is null");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    LaunchCounterScope counter(&f->n_launches);
    cudaStream_t st = f->stream;
    const int n_terms = (int)f->terms.size();
    if (mode == TAD_MODE_SECOND || f->comm) TAD_TRY(ensure_pattern(f));   // partitioned: ownership and exchange lists come with the pattern
    const bool partitioned = f->comm && f->halo.ready;
    TAD_TRY(wait_for_caller(f));
    TAD_CUDA(cudaMemsetAsync(f->err.p, 0, 8 * sizeof(int32_t), st));
    TAD_CUDA(f->fterm.ensure((size_t)std::max(n_terms, 1)));
    const bool gather = (f->assembly == TAD_ASSEMBLY_GATHER) && mode == TAD_MODE_SECOND;
    if (mode >= TAD_MODE_FIRST && !gather) TAD_CUDA(cudaMemsetAsync(g, 0, (size_t)f->n_vars * sizeof(double), st));
    if (mode == TAD_MODE_SECOND && !gather && f->nnz) TAD_CUDA(cudaMemsetAsync(Hv, 0, (size_t)f->nnz * sizeof(double), st));
    tad_function_s::Schedule* schedule = nullptr;
    TAD_TRY(build_schedule(f, mode, gather, hc != nullptr, &schedule));
    const std::vector<Slab>& sched = schedule->slabs;
    const int n_slabs = (int)sched.size();
    static const int env_lanes = [] { const char* e = getenv("TAD_LANES"); return e ? atoi(e) : 0; }();
    const int want_lanes = env_lanes > 0 ? std::min(env_lanes, 4) : f->n_lanes;
    const int n_lanes = std::max(1, std::min((f->timing || gather) ? 1 : want_lanes, std::max(n_slabs, 1)));
    TAD_TRY(ensure_lanes(f, n_lanes));
    TAD_TRY(ensure_slab_events(f, (size_t)n_slabs));
    // partitioned: the first n_first slabs (the halo slabs; at least one, so that all ranks reach the exchange together) on the priority lane
    // (opt-in, TAD_PRIORITY_LANE=1: it wins with three slabs per rank -- C5 on 2 GPUs 15.3 vs 16.2 ms -- but loses with two -- C5 on
    // 8 GPUs 4.28 vs 4.11 ms: two slabs interleaved on equal-priority lanes use the GPU better than one after the other)
    static const bool prio_enabled = [] { const char* e = getenv("TAD_PRIORITY_LANE"); return e && atoi(e) != 0; }();
    const int n_first = (partitioned && prio_enabled && !f->timing && !gather && n_slabs > 1 && mode >= TAD_MODE_FIRST)
                            ? std::min(n_slabs - 1, std::max(1, schedule->n_halo_slabs)) : 0;
    if (n_first > 0) TAD_TRY(ensure_prio_lane(f));
    std::vector<Lane*> active_lanes;
    for (int l = 0; l < n_lanes; ++l) active_lanes.push_back(&f->lanes[(size_t)l]);
    if (n_first > 0) active_lanes.push_back(&f->prio_lane[0]);

    // partial sums of f: one per 256 elements, laid out term by term (independent of the slab size)
    std::vector<int64_t> part_off((size_t)n_terms + 1, 0);
    for (int ti = 0; ti < n_terms; ++ti) part_off[(size_t)ti + 1] = part_off[(size_t)ti] + (f->terms[(size_t)ti].n + 255) / 256;
    TAD_CUDA(f->fpart.ensure((size_t)std::max<int64_t>(1, part_off[(size_t)n_terms])));

    TAD_CUDA(cudaEventRecord(f->ev[0], st));
    for (Lane* Lp : active_lanes)
    {
        Lane& L = *Lp;
        TAD_CUDA(cudaStreamWaitEvent(L.stream, f->ev[0], 0));
        if (mode == TAD_MODE_SECOND && project) TAD_CUDA(cudaMemsetAsync(L.counts.p, 0, 4 * sizeof(unsigned long long), L.stream));
    }
    float ms_eval = 0, ms_proj = 0, ms_asm = 0;
    for (int q = 0; q < n_slabs; ++q)
    {
        const Slab& sl = sched[(size_t)q];
        Term& t = f->terms[(size_t)sl.term];
        Lane& L = q < n_first ? f->prio_lane[0] : f->lanes[(size_t)((q - n_first) % n_lanes)];
        cudaStream_t ls = L.stream;
        const int64_t sstride = ((sl.n + 31) / 32) * 32;
        DevBuf<double>& stage = gather ? t.stage : L.stage;
        static const bool fused_enabled = [] { const char* e = getenv("TAD_FUSED_SMALL_K"); return !e || atoi(e) != 0; }();
        const bool try_fused = mode == TAD_MODE_SECOND && !gather && t.M == 0 && t.fused != 0 && fused_enabled && !(project && f->projection_full) &&
                               t.blockbase.p != nullptr;
        // a term known to run fused stages nothing but its values
        TAD_CUDA(stage.ensure(stage_doubles(t, (try_fused && t.fused == 1) ? TAD_MODE_PASSIVE : mode, sstride)));
        tad_launch_args a;
        fill_launch_args(f, t, mode, x, stage.p, sl.e_begin, sl.n, sstride, ls, a);
        if (f->timing) cudaEventRecord(L.tev[0], ls);
        // few variables per element (Double<6> triangles ...): evaluation, projection and assembly in ONE kernel of the user's
        // translation unit, nothing staged but the values (TAD_MODE_SECOND_FUSED); launchers without such a kernel decline once
        bool fused_done = false;
        if (try_fused)
        {
            tad_launch_args af = a;
            af.mode = TAD_MODE_SECOND_FUSED;
            af.blockbase = t.blockbase.p;
            af.rstride = t.rstride.p;
            af.g = g;
            af.H_values = Hv;
            af.project = project ? 1 : 0;
            af.eps = eps;
            af.counts = project ? L.counts.p : nullptr;
            const int s = t.launch(t.user, &af);
            if (s == TAD_OK) { fused_done = true; t.fused = 1; }
            else if (s == TAD_NOT_SUPPORTED && t.fused < 0) t.fused = 0;
            else return fail(s, "element kernel launch failed");
        }
        if (!fused_done)
        {
            const int s = t.launch(t.user, &a);
            if (s != TAD_OK) return fail(s, "element kernel launch failed");
        }
        if (f->timing) cudaEventRecord(L.tev[1], ls);
        count_launch();
        reduce_stage1<false><<<(unsigned)((sl.n + 255) / 256), 256, 0, ls>>>(a.val, sl.n, sstride, 1, f->fpart.p + part_off[(size_t)sl.term] + sl.e_begin / 256);
        // atomic mode + fast projection: the last projection phase (low-rank update) is fused with the scatter
        static const bool deflate_translations = [] { const char* e = getenv("TAD_DEFLATE_TRANSLATIONS"); return !e || atoi(e) != 0; }();
        // reduced pipeline for four-handle elements (tets): phases B1 / B2 / C on the complement of the translations (K - d instead of K);
        // TAD_REDUCED_PIPELINE=0 switches it off (A/B: C5 26.9 -> 23.7 ms)
        static const bool reduce_translations = [] { const char* e = getenv("TAD_REDUCED_PIPELINE"); return !e || atoi(e) != 0; }();
        const bool fuse = !fused_done && mode == TAD_MODE_SECOND && project && !gather && !f->projection_full && fused_c_assemble_supported(f->d, t.N);
        ProjScratch fused_sc;
        if (mode == TAD_MODE_SECOND && project && !fused_done)
        {
            TAD_CUDA(L.proj_list.ensure((size_t)sstride));
            TAD_CUDA(L.proj_codes.ensure((size_t)sstride));
            if (!f->projection_full || !project_has_instance(t.k))
                TAD_CUDA(L.proj_scratch.ensure(std::max<size_t>(1, project_scratch_doubles_rt(t.k, sstride))));
            TAD_CUDA(cudaMemsetAsync(L.counts.p + 2, 0, sizeof(unsigned long long), ls));  // the list of the full solver is per slab
            TAD_TRY(project_dispatch(t.k, a.hess, sl.n, sstride, eps, L.counts.p, L.proj_scratch.p, L.proj_codes.p, L.proj_list.p,
                                     f->projection_full != 0, fuse ? &fused_sc : nullptr, fuse ? &L.side : nullptr, ls,
                                     (deflate_translations && t.k == f->d * t.N) ? (f->d | ((reduce_translations && fuse && t.N == 4) ? 64 : 0)) : 0));
        }
        if (f->timing) cudaEventRecord(L.tev[2], ls);
        const SlabMaps maps{t.rec_handles.p + sl.e_begin, t.blockbase.p ? t.blockbase.p + sl.e_begin : nullptr,
                            t.rstride.p ? t.rstride.p + sl.e_begin : nullptr, t.stride};
        if (fuse)
            TAD_TRY(c_assemble(f->d, t.N, maps, a.grad, a.hess, sl.n, sstride, eps, fused_sc, g, Hv, f->err.p, L.counts.p, &L.side, ls));
        else if (mode >= TAD_MODE_FIRST && !gather && !fused_done)
            TAD_TRY(assemble_atomic(f->d, t.N, maps, a.grad, mode == TAD_MODE_SECOND ? a.hess : nullptr, sl.n, sstride, g, Hv, f->err.p, ls));
        TAD_CUDA(cudaEventRecord(f->slab_events[(size_t)q], ls));
        if (f->timing)
        {
            cudaEventRecord(L.tev[3], ls);
            cudaEventSynchronize(L.tev[3]);
            float m = 0;
            cudaEventElapsedTime(&m, L.tev[0], L.tev[1]); ms_eval += m;
            cudaEventElapsedTime(&m, L.tev[1], L.tev[2]); ms_proj += m;
            cudaEventElapsedTime(&m, L.tev[2], L.tev[3]); ms_asm += m;
        }
    }
    // multi-GPU: as soon as the slabs that touch halo rows are assembled (scheduled first), their H values and g entries travel to
    // the owners on the communication stream, and what the peers send is added into this rank's rows -- next to the assembly of
    // the remaining slabs (atomics on both sides)
    // TAD_COMM_TRACE=1 (development): timeline of the exchange of every partitioned second-order evaluation on stderr
    static const bool comm_trace_env = [] { const char* e = getenv("TAD_COMM_TRACE"); return e && atoi(e) != 0; }();
    const bool comm_trace = comm_trace_env && partitioned && mode == TAD_MODE_SECOND && !gather;
    if (comm_trace && !f->ev_trace[0])
        for (auto& e : f->ev_trace) TAD_CUDA(cudaEventCreate(&e));
    const bool halo_g = partitioned && mode >= TAD_MODE_FIRST && !f->replicate_gradient;
    const bool halo_h = partitioned && mode == TAD_MODE_SECOND;
    if ((halo_g || halo_h) && !gather)
    {
        const HaloPlan& P = f->halo;
        cudaStream_t cs = f->comm_stream;
        const int W = f->comm->world;
        TAD_CUDA(cudaStreamWaitEvent(cs, f->ev[0], 0));
        // every rank waits for its halo slabs and at least for its first slab: a rank with nothing to send would otherwise post its
        // receive at t = 0, and the spinning NCCL kernel takes SM resources from the assembly until the sender is ready
        for (int q = 0; q < std::min(std::max(schedule->n_halo_slabs, 1), n_slabs); ++q) TAD_CUDA(cudaStreamWaitEvent(cs, f->slab_events[(size_t)q], 0));
        if (comm_trace) cudaEventRecord(f->ev_trace[0], cs);   // halo slabs assembled
        TAD_TRY(halo_pack(halo_h ? Hv : nullptr, halo_g ? g : nullptr, P.send_base.p, P.send_rs.p, P.send_blk_off[(size_t)W], P.send_vtx.p,
                          P.send_vtx_off[(size_t)W], f->d, P.send_h.p, P.send_g.p, cs));
        TAD_TRY(comm_group_begin());
        if (halo_h) TAD_TRY(comm_exchange(f->comm, P.send_h.p, P.send_h_off.data(), P.recv_h.p, P.recv_h_off.data(), (int)sizeof(double), cs));
        if (halo_g) TAD_TRY(comm_exchange(f->comm, P.send_g.p, P.send_g_off.data(), P.recv_g.p, P.recv_g_off.data(), (int)sizeof(double), cs));
        if (comm_trace) cudaEventRecord(f->ev_trace[1], cs);   // packed
        TAD_TRY(comm_group_end());
        if (comm_trace) cudaEventRecord(f->ev_trace[2], cs);   // exchanged
        TAD_TRY(halo_add(halo_h ? Hv : nullptr, halo_g ? g : nullptr, P.recv_base.p, P.recv_rs.p, P.recv_blk_off[(size_t)W], P.recv_vtx.p,
                         P.recv_vtx_off[(size_t)W], f->d, P.recv_h.p, P.recv_g.p, cs));
        if (comm_trace) cudaEventRecord(f->ev_trace[3], cs);   // added
    }
    // pipelined D2H of the rows that are final (all slabs are queued by now, so a pageable destination, whose copies block the
    // host, cannot starve the GPU)
    int64_t copied = 0;
    if (hc && hc->H_host && mode == TAD_MODE_SECOND && !gather)
        for (int q = 0; q < n_slabs; ++q)
        {
            TAD_CUDA(cudaStreamWaitEvent(f->copy_stream, f->slab_events[(size_t)q], 0));
            const int64_t fin = std::min(sched[(size_t)q].final_values, f->nnz);
            if (fin > copied)
            {
                TAD_CUDA(cudaMemcpyAsync(hc->H_host + copied, Hv + copied, (size_t)(fin - copied) * sizeof(double), cudaMemcpyDeviceToHost, f->copy_stream));
                copied = fin;
            }
        }
    // the main stream continues after the last slab of every lane
    for (int q = std::max(0, n_slabs - n_lanes); q < n_slabs; ++q) TAD_CUDA(cudaStreamWaitEvent(st, f->slab_events[(size_t)q], 0));
    if (n_first > 0) TAD_CUDA(cudaStreamWaitEvent(st, f->slab_events[(size_t)n_first - 1], 0));   // ... and of the priority lane
    for (int ti = 0; ti < n_terms; ++ti)
    {
        const int64_t nb = part_off[(size_t)ti + 1] - part_off[(size_t)ti];
        if (nb > 0) { count_launch(); reduce_stage2<<<1, 1024, 0, st>>>(f->fpart.p + part_off[(size_t)ti], nb, f->fterm.p + ti); }
        else TAD_CUDA(cudaMemsetAsync(f->fterm.p + ti, 0, sizeof(double), st));
    }
    if (gather)
    {
        for (const Term& t : f->terms)
            if (t.k > 32) return fail(TAD_NOT_SUPPORTED, "gather assembly supports at most 32 variables per element");
        cudaEvent_t g0 = f->lanes[0].tev[0], g1 = f->lanes[0].tev[1];
        if (f->timing) cudaEventRecord(g0, st);
        TAD_TRY(upload_terms_dev(f, true, mode));
        TAD_TRY(gather_assemble(f->block_ptr.p, f->contrib.p, f->block_key.p, f->vrow.p, f->terms_dev.p, n_terms,
                                reinterpret_cast<const SeqTable*>(f->seqs_dev.p), f->n_blocks, f->n_handles, f->n_vars, f->d, g, Hv, f->err.p, st));
        if (f->timing)
        {
            cudaEventRecord(g1, st);
            cudaEventSynchronize(g1);
            float m = 0;
            cudaEventElapsedTime(&m, g0, g1); ms_asm += m;
        }
    }
    if (partitioned)
    {
        // the per-term sums of f (and, on request, all of g) are all-reduced on the communication stream, after the halo exchange
        const HaloPlan& P = f->halo;
        cudaStream_t cs = f->comm_stream;
        const int W = f->comm->world;
        TAD_CUDA(cudaEventRecord(f->ev_comm[0], st));
        if (comm_trace) cudaEventRecord(f->ev_trace[4], st);   // all slabs assembled, f reduced
        TAD_CUDA(cudaStreamWaitEvent(cs, f->ev_comm[0], 0));
        if ((halo_g || halo_h) && gather)
        {
            // gather assembly writes g and the CSR values at the end of the evaluation: the exchange follows it (no overlap)
            TAD_TRY(halo_pack(halo_h ? Hv : nullptr, halo_g ? g : nullptr, P.send_base.p, P.send_rs.p, P.send_blk_off[(size_t)W], P.send_vtx.p,
                              P.send_vtx_off[(size_t)W], f->d, P.send_h.p, P.send_g.p, cs));
            TAD_TRY(comm_group_begin());
            if (halo_h) TAD_TRY(comm_exchange(f->comm, P.send_h.p, P.send_h_off.data(), P.recv_h.p, P.recv_h_off.data(), (int)sizeof(double), cs));
            if (halo_g) TAD_TRY(comm_exchange(f->comm, P.send_g.p, P.send_g_off.data(), P.recv_g.p, P.recv_g_off.data(), (int)sizeof(double), cs));
            TAD_TRY(comm_group_end());
            TAD_TRY(halo_add(halo_h ? Hv : nullptr, halo_g ? g : nullptr, P.recv_base.p, P.recv_rs.p, P.recv_blk_off[(size_t)W], P.recv_vtx.p,
                             P.recv_vtx_off[(size_t)W], f->d, P.recv_h.p, P.recv_g.p, cs));
        }
        if (n_terms) TAD_TRY(comm_allreduce_sum_f64(f->comm, f->fterm.p, n_terms, cs));
        if (mode >= TAD_MODE_FIRST && f->replicate_gradient) TAD_TRY(comm_allreduce_sum_f64(f->comm, g, f->n_vars, cs));
        if (comm_trace) cudaEventRecord(f->ev_trace[5], cs);   // f all-reduced
        TAD_CUDA(cudaEventRecord(f->ev_comm[1], cs));
        TAD_CUDA(cudaStreamWaitEvent(st, f->ev_comm[1], 0));
    }
    TAD_CUDA(cudaEventRecord(f->ev[1], st));
    if (hc)
    {
        // what is left: g (complete only now) and any CSR values not final before the end
        TAD_CUDA(cudaStreamWaitEvent(f->copy_stream, f->ev[1], 0));
        if (hc->g_host && mode >= TAD_MODE_FIRST)
            TAD_CUDA(cudaMemcpyAsync(hc->g_host, g, (size_t)f->n_vars * sizeof(double), cudaMemcpyDeviceToHost, f->copy_stream));
        if (hc->H_host && mode == TAD_MODE_SECOND && f->nnz > copied)
            TAD_CUDA(cudaMemcpyAsync(hc->H_host + copied, Hv + copied, (size_t)(f->nnz - copied) * sizeof(double), cudaMemcpyDeviceToHost, f->copy_stream));
    }
    // f = sum over terms in order, with the reference's INFINITY short-circuit for eval() (ScalarFunctionImpl.hh:265-270)
    std::vector<double> fterm((size_t)std::max(n_terms, 1), 0.0);
    if (n_terms)
        TAD_CUDA(cudaMemcpyAsync(fterm.data(), f->fterm.p, (size_t)n_terms * sizeof(double), cudaMemcpyDeviceToHost, st));
    unsigned long long lane_counts[5][4] = {};
    if (mode == TAD_MODE_SECOND && project)
        for (size_t l = 0; l < active_lanes.size(); ++l)
            TAD_CUDA(cudaMemcpyAsync(lane_counts[l], active_lanes[l]->counts.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    TAD_CUDA(cudaStreamSynchronize(st));
    if (hc) TAD_CUDA(cudaStreamSynchronize(f->copy_stream));
    if (mode == TAD_MODE_SECOND && project)
    {
        f->last_proj[0] = f->last_proj[1] = f->last_proj[2] = 0;
        for (size_t l = 0; l < active_lanes.size(); ++l)
        {
            f->last_proj[0] += (int64_t)lane_counts[l][0];
            f->last_proj[1] += (int64_t)lane_counts[l][1];
            f->last_proj[2] += (int64_t)lane_counts[l][3];
        }
    }
    if (comm_trace)
    {
        float t[6] = {0, 0, 0, 0, 0, 0}, total = 0;
        for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&t[i], f->ev[0], f->ev_trace[i]);
        cudaEventElapsedTime(&total, f->ev[0], f->ev[1]);
        fprintf(stderr, "[tad comm trace] rank %d: halo slabs done %.3f | packed %.3f | exchanged %.3f | added %.3f | all slabs + f %.3f | f all-reduced %.3f | end %.3f ms (slabs %d, halo slabs %d)\n",
                f->comm->rank, t[0], t[1], t[2], t[3], t[4], t[5], total, n_slabs, schedule->n_halo_slabs);
    }
    if (f->timing)
    {
        cudaEventElapsedTime(&f->last_ms[3], f->ev[0], f->ev[1]);
        f->last_ms[0] = ms_eval; f->last_ms[1] = ms_proj; f->last_ms[2] = ms_asm;
    }
    double fv = 0.0;
    for (int ti = 0; ti < n_terms; ++ti)
    {
        if (mode == TAD_MODE_PASSIVE && fv == INFINITY) break;
        fv += fterm[(size_t)ti];
    }
    if (f_host) *f_host = fv;
    const int es = check_error_word(f, true);
    if (es == TAD_NONFINITE_DERIVATIVE && mode == TAD_MODE_PASSIVE) return TAD_OK;
    return es;
}

// First residual-Hessian block of every term in the output of tad_veval_with_derivatives (in doubles), and the total.
std::vector<int64_t> residual_hessian_offsets(tad_function f)
{
    std::vector<int64_t> off(f->terms.size() + 1, 0);
    for (size_t ti = 0; ti < f->terms.size(); ++ti)
        off[ti + 1] = off[ti] + f->terms[ti].n * f->terms[ti].M * (int64_t)f->terms[ti].k * f->terms[ti].k;
    return off;
}

int eval_vector(tad_function f, int what, const double* x, double* f_host, double* g, double* r, double* Jv, double* Hblocks = nullptr)
{
    // what: 0 eval (r), 1 jacobian (r, J), 2 sum of squares (f), 3 sum of squares with derivatives (f, g, r, J),
    //       4 derivatives (r, J, one dense k x k Hessian block per residual)
    if (!f->is_vector) return fail(TAD_INVALID_ARGUMENT, "vector evaluation called on a scalar function");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    LaunchCounterScope counter(&f->n_launches);
    cudaStream_t st = f->stream;
    TAD_TRY(ensure_pattern(f));
    TAD_TRY(wait_for_caller(f));
    const int n_terms = (int)f->terms.size();
    const int mode = what == 4 ? TAD_MODE_SECOND : ((what == 1 || what == 3) ? TAD_MODE_FIRST : TAD_MODE_PASSIVE);
    TAD_CUDA(cudaMemsetAsync(f->err.p, 0, 8 * sizeof(int32_t), st));
    TAD_CUDA(f->fterm.ensure((size_t)std::max(n_terms, 1) + 1));
    const std::vector<int64_t> hoff = residual_hessian_offsets(f);
    for (int ti = 0; ti < n_terms; ++ti)
    {
        Term& t = f->terms[ti];
        TAD_CUDA(f->stage.ensure(stage_doubles(t, mode)));
        tad_launch_args a;
        fill_launch_args(f, t, mode, x, f->stage.p, a);
        if (t.n > 0)
        {
            const int s = t.launch(t.user, &a);
            if (s != TAD_OK) return fail(s, "element kernel launch failed");
        }
        if (what == 2) TAD_TRY(sum_to(f, a.val, t.n, t.stride, t.M, true, f->fterm.p + ti));
        if (what != 2 && t.n > 0)
        {
            count_launch();
            scatter_residuals<<<blocks_for(t.n * t.M, 256), 256, 0, st>>>(a.val, t.n, t.stride, t.M, t.out_offset, r);
        }
        if (mode >= TAD_MODE_FIRST && t.n > 0)
        {
            count_launch();
            scatter_jacobian<<<blocks_for(t.n, 128), 128, 0, st>>>(a.grad, t.jslot.p, t.n, t.stride, t.M * t.k, Jv, f->err.p);
        }
        if (what == 4 && t.n > 0)
        {
            if (t.k > 32) return fail(TAD_NOT_SUPPORTED, "per-residual Hessians support at most 32 variables per element");
            SeqTable seq;
            for (int i = 0; i < t.k; ++i)
                for (int j = 0; j < t.k; ++j) seq.idx[i * t.k + j] = (int16_t)hess_seq_index(t.k, i, j);
            count_launch();
            scatter_hess_blocks<<<blocks_for(t.n * t.M * t.k * t.k, 256), 256, 0, st>>>(a.hess, t.n, t.stride, t.M, t.k, hess_size(t.k), seq,
                                                                                        Hblocks + hoff[(size_t)ti], f->err.p);
        }
    }
    double fv = 0.0;
    if (what == 2)
    {
        std::vector<double> fterm((size_t)std::max(n_terms, 1), 0.0);
        if (n_terms) TAD_CUDA(cudaMemcpyAsync(fterm.data(), f->fterm.p, (size_t)n_terms * sizeof(double), cudaMemcpyDeviceToHost, st));
        TAD_CUDA(cudaStreamSynchronize(st));
        for (int ti = 0; ti < n_terms; ++ti) fv += fterm[(size_t)ti];
    }
    else if (what == 3)
    {
        TAD_TRY(sum_to(f, r, f->n_outputs, f->n_outputs, 1, true, f->fterm.p + n_terms));
        count_launch();
        jt_r<<<blocks_for(f->n_vars, 128), 128, 0, st>>>(f->outer.p, f->inner.p, Jv, r, f->n_vars, g);
        TAD_CUDA(cudaMemcpyAsync(&fv, f->fterm.p + n_terms, sizeof(double), cudaMemcpyDeviceToHost, st));
        TAD_CUDA(cudaStreamSynchronize(st));
    }
    else
        TAD_CUDA(cudaStreamSynchronize(st));
    TAD_CUDA(cudaGetLastError());
    if (f_host) *f_host = fv;
    const int es = check_error_word(f, true);
    if (es == TAD_NONFINITE_DERIVATIVE && mode == TAD_MODE_PASSIVE) return TAD_OK;
    return es;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* tad_last_error(void) { return g_last_error.c_str(); }
void tad_set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }  // for the other translation units of this library
int tad_function_variable_dimension(tad_function f) { return f ? f->d : 0; }

int tad_device_count(int* count)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { *count = 0; return fail(TAD_CUDA_ERROR, "no CUDA device available (there is no CPU fallback)"); }
    *count = c;
    return TAD_OK;
}

int tad_function_create(int variable_dimension, int64_t n_handles, int is_vector_function, int device, tad_function* out)
{
    if (!out) return fail(TAD_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (variable_dimension < 1 || n_handles < 1) return fail(TAD_INVALID_ARGUMENT, "variable_dimension and n_handles must be >= 1");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(TAD_CUDA_ERROR, "no CUDA device available: tinyad_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail(TAD_INVALID_ARGUMENT, "bad device ordinal");
    if ((int64_t)variable_dimension * n_handles >= (int64_t)INT32_MAX) return fail(TAD_NOT_SUPPORTED, "n_vars must fit int32 (Eigen StorageIndex)");
    DeviceGuard guard(device);
    tad_function f = new tad_function_s();
    f->d = variable_dimension;
    f->n_handles = n_handles;
    f->n_vars = (int64_t)variable_dimension * n_handles;
    f->is_vector = is_vector_function != 0;
    f->device = device;
    if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess || f->err.ensure(8) != cudaSuccess ||
        cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&f->ev[0]) != cudaSuccess ||
        cudaEventCreate(&f->ev[1]) != cudaSuccess || cudaEventCreateWithFlags(&f->ev_caller, cudaEventDisableTiming) != cudaSuccess ||
        [&] { int least = 0, greatest = 0; cudaDeviceGetStreamPriorityRange(&least, &greatest);
              return cudaStreamCreateWithPriority(&f->comm_stream, cudaStreamNonBlocking, greatest); }() != cudaSuccess ||   // pack / add kernels must not queue behind the assembly
        cudaEventCreateWithFlags(&f->ev_comm[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&f->ev_comm[1], cudaEventDisableTiming) != cudaSuccess)
    {
        tad_function_destroy(f);
        return fail(TAD_CUDA_ERROR, "stream / buffer creation failed");
    }
    *out = f;
    return TAD_OK;
}

void tad_function_destroy(tad_function f)
{
    if (!f) return;
    DeviceGuard guard(f->device);
    if (f->stream) cudaStreamSynchronize(f->stream);
    for (auto& t : f->terms)
        if (t.user_free && t.user) t.user_free(t.user);
    f->terms.clear();
    for (auto* lanes : {&f->lanes, &f->prio_lane})
      for (auto& L : *lanes)
    {
        if (L.stream) { cudaStreamSynchronize(L.stream); cudaStreamDestroy(L.stream); }
        if (L.side.ev_b) cudaEventDestroy(L.side.ev_b);
        if (L.side.ev_list) cudaEventDestroy(L.side.ev_list);
        if (L.side.stream) cudaStreamDestroy(L.side.stream);
        for (auto& e : L.tev) if (e) cudaEventDestroy(e);
    }
    for (auto& e : f->slab_events) cudaEventDestroy(e);
    for (auto& e : f->ev) if (e) cudaEventDestroy(e);
    if (f->ev_caller) cudaEventDestroy(f->ev_caller);
    if (f->copy_stream) cudaStreamDestroy(f->copy_stream);
    if (f->comm_stream) { cudaStreamSynchronize(f->comm_stream); cudaStreamDestroy(f->comm_stream); }
    for (auto& e : f->ev_comm) if (e) cudaEventDestroy(e);
    for (auto& e : f->ev_trace) if (e) cudaEventDestroy(e);
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;  // device buffers are freed here, still under the device guard
}

int tad_function_set_option(tad_function f, int option, int64_t value)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "null function");
    switch (option)
    {
    case TAD_OPT_ASSEMBLY:
        if (value != TAD_ASSEMBLY_ATOMIC && value != TAD_ASSEMBLY_GATHER) return fail(TAD_INVALID_ARGUMENT, "bad assembly mode");
        f->assembly = (int)value;
        return TAD_OK;
    case TAD_OPT_CHUNK_ELEMENTS: { std::lock_guard<std::recursive_mutex> lock(f->mtx); f->chunk = value; return TAD_OK; }
    case TAD_OPT_PROJECTION: f->projection_full = value != 0; return TAD_OK;
    case TAD_OPT_LANES:
        if (value < 1 || value > 4) return fail(TAD_INVALID_ARGUMENT, "lanes must be in 1..4");
        f->n_lanes = (int)value;
        return TAD_OK;
    case TAD_OPT_REPLICATE_GRADIENT: f->replicate_gradient = value != 0; return TAD_OK;
    default: return fail(TAD_INVALID_ARGUMENT, "unknown option");
    }
}

int tad_function_set_comm(tad_function f, tad_comm c)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "null function");
    if (c && f->is_vector) return fail(TAD_NOT_SUPPORTED, "partitioned evaluation is implemented for scalar functions");
    if (c && c->device != f->device) return fail(TAD_INVALID_ARGUMENT, "the communicator lives on another device than the function");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    f->comm = c;
    f->halo.clear();
    f->pattern_built = false;   // ownership, the peers' blocks and the exchange lists are built with the pattern
    return TAD_OK;
}

int tad_function_vertex_owner(tad_function f, int32_t* owner_host)
{
    if (!f || !owner_host) return fail(TAD_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    if (!f->comm) { for (int64_t v = 0; v < f->n_handles; ++v) owner_host[v] = 0; return TAD_OK; }
    TAD_TRY(ensure_pattern(f));
    std::memcpy(owner_host, f->halo.owner_host.data(), (size_t)f->n_handles * sizeof(int32_t));
    return TAD_OK;
}

int tad_function_halo_bytes(tad_function f, int64_t* sent_per_evaluation)
{
    if (!f || !sent_per_evaluation) return fail(TAD_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    *sent_per_evaluation = 0;
    if (!f->comm) return TAD_OK;
    TAD_TRY(ensure_pattern(f));
    const size_t W = (size_t)f->comm->world;
    *sent_per_evaluation = (int64_t)sizeof(double) * (f->halo.send_h_off[W] + f->halo.send_g_off[W]);
    return TAD_OK;
}

int tad_function_get_stream(tad_function f, void** stream)
{
    if (!f || !stream) return fail(TAD_INVALID_ARGUMENT, "null argument");
    *stream = f->stream;
    return TAD_OK;
}

int tad_function_set_caller_stream(tad_function f, void* stream, int enabled)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "null function");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    f->caller_stream = static_cast<cudaStream_t>(stream);
    f->wait_caller = enabled != 0;
    return TAD_OK;
}

int tad_function_launch_count(tad_function f, int64_t* count)
{
    if (!f || !count) return fail(TAD_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    *count = f->n_launches;
    return TAD_OK;
}

int tad_function_add_term(tad_function f, int valence, int outputs_per_element, int64_t n_elements, const int64_t* elem_handles_host,
                          tad_launch_fn launch, void* user, void (*user_free)(void*))
{
    auto bail = [&](int s, const char* m) { if (user_free && user) user_free(user); return fail(s, m); };
    if (!f || !launch) return bail(TAD_INVALID_ARGUMENT, "null function or launcher");
    if (valence < 0 || n_elements < 0) return bail(TAD_INVALID_ARGUMENT, "negative valence or element count");
    if (f->is_vector != (outputs_per_element > 0)) return bail(TAD_INVALID_ARGUMENT, "outputs_per_element must be > 0 exactly for vector functions");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    Term t;
    t.N = valence; t.M = outputs_per_element; t.k = f->d * valence;
    t.n = n_elements;
    t.stride = ((n_elements + 31) / 32) * 32;
    t.launch = launch; t.user = user; t.user_free = user_free;
    auto cuda_bail = [&](cudaError_t e) { if (user_free && user) user_free(user); t.user = nullptr; return fail(TAD_CUDA_ERROR, std::string("CUDA error in add_term: ") + cudaGetErrorString(e)); };
    cudaError_t ce;
    if (elem_handles_host && n_elements > 0)
    {
        if ((ce = t.elem_handles.ensure((size_t)n_elements)) != cudaSuccess) return cuda_bail(ce);
        if ((ce = cudaMemcpy(t.elem_handles.p, elem_handles_host, (size_t)n_elements * sizeof(int64_t), cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_bail(ce);
        t.has_handles = true;
    }
    if ((ce = t.rec_handles.ensure((size_t)std::max<int64_t>(1, (int64_t)valence * t.stride))) != cudaSuccess) return cuda_bail(ce);
    if ((ce = t.rec_counts.ensure((size_t)std::max<int64_t>(1, t.stride))) != cudaSuccess) return cuda_bail(ce);
    // recording pass
    cudaMemsetAsync(f->err.p, 0, 8 * sizeof(int32_t), f->stream);
    if (valence > 0 && t.stride > 0) fill_i32<<<blocks_for((int64_t)valence * t.stride, 256), 256, 0, f->stream>>>(t.rec_handles.p, (int64_t)valence * t.stride, -1);
    tad_launch_args a;
    fill_launch_args(f, t, TAD_MODE_RECORD, nullptr, nullptr, a);
    if (n_elements > 0)
    {
        const int s = launch(user, &a);
        if (s != TAD_OK) { if (user_free && user) user_free(user); return fail(s, "record kernel launch failed"); }
        scan_counts<<<blocks_for(n_elements, 256), 256, 0, f->stream>>>(t.rec_counts.p, n_elements, f->err.p + 1);
        f->n_launches += 1;
    }
    int32_t h_err[2] = {0, 0};
    if ((ce = cudaStreamSynchronize(f->stream)) != cudaSuccess) return cuda_bail(ce);
    if ((ce = cudaMemcpy(h_err, f->err.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_bail(ce);
    if (h_err[0] & ERR_RANGE) return bail(TAD_INDEX_OUT_OF_RANGE, "variable handle out of range in element.variables(...)");
    if (h_err[0] & ERR_TOO_MANY) return bail(TAD_TOO_MANY_VARIABLES, "Too many variables requested via element.variables(...).");
    t.dedup = (h_err[1] & 1) != 0;
    f->n_elements += n_elements;
    f->n_outputs += (int64_t)outputs_per_element * n_elements;
    f->terms.push_back(std::move(t));
    f->pattern_built = false;
    f->sched[0].clear();
    f->sched[1].clear();
    return TAD_OK;
}

int tad_function_add_pattern_blocks(tad_function f, int64_t n_blocks, const int64_t* vi_host, const int64_t* vj_host)
{
    if (!f || n_blocks < 0 || (n_blocks > 0 && (!vi_host || !vj_host))) return fail(TAD_INVALID_ARGUMENT, "bad pattern blocks");
    if (f->is_vector) return fail(TAD_INVALID_ARGUMENT, "pattern blocks apply to scalar functions");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    for (int64_t i = 0; i < n_blocks; ++i)
    {
        if (vi_host[i] < 0 || vi_host[i] >= f->n_handles || vj_host[i] < 0 || vj_host[i] >= f->n_handles)
            return fail(TAD_INDEX_OUT_OF_RANGE, "pattern block handle out of range");
        f->extra_keys.push_back(vi_host[i] * f->n_handles + vj_host[i]);
    }
    f->pattern_built = false;
    return TAD_OK;
}

int64_t tad_function_n_vars(tad_function f) { return f ? f->n_vars : 0; }
int64_t tad_function_n_elements(tad_function f) { return f ? f->n_elements : 0; }
int64_t tad_function_n_outputs(tad_function f) { return f ? f->n_outputs : 0; }

int tad_function_pattern(tad_function f, int64_t* n_outer, int64_t* nnz)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "null function");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    TAD_TRY(ensure_pattern(f));
    if (n_outer) *n_outer = f->n_outer;
    if (nnz) *nnz = f->nnz;
    return TAD_OK;
}

int tad_function_pattern_copy(tad_function f, int32_t* outer_host, int32_t* inner_host)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "null function");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    TAD_TRY(ensure_pattern(f));
    if (outer_host) TAD_CUDA(cudaMemcpy(outer_host, f->outer.p, ((size_t)f->n_outer + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (inner_host && f->nnz) TAD_CUDA(cudaMemcpy(inner_host, f->inner.p, (size_t)f->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return TAD_OK;
}

int tad_function_pattern_device(tad_function f, const int32_t** outer_dev, const int32_t** inner_dev)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "null function");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    TAD_TRY(ensure_pattern(f));
    if (outer_dev) *outer_dev = f->outer.p;
    if (inner_dev) *inner_dev = f->inner.p;
    return TAD_OK;
}

int tad_function_term_table(tad_function f, int term, int32_t* handles_host)
{
    if (!f || term < 0 || term >= (int)f->terms.size() || !handles_host) return fail(TAD_INVALID_ARGUMENT, "bad term index");
    DeviceGuard guard(f->device);
    const Term& t = f->terms[(size_t)term];
    for (int j = 0; j < t.N; ++j)
        if (t.n) TAD_CUDA(cudaMemcpy(handles_host + (int64_t)j * t.n, t.rec_handles.p + (int64_t)j * t.stride, (size_t)t.n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return TAD_OK;
}

int tad_eval(tad_function f, const double* x_dev, double* f_host)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "null function");
    return eval_scalar(f, TAD_MODE_PASSIVE, x_dev, f_host, nullptr, nullptr, false, 0.0, nullptr);
}

int tad_eval_with_gradient(tad_function f, const double* x_dev, double* f_host, double* g_dev)
{
    if (!f || !g_dev) return fail(TAD_INVALID_ARGUMENT, "null argument");
    return eval_scalar(f, TAD_MODE_FIRST, x_dev, f_host, g_dev, nullptr, false, 0.0, nullptr);
}

int tad_eval_with_derivatives(tad_function f, const double* x_dev, double* f_host, double* g_dev, double* H_values_dev,
                              int project_hessian, double projection_eps)
{
    if (!f || !g_dev) return fail(TAD_INVALID_ARGUMENT, "null argument");
    return eval_scalar(f, TAD_MODE_SECOND, x_dev, f_host, g_dev, H_values_dev, project_hessian != 0, projection_eps, nullptr);
}

// Host-buffer entry points.  The function's own device buffers (x_dev / g_dev / H_dev) are shared state, so the function
// mutex is held from the H2D of x to the last D2H: concurrent calls on one function object (the reference's scenario,
// tests/ScalarFunctionTest.cc:255-291) are serialised as a whole and cannot see each other's x or results.
int tad_eval_host(tad_function f, const double* x_host, double* f_host)
{
    if (!f || !x_host) return fail(TAD_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    TAD_CUDA(f->x_dev.ensure((size_t)f->n_vars));
    TAD_CUDA(cudaMemcpyAsync(f->x_dev.p, x_host, (size_t)f->n_vars * sizeof(double), cudaMemcpyHostToDevice, f->stream));
    return eval_scalar(f, TAD_MODE_PASSIVE, f->x_dev.p, f_host, nullptr, nullptr, false, 0.0, nullptr);
}

int tad_eval_with_gradient_host(tad_function f, const double* x_host, double* f_host, double* g_host)
{
    if (!f || !x_host || !g_host) return fail(TAD_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    TAD_CUDA(f->x_dev.ensure((size_t)f->n_vars));
    TAD_CUDA(f->g_dev.ensure((size_t)f->n_vars));
    TAD_CUDA(cudaMemcpyAsync(f->x_dev.p, x_host, (size_t)f->n_vars * sizeof(double), cudaMemcpyHostToDevice, f->stream));
    HostCopy hc;
    hc.g_host = g_host;
    return eval_scalar(f, TAD_MODE_FIRST, f->x_dev.p, f_host, f->g_dev.p, nullptr, false, 0.0, &hc);
}

int tad_eval_with_derivatives_host(tad_function f, const double* x_host, double* f_host, double* g_host, double* H_values_host,
                                   int project_hessian, double projection_eps)
{
    if (!f || !x_host || !g_host) return fail(TAD_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::recursive_mutex> lock(f->mtx);
    DeviceGuard guard(f->device);
    int64_t nnz = 0;
    TAD_TRY(tad_function_pattern(f, nullptr, &nnz));
    if (nnz && !H_values_host) return fail(TAD_INVALID_ARGUMENT, "H_values_host is null");
    TAD_CUDA(f->x_dev.ensure((size_t)f->n_vars));
    TAD_CUDA(f->g_dev.ensure((size_t)f->n_vars));
    TAD_CUDA(f->H_dev.ensure((size_t)std::max<int64_t>(nnz, 1)));
    TAD_CUDA(cudaMemcpyAsync(f->x_dev.p, x_host, (size_t)f->n_vars * sizeof(double), cudaMemcpyHostToDevice, f->stream));
    HostCopy hc;
    hc.g_host = g_host;
    hc.H_host = H_values_host;
    // the D2H of the CSR values is pipelined with the evaluation inside eval_scalar
    return eval_scalar(f, TAD_MODE_SECOND, f->x_dev.p, f_host, f->g_dev.p, f->H_dev.p, project_hessian != 0, projection_eps, &hc);
}

int tad_veval(tad_function f, const double* x_dev, double* r_dev)
{
    if (!f || !r_dev) return fail(TAD_INVALID_ARGUMENT, "null argument");
    return eval_vector(f, 0, x_dev, nullptr, nullptr, r_dev, nullptr);
}
int tad_veval_with_jacobian(tad_function f, const double* x_dev, double* r_dev, double* J_values_dev)
{
    if (!f || !r_dev || !J_values_dev) return fail(TAD_INVALID_ARGUMENT, "null argument");
    return eval_vector(f, 1, x_dev, nullptr, nullptr, r_dev, J_values_dev);
}
int tad_veval_sum_of_squares(tad_function f, const double* x_dev, double* f_host)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "null argument");
    return eval_vector(f, 2, x_dev, f_host, nullptr, nullptr, nullptr);
}
int tad_veval_sum_of_squares_with_derivatives(tad_function f, const double* x_dev, double* f_host, double* g_dev, double* r_dev,
                                              double* J_values_dev)
{
    if (!f || !g_dev || !r_dev || !J_values_dev) return fail(TAD_INVALID_ARGUMENT, "null argument");
    return eval_vector(f, 3, x_dev, f_host, g_dev, r_dev, J_values_dev);
}

int tad_veval_with_derivatives(tad_function f, const double* x_dev, double* r_dev, double* J_values_dev, double* H_blocks_dev)
{
    if (!f || !r_dev || !J_values_dev || !H_blocks_dev) return fail(TAD_INVALID_ARGUMENT, "null argument");
    return eval_vector(f, 4, x_dev, nullptr, nullptr, r_dev, J_values_dev, H_blocks_dev);
}

int tad_function_residual_hessian_layout(tad_function f, int term, int64_t* offset, int* k, int64_t* n_residuals, int64_t* total)
{
    if (!f || !f->is_vector) return fail(TAD_INVALID_ARGUMENT, "residual Hessians belong to vector functions");
    const std::vector<int64_t> off = residual_hessian_offsets(f);
    if (total) *total = off.back();
    if (term < 0) return TAD_OK;
    if (term >= (int)f->terms.size()) return fail(TAD_INVALID_ARGUMENT, "bad term index");
    if (offset) *offset = off[(size_t)term];
    if (k) *k = f->terms[(size_t)term].k;
    if (n_residuals) *n_residuals = f->terms[(size_t)term].n * f->terms[(size_t)term].M;
    return TAD_OK;
}

int tad_project_batch(int k, int64_t n, int64_t stride, double* hess_dev, double eps, int method, int64_t* counts_dev, void* stream)
{
    if (k < 1 || n < 0 || stride < n || !hess_dev) return fail(TAD_INVALID_ARGUMENT, "bad projection arguments");
    // counts_dev: optional device int64[4] {decomposed, rebuilt, fallback, -}, zeroed by the caller
    DevBuf<unsigned long long> local_counts;
    DevBuf<int64_t> list;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(counts_dev);
    if (!counts)
    {
        TAD_CUDA(local_counts.ensure(4));
        TAD_CUDA(cudaMemsetAsync(local_counts.p, 0, 4 * sizeof(unsigned long long), st));
        counts = local_counts.p;
    }
    DevBuf<double> scratch;
    DevBuf<int32_t> codes;
    TAD_CUDA(list.ensure((size_t)std::max<int64_t>(stride, 1)));
    TAD_CUDA(codes.ensure((size_t)std::max<int64_t>(stride, 1)));
    if (method != 1 || !project_has_instance(k)) TAD_CUDA(scratch.ensure(std::max<size_t>(1, project_scratch_doubles_rt(k, stride))));
    TAD_TRY(project_dispatch(k, hess_dev, n, stride, eps, counts, scratch.p, codes.p, list.p, method == 1, nullptr, nullptr, st));
    TAD_CUDA(cudaStreamSynchronize(st));  // the scratch buffers above are locals
    return TAD_OK;
}

int tad_bench_fp64_peak(int device, double seconds, double* tflops)
{
    if (!tflops) return fail(TAD_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    TAD_CUDA(cudaGetDeviceProperties(&prop, device));
    DevBuf<double> out;
    TAD_CUDA(out.ensure(1));
    cudaEvent_t e0, e1;
    TAD_CUDA(cudaEventCreate(&e0));
    TAD_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    const double flop_per_launch = 2.0 * 64.0 * iters * 256.0 * blocks;
    fp64_peak_kernel<<<blocks, 256>>>(out.p, iters, 1.0000001, 1e-9);  // warm-up
    TAD_CUDA(cudaDeviceSynchronize());
    double best = 0.0, elapsed = 0.0;
    while (elapsed < seconds)
    {
        cudaEventRecord(e0);
        for (int r = 0; r < 4; ++r) fp64_peak_kernel<<<blocks, 256>>>(out.p, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        TAD_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::max(best, 4.0 * flop_per_launch / (ms * 1e-3) / 1e12);
        elapsed += ms * 1e-3;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return TAD_OK;
}

int tad_function_projection_stats(tad_function f, int64_t* stats2)
{
    if (!f || !stats2) return fail(TAD_INVALID_ARGUMENT, "null argument");
    stats2[0] = f->last_proj[0];
    stats2[1] = f->last_proj[1];
    stats2[2] = f->last_proj[2];
    return TAD_OK;
}

int tad_function_set_timing(tad_function f, int enabled)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "null function");
    f->timing = enabled != 0;
    return TAD_OK;
}

int tad_function_last_timings(tad_function f, float* ms4)
{
    if (!f || !ms4) return fail(TAD_INVALID_ARGUMENT, "null argument");
    for (int i = 0; i < 4; ++i) ms4[i] = f->last_ms[i];
    return TAD_OK;
}

}  // extern "C"
