// tinyad_b200 runtime -- declarations shared by the translation units of libtinyad_b200.so (runtime.cu: function object, pattern,
// evaluation, C ABI; projection.cu: batched PSD projection, one object per group of K; assembly.cu: scatter / gather assembly, one
// object per variable dimension; comm.cu: multi-GPU exchange).  Internal: nothing here crosses the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include <TinyAD/Detail/HessLayout.hh>
#include <TinyAD/Detail/Projection.hh>
#include <tinyad_b200.h>

using TinyAD::detail::hess_seq_index;
using TinyAD::detail::hess_seq_rc;
using TinyAD::detail::hess_size;

// One rank's handle on the GPU group (tad_comm of the C ABI); `nccl` is an ncclComm_t.
struct tad_comm_s
{
    void* nccl = nullptr;
    int rank = 0, world = 1, device = 0;
    bool owned = false;
};

namespace tadrt
{

extern thread_local std::string g_last_error;   // defined in runtime.cu

inline int fail(int status, const std::string& msg)
{
    g_last_error = msg;
    return status;
}

// Function attributes (dynamic shared-memory opt-in, carve-out) are per device: run(fn) executes fn exactly once per device
// (std::call_once: a second thread that arrives while the first is still configuring waits for it, so no launch can overtake
// the opt-in).  One object per call site; processes normally drive one GPU, but nothing here assumes it.
struct PerDeviceOnce
{
    std::once_flag flags[64];
    template <class Fn>
    void run(Fn&& fn)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { fn(); return; }
        std::call_once(flags[dev], fn);
    }
};

// Kernel-launch accounting (tad_function_launch_count): the evaluation entry points point this at the function's counter.
extern thread_local int64_t* tl_launch_counter;   // defined in runtime.cu
inline void count_launch(int n = 1) { if (tl_launch_counter) *tl_launch_counter += n; }
struct LaunchCounterScope
{
    int64_t* prev;
    explicit LaunchCounterScope(int64_t* c) : prev(tl_launch_counter) { tl_launch_counter = c; }
    ~LaunchCounterScope() { tl_launch_counter = prev; }
};

#define TAD_CUDA(expr)                                                                                   \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? TAD_OUT_OF_MEMORY : TAD_CUDA_ERROR,            \
                        std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);          \
    } while (0)

#define TAD_TRY(expr)                    \
    do                                   \
    {                                    \
        int _s = (expr);                 \
        if (_s != TAD_OK) return _s;     \
    } while (0)

constexpr int ERR_NONFINITE = 1 << TAD_NONFINITE_DERIVATIVE;
constexpr int ERR_TOO_MANY = 1 << TAD_TOO_MANY_VARIABLES;
constexpr int ERR_RANGE = 1 << TAD_INDEX_OUT_OF_RANGE;
constexpr int ERR_PATTERN = 1 << TAD_PATTERN_MISMATCH;

template <class T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; return *this; }
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t ensure(size_t count)
    {
        if (count <= n) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
};

struct Term
{
    int N = 0, M = 0, k = 0;
    int64_t n = 0, stride = 0;
    tad_launch_fn launch = nullptr;
    void* user = nullptr;
    void (*user_free)(void*) = nullptr;
    bool dedup = false;
    int fused = -1;               // TAD_MODE_SECOND_FUSED: -1 not tried yet, 0 the launcher has no such kernel, 1 available
    DevBuf<int64_t> elem_handles;
    bool has_handles = false;
    DevBuf<int32_t> rec_handles;  // [N][stride]
    DevBuf<int32_t> rec_counts;   // [n]
    // scalar functions: scatter map
    DevBuf<int32_t> blockbase;    // [N*N][stride]  CSR value index of entry (0,0) of block (bi,bj); -1 = unused
    DevBuf<int32_t> rstride;      // [N][stride]    distance between consecutive rows of that block row (= d * deg(vertex))
    int64_t contrib_offset = 0;
    // vector functions
    int64_t out_offset = 0;
    DevBuf<int32_t> jslot;        // [M*k][stride]  CSC value index; -1 = unused
    // staging (gather mode keeps one per term)
    DevBuf<double> stage;
};

struct TermDev  // device-visible description used by pattern / gather kernels
{
    int64_t off;      // first contribution id
    int64_t n, stride;
    int N, M, k;
    const int32_t* rec;
    int32_t* blockbase;
    int32_t* rstride;
    int32_t* jslot;
    int64_t out_offset;
    const double* grad;  // staging pointers (gather mode)
    const double* hess;
};

// Side stream of the fused path: the full solver of the few listed elements runs next to the fused phase C / assembly
// kernel (its cost is the serial latency of one full eigensolve, ~0.15 ms, not throughput).
struct ProjSide
{
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_b = nullptr, ev_list = nullptr;
};

// One element slab in flight: its own stream, staging, projection scratch and counters.  Consecutive slabs of an evaluation
// alternate between the lanes, so the tail of one slab's kernels overlaps the head of the next slab's, and the memory an
// evaluation needs is bounded by lanes x slab size instead of the term size.
struct Lane
{
    cudaStream_t stream = nullptr;
    ProjSide side;
    cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};  // timing mode: before element / after element / after projection / after assembly
    DevBuf<double> stage;                  // val / grad / hess of the slab, SoA with the slab's stride
    DevBuf<double> proj_scratch;           // R and W of the fast projection path
    DevBuf<int32_t> proj_codes;
    DevBuf<int64_t> proj_list;             // elements handed to the full eigensolver
    DevBuf<unsigned long long> counts;     // [4]: decomposed, rebuilt, listed in this slab, listed in this evaluation
};

// A slab of the evaluation schedule.
struct Slab
{
    int term;
    int64_t e_begin, n;
    int64_t final_values;  // leading CSR values that can no longer change once this slab and all earlier ones are complete
};

// Destination of the pipelined device -> host copies of the host-buffer entry points.
struct HostCopy
{
    double* g_host = nullptr;
    double* H_host = nullptr;
};


// ---- projection (projection.cu) ----
// Scratch of the fast projection path (structure-of-arrays over the elements of a slab).
struct ProjScratch
{
    double* R;       // [ProjLayout<K>::nR][stride]
    double* W;       // [ProjLayout<K>::nW][stride]
    int32_t* codes;  // [stride] ProjectCode per element
    int64_t* list;   // elements handed to the full solver
    int reduced = 0; // 1: elements with code bit PROJ_REDUCED_BIT went through the reduced pipeline (R / W in ProjLayout<K - D> order)
};
constexpr int kListBlocks = 148, kListThreads = 32;
template <int K> size_t project_scratch_doubles(int64_t stride);
template <int K>
int launch_project(double* hess, int64_t n, int64_t stride, double eps, unsigned long long* counts, double* scratch_d, int32_t* codes,
                   int64_t* list, bool full_only, ProjScratch* fuse_out, const ProjSide* side, cudaStream_t st, int tdim);
size_t project_scratch_doubles_generic(int k);   // k without a dedicated instantiation (run-time k Jacobi, projection.cu)
int launch_project_generic(int k, double* hess, int64_t n, int64_t stride, double eps, unsigned long long* counts, double* scratch_d, cudaStream_t st);
size_t project_scratch_doubles_rt(int k, int64_t stride);
int project_dispatch(int k, double* hess, int64_t n, int64_t stride, double eps, unsigned long long* counts, double* scratch_d,
                     int32_t* codes, int64_t* list, bool full_only, ProjScratch* fuse_out, const ProjSide* side, cudaStream_t st,
                     int tdim = 0);   // tdim: variable dimension of the term (translation null-space deflation), 0 = none

// ---- assembly (assembly.cu) ----
struct SeqTable { int16_t idx[32 * 32]; };   // packed (tile-order) index of entry (i, j), row-major k x k, k <= 32
// The scatter maps of one slab: the term's maps offset to the slab's first element (leading dimension mstride).
struct SlabMaps
{
    const int32_t* rec;
    const int32_t* blockbase;
    const int32_t* rstride;
    int64_t mstride;
};
bool fused_c_assemble_supported(int d, int N);
template <int D>
int c_assemble_d(int N, const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double eps, ProjScratch sc,
                 double* g, double* Hv, int32_t* err, const unsigned long long* counts, const ProjSide* side, cudaStream_t st);
int c_assemble(int d, int N, const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double eps, ProjScratch sc,
               double* g, double* Hv, int32_t* err, const unsigned long long* counts, const ProjSide* side, cudaStream_t st);
template <int D>
bool assemble_atomic_d(int N, const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double* g, double* Hv,
                       int32_t* err, cudaStream_t st);
int assemble_atomic(int d, int N, const SlabMaps& m, const double* grad, const double* hess, int64_t n, int64_t stride, double* g, double* Hv,
                    int32_t* err, cudaStream_t st);
int gather_assemble(const int64_t* block_ptr, const int32_t* contrib, const int64_t* block_key, const int64_t* vrow, const TermDev* terms,
                    int n_terms, const SeqTable* seqs, int64_t n_blocks, int64_t n_handles, int64_t n_vars, int d, double* g, double* Hv,
                    int32_t* err, cudaStream_t st);

// ---- multi-GPU exchange (comm.cu) ----
// Who sends which blocks / gradient entries to whom (built with the pattern, runtime.cu build_halo_plan).  Blocks and vertices are
// stored peer by peer: segment [blk_off[p], blk_off[p+1]) of the send lists goes to rank p, likewise for the receive lists.
struct HaloPlan
{
    bool ready = false;
    std::vector<int32_t> owner_host;                 // per vertex; world = untouched
    DevBuf<int32_t> owner;
    std::vector<int64_t> send_blk_off, recv_blk_off; // [world + 1], in blocks
    std::vector<int64_t> send_vtx_off, recv_vtx_off; // [world + 1], in vertices
    std::vector<int64_t> send_h_off, recv_h_off;     // the same in doubles (blocks * d * d)
    std::vector<int64_t> send_g_off, recv_g_off;     // (vertices * d)
    DevBuf<int32_t> send_base, send_rs, recv_base, recv_rs;   // CSR value index of entry (0,0) and row stride of every block
    DevBuf<int32_t> send_vtx, recv_vtx;
    DevBuf<double> send_h, recv_h, send_g, recv_g;
    std::vector<int64_t> keys_from_peers;            // blocks other ranks send here: structural slots of the local pattern
    int64_t recv_min_value = INT64_MAX;              // CSR values from this index on may receive halo contributions
    void clear() { *this = HaloPlan(); }
};
int comm_allreduce_min_i32(tad_comm c, int32_t* buf_dev, int64_t n, cudaStream_t st);
int comm_allreduce_sum_f64(tad_comm c, double* buf_dev, int64_t n, cudaStream_t st);
int comm_allgather_i64(tad_comm c, const int64_t* send_dev, int64_t* recv_dev, int64_t n_per_rank, cudaStream_t st);
int comm_group_begin();
int comm_group_end();
int comm_exchange(tad_comm c, const void* sendbuf, const int64_t* off, void* recvbuf, const int64_t* roff, int bytes_per_element, cudaStream_t st);
int halo_pack(const double* Hv, const double* g, const int32_t* base, const int32_t* rs, int64_t n_blocks, const int32_t* vtx, int64_t n_vtx, int d,
              double* bufH, double* bufG, cudaStream_t st);
int halo_add(double* Hv, double* g, const int32_t* base, const int32_t* rs, int64_t n_blocks, const int32_t* vtx, int64_t n_vtx, int d,
             const double* bufH, const double* bufG, cudaStream_t st);

inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); }

}  // namespace tadrt
