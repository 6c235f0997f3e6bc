// tinyad_b200 runtime -- multi-GPU exchange: one process per GPU, NCCL over NVLink.
//
// The reference reduces element results into shared rows of g and H in one serial loop (Detail/ScalarObjectiveTerm.hh:256-277).
// With the elements partitioned over the ranks, that reduction crosses ranks for the rows of interface vertices: every rank
// assembles the rows of all vertices its elements touch, a vertex is OWNED by the lowest rank touching it, and the values of
// halo rows (rows of vertices owned elsewhere) and halo gradient entries are sent to the owner and added there.
//
// This file holds what does not depend on the function object: the communicator (NCCL is loaded at run time with dlopen, so a
// single-GPU user needs no NCCL at all), the collectives the pattern construction uses, and the pack / add kernels of the
// per-evaluation exchange.  The halo plan itself (who sends which blocks to whom) is built in runtime.cu with the pattern.
#include "rt_common.cuh"

#include <dlfcn.h>
#include <nccl.h>

namespace tadrt
{

namespace
{

struct NcclApi
{
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    std::string error;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // libnccl.so.2 resolves to the copy already mapped into the process (e.g. the one PyTorch ships) or to the system library
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names)
        {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define TAD_NCCL_SYM(name)                                                                   \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.lib, "nccl" #name));           \
    if (!api.name) { api.error = "libnccl.so.2 lacks nccl" #name; return; }
        TAD_NCCL_SYM(GetUniqueId) TAD_NCCL_SYM(CommInitRank) TAD_NCCL_SYM(CommDestroy) TAD_NCCL_SYM(GetErrorString) TAD_NCCL_SYM(AllReduce)
        TAD_NCCL_SYM(AllGather) TAD_NCCL_SYM(Send) TAD_NCCL_SYM(Recv) TAD_NCCL_SYM(GroupStart) TAD_NCCL_SYM(GroupEnd)
#undef TAD_NCCL_SYM
    });
    return api;
}

int nccl_ready()
{
    NcclApi& a = nccl();
    if (!a.error.empty()) return fail(TAD_COMM_ERROR, a.error);
    return TAD_OK;
}

#define TAD_NCCL(expr)                                                                                              \
    do                                                                                                              \
    {                                                                                                               \
        ncclResult_t _r = (expr);                                                                                   \
        if (_r != ncclSuccess) return fail(TAD_COMM_ERROR, std::string("NCCL error: ") + nccl().GetErrorString(_r) + " at " #expr); \
    } while (0)

__global__ void __launch_bounds__(256) pack_blocks_kernel(const double* __restrict__ Hv, const int32_t* __restrict__ base, const int32_t* __restrict__ rs,
                                                          int64_t n_blocks, int d, double* __restrict__ buf)
{
    const int dd = d * d;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_blocks * dd) return;
    const int64_t p = i / dd;
    const int ab = (int)(i % dd), a = ab / d, b = ab % d;
    buf[i] = Hv[(int64_t)base[p] + (int64_t)a * rs[p] + b];
}

__global__ void __launch_bounds__(256) add_blocks_kernel(double* __restrict__ Hv, const int32_t* __restrict__ base, const int32_t* __restrict__ rs,
                                                         int64_t n_blocks, int d, const double* __restrict__ buf)
{
    const int dd = d * d;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_blocks * dd) return;
    const int64_t p = i / dd;
    const int ab = (int)(i % dd), a = ab / d, b = ab % d;
    // atomic: the local assembly of later slabs may be adding to the same entries at the same time
    atomicAdd(&Hv[(int64_t)base[p] + (int64_t)a * rs[p] + b], buf[i]);
}

__global__ void __launch_bounds__(256) pack_vertices_kernel(const double* __restrict__ g, const int32_t* __restrict__ vtx, int64_t n_vtx, int d,
                                                            double* __restrict__ buf)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vtx * d) return;
    buf[i] = g[(int64_t)d * vtx[i / d] + (i % d)];
}

__global__ void __launch_bounds__(256) add_vertices_kernel(double* __restrict__ g, const int32_t* __restrict__ vtx, int64_t n_vtx, int d,
                                                           const double* __restrict__ buf)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vtx * d) return;
    atomicAdd(&g[(int64_t)d * vtx[i / d] + (i % d)], buf[i]);
}

}  // namespace

int comm_allreduce_min_i32(tad_comm c, int32_t* buf_dev, int64_t n, cudaStream_t st)
{
    TAD_TRY(nccl_ready());
    TAD_NCCL(nccl().AllReduce(buf_dev, buf_dev, (size_t)n, ncclInt32, ncclMin, static_cast<ncclComm_t>(c->nccl), st));
    return TAD_OK;
}

int comm_allreduce_sum_f64(tad_comm c, double* buf_dev, int64_t n, cudaStream_t st)
{
    TAD_TRY(nccl_ready());
    TAD_NCCL(nccl().AllReduce(buf_dev, buf_dev, (size_t)n, ncclDouble, ncclSum, static_cast<ncclComm_t>(c->nccl), st));
    return TAD_OK;
}

int comm_allgather_i64(tad_comm c, const int64_t* send_dev, int64_t* recv_dev, int64_t n_per_rank, cudaStream_t st)
{
    TAD_TRY(nccl_ready());
    TAD_NCCL(nccl().AllGather(send_dev, recv_dev, (size_t)n_per_rank, ncclInt64, static_cast<ncclComm_t>(c->nccl), st));
    return TAD_OK;
}

// One grouped neighbour exchange: segment [off[p], off[p+1]) of sendbuf goes to rank p, segment [roff[p], roff[p+1]) of recvbuf
// comes from rank p (units: elements of `bytes_per_element` bytes).  Several exchanges can share one NCCL group (begin / end).
int comm_group_begin()
{
    TAD_TRY(nccl_ready());
    TAD_NCCL(nccl().GroupStart());
    return TAD_OK;
}
int comm_group_end()
{
    TAD_NCCL(nccl().GroupEnd());
    return TAD_OK;
}
int comm_exchange(tad_comm c, const void* sendbuf, const int64_t* off, void* recvbuf, const int64_t* roff, int bytes_per_element, cudaStream_t st)
{
    ncclComm_t nc = static_cast<ncclComm_t>(c->nccl);
    for (int p = 0; p < c->world; ++p)
    {
        if (p == c->rank) continue;
        const int64_t ns = off[p + 1] - off[p], nr = roff[p + 1] - roff[p];
        if (ns > 0)
            TAD_NCCL(nccl().Send(static_cast<const char*>(sendbuf) + off[p] * bytes_per_element, (size_t)(ns * bytes_per_element), ncclChar, p, nc, st));
        if (nr > 0)
            TAD_NCCL(nccl().Recv(static_cast<char*>(recvbuf) + roff[p] * bytes_per_element, (size_t)(nr * bytes_per_element), ncclChar, p, nc, st));
    }
    return TAD_OK;
}

int halo_pack(const double* Hv, const double* g, const int32_t* base, const int32_t* rs, int64_t n_blocks, const int32_t* vtx, int64_t n_vtx, int d,
              double* bufH, double* bufG, cudaStream_t st)
{
    if (Hv && n_blocks > 0)
    {
        count_launch();
        pack_blocks_kernel<<<blocks_for(n_blocks * d * d, 256), 256, 0, st>>>(Hv, base, rs, n_blocks, d, bufH);
    }
    if (g && n_vtx > 0)
    {
        count_launch();
        pack_vertices_kernel<<<blocks_for(n_vtx * d, 256), 256, 0, st>>>(g, vtx, n_vtx, d, bufG);
    }
    return cudaGetLastError() == cudaSuccess ? TAD_OK : fail(TAD_CUDA_ERROR, "halo pack launch failed");
}

int halo_add(double* Hv, double* g, const int32_t* base, const int32_t* rs, int64_t n_blocks, const int32_t* vtx, int64_t n_vtx, int d,
             const double* bufH, const double* bufG, cudaStream_t st)
{
    if (Hv && n_blocks > 0)
    {
        count_launch();
        add_blocks_kernel<<<blocks_for(n_blocks * d * d, 256), 256, 0, st>>>(Hv, base, rs, n_blocks, d, bufH);
    }
    if (g && n_vtx > 0)
    {
        count_launch();
        add_vertices_kernel<<<blocks_for(n_vtx * d, 256), 256, 0, st>>>(g, vtx, n_vtx, d, bufG);
    }
    return cudaGetLastError() == cudaSuccess ? TAD_OK : fail(TAD_CUDA_ERROR, "halo add launch failed");
}

}  // namespace tadrt

using namespace tadrt;

extern "C" {

int tad_comm_unique_id(void* id_out)
{
    if (!id_out) return fail(TAD_INVALID_ARGUMENT, "null argument");
    TAD_TRY(nccl_ready());
    static_assert(sizeof(ncclUniqueId) == TAD_COMM_ID_BYTES, "TAD_COMM_ID_BYTES must match ncclUniqueId");
    ncclUniqueId id;
    TAD_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof(id));
    return TAD_OK;
}

int tad_comm_create(const void* id, int rank, int world, int device, tad_comm* out)
{
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(TAD_INVALID_ARGUMENT, "bad communicator arguments");
    if (world > 64) return fail(TAD_NOT_SUPPORTED, "at most 64 ranks");
    TAD_TRY(nccl_ready());
    int prev = -1;
    TAD_CUDA(cudaGetDevice(&prev));
    TAD_CUDA(cudaSetDevice(device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t nc = nullptr;
    const ncclResult_t r = nccl().CommInitRank(&nc, world, uid, rank);
    cudaSetDevice(prev);
    if (r != ncclSuccess) return fail(TAD_COMM_ERROR, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
    tad_comm c = new tad_comm_s();
    c->nccl = nc; c->rank = rank; c->world = world; c->device = device; c->owned = true;
    *out = c;
    return TAD_OK;
}

int tad_comm_adopt(void* nccl_comm, int rank, int world, int device, tad_comm* out)
{
    if (!nccl_comm || !out || world < 1 || rank < 0 || rank >= world) return fail(TAD_INVALID_ARGUMENT, "bad communicator arguments");
    if (world > 64) return fail(TAD_NOT_SUPPORTED, "at most 64 ranks");
    TAD_TRY(nccl_ready());
    tad_comm c = new tad_comm_s();
    c->nccl = nccl_comm; c->rank = rank; c->world = world; c->device = device; c->owned = false;
    *out = c;
    return TAD_OK;
}

void tad_comm_destroy(tad_comm c)
{
    if (!c) return;
    if (c->owned && c->nccl && nccl().CommDestroy) nccl().CommDestroy(static_cast<ncclComm_t>(c->nccl));
    delete c;
}

int tad_comm_rank(tad_comm c) { return c ? c->rank : -1; }
int tad_comm_world(tad_comm c) { return c ? c->world : 0; }

}  // extern "C"
