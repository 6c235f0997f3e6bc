// Development tool: FP64 RED throughput of the L2 for the access shapes of the assembly kernel.
//   mode 0: every lane its own pseudo-random double (what thread = element does)
//   mode 1: groups of 3 consecutive lanes hit 3 consecutive doubles (one 3-wide block row per group)
//   mode 2: groups of 9 lanes hit a 3 x 3 block (rows `rs` doubles apart)
//   mode 3: like 0 but plain stores (upper bound of the request path without the read-modify-write)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// window > 0: the targets of a thread block lie in a window of `window` doubles that slides with the block index (L2-resident,
// like the real assembly where consecutive elements hit nearby CSR rows); window == 0: anywhere in the array (DRAM-bound)
template <int MODE>
__global__ void __launch_bounds__(128) k(double* out, int64_t n_dbl, int per_thread, int rs, int window)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t span = window > 0 ? (uint32_t)window : (uint32_t)(n_dbl - 3 * rs - 8);
    const int64_t base = window > 0 ? (int64_t)((double)blockIdx.x / gridDim.x * (double)(n_dbl - window - 3 * rs - 8)) : 0;
    for (int i = 0; i < per_thread; ++i)
    {
        int64_t idx;
        if (MODE == 0 || MODE == 3) idx = base + hash(tid * 977u + i) % span;
        else if (MODE == 1) idx = base + (int64_t)(hash((tid / 3) * 977u + i) % span) + (lane % 3);
        else { const uint32_t g = tid / 9, w = lane % 9; idx = base + (int64_t)(hash(g * 977u + i) % span) + (w / 3) * rs + (w % 3); }
        if (MODE == 3) out[idx] = 1.0;
        else atomicAdd(out + idx, 1.0);
    }
}
int main()
{
    const int64_t n_dbl = 23036814;  // nnz of C2
    double* out; cudaMalloc(&out, n_dbl * 8); cudaMemset(out, 0, n_dbl * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 998250, per = 156;
    for (int window : {0, 1 << 16})
    for (int mode = 0; mode < 4; ++mode)
    {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep)
        {
            cudaEventRecord(e0);
            const unsigned g = (threads + 127) / 128;
            if (mode == 0) k<0><<<g, 128>>>(out, n_dbl, per, 207, window);
            if (mode == 1) k<1><<<g, 128>>>(out, n_dbl, per, 207, window);
            if (mode == 2) k<2><<<g, 128>>>(out, n_dbl, per, 207, window);
            if (mode == 3) k<3><<<g, 128>>>(out, n_dbl, per, 207, window);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("window %d mode %d: %.3f ms for %.1f M ops -> %.1f G ops/s (%s)\n", window, mode, best, threads * (double)per / 1e6, threads * (double)per / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
