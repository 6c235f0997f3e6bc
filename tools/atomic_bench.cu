// Development tool: FP64 RED throughput of the L2 for the access shapes of the assembly kernel.
//   mode 0: every lane its own pseudo-random double (what thread = element does)
//   mode 1: groups of 3 consecutive lanes hit 3 consecutive doubles (one 3-wide block row per group)
//   mode 2: groups of 9 lanes hit a 3 x 3 block (rows `rs` doubles apart)
//   mode 3: like 0 but plain stores (upper bound of the request path without the read-modify-write)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
template <int MODE>
__global__ void __launch_bounds__(128) k(double* out, int64_t n_dbl, int per_thread, int rs)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    for (int i = 0; i < per_thread; ++i)
    {
        int64_t idx;
        if (MODE == 0 || MODE == 3) idx = hash(tid * 977u + i) % (uint32_t)n_dbl;
        else if (MODE == 1) idx = (int64_t)(hash((tid / 3) * 977u + i) % (uint32_t)(n_dbl - 4)) + (lane % 3);
        else { const uint32_t g = tid / 9, w = lane % 9; idx = (int64_t)(hash(g * 977u + i) % (uint32_t)(n_dbl - 3 * rs - 4)) + (w / 3) * rs + (w % 3); }
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
    for (int mode = 0; mode < 4; ++mode)
    {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep)
        {
            cudaEventRecord(e0);
            const unsigned g = (threads + 127) / 128;
            if (mode == 0) k<0><<<g, 128>>>(out, n_dbl, per, 207);
            if (mode == 1) k<1><<<g, 128>>>(out, n_dbl, per, 207);
            if (mode == 2) k<2><<<g, 128>>>(out, n_dbl, per, 207);
            if (mode == 3) k<3><<<g, 128>>>(out, n_dbl, per, 207);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("mode %d: %.3f ms for %.1f M ops -> %.1f G ops/s (%s)\n", mode, best, threads * (double)per / 1e6, threads * (double)per / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
