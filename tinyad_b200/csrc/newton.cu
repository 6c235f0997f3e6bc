// tinyad_b200 -- the callers of the hot path: projected-Newton utilities on the device.
//
//   tad_newton_direction   <- Utils/NewtonDirection.hh:25-48   d = -(H_proj + w_identity I)^-1 g
//   tad_newton_decrement   <- Utils/NewtonDecrement.hh:20-26    -0.5 d.g
//   tad_line_search        <- Utils/LineSearch.hh:14-65         backtracking Armijo search, value-only evaluations
//   tad_gauss_newton_direction <- Utils/GaussNewtonDirection.hh:24-47  d = -(J^T J + w_identity I)^-1 J^T r, matrix-free
//   tad_pcg_solve                                               the linear solver behind tad_newton_direction
//
// The reference factorises H with Eigen::SimplicialLDLT (Utils/LinearSolver.hh:12-19), BASELINE.json names cuDSS for
// the GPU; neither a sparse direct solver for the device nor Eigen exists in this image.  The stand-in, clearly
// labelled as such, is a preconditioned conjugate gradient on the fixed CSR pattern: block-Jacobi preconditioner
// (the d x d diagonal blocks of the vertex-block matrix), CSR SpMV with 8 lanes per row, all CG scalars kept on
// the device (no host synchronisation inside an iteration; the residual is read back every few iterations).  Dot products
// are reduced in a fixed order (block partials + one finishing block), so the Newton solve is bitwise reproducible.
// It is written against the public C ABI only (pattern_device / get_stream / eval), so it is independent of runtime.cu.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "tinyad_b200.h"

extern "C" void tad_set_last_error(const char* msg);

namespace
{

int fail(int status, const std::string& msg)
{
    tad_set_last_error(msg.c_str());
    return status;
}

#define NT_CUDA(expr)                                                                                         \
    do                                                                                                        \
    {                                                                                                         \
        cudaError_t e_ = (expr);                                                                              \
        if (e_ != cudaSuccess) return fail(TAD_CUDA_ERROR, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct Buf
{
    void* p = nullptr;
    ~Buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
    template <class T> T* as() { return (T*)p; }
};

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double sm[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sm[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
    if (w == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    return v;  // valid in thread 0
}

// Inverse of the D x D diagonal blocks of A + w I (row-major, D*D doubles per block row); blocks that are not fully
// present in the pattern or are singular fall back to the scalar Jacobi entries.
template <int D>
__global__ void __launch_bounds__(128) block_jacobi(int64_t n_blocks, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                                                    const double* __restrict__ vals, double w, double* __restrict__ minv)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_blocks) return;
    double B[D][D];
    bool complete = true;
    for (int a = 0; a < D; ++a)
    {
        const int64_t row = D * v + a;
        int lo = outer[row], hi = outer[row + 1];
        const int32_t want = (int32_t)(D * v);
        while (lo < hi)  // first entry with inner >= want
        {
            const int mid = (lo + hi) >> 1;
            if (inner[mid] < want) lo = mid + 1;
            else hi = mid;
        }
        for (int b = 0; b < D; ++b)
        {
            double x = 0.0;
            bool found = false;
            for (int q = lo; q < outer[row + 1] && inner[q] <= want + b; ++q)
                if (inner[q] == want + b) { x = vals[q]; found = true; }
            if (!found && a == b) complete = false;
            B[a][b] = x + (a == b ? w : 0.0);
        }
    }
    double M[D][D];
    bool ok = complete;
    if constexpr (D == 1) { ok = ok && B[0][0] != 0.0; M[0][0] = 1.0 / B[0][0]; }
    else if constexpr (D == 2)
    {
        const double det = B[0][0] * B[1][1] - B[0][1] * B[1][0];
        ok = ok && fabs(det) > 1e-300;
        const double id = 1.0 / det;
        M[0][0] = B[1][1] * id; M[0][1] = -B[0][1] * id; M[1][0] = -B[1][0] * id; M[1][1] = B[0][0] * id;
    }
    else
    {
        const double c00 = B[1][1] * B[2][2] - B[1][2] * B[2][1], c01 = B[1][2] * B[2][0] - B[1][0] * B[2][2],
                     c02 = B[1][0] * B[2][1] - B[1][1] * B[2][0];
        const double det = B[0][0] * c00 + B[0][1] * c01 + B[0][2] * c02;
        ok = ok && fabs(det) > 1e-300;
        const double id = 1.0 / det;
        M[0][0] = c00 * id; M[1][0] = c01 * id; M[2][0] = c02 * id;
        M[0][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) * id;
        M[1][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) * id;
        M[2][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) * id;
        M[0][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) * id;
        M[1][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) * id;
        M[2][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) * id;
    }
    for (int a = 0; a < D; ++a)
        for (int b = 0; b < D; ++b)
        {
            double m = ok ? M[a][b] : ((a == b && B[a][a] > 0.0) ? 1.0 / B[a][a] : (a == b ? 1.0 : 0.0));
            if (!isfinite(m)) m = (a == b) ? 1.0 : 0.0;
            minv[(v * D + a) * D + b] = m;
        }
}

// y = (A + w I) p, pAp += p.y   (8 lanes per row)
__global__ void __launch_bounds__(256) spmv_dot(int64_t n, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                                                const double* __restrict__ vals, double w, const double* __restrict__ p,
                                                double* __restrict__ y, double* pAp)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    double acc = 0.0;
    if (row < n)
    {
        const int lo = outer[row], hi = outer[row + 1];
        for (int q = lo + sub; q < hi; q += 8) acc = fma(vals[q], p[inner[q]], acc);
    }
    acc += __shfl_down_sync(0xffffffffu, acc, 4, 8);
    acc += __shfl_down_sync(0xffffffffu, acc, 2, 8);
    acc += __shfl_down_sync(0xffffffffu, acc, 1, 8);
    double contrib = 0.0;
    if (row < n && sub == 0)
    {
        const double pr = p[row];
        acc = fma(w, pr, acc);
        y[row] = acc;
        contrib = pr * acc;
    }
    const double s = block_sum(contrib);
    if (threadIdx.x == 0) pAp[blockIdx.x] = s;  // block partial; finish_dots adds the partials in a fixed order (deterministic)
}

// Adds the block partials of up to two dot products in a fixed order: out_a = sum part_a[0..n_a), out_b = sum part_b[0..n_b).
// One block; every thread accumulates a strided subsequence, then a fixed-shape tree -- the same order in every run, so the whole
// solve is bitwise reproducible (the reference's direct solve is; tests/NewtonTest.cc:97-111 relies on it).
__global__ void __launch_bounds__(256) finish_dots(const double* part_a, int n_a, double* out_a, const double* part_b, int n_b, double* out_b)
{
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n_a; i += blockDim.x) a += part_a[i];
    for (int i = threadIdx.x; i < n_b; i += blockDim.x) b += part_b[i];
    const double ta = block_sum(a);
    const double tb = block_sum(b);
    if (threadIdx.x == 0)
    {
        *out_a = ta;
        if (out_b) *out_b = tb;
    }
}

// z = Minv r for block row v (D rows)
template <int D>
__device__ __forceinline__ void apply_minv(const double* __restrict__ minv, int64_t v, const double (&r)[D], double (&z)[D])
{
    for (int a = 0; a < D; ++a)
    {
        double s = 0.0;
        for (int b = 0; b < D; ++b) s = fma(minv[(v * D + a) * D + b], r[b], s);
        z[a] = s;
    }
}

// r = scale * b, x = 0, z = Minv r, p = z, S.rz[0] = r.z, S.rr[0] = r.r
template <int D>
__global__ void __launch_bounds__(256) pcg_init(int64_t n_blocks, const double* __restrict__ b, double scale, const double* __restrict__ minv,
                                                double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, double* __restrict__ p,
                                                double* rz, double* rr)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s_rz = 0.0, s_rr = 0.0;
    if (v < n_blocks)
    {
        double rl[D], zl[D];
        for (int a = 0; a < D; ++a) rl[a] = scale * b[v * D + a];
        apply_minv<D>(minv, v, rl, zl);
        for (int a = 0; a < D; ++a)
        {
            x[v * D + a] = 0.0;
            r[v * D + a] = rl[a];
            z[v * D + a] = zl[a];
            p[v * D + a] = zl[a];
            s_rz = fma(rl[a], zl[a], s_rz);
            s_rr = fma(rl[a], rl[a], s_rr);
        }
    }
    const double t1 = block_sum(s_rz);
    const double t2 = block_sum(s_rr);
    if (threadIdx.x == 0) { rz[blockIdx.x] = t1; rr[blockIdx.x] = t2; }  // block partials, see finish_dots
}

// alpha = rz_k / pAp_k; x += alpha p; r -= alpha y; z = Minv r; rz_{k+1} += r.z; rr_{k+1} += r.r
template <int D>
__global__ void __launch_bounds__(256) pcg_update_xr(int64_t n_blocks, const double* __restrict__ minv, const double* __restrict__ p,
                                                     const double* __restrict__ y, double* __restrict__ x, double* __restrict__ r,
                                                     double* __restrict__ z, const double* rz_k, const double* pAp_k, double* rz_next,
                                                     double* rr_next, int* breakdown)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double pap = *pAp_k;
    const double alpha = (pap > 0.0) ? *rz_k / pap : 0.0;
    if (!(pap > 0.0) && v == 0 && *rz_k != 0.0) *breakdown = 1;  // not positive definite (or NaN)
    double s_rz = 0.0, s_rr = 0.0;
    if (v < n_blocks)
    {
        double rl[D], zl[D];
        for (int a = 0; a < D; ++a)
        {
            const int64_t i = v * D + a;
            x[i] = fma(alpha, p[i], x[i]);
            rl[a] = fma(-alpha, y[i], r[i]);
            r[i] = rl[a];
        }
        apply_minv<D>(minv, v, rl, zl);
        for (int a = 0; a < D; ++a)
        {
            z[v * D + a] = zl[a];
            s_rz = fma(rl[a], zl[a], s_rz);
            s_rr = fma(rl[a], rl[a], s_rr);
        }
    }
    const double t1 = block_sum(s_rz);
    const double t2 = block_sum(s_rr);
    if (threadIdx.x == 0) { rz_next[blockIdx.x] = t1; rr_next[blockIdx.x] = t2; }  // block partials, see finish_dots
}

// beta = rz_{k+1} / rz_k; p = z + beta p
__global__ void __launch_bounds__(256) pcg_update_p(int64_t n, const double* __restrict__ z, double* __restrict__ p, const double* rz_k,
                                                    const double* rz_next)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double den = *rz_k;
    const double beta = den != 0.0 ? *rz_next / den : 0.0;
    p[i] = fma(beta, p[i], z[i]);
}

__global__ void __launch_bounds__(256) axpy_kernel(int64_t n, const double* __restrict__ x0, double s, const double* __restrict__ d,
                                                   double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fma(s, d[i], x0[i]);
}

// fixed-order two-stage dot product (deterministic): partial[blockIdx] then a single block
__global__ void __launch_bounds__(256) dot_stage1(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* partial)
{
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fma(a[i], b[i], s);
    const double t = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void __launch_bounds__(256) dot_stage2(int nb, const double* partial, double* out)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[i];
    const double t = block_sum(s);
    if (threadIdx.x == 0) *out = t;
}

unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

int device_dot(int64_t n, const double* a, const double* b, double* out_host, cudaStream_t st)
{
    const int nb = (int)std::min<int64_t>(1024, std::max<int64_t>(1, (n + 255) / 256));
    Buf part;
    NT_CUDA(part.alloc((size_t)(nb + 1) * sizeof(double)));
    dot_stage1<<<nb, 256, 0, st>>>(n, a, b, part.as<double>());
    dot_stage2<<<1, 256, 0, st>>>(nb, part.as<double>(), part.as<double>() + nb);
    NT_CUDA(cudaMemcpyAsync(out_host, part.as<double>() + nb, sizeof(double), cudaMemcpyDeviceToHost, st));
    NT_CUDA(cudaStreamSynchronize(st));
    return TAD_OK;
}

// ---- operators ----------------------------------------------------------------------------------------------------
// A + w I, A symmetric in CSR (== CSC)
struct SymCsrOp
{
    int64_t n;
    const int32_t *outer, *inner;
    const double* vals;
    double w;
    int dot_blocks() const { return (int)blocks_for(n * 8, 256); }
    // y = (A + w I) p; pAp_part[dot_blocks()] receives the block partials of p.y
    int apply(const double* p, double* y, double* pAp_part, cudaStream_t st) const
    {
        spmv_dot<<<blocks_for(n * 8, 256), 256, 0, st>>>(n, outer, inner, vals, w, p, y, pAp_part);
        return TAD_OK;
    }
    template <int D>
    int preconditioner(double* minv, cudaStream_t st) const
    {
        block_jacobi<D><<<blocks_for(n / D, 128), 128, 0, st>>>(n / D, outer, inner, vals, w, minv);
        return TAD_OK;
    }
};

// y_out = J p (column-compressed J: scatter with atomics, 8 lanes per column)
__global__ void __launch_bounds__(256) csc_scatter(int64_t n_cols, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                                                   const double* __restrict__ vals, const double* __restrict__ p, double* __restrict__ y_out)
{
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (c >= n_cols) return;
    const double pc = p[c];
    for (int q = outer[c] + (threadIdx.x & 7); q < outer[c + 1]; q += 8) atomicAdd(&y_out[inner[q]], vals[q] * pc);
}

// z = J^T y_out + w p, pAp += p.z   (8 lanes per column); p == nullptr: plain z = J^T y_out
__global__ void __launch_bounds__(256) csc_gather_dot(int64_t n_cols, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                                                      const double* __restrict__ vals, const double* __restrict__ y_out, double w,
                                                      const double* __restrict__ p, double* __restrict__ z, double* pAp)
{
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    double acc = 0.0;
    if (c < n_cols)
        for (int q = outer[c] + sub; q < outer[c + 1]; q += 8) acc = fma(vals[q], y_out[inner[q]], acc);
    acc += __shfl_down_sync(0xffffffffu, acc, 4, 8);
    acc += __shfl_down_sync(0xffffffffu, acc, 2, 8);
    acc += __shfl_down_sync(0xffffffffu, acc, 1, 8);
    double contrib = 0.0;
    if (c < n_cols && sub == 0)
    {
        if (p)
        {
            const double pc = p[c];
            acc = fma(w, pc, acc);
            contrib = pc * acc;
        }
        z[c] = acc;
    }
    if (pAp)
    {
        const double s = block_sum(contrib);
        if (threadIdx.x == 0) pAp[blockIdx.x] = s;  // block partial, see finish_dots
    }
}

// 1 / (|J e_c|^2 + w): Jacobi preconditioner of J^T J + w I
__global__ void __launch_bounds__(256) csc_colnorm_inv(int64_t n_cols, const int32_t* __restrict__ outer, const double* __restrict__ vals, double w,
                                                       double* __restrict__ minv)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    double s = w;
    for (int q = outer[c]; q < outer[c + 1]; ++q) s = fma(vals[q], vals[q], s);
    minv[c] = s > 0.0 ? 1.0 / s : 1.0;
}

// J^T J + w I without forming the product (the reference multiplies the sparse matrices, GaussNewtonDirection.hh:31)
struct NormalOp
{
    int64_t n, n_out;  // n = columns of J (variables)
    const int32_t *outer, *inner;
    const double* vals;
    double w;
    double* y_out;  // n_out doubles of scratch
    int dot_blocks() const { return (int)blocks_for(n * 8, 256); }
    // (the scatter J p accumulates with atomics: the Gauss-Newton solve is not bitwise reproducible, unlike the Newton solve)
    int apply(const double* p, double* y, double* pAp_part, cudaStream_t st) const
    {
        if (cudaMemsetAsync(y_out, 0, (size_t)n_out * sizeof(double), st) != cudaSuccess) return fail(TAD_CUDA_ERROR, "memset failed");
        csc_scatter<<<blocks_for(n * 8, 256), 256, 0, st>>>(n, outer, inner, vals, p, y_out);
        csc_gather_dot<<<blocks_for(n * 8, 256), 256, 0, st>>>(n, outer, inner, vals, y_out, w, p, y, pAp_part);
        return TAD_OK;
    }
    template <int D>
    int preconditioner(double* minv, cudaStream_t st) const
    {
        static_assert(D == 1, "scalar Jacobi only");
        csc_colnorm_inv<<<blocks_for(n, 256), 256, 0, st>>>(n, outer, vals, w, minv);
        return TAD_OK;
    }
};

template <int D, class Op>
int pcg_run(const Op& op, const double* b, double scale, double* x, double rel_tol, int max_iters, int* iters_out, double* rel_res_out,
            cudaStream_t st)
{
    const int64_t n = op.n;
    const int64_t nb = n / D;
    Buf minv, r, z, p, y, scal, flag;
    NT_CUDA(minv.alloc((size_t)nb * D * D * sizeof(double)));
    NT_CUDA(r.alloc((size_t)n * sizeof(double)));
    NT_CUDA(z.alloc((size_t)n * sizeof(double)));
    NT_CUDA(p.alloc((size_t)n * sizeof(double)));
    NT_CUDA(y.alloc((size_t)n * sizeof(double)));
    // device scalars, one slot per iteration: rz[k], rr[k], pAp[k]
    const size_t slots = (size_t)max_iters + 2;
    NT_CUDA(scal.alloc(3 * slots * sizeof(double)));
    NT_CUDA(flag.alloc(sizeof(int)));
    NT_CUDA(cudaMemsetAsync(scal.p, 0, 3 * slots * sizeof(double), st));
    NT_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    double* rz = scal.as<double>();
    double* rr = rz + slots;
    double* pAp = rr + slots;
    // block partials of the dot products (summed in a fixed order by finish_dots: deterministic)
    const int nbv = (int)blocks_for(nb, 256), nbs = op.dot_blocks();
    Buf parts;
    NT_CUDA(parts.alloc((size_t)(2 * nbv + nbs) * sizeof(double)));
    double* part_rz = parts.as<double>();
    double* part_rr = part_rz + nbv;
    double* part_pAp = part_rr + nbv;
    int s_ = op.template preconditioner<D>(minv.as<double>(), st);
    if (s_ != TAD_OK) return s_;
    pcg_init<D><<<nbv, 256, 0, st>>>(nb, b, scale, minv.as<double>(), x, r.as<double>(), z.as<double>(), p.as<double>(), part_rz, part_rr);
    finish_dots<<<1, 256, 0, st>>>(part_rz, nbv, rz, part_rr, nbv, rr);
    double rr0 = 0.0;
    NT_CUDA(cudaMemcpyAsync(&rr0, rr, sizeof(double), cudaMemcpyDeviceToHost, st));
    NT_CUDA(cudaStreamSynchronize(st));
    if (!(rr0 == rr0) || std::isinf(rr0)) return fail(TAD_SOLVER_FAILED, "Linear solve failed: right-hand side is not finite.");
    int k = 0;
    double rel = 0.0;
    if (rr0 > 0.0)
    {
        rel = 1.0;
        const int check_every = 8;
        while (k < max_iters)
        {
            s_ = op.apply(p.as<double>(), y.as<double>(), part_pAp, st);
            if (s_ != TAD_OK) return s_;
            finish_dots<<<1, 256, 0, st>>>(part_pAp, nbs, pAp + k, nullptr, 0, nullptr);
            pcg_update_xr<D><<<nbv, 256, 0, st>>>(nb, minv.as<double>(), p.as<double>(), y.as<double>(), x, r.as<double>(), z.as<double>(), rz + k,
                                                  pAp + k, part_rz, part_rr, flag.as<int>());
            finish_dots<<<1, 256, 0, st>>>(part_rz, nbv, rz + k + 1, part_rr, nbv, rr + k + 1);
            pcg_update_p<<<blocks_for(n, 256), 256, 0, st>>>(n, z.as<double>(), p.as<double>(), rz + k, rz + k + 1);
            ++k;
            if (k % check_every == 0 || k == max_iters || k <= 2)
            {
                double rrk = 0.0;
                int bd = 0;
                NT_CUDA(cudaMemcpyAsync(&rrk, rr + k, sizeof(double), cudaMemcpyDeviceToHost, st));
                NT_CUDA(cudaMemcpyAsync(&bd, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                NT_CUDA(cudaStreamSynchronize(st));
                if (bd || !(rrk == rrk)) return fail(TAD_SOLVER_FAILED, "Linear solve failed: matrix is not positive definite (CG breakdown).");
                rel = std::sqrt(rrk / rr0);
                if (rel <= rel_tol) break;
            }
        }
    }
    else
        NT_CUDA(cudaStreamSynchronize(st));
    NT_CUDA(cudaGetLastError());
    if (iters_out) *iters_out = k;
    if (rel_res_out) *rel_res_out = rel;
    if (rel > rel_tol) return fail(TAD_SOLVER_FAILED, "Linear solve failed: PCG did not reach the tolerance within max_iters.");
    return TAD_OK;
}

}  // namespace

extern "C" {

int tad_pcg_solve(int64_t n, int block_dim, const int32_t* outer_dev, const int32_t* inner_dev, const double* values_dev, double w_identity,
                  const double* b_dev, double b_scale, double* x_dev, double rel_tol, int max_iters, int* iters_out, double* rel_residual_out,
                  void* stream)
{
    if (n < 0 || !outer_dev || (!b_dev && n) || (!x_dev && n)) return fail(TAD_INVALID_ARGUMENT, "tad_pcg_solve: null argument");
    if (max_iters <= 0) max_iters = 10000;
    if (!(rel_tol > 0.0)) rel_tol = 1e-10;
    if (n == 0) { if (iters_out) *iters_out = 0; if (rel_residual_out) *rel_residual_out = 0.0; return TAD_OK; }
    cudaStream_t st = (cudaStream_t)stream;
    const SymCsrOp op{n, outer_dev, inner_dev, values_dev, w_identity};
    if (block_dim == 3 && n % 3 == 0) return pcg_run<3>(op, b_dev, b_scale, x_dev, rel_tol, max_iters, iters_out, rel_residual_out, st);
    if (block_dim == 2 && n % 2 == 0) return pcg_run<2>(op, b_dev, b_scale, x_dev, rel_tol, max_iters, iters_out, rel_residual_out, st);
    return pcg_run<1>(op, b_dev, b_scale, x_dev, rel_tol, max_iters, iters_out, rel_residual_out, st);
}

int tad_gauss_newton_direction(tad_function f, const double* r_dev, const double* J_values_dev, double w_identity, double rel_tol,
                               int max_iters, double* d_dev, int* iters_out, double* rel_residual_out)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "tad_gauss_newton_direction: null function");
    const int64_t n_out = tad_function_n_outputs(f);
    const int64_t n = tad_function_n_vars(f);
    if (n_out <= 0) return fail(TAD_INVALID_ARGUMENT, "tad_gauss_newton_direction needs a vector function with residuals");
    const int32_t *outer = nullptr, *inner = nullptr;
    int s = tad_function_pattern_device(f, &outer, &inner);
    if (s != TAD_OK) return s;
    void* stv = nullptr;
    s = tad_function_get_stream(f, &stv);
    if (s != TAD_OK) return s;
    cudaStream_t st = (cudaStream_t)stv;
    if (max_iters <= 0) max_iters = 10000;
    if (!(rel_tol > 0.0)) rel_tol = 1e-10;
    Buf y_out, rhs;
    NT_CUDA(y_out.alloc((size_t)n_out * sizeof(double)));
    NT_CUDA(rhs.alloc((size_t)n * sizeof(double)));
    // b = J^T r; solve (J^T J + w I) d = -b
    csc_gather_dot<<<blocks_for(n * 8, 256), 256, 0, st>>>(n, outer, inner, J_values_dev, r_dev, 0.0, nullptr, rhs.as<double>(), nullptr);
    const NormalOp op{n, n_out, outer, inner, J_values_dev, w_identity, y_out.as<double>()};
    s = pcg_run<1>(op, rhs.as<double>(), -1.0, d_dev, rel_tol, max_iters, iters_out, rel_residual_out, st);
    if (s != TAD_OK) return s;
    double dd = 0.0;
    s = device_dot(n, d_dev, d_dev, &dd, st);
    if (s != TAD_OK) return s;
    if (!std::isfinite(dd)) return fail(TAD_SOLVER_FAILED, "Linear solve failed: direction is not finite.");
    return TAD_OK;
}

int tad_newton_direction(tad_function f, const double* g_dev, const double* H_values_dev, double w_identity, double rel_tol, int max_iters,
                         double* d_dev, int* iters_out, double* rel_residual_out)
{
    if (!f) return fail(TAD_INVALID_ARGUMENT, "tad_newton_direction: null function");
    const int32_t *outer = nullptr, *inner = nullptr;
    int s = tad_function_pattern_device(f, &outer, &inner);
    if (s != TAD_OK) return s;
    void* st = nullptr;
    s = tad_function_get_stream(f, &st);
    if (s != TAD_OK) return s;
    const int64_t n = tad_function_n_vars(f);
    const int d = tad_function_variable_dimension(f);
    s = tad_pcg_solve(n, d, outer, inner, H_values_dev, w_identity, g_dev, -1.0, d_dev, rel_tol, max_iters, iters_out, rel_residual_out, st);
    if (s != TAD_OK) return s;
    // TINYAD_ASSERT_FINITE_MAT(d) (NewtonDirection.hh:46): d.d is finite iff every entry is
    double dd = 0.0;
    s = device_dot(n, d_dev, d_dev, &dd, (cudaStream_t)st);
    if (s != TAD_OK) return s;
    if (!std::isfinite(dd)) return fail(TAD_SOLVER_FAILED, "Linear solve failed: direction is not finite.");
    return TAD_OK;
}

int tad_newton_decrement(tad_function f, const double* d_dev, const double* g_dev, double* out_host)
{
    if (!f || !out_host) return fail(TAD_INVALID_ARGUMENT, "tad_newton_decrement: null argument");
    void* st = nullptr;
    int s = tad_function_get_stream(f, &st);
    if (s != TAD_OK) return s;
    double dg = 0.0;
    s = device_dot(tad_function_n_vars(f), d_dev, g_dev, &dg, (cudaStream_t)st);
    if (s != TAD_OK) return s;
    *out_host = -0.5 * dg;
    return TAD_OK;
}

int tad_line_search(tad_function f, const double* x0_dev, const double* d_dev, double f0, const double* g_dev, double s_max, double shrink,
                    int max_iters, double armijo_const, double* x_new_dev, double* f_new_host, double* step_host, int* n_evals)
{
    if (!f || !x_new_dev) return fail(TAD_INVALID_ARGUMENT, "tad_line_search: null argument");
    if (s_max <= 0.0) return fail(TAD_INVALID_ARGUMENT, "Max step size not positive.");  // LineSearch.hh:41-42
    void* stv = nullptr;
    int st_ = tad_function_get_stream(f, &stv);
    if (st_ != TAD_OK) return st_;
    cudaStream_t st = (cudaStream_t)stv;
    const int64_t n = tad_function_n_vars(f);
    double dg = 0.0;
    st_ = device_dot(n, d_dev, g_dev, &dg, st);
    if (st_ != TAD_OK) return st_;
    const bool try_one = s_max > 1.0;  // also try a step size of 1.0 (LineSearch.hh:44-45)
    const bool is_vector = tad_function_n_outputs(f) > 0;  // objective of a vector function: eval_sum_of_squares (GaussNewtonTest.cc:142-145)
    double s = s_max;
    int evals = 0;
    for (int i = 0; i < max_iters; ++i)
    {
        axpy_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, x0_dev, s, d_dev, x_new_dev);
        double f_new = 0.0;
        st_ = is_vector ? tad_veval_sum_of_squares(f, x_new_dev, &f_new) : tad_eval(f, x_new_dev, &f_new);
        ++evals;
        if (st_ != TAD_OK) return st_;
        if (f_new != f_new) return fail(TAD_INVALID_ARGUMENT, "line_search: objective is NaN");  // TINYAD_ASSERT_EQ(f_new, f_new), :53
        if (f_new <= f0 + armijo_const * s * dg)  // armijo_condition, :14-24
        {
            if (f_new_host) *f_new_host = f_new;
            if (step_host) *step_host = s;
            if (n_evals) *n_evals = evals;
            return TAD_OK;
        }
        if (try_one && s > 1.0 && s * shrink < 1.0) s = 1.0;
        else s *= shrink;
    }
    // "Line search couldn't find improvement": return x0 (:62-64)
    NT_CUDA(cudaMemcpyAsync(x_new_dev, x0_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    NT_CUDA(cudaStreamSynchronize(st));
    if (f_new_host) *f_new_host = f0;
    if (step_host) *step_host = 0.0;
    if (n_evals) *n_evals = evals;
    return TAD_OK;
}

}  // extern "C"
