// tinyad_b200 -- element kernels (sm_100a).  These templates contain the USER's element
// functor, so they are instantiated by nvcc in the user's translation unit and handed to the
// prebuilt runtime (libtinyad_b200.so, include/tinyad_b200.h) as one launch function per term.
//
// They replace the body of the reference's parallel_for loops
//   ScalarObjectiveTerm::eval / eval_with_gradient_add / eval_with_derivatives_add
//     (include/TinyAD/Detail/ScalarObjectiveTerm.hh:162-278)
//   VectorObjectiveTerm::eval / eval_with_jacobian_add / eval_sum_of_squares
//     (include/TinyAD/Detail/VectorObjectiveTerm.hh:158-243,326-351)
// up to the per-element result; projection and assembly are the runtime's kernels.
//
// Thread mapping of the second-order kernels: thread = element (32 consecutive elements per
// warp, so every staging access is a fully coalesced 256-byte row).  For large k the packed
// Hessian is cut into NP parts and one kernel per part is launched, each instantiated for
// Scalar<k, true, NP, P>, which carries val, the gradient and 1/NP of the Hessian in registers;
// there is no divergence and no inter-thread traffic.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include <TinyAD/Detail/Projection.hh>
#include <TinyAD/Element.hh>
#include <tinyad_b200.h>

namespace TinyAD
{
namespace detail
{

// Number of cooperating warps per element for a k-variable second-order scalar.
TINYAD_HD constexpr int default_parts(int k) { return k <= 6 ? 1 : (k <= 9 ? 2 : 4); }

template <class Functor, typename = void>
struct functor_parts { static constexpr int value = 0; };
template <class Functor>
struct functor_parts<Functor, std::void_t<decltype(Functor::tinyad_parts)>> { static constexpr int value = Functor::tinyad_parts; };

TINYAD_HD TINYAD_INLINE int64_t elem_handle(const tad_launch_args& a, int64_t e) { return a.elem_handles ? a.elem_handles[e] : e; }

// This thread's position i inside the launched slab [e_begin, e_begin + n_elements); false if it has no element.
__device__ TINYAD_INLINE bool slab_index(const tad_launch_args& a, int64_t& i)
{
    i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    return i < a.n_elements;
}
// Column of element e in the recorded element -> handle table (nullptr: no table, e.g. while recording).
TINYAD_HD TINYAD_INLINE const int32_t* rec_column(const tad_launch_args& a, int64_t e) { return a.rec_handles ? a.rec_handles + e : nullptr; }

// Which part writes gradient component i: the one that owns Hessian entry (i, i) -- it needs grad[i] anyway.
template <int k, int NP>
TINYAD_HD constexpr int grad_owner(int i)
{
    const int s = hess_seq_index(k, i, i);
    int p = 0;
    while (p + 1 < NP && hess_part_begin(k, NP, p + 1) <= s) ++p;
    return p;
}

template <class Functor, int d, int N, int M>
__global__ void __launch_bounds__(128) record_kernel(Functor f, tad_launch_args a)
{
    int64_t i;
    if (!slab_index(a, i)) return;
    const int64_t e = a.e_begin + i;
    RecorderElement<d, N, M> el(elem_handle(a, e), a.n_handles, a.error_flags);
    (void)f(el);
    for (int j = 0; j < N; ++j) a.rec_handles[j * a.rec_stride + e] = (j < el.n_used) ? (int32_t)el.seen[j] : -1;
    // count < 0 marks "a handle was requested more than once" (element kernels then need the Dedup variant)
    a.rec_counts[e] = (el.n_calls != el.n_used) ? -el.n_used - 1 : el.n_used;
}

template <class Functor, int d, int N, int M, bool Dedup>
__global__ void __launch_bounds__(128) passive_kernel(Functor f, tad_launch_args a)
{
    int64_t i;
    if (!slab_index(a, i)) return;
    const int64_t e = a.e_begin + i;
    Element<d, N, M, double, false, Dedup> el(elem_handle(a, e), a.x, a.n_handles, a.error_flags, rec_column(a, e), a.rec_stride);
    if constexpr (M == 0)
    {
        const double r = f(el);
        a.val[i] = r;
    }
    else
    {
        const Vec<double, M> r = f(el);
        static_for<M>([&](auto mc) TINYAD_LAMBDA_INLINE { constexpr int m = decltype(mc)::value; a.val[m * a.stride + i] = r.a[m]; });
    }
    if (a.rec_counts) el.check_recorded_count(a.rec_counts[e]);
}

template <class Functor, int d, int N, int M, bool Dedup>
__global__ void __launch_bounds__(128) first_order_kernel(Functor f, tad_launch_args a)
{
    constexpr int k = d * N;
    using T = Scalar<k, false>;
    int64_t si;
    if (!slab_index(a, si)) return;
    const int64_t e = a.e_begin + si;
    Element<d, N, M, T, true, Dedup> el(elem_handle(a, e), a.x, a.n_handles, a.error_flags, rec_column(a, e), a.rec_stride);
    if constexpr (M == 0)
    {
        const T r = f(el);
        a.val[si] = r.val;
        static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; a.grad[i * a.stride + si] = r.grad[i]; });
    }
    else
    {
        const Vec<T, M> r = f(el);
        static_for<M>([&](auto mc) TINYAD_LAMBDA_INLINE {
            constexpr int m = decltype(mc)::value;
            a.val[m * a.stride + si] = r.a[m].val;
            static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE {
                constexpr int i = decltype(ic)::value;
                a.grad[(m * k + i) * a.stride + si] = r.a[m].grad[i];
            });
        });
    }
    if (a.rec_counts) el.check_recorded_count(a.rec_counts[e]);
}

// si: position inside the slab (staging index), e = a.e_begin + si: the element.  Only part 0 validates the requested handles.
template <class Functor, int d, int N, int NP, int P, bool Dedup>
__device__ TINYAD_INLINE void second_order_part(const Functor& f, const tad_launch_args& a, int64_t si)
{
    constexpr int k = d * N;
    using T = Scalar<k, true, NP, P>;
    const int64_t e = a.e_begin + si;
    Element<d, N, 0, T, true, Dedup> el(elem_handle(a, e), a.x, a.n_handles, a.error_flags, rec_column(a, e), a.rec_stride);   // every part addresses x through the recorded handles
    const T r = f(el);
    if constexpr (P == 0)
    {
        a.val[si] = r.val;
        if (a.rec_counts) el.check_recorded_count(a.rec_counts[e]);
    }
    static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE {
        constexpr int i = decltype(ic)::value;
        if constexpr (grad_owner<k, NP>(i) == P) a.grad[i * a.stride + si] = r.grad[i];
    });
    static_for<T::nh>([&](auto ic) TINYAD_LAMBDA_INLINE {
        constexpr int s = decltype(ic)::value;
        a.hess[(int64_t)(T::h_begin + s) * a.stride + si] = r.hess[s];
    });
}

// TAD_MODE_SECOND_FUSED: evaluation, PSD projection and assembly of one element in one thread, nothing staged but the value.
// For k <= TINYAD_FUSED_MAX_K the value, the gradient and the packed Hessian (1 + k + k(k+1)/2 doubles: 28 for Double<6>) live in
// registers from the functor call to the FP64 atomics; the projection is the same three-phase routine the staged path runs
// (Detail/Projection.hh project_element: tridiagonalise, eigenvalues, inverse iteration for the moved eigenpairs, low-rank update),
// so both paths give the same values, with a thread-local cyclic Jacobi for the few elements per million it hands back.
// Replaces ScalarObjectiveTerm.hh:242-277 (element evaluation, project_positive_definite, accumulation) in one pass.
#ifndef TINYAD_FUSED_MAX_K
#define TINYAD_FUSED_MAX_K 6
#endif
#ifndef TINYAD_FUSED_MIN_BLOCKS
#define TINYAD_FUSED_MIN_BLOCKS 2   // blocks of 128 threads per SM the register allocation aims at
#endif
template <class Functor, int d, int N, bool Dedup>
__global__ void __launch_bounds__(128, TINYAD_FUSED_MIN_BLOCKS) second_order_fused_kernel(Functor f, tad_launch_args a)
{
    constexpr int k = d * N;
    using T = Scalar<k, true, 1, 0>;
    constexpr int nh = T::nh;
    int64_t si;
    if (!slab_index(a, si)) return;
    const int64_t e = a.e_begin + si;
    Element<d, N, 0, T, true, Dedup> el(elem_handle(a, e), a.x, a.n_handles, a.error_flags, rec_column(a, e), a.rec_stride);
    const T r = f(el);
    a.val[si] = r.val;
    if (a.rec_counts) el.check_recorded_count(a.rec_counts[e]);
    double h[nh];
    static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int s = decltype(ic)::value; h[s] = r.hess[s]; });
    bool finite = true;
    static_for<nh>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int s = decltype(ic)::value; finite = finite && isfinite(h[s]); });
    if (a.project && finite)
    {
        int code = project_element<k, d>([&](int s) { return h[s]; }, [&](int s, double v) { h[s] = v; }, a.eps);
        if (code == PROJ_FALLBACK)
        {
            code = project_full_jacobi<k>([&](int s) { return h[s]; }, [&](int s, double v) { h[s] = v; }, a.eps);
            if (a.counts) atomicAdd(&a.counts[3], 1ull);
        }
        if (a.counts && code != PROJ_DOMINANT) atomicAdd(&a.counts[0], 1ull);
        if (a.counts && code == PROJ_REBUILT) atomicAdd(&a.counts[1], 1ull);
    }
    // assembly: FP64 atomics on g and the fixed CSR values through the block scatter map
    const int32_t* rec = a.rec_handles + e;
    static_for<N>([&](auto bc) TINYAD_LAMBDA_INLINE {
        constexpr int bi = decltype(bc)::value;
        const int32_t vi = rec[(int64_t)bi * a.rec_stride];
        if (vi >= 0)
            static_for<d>([&](auto ac) TINYAD_LAMBDA_INLINE {
                constexpr int c = decltype(ac)::value;
                const double v = r.grad[d * bi + c];
                finite = finite && isfinite(v);
                atomicAdd(&a.g[(int64_t)d * vi + c], v);
            });
    });
    const int32_t* bb = a.blockbase + e;
    const int32_t* rsp = a.rstride + e;
    static_for<N>([&](auto bc) TINYAD_LAMBDA_INLINE {
        constexpr int bi = decltype(bc)::value;
        const int32_t rs = rsp[(int64_t)bi * a.rec_stride];
        static_for<N>([&](auto cc) TINYAD_LAMBDA_INLINE {
            constexpr int bj = decltype(cc)::value;
            const int32_t base = bb[(int64_t)(bi * N + bj) * a.rec_stride];
            if (base >= 0)
                static_for<d * d>([&](auto qc) TINYAD_LAMBDA_INLINE {
                    constexpr int q = decltype(qc)::value, ra = q / d, cb = q % d;
                    const double v = h[hess_seq_index(k, d * bi + ra, d * bj + cb)];
                    finite = finite && isfinite(v);
                    atomicAdd(&a.H_values[(int64_t)base + (int64_t)ra * rs + cb], v);
                });
        });
    });
    if (!finite) atomicOr(a.error_flags, 1 << TAD_NONFINITE_DERIVATIVE);
}

// Per-residual second derivatives of a VectorFunction element (VectorObjectiveTerm.hh:245-324, eval_with_derivatives): every
// residual r_m of the element comes back with its gradient (the Jacobian rows) and its packed k x k Hessian.  Staging layout:
// val[m], grad[m * k + i], hess[m * k(k+1)/2 + s] (tile order), leading dimension stride.  One thread per element, so this is
// for elements with few variables (NP == 1, k <= 6); larger elements report TAD_NOT_SUPPORTED.
template <class Functor, int d, int N, int M, bool Dedup>
__global__ void __launch_bounds__(128) second_order_vector_kernel(Functor f, tad_launch_args a)
{
    constexpr int k = d * N;
    using T = Scalar<k, true, 1, 0>;
    constexpr int nh = T::nh;
    int64_t si;
    if (!slab_index(a, si)) return;
    const int64_t e = a.e_begin + si;
    Element<d, N, M, T, true, Dedup> el(elem_handle(a, e), a.x, a.n_handles, a.error_flags, rec_column(a, e), a.rec_stride);
    const Vec<T, M> r = f(el);
    static_for<M>([&](auto mc) TINYAD_LAMBDA_INLINE {
        constexpr int m = decltype(mc)::value;
        a.val[m * a.stride + si] = r.a[m].val;
        static_for<k>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            a.grad[(m * k + i) * a.stride + si] = r.a[m].grad[i];
        });
        static_for<nh>([&](auto sc) TINYAD_LAMBDA_INLINE {
            constexpr int s = decltype(sc)::value;
            a.hess[(int64_t)(m * nh + s) * a.stride + si] = r.a[m].hess[s];
        });
    });
    if (a.rec_counts) el.check_recorded_count(a.rec_counts[e]);
}

inline int check_launch();

// One kernel per Hessian part: all threads of a launch run the same instantiation
// Scalar<k, true, NP, P>; thread = element, so staging accesses are coalesced 256-byte rows.
template <class Functor, int d, int N, int NP, int P, bool Dedup>
__global__ void __launch_bounds__(128) second_order_part_kernel(Functor f, tad_launch_args a)
{
    int64_t si;
    if (!slab_index(a, si)) return;
    second_order_part<Functor, d, N, NP, P, Dedup>(f, a, si);
}

// Host-side launch of one part.  A plain function template so that a heavy functor can spread its NP
// instantiations over several translation units (`extern template` here, explicit instantiation there)
// and compile them in parallel -- see csrc/energies_tet_part.cu.
template <class Functor, int d, int N, int NP, int P, bool Dedup>
int launch_second_order_part(const Functor& f, const tad_launch_args& a);

inline int check_launch()
{
    return cudaGetLastError() == cudaSuccess ? (int)TAD_OK : (int)TAD_CUDA_ERROR;
}

template <class Functor, int d, int N, int NP, int P, bool Dedup>
int launch_second_order_part(const Functor& f, const tad_launch_args& a)
{
    cudaStream_t st = static_cast<cudaStream_t>(a.stream);
    second_order_part_kernel<Functor, d, N, NP, P, Dedup><<<(unsigned)((a.n_elements + 127) / 128), 128, 0, st>>>(f, a);
    return check_launch();
}

// A functor may declare `static constexpr bool tinyad_unique_handles = true;` to promise that no element
// requests the same handle twice; the run-time-search (Dedup) kernel variants are then not compiled.
template <class Functor, typename = void>
struct functor_unique_handles { static constexpr bool value = false; };
template <class Functor>
struct functor_unique_handles<Functor, std::void_t<decltype(Functor::tinyad_unique_handles)>> { static constexpr bool value = Functor::tinyad_unique_handles; };

}  // namespace detail

// One objective term = functor + its launch function (the type-erased LambdaImpl of
// ScalarObjectiveTerm.hh:78-115, with the three deferred instantiations done eagerly by nvcc).
// ForceDense: always use the run-time-indexed element (add_elements_dynamic: the number of handles an element touches, and
// with it the slot of each variables() call, is a run-time quantity).
template <class Functor, int d, int N, int M, bool ForceDense = false>
struct TermLauncher
{
    static constexpr int k = d * N;
    static constexpr int NP = detail::functor_parts<Functor>::value > 0 ? detail::functor_parts<Functor>::value
                                                                         : detail::default_parts(k);

    Functor f;

    static void destroy(void* user) { delete static_cast<TermLauncher*>(user); }

    template <bool Dedup>
    static int launch_eval(const TermLauncher* self, const tad_launch_args* a)
    {
        cudaStream_t st = static_cast<cudaStream_t>(a->stream);
        const int64_t n = a->n_elements;
        const unsigned g128 = (unsigned)((n + 127) / 128);
        if (a->mode == TAD_MODE_SECOND_FUSED)
        {
            if constexpr (M == 0 && NP == 1 && k <= TINYAD_FUSED_MAX_K)
            {
                if (a->launch_counter) *a->launch_counter += 1;
                detail::second_order_fused_kernel<Functor, d, N, Dedup><<<g128, 128, 0, st>>>(self->f, *a);
                return detail::check_launch();
            }
            else
                return TAD_NOT_SUPPORTED;   // the runtime falls back to TAD_MODE_SECOND + its projection / assembly kernels
        }
        if (a->launch_counter) *a->launch_counter += (a->mode == TAD_MODE_SECOND) ? NP : 1;
        switch (a->mode)
        {
        case TAD_MODE_PASSIVE:
            detail::passive_kernel<Functor, d, N, M, Dedup><<<g128, 128, 0, st>>>(self->f, *a);
            break;
        case TAD_MODE_FIRST:
            detail::first_order_kernel<Functor, d, N, M, Dedup><<<g128, 128, 0, st>>>(self->f, *a);
            break;
        case TAD_MODE_SECOND:
            if constexpr (M == 0)
            {
                int status = TAD_OK;
                detail::static_for<NP>([&](auto pc) TINYAD_LAMBDA_INLINE {
                    constexpr int P = decltype(pc)::value;
                    if (status == TAD_OK) status = detail::launch_second_order_part<Functor, d, N, NP, P, Dedup>(self->f, *a);
                });
                return status;
            }
            else if constexpr (NP == 1)
                detail::second_order_vector_kernel<Functor, d, N, M, Dedup><<<g128, 128, 0, st>>>(self->f, *a);
            else
                return TAD_NOT_SUPPORTED;  // per-residual Hessians (VectorObjectiveTerm.hh:245-324) of elements with k > 6
            break;
        default:
            return TAD_INVALID_ARGUMENT;
        }
        return detail::check_launch();
    }

    static int launch(void* user, const tad_launch_args* a)
    {
        const TermLauncher* self = static_cast<const TermLauncher*>(user);
        if (a->n_elements <= 0) return TAD_OK;
        if (a->mode == TAD_MODE_RECORD)
        {
            cudaStream_t st = static_cast<cudaStream_t>(a->stream);
            if (a->launch_counter) *a->launch_counter += 1;
            detail::record_kernel<Functor, d, N, M><<<(unsigned)((a->n_elements + 127) / 128), 128, 0, st>>>(self->f, *a);
            return detail::check_launch();
        }
        if constexpr (ForceDense) return launch_eval<true>(self, a);
        else
        {
            if (!a->dedup) return launch_eval<false>(self, a);
            if constexpr (detail::functor_unique_handles<Functor>::value) return TAD_NOT_SUPPORTED;
            else return launch_eval<true>(self, a);
        }
    }
};

}  // namespace TinyAD
