/* tinyad_b200 -- C ABI of the B200 runtime behind TinyAD's per-element derivative path.
 *
 * This is the drop-in boundary.  In the reference the seam is the
 * objective-term interface that ScalarFunction / VectorFunction call
 *   ScalarObjectiveTermBase  include/TinyAD/Detail/ScalarObjectiveTerm.hh:22-44
 *   VectorObjectiveTermBase  include/TinyAD/Detail/VectorObjectiveTerm.hh:22-45
 * and everything below it (parallel_for over elements, serial accumulation, triplets,
 * setFromTriplets).  Here everything below that seam runs on the GPU:
 *   - the element kernels that contain the user's element functor are templates instantiated
 *     by nvcc in the USER's translation unit (tinyad_b200/include/TinyAD/Kernels.cuh) and are handed
 *     to this runtime as one plain function pointer per term (tad_launch_fn);
 *   - this library (libtinyad_b200.so) is functor-independent: it owns device memory, the
 *     one-time recorded element->variable table, the fixed CSR pattern and scatter maps, the
 *     batched PSD projection, the assembly kernels and the reductions.
 * No C++ or torch types cross this boundary; all functions return a tad_status.
 * There is no CPU fallback: every entry point that computes needs a CUDA device.
 */
#ifndef TINYAD_B200_H
#define TINYAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tad_status
{
    TAD_OK = 0,
    TAD_INVALID_ARGUMENT = 1,      /* size mismatch etc.  (TINYAD_ASSERT_EQ in ScalarFunctionImpl.hh:262,292,325,388) */
    TAD_NONFINITE_DERIVATIVE = 2,  /* non-finite element gradient / Hessian (ScalarObjectiveTerm.hh:210,252-253) */
    TAD_CUDA_ERROR = 3,
    TAD_TOO_MANY_VARIABLES = 4,    /* "Too many variables requested via element.variables(...)" (Element.hh:237-238) */
    TAD_INDEX_OUT_OF_RANGE = 5,    /* variable handle outside [0, n_handles) (Element.hh:159-170) */
    TAD_NOT_SUPPORTED = 6,
    TAD_OUT_OF_MEMORY = 7,
    TAD_SOLVER_FAILED = 8,         /* "Linear solve failed." (Utils/NewtonDirection.hh:43-44) */
    TAD_PATTERN_MISMATCH = 9,      /* an element requested other variable handles at evaluation time than when it was added: the
                                      fixed pattern (recorded once; the reference rediscovers it at every evaluation,
                                      Element.hh:208-260) does not cover this x.  SURVEY.md App. E 3 */
    TAD_COMM_ERROR = 10            /* NCCL failure in the multi-GPU exchange */
} tad_status;

/* What an element kernel launch computes. */
typedef enum tad_mode
{
    TAD_MODE_RECORD = 0,  /* run the functor on a recorder element: which handles does each element touch, in first-access order (RecorderElement, ScalarFunctionImpl.hh:106-129) */
    TAD_MODE_PASSIVE = 1, /* plain double        (ScalarObjectiveTerm.hh:162-186) */
    TAD_MODE_FIRST = 2,   /* Scalar<k,false>     (ScalarObjectiveTerm.hh:188-222, VectorObjectiveTerm.hh:180-243) */
    TAD_MODE_SECOND = 3,  /* Scalar<k,true>      (ScalarObjectiveTerm.hh:224-278) */
    TAD_MODE_SECOND_FUSED = 4 /* Scalar<k,true> + PSD projection (HessianProjection.hh:48-101) + assembly (ScalarObjectiveTerm.hh:256-277)
                             in ONE kernel, nothing but `val` staged: for terms with few variables per element (k <= 6, e.g. Double<6>
                             triangles), whose value, gradient and packed Hessian fit one thread.  A launcher that has no such kernel
                             returns TAD_NOT_SUPPORTED and the runtime falls back to TAD_MODE_SECOND + its own projection / assembly kernels. */
} tad_mode;

/* Arguments of one element-kernel launch.  All pointers are device pointers.  A launch covers the element SLAB
 * [e_begin, e_begin + n_elements) of its term (TAD_OPT_CHUNK_ELEMENTS); the per-element outputs are written
 * structure-of-arrays, indexed by the position i = e - e_begin inside the slab, with leading dimension `stride`
 * (>= n_elements):
 *   val [m * stride + i]                       m < max(1, outputs_per_element)
 *   grad[(m * k + i') * stride + i]            i' < k = d * valence
 *   hess[s * stride + i]                       s < k(k+1)/2, tile order (Detail/HessLayout.hh), scalar terms only
 * The recorded element -> handle table is indexed by the element itself with leading dimension `rec_stride`:
 *   rec_handles[j * rec_stride + e], rec_counts[e]   j < valence, -1 = unused
 * It is written by TAD_MODE_RECORD and read back by the other modes, which check that the functor requests the same
 * handles in the same order as recorded (else TAD_PATTERN_MISMATCH).
 */
typedef struct tad_launch_args
{
    int32_t mode;               /* tad_mode */
    int32_t dedup;              /* 1: some element of this term touches a handle twice -> element kernels search */
    int64_t n_elements;
    int64_t stride;
    const int64_t* elem_handles; /* element handle values, NULL = identity 0..n-1 (TinyAD::range) */
    const double* x;            /* n_vars doubles, x[d * handle + i] (Element.hh:165) */
    int64_t n_handles;          /* number of variable handles (n_vars / d) */
    double* val;
    double* grad;
    double* hess;
    int32_t* rec_handles;
    int32_t* rec_counts;
    int32_t* error_flags;       /* device int32[8], OR-ed: bit per tad_status */
    void* stream;               /* cudaStream_t */
    int64_t e_begin;            /* first element of the slab */
    int64_t rec_stride;         /* leading dimension of rec_handles */
    int64_t* launch_counter;    /* HOST counter (may be NULL): the launcher adds the number of kernels it launched */
    /* TAD_MODE_SECOND_FUSED only: where the element kernel scatters to.  The maps are the term's scatter maps, indexed by the element
     * itself like rec_handles (leading dimension rec_stride): blockbase[(bi * N + bj) * rec_stride + e] = CSR value index of entry (0,0)
     * of the d x d block of vertex pair (bi, bj), -1 = unused; rstride[bi * rec_stride + e] = distance between the rows of that block row. */
    const int32_t* blockbase;
    const int32_t* rstride;
    double* g;                  /* n_vars, accumulated with FP64 atomics */
    double* H_values;           /* nnz, accumulated with FP64 atomics */
    int32_t project;            /* 1: project every element Hessian to PSD first */
    double eps;                 /* projection_eps */
    unsigned long long* counts; /* device uint64[4], may be NULL: [0] += elements decomposed, [1] += rebuilt, [3] += finished by the full solver */
} tad_launch_args;

/* Launches the element kernel of one term; returns a tad_status (TAD_CUDA_ERROR on launch failure). */
typedef int (*tad_launch_fn)(void* user, const tad_launch_args* args);

typedef struct tad_function_s* tad_function;
typedef struct tad_comm_s* tad_comm;   /* one rank's view of the group of GPUs that share a partitioned function */

/* Options (tad_function_set_option). */
#define TAD_OPT_ASSEMBLY 1      /* 0 = FP64 atomics (default), 1 = deterministic gather in element order */
#define TAD_OPT_CHUNK_ELEMENTS 2 /* elements per slab (rounded up to a multiple of 256): a term is evaluated, projected and assembled slab
                                    by slab, so staging + projection scratch are bounded by `lanes` slabs (2.3 KB per tet) instead of the
                                    whole term, consecutive slabs overlap on different streams, and the host-buffer entry points copy
                                    finished CSR rows to the host while later slabs are still being assembled.
                                    0 = default (half of the function, 512 k .. 2 M elements, for the device-pointer entry points; about a nineteenth, 128 k .. 1 M, for the host-buffer ones), < 0 = whole term in one slab.  Gather assembly always stages whole terms. */
#define TAD_OPT_PROJECTION 3    /* 0 = low-rank update via selected eigenvectors (default), 1 = full eigendecomposition */
#define TAD_OPT_LANES 4         /* number of slabs in flight (1..4, default 2) */
#define TAD_OPT_REPLICATE_GRADIENT 5 /* multi-GPU: 0 = halo-only exchange, every rank ends with the complete g entries of the vertices it
                                        OWNS (default); 1 = all-reduce, every rank ends with the complete global g */
#define TAD_ASSEMBLY_ATOMIC 0
#define TAD_ASSEMBLY_GATHER 1

const char* tad_last_error(void);
void tad_set_last_error(const char* msg); /* used by the library's own translation units */
int tad_device_count(int* count);

/* scalar_function<d>(range(n_handles)) / vector_function<d>(...)   ScalarFunctionImpl.hh:418-442, VectorFunctionImpl.hh:303-323 */
int tad_function_create(int variable_dimension, int64_t n_handles, int is_vector_function, int device, tad_function* out);
void tad_function_destroy(tad_function f);
int tad_function_set_option(tad_function f, int option, int64_t value);
/* Stream contract of the device-pointer entry points (tad_eval*, tad_veval*, tad_newton_direction, tad_line_search ...): the
 * work runs on the function's own non-blocking streams (tad_function_get_stream returns the main one) and is COMPLETE when
 * the call returns (the call synchronises), so outputs may be consumed from any stream afterwards.  Inputs (x_dev, g_dev,
 * H_values_dev ...) must be complete before the call, OR the stream that produces them is registered once with
 * tad_function_set_caller_stream: every evaluation then first waits (event record + stream wait, no host sync) for the work
 * queued on that stream at the time of the call.  stream = NULL registers the legacy default stream; call with
 * enabled = 0 to unregister. */
int tad_function_get_stream(tad_function f, void** stream);
int tad_function_set_caller_stream(tad_function f, void* stream, int enabled);

/* add_elements<N>(range, functor) / add_elements<N, M>(...)   ScalarFunctionImpl.hh:63-100, VectorFunctionImpl.hh:64-101
 * elem_handles_host may be NULL (identity).  The term keeps `user` and calls user_free(user) on destroy.
 * Runs the recording pass (TAD_MODE_RECORD) immediately. */
int tad_function_add_term(tad_function f, int valence, int outputs_per_element, int64_t n_elements,
                          const int64_t* elem_handles_host, tad_launch_fn launch, void* user, void (*user_free)(void*));

/* Multi-GPU: inject structural-only d x d blocks (handle pairs vi, vj) into the Hessian pattern, e.g. the halo-row
 * blocks other ranks will send to this rank.  They get explicit zero slots; the pattern is rebuilt on next use. */
int tad_function_add_pattern_blocks(tad_function f, int64_t n_blocks, const int64_t* vi_host, const int64_t* vj_host);

/* ---- multi-GPU: one process per GPU, the ELEMENTS are partitioned over the ranks --------------------------
 * In the reference the reduction of element results into shared rows of g and H is one serial loop (ScalarObjectiveTerm.hh:256-277);
 * here it crosses ranks for interface vertices.  Every rank creates the same function (same d, same n_handles: handle numbering stays
 * GLOBAL), adds ITS elements with the same sequence of add_term calls, and attaches the communicator before the first evaluation.
 * The runtime then
 *   - assigns every vertex to the lowest rank that touches it (ncclAllReduce MIN) -- tad_function_vertex_owner;
 *   - adds to each rank's pattern the blocks of its owned rows that other ranks contribute to, so that the rows a rank owns have
 *     exactly the columns of the single-rank pattern (bit-exact per row);
 *   - inside every evaluation, as soon as the slabs that touch halo rows are assembled (they are scheduled first), packs the
 *     halo-row H values and halo g entries, sends them to their owners (grouped ncclSend / ncclRecv over NVLink) and adds what it
 *     receives into its own rows, while the remaining slabs are still being assembled; f is all-reduced.
 * After eval_with_derivatives every rank holds: f (global), the complete CSR rows of the vertices it owns (a row-distributed
 * matrix), the complete g entries of those vertices (TAD_OPT_REPLICATE_GRADIENT = 1: all of g).  Rows / entries of vertices owned
 * elsewhere hold this rank's partial sums.  All ranks must call the same evaluation functions in the same order.
 * The communicator wraps NCCL, which is loaded at run time (dlopen libnccl.so.2): tad_comm_unique_id on one rank, the 128 bytes
 * travel to the others by the caller's means (MPI, torch.distributed, a file), tad_comm_create on every rank.  tad_comm_adopt
 * wraps an existing ncclComm_t instead (not destroyed by tad_comm_destroy). */
#define TAD_COMM_ID_BYTES 128
int tad_comm_unique_id(void* id_out /* TAD_COMM_ID_BYTES */);
int tad_comm_create(const void* id, int rank, int world, int device, tad_comm* out);
int tad_comm_adopt(void* nccl_comm, int rank, int world, int device, tad_comm* out);
void tad_comm_destroy(tad_comm c);   /* after the functions that use it have been destroyed or detached (tad_function_set_comm(f, NULL)) */
int tad_comm_rank(tad_comm c);
int tad_comm_world(tad_comm c);
int tad_function_set_comm(tad_function f, tad_comm c);                 /* c = NULL detaches; the pattern is rebuilt on next use */
int tad_function_vertex_owner(tad_function f, int32_t* owner_host);     /* n_handles entries; `world` = touched by no rank */
int tad_function_halo_bytes(tad_function f, int64_t* sent_per_evaluation); /* bytes this rank sends per eval_with_derivatives */

int tad_function_variable_dimension(tad_function f);
int64_t tad_function_n_vars(tad_function f);
int64_t tad_function_n_elements(tad_function f);
int64_t tad_function_n_outputs(tad_function f); /* vector functions: number of residuals */

/* Fixed sparsity pattern (built on first use).  Scalar functions: Hessian, n_vars x n_vars, compressed
 * rows == compressed columns (structurally symmetric), inner indices ascending, explicit zeros kept --
 * the arrays Eigen's setFromTriplets produces (ScalarFunctionImpl.hh:398).  Vector functions: the Jacobian,
 * n_outputs x n_vars, compressed COLUMN storage like the reference's SparseMatrix (VectorFunctionImpl.hh:186-187). */
int tad_function_pattern(tad_function f, int64_t* n_outer, int64_t* nnz);
int tad_function_pattern_copy(tad_function f, int32_t* outer_host, int32_t* inner_host);
int tad_function_pattern_device(tad_function f, const int32_t** outer_dev, const int32_t** inner_dev);

/* Recorded element -> handle table of term t (host copy, valence x n_elements row-major by slot; -1 unused). */
int tad_function_term_table(tad_function f, int term, int32_t* handles_host);

/* ---- ScalarFunction evaluation: device-resident x / g / H values ------------------------------------
 * eval                      ScalarFunctionImpl.hh:256-273   (f only; INFINITY short-circuit across terms)
 * eval_with_gradient        ScalarFunctionImpl.hh:284-299
 * eval_with_derivatives     ScalarFunctionImpl.hh:316-336   (project = 0)
 * eval_with_hessian_proj    ScalarFunctionImpl.hh:378-399   (project = 1, eps as HessianProjection.hh:48-101)
 * Outputs are overwritten.  f_host receives the value (one 8-byte D2H).  H_values_dev has nnz doubles. */
int tad_eval(tad_function f, const double* x_dev, double* f_host);
int tad_eval_with_gradient(tad_function f, const double* x_dev, double* f_host, double* g_dev);
int tad_eval_with_derivatives(tad_function f, const double* x_dev, double* f_host, double* g_dev, double* H_values_dev,
                              int project_hessian, double projection_eps);

/* Same with HOST buffers (H2D of x, D2H of g and H values inside the call).  The D2H of the CSR values is pipelined with the
 * evaluation: after every slab the prefix of the value array whose rows can no longer change is copied on a second stream
 * (fast for pinned host buffers; pageable buffers work, the copies are then staged by the driver). */
int tad_eval_host(tad_function f, const double* x_host, double* f_host);
int tad_eval_with_gradient_host(tad_function f, const double* x_host, double* f_host, double* g_host);
int tad_eval_with_derivatives_host(tad_function f, const double* x_host, double* f_host, double* g_host,
                                   double* H_values_host, int project_hessian, double projection_eps);

/* ---- VectorFunction evaluation ------------------------------------------------------------------------
 * eval                                   VectorFunctionImpl.hh:143-159     r (n_outputs)
 * eval_with_jacobian                     VectorFunctionImpl.hh:170-188     r, J values (fixed CSC pattern)
 * eval_sum_of_squares                    VectorFunctionImpl.hh:254-266     f = sum r^2
 * eval_sum_of_squares_with_derivatives   VectorFunctionImpl.hh:268-283     f = r.r, g = 2 J^T r, r, J */
int tad_veval(tad_function f, const double* x_dev, double* r_dev);
int tad_veval_with_jacobian(tad_function f, const double* x_dev, double* r_dev, double* J_values_dev);
int tad_veval_sum_of_squares(tad_function f, const double* x_dev, double* f_host);
int tad_veval_sum_of_squares_with_derivatives(tad_function f, const double* x_dev, double* f_host, double* g_dev,
                                              double* r_dev, double* J_values_dev);

/* eval_with_derivatives of a VectorFunction (VectorFunctionImpl.hh:203-236, VectorObjectiveTerm.hh:245-324): residuals, Jacobian
 * values and the Hessian of EVERY residual.  The reference returns one n_vars x n_vars sparse matrix per residual, each holding the
 * k x k entries of its element; here they come back as what they are: one dense row-major k x k block per residual in the
 * element's local variable order (local index = d * slot + component; slot -> handle: tad_function_term_table), term after term,
 * residual (e, m) of term t at H_blocks_dev[offset_t + ((M * e + m) * k) * k].  tad_function_residual_hessian_layout gives
 * offset_t, k and the number of residuals of a term (term < 0: only `total`, the length of H_blocks_dev in doubles).
 * Elements with more than 6 variables report TAD_NOT_SUPPORTED (one thread holds all M second-order scalars of an element). */
int tad_veval_with_derivatives(tad_function f, const double* x_dev, double* r_dev, double* J_values_dev, double* H_blocks_dev);
int tad_function_residual_hessian_layout(tad_function f, int term, int64_t* offset, int* k, int64_t* n_residuals, int64_t* total);

/* ---- building blocks exposed for tests / reuse ---- */
/* Batched in-place projection of n packed symmetric k x k matrices (SoA, tile order, leading dim stride):
 * project_positive_definite (Utils/HessianProjection.hh:48-101) incl. both early-outs and the eps < 0 mode.
 * method: 0 = fast path (tridiagonalisation + eigenvalues + inverse iteration for the clamped eigenpairs; the full
 * solver only for elements that do not converge), 1 = full eigendecomposition for every element.
 * counts_dev (optional, zeroed device int64[4]) receives #decomposed, #rebuilt, #handed to the full solver. */
int tad_project_batch(int k, int64_t n, int64_t stride, double* hess_dev, double eps, int method, int64_t* counts_dev, void* stream);

/* FP64 (DFMA) throughput of the device in TFLOP/s, best burst over ~`seconds` of back-to-back launches of an
 * FMA-chain kernel.  This is the measured denominator of the FP64 roofline (MEASURED_PEAKS.json has no FP64 figure). */
int tad_bench_fp64_peak(int device, double seconds, double* tflops);

/* Statistics of the last eval_with_derivatives (host int64[3]): elements eigendecomposed, rebuilt, handed to the full solver. */
int tad_function_projection_stats(tad_function f, int64_t* stats2);

/* Milliseconds of the phases of the last evaluation, measured with CUDA events on the function's stream:
 * [0] element kernels, [1] projection, [2] assembly, [3] total. */
int tad_function_last_timings(tad_function f, float* ms4);
int tad_function_set_timing(tad_function f, int enabled);
/* Number of CUDA kernels launched for this function since it was created (this library's kernels plus the element
 * kernels reported by the term launchers through tad_launch_args.launch_counter). */
int tad_function_launch_count(tad_function f, int64_t* count);

/* ---- callers of the hot path: projected-Newton utilities, all vectors device-resident ----
 * newton_direction   Utils/NewtonDirection.hh:25-48   d = -(H_proj + w_identity I)^-1 g on the function's fixed CSR pattern.
 *   The reference factorises with Eigen::SimplicialLDLT (Utils/LinearSolver.hh:12-19; BASELINE.json names cuDSS on the GPU);
 *   no sparse direct solver exists in this image, so the solve is a block-Jacobi preconditioned conjugate gradient
 *   (tad_pcg_solve) -- a labelled stand-in, timed separately from the assembly.  rel_tol <= 0 -> 1e-10, max_iters <= 0 -> 10000.
 *   TAD_SOLVER_FAILED ("Linear solve failed.") if H is not positive definite, d is not finite or the tolerance is not reached.
 * newton_decrement   Utils/NewtonDecrement.hh:20-26    -0.5 d.g
 * line_search        Utils/LineSearch.hh:26-65         x_new = x0 + s d with the first s in {s_max, s_max*shrink, ...} (and 1.0 when
 *   s_max > 1) that satisfies the Armijo condition f(x_new) <= f0 + armijo_const * s * d.g (:14-24); each trial is one value-only
 *   tad_eval (tad_veval_sum_of_squares for a vector function, GaussNewtonTest.cc:142-145); returns x0 (step 0) after max_iters failures like the reference; NaN objective -> TAD_INVALID_ARGUMENT (:53). */
int tad_pcg_solve(int64_t n, int block_dim, const int32_t* outer_dev, const int32_t* inner_dev, const double* values_dev, double w_identity,
                  const double* b_dev, double b_scale, double* x_dev, double rel_tol, int max_iters, int* iters_out, double* rel_residual_out,
                  void* stream);
int tad_newton_direction(tad_function f, const double* g_dev, const double* H_values_dev, double w_identity, double rel_tol, int max_iters,
                         double* d_dev, int* iters_out, double* rel_residual_out);
/* gauss_newton_direction  Utils/GaussNewtonDirection.hh:24-47: d = -(J^T J + w_identity I)^-1 J^T r for a
 * vector function, from its residuals r and the values of its fixed-pattern CSC Jacobian.  The reference forms J^T J with a sparse
 * product and factorises it; here the normal-equations operator is applied matrix-free (y = J p scattered, J^T y gathered) inside
 * the same PCG, Jacobi preconditioner = 1 / (column norms^2 + w). */
int tad_gauss_newton_direction(tad_function f, const double* r_dev, const double* J_values_dev, double w_identity, double rel_tol,
                               int max_iters, double* d_dev, int* iters_out, double* rel_residual_out);
int tad_newton_decrement(tad_function f, const double* d_dev, const double* g_dev, double* out_host);
int tad_line_search(tad_function f, const double* x0_dev, const double* d_dev, double f0, const double* g_dev, double s_max, double shrink,
                    int max_iters, double armijo_const, double* x_new_dev, double* f_new_host, double* step_host, int* n_evals);

#ifdef __cplusplus
}
#endif
#endif /* TINYAD_B200_H */
