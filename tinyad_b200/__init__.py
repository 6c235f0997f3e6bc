"""tinyad_b200 -- B200-native implementation of TinyAD's per-element sparse derivative path.

The product is C++/CUDA: a header facade with TinyAD's surface (tinyad_b200/include/TinyAD) over a
C-ABI runtime (include/tinyad_b200.h, libtinyad_b200.so).  This Python module is only the ctypes
harness that tests and bench.py use to drive it: it loads the two in-tree libraries, exposes the
C ABI one-to-one, and wraps the element functors compiled in csrc/energies.cu.

There is no CPU fallback: importing works without a GPU (so the symbol/ABI tests can run), but
creating a function raises unless a CUDA device is present.
"""
import ctypes
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

TAD_OK = 0
STATUS_NAMES = {0: "TAD_OK", 1: "TAD_INVALID_ARGUMENT", 2: "TAD_NONFINITE_DERIVATIVE", 3: "TAD_CUDA_ERROR",
                4: "TAD_TOO_MANY_VARIABLES", 5: "TAD_INDEX_OUT_OF_RANGE", 6: "TAD_NOT_SUPPORTED", 7: "TAD_OUT_OF_MEMORY",
                8: "TAD_SOLVER_FAILED", 9: "TAD_PATTERN_MISMATCH", 10: "TAD_COMM_ERROR"}
ASSEMBLY_ATOMIC, ASSEMBLY_GATHER = 0, 1
OPT_ASSEMBLY, OPT_CHUNK_ELEMENTS, OPT_PROJECTION, OPT_LANES, OPT_REPLICATE_GRADIENT = 1, 2, 3, 4, 5
COMM_ID_BYTES = 128

# term kinds of csrc/energies.cu
SYMDIRICHLET2D, PENALTY2D, SYMDIRICHLET3D, PENALTY3D = 1, 2, 3, 4
EDGE_DIRICHLET1D, QUADRATIC2D, REPEATED_HANDLE, TRIG_MIX2D, SQRT1D = 5, 6, 7, 8, 9
BRANCH_ON_X1D = 13                      # variables() calls depend on x -> TAD_PATTERN_MISMATCH
ARAP2D = 12                             # w |J - closest_orthogonal(J)|^2 (Operations/SVD.hh inside an element functor)
DYN_SUM_SQR2D, DYN_ONERING1D = 10, 11   # add_elements_dynamic (tests/DynamicElementsTest.cc)
SOS_SYMDIRICHLET2D, SOS_PENALTY2D, SOS_POLYCURL2D = 101, 102, 103
SOS_TEST1D_A, SOS_TEST1D_B = 104, 105     # tests/VectorFunctionTest.cc:73-93

# every symbol include/tinyad_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "tad_last_error", "tad_device_count", "tad_function_create", "tad_function_destroy", "tad_function_set_option",
    "tad_function_get_stream", "tad_function_set_caller_stream", "tad_function_launch_count", "tad_function_add_term", "tad_function_add_pattern_blocks", "tad_function_n_vars", "tad_function_n_elements",
    "tad_function_n_outputs", "tad_function_pattern", "tad_function_pattern_copy", "tad_function_pattern_device",
    "tad_function_term_table", "tad_eval", "tad_eval_with_gradient", "tad_eval_with_derivatives", "tad_eval_host",
    "tad_eval_with_gradient_host", "tad_eval_with_derivatives_host", "tad_veval", "tad_veval_with_jacobian",
    "tad_veval_sum_of_squares", "tad_veval_sum_of_squares_with_derivatives", "tad_veval_with_derivatives",
    "tad_function_residual_hessian_layout", "tad_project_batch",
    "tad_function_projection_stats", "tad_function_last_timings", "tad_function_set_timing", "tad_bench_fp64_peak",
    "tad_set_last_error", "tad_function_variable_dimension",
    "tad_comm_unique_id", "tad_comm_create", "tad_comm_adopt", "tad_comm_destroy", "tad_comm_rank", "tad_comm_world",
    "tad_function_set_comm", "tad_function_vertex_owner", "tad_function_halo_bytes",
    "tad_pcg_solve", "tad_newton_direction", "tad_gauss_newton_direction", "tad_newton_decrement", "tad_line_search",
]

_rt = None
_en = None


class TinyADError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


def runtime():
    """ctypes handle of libtinyad_b200.so (the C ABI).  Fails loudly if the CUDA library is missing."""
    global _rt
    if _rt is None:
        so = _build.RUNTIME_SO
        if not os.path.exists(so):
            raise ImportError(f"{so} is missing: run `python -m tinyad_b200.build` (needs nvcc); there is no CPU fallback")
        L = ctypes.CDLL(so, mode=ctypes.RTLD_GLOBAL)
        vp, i64, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double
        L.tad_last_error.restype = ctypes.c_char_p
        L.tad_function_create.argtypes = [ctypes.c_int, i64, ctypes.c_int, ctypes.c_int, vp]
        L.tad_function_destroy.argtypes = [vp]
        L.tad_function_destroy.restype = None
        L.tad_function_set_option.argtypes = [vp, ctypes.c_int, i64]
        L.tad_function_get_stream.argtypes = [vp, vp]
        L.tad_function_set_caller_stream.argtypes = [vp, vp, ctypes.c_int]
        L.tad_function_launch_count.argtypes = [vp, vp]
        for n in ("n_vars", "n_elements", "n_outputs"):
            fn = getattr(L, "tad_function_" + n)
            fn.restype = i64
            fn.argtypes = [vp]
        L.tad_function_pattern.argtypes = [vp, vp, vp]
        L.tad_function_add_pattern_blocks.argtypes = [vp, i64, vp, vp]
        L.tad_function_pattern_copy.argtypes = [vp, vp, vp]
        L.tad_function_pattern_device.argtypes = [vp, vp, vp]
        L.tad_function_term_table.argtypes = [vp, ctypes.c_int, vp]
        L.tad_eval.argtypes = [vp, vp, vp]
        L.tad_eval_with_gradient.argtypes = [vp, vp, vp, vp]
        L.tad_eval_with_derivatives.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, dbl]
        L.tad_eval_host.argtypes = [vp, vp, vp]
        L.tad_eval_with_gradient_host.argtypes = [vp, vp, vp, vp]
        L.tad_eval_with_derivatives_host.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, dbl]
        L.tad_veval.argtypes = [vp, vp, vp]
        L.tad_veval_with_jacobian.argtypes = [vp, vp, vp, vp]
        L.tad_veval_sum_of_squares.argtypes = [vp, vp, vp]
        L.tad_veval_sum_of_squares_with_derivatives.argtypes = [vp, vp, vp, vp, vp, vp]
        L.tad_veval_with_derivatives.argtypes = [vp, vp, vp, vp, vp]
        L.tad_function_residual_hessian_layout.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
        L.tad_project_batch.argtypes = [ctypes.c_int, i64, i64, vp, dbl, ctypes.c_int, vp, vp]
        L.tad_function_projection_stats.argtypes = [vp, vp]
        L.tad_function_last_timings.argtypes = [vp, vp]
        L.tad_function_set_timing.argtypes = [vp, ctypes.c_int]
        L.tad_device_count.argtypes = [vp]
        L.tad_bench_fp64_peak.argtypes = [ctypes.c_int, dbl, vp]
        L.tad_function_variable_dimension.argtypes = [vp]
        L.tad_comm_unique_id.argtypes = [vp]
        L.tad_comm_create.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        L.tad_comm_adopt.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        L.tad_comm_destroy.argtypes = [vp]
        L.tad_comm_destroy.restype = None
        L.tad_comm_rank.argtypes = [vp]
        L.tad_comm_world.argtypes = [vp]
        L.tad_function_set_comm.argtypes = [vp, vp]
        L.tad_function_vertex_owner.argtypes = [vp, vp]
        L.tad_function_halo_bytes.argtypes = [vp, vp]
        L.tad_pcg_solve.argtypes = [i64, ctypes.c_int, vp, vp, vp, dbl, vp, dbl, vp, dbl, ctypes.c_int, vp, vp, vp]
        L.tad_newton_direction.argtypes = [vp, vp, vp, dbl, dbl, ctypes.c_int, vp, vp, vp]
        L.tad_gauss_newton_direction.argtypes = [vp, vp, vp, dbl, dbl, ctypes.c_int, vp, vp, vp]
        L.tad_newton_decrement.argtypes = [vp, vp, vp, vp]
        L.tad_line_search.argtypes = [vp, vp, vp, dbl, vp, dbl, dbl, ctypes.c_int, dbl, vp, vp, vp, vp]
        _rt = L
    return _rt


def energies():
    """ctypes handle of libtinyad_b200_energies.so (element functors of the tests / benchmark)."""
    global _en
    if _en is None:
        runtime()
        so = os.environ.get("TINYAD_ENERGIES_SO", _build.ENERGIES_SO)   # override: experiment builds of the functor library
        if not os.path.exists(so):
            raise ImportError(f"{so} is missing: run `python -m tinyad_b200.build`")
        L = ctypes.CDLL(so)
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.tadx_last_error.restype = ctypes.c_char_p
        L.tadx_create.argtypes = [ctypes.c_int, i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        L.tadx_destroy.argtypes = [vp]
        L.tadx_destroy.restype = None
        L.tadx_handle.argtypes = [vp]
        L.tadx_handle.restype = vp
        L.tadx_add_term.argtypes = [vp, ctypes.c_int, i64, vp, ctypes.c_int, vp, ctypes.c_int]
        L.tadx_scalar_case.argtypes = [ctypes.c_int, vp, vp, ctypes.c_int]
        L.tadx_selftest.argtypes = [ctypes.c_int]
        _en = L
    return _en


def _check(status):
    if status != TAD_OK:
        raise TinyADError(status, runtime().tad_last_error().decode())


def _ptr(a):
    """Device or host pointer of a torch tensor / numpy array / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


class Comm:
    """One rank's handle on the group of GPUs that share a partitioned function (tad_comm of the C ABI; NCCL underneath)."""

    def __init__(self, id_bytes, rank, world, device):
        self.rank, self.world, self.device = rank, world, device
        self.h = ctypes.c_void_p()
        buf = (ctypes.c_ubyte * COMM_ID_BYTES).from_buffer_copy(bytes(id_bytes))
        _check(runtime().tad_comm_create(buf, rank, world, device, ctypes.byref(self.h)))

    @staticmethod
    def unique_id():
        buf = (ctypes.c_ubyte * COMM_ID_BYTES)()
        _check(runtime().tad_comm_unique_id(buf))
        return bytes(buf)

    @staticmethod
    def from_torch_distributed(device, group=None):
        """The 128-byte NCCL id is created on rank 0 and broadcast with torch.distributed (plumbing only; the exchange itself runs
        inside the runtime on its own communicator)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return Comm(box[0], rank, world, device)

    def close(self):
        if self.h:
            runtime().tad_comm_destroy(self.h)
            self.h = ctypes.c_void_p()


class Function:
    """A ScalarFunction<d> / VectorFunction<d> built from the functors in csrc/energies.cu.

    Thin mirror of the C++ facade (tinyad_b200/include/TinyAD/ScalarFunction.hh); every method is one call
    through the C ABI.  `*_host` variants take / return numpy arrays, the others device pointers
    (torch CUDA tensors)."""

    def __init__(self, d, n_vertices, is_vector=False, device=0, assembly=ASSEMBLY_ATOMIC):
        self.d, self.n_vertices, self.is_vector = d, n_vertices, is_vector
        self.n_vars = d * n_vertices
        self._x = ctypes.c_void_p()
        E = energies()
        if E.tadx_create(d, n_vertices, int(is_vector), device, assembly, ctypes.byref(self._x)) != 0:
            raise TinyADError(-1, E.tadx_last_error().decode())
        self.h = E.tadx_handle(self._x)
        self._pattern = None

    def close(self):
        if self._x:
            energies().tadx_destroy(self._x)
            self._x = ctypes.c_void_p()
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_term(self, kind, conn, data):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        conn = conn.reshape(len(conn), -1)
        data = np.ascontiguousarray(data, dtype=np.float64).reshape(len(conn), -1)
        E = energies()
        if E.tadx_add_term(self._x, kind, conn.shape[0], conn.ctypes.data, conn.shape[1], data.ctypes.data, data.shape[1]) != 0:
            raise TinyADError(-1, E.tadx_last_error().decode())
        self._pattern = None
        return self

    def add_pattern_blocks(self, vi, vj):
        """Structural-only d x d blocks (vertex pairs) for halo rows received from other ranks."""
        vi = np.ascontiguousarray(vi, dtype=np.int64)
        vj = np.ascontiguousarray(vj, dtype=np.int64)
        _check(runtime().tad_function_add_pattern_blocks(self.h, len(vi), vi.ctypes.data, vj.ctypes.data))
        self._pattern = None

    def set_comm(self, comm):
        """Partitioned evaluation: this rank's terms hold ITS elements (global handles); see include/tinyad_b200.h."""
        _check(runtime().tad_function_set_comm(self.h, comm.h if comm is not None else None))
        self._comm = comm
        self._pattern = None

    def vertex_owner(self):
        """int32 per vertex: the rank that owns its rows after the exchange (lowest rank touching it)."""
        o = np.empty(self.n_vertices, dtype=np.int32)
        _check(runtime().tad_function_vertex_owner(self.h, o.ctypes.data))
        return o

    def halo_bytes(self):
        n = ctypes.c_int64()
        _check(runtime().tad_function_halo_bytes(self.h, ctypes.byref(n)))
        return n.value

    def set_option(self, opt, value):
        _check(runtime().tad_function_set_option(self.h, opt, value))

    def launch_count(self):
        """CUDA kernels launched for this function so far."""
        n = ctypes.c_int64()
        _check(runtime().tad_function_launch_count(self.h, ctypes.byref(n)))
        return n.value

    def set_caller_stream(self, stream_ptr, enabled=True):
        """Evaluations first wait for the work queued on this CUDA stream (0 / None = legacy default stream)."""
        _check(runtime().tad_function_set_caller_stream(self.h, stream_ptr, int(enabled)))

    def set_timing(self, on=True):
        _check(runtime().tad_function_set_timing(self.h, int(on)))

    def last_timings(self):
        ms = (ctypes.c_float * 4)()
        _check(runtime().tad_function_last_timings(self.h, ms))
        return {"element_ms": ms[0], "projection_ms": ms[1], "assembly_ms": ms[2], "total_ms": ms[3]}

    def projection_stats(self):
        s = (ctypes.c_int64 * 3)()
        _check(runtime().tad_function_projection_stats(self.h, s))
        return {"decomposed": s[0], "rebuilt": s[1], "full_solver": s[2]}

    @property
    def n_elements(self):
        return runtime().tad_function_n_elements(self.h)

    @property
    def n_outputs(self):
        return runtime().tad_function_n_outputs(self.h)

    def pattern(self):
        """(outer, inner) int32 numpy arrays of the fixed pattern (Hessian CSR==CSC, or Jacobian CSC)."""
        if self._pattern is None:
            n_outer, nnz = ctypes.c_int64(), ctypes.c_int64()
            _check(runtime().tad_function_pattern(self.h, ctypes.byref(n_outer), ctypes.byref(nnz)))
            outer = np.empty(n_outer.value + 1, dtype=np.int32)
            inner = np.empty(nnz.value, dtype=np.int32)
            _check(runtime().tad_function_pattern_copy(self.h, outer.ctypes.data, inner.ctypes.data))
            self._pattern = (outer, inner)
        return self._pattern

    @property
    def nnz(self):
        return len(self.pattern()[1])

    def term_table(self, term, valence, n_elements):
        t = np.empty((valence, n_elements), dtype=np.int32)
        _check(runtime().tad_function_term_table(self.h, term, t.ctypes.data))
        return t

    # ---- host-buffer calls (H2D / D2H inside the C call) ----
    def eval_host(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = ctypes.c_double()
        _check(runtime().tad_eval_host(self.h, x.ctypes.data, ctypes.byref(f)))
        return f.value

    def eval_with_gradient_host(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = ctypes.c_double()
        g = np.empty(self.n_vars)
        _check(runtime().tad_eval_with_gradient_host(self.h, x.ctypes.data, ctypes.byref(f), g.ctypes.data))
        return f.value, g

    def eval_with_derivatives_host(self, x, project=False, eps=1e-9, out_g=None, out_H=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = ctypes.c_double()
        g = np.empty(self.n_vars) if out_g is None else out_g
        H = np.empty(self.nnz) if out_H is None else out_H
        _check(runtime().tad_eval_with_derivatives_host(self.h, _ptr(x), ctypes.byref(f), _ptr(g), _ptr(H), int(project), eps))
        return f.value, g, H

    def eval_with_hessian_proj_host(self, x, eps=1e-9, **kw):
        return self.eval_with_derivatives_host(x, True, eps, **kw)

    # ---- device-pointer calls ----
    def eval(self, x_dev):
        f = ctypes.c_double()
        _check(runtime().tad_eval(self.h, _ptr(x_dev), ctypes.byref(f)))
        return f.value

    def eval_with_gradient(self, x_dev, g_dev):
        f = ctypes.c_double()
        _check(runtime().tad_eval_with_gradient(self.h, _ptr(x_dev), ctypes.byref(f), _ptr(g_dev)))
        return f.value

    def eval_with_derivatives(self, x_dev, g_dev, H_dev, project=False, eps=1e-9):
        f = ctypes.c_double()
        _check(runtime().tad_eval_with_derivatives(self.h, _ptr(x_dev), ctypes.byref(f), _ptr(g_dev), _ptr(H_dev), int(project), eps))
        return f.value

    def eval_with_hessian_proj(self, x_dev, g_dev, H_dev, eps=1e-9):
        return self.eval_with_derivatives(x_dev, g_dev, H_dev, True, eps)

    # ---- projected-Newton utilities (Utils/NewtonDirection.hh, NewtonDecrement.hh, LineSearch.hh), device pointers ----
    def newton_direction(self, g_dev, H_dev, d_dev, w_identity=0.0, rel_tol=1e-10, max_iters=10000):
        """d = -(H + w_identity I)^-1 g by block-Jacobi PCG on the fixed pattern.  Returns (iterations, relative residual)."""
        it, rel = ctypes.c_int(), ctypes.c_double()
        _check(runtime().tad_newton_direction(self.h, _ptr(g_dev), _ptr(H_dev), w_identity, rel_tol, max_iters, _ptr(d_dev),
                                              ctypes.byref(it), ctypes.byref(rel)))
        return it.value, rel.value

    def gauss_newton_direction(self, r_dev, J_dev, d_dev, w_identity=0.0, rel_tol=1e-10, max_iters=10000):
        """d = -(J^T J + w_identity I)^-1 J^T r (Utils/GaussNewtonDirection.hh:24-47), matrix-free PCG.  Returns (iterations, rel. residual)."""
        it, rel = ctypes.c_int(), ctypes.c_double()
        _check(runtime().tad_gauss_newton_direction(self.h, _ptr(r_dev), _ptr(J_dev), w_identity, rel_tol, max_iters, _ptr(d_dev),
                                                    ctypes.byref(it), ctypes.byref(rel)))
        return it.value, rel.value

    def newton_decrement(self, d_dev, g_dev):
        out = ctypes.c_double()
        _check(runtime().tad_newton_decrement(self.h, _ptr(d_dev), _ptr(g_dev), ctypes.byref(out)))
        return out.value

    def line_search(self, x0_dev, d_dev, f0, g_dev, x_new_dev, s_max=1.0, shrink=0.8, max_iters=64, armijo_const=1e-4):
        """Backtracking Armijo line search; writes x_new_dev, returns (f_new, step, n_evals); step 0 = no improvement found."""
        f_new, step, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        _check(runtime().tad_line_search(self.h, _ptr(x0_dev), _ptr(d_dev), f0, _ptr(g_dev), s_max, shrink, max_iters, armijo_const,
                                         _ptr(x_new_dev), ctypes.byref(f_new), ctypes.byref(step), ctypes.byref(n)))
        return f_new.value, step.value, n.value

    # ---- vector functions (device pointers) ----
    def veval(self, x_dev, r_dev):
        _check(runtime().tad_veval(self.h, _ptr(x_dev), _ptr(r_dev)))

    def veval_with_jacobian(self, x_dev, r_dev, J_dev):
        _check(runtime().tad_veval_with_jacobian(self.h, _ptr(x_dev), _ptr(r_dev), _ptr(J_dev)))

    def veval_sum_of_squares(self, x_dev):
        f = ctypes.c_double()
        _check(runtime().tad_veval_sum_of_squares(self.h, _ptr(x_dev), ctypes.byref(f)))
        return f.value

    def veval_sum_of_squares_with_derivatives(self, x_dev, g_dev, r_dev, J_dev):
        f = ctypes.c_double()
        _check(runtime().tad_veval_sum_of_squares_with_derivatives(self.h, _ptr(x_dev), ctypes.byref(f), _ptr(g_dev), _ptr(r_dev), _ptr(J_dev)))
        return f.value


def _residual_hessian_layout(fn, term):
    off, k, n_res, total = ctypes.c_int64(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
    _check(runtime().tad_function_residual_hessian_layout(fn.h, term, ctypes.byref(off), ctypes.byref(k), ctypes.byref(n_res), ctypes.byref(total)))
    return off.value, k.value, n_res.value, total.value


def _veval_with_derivatives(fn, x_dev, r_dev, J_dev, Hb_dev):
    _check(runtime().tad_veval_with_derivatives(fn.h, _ptr(x_dev), _ptr(r_dev), _ptr(J_dev), _ptr(Hb_dev)))


Function.residual_hessian_layout = _residual_hessian_layout      # (offset, k, n_residuals, total doubles) of a term (term < 0: total only)
Function.veval_with_derivatives = _veval_with_derivatives        # r, J values, one dense k x k Hessian block per residual


def project_batch(k, hess_dev, n, stride, eps=1e-9, method=0, counts_dev=None, stream=None):
    """tad_project_batch on a device SoA buffer (torch tensor); counts_dev: zeroed int64[4]."""
    _check(runtime().tad_project_batch(k, n, stride, _ptr(hess_dev), eps, method, _ptr(counts_dev), stream))


def fp64_peak_tflops(device=0, seconds=0.5):
    """Measured DFMA throughput (TFLOP/s) -- the FP64 roofline denominator."""
    t = ctypes.c_double()
    _check(runtime().tad_bench_fp64_peak(device, seconds, ctypes.byref(t)))
    return t.value


# order of enum ScalarCase in csrc/energies.cuh
SCALAR_CASES = [
    "neg", "sqrt", "sqr", "fabs", "abs", "exp", "log", "log2", "log10", "sin", "cos", "tan", "asin", "acos", "atan",
    "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "pow_int", "pow_real",
    "add", "sub", "mul", "div", "add_s", "s_add", "sub_s", "s_sub", "mul_s", "s_mul", "div_s", "s_div",
    "iadd", "isub", "imul", "idiv", "iadd_s", "isub_s", "imul_s", "idiv_s", "min", "max", "clamp", "quadratic", "atan2_1",
    "sqr_pow_mul", "atan2_const", "atan2_2", "hypot", "div2d", "div2d_2", "plus_minus_mult_div_2d", "sphere",
    "c_mul", "c_mul_d", "c_d_mul", "c_div", "c_div_d", "c_add", "c_sub", "c_sqr", "c_conj", "c_abs", "c_arg", "symm_dirich6",
    "svd2", "closest_orthogonal2",
    "fmin", "fmax", "clamp_d", "cmp", "isnan_isinf",
    "hess_block_issue13", "hess_block_symdir",
]


def scalar_case(name, params, k, on_device=False):
    """Known-answer case of the reference's Scalar tests on the product's Scalar (host build or a 1-thread kernel).
    Returns a list of (val, grad[k], Hess[k, k])."""
    p = np.zeros(16)
    p[:len(params)] = params
    out = np.zeros(256)
    n = energies().tadx_scalar_case(SCALAR_CASES.index(name), p.ctypes.data, out.ctypes.data, int(on_device))
    if n < 0:
        raise RuntimeError(f"scalar case {name}: code {n}")
    res, o = [], 0
    for _ in range(n):
        res.append((out[o], out[o + 1:o + 1 + k].copy(), out[o + 1 + k:o + 1 + k + k * k].reshape(k, k).copy()))
        o += 1 + k + k * k
    return res


def selftest(device=0):
    """C++ facade semantics (default-constructed / moved functions, errors leave the object usable, x_from_data ...)."""
    code = energies().tadx_selftest(device)
    if code != 0:
        raise AssertionError(f"tadx_selftest failed at step {code}: {energies().tadx_last_error().decode()}")
