"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/tinyad_oracle.hh).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
The product package tinyad_b200 never does.
"""
import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# term kinds (keep in sync with oracle_capi.cc)
SYMDIRICHLET2D, PENALTY2D, SYMDIRICHLET3D, PENALTY3D = 1, 2, 3, 4
EDGE_DIRICHLET1D, QUADRATIC2D, REPEATED_HANDLE, TRIG_MIX2D = 5, 6, 7, 8
ARAP2D = 12
DYN_SUM_SQR2D, DYN_ONERING1D = 10, 11   # add_elements_dynamic; ONERING: data must have as many columns as conn
SOS_SYMDIRICHLET2D, SOS_PENALTY2D, SOS_POLYCURL2D = 101, 102, 103


class _Term(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("n_elements", ctypes.c_int64),
                ("conn", ctypes.c_void_p), ("data", ctypes.c_void_p), ("n_data", ctypes.c_int)]


def build(force=False):
    """Compile liboracle.so (g++ -O3 -fopenmp); idempotent."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cc", "tinyad_oracle.hh", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.oracle_scalar_eval.restype = ctypes.c_void_p
        L.oracle_scalar_eval.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
        L.oracle_vector_eval.restype = ctypes.c_void_p
        L.oracle_vector_eval.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_int]
        L.oracle_last_error.restype = ctypes.c_char_p
        L.oracle_result_f.restype = ctypes.c_double
        L.oracle_result_f.argtypes = [ctypes.c_void_p]
        for n in ("nnz", "rows", "cols", "g_size", "r_size"):
            f = getattr(L, "oracle_result_" + n)
            f.restype = ctypes.c_int64
            f.argtypes = [ctypes.c_void_p]
        L.oracle_result_copy.argtypes = [ctypes.c_void_p] * 6
        L.oracle_result_phases.argtypes = [ctypes.c_void_p] * 3
        L.oracle_result_free.argtypes = [ctypes.c_void_p]
        L.oracle_project.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
        L.oracle_scalar_case.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


@dataclass
class Term:
    kind: int
    conn: np.ndarray   # (n_elements, valence) int32
    data: np.ndarray   # (n_elements, n_data) float64


@dataclass
class Result:
    f: float = 0.0
    g: np.ndarray = None
    r: np.ndarray = None
    outer: np.ndarray = None
    inner: np.ndarray = None
    values: np.ndarray = None
    shape: tuple = (0, 0)
    phases: dict = field(default_factory=dict)


def _terms_array(terms):
    keep = []
    arr = (_Term * len(terms))()
    for i, t in enumerate(terms):
        conn = np.ascontiguousarray(t.conn, dtype=np.int32)
        data = np.ascontiguousarray(t.data, dtype=np.float64)
        conn = conn if conn.ndim == 2 else conn.reshape(len(conn), -1)     # (0, valence) stays what it is: an empty element range
        data = data if data.ndim == 2 else data.reshape(len(conn), -1)
        keep += [conn, data]
        arr[i] = _Term(t.kind, conn.shape[0], conn.ctypes.data, data.ctypes.data, data.shape[1])
    return arr, keep


def _collect(L, h, want_matrix):
    if not h:
        raise RuntimeError(L.oracle_last_error().decode())
    try:
        res = Result(f=L.oracle_result_f(h))
        ng, nr = L.oracle_result_g_size(h), L.oracle_result_r_size(h)
        res.g = np.empty(ng)
        res.r = np.empty(nr)
        rows, cols, nnz = L.oracle_result_rows(h), L.oracle_result_cols(h), L.oracle_result_nnz(h)
        res.shape = (rows, cols)
        if want_matrix:
            res.outer = np.empty(cols + 1, dtype=np.int32)
            res.inner = np.empty(nnz, dtype=np.int32)
            res.values = np.empty(nnz)
        L.oracle_result_copy(h, res.g.ctypes.data, res.r.ctypes.data,
                             res.outer.ctypes.data if want_matrix else None,
                             res.inner.ctypes.data if want_matrix else None,
                             res.values.ctypes.data if want_matrix else None)
        t = (ctypes.c_double * 3)()
        n = (ctypes.c_int64 * 2)()
        L.oracle_result_phases(h, t, n)
        res.phases = {"eval_s": t[0], "accumulate_s": t[1], "compress_s": t[2],
                      "n_decomposed": n[0], "n_rebuilt": n[1]}
        return res
    finally:
        L.oracle_result_free(h)


EVAL, GRADIENT, DERIVATIVES, HESSIAN_PROJ = 0, 1, 2, 3
NORM_SUM = 16  # same, every contribution counted with its element's largest |entry| (element-level rounding scale)
ABS_SUM = 8   # OR-ed into DERIVATIVES / HESSIAN_PROJ: g and H values come back as sum |contributions| per entry (parity scale)


def scalar_eval(d, n_vertices, terms, mode, x, eps=1e-9, n_threads=-1):
    """Reference semantics of ScalarFunction::eval* (Detail/ScalarFunctionImpl.hh:256-416).
    H is returned as compressed-column arrays (== CSR of the structurally symmetric Hessian)."""
    L = lib()
    arr, keep = _terms_array(terms)
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.size == d * n_vertices
    h = L.oracle_scalar_eval(d, n_vertices, len(terms), ctypes.addressof(arr), mode, x.ctypes.data, eps, n_threads)
    return _collect(L, h, (mode & 7) >= 2)


V_EVAL, V_JACOBIAN, V_SOS, V_SOS_DERIVATIVES = 0, 1, 2, 3


def vector_eval(d, n_vertices, terms, mode, x, n_threads=-1):
    """Reference semantics of VectorFunction::eval* (Detail/VectorFunctionImpl.hh:143-301). J is CSC (m x n)."""
    L = lib()
    arr, keep = _terms_array(terms)
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = L.oracle_vector_eval(d, n_vertices, len(terms), ctypes.addressof(arr), mode, x.ctypes.data, n_threads)
    return _collect(L, h, mode in (1, 3))


def project(H, eps=1e-9):
    """project_positive_definite (Utils/HessianProjection.hh:48-101) on one dense symmetric matrix."""
    A = np.array(H, dtype=np.float64, order="C")
    code = lib().oracle_project(A.shape[0], A.ctypes.data, eps)
    if code < 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return A, code


def scalar_case(name, params, k, n_out_max=3):
    """Run one known-answer scalar case; returns list of (val, grad[k], Hess[k,k])."""
    p = np.zeros(16)
    p[:len(params)] = params
    out = np.zeros(n_out_max * (1 + k + k * k) + 64)
    n = lib().oracle_scalar_case(name.encode(), p.ctypes.data, out.ctypes.data)
    if n < 0:
        raise RuntimeError(f"scalar case {name}: code {n} {lib().oracle_last_error().decode()}")
    res, o = [], 0
    for _ in range(n):
        val = out[o]
        grad = out[o + 1:o + 1 + k].copy()
        hess = out[o + 1 + k:o + 1 + k + k * k].reshape(k, k).copy()
        res.append((val, grad, hess))
        o += 1 + k + k * k
    return res


def default_threads():
    return lib().oracle_default_threads()


# ---------------------------------------------------------------------------------------------------------------------
# oracle/_ref: the UNMODIFIED reference headers compiled over oracle/eigen_shim (ref_driver.cc).  Used to check the
# restatement above (tests/test_oracle_vs_reference.py) and as the `--impl reference` arm of bench.py.
# ---------------------------------------------------------------------------------------------------------------------
_REF = None
REFERENCE_ROOT = "/root/reference"


def ref_path():
    return os.path.join(_HERE, "_ref", "libtinyad_ref.so")


def build_ref(force=False):
    """Compile oracle/_ref/libtinyad_ref.so from the reference headers where they lie; needs /root/reference
    (the build container).  Elsewhere the prebuilt file is used as is.  Returns the path or None."""
    so = ref_path()
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "include", "TinyAD")):
        srcs = [os.path.join(_HERE, f) for f in ("ref_driver.cc", "eigen_shim/Eigen/src/Shim.h", "Makefile")]
        if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            try:
                subprocess.run(["make", "-C", _HERE, "_ref/libtinyad_ref.so"], check=True, capture_output=True)
            except (subprocess.CalledProcessError, OSError) as e:      # the checker's checker is optional: tests that need it skip
                import warnings
                warnings.warn(f"oracle/_ref could not be built: {getattr(e, 'stderr', b'')[-400:]!r}")
                return None
    return so if os.path.exists(so) else None


def ref_available():
    return os.path.exists(ref_path())


# Reference tests excluded from the run (none: all 404 pass over the shim).  History: while the shim solved SimplicialLDLT's
# systems with a dense LU, NewtonTest.2DDeformationDouble missed its absolute 1e-15 bound on the converged gradient by rounding
# (1.39e-15); with an actual L D L^T (what that solver computes) it passes.
REF_TESTS_SKIPPED = ()


def ref_tests_path():
    return os.path.join(_HERE, "_ref", "reference_tests")


def build_ref_tests(force=False):
    """Compile oracle/_ref/reference_tests: the reference's own tests/*.cc, where they lie, over oracle/eigen_shim and
    oracle/gtest_shim.  Needs /root/reference; elsewhere the prebuilt binary is used.  Returns the path or None."""
    exe = ref_tests_path()
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "tests")):
        srcs = [os.path.join(_HERE, f) for f in ("ref_tests_main.cc", "eigen_shim/Eigen/src/Shim.h", "gtest_shim/gtest/gtest.h", "Makefile")]
        if force or not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
            try:
                subprocess.run(["make", "-j8", "-C", _HERE, "_ref/reference_tests"], check=True, capture_output=True)
            except (subprocess.CalledProcessError, OSError) as e:
                import warnings
                warnings.warn(f"oracle/_ref/reference_tests could not be built: {getattr(e, 'stderr', b'')[-400:]!r}")
                return None
            import shutil
            shutil.rmtree(os.path.join(_HERE, "_ref", "obj"), ignore_errors=True)   # 14 MB of objects: keep the snapshot small
    return exe if os.path.exists(exe) else None


def run_ref_tests(name_filter=None, skip=REF_TESTS_SKIPPED, timeout=600):
    """Runs the reference's test suite; returns (return code, output)."""
    exe = build_ref_tests()
    if exe is None:
        raise RuntimeError("oracle/_ref/reference_tests is not built (make -C oracle _ref/reference_tests, needs /root/reference)")
    cmd = [exe] + (["--skip=" + ",".join(skip)] if skip else []) + ([name_filter] if name_filter else [])
    env = dict(os.environ)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    env["OMP_NUM_THREADS"] = str(max(2, cores))      # tests/OpenMPTest.cc:24 asserts omp_get_max_threads() >= 2, whatever the caller exported
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    return p.returncode, p.stdout + p.stderr


def ref_lib():
    global _REF
    if _REF is None:
        so = build_ref()
        if so is None:
            raise RuntimeError("oracle/_ref/libtinyad_ref.so is not built (make -C oracle _ref, needs /root/reference)")
        L = ctypes.CDLL(so)
        L.ref_scalar_eval.restype = ctypes.c_void_p
        L.ref_scalar_eval.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
        L.ref_vector_eval.restype = ctypes.c_void_p
        L.ref_vector_eval.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_int]
        L.ref_last_error.restype = ctypes.c_char_p
        L.ref_description.restype = ctypes.c_char_p
        for n in ("f", "seconds"):
            f = getattr(L, "ref_result_" + n)
            f.restype = ctypes.c_double
            f.argtypes = [ctypes.c_void_p]
        for n in ("nnz", "rows", "cols", "g_size", "r_size"):
            f = getattr(L, "ref_result_" + n)
            f.restype = ctypes.c_int64
            f.argtypes = [ctypes.c_void_p]
        L.ref_result_copy.argtypes = [ctypes.c_void_p] * 6
        L.ref_result_hess_count.restype = ctypes.c_int64
        L.ref_result_hess_count.argtypes = [ctypes.c_void_p]
        L.ref_result_hess_copy.argtypes = [ctypes.c_void_p] * 5
        L.ref_result_free.argtypes = [ctypes.c_void_p]
        L.ref_project.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
        _REF = L
    return _REF


def _ref_collect(L, h, want_matrix):
    if not h:
        raise RuntimeError(L.ref_last_error().decode())
    try:
        res = Result(f=L.ref_result_f(h))
        res.g = np.empty(L.ref_result_g_size(h))
        res.r = np.empty(L.ref_result_r_size(h))
        rows, cols, nnz = L.ref_result_rows(h), L.ref_result_cols(h), L.ref_result_nnz(h)
        res.shape = (rows, cols)
        if want_matrix:
            res.outer = np.empty(cols + 1, dtype=np.int32)
            res.inner = np.empty(nnz, dtype=np.int32)
            res.values = np.empty(nnz)
        L.ref_result_copy(h, res.g.ctypes.data, res.r.ctypes.data,
                          res.outer.ctypes.data if want_matrix else None,
                          res.inner.ctypes.data if want_matrix else None,
                          res.values.ctypes.data if want_matrix else None)
        res.phases = {"total_s": L.ref_result_seconds(h)}
        nh = L.ref_result_hess_count(h)
        if nh:
            hr, hi, hj = (np.empty(nh, dtype=np.int32) for _ in range(3))
            hv = np.empty(nh)
            L.ref_result_hess_copy(h, hr.ctypes.data, hi.ctypes.data, hj.ctypes.data, hv.ctypes.data)
            res.phases["residual_hessians"] = (hr, hi, hj, hv)      # (residual, row, col, value) of every stored entry
        return res
    finally:
        L.ref_result_free(h)


def ref_scalar_eval(d, n_vertices, terms, mode, x, eps=1e-9, n_threads=-1):
    """TinyAD::ScalarFunction::eval* of the reference itself (same arguments as scalar_eval)."""
    L = ref_lib()
    arr, keep = _terms_array(terms)
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.size == d * n_vertices
    h = L.ref_scalar_eval(d, n_vertices, len(terms), ctypes.addressof(arr), mode, x.ctypes.data, eps, n_threads)
    return _ref_collect(L, h, mode >= 2)


def ref_vector_eval(d, n_vertices, terms, mode, x, n_threads=-1):
    """TinyAD::VectorFunction::eval* of the reference itself (same arguments as vector_eval)."""
    L = ref_lib()
    arr, keep = _terms_array(terms)
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = L.ref_vector_eval(d, n_vertices, len(terms), ctypes.addressof(arr), mode, x.ctypes.data, n_threads)
    return _ref_collect(L, h, mode in (1, 3, 4))


V_DERIVATIVES = 4   # ref_vector_eval only: VectorFunction::eval_with_derivatives (r, J, Hessian of every residual)


def ref_scalar_case(name, params, k, n_out_max=3):
    """The scalar case `name` of the shared vocabulary on the reference's TinyAD::Scalar; same return value as scalar_case."""
    L = ref_lib()
    L.ref_scalar_case.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p]
    p = np.zeros(16)
    p[:len(params)] = params
    out = np.zeros(n_out_max * (1 + k + k * k) + 64)
    n = L.ref_scalar_case(name.encode(), p.ctypes.data, out.ctypes.data)
    if n < 0:
        raise RuntimeError(f"reference scalar case {name}: code {n} {L.ref_last_error().decode()}")
    res, o = [], 0
    for _ in range(n):
        res.append((out[o], out[o + 1:o + 1 + k].copy(), out[o + 1 + k:o + 1 + k + k * k].reshape(k, k).copy()))
        o += 1 + k + k * k
    return res


# ---------------------------------------------------------------------------------------------------------------------
# oracle/_ref/libtinyad_plugin.so: the reference's ScalarFunction with include/reference_binding/B200ObjectiveTerm.hh plugged
# into its objective_terms (oracle/ref_plugin_driver.cc) -- the drop-in exercised from the reference's side.
# ---------------------------------------------------------------------------------------------------------------------
_PLUGIN = None


class _PluginTerm(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("n_elements", ctypes.c_int64), ("conn", ctypes.c_void_p), ("data", ctypes.c_void_p),
                ("n_data", ctypes.c_int), ("valence", ctypes.c_int)]


def plugin_path():
    return os.path.join(_HERE, "_ref", "libtinyad_plugin.so")


def build_plugin(force=False):
    """Needs /root/reference and the built product libraries (python -m tinyad_b200.build); elsewhere the prebuilt file is used."""
    so = plugin_path()
    product = os.path.join(os.path.dirname(_HERE), "tinyad_b200", "libtinyad_b200.so")
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "include", "TinyAD")) and os.path.exists(product):
        srcs = [os.path.join(_HERE, f) for f in ("ref_plugin_driver.cc", "eigen_shim/Eigen/src/Shim.h", "Makefile",
                                                 "../include/reference_binding/B200ObjectiveTerm.hh", "../include/tinyad_b200.h")]
        if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            try:
                subprocess.run(["make", "-C", _HERE, "_ref/libtinyad_plugin.so"], check=True, capture_output=True)
            except (subprocess.CalledProcessError, OSError) as e:
                import warnings
                warnings.warn(f"oracle/_ref/libtinyad_plugin.so could not be built: {getattr(e, 'stderr', b'')[-400:]!r}")
                return None
    return so if os.path.exists(so) else None


def plugin_lib():
    global _PLUGIN
    if _PLUGIN is None:
        so = build_plugin()
        if so is None:
            raise RuntimeError("oracle/_ref/libtinyad_plugin.so is not built (make -C oracle _ref/libtinyad_plugin.so)")
        L = ctypes.CDLL(so)
        L.plugin_scalar_eval.restype = ctypes.c_void_p
        L.plugin_scalar_eval.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_double]
        L.plugin_last_error.restype = ctypes.c_char_p
        L.plugin_result_f.restype = ctypes.c_double
        L.plugin_result_f.argtypes = [ctypes.c_void_p]
        for n in ("nnz", "cols", "g_size"):
            f = getattr(L, "plugin_result_" + n)
            f.restype = ctypes.c_int64
            f.argtypes = [ctypes.c_void_p]
        L.plugin_result_copy.argtypes = [ctypes.c_void_p] * 5
        L.plugin_result_free.argtypes = [ctypes.c_void_p]
        _PLUGIN = L
    return _PLUGIN


def plugin_scalar_eval(d, n_vertices, terms, mode, x, eps=1e-9, assembly=0):
    """TinyAD::ScalarFunction::eval* of the REFERENCE with a B200ScalarObjectiveTerm in its objective_terms: the element functors
    run on cuda:0 through the product's C ABI, everything above the term interface is the reference's own code."""
    L = plugin_lib()
    keep = []
    arr = (_PluginTerm * len(terms))()
    for i, t in enumerate(terms):
        conn = np.ascontiguousarray(t.conn, dtype=np.int32).reshape(len(t.conn), -1)
        data = np.ascontiguousarray(t.data, dtype=np.float64).reshape(len(t.conn), -1)
        keep += [conn, data]
        arr[i] = _PluginTerm(t.kind, conn.shape[0], conn.ctypes.data, data.ctypes.data, data.shape[1], conn.shape[1])
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.size == d * n_vertices
    h = L.plugin_scalar_eval(d, n_vertices, len(terms), ctypes.addressof(arr), assembly, mode, x.ctypes.data, eps)
    if not h:
        raise RuntimeError(L.plugin_last_error().decode())
    try:
        res = Result(f=L.plugin_result_f(h))
        res.g = np.empty(L.plugin_result_g_size(h))
        res.r = np.empty(0)
        cols, nnz = L.plugin_result_cols(h), L.plugin_result_nnz(h)
        res.shape = (cols, cols)
        want = mode >= 2
        if want:
            res.outer = np.empty(cols + 1, dtype=np.int32)
            res.inner = np.empty(nnz, dtype=np.int32)
            res.values = np.empty(nnz)
        L.plugin_result_copy(h, res.g.ctypes.data, res.outer.ctypes.data if want else None, res.inner.ctypes.data if want else None,
                             res.values.ctypes.data if want else None)
        return res
    finally:
        L.plugin_result_free(h)


def ref_project(H, eps=1e-9):
    """TinyAD::project_positive_definite of the reference on one dense symmetric matrix."""
    A = np.array(H, dtype=np.float64, order="C")
    if ref_lib().ref_project(A.shape[0], A.ctypes.data, eps) < 0:
        raise RuntimeError(ref_lib().ref_last_error().decode())
    return A


def max_threads():
    return lib().oracle_max_threads()
