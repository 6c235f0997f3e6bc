"""CPU: the oracle (oracle/tinyad_oracle.hh, the restatement every GPU parity test is checked against) versus the reference
ITSELF: oracle/_ref/libtinyad_ref.so is the unmodified TinyAD headers of /root/reference compiled in place over
oracle/eigen_shim (oracle/ref_driver.cc, `make -C oracle _ref`).  The arithmetic of TinyAD::Scalar, Element, the objective
terms, ScalarFunction / VectorFunction and project_positive_definite's control flow is the reference's own code; the dense
and sparse primitives underneath are the shim's, not Eigen's (Eigen is not in the image).

Bars: sparsity patterns bit-exact; f, g, r, J and unprojected H within 1e-13 relative to the largest magnitude (they are
bit-identical on most inputs: same operation order); projected H within 1e-10 (two different eigen-solvers: cyclic Jacobi in
the oracle, Householder + QL in the shim).

Nothing here reads /root/reference at run time; the tests skip when the prebuilt library is absent.
"""
import os

import numpy as np
import pytest

import oracle
from problems import Problem, grid_problem, icosphere, one_ring_table, planar_newton_problem, tet_problem
from tinyad_b200 import meshes

pytestmark = pytest.mark.skipif(not (oracle.ref_available() or oracle.build_ref()), reason="oracle/_ref is not built")

TOL = 1e-13
TOL_PROJ = 1e-10


def close(a, b, tol):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    if a.size == 0:
        return 0.0
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    m = np.isfinite(a)
    scale = max(np.abs(b[m]).max(initial=0.0), 1e-300)
    err = np.abs(a[m] - b[m]).max(initial=0.0) / scale
    assert err <= tol, f"relative error {err:.3e} > {tol:.1e}"
    return err


def compare_scalar(p, x, eps=1e-9, modes=(0, 1, 2, 3), n_threads=-1):
    ot = p.oracle_terms()
    worst = 0.0
    for mode in modes:
        a = oracle.scalar_eval(p.d, p.n_vertices, ot, mode, x, eps=eps, n_threads=n_threads)
        b = oracle.ref_scalar_eval(p.d, p.n_vertices, ot, mode, x, eps=eps, n_threads=n_threads)
        if np.isfinite(b.f):
            assert abs(a.f - b.f) <= TOL * max(abs(b.f), 1e-300)
        else:
            assert a.f == b.f or (np.isnan(a.f) and np.isnan(b.f))
        if mode >= 1:
            close(a.g, b.g, TOL)
        if mode >= 2:
            assert a.shape == b.shape
            assert np.array_equal(a.outer, b.outer) and np.array_equal(a.inner, b.inner)      # bit-exact pattern
            worst = max(worst, close(a.values, b.values, TOL if mode == 2 else TOL_PROJ))
    return worst


def test_reference_library_identifies_itself():
    assert b"unmodified" in oracle.ref_lib().ref_description()
    assert oracle.ref_lib().ref_default_threads() == oracle.default_threads()                 # Detail/Parallel.hh:21-32


def test_planar_newton_fixture():
    """tests/NewtonTest.cc:12-58 on its own four-triangle mesh."""
    p, x = planar_newton_problem()
    compare_scalar(p, x)
    b = oracle.ref_scalar_eval(2, p.n_vertices, p.oracle_terms(), oracle.HESSIAN_PROJ, x)
    assert b.f == 24.5625 and len(b.inner) == 96               # tests/NewtonTest.cc:66-69: nnz == 4V + 8(V+F-1)


@pytest.mark.parametrize("N,seed", [(6, 0), (17, 3), (40, 11)])
def test_triangle_grids(N, seed):
    p, x = grid_problem(N, seed=seed, with_penalty=True)
    compare_scalar(p, x)


@pytest.mark.parametrize("n,seed", [(2, 0), (5, 1), (9, 4)])
def test_tet_cubes(n, seed):
    """BASELINE.json configs C2 / C5 in small: Double<12> symmetric Dirichlet on Kuhn cubes, most element Hessians indefinite."""
    p, x = tet_problem(n, seed=seed, with_penalty=True)
    compare_scalar(p, x)


@pytest.mark.parametrize("n", [30] + ([55] if os.environ.get("TAD_FULLSIZE") == "1" else []))
def test_large_tet_cube(n):
    """162 000 tets by default; TAD_FULLSIZE=1 adds BASELINE.json's C2 itself (n = 55: 998 250 tets, 23 036 814 nonzeros, ~40 s and
    ~10 GB).  Measured at n = 55: pattern bit-exact, f identical, g 5.8e-16, projected H 4.0e-15 -- the full-size GPU test
    (tests/test_fullsize_gpu.py) compares every CSR row of the same problem with the oracle."""
    p, x = tet_problem(n, seed=0)
    assert compare_scalar(p, x, modes=(3,)) <= 1e-13


def test_tets_strongly_deformed_and_abs_eps():
    """Larger deformation (more negative eigenvalues, some inverted elements -> f = inf) and the absolute-value strategy
    of Utils/HessianProjection.hh:72-80 (eps < 0)."""
    V, T = meshes.kuhn_cube(6, 6, 6)
    data = meshes.tet_rest_data(V, T)
    p = Problem(3, len(V), [(oracle.SYMDIRICHLET3D, T, data)])
    x = meshes.deform(V, 1.0 / 6, seed=2, noise=0.3).reshape(-1)
    compare_scalar(p, x, modes=(2, 3))
    compare_scalar(p, x, eps=-1.0, modes=(3,))
    compare_scalar(p, x, eps=1e-3, modes=(3,))
    rng = np.random.default_rng(0)
    xi = x + 0.5 / 6 * rng.standard_normal(x.shape)            # inverted tets: the lambda returns (T)INFINITY
    a = oracle.scalar_eval(3, len(V), p.oracle_terms(), 0, xi)
    b = oracle.ref_scalar_eval(3, len(V), p.oracle_terms(), 0, xi)
    assert np.isinf(b.f) and a.f == b.f


def test_single_thread_equals_many():
    p, x = tet_problem(4, seed=5)
    ot = p.oracle_terms()
    r1 = oracle.ref_scalar_eval(3, p.n_vertices, ot, 3, x, n_threads=1)
    r8 = oracle.ref_scalar_eval(3, p.n_vertices, ot, 3, x, n_threads=8)
    assert r1.f == r8.f and np.array_equal(r1.g, r8.g) and np.array_equal(r1.values, r8.values)   # serial accumulation


def test_misc_energies():
    """Every unary / binary operator family of Scalar.hh inside ScalarFunction, the repeated-handle case of
    tests/ScalarFunctionTest.cc:153-179, Operations/SVD.hh and the 1-D graph Laplacian."""
    rng = np.random.default_rng(5)
    nv = 40
    conn = np.stack([rng.permutation(nv)[:2] for _ in range(60)]).astype(np.int32)
    data = rng.random((60, 1)) + 0.5
    x = rng.random(2 * nv) * 2.0
    for kind in (oracle.TRIG_MIX2D, oracle.REPEATED_HANDLE):
        compare_scalar(Problem(2, nv, [(kind, conn, data)]), x)
    compare_scalar(Problem(1, nv, [(oracle.EDGE_DIRICHLET1D, conn, np.full((60, 1), 0.5))]), rng.random(nv))
    one = np.array([[0]], dtype=np.int32)
    for sign in (1.0, -1.0):
        compare_scalar(Problem(2, 1, [(oracle.QUADRATIC2D, one, np.array([[sign]]))]), np.array([1.0, 2.0]))
    V, F = meshes.grid_2d(10)
    p = Problem(2, len(V), [(oracle.ARAP2D, F, meshes.tri_rest_data(V, F))])
    compare_scalar(p, meshes.deform(V, 0.1, seed=7).reshape(-1))


def test_quadratic_known_answers_from_the_reference():
    """tests/ScalarFunctionTest.cc:72-146 evaluated by the reference: exact f, g, H at (1, 2)."""
    one = np.array([[0]], dtype=np.int32)
    r = oracle.ref_scalar_eval(2, 1, [oracle.Term(oracle.QUADRATIC2D, one, np.array([[1.0]]))], 2, np.array([1.0, 2.0]))
    assert r.f == 12.0 and np.array_equal(r.g, [9.0, 6.0]) and np.array_equal(r.values, [4.0, 2.0, 2.0, 2.0])


def test_dynamic_elements():
    """tests/DynamicElementsTest.cc:9-33 (<3, 1>, f = 28) and :92-141 (one-ring elements, groups by valence)."""
    dummy = np.zeros((4, 1), dtype=np.int32)
    p = Problem(2, 4, [(oracle.DYN_SUM_SQR2D, dummy, np.zeros((4, 1)))])
    compare_scalar(p, np.ones(8))
    assert oracle.ref_scalar_eval(2, 4, p.oracle_terms(), 0, np.ones(8)).f == 28.0
    V, F = icosphere(1)
    tab = one_ring_table(len(V), F)
    p = Problem(1, len(V), [(oracle.DYN_ONERING1D, tab, np.zeros(tab.shape))])
    compare_scalar(p, np.zeros(len(V)))
    compare_scalar(p, np.random.default_rng(1).standard_normal(len(V)))


def test_vector_functions():
    """VectorFunction::eval / eval_with_jacobian / eval_sum_of_squares / ..._with_derivatives (Detail/VectorFunctionImpl.hh:143-301)
    on tests/GaussNewtonTest.cc's terms and the complex-valued residual."""
    V_rest, V_init, F, b, bc = meshes.planar_test_mesh()
    cases = [(len(V_rest), [(oracle.SOS_SYMDIRICHLET2D, F, meshes.tri_rest_data(V_rest, F, weight=1.0 / np.sqrt(len(F)))),
                            (oracle.SOS_PENALTY2D, b.reshape(-1, 1), bc)], V_init.reshape(-1).copy())]
    N = 14
    V, F = meshes.grid_2d(N)
    bb = np.array([[0], [N], [(N + 1) * N]], dtype=np.int32)
    cases.append((len(V), [(oracle.SOS_SYMDIRICHLET2D, F, meshes.tri_rest_data(V, F, weight=1.0 / np.sqrt(len(F)))),
                           (oracle.SOS_PENALTY2D, bb, V[bb[:, 0]] + 0.02)], meshes.deform(V, 1.0 / N, seed=5).reshape(-1)))
    rng = np.random.default_rng(9)
    edges = np.unique(np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), axis=1), axis=0).astype(np.int32)
    e = V[edges[:, 1]] - V[edges[:, 0]]
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    cases.append((len(V), [(oracle.SOS_POLYCURL2D, edges, np.concatenate([e, rng.random((len(edges), 1)) + 0.5], axis=1))],
                  rng.standard_normal(2 * len(V))))
    for nv, terms, x in cases:
        ot = [oracle.Term(k, c, d) for k, c, d in terms]
        for mode in range(4):
            a = oracle.vector_eval(2, nv, ot, mode, x)
            r = oracle.ref_vector_eval(2, nv, ot, mode, x)
            if mode in (0, 1, 3):
                close(a.r, r.r, TOL)
            if mode in (2, 3):
                assert abs(a.f - r.f) <= TOL * abs(r.f)
            if mode in (1, 3):
                assert a.shape == r.shape and np.array_equal(a.outer, r.outer) and np.array_equal(a.inner, r.inner)
                close(a.values, r.values, TOL)
            if mode == 3:
                close(a.g, r.g, 1e-12)          # g = 2 J^T r: the summation order inside the sparse product is the shim's


@pytest.mark.parametrize("k", [2, 3, 4, 6, 9, 12])
def test_project_positive_definite(k):
    """Utils/HessianProjection.hh:48-101 on random symmetric matrices: indefinite, already PD (early out, bit-identical),
    diagonally dominant, rank deficient, and the eps < 0 strategy."""
    rng = np.random.default_rng(k)
    for trial in range(40):
        A = rng.standard_normal((k, k))
        H = A + A.T
        if trial % 5 == 1:
            H = A @ A.T + 0.1 * np.eye(k)                      # PD
        if trial % 5 == 2:
            H = np.diag(np.abs(H).sum(axis=1) + 1.0) + 0.5 * H  # positive diagonally dominant
        if trial % 5 == 3:
            v = rng.standard_normal((k, max(1, k // 2)))
            H = v @ v.T - 0.3 * np.outer(v[:, 0], v[:, 0])      # rank deficient, one negative direction
        for eps in (1e-9, -1.0):
            a, _ = oracle.project(H, eps)
            b = oracle.ref_project(H, eps)
            assert np.abs(a - b).max() <= TOL_PROJ * max(np.abs(H).max(), 1.0)
            w, Q = np.linalg.eigh(H)
            w2 = np.abs(w) if eps < 0 else np.maximum(w, eps)
            scale = max(np.abs(H).max(), 1.0)
            clearly_kept = w.min() > (0.0 if eps < 0 else eps) + 1e-10 * scale     # away from the threshold by more than rounding
            if clearly_kept or (np.diag(H) >= np.abs(H).sum(axis=1) - np.abs(np.diag(H)) + eps).all():
                assert np.array_equal(b, H)                    # untouched (HessianProjection.hh:62-63 and :92-93)
            assert np.abs(b - (Q * w2) @ Q.T).max() <= 1e-12 * scale


def test_edge_cases_and_error_behaviour():
    """Empty element range, a function without terms, and the two run-time errors of the path: an out-of-range variable handle
    (Detail/Element.hh:160-170) and non-finite derivatives (Detail/ScalarObjectiveTerm.hh:250-253) throw in both."""
    V, F = meshes.grid_2d(3)
    x = V.reshape(-1).copy()
    nv = len(V)
    empty = [oracle.Term(oracle.SYMDIRICHLET2D, np.zeros((0, 3), dtype=np.int32), np.zeros((0, 5)))]
    for terms in (empty, []):
        for mode in range(4):
            a = oracle.scalar_eval(2, nv, terms, mode, x)
            b = oracle.ref_scalar_eval(2, nv, terms, mode, x)
            assert a.f == b.f == 0.0
            if mode >= 1:
                assert np.array_equal(a.g, b.g) and not a.g.any() and len(a.g) == 2 * nv
            if mode >= 2:
                assert a.shape == b.shape == (2 * nv, 2 * nv) and len(a.inner) == len(b.inner) == 0 and np.array_equal(a.outer, b.outer)
    bad = [oracle.Term(oracle.PENALTY2D, np.array([[nv]], dtype=np.int32), np.zeros((1, 2)))]
    xn = x.copy()
    xn[3] = np.nan
    good = [oracle.Term(oracle.SYMDIRICHLET2D, F, meshes.tri_rest_data(V, F))]
    for fn in (oracle.scalar_eval, oracle.ref_scalar_eval):
        with pytest.raises(RuntimeError):
            fn(2, nv, bad, oracle.HESSIAN_PROJ, x)
        with pytest.raises(RuntimeError):
            fn(2, nv, good, oracle.HESSIAN_PROJ, xn)
        assert np.isnan(fn(2, nv, good, oracle.EVAL, xn).f)           # the passive pass does not check (ScalarObjectiveTerm.hh:162-187)
