"""CPU tests: pin the oracle against the reference's own known-answer tests and fixtures."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle
from problems import planar_newton_problem, Problem, icosphere, one_ring_table
import tinyad_b200 as tad
from tinyad_b200 import meshes

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "scalar_cases.json")))


def check_case(c, results):
    assert len(results) == c["n_out"]
    tol = max(c["tol"], 1e-15)
    for (val, grad, hess), exp in zip(results, c["expected"]):
        if exp["val"] is not None:
            assert abs(val - exp["val"]) <= tol, (c["name"], c["params"], val, exp["val"])
        for i, g in enumerate(exp["grad"]):
            assert abs(grad[i] - g) <= tol, (c["name"], c["params"], "grad", i, grad[i], g)
        for i, row in enumerate(exp["hess"]):
            for j, h in enumerate(row):
                assert abs(hess[i, j] - h) <= tol, (c["name"], c["params"], "hess", i, j, hess[i, j], h)
        assert np.abs(hess - hess.T).max() <= 1e-12 * max(1.0, np.abs(hess).max())  # TINYAD_ASSERT_SYMMETRIC


@pytest.mark.parametrize("c", CASES, ids=[f"{c['name']}-{i}" for i, c in enumerate(CASES)])
def test_scalar_known_answers(c):
    check_case(c, oracle.scalar_case(c["name"], c["params"], c["k"], n_out_max=c["n_out"]))


def test_complex_ops_against_numpy():
    # ComplexTest.cc: complex arithmetic on active scalars == derivatives of the real / imaginary parts
    import sympy as sy
    x, y = sy.symbols("x y", real=True)
    px, py, br, bi = 0.7, -1.3, 0.4, 2.5
    a = x + sy.I * y
    b = (br + 0.5 * x) + sy.I * (bi - 0.25 * y)
    bd = br + sy.I * bi
    exprs = {"c_mul": a * b, "c_mul_d": a * bd, "c_d_mul": bd * a, "c_div": a / b, "c_div_d": a / bd, "c_add": a + b,
             "c_sub": a - b, "c_sqr": a * a, "c_conj": sy.conjugate(a)}
    for name, e in exprs.items():
        res = oracle.scalar_case(name, [px, py, br, bi], 2, n_out_max=2)
        for part, (val, grad, hess) in zip((sy.re(sy.expand(e)), sy.im(sy.expand(e))), res):
            part = sy.simplify(part)
            sub = {x: px, y: py}
            assert abs(val - float(part.subs(sub))) < 1e-12
            for i, v in enumerate((x, y)):
                assert abs(grad[i] - float(sy.diff(part, v).subs(sub))) < 1e-12, name
                for j, w in enumerate((x, y)):
                    assert abs(hess[i, j] - float(sy.diff(part, v, w).subs(sub))) < 1e-11, name
    for name, e in {"c_abs": sy.sqrt(x * x + y * y), "c_arg": sy.atan2(y, x)}.items():
        (val, grad, hess), = oracle.scalar_case(name, [px, py, br, bi], 2, n_out_max=1)
        sub = {x: px, y: py}
        assert abs(val - float(e.subs(sub))) < 1e-12
        for i, v in enumerate((x, y)):
            assert abs(grad[i] - float(sy.diff(e, v).subs(sub))) < 1e-12
            for j, w in enumerate((x, y)):
                assert abs(hess[i, j] - float(sy.diff(e, v, w).subs(sub))) < 1e-12


def test_symmetric_dirichlet_triangle_vs_sympy():
    # symm_dirich at x_init = {10,1,15,3,2,2}, rest {1,1,2,1,1,2} (ScalarTestHessianBlock.cc:50-90, ScalarTestMisc.cc:121-149)
    import sympy as sy
    X = sy.symbols("x0:6", real=True)
    ar, br, cr = sy.Matrix([1, 1]), sy.Matrix([2, 1]), sy.Matrix([1, 2])
    Mr = sy.Matrix.hstack(br - ar, cr - ar)
    a, b, c = sy.Matrix(X[0:2]), sy.Matrix(X[2:4]), sy.Matrix(X[4:6])
    M = sy.Matrix.hstack(b - a, c - a)
    J = M * Mr.inv()
    E = (J.T * J).trace() + (J.inv().T * J.inv()).trace()
    pt = [10.0, 1.0, 15.0, 3.0, 2.0, 2.0]
    sub = dict(zip(X, pt))
    (val, grad, hess), = oracle.scalar_case("symm_dirich6", pt + [1, 1, 2, 1, 1, 2], 6, n_out_max=1)
    assert abs(val - float(E.subs(sub))) < 1e-12 * abs(val)
    G = [float(sy.diff(E, v).subs(sub)) for v in X]
    H = np.array([[float(sy.diff(E, v, w).subs(sub)) for w in X] for v in X])
    assert np.abs(grad - G).max() < 1e-12 * np.abs(G).max()
    assert np.abs(hess - H).max() < 1e-12 * np.abs(H).max()


def to_csc(r):
    return sp.csc_matrix((r.values, r.inner, r.outer), shape=r.shape)


def test_newton_fixture():
    """tests/NewtonTest.cc:12-88: nnz == 4V + 8(V+F-1); 10 projected-Newton iterations reach f = 4, g = 0."""
    p, x = planar_newton_problem()
    terms = p.oracle_terms()
    r = oracle.scalar_eval(2, 6, terms, oracle.HESSIAN_PROJ, x)
    assert r.values.size == 4 * 6 + 8 * (6 + 4 - 1)                      # NewtonTest.cc:65
    assert abs(r.f - 24.5625) < 1e-12
    for _ in range(10):
        r = oracle.scalar_eval(2, 6, terms, oracle.HESSIAN_PROJ, x)
        H = to_csc(r) + 1e-9 * sp.identity(12, format="csc")          # newton_direction's w_identity (Utils/NewtonDirection.hh:25-48)
        dx = spla.spsolve(H, -r.g)
        s, f0 = 1.0, r.f                                                # Armijo backtracking (Utils/LineSearch.hh:26-65)
        for _ in range(64):
            f1 = oracle.scalar_eval(2, 6, terms, oracle.EVAL, x + s * dx).f
            if f1 <= f0 + 1e-4 * s * r.g.dot(dx):
                break
            s *= 0.8
        x = x + s * dx
    r = oracle.scalar_eval(2, 6, terms, oracle.HESSIAN_PROJ, x)
    assert abs(r.f - 4.0) < 1e-12 and np.abs(r.g).max() < 1e-10          # NewtonTest.cc:82-88
    assert abs(oracle.scalar_eval(2, 6, terms, oracle.EVAL, x).f - r.f) < 1e-15


def test_scalar_function_fixtures():
    # tests/ScalarFunctionTest.cc:72-146
    conn = np.array([[0]], dtype=np.int32)
    x = np.array([1.0, 2.0])
    r = oracle.scalar_eval(2, 1, [oracle.Term(oracle.QUADRATIC2D, conn, np.array([[1.0]]))], oracle.DERIVATIVES, x)
    assert r.f == 12.0 and list(r.g) == [9.0, 6.0] and list(r.values) == [4.0, 2.0, 2.0, 2.0]
    rp = oracle.scalar_eval(2, 1, [oracle.Term(oracle.QUADRATIC2D, conn, np.array([[1.0]]))], oracle.HESSIAN_PROJ, x)
    assert np.abs(rp.values - r.values).max() <= 1e-16                  # convex: H_proj == H (ScalarFunctionTest.cc:102-105)
    rn = oracle.scalar_eval(2, 1, [oracle.Term(oracle.QUADRATIC2D, conn, np.array([[-1.0]]))], oracle.HESSIAN_PROJ, x)
    assert rn.f == -12.0 and np.linalg.eigvalsh(rn.values.reshape(2, 2)).min() > 0   # :143-146
    # repeated handle -> same local slots, pattern only contains accessed variables (ScalarFunctionTest.cc:153-179)
    conn2 = np.array([[0, 2]], dtype=np.int32)
    r = oracle.scalar_eval(2, 3, [oracle.Term(oracle.REPEATED_HANDLE, conn2, np.zeros((1, 1)))], oracle.DERIVATIVES, np.arange(6.0))
    assert r.values.size == 16 and list(np.diff(r.outer)) == [4, 4, 0, 0, 4, 4]


def test_gauss_newton_twins_agree():
    """tests/GaussNewtonTest.cc:113-138: scalar and sum-of-squares formulations agree in f and g to 1e-12."""
    p, x = planar_newton_problem()
    V_rest, _, F, b, bc = tad.meshes.planar_test_mesh()
    data = tad.meshes.tri_rest_data(V_rest, F, weight=1.0 / np.sqrt(len(F)))
    vt = [oracle.Term(oracle.SOS_SYMDIRICHLET2D, F, data), oracle.Term(oracle.SOS_PENALTY2D, b.reshape(-1, 1), bc)]
    rng = np.random.default_rng(0)
    for it in range(5):
        xi = x + 0.01 * it * rng.standard_normal(x.size)
        ref = oracle.scalar_eval(2, 6, p.oracle_terms(), oracle.GRADIENT, xi)
        assert abs(oracle.vector_eval(2, 6, vt, oracle.V_SOS, xi).f - ref.f) < 1e-12
        r = oracle.vector_eval(2, 6, vt, oracle.V_SOS_DERIVATIVES, xi)
        assert abs(r.f - ref.f) < 1e-12 and np.abs(r.g - ref.g).max() < 1e-12
        assert r.shape == (4 * 8 + 2 * 2, 12) and r.values.size == 4 * 8 * 6 + 2 * 2 * 2
        J = to_csc(r)
        assert np.abs(2.0 * J.T @ r.r - r.g).max() < 1e-12


def test_hessian_is_graph_laplacian():
    """tests/DynamicElementsTest.cc:60-91: Hessian of the edge Dirichlet energy == graph Laplacian, same nnz."""
    rng = np.random.default_rng(1)
    nv = 30
    edges = {tuple(sorted(e)) for e in rng.integers(0, nv, size=(80, 2)) if e[0] != e[1]}
    conn = np.array(sorted(edges), dtype=np.int32)
    r = oracle.scalar_eval(1, nv, [oracle.Term(oracle.EDGE_DIRICHLET1D, conn, np.full((len(conn), 1), 0.5))], oracle.DERIVATIVES, rng.random(nv))
    L = np.zeros((nv, nv))
    for a, b in conn:
        L[a, a] += 1; L[b, b] += 1; L[a, b] -= 1; L[b, a] -= 1
    assert np.abs(to_csc(r).toarray() - L).max() < 1e-12
    assert r.values.size == np.count_nonzero(L)


def test_projection_matches_numpy_eigh():
    """project_positive_definite vs numpy: V max(L, eps) V^T, both early-outs leave H bit-unchanged."""
    rng = np.random.default_rng(2)
    for k in (2, 3, 6, 9, 12):
        A = rng.standard_normal((k, k)); A = A + A.T
        P, code = oracle.project(A, 1e-9)
        w, V = np.linalg.eigh(A)
        assert code == 2 and np.abs(P - (V * np.maximum(w, 1e-9)) @ V.T).max() < 1e-12 * np.abs(A).max()
        P, code = oracle.project(A, -1.0)                                  # |lambda| mode
        assert np.abs(P - (V * np.abs(w)) @ V.T).max() < 1e-12 * np.abs(A).max()
        D = A + 100.0 * k * np.eye(k)                                      # diagonally dominant: early-out 1
        P, code = oracle.project(D, 1e-9)
        assert code == 0 and np.array_equal(P, D)
        B = A @ A.T + 0.5 * np.eye(k)                                       # PD but not dominant: early-out 2
        P, code = oracle.project(B, 1e-9)
        assert code in (0, 1) and np.array_equal(P, B)


def test_set_from_triplets_semantics():
    """Pattern / values of the oracle's COO->CSC == scipy's coo->csc with summed duplicates, zeros kept."""
    p, x = planar_newton_problem()
    r = oracle.scalar_eval(2, 6, p.oracle_terms(), oracle.DERIVATIVES, x)
    M = to_csc(r)
    assert M.has_sorted_indices and np.abs((M - M.T)).max() < 1e-12
    # explicit zeros stay: 96 stored entries although some couplings vanish numerically
    assert M.nnz == 96


def test_inverted_element_returns_infinity():
    p, x = planar_newton_problem()
    x = x.copy(); x[2], x[4] = x[4], x[2]
    assert oracle.scalar_eval(2, 6, p.oracle_terms(), oracle.EVAL, x).f == np.inf
    r = oracle.scalar_eval(2, 6, p.oracle_terms(), oracle.HESSIAN_PROJ, x)   # val = inf is not an error
    assert r.f == np.inf and np.all(np.isfinite(r.g)) and np.all(np.isfinite(r.values))


def reference_laplace(n_vertices, F):
    """DynamicElementsTest.cc:38-56."""
    L = sp.lil_matrix((n_vertices, n_vertices))
    for f in F:
        for i in range(3):
            v1, v2 = int(f[i]), int(f[(i + 1) % 3])
            L[v1, v1] += 1.0
            L[v1, v2] -= 1.0
    return L.tocsc()


def test_dynamic_elements_fixture():
    """tests/DynamicElementsTest.cc:9-33: add_elements_dynamic<3, 1>, element e accesses e variables; x = ones.
    Groups: valence 3 <- elements {2, 3}, valence 1 <- elements {0, 1}.  f = sum_e |e * (1,1)|^2 = 2 (0 + 1 + 4 + 9) = 28."""
    dummy = np.zeros((4, 1), dtype=np.int32)
    t = [oracle.Term(oracle.DYN_SUM_SQR2D, dummy, np.zeros((4, 1)))]
    x = np.ones(8)
    r = oracle.scalar_eval(2, 4, t, oracle.HESSIAN_PROJ, x)
    assert r.f == 28.0
    # gradient of |s|^2 w.r.t. every accessed variable is 2 s: vertex v is accessed by the elements e > v
    g = np.zeros(8)
    for e in range(4):
        for v in range(e):
            g[2 * v:2 * v + 2] += 2.0 * e
    assert np.array_equal(r.g, g)
    H = to_csc(r).toarray()
    assert np.allclose(H, H.T) and np.linalg.eigvalsh(H).min() > -1e-12      # PSD energy: projection leaves it PSD


def test_dynamic_one_ring_laplacian():
    """tests/DynamicElementsTest.cc:92-141: Hessian of the one-ring Dirichlet energy (dynamic vertex elements) == Laplacian to 1e-12;
    the nnz differ because padded elements add explicit zeros (comment at :140)."""
    V, F = icosphere(1)                     # closed mesh, 12 vertices of valence 5 and 30 of valence 6 -> two groups (6 and 7)
    tab = one_ring_table(len(V), F)
    assert sorted(set((tab >= 0).sum(axis=1))) == [5, 6]
    t = [oracle.Term(oracle.DYN_ONERING1D, tab, np.zeros(tab.shape))]
    r = oracle.scalar_eval(1, len(V), t, oracle.DERIVATIVES, np.zeros(len(V)))
    L = reference_laplace(len(V), F)
    assert abs(to_csc(r) - L).max() < 1e-12
    assert spla.norm(to_csc(r) - L) < 1e-12
