"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import oracle
import tinyad_b200 as tad
from conftest import TOL_H, TOL_H_PROJ, assert_f, assert_pattern, assert_vec
from problems import Problem, grid_problem, planar_newton_problem, tet_problem

pytestmark = pytest.mark.gpu


def check_all_modes(p, x, assembly, tol_proj=TOL_H_PROJ, eps=1e-9):
    ref_h = oracle.scalar_eval(p.d, p.n_vertices, p.oracle_terms(), oracle.DERIVATIVES, x)
    ref_p = oracle.scalar_eval(p.d, p.n_vertices, p.oracle_terms(), oracle.HESSIAN_PROJ, x, eps=eps)
    fn = p.gpu(assembly=assembly)
    try:
        # eval
        assert_f(fn.eval_host(x), ref_h.f)
        # eval_with_gradient
        f, g = fn.eval_with_gradient_host(x)
        assert_f(f, ref_h.f)
        assert_vec(g, ref_h.g)
        # pattern: bit-exact index arrays
        outer, inner = fn.pattern()
        assert_pattern(outer, inner, ref_h)
        # eval_with_derivatives
        f, g, H = fn.eval_with_derivatives_host(x)
        assert_f(f, ref_h.f)
        assert_vec(g, ref_h.g)
        assert_vec(H, ref_h.values, tol=TOL_H)
        # eval_with_hessian_proj
        f, g, H = fn.eval_with_hessian_proj_host(x, eps=eps)
        assert_f(f, ref_p.f)
        assert_vec(g, ref_p.g)
        assert_vec(H, ref_p.values, tol=tol_proj)
        st = fn.projection_stats()
        assert st["decomposed"] == ref_p.phases["n_decomposed"]
        if eps >= 0:  # in |lambda| mode the sign of a numerically zero eigenvalue decides, which no solver pins
            assert st["rebuilt"] == ref_p.phases["n_rebuilt"]
    finally:
        fn.close()


@pytest.mark.parametrize("assembly", [tad.ASSEMBLY_ATOMIC, tad.ASSEMBLY_GATHER])
def test_planar_newton_mesh(torch_cuda, assembly):
    p, x = planar_newton_problem()
    check_all_modes(p, x, assembly)
    fn = p.gpu(assembly=assembly)
    # golden value of the reference: nnz == 4V + 8(V+F-1)  (tests/NewtonTest.cc:65)
    assert fn.nnz == 4 * 6 + 8 * (6 + 4 - 1)
    fn.close()


@pytest.mark.parametrize("assembly", [tad.ASSEMBLY_ATOMIC, tad.ASSEMBLY_GATHER])
def test_grid_triangles(torch_cuda, assembly):
    p, x = grid_problem(24, seed=1, with_penalty=True)
    check_all_modes(p, x, assembly)


@pytest.mark.parametrize("assembly", [tad.ASSEMBLY_ATOMIC, tad.ASSEMBLY_GATHER])
def test_kuhn_tets(torch_cuda, assembly):
    p, x = tet_problem(5, seed=2, with_penalty=True)
    check_all_modes(p, x, assembly)


def test_abs_eigenvalue_mode(torch_cuda):
    # eps < 0: negative eigenvalues are replaced by their absolute value (HessianProjection.hh:73-81)
    p, x = grid_problem(8, seed=3)
    check_all_modes(p, x, tad.ASSEMBLY_ATOMIC, eps=-1.0)


def test_inverted_element_gives_infinity(torch_cuda):
    # return INFINITY is not an error; f propagates inf, derivatives of that element are zero
    p, x = planar_newton_problem()
    x = x.copy()
    x[2], x[4] = x[4], x[2]  # swap x of vertices 1 and 2 -> first triangle flips
    ref = oracle.scalar_eval(p.d, p.n_vertices, p.oracle_terms(), oracle.HESSIAN_PROJ, x)
    assert np.isinf(ref.f)
    fn = p.gpu()
    assert fn.eval_host(x) == np.inf
    f, g, H = fn.eval_with_hessian_proj_host(x)
    assert f == np.inf
    assert_vec(g, ref.g)
    assert_vec(H, ref.values, tol=TOL_H_PROJ)
    fn.close()


def test_misc_energies(torch_cuda):
    rng = np.random.default_rng(5)
    nv = 40
    conn = np.stack([rng.permutation(nv)[:2] for _ in range(60)]).astype(np.int32)
    data = rng.random((60, 1)) + 0.5
    x = rng.random(2 * nv) * 2.0
    for kind in (tad.TRIG_MIX2D, tad.REPEATED_HANDLE):
        p = Problem(2, nv, [(kind, conn, data)])
        check_all_modes(p, x, tad.ASSEMBLY_ATOMIC)
        check_all_modes(p, x, tad.ASSEMBLY_GATHER)
    # 1-D edge Dirichlet energy: Hessian == graph Laplacian (tests/DynamicElementsTest.cc:60-91)
    p = Problem(1, nv, [(tad.EDGE_DIRICHLET1D, conn, np.full((60, 1), 0.5))])
    xs = rng.random(nv)
    check_all_modes(p, xs, tad.ASSEMBLY_ATOMIC)
    fn = p.gpu()
    _, _, H = fn.eval_with_derivatives_host(xs)
    outer, inner = fn.pattern()
    import scipy.sparse as sp
    Hs = sp.csr_matrix((H, inner, outer), shape=(nv, nv)).toarray()
    L = np.zeros((nv, nv))
    for a, b in conn:
        L[a, a] += 1; L[b, b] += 1; L[a, b] -= 1; L[b, a] -= 1
    assert np.abs(Hs - L).max() < 1e-12
    fn.close()


def test_quadratic_known_answers(torch_cuda):
    # tests/ScalarFunctionTest.cc:72-146: exact f, g, H of a convex / non-convex quadratic at (1, 2)
    conn = np.array([[0]], dtype=np.int32)
    x = np.array([1.0, 2.0])
    fn = Problem(2, 1, [(tad.QUADRATIC2D, conn, np.array([[1.0]]))]).gpu()
    f, g, H = fn.eval_with_derivatives_host(x)
    assert f == 12.0 and np.array_equal(g, [9.0, 6.0]) and np.array_equal(H, [4.0, 2.0, 2.0, 2.0])
    f, g, Hp = fn.eval_with_hessian_proj_host(x)
    assert np.abs(Hp - H).max() <= 1e-16 * 0 + 1e-15  # already PD: H_proj == H (ScalarFunctionTest.cc:102-105)
    fn.close()
    fn = Problem(2, 1, [(tad.QUADRATIC2D, conn, np.array([[-1.0]]))]).gpu()
    f, g, H = fn.eval_with_derivatives_host(x)
    assert f == -12.0 and np.array_equal(g, [-9.0, -6.0]) and np.array_equal(H, [-4.0, -2.0, -2.0, -2.0])
    f, g, Hp = fn.eval_with_hessian_proj_host(x)
    w = np.linalg.eigvalsh(Hp.reshape(2, 2))
    assert w.min() > 0.0  # ScalarFunctionTest.cc:143-146
    fn.close()


def test_arap_with_closest_orthogonal(torch_cuda):
    """Operations/SVD.hh (closest_orthogonal) inside an element functor, Double<6>: the 2-D ARAP energy
    w |J - R(J)|^2 on a deformed grid; f, g, H and projected H against the oracle's restatement of Operations/SVD.hh:70-99."""
    from tinyad_b200 import meshes
    V, F = meshes.grid_2d(10)
    data = meshes.tri_rest_data(V, F)
    x = meshes.deform(V, 1.0 / 10, seed=7).reshape(-1)
    p = Problem(2, len(V), [(tad.ARAP2D, F, data)])
    for mode, project in ((oracle.DERIVATIVES, False), (oracle.HESSIAN_PROJ, True)):
        ref = oracle.scalar_eval(2, len(V), p.oracle_terms(), mode, x)
        fn = p.gpu()
        xd = torch_cuda.from_numpy(x).cuda()
        g = torch_cuda.empty(fn.n_vars, dtype=torch_cuda.float64, device="cuda")
        H = torch_cuda.empty(fn.nnz, dtype=torch_cuda.float64, device="cuda")
        f = fn.eval_with_derivatives(xd, g, H, project=project)
        outer, inner = fn.pattern()
        assert_pattern(outer, inner, ref)
        assert_f(f, ref.f)
        assert_vec(g.cpu().numpy(), ref.g)
        assert_vec(H.cpu().numpy(), ref.values, tol=TOL_H_PROJ if project else 1e-11)
        fn.close()
