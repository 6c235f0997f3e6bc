"""GPU: VectorFunction path (residuals, Jacobian in the reference's CSC layout, sum of squares, g = 2 J^T r) vs the oracle."""
import numpy as np
import pytest

import oracle
import tinyad_b200 as tad
from conftest import assert_f, assert_vec
from problems import Problem
from tinyad_b200 import meshes

pytestmark = pytest.mark.gpu


def sos_problem(N=None):
    if N is None:
        V_rest, V_init, F, b, bc = meshes.planar_test_mesh()      # tests/GaussNewtonTest.cc:12-72
        x = V_init.reshape(-1).copy()
    else:
        V_rest, F = meshes.grid_2d(N)
        x = meshes.deform(V_rest, 1.0 / N, seed=5).reshape(-1)
        b = np.array([0, N, (N + 1) * N], dtype=np.int32)
        bc = V_rest[b] + 0.02
    data = meshes.tri_rest_data(V_rest, F, weight=1.0 / np.sqrt(len(F)))
    terms = [(tad.SOS_SYMDIRICHLET2D, F, data), (tad.SOS_PENALTY2D, b.reshape(-1, 1), bc)]
    return Problem(2, len(V_rest), terms, is_vector=True), x


def polycurl_problem(N=12):
    V, F = meshes.grid_2d(N)
    rng = np.random.default_rng(9)
    # per "edge" element: two neighbouring faces' frame variables; here: variables live on vertices, elements on mesh edges
    edges = np.unique(np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), axis=1), axis=0).astype(np.int32)
    e = V[edges[:, 1]] - V[edges[:, 0]]
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    data = np.concatenate([e, rng.random((len(edges), 1)) + 0.5], axis=1)
    x = rng.standard_normal(2 * len(V))
    return Problem(2, len(V), [(tad.SOS_POLYCURL2D, edges, data)], is_vector=True), x


@pytest.mark.parametrize("make", [lambda: sos_problem(), lambda: sos_problem(16), polycurl_problem])
def test_vector_function(torch_cuda, make):
    torch = torch_cuda
    p, x = make()
    ot = p.oracle_terms()
    ref = oracle.vector_eval(2, p.n_vertices, ot, oracle.V_SOS_DERIVATIVES, x)
    fn = p.gpu()
    try:
        outer, inner = fn.pattern()
        assert np.array_equal(outer, ref.outer) and np.array_equal(inner, ref.inner)      # CSC pattern bit-exact
        m, nnz = fn.n_outputs, len(inner)
        assert m == len(ref.r)
        xd = torch.from_numpy(x).cuda()
        r = torch.empty(m, dtype=torch.float64, device="cuda")
        J = torch.empty(nnz, dtype=torch.float64, device="cuda")
        g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
        fn.veval(xd, r)
        assert_vec(r.cpu().numpy(), ref.r)
        r.zero_()
        fn.veval_with_jacobian(xd, r, J)
        assert_vec(r.cpu().numpy(), ref.r)
        assert_vec(J.cpu().numpy(), ref.values)
        assert_f(fn.veval_sum_of_squares(xd), oracle.vector_eval(2, p.n_vertices, ot, oracle.V_SOS, x).f)
        f = fn.veval_sum_of_squares_with_derivatives(xd, g, r, J)
        assert_f(f, ref.f)
        assert_vec(g.cpu().numpy(), ref.g)
        assert_vec(J.cpu().numpy(), ref.values)
    finally:
        fn.close()


def test_sos_equals_scalar_formulation(torch_cuda):
    """tests/GaussNewtonTest.cc:113-138: scalar and sum-of-squares formulations agree in f and g to 1e-12 (on the GPU)."""
    torch = torch_cuda
    from problems import planar_newton_problem
    ps, x = planar_newton_problem()
    pv, _ = sos_problem()
    fs, fv = ps.gpu(), pv.gpu()
    f_ref, g_ref = fs.eval_with_gradient_host(x)
    xd = torch.from_numpy(x).cuda()
    g = torch.empty(fv.n_vars, dtype=torch.float64, device="cuda")
    r = torch.empty(fv.n_outputs, dtype=torch.float64, device="cuda")
    J = torch.empty(fv.nnz, dtype=torch.float64, device="cuda")
    f = fv.veval_sum_of_squares_with_derivatives(xd, g, r, J)
    assert abs(f - f_ref) < 1e-12 and np.abs(g.cpu().numpy() - g_ref).max() < 1e-12
    fs.close(); fv.close()


def test_eval_with_derivatives_golden(torch_cuda):
    """tests/VectorFunctionTest.cc:73-140 (test_eval): R^2 -> R^3, residuals (2 x0, x0^2, x1^2) at x = (3, 4): r, J and the Hessian of
    every residual, exact values."""
    torch = torch_cuda
    fn = tad.Function(1, 2, is_vector=True)
    fn.add_term(tad.SOS_TEST1D_A, np.array([[0]], dtype=np.int32), np.zeros((1, 1)))
    fn.add_term(tad.SOS_TEST1D_B, np.array([[1]], dtype=np.int32), np.zeros((1, 1)))
    try:
        outer, inner = fn.pattern()
        assert fn.n_outputs == 3 and list(outer) == [0, 2, 3] and list(inner) == [0, 1, 2]
        total = fn.residual_hessian_layout(-1)[3]
        assert total == 3 and fn.residual_hessian_layout(0)[:3] == (0, 1, 2) and fn.residual_hessian_layout(1)[:3] == (2, 1, 1)
        xd = torch.tensor([3.0, 4.0], dtype=torch.float64, device="cuda")
        r = torch.empty(3, dtype=torch.float64, device="cuda")
        J = torch.empty(3, dtype=torch.float64, device="cuda")
        Hb = torch.empty(total, dtype=torch.float64, device="cuda")
        fn.veval_with_derivatives(xd, r, J, Hb)
        assert r.cpu().tolist() == [6.0, 9.0, 16.0]
        assert J.cpu().tolist() == [2.0, 6.0, 8.0]            # d r0 / d x0, d r1 / d x0, d r2 / d x1
        assert Hb.cpu().tolist() == [0.0, 2.0, 2.0]           # H_0 = 0, H_1(0,0) = 2, H_2(1,1) = 2
    finally:
        fn.close()


def test_per_residual_hessians_sum_to_the_scalar_hessian(torch_cuda):
    """H(sum_i r_i^2) = 2 (J^T J + sum_i r_i H_i): the per-residual Hessians of the sum-of-squares formulation
    (tests/GaussNewtonTest.cc:34-72, 8 residuals x Double<6> per triangle) must reproduce the unprojected Hessian of the scalar twin
    (tests/NewtonTest.cc:28-55), which the scalar tests compare with the oracle."""
    import scipy.sparse as sp
    torch = torch_cuda
    N = 16
    V, F = meshes.grid_2d(N)
    x = meshes.deform(V, 1.0 / N, seed=5).reshape(-1)
    s = 1.0 / np.sqrt(len(F))
    b = np.array([[0], [N], [(N + 1) * N]], dtype=np.int32)
    bc = V[b[:, 0]] + 0.02
    fv = tad.Function(2, len(V), is_vector=True)
    fv.add_term(tad.SOS_SYMDIRICHLET2D, F, meshes.tri_rest_data(V, F, weight=s))
    fv.add_term(tad.SOS_PENALTY2D, b, bc)
    fs = tad.Function(2, len(V))
    fs.add_term(tad.SYMDIRICHLET2D, F, meshes.tri_rest_data(V, F, weight=s * s))
    fs.add_term(tad.PENALTY2D, b, bc)
    try:
        n, m = fv.n_vars, fv.n_outputs
        outer, inner = fv.pattern()
        total = fv.residual_hessian_layout(-1)[3]
        xd = torch.from_numpy(x).cuda()
        r = torch.empty(m, dtype=torch.float64, device="cuda")
        J = torch.empty(len(inner), dtype=torch.float64, device="cuda")
        Hb = torch.empty(total, dtype=torch.float64, device="cuda")
        fv.veval_with_derivatives(xd, r, J, Hb)
        rh, Hbh = r.cpu().numpy(), Hb.cpu().numpy()
        Jm = sp.csc_matrix((J.cpu().numpy(), inner, outer), shape=(m, n))
        Hf = (2.0 * (Jm.T @ Jm)).toarray()
        row0 = 0
        for term, (conn, valence, M) in enumerate(((F, 3, 8), (b, 1, 2))):
            off, k, n_res, _ = fv.residual_hessian_layout(term)
            assert k == 2 * valence and n_res == M * len(conn)
            table = fv.term_table(term, valence, len(conn))                         # slot -> handle, per element
            blocks = Hbh[off:off + n_res * k * k].reshape(len(conn), M, k, k)
            assert np.abs(blocks - blocks.transpose(0, 1, 3, 2)).max() == 0.0       # packed symmetric storage
            for e in range(len(conn)):
                gv = (2 * table[:, e][:, None] + np.arange(2)[None, :]).reshape(-1)  # local index -> global variable
                w = rh[row0 + M * e: row0 + M * (e + 1)]
                Hf[np.ix_(gv, gv)] += 2.0 * np.einsum("m,mij->ij", w, blocks[e])
            row0 += n_res
        f, g, Hs = fs.eval_with_derivatives_host(x)
        so, si = fs.pattern()
        Hsd = sp.csr_matrix((Hs, si, so), shape=(n, n)).toarray()
        assert np.abs(Hf - Hsd).max() <= 1e-11 * np.abs(Hsd).max()
        assert abs(float(rh @ rh) - f) <= 1e-12 * abs(f)
    finally:
        fv.close(); fs.close()
