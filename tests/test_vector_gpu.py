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
