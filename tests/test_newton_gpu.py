"""GPU: projected-Newton utilities around the hot path through the C ABI.

newton_direction (Utils/NewtonDirection.hh:25-48), newton_decrement (Utils/NewtonDecrement.hh:20-26) and
line_search (Utils/LineSearch.hh:14-65) against scipy's sparse direct solve and a numpy restatement of the
reference loop; the reference's own end-to-end fixture tests/NewtonTest.cc:60-90 (f -> 4, g -> 0)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle
import tinyad_b200 as tad
from problems import planar_newton_problem, tet_problem, grid_problem

pytestmark = pytest.mark.gpu


def _csr(fn, H_dev):
    outer, inner = fn.pattern()
    return sp.csr_matrix((H_dev.cpu().numpy(), inner, outer), shape=(fn.n_vars, fn.n_vars))


def _line_search_ref(terms, d, nv, x0, dx, f0, g, s_max=1.0, shrink=0.8, max_iters=64, armijo=1e-4):
    """Utils/LineSearch.hh:26-65 on the CPU oracle."""
    try_one = s_max > 1.0
    s = s_max
    for i in range(max_iters):
        f1 = oracle.scalar_eval(d, nv, terms, oracle.EVAL, x0 + s * dx).f
        if f1 <= f0 + armijo * s * dx.dot(g):
            return x0 + s * dx, f1, s, i + 1
        if try_one and s > 1.0 and s * shrink < 1.0:
            s = 1.0
        else:
            s *= shrink
    return x0, f0, 0.0, max_iters


def test_newton_fixture_on_device(torch_cuda):
    """tests/NewtonTest.cc:60-90: 10 iterations of eval_with_hessian_proj -> newton_direction(w_identity=1e-9) -> line_search."""
    torch = torch_cuda
    p, x = planar_newton_problem()
    fn = p.gpu()
    xd = torch.from_numpy(x).cuda()
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    d = torch.empty_like(g)
    xn = torch.empty_like(g)
    assert fn.nnz == 4 * 6 + 8 * (6 + 4 - 1)
    for _ in range(10):
        f = fn.eval_with_hessian_proj(xd, g, H)
        its, rel = fn.newton_direction(g, H, d, w_identity=1e-9, rel_tol=1e-13)
        assert rel <= 1e-13
        dec = fn.newton_decrement(d, g)
        assert dec >= 0.0 and abs(dec + 0.5 * float(d.dot(g))) <= 1e-14 * max(1.0, abs(dec))
        f_new, step, n = fn.line_search(xd, d, f, g, xn)
        assert f_new <= f
        xd, xn = xn, xd
    f = fn.eval_with_hessian_proj(xd, g, H)
    assert abs(f - 4.0) < 1e-12 and float(g.abs().max()) < 1e-10      # NewtonTest.cc:82-88
    assert abs(fn.eval(xd) - f) < 1e-15
    fn.close()


@pytest.mark.parametrize("make", [lambda: grid_problem(12, seed=1, with_penalty=True), lambda: tet_problem(5, seed=2, with_penalty=True)])
def test_newton_direction_matches_direct_solve(torch_cuda, make):
    torch = torch_cuda
    p, x = make()
    fn = p.gpu()
    xd = torch.from_numpy(x).cuda()
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    d = torch.empty_like(g)
    fn.eval_with_hessian_proj(xd, g, H)
    w = 1e-6
    its, rel = fn.newton_direction(g, H, d, w_identity=w, rel_tol=1e-13, max_iters=20000)
    A = _csr(fn, H).tocsc() + w * sp.identity(fn.n_vars, format="csc")
    ref = spla.spsolve(A, -g.cpu().numpy())
    err = np.abs(d.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 1e-7, (err, its, rel)
    # residual of the device solution, measured independently
    res = np.linalg.norm(A @ d.cpu().numpy() + g.cpu().numpy()) / np.linalg.norm(g.cpu().numpy())
    assert res < 1e-11
    assert abs(fn.newton_decrement(d, g) + 0.5 * ref.dot(g.cpu().numpy())) <= 1e-7 * abs(ref.dot(g.cpu().numpy()))
    fn.close()


def test_newton_direction_rejects_indefinite(torch_cuda):
    """'Linear solve failed.' (NewtonDirection.hh:43-44): an unprojected, indefinite Hessian must not return a direction silently."""
    torch = torch_cuda
    p, x = tet_problem(4, seed=3)
    fn = p.gpu()
    xd = torch.from_numpy(x).cuda()
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    d = torch.empty_like(g)
    fn.eval_with_derivatives(xd, g, H)      # not projected
    H.mul_(-1.0)                            # negative definite on most of the space
    with pytest.raises(tad.TinyADError) as e:
        fn.newton_direction(g, H, d, w_identity=0.0, max_iters=200)
    assert e.value.status == 8
    # the function object stays usable (tests/ExceptionTest.cc:26-53)
    fn.eval_with_hessian_proj(xd, g, H)
    fn.newton_direction(g, H, d, w_identity=1e-6)
    fn.close()


@pytest.mark.parametrize("s_max", [1.0, 3.0])
def test_line_search_matches_reference_loop(torch_cuda, s_max):
    torch = torch_cuda
    p, x = tet_problem(4, seed=5, with_penalty=True)
    fn = p.gpu()
    terms = p.oracle_terms()
    xd = torch.from_numpy(x).cuda()
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    d = torch.empty_like(g)
    xn = torch.empty_like(g)
    f = fn.eval_with_hessian_proj(xd, g, H)
    fn.newton_direction(g, H, d, w_identity=1e-8, rel_tol=1e-12)
    d.mul_(4.0)                              # overshoot so that the search really backtracks (and inverts tets: f = inf trials)
    f_new, step, n = fn.line_search(xd, d, f, g, xn, s_max=s_max)
    x_ref, f_ref, s_ref, n_ref = _line_search_ref(terms, 3, p.n_vertices, x, d.cpu().numpy(), f, g.cpu().numpy(), s_max=s_max)
    assert n == n_ref and step == pytest.approx(s_ref, rel=1e-15)
    assert n > 1
    assert abs(f_new - f_ref) <= 1e-12 * abs(f_ref)
    assert np.abs(xn.cpu().numpy() - x_ref).max() <= 1e-15 * np.abs(x_ref).max()
    # no descent possible along +g: returns x0 like the reference (LineSearch.hh:62-64)
    f_new, step, n = fn.line_search(xd, g, f, g, xn, max_iters=6)
    assert step == 0.0 and n == 6 and f_new == f and torch.equal(xn, xd)
    with pytest.raises(tad.TinyADError):
        fn.line_search(xd, d, f, g, xn, s_max=0.0)
    fn.close()


def _gn_buffers(torch, fn):
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    r = torch.empty(fn.n_outputs, dtype=torch.float64, device="cuda")
    J = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    return g, r, J


def test_gauss_newton_direction_matches_normal_equations(torch_cuda):
    """gauss_newton_direction (Utils/GaussNewtonDirection.hh:24-47): d = -(J^T J + w I)^-1 J^T r, against scipy on the same J."""
    torch = torch_cuda
    from test_vector_gpu import sos_problem
    p, x = sos_problem(12)
    fn = p.gpu()
    xd = torch.from_numpy(x).cuda()
    g, r, J = _gn_buffers(torch, fn)
    d = torch.empty_like(g)
    fn.veval_sum_of_squares_with_derivatives(xd, g, r, J)
    w = 1e-8
    its, rel = fn.gauss_newton_direction(r, J, d, w_identity=w, rel_tol=1e-13, max_iters=20000)
    outer, inner = fn.pattern()
    Jm = sp.csc_matrix((J.cpu().numpy(), inner, outer), shape=(fn.n_outputs, fn.n_vars))
    A = (Jm.T @ Jm + w * sp.identity(fn.n_vars)).tocsc()
    ref = spla.spsolve(A, -(Jm.T @ r.cpu().numpy()))
    err = np.abs(d.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 1e-6, (err, its, rel)
    # g = 2 J^T r  =>  the Gauss-Newton direction is a descent direction with decrement -0.5 d.g = d.(J^T J + w) d
    assert fn.newton_decrement(d, g) > 0.0
    fn.close()


def test_gauss_newton_fixture_on_device(torch_cuda):
    """tests/GaussNewtonTest.cc:140-160: Gauss-Newton iterations (w_identity = 1e-12) + line search on eval_sum_of_squares
    reach the distortion minimum f = 4 (within 0.1) with a small gradient."""
    torch = torch_cuda
    from test_vector_gpu import sos_problem
    p, x = sos_problem()
    fn = p.gpu()
    xd = torch.from_numpy(x).cuda()
    g, r, J = _gn_buffers(torch, fn)
    d = torch.empty_like(g)
    xn = torch.empty_like(g)
    for _ in range(20):                                              # GaussNewtonTest.cc:113
        f = fn.veval_sum_of_squares_with_derivatives(xd, g, r, J)
        fn.gauss_newton_direction(r, J, d, w_identity=1e-12, rel_tol=1e-12)
        f_new, step, n = fn.line_search(xd, d, f, g, xn)
        assert f_new <= f
        xd, xn = xn, xd
    f = fn.veval_sum_of_squares_with_derivatives(xd, g, r, J)
    assert abs(f - 4.0) < 0.1 and float(g.abs().max()) < 0.1           # :152-159
    fn.close()
