"""GPU: the CUDA path (through the C ABI) against the REFERENCE ITSELF.

(1) tests/golden/reference_fixtures.npz: outputs of the unmodified TinyAD headers (oracle/_ref, see
    tests/golden/make_reference_fixtures.py) for twelve small problems covering every term kind -- committed, so nothing of the
    reference is needed at run time;
(2) where the prebuilt oracle/_ref/libtinyad_ref.so travelled with the snapshot: the same comparison live at sizes the reference
    finishes in seconds (29 k tets, 20 k triangles), in both assembly modes.
Bars as everywhere: index arrays bit-exact, f / g / H 1e-12, projected H 1e-10."""
import numpy as np
import pytest

import oracle
import tinyad_b200 as tad
from conftest import TOL_H, TOL_H_PROJ, assert_f, assert_vec
from problems import Problem, grid_problem, load_reference_fixtures, tet_problem

pytestmark = pytest.mark.gpu

CASES = load_reference_fixtures()


# gather-mode assembly is exercised on the mesh energies (as in tests/test_parity_gpu.py); the atomic mode on every case
GATHER_CASES = ("s/planar_newton/", "s/tri_grid7/", "s/tet_cube3/", "s/tet_slab_strong/", "s/trig_mix/", "s/repeated_handle/")
SCALAR_RUNS = [(k, tad.ASSEMBLY_ATOMIC) for k in sorted(CASES) if k.startswith("s/")] + [(k, tad.ASSEMBLY_GATHER) for k in GATHER_CASES]


@pytest.mark.parametrize("name,assembly", SCALAR_RUNS)
def test_scalar_functions_match_the_reference(torch_cuda, name, assembly):
    c = CASES[name]
    f0, f2, f3 = c["f"]
    fn = Problem(c["d"], c["n_vertices"], c["terms"]).gpu(assembly=assembly)
    try:
        outer, inner = fn.pattern()
        assert np.array_equal(outer, c["outer"]) and np.array_equal(inner, c["inner"])         # bit-exact pattern
        assert_f(fn.eval_host(c["x"]), f0)
        f, g = fn.eval_with_gradient_host(c["x"])
        assert_f(f, f2)
        assert_vec(g, c["g"])
        f, g, H = fn.eval_with_derivatives_host(c["x"])
        assert_f(f, f2)
        assert_vec(g, c["g"])
        assert_vec(H, c["H"], tol=TOL_H)
        f, g, H = fn.eval_with_hessian_proj_host(c["x"], eps=1e-9)
        assert_f(f, f3)
        assert_vec(g, c["g"])
        assert_vec(H, c["H_proj"], tol=TOL_H_PROJ)
    finally:
        fn.close()


@pytest.mark.parametrize("name", sorted(k for k in CASES if k.startswith("v/")))
def test_vector_functions_match_the_reference(torch_cuda, name):
    torch = torch_cuda
    c = CASES[name]
    fn = Problem(2, c["n_vertices"], c["terms"], is_vector=True).gpu()
    try:
        outer, inner = fn.pattern()
        assert np.array_equal(outer, c["outer"]) and np.array_equal(inner, c["inner"])
        xd = torch.from_numpy(c["x"]).cuda()
        r = torch.empty(fn.n_outputs, dtype=torch.float64, device="cuda")
        J = torch.empty(len(inner), dtype=torch.float64, device="cuda")
        g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
        f = fn.veval_sum_of_squares_with_derivatives(xd, g, r, J)
        assert_f(f, c["f"][0])
        assert_vec(r.cpu().numpy(), c["r"])
        assert_vec(J.cpu().numpy(), c["J"])
        assert_vec(g.cpu().numpy(), c["g"])
    finally:
        fn.close()


VALENCE_OUTPUTS = {tad.SOS_SYMDIRICHLET2D: (3, 8), tad.SOS_PENALTY2D: (1, 2), tad.SOS_POLYCURL2D: (2, 2)}


@pytest.mark.parametrize("name", sorted(k for k in CASES if k.startswith("v/")))
def test_per_residual_hessians_match_the_reference(torch_cuda, name):
    """VectorFunction::eval_with_derivatives (Detail/VectorFunctionImpl.hh:207-235): the reference returns one n x n sparse Hessian
    per residual; the product returns the dense k x k block of every residual (tad_veval_with_derivatives).  Scattered through the
    term tables, the blocks must reproduce the reference's matrices entry by entry."""
    torch = torch_cuda
    c = CASES[name]
    fn = Problem(2, c["n_vertices"], c["terms"], is_vector=True).gpu()
    try:
        n, m = fn.n_vars, fn.n_outputs
        outer, inner = fn.pattern()
        total = fn.residual_hessian_layout(-1)[3]
        xd = torch.from_numpy(c["x"]).cuda()
        r = torch.empty(m, dtype=torch.float64, device="cuda")
        J = torch.empty(len(inner), dtype=torch.float64, device="cuda")
        Hb = torch.empty(total, dtype=torch.float64, device="cuda")
        fn.veval_with_derivatives(xd, r, J, Hb)
        assert_vec(r.cpu().numpy(), c["r"])
        assert_vec(J.cpu().numpy(), c["J"])
        Hbh = Hb.cpu().numpy()
        mine = np.zeros((m, n, n))
        row0 = 0
        for term, (kind, conn, _) in enumerate(c["terms"]):
            valence, M = VALENCE_OUTPUTS[kind]
            off, k, n_res, _ = fn.residual_hessian_layout(term)
            assert k == 2 * valence and n_res == M * len(conn)
            table = fn.term_table(term, valence, len(conn))
            blocks = Hbh[off:off + n_res * k * k].reshape(len(conn), M, k, k)
            for e in range(len(conn)):
                gv = (2 * table[:, e][:, None] + np.arange(2)[None, :]).reshape(-1)
                for q in range(M):
                    np.add.at(mine[row0 + M * e + q], np.ix_(gv, gv), blocks[e, q])
            row0 += n_res
        ref = np.zeros((m, n, n))
        np.add.at(ref, (c["hess_res"], c["hess_row"], c["hess_col"]), c["hess_val"])
        assert np.abs(ref).max() > 0.0
        assert np.abs(mine - ref).max() <= 1e-12 * np.abs(ref).max()
    finally:
        fn.close()


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libtinyad_ref.so did not travel with this snapshot")
@pytest.mark.parametrize("make", [lambda: tet_problem(17, seed=6, with_penalty=True), lambda: grid_problem(100, seed=2, with_penalty=True)])
def test_live_against_the_reference_library(torch_cuda, make):
    p, x = make()
    ot = p.oracle_terms()
    ref_h = oracle.ref_scalar_eval(p.d, p.n_vertices, ot, oracle.DERIVATIVES, x)
    ref_p = oracle.ref_scalar_eval(p.d, p.n_vertices, ot, oracle.HESSIAN_PROJ, x, eps=1e-9)
    for assembly in (tad.ASSEMBLY_ATOMIC, tad.ASSEMBLY_GATHER):
        fn = p.gpu(assembly=assembly)
        try:
            outer, inner = fn.pattern()
            assert np.array_equal(outer, ref_h.outer) and np.array_equal(inner, ref_h.inner)
            f, g, H = fn.eval_with_derivatives_host(x)
            assert_f(f, ref_h.f)
            assert_vec(g, ref_h.g)
            assert_vec(H, ref_h.values, tol=TOL_H)
            f, g, H = fn.eval_with_hessian_proj_host(x, eps=1e-9)
            assert_f(f, ref_p.f)
            assert_vec(H, ref_p.values, tol=TOL_H_PROJ)
        finally:
            fn.close()
