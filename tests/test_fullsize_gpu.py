"""GPU, at BASELINE.json's full sizes.

Config 2 (Kuhn cube n = 55, 998,250 tets, Double<12>), config 1 (512 x 512 grid, 524,288 triangles, Double<6>) and a >= 200 k-element
slice of config 4 (VectorFunction residual Jacobians) are compared with the ORACLE IN FULL -- every CSR row: index arrays bit-exact,
f / g / unprojected H <= 1e-12, projected H <= 1e-10, in the global-max metric and in SURVEY.md 8(c)'s per-entry metric
(|delta| <= tol * sum over the contributing elements; conftest.entry_ratio).  The oracle evaluates config 2 in ~10 s on 16 cores.
On top of that, size-independent properties:
  * closed-form pattern size nnz = 9 (V + 2E), ascending inner indices, structural symmetry;
  * translation invariance of the energy: sum_v g_v = 0, H t = 0 for the three translations (unprojected);
  * H = H^T (1e-12 unprojected, 1e-10 projected), z^T H_proj z >= 0;
  * the deterministic gather assembly is bit-reproducible and agrees with the atomic assembly to 1e-12;
  * the CSR rows of the z = 0 lattice plane only receive contributions of the first cube layer: they must equal the ORACLE's rows
    for that layer alone (pattern of those rows bit-exact, values 1e-12 / 1e-10, same for g)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
import tinyad_b200 as tad
from conftest import TOL_H, TOL_H_PROJ, assert_entries, assert_f, assert_pattern, assert_vec
from tinyad_b200 import meshes

pytestmark = pytest.mark.gpu
N = 55


@pytest.fixture(scope="module")
def c2(torch_cuda):
    torch = torch_cuda
    V, T = meshes.kuhn_cube(N)
    data = meshes.tet_rest_data(V, T)
    x = meshes.deform(V, 1.0 / N, seed=0).reshape(-1)
    out = {"V": V, "T": T, "data": data, "x": x}
    xd = torch.from_numpy(x).cuda()
    for name, assembly in (("atomic", tad.ASSEMBLY_ATOMIC), ("gather", tad.ASSEMBLY_GATHER)):
        fn = tad.Function(3, len(V), assembly=assembly)
        fn.add_term(tad.SYMDIRICHLET3D, T, data)
        g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
        H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
        f = fn.eval_with_derivatives(xd, g, H, project=False)
        out[name] = {"f": f, "g": g.cpu().numpy(), "H": H.cpu().numpy()}
        fp = fn.eval_with_hessian_proj(xd, g, H)
        out[name + "_proj"] = {"f": fp, "g": g.cpu().numpy(), "H": H.cpu().numpy()}
        if name == "gather":
            fn.eval_with_hessian_proj(xd, g, H)
            out["gather_proj_again"] = H.cpu().numpy()
        out["pattern"] = fn.pattern()
        out["stats"] = fn.projection_stats()
        fn.close()
    return out


def _csr(c2, vals):
    outer, inner = c2["pattern"]
    n = len(outer) - 1
    return sp.csr_matrix((vals, inner, outer), shape=(n, n))


def test_pattern_closed_form(c2):
    outer, inner = c2["pattern"]
    n = N
    nv = (n + 1) ** 3
    ne = 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3          # axis + face-diagonal + body-diagonal edges (App. C)
    assert len(inner) == 9 * (nv + 2 * ne) == 23036814
    assert outer[0] == 0 and outer[-1] == len(inner) and np.all(np.diff(outer) > 0)
    rows = np.repeat(np.arange(len(outer) - 1), np.diff(outer))
    asc = (np.diff(inner) > 0) | (np.diff(rows) > 0)
    assert asc.all()                                                    # ascending inner indices inside every row
    P = sp.csr_matrix((np.ones(len(inner), dtype=np.int8), inner, outer))
    assert (P != P.T).nnz == 0                                          # structurally symmetric: CSR == CSC


def test_translation_invariance_and_symmetry(c2):
    r = c2["atomic"]
    g = r["g"].reshape(-1, 3)
    assert np.abs(g.sum(axis=0)).max() <= 1e-10 * np.abs(g).sum()
    H = _csr(c2, r["H"])
    hmax = np.abs(r["H"]).max()
    for a in range(3):
        t = np.zeros(H.shape[0]); t[a::3] = 1.0
        assert np.abs(H @ t).max() <= 1e-11 * hmax * 100
    assert abs(H - H.T).max() <= 1e-12 * hmax
    Hp = _csr(c2, c2["atomic_proj"]["H"])
    assert abs(Hp - Hp.T).max() <= 1e-10 * np.abs(c2["atomic_proj"]["H"]).max()
    rng = np.random.default_rng(0)
    for _ in range(4):
        z = rng.standard_normal(H.shape[0])
        assert z @ (Hp @ z) > 0.0                                       # every element block is PSD after projection
    assert c2["stats"]["decomposed"] == len(c2["T"])
    assert c2["atomic_proj"]["f"] == c2["atomic"]["f"] and np.isfinite(c2["atomic"]["f"])


def test_assembly_modes_agree_and_gather_is_deterministic(c2):
    for a, b, tol in (("atomic", "gather", 1e-12), ("atomic_proj", "gather_proj", 1e-10)):
        assert abs(c2[a]["f"] - c2[b]["f"]) <= 1e-12 * abs(c2[a]["f"])
        assert np.abs(c2[a]["g"] - c2[b]["g"]).max() <= 1e-12 * np.abs(c2[a]["g"]).max()
        assert np.abs(c2[a]["H"] - c2[b]["H"]).max() <= tol * np.abs(c2[a]["H"]).max()
    assert np.array_equal(c2["gather_proj"]["H"], c2["gather_proj_again"])   # bitwise run-to-run (NewtonTest.cc:97-111)


def test_bottom_plane_rows_equal_oracle_on_first_layer(c2):
    n = N
    layer = 6 * n * n                                                   # tets of the first cube layer (cell-major order)
    nv_plane = (n + 1) ** 2                                             # vertices with z = 0 come first
    T0, d0 = c2["T"][:layer], c2["data"][:layer]
    term = [oracle.Term(oracle.SYMDIRICHLET3D, T0, d0)]
    outer, inner = c2["pattern"]
    rows = 3 * nv_plane
    for mode, key, tol in ((oracle.DERIVATIVES, "atomic", 1e-12), (oracle.HESSIAN_PROJ, "atomic_proj", 1e-10)):
        ref = oracle.scalar_eval(3, len(c2["V"]), term, mode, c2["x"])
        # same rows: identical column sets and values
        assert np.array_equal(np.diff(outer[:rows + 1]), np.diff(ref.outer[:rows + 1]))
        assert np.array_equal(inner[:outer[rows]], ref.inner[:ref.outer[rows]])
        Hg = c2[key]["H"][:outer[rows]]
        Ho = ref.values[:ref.outer[rows]]
        assert np.abs(Hg - Ho).max() <= tol * np.abs(Ho).max()
        assert np.abs(c2[key]["g"][:rows] - ref.g[:rows]).max() <= 1e-12 * np.abs(ref.g[:rows]).max()


# ---- every row against the oracle -------------------------------------------------------------------------------------------
class _Ref:
    def __init__(self, outer, inner):
        self.outer, self.inner = outer, inner


def _full_parity(torch, d, kind, okind, V, conn, data, x, label):
    """All modes of one full-size scalar problem vs the oracle: pattern bit-exact, global-max and per-entry metrics."""
    terms = [oracle.Term(okind, conn, data)]
    ref_h = oracle.scalar_eval(d, len(V), terms, oracle.DERIVATIVES, x)
    abs_h = oracle.scalar_eval(d, len(V), terms, oracle.DERIVATIVES | oracle.ABS_SUM, x)
    nrm_h = oracle.scalar_eval(d, len(V), terms, oracle.DERIVATIVES | oracle.NORM_SUM, x)
    ref_p = oracle.scalar_eval(d, len(V), terms, oracle.HESSIAN_PROJ, x)
    nrm_p = oracle.scalar_eval(d, len(V), terms, oracle.HESSIAN_PROJ | oracle.NORM_SUM, x)
    fn = tad.Function(d, len(V))
    fn.add_term(kind, conn, data)
    try:
        outer, inner = fn.pattern()
        assert_pattern(outer, inner, ref_h)                                  # every index, bit-exact
        assert_pattern(outer, inner, ref_p)
        xd = torch.from_numpy(x).cuda()
        g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
        H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
        assert_f(fn.eval(xd), ref_h.f)
        assert_f(fn.eval_with_gradient(xd, g), ref_h.f)
        assert_vec(g.cpu().numpy(), ref_h.g)
        f = fn.eval_with_derivatives(xd, g, H, project=False)
        gh, Hh = g.cpu().numpy(), H.cpu().numpy()
        assert_f(f, ref_h.f)
        assert_vec(gh, ref_h.g)
        assert_vec(Hh, ref_h.values, tol=TOL_H)
        r_g = assert_entries(gh, ref_h.g, nrm_h.g, 1e-12)                    # per entry, element-norm scale
        r_H = assert_entries(Hh, ref_h.values, nrm_h.values, 1e-12)
        r_Ha = assert_entries(Hh, ref_h.values, abs_h.values, 1e-11)         # per entry, sum |contributions| (cancellation inside an
        r_ga = assert_entries(gh, ref_h.g, abs_h.g, 1e-11)                   # element is not bounded by its own entry: one digit looser)
        f = fn.eval_with_hessian_proj(xd, g, H)
        gp, Hp = g.cpu().numpy(), H.cpu().numpy()
        assert_f(f, ref_p.f)
        assert_vec(gp, ref_p.g)
        assert_vec(Hp, ref_p.values, tol=TOL_H_PROJ)
        r_Hp = assert_entries(Hp, ref_p.values, nrm_p.values, 1e-10)
        st = fn.projection_stats()
        assert st["decomposed"] == ref_p.phases["n_decomposed"] and st["rebuilt"] == ref_p.phases["n_rebuilt"]
        print(f"\n[{label}] per-entry ratios: g {r_g:.2e} (abs-sum scale {r_ga:.2e}), H {r_H:.2e} (abs-sum scale {r_Ha:.2e}), "
              f"H_proj {r_Hp:.2e}; rebuilt {st['rebuilt']} of {len(conn)}")
    finally:
        fn.close()


def test_c2_every_row_vs_oracle(torch_cuda, c2):
    """Config 2 in full (998,250 tets, 23,036,814 CSR entries)."""
    _full_parity(torch_cuda, 3, tad.SYMDIRICHLET3D, oracle.SYMDIRICHLET3D, c2["V"], c2["T"], c2["data"], c2["x"], "C2 n=55")


def test_c1_every_row_vs_oracle(torch_cuda):
    """Config 1 in full (512 x 512 grid, 524,288 triangles, Double<6>); nnz = 4V + 8(V + F - 1) (tests/NewtonTest.cc:65)."""
    V, F = meshes.grid_2d(512)
    x = meshes.deform(V, 1.0 / 512, seed=0).reshape(-1)
    _full_parity(torch_cuda, 2, tad.SYMDIRICHLET2D, oracle.SYMDIRICHLET2D, V, F, meshes.tri_rest_data(V, F), x, "C1 N=512")
    fn = tad.Function(2, len(V))
    fn.add_term(tad.SYMDIRICHLET2D, F, meshes.tri_rest_data(V, F))
    assert fn.nnz == 4 * len(V) + 8 * (len(V) + len(F) - 1) == 7352324
    fn.close()


def test_c4_slice_every_entry_vs_oracle(torch_cuda):
    """Config 4 (VectorFunction, complex-arithmetic residuals) at N = 300: 270,600 edge elements, 541,200 residuals."""
    torch = torch_cuda
    from test_vector_gpu import polycurl_problem
    p, x = polycurl_problem(300)
    assert len(p.terms[0][1]) >= 200000
    ot = p.oracle_terms()
    ref = oracle.vector_eval(2, p.n_vertices, ot, oracle.V_SOS_DERIVATIVES, x)
    fn = p.gpu()
    try:
        outer, inner = fn.pattern()
        assert np.array_equal(outer, ref.outer) and np.array_equal(inner, ref.inner)      # CSC pattern bit-exact
        xd = torch.from_numpy(x).cuda()
        r = torch.empty(fn.n_outputs, dtype=torch.float64, device="cuda")
        J = torch.empty(len(inner), dtype=torch.float64, device="cuda")
        g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
        f = fn.veval_sum_of_squares_with_derivatives(xd, g, r, J)
        assert_f(f, ref.f)
        assert_vec(r.cpu().numpy(), ref.r)
        assert_vec(J.cpu().numpy(), ref.values)
        assert_vec(g.cpu().numpy(), ref.g)
        # per entry: a residual / Jacobian entry is one element's output, so its scale is that element's largest |entry|
        Jh, rows = J.cpu().numpy(), inner
        row_scale = np.zeros(fn.n_outputs)
        np.maximum.at(row_scale, rows, np.abs(ref.values))
        el_scale = np.maximum(row_scale[0::2], row_scale[1::2]).repeat(2)                   # 2 residuals per element
        assert_entries(Jh, ref.values, el_scale[rows], 1e-12)
    finally:
        fn.close()
