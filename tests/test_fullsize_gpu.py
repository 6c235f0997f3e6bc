"""GPU, at BASELINE.json's full single-GPU size (config 2: Kuhn cube n = 55, 998,250 tets, Double<12>): the oracle cannot
evaluate a million tets in seconds, so parity at this size is checked through size-independent properties and through an exact
sub-problem:
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
