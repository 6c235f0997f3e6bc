"""GPU: add_elements_dynamic (ScalarFunction.hh:80-108, ScalarFunctionImpl.hh:134-214) against the oracle's
restatement of the same grouping and against the reference's own fixtures in tests/DynamicElementsTest.cc."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
import tinyad_b200 as tad
from conftest import assert_f, assert_vec, assert_pattern, TOL_H, TOL_H_PROJ
from problems import Problem, icosphere, one_ring_table

pytestmark = pytest.mark.gpu


def _run(torch, p, x, mode_project):
    fn = p.gpu()
    xd = torch.from_numpy(x).cuda()
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    f = fn.eval_with_derivatives(xd, g, H, project=mode_project)
    out = (f, g.cpu().numpy(), H.cpu().numpy(), fn.pattern(), fn.n_elements)
    fn.close()
    return out


def test_dynamic_elements_fixture(torch_cuda):
    """DynamicElementsTest.cc:9-33: add_elements_dynamic<3, 1> over range(4), element e accesses e variables, x = ones."""
    p = Problem(2, 4, [(tad.DYN_SUM_SQR2D, np.zeros((4, 1), dtype=np.int32), np.zeros((4, 1)))])
    x = np.ones(8)
    ref = oracle.scalar_eval(2, 4, p.oracle_terms(), oracle.HESSIAN_PROJ, x)
    f, g, H, (outer, inner), n_el = _run(torch_cuda, p, x, True)
    assert n_el == 4 and f == 28.0
    assert_pattern(outer, inner, ref)
    assert_vec(g, ref.g)
    assert_vec(H, ref.values, tol=TOL_H_PROJ)


@pytest.mark.parametrize("project", [False, True])
def test_dynamic_one_ring(torch_cuda, project):
    """DynamicElementsTest.cc:92-141: one-ring Dirichlet energy with vertex elements of run-time valence (two static groups).
    Pattern bit-exact vs the oracle (incl. the explicit zeros of padded elements), Hessian == Laplacian to 1e-12."""
    V, F = icosphere(2)
    tab = one_ring_table(len(V), F)
    p = Problem(1, len(V), [(tad.DYN_ONERING1D, tab, np.zeros(tab.shape))])
    x = np.random.default_rng(4).standard_normal(len(V))
    ref = oracle.scalar_eval(1, len(V), p.oracle_terms(), oracle.HESSIAN_PROJ if project else oracle.DERIVATIVES, x)
    f, g, H, (outer, inner), n_el = _run(torch_cuda, p, x, project)
    assert n_el == len(V)
    assert_pattern(outer, inner, ref)
    assert_f(f, ref.f)
    assert_vec(g, ref.g)
    assert_vec(H, ref.values, tol=TOL_H_PROJ if project else TOL_H)
    if not project:
        L = sp.lil_matrix((len(V), len(V)))
        for fc in F:
            for i in range(3):
                L[int(fc[i]), int(fc[i])] += 1.0
                L[int(fc[i]), int(fc[(i + 1) % 3])] -= 1.0
        Hm = sp.csc_matrix((H, inner, outer), shape=(len(V), len(V)))
        assert abs(Hm - L.tocsc()).max() < 1e-12


def test_dynamic_valence_too_large(torch_cuda):
    """ScalarFunctionImpl.hh:175-180: an element whose valence exceeds the largest static valence is an error (here: one-ring of
    11 handles against <4, 6, 7, 10>), and the failed call leaves nothing behind."""
    n = 12
    tab = -np.ones((n, 10), dtype=np.int32)
    tab[0, :] = np.arange(1, 11)          # vertex 0: itself + 10 neighbours = 11 handles
    fn = tad.Function(1, n)
    with pytest.raises(tad.TinyADError):
        fn.add_term(tad.DYN_ONERING1D, tab, np.zeros(tab.shape))
    assert fn.n_elements == 0
    fn.close()
