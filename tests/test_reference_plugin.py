"""The drop-in seen from the reference's side.

oracle/_ref/libtinyad_plugin.so (oracle/ref_plugin_driver.cc) is reference code: the UNMODIFIED TinyAD::ScalarFunction of
/root/reference, compiled in place, with include/reference_binding/B200ObjectiveTerm.hh -- a ScalarObjectiveTermBase, the reference's
own plugin interface for the path (Detail/ScalarObjectiveTerm.hh:21-45) -- pushed onto its `objective_terms`.  Calling the
reference's eval / eval_with_gradient / eval_with_derivatives / eval_with_hessian_proj then runs the element functors on the B200
through the C ABI, while x handling, term summation, setFromTriplets and the return types stay the reference's.

CPU: the library builds against the reference and the product, loads, and without a device the reference's facade reports the
product's loud failure (no CPU fallback).  GPU: the results equal those of the same ScalarFunction holding the reference's own CPU
terms (oracle/_ref/libtinyad_ref.so): pattern bit-exact, f / g / H 1e-12, projected H 1e-10."""
import numpy as np
import pytest

import oracle
from conftest import TOL_H, TOL_H_PROJ, assert_f, assert_vec, has_gpu
from problems import grid_problem, planar_newton_problem, tet_problem

pytestmark = pytest.mark.skipif(oracle.build_plugin() is None or not oracle.ref_available(), reason="oracle/_ref plugin library is not built")


def test_plugin_library_loads_and_fails_loudly_without_a_device():
    L = oracle.plugin_lib()
    assert hasattr(L, "plugin_scalar_eval")
    if has_gpu():
        return
    p, x = planar_newton_problem()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        oracle.plugin_scalar_eval(p.d, p.n_vertices, p.oracle_terms(), oracle.HESSIAN_PROJ, x)


@pytest.mark.gpu
@pytest.mark.parametrize("make", [planar_newton_problem, lambda: grid_problem(20, seed=4, with_penalty=True), lambda: tet_problem(6, seed=3, with_penalty=True)])
def test_reference_facade_with_the_b200_term_equals_the_reference(torch_cuda, make):
    p, x = make()
    ot = p.oracle_terms()
    for assembly in (0, 1):                                         # FP64 atomics / deterministic gather
        for mode in (oracle.EVAL, oracle.GRADIENT, oracle.DERIVATIVES, oracle.HESSIAN_PROJ):
            want = oracle.ref_scalar_eval(p.d, p.n_vertices, ot, mode, x, eps=1e-9)
            got = oracle.plugin_scalar_eval(p.d, p.n_vertices, ot, mode, x, eps=1e-9, assembly=assembly)
            assert_f(got.f, want.f)
            if mode >= oracle.GRADIENT:
                assert_vec(got.g, want.g)
            if mode >= oracle.DERIVATIVES:
                assert np.array_equal(got.outer, want.outer) and np.array_equal(got.inner, want.inner)     # Eigen::SparseMatrix pattern
                assert_vec(got.values, want.values, tol=TOL_H if mode == oracle.DERIVATIVES else TOL_H_PROJ)
