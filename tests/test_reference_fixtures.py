"""CPU: the oracle against golden vectors produced by the REFERENCE ITSELF (tests/golden/reference_fixtures.npz, generated
in the build container by tests/golden/make_reference_fixtures.py from oracle/_ref = the unmodified TinyAD headers).
Unlike tests/test_oracle_vs_reference.py this needs neither /root/reference nor oracle/_ref, so it also runs on the GPU box."""
import numpy as np
import pytest

import oracle
from conftest import TOL_H, TOL_H_PROJ, assert_f, assert_vec
from problems import load_reference_fixtures

CASES = load_reference_fixtures()


def test_fixture_inventory():
    assert len([k for k in CASES if k.startswith("s/")]) >= 10 and len([k for k in CASES if k.startswith("v/")]) >= 2
    c = CASES["s/planar_newton/"]
    assert c["f"][2] == 24.5625 and len(c["inner"]) == 96            # tests/NewtonTest.cc:65: nnz == 4V + 8(V+F-1)
    assert CASES["s/dyn_sum_sqr/"]["f"][0] == 28.0                     # tests/DynamicElementsTest.cc:9-33


@pytest.mark.parametrize("name", sorted(k for k in CASES if k.startswith("s/")))
def test_oracle_scalar_functions_match_the_reference(name):
    c = CASES[name]
    terms = [oracle.Term(k, conn, data) for k, conn, data in c["terms"]]
    f0, f2, f3 = c["f"]
    assert_f(oracle.scalar_eval(c["d"], c["n_vertices"], terms, oracle.EVAL, c["x"]).f, f0)
    r = oracle.scalar_eval(c["d"], c["n_vertices"], terms, oracle.DERIVATIVES, c["x"])
    assert_f(r.f, f2)
    assert np.array_equal(r.outer, c["outer"]) and np.array_equal(r.inner, c["inner"])     # bit-exact pattern
    assert_vec(r.g, c["g"])
    assert_vec(r.values, c["H"], tol=TOL_H)
    rp = oracle.scalar_eval(c["d"], c["n_vertices"], terms, oracle.HESSIAN_PROJ, c["x"], eps=1e-9)
    assert_f(rp.f, f3)
    assert np.array_equal(rp.outer, c["outer"]) and np.array_equal(rp.inner, c["inner"])
    assert_vec(rp.values, c["H_proj"], tol=TOL_H_PROJ)


@pytest.mark.parametrize("name", sorted(k for k in CASES if k.startswith("v/")))
def test_oracle_vector_functions_match_the_reference(name):
    c = CASES[name]
    terms = [oracle.Term(k, conn, data) for k, conn, data in c["terms"]]
    r = oracle.vector_eval(2, c["n_vertices"], terms, oracle.V_SOS_DERIVATIVES, c["x"])
    assert np.array_equal(r.outer, c["outer"]) and np.array_equal(r.inner, c["inner"])
    assert_f(r.f, c["f"][0])
    assert_vec(r.r, c["r"])
    assert_vec(r.values, c["J"])
    assert_vec(r.g, c["g"])
