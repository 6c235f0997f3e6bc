"""CPU: the reference's OWN unit tests, run here.

oracle/_ref/reference_tests is /root/reference/tests/*.cc (all 22 files, unmodified, compiled where they lie by
`make -C oracle _ref/reference_tests`) over oracle/eigen_shim (stand-in for the Eigen API) and oracle/gtest_shim (stand-in
for the googletest macros).  It is what makes oracle/_ref a legitimate checker for the oracle: with the shim underneath, the
reference still satisfies the known answers its authors wrote down for Scalar.hh (unary / binary operators, comparison,
constructors, Hessian blocks, complex numbers, custom derivatives), Element, ScalarFunction, VectorFunction, dynamic
elements, the handle types, exceptions, OpenMP evaluation, the closed-form SVD, Newton / Gauss-Newton / solver switching.

All 404 tests of the reference tree must pass (oracle.REF_TESTS_SKIPPED is empty).  The binary reads nothing from /root/reference
at run time.
"""
import re

import pytest

import oracle

pytestmark = pytest.mark.skipif(oracle.build_ref_tests() is None, reason="oracle/_ref/reference_tests is not built")


def test_reference_suite_passes_over_the_shim():
    rc, out = oracle.run_ref_tests()
    m = re.search(r"(\d+) tests ran, (\d+) passed, (\d+) failed", out)
    assert m, out[-2000:]
    ran, passed, failed = map(int, m.groups())
    assert rc == 0 and failed == 0 and passed == ran, out[-4000:]
    assert ran >= 404                                       # 404 TEST()s in the reference tree
    assert out.count("[ SKIPPED  ]") == len(oracle.REF_TESTS_SKIPPED)
    for suite in ("ScalarTestUnaryOperators", "ScalarTestBinaryOperators", "ScalarTestHessianBlock", "ScalarFunctionTest", "VectorFunctionTest",
                  "DynamicElementsTest", "GaussNewtonTest", "NewtonTest", "SVDTest", "HandleTypeTest", "ExceptionTest", "ComplexTest"):
        assert re.search(r"\[       OK \] %s\." % suite, out), suite
