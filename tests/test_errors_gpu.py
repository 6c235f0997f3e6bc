"""GPU: error convention of the C ABI (SURVEY.md 8(b) "Error convention"): status codes + message, and the function object stays
fully usable after an error (tests/ExceptionTest.cc:26-53)."""
import ctypes

import numpy as np
import pytest

import tinyad_b200 as tad

pytestmark = pytest.mark.gpu


def test_nonfinite_derivative_is_reported_and_object_stays_usable(torch_cuda):
    """Non-finite element gradient / Hessian -> error (ScalarObjectiveTerm.hh:210,252-253); val = inf alone is not (App. E 5, 6)."""
    torch = torch_cuda
    fn = tad.Function(1, 1)
    fn.add_term(tad.SQRT1D, np.array([[0]], dtype=np.int32), np.array([[1.0]]))
    g = torch.empty(1, dtype=torch.float64, device="cuda")
    H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    bad = torch.tensor([-1.0], dtype=torch.float64, device="cuda")      # sqrt(-1): NaN value and derivatives
    zero = torch.tensor([0.0], dtype=torch.float64, device="cuda")       # sqrt(0): finite value, infinite derivative
    for x in (bad, zero):
        with pytest.raises(tad.TinyADError) as e:
            fn.eval_with_gradient(x, g)
        assert e.value.status == 2                                        # TAD_NONFINITE_DERIVATIVE
        with pytest.raises(tad.TinyADError) as e:
            fn.eval_with_hessian_proj(x, g, H)
        assert e.value.status == 2
    assert fn.eval(zero) == 0.0                                           # the passive evaluation has no derivative to check
    good = torch.tensor([4.0], dtype=torch.float64, device="cuda")
    assert fn.eval_with_gradient(good, g) == 2.0 and g.item() == 0.25    # usable after the errors
    f = fn.eval_with_derivatives(good, g, H)
    assert f == 2.0 and abs(H.item() + 0.25 / 8.0) < 1e-15               # d2/dx2 sqrt(x) = -1/(4 x^1.5)
    fn.close()


def test_handle_out_of_range_is_rejected_at_add_elements(torch_cuda):
    """Variable handle outside [0, n_handles) (Element.hh:159-170): reported by the recording pass, nothing is added."""
    fn = tad.Function(1, 4)
    with pytest.raises(tad.TinyADError) as e:
        fn.add_term(tad.EDGE_DIRICHLET1D, np.array([[0, 1], [2, 7]], dtype=np.int32), np.ones((2, 1)))
    assert "out of range" in str(e.value)
    assert fn.n_elements == 0
    fn.add_term(tad.EDGE_DIRICHLET1D, np.array([[0, 1], [2, 3]], dtype=np.int32), np.ones((2, 1)))
    assert fn.n_elements == 2 and fn.nnz == 8
    fn.close()


def test_invalid_arguments(torch_cuda):
    torch = torch_cuda
    R = tad.runtime()
    fn = tad.Function(1, 2)
    fn.add_term(tad.EDGE_DIRICHLET1D, np.array([[0, 1]], dtype=np.int32), np.ones((1, 1)))
    f = ctypes.c_double()
    assert R.tad_eval(fn.h, None, ctypes.byref(f)) == 1                  # TAD_INVALID_ARGUMENT: x is null
    assert b"null" in R.tad_last_error()
    x = torch.zeros(2, dtype=torch.float64, device="cuda")
    r = torch.zeros(2, dtype=torch.float64, device="cuda")
    assert R.tad_veval(fn.h, x.data_ptr(), r.data_ptr()) == 1            # vector evaluation of a scalar function
    assert R.tad_gauss_newton_direction(fn.h, r.data_ptr(), r.data_ptr(), 0.0, 1e-10, 10, x.data_ptr(), None, None) == 1
    assert fn.eval(x) == 0.0                                              # still usable
    h = ctypes.c_void_p()
    assert R.tad_function_create(0, 4, 0, 0, ctypes.byref(h)) == 1       # variable dimension must be >= 1
    fn.close()


def test_functor_that_changes_its_handles_is_reported(torch_cuda):
    """the sparsity pattern is recorded once (the reference rediscovers it at every evaluation,
    Element.hh:208-260).  A functor whose variables() calls depend on x must be reported (TAD_PATTERN_MISMATCH), not silently
    assembled into the wrong rows; evaluations that stay on the recorded path work, and the object stays usable."""
    torch = torch_cuda
    fn = tad.Function(1, 3)
    # element 0: handles (0, 1), element 1: handles (1, 2); the functor returns before touching its second handle when x_a > 0.5
    fn.add_term(tad.BRANCH_ON_X1D, np.array([[0, 1], [1, 2]], dtype=np.int32), np.array([[0.5], [0.5]]))
    g = torch.empty(3, dtype=torch.float64, device="cuda")
    H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    ok = torch.tensor([0.2, 0.1, 0.4], dtype=torch.float64, device="cuda")
    f = fn.eval_with_derivatives(ok, g, H)
    assert abs(f - (0.1 ** 2 + 0.3 ** 2)) < 1e-15
    bad = torch.tensor([0.9, 0.1, 0.4], dtype=torch.float64, device="cuda")     # element 0 leaves early
    for call in (lambda: fn.eval(bad), lambda: fn.eval_with_gradient(bad, g), lambda: fn.eval_with_hessian_proj(bad, g, H)):
        with pytest.raises(tad.TinyADError) as e:
            call()
        assert e.value.status == 9                                              # TAD_PATTERN_MISMATCH
    assert abs(fn.eval(ok) - f) < 1e-15                                         # still usable
    fn.close()
