import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


# ---- parity metrics -------------------------------------------------------------------
TOL_F = 1e-12          # |df| / |f|
TOL_G = 1e-12          # ||dg||_inf / max(||g||_inf, scale)
TOL_H = 1e-12          # max|dH| / max|H|, unprojected
TOL_H_PROJ = 1e-10     # after projection (north_star)


def assert_f(f, f_ref, tol=TOL_F):
    if np.isinf(f_ref) or np.isnan(f_ref):
        assert f == f_ref or (np.isnan(f) and np.isnan(f_ref))
    else:
        assert abs(f - f_ref) <= tol * max(abs(f_ref), 1e-300), (f, f_ref)


def assert_vec(g, g_ref, tol=TOL_G, scale=0.0):
    g, g_ref = np.asarray(g), np.asarray(g_ref)
    assert g.shape == g_ref.shape
    denom = max(np.abs(g_ref).max(initial=0.0), scale, 1e-300)
    err = np.abs(g - g_ref).max(initial=0.0) / denom
    assert err <= tol, err


def entry_ratio(v, v_ref, scale):
    """SURVEY.md 8(c) per-entry metric: max over the entries of |v - v_ref| / scale_entry, where scale_entry is the sum over the
    contributing elements of |contribution| (oracle mode | ABS_SUM) or of the element's largest |entry| (| NORM_SUM: the rounding
    scale of one element's derivative block -- the right one after projection, whose error is relative to |H_e|, not to the entry).
    Entries whose scale is zero (explicit structural zeros, HessianProjection leaves them exact) must agree exactly."""
    v, v_ref, scale = np.asarray(v), np.asarray(v_ref), np.asarray(scale)
    assert v.shape == v_ref.shape == scale.shape
    delta = np.abs(v - v_ref)
    zero = scale == 0.0
    assert not delta[zero].any(), "an entry without contributions differs"
    return float((delta[~zero] / scale[~zero]).max(initial=0.0))


def assert_entries(v, v_ref, scale, tol):
    r = entry_ratio(v, v_ref, scale)
    assert r <= tol, f"per-entry |delta| / sum|contributions| = {r:.3e} > {tol:.1e}"
    return r


def assert_pattern(outer, inner, ref):
    assert outer.dtype == np.int32 and inner.dtype == np.int32
    assert np.array_equal(outer, ref.outer), "outer index array differs"
    assert np.array_equal(inner, ref.inner), "inner index array differs"
