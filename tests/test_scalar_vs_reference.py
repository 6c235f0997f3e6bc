"""CPU: Scalar arithmetic against the reference's TinyAD::Scalar ITSELF (oracle/_ref, `ref_scalar_case` in oracle/ref_driver.cc).

Three implementations of the shared case vocabulary (every unary / binary operator of Scalar.hh, compound assignment, comparisons,
min / max / clamp, atan2 / hypot, multi-variable expressions, complex arithmetic, the k = 6 triangle):
  reference  TinyAD::Double<k> of /root/reference, compiled in place            (oracle.ref_scalar_case)
  oracle     oracle/tinyad_oracle.hh                                             (oracle.scalar_case)
  product    tinyad_b200/include/TinyAD/Scalar.hh, host build                    (tad.scalar_case(..., on_device=False))
(a) on the golden parameters of tests/golden/scalar_cases.json the reference must return the transcribed expected values -- which
    checks the transcription; (b) on randomly perturbed parameters (the reference's tests hold one point per operator) oracle and
    product must follow the reference: 1e-13 relative per returned scalar (value, gradient, Hessian separately)."""
import zlib

import numpy as np
import pytest

import oracle
import tinyad_b200 as tad
from test_oracle_golden import CASES, check_case

pytestmark = pytest.mark.skipif(not (oracle.ref_available() or oracle.build_ref()), reason="oracle/_ref is not built")

COMPLEX = ["c_mul", "c_mul_d", "c_d_mul", "c_div", "c_div_d", "c_add", "c_sub", "c_sqr", "c_conj", "c_abs", "c_arg"]
DISCRETE = {"pow_int": (3,), "cmp": tuple(range(16)), "isnan_isinf": tuple(range(16)), "clamp": tuple(range(16)), "clamp_d": tuple(range(16)),
            "min": tuple(range(16)), "max": tuple(range(16)), "fmin": tuple(range(16)), "fmax": tuple(range(16)), "fabs": (), "abs": ()}
TOL = 1e-13


def agree(name, got, want, tol=TOL):
    assert len(got) == len(want), name
    for (v, g, h), (vr, gr, hr) in zip(got, want):
        for a, b, what in ((np.array([v]), np.array([vr]), "val"), (g, gr, "grad"), (h, hr, "hess")):
            a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
            assert np.array_equal(np.isnan(a), np.isnan(b)), (name, what)
            m = ~np.isnan(b)
            scale = max(np.abs(b[m]).max(initial=0.0), 1.0)
            assert np.abs(a[m] - b[m]).max(initial=0.0) <= tol * scale, (name, what, a, b)


@pytest.mark.parametrize("c", CASES, ids=[f"{c['name']}-{i}" for i, c in enumerate(CASES)])
def test_reference_returns_the_transcribed_golden_values(c):
    check_case(c, oracle.ref_scalar_case(c["name"], c["params"], c["k"]))


def perturbed(c, rng):
    p = np.array(c["params"], dtype=float)
    keep = DISCRETE.get(c["name"], ())
    q = p * (1.0 + 0.03 * rng.uniform(-1.0, 1.0, size=p.shape)) + 1e-3 * rng.uniform(-1.0, 1.0, size=p.shape) * (p != 0)
    for i in keep:
        if i < len(q):
            q[i] = p[i]
    return list(q)


@pytest.mark.parametrize("name", sorted(set(c["name"] for c in CASES)))
def test_oracle_and_product_follow_the_reference_on_perturbed_parameters(name):
    rng = np.random.default_rng(zlib.crc32(name.encode()))      # a stable seed per operator
    for c in [c for c in CASES if c["name"] == name]:
        for _ in range(6):
            p = perturbed(c, rng)
            want = oracle.ref_scalar_case(name, p, c["k"])
            if not all(np.isfinite(v) and np.isfinite(g).all() and np.isfinite(h).all() for v, g, h in want):
                continue                                    # left the operator's domain (acos, atanh, log ... near their limits)
            agree(name, oracle.scalar_case(name, p, c["k"]), want)
            agree(name, tad.scalar_case(name, p, c["k"], on_device=False), want, tol=1e-12)


@pytest.mark.parametrize("name", COMPLEX + ["symm_dirich6"])
def test_complex_and_triangle_cases_follow_the_reference(name):
    rng = np.random.default_rng(len(name))
    for _ in range(8):
        if name == "symm_dirich6":
            p = list(np.array([10.0, 1.0, 15.0, 3.0, 2.0, 2.0, 1, 1, 2, 1, 1, 2]) + 0.1 * rng.uniform(-1, 1, 12))
            k = 6
        else:
            p = list(rng.uniform(-2.0, 2.0, 4) + np.array([0.0, 0.0, 3.0, 0.0]))     # |b| stays away from 0 for the divisions
            k = 2
        want = oracle.ref_scalar_case(name, p, k)
        agree(name, oracle.scalar_case(name, p, k), want)
        agree(name, tad.scalar_case(name, p, k, on_device=False), want, tol=1e-12)


@pytest.mark.gpu
def test_device_scalar_follows_the_reference_on_perturbed_parameters(torch_cuda):
    """The same comparison for the product's Scalar running in a CUDA kernel (one thread per case): 1e-12 relative."""
    for name in sorted(set(c["name"] for c in CASES)):
        rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
        for c in [c for c in CASES if c["name"] == name]:
            for _ in range(3):
                p = perturbed(c, rng)
                want = oracle.ref_scalar_case(name, p, c["k"])
                if not all(np.isfinite(v) and np.isfinite(g).all() and np.isfinite(h).all() for v, g, h in want):
                    continue
                agree(name, tad.scalar_case(name, p, c["k"], on_device=True), want, tol=1e-12)
    rng = np.random.default_rng(5)
    for name in COMPLEX:
        for _ in range(4):
            p = list(rng.uniform(-2.0, 2.0, 4) + np.array([0.0, 0.0, 3.0, 0.0]))
            agree(name, tad.scalar_case(name, p, 2, on_device=True), oracle.ref_scalar_case(name, p, 2), tol=1e-12)
