"""Operations/SVD.hh of the reference (2 x 2 closed-form SVD, closest orthogonal matrix) on the product's Scalar<4>: host build of
the __host__ __device__ header (CPU) and a one-thread CUDA kernel (GPU).  Checks of tests/SVDTest.cc with its tolerances
(1e-12 value, 1e-8 gradient, 1e-4 Hessian), against numpy's SVD and analytic / finite-difference derivatives instead of
Eigen::JacobiSVD on active scalars (Eigen is not available)."""
import numpy as np
import pytest

import tinyad_b200 as tad


def _polar(A):
    U, _, Vt = np.linalg.svd(A)
    return U @ Vt


def _check(on_device):
    rng = np.random.default_rng(11)
    for _ in range(8):
        A = rng.uniform(-1.0, 1.0, (2, 2))           # Eigen::Matrix2::Random is uniform in [-1, 1]
        p = A.reshape(-1)
        res = tad.scalar_case("svd2", p, 4, on_device=on_device)
        # U * diag(S) * V^T recomposes A: values, identity Jacobian, zero Hessian (SVDTest.cc:30-46)
        for idx, (v, g, h) in enumerate(res[:4]):
            assert abs(v - p[idx]) < 1e-12
            assert np.abs(g - np.eye(4)[idx]).max() < 1e-8
            assert np.abs(h).max() < 1e-4
        # singular values and their analytic gradients dS_k/dA = u_k v_k^T
        U, S, Vt = np.linalg.svd(A)
        for k in range(2):
            v, g, h = res[4 + k]
            assert abs(v - S[k]) < 1e-12
            assert np.abs(g - np.outer(U[:, k], Vt[k, :]).reshape(-1)).max() < 1e-8
            assert np.abs(h - h.T).max() < 1e-10
        # closest orthogonal matrix (SVDTest.cc:62-88): value vs numpy, gradient vs central differences of numpy's polar factor,
        # Hessian vs central differences of the product's own gradients
        res = tad.scalar_case("closest_orthogonal2", p, 4, on_device=on_device)
        R = _polar(A).reshape(-1)
        hstep = 1e-6
        for idx, (v, g, h) in enumerate(res):
            assert abs(v - R[idx]) < 1e-12
            fd = np.zeros(4)
            for j in range(4):
                e = np.zeros(4); e[j] = hstep
                fd[j] = (_polar((p + e).reshape(2, 2)).reshape(-1)[idx] - _polar((p - e).reshape(2, 2)).reshape(-1)[idx]) / (2 * hstep)
            assert np.abs(g - fd).max() < 1e-7
            Hfd = np.zeros((4, 4))
            for j in range(4):
                e = np.zeros(4); e[j] = 1e-5
                gp = tad.scalar_case("closest_orthogonal2", p + e, 4, on_device=on_device)[idx][1]
                gm = tad.scalar_case("closest_orthogonal2", p - e, 4, on_device=on_device)[idx][1]
                Hfd[:, j] = (gp - gm) / 2e-5
            assert np.abs(h - Hfd).max() < 1e-4 * max(1.0, np.abs(Hfd).max())
            assert np.abs(h - h.T).max() < 1e-10 * max(1.0, np.abs(h).max())


def test_svd_host_build():
    _check(False)


@pytest.mark.gpu
def test_svd_on_device(torch_cuda):
    _check(True)
