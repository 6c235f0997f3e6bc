"""GPU: batched PSD projection kernel (tad_project_batch) vs the oracle / numpy for every instantiated k, and for sizes
without a dedicated instantiation (run-time k Jacobi kernel: any k <= 32, e.g. d = 2 with N = 7, hexahedra k = 24)."""
import numpy as np
import pytest

import oracle
import tinyad_b200 as tad

pytestmark = pytest.mark.gpu


def seq_rc(k):
    """tile order of Detail/HessLayout.hh (python restatement for the test)."""
    t = 3 if k % 3 == 0 else (2 if k % 2 == 0 else 1)
    out = []
    for bi in range(k // t):
        for bj in range(bi + 1):
            for r in range(t):
                for c in range(t):
                    if bi == bj and c > r:
                        continue
                    out.append((bi * t + r, bj * t + c))
    return out


def make_batch(k, n, rng):
    mats = []
    for i in range(n):
        A = rng.standard_normal((k, k))
        A = A + A.T
        kind = i % 5
        if kind == 1:
            A = A + 50.0 * k * np.eye(k)                 # diagonally dominant -> early-out 1
        elif kind == 2:
            A = A @ A.T + 0.5 * np.eye(k)                # PD, not dominant -> early-out 2
        elif kind == 3:
            B = rng.standard_normal((k, max(1, k - 3)))
            A = B @ B.T                                  # PSD with a null space (like an element Hessian)
        elif kind == 4:
            A = np.zeros((k, k))                         # val = inf element: zero Hessian -> eps * I
        mats.append(A)
    return np.array(mats)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 18, 11, 14, 24, 32])
@pytest.mark.parametrize("eps", [1e-9, -1.0])
@pytest.mark.parametrize("method", [0, 1])
def test_project_batch(torch_cuda, k, eps, method):
    torch = torch_cuda
    rng = np.random.default_rng(100 + k)
    n = 257
    mats = make_batch(k, n, rng)
    rc = seq_rc(k)
    assert len(rc) == k * (k + 1) // 2
    stride = ((n + 31) // 32) * 32
    packed = np.zeros((len(rc), stride))
    for s, (r, c) in enumerate(rc):
        packed[s, :n] = mats[:, r, c]
    dev = torch.from_numpy(packed).cuda()
    counts = torch.zeros(4, dtype=torch.int64, device="cuda")
    tad.project_batch(k, dev, n, stride, eps=eps, method=method, counts_dev=counts)
    torch.cuda.synchronize()
    out = dev.cpu().numpy()
    n_rebuilt = 0
    for i in range(n):
        ref, code = oracle.project(mats[i], eps)
        got = np.zeros((k, k))
        for s, (r, c) in enumerate(rc):
            got[r, c] = got[c, r] = out[s, i]
        scale = max(np.abs(mats[i]).max(), abs(eps), 1e-300)
        assert np.abs(got - ref).max() <= 2e-11 * scale, (k, i, code)   # an order below the 1e-10 parity bar
        if code < 2 and i % 5 in (1, 2):
            assert np.array_equal(got, mats[i])           # both early-outs leave H bit-unchanged
        n_rebuilt += code == 2
    c = counts.cpu().numpy()
    assert c[1] <= c[0] <= n
    assert c[2] <= n // 20                                  # the fast path rarely needs the full solver
    if eps > 0:
        assert abs(int(c[1]) - n_rebuilt) <= 2             # borderline eigenvalues may fall on either side of eps
