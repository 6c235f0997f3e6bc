"""The reference's Scalar known-answer tests (tests/ScalarTest*.cc, ComplexTest.cc) on the PRODUCT's Scalar:
host build of the __host__ __device__ header (CPU test) and a one-thread CUDA kernel (GPU test)."""
import json
import os

import numpy as np
import pytest

import oracle
import tinyad_b200 as tad
from test_oracle_golden import CASES, check_case

COMPLEX = ["c_mul", "c_mul_d", "c_d_mul", "c_div", "c_div_d", "c_add", "c_sub", "c_sqr", "c_conj", "c_abs", "c_arg"]


def run_all(on_device):
    for c in CASES:
        check_case(c, tad.scalar_case(c["name"], c["params"], c["k"], on_device=on_device))
    # complex arithmetic and the k = 6 triangle: against the (pinned) oracle, 1e-12
    for name in COMPLEX:
        params = [0.7, -1.3, 0.4, 2.5]
        for (v, g, h), (vo, go, ho) in zip(tad.scalar_case(name, params, 2, on_device=on_device), oracle.scalar_case(name, params, 2)):
            assert abs(v - vo) <= 1e-12 and np.abs(g - go).max() <= 1e-12 and np.abs(h - ho).max() <= 1e-11, name
    params = [10.0, 1.0, 15.0, 3.0, 2.0, 2.0, 1, 1, 2, 1, 1, 2]     # ScalarTestHessianBlock.cc:50-90
    (v, g, h), = tad.scalar_case("symm_dirich6", params, 6, on_device=on_device)
    (vo, go, ho), = oracle.scalar_case("symm_dirich6", params, 6)
    assert abs(v - vo) <= 1e-12 * abs(vo) and np.abs(g - go).max() <= 1e-12 * np.abs(go).max()
    assert np.abs(h - ho).max() <= 1e-12 * np.abs(ho).max() and np.array_equal(h, h.T)


BLOCKS = [(0, 0, 0, 0), (0, 0, 6, 6), (0, 0, 3, 1), (0, 0, 1, 3), (2, 2, 2, 2), (1, 4, 5, 2)]   # ScalarTestHessianBlock.cc:93-100


def check_block(full, block, r0, c0, nr, nc, k):
    """A truncated-Hessian scalar (Scalar.hh:24-38 of the reference) carries the same value and gradient as the full one and
    exactly the entries of its block (read through the symmetric accessor: the block and its mirror), nothing else."""
    (vf, gf, hf), (vb, gb, hb) = full, block
    assert vb == vf and np.array_equal(gb, gf)
    inside = np.zeros((k, k), dtype=bool)
    inside[r0:r0 + nr, c0:c0 + nc] = True
    inside |= inside.T
    assert np.array_equal(hb[inside], hf[inside])           # TINYAD_ASSERT_EPS_MAT(H_block, H_full.block(...), 1e-16)
    assert not hb[~inside].any()


def run_hessian_blocks(on_device):
    full, block = tad.scalar_case("hess_block_issue13", [1.0, 2.0, 3.0, 4.0, 5.0], 5, on_device=on_device)   # :24-45
    check_block(full, block, 0, 2, 2, 3, 5)
    assert full[2][0, 2] == -10.0 and full[2][0, 4] == -4.0                                                      # d2f/dx1dy1 = -2 y3, d2f/dx1dy3 = 2 (x1 - y1)
    params = [10.0, 1.0, 15.0, 3.0, 2.0, 2.0, 1.0, 1.0, 2.0, 1.0, 1.0, 2.0]                                     # :76-77
    (vo, go, ho), = oracle.scalar_case("symm_dirich6", params, 6)
    for i, (r0, c0, nr, nc) in enumerate(BLOCKS):
        full, block = tad.scalar_case("hess_block_symdir", params + [float(i)], 6, on_device=on_device)
        check_block(full, block, r0, c0, nr, nc, 6)
        assert np.abs(full[2] - ho).max() <= 1e-12 * np.abs(ho).max()


def test_truncated_hessian_blocks_host_build():
    run_hessian_blocks(False)


@pytest.mark.gpu
def test_truncated_hessian_blocks_on_device(torch_cuda):
    run_hessian_blocks(True)


def test_scalar_cases_host_build():
    run_all(False)


@pytest.mark.gpu
def test_scalar_cases_on_device(torch_cuda):
    run_all(True)


@pytest.mark.gpu
def test_cpp_facade_semantics(torch_cuda):
    tad.selftest(0)
