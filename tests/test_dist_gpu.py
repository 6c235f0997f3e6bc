"""GPU, 2 ranks over NCCL (skipped with fewer than 2 devices): CUDA assembly per rank + halo-row exchange vs the oracle."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import oracle
        import tinyad_b200 as tad
        from problems import tet_problem
        from tinyad_b200.dist import HaloPlan, slab_partition
        p, x = tet_problem(6, seed=3)
        k, conn, data = p.terms[0]
        ref = oracle.scalar_eval(3, p.n_vertices, p.oracle_terms(), oracle.HESSIAN_PROJ, x)
        lo, hi = slab_partition(len(conn), world)[rank]
        fn = tad.Function(3, p.n_vertices, device=rank)
        fn.add_term(k, conn[lo:hi], data[lo:hi])
        plan = HaloPlan(3, p.n_vertices, [conn[lo:hi]])
        fn.add_pattern_blocks(*plan.extra_pattern_blocks())
        outer, inner = fn.pattern()
        plan.finalize(outer, inner, device="cuda")
        x_dev = torch.from_numpy(x).cuda()
        g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
        H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
        f = torch.tensor([fn.eval_with_hessian_proj(x_dev, g, H)], dtype=torch.float64, device="cuda")
        plan.exchange(H, g)
        dist.all_reduce(f)
        Hh, gh = H.cpu().numpy(), g.cpu().numpy()
        ok = True
        for r in np.nonzero(plan.owned_row_mask())[0]:
            s, e = outer[r], outer[r + 1]
            rs, re = ref.outer[r], ref.outer[r + 1]
            ok &= np.array_equal(inner[s:e], ref.inner[rs:re])
            ok &= np.abs(Hh[s:e] - ref.values[rs:re]).max(initial=0.0) <= 1e-10 * np.abs(ref.values).max()
        ok &= abs(f.item() - ref.f) <= 1e-12 * abs(ref.f)
        ok &= np.abs(gh - ref.g).max() <= 1e-12 * np.abs(ref.g).max()
        ret[rank] = bool(ok)
        fn.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_assembly(torch_cuda):
    torch = torch_cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29733, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
