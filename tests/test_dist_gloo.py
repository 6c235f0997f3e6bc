"""CPU, world_size 2 over gloo: the multi-rank path (element partition, vertex ownership, halo-row exchange of the
Hessian values, gradient all-reduce) against the single-process oracle.  Per-rank local assembly is done by the CPU
oracle here; on GPUs the same HaloPlan runs over NCCL with the CUDA assembly (tests/test_dist_gpu.py, bench.py)."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _worker(rank, world, port, kind, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from problems import grid_problem, tet_problem
        from tinyad_b200.dist import HaloPlan, slab_partition
        p, x = tet_problem(4, seed=7) if kind == "tet" else grid_problem(10, seed=7)
        k, conn, data = p.terms[0]
        d, nv = p.d, p.d * p.n_vertices
        ref = oracle.scalar_eval(d, p.n_vertices, p.oracle_terms(), oracle.HESSIAN_PROJ, x, n_threads=1)
        lo, hi = slab_partition(len(conn), world)[rank]
        loc = oracle.scalar_eval(d, p.n_vertices, [oracle.Term(k, conn[lo:hi], data[lo:hi])], oracle.HESSIAN_PROJ, x, n_threads=1)
        plan = HaloPlan(d, p.n_vertices, [conn[lo:hi]])
        # local CSR + structural zeros for the blocks other ranks will send (the CUDA runtime does this with
        # tad_function_add_pattern_blocks)
        A = sp.csr_matrix((loc.values, loc.inner, loc.outer), shape=(nv, nv))
        vi, vj = plan.extra_pattern_blocks()
        a = np.arange(d)
        er = ((d * vi)[:, None, None] + a[None, :, None] + 0 * a[None, None, :]).ravel()
        ec = ((d * vj)[:, None, None] + 0 * a[None, :, None] + a[None, None, :]).ravel()
        ones = sp.csr_matrix((np.ones(nv * 0 + len(er)), (er, ec)), shape=(nv, nv))
        pat = (abs(A) + sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=(nv, nv)) + ones).tocsr()
        pat.sort_indices()
        # values of A on the union pattern
        M = sp.csr_matrix((np.zeros(pat.nnz), pat.indices, pat.indptr), shape=(nv, nv))
        rows = np.repeat(np.arange(nv), np.diff(pat.indptr))
        gk = rows.astype(np.int64) * nv + pat.indices
        ar = np.repeat(np.arange(nv), np.diff(A.indptr))
        pos = np.searchsorted(gk, ar.astype(np.int64) * nv + A.indices)
        vals = np.zeros(pat.nnz)
        vals[pos] = A.data
        outer, inner = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
        plan.finalize(outer, inner)
        Hv = torch.from_numpy(vals)
        g = torch.from_numpy(loc.g.copy())
        f = torch.tensor([loc.f], dtype=torch.float64)
        plan.exchange(Hv, g)
        dist.all_reduce(f)
        # owned rows: pattern identical to the single-process rows, values equal
        R = sp.csr_matrix((ref.values, ref.inner, ref.outer), shape=(nv, nv))
        mask = plan.owned_row_mask()
        ok_pat = ok_val = True
        for r in np.nonzero(mask)[0]:
            s, e = outer[r], outer[r + 1]
            rs, re = ref.outer[r], ref.outer[r + 1]
            same = np.array_equal(inner[s:e], ref.inner[rs:re])
            ok_pat &= same
            if same:
                ok_val &= np.abs(Hv.numpy()[s:e] - ref.values[rs:re]).max(initial=0.0) <= 1e-12 * np.abs(ref.values).max()
        ok_f = f.item() == ref.f or abs(f.item() - ref.f) <= 1e-12 * abs(ref.f)   # an inverted element gives inf on both sides
        ok_g = np.abs(g.numpy() - ref.g).max() <= 1e-12 * np.abs(ref.g).max()
        ok = ok_pat and ok_val and ok_f and ok_g
        if not ok:
            print("rank", rank, "pattern", ok_pat, "values", ok_val, "f", ok_f, f.item(), ref.f, "g", ok_g, flush=True)
        owned = torch.from_numpy(mask.astype(np.int64))
        dist.all_reduce(owned)
        touched_rows = np.diff(ref.outer) > 0
        ok &= bool(np.all(owned.numpy()[touched_rows] == 1))      # every assembled row has exactly one owner
        ret[rank] = (bool(ok), int(mask.sum()), int(plan.halo_bytes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("tri", 2), ("tet", 2), ("tet", 3)])
def test_halo_exchange(kind, world):
    """world 2: one interface; world 3: the middle rank owns one interface and sends the other (two neighbours)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500) + (7 if kind == "tet" else 0) + 13 * world
    mp.spawn(_worker, args=(world, port, kind, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, n_owned, halo = ret[r]
        assert ok, f"rank {r}"
        assert n_owned > 0
    # every rank but the first sends its lower interface rows to the rank below (the lowest rank touching a vertex owns it)
    assert ret[0][2] == 0 and all(ret[r][2] > 0 for r in range(1, world))
