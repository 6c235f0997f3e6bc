"""GPU, 2 ranks over NCCL (skipped with fewer than 2 devices): the partitioned evaluation INSIDE the runtime
(tad_function_set_comm: vertex ownership, halo blocks, overlapped halo exchange, f all-reduce) against the single-process oracle.
bench.py carries the same comparison in the `check` of every N > 1 line (the driver's test box has one GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

pytestmark = pytest.mark.gpu


def _check_owned(fn, rank, ref, H, g, tol_H):
    """Rows owned by `rank`: same columns as the oracle's rows (bit-exact), values / g entries within tolerance."""
    outer, inner = fn.pattern()
    d = fn.d
    owned = np.repeat(fn.vertex_owner() == rank, d)
    ok = True
    hmax, gmax = np.abs(ref.values).max(), np.abs(ref.g).max()
    for r in np.nonzero(owned)[0]:
        s, e = outer[r], outer[r + 1]
        rs, re = ref.outer[r], ref.outer[r + 1]
        ok &= np.array_equal(inner[s:e], ref.inner[rs:re])
        if H is not None:
            ok &= np.abs(H[s:e] - ref.values[rs:re]).max(initial=0.0) <= tol_H * hmax
    ok &= np.abs(g[owned] - ref.g[owned]).max() <= 1e-12 * gmax
    return bool(ok), owned


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
        from tinyad_b200.dist import slab_partition
        comm = tad.Comm.from_torch_distributed(rank)
        p, x = tet_problem(8, seed=3, with_penalty=True)
        ref_p = oracle.scalar_eval(3, p.n_vertices, p.oracle_terms(), oracle.HESSIAN_PROJ, x)
        ref_h = oracle.scalar_eval(3, p.n_vertices, p.oracle_terms(), oracle.DERIVATIVES, x)
        ok = True
        for assembly in (tad.ASSEMBLY_ATOMIC, tad.ASSEMBLY_GATHER):
            for chunk in (1024, -1):                                   # several slabs per rank (halo slabs first) / one slab
                fn = tad.Function(3, p.n_vertices, device=rank, assembly=assembly)
                for k, conn, data in p.terms:                          # every rank adds the same sequence of terms, with ITS elements
                    lo, hi = slab_partition(len(conn), world)[rank]
                    fn.add_term(k, conn[lo:hi], data[lo:hi])
                fn.set_option(tad.OPT_CHUNK_ELEMENTS, chunk)
                fn.set_comm(comm)
                x_dev = torch.from_numpy(x).cuda()
                g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
                H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
                # value only, gradient, derivatives, projected derivatives: f is global, owned g entries / rows are complete
                ok &= abs(fn.eval(x_dev) - ref_h.f) <= 1e-12 * abs(ref_h.f)
                f = fn.eval_with_gradient(x_dev, g)
                ok &= abs(f - ref_h.f) <= 1e-12 * abs(ref_h.f)
                ok &= _check_owned(fn, rank, ref_h, None, g.cpu().numpy(), 0)[0]
                f = fn.eval_with_derivatives(x_dev, g, H, project=False)
                ok &= abs(f - ref_h.f) <= 1e-12 * abs(ref_h.f)
                ok &= _check_owned(fn, rank, ref_h, H.cpu().numpy(), g.cpu().numpy(), 1e-12)[0]
                f = fn.eval_with_hessian_proj(x_dev, g, H)
                ok &= abs(f - ref_p.f) <= 1e-12 * abs(ref_p.f)
                good, owned = _check_owned(fn, rank, ref_p, H.cpu().numpy(), g.cpu().numpy(), 1e-10)
                ok &= good
                # host-buffer entry point (pipelined D2H must not copy rows that still receive halo values)
                f2, g2, H2 = fn.eval_with_hessian_proj_host(x)
                ok &= abs(f2 - ref_p.f) <= 1e-12 * abs(ref_p.f)
                ok &= _check_owned(fn, rank, ref_p, H2, g2, 1e-10)[0]
                # every vertex is owned by exactly one rank; replicated gradient on request
                cnt = torch.from_numpy(owned.astype(np.int32)).cuda()
                dist.all_reduce(cnt)
                ok &= bool((cnt == 1).all().item())
                fn.set_option(tad.OPT_REPLICATE_GRADIENT, 1)
                fn.eval_with_gradient(x_dev, g)
                ok &= np.abs(g.cpu().numpy() - ref_h.g).max() <= 1e-12 * np.abs(ref_h.g).max()
                ok &= fn.halo_bytes() > 0 if rank == 1 else True       # rank 1 sends its bottom-plane rows to rank 0
                fn.close()
        ret[rank] = bool(ok)
        comm.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_partitioned_evaluation(torch_cuda):
    torch = torch_cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29733, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_single_rank_communicator_is_a_no_op(torch_cuda):
    """world = 1: the partitioned code path (ownership, schedule, all-reduce of f) with nothing to exchange gives the plain result."""
    torch = torch_cuda
    import tinyad_b200 as tad
    from problems import tet_problem
    p, x = tet_problem(5, seed=1)
    comm = tad.Comm(tad.Comm.unique_id(), 0, 1, 0)
    fa, fb = p.gpu(), p.gpu()
    fb.set_comm(comm)
    xd = torch.from_numpy(x).cuda()
    ga, gb = (torch.empty(fa.n_vars, dtype=torch.float64, device="cuda") for _ in range(2))
    Ha, Hb = (torch.empty(fa.nnz, dtype=torch.float64, device="cuda") for _ in range(2))
    assert fa.nnz == fb.nnz and all(np.array_equal(a, b) for a, b in zip(fa.pattern(), fb.pattern()))
    va, vb = fa.eval_with_hessian_proj(xd, ga, Ha), fb.eval_with_hessian_proj(xd, gb, Hb)
    assert va == vb
    assert np.abs(ga.cpu().numpy() - gb.cpu().numpy()).max() <= 1e-13 * np.abs(ga.cpu().numpy()).max()
    assert np.abs(Ha.cpu().numpy() - Hb.cpu().numpy()).max() <= 1e-13 * np.abs(Ha.cpu().numpy()).max()
    assert (fb.vertex_owner() == 0).all() and fb.halo_bytes() == 0
    fa.close(); fb.close(); comm.close()
