"""Multi-GPU assembly: element partition + halo-row exchange.

One process per GPU.  Elements are partitioned (contiguous slabs); vertex numbering stays global.  Every rank
evaluates its own elements into a local CSR (rows of all vertices it touches).  A vertex is OWNED by the lowest
rank that touches it; after the exchange every rank holds the complete rows (pattern = union, values = sum) of
the vertices it owns -- a row-distributed CSR, which is what a distributed solver wants.  Exchange per evaluation:
  * halo-row Hessian values: each rank sends the values of the blocks (vi, vj) whose row vertex it does not own
    to the owner (point-to-point, one message per neighbour) and the owner adds them into its slots;
  * gradient: all-reduce;  f: all-reduce (done by the caller).
The plan is backend agnostic: it works on torch tensors with torch.distributed, so the same code runs over NCCL on
GPUs and over gloo on CPU (tests/test_dist_gloo.py runs it at world_size 2 against the oracle).
"""
import numpy as np
import torch
import torch.distributed as dist


def slab_partition(n_elements, world):
    """Contiguous element ranges [lo, hi) per rank."""
    return [((n_elements * r) // world, (n_elements * (r + 1)) // world) for r in range(world)]


class HaloPlan:
    def __init__(self, d, n_vertices, conns, rank=None, world=None, group=None):
        """conns: list of (n_el, valence) int arrays with GLOBAL vertex handles of this rank's elements."""
        self.d, self.n_vertices = d, n_vertices
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.group = group
        # vertex-pair keys of the local pattern
        keys = []
        touched = np.zeros(n_vertices, dtype=bool)
        for c in conns:
            c = np.asarray(c, dtype=np.int64)
            c = c.reshape(len(c), -1)
            touched[c.ravel()] = True
            N = c.shape[1]
            keys.append((c[:, :, None] * n_vertices + c[:, None, :]).reshape(-1) if N else np.zeros(0, np.int64))
        self.local_keys = np.unique(np.concatenate(keys)) if keys else np.zeros(0, np.int64)
        # owner = lowest rank touching the vertex
        own = torch.full((n_vertices,), self.world, dtype=torch.int32)
        own[torch.from_numpy(touched)] = self.rank
        own = self._to_comm(own)
        dist.all_reduce(own, op=dist.ReduceOp.MIN, group=group)
        self.owner = own.cpu().numpy()
        self.owned = self.owner == self.rank
        # halo blocks: row vertex owned elsewhere -> destination rank
        vi = self.local_keys // n_vertices
        dest = self.owner[vi]
        halo = dest != self.rank
        self.send_keys = {int(p): self.local_keys[halo & (dest == p)] for p in np.unique(dest[halo])}
        gathered = [None] * self.world
        dist.all_gather_object(gathered, self.send_keys, group=group)
        self.recv_keys = {src: g[self.rank] for src, g in enumerate(gathered) if src != self.rank and self.rank in g and len(g[self.rank])}
        self.send_idx = self.recv_idx = None

    def _to_comm(self, t):
        return t.cuda() if dist.get_backend(self.group) == "nccl" else t

    def extra_pattern_blocks(self):
        """(vi, vj) of the blocks other ranks will send here; inject them into the local pattern before building it."""
        if not self.recv_keys:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        k = np.unique(np.concatenate(list(self.recv_keys.values())))
        return k // self.n_vertices, k % self.n_vertices

    def finalize(self, outer, inner, device="cpu"):
        """Value positions of the exchanged entries in the local CSR (outer, inner: numpy int32 arrays)."""
        d, nv = self.d, self.d * self.n_vertices
        rows = np.repeat(np.arange(nv, dtype=np.int64), np.diff(outer).astype(np.int64))
        gkeys = rows * nv + inner.astype(np.int64)          # strictly increasing over the CSR arrays

        def positions(block_keys):
            vi, vj = block_keys // self.n_vertices, block_keys % self.n_vertices
            a = np.arange(d, dtype=np.int64)
            r = (d * vi)[:, None, None] + a[None, :, None]
            c = (d * vj)[:, None, None] + a[None, None, :]
            q = (r * nv + c).reshape(-1)
            pos = np.searchsorted(gkeys, q)
            assert np.all(pos < len(gkeys)) and np.array_equal(gkeys[pos], q), "halo block missing from the local pattern"
            return torch.from_numpy(pos).to(device)

        self.send_idx = {p: positions(k) for p, k in self.send_keys.items()}
        self.recv_idx = {p: positions(k) for p, k in self.recv_keys.items()}
        self.halo_bytes = 8 * sum(len(v) for v in self.send_idx.values())
        return self

    def exchange(self, H_values, g=None):
        """Add the halo-row values received from the other ranks into H_values (in place); all-reduce g."""
        ops, bufs = [], []
        for p, idx in self.send_idx.items():
            ops.append(dist.P2POp(dist.isend, H_values.index_select(0, idx).contiguous(), p, group=self.group))
        for p, idx in self.recv_idx.items():
            buf = torch.empty(len(idx), dtype=H_values.dtype, device=H_values.device)
            bufs.append((idx, buf))
            ops.append(dist.P2POp(dist.irecv, buf, p, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for idx, buf in bufs:
            H_values.index_add_(0, idx, buf)
        if g is not None:
            dist.all_reduce(g, group=self.group)
        return H_values

    def owned_row_mask(self):
        """Boolean mask over the n_vars rows this rank owns."""
        return np.repeat(self.owned, self.d)
