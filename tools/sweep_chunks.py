"""Dev tool: device-resident and host-buffer step time of eval_with_hessian_proj over slab sizes / lanes (C2 by default)."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tinyad_b200 as tad
from tinyad_b200 import meshes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
V, T = meshes.kuhn_cube(n)
data = meshes.tet_rest_data(V, T)
x = meshes.deform(V, 1.0 / n, seed=0).reshape(-1)
fn = tad.Function(3, len(V))
fn.add_term(tad.SYMDIRICHLET3D, T, data)
nnz = fn.nnz
xd = torch.from_numpy(x).cuda()
g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
H = torch.empty(nnz, dtype=torch.float64, device="cuda")
xh = torch.from_numpy(x.copy()).pin_memory()
gh = torch.empty(fn.n_vars, dtype=torch.float64).pin_memory()
Hh = torch.empty(nnz, dtype=torch.float64).pin_memory()
reps = 20 if n <= 60 else 8
configs = eval("[" + sys.argv[2] + "]") if len(sys.argv) > 2 else [(-1, 1), (524288, 1), (524288, 2), (262144, 1), (262144, 2), (262144, 3), (131072, 1), (131072, 2), (131072, 3), (65536, 2), (65536, 4)]
for chunk, lanes in configs:
    fn.set_option(tad.OPT_CHUNK_ELEMENTS, chunk)
    fn.set_option(tad.OPT_LANES, lanes)
    for _ in range(3):
        fn.eval_with_hessian_proj(xd, g, H)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn.eval_with_hessian_proj(xd, g, H)
    torch.cuda.synchronize()
    dev = (time.perf_counter() - t0) / reps
    for _ in range(2):
        fn.eval_with_hessian_proj_host(xh.numpy(), out_g=gh.numpy(), out_H=Hh.numpy())
    t0 = time.perf_counter()
    for _ in range(max(3, reps // 2)):
        fn.eval_with_hessian_proj_host(xh.numpy(), out_g=gh.numpy(), out_H=Hh.numpy())
    e2e = (time.perf_counter() - t0) / max(3, reps // 2)
    print(f"chunk {chunk:8d} lanes {lanes}: device {dev*1e3:7.3f} ms  e2e {e2e*1e3:7.3f} ms", flush=True)
