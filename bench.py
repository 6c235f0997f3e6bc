#!/usr/bin/env python
"""bench.py -- projected-Hessian assembly throughput of the tet Double<12> path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            # this framework, one B200
  torchrun ... bench.py --gpus N ...                        # N ranks, weak scaling over z-slabs
  python bench.py --impl reference ...                      # the reference's CPU/OpenMP algorithm (oracle port)

A "step" is one eval_with_hessian_proj over the whole mesh: x is resident in HBM, f / g / CSR values are
left in HBM (`value`); `e2e` times the same call through the host-buffer C ABI (pinned host x, g, H values;
H2D + D2H inside the timed region).  Pattern / scatter-map construction is one-time setup and reported apart.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "proj-Hessian assembled elements/s (tet Double<12>)"
UNIT = "elements/s"
# algorithmic work per tet (packed-symmetric flop count; compulsory HBM bytes)
FLOPS_AD_TET, FLOPS_PROJ_TET, BYTES_TET = 41823.0, 17712.0, 353.0
FLOPS_AD_TRI, FLOPS_PROJ_TRI, BYTES_TRI = 3943.0, 2268.0, 216.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c5", "c1", "small", "c3", "c3small", "c4", "c4small"],
                    help="c2: Kuhn cube n=55 (998,250 tets) per GPU [default]; c5: n=119 (10.1M tets) split over the GPUs; "
                         "c1: 512^2 triangle grid; small: n=16 smoke size; c3: projected-Newton loop on the C2 mesh, timed per phase "
                         "(--steps = Newton iterations, default 20); c4: VectorFunction residual Jacobians + Gauss-Newton on a 2M-face grid")
    ap.add_argument("--newton-tol", type=float, default=1e-6, help="c3: relative residual of the PCG solve (inexact Newton)")
    ap.add_argument("--assembly", default="atomic", choices=["atomic", "gather"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-n", type=int, default=22, help="cube edge of the CPU-baseline sample (22 -> 63,888 tets)")
    return ap.parse_args()


def workload_mesh(name, rank, world):
    """Returns (d, kind, V, conn, data, x, description).  For world > 1 each rank gets a z-slab."""
    import tinyad_b200 as tad
    from tinyad_b200 import meshes
    if name == "c1":
        V, F = meshes.grid_2d(512)
        return 2, tad.SYMDIRICHLET2D, V, F, meshes.tri_rest_data(V, F), meshes.deform(V, 1.0 / 512, seed=0), "C1: 512x512 grid, 524,288 triangles, Double<6>"
    n = {"c2": 55, "c5": 119, "small": 16}[name]
    if name == "c5":                      # strong: n=119 cube, z-layers split over the ranks
        lo = (119 * rank) // world
        hi = (119 * (rank + 1)) // world
        nz_total, z0, nz = 119, lo, hi - lo
        desc = f"C5: Kuhn cube n=119 (10,110,954 tets) split in {world} z-slab(s), Double<12>"
    else:                                 # weak: each rank owns an n x n x n block of an n x n x (n*world) lattice
        nz_total, z0, nz = n * world, n * rank, n
        desc = f"{name.upper()}: Kuhn cube n={n} ({6 * n ** 3:,} tets) per GPU, Double<12>"
    V, T = meshes.kuhn_cube(n, n, nz, z0=z0, nz_total=nz_total)
    x = meshes.deform(V, 1.0 / n, seed=0)
    return 3, tad.SYMDIRICHLET3D, V, T, meshes.tet_rest_data(V, T), x, desc


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  Started early (nvidia-smi needs a while
    to come up); stop(t0, t1) keeps the samples taken while the GPU was under load in [t0, t1] (wall clock)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= t1 + 0.15)]
        if not rows and self.rows:
            rows = [self.rows[-1][1]]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(sample_n, threads):
    """Times the CPU oracle (port of the reference's OpenMP eval_with_hessian_proj) on an n^3 Kuhn cube sample."""
    import oracle
    import tinyad_b200 as tad
    from tinyad_b200 import meshes
    V, T = meshes.kuhn_cube(sample_n)
    data = meshes.tet_rest_data(V, T)
    x = meshes.deform(V, 1.0 / sample_n, seed=0).reshape(-1)
    terms = [oracle.Term(oracle.SYMDIRICHLET3D, T, data)]
    oracle.scalar_eval(3, len(V), terms, oracle.HESSIAN_PROJ, x[: 3 * len(V)], n_threads=threads)  # warm-up
    times, phases = [], None
    for _ in range(3):
        t0 = time.perf_counter()
        r = oracle.scalar_eval(3, len(V), terms, oracle.HESSIAN_PROJ, x, n_threads=threads)
        times.append(time.perf_counter() - t0)
        phases = r.phases
    t = float(np.median(times))
    return {"value": len(T) / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"Kuhn cube n={sample_n} ({len(T):,} tets), same energy / eps, median of 3 after 1 warm-up; oracle = C++ restatement "
                      f"(reference needs Eigen, not installed)",
            "phases_s": {k: phases[k] for k in ("eval_s", "accumulate_s", "compress_s")}, "seconds": t}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, all host threads), same metric and config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = oracle.max_threads()
    n = args.cpu_sample_n
    oracle.lib()
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_baseline_step(n, threads) if i else cpu_baseline_step(n, threads)
        if i >= args.warmup:
            times.append(res[1])
    n_el = res[0]
    t = float(np.mean(times))
    value = n_el / t
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"bounded sample of C2: Kuhn cube n={n} ({n_el:,} tets) per step, eval_with_hessian_proj, eps=1e-9"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"n={n} cube per step; oracle port of the reference's OpenMP path (reference itself needs Eigen, absent here)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


_cpu_cache = {}


def cpu_baseline_step(n, threads):
    import oracle
    from tinyad_b200 import meshes
    if n not in _cpu_cache:
        V, T = meshes.kuhn_cube(n)
        _cpu_cache[n] = (len(V), [oracle.Term(oracle.SYMDIRICHLET3D, T, meshes.tet_rest_data(V, T))],
                         meshes.deform(V, 1.0 / n, seed=0).reshape(-1), len(T))
    nv, terms, x, nt = _cpu_cache[n]
    t0 = time.perf_counter()
    oracle.scalar_eval(3, nv, terms, oracle.HESSIAN_PROJ, x, n_threads=threads)
    return nt, time.perf_counter() - t0


def run_newton(args):
    """BASELINE.json configs[2]: projected-Newton loop (assembly + linear solve + line search) on the 1M-tet deformation,
    timed per phase.  The solve is the PCG stand-in of tad_newton_direction (no cuDSS in this image) and is reported apart
    from the assembly.  Loop shape: tests/NewtonTest.cc:68-75 of the reference."""
    import torch
    import tinyad_b200 as tad
    from tinyad_b200 import meshes
    n = 55 if args.workload == "c3" else 12
    V, T = meshes.kuhn_cube(n)
    x0 = meshes.deform(V, 1.0 / n, seed=0).reshape(-1)
    pins = np.array([[0], [n], [len(V) - 1], [len(V) - 1 - n]], dtype=np.int32)
    fn = tad.Function(3, len(V))
    fn.add_term(tad.SYMDIRICHLET3D, T, meshes.tet_rest_data(V, T))
    fn.add_term(tad.PENALTY3D, pins, V[pins[:, 0]])
    nnz = fn.nnz
    x = torch.from_numpy(x0.copy()).cuda()
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H = torch.empty(nnz, dtype=torch.float64, device="cuda")
    d = torch.empty_like(g)
    xn = torch.empty_like(g)
    iters = args.steps
    for _ in range(max(1, args.warmup)):   # warm-up (same x)
        fn.eval_with_hessian_proj(x, g, H)
    torch.cuda.synchronize()
    t_asm = t_sol = t_ls = 0.0
    log = []
    for it in range(iters):
        t0 = time.perf_counter()
        f = fn.eval_with_hessian_proj(x, g, H)
        t1 = time.perf_counter()
        cg_iters, rel = fn.newton_direction(g, H, d, w_identity=1e-9, rel_tol=args.newton_tol, max_iters=20000)
        dec = fn.newton_decrement(d, g)
        t2 = time.perf_counter()
        f_new, step, n_evals = fn.line_search(x, d, f, g, xn)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        t_asm += t1 - t0; t_sol += t2 - t1; t_ls += t3 - t2
        log.append({"f": f, "decrement": dec, "cg_iters": cg_iters, "step": step, "ls_evals": n_evals})
        x, xn = xn, x
    total = t_asm + t_sol + t_ls
    line = {"metric": "projected-Newton iteration time (1M-tet deformation), per phase", "value": total / iters * 1e3, "unit": "ms/iteration",
            "n_gpus": 1, "steps": iters, "warmup": args.warmup, "higher_is_better": False, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3: {iters} projected-Newton iterations, Kuhn cube n={n} ({len(T):,} tets), 4 pinned vertices, eps=1e-9, "
                                   f"w_identity=1e-9, PCG rel_tol={args.newton_tol}", "nnz": int(nnz)},
            "phases_ms_per_iteration": {"assembly_eval_with_hessian_proj": t_asm / iters * 1e3, "solve_pcg_standin_for_cudss": t_sol / iters * 1e3,
                                        "line_search": t_ls / iters * 1e3},
            "f_first": log[0]["f"], "f_last": log[-1]["f"], "iterations": log}
    print(json.dumps(line))


def run_gauss_newton(args):
    """BASELINE.json configs[3]: frame-field style Gauss-Newton via VectorFunction per-element residual Jacobians on a 2M-face
    synthetic triangle grid.  The real polycurl functor lives in TinyAD-Examples (not in the reference tree); the stand-in is the
    complex-arithmetic per-edge residual of csrc/energies.cuh (SOS_POLYCURL2D, 2 residuals x 4 variables per element,
    std::complex<Scalar> operators of Scalar.hh:1151-1320).  Timed: eval_sum_of_squares_with_derivatives (r, J values in the fixed
    CSC pattern, f, g = 2 J^T r), and apart from it the matrix-free Gauss-Newton direction and the line search."""
    import torch
    import tinyad_b200 as tad
    from test_vector_gpu import polycurl_problem
    N = 1000 if args.workload == "c4" else 64
    p, x0 = polycurl_problem(N)
    n_el = len(p.terms[0][1])
    t0 = time.perf_counter()
    fn = p.gpu()
    nnz = fn.nnz
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    x = torch.from_numpy(x0.copy()).cuda()
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    r = torch.empty(fn.n_outputs, dtype=torch.float64, device="cuda")
    J = torch.empty(nnz, dtype=torch.float64, device="cuda")
    d = torch.empty_like(g)
    xn = torch.empty_like(g)
    for _ in range(max(3, args.warmup)):
        fn.veval_sum_of_squares_with_derivatives(x, g, r, J)
    torch.cuda.synchronize()
    stream_ptr = tad.ctypes.c_void_p()
    tad.runtime().tad_function_get_stream(fn.h, tad.ctypes.byref(stream_ptr))
    ext = torch.cuda.ExternalStream(stream_ptr.value)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    for _ in range(args.steps):
        f = fn.veval_sum_of_squares_with_derivatives(x, g, r, J)
    ev1.record(ext)
    torch.cuda.synchronize()
    t_eval = ev0.elapsed_time(ev1) * 1e-3 / args.steps
    # a few Gauss-Newton iterations, phases timed with the host clock around the synchronous C calls
    t_dir = t_ls = 0.0
    log = []
    gn_iters = 3
    for _ in range(gn_iters):
        f = fn.veval_sum_of_squares_with_derivatives(x, g, r, J)
        t1 = time.perf_counter()
        cg, rel = fn.gauss_newton_direction(r, J, d, w_identity=1e-6, rel_tol=1e-6, max_iters=5000)
        t2 = time.perf_counter()
        f_new, step, n_evals = fn.line_search(x, d, f, g, xn)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        t_dir += t2 - t1; t_ls += t3 - t2
        log.append({"f": f, "f_new": f_new, "cg_iters": cg, "step": step, "ls_evals": n_evals})
        x, xn = xn, x
    line = {"metric": "VectorFunction residual-Jacobian elements/s (eval_sum_of_squares_with_derivatives)", "value": n_el / t_eval,
            "unit": "elements/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_eval * 1e3, "higher_is_better": True,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4: {N}x{N} grid ({2 * N * N:,} faces, {n_el:,} edge elements, 2 residuals x 4 variables each), polycurl-style "
                                   f"complex residual stand-in", "n_outputs": int(fn.n_outputs), "nnz_jacobian": int(nnz), "setup_s_pattern": setup_s},
            "gauss_newton_ms_per_iteration": {"direction_matrix_free_pcg": t_dir / gn_iters * 1e3, "line_search": t_ls / gn_iters * 1e3},
            "iterations": log}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload in ("c3", "c3small"):
        return run_newton(args)
    if args.workload in ("c4", "c4small"):
        return run_gauss_newton(args)

    import torch
    import torch.distributed as dist
    import tinyad_b200 as tad

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tinyad_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sampler = ClockSampler(local_rank)
    sampler.start()
    d, kind, V, conn, data, X, desc = workload_mesh(args.workload, rank, world)
    n_el = len(conn)
    assembly = tad.ASSEMBLY_GATHER if args.assembly == "gather" else tad.ASSEMBLY_ATOMIC

    t0 = time.perf_counter()
    fn = tad.Function(d, len(V), device=local_rank, assembly=assembly)
    fn.add_term(kind, conn, data)
    plan = None
    if world > 1:
        # vertex ownership + halo rows: the owner's pattern gets slots for the blocks its neighbours send
        from tinyad_b200.dist import HaloPlan
        plan = HaloPlan(d, len(V), [conn])
        fn.add_pattern_blocks(*plan.extra_pattern_blocks())
    nnz = fn.nnz                       # builds the pattern + scatter maps
    if plan is not None:
        plan.finalize(*fn.pattern(), device="cuda")
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    x_host = torch.from_numpy(X.reshape(-1).copy()).pin_memory()
    x_dev = x_host.cuda()
    g_dev = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H_dev = torch.empty(nnz, dtype=torch.float64, device="cuda")
    g_host = torch.empty(fn.n_vars, dtype=torch.float64).pin_memory()
    H_host = torch.empty(nnz, dtype=torch.float64).pin_memory()
    fsum = torch.zeros(1, dtype=torch.float64, device="cuda")

    def step():
        f = fn.eval_with_hessian_proj(x_dev, g_dev, H_dev)
        if world > 1:                   # exchange step: halo-row H values to the owners, g and f all-reduced (NCCL)
            plan.exchange(H_dev, g_dev)
            fsum[0] = f
            dist.all_reduce(fsum)
            torch.cuda.current_stream().synchronize()   # H_dev / g_dev are reused by the next step
        return f

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    load_t0 = time.time()
    for _ in range(args.warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_ptr = tad.ctypes.c_void_p()
    tad.runtime().tad_function_get_stream(fn.h, tad.ctypes.byref(stream_ptr))
    ext = torch.cuda.ExternalStream(stream_ptr.value) if world == 1 else torch.cuda.current_stream()
    t0 = time.perf_counter()
    ev0.record(ext)
    for _ in range(args.steps):
        f = step()
    ev1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    load_t1 = time.time()
    t_step = torch.tensor([dev_s / args.steps], dtype=torch.float64, device="cuda")
    n_total = torch.tensor([float(n_el)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_total)
    t_step, n_total = t_step.item(), n_total.item()

    # ---- e2e through the host-buffer C ABI (pinned buffers; H2D x, D2H g + H values inside) ----
    for _ in range(2):
        fn.eval_with_hessian_proj_host(x_host.numpy(), out_g=g_host.numpy(), out_H=H_host.numpy())
    barrier()
    te = time.perf_counter()
    e2e_steps = max(3, args.steps // 4)
    for _ in range(e2e_steps):
        fn.eval_with_hessian_proj_host(x_host.numpy(), out_g=g_host.numpy(), out_H=H_host.numpy())
    barrier()
    e2e_t = torch.tensor([(time.perf_counter() - te) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_t = e2e_t.item()

    # ---- per-kernel durations (CUDA events on the function's stream, separate short pass) ----
    fn.set_timing(True)
    phase = {"element_ms": [], "projection_ms": [], "assembly_ms": [], "total_ms": []}
    for _ in range(5):
        fn.eval_with_hessian_proj(x_dev, g_dev, H_dev)
        for k, v in fn.last_timings().items():
            phase[k].append(v)
    fn.set_timing(False)
    clocks = sampler.stop(load_t0, time.time())      # warm-up, timed steps, e2e and the per-kernel pass: all under load
    phase = {k: float(np.median(v)) for k, v in phase.items()}
    stats = fn.projection_stats()
    phi = stats["rebuilt"] / max(1, n_el)

    if rank == 0:
        fp64_peak = tad.fp64_peak_tflops(local_rank, 0.5)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tet = d == 3
        flops_el = (FLOPS_AD_TET + phi * FLOPS_PROJ_TET) if tet else (FLOPS_AD_TRI + phi * FLOPS_PROJ_TRI)
        bytes_el = BYTES_TET if tet else BYTES_TRI
        k_ms = {"element": phase["element_ms"], "projection": phase["projection_ms"], "assembly": phase["assembly_ms"]}
        dominant = max(k_ms, key=k_ms.get)
        dom_flops = {"element": FLOPS_AD_TET if tet else FLOPS_AD_TRI, "projection": phi * (FLOPS_PROJ_TET if tet else FLOPS_PROJ_TRI),
                     "assembly": 0.0}[dominant]
        dom_s = k_ms[dominant] * 1e-3
        if dominant == "assembly":      # pure scatter: HBM-side roofline
            achieved = bytes_el * n_el / dom_s / 1e9
            roof = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak}
        else:
            achieved = dom_flops * n_el / dom_s / 1e12
            roof = {"bound": "fp64", "kernel": dominant, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak}
        # dram__bytes_read + dram__bytes_write of the projection kernels (A + B1 + B2) from the ncu --set full capture in
        # profiles/r01_v5_ncu_full_hot_kernels.csv (C2 workload; per step = per launch of each of the three kernels)
        roof["traffic"] = 2.48e9 if (args.workload == "c2" and dominant == "projection") else None
        roof["peak_source"] = "FP64: DFMA-chain microbenchmark run in this process (measured); HBM: MEASURED_PEAKS.json" if peaks else "HBM fallback 6650 GB/s"
        step_tflops = flops_el * n_el / t_step / 1e12
        roof["whole_step"] = {"fp64_tflops": step_tflops, "frac_of_fp64_peak": step_tflops / fp64_peak,
                              "hbm_gbs": bytes_el * n_el / t_step / 1e9, "frac_of_hbm_peak": bytes_el * n_el / t_step / 1e9 / hbm_peak,
                              "algorithmic_flops_per_element": flops_el, "algorithmic_bytes_per_element": bytes_el, "phi_projected": phi}
        roof["kernel_ms"] = k_ms
        line = {
            "metric": METRIC, "value": n_total / t_step, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong" if args.workload == "c5" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "elements_total": int(n_total), "elements_per_gpu": n_el, "nnz_per_gpu": int(nnz),
                       "eps": 1e-9, "assembly": args.assembly,
                       "partition": "z-slabs, one per rank; vertex owner = lowest rank; halo-row H values sent to the owner, g/f all-reduced"
                       if world > 1 else "none", "halo_bytes_rank0": (plan.halo_bytes if plan else 0),
                       "l2": "no flush: per-step working set (staging + CSR values) exceeds the 126 MB L2" if n_el > 200000 else "small workload, L2-resident",
                       "setup_s_pattern_and_maps": setup_s},
            "e2e": {"value": n_total / e2e_t, "unit": UNIT, "h2d_bytes_per_step": int(8 * fn.n_vars), "d2h_bytes_per_step": int(8 * (fn.n_vars + nnz + 1)),
                    "ms_per_step": e2e_t * 1e3},
            # per step: 4 element kernels (one per Hessian part; 1 for triangles), 2 reduction, 4 projection (A, B1, B2, fallback list),
            # 2 fused projection-C + assembly (all elements / the listed ones)
            "gpu_launches": (12 if d == 3 else 9) * args.steps,
            "clocks": clocks, "projection_stats": stats, "roofline": roof, "wall_s_timed_region": wall, "f": f,
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle
            line["cpu_baseline"] = cpu_baseline(args.cpu_sample_n, oracle.max_threads())
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
