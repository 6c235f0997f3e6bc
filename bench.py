#!/usr/bin/env python
"""bench.py -- projected-Hessian assembly throughput of the tet Double<12> path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            # this framework, one B200: C5 (10.1 M tets), the north-star configuration
  torchrun ... bench.py --gpus N ...                        # N ranks: C5 split in N z-slabs (strong scaling), exchange inside the runtime
  python bench.py --impl reference ...                      # the reference's own CPU/OpenMP code (oracle/_ref; else the oracle port) on the same workload

A "step" is one eval_with_hessian_proj over the whole mesh: x is resident in HBM, f / g / CSR values are
left in HBM (`value`); `e2e` times the same call through the host-buffer C ABI (pinned host x, g, H values;
H2D + D2H inside the timed region).  Pattern / scatter-map construction is one-time setup and reported apart.
Every line carries a `check` (against the CPU oracle at N = 1, against a single-rank evaluation of the whole mesh at N > 1);
the C2 (and, at N = 1, C1) lines ride along under `also`.  Prints ONE JSON line on rank 0.
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
METRIC_TRI = "proj-Hessian assembled elements/s (triangle Double<6>)"
UNIT = "elements/s"
# algorithmic work per tet (packed-symmetric flop count; compulsory HBM bytes)
FLOPS_AD_TET, FLOPS_PROJ_TET, BYTES_TET = 41823.0, 17712.0, 353.0
FLOPS_AD_TRI, FLOPS_PROJ_TRI, BYTES_TRI = 3943.0, 2268.0, 216.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c2", "c5", "c1", "small", "c3", "c3small", "c4", "c4small"],
                    help="c5: Kuhn cube n=119 (10,110,954 tets), on one GPU or split over the GPUs in z-slabs (strong scaling) [default: the "
                         "north-star configuration]; c2: Kuhn cube n=55 (998,250 tets) per GPU (weak); "
                         "c1: 512^2 triangle grid; small: n=16 smoke size; c3: projected-Newton loop on the C2 mesh, timed per phase "
                         "(--steps = Newton iterations, default 20); c4: VectorFunction residual Jacobians + Gauss-Newton on a 2M-face grid")
    ap.add_argument("--newton-tol", type=float, default=1e-6, help="c3: relative residual of the PCG solve (inexact Newton)")
    ap.add_argument("--assembly", default="atomic", choices=["atomic", "gather"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C2 / C1 lines that ride along under `also`")
    return ap.parse_args()


def workload_mesh(name, rank, world, world_for_shape=None):
    """Returns (d, kind, V, conn, data, x, description).  For world > 1 each rank gets a z-slab.  world_for_shape: the lattice of a
    weak-scaling workload is n x n x (n * world_for_shape); with rank 0 of world 1 this returns the WHOLE mesh of an N-rank run."""
    ws = world if world_for_shape is None else world_for_shape
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
        nz_total = n * ws
        z0, nz = (n * rank, n) if world_for_shape is None else (0, nz_total)
        desc = f"{name.upper()}: Kuhn cube n={n} ({6 * n ** 3:,} tets) per GPU (n x n x {nz_total} lattice), Double<12>, eval_with_hessian_proj, eps=1e-9"
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


def host_threads():
    """Host cores this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


WORKLOADS = {
    # name: (kind, lattice edge, description of the FULL workload)
    "c5": ("tet", 119, "C5: Kuhn cube n=119 (10,110,954 tets), Double<12>, eval_with_hessian_proj, eps=1e-9"),
    "c2": ("tet", 55, "C2: Kuhn cube n=55 (998,250 tets), Double<12>, eval_with_hessian_proj, eps=1e-9"),
    "small": ("tet", 16, "small: Kuhn cube n=16 (24,576 tets), Double<12>, eval_with_hessian_proj, eps=1e-9"),
    "c1": ("tri", 512, "C1: 512x512 grid (524,288 triangles), Double<6>, eval_with_hessian_proj, eps=1e-9"),
}


class CpuSample:
    """A bounded sample of a workload for the CPU arm: the first `layers` cube layers (tets) / grid rows (triangles) of the SAME
    mesh, same x, same energy, same eps.  elements/s does not depend on the number of layers (every layer has the same work)."""

    def __init__(self, workload, layers):
        import oracle
        from tinyad_b200 import meshes
        kind, n, _ = WORKLOADS[workload]
        layers = max(1, min(layers, n))
        self.workload, self.layers, self.n = workload, layers, n
        if kind == "tet":
            V, T = meshes.kuhn_cube(n, n, layers, z0=0, nz_total=n)
            self.d, self.nv = 3, len(V)
            self.terms = [oracle.Term(oracle.SYMDIRICHLET3D, T, meshes.tet_rest_data(V, T))]
            self.n_el, self.per_layer = len(T), 6 * n * n
            self.what = f"first {layers} of {n} cube layers of the n={n} lattice ({len(T):,} tets)"
        else:
            V, F = meshes.grid_2d(n)
            F = F[: 2 * n * layers]
            self.d, self.nv = 2, len(V)
            self.terms = [oracle.Term(oracle.SYMDIRICHLET2D, F, meshes.tri_rest_data(V, F))]
            self.n_el, self.per_layer = len(F), 2 * n
            self.what = f"first {layers} of {n} grid rows of the {n}x{n} grid ({len(F):,} triangles)"
        self.x = meshes.deform(V, 1.0 / n, seed=0).reshape(-1)

    def step(self, threads, engine="port"):
        """engine "reference": oracle/_ref (the unmodified TinyAD headers over oracle/eigen_shim); "port": the oracle restatement."""
        import oracle
        fn = oracle.ref_scalar_eval if engine == "reference" else oracle.scalar_eval
        t0 = time.perf_counter()
        r = fn(self.d, self.nv, self.terms, oracle.HESSIAN_PROJ, self.x, n_threads=threads)
        return time.perf_counter() - t0, r.phases


def cpu_engine():
    """("reference", description) when oracle/_ref/libtinyad_ref.so exists (built in the build container from /root/reference, shipped
    prebuilt to the GPU box), else ("port", description)."""
    import oracle
    if oracle.ref_available():
        return "reference", ("oracle/_ref = the reference's own headers (TinyAD::ScalarFunction::eval_with_hessian_proj, unmodified, compiled "
                             "in place) over oracle/eigen_shim, a minimal eager stand-in for the Eigen API (Eigen is not in the image; real "
                             "Eigen would vectorise the fixed-size Hessian updates)")
    return "port", "oracle = C++/OpenMP restatement of the reference path (oracle/_ref is not built on this machine)"


def cpu_sample_for(workload, budget_s, threads, engine="port"):
    """Chooses the number of layers so that one step takes about budget_s seconds on `threads` threads (calibrated on one layer)."""
    probe = CpuSample(workload, 1)
    probe.step(threads, engine)                 # warm-up (OpenMP pool, page faults)
    t1, _ = probe.step(threads, engine)
    layers = int(max(1, min(probe.n, budget_s / max(t1, 1e-4))))
    if engine == "reference":
        # the reference keeps every element's Scalar (1.3 kB for Double<12>) and 144 triplets (2.3 kB) alive until setFromTriplets:
        # bound one step to ~600 k elements (~6 GB of host memory)
        layers = max(1, min(layers, 600000 // probe.per_layer))
    return probe if layers == 1 else CpuSample(workload, layers)


def cpu_baseline(workload, budget_s=4.0):
    """cpu_baseline leg of the product line: the reference's OpenMP eval_with_hessian_proj (oracle/_ref, else the oracle port) on a bounded
    sample of the same workload, all host threads and the reference's default (max - 1, Detail/Parallel.hh:27)."""
    threads = host_threads()
    engine, what = cpu_engine()
    smp = cpu_sample_for(workload, budget_s, threads, engine)
    times, phases = [], None
    for _ in range(3):
        t, phases = smp.step(threads, engine)
        times.append(t)
    t = float(np.median(times))
    out = {"value": smp.n_el / t, "unit": UNIT, "cores": threads, "kind": engine,
           "sample": f"{smp.what}, same x / energy / eps as the GPU arm, median of 3 steps after warm-up; {what}",
           "phases_s": phases, "seconds_per_step": t}
    if threads > 1:
        td, _ = smp.step(threads - 1, engine)
        out["reference_default_threads"] = {"cores": threads - 1, "value": smp.n_el / td, "note": "Detail/Parallel.hh:27: omp_get_max_threads() - 1"}
    if engine == "reference":
        tp, pp = smp.step(threads, "port")
        out["oracle_port"] = {"value": smp.n_el / tp, "cores": threads, "phases_s": {k: pp[k] for k in ("eval_s", "accumulate_s", "compress_s")},
                              "note": "the oracle restatement (the checker of the parity tests) on the same sample"}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU path (oracle/_ref: the unmodified TinyAD headers over an Eigen-API shim; the oracle
    port where that library is absent) on the box's host cores,
    thread count set explicitly (torchrun exports OMP_NUM_THREADS=1), on the SAME workload as the product arm.  Each step is a
    bounded sample of that workload (whole cube layers of the same mesh, same x) sized so that warm-up + steps end within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.lib()
    threads = host_threads()
    engine, what = cpu_engine()
    workload = args.workload if args.workload in WORKLOADS else "c5"
    total = max(1, args.steps + args.warmup)
    budget = min(6.0, max(0.5, 150.0 / total))          # seconds per step: the whole run stays below ~3 minutes
    smp = cpu_sample_for(workload, budget, threads, engine)
    times = []
    for i in range(args.warmup + args.steps):
        t, phases = smp.step(threads, engine)
        if i >= args.warmup:
            times.append(t)
    t = float(np.mean(times))
    value = smp.n_el / t
    extra = {}
    if threads > 1:
        td, _ = smp.step(threads - 1, engine)
        extra = {"reference_default_threads": {"cores": threads - 1, "value": smp.n_el / td, "note": "Detail/Parallel.hh:27: omp_get_max_threads() - 1"}}
    if engine == "reference":
        tp, pp = smp.step(threads, "port")
        extra["oracle_port"] = {"value": smp.n_el / tp, "cores": threads, "phases_s": {k: pp[k] for k in ("eval_s", "accumulate_s", "compress_s")},
                                "note": "the oracle restatement (the checker of the parity tests) on the same sample"}
    full_n = {"c5": 10110954, "c2": 998250, "small": 24576, "c1": 524288}[workload]
    line = {"metric": METRIC if WORKLOADS[workload][0] == "tet" else METRIC_TRI, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong" if workload == "c5" else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOADS[workload][2], "elements_total": full_n, "eps": 1e-9,
                       "elements_per_step": smp.n_el, "seconds_full_workload_extrapolated": full_n / value},
            "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": threads, "kind": engine,
                                  "sample": f"each step: {smp.what} of the same workload (same x, energy, eps); {what}; {threads} threads set explicitly",
                                  "phases_s": phases}, **extra),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


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


def plane_check(fn, name, x, g, H):
    """Correctness bit of an N = 1 line: the CSR rows of the first lattice plane (z = 0 vertices; first grid row for triangles) only
    receive contributions from the first cube layer, so they must equal the ORACLE's rows for that layer alone -- index arrays
    bit-exact, values within north_star's tolerance.  (tests/test_fullsize_gpu.py compares EVERY row of C1 / C2 with the oracle.)"""
    import oracle
    kind, n, _ = WORKLOADS[name]
    smp = CpuSample(name, 1)
    ref = oracle.scalar_eval(smp.d, smp.nv, smp.terms, oracle.HESSIAN_PROJ, smp.x, n_threads=host_threads())
    rows = smp.d * ((n + 1) ** 2 if kind == "tet" else (n + 1))
    outer, inner = fn.pattern()
    end, rend = int(outer[rows]), int(ref.outer[rows])
    pattern_equal = bool(end == rend and np.array_equal(outer[:rows + 1], ref.outer[:rows + 1]) and np.array_equal(inner[:end], ref.inner[:rend]))
    Hg, gg = H[:end].cpu().numpy(), g[:rows].cpu().numpy()
    rel_H = float(np.abs(Hg - ref.values[:rend]).max() / np.abs(ref.values[:rend]).max()) if pattern_equal else float("nan")
    rel_g = float(np.abs(gg - ref.g[:rows]).max() / np.abs(ref.g[:rows]).max())
    return {"against": f"CPU oracle on the first layer: the {rows:,} CSR rows ({end:,} values) of the first lattice plane", "pattern_equal": pattern_equal,
            "max_rel_H": rel_H, "max_rel_g": rel_g, "tol_H_proj": 1e-10, "tol_g": 1e-12,
            "ok": bool(pattern_equal and rel_H <= 1e-10 and rel_g <= 1e-12)}


def distributed_check(torch, dist, tad, fn, name, world, local_rank, g, H, f):
    """Correctness bit of an N > 1 line: every rank also evaluates the WHOLE mesh alone on its GPU (the 1-rank path is what the
    GPU tests compare with the oracle row by row) and compares the rows it owns after the exchange: index arrays of those rows
    bit-exact, values 1e-10 (projected), owned g entries 1e-12, f 1e-12.  Max over ranks."""
    d, kind, V, conn, data, X, _ = workload_mesh(name, 0, 1, world_for_shape=world)
    ref = tad.Function(d, len(V), device=local_rank)
    ref.add_term(kind, conn, data)
    ro, ri = ref.pattern()
    xr = torch.from_numpy(X.reshape(-1).copy()).cuda()
    gr = torch.empty(ref.n_vars, dtype=torch.float64, device="cuda")
    Hr = torch.empty(ref.nnz, dtype=torch.float64, device="cuda")
    fr = ref.eval_with_hessian_proj(xr, gr, Hr)
    owned = np.repeat(fn.vertex_owner() == dist.get_rank(), d)            # rows this rank owns
    lo, li = fn.pattern()
    rows = np.nonzero(owned)[0]
    cnt_l, cnt_r = np.diff(lo)[rows], np.diff(ro)[rows]
    pattern_equal = bool(np.array_equal(cnt_l, cnt_r))
    rel_H = rel_g = float("nan")
    if pattern_equal:
        def gather_idx(outer):
            starts = outer[rows].astype(np.int64)
            return np.repeat(starts - np.concatenate([[0], np.cumsum(cnt_l)[:-1]]), cnt_l) + np.arange(int(cnt_l.sum()))
        il, ir = gather_idx(lo), gather_idx(ro)
        pattern_equal = bool(np.array_equal(li[il], ri[ir]))
        Hl = H.cpu().numpy()[il]
        Hf = Hr.cpu().numpy()
        rel_H = float(np.abs(Hl - Hf[ir]).max() / np.abs(Hf).max())
        gl, gf = g.cpu().numpy()[rows], gr.cpu().numpy()
        rel_g = float(np.abs(gl - gf[rows]).max() / np.abs(gf).max())
    ref.close()
    agg = torch.tensor([0.0 if pattern_equal else 1.0, rel_H if rel_H == rel_H else 1.0, rel_g if rel_g == rel_g else 1.0,
                        abs(f - fr) / abs(fr)], dtype=torch.float64, device="cuda")
    dist.all_reduce(agg, op=dist.ReduceOp.MAX)
    bad, rel_H, rel_g, rel_f = agg.tolist()
    return {"against": "single-rank evaluation of the whole mesh on every rank's GPU, rows owned after the exchange (max over ranks)",
            "pattern_equal": bad == 0.0, "max_rel_H": rel_H, "max_rel_g": rel_g, "rel_f": rel_f, "tol_H_proj": 1e-10, "tol_g": 1e-12,
            "ok": bool(bad == 0.0 and rel_H <= 1e-10 and rel_g <= 1e-12 and rel_f <= 1e-12)}


def kernel_traffic(name):
    """DRAM bytes per launch of the hot kernels from the committed ncu --set full capture of this workload (profiles/), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if name in t:
            return t[name]
        if name in ("c5", "small") and "c2" in t:      # same kernels, same bytes per tet: the capture was taken on the C2 mesh
            e = dict(t["c2"])
            e["source"] = e.get("source", "") + "; per-element bytes measured on the C2 mesh, scaled to this mesh"
            return e
        return None
    except Exception:
        return None


def run_workload(args, name, torch, dist, tad, rank, world, local_rank, comm, sampler_index=None, steps=None, warmup=None, with_cpu=True):
    """One bench line (dict) for one workload on `world` GPUs; only rank 0's return value is complete."""
    steps = args.steps if steps is None else steps
    warmup = max(3, args.warmup if warmup is None else warmup)
    d, kind, V, conn, data, X, desc = workload_mesh(name, rank, world)
    n_el = len(conn)
    assembly = tad.ASSEMBLY_GATHER if args.assembly == "gather" else tad.ASSEMBLY_ATOMIC

    t0 = time.perf_counter()
    fn = tad.Function(d, len(V), device=local_rank, assembly=assembly)
    fn.add_term(kind, conn, data)
    if world > 1:
        fn.set_comm(comm)              # vertex ownership, halo rows and the exchange lists are built with the pattern, inside the runtime
    nnz = fn.nnz                       # builds the pattern + scatter maps
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    x_host = torch.from_numpy(X.reshape(-1).copy()).pin_memory()
    x_dev = x_host.cuda()
    g_dev = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H_dev = torch.empty(nnz, dtype=torch.float64, device="cuda")
    g_host = torch.empty(fn.n_vars, dtype=torch.float64).pin_memory()
    H_host = torch.empty(nnz, dtype=torch.float64).pin_memory()

    def step():
        # one eval_with_hessian_proj; at N > 1 the call contains the halo exchange (NCCL send/recv of halo-row H values and halo g
        # entries to the owners, f all-reduced), overlapped with the assembly of the remaining slabs
        return fn.eval_with_hessian_proj(x_dev, g_dev, H_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    load_t0 = time.time()
    for _ in range(warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_ptr = tad.ctypes.c_void_p()
    tad.runtime().tad_function_get_stream(fn.h, tad.ctypes.byref(stream_ptr))
    ext = torch.cuda.ExternalStream(stream_ptr.value)      # the stream the kernels are launched on
    launches0 = fn.launch_count()
    t0 = time.perf_counter()
    ev0.record(ext)
    for _ in range(steps):
        f = step()
    ev1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    launches = fn.launch_count() - launches0
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    t_step = torch.tensor([dev_s / steps], dtype=torch.float64, device="cuda")
    n_total = torch.tensor([float(n_el)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_total)
    t_step, n_total = t_step.item(), n_total.item()

    # ---- e2e through the host-buffer C ABI (pinned buffers; H2D of x, D2H of g + H values inside; at N > 1 the exchange too) ----
    for _ in range(2):
        fn.eval_with_hessian_proj_host(x_host.numpy(), out_g=g_host.numpy(), out_H=H_host.numpy())
    barrier()
    te = time.perf_counter()
    e2e_steps = max(3, steps // 4)
    for _ in range(e2e_steps):
        fn.eval_with_hessian_proj_host(x_host.numpy(), out_g=g_host.numpy(), out_H=H_host.numpy())
    barrier()
    e2e_t = torch.tensor([(time.perf_counter() - te) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_t = e2e_t.item()

    # ---- per-kernel durations (CUDA events on the function's stream, separate short pass, one slab lane) ----
    fn.set_timing(True)
    phase = {"element_ms": [], "projection_ms": [], "assembly_ms": [], "total_ms": []}
    for _ in range(5):
        fn.eval_with_hessian_proj(x_dev, g_dev, H_dev)
        for k, v in fn.last_timings().items():
            phase[k].append(v)
    fn.set_timing(False)
    load_t1 = time.time()
    phase = {k: float(np.median(v)) for k, v in phase.items()}
    stats = fn.projection_stats()
    phi = stats["rebuilt"] / max(1, n_el)

    # ---- correctness bit ----
    f = step()
    if world > 1:
        check = distributed_check(torch, dist, tad, fn, name, world, local_rank, g_dev, H_dev, f)
        halo = torch.tensor([float(fn.halo_bytes())], dtype=torch.float64, device="cuda")
        dist.all_reduce(halo, op=dist.ReduceOp.MAX)
        halo_bytes = int(halo.item())
    else:
        check = plane_check(fn, name, X, g_dev, H_dev)
        halo_bytes = 0

    line = None
    if rank == 0:
        fp64_peak = tad.fp64_peak_tflops(local_rank, 0.5)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tet = d == 3
        f_ad, f_proj = (FLOPS_AD_TET, FLOPS_PROJ_TET) if tet else (FLOPS_AD_TRI, FLOPS_PROJ_TRI)
        flops_el = f_ad + phi * f_proj
        bytes_el = BYTES_TET if tet else BYTES_TRI
        k_ms = {"element": phase["element_ms"], "projection": phase["projection_ms"], "assembly": phase["assembly_ms"]}
        dominant = max(k_ms, key=k_ms.get)
        dom_flops = {"element": f_ad, "projection": phi * f_proj, "assembly": 0.0}[dominant]
        dom_s = k_ms[dominant] * 1e-3
        if dominant == "assembly":      # scatter: HBM-side roofline
            achieved = bytes_el * n_el / dom_s / 1e9
            roof = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak}
        else:
            achieved = dom_flops * n_el / dom_s / 1e12
            roof = {"bound": "fp64", "kernel": dominant, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak}
        tr = kernel_traffic(name)
        roof["traffic"] = (tr or {}).get(dominant + "_bytes_per_element", None)
        if roof["traffic"] is not None:
            roof["traffic"] *= n_el
            roof["traffic_source"] = tr.get("source")
        roof["peak_source"] = ("FP64: DFMA-chain microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 figure); HBM: MEASURED_PEAKS.json"
                               if peaks else "FP64: DFMA-chain microbenchmark run in this process; HBM: fallback 6650 GB/s (B200_PROFILING.md)")
        step_tflops = flops_el * n_el / t_step / 1e12 if world == 1 else flops_el * n_total / t_step / 1e12 / world
        roof["whole_step"] = {"fp64_tflops_per_gpu": step_tflops, "frac_of_fp64_peak": step_tflops / fp64_peak,
                              "hbm_gbs_per_gpu": bytes_el * n_total / world / t_step / 1e9,
                              "frac_of_hbm_peak": bytes_el * n_total / world / t_step / 1e9 / hbm_peak,
                              "algorithmic_flops_per_element": flops_el, "algorithmic_bytes_per_element": bytes_el, "phi_projected": phi,
                              "note": "algorithmic (dense packed) flop count of SURVEY.md 8(d); the kernels execute fewer FP64 instructions "
                                      "(structural sparsity masks, low-rank projection): measured FP64-pipe utilisation is in profiles/"}
        roof["kernel_ms"] = k_ms
        line = {
            "metric": METRIC if tet else METRIC_TRI, "value": n_total / t_step, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong" if name == "c5" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[name][2] if world == 1 or name == "c5" else desc, "elements_total": int(n_total),
                       "elements_rank0": n_el, "nnz_rank0": int(nnz), "eps": 1e-9, "assembly": args.assembly,
                       "partition": (f"{world} z-slabs of cube layers, one per rank; vertex owner = lowest rank touching it; inside the call: halo-row H "
                                     f"values and halo g entries sent to the owner (NCCL send/recv, overlapped with the remaining slabs), f all-reduced")
                       if world > 1 else "none", "halo_bytes_max_rank": halo_bytes,
                       "l2": "no flush: per-step working set (staging + CSR values) exceeds the 126 MB L2" if n_el > 200000 else "small workload, L2-resident",
                       "setup_s_pattern_and_maps": setup_s},
            "e2e": {"value": n_total / e2e_t, "unit": UNIT, "h2d_bytes_per_step": int(8 * fn.n_vars), "d2h_bytes_per_step": int(8 * (fn.n_vars + nnz + 1)),
                    "ms_per_step": e2e_t * 1e3, "note": "per rank; tad_eval_with_derivatives_host on pinned host buffers" + (", exchange included" if world > 1 else "")},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / steps,
            "check": check, "projection_stats": stats, "roofline": roof, "wall_s_timed_region": wall, "f": f,
            "_load_window": (load_t0, load_t1),
        }
        if with_cpu and not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(name)
    fn.close()
    del x_dev, g_dev, H_dev, g_host, H_host
    torch.cuda.empty_cache()
    return line


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
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = tad.Comm.from_torch_distributed(local_rank)

    sampler = ClockSampler(local_rank)
    sampler.start()
    line = run_workload(args, args.workload, torch, dist, tad, rank, world, local_rank, comm)
    # the other single-GPU configurations ride along under `also` (short runs; the headline is the line itself)
    also = {}
    if not args.no_extra and args.workload == "c5":
        extra = ["c2", "c1"] if world == 1 else ["c2"]
        for name in extra:
            sub = run_workload(args, name, torch, dist, tad, rank, world, local_rank, comm, steps=max(5, min(args.steps, 20)), warmup=3, with_cpu=(name == "c1"))
            if rank == 0:
                sub.pop("_load_window", None)
                also[name] = sub
    if rank == 0:
        w0, w1 = line.pop("_load_window")
        line["clocks"] = sampler.stop(w0, w1)
        if also:
            line["also"] = also
        print(json.dumps(line))
    else:
        sampler.stop()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
