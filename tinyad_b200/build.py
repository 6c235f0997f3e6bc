"""Builds the in-tree CUDA libraries for sm_100a with nvcc (cross-compiles without a GPU).

  libtinyad_b200.so            csrc/runtime.cu, projection.cu, assembly.cu, comm.cu, newton.cu    the C-ABI runtime (include/tinyad_b200.h)
  libtinyad_b200_energies.so   csrc/energies*.cu  element functors of the tests / benchmark (a "user TU");
                                                  the Double<12> tet kernel is split into one object per Hessian part
                                                  so that the heavy instantiations compile in parallel.
"""
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = os.environ.get("TINYAD_NVCC", "/usr/local/cuda/bin/nvcc")
TET_PARTS = int(os.environ.get("TADX_TET_PARTS", "4"))
FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-ccbin", "/usr/bin/g++",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "include"),
]

RUNTIME_SO = os.path.join(HERE, "libtinyad_b200.so")
ENERGIES_SO = os.path.join(HERE, "libtinyad_b200_energies.so")
OBJ_DIR = os.path.join(HERE, "build")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    out = [os.path.join(ROOT, "include", "tinyad_b200.h"), os.path.join(HERE, "csrc", "energies.cuh"), os.path.abspath(__file__)]
    for base, _, files in os.walk(os.path.join(HERE, "include")):
        out += [os.path.join(base, f) for f in files]
    return out


def _run(cmd, verbose):
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    # stamp the output with the START time of the compilation: a source edited while nvcc was running is then newer than
    # the object and gets rebuilt next time (nvcc reads its inputs at the start)
    out = cmd[cmd.index("-o") + 1]
    if os.path.exists(out):
        os.utime(out, (t0, t0))


def build(force=False, verbose=False):
    hdrs = _headers()
    os.makedirs(OBJ_DIR, exist_ok=True)
    ptxas = ["-Xptxas", "-v"] if verbose else []
    jobs = []
    inc = os.path.join(HERE, "include", "TinyAD")
    api = hdrs[0]
    # runtime library: one object per translation unit (the projection kernels once per group of K, the assembly kernels once per
    # variable dimension, so that the heavy instantiations compile in parallel), linked into libtinyad_b200.so
    common = os.path.join(HERE, "csrc", "rt_common.cuh")
    proj_deps = [common, os.path.join(inc, "Scalar.hh"), os.path.join(inc, "Detail", "HessLayout.hh"), os.path.join(inc, "Detail", "Projection.hh")]
    rt_units = [("runtime.o", "runtime.cu", proj_deps, []), ("newton.o", "newton.cu", [], []), ("comm.o", "comm.cu", [common], [])]
    rt_units += [(f"projection{p}.o", "projection.cu", proj_deps, [f"-DTAD_PROJ_PART={p}"]) for p in range(6)]
    rt_units += [(f"assembly{p}.o", "assembly.cu", proj_deps, [f"-DTAD_ASM_PART={p}"]) for p in range(3)]
    rt_objs, relink_runtime = [], not os.path.exists(RUNTIME_SO)
    for obj, cu, deps, defs in rt_units:
        o = os.path.join(OBJ_DIR, obj)
        src = os.path.join(HERE, "csrc", cu)
        if not os.path.exists(src):
            continue
        rt_objs.append(o)
        if force or _newer(o, [src, api, os.path.abspath(__file__)] + deps):
            jobs.append([NVCC] + FLAGS + ptxas + defs + ["-c", "-o", o, src])
            relink_runtime = True
    hdrs = [h for h in hdrs if not h.endswith("Projection.hh")]   # only the runtime includes the projection routines
    objs = []
    units = [("energies.o", "energies.cu", [])] + [(f"energies_tet_part{p}.o", "energies_tet_part.cu", [f"-DTADX_PART={p}"]) for p in range(TET_PARTS)]
    for obj, cu, defs in units:
        o = os.path.join(OBJ_DIR, obj)
        objs.append(o)
        src = os.path.join(HERE, "csrc", cu)
        if force or _newer(o, [src] + hdrs):
            jobs.append([NVCC] + FLAGS + ptxas + [f"-DTADX_TET_PARTS={TET_PARTS}"] + defs + ["-c", "-o", o, src])
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    if relink_runtime:
        _run([NVCC] + FLAGS + ["-shared", "-o", RUNTIME_SO] + rt_objs, verbose)
    if jobs or not os.path.exists(ENERGIES_SO):
        _run([NVCC] + FLAGS + ["-shared", "-o", ENERGIES_SO] + objs + ["-L", HERE, "-ltinyad_b200", "-Xlinker", "-rpath=$ORIGIN"], verbose)
    return RUNTIME_SO, ENERGIES_SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", RUNTIME_SO, ENERGIES_SO)
