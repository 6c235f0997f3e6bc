"""Builds the in-tree CUDA libraries for sm_100a with nvcc (cross-compiles without a GPU).

  libtinyad_b200.so            csrc/runtime.cu   the C-ABI runtime (include/tinyad_b200.h)
  libtinyad_b200_energies.so   csrc/energies.cu  element functors of the tests / benchmark (a "user TU")
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = os.environ.get("TINYAD_NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = [
    NVCC, "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-ccbin", "/usr/bin/g++",
    "-Xcompiler", "-fPIC", "-shared",
    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "include"),
]

RUNTIME_SO = os.path.join(HERE, "libtinyad_b200.so")
ENERGIES_SO = os.path.join(HERE, "libtinyad_b200_energies.so")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    out = [os.path.join(ROOT, "include", "tinyad_b200.h")]
    for base, _, files in os.walk(os.path.join(HERE, "include")):
        out += [os.path.join(base, f) for f in files]
    return out


def build(force=False, verbose=False):
    hdrs = _headers()
    jobs = []
    src = os.path.join(HERE, "csrc", "runtime.cu")
    if force or _newer(RUNTIME_SO, [src] + hdrs):
        jobs.append(COMMON + ["-o", RUNTIME_SO, src])
    src = os.path.join(HERE, "csrc", "energies.cu")
    if force or _newer(ENERGIES_SO, [src, RUNTIME_SO] + hdrs) or jobs:
        jobs.append(COMMON + ["-Xptxas", "-v" if verbose else "-O3", "-o", ENERGIES_SO, src,
                              "-L", HERE, "-ltinyad_b200", "-Xlinker", "-rpath=$ORIGIN"])
    for cmd in jobs:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return RUNTIME_SO, ENERGIES_SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", RUNTIME_SO, ENERGIES_SO)
