"""CPU: the C-ABI library loads and exports every symbol include/tinyad_b200.h declares (no compute calls)."""
import os
import re

import tinyad_b200 as tad

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tinyad_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tad_[a-z0-9_]+)\s*\(", text)) - {"tad_launch_fn"})


def test_library_exports_every_declared_symbol():
    L = tad.runtime()
    syms = header_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(L, name), f"{name} declared in include/tinyad_b200.h but not exported"
    assert sorted(tad.ABI_SYMBOLS) == syms


def test_energies_library_loads():
    E = tad.energies()
    for name in ("tadx_create", "tadx_destroy", "tadx_handle", "tadx_add_term", "tadx_last_error"):
        assert hasattr(E, name)


def test_no_cpu_fallback():
    """Without a device the product path fails loudly instead of computing on the CPU."""
    import ctypes
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    status = tad.runtime().tad_function_create(2, 4, 0, 0, ctypes.byref(h))
    assert status == 3 and not h.value            # TAD_CUDA_ERROR
    assert b"no CPU fallback" in tad.runtime().tad_last_error()


def test_product_does_not_import_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "tinyad_b200")):
        for f in files:
            if f.endswith((".py", ".hh", ".cuh", ".cu", ".h")):
                src = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in src and "oracle/" not in src.replace("tests/oracle", ""), os.path.join(base, f)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/tinyad_b200.h must compile as C99 (-pedantic) and as C++17, with no other include path."""
    import subprocess
    import tempfile
    inc = os.path.join(ROOT, "include")
    with tempfile.TemporaryDirectory() as d:
        for name, compiler, std in (("abi.c", "/usr/bin/gcc", "-std=c99"), ("abi.cc", "/usr/bin/g++", "-std=c++17")):
            path = os.path.join(d, name)
            with open(path, "w") as fh:
                fh.write('#include "tinyad_b200.h"\nint main(void) { tad_function* f = 0; (void)f; return (int)TAD_OK; }\n')
            p = subprocess.run([compiler, std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", path], capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
