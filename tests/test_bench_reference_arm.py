"""CPU: the contract of `bench.py --impl reference` (the reference's own CPU path timed on the host cores): one JSON line with the
product arm's metric / unit / config, `impl: "reference"`, a `cpu_baseline` describing the run and an `e2e` equal to the line's value;
under torchrun only rank 0 works.  Runs the `small` workload so that it takes seconds."""
import json
import os
import subprocess
import sys

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env=None, args=("--workload", "small", "--steps", "1", "--warmup", "1")):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_line():
    out = run({"OMP_NUM_THREADS": "1"})                      # torchrun exports OMP_NUM_THREADS=1: the arm must set its thread count itself
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "elements/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    assert cb["kind"] == ("reference" if oracle.ref_available() else "port")
    assert cb["cores"] == max(1, len(os.sched_getaffinity(0)))                     # not the 1 thread OMP_NUM_THREADS asked for
    if cb["kind"] == "reference":
        assert cb["oracle_port"]["value"] > 0                                      # the oracle restatement timed beside it
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_exit_without_work():
    assert run({"RANK": "1", "WORLD_SIZE": "2"}).strip() == ""
