#!/bin/bash
# final single-GPU evidence of the round: ncu full capture (C2 hot kernels), launch list of the default bench command, the driver's bench commands
cd "$(dirname "$0")/.."
O=gpurun_out
TAD_CHUNK_ELEMENTS=-1 ncu --set full --clock-control none --import-source on -k "regex:second_order_part_kernel|project_kernel|project_c_assemble" -s 30 -c 11 -f -o $O/r02q_full_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/r02q_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file $O/r02q_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/r02q_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02q_bench_reference.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > $O/r02q_bench_c5_c2_c1.json 2> $O/r02q_bench.err; echo "bench rc=$?"
python bench.py --workload c4 --steps 20 --warmup 5 > $O/r02q_bench_c4.json 2>/dev/null; echo "c4 rc=$?"
python bench.py --workload c3 --steps 5 --warmup 2 > $O/r02q_bench_c3.json 2>/dev/null; echo "c3 rc=$?"
python - <<P
import json
d=json.loads(open("$O/r02q_bench_c5_c2_c1.json").read().strip().split(chr(10))[-1])
def show(d): print(d["config"]["workload"][:30], round(d["ms_per_step"],4), round(d["value"]/1e6,1), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, d["check"]["ok"], d["gpu_launches_per_step"], round(d["roofline"]["whole_step"]["frac_of_fp64_peak"],4), round(d["roofline"]["frac"],3), d["roofline"]["kernel"])
show(d); [show(v) for v in d["also"].values()]; print(d["clocks"]); print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
r=json.loads(open("$O/r02q_bench_reference.json").read().strip().split(chr(10))[-1]); print("reference", r["value"], r["cpu_baseline"]["cores"], r["config"]["elements_per_step"])
c=json.loads(open("$O/r02q_bench_c4.json").read().strip().split(chr(10))[-1]); print("c4", c["ms_per_step"], c["value"])
c=json.loads(open("$O/r02q_bench_c3.json").read().strip().split(chr(10))[-1]); print("c3", c["phases_ms_per_iteration"], c["f_first"], c["f_last"])
P
