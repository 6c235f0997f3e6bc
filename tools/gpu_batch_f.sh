#!/bin/bash
# dev batch: topology probe, selftest, fused-kernel launch-bounds variants, ncu captures (C2 hot kernels, C1 fused kernel, launch list of the default bench command)
cd "$(dirname "$0")/.."
O=gpurun_out
{ lscpu | grep -iE "numa|socket|model name|^CPU\(s\)"; numactl -H 2>/dev/null | head -20; nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; nvidia-smi --query-gpu=index,pci.bus_id --format=csv; free -g | head -2; } > $O/r02f_topology.txt 2>&1
python -m pytest tests/test_scalar_golden.py -m gpu -q -x 2>&1 | tail -3
for so in "" tools/pb_energies_mb3.so tools/pb_energies_mb4.so; do
  TINYAD_ENERGIES_SO=${so:-tinyad_b200/libtinyad_b200_energies.so} python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('energies=$so', d['ms_per_step'], d['roofline']['kernel_ms'], d['check']['ok'])"
done 2>&1 | tee $O/r02f_fused_launch_bounds.txt
TAD_CHUNK_ELEMENTS=-1 ncu --set full --clock-control none --import-source on -k "regex:second_order_part_kernel|project_kernel|project_c_assemble" -s 30 -c 11 -f -o $O/r02f_full_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/r02f_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:second_order_fused" -s 3 -c 1 -f -o $O/r02f_full_c1 python bench.py --workload c1 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/r02f_ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02f_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/r02f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ls -la $O | tail -8
