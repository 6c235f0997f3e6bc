#!/bin/bash
# dev batch: B1 occupancy variants, final ncu captures (C2 hot kernels, launch list of the default bench command), sanitizer
cd "$(dirname "$0")/.."
O=gpurun_out
for mb in 1 7 8 10; do b=tools/pb_proj_bench_b1mb$mb; [ $mb = 1 ] && b=tools/pb_proj_bench; echo -n "B1_MINB=$mb "; $b 998250 7.0 128 128 128; done 2>&1 | tee $O/r02k_projbench_b1_occupancy.txt
TAD_CHUNK_ELEMENTS=-1 ncu --set full --clock-control none --import-source on -k "regex:second_order_part_kernel|project_kernel|project_c_assemble" -s 30 -c 11 -f -o $O/r02k_full_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/r02k_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file $O/r02k_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/r02k_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02k_sanitizer_memcheck_smoke.log 2>&1; echo "sanitizer smoke rc=$?"; tail -3 $O/r02k_sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_vector_gpu.py tests/test_dynamic_gpu.py -q -x -k "golden or fixture or planar or dynamic" > $O/r02k_sanitizer_memcheck_vector_dynamic.log 2>&1; echo "sanitizer tests rc=$?"; tail -3 $O/r02k_sanitizer_memcheck_vector_dynamic.log
