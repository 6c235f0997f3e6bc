#!/bin/bash
cd "$(dirname "$0")/.."
for pl in 1 0 1 0; do
TAD_PRIORITY_LANE=$pl TAD_COMM_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-extra > gpurun_out/r02o_n8_prio$pl.json 2> gpurun_out/r02o_n8_prio$pl.err
python -c "
import json
d=json.loads(open('gpurun_out/r02o_n8_prio$pl.json').read().strip().split(chr(10))[-1]); print('PRIORITY_LANE=$pl', round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['check']['ok'], 'e2e', round(d['e2e']['ms_per_step'],2))"
grep "comm trace" gpurun_out/r02o_n8_prio$pl.err | grep "slabs 2" | sed -n '80,87p' | cut -c1-230
done
