#!/bin/bash
cd "$(dirname "$0")/.."
N=$1
TAD_COMM_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 --no-extra > gpurun_out/r02n_trace_n$N.json 2> gpurun_out/r02n_trace_n$N.err
grep "comm trace" gpurun_out/r02n_trace_n$N.err | sed -n '10,19p'
python -c "
import json
d=json.loads(open('gpurun_out/r02n_trace_n$N.json').read().strip().split(chr(10))[-1]); print('N=$N', d['ms_per_step'], d['roofline']['kernel_ms'], d['check']['ok'])"
