#!/bin/bash
# bench.py at N GPUs the way the driver launches it
cd "$(dirname "$0")/.."
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02m_scale_n$N.json 2> gpurun_out/r02m_scale_n$N.err; echo "rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/r02m_scale_n$N.json").read().strip().split(chr(10))[-1])
print("N=$N", round(d["ms_per_step"],4), round(d["value"]/1e6,1), d["check"]["ok"], "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, d["gpu_launches_per_step"])
for k,v in d.get("also",{}).items(): print("   also", k, round(v["ms_per_step"],4), round(v["value"]/1e6,1), v["check"]["ok"], round(v["e2e"]["ms_per_step"],3))
P
