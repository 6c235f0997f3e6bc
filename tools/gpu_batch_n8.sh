#!/bin/bash
# dev batch on 8 GPUs: C5 strong with several slab sizes (TAD_CHUNK_ELEMENTS), default last with the C2 weak line riding along
cd "$(dirname "$0")/.."
O=gpurun_out
run() { # $1 = label, $2 = chunk env (empty = default), rest = extra flags
  local label=$1 chunk=$2; shift 2
  env ${chunk:+TAD_CHUNK_ELEMENTS=$chunk} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 "$@" > $O/r02i_n8_$label.json 2> $O/r02i_n8_$label.err
  python - <<P
import json
d=json.loads(open("$O/r02i_n8_$label.json").read().strip().split(chr(10))[-1])
print("$label", d["ms_per_step"], d["value"], d["check"]["ok"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches_per_step"])
for k,v in d.get("also",{}).items(): print("   also", k, v["ms_per_step"], v["value"], v["check"]["ok"], v["e2e"]["ms_per_step"])
P
}
run chunk640k 655360 --no-extra
run chunk1280k 1310720 --no-extra
run chunk320k 327680 --no-extra
run default "" 
