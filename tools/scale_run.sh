#!/bin/bash
# Multi-GPU bench lines on one box.  usage: tools/scale_run.sh <tag> N [N ...]   (every run under its own timeout)
cd "$(dirname "$0")/.."
tag=$1; shift
for N in "$@"; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29610 + N)) \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> gpurun_out/scale_${tag}_${N}.err | grep '^{' | tail -1 > gpurun_out/scale_${tag}_${N}.json
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/scale_${tag}_${N}.json"))
    print("N=$N value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "sharded", {k: d.get("sharded", {}).get(k) for k in ("ms_per_iteration", "speedup_vs_1gpu", "sharded_check_rel_err")})
except Exception as e:
    print("N=$N failed:", e)
P
done
