#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ... : bench once per value of an environment tuning variable
cd "$(dirname "$0")/.."
var=$1; shift
for v in "$@"; do env $var=$v $EXTRA_ENV python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-sharded 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$var=$v', 'step_ms=%.4f' % d['ms_per_step'], 'e2e_ms=%.3f' % d['e2e']['ms_per_step'], {k: round(v, 1) for k, v in d['roofline']['kernel_avg_us'].items()})
"; done
