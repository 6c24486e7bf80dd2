#!/bin/bash
# Sweep the per-window Schur chunk target (HSLAM_SC_BLOCKS) at the default top/lineariser block target.
cd "$(dirname "$0")/.."
for sc in ${@:-14 21 28 35 42 49 56 70 84}; do HSLAM_SC_BLOCKS=$sc python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-sharded 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('sc=$sc', 'step_ms=%.4f' % d['ms_per_step'], 'e2e_ms=%.3f' % d['e2e']['ms_per_step'], {k: round(v, 1) for k, v in d['roofline']['kernel_avg_us'].items()})
"; done
