#!/bin/bash
# Run bench.py once per tuning variant in build/variants/ and print per-kernel average launch times.
# usage: tools/sweep.sh [extra bench args]
cd "$(dirname "$0")/.."
for so in build/variants/libhslam_b200_*.so; do
  name=$(basename $so .so); name=${name#libhslam_b200_}
  HSLAM_B200_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$name', 'step_ms=%.4f' % d['ms_per_step'], 'e2e_ms=%.3f' % d['e2e']['ms_per_step'], 'frac=%.3f' % d['roofline']['frac'], {k: round(v, 1) for k, v in d['roofline']['kernel_avg_us'].items()})
"
done
