#!/bin/bash
# Tracking latency per tuning variant in build/variants/ (bench extras only).
cd "$(dirname "$0")/.."
for so in build/variants/libhslam_b200_*.so; do
  name=$(basename $so .so); name=${name#libhslam_b200_}
  HSLAM_B200_LIB=$PWD/$so python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)['extra']
        print('$name', {k: round(v, 4) for k, v in d.items() if k.startswith('tracking') and isinstance(v, float)})
"
done
