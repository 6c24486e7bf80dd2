for tb in 5 8 9 10 14 18 28; do HSLAM_BATCH_BLOCKS=$tb python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('tb=$tb', 'step_ms=%.4f' % d['ms_per_step'], 'e2e_ms=%.3f' % d['e2e']['ms_per_step'], {k: round(v, 1) for k, v in d['roofline']['kernel_avg_us'].items()})
"; done
