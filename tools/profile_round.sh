#!/bin/bash
# Round-end evidence on the GPU box.  usage: tools/profile_round.sh <tag> bench|list|full
#   bench : full GPU test suite, default bench (both arms) -> gpurun_out/bench_<tag>.json, bench_ref_<tag>.json
#   list  : ncu launch list (gpu__time_duration) of the short bench -> gpurun_out/launches_<tag>.csv
#   full  : ncu --set full of one launch of each BA kernel -> gpurun_out/prof_ba_<tag>.ncu-rep
cd "$(dirname "$0")/.."
tag=$1; what=$2
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-sharded"
KERN='regex:ba_(linearize|apply|fixlin|top|point_sum|schur|fold_sc|stitch|energy)'   # not the per-window import / permute launches of the set-up
case $what in
  bench)
    timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
    timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.json
    timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; tail -c 400 gpurun_out/bench_ref_$tag.json ;;
  list)
    timeout 300 $SHORT > gpurun_out/short_$tag.log 2>&1 || exit 1
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 1500 --csv \
      --log-file gpurun_out/launches_$tag.csv $SHORT > gpurun_out/ncu_list_$tag.log 2>&1; tail -3 gpurun_out/ncu_list_$tag.log ;;
  full)
    timeout 300 $SHORT > gpurun_out/short_$tag.log 2>&1 || exit 1
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"ba_(linearize|top|schur|fold_sc|stitch)" \
      --launch-skip 101 --launch-count 6 -f -o gpurun_out/prof_ba_$tag $SHORT > gpurun_out/ncu_full_$tag.log 2>&1; tail -3 gpurun_out/ncu_full_$tag.log ;;
esac
