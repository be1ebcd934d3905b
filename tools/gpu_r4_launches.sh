#!/bin/bash
# ncu launch lists (warm caches: --cache-control none) of one steady-state iteration at 256^2 / 512^2 / 1024^2
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r04f}; mkdir -p $OUT
for S in ${2:-256 512 1024}; do
  mkdir -p $OUT/l$S
  timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv --log-file /tmp/launches_$S.csv \
     python bench.py --size $S --steps 2 --warmup 1 --history-prefill 8 --no-cpu-baseline --no-multires --no-extras > $OUT/ncu_$S.log 2>&1
  python tools/launch_summary.py /tmp/launches_$S.csv $OUT/l$S > $OUT/l$S/summary.txt 2>&1
  tail -45 $OUT/l$S/summary.txt
done
