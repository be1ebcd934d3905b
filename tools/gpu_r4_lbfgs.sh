#!/bin/bash
# L-BFGS kernels: parity test, bench legs, per-kernel durations at full history (ncu, warm caches)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r04h}; mkdir -p $OUT
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -x -k lbfgs 2>&1 | tail -2 | tee $OUT/tests.txt
for S in 256 512 1024; do
  timeout 150 python bench.py --size $S --steps 40 --warmup 5 --no-cpu-baseline --no-multires --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['image'], round(d['ms_per_step'],4), round(d['value'],1), d['clocks']['sm_mhz'], d.get('lbfgs_state'))" | tee -a $OUT/bench.txt
done
for S in ${2:-256 1024}; do
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:lbfgs -s 440 -c 4 --csv --log-file $OUT/ncu_$S.csv python bench.py --size $S --steps 8 --warmup 3 --no-cpu-baseline --no-multires --no-extras > /dev/null 2>&1
python - $OUT/ncu_$S.csv $S <<'PY' | tee -a $OUT/lbfgs_kernels.txt
import csv, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.reader(rows); hdr = next(rd)
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
print('size', sys.argv[2], '  '.join(f"{r[ki].split('(')[0].split('::')[-1]} {r[vi]} us" for r in rd))
PY
done
