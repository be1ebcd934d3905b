#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r02s}; mkdir -p $OUT
for S in 256 512 1024; do
  for M in 0 25 29 31 27 0 25 29; do
    MAUA_PDL=$M timeout 150 python bench.py --size $S --steps 40 --warmup 5 --no-cpu-baseline --no-multires --no-extras > $OUT/b_${S}_$M.json 2> $OUT/b_${S}_$M.err
    python - $OUT/b_${S}_$M.json $S $M <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"size {sys.argv[2]} MAUA_PDL={sys.argv[3]:>2}: {d['ms_per_step']:.3f} ms  {d['value']:.1f} it/s  clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
  done
done
