#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r02y}; mkdir -p $OUT
for S in 1024 512 256; do
  for M in 0 2 0 2; do
    MAUA_CONV_TAIL=$M timeout 150 python bench.py --size $S --steps 40 --warmup 5 --no-cpu-baseline --no-multires --no-extras --profile-out $OUT/prof_${S}_$M.json > $OUT/b_${S}_$M.json 2> $OUT/b_${S}_$M.err
    python - $OUT/b_${S}_$M.json $S $M <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d['kernel_breakdown_ms']
print(f"size {sys.argv[2]} MAUA_CONV_TAIL={sys.argv[3]}: {d['ms_per_step']:.3f} ms  {d['value']:.1f} it/s  conv fwd {k['conv_fwd']:.4f} dgrad {k['conv_dgrad']:.4f}  clk {d['clocks']['sm_mhz']}")
PY
  done
done
