#!/bin/bash
# un-pooling from arg-max codes: bit-identity tests, plan / arch / fullsize tests, A/B at 512^2 / 1024^2 / 2048^2
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r04p}; mkdir -p $OUT
timeout 500 python -m pytest tests/test_zz_arch_gpu.py tests/test_plan_gpu.py tests/test_fullsize_gpu.py tests/test_video_gpu.py -q -x 2>&1 | tail -3 | tee $OUT/tests.txt
for S in 512 1024 2048; do
  for M in 0 1 0 1; do
    MAUA_POOL_CODES=$M timeout 150 python bench.py --size $S --steps 20 --warmup 5 --no-cpu-baseline --no-multires --no-extras > $OUT/b_${S}_$M.json 2> $OUT/b_${S}_$M.err
    python - $OUT/b_${S}_$M.json $S $M <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"size {sys.argv[2]} MAUA_POOL_CODES={sys.argv[3]}: {d['ms_per_step']:.3f} ms  {d['value']:.1f} it/s  pool_bwd {d['kernel_breakdown_ms'].get('pool_bwd')} clk {d['clocks']['sm_mhz']}")
PY
  done
done
