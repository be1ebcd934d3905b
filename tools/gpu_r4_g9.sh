#!/bin/bash
# nine-tap geometry: kernel / plan tests, sweeps with the forced shape, bench legs
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r04d}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_plan_gpu.py -q -x 2>&1 | tail -15 | tee $OUT/tests.txt
SWEEP_BIG=1 SWEEP_TOP=8 timeout 120 python tools/sweep_conv_small.py 7 2>&1 | tee $OUT/sweep_big.txt
SWEEP_TOP=8 timeout 120 python tools/sweep_conv_small.py 9 2>&1 | tee $OUT/sweep.txt
for S in 256 512 1024; do
  timeout 150 python bench.py --size $S --steps 40 --warmup 5 --no-cpu-baseline --no-multires --no-extras --profile-out $OUT/prof_${S}.json > $OUT/b_${S}.json 2> $OUT/b_${S}.err
  python - $OUT/b_${S}.json $S <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d['kernel_breakdown_ms']
print(f"size {sys.argv[2]}: {d['ms_per_step']:.3f} ms  {d['value']:.1f} it/s  conv fwd {k['conv_fwd']:.4f} dgrad {k['conv_dgrad']:.4f} clk {d['clocks']['sm_mhz']} roofline {d['roofline']['achieved']:.0f} TF/s")
PY
done
