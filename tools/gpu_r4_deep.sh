#!/bin/bash
# Deeper operand rings for the narrow conv tiles: kernel tests, the small-shape sweep and the 256^2 / 512^2 / 1024^2 bench legs
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r04a}; mkdir -p $OUT
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_plan_gpu.py -q -x 2>&1 | tail -5
timeout 60 python tools/exp_conv_chain.py 2>/dev/null | tee $OUT/chain.txt
timeout 120 python tools/sweep_conv_small.py 15 2>&1 | tee $OUT/sweep.txt
for S in 256 512 1024; do
  timeout 150 python bench.py --size $S --steps 40 --warmup 5 --no-cpu-baseline --no-multires --no-extras --profile-out $OUT/prof_${S}.json > $OUT/b_${S}.json 2> $OUT/b_${S}.err
  python - $OUT/b_${S}.json $S <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d['kernel_breakdown_ms']
print(f"size {sys.argv[2]}: {d['ms_per_step']:.3f} ms  {d['value']:.1f} it/s  conv fwd {k['conv_fwd']:.4f} dgrad {k['conv_dgrad']:.4f} clk {d['clocks']['sm_mhz']} lbfgs {d.get('lbfgs_state')}")
PY
done
