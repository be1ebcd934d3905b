#!/bin/bash
# A/B of conv tail mode 3 (K-split + reduce kernel for launches with fewer tiles than SMs) against the default mode 2
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r03t}; mkdir -p $OUT
MAUA_CONV_TAIL=3 timeout 600 python -m pytest tests/test_plan_gpu.py tests/test_zz_arch_gpu.py tests/test_exact_gpu.py tests/test_fullsize_gpu.py tests/test_video_gpu.py tests/test_image_gpu.py -q -s 2>&1 | grep -E "PSNR|passed|failed|FAILED|assert" | tail -40
for S in 256 512 1024; do
  for M in 2 3 2 3; do
    MAUA_CONV_TAIL=$M timeout 150 python bench.py --size $S --steps 40 --warmup 5 --no-cpu-baseline --no-multires --no-extras --profile-out $OUT/prof_${S}_$M.json > $OUT/b_${S}_$M.json 2> $OUT/b_${S}_$M.err
    python - $OUT/b_${S}_$M.json $S $M <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d['kernel_breakdown_ms']
print(f"size {sys.argv[2]} MAUA_CONV_TAIL={sys.argv[3]}: {d['ms_per_step']:.3f} ms  {d['value']:.1f} it/s  conv fwd {k['conv_fwd']:.4f} dgrad {k['conv_dgrad']:.4f} pool_fwd {k.get('pool_fwd', 0):.4f} clk {d['clocks']['sm_mhz']}")
PY
  done
done
