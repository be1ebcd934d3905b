#!/bin/bash
# loss modules on a side stream: bit-identity test + A/B at 256^2 / 512^2 / 1024^2
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r04e}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_zz_arch_gpu.py -q -x -k "side_stream or fused_pool" 2>&1 | tail -5 | tee $OUT/tests.txt
for S in 256 512 1024; do
  for M in 0 1 0 1; do
    MAUA_SIDE_STREAM=$M timeout 150 python bench.py --size $S --steps 40 --warmup 5 --no-cpu-baseline --no-multires --no-extras > $OUT/b_${S}_$M.json 2> $OUT/b_${S}_$M.err
    python - $OUT/b_${S}_$M.json $S $M <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"size {sys.argv[2]} MAUA_SIDE_STREAM={sys.argv[3]}: {d['ms_per_step']:.3f} ms  {d['value']:.1f} it/s  e2e {d['e2e']['value']:.1f} clk {d['clocks']['sm_mhz']} graph {d.get('cuda_graph')}")
PY
  done
done
