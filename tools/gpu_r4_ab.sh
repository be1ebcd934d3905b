#!/bin/bash
# A/B of one environment switch (values 0 / 1) at 256^2 / 512^2 / 1024^2:   bash tools/gpu_r4_ab.sh TAG ENVVAR [pytest -k expr]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r04g}; mkdir -p $OUT
VAR=${2:-MAUA_PREFETCH_W}
V0=${4:-0}; V1=${5:-1}
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_plan_gpu.py tests/test_zz_arch_gpu.py -q -x ${3:+-k "$3"} 2>&1 | tail -4 | tee $OUT/tests.txt
for S in 256 512 1024; do
  for M in $V0 $V1 $V0 $V1; do
    env $VAR=$M timeout 150 python bench.py --size $S --steps 40 --warmup 5 --no-cpu-baseline --no-multires --no-extras > $OUT/b_${S}_$M.json 2> $OUT/b_${S}_$M.err
    python - $OUT/b_${S}_$M.json $S $VAR=$M <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"size {sys.argv[2]} {sys.argv[3]}: {d['ms_per_step']:.3f} ms  {d['value']:.1f} it/s  e2e {d['e2e']['value']:.1f} clk {d['clocks']['sm_mhz']} graph {d.get('cuda_graph')} lbfgs {d.get('lbfgs_state', {}).get('history_len')}")
PY
  done
done
