#!/bin/bash
# Generic A/B of one environment switch: bash tools/gpu_r2_ab.sh <tag> <ENVVAR> [sizes...]   (values 0 and 1)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=$1; VAR=$2; shift 2
SIZES=${@:-"1024 512 256"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $OUT/tests.log 2>&1
echo tests $?; tail -3 $OUT/tests.log
for S in $SIZES; do
  for V in 0 1; do
    env $VAR=$V timeout 150 python bench.py --size $S --steps 40 --warmup 5 --no-cpu-baseline --no-multires --no-extras \
      --profile-out $OUT/prof_${S}_$V.json > $OUT/bench_${S}_$V.json 2> $OUT/bench_${S}_$V.err
    python - $OUT/bench_${S}_$V.json $S $VAR $V <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"size {sys.argv[2]} {sys.argv[3]}={sys.argv[4]}:", round(d["value"], 1), "it/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1),
          " conv", round(d["roofline"]["achieved"], 1), "TF/s  clk", d["clocks"]["sm_mhz"], "graph", d["cuda_graph"], d["kernel_breakdown_ms"])
except Exception as e:
    print("parse failed", sys.argv[1], e)
PY
    tail -2 $OUT/bench_${S}_$V.err | cut -c1-300
  done
done
