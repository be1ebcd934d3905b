#!/bin/bash
# Split-K A/B (MAUA_SPLITK=0/1) at four sizes + parity suites + the tile / split plan of every conv launch.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_plan_gpu.py tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_zz_arch_gpu.py -q -m gpu -x -p no:cacheprovider > $OUT/tests.log 2>&1
echo tests $?; tail -4 $OUT/tests.log
for S in 1024 512 256 2048; do
  for SK in 0 1; do
    MAUA_SPLITK=$SK timeout 120 python bench.py --size $S --steps 30 --warmup 5 --no-cpu-baseline --no-multires --no-extras \
      --profile-out $OUT/prof_${S}_sk$SK.json > $OUT/bench_${S}_sk$SK.json 2> $OUT/bench_${S}_sk$SK.err
    python - $OUT/bench_${S}_sk$SK.json $S $SK <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"size {sys.argv[2]} splitk {sys.argv[3]}:", round(d["value"], 1), "it/s", round(d["ms_per_step"], 3), "ms  conv",
      round(d["roofline"]["achieved"], 1), "TF/s  clk", d["clocks"]["sm_mhz"], d["kernel_breakdown_ms"])
PY
  done
done
MAUA_CONV_DEBUG=1 timeout 100 python bench.py --size 1024 --steps 1 --warmup 3 --history-prefill 0 --no-cpu-baseline --no-multires \
  --no-extras 2>&1 | grep "^conv_tc" | sort | uniq -c | sort -rn | head -40
