#!/bin/bash
# compute-sanitizer over the kernel-level tests (memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory
# hazards in the SIMT kernels).  Slow (10-50x): run on the small-shape tests only.
# Usage (under gpurun): bash tools/gpu_sanitize.sh [tag] [pytest -k expression]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-sanitize}
K=${2:-"conv_first or pool or gram or content_tv_adam or lbfgs"}
mkdir -p $OUT
for tool in memcheck racecheck; do
  timeout -k 5 600 compute-sanitizer --tool $tool --error-exitcode 9 --target-processes all \
    python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "$K" --timeout 500 -p no:cacheprovider > $OUT/$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Race" $OUT/$tool.log | tail -8
done
