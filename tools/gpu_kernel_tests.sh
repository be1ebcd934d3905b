#!/bin/bash
# Run the per-kernel GPU parity tests group by group, each in its own process with a hard timeout, so that a
# hung tcgen05 kernel only loses its own group.  Logs go to gpurun_out/.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, -k expression
  timeout -k 10 420 python -m pytest tests/test_kernels_gpu.py -q -k "$2" --timeout 180 -p no:cacheprovider > "gpurun_out/kt_$1.log" 2>&1
  echo "== $1: exit $? =="; tail -n 6 "gpurun_out/kt_$1.log"
}
run ref "ref or first or pool or style_loss or content_tv or lbfgs"
run tc_fwd "test_conv3x3_fwd and tc"
run tc_dgrad "test_conv3x3_dgrad and tc"
run tc_aux "test_dgrad_with_style and tc"
run tc_gram "test_gram and tc"
