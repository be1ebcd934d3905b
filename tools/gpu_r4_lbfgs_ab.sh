#!/bin/bash
# same-box A/B of the L-BFGS step: previous library (gpurun_tmp/libmaua_old.so) vs the current one, automatic regime and forced
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
cp maua_style_b200/libmaua_b200.so /tmp/new.so
SZ="${1:-256 320 384 512 1024}"
cp gpurun_tmp/libmaua_old.so maua_style_b200/libmaua_b200.so; echo OLD; timeout 120 python tools/bench_lbfgs.py $SZ
cp /tmp/new.so maua_style_b200/libmaua_b200.so
echo NEW auto; timeout 120 python tools/bench_lbfgs.py $SZ
echo NEW small=1; MAUA_LBFGS_SMALL=1 timeout 120 python tools/bench_lbfgs.py $SZ
echo NEW small=0; MAUA_LBFGS_SMALL=0 timeout 120 python tools/bench_lbfgs.py $SZ
timeout 100 python -m pytest tests/test_kernels_gpu.py -q -x -k lbfgs 2>&1 | tail -1
