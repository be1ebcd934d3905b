#!/bin/bash
# same-box A/B of the Gram SYRK + finalize: previous library (gpurun_tmp/libmaua_old.so) vs the current one
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
cp maua_style_b200/libmaua_b200.so /tmp/new.so
for rep in 1 2; do
cp gpurun_tmp/libmaua_old.so maua_style_b200/libmaua_b200.so; echo OLD; timeout 60 python tools/bench_gram.py 1024 512 256
cp /tmp/new.so maua_style_b200/libmaua_b200.so; echo NEW; timeout 60 python tools/bench_gram.py 1024 512 256
done
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "gram or style" 2>&1 | tail -1
