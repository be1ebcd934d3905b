#!/bin/bash
# Round-2 (fourth session) GPU visits: the video drivers around the path and the standalone loss modules.
#   bash tools/gpu_r5_video.sh [tag] [phases: v l t b]
#     v  tests/test_vid_driver_gpu.py with the PSNR lines   -> profiles/r05_vid_driver.txt
#     l  tests/test_loss_modules_gpu.py                      -> profiles/r05_loss_modules.txt
#     t  the whole GPU suite (tools/gpu_r4_final.sh <tag> t) -> profiles/r05_gpu_tests.txt, r05_parity.txt
#     b  the measured video job, stand-alone and as the bench leg -> profiles/r05_bench_video*.json
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r05}; PH=${2:-vlb}
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [[ $PH == *v* ]]; then
  timeout 150 python -m pytest tests/test_vid_driver_gpu.py -q -s -p no:cacheprovider 2>&1 | grep -a "PSNR\|passed\|failed" | tee $OUT/vid_driver.txt
fi
if [[ $PH == *l* ]]; then
  timeout 100 python -m pytest tests/test_loss_modules_gpu.py -q -p no:cacheprovider 2>&1 | tail -5 | tee $OUT/loss_modules.txt
fi
if [[ $PH == *t* ]]; then
  bash tools/gpu_r4_final.sh $TAG t
fi
if [[ $PH == *b* ]]; then
  timeout 95 python tools/bench_video.py 2> $OUT/bench_video.err | tee $OUT/bench_video.json
  timeout 95 python bench.py --video-leg 2> $OUT/video_leg.err | tee $OUT/video_leg.json
fi
