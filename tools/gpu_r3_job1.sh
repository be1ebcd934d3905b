#!/bin/bash
# second-session GPU check: graph-capture path, multi-res trace, NIN / pruned legs at large sizes
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r03h}; mkdir -p $OUT
python -m pytest tests/test_plan_gpu.py tests/test_zz_arch_gpu.py tests/test_image_gpu.py tests/test_video_gpu.py tests/test_dropin_gpu.py tests/test_kernels_gpu.py -q -x -k "not gram" 2>&1 | tail -4
python tools/dbg_small_scale.py 256 2>&1 | grep call
MAUA_TRACE=2 python tools/trace_multires.py > $OUT/multires_trace2.txt 2>&1; grep -E "run|load_model|1\.\.100|rror" $OUT/multires_trace2.txt | tail -16
python - > $OUT/arch_legs.jsonl <<'PY'
import json, torch, sys
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0); torch.cuda.set_device(dev)
pk = bench.peaks()
for name, fn in [("nin_4096_adam", lambda: bench.side_leg(4096, "adam", dev, 5, 2, pk, arch="nin")),
                 ("nin_2048_adam", lambda: bench.side_leg(2048, "adam", dev, 5, 2, pk, arch="nin")),
                 ("nin_5312_adam", lambda: bench.side_leg(5312, "adam", dev, 3, 2, pk, arch="nin")),
                 ("pruned_4096_adam", lambda: bench.side_leg(4096, "adam", dev, 5, 2, pk, arch="prune"))]:
    try:
        print(json.dumps({name: fn()}), flush=True)
    except Exception as e:
        print(json.dumps({name: {"error": f"{type(e).__name__}: {str(e)[:300]}"}}), flush=True)
PY
cut -c1-1000 $OUT/arch_legs.jsonl
