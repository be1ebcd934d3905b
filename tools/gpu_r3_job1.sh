mkdir -p gpurun_out/r03h
python -m pytest tests/test_plan_gpu.py tests/test_zz_arch_gpu.py tests/test_image_gpu.py tests/test_video_gpu.py tests/test_dropin_gpu.py -q -x 2>&1 | tail -4
python tools/dbg_small_scale.py 256 2>&1 | grep call
MAUA_TRACE=2 python tools/trace_multires.py 2>&1 | grep -E "run|load_model|1\.\.100" | tail -12
python - <<'PY'
import json, torch, sys
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0); torch.cuda.set_device(dev)
pk = bench.peaks()
for name, fn in [("nin_4096_adam", lambda: bench.side_leg(4096, "adam", dev, 5, 2, pk, arch="nin")),
                 ("nin_2048_adam", lambda: bench.side_leg(2048, "adam", dev, 5, 2, pk, arch="nin")),
                 ("pruned_4096_adam", lambda: bench.side_leg(4096, "adam", dev, 5, 2, pk, arch="prune"))]:
    try:
        print(name, json.dumps(fn()))
    except Exception as e:
        print(name, "ERROR", type(e).__name__, str(e)[:300])
PY
