"""Where does a short conv launch spend its time?  Back-to-back launches (host overhead hidden behind the queue, operands
L2-warm like inside an iteration) of the deep-layer shapes of the small scales with parts of the pipeline switched off
(MAUA_CONV_DBG: 1 no MMAs, 2 no activation loads, 4 no weight loads; results are garbage, only the time matters).
  python tools/exp_conv_chain.py [reps]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from maua_style_b200 import _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lib = _lib.load()
_lib.require_gpu()
SHAPES = [tuple(int(v) for v in sh.split(",")) for sh in os.environ["EXP_SHAPES"].split(";")] if os.environ.get("EXP_SHAPES") else [(512, 512, 16, 16), (512, 512, 32, 32), (256, 512, 32, 32), (256, 256, 64, 64), (128, 128, 128, 128), (64, 64, 256, 256),
          (512, 512, 64, 64)]
MODES = [int(m) for m in os.environ.get("EXP_MODES", "0").split(",")]  # non-zero modes need a build with the MAUA_CONV_DBG knob (see git history)
for cin, cout, h, w in SHAPES:
    x = torch.randn(1, h, w, cin, device="cuda")
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    b = torch.randn(cout, device="cuda")
    wg = torch.empty(cout, 9 * cin, device="cuda")
    _lib.check(lib.maua_prep_conv_weights(_lib.ptr(wt), _lib.ptr(wg), cout, cin, 0, _lib.stream_ptr()))
    y = torch.empty(1, h, w, cout, device="cuda")
    out = []
    for force in (None,) + tuple(os.environ.get("EXP_FORCE", "").split(";")) if os.environ.get("EXP_FORCE") else (None,):
        if force:
            os.environ["MAUA_CONV_FORCE"] = force
        else:
            os.environ.pop("MAUA_CONV_FORCE", None)
        for mode in MODES:
            os.environ["MAUA_CONV_DBG"] = str(mode)
            for timed in (0, 1):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    _lib.check(lib.maua_conv3x3_fwd(_lib.ptr(x), _lib.ptr(wg), _lib.ptr(b), _lib.ptr(y), 1, h, w, cin, cout, 1, 0, _lib.stream_ptr()))
                e1.record()
                e1.synchronize()
            out.append(f"{force or 'auto'} dbg{mode}: {e0.elapsed_time(e1) * 1e3 / reps:.1f}")
            print(out[-1], file=sys.stderr, flush=True)
    print(f"conv {cin}->{cout} {h}x{w}: " + "  ".join(out), flush=True)
os.environ.pop("MAUA_CONV_DBG", None)
