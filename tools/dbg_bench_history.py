"""Does bench.py's Job really run at full L-BFGS history?  History length / halted flag after the prefill, per size, and the
per-iteration time of the same Job next to an optimize_device loop on the same inputs."""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from maua_style_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
for size in (256, 512, 1024):
    job = bench.Job(size, "lbfgs", dev, 0)
    n, h, halted = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.maua_lbfgs_query(job.opt._state, C.byref(n), C.byref(h), C.byref(halted), _lib.stream_ptr()))
    ms = job.timed(40, 5) / 40
    n2, h2, halted2 = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.maua_lbfgs_query(job.opt._state, C.byref(n2), C.byref(h2), C.byref(halted2), _lib.stream_ptr()))
    print(f"size {size}: after prefill n_iter {n.value} history {h.value} halted {halted.value}; {ms:.3f} ms/iteration; "
          f"after timing n_iter {n2.value} history {h2.value} halted {halted2.value}", flush=True)
    job.close()
    del job
    torch.cuda.empty_cache()
