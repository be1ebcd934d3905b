import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


GOLDEN_CPU_THREADS = 8  # the thread count of the container in which tests/golden/make_golden*.py ran the reference


def pytest_configure(config):
    # torch's CPU convolutions sum in an order that depends on the number of threads; under max pooling that moves near-tie arg-max
    # decisions, and the fp32 image gradient with them (1e-4 relative between 1 and 8 threads).  The CPU suite compares the oracle
    # with goldens of the reference to 1e-5, so it pins the thread count the goldens were made with.  Not applied on a GPU box:
    # there the CPU oracle is only the checker of the -m gpu tests, whose bounds do not depend on it.
    try:
        import torch

        if not torch.cuda.is_available():
            torch.set_num_threads(GOLDEN_CPU_THREADS)
    except Exception:  # pragma: no cover
        pass
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100a); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "multigpu: needs two B200s on one box (gpurun --gpus 2 -- pytest -m multigpu); not part of -m gpu")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected explicitly with -m gpu; when collected without a GPU they are skipped, not failed.
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords or "multigpu" in item.keywords:
            item.add_marker(skip)
