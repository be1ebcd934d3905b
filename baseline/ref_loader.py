"""Locate, install and import the UNMODIFIED reference (JCBrouwer/maua-style) for the reference arm of bench.py and for
the drop-in tests.

The reference is a directory of Python scripts (no package, no setup.py), so `pip install --target baseline/_ref
/root/reference` has nothing to install; `install()` instead copies the scripts of the hot path and their JSON presets,
byte for byte, from the read-only mount into the git-ignored `baseline/_ref/` (the driver's build step runs it where
`/root/reference` exists; the directory then travels to the GPU box with the repo snapshot).  Nothing under
`baseline/_ref/` is tracked, and nothing in `maua_style_b200/` imports it.

    import_reference("stock")  ->  the reference's own loss / models / optim (torch CPU or cuDNN arithmetic)
    import_reference("b200")   ->  the reference's style.py / config.py / load.py / utils.py on top of
                                   maua_style_b200.{loss,models,optim}: exactly the sys.modules shim of INTEGRATION.md
                                   section 1
"""
from __future__ import annotations

import os
import shutil
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
SOURCE = Path(os.environ.get("MAUA_REF", "/root/reference"))
FILES = ["loss.py", "models.py", "optim.py", "config.py", "style.py", "load.py", "utils.py"]
BARE_MODULES = ["loss", "models", "optim", "config", "style", "load", "utils", "flow"]
STUBS = ["gdown", "skvideo", "skvideo.io", "ffmpeg", "flow"]  # optional imports the hot path never calls


def install(force: bool = False) -> Path | None:
    """Copy the reference's hot-path scripts + config presets into baseline/_ref (no-op without the source mount)."""
    if not (SOURCE / "optim.py").exists():
        return REF_DIR if (REF_DIR / "optim.py").exists() else None
    REF_DIR.mkdir(parents=True, exist_ok=True)
    for name in FILES:
        dst = REF_DIR / name
        if force or not dst.exists() or dst.read_bytes() != (SOURCE / name).read_bytes():
            shutil.copyfile(SOURCE / name, dst)
    cfg = REF_DIR / "config"
    cfg.mkdir(exist_ok=True)
    for f in (SOURCE / "config").glob("*.json"):
        if force or not (cfg / f.name).exists():
            shutil.copyfile(f, cfg / f.name)
    return REF_DIR


def ref_dir() -> Path | None:
    if (REF_DIR / "optim.py").exists():
        return REF_DIR
    if (SOURCE / "optim.py").exists():
        return SOURCE
    return None


def unload() -> None:
    """Forget every bare-name module of the reference (and the backend aliases) so that another flavour can be imported."""
    for name in BARE_MODULES + STUBS:
        sys.modules.pop(name, None)
    for p in list(sys.path):
        if Path(p) in (REF_DIR, SOURCE):
            sys.path.remove(p)


def import_reference(backend: str = "stock") -> types.SimpleNamespace:
    d = ref_dir()
    if d is None:
        raise FileNotFoundError("the reference is neither installed under baseline/_ref nor mounted at " + str(SOURCE))
    unload()
    for name in STUBS:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["skvideo"].io = sys.modules["skvideo.io"]
    if backend == "b200":
        import maua_style_b200.loss as _loss
        import maua_style_b200.models as _models
        import maua_style_b200.optim as _optim

        sys.modules["loss"], sys.modules["models"], sys.modules["optim"] = _loss, _models, _optim
    elif backend != "stock":
        raise ValueError(f"unknown backend {backend!r}")
    sys.path.insert(0, str(d))
    import config as rconfig  # noqa: E402
    import load as rload  # noqa: E402
    import loss as rloss  # noqa: E402
    import models as rmodels  # noqa: E402
    import optim as roptim  # noqa: E402
    import style as rstyle  # noqa: E402
    import utils as rutils  # noqa: E402

    return types.SimpleNamespace(config=rconfig, load=rload, loss=rloss, models=rmodels, optim=roptim, style=rstyle,
                                 utils=rutils, dir=d, backend=backend)


def reference_args(ref, workdir: Path, ckpt: Path, gpu: str = "c", **over):
    """args as `python style.py --load_args config/args-img.json ...` builds them (config.py:10-131 + postprocess), with a
    scaling JSON that pins the given checkpoint / optimizer for every size (the stock one names downloadable models)."""
    import argparse
    import json

    args = argparse.Namespace()
    with open(Path(ref.dir) / "config" / "args-img.json") as f:
        args.__dict__ = json.load(f)
    optimizer = over.pop("optimizer", "adam")
    scaling = Path(workdir) / "scaling.json"
    scaling.write_text(json.dumps({"100000": {"model_file": str(ckpt), "optimizer": optimizer, "multidevice": False, "gpu": gpu}}))
    args.__dict__.update(dict(gpu=gpu, backend="mkl" if gpu == "c" else "cudnn", scaling_args=str(scaling), model_file=str(ckpt),
                              disable_check=True, no_hist_match=True, content_weight=5.0, optimizer=optimizer,
                              style=["s"] * over.pop("n_styles", 1), content="c", output_dir=str(workdir)))
    args.__dict__.update(over)
    return ref.config.postprocess(args)
