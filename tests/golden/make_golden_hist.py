"""Golden vectors for histogram matching, produced by the reference's own function:

    python tests/golden/make_golden_hist.py   ->  tests/golden/hist_match.npz

`utils.match_histogram` (utils.py:88-151) is imported UNMODIFIED from /root/reference.  It calls the third-party
`torch.symeig`, which the pinned torch 1.8.1 has and torch >= 1.9 removed (the call raises, the reference catches the
RuntimeError and silently skips the matching).  To run the function as its author's environment did, the removed torch
entry point is re-provided for the duration of this script as what its documentation defined it to be:
`symeig(A, eigenvectors=True, upper=True)` = eigenvalues ascending + eigenvectors of the symmetric matrix read from
its upper triangle = `torch.linalg.eigh(A, UPLO="U")`.  The reference code itself is not touched.

The function adds `1e-3 * randn` noise to both images before taking their statistics and to the image it transforms
(utils.py:120-121), seeded here with torch.manual_seed; the noise-free transform differs from these outputs by about
1e-3 * |M| per pixel, which is what the tolerances in the tests allow for.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference  # noqa: E402

CASES = [  # (name, target hw, [source hw ...], mode, seed)
    ("one_source", (40, 56), [(30, 44)], True, 1),
    ("two_sources", (37, 53), [(48, 32), (25, 61)], True, 2),
    ("avg_mode", (32, 32), [(40, 24)], "avg", 3),
    ("flat_target", (24, 40), [(32, 32)], True, 4),
]


def photo_like(h, w, seed, gain=(1.0, 1.0, 1.0), offset=(0.0, 0.0, 0.0)):
    """Smooth, channel-correlated image in the pipeline's value range (BGR, 0-255, mean-subtracted)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 3, max(h // 6, 2), max(w // 6, 2), generator=g)
    mix = torch.tensor([[0.7, 0.2, 0.1], [0.25, 0.6, 0.15], [0.1, 0.3, 0.6]])
    low = torch.einsum("dc,bchw->bdhw", mix, low)
    img = F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False) * 255
    img = img + torch.randn(1, 3, h, w, generator=g) * 4
    img = img * torch.tensor(gain).view(1, 3, 1, 1) + torch.tensor(offset).view(1, 3, 1, 1)
    return (img - torch.tensor([103.939, 116.779, 123.68]).view(1, 3, 1, 1)).contiguous()


def case_inputs(name, thw, shws, seed):
    if name == "flat_target":  # nearly constant target: Ct ~ eps * I, the transform is dominated by eps
        t = photo_like(*thw, seed=100 + seed) * 0.002 + 17.0
    else:
        t = photo_like(*thw, seed=100 + seed, gain=(0.6, 0.8, 0.5), offset=(20.0, -10.0, 5.0))
    ss = [photo_like(h, w, seed=200 + 10 * seed + i, gain=(1.1, 0.7, 0.9), offset=(-15.0 * i, 8.0, 12.0)) for i, (h, w) in enumerate(shws)]
    return t, ss


def main():
    import_reference()
    import utils as rutils  # noqa  (reference utils.py)

    if True:  # the removed third-party entry point, see the module docstring
        def symeig(A, eigenvectors=False, upper=True):
            vals, vecs = torch.linalg.eigh(A, UPLO="U" if upper else "L")
            return vals, vecs

        torch.symeig = symeig
    out = {"cases": json.dumps(CASES)}
    for name, thw, shws, mode, seed in CASES:
        t, ss = case_inputs(name, thw, shws, seed)
        torch.manual_seed(seed)
        np.random.seed(seed)
        res = rutils.match_histogram(t.clone(), [s.clone() for s in ss], mode=mode)
        assert not torch.equal(res, t), "the reference skipped the matching (symeig shim not in effect?)"
        out[f"{name}_target"] = t.numpy()
        for i, src in enumerate(ss):
            out[f"{name}_source{i}"] = src.numpy()
        out[f"{name}_out"] = res.numpy()
        print(name, tuple(res.shape), "mean", res.mean(dim=(0, 2, 3)).tolist())
    np.savez_compressed(HERE / "hist_match.npz", **out)


if __name__ == "__main__":
    main()
