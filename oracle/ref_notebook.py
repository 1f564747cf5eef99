"""Mint golden vectors for the multi-scale conditioning pyramid (row f-4) from the reference's OWN notebook cells.

TEST INFRASTRUCTURE, NOT PRODUCT. Runs only in the build container. ``improv_experiments.ipynb`` cell 3 defines
``soft_fuse`` and cell 5 is the multi-scale loop (resize frames and flows, soft-splat both frames, fuse). Exactly that
source text is executed here: ``soft_fuse`` is lifted out of cell 3 with ``ast``; cell 5 runs unmodified with
``softsplat`` bound to the reference's own ``controlnet/softsplat.py`` (its kernel text compiled for the CPU by
``oracle/ref_emulation.py``), ``device = 'cpu'`` and ``display_fusion_results`` bound to a recorder instead of matplotlib.
Nothing of the reference is copied into the repository; only seeded inputs and the recorded outputs are stored, as
``tests/golden/ref_notebook_pyramid.npz``.

Usage:  python oracle/ref_notebook.py
"""
from __future__ import annotations

import ast
import json
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
_ROOT = os.environ.get("DCB_REFERENCE_ROOT", "/root/reference")
_GOLD = os.path.join(os.path.dirname(_HERE), "tests", "golden", "ref_notebook_pyramid.npz")

SIZES = [48, 24, 12]           # cell 4 has target_sizes = [128, 64, 32] on 512 x 512 frames: the same 4 : 2 : 1 ladder


def make_inputs(seed: int = 5, res: int = 96):
    """Two frames and two roughly inverse smooth flows at `res` x `res` (the notebook loads 512 x 512 UVG frames)."""
    g = torch.Generator().manual_seed(seed)
    img1 = torch.randint(0, 256, (1, 3, res, res), generator=g).float() / 255.0        # 8-bit frames: stored as uint8 in the fixture
    img2 = torch.randint(0, 256, (1, 3, res, res), generator=g).float() / 255.0
    coarse = torch.randn(1, 2, res // 16, res // 16, generator=g) * 6.0
    flow1 = F.interpolate(coarse, size=(res, res), mode="bicubic", align_corners=False)
    flow2 = -flow1 + 0.5 * F.interpolate(torch.randn(1, 2, res // 8, res // 8, generator=g), size=(res, res), mode="bilinear", align_corners=False)
    return img1, img2, flow1, flow2


def notebook_cells():
    nb = json.load(open(os.path.join(_ROOT, "improv_experiments.ipynb")))
    cells = ["".join(c["source"]) for c in nb["cells"]]
    fuse_cell = next(c for c in cells if "def soft_fuse" in c)
    loop_cell = next(c for c in cells if c.lstrip().startswith("for size in target_sizes"))
    fuse_def = next(n for n in ast.parse(fuse_cell).body if isinstance(n, ast.FunctionDef) and n.name == "soft_fuse")
    return ast.Module(body=[fuse_def], type_ignores=[]), loop_cell


def run_reference(img1, img2, flow1, flow2):
    from oracle import ref_emulation
    ref = ref_emulation.load_reference("fast")
    fuse_mod, loop_src = notebook_cells()
    recorded = []
    ns = {"torch": torch, "F": F, "np": np, "softsplat": ref.softsplat, "device": "cpu", "dtype": torch.float32,
          "img1": img1, "img2": img2, "flow1": flow1, "flow2": flow2, "target_sizes": list(SIZES),
          "display_fusion_results": lambda a, b, c, main_title="": recorded.append((a.clone(), b.clone(), c.clone())),
          "print": lambda *a, **k: None}
    exec(compile(fuse_mod, "improv_experiments.ipynb:cell3", "exec"), ns)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with ref_emulation.pretend_cuda():
            exec(compile(loop_src, "improv_experiments.ipynb:cell5", "exec"), ns)
    assert len(recorded) == len(SIZES)
    return recorded


def main():
    img1, img2, flow1, flow2 = make_inputs()
    got = run_reference(img1, img2, flow1, flow2)
    blob = {"img1_u8": (img1 * 255.0).round().to(torch.uint8).numpy(), "img2_u8": (img2 * 255.0).round().to(torch.uint8).numpy(), "flow1": flow1.numpy(), "flow2": flow2.numpy(), "sizes": np.asarray(SIZES)}
    for size, (w1, w2, fused) in zip(SIZES, got):
        blob[f"warped1_{size}"], blob[f"warped2_{size}"], blob[f"fused_{size}"] = w1.numpy(), w2.numpy(), fused.numpy()
        print(size, tuple(fused.shape), float(fused.abs().mean()))
    np.savez_compressed(_GOLD, **blob)
    print("wrote", _GOLD, f"{os.path.getsize(_GOLD) / 1024:.0f} KB")


if __name__ == "__main__":
    sys.exit(main())
