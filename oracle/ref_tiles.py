"""Mint golden vectors for the latent tile merge from the reference's OWN function.

TEST INFRASTRUCTURE, NOT PRODUCT. Runs only in the build container: it reads
``/root/reference/patch_utils.py`` (which cannot be imported as a module here -- it needs cv2,
PIL and ``test_utils``), lifts the two function definitions it needs out of the file with ``ast``
and executes exactly that source text with ``torch`` / ``F`` / ``np`` in scope. Nothing of the
reference is copied into the repository; only inputs and outputs are stored, as
``tests/golden/ref_tiles_*.npz``.

Usage:  python oracle/ref_tiles.py
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(os.environ.get("DCB_REFERENCE_ROOT", "/root/reference"), "patch_utils.py")
_GOLD = os.path.join(os.path.dirname(_HERE), "tests", "golden")


def reference_functions():
    src = open(_REF).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": F, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("merge_latent_tiles_from_pixel_coords", "crop_into_tiles"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), _REF, "exec"), ns)
    return ns["merge_latent_tiles_from_pixel_coords"], ns["crop_into_tiles"]


def cases():
    """name -> (latents, pixel_coords, full_latent_shape, original_image_size)."""
    g = torch.Generator().manual_seed(1234)
    out = {}
    # (a) 1080p frame cut into 512^2 tiles with 64 px overlap (patch_exp.ipynb cell 3), latents 8x smaller.
    #     The coordinate tuples go in exactly as crop_into_tiles() returns them.
    _, crop = reference_functions()
    img = np.zeros((4, 1080, 1920), np.float32)
    _, coords, _ = crop(img, (512, 512), overlap=64, order="chw")
    lat = [torch.randn(1, 4, (y2 - y) // 8, (x2 - x) // 8, generator=g) for (y, y2, x, x2) in coords]
    out["uvg1080p_lat8"] = (lat, coords, (1, 4, 135, 240), (1080, 1920))
    # (b) square canvas, tuples in the order the function unpacks them (x1, x2, y1, y2): exact fits, no resize
    coords_b = [(x, min(x + 96, 256), y, min(y + 96, 256)) for y in range(0, 256, 80) for x in range(0, 256, 80)]
    lat_b = [torch.randn(1, 3, (c[3] - c[2]) // 4, (c[1] - c[0]) // 4, generator=g) for c in coords_b]
    out["square_fit"] = (lat_b, coords_b, (1, 3, 64, 64), (256, 256))
    # (c) odd sizes: rounding to even, 1-pixel rectangles, a tile that lands outside the canvas, resizes
    coords_c = [(0, 50, 0, 30), (45, 100, 0, 31), (0, 49, 25, 60), (40, 100, 29, 60), (98, 100, 58, 60), (100, 130, 0, 10), (10, 11, 10, 11)]
    shapes_c = [(7, 9), (8, 14), (9, 12), (5, 5), (2, 2), (3, 3), (1, 1)]
    lat_c = [torch.randn(1, 5, h, w, generator=g) for (h, w) in shapes_c]
    out["ragged"] = (lat_c, coords_c, (1, 5, 15, 25), (60, 100))
    return out


def main():
    merge, _ = reference_functions()
    os.makedirs(_GOLD, exist_ok=True)
    for name, (lat, coords, full, size) in cases().items():
        got = merge([t.clone() for t in lat], coords, full, size)
        blob = {"out": got.numpy(), "coords": np.asarray(coords, np.int64), "full": np.asarray(full, np.int64),
                "size": np.asarray(size, np.int64)}
        for i, t in enumerate(lat):
            blob[f"tile_{i:03d}"] = t.numpy()
        path = os.path.join(_GOLD, f"ref_tiles_{name}.npz")
        np.savez_compressed(path, **blob)
        print(name, tuple(got.shape), f"{os.path.getsize(path) / 1024:.0f} KB", "zeros:", int((got == 0).sum()))


if __name__ == "__main__":
    sys.exit(main())
