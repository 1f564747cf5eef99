"""Install this package in place of the reference's hot-path modules.

After ``install()``, ``from controlnet.softsplat import softsplat`` and
``from controlnet.control_utils import compute_mask, FeatureWarperSoftsplat, ...`` and
``from controlnet.extractors import Bi_Dir_FeatureExtractor`` (the imports used by the reference's
``extractors.py:4``, ``dataset.py:11-12``, ``residual_utils.py:10-13``, ``flownet.py:8``) resolve to the
B200 implementations; the rest of the pipeline is untouched.
"""
from __future__ import annotations

import importlib
import sys
import types

__all__ = ["install"]


def install(force: bool = True) -> None:
    # the package re-exports the function `softsplat` under the submodule's name: go through importlib
    softsplat = importlib.import_module(__package__ + ".softsplat")
    control_utils = importlib.import_module(__package__ + ".control_utils")
    warp = importlib.import_module(__package__ + ".warp")
    extractors = importlib.import_module(__package__ + ".extractors")

    try:
        pkg = importlib.import_module("controlnet")
    except Exception:
        pkg = types.ModuleType("controlnet")
        pkg.__path__ = []
        sys.modules["controlnet"] = pkg
    # controlnet.extractors too (flownet.py:8 imports Bi_Dir_FeatureExtractor from it): same classes and state_dict keys, the
    # per-scale motion compensation fused into one call and the `holes.any()` host sync gone
    for name, mod in (("softsplat", softsplat), ("control_utils", control_utils), ("extractors", extractors)):
        full = f"controlnet.{name}"
        if force or full not in sys.modules:
            sys.modules[full] = mod
            setattr(pkg, name, mod)
    # the dormant bilinear backward warp lives under cmp.models.modules.warp in the reference
    full = "cmp.models.modules.warp"
    if full in sys.modules:
        sys.modules[full].WarpingLayerBWFlow = warp.WarpingLayerBWFlow
    # patch_utils.py is a top-level script module of the reference (cv2 / PIL imports): patch it if it is loaded
    if "patch_utils" in sys.modules:
        tiles = importlib.import_module(__package__ + ".patch_utils")
        sys.modules["patch_utils"].merge_latent_tiles_from_pixel_coords = tiles.merge_latent_tiles_from_pixel_coords
