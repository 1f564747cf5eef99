"""Import alias: ``import diffcodec_b200`` loads the package whose directory carries the full
(hyphenated, hence not a Python identifier) project name."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("diffcodec-controlling-latent-diffusion-for-perceptual-video-compression_b200")
sys.modules[__name__] = _pkg
