"""B-ref-gpu: the reference's own, unmodified ``controlnet/softsplat.py`` running on this GPU.

BASELINE / TEST INFRASTRUCTURE, NOT PRODUCT. The file itself is never committed: ``stage()`` copies it from
``/root/reference`` into the git-ignored ``baseline/_ref/`` (it travels to the GPU box with the snapshot, like the
built ``.so`` files; ``__graft_entry__.build()`` calls ``stage()`` when ``/root/reference`` is present). ``load()``
imports that copy behind ``baseline/cupy_shim.py`` and returns the module: ``load().softsplat(tenIn, tenFlow,
tenMetric, strMode)`` is the reference's public entry point (its eager pre/post ops + its NVRTC-compiled kernels).
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SRC = os.path.join(os.environ.get("DCB_REFERENCE_ROOT", "/root/reference"), "controlnet", "softsplat.py")
_REF_DST = os.path.join(_HERE, "_ref", "controlnet", "softsplat.py")
_module = None


def stage() -> bool:
    """Copy the reference file into baseline/_ref/ (git-ignored). Returns False when /root/reference is absent."""
    if not os.path.exists(_REF_SRC):
        return os.path.exists(_REF_DST)
    os.makedirs(os.path.dirname(_REF_DST), exist_ok=True)
    shutil.copyfile(_REF_SRC, _REF_DST)
    return True


def available() -> str | None:
    """None if the reference can run here, else the reason it cannot."""
    if not os.path.exists(_REF_DST):
        return "baseline/_ref/controlnet/softsplat.py is absent (stage() needs /root/reference)"
    try:
        from cuda.bindings import driver, nvrtc  # noqa: F401
    except Exception as e:
        return f"cuda-python is not importable: {e!r}"
    import torch
    if not torch.cuda.is_available():
        return "no CUDA device"
    return None


def load():
    global _module
    if _module is None:
        why = available()
        if why:
            raise RuntimeError("reference-on-GPU unavailable: " + why)
        if _HERE not in sys.path:
            sys.path.insert(0, _HERE)
        import cupy_shim
        cupy_shim.install()
        spec = importlib.util.spec_from_file_location("dcb_reference_softsplat", _REF_DST)
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")              # torch.cuda.amp.custom_fwd / custom_bwd deprecation notes
            spec.loader.exec_module(mod)
        _module = mod
    return _module


if __name__ == "__main__":
    print("staged" if stage() else "nothing to stage")
