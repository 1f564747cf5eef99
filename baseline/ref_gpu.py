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
_REF_ROOT = os.path.join(os.environ.get("DCB_REFERENCE_ROOT", "/root/reference"), "controlnet")
_REF_SRC = os.path.join(_REF_ROOT, "softsplat.py")
_REF_DST = os.path.join(_HERE, "_ref", "controlnet", "softsplat.py")
_STAGED = ("softsplat.py", "control_utils.py", "extractors.py")     # the op, its helpers, and the live consumer
_module = None
_package = None


def stage() -> bool:
    """Copy the reference files into baseline/_ref/ (git-ignored). Returns False when /root/reference is absent."""
    if not os.path.exists(_REF_SRC):
        return os.path.exists(_REF_DST)
    os.makedirs(os.path.dirname(_REF_DST), exist_ok=True)
    for name in _STAGED:
        shutil.copyfile(os.path.join(_REF_ROOT, name), os.path.join(os.path.dirname(_REF_DST), name))
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


def load_package():
    """The reference's ``controlnet.softsplat`` / ``control_utils`` / ``extractors`` as a namespace of module objects
    (``.softsplat``, ``.control_utils``, ``.extractors``), imported from baseline/_ref behind the CuPy stand-in. The
    ``controlnet*`` entries are removed from sys.modules again, so this never collides with diffcodec_b200.install()."""
    global _package
    if _package is None:
        why = available()
        if why:
            raise RuntimeError("reference-on-GPU unavailable: " + why)
        if _HERE not in sys.path:
            sys.path.insert(0, _HERE)
        import cupy_shim
        import types
        cupy_shim.install()
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "controlnet" or k.startswith("controlnet.")}
        root = os.path.join(_HERE, "_ref")
        sys.path.insert(0, root)
        try:
            import importlib
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mods = {n: importlib.import_module("controlnet." + n) for n in ("softsplat", "control_utils", "extractors")}
        finally:
            sys.path.remove(root)
            for k in [k for k in sys.modules if k == "controlnet" or k.startswith("controlnet.")]:
                del sys.modules[k]
            sys.modules.update(saved)
        _package = types.SimpleNamespace(**mods)
    return _package


if __name__ == "__main__":
    print("staged" if stage() else "nothing to stage")
