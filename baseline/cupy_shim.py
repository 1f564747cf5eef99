"""A stand-in for the five CuPy symbols the reference's ``controlnet/softsplat.py`` touches, built on
cuda-python (NVRTC + the driver API), so that the UNMODIFIED reference file runs its own three kernel
strings on the GPU of this box (BASELINE.md section 5 "B-ref-gpu").

BASELINE / TEST INFRASTRUCTURE, NOT PRODUCT: only ``bench.py``'s reference-on-GPU leg, ``baseline/ref_gpu.py``
and the tests that mint / check golden vectors import it. CuPy itself is not installed in this image and
cannot be fetched (no network). What the reference uses (``controlnet/softsplat.py`` lines):

    cupy.int32, cupy.float32                       :18, :23       -> numpy scalars
    cupy.memoize(for_each_device=True)             :219           -> a per-device dict cache
    cupy.cuda.get_cuda_path()                      :222           -> $CUDA_HOME or /usr/local/cuda
    cupy.cuda.compile_with_cache(src, opts)        :225           -> nvrtcCompileProgram for the current device's arch
        .get_function(name)(grid=, block=, args=, stream=)  :340-345, :430-435, :519-524 -> cuLaunchKernel

Like CuPy, compilation targets the compute capability of the current device and results are cached per source.
"""
from __future__ import annotations

import ctypes
import os
import sys
import types

import numpy as np


def _nvrtc():
    from cuda.bindings import nvrtc
    return nvrtc


def _driver():
    from cuda.bindings import driver
    return driver


def _check(res, what):
    code = res[0] if isinstance(res, tuple) else res
    if int(code) != 0:
        raise RuntimeError(f"{what} failed: {code!r}")
    return res[1:] if isinstance(res, tuple) and len(res) > 1 else None


def compile_to_cubin(source: str, options=(), arch: str | None = None) -> bytes:
    """NVRTC-compile CUDA C to a cubin for `arch` (default: the current device, e.g. 'sm_100')."""
    nvrtc = _nvrtc()
    if arch is None:
        import torch
        major, minor = torch.cuda.get_device_capability()
        arch = f"sm_{major}{minor}"
    opts = [f"--gpu-architecture={arch}"]
    for o in options:                      # the reference passes '-I <dir>' as ONE string; NVRTC wants '-I<dir>'
        o = o.strip()
        if o.startswith("-I "):
            o = "-I" + o[3:].strip()
        opts.append(o)
    (prog,) = _check(nvrtc.nvrtcCreateProgram(source.encode(), b"kernel.cu", 0, [], []), "nvrtcCreateProgram")
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), [o.encode() for o in opts])
    if int(res[0]) != 0:
        (n,) = _check(nvrtc.nvrtcGetProgramLogSize(prog), "nvrtcGetProgramLogSize")
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError("NVRTC compilation of the reference kernel failed:\n" + log.decode(errors="replace"))
    (n,) = _check(nvrtc.nvrtcGetCUBINSize(prog), "nvrtcGetCUBINSize")
    cubin = b" " * n
    _check(nvrtc.nvrtcGetCUBIN(prog, cubin), "nvrtcGetCUBIN")
    nvrtc.nvrtcDestroyProgram(prog)
    return cubin


class _Function:
    def __init__(self, module, name: str):
        (self._fn,) = _check(_driver().cuModuleGetFunction(module, name.encode()), "cuModuleGetFunction")

    def __call__(self, grid, block, args, stream=None, shared_mem=0):
        drv = _driver()
        values, types_ = [], []
        for a in args:
            if isinstance(a, np.int32):
                values.append(int(a)); types_.append(ctypes.c_int)
            elif isinstance(a, np.float32):
                values.append(float(a)); types_.append(ctypes.c_float)
            else:                            # tensor.data_ptr(); None is a null pointer (softsplat.py:433, :523)
                values.append(0 if a is None else int(a)); types_.append(ctypes.c_void_p)
        ptr = getattr(stream, "ptr", 0) if stream is not None else 0
        grid = tuple(grid) + (1,) * (3 - len(grid)); block = tuple(block) + (1,) * (3 - len(block))
        _check(drv.cuLaunchKernel(self._fn, grid[0], grid[1], grid[2], block[0], block[1], block[2], shared_mem,
                                  drv.CUstream(ptr), (tuple(values), tuple(types_)), 0), "cuLaunchKernel")


class _Module:
    _cache: dict = {}

    def __init__(self, source: str, options):
        import torch
        torch.cuda.init()
        torch.zeros(1, device="cuda")        # make torch's primary context current on this thread
        key = (source, tuple(options), torch.cuda.current_device())
        mod = _Module._cache.get(key)
        if mod is None:
            cubin = compile_to_cubin(source, options)
            (mod,) = _check(_driver().cuModuleLoadData(cubin), "cuModuleLoadData")
            _Module._cache[key] = mod
        self._mod = mod

    def get_function(self, name: str) -> _Function:
        return _Function(self._mod, name)


def _memoize(for_each_device: bool = False):
    def deco(fn):
        cache = {}

        def wrapper(*args):
            import torch
            key = (torch.cuda.current_device() if for_each_device else None,) + args
            if key not in cache:
                cache[key] = fn(*args)
            return cache[key]
        return wrapper
    return deco


def install() -> types.ModuleType:
    """Put the stand-in into sys.modules as ``cupy`` (no-op if a real CuPy is importable)."""
    try:
        import cupy  # noqa: F401
        if not getattr(cupy, "_dcb_shim", False):
            return cupy
    except Exception:
        pass
    cupy = types.ModuleType("cupy")
    cupy._dcb_shim = True
    cupy.int32, cupy.float32 = np.int32, np.float32
    cupy.memoize = _memoize
    cuda = types.ModuleType("cupy.cuda")
    cuda.get_cuda_path = lambda: os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cuda.compile_with_cache = lambda source, options=(), **kw: _Module(source, options)
    cupy.cuda = cuda
    sys.modules["cupy"] = cupy
    sys.modules["cupy.cuda"] = cuda
    return cupy
