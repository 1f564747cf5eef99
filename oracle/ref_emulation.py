"""Run the reference's *own* kernel text on the CPU and mint golden vectors from it.

TEST INFRASTRUCTURE, NOT PRODUCT. Runs only in the build container (it imports
``/root/reference``, which does not exist on the GPU box); its outputs are the small
committed fixtures ``tests/golden/ref_emu_*.npz`` that pin ``oracle/`` (and through
it the CUDA path) to the reference.

How: the reference has no CPU path (``controlnet/softsplat.py:347-348``) and needs
CuPy, which is not installed. We

1. put a stub ``cupy`` in ``sys.modules`` exposing the five symbols the file uses
   (``int32``, ``float32``, ``memoize``, ``cuda.get_cuda_path``,
   ``cuda.compile_with_cache``; softsplat.py:18,23,219,222,225);
2. import the unmodified ``controlnet/softsplat.py`` from ``/root/reference``; its own
   ``cuda_kernel()`` templates the CUDA-C string (sizes, strides, dtype) exactly as on a GPU;
3. implement ``compile_with_cache`` by compiling that templated string with ``g++`` behind a
   25-line prelude that defines ``blockIdx/threadIdx/...``, ``__global__`` and a non-inlined
   ``atomicAdd`` (so the product is rounded before the add, as on the device), and
   ``get_function(name)(grid=, block=, args=)`` by running every block/thread sequentially;
4. make CPU tensors answer ``is_cuda == True`` while the reference runs.

The generated ``.cpp``/``.so`` land in ``oracle/_ref/`` (git-ignored). Reference sources
are never copied into the repository.

Two builds are produced: ``-ffp-contract=off`` and ``-ffp-contract=fast -mfma`` (the latter
imitates NVRTC's default ``-fmad=true``, which contracts the gather kernels' ``acc += g*w``).

Usage:  python oracle/ref_emulation.py            # regenerate tests/golden/ref_emu_*.npz
"""
from __future__ import annotations

import collections
import contextlib
import ctypes
import hashlib
import os
import re
import subprocess
import sys
import types

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_ROOT = os.environ.get("DCB_REFERENCE_ROOT", "/root/reference")
_OUT = os.path.join(_HERE, "_ref")

_PRELUDE = r"""
#include <math.h>
#include <cmath>
#include <cassert>
struct emu_dim3 { unsigned x, y, z; };
static emu_dim3 blockIdx, blockDim, threadIdx, gridDim;
#define __global__
#define __launch_bounds__(x)
using std::isfinite;
template <class T> __attribute__((noinline)) static T atomicAdd(T* p, T v) {
    volatile T rounded = v;          /* product is rounded before it reaches the adder */
    T old = *p; *p = old + rounded; return old;
}
"""


def _params(name: str, src: str):
    sig = re.search(name + r"\s*\(([^)]*)\)", src).group(1)
    return [p.strip() for p in sig.split(",")]


def _launcher(name: str, src: str) -> str:
    params = _params(name, src)
    decl, call = [], []
    for i, p in enumerate(params):
        if "*" in p:
            ty = p[: p.rindex("*") + 1].replace("__restrict__", "").strip()
            decl.append(f"void* a{i}")
            call.append(f"({ty}) a{i}")
        else:
            decl.append(f"int a{i}")
            call.append(f"a{i}")
    return (
        f'\nextern "C" void emu_launch_{name}(unsigned grid, unsigned block, {", ".join(decl)}) {{\n'
        "  gridDim = {grid, 1, 1}; blockDim = {block, 1, 1};\n"
        "  for (unsigned b = 0; b < grid; ++b) for (unsigned t = 0; t < block; ++t) {\n"
        "    blockIdx = {b, 0, 0}; threadIdx = {t, 0, 0};\n"
        f"    {name}({', '.join(call)});\n"
        "  }\n}\n"
    )


class _Module:
    def __init__(self, src: str, contract: str):
        self.src, self.contract = src, contract

    def get_function(self, name: str):
        os.makedirs(_OUT, exist_ok=True)
        full = _PRELUDE + self.src + _launcher(name, self.src)
        tag = hashlib.sha1((full + self.contract).encode()).hexdigest()[:16]
        cpp, so = os.path.join(_OUT, f"emu_{tag}.cpp"), os.path.join(_OUT, f"emu_{tag}.so")
        if not os.path.exists(so):
            with open(cpp, "w") as f:
                f.write(full)
            flags = ["-O2", "-fPIC", "-shared", "-fno-fast-math"]
            flags += ["-ffp-contract=fast", "-mfma"] if self.contract == "fast" else ["-ffp-contract=off"]
            subprocess.check_call(["g++", *flags, "-o", so, cpp])
        fn = getattr(ctypes.CDLL(so), f"emu_launch_{name}")

        is_ptr = ["*" in p for p in _params(name, self.src)]

        def launch(grid, block, args, stream=None):
            cargs = [ctypes.c_uint(grid[0]), ctypes.c_uint(block[0])]
            assert len(args) == len(is_ptr)
            for a, ptr in zip(args, is_ptr):
                cargs.append(ctypes.c_void_p(a) if ptr else ctypes.c_int(int(a)))
            fn(*cargs)

        return launch


def _install_stub_cupy(contract: str):
    cupy = types.ModuleType("cupy")
    cupy.int32 = lambda v: int(v)
    cupy.float32 = lambda v: float(v)

    def memoize(for_each_device=False):
        def deco(f):
            cache = {}

            def wrapped(key):
                if key not in cache:
                    cache[key] = f(key)
                return cache[key]

            return wrapped

        return deco

    cupy.memoize = memoize
    cupy.cuda = types.ModuleType("cupy.cuda")
    cupy.cuda.get_cuda_path = lambda: "/usr/local/cuda"
    cupy.cuda.compile_with_cache = lambda src, opts=(): _Module(src, contract)
    sys.modules["cupy"] = cupy
    sys.modules["cupy.cuda"] = cupy.cuda


def load_reference(contract: str = "fast"):
    """Import the unmodified reference module with the emulation back-end."""
    _install_stub_cupy(contract)
    for k in [k for k in sys.modules if k == "controlnet" or k.startswith("controlnet.")]:
        del sys.modules[k]
    sys.path.insert(0, _REF_ROOT)
    try:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import controlnet.softsplat as ref  # noqa: the reference, unmodified
    finally:
        sys.path.remove(_REF_ROOT)
    ref.objCudacache["device"] = "cpu-emulation"  # softsplat.py:28-30 would query a GPU name
    return ref


@contextlib.contextmanager
def pretend_cuda():
    """CPU tensors answer is_cuda=True; torch.cuda.current_stream() returns a null stream."""
    saved_stream = torch.cuda.current_stream
    torch.Tensor.is_cuda = property(lambda self: True)
    torch.cuda.current_stream = lambda *a, **k: collections.namedtuple("S", "cuda_stream")(0)
    try:
        yield
    finally:
        del torch.Tensor.is_cuda
        torch.cuda.current_stream = saved_stream


# ---------------------------------------------------------------------------------------------
# fixture generation
# ---------------------------------------------------------------------------------------------

def make_case(seed: int, n: int, c: int, h: int, w: int, flow_scale: float, dtype=torch.float32, special=False):
    g = torch.Generator().manual_seed(seed)
    tin = torch.randn(n, c, h, w, generator=g, dtype=dtype)
    flow = torch.randn(n, 2, h, w, generator=g, dtype=dtype) * flow_scale
    metric = torch.randn(n, 1, h, w, generator=g, dtype=dtype) * 0.5
    gout = torch.randn(n, c, h, w, generator=g, dtype=dtype)
    if special:  # integer / half-integer / out-of-frame / non-finite flows (SURVEY.md App. C 2,3,7)
        flow[0, :, 0, 0] = 0.0
        flow[0, 0, 0, 1], flow[0, 1, 0, 1] = 2.0, 1.0
        flow[0, 0, 1, 0], flow[0, 1, 1, 0] = 0.5, 0.0
        flow[0, 0, 1, 1], flow[0, 1, 1, 1] = -100.0, 3.0
        flow[0, 0, 2, 2] = float("nan")
        flow[0, 1, 2, 3] = float("inf")
        flow[0, 0, 3, 3] = -float("inf")
        flow[0, 0, 3, 0], flow[0, 1, 3, 0] = 1e30, -1e30
        flow[0, 0, 0, 3], flow[0, 1, 0, 3] = -0.25, -0.75
    return tin, flow, metric, gout


CASES = [
    # name, seed, N, C, H, W, flow_scale, special
    ("small_a", 11, 2, 3, 9, 13, 1.5, True),
    ("small_b", 12, 1, 5, 16, 11, 4.0, False),
    ("collide", 13, 1, 2, 8, 8, 0.0, False),   # flows overwritten below: all to one pixel
]
MODES = ["sum", "avg", "linear", "soft", "avg-zeroeps", "linear-clipeps", "soft-zeroeps", "soft-clipeps", "soft-addeps"]


def run_reference(ref, tin, flow, metric, gout, mode):
    """Forward + backward through the reference's softsplat(); returns numpy dict."""
    tin = tin.clone().requires_grad_(True)
    flow = flow.clone().requires_grad_(True)
    needs_metric = mode.split("-")[0] in ("linear", "soft")
    met = metric.clone().requires_grad_(True) if needs_metric else None
    with pretend_cuda():
        out = ref.softsplat(tenIn=tin, tenFlow=flow, tenMetric=met, strMode=mode)
        out.backward(gout[:, : out.shape[1]])  # 'avg-<eps>' drops a channel (exact-match quirk, softsplat.py:240)
    res = {"out": out.detach().numpy(), "gin": tin.grad.numpy(), "gflow": flow.grad.numpy()}
    if met is not None:
        res["gmetric"] = met.grad.numpy()
    return res


def generate(out_dir: str):
    os.makedirs(out_dir, exist_ok=True)
    for contract in ("fast", "off"):
        ref = load_reference(contract)
        for name, seed, n, c, h, w, scale, special in CASES:
            for dtype, dname in ((torch.float32, "f32"), (torch.float64, "f64")):
                tin, flow, metric, gout = make_case(seed, n, c, h, w, scale, dtype, special)
                if name == "collide":
                    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
                    flow[0, 0] = (3.25 - xs).to(dtype)
                    flow[0, 1] = (4.5 - ys).to(dtype)
                blob = {"tin": tin.numpy(), "flow": flow.numpy(), "metric": metric.numpy(), "gout": gout.numpy()}
                for mode in MODES:
                    r = run_reference(ref, tin, flow, metric, gout, mode)
                    for k, v in r.items():
                        blob[f"{mode}/{k}"] = v
                # func-level (softsplat_func.apply) on the raw tensors, both grads
                ti = tin.clone().requires_grad_(True)
                fl = flow.clone().requires_grad_(True)
                with pretend_cuda():
                    o = ref.softsplat_func.apply(ti, fl)
                    o.backward(gout)
                blob["func/out"], blob["func/gin"], blob["func/gflow"] = o.detach().numpy(), ti.grad.numpy(), fl.grad.numpy()
                path = os.path.join(out_dir, f"ref_emu_{name}_{dname}_{contract}.npz")
                np.savez_compressed(path, **blob)
                print("wrote", path, {k: v.shape for k, v in list(blob.items())[:2]})


if __name__ == "__main__":
    generate(os.path.join(os.path.dirname(_HERE), "tests", "golden"))
