"""Host-buffer entry point: splat frames that live in (pinned) host memory.

``softsplat_host`` is ``softsplat`` for inputs and outputs on the HOST: it streams the batch
through the GPU in chunks of a few frames over three CUDA streams (H2D copy / kernels / D2H copy),
so the uploads of chunk i+1 and the download of chunk i-1 overlap the kernels of chunk i. The
arithmetic is exactly ``softsplat`` on each chunk (frames are independent: a splat never crosses
a frame); nothing is computed on the CPU. This is what ``bench.py`` times as ``e2e``.
"""
from __future__ import annotations

import torch

from .softsplat import softsplat

__all__ = ["softsplat_host"]


def softsplat_host(tenIn: torch.Tensor, tenFlow: torch.Tensor, tenMetric, strMode: str, out: torch.Tensor | None = None,
                   device=None, chunk_frames: int = 4) -> torch.Tensor:
    """tenIn [N,C,H,W], tenFlow [N,2,H,W], tenMetric [N,1,H,W] or None: CPU tensors (pin them for
    full PCIe speed). Returns a CPU tensor [N,C,H,W] (``out`` if given; pinned if it was allocated here).
    No autograd (host tensors); same modes and asserts as ``softsplat``."""
    assert not tenIn.is_cuda and not tenFlow.is_cuda, "softsplat_host takes host tensors; use softsplat for device tensors"
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = tenIn.shape[0]
    if out is None:
        out = torch.empty(tenIn.shape, dtype=tenIn.dtype).pin_memory()
    if n == 0:
        return out
    k = max(1, min(chunk_frames, n))
    up, run, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    slots = 3                                                   # device staging buffers in rotation
    bufs = [None] * slots
    freed = [None] * slots                                      # event: the slot's previous result has left the device
    with torch.no_grad():
        for i, lo in enumerate(range(0, n, k)):
            hi = min(n, lo + k)
            s = i % slots
            with torch.cuda.stream(up):
                if freed[s] is not None:
                    up.wait_event(freed[s])                     # do not overwrite inputs still in use
                d_in = tenIn[lo:hi].to(dev, non_blocking=True)
                d_fl = tenFlow[lo:hi].to(dev, non_blocking=True)
                d_me = tenMetric[lo:hi].to(dev, non_blocking=True) if tenMetric is not None else None
                uploaded = torch.cuda.Event(); uploaded.record(up)
            with torch.cuda.stream(run):
                run.wait_event(uploaded)
                d_out = softsplat(d_in, d_fl, d_me, strMode)
                for t in (d_in, d_fl, d_me, d_out):
                    if t is not None:
                        t.record_stream(run)
                computed = torch.cuda.Event(); computed.record(run)
            with torch.cuda.stream(down):
                down.wait_event(computed)
                out[lo:hi].copy_(d_out, non_blocking=True)
                d_out.record_stream(down)
                freed[s] = torch.cuda.Event(); freed[s].record(down)
            bufs[s] = (d_in, d_fl, d_me, d_out)
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(down)                                   # the caller's stream sees the finished result
    return out
