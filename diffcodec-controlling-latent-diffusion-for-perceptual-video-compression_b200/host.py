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


_staging: dict = {}
_streams: dict = {}


def _pipeline_streams(dev):
    """The three streams (upload / kernels / download) are created once per device: the library's
    accumulator workspace is cached per stream, so a fresh stream per call would re-create it."""
    st = _streams.get(dev)
    if st is None:
        st = _streams[dev] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return st


def _chunk_schedule(n: int, k: int) -> list[int]:
    """Frames per chunk: 1, 2, 4 ... up to k, k ... k, then back down to 1, so that neither the first
    upload nor the last download (the two transfers nothing can hide) is a full-size chunk."""
    k = max(1, k)
    ramp = []
    s = 1
    while s < k:
        ramp.append(s)
        s *= 2
    if n < 2 * sum(ramp) + k:                                   # short batch: small equal chunks
        c = max(1, min(k, n // 6 or 1))
        return [c] * (n // c) + ([n % c] if n % c else [])
    mid = n - 2 * sum(ramp)
    return ramp + [k] * (mid // k) + ([mid % k] if mid % k else []) + ramp[::-1]


def _stage_buffers(dev, slots: int, k: int, tenIn, tenFlow, tenMetric):
    """Persistent device staging buffers (inputs only), one set per slot: no allocator traffic
    and no cross-stream allocator bookkeeping per chunk."""
    key = (dev, slots, k, tuple(tenIn.shape[1:]), tenIn.dtype, tenFlow.dtype,
           None if tenMetric is None else (tuple(tenMetric.shape[1:]), tenMetric.dtype))
    bufs = _staging.get(key)
    if bufs is None:
        _staging.clear()                                        # one live configuration at a time
        def mk(t):
            return None if t is None else torch.empty((k,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        bufs = _staging[key] = [(mk(tenIn), mk(tenFlow), mk(tenMetric)) for _ in range(slots)]
    return bufs


def softsplat_host(tenIn: torch.Tensor, tenFlow: torch.Tensor, tenMetric, strMode: str, out: torch.Tensor | None = None,
                   device=None, chunk_frames: int = 8) -> torch.Tensor:
    """tenIn [N,C,H,W], tenFlow [N,2,H,W], tenMetric [N,1,H,W] or None: CPU tensors (pin them for
    full PCIe speed). Returns a CPU tensor [N,C,H,W] (``out`` if given; pinned if it was allocated here).
    No autograd (host tensors); same modes and asserts as ``softsplat``. ``chunk_frames`` is the
    steady-state chunk; the first and last chunks ramp 1, 2, 4 ... frames."""
    assert not tenIn.is_cuda and not tenFlow.is_cuda, "softsplat_host takes host tensors; use softsplat for device tensors"
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = tenIn.shape[0]
    if out is None:
        out = torch.empty(tenIn.shape, dtype=tenIn.dtype).pin_memory()
    if n == 0:
        return out
    sizes = _chunk_schedule(n, min(chunk_frames, n))
    slots = 3
    stage = _stage_buffers(dev, slots, max(sizes), tenIn, tenFlow, tenMetric)
    up, run, down = _pipeline_streams(dev)
    consumed = [None] * slots                                   # event: the kernels that read this slot are done
    cur = torch.cuda.current_stream(dev)
    up.wait_stream(cur)
    up.wait_stream(run)                                         # staging buffers may still be read by an earlier call
    with torch.no_grad():
        lo = 0
        for i, c in enumerate(sizes):
            hi = lo + c
            s = i % slots
            b_in, b_fl, b_me = stage[s]
            with torch.cuda.stream(up):
                if consumed[s] is not None:
                    up.wait_event(consumed[s])                  # do not overwrite inputs still in use
                d_in = b_in[:c]; d_in.copy_(tenIn[lo:hi], non_blocking=True)
                d_fl = b_fl[:c]; d_fl.copy_(tenFlow[lo:hi], non_blocking=True)
                d_me = None
                if tenMetric is not None:
                    d_me = b_me[:c]; d_me.copy_(tenMetric[lo:hi], non_blocking=True)
                uploaded = torch.cuda.Event(); uploaded.record(up)
            with torch.cuda.stream(run):
                run.wait_event(uploaded)
                d_out = softsplat(d_in, d_fl, d_me, strMode)
                consumed[s] = torch.cuda.Event(); consumed[s].record(run)
            with torch.cuda.stream(down):
                down.wait_event(consumed[s])
                out[lo:hi].copy_(d_out, non_blocking=True)
                d_out.record_stream(down)                       # allocated on `run`, last read on `down`
            lo = hi
        cur.wait_stream(down)                                   # the caller's stream sees the finished result
        cur.wait_stream(run)
    return out
