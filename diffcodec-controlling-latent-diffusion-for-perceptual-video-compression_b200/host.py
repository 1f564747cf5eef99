"""Host-buffer entry point: splat frames that live in (pinned) host memory.

``softsplat_host`` is ``softsplat`` for inputs and outputs on the HOST: it streams the batch
through the GPU in chunks of a few frames over three CUDA streams (H2D copy / kernels / D2H copy),
so the uploads of chunk i+1 and the download of chunk i-1 overlap the kernels of chunk i. The
arithmetic is exactly ``softsplat`` on each chunk (frames are independent: a splat never crosses
a frame); nothing is computed on the CPU. This is what ``bench.py`` times as ``e2e``.
"""
from __future__ import annotations

import torch

from . import _lib
from .softsplat import softsplat

__all__ = ["softsplat_host", "bind_to_gpu_numa"]


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device=None) -> dict:
    """Pin the calling process to the host cores of the NUMA node its GPU hangs off (sysfs: the PCI device's
    ``numa_node``, the node's ``cpulist``), so that pinned buffers allocated AFTERWARDS are first-touched on that node and
    the copy threads run next to the GPU's PCIe root. One process per GPU calls this right after choosing its device;
    with 8 ranks on a two-socket box the default placement puts every rank's buffers on whatever node the launcher ran
    on. Returns what it found / did; never raises (containers often hide the topology)."""
    import os
    info = {"bound": False}
    try:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        props = torch.cuda.get_device_properties(dev)
        bus = f"{int(getattr(props, 'pci_domain_id', 0)):04x}:{int(props.pci_bus_id):02x}:{int(props.pci_device_id):02x}.0"
        info["pci"] = bus
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as e:                                      # no sysfs, no permission, unknown layout: run unbound
        info["error"] = repr(e)
    return info


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


def _stage_buffers(dev, slots: int, k: int, tenIn, tenFlow, tenMetric, out_dtype):
    """Persistent device staging buffers, one set per slot: no allocator traffic and no cross-stream allocator
    bookkeeping per chunk. Per slot: the uploaded tensors in their HOST element types, fp32 copies for the ones that
    are converted on the device (8-bit frames, half-precision flows / metrics), and a narrow result buffer when the
    caller wants a bf16 result."""
    key = (dev, slots, k, tuple(tenIn.shape[1:]), tenIn.dtype, tenFlow.dtype,
           None if tenMetric is None else (tuple(tenMetric.shape[1:]), tenMetric.dtype), out_dtype)
    bufs = _staging.get(key)
    if bufs is None:
        _staging.clear()                                        # one live configuration at a time
        def raw(t):
            return None if t is None else torch.empty((k,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        def f32(t):
            return None if (t is None or t.dtype == torch.float32) else torch.empty((k,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)
        def res():
            return None if out_dtype == torch.float32 else torch.empty((k,) + tuple(tenIn.shape[1:]), dtype=out_dtype, device=dev)
        bufs = _staging[key] = [(raw(tenIn), raw(tenFlow), raw(tenMetric), f32(tenIn), f32(tenFlow), f32(tenMetric), res()) for _ in range(slots)]
    return bufs


_HOST_IN = (torch.uint8, torch.float16, torch.bfloat16, torch.float32)
_HOST_AUX = (torch.float16, torch.bfloat16, torch.float32)


def softsplat_host(tenIn: torch.Tensor, tenFlow: torch.Tensor, tenMetric, strMode: str, out: torch.Tensor | None = None,
                   device=None, chunk_frames: int = 8, in_scale: float | None = None, sync: bool = True) -> torch.Tensor:
    """tenIn [N,C,H,W], tenFlow [N,2,H,W], tenMetric [N,1,H,W] or None: CPU tensors (pin them for full PCIe speed).
    Returns a CPU tensor [N,C,H,W] (``out`` if given; pinned if it was allocated here). No autograd (host tensors); same
    modes and asserts as ``softsplat``. ``chunk_frames`` is the steady-state chunk; the first and last chunks ramp
    1, 2, 4 ... frames.

    Element types. fp32 tensors are uploaded and splatted as they are. To cut PCIe bytes (the bound of this entry point),
    ``tenIn`` may be uint8 (decoded video; multiplied by ``in_scale``, default 1/255) or fp16 / bf16, and ``tenFlow`` /
    ``tenMetric`` may be fp16 / bf16: they are uploaded in that type and widened to fp32 ON THE DEVICE (dcb_convert); the
    splat always runs in fp32. ``out`` may be float32 (default) or bfloat16 (rounded once, on the device, before the
    download): 9 + 6 instead of 24 + 12 bytes per pixel for a 3-channel frame.

    Synchronisation. With ``sync=True`` (default) the call returns after the last download has landed: the result can be
    read and the input buffers reused right away. With ``sync=False`` it returns as soon as the work is enqueued: the
    caller's current CUDA stream is ordered behind the downloads, but the HOST must not read ``out`` nor modify the
    inputs before ``torch.cuda.current_stream().synchronize()`` (or any later synchronisation of that stream)."""
    assert not tenIn.is_cuda and not tenFlow.is_cuda, "softsplat_host takes host tensors; use softsplat for device tensors"
    assert tenIn.dtype in _HOST_IN and tenFlow.dtype in _HOST_AUX, "softsplat_host: uint8 / fp16 / bf16 / fp32 frames, fp16 / bf16 / fp32 flows"
    assert tenMetric is None or tenMetric.dtype in _HOST_AUX
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = tenIn.shape[0]
    if out is None:
        out = torch.empty(tenIn.shape, dtype=torch.float32).pin_memory()
    assert out.dtype in (torch.float32, torch.bfloat16) and tuple(out.shape) == tuple(tenIn.shape)
    if n == 0:
        return out
    if in_scale is None:
        in_scale = 1.0 / 255.0 if tenIn.dtype == torch.uint8 else 1.0
    sizes = _chunk_schedule(n, min(chunk_frames, n))
    slots = 3
    stage = _stage_buffers(dev, slots, max(sizes), tenIn, tenFlow, tenMetric, out.dtype)
    up, run, down = _pipeline_streams(dev)
    consumed = [None] * slots                                   # event: the kernels that read this slot are done
    drained = [None] * slots                                    # event: the narrow result buffer of this slot has been downloaded
    cur = torch.cuda.current_stream(dev)
    up.wait_stream(cur)
    up.wait_stream(run)                                         # staging buffers may still be read by an earlier call
    run.wait_stream(down)
    with torch.no_grad():
        lo = 0
        for i, c in enumerate(sizes):
            hi = lo + c
            s = i % slots
            r_in, r_fl, r_me, w_in, w_fl, w_me, b_out = stage[s]
            with torch.cuda.stream(up):
                if consumed[s] is not None:
                    up.wait_event(consumed[s])                  # do not overwrite inputs still in use
                d_in = r_in[:c]; d_in.copy_(tenIn[lo:hi], non_blocking=True)
                d_fl = r_fl[:c]; d_fl.copy_(tenFlow[lo:hi], non_blocking=True)
                d_me = None
                if tenMetric is not None:
                    d_me = r_me[:c]; d_me.copy_(tenMetric[lo:hi], non_blocking=True)
                uploaded = torch.cuda.Event(); uploaded.record(up)
            with torch.cuda.stream(run):
                run.wait_event(uploaded)
                if w_in is not None:
                    d_in = _lib.convert(d_in, w_in[:c], in_scale)
                elif in_scale != 1.0:
                    d_in = _lib.convert(d_in, d_in, in_scale)
                if w_fl is not None:
                    d_fl = _lib.convert(d_fl, w_fl[:c])
                if w_me is not None:
                    d_me = _lib.convert(d_me, w_me[:c])
                d_out = softsplat(d_in, d_fl, d_me, strMode)
                if b_out is not None:
                    if drained[s] is not None:
                        run.wait_event(drained[s])
                    d_out = _lib.convert(d_out, b_out[:c])
                consumed[s] = torch.cuda.Event(); consumed[s].record(run)
            with torch.cuda.stream(down):
                down.wait_event(consumed[s])
                out[lo:hi].copy_(d_out, non_blocking=True)
                if b_out is None:
                    d_out.record_stream(down)                   # allocated on `run`, last read on `down`
                else:
                    drained[s] = torch.cuda.Event(); drained[s].record(down)
            lo = hi
        finished = torch.cuda.Event(); finished.record(down)
        cur.wait_stream(down)                                   # the caller's stream sees the finished result
        cur.wait_stream(run)
    if sync:
        finished.synchronize()                                  # the host may read `out` and reuse the inputs now
    return out
