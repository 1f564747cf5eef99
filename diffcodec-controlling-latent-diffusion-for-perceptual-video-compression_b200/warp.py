"""Drop-in replacement for ``WarpingLayerBWFlow`` (reference ``cmp/models/modules/warp.py:4-25``).

The reference builds a normalised sampling grid with zeros_like / linspace / cat / .cuda() /
permute on every call and hands it to ``F.grid_sample``. Here the same fp32 operation sequence
is evaluated in registers by one gather kernel (dcb_backwarp_fwd), optionally fused with the
residual ``gt - warped`` (reference ``controlnet/residual_utils.py:199``); its autograd backward
is dcb_backwarp_bwd.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

__all__ = ["WarpingLayerBWFlow", "backwarp", "backwarp_residual"]


def _match_flow(image, flow):
    if flow.dtype == image.dtype or (image.dtype == torch.bfloat16 and flow.dtype == torch.float32):
        return flow
    return flow.to(image.dtype)


class _backwarp_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, flow, gt, align_corners):
        assert image.dim() == 4 and flow.dim() == 4 and flow.shape[1] == 2
        assert image.is_cuda and flow.is_cuda, "backwarp has no CPU path"
        flow = _match_flow(image, flow)
        lib = _lib.lib()
        dev = image.device
        warped = torch.empty(image.shape, dtype=image.dtype, device=dev)
        residual = torch.empty_like(warped) if gt is not None else None
        with _lib.on_device(dev):
            rc = lib.dcb_backwarp_fwd(_lib.desc(image), _lib.desc(flow), _lib.desc(gt), _lib.desc(warped),
                                      _lib.desc(residual), int(bool(align_corners)), _lib.stream_ptr(dev))
        _lib.check(rc, "dcb_backwarp_fwd")
        ctx.align = int(bool(align_corners))
        ctx.has_gt = gt is not None
        ctx.save_for_backward(image, flow)
        if gt is None:
            return warped
        return warped, residual

    @staticmethod
    def backward(ctx, gwarped, gresidual=None):
        image, flow = ctx.saved_tensors
        g = gwarped
        if ctx.has_gt and gresidual is not None:
            g = gresidual.neg() if g is None else g - gresidual           # residual = gt - warped
        lib = _lib.lib()
        dev = image.device
        n, c, h, w = image.shape
        gimage = torch.empty_like(image, memory_format=torch.contiguous_format) if ctx.needs_input_grad[0] else None
        gflow = torch.empty((n, 2, h, w), dtype=flow.dtype, device=dev) if ctx.needs_input_grad[1] else None
        need = lib.dcb_backwarp_bwd_workspace_bytes(n, c, h, w, _lib._DTYPES[image.dtype]) if gimage is not None else 0
        ws = _lib.workspace(dev, need, "scratch") if need > 0 else None
        with _lib.on_device(dev):
            rc = lib.dcb_backwarp_bwd(_lib.desc(g), _lib.desc(image), _lib.desc(flow), _lib.desc(gimage),
                                      _lib.desc(gflow), ctx.align, ws.data_ptr() if ws is not None else None,
                                      ws.numel() if ws is not None else 0, _lib.stream_ptr(dev))
        _lib.check(rc, "dcb_backwarp_bwd")
        ggt = None
        if ctx.has_gt and ctx.needs_input_grad[2]:
            ggt = gresidual
        return gimage, gflow, ggt, None


def backwarp(image, flow, align_corners: bool = False):
    """``out[n,c,y,x] = bilinear(image[n,c], (x, y) + flow)`` with zeros padding.

    align_corners=False reproduces the reference layer as executed by current torch
    (``grid_sample`` default): sx = (x + fx) * W/(W-1) - 0.5; True is the convention its flow
    normalisation was written for: sx = x + fx."""
    return _backwarp_func.apply(image, flow, None, align_corners)


def backwarp_residual(image, flow, gt, align_corners: bool = False):
    """Returns ``(warped, gt - warped)`` from one pass over the data."""
    return _backwarp_func.apply(image, flow, gt, align_corners)


class WarpingLayerBWFlow(nn.Module):
    """Same constructor and ``forward(image, flow)`` as the reference class."""

    def __init__(self, align_corners: bool = False):
        super().__init__()
        self.align_corners = align_corners

    def forward(self, image, flow):
        return backwarp(image, flow, self.align_corners)
