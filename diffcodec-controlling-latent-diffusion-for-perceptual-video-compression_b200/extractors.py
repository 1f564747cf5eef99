"""The motion-compensation part of a bi-directional conditioning block.

The reference's ``Bi_Dir_FeatureExtractor.forward`` (``controlnet/extractors.py:280-314``) does, per
pyramid scale: two occlusion masks, two soft splats with ``(1 - mask)``, a confidence fusion
(cat / clamp / sum / divide / two products / add) and a double-hole fill guarded by
``holes.any()`` -- a device->host sync four times per ControlNet forward. Its conv stacks are
cuDNN business and stay where they are; this module provides the rest as three fused calls with
no host sync:

    occ_fwd, occ_bwd = compute_mask(...)                (dcb_occlusion_mask, x2)
    warped, conf     = warper(feat, flow, mask=occ)     (dcb_splat_fwd with mask, x2)
    fused            = bidir_fuse(...)                  (dcb_bidir_fuse_fwd / _bwd)
"""
from __future__ import annotations

import torch

from . import _lib
from .control_utils import compute_mask

__all__ = ["bidir_fuse", "bidirectional_warp_fuse"]


def _fuse_forward(A, B, conf_a, conf_b, occ_a, occ_b):
    assert A.is_cuda and A.dim() == 4 and A.shape == B.shape, "bidir_fuse expects two [N,C,H,W] CUDA tensors"
    dt = A.dtype
    if B.dtype != dt: B = B.to(dt)
    if conf_a.dtype != dt: conf_a = conf_a.to(dt)
    if conf_b.dtype != dt: conf_b = conf_b.to(dt)
    if occ_a is not None and (occ_a.dtype != dt or occ_b.dtype != dt):
        occ_a, occ_b = occ_a.to(dt), occ_b.to(dt)
    lib = _lib.lib()
    dev = A.device
    fused = torch.empty(A.shape, dtype=dt, device=dev)
    with _lib.on_device(dev):
        rc = lib.dcb_bidir_fuse_fwd(_lib.desc(A), _lib.desc(B), _lib.desc(conf_a), _lib.desc(conf_b),
                                    _lib.desc(occ_a), _lib.desc(occ_b), _lib.desc(fused), _lib.stream_ptr(dev))
    _lib.check(rc, "dcb_bidir_fuse_fwd")
    return fused, (A, B, conf_a, conf_b, occ_a, occ_b)


class _bidir_fuse_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, conf_a, conf_b, occ_a, occ_b):
        fused, (A, B, conf_a, conf_b, occ_a, occ_b) = _fuse_forward(A, B, conf_a, conf_b, occ_a, occ_b)
        ctx.has_occ = occ_a is not None
        ctx.save_for_backward(*([A, B, conf_a, conf_b] + ([occ_a, occ_b] if occ_a is not None else [])))
        return fused

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        A, B, conf_a, conf_b = saved[:4]
        occ_a, occ_b = (saved[4], saved[5]) if ctx.has_occ else (None, None)
        need = ctx.needs_input_grad
        lib = _lib.lib()
        dev = A.device
        gA = torch.empty_like(A, memory_format=torch.contiguous_format) if need[0] else None
        gB = torch.empty_like(A, memory_format=torch.contiguous_format) if need[1] else None
        gca = torch.empty(conf_a.shape, dtype=A.dtype, device=dev) if need[2] else None
        gcb = torch.empty(conf_a.shape, dtype=A.dtype, device=dev) if need[3] else None
        with _lib.on_device(dev):
            rc = lib.dcb_bidir_fuse_bwd(_lib.desc(g.to(A.dtype)), _lib.desc(A), _lib.desc(B), _lib.desc(conf_a),
                                        _lib.desc(conf_b), _lib.desc(occ_a), _lib.desc(occ_b), _lib.desc(gA),
                                        _lib.desc(gB), _lib.desc(gca), _lib.desc(gcb), _lib.stream_ptr(dev))
        _lib.check(rc, "dcb_bidir_fuse_bwd")
        return gA, gB, gca, gcb, None, None


def bidir_fuse(warped_a, warped_b, conf_a, conf_b, occ_a=None, occ_b=None):
    """``w = clamp(conf, 0) / (sum + 1e-6); fused = w0*A + w1*B``; where both are occluded
    (``occ_a + occ_b > 1.5``) ``0.5*(A + B)`` -- reference ``extractors.py:298-310`` -- in one kernel,
    differentiable w.r.t. both maps and both confidences, without the ``holes.any()`` host sync."""
    assert (occ_a is None) == (occ_b is None)
    if torch.is_grad_enabled() and any(t.requires_grad for t in (warped_a, warped_b, conf_a, conf_b)):
        return _bidir_fuse_func.apply(warped_a, warped_b, conf_a, conf_b, occ_a, occ_b)
    return _fuse_forward(warped_a, warped_b, conf_a, conf_b, occ_a, occ_b)[0]     # inference: no autograd.Function bookkeeping


def bidirectional_warp_fuse(first_features, last_features, flow_f, flow_b, warper):
    """The per-scale block of ``Bi_Dir_FeatureExtractor.forward`` between the conv stacks
    (``extractors.py:289-310``): masks, both warps, fusion. ``warper`` is a ``FeatureWarperSoftsplat``."""
    occ_fwd = compute_mask(flow_f, flow_b)
    occ_bwd = compute_mask(flow_b, flow_f)
    warped_first, conf_fwd = warper(first_features, flow_f, mask=occ_fwd)
    warped_last, conf_bwd = warper(last_features, flow_b, mask=occ_bwd)
    return bidir_fuse(warped_first, warped_last, conf_fwd, conf_bwd, occ_fwd, occ_bwd)
