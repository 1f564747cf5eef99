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


class _bidir_block_func(torch.autograd.Function):
    """The whole block as ONE autograd node (metric = ones, flows without gradient): the same five
    library calls forward and three backward as the composition below, without four of its five
    autograd nodes and their bookkeeping -- the block is bound by Python, not by its kernels."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)       # the splats run in fp32 (control_utils.py:61)
    def forward(ctx, first, last, flow_f, flow_b):
        import importlib
        ss = importlib.import_module(__package__ + ".softsplat")
        dt = first.dtype
        occ_fwd = compute_mask(flow_f, flow_b)
        occ_bwd = compute_mask(flow_b, flow_f)
        mf, mb = (occ_fwd, occ_bwd) if dt == torch.float32 else (occ_fwd.to(dt), occ_bwd.to(dt))
        metric = torch.ones_like(flow_f[:, :1], dtype=dt)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        ss._check_inputs(first, flow_f); ss._check_inputs(last, flow_b)
        ff, fb = ss._match_flow(first, flow_f), ss._match_flow(last, flow_b)
        det = ss.is_deterministic()
        w1, n1 = ss._forward(first, ff, metric, mf, _lib.MODE_SOFT, _lib.EPS_ADD, det, need)
        w2, n2 = ss._forward(last, fb, metric, mb, _lib.MODE_SOFT, _lib.EPS_ADD, det, need)
        fused, _ = _fuse_forward(w1, w2, metric, metric, mf, mb)
        if need:
            ctx.save_for_backward(first, last, ff, fb, metric, mf, mb, w1, n1, w2, n2)
        return fused

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        import importlib
        ss = importlib.import_module(__package__ + ".softsplat")
        first, last, ff, fb, metric, mf, mb, w1, n1, w2, n2 = ctx.saved_tensors
        lib = _lib.lib()
        dev = first.device
        need = ctx.needs_input_grad
        g = g.to(first.dtype)
        gA = torch.empty_like(w1) if need[0] else None
        gB = torch.empty_like(w2) if need[1] else None
        with _lib.on_device(dev):
            rc = lib.dcb_bidir_fuse_bwd(_lib.desc(g), _lib.desc(w1), _lib.desc(w2), _lib.desc(metric), _lib.desc(metric),
                                        _lib.desc(mf), _lib.desc(mb), _lib.desc(gA), _lib.desc(gB), None, None, _lib.stream_ptr(dev))
        _lib.check(rc, "dcb_bidir_fuse_bwd")
        g1 = ss._backward(gA, first, ff, metric, w1, n1, mf, _lib.MODE_SOFT, _lib.EPS_ADD, (True, False, False))[0] if need[0] else None
        g2 = ss._backward(gB, last, fb, metric, w2, n2, mb, _lib.MODE_SOFT, _lib.EPS_ADD, (True, False, False))[0] if need[1] else None
        return g1, g2, None, None


def bidirectional_warp_fuse(first_features, last_features, flow_f, flow_b, warper):
    """The per-scale block of ``Bi_Dir_FeatureExtractor.forward`` between the conv stacks
    (``extractors.py:289-310``): masks, both warps, fusion. ``warper`` is a ``FeatureWarperSoftsplat``."""
    if (not getattr(warper, "with_learnable_metric", True) and not flow_f.requires_grad and not flow_b.requires_grad
            and first_features.is_cuda and first_features.dtype == last_features.dtype
            and flow_f.dtype in (torch.float32, torch.bfloat16) and flow_b.dtype == flow_f.dtype
            and first_features.shape == last_features.shape):
        return _bidir_block_func.apply(first_features, last_features, flow_f, flow_b)
    occ_fwd = compute_mask(flow_f, flow_b)
    occ_bwd = compute_mask(flow_b, flow_f)
    warped_first, conf_fwd = warper(first_features, flow_f, mask=occ_fwd)
    warped_last, conf_bwd = warper(last_features, flow_b, mask=occ_bwd)
    return bidir_fuse(warped_first, warped_last, conf_fwd, conf_bwd, occ_fwd, occ_bwd)
