"""Drop-in replacement for the hot-path part of the reference's ``controlnet/control_utils.py``.

Same names and call signatures: ``compute_mask`` (reference ``control_utils.py:11-17``),
``FeatureWarperSoftsplat`` (``:36-72``), ``resize_and_normalize_flow_batched`` (``:74-97``).
``zero_module`` and ``FDN`` (``:6-9, 19-34``) are not on the motion-compensation path (plain
conv / GroupNorm modules); they are re-stated only so that ``from controlnet.control_utils
import ...`` keeps working when this module is installed in its place.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .softsplat import softsplat, _splat_normalised, is_deterministic

__all__ = ["zero_module", "compute_mask", "FDN", "FeatureWarperSoftsplat", "resize_and_normalize_flow_batched"]


def zero_module(module):
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


def compute_mask(flow_bwd_tensor, flow_fwd_tensor):
    """Occlusion mask: the FIRST argument is splatted by the SECOND (names as in the reference,
    whose call sites pass them swapped -- ``extractors.py:290-291``); then
    ``||second + splat||_2 > 0.3`` as float. One scatter + one epilogue launch (dcb_occlusion_mask)
    instead of ~12 eager kernels; no gradient flows (the reference's comparison blocks it too).
    """
    a, b = flow_bwd_tensor, flow_fwd_tensor
    assert a.dim() == 4 and a.shape[1] == 2 and a.shape == b.shape, "compute_mask expects two [N,2,H,W] flows"
    with torch.no_grad():
        if a.dtype not in (torch.float32, torch.bfloat16) or is_deterministic():
            # fp64 / deterministic: compose from the general op (same arithmetic, more launches)
            metric = torch.ones_like(b[:, :1])
            warped = softsplat(tenIn=a, tenFlow=b, tenMetric=metric, strMode="soft")
            return (torch.norm(b + warped, p=2, dim=1, keepdim=True) > 0.3).float()
        if b.dtype != a.dtype:
            b = b.to(a.dtype)
        lib = _lib.lib()
        n, _, h, w = a.shape
        dev = a.device
        mask = torch.empty((n, 1, h, w), dtype=a.dtype, device=dev)
        need = lib.dcb_occlusion_mask_workspace_bytes(n, h, w)
        scratch = _lib.fwd_is_scratch(n, 2, h, w, _lib._DTYPES[a.dtype], _lib.MODE_SOFT)
        ws = _lib.workspace(dev, need, "scratch" if scratch else "acc")
        with _lib.on_device(dev):
            rc = lib.dcb_occlusion_mask(_lib.desc(a), _lib.desc(b), _lib.desc(mask), ws.data_ptr(), ws.numel(),
                                        0 if scratch else _lib.FLAG_WS_CLEAN, _lib.stream_ptr(dev))
        if rc != 0:
            _lib.invalidate_acc(dev)
        _lib.check(rc, "dcb_occlusion_mask")
        return mask.float()                                   # reference returns float32 (.float())


class FDN(nn.Module):
    """Feature-denormalisation block (reference ``control_utils.py:19-34``); plain torch, not on the path."""

    def __init__(self, norm_nc, label_nc):
        super().__init__()
        self.param_free_norm = nn.GroupNorm(32, norm_nc, affine=False)
        self.conv_gamma = nn.Conv2d(label_nc, norm_nc, kernel_size=3, padding=1)
        self.conv_beta = nn.Conv2d(label_nc, norm_nc, kernel_size=3, padding=1)

    def forward(self, x, local_features):
        assert local_features.size()[2:] == x.size()[2:]
        normalized = self.param_free_norm(x)
        return normalized * (1 + self.conv_gamma(local_features)) + self.conv_beta(local_features)


class FeatureWarperSoftsplat(nn.Module):
    """Soft-splat a feature map, optionally with a learned metric, then apply ``(1 - mask)``.

    Parameter names match the reference (``metric_net.{0,2}.{weight,bias}``) so checkpoints load.
    The metric net is a cuDNN conv stack (not this library's business); the splat, the
    normalisation and the mask product are one fused pass, with a fused backward into the
    features and the learned metric.
    """

    def __init__(self, with_learnable_metric=False, in_channels=128):
        super().__init__()
        self.with_learnable_metric = with_learnable_metric
        if with_learnable_metric:
            self.metric_net = nn.Sequential(
                nn.Conv2d(in_channels, 64, kernel_size=3, padding=1),
                nn.SiLU(inplace=True),
                nn.Conv2d(64, 1, kernel_size=3, padding=1),
            )

    def forward(self, feat_ref, flow, mask=None):
        if self.with_learnable_metric:
            metric = self.metric_net(feat_ref)
        else:
            metric = torch.ones_like(flow[:, :1])
        with torch.autocast(device_type="cuda", enabled=False):               # control_utils.py:61
            if mask is not None and not mask.requires_grad and mask.shape == metric.shape:
                warped = _splat_normalised(feat_ref, flow, metric, _lib.MODE_SOFT, _lib.EPS_ADD, mask=mask)
            else:
                warped = softsplat(tenIn=feat_ref, tenFlow=flow, tenMetric=metric, strMode="soft")
                if mask is not None:
                    warped = warped * (1 - mask)
        return warped, metric


def resize_and_normalize_flow_batched(flow_tensor: torch.Tensor, target_h: int, target_w: int) -> torch.Tensor:
    """Bilinear resize (align_corners=False) then u /= (w-1)/2, v /= (h-1)/2 (reference
    ``control_utils.py:74-97``; note: no magnitude rescale for the resolution change)."""
    resized = F.interpolate(flow_tensor, size=(target_h, target_w), mode="bilinear", align_corners=False)
    u = resized[:, 0] / ((target_w - 1) / 2.0)
    v = resized[:, 1] / ((target_h - 1) / 2.0)
    return torch.stack([u, v], dim=1)
