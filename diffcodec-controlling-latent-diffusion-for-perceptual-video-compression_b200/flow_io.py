"""Flow ingest -- the step immediately before the hot path.

File formats and resize conventions the reference feeds into the splat kernels, so that the
units reaching ``softsplat`` are the ones the reference produces:

* Middlebury ``.flo``: float32 magic 202021.25, int32 width, int32 height, then ``h*w*2`` float32
  values INTERLEAVED as (u, v) per pixel (reference reader ``controlnet/utils.py:10-19``).
  ``controlnet/dataset.py:15-24`` reshapes the same payload as planar ``(2,h,w)`` with
  ``np.resize`` -- a bug that scrambles u and v; ``read_flo(..., planar_quirk=True)``
  reproduces it for anyone who needs bit-compatibility with checkpoints trained that way.
* ``.npy`` cache next to the ``.flo`` (``dataset.py:52-59``).
* ``resize_flow_to``: bilinear, ``align_corners=True``, vectors RESCALED to the new resolution
  (``controlnet/utils.py:21-28``).
* ``fast_downsample_flow``: adaptive average pooling, vectors NOT rescaled (``dataset.py:43-50``).

Plain NumPy / torch on the host; no arithmetic of the hot path happens here.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

__all__ = ["FLO_MAGIC", "read_flo", "write_flo", "resize_flow_to", "fast_downsample_flow", "load_flow_cached"]

FLO_MAGIC = 202021.25


def read_flo(path: str, planar_quirk: bool = False) -> np.ndarray:
    """Returns [H,W,2] float32 in pixel units; with ``planar_quirk`` the reference dataset's [2,H,W] mis-reshape."""
    with open(path, "rb") as f:
        magic = np.frombuffer(f.read(4), np.float32, 1)[0]
        if magic != np.float32(FLO_MAGIC):
            raise ValueError(f"Invalid .flo file: {path} (magic={magic})")
        w, h = (int(v) for v in np.frombuffer(f.read(8), np.int32, 2))
        data = np.frombuffer(f.read(8 * w * h), np.float32, 2 * w * h)
    if planar_quirk:
        return np.resize(data, (2, h, w)).copy()
    return data.reshape(h, w, 2).copy()


def write_flo(path: str, flow_hw2: np.ndarray) -> None:
    h, w, two = flow_hw2.shape
    assert two == 2
    with open(path, "wb") as f:
        np.array([FLO_MAGIC], np.float32).tofile(f)
        np.array([w, h], np.int32).tofile(f)
        np.ascontiguousarray(flow_hw2, np.float32).tofile(f)


def resize_flow_to(flow_hw2: np.ndarray, target_h: int, target_w: int) -> torch.Tensor:
    """[H,W,2] -> [1,2,target_h,target_w], bilinear (align_corners=True), vectors scaled to stay in pixel units."""
    ft = torch.from_numpy(np.ascontiguousarray(flow_hw2)).permute(2, 0, 1).unsqueeze(0)
    _, _, h, w = ft.shape
    ft = F.interpolate(ft, size=(target_h, target_w), mode="bilinear", align_corners=True)
    ft[:, 0] *= target_w / max(w, 1)
    ft[:, 1] *= target_h / max(h, 1)
    return ft


def fast_downsample_flow(flow, target_h: int = 128, target_w: int = 128) -> np.ndarray:
    """[2,H,W] (or [1,2,H,W]) -> [2,target_h,target_w] by adaptive average pooling; magnitudes are NOT rescaled."""
    if isinstance(flow, np.ndarray):
        flow = torch.from_numpy(flow)
    if flow.ndim == 3:
        flow = flow.unsqueeze(0)
    return F.adaptive_avg_pool2d(flow.float(), (target_h, target_w)).squeeze(0).numpy()


def load_flow_cached(path, target_h: int = 128, target_w: int = 128, planar_quirk: bool = True) -> np.ndarray:
    """``.npy`` cache if present, else the ``.flo`` read the way the reference dataset reads it (planar quirk)."""
    npy = str(path).replace(".flo", ".npy")
    if os.path.exists(npy):
        flow = np.load(npy)
    else:
        flow = read_flo(str(path), planar_quirk=planar_quirk)
        if not planar_quirk:
            flow = flow.transpose(2, 0, 1)
    return fast_downsample_flow(flow, target_h, target_w)
