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

Two families of entry points:

* host helpers with the reference's names and signatures (``read_flo``, ``resize_flow_to``, ``fast_downsample_flow``,
  ``load_flow_cached``): plain NumPy / torch on the host, for callers that want exactly the reference's objects;
* the DEVICE path (``flo_to_device``, ``resize_flow_device``, ``resize_and_normalize_flow_device``): the raw ``.flo``
  payload is uploaded as it is (pinned staging, one cudaMemcpyAsync) and ONE kernel of the native library
  (``dcb_flow_resize``, csrc/flow_ingest.cu) de-interleaves (or applies the planar quirk -- only the strides of the view
  differ), resamples with the chosen convention and writes the planar ``[N,2,h,w]`` tensor the splat kernels take.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

__all__ = ["FLO_MAGIC", "read_flo", "write_flo", "resize_flow_to", "fast_downsample_flow", "load_flow_cached",
           "flo_to_device", "resize_flow_device", "resize_and_normalize_flow_device"]

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


# -------------------------------------------------------------------------------------------------
# device path
# -------------------------------------------------------------------------------------------------
_CONVENTIONS = {"bilinear_rescale": 0, "adaptive_avg": 1, "bilinear_normalize": 2}


def resize_flow_device(flow: torch.Tensor, target_h: int, target_w: int, convention: str = "bilinear_rescale",
                       out_dtype: torch.dtype | None = None) -> torch.Tensor:
    """Resample a CUDA flow field ``[N,2,H,W]`` (any strides) to ``[N,2,target_h,target_w]`` in one kernel.

    convention: ``bilinear_rescale`` = ``resize_flow_to`` (utils.py:21-28), ``adaptive_avg`` = ``fast_downsample_flow``
    (dataset.py:43-50), ``bilinear_normalize`` = ``resize_and_normalize_flow_batched`` (control_utils.py:74-97)."""
    from . import _lib
    assert flow.is_cuda and flow.dim() == 4 and flow.shape[1] == 2, "resize_flow_device expects a CUDA [N,2,H,W] flow"
    assert flow.dtype in (torch.float32, torch.bfloat16)
    out = torch.empty((flow.shape[0], 2, int(target_h), int(target_w)), dtype=out_dtype or flow.dtype, device=flow.device)
    with _lib.on_device(flow.device):
        rc = _lib.lib().dcb_flow_resize(_lib.desc(flow), _lib.desc(out), _CONVENTIONS[convention], _lib.stream_ptr(flow.device))
    _lib.check(rc, "dcb_flow_resize")
    return out


def resize_and_normalize_flow_device(flow: torch.Tensor, target_h: int, target_w: int) -> torch.Tensor:
    """``resize_and_normalize_flow_batched`` (control_utils.py:74-97) as one kernel instead of interpolate + 2 divisions + stack."""
    return resize_flow_device(flow, target_h, target_w, "bilinear_normalize")


def flo_to_device(path_or_bytes, device="cuda", target_hw=None, convention: str = "bilinear_rescale", planar_quirk: bool = False,
                  out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Middlebury ``.flo`` -> CUDA ``[1,2,h,w]``. The header is parsed on the host (12 bytes); the payload is uploaded RAW and
    read on the device through a strided view: interleaved (u, v) per pixel as ``utils.py:10-19`` reads it, or -- with
    ``planar_quirk`` -- as the dataset's ``np.resize(data, (2, h, w))`` mis-reshape (``dataset.py:15-24``). ``target_hw=None``
    keeps the resolution (a pure de-interleave)."""
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        blob = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            blob = f.read()
    magic = np.frombuffer(blob[:4], np.float32, 1)[0]
    if magic != np.float32(FLO_MAGIC):
        raise ValueError(f"Invalid .flo data (magic={magic})")
    w, h = (int(v) for v in np.frombuffer(blob[4:12], np.int32, 2))
    payload = torch.frombuffer(bytearray(blob[12:12 + 8 * w * h]), dtype=torch.float32)
    dev = torch.device(device)
    raw = payload.pin_memory().to(dev, non_blocking=True)
    if planar_quirk:
        view = raw.as_strided((1, 2, h, w), (2 * h * w, h * w, w, 1))
    else:
        view = raw.as_strided((1, 2, h, w), (2 * h * w, 1, 2 * w, 2))
    th, tw = (h, w) if target_hw is None else target_hw
    if target_hw is None:
        # same resolution: bilinear with align_corners=True at scale 1 is the identity, times (1, 1)
        return resize_flow_device(view, h, w, "bilinear_rescale", out_dtype)
    return resize_flow_device(view, th, tw, convention, out_dtype)
