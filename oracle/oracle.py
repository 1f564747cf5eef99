"""CPU oracle for the motion-compensation hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module. The product path (the CUDA
extension behind ``include/diffcodec_b200.h``) never does, and fails loudly when
its shared library is missing instead of falling back to anything here.

What is restated, and from where (paths relative to the upstream repository):

* ``splat_fwd / splat_ingrad / splat_flowgrad`` -> ``liboracle.so``
  (``softsplat_oracle.c``): the three kernels of ``controlnet/softsplat.py:284-335,
  368-423, 439-512`` in sequential canonical order.
* ``OracleSplatFunc``  -> ``softsplat_func`` (``controlnet/softsplat.py:277-528``).
* ``softsplat``        -> the mode wrapper ``controlnet/softsplat.py:232-274``.
* ``compute_mask / feature_warper / resize_and_normalize_flow_batched``
                       -> ``controlnet/control_utils.py:11-17, 49-72, 74-97``.
* ``residual_recipe``  -> ``controlnet/residual_utils.py:151-199`` and
                          ``controlnet/dataset.py:224-265``.
* ``backwarp``         -> ``cmp/models/modules/warp.py:9-25`` (calls torch's own CPU
                          ``grid_sample``, the reference's actual callee).

Parity pin: the reference has no tests or golden vectors for this path. The
restatement is pinned against the reference's *own kernel text*, templated by its
own ``cuda_kernel()`` and executed sequentially on the CPU, see
``oracle/ref_emulation.py`` and ``tests/golden/ref_emu_*.npz``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "softsplat_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []))
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_max_threads.restype = ctypes.c_int
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _suffix(dtype) -> str:
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _prep(*arrays):
    dt = arrays[0].dtype
    return [np.ascontiguousarray(a, dtype=dt) for a in arrays]


def splat_fwd(tin: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """Sum-splat of ``tin [N,C,H,W]`` by ``flow [N,2,H,W]`` (softsplat.py:290-335)."""
    tin, flow = _prep(tin, flow)
    n, c, h, w = tin.shape
    assert flow.shape == (n, 2, h, w)
    out = np.zeros_like(tin)
    getattr(lib(), "orc_splat_fwd_" + _suffix(tin.dtype))(_ptr(tin), _ptr(flow), _ptr(out), n, c, h, w)
    return out


def splat_fwd_mt(tin: np.ndarray, flow: np.ndarray, threads: int = 0) -> np.ndarray:
    """Frame-parallel float32 forward for the CPU baseline (bit-identical to splat_fwd)."""
    tin, flow = _prep(tin.astype(np.float32, copy=False), flow.astype(np.float32, copy=False))
    n, c, h, w = tin.shape
    out = np.zeros_like(tin)
    lib().orc_splat_fwd_f32_mt(_ptr(tin), _ptr(flow), _ptr(out), n, c, h, w, int(threads))
    return out


def max_threads() -> int:
    return int(lib().orc_max_threads())


def splat_ingrad(flow: np.ndarray, outgrad: np.ndarray) -> np.ndarray:
    """softsplat.py:376-423."""
    outgrad, flow = _prep(outgrad, flow)
    n, c, h, w = outgrad.shape
    ingrad = np.zeros_like(outgrad)
    getattr(lib(), "orc_splat_ingrad_" + _suffix(outgrad.dtype))(_ptr(flow), _ptr(outgrad), _ptr(ingrad), n, c, h, w)
    return ingrad


def splat_flowgrad(tin: np.ndarray, flow: np.ndarray, outgrad: np.ndarray) -> np.ndarray:
    """softsplat.py:447-512."""
    tin, flow, outgrad = _prep(tin, flow, outgrad)
    n, c, h, w = tin.shape
    flowgrad = np.zeros_like(flow)
    getattr(lib(), "orc_splat_flowgrad_" + _suffix(tin.dtype))(
        _ptr(tin), _ptr(flow), _ptr(outgrad), _ptr(flowgrad), n, c, h, w)
    return flowgrad


class OracleSplatFunc(torch.autograd.Function):
    """CPU stand-in for ``softsplat_func`` (controlnet/softsplat.py:277-528)."""

    @staticmethod
    def forward(ctx, tenIn, tenFlow):
        out = torch.from_numpy(splat_fwd(tenIn.detach().numpy(), tenFlow.detach().numpy()))
        ctx.save_for_backward(tenIn, tenFlow)
        return out

    @staticmethod
    def backward(ctx, tenOutgrad):
        tenIn, tenFlow = ctx.saved_tensors
        g = tenOutgrad.contiguous().numpy()
        gi = gf = None
        if ctx.needs_input_grad[0]:
            gi = torch.from_numpy(splat_ingrad(tenFlow.detach().numpy(), g))
        if ctx.needs_input_grad[1]:
            gf = torch.from_numpy(splat_flowgrad(tenIn.detach().numpy(), tenFlow.detach().numpy(), g))
        return gi, gf


def exp_det(t: torch.Tensor) -> torch.Tensor:
    """exp() as the library's DETERMINISTIC mode computes it (csrc/dcb_common.cuh exp_det: IEEE double operations only,
    rounded once to the tensor's type): what makes the deterministic `soft` mode bit-comparable across CPU and GPU.
    Within an ulp of torch's own exp (checked in tests/test_oracle_golden.py)."""
    a = np.ascontiguousarray(t.detach().numpy())
    out = np.empty_like(a)
    fn = getattr(lib(), "orc_exp_det_" + _suffix(a.dtype))
    fn(_ptr(a), _ptr(out), ctypes.c_longlong(a.size))
    return torch.from_numpy(out)


def softsplat(tenIn, tenFlow, tenMetric, strMode: str, exp_fn=None):
    """Mode wrapper, restating controlnet/softsplat.py:232-274 on CPU tensors. `exp_fn` (default: torch's exp, as the
    reference) lets the deterministic-mode tests substitute `exp_det`."""
    base = strMode.split("-")[0]
    assert base in ("sum", "avg", "linear", "soft")
    if strMode in ("sum", "avg"):
        assert tenMetric is None
    if base in ("linear", "soft"):
        assert tenMetric is not None

    if strMode == "avg":
        ones = tenIn.new_ones([tenIn.shape[0], 1, tenIn.shape[2], tenIn.shape[3]])
        tenIn = torch.cat([tenIn, ones], 1)
    elif base == "linear":
        tenIn = torch.cat([tenIn * tenMetric, tenMetric], 1)
    elif base == "soft":
        e = tenMetric.exp() if exp_fn is None else exp_fn(tenMetric)
        tenIn = torch.cat([tenIn * e, e], 1)

    tenOut = OracleSplatFunc.apply(tenIn, tenFlow)

    if base in ("avg", "linear", "soft"):
        norm = tenOut[:, -1:, :, :]
        parts = strMode.split("-")
        if len(parts) == 1 or parts[1] == "addeps":
            norm = norm + 0.0000001
        elif parts[1] == "zeroeps":
            norm = torch.where(norm == 0.0, torch.ones_like(norm), norm)  # values of :263, no aliasing
        elif parts[1] == "clipeps":
            norm = norm.clip(0.0000001, None)
        tenOut = tenOut[:, :-1, :, :] / norm
    return tenOut


def compute_mask(flow_a, flow_b):
    """controlnet/control_utils.py:11-17 -- first argument is splatted by the second."""
    metric = torch.ones_like(flow_b[:, :1])
    warped = softsplat(flow_a, flow_b, metric, "soft")
    diff = flow_b + warped
    return (torch.norm(diff, p=2, dim=1, keepdim=True) > 0.3).float()


def feature_warper(feat, flow, metric=None, mask=None):
    """controlnet/control_utils.py:49-72 with the metric supplied by the caller."""
    if metric is None:
        metric = torch.ones_like(flow[:, :1])
    warped = softsplat(feat, flow, metric, "soft")
    if mask is not None:
        warped = warped * (1 - mask)
    return warped, metric


def resize_and_normalize_flow_batched(flow, target_h: int, target_w: int):
    """controlnet/control_utils.py:74-97."""
    resized = torch.nn.functional.interpolate(flow, size=(target_h, target_w), mode="bilinear", align_corners=False)
    u = resized[:, 0] / ((target_w - 1) / 2.0)
    v = resized[:, 1] / ((target_h - 1) / 2.0)
    return torch.stack([u, v], dim=1)


def fuse(warped_a, warped_b, conf_a, conf_b, occ_a=None, occ_b=None):
    """Confidence fusion (+ double-hole fill), extractors.py:298-310 / residual_utils.py:181-193."""
    conf = torch.clamp(torch.cat([conf_a, conf_b], dim=1), min=0)
    w = conf / (conf.sum(dim=1, keepdim=True) + 1e-6)
    fused = w[:, :1] * warped_a + w[:, 1:] * warped_b
    if occ_a is not None:
        holes = (occ_a + occ_b) > 1.5
        if holes.any():
            fused = torch.where(holes.expand_as(fused), 0.5 * (warped_a + warped_b), fused)
    return fused


def soft_fuse(W_fwd, W_bwd, M_fwd, M_bwd, conf_fwd=None, conf_bwd=None, eps: float = 1e-6):
    """improv_experiments.ipynb cell 3: masks are VALID masks here (1 valid), holes where both are invalid."""
    conf_fwd = M_fwd if conf_fwd is None else conf_fwd
    conf_bwd = M_bwd if conf_bwd is None else conf_bwd
    w = torch.clamp(torch.cat([conf_fwd, conf_bwd], dim=1), min=0)
    w_norm = w / (w.sum(dim=1, keepdim=True) + eps)
    out = w_norm[:, :1] * W_fwd + w_norm[:, 1:] * W_bwd
    holes = (M_fwd + M_bwd) < 0.5
    if holes.any():
        out = torch.where(holes.expand_as(out), 0.5 * (W_fwd + W_bwd), out)
    return out


def pyramid_conditioning(img1, img2, flow1, flow2, sizes=(128, 64, 32)):
    """improv_experiments.ipynb cell 5, restated on CPU tensors: per size both frames and
    both flows resized (bilinear, align_corners=False; flows times size / W), both frames soft-splatted with an all-ones
    metric, fused with identity masks. Returns [(warped1, warped2, fused), ...]."""
    F = torch.nn.functional
    out = []
    for size in sizes:
        a = F.interpolate(img1, size=(size, size), mode="bilinear", align_corners=False)
        b = F.interpolate(img2, size=(size, size), mode="bilinear", align_corners=False)
        k = float(size) / float(flow1.shape[-1])
        f1 = F.interpolate(flow1, size=(size, size), mode="bilinear", align_corners=False) * k
        f2 = F.interpolate(flow2, size=(size, size), mode="bilinear", align_corners=False) * k
        ones = torch.ones(img1.shape[0], 1, size, size, dtype=img1.dtype)
        w1 = softsplat(a, f1, ones, "soft")
        w2 = softsplat(b, f2, ones, "soft")
        out.append((w1, w2, soft_fuse(w1, w2, ones, ones)))
    return out


def residual_recipe(image1, flow1, flow2, gt, variant: str):
    """Conditioning builder.

    variant "dataset": controlnet/dataset.py:233-265 (occlusion masks are the fusion weights,
    no hole branch). variant "wrapper": controlnet/residual_utils.py:159-199 (ones metrics
    are the weights, double-hole average fill). Both splat image1 by flow1 twice (Appendix
    B-6 of SURVEY.md) -- restated as the reference computes it, not as it was intended.
    Returns (fused, residual, occ_fwd, occ_bwd).
    """
    metric = torch.ones_like(flow1[:, :1])
    warped1 = softsplat(image1, flow1, metric, "soft")
    warped2 = softsplat(image1, flow1, metric, "soft")
    occ_fwd = compute_mask(flow1, flow2)
    occ_bwd = compute_mask(flow2, flow1)
    if variant == "dataset":
        fused = fuse(warped1, warped2, occ_fwd, occ_bwd)
    elif variant == "wrapper":
        fused = fuse(warped1, warped2, metric, metric, occ_fwd, occ_bwd)
    else:
        raise ValueError(variant)
    return fused, gt - fused, occ_fwd, occ_bwd


def backwarp(image, flow, align_corners=None):
    """cmp/models/modules/warp.py:9-25 on CPU tensors (torch's own grid_sample)."""
    fg = torch.zeros_like(flow)
    fg[:, 0] = flow[:, 0] / ((flow.size(3) - 1.0) / 2.0)
    fg[:, 1] = flow[:, 1] / ((flow.size(2) - 1.0) / 2.0)
    n, _, h, w = image.shape
    hor = torch.linspace(-1.0, 1.0, w, dtype=image.dtype).view(1, 1, 1, w).expand(n, 1, h, w)
    ver = torch.linspace(-1.0, 1.0, h, dtype=image.dtype).view(1, 1, h, 1).expand(n, 1, h, w)
    grid = (torch.cat([hor, ver], 1) + fg).permute(0, 2, 3, 1)
    if align_corners is None:
        return torch.nn.functional.grid_sample(image, grid)  # as executed: default False
    return torch.nn.functional.grid_sample(image, grid, align_corners=bool(align_corners))


# ---------------------------------------------------------------------------------------------
# f-3: latent tile merge (patch_utils.py:83-174) and tile cropping (patch_utils.py:189-209)
# ---------------------------------------------------------------------------------------------
def _hann_2d(h: int, w: int, dtype):
    """patch_utils.py:121-133: outer product of two non-periodic Hann windows, max-normalised."""
    wy = torch.ones(1, dtype=dtype) if h <= 1 else torch.hann_window(h, periodic=False, dtype=dtype)
    wx = torch.ones(1, dtype=dtype) if w <= 1 else torch.hann_window(w, periodic=False, dtype=dtype)
    m = wy.unsqueeze(1) * wx.unsqueeze(0)
    return m / (m.max() + 1e-12)


def tile_rects(pixel_coords, full_latent_shape, original_image_size):
    """patch_utils.py:135-154. The coordinate tuple is unpacked as (x1, x2, y1, y2) although
    crop_into_tiles() produces (y1, y2, x1, x2): positions 2,3 are scaled by the HEIGHT ratio and
    positions 0,1 by the WIDTH ratio, as executed. Returns (ly1, ly2, lx1, lx2) per tile, clamped;
    Python's round() is round-half-to-even."""
    _, _, h_lat, w_lat = full_latent_shape
    h_px, w_px = original_image_size
    rects = []
    for (p0, p1, p2, p3) in pixel_coords:
        ly1 = int(round(p2 * (h_lat / float(h_px)))); ly2 = int(round(p3 * (h_lat / float(h_px))))
        lx1 = int(round(p0 * (w_lat / float(w_px)))); lx2 = int(round(p1 * (w_lat / float(w_px))))
        ly1 = max(0, min(ly1, h_lat)); ly2 = max(0, min(ly2, h_lat))
        lx1 = max(0, min(lx1, w_lat)); lx2 = max(0, min(lx2, w_lat))
        rects.append((ly1, ly2, lx1, lx2))
    return rects


def merge_latent_tiles_from_pixel_coords(latents, pixel_coords, full_latent_shape, original_image_size, eps: float = 1e-8):
    """Restatement of patch_utils.py:83-174 with torch CPU ops in the reference's order: per tile
    (list order) optional bilinear resize (align_corners=False) to its latent rectangle, Hann
    mask, out += tile * mask, weight += mask; merged = out / max(weight, eps)."""
    assert len(latents) == len(pixel_coords), "latents and coords length mismatch"
    dtype = latents[0].dtype
    out = torch.zeros(full_latent_shape, dtype=dtype)
    weight = torch.zeros_like(out)
    for tile, (ly1, ly2, lx1, lx2) in zip(latents, tile_rects(pixel_coords, full_latent_shape, original_image_size)):
        th_, tw_ = ly2 - ly1, lx2 - lx1
        if th_ <= 0 or tw_ <= 0:
            continue                                              # patch_utils.py:149-151
        assert tile.dim() == 4 and tile.size(0) == 1, "expected tile shape (1,C,H,W)"
        if tile.shape[-2] != th_ or tile.shape[-1] != tw_:
            tile = torch.nn.functional.interpolate(tile, size=(th_, tw_), mode="bilinear", align_corners=False)
        mask = _hann_2d(th_, tw_, dtype).unsqueeze(0).unsqueeze(0).expand(1, tile.size(1), th_, tw_)
        out[:, :, ly1:ly2, lx1:lx2] += tile * mask
        weight[:, :, ly1:ly2, lx1:lx2] += mask
    return out / torch.maximum(weight, torch.tensor(eps, dtype=dtype))


def crop_into_tiles(img, tile_size, overlap: int = 0, order: str = "hwc"):
    """patch_utils.py:189-209: row-major overlapping crops; coords are (y, y2, x, x2)."""
    h, w = (img.shape[0], img.shape[1]) if order == "hwc" else (img.shape[1], img.shape[2])
    sy, sx = tile_size[0] - overlap, tile_size[1] - overlap
    tiles, coords = [], []
    for y in range(0, h, sy):
        for x in range(0, w, sx):
            y2, x2 = min(y + tile_size[0], h), min(x + tile_size[1], w)
            tiles.append(img[y:y2, x:x2, :] if order == "hwc" else img[:, y:y2, x:x2])
            coords.append((y, y2, x, x2))
    return tiles, coords, (h, w)
