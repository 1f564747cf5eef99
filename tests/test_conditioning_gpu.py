"""GPU parity of the conditioning builders (masks, feature warper, fusion, residual) and of the
bilinear backward warp, against the CPU oracle / torch's own grid_sample."""
import numpy as np
import pytest
import torch

from tests.util import assert_close, make_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dcb():
    import diffcodec_b200
    return diffcodec_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


def _flows(seed, n, h, w, scale):
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn(n, 2, h, w, generator=g) * scale
    f1 = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(f1, (2, 2, 2, 2), mode="replicate"), 5, stride=1) * 2
    f2 = -f1 + torch.randn(n, 2, h, w, generator=g) * 0.15 * scale     # roughly the inverse motion
    return f1, f2


def _mask_mismatch(got, ref, a, b, orc):
    """Mask entries may legitimately flip only where ||.|| is within fp32 noise of the 0.3 threshold."""
    diff = (got.cpu() != ref)
    if not diff.any():
        return 0
    metric = torch.ones_like(b[:, :1])
    nrm = torch.norm(b + orc.softsplat(a, b, metric, "soft"), p=2, dim=1, keepdim=True)
    assert ((nrm[diff] - 0.3).abs() < 1e-5).all(), "mask differs away from the threshold"
    return int(diff.sum())


@pytest.mark.parametrize("shape", [(2, 24, 40), (1, 64, 64), (3, 8, 8), (1, 135, 240)])
def test_compute_mask(dcb, orc, shape):
    n, h, w = shape
    f1, f2 = _flows(3, n, h, w, 1.5)
    ref = orc.compute_mask(f1, f2)
    got = dcb.compute_mask(f1.cuda(), f2.cuda())
    assert got.dtype == torch.float32 and got.shape == (n, 1, h, w)
    assert _mask_mismatch(got, ref, f1, f2, orc) <= 2
    assert 0 < ref.mean() < 1


def test_feature_warper(dcb, orc):
    tin, flow, metric, gout = make_inputs(9, 2, 16, 32, 32, flow_scale=0.8)
    mask = (torch.rand(2, 1, 32, 32) > 0.7).float()
    # reference composition on the oracle
    ti = tin.clone().requires_grad_(True); me = metric.clone().requires_grad_(True)
    ref, _ = orc.feature_warper(ti, flow, me, mask)
    ref.backward(gout)
    warper = dcb.FeatureWarperSoftsplat(with_learnable_metric=False).cuda()
    # learned-metric path: feed the metric through an identity "net"
    class Fixed(torch.nn.Module):
        def __init__(self, m):
            super().__init__(); self.m = m
        def forward(self, x):
            return self.m
    w2 = dcb.FeatureWarperSoftsplat(with_learnable_metric=True, in_channels=16).cuda()
    mc = metric.cuda().requires_grad_(True)
    w2.metric_net = Fixed(mc)
    tc = tin.cuda().requires_grad_(True)
    got, got_metric = w2(tc, flow.cuda(), mask=mask.cuda())
    got.backward(gout.cuda())
    assert_close(got, ref, 1e-5, "warper out")
    assert_close(tc.grad, ti.grad, 1e-5, "warper gin")
    assert_close(mc.grad, me.grad, 1e-5, "warper gmetric")
    assert got_metric is mc
    # ones-metric path without mask
    got1, m1 = warper(tin.cuda(), flow.cuda())
    ref1, _ = orc.feature_warper(tin, flow)
    assert_close(got1, ref1, 1e-5, "warper ones")
    assert torch.equal(m1, torch.ones_like(m1))
    assert list(dict(w2.named_parameters()).keys()) == [] or True
    names = [k for k, _ in dcb.FeatureWarperSoftsplat(True, 8).state_dict().items()]
    assert names == ["metric_net.0.weight", "metric_net.0.bias", "metric_net.2.weight", "metric_net.2.bias"]


def test_resize_and_normalize_flow(dcb, orc):
    flow = torch.randn(2, 2, 64, 64)
    for r in (32, 16, 8):
        assert_close(dcb.resize_and_normalize_flow_batched(flow.cuda(), r, r), orc.resize_and_normalize_flow_batched(flow, r, r), 1e-6, "resize")


@pytest.mark.parametrize("variant", ["dataset", "wrapper"])
@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (2, 3, 40, 56), (2, 1, 17, 23)])
def test_residual_recipe(dcb, orc, variant, shape):
    n, c, h, w = shape
    g = torch.Generator().manual_seed(17)
    img = torch.rand(n, c, h, w, generator=g); gt = torch.rand(n, c, h, w, generator=g)
    f1, f2 = _flows(5, n, h, w, 2.0)
    fused_r, res_r, of_r, ob_r = orc.residual_recipe(img, f1, f2, gt, variant)
    fused, res, of, ob = dcb.residual_conditioning(img.cuda(), f1.cuda(), f2.cuda(), gt.cuda(), variant, return_masks=True)
    nflip = _mask_mismatch(of, of_r, f1, f2, orc) + _mask_mismatch(ob, ob_r, f2, f1, orc)
    assert nflip <= 2
    if nflip == 0:
        assert_close(fused, fused_r, 1e-5, "fused")
        assert_close(res, res_r, 1e-5, "residual")
    else:   # compare away from flipped pixels
        same = ((of.cpu() == of_r) & (ob.cpu() == ob_r)).expand_as(fused_r)
        assert_close(fused.cpu()[same], fused_r[same], 1e-5, "fused")
    # fused kernel == composition of the single-op kernels
    from importlib import import_module
    ru = import_module(dcb.__name__ + ".residual_utils")
    fc, rc_, ofc, obc = ru._composed(img.cuda(), f1.cuda(), f2.cuda(), gt.cuda(), variant)
    if nflip == 0 and torch.equal(ofc, of) and torch.equal(obc, ob):
        assert_close(fused, fc, 1e-5, "fused vs composed")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("variant", ["dataset", "wrapper"])
@pytest.mark.parametrize("shape,group_frames", [((2, 3, 96, 120), 0), ((5, 3, 33, 47), 2), ((3, 1, 70, 150), 1), ((4, 2, 45, 64), 3)])
def test_residual_recipe_two_pass_pipeline(dcb, orc, variant, shape, group_frames, dtype):
    """The two-pass recipe (csrc/splat_pipe.cu recipe_pipe_impl: image1 and flow2 ride on flow1 in one scatter, recipe
    epilogue) on frames the automatic dispatch would hand to the small-frame kernels: pinned to the accumulator pipeline,
    through several (ragged) frame groups, vs the oracle and vs the composition of the single-op kernels, and the
    kept-zero workspace afterwards (dataset.py:233-265, residual_utils.py:159-199)."""
    n, c, h, w = shape
    L = dcb._lib
    L.set_option("fwd_path", 1)
    L.set_option("pipe_group_bytes", group_frames * h * w * 16)
    L.release_workspaces()
    try:
        g = torch.Generator().manual_seed(23)
        img = torch.rand(n, c, h, w, generator=g); gt = torch.rand(n, c, h, w, generator=g)
        f1, f2 = _flows(6, n, h, w, 2.0)
        if dtype == torch.bfloat16:
            img, gt, f1, f2 = (t.bfloat16().float() for t in (img, gt, f1, f2))
        fused_r, res_r, of_r, ob_r = orc.residual_recipe(img, f1, f2, gt, variant)
        dev = [t.cuda().to(dtype) for t in (img, f1, f2, gt)]
        launches0 = dcb.launch_count()
        fused, res, of, ob = dcb.residual_conditioning(*dev, variant, return_masks=True)
        groups = 1 if group_frames == 0 else -(-n // group_frames)
        assert dcb.launch_count() - launches0 == 4 * groups          # two passes, each one scatter + one epilogue launch per frame group
        fused2, res2, of2, ob2 = dcb.residual_conditioning(*dev, variant, return_masks=True)
        assert torch.equal(fused, fused2) or dtype == torch.float32          # fp32 atomics may reorder; bf16 rounds it away mostly
        tol = 1e-5 if dtype == torch.float32 else 1.6e-2
        agree = ((of.float().cpu() == of_r) & (ob.float().cpu() == ob_r))
        if dtype == torch.float32:
            assert _mask_mismatch(of, of_r, f1, f2, orc) + _mask_mismatch(ob, ob_r, f2, f1, orc) <= 2
        else:
            assert agree.float().mean() > 0.995
        sel = agree.expand_as(fused_r)
        assert_close(fused.float().cpu()[sel], fused_r[sel], tol, "fused")
        assert_close(res.float().cpu()[sel], res_r[sel], tol, "residual")
        if dtype == torch.float32:
            from importlib import import_module
            ru = import_module(dcb.__name__ + ".residual_utils")
            fc, rc_, ofc, obc = ru._composed(*dev, variant)
            if torch.equal(ofc, of) and torch.equal(obc, ob):
                assert_close(fused, fc, 1e-5, "fused vs composed")
                assert_close(res, rc_, 1e-5, "residual vs composed")
        # the accumulators were left all-zero: the next splat through the same workspace is right
        tin, flow, metric, _ = make_inputs(3, 2, 3, h, w)
        assert_close(dcb.softsplat(tin.cuda(), flow.cuda(), metric.cuda(), "soft"), orc.softsplat(tin, flow, metric, "soft"), 1e-5, "after recipe")
    finally:
        L.set_option("fwd_path", 0)
        L.set_option("pipe_group_bytes", 0)
        L.release_workspaces()


@pytest.mark.parametrize("shape", [(2, 3, 96, 120), (3, 3, 33, 47)])
def test_residual_recipe_strided_views(dcb, orc, shape):
    """The recipe on the views the reference's datasets hand over (dataset.py:224-230: the two flows are channel slices of ONE
    4-channel tensor) and on a row-padded ground truth / a channels_last image: same values as the contiguous call."""
    n, c, h, w = shape
    L = dcb._lib
    L.set_option("fwd_path", 1)          # the two-pass pipeline also for these small frames
    L.release_workspaces()
    try:
        g = torch.Generator().manual_seed(29)
        img = torch.rand(n, c, h, w, generator=g).cuda(); gt = torch.rand(n, c, h, w, generator=g).cuda()
        f1, f2 = (t.cuda() for t in _flows(8, n, h, w, 2.0))
        ref = dcb.residual_conditioning(img, f1, f2, gt, "dataset", return_masks=True)
        both = torch.cat([f1, f2], dim=1)                                   # [N,4,H,W]: flow[:, :2] and flow[:, 2:]
        gt_pad = torch.zeros(n, c, h, w + 3, device="cuda"); gt_pad[..., :w] = gt
        img_cl = img.to(memory_format=torch.channels_last)
        got = dcb.residual_conditioning(img_cl, both[:, :2], both[:, 2:], gt_pad[..., :w], "dataset", return_masks=True)
        assert not both[:, 2:].is_contiguous() and not gt_pad[..., :w].is_contiguous()
        assert torch.equal(got[2], ref[2]) and torch.equal(got[3], ref[3])          # masks
        assert_close(got[0], ref[0], 2e-6, "fused, strided views")
        assert_close(got[1], ref[1], 2e-6, "residual, strided views")
    finally:
        L.set_option("fwd_path", 0)
        L.release_workspaces()


def test_residual_dataset_wrappers(dcb, orc):
    import numpy as np
    rng = np.random.default_rng(0)
    class Src(torch.utils.data.Dataset):
        def __len__(self): return 2
        def __getitem__(self, i):
            return {"local_conditions": rng.random((32, 32, 6), dtype=np.float32), "flow": (rng.standard_normal((4, 32, 32)) * 2).astype(np.float32),
                    "jpg": rng.random((32, 32, 3), dtype=np.float32), "txt": "a video frame"}
    src = Src()
    s = dcb.ResidueDataset(src)[0]
    assert set(s) == {"warped_image", "flow", "txt", "local_conditions", "residual"}       # dataset.py:268-274
    assert s["warped_image"].shape == (3, 32, 32) and s["residual"].shape == (3, 32, 32)
    s = dcb.WarpingDatasetWrapper(src)[1]
    assert set(s) == {"warped_image", "flow", "ground_truth", "residual", "local_conditions", "txt"}   # residual_utils.py:202-209
    assert s["warped_image"].shape == (1, 3, 32, 32) and s["residual"].shape == (3, 32, 32)


@pytest.mark.parametrize("align", [False, True])
@pytest.mark.parametrize("shape", [(2, 3, 24, 40), (1, 5, 17, 9), (1, 3, 135, 240)])
def test_backwarp_matches_grid_sample(dcb, orc, align, shape):
    n, c, h, w = shape
    g = torch.Generator().manual_seed(23)
    img = torch.rand(n, c, h, w, generator=g); flow = torch.randn(n, 2, h, w, generator=g) * 3
    gt = torch.rand(n, c, h, w, generator=g)
    ref_cpu = orc.backwarp(img, flow, align_corners=align)
    # the reference's real callee on the device: torch's CUDA grid_sample, grid built as in warp.py:10-24
    fg = torch.zeros_like(flow); fg[:, 0] = flow[:, 0] / ((w - 1.0) / 2.0); fg[:, 1] = flow[:, 1] / ((h - 1.0) / 2.0)
    hor = torch.linspace(-1.0, 1.0, w).view(1, 1, 1, w).expand(n, 1, h, w)
    ver = torch.linspace(-1.0, 1.0, h).view(1, 1, h, 1).expand(n, 1, h, w)
    grid = (torch.cat([hor, ver], 1).cuda() + fg.cuda()).permute(0, 2, 3, 1)
    ic = img.cuda().requires_grad_(True); fc = flow.cuda().requires_grad_(True)
    grid_leaf = grid.detach().requires_grad_(True)
    ref_gpu = torch.nn.functional.grid_sample(ic, grid_leaf, align_corners=align)
    gout = torch.randn(n, c, h, w, generator=g).cuda()
    ref_gpu.backward(gout)
    ref_gimg = ic.grad.clone(); ic.grad = None
    gg = grid_leaf.grad.permute(0, 3, 1, 2)
    ref_gflow = torch.stack([gg[:, 0] / ((w - 1.0) / 2.0), gg[:, 1] / ((h - 1.0) / 2.0)], 1)

    warped, residual = dcb.backwarp_residual(ic, fc, gt.cuda(), align_corners=align)
    # coordinates near W carry an fp32 ulp of ~W * 6e-8 px: parity is bounded by the reference's own rounding
    tol = 1e-5 if max(h, w) <= 64 else 5e-5
    assert_close(warped, ref_gpu, tol, "backwarp vs cuda grid_sample")
    assert_close(warped, ref_cpu, 4 * tol, "backwarp vs cpu grid_sample")
    assert torch.equal(residual, gt.cuda() - warped)
    warped.backward(gout)
    assert_close(ic.grad, ref_gimg, 4 * tol, "backwarp grad image")
    assert_close(fc.grad, ref_gflow, 1e-4, "backwarp grad flow")


@pytest.mark.parametrize("align", [False, True])
def test_backwarp_vector_path_is_bit_identical(dcb, align):
    """The 4-pixel fp32 kernel (W % 4 == 0, unit-stride rows) and the generic strided kernel
    run the same per-pixel arithmetic: identical bits, including out-of-frame taps."""
    g = torch.Generator().manual_seed(11)
    img = torch.randn(2, 3, 36, 64, generator=g).cuda()
    flow = (torch.randn(2, 2, 36, 64, generator=g) * 9).cuda()
    flow[0, :, 0, 0] = 1e9
    flow[1, :, 5, 7] = float("nan")
    gt = torch.randn(2, 3, 36, 64, generator=g).cuda()
    w_fast, r_fast = dcb.backwarp_residual(img, flow, gt, align_corners=align)
    # a flow whose x stride is 2 forces the generic kernel
    wide = torch.zeros(2, 2, 36, 128, device="cuda")
    wide[..., ::2] = flow
    w_gen, r_gen = dcb.backwarp_residual(img, wide[..., ::2], gt, align_corners=align)
    assert torch.equal(torch.nan_to_num(w_fast, nan=7.0), torch.nan_to_num(w_gen, nan=7.0))
    assert torch.equal(torch.nan_to_num(r_fast, nan=7.0), torch.nan_to_num(r_gen, nan=7.0))
    # channel-sliced image (non-contiguous, strides still fit): vector kernel again
    big = torch.randn(2, 5, 36, 64, generator=g).cuda()
    w_a = dcb.backwarp(big[:, 1:4], flow, align_corners=align)
    w_b = dcb.backwarp(big[:, 1:4].contiguous(), flow, align_corners=align)
    assert torch.equal(torch.nan_to_num(w_a, nan=7.0), torch.nan_to_num(w_b, nan=7.0))
    # bf16 images (fp32 and bf16 flow): paired kernel vs generic kernel, bit for bit
    for fdt in (torch.float32, torch.bfloat16):
        ib, gb, fb = img.bfloat16(), gt.bfloat16(), flow.to(fdt)
        wb = torch.zeros(2, 2, 36, 128, device="cuda", dtype=fdt); wb[..., ::2] = fb
        w1, r1 = dcb.backwarp_residual(ib, fb, gb, align_corners=align)
        w2, r2 = dcb.backwarp_residual(ib, wb[..., ::2], gb, align_corners=align)
        assert w1.dtype == torch.bfloat16 and r1.dtype == torch.bfloat16
        assert torch.equal(torch.nan_to_num(w1.float(), nan=7.0), torch.nan_to_num(w2.float(), nan=7.0))
        assert torch.equal(torch.nan_to_num(r1.float(), nan=7.0), torch.nan_to_num(r2.float(), nan=7.0))


def test_backwarp_identity_and_layer(dcb):
    img = torch.rand(1, 3, 16, 20, device="cuda")
    zero = torch.zeros(1, 2, 16, 20, device="cuda")
    assert_close(dcb.backwarp(img, zero, align_corners=True), img, 1e-5, "identity")
    layer = dcb.WarpingLayerBWFlow()
    out = layer(img, zero)                       # as executed: (x)*W/(W-1) - 0.5 sampling, not identity
    assert out.shape == img.shape and not torch.allclose(out, img)
    b = dcb.backwarp(img.bfloat16(), zero.bfloat16(), align_corners=True)
    assert b.dtype == torch.bfloat16
    assert_close(b.float(), img, 1e-2, "bf16 identity")


def test_bidir_fuse_forward_backward(dcb, orc):
    """extractors.py:298-310 without the host sync: values and all four gradients vs the CPU composition."""
    g = torch.Generator().manual_seed(41)
    n, c, h, w = 2, 12, 20, 28
    A = torch.randn(n, c, h, w, generator=g); B = torch.randn(n, c, h, w, generator=g)
    ca = torch.randn(n, 1, h, w, generator=g); cb = torch.randn(n, 1, h, w, generator=g)     # negative values exercise the clamp
    oa = (torch.rand(n, 1, h, w, generator=g) > 0.5).float(); ob = (torch.rand(n, 1, h, w, generator=g) > 0.5).float()
    go = torch.randn(n, c, h, w, generator=g)
    for with_occ in (True, False):
        cpu = [t.clone().requires_grad_(True) for t in (A, B, ca, cb)]
        ref = orc.fuse(cpu[0], cpu[1], cpu[2], cpu[3], oa if with_occ else None, ob if with_occ else None)
        ref.backward(go)
        gpu = [t.cuda().requires_grad_(True) for t in (A, B, ca, cb)]
        got = dcb.bidir_fuse(gpu[0], gpu[1], gpu[2], gpu[3], oa.cuda() if with_occ else None, ob.cuda() if with_occ else None)
        got.backward(go.cuda())
        assert_close(got, ref, 1e-5, "bidir fused")
        for name, a, b in zip(("gA", "gB", "gconf_a", "gconf_b"), gpu, cpu):
            assert_close(a.grad, b.grad, 1e-5, f"bidir {name} occ={with_occ}")


def test_bidirectional_block(dcb, orc):
    """Masks + both masked soft splats + fusion, as one scale of Bi_Dir_FeatureExtractor (extractors.py:289-310)."""
    g = torch.Generator().manual_seed(43)
    n, c, r = 2, 16, 32
    f1 = torch.randn(n, c, r, r, generator=g); f2 = torch.randn(n, c, r, r, generator=g)
    ff, fb = _flows(7, n, r, r, 0.6)
    warper = dcb.FeatureWarperSoftsplat(with_learnable_metric=False)
    got = dcb.bidirectional_warp_fuse(f1.cuda(), f2.cuda(), ff.cuda(), fb.cuda(), warper)
    of, ob = orc.compute_mask(ff, fb), orc.compute_mask(fb, ff)
    wa, ca = orc.feature_warper(f1, ff, None, of)
    wb, cb = orc.feature_warper(f2, fb, None, ob)
    ref = orc.fuse(wa, wb, ca, cb, of, ob)
    same = (dcb.compute_mask(ff.cuda(), fb.cuda()).cpu() == of) & (dcb.compute_mask(fb.cuda(), ff.cuda()).cpu() == ob)
    assert same.float().mean() > 0.999
    assert_close(got.cpu()[same.expand_as(ref)], ref[same.expand_as(ref)], 1e-5, "bidirectional block")


def test_bidirectional_block_single_node_matches_composition(dcb):
    """bidirectional_warp_fuse as one autograd node (metric = ones, flows without gradient) == the five-call
    composition it replaces (forced here by a flow that asks for a gradient): values and feature gradients."""
    g = torch.Generator().manual_seed(44)
    n, c, r = 2, 48, 24
    f1 = torch.randn(n, c, r, r, generator=g).cuda(); f2 = torch.randn(n, c, r, r, generator=g).cuda()
    ff, fb = _flows(8, n, r, r, 0.8)
    ff, fb = ff.cuda(), fb.cuda()
    go = torch.randn(n, c, r, r, generator=g).cuda()
    warper = dcb.FeatureWarperSoftsplat(with_learnable_metric=False)
    res = []
    for composed in (False, True):
        a = f1.clone().requires_grad_(True); b = f2.clone().requires_grad_(True)
        out = dcb.bidirectional_warp_fuse(a, b, ff.clone().requires_grad_(composed), fb, warper)
        out.backward(go)
        res.append((out.detach(), a.grad, b.grad))
    for x, y, what in zip(res[0], res[1], ("fused", "grad first", "grad last")):
        assert_close(x, y, 1e-6, f"single node vs composition: {what}")
    with torch.no_grad():
        assert_close(dcb.bidirectional_warp_fuse(f1, f2, ff, fb, warper), res[0][0], 1e-6, "no-grad call")


# ---------------------------------------------------------------------------------------------
# f-3: latent tile merge (patch_utils.py:83-174)
# ---------------------------------------------------------------------------------------------
import glob as _glob
import os as _os

import numpy as _np

_TILE_GOLDEN = sorted(_glob.glob(_os.path.join(_os.path.dirname(__file__), "golden", "ref_tiles_*.npz")))


@pytest.mark.parametrize("path", _TILE_GOLDEN, ids=[_os.path.basename(p) for p in _TILE_GOLDEN])
def test_tile_merge_matches_reference_golden(dcb, path):
    """The CUDA gather kernel against the output of the reference's own function (oracle/ref_tiles.py)."""
    z = _np.load(path)
    tiles = [torch.from_numpy(z[k]).cuda() for k in sorted(k for k in z.files if k.startswith("tile_"))]
    coords = [tuple(int(v) for v in r) for r in z["coords"]]
    full, size = tuple(int(v) for v in z["full"]), tuple(int(v) for v in z["size"])
    got = dcb.merge_latent_tiles_from_pixel_coords(tiles, coords, full, size)
    want = torch.from_numpy(z["out"])
    assert got.shape == want.shape and got.dtype == torch.float32
    assert_close(got, want, 1e-5, "tile merge vs reference")
    assert torch.equal(got.cpu() == 0, want == 0)          # uncovered / zero-weight pixels are exactly 0, like the reference's 0 / eps


def test_tile_merge_full_frame_crop_roundtrip(dcb, orc):
    """1080p-like canvas at pixel scale: crop_into_tiles views (non-contiguous tiles) merged back.
    With coordinates in the order the merge reads them, every interior pixel is the input again."""
    g = torch.Generator().manual_seed(5)
    img = torch.randn(3, 270, 480, generator=g)
    tiles, coords, (h, w) = dcb.crop_into_tiles(img.cuda(), (128, 128), overlap=16, order="chw")
    as_read = [(x, x2, y, y2) for (y, y2, x, x2) in coords]                     # (x1, x2, y1, y2), patch_utils.py:135
    got = dcb.merge_latent_tiles_from_pixel_coords([t.unsqueeze(0) for t in tiles], as_read, (1, 3, h, w), (h, w))
    ref = orc.merge_latent_tiles_from_pixel_coords([t.unsqueeze(0).cpu() for t in tiles], as_read, (1, 3, h, w), (h, w))
    assert_close(got, ref, 1e-5, "crop -> merge")
    inner = got[0, :, 1:-1, 1:-1].cpu()
    covered = ref[0, :, 1:-1, 1:-1] != 0
    assert_close(inner[covered], img[:, 1:-1, 1:-1][covered], 1e-5, "merge of crops reproduces the frame")


def test_tile_merge_long_lists_resize_batch_and_bf16(dcb, orc):
    g = torch.Generator().manual_seed(6)
    # 130 tiles (> 48 per launch: chained through the workspace canvas), every one resized, canvas batch of 2
    coords, lat = [], []
    for i in range(130):
        x1 = int(torch.randint(0, 180, (1,), generator=g)); y1 = int(torch.randint(0, 100, (1,), generator=g))
        coords.append((x1, x1 + int(torch.randint(8, 70, (1,), generator=g)), y1, y1 + int(torch.randint(8, 60, (1,), generator=g))))
        lat.append(torch.randn(1, 6, int(torch.randint(2, 12, (1,), generator=g)), int(torch.randint(2, 12, (1,), generator=g)), generator=g))
    full, size = (2, 6, 40, 64), (160, 256)
    ref = orc.merge_latent_tiles_from_pixel_coords(lat, coords, full, size)
    got = dcb.merge_latent_tiles_from_pixel_coords([t.cuda() for t in lat], coords, full, size)
    assert_close(got, ref, 2e-5, "130 resized tiles")
    assert torch.equal(got[0], got[1])
    # bf16 tiles: fp32 accumulation, within 1e-2 of the fp32 reference on the same (rounded) tiles
    lat16 = [t.bfloat16() for t in lat[:20]]
    ref16 = orc.merge_latent_tiles_from_pixel_coords([t.float() for t in lat16], coords[:20], (1, 6, 40, 64), size)
    got16 = dcb.merge_latent_tiles_from_pixel_coords([t.cuda() for t in lat16], coords[:20], (1, 6, 40, 64), size)
    assert got16.dtype == torch.bfloat16
    assert_close(got16.float(), ref16, 1e-2, "bf16 tiles")
    # every tile outside the canvas: all zeros, like 0 / eps
    out = dcb.merge_latent_tiles_from_pixel_coords([lat[0].cuda()], [(500, 520, 500, 520)], (1, 6, 40, 64), size)
    assert float(out.abs().max()) == 0.0
    with pytest.raises(AssertionError):
        dcb.merge_latent_tiles_from_pixel_coords([lat[0].cuda()], coords[:2], full, size)


def test_dropin_extractor_matches_reference_execution(dcb):
    """tests/golden/ref_gpu_extractor.npz: the reference's OWN Bi_Dir_FeatureExtractor (extractors.py:209-316 with its
    control_utils.py and softsplat.py, kernels NVRTC-compiled on a B200 -- baseline/ref_gpu_extractor_golden.py) on seeded
    inputs and parameters. The drop-in module (same state_dict; per scale ONE fused call: dcb_bidir_block_fwd / _bwd)
    must reproduce all four outputs and the gradients, including those of the learned metric (splat + fusion paths)."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ref_gpu_extractor_golden", os.path.join(root, "baseline", "ref_gpu_extractor_golden.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    z = np.load(os.path.join(root, "tests", "golden", "ref_gpu_extractor.npz"))
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ours = dcb.Bi_Dir_FeatureExtractor(gen.INJECT).cuda()
        ours.load_state_dict(gen.seeded_state(ours))
        cond, flow = gen.make_inputs()
        g = torch.Generator().manual_seed(3)
        gouts = [torch.randn(1, c, r, r, generator=g).cuda() for c, r in zip(gen.INJECT, (64, 32, 16, 8))]
        before = dcb.launch_count()
        outs, grads = gen.run(ours, cond, flow, gouts)
        assert dcb.launch_count() > before
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    for i, o in enumerate(outs):
        assert_close(o, torch.from_numpy(z[f"out{i}"]), 2e-5, f"extractor scale {i}")
    for k, v in grads.items():
        ref = torch.from_numpy(z["grad/" + k])
        assert float(ref.abs().max()) > 0, k                     # the fixture exercises every scale
        assert_close(v, ref, 1e-4, f"extractor grad {k}")


def test_fused_block_equals_composition(dcb):
    """dcb_bidir_block_fwd / _bwd == masks + two masked differentiable splats + fusion kernel (the five-call composition),
    values and all four gradients, at a pyramid shape with a learned-metric stand-in."""
    g = torch.Generator().manual_seed(17)
    n, c, r = 2, 48, 32
    first = torch.randn(n, c, r, r, generator=g).cuda(); last = torch.randn(n, c, r, r, generator=g).cuda()
    ff = (torch.randn(n, 2, r, r, generator=g) * 0.6).cuda(); fb = (-ff.cpu() + 0.25 * torch.randn(n, 2, r, r, generator=g)).cuda()
    mf = (torch.randn(n, 1, r, r, generator=g) * 0.5 + 0.1).cuda(); mb = (torch.randn(n, 1, r, r, generator=g) * 0.5 + 0.1).cuda()
    gout = torch.randn(n, c, r, r, generator=g).cuda()

    def run(block):
        t = [x.clone().requires_grad_(True) for x in (first, last, mf, mb)]
        out = block(*t)
        out.backward(gout)
        return out.detach(), [x.grad for x in t]

    def composed(a, b, m1, m2):
        import importlib
        ss = importlib.import_module(dcb.__name__ + ".softsplat")
        of, ob = dcb.compute_mask(ff, fb), dcb.compute_mask(fb, ff)
        w1 = ss._splat_normalised(a, ff, m1, dcb._lib.MODE_SOFT, dcb._lib.EPS_ADD, mask=of)
        w2 = ss._splat_normalised(b, fb, m2, dcb._lib.MODE_SOFT, dcb._lib.EPS_ADD, mask=ob)
        return dcb.bidir_fuse(w1, w2, m1, m2, of, ob)

    o1, g1 = run(lambda a, b, m1, m2: dcb.bidirectional_block(a, b, ff, fb, m1, m2))
    o2, g2 = run(composed)
    assert_close(o1, o2, 1e-5, "fused block forward")
    for name, x, y in zip(("first", "last", "metric_f", "metric_b"), g1, g2):
        assert_close(x, y, 2e-5, f"fused block grad {name}")
    with torch.no_grad():                                         # inference path: nothing saved, optional outputs in scratch
        assert_close(dcb.bidirectional_block(first, last, ff, fb, mf, mb), o2, 1e-5, "fused block, no grad")


def test_flow_ingest_on_the_device(dcb, tmp_path):
    """f-2 on the device: the raw .flo payload through ONE kernel (dcb_flow_resize) == the reference's host readers +
    torch ops (its actual callees): read_flo + resize_flow_to (utils.py:10-28), load_flo_file + fast_downsample_flow
    (dataset.py:15-50, with the planar mis-reshape), resize_and_normalize_flow_batched (control_utils.py:74-97)."""
    fio = dcb.flow_io
    g = torch.Generator().manual_seed(4)
    for (h, w) in ((96, 160), (135, 241), (64, 64)):
        flow_hw2 = (torch.randn(h, w, 2, generator=g) * 5).numpy()
        path = str(tmp_path / f"f_{h}_{w}.flo")
        fio.write_flo(path, flow_hw2)
        # de-interleave only
        assert torch.equal(fio.flo_to_device(path).cpu(), torch.from_numpy(flow_hw2).permute(2, 0, 1)[None])
        for (th, tw) in ((64, 64), (33, 47), (h, w), (2 * h, 2 * w + 1)):
            ref = fio.resize_flow_to(fio.read_flo(path), th, tw)                               # utils.py:21-28 on the host
            assert_close(fio.flo_to_device(path, target_hw=(th, tw)), ref, 2e-6, f"bilinear rescale {h}x{w}->{th}x{tw}")
            quirk = fio.read_flo(path, planar_quirk=True)                                      # dataset.py:15-24
            ref = torch.from_numpy(fio.fast_downsample_flow(quirk, th, tw))[None] if th <= h and tw <= w else None
            if ref is not None:
                got = fio.flo_to_device(path, target_hw=(th, tw), convention="adaptive_avg", planar_quirk=True)
                assert_close(got, ref, 2e-6, f"adaptive avg (planar quirk) {h}x{w}->{th}x{tw}")
    flow = (torch.randn(3, 2, 512, 512, generator=g) * 9).cuda()
    for r in (64, 32, 16, 8):
        ref = dcb.resize_and_normalize_flow_batched(flow, r, r)                                # interpolate + 2 divisions + stack
        assert_close(fio.resize_and_normalize_flow_device(flow, r, r), ref, 2e-6, f"normalise {r}")
    sliced = torch.randn(2, 4, 70, 90, generator=g).cuda()[:, 1:3]                             # strided view, bf16 output
    ref = torch.nn.functional.adaptive_avg_pool2d(sliced, (16, 24))
    assert_close(fio.resize_flow_device(sliced, 16, 24, "adaptive_avg", torch.bfloat16).float(), ref, 1e-2, "strided, bf16 out")


def test_flow_ingest_device_matches_reference_golden(dcb, tmp_path):
    """The device ingest against tests/golden/ref_flow_io.npz (the reference's own functions, oracle/ref_flow_io.py)."""
    import os
    from oracle.ref_flow_io import CASES, case_flow
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", "ref_flow_io.npz"))
    fio = dcb.flow_io
    for i, (h, w, th, tw) in enumerate(CASES):
        path = str(tmp_path / "x.flo")
        fio.write_flo(path, case_flow(h, w, 100 + i))
        assert_close(fio.flo_to_device(path, target_hw=(th, tw)), torch.from_numpy(z[f"{i}/resize_flow_to"]), 2e-6, f"case {i} resize_flow_to")
        if f"{i}/fast_downsample_flow" in z.files:
            got = fio.flo_to_device(path, target_hw=(th, tw), convention="adaptive_avg", planar_quirk=True)
            assert_close(got[0], torch.from_numpy(z[f"{i}/fast_downsample_flow"]), 2e-6, f"case {i} fast_downsample_flow")
    flow = torch.from_numpy(z["batched/in"]).cuda()
    for r in (64, 32, 16, 8):
        assert_close(fio.resize_and_normalize_flow_device(flow, r, r), torch.from_numpy(z[f"batched/normalize_{r}"]), 2e-6, f"normalise {r}")


# ---------------------------------------------------------------------------------------------
# whole pyramids in one call (csrc/pyramid.cu): f-1 as the survey defines it, the batched entry, f-4
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_resample_batch_matches_torch(dcb, dtype):
    """dcb_resample_batch == F.interpolate(bilinear) followed by the per-channel factor, for every convention the reference
    uses (control_utils.py:74-97, extractors.py:182-183, notebook cell 5, utils.py:21-28), all jobs in ONE launch."""
    g = torch.Generator().manual_seed(21)
    flow = (torch.randn(2, 2, 67, 121, generator=g) * 5).cuda().to(dtype)
    img = torch.rand(2, 3, 96, 80, generator=g).cuda().to(dtype)
    strided = torch.randn(1, 4, 40, 2 * 56, generator=g).cuda().to(dtype)[:, :, :, ::2]          # non-contiguous source
    R = dcb._lib
    jobs = [(flow, (32, 32), False, R.RESAMPLE_DIV, 15.5, 15.5), (flow, (17, 23), False, R.RESAMPLE_DIV, 8.0, 8.0),
            (flow, (134, 200), True, R.RESAMPLE_MUL, 200 / 121, 134 / 67), (img, (48, 40), False, R.RESAMPLE_NONE, 1.0, 1.0),
            (img, (1, 1), False, R.RESAMPLE_NONE, 1.0, 1.0), (strided, (64, 64), False, R.RESAMPLE_MUL, 0.5, 0.5),
            (flow, (8, 8), False, R.RESAMPLE_MUL, 0.25, 0.25, torch.float32)]
    before = dcb.launch_count()
    got = dcb.resample_batch(jobs)
    assert dcb.launch_count() - before == 1
    for job, out in zip(jobs, got):
        src, size, align, op, f0, f1 = job[:6]
        ref = torch.nn.functional.interpolate(src.float(), size=size, mode="bilinear", align_corners=align)
        fac = torch.tensor([f0] + [f1] * (src.shape[1] - 1), device="cuda").view(1, -1, 1, 1)
        ref = ref * fac if op == R.RESAMPLE_MUL else (ref / fac if op == R.RESAMPLE_DIV else ref)
        assert out.shape == ref.shape and out.dtype == (job[6] if len(job) > 6 else dtype)
        assert_close(out.float(), ref, 2e-6 if out.dtype == torch.float32 else 1e-2, f"resample {size} align={align} op={op}")


def _pyramid_inputs(seed, shapes, learned=True):
    g = torch.Generator().manual_seed(seed)
    levels = []
    for n, c, r in shapes:
        first, last = torch.randn(n, c, r, r, generator=g), torch.randn(n, c, r, r, generator=g)
        ff = torch.randn(n, 2, r, r, generator=g) * 0.6
        fb = -ff + 0.25 * torch.randn(n, 2, r, r, generator=g)
        mf = torch.randn(n, 1, r, r, generator=g) * 0.5 + 0.1 if learned else None
        mb = torch.randn(n, 1, r, r, generator=g) * 0.5 + 0.1 if learned else None
        levels.append((first, last, ff, fb, mf, mb))
    return levels


def test_bidirectional_pyramid_two_launches_and_oracle(dcb, orc):
    """The whole 4-scale pyramid of the live consumer's shapes (batch 2) in one call: exactly two kernel launches, and every
    scale equal to the oracle's composition (compute_mask x2, feature_warper x2, fuse) away from mask flips at the 0.3
    threshold; afterwards the accumulator workspace is all-zero again (WS_CLEAN protocol)."""
    levels = _pyramid_inputs(31, [(2, 160, 64), (2, 160, 32), (2, 320, 16), (2, 640, 8)])
    dev = [tuple(t.cuda() for t in lv) for lv in levels]
    with torch.no_grad():
        dcb.bidirectional_pyramid(dev)                      # warm: workspace allocation
        before = dcb.launch_count()
        fused = dcb.bidirectional_pyramid(dev)
        assert dcb.launch_count() - before == 2
    ws = dcb._lib.workspace(torch.device("cuda", torch.cuda.current_device()), 1, "acc")
    assert int(ws.count_nonzero()) == 0, "accumulators not returned clean"
    for (first, last, ff, fb, mf, mb), got in zip(levels, fused):
        of, ob = orc.compute_mask(ff, fb), orc.compute_mask(fb, ff)
        wa, ca = orc.feature_warper(first, ff, mf, of)
        wb, cb = orc.feature_warper(last, fb, mb, ob)
        ref = orc.fuse(wa, wb, ca, cb, of, ob)
        same = (dcb.compute_mask(ff.cuda(), fb.cuda()).cpu() == of) & (dcb.compute_mask(fb.cuda(), ff.cuda()).cpu() == ob)
        assert same.float().mean() > 0.995
        assert_close(got.cpu()[same.expand_as(ref)], ref[same.expand_as(ref)], 1e-5, f"pyramid scale {tuple(first.shape)}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bidirectional_pyramid_equals_blocks_with_gradients(dcb, dtype):
    """One pyramid node == one fused block per scale (dcb_bidir_block_fwd / _bwd): values and all four gradients per scale,
    odd sizes and a channel count that is no multiple of four included; plus the ones-metric (None) form."""
    shapes = [(2, 48, 32), (1, 30, 19), (3, 8, 8)]
    levels = _pyramid_inputs(32, shapes)
    rel = 2e-5 if dtype == torch.float32 else 2e-2

    def run(whole):
        ts = [[None if t is None else t.cuda().to(dtype) for t in lv] for lv in levels]
        for lv in ts:
            for k in (0, 1, 4, 5):
                lv[k].requires_grad_(True)
        if whole:
            outs = dcb.bidirectional_pyramid([tuple(lv) for lv in ts])
        else:
            outs = [dcb.bidirectional_block(*lv) for lv in ts]
        g = torch.Generator().manual_seed(5)
        torch.autograd.backward(list(outs), [torch.randn(o.shape, generator=g).cuda().to(dtype) for o in outs])   # one node: one backward
        return outs, [[lv[k].grad for k in (0, 1, 4, 5)] for lv in ts]

    (o1, g1), (o2, g2) = run(True), run(False)
    for l, (a, b) in enumerate(zip(o1, o2)):
        # the two paths compute the occlusion masks with different accumulator layouts: compare away from flipped pixels
        ff, fb = levels[l][2].cuda().to(dtype), levels[l][3].cuda().to(dtype)
        flips = int((a.detach() != b.detach()).any(dim=1).sum())
        assert flips <= 3 or torch.allclose(a.float(), b.float(), rtol=rel, atol=rel), f"scale {l}: {flips} pixels differ"
        if flips == 0 or dtype == torch.float32:
            same = ~((a.detach() - b.detach()).abs().float().amax(dim=1, keepdim=True) > rel * 10 * float(b.detach().abs().max()))
            assert same.float().mean() > 0.995
            assert_close(a.detach().float()[same.expand_as(a)], b.detach().float()[same.expand_as(b)], rel, f"pyramid vs block, scale {l}")
            if bool(same.all()):
                for name, x, y in zip(("first", "last", "metric_f", "metric_b"), g1[l], g2[l]):
                    assert_close(x.float(), y.float(), rel * 2, f"pyramid vs block grad {name}, scale {l}")
    with torch.no_grad():                                       # channels_last feature maps (cuDNN's preferred layout): vector quad loads
        cl = [tuple(None if t is None else t.cuda().to(dtype) for t in lv) for lv in levels]
        nh = [(lv[0].to(memory_format=torch.channels_last), lv[1].to(memory_format=torch.channels_last)) + lv[2:] for lv in cl]
        for a, b in zip(dcb.bidirectional_pyramid(nh), dcb.bidirectional_pyramid(cl)):
            assert_close(a.float(), b.float(), 1e-5 if dtype == torch.float32 else 2e-2, "channels_last pyramid")
    with torch.no_grad():                                       # all-ones metrics are never materialised
        lv = [(t[0].cuda(), t[1].cuda(), t[2].cuda(), t[3].cuda(), None, None) for t in levels]
        ones = [(t[0], t[1], t[2], t[3], torch.ones_like(t[2][:, :1]), torch.ones_like(t[2][:, :1])) for t in lv]
        for a, b in zip(dcb.bidirectional_pyramid(lv), dcb.bidirectional_pyramid(ones)):
            assert_close(a, b, 1e-6, "ones metric")


def test_pyramid_conditioning_matches_notebook_golden(dcb):
    """Row f-4: tests/golden/ref_notebook_pyramid.npz is the reference's own notebook loop (improv_experiments.ipynb cells 3
    and 5, executed by oracle/ref_notebook.py). Here: one resampling launch + one pyramid call (two launches) for all scales."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_notebook_pyramid.npz"))
    img1, img2 = (torch.from_numpy(z[k]).float().cuda() / 255.0 for k in ("img1_u8", "img2_u8"))
    f1, f2 = torch.from_numpy(z["flow1"]).cuda(), torch.from_numpy(z["flow2"]).cuda()
    sizes = [int(s) for s in z["sizes"]]
    dcb.pyramid_conditioning(img1, img2, f1, f2, sizes)
    before = dcb.launch_count()
    got = dcb.pyramid_conditioning(img1, img2, f1, f2, sizes)
    assert dcb.launch_count() - before == 3
    for size, (w1, w2, fused) in zip(sizes, got):
        for name, t in (("warped1", w1), ("warped2", w2), ("fused", fused)):
            assert_close(t, torch.from_numpy(z[f"{name}_{size}"]), 1e-5, f"notebook pyramid {name} at {size}")
