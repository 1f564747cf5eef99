"""Parity at the configurations BASELINE.json names (C1-C4, the 64-frame headline's code path, the
conditioning recipe), CUDA path vs the CPU oracle on the same seeded inputs, through both forward
kernel families:

  * "owner": target-tile-owner kernels (splat_owner.cu), with several frame groups (pre-pass + owner
    launch pairs) forced by a small ``owner_group_bytes``;
  * "pipe":  the round-1 accumulator pipelines (splat_pipe.cu / splat_planar.cu): >= 3 ring groups,
    both by real size (3 x 1080p) and by a shrunken ``pipe_group_bytes`` on small tensors.

Each test appends one JSON line per compared tensor to ``gpurun_out/parity_report.jsonl``: the max
PLAIN relative error |a - r| / max|r| , the max element-wise relative error where |r| > 1e-3 max|r|,
the same for the numerator S and the normaliser D separately, and how many
elements needed the fp64-truth widening of tests/util.assert_close.

Reference semantics: controlnet/softsplat.py:232-274 (modes), :284-335 (kernel).
"""
import json
import os

import pytest
import torch

from tests.util import assert_close, cuda_run, oracle_run

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


@pytest.fixture(scope="module")
def dcb():
    import diffcodec_b200
    return diffcodec_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(params=["owner", "pipe", "pipe-ring2"])
def path(request, dcb):
    """Pin the forward kernel family; restore the library's own dispatch afterwards."""
    L = dcb._lib
    L.set_option("fwd_path", 2 if request.param == "owner" else 1)
    L.set_option("pipe_ring_slots", 2 if request.param == "pipe-ring2" else 1)
    L.release_workspaces()
    yield request.param
    L.set_option("fwd_path", 0)
    L.set_option("pipe_ring_slots", 1)
    L.set_option("pipe_group_bytes", 0)
    L.set_option("owner_group_bytes", 0)
    L.set_option("bwd_group_bytes", 0)
    L.release_workspaces()


def smooth_flow(n, h, w, amp, seed):
    """bench.py's synthetic flow: low-resolution noise, bicubic upsampling, ~amp px."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(n, 2, max(h // 32, 2), max(w // 32, 2), generator=g)
    return torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=False) * amp


def report(case, name, got, ref, truth=None, rel=1e-5):
    a, r = got.detach().double().cpu(), ref.detach().double().cpu()
    scale = float(r.abs().max()) or 1.0
    err = (a - r).abs()
    big = r.abs() > 1e-3 * scale
    rec = {"case": case, "tensor": name, "numel": a.numel(), "max_abs_err_over_max_ref": float(err.max() / scale),
           "max_elementwise_rel_err_where_ref_gt_1e-3_max": float((err[big] / r.abs()[big]).max()) if big.any() else 0.0}
    if truth is not None:
        plain_tol = rel * (r.abs() + scale)
        rec["elements_needing_truth_widening"] = int((err > plain_tol).sum())
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(rec) + "\n")
    return rec


def numerator_and_normaliser(dcb, orc, case, tin, flow, metric, mode):
    """S and D separately (App. C-9): sum-splat of X = cat(in * g, g) on both sides."""
    g = metric.exp() if mode == "soft" else (metric if mode == "linear" else torch.ones_like(tin[:, :1]))
    x = torch.cat([tin * g, g], 1)
    s_ref = orc.softsplat(x, flow, None, "sum")
    s_got = dcb.softsplat(x.cuda(), flow.cuda(), None, "sum").cpu()
    c = tin.shape[1]
    assert_close(s_got[:, :c], s_ref[:, :c], 1e-5, f"{case} numerator S")
    assert_close(s_got[:, c:], s_ref[:, c:], 1e-5, f"{case} normaliser D")
    report(case, "S", s_got[:, :c], s_ref[:, :c])
    report(case, "D", s_got[:, c:], s_ref[:, c:])


@pytest.mark.parametrize("mode", ["soft", "avg"])
def test_three_1080p_frames_forward(dcb, orc, path, mode):
    """3 x 3 x 1080 x 1920 fp32 (C1 / headline frames): pipe -> 3 ring groups through the two-slot ring with PDL;
    owner -> 3 frame groups of (pre-pass, owner) launches. Twice, to catch workspace leftovers."""
    g = torch.Generator().manual_seed(11)
    tin = torch.rand(3, 3, 1080, 1920, generator=g)
    metric = -torch.rand(3, 1, 1080, 1920, generator=g)
    flow = smooth_flow(3, 1080, 1920, 8.0, 12)
    flow[1, :, 100:140, 300:340] = float("nan"); flow[2, 0, 500:520, :64] = -4000.0     # dead pixels and a hole
    me = metric if mode == "soft" else None
    if path == "owner":
        dcb._lib.set_option("owner_group_bytes", 1080 * 1920 * 8)      # one frame per launch pair
    ref = orc.softsplat(tin, flow, me, mode)
    for rep in range(2):
        got = dcb.softsplat(tin.cuda(), flow.cuda(), None if me is None else me.cuda(), mode)
        assert_close(got, ref, 1e-5, f"3x1080p {mode} {path} run {rep}")
    report(f"3x3x1080x1920 {mode} {path}", "out", got, ref)
    if mode == "soft":
        numerator_and_normaliser(dcb, orc, f"3x3x1080x1920 {mode} {path}", tin, flow, metric, mode)


def test_c1_single_frame_avg(dcb, orc, path):
    """BASELINE C1 exactly: avg forward of one 1 x 3 x 1080 x 1920 fp32 frame."""
    g = torch.Generator().manual_seed(0)
    tin = torch.rand(1, 3, 1080, 1920, generator=g)
    flow = smooth_flow(1, 1080, 1920, 8.0, 1)
    got = dcb.softsplat(tin.cuda(), flow.cuda(), None, "avg")
    ref = orc.softsplat(tin, flow, None, "avg")
    assert_close(got, ref, 1e-5, f"C1 {path}")
    report(f"C1 1x3x1080x1920 avg {path}", "out", got, ref)
    # the adversarial set of SURVEY.md 8d: unsmoothed randn * 32 px
    rough = torch.randn(1, 2, 1080, 1920, generator=g) * 32
    got = dcb.softsplat(tin.cuda(), rough.cuda(), None, "avg")
    ref = orc.softsplat(tin, rough, None, "avg")
    assert_close(got, ref, 1e-5, f"C1 rough {path}")
    report(f"C1 1x3x1080x1920 avg rough-32px {path}", "out", got, ref)


@pytest.mark.parametrize("flow_fp32", [False, True])
def test_c2_latents_bf16(dcb, orc, path, flow_fp32):
    """BASELINE C2 exactly: soft forward of SD latents 4 x 4 x 135 x 240 bf16 (bf16 and fp32 flow)."""
    g = torch.Generator().manual_seed(1)
    lat = (torch.randn(4, 4, 135, 240, generator=g) * 0.18215).bfloat16()
    met = (-torch.randn(4, 1, 135, 240, generator=g).abs()).bfloat16()
    fl = torch.randn(4, 2, 135, 240, generator=g)
    fb = fl if flow_fp32 else fl.bfloat16()
    ref = orc.softsplat(lat.float(), fb.float(), met.float(), "soft")
    got = dcb.softsplat(lat.cuda(), fb.cuda(), met.cuda(), "soft")
    assert got.dtype == torch.bfloat16
    assert_close(got.float(), ref, 1e-2, f"C2 {path} flow_fp32={flow_fp32}")
    report(f"C2 4x4x135x240 bf16 flow_fp32={flow_fp32} {path}", "out", got.float(), ref, rel=1e-2)


def test_c4_features_forward_backward(dcb, orc):
    """BASELINE C4 exactly: soft forward + all three gradients on 8 x 64 x 256 x 256 fp32 (list path + gather backward)."""
    g = torch.Generator().manual_seed(3)
    tin = torch.randn(8, 64, 256, 256, generator=g)
    metric = torch.randn(8, 1, 256, 256, generator=g) * 0.5
    flow = smooth_flow(8, 256, 256, 4.0, 4)
    gout = torch.randn(8, 64, 256, 256, generator=g)
    ref = oracle_run(orc, tin, flow, metric, gout, "soft")
    truth = oracle_run(orc, tin[:2].double(), flow[:2].double(), metric[:2].double(), gout[:2].double(), "soft")   # fp64 on 2 of the 8 frames
    got = cuda_run(dcb.softsplat, tin, flow, metric, gout, "soft")
    for k in ("out", "gin", "gflow", "gmetric"):
        assert_close(got[k][:2], ref[k][:2], 2e-5, f"C4 {k} (frames 0-1)", truth=truth[k])
        report("C4 8x64x256x256 soft", k, got[k][:2], ref[k][:2], truth=truth[k], rel=2e-5)
    # the remaining frames without the widening: forward at 1e-5, gradients at the tolerance the fp32 reference itself meets
    assert_close(got["out"], ref["out"], 1e-5, "C4 out")
    assert_close(got["gin"], ref["gin"], 1e-4, "C4 gin")


def test_many_channel_ring_groups(dcb, orc):
    """Planar (channel-quad) pipeline through >= 3 ring groups: 4 x 8 x 540 x 960 by size, and a small
    tensor with the ring slots shrunk to one frame."""
    L = dcb._lib
    try:
        g = torch.Generator().manual_seed(5)
        tin = torch.randn(4, 12, 540, 960, generator=g)
        metric = torch.randn(4, 1, 540, 960, generator=g) * 0.5
        flow = smooth_flow(4, 540, 960, 6.0, 6)
        L.set_option("pipe_group_bytes", 540 * 960 * 16 * 3)          # one frame (3 channel quads) per slot -> 4 groups
        L.release_workspaces()
        ref = orc.softsplat(tin, flow, metric, "soft")
        for rep in range(2):
            got = dcb.softsplat(tin.cuda(), flow.cuda(), metric.cuda(), "soft")
            assert_close(got, ref, 1e-5, f"planar 4 groups run {rep}")
        report("planar 4x12x540x960 soft, 4 ring groups", "out", got, ref)
        tin, flow, metric, gout = (t[:, :, :70, :150].contiguous() for t in (tin, flow, metric, torch.randn(4, 12, 540, 960, generator=g)))
        tin = torch.cat([tin, tin.flip(0), tin * 0.5], 0)[:7]; flow = torch.cat([flow, -flow, flow * 0.5], 0)[:7]
        metric = torch.cat([metric, metric, metric], 0)[:7]; gout = torch.cat([gout, gout, gout], 0)[:7]
        L.set_option("pipe_group_bytes", 70 * 150 * 16 * 3)
        L.release_workspaces()
        ref = oracle_run(orc, tin, flow, metric, gout, "soft")
        truth = oracle_run(orc, tin.double(), flow.double(), metric.double(), gout.double(), "soft")
        got = cuda_run(dcb.softsplat, tin, flow, metric, gout, "soft")
        for k in ("out", "gin", "gflow", "gmetric"):
            assert_close(got[k], ref[k], 1e-5, f"planar 7 groups {k}", truth=truth[k])
    finally:
        L.set_option("pipe_group_bytes", 0)
        L.release_workspaces()


def test_small_frames_many_ring_groups(dcb, orc, path):
    """7 x 3 x 70 x 150 with one frame per ring slot / per launch pair: S0 | S1,N0 | S2,N1 | ... with the slot
    re-zeroing and the PDL chain of the real headline, at a size the oracle checks in full (forward + gradients),
    followed by the mask epilogue and a differently shaped call on the same workspace."""
    from tests.util import make_inputs
    L = dcb._lib
    tin, flow, metric, gout = make_inputs(31, 7, 3, 70, 150, flow_scale=3.0)
    L.set_option("pipe_group_bytes", 70 * 150 * 16)
    L.set_option("owner_group_bytes", 70 * 150 * 8)
    L.set_option("bwd_group_bytes", 70 * 150 * 16 * 2)      # packed backward: two frames per group -> 4 groups, the last one ragged
    L.release_workspaces()
    ref = oracle_run(orc, tin, flow, metric, gout, "soft")
    truth = oracle_run(orc, tin.double(), flow.double(), metric.double(), gout.double(), "soft")
    for rep in range(3):
        got = cuda_run(dcb.softsplat, tin, flow, metric, gout, "soft")
        for k in ("out", "gin", "gflow", "gmetric"):
            assert_close(got[k], ref[k], 1e-5, f"7 groups {path} run {rep} {k}", truth=truth[k])
    report(f"7x3x70x150 soft, one frame per group, {path}", "out", got["out"], ref["out"], truth=truth["out"])
    fa, fb = flow, -flow + 0.3 * torch.randn_like(flow)
    assert torch.equal(dcb.compute_mask(fa.cuda(), fb.cuda()).cpu(), orc.compute_mask(fa, fb)) or \
        (dcb.compute_mask(fa.cuda(), fb.cuda()).cpu() != orc.compute_mask(fa, fb)).float().mean() < 1e-4
    assert_close(dcb.softsplat(tin[:2, :2].cuda(), flow[:2].cuda(), None, "avg"), orc.softsplat(tin[:2, :2], flow[:2], None, "avg"), 1e-5, "after groups")


def test_recipe_two_1080p_frames(dcb, orc, path):
    """The conditioning recipe (dcb_residual_fused) at 2 x 3 x 1080 x 1920, both variants."""
    g = torch.Generator().manual_seed(7)
    img = torch.rand(2, 3, 1080, 1920, generator=g); gt = torch.rand(2, 3, 1080, 1920, generator=g)
    f1 = smooth_flow(2, 1080, 1920, 8.0, 8); f2 = -f1 + 0.5 * smooth_flow(2, 1080, 1920, 1.0, 9)
    if path == "owner":
        dcb._lib.set_option("owner_group_bytes", 1080 * 1920 * 8)
    for variant in ("dataset", "wrapper"):
        fused, res, of, ob = dcb.residual_conditioning(img.cuda(), f1.cuda(), f2.cuda(), gt.cuda(), variant, return_masks=True)
        fr, rr, ofr, obr = orc.residual_recipe(img, f1, f2, gt, variant)
        # a mask pixel may flip where ||.|| is within rounding of the 0.3 threshold: compare where the masks agree
        agree = ((of.cpu() == ofr) & (ob.cpu() == obr)).expand_as(fr)
        assert agree.float().mean() > 0.9999, f"recipe {variant} {path}: masks differ on {(~agree).float().mean():.2e} of the pixels"
        assert_close(fused.cpu()[agree], fr[agree], 1e-5, f"recipe {variant} {path} fused")
        assert_close(res.cpu()[agree], rr[agree], 1e-5, f"recipe {variant} {path} residual")
        report(f"recipe 2x3x1080x1920 {variant} {path}", "residual", res.cpu()[agree], rr[agree])


def test_two_1080p_frames_forward_backward(dcb, orc, path):
    """soft forward + all gradients at 2 x 3 x 1080 x 1920 fp32 (the packed few-channel backward at frame size)."""
    g = torch.Generator().manual_seed(13)
    tin = torch.rand(2, 3, 1080, 1920, generator=g)
    metric = -torch.rand(2, 1, 1080, 1920, generator=g)
    flow = smooth_flow(2, 1080, 1920, 8.0, 14)
    gout = torch.randn(2, 3, 1080, 1920, generator=g)
    ref = oracle_run(orc, tin, flow, metric, gout, "soft")
    truth = oracle_run(orc, tin.double(), flow.double(), metric.double(), gout.double(), "soft")
    got = cuda_run(dcb.softsplat, tin, flow, metric, gout, "soft")
    for k in ("out", "gin", "gflow", "gmetric"):
        assert_close(got[k], ref[k], 1e-5, f"2x1080p fwd+bwd {path} {k}", truth=truth[k])
        report(f"2x3x1080x1920 soft fwd+bwd {path}", k, got[k], ref[k], truth=truth[k])


def test_owner_kernel_edge_cases(dcb, orc):
    """Target-tile-owner specifics: frames that are multiples of neither the 64 x 32 tile nor the 32 x 4 strip,
    every pixel landing in ONE cell (all losers but one per batch), flows that throw whole strips across the
    frame (boxes spanning every tile), 5..8 accumulated channels (two float4 per cell), self-derived boxes vs pre-pass."""
    from tests.util import make_inputs
    L = dcb._lib
    L.set_option("fwd_path", 2)
    try:
        for (n, c, h, w, scale) in [(2, 3, 37, 71, 2.0), (1, 1, 130, 257, 5.0), (3, 4, 135, 240, 1.0), (2, 7, 33, 65, 3.0),
                                    (1, 5, 260, 300, 40.0), (2, 2, 1, 300, 2.0), (2, 3, 300, 1, 2.0)]:
            tin, flow, metric, gout = make_inputs(n * 100 + c, n, c, h, w, flow_scale=scale)
            for mode in ("sum", "avg", "soft", "linear-zeroeps"):
                if mode == "sum" and c > 8 or mode != "sum" and c + 1 > 8:
                    continue
                me = metric.abs() + 0.1 if mode.startswith("linear") else metric
                ref = orc.softsplat(tin, flow, me if mode.split("-")[0] in ("soft", "linear") else None, mode)
                got = dcb.softsplat(tin.cuda(), flow.cuda(), me.cuda() if mode.split("-")[0] in ("soft", "linear") else None, mode)
                assert_close(got, ref, 1e-5, f"owner {n}x{c}x{h}x{w} {mode}")
        # all-to-one collision at 96 x 160 (15360 sources on one cell, fractional landing point)
        h, w = 96, 160
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        flow = torch.stack([(70.25 - xs).float(), (40.5 - ys).float()])[None]
        tin = torch.rand(1, 3, h, w)
        ref = orc.softsplat(tin, flow, None, "avg")
        got = dcb.softsplat(tin.cuda(), flow.cuda(), None, "avg")
        assert_close(got, ref, 2e-5, "owner all-to-one")
        # transposition-like flow: every strip's box spans the frame
        flow = torch.stack([(w - 1 - 2 * xs).float() + 0.3, (h - 1 - 2 * ys).float() + 0.6])[None]
        ref = orc.softsplat(tin, flow, None, "avg")
        got = dcb.softsplat(tin.cuda(), flow.cuda(), None, "avg")
        assert_close(got, ref, 1e-5, "owner mirrored frame")
    finally:
        L.set_option("fwd_path", 0)
