"""GPU parity tests of the forward splat and its backward, through the C ABI, against the CPU
oracle (oracle/) and the reference-derived golden fixtures (tests/golden/ref_emu_*.npz).

Tolerances (BASELINE.json north_star): 1e-5 relative in fp32 (atomic order is not deterministic),
1e-2 in bf16, bit-exact in deterministic mode."""
import glob
import os
import zlib

import numpy as np
import pytest
import torch

from tests.util import assert_close, cuda_run, make_inputs, oracle_run

pytestmark = pytest.mark.gpu

MODES = ["sum", "avg", "linear", "soft", "linear-zeroeps", "linear-clipeps", "soft-zeroeps", "soft-clipeps", "soft-addeps"]
SHAPES = [(1, 3, 9, 13), (2, 1, 16, 16), (2, 4, 33, 47), (1, 7, 40, 24), (3, 2, 5, 64), (1, 65, 12, 12)]


@pytest.fixture(scope="module")
def dcb():
    import diffcodec_b200
    return diffcodec_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(params=["auto", "pipelines"], autouse=True)
def kernel_family(request, dcb):
    """Every test of this module runs twice: with the library's own dispatch (small frames take the single-launch cluster
    kernel, csrc/splat_small.cu) and with the accumulator pipelines forced (dcb_set_option("fwd_path", 1): splat_pipe.cu /
    splat_planar.cu / splat_lists.cu), so that both kernel families see every mode, eps rule, layout and special flow."""
    L = dcb._lib
    L.set_option("fwd_path", 0 if request.param == "auto" else 1)
    L.release_workspaces()
    yield request.param
    L.set_option("fwd_path", 0)
    L.release_workspaces()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_forward_backward_fp32(dcb, orc, mode, shape):
    seed = zlib.crc32(repr((mode, shape)).encode()) % 1000          # stable across processes (str hash is salted)
    tin, flow, metric, gout = make_inputs(seed, *shape, flow_scale=2.5)
    if mode.startswith("linear"):
        metric = metric.abs() + 0.1          # keep the normaliser away from cancellation (tolerance is relative)
    ref = oracle_run(orc, tin, flow, metric, gout, mode)
    truth = oracle_run(orc, tin.double(), flow.double(), metric.double(), gout.double(), mode)
    got = cuda_run(dcb.softsplat, tin, flow, metric, gout, mode)
    for k in ("out", "gin", "gflow", "gmetric"):
        if ref[k] is not None:
            assert_close(got[k], ref[k], 1e-5, f"{mode} {shape} {k}", truth=truth[k])
    assert got["out"].is_contiguous() and got["out"].dtype == torch.float32


@pytest.mark.parametrize("mode", ["sum", "avg", "soft", "linear"])
def test_forward_backward_fp64(dcb, orc, mode):
    tin, flow, metric, gout = make_inputs(5, 2, 3, 17, 19, flow_scale=3.0, dtype=torch.float64)
    ref = oracle_run(orc, tin, flow, metric, gout, mode)
    got = cuda_run(dcb.softsplat, tin, flow, metric, gout, mode)
    for k in ("out", "gin", "gflow", "gmetric"):
        if ref[k] is not None:
            assert_close(got[k], ref[k], 1e-12, f"{mode} fp64 {k}")


FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_emu_*_fast.npz")) +
                  glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_gpu_*_f??.npz")))     # CPU emulation of, and a real B200 run of, the reference
GOLDEN_MODES = ["sum", "avg", "linear", "soft", "avg-zeroeps", "linear-clipeps", "soft-zeroeps", "soft-clipeps", "soft-addeps"]


@pytest.mark.parametrize("mode", GOLDEN_MODES)
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[4:-4] for p in FIXTURES])
def test_golden_reference_vectors(dcb, path, mode):
    """Outputs of the reference's own kernel text (sequential CPU emulation): includes integer /
    half-integer / out-of-frame / NaN / +-Inf / 1e30 flows and an all-to-one-pixel collision."""
    z = np.load(path)
    f64 = z["tin"].dtype == np.float64
    t = lambda k: torch.from_numpy(z[k])
    got = cuda_run(dcb.softsplat, t("tin"), t("flow"), t("metric"), t("gout"), mode)
    rel = 1e-12 if f64 else 1e-5
    truth = None
    if not f64:   # exact answer on the same fp32 inputs: bounds the reference's own conditioning error
        from oracle import oracle as orc
        truth = oracle_run(orc, t("tin").double(), t("flow").double(), t("metric").double(), t("gout").double(), mode)
    for k in ("out", "gin", "gflow", "gmetric"):
        key = f"{mode}/{k}"
        if key in z.files:
            assert_close(got[k], t(key), rel, f"{os.path.basename(path)} {key}", truth=None if truth is None else truth[k])


def test_func_level_matches_golden(dcb):
    for path in FIXTURES:
        z = np.load(path)
        ti = torch.from_numpy(z["tin"]).cuda().requires_grad_(True)
        fl = torch.from_numpy(z["flow"]).cuda().requires_grad_(True)
        out = dcb.softsplat_func.apply(ti, fl)
        out.backward(torch.from_numpy(z["gout"]).cuda())
        rel = 1e-12 if z["tin"].dtype == np.float64 else 1e-5
        assert_close(out, torch.from_numpy(z["func/out"]), rel, "func out")
        assert_close(ti.grad, torch.from_numpy(z["func/gin"]), rel, "func gin")
        assert_close(fl.grad, torch.from_numpy(z["func/gflow"]), rel, "func gflow")


@pytest.mark.parametrize("mode", ["sum", "avg", "linear", "soft"])
@pytest.mark.parametrize("flow_fp32", [False, True])
def test_bf16_within_1e2(dcb, orc, mode, flow_fp32):
    """bf16 semantics: inputs bf16, positions/weights/accumulation fp32, one rounding at the output.
    Oracle = fp32 reference on the up-cast inputs."""
    # flows below ~half a pixel keep every normaliser O(1): with a tiny normaliser the saved bf16
    # output (4e-3 relative) is amplified by 1/D in the gradients, for the reference as for us
    tin, flow, metric, gout = make_inputs(7, 2, 4, 24, 40, flow_scale=0.4)
    if mode == "linear":
        metric = metric.abs() + 0.25
    tb, mb, gb = tin.bfloat16(), metric.bfloat16(), gout.bfloat16()
    fb = flow if flow_fp32 else flow.bfloat16()
    ref = oracle_run(orc, tb.float(), fb.float(), mb.float(), gb.float(), mode)
    got = cuda_run(dcb.softsplat, tb, fb, mb, gb, mode)
    assert got["out"].dtype == torch.bfloat16
    for k in ("out", "gin", "gflow", "gmetric"):
        if ref[k] is not None:
            assert_close(got[k].float(), ref[k], 1e-2, f"bf16 {mode} {k}")


def test_bf16_accumulates_in_fp32(dcb):
    """4096 sources land on one pixel: a bf16 accumulator would stall at 256."""
    h = w = 64
    tin = torch.ones(1, 1, h, w, dtype=torch.bfloat16, device="cuda")
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    flow = torch.stack([(5 - xs).float(), (7 - ys).float()])[None].cuda()
    out = dcb.softsplat(tenIn=tin, tenFlow=flow, tenMetric=None, strMode="sum")
    assert out[0, 0, 7, 5].item() == 4096.0
    assert out.sum().item() == 4096.0


@pytest.mark.parametrize("mode", ["sum", "avg", "linear", "soft", "soft-clipeps"])
def test_deterministic_bit_exact(dcb, orc, mode):
    """sort-then-reduce mode: identical bits to the sequential oracle and from run to run -- `soft` included: the
    deterministic kernels evaluate exp() with IEEE double operations only (exp_det), and so does the oracle here."""
    tin, flow, metric, gout = make_inputs(21, 2, 3, 31, 45, flow_scale=4.0)
    flow[0, :, :8, :8] = 0.0
    ys, xs = torch.meshgrid(torch.arange(31), torch.arange(45), indexing="ij")
    flow[1, 0] = (10.25 - xs).float()                       # frame 1: 1395-way collision
    flow[1, 1] = (9.5 - ys).float()
    me = metric if mode.split("-")[0] in ("linear", "soft") else None
    ref = orc.softsplat(tin, flow, me, mode, exp_fn=orc.exp_det)
    with dcb.deterministic(True):
        a = dcb.softsplat(tenIn=tin.cuda(), tenFlow=flow.cuda(), tenMetric=None if me is None else me.cuda(), strMode=mode)
        b = dcb.softsplat(tenIn=tin.cuda(), tenFlow=flow.cuda(), tenMetric=None if me is None else me.cuda(), strMode=mode)
    assert torch.equal(a, b)
    assert torch.equal(a.cpu(), ref), f"max diff {(a.cpu() - ref).abs().max().item()}"


def test_deterministic_soft_run_to_run(dcb, orc):
    tin, flow, metric, _ = make_inputs(22, 1, 3, 64, 64, flow_scale=6.0)
    with dcb.deterministic(True):
        outs = [dcb.softsplat(tenIn=tin.cuda(), tenFlow=flow.cuda(), tenMetric=metric.cuda(), strMode="soft") for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert_close(outs[0], orc.softsplat(tin, flow, metric, "soft"), 1e-5, "det soft")


def test_known_answers(dcb):
    dev = "cuda"
    tin = torch.randn(2, 3, 12, 20, device=dev)
    zero = torch.zeros(2, 2, 12, 20, device=dev)
    # C-1: zero flow, sum -> identity, bit exact
    assert torch.equal(dcb.softsplat(tin, zero, None, "sum"), tin)
    # C-2: integer flow -> exact shift with zero fill
    fl = zero.clone(); fl[:, 0] = 3.0; fl[:, 1] = -2.0
    out = dcb.softsplat(tin, fl, None, "sum")
    exp = torch.zeros_like(tin); exp[:, :, :-2, 3:] = tin[:, :, 2:, :-3]
    assert torch.equal(out, exp)
    # C-3: half-pixel flow -> average of neighbours (exact in fp32)
    fl = zero.clone(); fl[:, 0] = 0.5
    out = dcb.softsplat(tin, fl, None, "sum")
    exp = 0.5 * tin; exp[:, :, :, 1:] += 0.5 * tin[:, :, :, :-1]
    assert torch.equal(out, exp)
    # C-4: avg under zero flow -> in / (1 + 1e-7); holes -> 0
    out = dcb.softsplat(tin, zero, None, "avg")
    # (the fast path multiplies by one correctly-rounded reciprocal: <= 1 ulp from the true quotient)
    assert_close(out, tin / (torch.ones_like(tin[:, :1]) + 0.0000001), 2e-7, "avg under zero flow")
    fl = zero.clone(); fl[:, 0] = 1000.0
    assert torch.count_nonzero(dcb.softsplat(tin, fl, None, "avg")) == 0
    # C-7: non-finite flow contributes nothing and gets zero gradients
    fl = (torch.rand(2, 2, 12, 20, device=dev) * 2).requires_grad_(True)
    with torch.no_grad():
        fl[0, 0, 3, 4] = float("nan"); fl[1, 1, 5, 6] = float("inf")
    ti = tin.clone().requires_grad_(True)
    out = dcb.softsplat(ti, fl, None, "sum")
    assert torch.isfinite(out).all()
    out.sum().backward()
    assert ti.grad[0, :, 3, 4].abs().sum() == 0 and ti.grad[1, :, 5, 6].abs().sum() == 0
    assert fl.grad[0, :, 3, 4].abs().sum() == 0 and fl.grad[1, :, 5, 6].abs().sum() == 0
    # C-8: soft with a constant metric == avg (up to the eps scaling); metric -inf removes a pixel
    m = torch.full((2, 1, 12, 20), 0.7, device=dev)
    fl = torch.randn(2, 2, 12, 20, device=dev)
    a = dcb.softsplat(tin, fl, None, "avg"); s = dcb.softsplat(tin, fl, m, "soft")
    solid = (dcb.softsplat(torch.ones_like(tin[:, :1]), fl, None, "sum") > 1e-2).expand_as(a)   # eps negligible there
    assert_close(s[solid], a[solid], 1e-5, "soft(const) vs avg")
    m2 = torch.zeros(2, 1, 12, 20, device=dev); m2[:, :, 0, 0] = -float("inf")
    t2 = tin.clone(); t2[:, :, 0, 0] = 12345.0
    assert dcb.softsplat(t2, zero, m2, "soft")[:, :, 0, 0].abs().max() == 0


def test_mass_conservation_1080p(dcb):
    """Size-independent property at BASELINE's full frame size: with every corner in range,
    sum(out) == sum(in) to fp32 reassociation error."""
    g = torch.Generator(device="cuda").manual_seed(0)
    tin = torch.rand(1, 3, 1080, 1920, device="cuda", generator=g)
    flow = torch.randn(1, 2, 1080, 1920, device="cuda", generator=g) * 8
    ys, xs = torch.meshgrid(torch.arange(1080, device="cuda"), torch.arange(1920, device="cuda"), indexing="ij")
    flow[:, 0] = torch.minimum(torch.maximum(flow[:, 0], 0.25 - xs), 1918.5 - xs)
    flow[:, 1] = torch.minimum(torch.maximum(flow[:, 1], 0.25 - ys), 1078.5 - ys)
    out = dcb.softsplat(tin, flow, None, "sum")
    assert abs(out.double().sum().item() - tin.double().sum().item()) < 1e-6 * tin.double().sum().item()
    # avg: where anything landed, a constant image stays constant
    ones = torch.full_like(tin, 3.0)
    avg = dcb.softsplat(ones, flow, None, "avg")
    hit = dcb.softsplat(torch.ones_like(tin[:, :1]), flow, None, "sum") > 0.05   # the 1e-7 of the normaliser is a 2e-6 relative effect there
    assert_close(avg[:, :1][hit], torch.full_like(avg[:, :1][hit], 3.0), 1e-4, "avg of constant")


def test_stride_contract(dcb):
    """channels_last and sliced inputs give the same result as .contiguous(); output is NCHW-contiguous."""
    tin = torch.randn(2, 6, 20, 28, device="cuda")
    flow = torch.randn(2, 4, 20, 28, device="cuda") * 2
    metric = torch.randn(2, 3, 20, 28, device="cuda") * 0.3
    with dcb.deterministic(True):
        base = dcb.softsplat(tin[:, :3].contiguous(), flow[:, 1:3].contiguous(), metric[:, 1:2].contiguous(), "soft")
        a = dcb.softsplat(tin[:, :3], flow[:, 1:3], metric[:, 1:2], "soft")
        b = dcb.softsplat(tin[:, :3].contiguous(memory_format=torch.channels_last), flow[:, 1:3], metric[:, 1:2], "soft")
        c = dcb.softsplat(tin.permute(0, 1, 3, 2)[:, :3].permute(0, 1, 3, 2), flow[:, 1:3], metric[:, 1:2], "soft")
    assert a.is_contiguous() and torch.equal(a, base) and torch.equal(b, base) and torch.equal(c, base)
    # strided upstream gradient
    ti = tin[:, :3].clone().requires_grad_(True)
    out = dcb.softsplat(ti, flow[:, :2], None, "avg")
    g = torch.randn(2, 3, 28, 20, device="cuda").permute(0, 1, 3, 2)
    out.backward(g)
    ti2 = tin[:, :3].clone().requires_grad_(True)
    dcb.softsplat(ti2, flow[:, :2].contiguous(), None, "avg").backward(g.contiguous())
    assert_close(ti.grad, ti2.grad, 1e-6, "strided grad")


def test_needs_input_grad_gating(dcb):
    tin = torch.randn(1, 3, 8, 8, device="cuda", requires_grad=True)
    flow = torch.randn(1, 2, 8, 8, device="cuda")
    metric = torch.randn(1, 1, 8, 8, device="cuda", requires_grad=True)
    dcb.softsplat(tin, flow, metric, "soft").sum().backward()
    assert tin.grad is not None and metric.grad is not None and flow.grad is None


def test_autocast_casts_to_fp32(dcb):
    tin = torch.randn(1, 3, 8, 8, device="cuda", dtype=torch.float16)
    flow = torch.randn(1, 2, 8, 8, device="cuda", dtype=torch.float16)
    with torch.autocast("cuda", dtype=torch.float16):
        out = dcb.softsplat(tin, flow, None, "avg")
    assert out.dtype == torch.float32          # custom_fwd(cast_inputs=float32), softsplat.py:279


def test_asserts_like_reference(dcb):
    tin = torch.randn(1, 3, 8, 8, device="cuda"); flow = torch.zeros(1, 2, 8, 8, device="cuda"); m = torch.zeros(1, 1, 8, 8, device="cuda")
    for args in [(tin, flow, m, "sum"), (tin, flow, m, "avg"), (tin, flow, None, "soft"), (tin, flow, None, "linear"), (tin, flow, None, "max")]:
        with pytest.raises(AssertionError):
            dcb.softsplat(*args)
    with pytest.raises(AssertionError):
        dcb.softsplat(tin.cpu(), flow.cpu(), None, "sum")            # no CPU path, like the reference
    with pytest.raises(AssertionError):
        dcb.softsplat(tin, torch.zeros(1, 2, 8, 9, device="cuda"), None, "sum")


def test_empty_and_ragged(dcb):
    out = dcb.softsplat(torch.zeros(0, 3, 8, 8, device="cuda"), torch.zeros(0, 2, 8, 8, device="cuda"), None, "avg")
    assert out.shape == (0, 3, 8, 8)
    for shape in [(1, 1, 1, 1), (1, 3, 1, 37), (2, 2, 37, 1), (1, 5, 3, 257)]:
        tin = torch.randn(*shape, device="cuda"); fl = torch.randn(shape[0], 2, shape[2], shape[3], device="cuda")
        from oracle import oracle as orc
        assert_close(dcb.softsplat(tin, fl, None, "avg"), orc.softsplat(tin.cpu(), fl.cpu(), None, "avg"), 1e-5, str(shape))


def test_small_spatial_many_channels(dcb, orc):
    """ControlNet pyramid shapes (extractors.py:245-248): channel-sliced backward path."""
    for (n, c, r) in [(2, 160, 32), (2, 320, 16), (2, 640, 8)]:
        tin, flow, metric, gout = make_inputs(c, n, c, r, r, flow_scale=0.7)
        ref = oracle_run(orc, tin, flow, metric, gout, "soft")
        truth = oracle_run(orc, tin.double(), flow.double(), metric.double(), gout.double(), "soft")
        got = cuda_run(dcb.softsplat, tin, flow, metric, gout, "soft")
        for k in ("out", "gin", "gflow", "gmetric"):
            assert_close(got[k], ref[k], 2e-5, f"pyramid {c}x{r} {k}", truth=truth[k])


@pytest.mark.parametrize("mode", ["sum", "avg", "linear", "soft", "avg-zeroeps", "soft-clipeps"])
def test_large_many_channel_tensors(dcb, orc, mode):
    """>= 16 MB with C + 1 > 4: the per-target list path (count / scan / fill / gather). Collisions,
    holes and out-of-frame corners from a rough flow; forward and all gradients vs the oracle."""
    tin, flow, metric, gout = make_inputs(77, 2, 35, 192, 320, flow_scale=4.0)
    flow[0, :, 3, 5] = float("nan"); flow[1, 0, 7, 9] = 1e9; flow[1, :, 100:110, 200:210] = -500.0
    ref = oracle_run(orc, tin, flow, metric, gout, mode)
    truth = oracle_run(orc, tin.double(), flow.double(), metric.double(), gout.double(), mode)
    got = cuda_run(dcb.softsplat, tin, flow, metric, gout, mode)
    for k in ("out", "gin", "gflow", "gmetric"):
        if ref[k] is not None:
            assert_close(got[k], ref[k], 2e-5, f"lists {mode} {k}", truth=truth[k])
    # the shared all-zero workspace must still be clean for the accumulator paths
    small = make_inputs(78, 1, 3, 24, 40, flow_scale=2.0)
    assert_close(dcb.softsplat(small[0].cuda(), small[1].cuda(), small[2].cuda(), "soft"),
                 orc.softsplat(small[0], small[1], small[2], "soft"), 1e-5, "pipe after lists")
    assert_close(dcb.softsplat(tin[:1, :9, :32, :32].cuda(), flow[:1, :, :32, :32].cuda(), None, "avg"),
                 orc.softsplat(tin[:1, :9, :32, :32], flow[:1, :, :32, :32], None, "avg"), 1e-5, "planar after lists")


def test_large_many_channel_bf16_and_frame_groups(dcb, orc):
    """bf16 through the list path, and more frames than one 64 MB list group (two groups at 512 x 768)."""
    tin, flow, metric, _ = make_inputs(79, 2, 70, 192, 320, flow_scale=3.0)
    tb, mb = tin.bfloat16(), metric.bfloat16()
    ref = orc.softsplat(tb.float(), flow, mb.float(), "soft")
    got = dcb.softsplat(tb.cuda(), flow.cuda(), mb.cuda(), "soft")
    assert got.dtype == torch.bfloat16
    assert_close(got.float(), ref, 1e-2, "lists bf16 (fp32 flow)")
    tin, flow, metric, _ = make_inputs(80, 3, 16, 512, 768, flow_scale=3.0)
    assert_close(dcb.softsplat(tin.cuda(), flow.cuda(), metric.cuda(), "soft"), orc.softsplat(tin, flow, metric, "soft"), 1e-5, "lists groups")


def test_list_path_agrees_with_accumulator_path(dcb):
    """Two independent GPU implementations of the many-channel forward: the whole batch goes through the
    per-target lists (>= 16 MB), single frames of it through the channel-quad accumulators."""
    g = torch.Generator().manual_seed(81)
    tin = torch.randn(4, 24, 192, 256, generator=g).cuda()
    flow = (torch.randn(4, 2, 192, 256, generator=g) * 5).cuda()
    metric = (torch.randn(4, 1, 192, 256, generator=g) * 0.5).cuda()
    for mode in ("sum", "avg", "soft"):
        me = metric if mode == "soft" else None
        whole = dcb.softsplat(tin, flow, me, mode)
        for n in range(4):
            part = dcb.softsplat(tin[n:n + 1], flow[n:n + 1], None if me is None else me[n:n + 1], mode)
            assert_close(whole[n:n + 1], part, 1e-5, f"lists vs accumulators {mode} frame {n}")


def test_list_path_ragged_sizes(dcb, orc):
    """Frame sizes that are multiples of neither the 32 x 8 gather tile nor the 256-pixel CTAs of the
    list builders (203 x 331), odd channel count (21: a short last channel block), three frames."""
    tin, flow, metric, gout = make_inputs(84, 3, 21, 203, 331, flow_scale=3.5)
    assert tin.numel() * 4 >= 16 << 20
    for mode in ("avg", "soft"):
        ref = oracle_run(orc, tin, flow, metric, gout, mode)
        truth = oracle_run(orc, tin.double(), flow.double(), metric.double(), gout.double(), mode)
        got = cuda_run(dcb.softsplat, tin, flow, metric, gout, mode)
        for k in ("out", "gin", "gflow", "gmetric"):
            if ref[k] is not None:
                assert_close(got[k], ref[k], 2e-5, f"ragged lists {mode} {k}", truth=truth[k])


def test_list_path_mask_strides_and_saved_normaliser(dcb, orc):
    """The list path behind FeatureWarperSoftsplat's call shape: (1 - mask) product, a channel-sliced
    (non-contiguous) input, and the gradient w.r.t. everything through the saved normaliser."""
    g = torch.Generator().manual_seed(83)
    n, c, h, w = 2, 24, 192, 256
    big = torch.randn(n, c + 8, h, w, generator=g)
    flow = torch.randn(n, 2, h, w, generator=g) * 3
    metric = torch.randn(n, 1, h, w, generator=g) * 0.5
    mask = (torch.rand(n, 1, h, w, generator=g) > 0.7).float()
    gout = torch.randn(n, c, h, w, generator=g)
    # oracle: soft splat of the channel slice, times (1 - mask)   (control_utils.py:62-70)
    ti = big[:, 4:4 + c].clone().requires_grad_(True); fl = flow.clone().requires_grad_(True); me = metric.clone().requires_grad_(True)
    ref = orc.softsplat(ti, fl, me, "soft") * (1 - mask)
    ref.backward(gout)
    bigc = big.cuda()
    tic = bigc[:, 4:4 + c].requires_grad_(True)              # a view: strides of the 32-channel tensor
    assert not tic.is_contiguous()
    flc = flow.cuda().requires_grad_(True); mec = metric.cuda().requires_grad_(True)
    import importlib
    impl = importlib.import_module(dcb.__name__ + ".softsplat")      # the submodule (the package re-exports the function under its name)
    got = impl._splat_normalised(tic, flc, mec, dcb._lib.MODE_SOFT, dcb._lib.EPS_ADD, mask=mask.cuda())
    got.backward(gout.cuda())
    assert_close(got, ref, 1e-5, "lists + mask + strides")
    assert_close(tic.grad, ti.grad, 2e-5, "gin")
    assert_close(flc.grad, fl.grad, 1e-4, "gflow")
    assert_close(mec.grad, me.grad, 1e-4, "gmetric")


def test_mass_conservation_c4_shape(dcb):
    """BASELINE config C4 (8 x 64 x 256 x 256) through the list path: a 'sum' splat whose footprints all stay
    inside the frame moves mass without creating or losing any; 'avg' of a constant is that constant."""
    g = torch.Generator(device="cuda").manual_seed(82)
    tin = torch.rand(8, 64, 256, 256, device="cuda", generator=g)
    low = torch.randn(8, 2, 8, 8, device="cuda", generator=g)
    flow = torch.nn.functional.interpolate(low, size=(256, 256), mode="bicubic") * 3
    border = torch.zeros(1, 1, 256, 256, device="cuda"); border[..., 12:-12, 12:-12] = 1
    flow = flow * border                                   # nothing leaves the frame
    out = dcb.softsplat(tin, flow, None, "sum")
    assert_close(out.double().sum(dim=(2, 3)), tin.double().sum(dim=(2, 3)), 1e-6, "mass per (frame, channel)")
    const = torch.full_like(tin, 0.75)
    avg = dcb.softsplat(const, flow, None, "avg")
    covered = dcb.softsplat(torch.ones(8, 1, 256, 256, device="cuda"), flow, None, "sum") > 0.05   # the 1e-7 of the normaliser is a 2e-6 relative effect there
    assert_close(avg[covered.expand_as(avg)], const[covered.expand_as(avg)], 1e-5, "avg of a constant")


def test_bf16_forward_wild_flows(dcb, orc):
    """bf16 forward on large flows (holes, collisions): values within 1e-2 of the fp32 reference."""
    tin, flow, metric, _ = make_inputs(8, 2, 3, 40, 56, flow_scale=6.0)
    tb, fb, mb = tin.bfloat16(), flow.bfloat16(), metric.bfloat16()
    for mode in ("sum", "avg", "soft"):
        me = mb if mode == "soft" else None
        ref = orc.softsplat(tb.float(), fb.float(), None if me is None else me.float(), mode)
        got = dcb.softsplat(tb.cuda(), fb.cuda(), None if me is None else me.cuda(), mode)
        assert_close(got.float(), ref, 1e-2, f"bf16 wild {mode}")


def test_multi_frame_pipeline_many_frames(dcb, orc):
    """Seven small frames, forward + all gradients, three times on the self-cleaning workspace. (At this size all
    seven frames form ONE frame group; the many-group schedules are driven by tests/test_baseline_shapes_gpu.py.)"""
    tin, flow, metric, gout = make_inputs(31, 7, 3, 70, 150, flow_scale=3.0)
    ref = oracle_run(orc, tin, flow, metric, gout, "soft")
    truth = oracle_run(orc, tin.double(), flow.double(), metric.double(), gout.double(), "soft")
    for _ in range(3):     # repeated calls reuse the self-cleaning workspace
        got = cuda_run(dcb.softsplat, tin, flow, metric, gout, "soft")
        for k in ("out", "gin", "gflow", "gmetric"):
            assert_close(got[k], ref[k], 1e-5, f"7 frames {k}", truth=truth[k])
    assert_close(dcb.softsplat(tin[:2].cuda(), flow[:2].cuda(), None, "avg"), orc.softsplat(tin[:2], flow[:2], None, "avg"), 1e-5, "2 frames avg")


def test_softsplat_host_matches_device(dcb):
    """Host-buffer entry point (chunked, three streams) == the device op on the same frames."""
    tin, flow, metric, _ = make_inputs(51, 7, 3, 60, 90, flow_scale=3.0)
    for mode, me in (("soft", metric), ("avg", None)):
        ref = dcb.softsplat(tin.cuda(), flow.cuda(), None if me is None else me.cuda(), mode).cpu()
        got = dcb.softsplat_host(tin.pin_memory(), flow.pin_memory(), None if me is None else me.pin_memory(), mode, chunk_frames=3)
        torch.cuda.synchronize()
        assert not got.is_cuda and got.shape == ref.shape
        assert_close(got, ref, 1e-5, f"host {mode}")
    # ramped schedule (1, 2, 4, 4 ..., 2, 1), staging buffers reused by a second call with new data
    for seed in (52, 53):
        tin, flow, metric, _ = make_inputs(seed, 23, 3, 24, 40, flow_scale=3.0)
        ref = dcb.softsplat(tin.cuda(), flow.cuda(), metric.cuda(), "soft").cpu()
        got = dcb.softsplat_host(tin.pin_memory(), flow.pin_memory(), metric.pin_memory(), "soft", chunk_frames=4)
        torch.cuda.synchronize()
        assert_close(got, ref, 1e-5, "host ramped")


def test_softsplat_host_compact_element_types(dcb, orc):
    """8-bit frames, half-precision flow / metric, bf16 result: uploaded as they are, widened to fp32 on the device
    (dcb_convert); within 1e-2 of the fp32 oracle on the same (up-cast) values. Also the synchronous return: the result
    is complete when the call returns, without any torch.cuda.synchronize()."""
    g = torch.Generator().manual_seed(61)
    frames = torch.randint(0, 256, (9, 3, 40, 56), generator=g, dtype=torch.uint8)
    flow = (torch.randn(9, 2, 40, 56, generator=g) * 2).half()
    metric = (-torch.rand(9, 1, 40, 56, generator=g)).half()
    ref = orc.softsplat(frames.float() / 255.0, flow.float(), metric.float(), "soft")
    out = torch.empty(9, 3, 40, 56, dtype=torch.bfloat16).pin_memory()
    got = dcb.softsplat_host(frames.pin_memory(), flow.pin_memory(), metric.pin_memory(), "soft", out=out, chunk_frames=4)
    assert got is out                                   # no synchronize here on purpose
    assert_close(got.float(), ref, 1e-2, "host compact bf16 result")
    got32 = dcb.softsplat_host(frames.pin_memory(), flow.bfloat16().pin_memory(), None, "avg", chunk_frames=2)
    assert got32.dtype == torch.float32
    assert_close(got32, orc.softsplat(frames.float() / 255.0, flow.bfloat16().float(), None, "avg"), 1e-5, "host u8 -> fp32 result")


def test_convert_kernel(dcb):
    L = dcb._lib
    g = torch.Generator(device="cuda").manual_seed(3)
    for n in (1, 7, 8, 1000, 4099):
        src = torch.randint(0, 256, (n,), device="cuda", generator=g, dtype=torch.uint8)
        assert torch.equal(L.convert(src, torch.empty(n, device="cuda")), src.float())
        assert torch.equal(L.convert(src, torch.empty(n, device="cuda"), 1 / 255.0), src.float() * (1 / 255.0))
        x = torch.randn(n, device="cuda", generator=g)
        assert torch.equal(L.convert(x, torch.empty(n, device="cuda", dtype=torch.bfloat16)), x.bfloat16())
        assert torch.equal(L.convert(x.half(), torch.empty(n, device="cuda")), x.half().float())
        assert torch.equal(L.convert(x[1:], torch.empty(max(n - 1, 0), device="cuda", dtype=torch.float16)), x[1:].half())   # unaligned source


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 32, 40, 56), (1, 160, 32, 32), (3, 8, 17, 23)])
def test_channels_last_vector_path(dcb, orc, dtype, shape):
    """NHWC (channels_last) feature maps: the many-channel scatter reads each channel quad with ONE 16-byte (fp32) / 8-byte
    (bf16) load instead of four strided scalar loads (north_star: coalesced float4 / bf16 NHWC loads; the reference accepts
    any strides, softsplat.py:170-207). Same values as the NCHW call and as the oracle; a sliced view whose quads are not
    aligned falls back to the scalar loads."""
    n, c, h, w = shape
    tin, flow, metric, _ = make_inputs(71, n, c, h, w, flow_scale=1.5)
    tin, flow, metric = (t.to(dtype).float() for t in (tin, flow, metric))        # the oracle sees the values the kernels see
    ref = orc.softsplat(tin, flow, metric, "soft")
    rel = 1e-5 if dtype == torch.float32 else 1e-2
    x = tin.cuda().to(dtype)
    fl, me = flow.cuda().to(dtype), metric.cuda().to(dtype)
    nhwc = x.to(memory_format=torch.channels_last)
    assert nhwc.stride(1) == 1
    a = dcb.softsplat(x, fl, me, "soft")
    b = dcb.softsplat(nhwc, fl, me, "soft")
    assert b.is_contiguous()
    assert_close(b.float(), a.float(), 2e-6 if dtype == torch.float32 else 1e-2, "channels_last vs NCHW")
    assert_close(b.float().cpu(), ref, rel, "channels_last vs oracle")
    odd = torch.zeros(n, h, w, c + 1, device="cuda", dtype=dtype)[..., 1:].permute(0, 3, 1, 2)      # quads start 1 element off: not aligned
    odd.copy_(x)
    assert_close(dcb.softsplat(odd, fl, me, "soft").float(), a.float(), 2e-6 if dtype == torch.float32 else 1e-2, "unaligned NHWC view")


@pytest.mark.parametrize("shape,dtype,mode", [((2, 64, 192, 200), torch.float32, "soft"), ((2, 64, 192, 200), torch.float32, "sum"),
                                              ((3, 160, 150, 152), torch.float32, "avg"), ((2, 40, 230, 250), torch.float32, "linear"),
                                              ((2, 96, 256, 260), torch.bfloat16, "soft")])
def test_channels_last_list_gather(dcb, orc, shape, dtype, mode):
    """Many channels on large tensors (the per-target list path, csrc/splat_lists.cu): channels_last input takes the
    channel-quad gather (k_list_gather_nhwc: sub-warp groups over the channels of one target, NCHW output through a
    shared-memory transpose). Same values as the NCHW call, as the NCHW gather on the same strided view, and as the oracle;
    covers ragged row tiles (W % 32 != 0), quad counts that are no power of two and more channels than one transpose pass."""
    n, c, h, w = shape
    tin, flow, metric, _ = make_inputs(83, n, c, h, w, flow_scale=2.5, smooth=True)
    tin, flow, metric = (t.to(dtype).float() for t in (tin, flow, metric))
    if mode == "linear":
        metric = metric.abs() + 0.1
    me_ref = metric if mode in ("soft", "linear") else None
    ref = orc.softsplat(tin, flow, me_ref, mode)
    rel = 1e-5 if dtype == torch.float32 else 1e-2
    x = tin.cuda().to(dtype)
    fl = flow.cuda().to(dtype)
    me = metric.cuda().to(dtype) if me_ref is not None else None
    nhwc = x.to(memory_format=torch.channels_last)
    assert nhwc.stride(1) == 1 and not nhwc.is_contiguous()
    a = dcb.softsplat(x, fl, me, mode)
    before = dcb.launch_count()
    b = dcb.softsplat(nhwc, fl, me, mode)
    assert dcb.launch_count() - before == 4          # count, alloc, fill, gather: the list path
    assert b.is_contiguous()
    dcb._lib.set_option("lists_nhwc", 0)
    try:
        b0 = dcb.softsplat(nhwc, fl, me, mode)
    finally:
        dcb._lib.set_option("lists_nhwc", 1)
    tight = 2e-6 if dtype == torch.float32 else 1e-2
    assert_close(b.float(), a.float(), tight, "channels_last vs NCHW")
    assert_close(b.float(), b0.float(), tight, "quad gather vs NCHW gather on the same view")
    assert_close(b.float().cpu(), ref, rel, "channels_last vs oracle")


@pytest.mark.parametrize("shape,dtype,mode", [((4, 4, 135, 240), torch.bfloat16, "soft"), ((2, 12, 40, 56), torch.float32, "avg"),
                                              ((1, 160, 32, 32), torch.float32, "soft"), ((3, 7, 33, 47), torch.float32, "sum")])
def test_planar_single_launch_kernel(dcb, orc, shape, dtype, mode):
    """The opt-in single-launch form of the many-channel pipeline (k_planar_one: scatter, grid barrier on two counters in the
    workspace, normalise): ONE launch, the same values as the two-launch pipeline and the oracle, and the counters back at
    zero afterwards (a second call and a different call through the same kept-zero workspace are right)."""
    n, c, h, w = shape
    tin, flow, metric, _ = make_inputs(91, n, c, h, w, flow_scale=1.5)
    tin, flow, metric = (t.to(dtype).float() for t in (tin, flow, metric))
    me_ref = metric if mode in ("soft", "linear") else None
    ref = orc.softsplat(tin, flow, me_ref, mode)
    x, fl = tin.cuda().to(dtype), flow.cuda().to(dtype)
    me = metric.cuda().to(dtype) if me_ref is not None else None
    L = dcb._lib
    L.set_option("fwd_path", 1)               # keep small frames off the cluster kernel
    try:
        two = dcb.softsplat(x, fl, me, mode)
        L.set_option("planar_one_launch", 1)
        before = dcb.launch_count()
        one = dcb.softsplat(x, fl, me, mode)
        assert dcb.launch_count() - before == 1
        again = dcb.softsplat(x, fl, me, mode)
        L.set_option("planar_one_launch", 0)
        after = dcb.softsplat(x, fl, me, mode)
    finally:
        L.set_option("planar_one_launch", 0)
        L.set_option("fwd_path", 0)
    rel = 1e-5 if dtype == torch.float32 else 1e-2
    assert_close(one.float().cpu(), ref, rel, "single launch vs oracle")
    assert_close(again.float().cpu(), ref, rel, "second single launch vs oracle")
    assert_close(after.float().cpu(), ref, rel, "two launches after the single-launch calls")
    assert_close(one.float(), two.float(), 2e-6 if dtype == torch.float32 else 1e-2, "single launch vs two launches")


def test_more_frames_than_a_grid_dimension(dcb, orc):
    """70001 frames of 2 x 3 pixels: the forward's and the packed backward's 3-d grids put frames on gridDim.z (<= 65535), so
    both must cut the batch into several launches (softsplat.py accepts any N)."""
    n, c, h, w = 70001, 2, 2, 3
    tin, flow, metric, gout = make_inputs(5, n, c, h, w, flow_scale=0.8)
    ref = oracle_run(orc, tin, flow, metric, gout, "soft")
    got = cuda_run(dcb.softsplat, tin, flow, metric, gout, "soft")
    for k in ("out", "gin", "gflow", "gmetric"):
        assert_close(got[k], ref[k], 1e-5, k)


@pytest.mark.parametrize("mode", ["soft", "sum"])
def test_list_path_long_lists(dcb, orc, mode):
    """A strongly compressive flow (every pixel moves 60 % of the way to the frame centre: ~25 sources per target there, none
    elsewhere) on a tensor that takes the per-target list path: long and empty lists side by side (csrc/splat_lists.cu).
    NCHW and channels_last gathers vs the oracle."""
    n, c, h, w = 2, 64, 192, 200
    tin, _, metric, _ = make_inputs(97, n, c, h, w)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    flow = torch.stack([-0.6 * (xs - w / 2 + 0.3), -0.6 * (ys - h / 2 - 0.2)], 0).unsqueeze(0).repeat(n, 1, 1, 1)
    flow = flow + 0.05 * torch.randn(n, 2, h, w, generator=torch.Generator().manual_seed(3))
    me_ref = metric if mode == "soft" else None
    ref = orc.softsplat(tin, flow, me_ref, mode)
    x, fl = tin.cuda(), flow.cuda()
    me = metric.cuda() if me_ref is not None else None
    before = dcb.launch_count()
    a = dcb.softsplat(x, fl, me, mode)
    assert dcb.launch_count() - before == 4          # count, alloc, fill, gather
    b = dcb.softsplat(x.to(memory_format=torch.channels_last), fl, me, mode)
    # long lists are summed in a different order than the oracle's: 25 terms of fp32
    assert_close(a.cpu(), ref, 2e-5, "NCHW gather, long lists")
    assert_close(b.cpu(), ref, 2e-5, "channels_last gather, long lists")
