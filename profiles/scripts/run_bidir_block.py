"""One bi-directional block (masks, both masked splats, fusion) forward + backward at the four ControlNet
pyramid shapes; meant to be run under `ncu --metrics gpu__time_duration.sum` to see every kernel's GPU time
(the event timings of such small calls are dominated by Python + autograd)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
g = torch.Generator(device="cuda").manual_seed(9)
warper = d.FeatureWarperSoftsplat(with_learnable_metric=False).cuda()
for (n, c, r) in [(2, 1280, 8), (2, 640, 16), (2, 320, 32), (2, 320, 64)]:
    fa = torch.randn(n, c, r, r, device="cuda", generator=g, requires_grad=True); fb = torch.randn(n, c, r, r, device="cuda", generator=g, requires_grad=True)
    ff = torch.randn(n, 2, r, r, device="cuda", generator=g) * 0.7; fbw = -ff + 0.1 * torch.randn(n, 2, r, r, device="cuda", generator=g)
    go = torch.randn(n, c, r, r, device="cuda", generator=g)
    for i in range(3):
        fa.grad = fb.grad = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = d.bidirectional_warp_fuse(fa, fb, ff, fbw, warper); out.backward(go); b.record(); torch.cuda.synchronize()
    print(f"{n}x{c}x{r}x{r}: block fwd+bwd {a.elapsed_time(b)*1e3:.0f} us")
