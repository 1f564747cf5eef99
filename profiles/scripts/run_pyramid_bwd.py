"""Soft splat forward / forward + backward at ControlNet pyramid shapes (us per call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
g = torch.Generator(device="cuda").manual_seed(9)
for (n, c, r) in [(2, 1280, 8), (2, 640, 16), (2, 320, 32), (2, 320, 64)]:
    ti = torch.randn(n, c, r, r, device="cuda", generator=g, requires_grad=True)
    fl = (torch.randn(n, 2, r, r, device="cuda", generator=g) * 0.7).requires_grad_(True)
    me = (torch.randn(n, 1, r, r, device="cuda", generator=g) * 0.5).requires_grad_(True)
    go = torch.randn(n, c, r, r, device="cuda", generator=g)
    def fwd():
        with torch.no_grad(): return d.softsplat(ti, fl, me, "soft")
    def fb():
        ti.grad = fl.grad = me.grad = None
        d.softsplat(ti, fl, me, "soft").backward(go)
    res = []
    for f in (fwd, fb):
        for _ in range(5): f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): f()
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 50 * 1e3)
    print(f"{n}x{c}x{r}x{r}: fwd {res[0]:7.1f} us   fwd+bwd {res[1]:7.1f} us")
