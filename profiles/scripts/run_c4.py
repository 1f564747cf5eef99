"""Profiling driver: C4 shape (8x64x256x256 fp32 soft), forward only or forward+backward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
g = torch.Generator(device="cuda").manual_seed(3)
ti = torch.randn(8, 64, 256, 256, device="cuda", generator=g).requires_grad_(bwd)
me = (torch.randn(8, 1, 256, 256, device="cuda", generator=g) * 0.5).requires_grad_(bwd)
low = torch.randn(8, 2, 8, 8, device="cuda", generator=g)
fl = (torch.nn.functional.interpolate(low, size=(256, 256), mode="bicubic") * 4).requires_grad_(bwd)
go = torch.randn(8, 64, 256, 256, device="cuda", generator=g)
for i in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = d.softsplat(ti, fl, me, "soft")
    if bwd:
        ti.grad = me.grad = fl.grad = None
        out.backward(go)
    b.record(); torch.cuda.synchronize()
    print(f"iter {i}: {a.elapsed_time(b)*1e3:.1f} us")
