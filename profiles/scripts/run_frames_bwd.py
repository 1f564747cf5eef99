"""Forward + backward of the soft splat on 1080p frames (C = 3), 16 frames: us per frame and GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator(device="cuda").manual_seed(0)
tin = torch.rand(F, 3, 1080, 1920, device="cuda", generator=g).requires_grad_(True)
low = torch.randn(F, 2, 34, 60, device="cuda", generator=g)
flow = (torch.nn.functional.interpolate(low, size=(1080, 1920), mode="bicubic") * 8).requires_grad_(True)
metric = (-torch.rand(F, 1, 1080, 1920, device="cuda", generator=g)).requires_grad_(True)
go = torch.randn(F, 3, 1080, 1920, device="cuda", generator=g)
px = F * 1080 * 1920
for i in range(4):
    tin.grad = flow.grad = metric.grad = None
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record(); out = d.softsplat(tin, flow, metric, "soft"); b.record(); out.backward(go); c.record(); torch.cuda.synchronize()
    tf, tb = a.elapsed_time(b), b.elapsed_time(c)
    print(f"iter {i}: fwd {tf*1e3/F:.1f} us/frame, bwd {tb*1e3/F:.1f} us/frame ({76*px/tb/1e6:.0f} GB/s of (4C+7)*4 = 76 B/px), fwd+bwd {112*px/(tf+tb)/1e6:.0f} GB/s")
