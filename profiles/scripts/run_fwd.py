"""Tiny driver for profiling: a few forward splats of 1080p frames (used under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = sys.argv[2] if len(sys.argv) > 2 else "soft"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(0)
tin = torch.rand(frames, 3, 1080, 1920, device="cuda", generator=g)
low = torch.randn(frames, 2, 34, 60, device="cuda", generator=g)
flow = torch.nn.functional.interpolate(low, size=(1080, 1920), mode="bicubic") * 8
metric = -torch.rand(frames, 1, 1080, 1920, device="cuda", generator=g) if mode in ("soft", "linear") else None
torch.cuda.synchronize()
for i in range(iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = d.softsplat(tin, flow, metric, mode)
    b.record()
    torch.cuda.synchronize()
    print(f"iter {i}: {a.elapsed_time(b) * 1e3:.1f} us for {frames} frame(s) -> {frames * 1080 * 1920 / a.elapsed_time(b) / 1e3:.0f} Mpx/s")
print("checksum", float(out.double().sum()))
