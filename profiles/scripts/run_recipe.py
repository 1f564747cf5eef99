"""Tiny driver for profiling: the conditioning recipe (dcb_residual_fused) on 1080p frames (used under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
variant = sys.argv[3] if len(sys.argv) > 3 else "dataset"
g = torch.Generator(device="cuda").manual_seed(0)
H, W = 1080, 1920


def smooth(amp):
    low = torch.randn(frames, 2, H // 32, W // 32, device="cuda", generator=g)
    return torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False) * amp


img = torch.rand(frames, 3, H, W, device="cuda", generator=g); gt = torch.rand(frames, 3, H, W, device="cuda", generator=g)
f1 = smooth(8.0); f2 = -f1 + 0.5 * smooth(1.0)
torch.cuda.synchronize()
for i in range(iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fused, res = d.residual_conditioning(img, f1, f2, gt, variant)
    b.record()
    torch.cuda.synchronize()
    print(f"iter {i}: {a.elapsed_time(b) * 1e3 / frames:.1f} us per frame ({frames} frames)")
print("checksum", float(fused.double().sum()), float(res.double().sum()))
