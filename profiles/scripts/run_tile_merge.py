"""Profiling driver: Hann-window tile merge, 15 tiles of 512^2 (64 px overlap) -> 3 x 1080 x 1920 fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
H, W, ts, ov, ch = 1080, 1920, 512, 64, 3
g = torch.Generator(device="cuda").manual_seed(4)
rects = [(y, min(y + ts, H), x, min(x + ts, W)) for y in range(0, H, ts - ov) for x in range(0, W, ts - ov)]
tiles = [torch.randn(1, ch, y2 - y1, x2 - x1, device="cuda", generator=g) for (y1, y2, x1, x2) in rects]
as_read = [(x1, x2, y1, y2) for (y1, y2, x1, x2) in rects]
for i in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = d.merge_latent_tiles_from_pixel_coords(tiles, as_read, (1, ch, H, W), (H, W)); b.record(); torch.cuda.synchronize()
    print(f"iter {i}: {a.elapsed_time(b)*1e3:.1f} us")
