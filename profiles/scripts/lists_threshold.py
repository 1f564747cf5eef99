"""Where do the per-target lists beat the accumulator pipeline? Forward soft splat, fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
shapes = [(2, 320, 32, 32), (2, 320, 64, 64), (1, 64, 128, 128), (2, 64, 128, 128), (4, 64, 128, 128), (8, 32, 128, 128),
          (2, 640, 32, 32), (1, 64, 256, 256), (2, 64, 256, 256), (1, 16, 512, 512), (1, 8, 1080, 1920), (1, 16, 1080, 1920), (4, 8, 540, 960), (1, 32, 540, 960), (16, 4, 256, 256), (16, 8, 256, 256), (16, 16, 256, 256)]
g = torch.Generator(device="cuda").manual_seed(5)
for (n, c, h, w) in shapes:
    ti = torch.randn(n, c, h, w, device="cuda", generator=g)
    me = torch.randn(n, 1, h, w, device="cuda", generator=g) * 0.5
    fl = torch.randn(n, 2, h, w, device="cuda", generator=g) * 2
    for _ in range(5): d.softsplat(ti, fl, me, "soft")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): d.softsplat(ti, fl, me, "soft")
    b.record(); torch.cuda.synchronize()
    print(f"{n}x{c}x{h}x{w} ({n*c*h*w*4/2**20:6.1f} MB): {a.elapsed_time(b)/20*1e3:8.1f} us")
