"""Latency of small forward splats: eager loop and CUDA-graph replay (DCB_NO_SMALL=1 disables the single-launch path)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
g = torch.Generator(device="cuda").manual_seed(1)
cases = [("pipe 4x3x135x240 f32", (4, 3, 135, 240), torch.float32), ("planar 4x4x135x240 bf16", (4, 4, 135, 240), torch.bfloat16),
         ("planar 2x320x32x32 f32", (2, 320, 32, 32), torch.float32), ("planar 2x640x16x16 f32", (2, 640, 16, 16), torch.float32),
         ("pipe 1x3x256x256 f32", (1, 3, 256, 256), torch.float32), ("pipe 1x3x540x960 f32", (1, 3, 540, 960), torch.float32)]
for name, (n, c, h, w), dt in cases:
    ti = torch.randn(n, c, h, w, device="cuda", generator=g).to(dt)
    me = (-torch.rand(n, 1, h, w, device="cuda", generator=g)).to(dt)
    fl = (torch.randn(n, 2, h, w, device="cuda", generator=g)).to(dt)
    f = lambda: d.softsplat(ti, fl, me, "soft")
    for _ in range(50): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2000): f()
    torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / 2000 * 1e6
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): f()
        torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(20): out = f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(100): gr.replay()
    torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / 2000 * 1e6
    print(f"{name:28s} eager {eager:6.1f} us/call   graph {graph:6.1f} us/call")
