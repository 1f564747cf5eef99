"""Latency of small / single-frame forward calls per kernel family (fwd_path 0 = automatic: cluster kernel for small frames, 1 = accumulator pipelines):
C2 latents, C1 single 1080p frame, occlusion masks and feature splats at the ControlNet pyramid sizes. Eager and CUDA-graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def graphed(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            fn()
    torch.cuda.synchronize()
    return gr.replay


cases = {}
lat = (torch.randn(4, 4, 135, 240, device=dev, generator=g) * 0.18215).bfloat16()
met = (-torch.randn(4, 1, 135, 240, device=dev, generator=g).abs()).bfloat16()
fl = torch.randn(4, 2, 135, 240, device=dev, generator=g).bfloat16()
cases["C2 soft 4x4x135x240 bf16"] = lambda: d.softsplat(lat, fl, met, "soft")
pool = [(torch.rand(1, 3, 1080, 1920, device=dev, generator=g),
         torch.nn.functional.interpolate(torch.randn(1, 2, 34, 60, device=dev, generator=g), size=(1080, 1920), mode="bicubic") * 8) for _ in range(16)]
it = [0]
def c1():
    t, f = pool[it[0] % 16]; it[0] += 1
    d.softsplat(t, f, None, "avg")
cases["C1 avg 1x3x1080x1920 f32 (pool of 16)"] = c1
for r in (64, 32, 16, 8):
    fa = torch.randn(2, 2, r, r, device=dev, generator=g) * 0.3
    cases[f"mask 2x2x{r}x{r}"] = (lambda fa=fa: d.compute_mask(fa, -fa))
f3 = torch.rand(4, 3, 256, 256, device=dev, generator=g); fl3 = torch.randn(4, 2, 256, 256, device=dev, generator=g) * 3
cases["avg 4x3x256x256 f32"] = lambda: d.softsplat(f3, fl3, None, "avg")
f5 = torch.rand(1, 3, 540, 960, device=dev, generator=g); fl5 = torch.randn(1, 2, 540, 960, device=dev, generator=g) * 3
cases["avg 1x3x540x960 f32"] = lambda: d.softsplat(f5, fl5, None, "avg")

for path in (0, 1):
    d._lib.set_option("fwd_path", path)
    d._lib.release_workspaces()
    for name, fn in cases.items():
        with torch.no_grad():
            e = timeit(fn)
            try:
                gq = timeit(graphed(fn)) if "pool" not in name else float("nan")
            except Exception as ex:
                gq = float("nan")
        print(f"path {path}  {name:40s} eager {e:7.1f} us   graph replay {gq:7.1f} us")
