"""Is the step pipeline host-launch-bound? Times the same 64-frame call (a) eagerly, (b) replayed from a CUDA graph."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device="cuda").manual_seed(0)
tin = torch.rand(frames, 3, 1080, 1920, device="cuda", generator=g)
low = torch.randn(frames, 2, 34, 60, device="cuda", generator=g)
flow = torch.nn.functional.interpolate(low, size=(1080, 1920), mode="bicubic") * 8
metric = -torch.rand(frames, 1, 1080, 1920, device="cuda", generator=g)

def run():
    return d.softsplat(tin, flow, metric, "soft")

for _ in range(3):
    run()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    run()
t_host = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    run()
b.record(); torch.cuda.synchronize()
print(f"eager: host issue time {t_host*1e6:.0f} us/call, gpu {a.elapsed_time(b)/5*1e3:.0f} us/call")

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    run()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        out = run()
torch.cuda.synchronize()
for _ in range(3):
    gr.replay()
torch.cuda.synchronize()
a.record()
for _ in range(5):
    gr.replay()
b.record(); torch.cuda.synchronize()
print(f"graph replay: gpu {a.elapsed_time(b)/5*1e3:.0f} us/call -> {frames*1080*1920/(a.elapsed_time(b)/5)/1e3:.0f} Mpx/s")
