"""Where do the host microseconds of a latent-sized softsplat() call go? (C2 shape, bf16)"""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
g = torch.Generator(device="cuda").manual_seed(1)
tin = (torch.randn(4, 4, 135, 240, device="cuda", generator=g) * 0.18215).bfloat16()
me = (-torch.randn(4, 1, 135, 240, device="cuda", generator=g).abs()).bfloat16()
fl = torch.randn(4, 2, 135, 240, device="cuda", generator=g).bfloat16()
for _ in range(200): d.softsplat(tin, fl, me, "soft")
torch.cuda.synchronize()
n = 5000
t0 = time.perf_counter()
for _ in range(n): d.softsplat(tin, fl, me, "soft")
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue time {1e6*(t1-t0)/n:.1f} us/call, with drain {1e6*(t2-t0)/n:.1f} us/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(n): d.softsplat(tin, fl, me, "soft")
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
