"""PCIe ceiling for the e2e number: pinned H2D, D2H and both at once; then softsplat_host."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
nb = 1 << 30
h_in = torch.empty(nb, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nb, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nb, dtype=torch.uint8, device="cuda"); d_out = torch.empty(nb, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
print(f"H2D {nb/t(h2d)/1e9:.1f} GB/s   D2H {nb/t(d2h)/1e9:.1f} GB/s   both: {nb/t(both)/1e9:.1f} GB/s each way")
import diffcodec_b200 as d
F, H, W = 64, 1080, 1920
x = torch.rand(F, 3, H, W).pin_memory(); fl = (torch.randn(F, 2, H, W) * 4).pin_memory(); m = torch.randn(F, 1, H, W).pin_memory()
out = torch.empty(F, 3, H, W).pin_memory()
for chunk in (2, 4, 8, 16):
    try:
        dt = t(lambda: d.softsplat_host(x, fl, m, "soft", out=out, chunk_frames=chunk), 3)
        print(f"softsplat_host chunk={chunk}: {dt*1e3:.1f} ms  {F*H*W/dt/1e9:.2f} Gpx/s  H2D {24*F*H*W/dt/1e9:.1f} GB/s")
    except TypeError as e:
        print("signature:", e); break
