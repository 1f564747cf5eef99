"""C4 shape (8x64x256x256 fp32 soft) forward: NCHW input vs channels_last input (k_list_gather_nhwc), CUDA-event timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
g = torch.Generator(device="cuda").manual_seed(3)
warm = torch.randn(4096, 4096, device="cuda")
for _ in range(200): warm = warm @ warm * 1e-3          # clocks up before the first timed loop
torch.cuda.synchronize()
for dt in (torch.float32, torch.bfloat16):
    ti = torch.randn(8, 64, 256, 256, device="cuda", generator=g).to(dt)
    me = (torch.randn(8, 1, 256, 256, device="cuda", generator=g) * 0.5).to(dt)
    low = torch.randn(8, 2, 8, 8, device="cuda", generator=g)
    fl = (torch.nn.functional.interpolate(low, size=(256, 256), mode="bicubic") * 4).to(dt)
    cl = ti.to(memory_format=torch.channels_last)
    for name, x in (("NCHW", ti), ("channels_last", cl)):
        for opt in ((1,) if name == "NCHW" else (1, 8, 0)):
            d._lib.set_option("lists_nhwc", opt)
            with torch.no_grad():
                for _ in range(5): d.softsplat(x, fl, me, "soft")
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                for _ in range(20): out = d.softsplat(x, fl, me, "soft")
                b.record(); torch.cuda.synchronize()
            print(f"{dt} {name:14s} lists_nhwc={opt}: {a.elapsed_time(b) / 20 * 1e3:7.1f} us per forward")
    d._lib.set_option("lists_nhwc", 1)
