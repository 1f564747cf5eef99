"""Profiling driver: backward warp + residual on 16 x 3 x 1080 x 1920 fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
F = 16
g = torch.Generator(device="cuda").manual_seed(2)
img = torch.rand(F, 3, 1080, 1920, device="cuda", generator=g); gt = torch.rand(F, 3, 1080, 1920, device="cuda", generator=g)
low = torch.randn(F, 2, 34, 60, device="cuda", generator=g)
flow = torch.nn.functional.interpolate(low, size=(1080, 1920), mode="bicubic") * 8
for i in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); w, r = d.backwarp_residual(img, flow, gt); b.record(); torch.cuda.synchronize()
    print(f"iter {i}: {a.elapsed_time(b)*1e3:.1f} us  {56*F*1080*1920/a.elapsed_time(b)/1e6:.0f} GB/s")
ib, gb = img.bfloat16(), gt.bfloat16()
for fl, tag in ((flow, "bf16 image, fp32 flow (32 B/px)"), (flow.bfloat16(), "bf16 image, bf16 flow (28 B/px)")):
    byts = (2 * 3 * 4 + fl.element_size() * 2) * F * 1080 * 1920
    for i in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); w, r = d.backwarp_residual(ib, fl, gb); b.record(); torch.cuda.synchronize()
    print(f"{tag}: {a.elapsed_time(b)*1e3:.1f} us  {byts/a.elapsed_time(b)/1e6:.0f} GB/s")
