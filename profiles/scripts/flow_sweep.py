"""Forward throughput vs flow roughness: same amplitude (~8 px), coarser noise grid = smoother flow."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
F = 32
# argv[1]: forward kernel family (0 automatic, 1 round-1 accumulator pipeline, 2 target-tile owner); argv[2]: owner_group_bytes
if len(sys.argv) > 1:
    d._lib.set_option("fwd_path", int(sys.argv[1]))
if len(sys.argv) > 2:
    d._lib.set_option("owner_group_bytes", int(sys.argv[2]))
print("fwd_path", sys.argv[1] if len(sys.argv) > 1 else 0, "owner_group_bytes", sys.argv[2] if len(sys.argv) > 2 else "default")
g = torch.Generator(device="cuda").manual_seed(0)
tin = torch.rand(F, 3, 1080, 1920, device="cuda", generator=g)
metric = -torch.rand(F, 1, 1080, 1920, device="cuda", generator=g)
for cell in (32, 64, 128, 256, 0):
    if cell:
        low = torch.randn(F, 2, max(1080 // cell, 2), max(1920 // cell, 2), device="cuda", generator=g)
        flow = torch.nn.functional.interpolate(low, size=(1080, 1920), mode="bicubic") * 8
    else:
        flow = torch.zeros(F, 2, 1080, 1920, device="cuda") + torch.tensor([3.3, -2.6], device="cuda").view(1, 2, 1, 1)
    gx = (flow[:, :, :, 1:] - flow[:, :, :, :-1]).abs().mean().item()
    for _ in range(3):
        d.softsplat(tin, flow, metric, "soft")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5):
        d.softsplat(tin, flow, metric, "soft")
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"noise cell {cell:4d} px  mean|dflow/dx| {gx:.3f} px/px : {ms*1e3/F:6.1f} us/frame  {F*1080*1920/ms/1e3:8.0f} Mpx/s  {36*F*1080*1920/ms/1e6/6552.6*100:5.1f} % of HBM peak")
