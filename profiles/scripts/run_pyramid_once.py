"""One whole-pyramid call (live consumer's shapes, batch 2) a few times: run under `ncu --metrics gpu__time_duration.sum`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
g = torch.Generator(device="cuda").manual_seed(9)
levels = []
for ch, r in ((160, 64), (160, 32), (320, 16), (640, 8)):
    feat = torch.randn(2, ch, r, r, device="cuda", generator=g); f_ = torch.randn(2, 2, r, r, device="cuda", generator=g) * 0.3
    m_ = torch.randn(2, 1, r, r, device="cuda", generator=g) * 0.1
    levels.append((feat, feat, f_, -f_, m_, m_))
with torch.no_grad():
    for _ in range(6):
        d.bidirectional_pyramid(levels)
torch.cuda.synchronize()
