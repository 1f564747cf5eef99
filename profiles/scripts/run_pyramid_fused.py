"""The four fused bi-directional blocks of one ControlNet forward (live consumer's shapes, batch 2, learned-metric stand-in):
forward only (inference) -- run under `ncu --metrics gpu__time_duration.sum` for the GPU time of every kernel -- and
forward + backward; prints event timings and the host time per call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
g = torch.Generator(device="cuda").manual_seed(9)
pyr = []
for ch, r in ((160, 64), (160, 32), (320, 16), (640, 8)):
    feat = torch.randn(2, ch, r, r, device="cuda", generator=g); f_ = torch.randn(2, 2, r, r, device="cuda", generator=g) * 0.3
    m_ = torch.randn(2, 1, r, r, device="cuda", generator=g) * 0.1
    pyr.append((feat, f_, -f_, m_))
def fwd():
    for feat, ff, fb, m_ in pyr:
        d.bidirectional_block(feat, feat, ff, fb, m_, m_)
with torch.no_grad():
    for _ in range(5): fwd()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): fwd()
    b.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): fwd()
    host = (time.perf_counter() - t0) / 20
    torch.cuda.synchronize()
print(f"pyramid forward, 4 fused blocks: {a.elapsed_time(b) / 20 * 1e3:.1f} us per forward (device-timed), host enqueue time {host * 1e6:.1f} us")

# where the host time goes: launches per block and the time of the C call alone
lib = d._lib.lib()
for feat, ff, fb, m_ in pyr:
    n, c, h, w = feat.shape
    fused = torch.empty_like(feat)
    sizes = [int(lib.dcb_bidir_block_workspace_bytes(n, c, h, w, 0, k)) for k in (0, 1, 2)]
    ws_a = d._lib.workspace(feat.device, sizes[0], "acc"); ws_s = d._lib.workspace(feat.device, sizes[1], "scratch")
    D = d._lib.desc
    args = (D(feat), D(feat), D(ff), D(fb), D(m_), D(m_), D(fused), None, None, None, None, None, None,
            ws_a.data_ptr(), ws_a.numel(), ws_s.data_ptr(), ws_s.numel(), 2, d._lib.stream_ptr(feat.device))
    l0 = d.launch_count()
    lib.dcb_bidir_block_fwd(*args)
    per = d.launch_count() - l0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        lib.dcb_bidir_block_fwd(*args)
    dt = (time.perf_counter() - t0) / 200
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        args2 = (D(feat), D(feat), D(ff), D(fb), D(m_), D(m_), D(fused))
    dd = (time.perf_counter() - t0) / 200
    print(f"{tuple(feat.shape)}: {per} launches per block, C call {dt * 1e6:.1f} us, 7 descriptors {dd * 1e6:.1f} us")
