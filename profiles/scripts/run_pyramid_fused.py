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

# the same four scales as ONE library call (dcb_bidir_pyramid_fwd: two launches + one memset)
levels = [(feat, feat, ff, fb, m_, m_) for feat, ff, fb, m_ in pyr]
def whole():
    d.bidirectional_pyramid(levels)
with torch.no_grad():
    for _ in range(5): whole()
    l0 = d.launch_count(); whole(); per = d.launch_count() - l0
    torch.cuda.synchronize(); a.record()
    for _ in range(50): whole()
    b.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): whole()
    host = (time.perf_counter() - t0) / 50
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        whole()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            whole()
    torch.cuda.synchronize(); a.record()
    for _ in range(50): gr.replay()
    b2 = torch.cuda.Event(enable_timing=True); b2.record(); torch.cuda.synchronize()
print(f"pyramid forward, ONE call: {per} launches, {a.elapsed_time(b2) / 50 * 1e3:.1f} us per CUDA-graph replay; eager: host enqueue time {host * 1e6:.1f} us")
with torch.no_grad():
    torch.cuda.synchronize(); a.record()
    for _ in range(50): whole()
    b.record(); torch.cuda.synchronize()
print(f"pyramid forward, ONE call, eager: {a.elapsed_time(b) / 50 * 1e3:.1f} us per forward (device-timed)")
# with the eight flow resizes in front (one launch) -- what Bi_Dir_FeatureExtractor.forward does between the conv stacks
flow_full = torch.randn(2, 4, 512, 512, device="cuda", generator=g) * 3
def with_resize():
    jobs = []
    for res in (64, 32, 16, 8):
        for fl in (flow_full[:, :2], flow_full[:, 2:]):
            jobs.append((fl, (res, res), False, d._lib.RESAMPLE_DIV, (res - 1) / 2.0, (res - 1) / 2.0))
    fl = d.resample_batch(jobs)
    d.bidirectional_pyramid([(feat, feat, fl[2 * i], fl[2 * i + 1], m_, m_) for i, (feat, _, _, m_) in enumerate(pyr)])
with torch.no_grad():
    for _ in range(5): with_resize()
    torch.cuda.synchronize(); a.record()
    for _ in range(50): with_resize()
    b.record(); torch.cuda.synchronize()
print(f"8 flow resizes + pyramid forward: {a.elapsed_time(b) / 50 * 1e3:.1f} us per forward (device-timed, eager)")

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
