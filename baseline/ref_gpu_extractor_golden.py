"""Run ON THE GPU BOX: the reference's own Bi_Dir_FeatureExtractor (controlnet/extractors.py:209-316, with its own
control_utils.py and softsplat.py, kernels NVRTC-compiled through baseline/cupy_shim.py) on seeded 512 x 512 inputs and
seeded non-zero parameters; outputs of all four scales and a few gradients -> gpurun_out/ref_gpu_extractor.npz.
Committed as tests/golden/ref_gpu_extractor.npz it pins compute_mask / FeatureWarperSoftsplat / the fusion / the
hole fill -- and this library's drop-in module with the fused block -- to an execution of the reference.
Also prints the forward / forward+backward times of both implementations."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_gpu  # noqa: E402

INJECT = [16, 16, 32, 64]


def make_inputs(seed=5, batch=1, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    cond = torch.rand(batch, 6, 512, 512, generator=g)
    low = torch.randn(batch, 4, 16, 16, generator=g)
    # raw 512-px flows of a few pixels: the extractor feeds them, divided by (res - 1) / 2, to the splat as PIXELS of the res x res
    # grid, i.e. 0.1 .. 1.5 cells here; backward = -forward + a perturbation, so that the
    # forward-backward check (> 0.3) marks part of every scale as occluded and leaves the rest visible
    flow = torch.nn.functional.interpolate(low, size=(512, 512), mode="bicubic", align_corners=False) * 4.0
    flow[:, 2:] = -flow[:, :2] + 1.2 * torch.nn.functional.interpolate(torch.randn(batch, 2, 8, 8, generator=g), size=(512, 512), mode="bicubic")
    return cond.to(device), flow.to(device)


def seeded_state(module, seed=9):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in module.state_dict().items():
        scale = 0.5 / max(1.0, float(v[0].numel()) ** 0.5) if v.dim() > 1 else 0.1
        sd[k] = torch.randn(v.shape, generator=g) * scale
        if k.endswith("metric_net.2.bias"):
            sd[k] = sd[k] * 0 + 0.15          # confidences on both sides of the clamp at 0 (extractors.py:301)
    return sd


def run(module, cond, flow, gouts):
    cond = cond.clone().requires_grad_(True)
    outs = module(cond, flow)
    loss = sum((o * g).sum() for o, g in zip(outs, gouts))
    loss.backward()
    grads = {"cond_sub": cond.grad[:, :, ::8, ::8].contiguous()}
    for name, p in module.named_parameters():
        if "metric_net" in name or name in ("first_pre_extractor.0.weight", "extractors_last.3.0.weight"):
            grads[name] = p.grad.detach().clone()
    return [o.detach() for o in outs], grads


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    ref = ref_gpu.load_package()
    import diffcodec_b200 as d
    theirs = ref.extractors.Bi_Dir_FeatureExtractor(INJECT).cuda()
    ours = d.Bi_Dir_FeatureExtractor(INJECT).cuda()
    sd = seeded_state(theirs)
    theirs.load_state_dict(sd); ours.load_state_dict(sd)
    cond, flow = make_inputs()
    g = torch.Generator().manual_seed(3)
    gouts = [torch.randn(1, c, r, r, generator=g).cuda() for c, r in zip(INJECT, (64, 32, 16, 8))]
    o_ref, g_ref = run(theirs, cond, flow, gouts)
    o_our, g_our = run(ours, cond, flow, gouts)
    blob = {}
    for i, (a, b) in enumerate(zip(o_ref, o_our)):
        blob[f"out{i}"] = a.cpu().numpy()
        print(f"scale {i}: max|ref| {a.abs().max().item():.4g}  max|ours - ref| {(a - b).abs().max().item():.3g}")
    for k in g_ref:
        blob["grad/" + k] = g_ref[k].cpu().numpy()
        print(f"grad {k}: max|ref| {g_ref[k].abs().max().item():.4g}  max|ours - ref| {(g_ref[k] - g_our[k]).abs().max().item():.3g}")
    out = os.path.join(ROOT, "gpurun_out", "ref_gpu_extractor.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")
    # timing at the live consumer's sizes: inject_channels (320, 320, 640, 1280), batch 2
    big_ref = ref.extractors.Bi_Dir_FeatureExtractor([320, 320, 640, 1280]).cuda()
    big_our = d.Bi_Dir_FeatureExtractor([320, 320, 640, 1280]).cuda()
    big_our.load_state_dict(big_ref.state_dict())
    cond2, flow2 = make_inputs(seed=6, batch=2)
    with torch.no_grad():
        t_ref = timeit(lambda: big_ref(cond2, flow2)); t_our = timeit(lambda: big_our(cond2, flow2))
    print(f"Bi_Dir_FeatureExtractor forward, batch 2, inject (320,320,640,1280): reference {t_ref:.0f} us, ours {t_our:.0f} us")
    c2 = cond2.clone().requires_grad_(True)
    def fb(m):
        def f():
            m.zero_grad(set_to_none=True); c2.grad = None
            sum(o.sum() for o in m(c2, flow2)).backward()
        return f
    t_ref = timeit(fb(big_ref), 10, 3); t_our = timeit(fb(big_our), 10, 3)
    print(f"Bi_Dir_FeatureExtractor forward + backward: reference {t_ref:.0f} us, ours {t_our:.0f} us")


if __name__ == "__main__":
    main()
