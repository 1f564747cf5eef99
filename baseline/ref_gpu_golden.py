"""Run ON THE GPU BOX: the unmodified reference (baseline/ref_gpu.py) on the inputs of the committed emulation
fixtures (oracle/ref_emulation.py: same CASES, MODES and seeds), forward + all gradients, written to
gpurun_out/ref_gpu_<case>_<dtype>.npz. Copied into tests/golden/ they pin the CPU emulation -- and with it the
oracle and the CUDA path -- to a REAL execution of the reference's kernels (NVRTC-compiled for this GPU).

Also compares, right here, the reference-on-GPU with this library on the same inputs and prints the differences.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_gpu  # noqa: E402
from oracle.ref_emulation import CASES, MODES, make_case  # noqa: E402  (pure helpers; nothing of /root/reference is touched)


def run(fn, apply_fn, tin, flow, metric, gout, mode):
    ti = tin.clone().cuda().requires_grad_(True); fl = flow.clone().cuda().requires_grad_(True)
    me = metric.clone().cuda().requires_grad_(True) if mode.split("-")[0] in ("linear", "soft") else None
    out = fn(tenIn=ti, tenFlow=fl, tenMetric=me, strMode=mode)
    out.backward(gout[:, : out.shape[1]].cuda())
    torch.cuda.synchronize()
    res = {"out": out.detach().cpu().numpy(), "gin": ti.grad.cpu().numpy(), "gflow": fl.grad.cpu().numpy()}
    if me is not None:
        res["gmetric"] = me.grad.cpu().numpy()
    return res


def main():
    ref = ref_gpu.load()
    import diffcodec_b200 as d
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    worst = 0.0
    for name, seed, n, c, h, w, scale, special in CASES:
        for dtype, dname in ((torch.float32, "f32"), (torch.float64, "f64")):
            tin, flow, metric, gout = make_case(seed, n, c, h, w, scale, dtype, special)
            if name == "collide":
                ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
                flow[0, 0] = (3.25 - xs).to(dtype); flow[0, 1] = (4.5 - ys).to(dtype)
            blob = {"tin": tin.numpy(), "flow": flow.numpy(), "metric": metric.numpy(), "gout": gout.numpy()}
            for mode in MODES:
                r = run(ref.softsplat, None, tin, flow, metric, gout, mode)
                o = run(d.softsplat, None, tin, flow, metric, gout, mode)
                for k, v in r.items():
                    blob[f"{mode}/{k}"] = v
                    den = max(float(np.nanmax(np.abs(v))), 1e-30)
                    diff = float(np.nanmax(np.abs(v.astype(np.float64) - o[k].astype(np.float64)))) / den
                    worst = max(worst, diff if dname == "f32" else 0.0)
                    if diff > (1e-5 if dname == "f32" else 1e-12):
                        print(f"  note: {name} {dname} {mode}/{k}: ours vs reference-on-GPU differ by {diff:.3g} of max|ref|")
            ti = tin.clone().cuda().requires_grad_(True); fl = flow.clone().cuda().requires_grad_(True)
            o = ref.softsplat_func.apply(ti, fl); o.backward(gout.cuda()); torch.cuda.synchronize()
            blob["func/out"], blob["func/gin"], blob["func/gflow"] = o.detach().cpu().numpy(), ti.grad.cpu().numpy(), fl.grad.cpu().numpy()
            path = os.path.join(out_dir, f"ref_gpu_{name}_{dname}.npz")
            np.savez_compressed(path, **blob)
            print("wrote", path)
    print(f"worst fp32 difference ours vs reference-on-GPU (max-normalised): {worst:.3g}; device {torch.cuda.get_device_name()}")


if __name__ == "__main__":
    main()
