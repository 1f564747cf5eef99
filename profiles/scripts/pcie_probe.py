"""PCIe ceiling of the e2e number, at 1..8 ranks (run under torchrun for N > 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/scripts/pcie_probe.py [--bind]

Every rank copies 1 GiB pinned host <-> its GPU with plain cudaMemcpyAsync (torch copy_): H2D alone, D2H alone, both at
once; ranks start together (barrier) and rank 0 prints the per-rank minimum and the aggregate GB/s. --bind pins each rank
to the NUMA node of its GPU first (diffcodec_b200.bind_to_gpu_numa). Then softsplat_host on 32 frames, fp32 and compact."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
import diffcodec_b200 as d
bind = d.bind_to_gpu_numa() if "--bind" in sys.argv else {"bound": False}
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

nb = 1 << 30
h_in = torch.empty(nb, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nb, dtype=torch.uint8).pin_memory()
h_in.fill_(1); h_out.fill_(0)
d_in = torch.empty(nb, dtype=torch.uint8, device="cuda"); d_out = torch.empty(nb, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def t(fn, n=4):
    fn(); barrier(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    v = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v.item())


def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()

a, b, c = t(h2d), t(d2h), t(both)
if rank == 0:
    print(f"ranks {world} bind {bind}: H2D {nb/a/1e9:.1f} GB/s per rank = {world*nb/a/1e9:.1f} aggregate | D2H {nb/b/1e9:.1f} = {world*nb/b/1e9:.1f} | both {nb/c/1e9:.1f} each way per rank = {2*world*nb/c/1e9:.1f} aggregate")
del h_in, h_out, d_in, d_out
F, H, W = 32, 1080, 1920
x = torch.rand(F, 3, H, W).pin_memory(); fl = (torch.randn(F, 2, H, W) * 4).pin_memory(); m = (-torch.rand(F, 1, H, W)).pin_memory()
out = torch.empty(F, 3, H, W).pin_memory()
dt = t(lambda: d.softsplat_host(x, fl, m, "soft", out=out, chunk_frames=8), 3)
if rank == 0:
    print(f"softsplat_host fp32    : {world*F*H*W/dt/1e9:.2f} Gpx/s whole job ({24*world*F*H*W/dt/1e9:.1f} GB/s up, {12*world*F*H*W/dt/1e9:.1f} GB/s down)")
x8 = (x * 255).to(torch.uint8).pin_memory(); fl16 = fl.half().pin_memory(); m16 = m.half().pin_memory(); o16 = torch.empty(F, 3, H, W, dtype=torch.bfloat16).pin_memory()
dt = t(lambda: d.softsplat_host(x8, fl16, m16, "soft", out=o16, chunk_frames=8), 3)
if rank == 0:
    print(f"softsplat_host compact : {world*F*H*W/dt/1e9:.2f} Gpx/s whole job ({9*world*F*H*W/dt/1e9:.1f} GB/s up, {6*world*F*H*W/dt/1e9:.1f} GB/s down)")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
