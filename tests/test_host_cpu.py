"""CPU-only checks: the C-ABI library loads and exports exactly what include/*.h declares, the
host logic (mode parsing, asserts, sharding) behaves, the N>1 gather path works over gloo."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "diffcodec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import diffcodec_b200
    L = diffcodec_b200._lib
    lib = L.lib()                                   # loads without a GPU (static cudart, no compute calls)
    declared = _header_symbols()
    assert declared == sorted(L.SYMBOLS), "ctypes table and header disagree"
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dcb_version() == 100
    assert b"sm_100a" in lib.dcb_build_info()
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (dcb_[a-z0-9_]+)", out)))
    assert exported == declared


def test_workspace_queries_need_no_gpu():
    import diffcodec_b200
    L = diffcodec_b200._lib
    lib = L.lib()
    # C+1 <= 4: ONE L2-sized slot of float4 accumulators (a slot = one 1080p frame); round 1's two-slot ring is an option
    assert lib.dcb_splat_fwd_workspace_bytes(1, 3, 1080, 1920, L.DCB_F32, L.MODE_SOFT, 0) == 1080 * 1920 * 16
    assert lib.dcb_splat_fwd_workspace_bytes(64, 3, 1080, 1920, L.DCB_F32, L.MODE_SOFT, 0) == 1080 * 1920 * 16
    L.set_option("pipe_ring_slots", 2)
    assert lib.dcb_splat_fwd_workspace_bytes(64, 3, 1080, 1920, L.DCB_F32, L.MODE_SOFT, 0) == 2 * 1080 * 1920 * 16
    L.set_option("pipe_ring_slots", 1)
    assert lib.dcb_splat_workspace_bytes(1, 8, 64, 64, L.DCB_F32, L.MODE_SUM, 0) == 2 * 64 * 64 * 16   # two channel quads
    assert lib.dcb_splat_fwd_workspace_is_scratch(1, 3, 1080, 1920, L.DCB_F32, L.MODE_SOFT, 0) == 0
    assert lib.dcb_splat_fwd_workspace_is_scratch(1, 3, 1080, 1920, L.DCB_F32, L.MODE_SOFT, L.FLAG_DETERMINISTIC) == 1
    L.set_option("pipe_group_bytes", 1 << 20)          # tests shrink the ring slots: many groups on small tensors
    try:
        assert lib.dcb_splat_fwd_workspace_bytes(64, 3, 256, 256, L.DCB_F32, L.MODE_SOFT, 0) == 256 * 256 * 16
    finally:
        L.set_option("pipe_group_bytes", 0)
    L.set_option("fwd_path", 2)      # opt-in target-tile owner kernels (C + 1 <= 4): the workspace only holds the landing boxes
    try:                             # of the 32 x 4 source strips (8 B per strip + 8 B per strip row); small frames need none
        boxes = 270 * (60 + 1) * 8
        assert lib.dcb_splat_fwd_workspace_bytes(1, 3, 1080, 1920, L.DCB_F32, L.MODE_SOFT, 0) == (boxes + 255) // 256 * 256
        assert lib.dcb_splat_fwd_workspace_bytes(64, 3, 1080, 1920, L.DCB_F32, L.MODE_SOFT, 0) == 64 * boxes
        assert lib.dcb_splat_fwd_workspace_bytes(4, 3, 135, 240, L.DCB_BF16, L.MODE_SOFT, 0) == 0
        assert lib.dcb_splat_fwd_workspace_is_scratch(1, 3, 1080, 1920, L.DCB_F32, L.MODE_SOFT, 0) == 1
        assert lib.dcb_splat_fwd_workspace_is_scratch(8, 64, 256, 256, L.DCB_F32, L.MODE_SOFT, 0) == 1     # per-target lists: rebuilt by every call
        assert lib.dcb_splat_fwd_workspace_is_scratch(2, 64, 32, 32, L.DCB_F32, L.MODE_SOFT, 0) == 0       # planar accumulators
        assert lib.dcb_occlusion_mask_workspace_bytes(2, 64, 64) == 0
    finally:
        L.set_option("fwd_path", 0)
    with pytest.raises(AssertionError):
        L.set_option("no_such_option", 1)
    assert lib.dcb_splat_workspace_bytes(1, 8, 64, 64, L.DCB_F64, L.MODE_SUM, 0) == 0        # fp64 reds go straight into out
    # many channels: planar accumulators, 2 slots x 2 frames x 64 planes + 3 slots x 2 normaliser planes
    assert lib.dcb_splat_workspace_bytes(8, 64, 256, 256, L.DCB_F32, L.MODE_SOFT, 0) == (2 * 2 * 64 + 3 * 2) * 256 * 256 * 4 + 256    # + the counters of the single-launch kernel
    assert lib.dcb_splat_bwd_workspace_bytes(64, 3, 1080, 1920, L.DCB_F32, L.MODE_SOFT, 0) == 64 * 1080 * 1920 * 16      # one packed float4 per target
    assert lib.dcb_splat_bwd_workspace_bytes(8, 64, 256, 256, L.DCB_F32, L.MODE_SOFT, 0) == 8 * 256 * 256 * 8
    assert lib.dcb_occlusion_mask_workspace_bytes(2, 64, 64) == 2 * 64 * 64 * 16
    assert lib.dcb_residual_workspace_bytes(1, 3, 1080, 1920) == 1080 * 1920 * (24 + 8)      # float4 + float2 cells, two mask planes


def test_cpu_tensors_are_refused_like_the_reference():
    import diffcodec_b200 as d
    tin, flow = torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4)
    with pytest.raises(AssertionError):
        d.softsplat(tin, flow, None, "sum")          # reference: assert(False) on non-CUDA, softsplat.py:347-348
    with pytest.raises(AssertionError):
        d.softsplat(tin, flow, None, "soft")         # metric required, softsplat.py:238
    with pytest.raises(AssertionError):
        d.softsplat(tin, flow, torch.zeros(1, 1, 4, 4), "avg")
    with pytest.raises(AssertionError):
        d.softsplat(tin, flow, None, "median")


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import diffcodec_b200 as d
    L = d._lib
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        L.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "diffcodec-controlling-latent-diffusion-for-perceptual-video-compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_dropin_install():
    import diffcodec_b200 as d
    d.install()
    from controlnet.softsplat import softsplat                      # noqa: the reference's import lines
    from controlnet.control_utils import zero_module, resize_and_normalize_flow_batched, FeatureWarperSoftsplat, compute_mask, FDN  # noqa
    assert softsplat is d.softsplat and compute_mask is d.compute_mask


def test_gop_sharding():
    import diffcodec_b200 as d
    units = d.enumerate_gops()
    assert len(units) == 975 and sum(u.inter_frames for u in units) == 2925
    for world in (1, 2, 4, 8):
        shards = [d.shard_units(units, r, world) for r in range(world)]
        assert sorted(sum(shards, []), key=lambda u: (u.sequence, u.index)) == sorted(units, key=lambda u: (u.sequence, u.index))
        assert max(map(len, shards)) - min(map(len, shards)) <= 1
    assert units[0].seed() == d.enumerate_gops()[0].seed() and units[0].seed() != units[1].seed()
    assert d.enumerate_gops([("x", 9)], gop=8)[0].inter_frames == 7 and len(d.enumerate_gops([("x", 9)], gop=8)) == 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DCB_ROOT"])
import diffcodec_b200 as d
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
units = d.shard_units(d.enumerate_gops([("a", 20), ("b", 12)]), rank, world)
table = torch.stack([d.checksum(torch.full((4, 4), float(u.seed() % 97))) for u in units])
allsums = d.gather_checksums(table)
assert allsums.shape[0] == world and allsums.shape[2] == 3
total = float(allsums[..., 2].sum())
assert total == 16 * 8, total                      # 8 GOPs, 16 elements each, none lost or duplicated
outs = d.gather_outputs(torch.full((2, 3), float(rank)), dst=0)
if rank == 0:
    assert [float(o[0, 0]) for o in outs] == [0.0, 1.0]
else:
    assert outs is None
dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_over_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, DCB_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


def test_flow_ingest(tmp_path):
    """f-2: .flo round trip (interleaved), the dataset's planar mis-reshape, and both resize conventions."""
    import numpy as np
    import diffcodec_b200 as d
    fio = d.flow_io
    rng = np.random.default_rng(3)
    flow = rng.standard_normal((6, 10, 2)).astype(np.float32) * 5
    path = str(tmp_path / "flow_0000_0003.flo")
    fio.write_flo(path, flow)
    assert os.path.getsize(path) == 12 + 6 * 10 * 2 * 4                      # header + payload (512x512 -> 2,097,164 B)
    back = fio.read_flo(path)
    assert back.shape == (6, 10, 2) and np.array_equal(back, flow)
    quirk = fio.read_flo(path, planar_quirk=True)                            # dataset.py:15-24
    assert quirk.shape == (2, 6, 10) and np.array_equal(quirk.ravel(), flow.ravel()) and not np.array_equal(quirk[0], flow[..., 0])
    with open(path, "r+b") as f:
        f.write(b"\x00\x00\x00\x00")
    with pytest.raises(ValueError):
        fio.read_flo(path)
    up = fio.resize_flow_to(flow, 12, 20)                                    # utils.py:21-28: rescaled vectors
    ref = torch.nn.functional.interpolate(torch.from_numpy(flow).permute(2, 0, 1)[None], size=(12, 20), mode="bilinear", align_corners=True)
    assert torch.allclose(up[:, 0], ref[:, 0] * 2.0) and torch.allclose(up[:, 1], ref[:, 1] * 2.0)
    down = fio.fast_downsample_flow(flow.transpose(2, 0, 1), 3, 5)           # dataset.py:43-50: no rescale
    assert down.shape == (2, 3, 5) and np.allclose(down[0, 0, 0], flow[:2, :2, 0].mean(), atol=1e-6)
    np.save(str(tmp_path / "cached.npy"), flow.transpose(2, 0, 1))
    assert np.allclose(fio.load_flow_cached(str(tmp_path / "cached.flo"), 3, 5), down)


def test_host_chunk_schedule():
    """softsplat_host's chunk plan covers every frame once, never exceeds the steady chunk, and
    starts / ends with single frames when the batch is long enough to ramp."""
    from diffcodec_b200.host import _chunk_schedule
    for n in range(1, 80):
        for k in (1, 2, 3, 4, 8, 16):
            plan = _chunk_schedule(n, min(k, n))
            assert sum(plan) == n and min(plan) >= 1 and max(plan) <= max(1, min(k, n)), (n, k, plan)
    assert _chunk_schedule(64, 8)[:4] == [1, 2, 4, 8] and _chunk_schedule(64, 8)[-1] == 1


def test_dropin_extractors_have_the_reference_state_dict():
    """tests/golden/ref_extractors_state_dict.json: parameter names and shapes of the reference's classes
    (controlnet/extractors.py, instantiated with inject_channels [32, 32, 64, 128]): checkpoints must load unchanged."""
    import json
    import diffcodec_b200 as d
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_extractors_state_dict.json")))
    for name, mod in (("Bi_Dir_FeatureExtractor", d.Bi_Dir_FeatureExtractor([32, 32, 64, 128])),
                      ("Bi_Dir_ResidueExtractor", d.Bi_Dir_ResidueExtractor([32, 32, 64, 128])), ("WarpExtractor", d.WarpExtractor())):
        got = {k: list(v.shape) for k, v in mod.state_dict().items()}
        assert got == want[name], name
    d.install()
    from controlnet.extractors import Bi_Dir_FeatureExtractor      # flownet.py:8
    assert Bi_Dir_FeatureExtractor is d.Bi_Dir_FeatureExtractor


def test_flow_io_host_helpers_match_the_reference_functions(tmp_path):
    """tests/golden/ref_flow_io.npz: outputs of the reference's OWN read_flo / resize_flow_to (controlnet/utils.py:10-28),
    load_flo_file / fast_downsample_flow (controlnet/dataset.py:15-50) and resize_and_normalize_flow_batched
    (controlnet/control_utils.py:74-97), executed from their source text by oracle/ref_flow_io.py."""
    import numpy as np
    import diffcodec_b200 as d
    from oracle.ref_flow_io import CASES, case_flow
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_flow_io.npz"))
    fio = d.flow_io
    for i, (h, w, th, tw) in enumerate(CASES):
        path = str(tmp_path / "x.flo")
        fio.write_flo(path, case_flow(h, w, 100 + i))
        assert np.array_equal(fio.read_flo(path), z[f"{i}/read_flo"])
        assert np.array_equal(fio.read_flo(path, planar_quirk=True), z[f"{i}/load_flo_file"])
        assert np.array_equal(fio.resize_flow_to(fio.read_flo(path), th, tw).numpy(), z[f"{i}/resize_flow_to"])
        if f"{i}/fast_downsample_flow" in z.files:
            assert np.array_equal(fio.fast_downsample_flow(fio.read_flo(path, planar_quirk=True), th, tw), z[f"{i}/fast_downsample_flow"])
    for r in (64, 32, 16, 8):
        got = d.resize_and_normalize_flow_batched(torch.from_numpy(z["batched/in"]), r, r).numpy()
        assert np.array_equal(got, z[f"batched/normalize_{r}"])
