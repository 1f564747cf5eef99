"""Mint golden vectors for the flow ingest (row f-2) from the reference's OWN readers and resizers.

TEST INFRASTRUCTURE, NOT PRODUCT. Runs only in the build container: ``controlnet/utils.py`` and
``controlnet/dataset.py`` cannot be imported here (PIL / torchvision / albumentations), so the four function
definitions are lifted out of the files with ``ast`` and exactly that source text is executed:

    read_flo, resize_flow_to                (controlnet/utils.py:10-28)
    load_flo_file, fast_downsample_flow     (controlnet/dataset.py:15-24, 43-50)
    resize_and_normalize_flow_batched       (controlnet/control_utils.py:74-97)

Nothing of the reference is copied into the repository; inputs are seeded, outputs go to tests/golden/ref_flow_io.npz.

Usage:  python oracle/ref_flow_io.py
"""
from __future__ import annotations

import ast
import os
import struct
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.environ.get("DCB_REFERENCE_ROOT", "/root/reference")
_GOLD = os.path.join(os.path.dirname(_HERE), "tests", "golden", "ref_flow_io.npz")

CASES = [(48, 80, 32, 32), (67, 121, 17, 23), (32, 32, 64, 65)]     # (H, W, target_h, target_w)


def case_flow(h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(h, w, 2, generator=g) * 5).numpy()


def reference_functions():
    ns = {"torch": torch, "F": F, "np": np, "struct": struct}
    want = {"utils.py": ("read_flo", "resize_flow_to"), "dataset.py": ("load_flo_file", "fast_downsample_flow"),
            "control_utils.py": ("resize_and_normalize_flow_batched",)}
    for fname, names in want.items():
        path = os.path.join(_ROOT, "controlnet", fname)
        for node in ast.parse(open(path).read()).body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns


def write_flo(path, flow_hw2):
    h, w, _ = flow_hw2.shape
    with open(path, "wb") as f:
        f.write(b"PIEH"); f.write(struct.pack("ii", w, h)); f.write(np.ascontiguousarray(flow_hw2, np.float32).tobytes())


def main():
    ref = reference_functions()
    blob = {}
    with tempfile.TemporaryDirectory() as tmp:
        for i, (h, w, th, tw) in enumerate(CASES):
            flow = case_flow(h, w, 100 + i)
            path = os.path.join(tmp, "x.flo")
            write_flo(path, flow)
            hw2 = ref["read_flo"](path)
            assert np.array_equal(hw2, flow)
            blob[f"{i}/read_flo"] = hw2
            blob[f"{i}/resize_flow_to"] = ref["resize_flow_to"](hw2, th, tw).numpy()
            planar = ref["load_flo_file"](path)                          # the (2, h, w) mis-reshape of dataset.py:23
            blob[f"{i}/load_flo_file"] = planar
            if th <= h and tw <= w:
                blob[f"{i}/fast_downsample_flow"] = ref["fast_downsample_flow"](planar, th, tw)
    g = torch.Generator().manual_seed(7)
    batched = torch.randn(2, 2, 72, 72, generator=g) * 9
    blob["batched/in"] = batched.numpy()
    for r in (64, 32, 16, 8):
        blob[f"batched/normalize_{r}"] = ref["resize_and_normalize_flow_batched"](batched, r, r).numpy()
    np.savez_compressed(_GOLD, **blob)
    print("wrote", _GOLD, os.path.getsize(_GOLD), "bytes;", len(blob), "arrays")


if __name__ == "__main__":
    main()
