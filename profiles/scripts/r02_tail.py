"""Latency of small pipeline calls against the strip height: pipe_tail_percent 0 (32 x 8 strips) vs 100 (32 x 4 single-pass strips)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffcodec_b200 as d
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from r02_small import timeit, graphed, cases          # noqa: E402  (r02_small runs its own table first)

for tail in (0, 100):
    d._lib.set_option("fwd_path", 0)
    d._lib.set_option("pipe_tail_percent", tail)
    d._lib.release_workspaces()
    for name, fn in cases.items():
        if name.startswith("mask"):
            continue
        with torch.no_grad():
            e = timeit(fn)
            try:
                gq = timeit(graphed(fn))
            except Exception as ex:
                gq = float("nan")
        print(f"tail {tail:3d}  {name:40s} eager {e:7.1f} us   graph replay {gq:7.1f} us")
