"""Frame / GOP sharding across the GPUs of one box.

Every frame -- and every GOP: two intra frames, the inter frames between them and their
forward / backward flows -- is independent: a splat never crosses a frame. So the path shards
with NO data-path collective: one process per GPU, GOP ``g`` of the flattened (sequence, GOP)
list goes to rank ``g mod world``. NCCL (or gloo in the CPU tests) is used only AFTER the sweep,
to gather per-GOP checksums or, on request, the output tensors.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["UVG_SEQUENCES", "GopUnit", "enumerate_gops", "shard_units", "gather_checksums", "gather_outputs", "checksum"]

# UVG 1080p sequences evaluated by the reference (test.sh:6) with the dataset's frame counts.
UVG_SEQUENCES: Tuple[Tuple[str, int], ...] = (
    ("Beauty", 600), ("Bosphorus", 600), ("HoneyBee", 600), ("Jockey", 600),
    ("ReadySteadyGo", 600), ("ShakeNDry", 300), ("YachtRide", 600),
)


@dataclass(frozen=True)
class GopUnit:
    """One GOP: intra frames at `first` and `first + gop`; inter frames first+1 .. first+gop-1."""
    sequence: str
    index: int      # GOP number inside the sequence
    first: int      # frame number of the leading intra frame
    gop: int

    @property
    def inter_frames(self) -> int:
        return self.gop - 1

    def seed(self) -> int:
        # stable across processes (Python's hash() is salted)
        h = 1469598103934665603
        for ch in f"{self.sequence}:{self.index}".encode():
            h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h & 0x7FFFFFFF


def enumerate_gops(sequences: Sequence[Tuple[str, int]] = UVG_SEQUENCES, gop: int = 4) -> List[GopUnit]:
    units = []
    for name, frames in sequences:
        for k in range(frames // gop):
            units.append(GopUnit(name, k, k * gop, gop))
    return units


def shard_units(units: Sequence[GopUnit], rank: int, world: int) -> List[GopUnit]:
    """Round-robin: balances the short sequence and keeps every rank's share within one unit."""
    assert 0 <= rank < world
    return [u for i, u in enumerate(units) if i % world == rank]


def checksum(t: torch.Tensor) -> torch.Tensor:
    """Order-insensitive fp64 digest [sum, sum of squares, count] of a tensor (stays on its device)."""
    x = t.detach().to(torch.float64)
    return torch.stack([x.sum(), (x * x).sum(), torch.tensor(float(x.numel()), dtype=torch.float64, device=t.device)])


def gather_checksums(local: torch.Tensor) -> torch.Tensor:
    """All-gather a per-rank [k,3] checksum table -> [world,k,3] (pads ragged k with zeros)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.unsqueeze(0)
    world = dist.get_world_size()
    k = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    ks = [torch.zeros_like(k) for _ in range(world)]
    dist.all_gather(ks, k)
    kmax = int(max(int(v.item()) for v in ks))
    padded = torch.zeros((kmax, 3), dtype=torch.float64, device=local.device)
    padded[: local.shape[0]] = local
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    return torch.stack(out)


def gather_outputs(local: torch.Tensor, dst: int = 0):
    """Gather equally-shaped per-rank output tensors to `dst` (off the timed path: a 64-frame fp32
    1080p batch is 1.6 GB per rank). Returns the list on `dst`, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    world = dist.get_world_size()
    bufs = [torch.empty_like(local) for _ in range(world)] if dist.get_rank() == dst else None
    dist.gather(local.contiguous(), bufs, dst=dst)
    return bufs
