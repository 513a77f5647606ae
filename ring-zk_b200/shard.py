"""Batch sharding across GPUs: every commitment / proof instance is independent
given the shared key (commit.rs:88-128 reads only self, params, x, r), so a batch is split into
contiguous item ranges, one per rank, with no exchange during compute.  The only collective is the
gather of the per-shard ok / verify bitmaps (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int):
    """Contiguous range [lo, hi) of rank `rank`; shard starts are multiples of 8 so that every
    shard's bitmap starts on a byte boundary of the global bitmap."""
    per = -(-B // world)            # ceil
    per = -(-per // 8) * 8          # round up to a multiple of 8
    lo = min(B, rank * per)
    hi = min(B, lo + per)
    return lo, hi


def shard_bytes(B: int, world: int) -> int:
    """bitmap bytes per rank in the gathered buffer (same for every rank)"""
    lo, hi = shard_range(B, 0, world)
    return (hi - lo + 7) // 8


def gather_bitmaps(local_bitmap: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """All-gather the per-shard bitmaps (uint8, LSB-first) into the global bitmap of B bits.
    `local_bitmap` holds the bits of this rank's shard_range; works on any backend."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_bitmap[: (B + 7) // 8].clone()
    nb = shard_bytes(B, world)
    send = torch.zeros(nb, dtype=torch.uint8, device=local_bitmap.device)
    n = min(nb, local_bitmap.numel())
    send[:n] = local_bitmap[:n]
    out = torch.empty(world * nb, dtype=torch.uint8, device=local_bitmap.device)
    dist.all_gather_into_tensor(out, send, group=group)
    return out[: (B + 7) // 8]


def bitmap_to_bool(bitmap, B: int) -> np.ndarray:
    return np.unpackbits(np.asarray(bitmap.cpu() if hasattr(bitmap, "cpu") else bitmap, dtype=np.uint8),
                         bitorder="little")[:B].astype(bool)
