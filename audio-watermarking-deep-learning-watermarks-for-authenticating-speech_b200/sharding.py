"""Multi-GPU plan for the embed+detect path: clips are independent (fresh LSTM state and
message per 1 s segment, py/main16.py:996-1009), so the path shards by contiguous clip
range with NO data-path collective; only file-level aggregates (a handful of floats) are
reduced.  One process per GPU, torch.distributed for the plumbing (NCCL on GPUs, gloo in
the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `n_items` owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_file_stats(sum_prob: float, n_samples: int, sum_msg_logits: torch.Tensor, n_segments: int,
                      group=None):
    """All-reduce the partial sums detect_watermark needs (py/main16.py:1170-1187):
    sum of per-sample probabilities / sample count, sum of per-segment mean message
    logits / segment count.  Works on any backend; tensors stay on `sum_msg_logits.device`."""
    import torch.distributed as dist
    dev = sum_msg_logits.device
    buf = torch.cat([torch.tensor([sum_prob, float(n_samples), float(n_segments)], dtype=torch.float64, device=dev),
                     sum_msg_logits.to(torch.float64)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    mean_prob = (buf[0] / buf[1].clamp(min=1)).item()
    mean_logits = (buf[3:] / buf[2].clamp(min=1)).to(torch.float32)
    return mean_prob, mean_logits
