"""Frame-parallel multi-GPU plumbing: one process per GPU, Gaussians replicated,
frame i -> rank i mod N, no collective on the raster path; the only exchange is an all-reduce of
the scalar loss (NCCL on GPUs, gloo in the CPU tests).  The reference has no multi-GPU code at all
(it pins device 0, utils/general_utils.py:161)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_frames(num_frames: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment used by the render sweep (render.py:64-88 sharded 8 frames/iter)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, num_frames, world_size))


def allreduce_loss(loss: torch.Tensor, group: Optional[dist.ProcessGroup] = None, average: bool = False) -> torch.Tensor:
    """Sum (or mean) of the per-rank scalar losses; identity when not distributed."""
    out = loss.detach().clone().reshape(1)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        if average:
            out /= dist.get_world_size(group)
    return out


def run_sharded(frames: Sequence, step: Callable[[object], torch.Tensor], rank: int, world_size: int,
                group: Optional[dist.ProcessGroup] = None):
    """Run `step(frame) -> scalar loss` on this rank's share, in lock-step iterations of
    `world_size` frames; returns (list of per-iteration global loss sums, local frame indices).
    Ranks whose share is exhausted contribute 0 so the collective stays matched."""
    mine = shard_frames(len(frames), rank, world_size)
    iters = (len(frames) + world_size - 1) // world_size
    sums = []
    for it in range(iters):
        idx = it * world_size + rank
        if idx < len(frames):
            loss = step(frames[idx])
        else:
            loss = torch.zeros((), device=sums[-1].device if sums else "cpu")
        sums.append(allreduce_loss(loss, group))
    return sums, mine
