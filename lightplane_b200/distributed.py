"""Ray-parallel multi-GPU support (new work: the reference has no distributed code at all).

Rays are independent, so the ray batch is split across ranks; every rank holds a full replica of
the feature grid and of the MLP parameters.  The only exchange steps are
  * Renderer backward: SUM all-reduce of grad(feature grid), grad(mlp_params) (one flat bucket);
  * Splatter forward: SUM all-reduce of the UN-normalised feature grid and of the weight grid
    *before* `feature / clamp(weight, 1e-5)` (normalising per rank first would be wrong) -- done
    inside `LightplaneSplatterFunction` when a `process_group` is passed;
  * Splatter backward: none.
Backend: `torch.distributed` (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).
"""

from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def is_distributed(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def shard_bounds(num_items: int, rank: int, world_size: int, multiple_of: int = 1) -> Tuple[int, int]:
    """Contiguous `[start, end)` slice of `num_items` for `rank`; all but the last shard are
    multiples of `multiple_of` long and sizes differ by at most one block."""
    blocks = (num_items + multiple_of - 1) // multiple_of
    base, rem = divmod(blocks, world_size)
    start_b = rank * base + min(rank, rem)
    end_b = start_b + base + (1 if rank < rem else 0)
    return min(start_b * multiple_of, num_items), min(end_b * multiple_of, num_items)


def shard_rays(rays, rank: Optional[int] = None, world_size: Optional[int] = None, group=None):
    """This rank's contiguous shard of a `Rays` batch (every field sliced)."""
    if rank is None:
        rank = dist.get_rank(group) if is_distributed(group) else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if is_distributed(group) else 1
    lo, hi = shard_bounds(int(rays.directions.shape[0]), rank, world_size, multiple_of=32)
    return rays[lo:hi]


def all_reduce_sum_(tensors: Sequence[torch.Tensor], group=None) -> None:
    """In-place SUM all-reduce of several tensors through ONE flat bucket (one collective: the
    messages here are small -- 0.8-6 MB triplanes + 17 KB of MLP gradients -- so launch latency,
    not bandwidth, is what matters)."""
    tensors = [t for t in tensors if t is not None]
    if not tensors or not is_distributed(group):
        return
    if len(tensors) == 1 and tensors[0].is_contiguous():
        dist.all_reduce(tensors[0], op=dist.ReduceOp.SUM, group=group)
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    pos = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[pos : pos + n].view_as(t))
        pos += n


def all_reduce_gradients(params: Iterable[torch.Tensor], group=None) -> None:
    """SUM all-reduce of `.grad` of the given leaves (feature grids, `mlp_params`, ...)."""
    all_reduce_sum_([p.grad for p in params if p is not None and p.grad is not None], group)


def broadcast_(tensors: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Broadcast grid / parameter replicas from `src` (once, before training)."""
    if not is_distributed(group):
        return
    for t in tensors:
        dist.broadcast(t, src=src, group=group)
