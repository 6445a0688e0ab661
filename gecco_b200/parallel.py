"""Multi-GPU plumbing for sampling: one process per GPU, the batch of independent clouds is split on dim 0 into
contiguous slices (like the reference's JAX twin shards its batch, gecco-jax types.py:53-60 / training.py:61-63),
every rank samples its slice with no data-path collective, and one final all-gather assembles the result
(SURVEY.md §8e).  Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of this rank's clouds; the batch must divide evenly (same assertion as the reference twin)."""
    if total % world_size != 0:
        raise ValueError(f"batch of {total} clouds does not divide over {world_size} ranks")
    per = total // world_size
    return rank * per, (rank + 1) * per


def shard_context(context, start: int, stop: int):
    """Slice of a Context3d (or None) for clouds [start, stop)."""
    if context is None:
        return None
    return context.apply_to_tensors(lambda t: t[start:stop])


def gather_clouds(local: Tensor) -> Tensor:
    """Concatenates the per-rank [B/g, N, 3] results on dim 0 on every rank (the only collective of sampling)."""
    rank, ws = world()
    if ws == 1:
        return local
    out = torch.empty((ws * local.shape[0], *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Max of a per-rank scalar (device-side elapsed time) over all ranks."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sample_sharded(sample_fn: Callable[[tuple, object, int], Tensor], shape, context, seed: int = 42) -> Tensor:
    """Runs `sample_fn(local_shape, local_context, local_seed)` on this rank's slice of the batch and gathers.
    `sample_fn` is typically `lambda s, c, sd: model.sample_stochastic(s, c, rng=torch.Generator(dev).manual_seed(sd))`;
    rank r draws from seed + r so that the clouds of different ranks are independent."""
    rank, ws = world()
    start, stop = shard_range(shape[0], rank, ws)
    local = sample_fn((stop - start, *shape[1:]), shard_context(context, start, stop), seed + rank)
    return gather_clouds(local)
