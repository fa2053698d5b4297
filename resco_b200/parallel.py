"""Multi-GPU plumbing: one process per GPU, instances sharded by contiguous global id.

Instances never interact (each reference env is its own SUMO process, multi_signal.py), so the
simulation needs NO data-path collective.  The per-instance RNG is keyed by the global instance id,
which makes results invariant to the number of ranks.  The only collective is the all-gather of
the local observation tensor when a shared-policy agent (MPLight, agents/mplight.py) evaluates the
whole batch on one rank -- NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """(first global id, count) of the contiguous block owned by `rank`."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allgather_obs(local: torch.Tensor) -> torch.Tensor:
    """[n_local, ...] on every rank -> [sum n_local, ...] on every rank (equal shards)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def broadcast_actions(actions: torch.Tensor, src: int = 0) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(actions, src)
    return actions
