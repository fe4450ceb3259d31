"""Data-parallel plumbing for the policy path: episodes shard across ranks (weights replicated), the only exchange is
the all-gather of action logits (BASELINE.json north_star).  One process per GPU, `torch.distributed`
(NCCL over NVLink on the GPU box; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `n_total` episodes owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_logits(local_logits: torch.Tensor, out: Optional[torch.Tensor] = None, group=None) -> torch.Tensor:
    """[B_local, 700] fp32 per rank -> [world * B_local, 700] in rank order (equal B_local on every rank)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_logits
    world = dist.get_world_size(group)
    x = local_logits.contiguous()
    if out is None:
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    return out
