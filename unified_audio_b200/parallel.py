"""Multi-GPU plumbing of the hot path: clips are independent, so ranks own contiguous batch
shards with replicated weights and NO data-path collective; the single exchange step is one all-gather of the
int64 token tensors.  Works with backend "nccl" (GPU) and "gloo" (CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `n_items` clips owned by `rank` (first `n % world` ranks get one extra)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_tokens(tokens: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather int64 tokens [b_local, ...] from every rank into [n_total, ...] (ragged shards padded)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tokens
    world = dist.get_world_size(group)
    b_max = (n_total + world - 1) // world
    pad = torch.zeros((b_max,) + tuple(tokens.shape[1:]), dtype=tokens.dtype, device=tokens.device)
    pad[: tokens.shape[0]] = tokens
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(out[r][: hi - lo])
    return torch.cat(parts, 0)
