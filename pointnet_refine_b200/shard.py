"""Data-parallel sharding of independent lane segments across ranks (one process per GPU).

Segments never interact in eval mode (BatchNorm uses running statistics; pooling and attention are
per segment -- SURVEY.md section 8e), so inference shards contiguous segment ranges and needs no
data-path collective; results are gathered only if the caller wants them in one place."""
from __future__ import annotations

import torch
import torch.distributed as dist


def segment_shard(num_segments: int, rank: int, world: int) -> range:
    """Contiguous range of ceil(B / world) segments owned by `rank` (possibly empty for trailing ranks)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    per = -(-num_segments // world)
    lo = min(rank * per, num_segments)
    return range(lo, min(lo + per, num_segments))


def gather_segments(local: torch.Tensor, num_segments: int, group=None) -> torch.Tensor:
    """All-gather per-rank results (first dim = local segments) back into segment order."""
    world = dist.get_world_size(group)
    per = -(-num_segments // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat(parts, dim=0)[:num_segments]
