"""Data-parallel tagging over the GPUs of one node: clips are independent, so rank r tags its own
contiguous shard with replicated weights and the only collective is one all_gather of the fp32 logits.
One process per GPU (torchrun); backend NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def tag_sharded(tag_fn: Callable[[torch.Tensor], torch.Tensor], audio: torch.Tensor, group=None) -> torch.Tensor:
    """Every rank passes the SAME [N, n_samples] batch (or a view of it); rank r runs `tag_fn` (e.g.
    `lambda a: model.tag_batch(a, at_time_res)`) on its shard and all ranks return the full [N, S, C] logits in clip
    order.  Uneven shards are padded to the largest shard for the gather and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        return tag_fn(audio)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = audio.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    max_len = shard_bounds(n, 0, world)[1]
    local = tag_fn(audio[lo:hi]) if hi > lo else None
    probe = local if local is not None else tag_fn(audio[:1])[:0]
    S, Cn = probe.shape[1], probe.shape[2]
    buf = torch.zeros((max_len, S, Cn), dtype=probe.dtype, device=probe.device)
    if local is not None:
        buf[:hi - lo] = local
    out = torch.empty((world * max_len, S, Cn), dtype=probe.dtype, device=probe.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n, r, world)
        parts.append(out[r * max_len:r * max_len + (b - a)])
    return torch.cat(parts, dim=0)
