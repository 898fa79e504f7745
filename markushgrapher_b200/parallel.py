"""Multi-GPU host logic: the path shards by image (independent units, full weight replica per GPU).

One process per GPU (torchrun). The only exchange is an all-gather of decoded token ids so that every rank ends
up with the ids of the whole batch (north star: NCCL all-gather of token ids). Works with any
torch.distributed backend, so the CPU test suite covers it with gloo / world_size 2.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous split of n_items over `world` ranks; the first (n_items % world) ranks take one extra item"""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rows(n_items: int, world: int) -> int:
    """rows every rank brings to the library's sharded generate (mg_generate_dist needs equal shards): the largest shard"""
    return (n_items + world - 1) // world


def pad_shard(local: dict, rows: int) -> dict:
    """pads this rank's shard to `rows` images by repeating its last image (the padding rows are decoded and thrown
    away by unpad_gathered); a rank whose shard is empty cannot be padded from its own rows"""
    n = local["input_ids"].shape[0]
    if n == rows:
        return local
    if n == 0:
        raise ValueError("empty shard: fewer images than ranks (give every rank at least one image)")
    out = {}
    for k, v in local.items():
        out[k] = None if v is None else torch.cat([v, v[-1:].expand(rows - n, *v.shape[1:])], dim=0).contiguous()
    return out


def unpad_gathered(all_ids: torch.Tensor, n_items: int, world: int) -> torch.Tensor:
    """(world * rows, T) ids gathered from equal padded shards -> (n_items, T) in global image order"""
    rows = shard_rows(n_items, world)
    keep = []
    for r in range(world):
        lo, hi = shard_range(n_items, world, r)
        keep.append(all_ids[r * rows: r * rows + (hi - lo)])
    return torch.cat(keep, dim=0)


def gather_token_ids(local_ids: torch.Tensor, n_total: int, pad_id: int = 0, group=None) -> torch.Tensor:
    """local_ids (b_local, T_local) int64 -> (n_total, T_max) on every rank, rows in global image order.
    Rows are right-padded with pad_id to the longest rank's width (ranks may stop at different steps)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local_ids.device
    shape = torch.tensor([local_ids.shape[0], local_ids.shape[1]], device=dev, dtype=torch.int64)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    b_max = int(max(s[0] for s in shapes))
    t_max = int(max(s[1] for s in shapes))
    buf = torch.full((b_max, t_max), pad_id, device=dev, dtype=torch.int64)
    buf[: local_ids.shape[0], : local_ids.shape[1]] = local_ids
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    rows = [bufs[r][: int(shapes[r][0])] for r in range(world)]
    out = torch.cat(rows, dim=0)
    assert out.shape[0] == n_total, (out.shape, n_total)
    _ = rank
    return out


def generate_sharded(engine_generate, batch: dict, max_length: int, pad_id: int = 0, group=None) -> torch.Tensor:
    """Run `engine_generate(**local_batch, max_length=...)` on this rank's contiguous shard of the batch and
    all-gather the ids. `batch` holds the FULL batch on every rank (or at least this rank's rows)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = batch["input_ids"].shape[0]
    lo, hi = shard_range(n, world, rank)
    local = {k: (v[lo:hi] if v is not None else None) for k, v in batch.items()}
    ids = engine_generate(**local, max_length=max_length)
    return gather_token_ids(ids, n, pad_id, group)
