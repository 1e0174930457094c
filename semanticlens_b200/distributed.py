"""Image-sharded data parallelism for the sweep and the embed stage (one process per GPU).

The reference is single-process; this module is the one exchange step the sharded path needs:

* rank r of R owns the contiguous dataset range ``[r*ceil(N/R), min(N, (r+1)*ceil(N/R)))`` of *both* datasets, so
  a sample id is still "position in iteration order" (reference activation_caching.py:410-413) plus the shard offset;
* after the sweep every rank holds, per layer, a sorted ``(C, k)`` top-k of its shard. ONE ``all_gather_into_tensor``
  of a packed byte buffer ``[all layers' bf16 values | all layers' int64 ids]`` (NCCL over NVLink/NVSwitch; gloo in
  the CPU tests) gives every rank all R lists, and the K2 list-merge kernel reduces them to the global top-k — the
  canonical (value desc, id asc) order makes the result identical to a single-process sweep;
* embeddings: each rank embeds its shard; one all-gather of the ``(ceil(N/R), D)`` fp32 shards.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    lo: int
    hi: int
    per: int  # rows per rank (last ranks may own fewer)


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def image_shard(n: int, rank: int | None = None, world_size: int | None = None) -> Shard:
    """Contiguous block partition of ``range(n)``."""
    if rank is None or world_size is None:
        rank, world_size = world()
    per = -(-n // world_size) if world_size > 0 else n
    lo = min(n, rank * per)
    hi = min(n, (rank + 1) * per)
    return Shard(rank, world_size, lo, hi, per)


def pack_states(states: list[tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
    """[(vals (C,k) bf16, ids (C,k) i64), ...] -> one uint8 buffer [all vals | pad to 8 | all ids]."""
    vals = torch.cat([v.reshape(-1) for v, _ in states]) if states else torch.empty(0, dtype=torch.bfloat16)
    ids = torch.cat([i.reshape(-1) for _, i in states]) if states else torch.empty(0, dtype=torch.int64)
    vb = vals.contiguous().view(torch.uint8)
    pad = (-vb.numel()) % 8
    if pad:
        vb = torch.cat([vb, torch.zeros(pad, dtype=torch.uint8, device=vb.device)])
    return torch.cat([vb, ids.contiguous().view(torch.uint8)])


def unpack_states(buf: torch.Tensor, shapes: list[tuple[int, int]]) -> list[tuple[torch.Tensor, torch.Tensor]]:
    """Inverse of :func:`pack_states` for one rank's buffer."""
    n = sum(c * k for c, k in shapes)
    vbytes = n * 2
    voff = vbytes + ((-vbytes) % 8)
    vals = buf[:vbytes].view(torch.bfloat16)
    ids = buf[voff : voff + n * 8].view(torch.int64)
    out, o = [], 0
    for c, k in shapes:
        out.append((vals[o : o + c * k].view(c, k), ids[o : o + c * k].view(c, k)))
        o += c * k
    return out


def exchange_states(
    states: list[tuple[torch.Tensor, torch.Tensor]], group=None
) -> list[tuple[torch.Tensor, torch.Tensor]]:
    """One all-gather: per layer (C,k) states on every rank -> per layer (R, C, k) stacks on every rank."""
    _, R = world()
    shapes = [tuple(v.shape) for v, _ in states]
    mine = pack_states(states)
    flat = torch.empty(R * mine.numel(), dtype=torch.uint8, device=mine.device)  # 1-D: valid for nccl and gloo
    dist.all_gather_into_tensor(flat, mine, group=group)
    gathered = flat.view(R, mine.numel())
    per_rank = [unpack_states(gathered[r], shapes) for r in range(R)]
    out = []
    for li in range(len(states)):
        out.append(
            (torch.stack([per_rank[r][li][0] for r in range(R)]), torch.stack([per_rank[r][li][1] for r in range(R)]))
        )
    return out


def merge_actmax_across_ranks(actmax_cache, device) -> None:
    """Exchange + K2 list merge; afterwards every rank's ``ActMax`` holds the global top-k."""
    from . import ops

    layers = list(actmax_cache.cache.keys())
    for name in layers:
        if not actmax_cache.cache[name].is_setup:
            raise ValueError(
                f"rank {world()[0]} collected nothing for layer '{name}': the dataset must have at least one item "
                "per rank"
            )
    states = [actmax_cache.cache[name].device_tensors(device) for name in layers]
    stacks = exchange_states(states)
    for name, (vals, ids) in zip(layers, stacks):
        mv, mi = ops.topk_merge_lists(vals, ids)
        am = actmax_cache.cache[name]
        am._dev_vals, am._dev_ids = mv, mi
        am._cpu_vals = am._cpu_ids = None
        am.finalize()


def all_gather_rows(local: torch.Tensor, shard: Shard, n_total: int, group=None) -> torch.Tensor:
    """(n_local, D) shards -> (n_total, D) on every rank (block partition: only trailing ranks are short)."""
    R, per = shard.world, shard.per
    D = local.shape[1]
    padded = local
    if local.shape[0] < per:
        padded = torch.zeros((per, D), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty(R * per * D, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous().view(-1), group=group)
    return out.view(R * per, D)[:n_total]
