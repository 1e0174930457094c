"""Image-sharded data parallelism for the sweep and the embed stage (one process per GPU).

The reference is single-process; this module is the exchange the sharded path needs:

* rank r of R owns a contiguous, balanced range of *both* datasets (the first ``N mod R`` ranks own one more item), so a
  sample id is still "position in iteration order" (reference activation_caching.py:410-413) plus the shard offset and
  no rank is ever empty when ``N >= R``; ``N < R`` is rejected identically on every rank (no collective is entered);
* after the sweep every rank holds, per layer, a sorted ``(C, k)`` top-k of its shard. ONE ``all_gather_into_tensor``
  of a packed byte buffer ``[all layers' bf16 values | all layers' int64 ids]`` (NCCL over NVLink/NVSwitch; gloo in
  the CPU tests) gives every rank all R lists, and the K2 list-merge kernel reduces them to the global top-k — the
  canonical (value desc, id asc) order makes the result identical to a single-process sweep;
* embeddings: each rank embeds its shard. The concept DB only needs the rows the merged top-k refers to
  (reference activation_based.py:385-390 indexes ``embeds[sample_ids]``), so every rank contributes just the winners it
  owns: the winner set is known to all ranks after the merge, which makes the row counts of the second all-gather a
  pure function of shared data (:func:`exchange_winner_rows`). ``all_gather_rows`` (the whole table) is kept for callers
  that want every embedding everywhere;
* decisions that depend on a rank's filesystem view (cache hit or miss) are taken by rank 0 and broadcast
  (:func:`agree`), so no rank can skip a collective the others enter.
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
    per: int  # the largest shard (rows every rank pads to in an all-gather)
    n: int = 0  # total number of items

    def bounds(self, r: int) -> tuple[int, int]:
        """[lo, hi) of rank r under the same partition."""
        return shard_bounds(self.n, r, self.world)


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> tuple[int, int]:
    base, extra = divmod(n, max(world_size, 1))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def image_shard(n: int, rank: int | None = None, world_size: int | None = None) -> Shard:
    """Contiguous balanced partition of ``range(n)``: shard sizes differ by at most one."""
    if rank is None or world_size is None:
        rank, world_size = world()
    lo, hi = shard_bounds(n, rank, world_size)
    per = -(-n // world_size) if world_size > 0 else n
    return Shard(rank, world_size, lo, hi, per, n)


def require_items_per_rank(n: int, what: str = "dataset") -> None:
    """Every rank evaluates the same predicate on the same numbers, so either all raise or none does."""
    _, R = world()
    if R > 1 and n < R:
        raise ValueError(f"the {what} has {n} item(s) but {R} ranks: every rank needs at least one item")


def agree(flag: bool, device=None) -> bool:
    """Rank 0's ``flag`` on every rank (a cache hit seen by rank 0 only must not split the control flow)."""
    _, R = world()
    if R == 1:
        return bool(flag)
    on_gpu = dist.get_backend() == "nccl"
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=(device or "cuda") if on_gpu else "cpu")
    dist.broadcast(t, src=0)
    return bool(int(t.item()))


def barrier() -> None:
    if world()[1] > 1:
        dist.barrier()


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


def share_actmax_from_rank0(actmax_cache) -> None:
    """Rank 0 loaded the act-max cache from disk; give every other rank the same per-layer states (a few MB)."""
    rank, R = world()
    if R == 1:
        return
    payload = [None]
    if rank == 0:
        payload[0] = {name: (am.n_collect, am.n_latents, am.activations.view(torch.int16).numpy(), am.sample_ids.numpy())
                      for name, am in actmax_cache.cache.items() if am.is_setup}
    dist.broadcast_object_list(payload, src=0)
    if rank != 0:
        for name, (k, c, bits, ids) in payload[0].items():
            am = type(actmax_cache.cache[name])(n_collect=k, n_latents=c)
            am.activations = torch.from_numpy(bits.copy()).view(torch.bfloat16)
            am.sample_ids = torch.from_numpy(ids.copy())
            actmax_cache.cache[name] = am


def all_gather_rows(local: torch.Tensor, shard: Shard, n_total: int, group=None) -> torch.Tensor:
    """(n_local, D) shards -> (n_total, D) on every rank."""
    R, per = shard.world, shard.per
    D = local.shape[1]
    padded = local
    if local.shape[0] < per:
        padded = torch.zeros((per, D), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty(R * per * D, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous().view(-1), group=group)
    out = out.view(R, per, D)
    if n_total == R * per:
        return out.view(R * per, D)
    sizes = [hi - lo for lo, hi in (shard_bounds(n_total, r, R) for r in range(R))]
    return torch.cat([out[r, : sizes[r]] for r in range(R)])


def winner_plan(id_tensors: list[torch.Tensor], n_total: int, world_size: int):
    """What the winners-only exchange moves, computed identically on every rank from the merged (global) top-k ids.

    Returns (winners, starts): ``winners`` = sorted unique non-negative dataset indices any layer refers to (the
    placeholder id -1 indexes like python, i.e. the LAST item: reference activation_based.py:389), ``starts[r]`` =
    offset of rank r's first winner in that list (length R + 1).
    """
    flat = torch.cat([t.reshape(-1) for t in id_tensors]) if id_tensors else torch.empty(0, dtype=torch.int64)
    flat = torch.where(flat < 0, flat + n_total, flat)
    winners = torch.unique(flat)  # sorted
    edges = torch.tensor([shard_bounds(n_total, r, world_size)[0] for r in range(world_size)] + [n_total],
                         dtype=torch.int64, device=winners.device)
    starts = torch.searchsorted(winners, edges).tolist()
    return winners, starts


def exchange_winner_rows(local: torch.Tensor, shard: Shard, id_tensors: list[torch.Tensor], group=None):
    """Second (and last) collective of the sharded concept-DB build: only the embeddings the top-k refers to travel.

    local       (n_local, D) embeddings of this rank's shard (row j = dataset item shard.lo + j)
    id_tensors  the merged per-layer (C, k) int64 ids, identical on every rank, on ``local``'s device
    Returns (table, remapped): ``table`` (n_winners, D) identical on every rank and ``remapped[i]`` = ``id_tensors[i]``
    rewritten as row numbers of ``table``, so ``table[remapped[i]] == embeds[id_tensors[i]]`` of the full table.
    """
    R, n = shard.world, shard.n
    winners, starts = winner_plan(id_tensors, n, R)
    counts = [starts[r + 1] - starts[r] for r in range(R)]
    most = max(counts) if counts else 0
    D = local.shape[1]
    mine = winners[starts[shard.rank] : starts[shard.rank + 1]] - shard.lo
    send = torch.zeros((most, D), dtype=local.dtype, device=local.device)
    if mine.numel():
        send[: mine.numel()] = local.index_select(0, mine)
    recv = torch.empty(R * most * D, dtype=local.dtype, device=local.device)
    if most:
        dist.all_gather_into_tensor(recv, send.view(-1), group=group)
    recv = recv.view(R, most, D)
    table = torch.cat([recv[r, : counts[r]] for r in range(R)]) if most else recv.view(0, D)
    remapped = []
    for t in id_tensors:
        w = torch.where(t < 0, t + n, t)
        remapped.append(torch.searchsorted(winners, w.reshape(-1)).view(t.shape))
    return table, remapped
