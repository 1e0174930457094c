"""The image-sharded (N > 1) host path on CPU: world_size-2 `gloo` processes exercise the shard arithmetic, the packed
per-rank top-k exchange (ONE all_gather_into_tensor) and the embedding-shard all-gather of semanticlens_b200.distributed.
The K2 list-merge kernel itself needs a GPU (tests/test_collect_gpu.py); here the gathered (R, C, k) stacks are merged by
the oracle and must equal a single-process sweep over the whole dataset."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import collect as oc
from semanticlens_b200 import distributed as sdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bits(t):
    return t.view(torch.int16).numpy().view(np.uint16)


def _maps(n, C, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, C, 3, 3)).astype(np.float32)


LAYERS = [("a", 5, 4), ("b", 3, 7)]  # (name, C, k): two layers with different shapes, packed into one buffer
N_IMAGES = 23  # not divisible by the world size: the last rank owns a short shard


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = sdist.image_shard(N_IMAGES)
        assert (shard.rank, shard.world) == (rank, world)
        states = []
        for li, (name, C, k) in enumerate(LAYERS):
            maps = _maps(N_IMAGES, C, 100 + li)[shard.lo : shard.hi]
            st = oc.sweep([maps[i : i + 4] for i in range(0, len(maps), 4)], "mean", "conv", k, id_base=shard.lo)
            vals = torch.from_numpy(st.bits.view(np.int16).copy()).view(torch.bfloat16)
            states.append((vals, torch.from_numpy(st.ids.copy())))
        stacks = sdist.exchange_states(states)
        merged = []
        for (vals, ids), (_, C, k) in zip(stacks, LAYERS):
            assert vals.shape == (world, C, k) and ids.shape == (world, C, k)
            mb, mi = oc.merge_lists(_bits(vals), ids.numpy())
            merged.append((mb, mi))
        # embedding shards: row i of the gathered table must be image i's embedding on every rank
        emb_local = torch.arange(shard.lo, shard.hi, dtype=torch.float32).unsqueeze(1).repeat(1, 6)
        table = sdist.all_gather_rows(emb_local, shard, N_IMAGES)
        assert table.shape == (N_IMAGES, 6) and torch.equal(table[:, 0], torch.arange(N_IMAGES, dtype=torch.float32))
        # winners-only exchange: only the rows the merged top-k refers to travel; indexing the compact table with the
        # remapped ids must equal indexing the full table with the original ids (id -1 = last item, like python)
        id_tensors = [torch.from_numpy(mi.copy()) for _, mi in merged]
        id_tensors[0][0, -1] = -1
        compact, remapped = sdist.exchange_winner_rows(emb_local, shard, id_tensors)
        winners, starts = sdist.winner_plan(id_tensors, N_IMAGES, world)
        assert compact.shape[0] == winners.numel() < N_IMAGES and starts[0] == 0 and starts[-1] == winners.numel()
        for ids, rm in zip(id_tensors, remapped):
            assert torch.equal(compact[rm], table[ids])
        # rank 0's filesystem view decides for everyone
        assert sdist.agree(rank == 0) is True and sdist.agree(rank != 0) is False
        # a cache only rank 0 could read reaches the other ranks
        from semanticlens_b200.component_visualization import aggregators as A
        from semanticlens_b200.component_visualization.activation_caching import ActMax, ActMaxCache

        cache = ActMaxCache([n for n, _, _ in LAYERS], A.aggregate_conv_mean, 4)
        if rank == 0:
            for (name, C, k), (mb, mi) in zip(LAYERS, merged):
                am = ActMax(n_collect=k, n_latents=C)
                am.activations = torch.from_numpy(mb.view(np.int16).copy()).view(torch.bfloat16)
                am.sample_ids = torch.from_numpy(mi.copy())
                cache.cache[name] = am
        sdist.share_actmax_from_rank0(cache)
        for (name, C, k), (mb, mi) in zip(LAYERS, merged):
            assert (_bits(cache.cache[name].activations) == mb).all() and (cache.cache[name].sample_ids.numpy() == mi).all()
        out[rank] = merged
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_exchange_equals_single_process_sweep():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    for li, (name, C, k) in enumerate(LAYERS):
        maps = _maps(N_IMAGES, C, 100 + li)
        ref = oc.sweep([maps[i : i + 5] for i in range(0, N_IMAGES, 5)], "mean", "conv", k)  # other batching on purpose
        for rank in range(world):
            mb, mi = out[rank][li]
            assert (mb == ref.bits).all(), f"layer {name}, rank {rank}: merged values differ from the single-process sweep"
            assert (mi == ref.ids).all(), f"layer {name}, rank {rank}: merged ids differ"


def test_image_shard_partition():
    for n in (0, 1, 7, 8, 1000, 1281167):
        for world in (1, 2, 3, 4, 8):
            shards = [sdist.image_shard(n, r, world) for r in range(world)]
            assert shards[0].lo == 0 and shards[-1].hi == n
            for a, b in zip(shards, shards[1:]):
                assert a.hi == b.lo and a.lo <= a.hi
            assert all(s.hi - s.lo <= s.per for s in shards)
            sizes = [s.hi - s.lo for s in shards]
            assert max(sizes) - min(sizes) <= 1  # balanced: no rank is empty unless n < world
            assert all(s.bounds(r) == (shards[r].lo, shards[r].hi) for s in shards for r in range(world))


def test_too_few_items_is_the_same_error_on_every_rank(monkeypatch):
    monkeypatch.setattr(sdist, "world", lambda: (3, 4))
    with pytest.raises(ValueError, match="at least one item"):
        sdist.require_items_per_rank(3)
    sdist.require_items_per_rank(4)
    monkeypatch.setattr(sdist, "world", lambda: (0, 1))
    sdist.require_items_per_rank(0)


def test_pack_unpack_roundtrip():
    g = torch.Generator().manual_seed(0)
    states = [(torch.randn(c, k, generator=g).bfloat16(), torch.randint(-1, 10**9, (c, k), generator=g)) for c, k in
              ((5, 3), (2, 7), (4, 0), (1, 1))]
    buf = sdist.pack_states(states)
    assert buf.dtype == torch.uint8
    back = sdist.unpack_states(buf, [tuple(v.shape) for v, _ in states])
    for (v, i), (v2, i2) in zip(states, back):
        assert torch.equal(v.view(torch.int16), v2.view(torch.int16)) and torch.equal(i, i2)


def test_prefetcher_is_a_passthrough_without_cuda():
    from semanticlens_b200.component_visualization.activation_based import _DevicePrefetcher

    batches = [(torch.full((2, 3, 4, 4), float(i)), torch.zeros(2)) for i in range(5)]
    got = list(_DevicePrefetcher(batches, "cpu"))
    assert len(got) == 5 and all(torch.equal(a[0], b[0]) for a, b in zip(got, batches))
    assert list(_DevicePrefetcher([], "cpu")) == []
