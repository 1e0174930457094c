"""The collect oracle (oracle/collect.py) against the reference's own outputs (tests/golden, made by
oracle/make_golden.py from the imported reference) and the reference's known-answer test."""

import numpy as np
import pytest

from oracle import collect as oc

AGGS = {
    "aggregate_conv_mean": ("mean", "conv"),
    "aggregate_conv_max": ("max", "conv"),
    "aggregate_transformer_mean": ("mean", "tokens"),
    "aggregate_transformer_absmean": ("absmean", "tokens"),
    "aggregate_transformer_max": ("max", "tokens"),
    "aggregate_transformer_absmax": ("absmax", "tokens"),
    "aggregate_transformer_special_token": ("token", "tokens"),
}


def load_maps(z):
    return [z[f"map{i}"] for i in range(int(z["n_batches"]))]


def test_reference_kat(golden):
    """tests/component_visualization/test_activation_caching.py:14-30 of the reference."""
    z = np.load(golden / "actmax_kat.npz")
    st = oc.ActMaxOracle(5, 3)
    st.update(z["acts1"], [0, 1])
    st.update(z["acts2"], [2, 3])
    assert (st.bits == z["ref_bits"]).all()
    assert (st.ids == z["ref_ids"]).all()
    assert st.ids[0].tolist() == [2, 3, 1, 0, -1]
    np.testing.assert_allclose(st.activations[0], np.array([0.9, 0.8, 0.2, 0.1, 0.0]), atol=4e-3)


@pytest.mark.parametrize("name", sorted(AGGS))
def test_aggregates_match_reference(golden, name):
    op, kind = AGGS[name]
    z = np.load(golden / f"collect_{name}.npz")
    for i, m in enumerate(load_maps(z)):
        ref = z[f"agg{i}"]
        exact = oc.aggregate_exact(m, op, kind)
        canon = oc.aggregate_canonical(m, op, kind)
        tol = 0 if op in ("max", "absmax", "token") else 4 * np.finfo(np.float32).eps * np.abs(m).mean() * 4
        np.testing.assert_allclose(exact, ref, rtol=0, atol=tol)
        np.testing.assert_allclose(canon, ref, rtol=0, atol=tol)


@pytest.mark.parametrize("name", sorted(AGGS))
def test_sweep_tie_aware_vs_reference(golden, name):
    op, kind = AGGS[name]
    z = np.load(golden / f"collect_{name}.npz")
    maps = load_maps(z)
    # feed the reference's own fp32 aggregates so that only the top-k semantics are under test
    st = oc.ActMaxOracle(int(z["k"]))
    n = 0
    cand = []
    for i in range(len(maps)):
        a = z[f"agg{i}"]
        st.update(a, np.arange(n, n + len(a)))
        cand.append(oc.f32_to_bf16_bits(a))
        n += len(a)
    cand = np.concatenate(cand).T  # (C, N)
    errs = oc.check_tie_aware(st.bits, st.ids, z["ref_bits"], z["ref_ids"], cand)
    assert not errs, errs


def test_sweep_tiefree_ids_exact(golden):
    z = np.load(golden / "collect_tiefree.npz")
    st = oc.sweep(load_maps(z), "max", "conv", int(z["k"]))
    assert (st.bits == z["ref_bits"]).all()
    assert (st.ids == z["ref_ids"]).all()


def test_edge_cases(golden):
    z = np.load(golden / "collect_edge.npz")
    st = oc.ActMaxOracle(8)
    st.update(z["nlk_acts"], np.arange(3))
    assert oc.values_equal(st.bits, z["nlk_bits"]).all()
    # all-negative latent keeps placeholders; NaN sorts first
    assert (st.ids[1] == -1).all()
    assert st.ids[3, 0] == 0 and (st.bits[3, 0] & 0x7FFF) > 0x7F80
    assert not oc.check_tie_aware(st.bits, st.ids, z["nlk_bits"], z["nlk_ids"])
    assert z["k0_shape"].tolist() == [3, 0]
    st0 = oc.ActMaxOracle(0)
    st0.update(np.random.randn(4, 3).astype(np.float32), np.arange(4))
    assert st0.bits.shape == (3, 0)


def test_key_roundtrip_and_order():
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 1 << 16, size=4096, dtype=np.uint16)
    bits = np.where((bits & 0x7FFF) > 0x7F80, 0x7FC0, bits).astype(np.uint16)
    ids = rng.integers(-1, 1 << 40, size=4096, dtype=np.int64)
    b2, i2 = oc.topk_unkey(oc.topk_key(bits, ids))
    assert (b2 == bits).all() and (i2 == ids).all()
    # order: value desc, then id asc, placeholders last
    k = oc.topk_key(np.array([0x3F80, 0x3F80, 0x3F80, 0x8000, 0x0000, 0xBF80], np.uint16), np.array([5, 2, -1, -1, 7, 0]))
    order = np.argsort(-k.astype(np.float64), kind="stable")  # coarse check; exact compare below
    assert k[1] > k[0] > k[2] > k[4] > k[3] > k[5]


def test_merge_lists_equals_single_sweep():
    rng = np.random.default_rng(1)
    acts = rng.standard_normal((64, 12)).astype(np.float32)
    whole = oc.ActMaxOracle(7)
    whole.update(acts, np.arange(64))
    parts = []
    for r in range(4):
        p = oc.ActMaxOracle(7)
        p.update(acts[r * 16 : (r + 1) * 16], np.arange(r * 16, (r + 1) * 16))
        parts.append(p)
    bits, ids = oc.merge_lists(np.stack([p.bits for p in parts]), np.stack([p.ids for p in parts]))
    assert (bits == whole.bits).all() and (ids == whole.ids).all()


def test_full_size_reference_fixture(golden):
    """k = 20, batch 256, C = 2048 (ResNet-50 layer4 geometry): the reference's hooks ran over tests/collect_cases.py's
    seeded maps (oracle/make_golden.py:gen_collect_large); the oracle must meet the tie-aware contract against it."""
    from tests.collect_cases import LARGE, large_maps

    z = np.load(golden / "collect_large.npz")
    maps = large_maps()
    np.testing.assert_allclose([float(m.astype(np.float64).sum()) for m in maps], z["checksum"], rtol=1e-12)
    st = oc.sweep(maps, "mean", "conv", LARGE["k"], exact=True)
    cand = np.concatenate([oc.f32_to_bf16_bits(oc.aggregate_exact(m, "mean", "conv")) for m in maps]).T
    errs = oc.check_tie_aware(st.bits, st.ids, z["ref_bits"], z["ref_ids"].astype(np.int64), cand)
    assert not errs, errs[:5]
    assert (st.ids >= 0).all() and ((st.bits & 0x7FFF) == 0).any()  # dead channels: real ids on zero values
