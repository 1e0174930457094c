"""GPU parity: K1 (aggregate) and K2 (top-k) through the C-ABI vs the oracle and the reference fixtures."""

import numpy as np
import pytest
import torch

from oracle import collect as oc

pytestmark = pytest.mark.gpu

OPS = {"mean": 0, "max": 1, "absmean": 2, "absmax": 3, "token": 4}
AGGS = {
    "aggregate_conv_mean": ("mean", "conv"),
    "aggregate_conv_max": ("max", "conv"),
    "aggregate_transformer_mean": ("mean", "tokens"),
    "aggregate_transformer_absmean": ("absmean", "tokens"),
    "aggregate_transformer_max": ("max", "tokens"),
    "aggregate_transformer_absmax": ("absmax", "tokens"),
    "aggregate_transformer_special_token": ("token", "tokens"),
}


def bits_of(t):
    return t.cpu().view(torch.int16).numpy().view(np.uint16)


def f32_bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_f32_bitexact(got, want):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    same = (f32_bits(got) == f32_bits(want)) | ((got == 0) & (want == 0)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), f"{(~same).sum()} of {same.size} differ; first {np.argwhere(~same)[0]}: {got[~same][0]} vs {want[~same][0]}"


@pytest.fixture(scope="module")
def ops():
    from semanticlens_b200 import ops

    return ops


# ---------------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------------
CONV_SHAPES = [
    (3, 5, 7, 7),  # L = 49: staged/small with unaligned rows (ResNet layer4)
    (4, 16, 14, 14),  # L = 196 (layer3)
    (2, 8, 28, 28),  # L = 784 (layer2)
    (2, 4, 56, 56),  # L = 3136: staged/large, one chunk per row (layer1)
    (2, 3, 112, 112),  # L = 12544: staged/large, two chunks per row (conv1)
    (1, 1, 1, 1),
    (5, 3, 1, 3),
    (2, 2, 33, 31),  # L = 1023, odd -> rows misaligned
    (1, 2, 90, 91),  # L = 8190 not a multiple of 4 -> direct CTA path
    (37, 19, 5, 5),
    (300, 64, 7, 7),  # many tiles, tail tile
    (40, 16, 4, 4),  # L = 16: sub-warp rows, one accumulator column
    (9, 33, 8, 8),  # L = 64: sub-warp rows, 8 lanes per row, two columns
    (7, 11, 9, 9),  # L = 81: 16 lanes per row, three columns
    (5, 13, 10, 10),  # L = 100
    (6, 10, 8, 16),  # L = 128
    (5, 9, 12, 12),  # L = 144: five columns
    (3, 17, 13, 13),  # L = 169
    (4, 6, 15, 15),  # L = 225
    (2, 8, 16, 16),  # L = 256: the longest sub-warp row
    (3, 5, 52, 52),  # L = 2704: 3 whole rows per stage, row count not a multiple of 3
    (2, 7, 32, 64),  # L = 2048: 4 rows per stage
    (40, 32, 56, 56),  # L = 3136, 1280 rows: more 2-row tiles than resident CTAs
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("op", ["mean", "max", "absmean", "absmax"])
def test_k1_conv_bitexact_vs_canonical_oracle(ops, shape, op):
    g = torch.Generator().manual_seed(hash((shape, op)) % 2**31)
    x = torch.randn(*shape, generator=g)
    got = ops.agg_reduce(x.cuda(), OPS[op], "conv").cpu().numpy()
    assert_f32_bitexact(got, oc.aggregate_canonical(x.numpy(), op, "conv"))
    exact = oc.aggregate_exact(x.numpy(), op, "conv")
    np.testing.assert_allclose(got, exact, rtol=0, atol=1e-6)


@pytest.mark.parametrize("shape", [(4, 6, 7, 7), (3, 4, 10, 10), (2, 3, 56, 56)])
def test_k1_signed_zero_rows_follow_the_canonical_fold(ops, shape):
    """All-(-0.0) and mixed-zero rows: the identity folds of the canonical order decide the sign of a zero result."""
    x = torch.zeros(*shape)
    x[::2] = -0.0
    x[0, 0].view(-1)[::3] = -0.0
    for op in ("mean", "max", "absmax"):
        got = ops.agg_reduce(x.cuda(), OPS[op], "conv").cpu().numpy()
        assert_f32_bitexact(got, oc.aggregate_canonical(x.numpy(), op, "conv"))


def test_k1_conv_staged_equals_direct(ops, monkeypatch):
    """Unaligned base pointer forces the direct kernels: same canonical order, same bits."""
    x = torch.randn(6, 12, 14, 14, device="cuda")
    a = ops.agg_reduce(x, 0, "conv")
    buf = torch.empty(x.numel() + 1, device="cuda")
    y = buf[1:].view_as(x)  # 4-byte aligned only
    y.copy_(x)
    assert y.data_ptr() % 16 != 0
    b = ops.agg_reduce(y, 0, "conv")
    assert torch.equal(a, b)


TOK_SHAPES = [(3, 13, 40), (2, 197, 768), (4, 50, 768), (2, 7, 5), (3, 257, 1024), (2, 70, 300), (1, 1, 1), (2, 65, 260)]


@pytest.mark.parametrize("shape", TOK_SHAPES)
@pytest.mark.parametrize("op", ["mean", "max", "absmean", "absmax"])
def test_k1_tokens_bitexact_vs_canonical_oracle(ops, shape, op):
    g = torch.Generator().manual_seed(hash((shape, op)) % 2**31)
    x = torch.randn(*shape, generator=g)
    got = ops.agg_reduce(x.cuda(), OPS[op], "tokens").cpu().numpy()
    assert_f32_bitexact(got, oc.aggregate_canonical(x.numpy(), op, "tokens"))


# the (B, T, F) kernel picks its per-lane vector width from F, the alignment and the CTA count, and stages block partials in
# chunks: long chains (several chunks), wide batches (16-byte lanes), odd widths (scalar lanes), 16-bit inputs, unaligned bases
@pytest.mark.parametrize("shape,dtype,offset", [
    ((1, 17000, 32), torch.float32, 0),      # 266 blocks of 64 tokens: two partial chunks at one element per lane
    ((2, 12544, 64), torch.float32, 0),      # the channels-last conv1 map of ResNet-50
    ((320, 70, 128), torch.float32, 0),      # enough CTAs for 16-byte lanes
    ((320, 130, 130), torch.float32, 0),     # F % 4 != 0: 8-byte lanes, a partial feature group
    ((3, 100, 37), torch.float32, 0),        # odd width: scalar lanes
    ((40, 90, 64), torch.float32, 1),        # base pointer 4-byte aligned only
    ((700, 66, 256), torch.float16, 0),      # 16-byte lanes of eight halves
    ((5, 200, 30), torch.float16, 0),
    ((6, 129, 72), torch.float16, 1),        # 2-byte aligned base
])
@pytest.mark.parametrize("op", ["mean", "absmax"])
def test_k1_tokens_vector_widths_and_chunks(ops, shape, dtype, offset, op):
    g = torch.Generator().manual_seed(hash((shape, op)) % 2**31)
    x = torch.randn(*shape, generator=g).to(dtype)
    buf = torch.empty(x.numel() + 8, dtype=dtype, device="cuda")
    xd = buf[offset : offset + x.numel()].view(shape)
    xd.copy_(x)
    got = ops.agg_reduce(xd, OPS[op], "tokens").cpu().numpy()
    assert_f32_bitexact(got, oc.aggregate_canonical(x.numpy(), op, "tokens"))


@pytest.mark.parametrize("pos", [0, 3, -1])
def test_k1_special_token(ops, pos):
    x = torch.randn(4, 9, 33)
    got = ops.agg_reduce(x.cuda(), 4, "tokens", pos).cpu()
    assert torch.equal(got, x[:, pos])


def test_k1_layout_variants(ops):
    x = torch.randn(4, 24, 10, 12, device="cuda")
    base = ops.agg_reduce(x, 1, "conv")
    cl = x.contiguous(memory_format=torch.channels_last)
    assert torch.equal(ops.agg_reduce(cl, 1, "conv"), base)  # max is order independent
    np.testing.assert_allclose(
        ops.agg_reduce(cl, 0, "conv").cpu().numpy(), oc.aggregate_exact(x.cpu().numpy(), "mean", "conv"), atol=1e-6
    )
    sl = x[:, ::2]  # non-contiguous -> copied
    assert torch.equal(ops.agg_reduce(sl, 1, "conv"), base[:, ::2])
    t = torch.randn(3, 40, 17, device="cuda").transpose(1, 2)  # (B, T=17, F=40) stored as (B, F, T)
    assert torch.equal(ops.agg_reduce(t, 1, "tokens"), t.amax(1))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_k1_half_inputs(ops, dtype):
    x = torch.randn(3, 10, 14, 14).to(dtype)
    got = ops.agg_reduce(x.cuda(), 0, "conv").cpu()
    ref = x.float().flatten(2).double().mean(-1).float().to(dtype).float()
    # one rounding to the input dtype; fp32 accumulation order may move a value across a rounding boundary
    assert (got - ref).abs().max() <= 2 * torch.finfo(dtype).eps * ref.abs().max()
    assert torch.equal(got.to(dtype).float(), got)
    xm = ops.agg_reduce(x.cuda(), 1, "conv").cpu()
    assert torch.equal(xm, x.float().flatten(2).amax(-1))


def test_k1_nan_inf_propagation(ops):
    x = torch.randn(2, 4, 6, 6)
    x[0, 0, 2, 3] = float("nan")
    x[1, 1, 0, 0] = float("inf")
    x[1, 2] = float("-inf")
    for op, name in ((0, "mean"), (1, "max")):
        got = ops.agg_reduce(x.cuda(), op, "conv").cpu().numpy()
        want = oc.aggregate_exact(x.numpy(), name, "conv")
        assert np.isnan(got[0, 0]) and np.isnan(want[0, 0])
        assert got[1, 1] == np.inf and got[1, 2] == -np.inf


def test_k1_exact_arithmetic_inputs_match_torch(ops):
    """Sums of small integers are exact in any order: K1 must equal torch's CPU mean bit for bit."""
    g = torch.Generator().manual_seed(5)
    x = torch.randint(-64, 64, (8, 32, 14, 14), generator=g).float()
    assert torch.equal(ops.agg_reduce(x.cuda(), 0, "conv").cpu(), x.flatten(2).mean(-1))
    t = torch.randint(-64, 64, (8, 50, 96), generator=g).float()
    assert torch.equal(ops.agg_reduce(t.cuda(), 0, "tokens").cpu(), t.mean(1))
    assert torch.equal(ops.agg_reduce(t.cuda(), 2, "tokens").cpu(), t.abs().mean(1))


def test_k1_batch_invariance(ops):
    x = torch.randn(64, 48, 7, 7, device="cuda")
    whole = ops.agg_reduce(x, 0, "conv")
    parts = torch.cat([ops.agg_reduce(x[i : i + 5].contiguous(), 0, "conv") for i in range(0, 64, 5)])
    assert torch.equal(whole, parts)


def test_k1_wrong_rank_raises():
    from semanticlens_b200.component_visualization import aggregators as A

    with pytest.raises(ValueError, match="Input tensor should be 4D"):
        A.aggregate_conv_mean(torch.randn(2, 4, 8))
    with pytest.raises(ValueError, match="Input tensor should be 3D"):
        A.aggregate_transformer_max(torch.randn(2, 10, 16, 1))
    assert A.aggregate_conv_mean(torch.randn(2, 4, 8, 8)).shape == (2, 4)
    assert A.aggregate_transformer_mean(torch.randn(2, 10, 16)).shape == (2, 16)
    assert not A.aggregate_conv_max(torch.randn(2, 4, 8, 8)).is_cuda


# ---------------------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------------------
def fresh_state(C, k):
    v = (-torch.zeros(C, k, dtype=torch.bfloat16)).cuda()
    i = (-torch.ones(C, k, dtype=torch.int64)).cuda()
    return v, i


@pytest.mark.parametrize("C,k,batches", [(3, 5, (2, 2)), (64, 20, (64, 64, 17)), (130, 1, (7, 300)), (33, 256, (256, 256, 40)), (16, 7, (1000,)), (9, 300, (100, 50))])
def test_k2_matches_oracle_exactly(ops, C, k, batches):
    rng = np.random.default_rng(C * 1000 + k)
    v, i = fresh_state(C, k)
    st = oc.ActMaxOracle(k, C)
    n = 0
    for b in batches:
        acts = rng.standard_normal((b, C)).astype(np.float32)
        acts[rng.random((b, C)) < 0.1] = 0.0  # exact zeros tie with the -0.0 placeholders
        acts = np.round(acts * 8) / 8  # coarse grid => many bf16 ties
        ops.topk_update(torch.from_numpy(acts).cuda(), v, i, None, n)
        st.update(acts, np.arange(n, n + b))
        n += b
        assert (bits_of(v) == st.bits).all()
        assert (i.cpu().numpy() == st.ids).all()


def test_k2_explicit_ids_and_bf16_candidates(ops):
    rng = np.random.default_rng(3)
    C, k, B = 10, 6, 40
    acts = torch.from_numpy(rng.standard_normal((B, C)).astype(np.float32))
    ids = torch.from_numpy(rng.permutation(1000)[:B].astype(np.int64))
    v, i = fresh_state(C, k)
    ops.topk_update(acts.to(torch.bfloat16).cuda(), v, i, ids.cuda(), 0)
    st = oc.ActMaxOracle(k, C)
    st.update(acts.numpy(), ids.numpy())
    assert (bits_of(v) == st.bits).all() and (i.cpu().numpy() == st.ids).all()


def test_k2_reference_kat(golden):
    """The reference's only numeric KAT, through the drop-in ActMax class."""
    from semanticlens_b200.component_visualization.activation_caching import ActMax

    z = np.load(golden / "actmax_kat.npz")
    am = ActMax(n_collect=5, n_latents=3)
    assert am.is_setup
    am.update(torch.from_numpy(z["acts1"]), torch.tensor([0, 1]))
    am.update(torch.from_numpy(z["acts2"]), torch.tensor([2, 3]))
    assert torch.allclose(am.activations[0], torch.tensor([0.9, 0.8, 0.2, 0.1, 0.0]).to(torch.bfloat16))
    assert torch.allclose(am.sample_ids[0], torch.tensor([2, 3, 1, 0, -1]))
    assert (bits_of(am.activations) == z["ref_bits"]).all()
    assert (am.sample_ids.numpy() == z["ref_ids"]).all()


def test_k2_edge_cases(golden):
    from semanticlens_b200.component_visualization.activation_caching import ActMax

    z = np.load(golden / "collect_edge.npz")
    am = ActMax(n_collect=8)
    am.update(torch.from_numpy(z["nlk_acts"]), torch.arange(3))
    errs = oc.check_tie_aware(bits_of(am.activations), am.sample_ids.numpy(), z["nlk_bits"], z["nlk_ids"])
    assert not errs, errs
    assert (am.sample_ids[1] == -1).all()  # all-negative latent keeps its placeholders
    assert torch.isnan(am.activations[3, 0].float()) and am.sample_ids[3, 0] == 0  # NaN sorts first
    a0 = ActMax(n_collect=0)
    a0.update(torch.randn(4, 3), torch.arange(4))
    assert tuple(a0.activations.shape) == (3, 0) and tuple(a0.sample_ids.shape) == (3, 0)
    assert a0.alive_latents.numel() == 0


def test_k2_merge_lists(ops):
    rng = np.random.default_rng(11)
    C, k, R = 50, 20, 8
    acts = np.round(rng.standard_normal((R * 32, C)).astype(np.float32) * 4) / 4
    whole = oc.ActMaxOracle(k, C)
    whole.update(acts, np.arange(R * 32))
    vs, is_ = [], []
    for r in range(R):
        v, i = fresh_state(C, k)
        ops.topk_update(torch.from_numpy(acts[r * 32 : (r + 1) * 32]).cuda(), v, i, None, r * 32)
        vs.append(v)
        is_.append(i)
    mv, mi = ops.topk_merge_lists(torch.stack(vs), torch.stack(is_))
    assert (bits_of(mv) == whole.bits).all() and (mi.cpu().numpy() == whole.ids).all()


def test_k5_gather_rows(ops):
    table = torch.randn(37, 24, device="cuda")
    idx = torch.tensor([[0, -1, 5], [36, 7, -1]])
    assert torch.equal(ops.gather_rows(table, idx).cpu(), table.cpu()[idx])
    t2 = torch.randn(9, 7, device="cuda")
    assert torch.equal(ops.gather_rows(t2, torch.tensor([8, -9, 3])).cpu(), t2.cpu()[torch.tensor([8, -9, 3])])


# ---------------------------------------------------------------------------------------------------
# hook path vs the reference fixtures
# ---------------------------------------------------------------------------------------------------
class _Emit(torch.nn.Module):
    def forward(self, x):
        return x


def run_hooks(maps, agg_fn, k):
    from semanticlens_b200.component_visualization.activation_caching import ActMaxCache

    model = torch.nn.Sequential()
    model.add_module("probe", _Emit())
    cache = ActMaxCache(["probe"], agg_fn, k)
    with cache.hook_context(model):
        for m in maps:
            model(torch.from_numpy(m).cuda())
    return cache.cache["probe"]


def get_agg(name):
    from semanticlens_b200.component_visualization import aggregators as A

    if name == "aggregate_transformer_special_token":
        return A.get_aggregate_transformer_special_token(0)
    return getattr(A, name)


@pytest.mark.parametrize("name", sorted(AGGS))
def test_hook_path_vs_reference_fixture(golden, name):
    op, kind = AGGS[name]
    z = np.load(golden / f"collect_{name}.npz")
    maps = [z[f"map{i}"] for i in range(int(z["n_batches"]))]
    am = run_hooks(maps, get_agg(name), int(z["k"]))
    assert not am.activations.is_cuda and am.activations.dtype == torch.bfloat16
    # exact against the oracle in canonical order
    st = oc.sweep(maps, op, kind, int(z["k"]))
    assert (bits_of(am.activations) == st.bits).all()
    assert (am.sample_ids.numpy() == st.ids).all()
    # tie-aware against the reference's own output; values whose fp32 aggregate sits within 4 ulp of a bf16
    # rounding midpoint would be excused (none occur in the fixture)
    cand = np.concatenate([oc.f32_to_bf16_bits(oc.aggregate_canonical(m, op, kind)) for m in maps]).T
    errs = oc.check_tie_aware(bits_of(am.activations), am.sample_ids.numpy(), z["ref_bits"], z["ref_ids"], cand)
    assert not errs, errs


def test_hook_path_full_size_reference_fixture(golden):
    """k = 20, batch 256, C = 2048 (ResNet-50 layer4 geometry), 3 batches: the reference's hooks ran over the same seeded
    maps (oracle/make_golden.py:gen_collect_large). Values bit-exact, ids under the tie-aware contract, and exact against
    the canonical oracle."""
    from tests.collect_cases import LARGE, large_maps

    z = np.load(golden / "collect_large.npz")
    maps = large_maps()
    np.testing.assert_allclose([float(m.astype(np.float64).sum()) for m in maps], z["checksum"], rtol=1e-12)
    am = run_hooks(maps, get_agg("aggregate_conv_mean"), LARGE["k"])
    got_bits, got_ids = bits_of(am.activations), am.sample_ids.numpy()
    assert oc.values_equal(got_bits, z["ref_bits"]).all()  # bit-exact bf16 values against the reference
    cand = np.concatenate([oc.f32_to_bf16_bits(oc.aggregate_canonical(m, "mean", "conv")) for m in maps]).T
    errs = oc.check_tie_aware(got_bits, got_ids, z["ref_bits"], z["ref_ids"].astype(np.int64), cand)
    assert not errs, errs[:5]
    st = oc.sweep(maps, "mean", "conv", LARGE["k"])
    assert (got_bits == st.bits).all() and (got_ids == st.ids).all()


def test_hook_path_special_token_on_a_permuted_map():
    """A (B, T, F) map whose memory is (B, F, T) (e.g. the output of a permute): the reference indexes it like any other
    tensor; the token select copies it once instead of failing."""
    from semanticlens_b200.component_visualization import aggregators as A
    from semanticlens_b200.component_visualization.activation_caching import ActMaxCache

    class Perm(torch.nn.Module):
        def forward(self, x):
            return x.permute(0, 2, 1)

    model = torch.nn.Sequential()
    model.add_module("probe", Perm())
    x = torch.randn(6, 24, 9, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))  # (B, F, T)
    cache = ActMaxCache(["probe"], A.get_aggregate_transformer_special_token(-2), 4)
    with cache.hook_context(model):
        model(x)
    st = oc.sweep([x.permute(0, 2, 1).cpu().numpy()], "token", "tokens", 4, token=7)
    assert (bits_of(cache.cache["probe"].activations) == st.bits).all()
    assert (cache.cache["probe"].sample_ids.numpy() == st.ids).all()
    bad = ActMaxCache(["probe"], A.get_aggregate_transformer_special_token(9), 4)
    with pytest.raises(IndexError), bad.hook_context(model):
        model(x)
    # float64 maps are cast, not rejected
    out = A.aggregate_conv_mean(torch.randn(2, 3, 4, 4, dtype=torch.float64).cuda())
    assert out.shape == (2, 3)


def test_hook_path_tiefree_ids_exact(golden):
    z = np.load(golden / "collect_tiefree.npz")
    maps = [z[f"map{i}"] for i in range(int(z["n_batches"]))]
    am = run_hooks(maps, get_agg("aggregate_conv_max"), int(z["k"]))
    assert (bits_of(am.activations) == z["ref_bits"]).all()
    assert (am.sample_ids.numpy() == z["ref_ids"]).all()


def test_hook_path_custom_aggregation_fn():
    """A user-defined aggregation function goes through ActMax.update (K2 only)."""

    def my_agg(t):
        return t.flatten(2).amax(-1).cpu()

    maps = [np.random.default_rng(i).standard_normal((6, 5, 4, 4)).astype(np.float32) for i in range(3)]
    am = run_hooks(maps, my_agg, 4)
    st = oc.sweep(maps, "max", "conv", 4)
    assert (bits_of(am.activations) == st.bits).all() and (am.sample_ids.numpy() == st.ids).all()


def test_batch_size_invariance_of_ids():
    """Unlike the reference the ids do not depend on the batch size."""
    rng = np.random.default_rng(0)
    data = np.maximum(rng.standard_normal((200, 16, 3, 3)), 0).astype(np.float32)  # post-ReLU: many ties
    outs = []
    for bs in (16, 64, 200):
        maps = [data[i : i + bs] for i in range(0, 200, bs)]
        am = run_hooks(maps, get_agg("aggregate_conv_mean"), 20)
        outs.append((bits_of(am.activations).copy(), am.sample_ids.numpy().copy()))
    for b, i in outs[1:]:
        assert (b == outs[0][0]).all() and (i == outs[0][1]).all()


def test_store_load_roundtrip_and_reference_cache(tmp_path, golden):
    from semanticlens_b200.component_visualization import aggregators as A
    from semanticlens_b200.component_visualization.activation_caching import ActMaxCache

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.ReLU(), torch.nn.Conv2d(8, 16, 3)).cuda()
    cache = ActMaxCache(["0", "2"], A.aggregate_conv_mean, 10)
    with cache.hook_context(model):
        model(torch.randn(4, 3, 32, 32, device="cuda"))
    assert cache.cache["0"].activations.shape == (8, 10) and cache.cache["2"].activations.shape == (16, 10)
    cache.store(tmp_path / "c")
    assert sorted(p.name for p in (tmp_path / "c").iterdir()) == [
        "aggregate_conv_mean-10-0.safetensors",
        "aggregate_conv_mean-10-2.safetensors",
    ]
    again = ActMaxCache(["0", "2"], A.aggregate_conv_mean, 10)
    again.load(tmp_path / "c")
    assert torch.equal(again.cache["0"].activations, cache.cache["0"].activations)
    assert torch.equal(again.cache["2"].sample_ids, cache.cache["2"].sample_ids)
    # a cache directory written by the reference loads as-is
    ref = ActMaxCache(["0", "2"], A.aggregate_conv_mean, 3)
    ref.load(golden / "cache_format")
    assert ref.cache["0"].activations.shape == (4, 3) and ref.cache["2"].sample_ids.dtype == torch.int64
    with pytest.raises(FileNotFoundError):
        ActMaxCache(["0"], A.aggregate_conv_max, 3).load(golden / "cache_format")


# ---------------------------------------------------------------------------------------------------
# BASELINE cfg-2 sizes: size-independent properties (the oracle is too slow here)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(256, 64, 112, 112), (256, 256, 56, 56), (256, 512, 28, 28), (256, 1024, 14, 14),
                                   (256, 2048, 7, 7)])
def test_full_size_layer_properties(ops, shape):
    """One ResNet-50 layer of cfg 2 at batch 256, three batches, k = 20:
    (1) K1 equals a float64 mean to fp32 round-off; (2) the streamed top-k state equals torch.topk over ALL candidates
    (values bit-exact: the value multiset is well defined); (3) rows are sorted, ids are unique, every id's own
    aggregate rounds to the stored value; (4) a sweep fed in a different batch split gives the identical state."""
    from semanticlens_b200.component_visualization.activation_caching import ActMax

    B, C = shape[:2]
    k = 20
    g = torch.Generator(device="cuda").manual_seed(C)
    maps = [torch.relu(torch.randn(*shape, device="cuda", generator=g)) for _ in range(3)]
    aggs = [ops.agg_reduce(m, 0, "conv") for m in maps]
    for m, a in zip(maps, aggs):
        ref = m.double().flatten(2).mean(-1)
        assert (a.double() - ref).abs().max() <= 2e-7 * ref.abs().max()
    am = ActMax(k)
    am2 = ActMax(k)
    for i, m in enumerate(maps):
        am.update_from_map(m, 0, "conv", 0, i * B)
    half = B // 2
    for i, m in enumerate(maps):  # same data, batches of 128
        am2.update_from_map(m[:half], 0, "conv", 0, i * B)
        am2.update_from_map(m[half:], 0, "conv", 0, i * B + half)
    vals, ids = am.activations, am.sample_ids
    assert torch.equal(vals.view(torch.int16), am2.activations.view(torch.int16)) and torch.equal(ids, am2.sample_ids)
    cand = torch.cat(aggs).T.to(torch.bfloat16).cpu()  # (C, 3B)
    want = torch.topk(cand.float(), k, dim=1).values.to(torch.bfloat16)
    assert torch.equal(vals.view(torch.int16), want.view(torch.int16))
    v32 = vals.float()
    assert (v32[:, :-1] >= v32[:, 1:]).all()
    assert all(len(set(r.tolist())) == k for r in ids[:64])
    assert torch.equal(torch.gather(cand, 1, ids).view(torch.int16), vals.view(torch.int16))
