"""ORACLE tooling — generate tests/golden/*.npz by running the REAL reference (imported read-only from
/root/reference with crp/zennit stubbed, see oracle/refshim.py) on small seeded inputs.

Run in the build container (the reference does not travel to the GPU box):

    python -m oracle.make_golden

Fixtures (inputs + the reference's outputs):
  collect_<agg>.npz     hooked maps for 3 batches -> ActMaxCache state (bf16 bits, ids) via the reference's hooks
  collect_tiefree.npz   same with aggregates that are pairwise distinct in bf16 (ids must match exactly)
  collect_edge.npz      N < k, all-negative channel, exact zeros, NaN, k = 0
  actmax_kat.npz        the reference's own known-answer test (tests/component_visualization/test_activation_caching.py:14-30)
  scores.npz            clarity / similarity (all shape branches) / polysemanticity (incl. the small-cluster fallback)
  scores_poly.npz       reference polysemanticity / clarity of the seeded cases of tests/polysem_cases.py (outputs only)
  scores_poly_general.npz  the same for n_clusters = 2..8 and up to 520 examples per neuron (GENERAL_CASES, outputs only)
  cache_format/         one ActMaxCache.store() directory written by the reference (file names, keys, metadata)
  collect_large.npz     full-size hook fixture: 3 batches of (256, 2048, 7, 7) post-ReLU maps (regenerated from a seed by
                        tests/collect_cases.py; only a checksum of the inputs and the reference's outputs are stored), k = 20
  text_probe.npz        lens._embed_text_probes with and without templates over a deterministic stand-in FM
                        (tests/collect_cases.py:FakeTextFM): pins the template-major / (q t) regrouping
"""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import refshim  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def bits_of(t: torch.Tensor) -> np.ndarray:
    return t.view(torch.int16).numpy().view(np.uint16).copy()


class _Emit(torch.nn.Module):
    """Identity module whose output is hooked; lets us feed arbitrary maps through the reference's hook path."""

    def forward(self, x):
        return x


def run_reference_hooks(maps, agg_fn, k):
    from semanticlens.component_visualization.activation_caching import ActMaxCache

    model = torch.nn.Sequential()
    model.add_module("probe", _Emit())
    cache = ActMaxCache(["probe"], agg_fn, k)
    with cache.hook_context(model):
        for m in maps:
            model(m)
    am = cache.cache["probe"]
    return am


def gen_collect():
    from semanticlens.component_visualization import aggregators as A

    g = torch.Generator().manual_seed(1234)
    conv_maps = [torch.randn(b, 24, 7, 9, generator=g) for b in (8, 8, 5)]
    conv_maps[1][:, 3] = -conv_maps[1][:, 3].abs() - 0.5  # channel 3 negative in batch 1
    for m in conv_maps:
        m[:, 5] = -m[:, 5].abs() - 0.1  # channel 5 always negative -> placeholders survive under mean
        m[:, 6] = torch.relu(m[:, 6] - 3.0)  # channel 6 mostly exact zeros (post-ReLU dead-ish channel)
    tok_maps = [torch.randn(b, 13, 40, generator=g) for b in (8, 8, 5)]
    cases = {
        "aggregate_conv_mean": (A.aggregate_conv_mean, conv_maps),
        "aggregate_conv_max": (A.aggregate_conv_max, conv_maps),
        "aggregate_transformer_mean": (A.aggregate_transformer_mean, tok_maps),
        "aggregate_transformer_absmean": (A.aggregate_transformer_absmean, tok_maps),
        "aggregate_transformer_max": (A.aggregate_transformer_max, tok_maps),
        "aggregate_transformer_absmax": (A.aggregate_transformer_absmax, tok_maps),
        "aggregate_transformer_special_token": (A.get_aggregate_transformer_special_token(0), tok_maps),
    }
    for name, (fn, maps) in cases.items():
        am = run_reference_hooks(maps, fn, 6)
        aggs = [fn(m).numpy() for m in maps]
        np.savez_compressed(
            GOLD / f"collect_{name}.npz",
            **{f"map{i}": m.numpy() for i, m in enumerate(maps)},
            **{f"agg{i}": a for i, a in enumerate(aggs)},
            n_batches=len(maps),
            k=6,
            ref_bits=bits_of(am.activations),
            ref_ids=am.sample_ids.numpy(),
        )

    # tie-free: per-(image, channel) maxima are distinct bf16 values
    n_img, C = 40, 16
    vals = (torch.arange(n_img * C, dtype=torch.int32) + 0x3C00).to(torch.int16).view(torch.bfloat16).float()
    vals = vals.reshape(n_img, C)  # consecutive bf16 bit patterns from 2^-7 upwards: exact and pairwise distinct
    perm = torch.stack([torch.randperm(n_img, generator=g) for _ in range(C)], dim=1)
    vals = torch.gather(vals, 0, perm)
    assert len(np.unique(bits_of(vals.to(torch.bfloat16)))) == n_img * C
    maps = []
    for b0 in range(0, n_img, 16):
        v = vals[b0 : b0 + 16]
        m = torch.rand(v.shape[0], C, 5, 5, generator=g) * 0.005  # below every planted maximum
        m[:, :, 2, 3] = v
        maps.append(m)
    am = run_reference_hooks(maps, A.aggregate_conv_max, 10)
    np.savez_compressed(
        GOLD / "collect_tiefree.npz",
        **{f"map{i}": m.numpy() for i, m in enumerate(maps)},
        n_batches=len(maps),
        k=10,
        ref_bits=bits_of(am.activations),
        ref_ids=am.sample_ids.numpy(),
    )

    # edge cases through ActMax.update directly
    from semanticlens.component_visualization.activation_caching import ActMax

    edge = {}
    a = ActMax(n_collect=8)  # N < k
    acts = torch.tensor([[0.5, -1.0, 0.0, float("nan")], [0.25, -2.0, 0.0, 1.0], [1.5, -3.0, -0.0, 2.0]])
    a.update(acts, torch.arange(3))
    edge["nlk_acts"], edge["nlk_bits"], edge["nlk_ids"] = acts.numpy(), bits_of(a.activations), a.sample_ids.numpy()
    a0 = ActMax(n_collect=0)  # k = 0
    a0.update(torch.randn(4, 3), torch.arange(4))
    edge["k0_shape"] = np.array(a0.activations.shape)
    np.savez_compressed(GOLD / "collect_edge.npz", **edge)

    # the reference's own KAT
    am = ActMax(n_collect=5, n_latents=3)
    a1 = torch.tensor([[0.1, 0.9, 0.3], [0.2, 0.8, 0.4]])
    a2 = torch.tensor([[0.9, 0.1, 0.5], [0.8, 0.2, 0.6]])
    am.update(a1, torch.tensor([0, 1]))
    am.update(a2, torch.tensor([2, 3]))
    np.savez_compressed(
        GOLD / "actmax_kat.npz", acts1=a1.numpy(), acts2=a2.numpy(), ref_bits=bits_of(am.activations),
        ref_ids=am.sample_ids.numpy(),
    )


def gen_cache_format():
    from semanticlens.component_visualization import aggregators as A
    from semanticlens.component_visualization.activation_caching import ActMaxCache

    torch.manual_seed(7)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.ReLU(), torch.nn.Conv2d(4, 6, 3))
    cache = ActMaxCache(["0", "2"], A.aggregate_conv_mean, 3)
    with cache.hook_context(model):
        model(torch.randn(5, 3, 8, 8))
    d = GOLD / "cache_format"
    d.mkdir(exist_ok=True)
    for f in d.glob("*.safetensors"):
        f.unlink()
    cache.store(d)


def gen_scores():
    from semanticlens import scores as S

    g = torch.Generator().manual_seed(99)
    out = {}
    V = torch.randn(12, 16, 32, generator=g)
    out["V"] = V.numpy()
    out["clarity"] = S.clarity_score(V).numpy()
    # similarity: general (Q,D)x(C,D), the C == D branch (no transpose), and the equal-shape row-wise branch
    x, y = torch.randn(5, 32, generator=g), torch.randn(9, 32, generator=g)
    out["sim_x"], out["sim_y"], out["sim_xy"] = x.numpy(), y.numpy(), S.similarity_score(x, y).numpy()
    y2 = torch.randn(32, 32, generator=g)
    out["sim_y2"], out["sim_xy2"] = y2.numpy(), S.similarity_score(x, y2).numpy()
    x3 = torch.randn(9, 32, generator=g)
    out["sim_x3"], out["sim_x3y"] = x3.numpy(), S.similarity_score(x3, y).numpy()
    # polysemanticity: gaussian, planted 2 clusters, and rows that force the "< 2 members" fallback
    P = torch.randn(10, 24, 16, generator=g)
    shift = torch.randn(10, 1, 16, generator=g) * 3
    P[:5, :12] += shift[:5]
    P[8, 1:] = P[8, 1:2]  # one outlier + 23 identical rows -> a cluster with a single member
    P[9] = P[9, :1]  # all rows identical -> KMeans finds one distinct label
    out["P"] = P.numpy()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["poly"] = S.polysemanticity_score(P).numpy()
        out["poly_noreplace"] = S.polysemanticity_score(P, replace_empty_clusters=False).numpy()
    R = torch.randn(10, 15, 128, generator=g)
    out["red_in"], out["red"] = R.numpy(), S.redundancy_score(R).numpy()
    R2 = torch.randn(40, 64, generator=g)
    out["red2_in"], out["red2"] = R2.numpy(), S.redundancy_score(R2).numpy()
    np.savez_compressed(GOLD / "scores.npz", **out)


def gen_scores_poly():
    """Larger polysemanticity cases, inputs regenerated from seeds by the tests (only the reference's outputs are
    stored): see tests/polysem_cases.py for the generators."""
    import warnings

    from semanticlens import scores as S
    from tests.polysem_cases import CASES, make_case

    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name in CASES:
            V = torch.from_numpy(make_case(name))
            out[f"{name}.poly"] = S.polysemanticity_score(V).numpy()
            out[f"{name}.poly_noreplace"] = S.polysemanticity_score(V, replace_empty_clusters=False).numpy()
            out[f"{name}.clarity"] = S.clarity_score(V).numpy()
    np.savez_compressed(GOLD / "scores_poly.npz", **out)


def gen_scores_poly_general():
    """polysemanticity_score with n_clusters != 2 and / or more than 256 examples per neuron (the general kernel K8g)."""
    import warnings

    from semanticlens import scores as S
    from tests.polysem_cases import GENERAL_CASES, make_general_case

    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, spec in GENERAL_CASES.items():
            V = torch.from_numpy(make_general_case(name))
            out[f"{name}.poly"] = S.polysemanticity_score(V, n_clusters=spec[4]).numpy()
            out[f"{name}.poly_noreplace"] = S.polysemanticity_score(V, replace_empty_clusters=False, n_clusters=spec[4]).numpy()
    np.savez_compressed(GOLD / "scores_poly_general.npz", **out)


def gen_collect_large():
    """k = 20, batch 256, C = 2048 (ResNet-50 layer4 geometry) through the reference's hooks."""
    from semanticlens.component_visualization import aggregators as A
    from tests.collect_cases import LARGE, large_maps

    maps = [torch.from_numpy(m) for m in large_maps()]
    am = run_reference_hooks(maps, A.aggregate_conv_mean, LARGE["k"])
    np.savez_compressed(
        GOLD / "collect_large.npz",
        checksum=np.array([float(m.double().sum()) for m in maps]),
        ref_bits=bits_of(am.activations),
        ref_ids=am.sample_ids.numpy().astype(np.int32),
    )


def gen_text_probe():
    from semanticlens import lens as L
    from tests.collect_cases import TEXT_PROBE, FakeTextFM

    fm = FakeTextFM()
    out = {}
    for name, (queries, templates, bs) in TEXT_PROBE.items():
        out[name] = L._embed_text_probes(fm, list(queries), list(templates) if templates else None, bs).numpy()
    np.savez_compressed(GOLD / "text_probe.npz", **out)


def main():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    refshim.import_reference()
    GOLD.mkdir(parents=True, exist_ok=True)
    gen_collect()
    gen_cache_format()
    gen_scores()
    gen_scores_poly()
    gen_scores_poly_general()
    gen_collect_large()
    gen_text_probe()
    for f in sorted(GOLD.rglob("*")):
        if f.is_file():
            print(f"{f.relative_to(ROOT)}  {f.stat().st_size} B")


if __name__ == "__main__":
    main()
