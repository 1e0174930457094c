"""BASELINE.json configs[4] at FULL size inside the GPU suite: a 65 536-neuron x 256-example x 512-d concept DB (34.4 GB in
HBM) scored by K6 / K7 / K8, checked against the torch-CPU port of the reference (sklearn KMeans) on neurons sampled
across the whole range, and through size-independent properties (a neuron's score does not depend on which other
neurons share the launch; ranges; symmetry of the cosine matrix)."""

import warnings

import numpy as np
import pytest
import torch

from oracle import ref_port as rp

pytestmark = pytest.mark.gpu

C, K, D, Q = 65536, 256, 512, 10000


@pytest.fixture(scope="module")
def S():
    from semanticlens_b200 import scores

    return scores


def _concept_db(kind):
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~45 GB of free HBM")
    g = torch.Generator(device="cuda").manual_seed(2)
    V = torch.randn(C, K, D, device="cuda", generator=g)
    if kind == "planted":
        V[:, ::2] += 2 * torch.randn(C, 1, D, device="cuda", generator=g)  # two clusters per neuron
    return V


@pytest.mark.parametrize("kind", ["planted", "gaussian"])
def test_polysemanticity_and_clarity_full_size(S, kind):
    V = _concept_db(kind)
    poly = S.polysemanticity_score(V)
    clar = S.clarity_score(V)
    assert poly.shape == (C,) and poly.dtype == torch.float64 and clar.shape == (C,) and clar.dtype == torch.float32
    assert torch.isfinite(poly).all() and torch.isfinite(clar).all()
    assert (poly >= -1e-9).all() and (poly <= 2 + 1e-9).all()
    assert (clar >= -1.0 / (K - 1) - 1e-6).all() and (clar <= 1 + 1e-6).all()
    # neurons sampled across the whole range against the reference's op sequence on the CPU
    idx = torch.arange(0, C, C // 24, device="cuda")[:24]
    Vs = V[idx].cpu()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_poly = rp.polysemanticity_score(Vs)
    d = (poly[idx].cpu() - ref_poly).abs()
    mism = int((d > 1e-6).sum())
    print(f"{kind}: max |poly - reference| = {d.max():.2e}, neurons in a different optimum: {mism}/24")
    assert d.max() < 1e-4 and mism == 0
    assert (clar[idx].cpu() - rp.clarity_score(Vs)).abs().max() < 1e-5
    # launch-composition invariance: the same neurons scored alone give the same bits
    alone_p, alone_c = S.polysemanticity_score(V[idx].contiguous()), S.clarity_score(V[idx].contiguous())
    assert torch.equal(alone_p, poly[idx]) and torch.equal(alone_c, clar[idx])
    if kind == "planted":
        assert poly.mean() > 0.5  # two well separated clusters per neuron
    del V
    torch.cuda.empty_cache()


def test_text_probing_matmul_full_size(S):
    g = torch.Generator(device="cuda").manual_seed(3)
    text = torch.randn(Q, D, device="cuda", generator=g)
    agg = torch.randn(C, D, device="cuda", generator=g)
    sim = S.similarity_score(text, agg)
    assert sim.shape == (Q, C) and sim.dtype == torch.float32
    assert sim.abs().max() <= 1 + 1e-5
    # slices against the reference port (rows of a cosine matrix are independent)
    for q0, c0 in ((0, 0), (Q - 200, C - 3000), (5000, 30000)):
        ref = rp.similarity_score(text[q0:q0 + 200].cpu(), agg[c0:c0 + 3000].cpu())
        err = (sim[q0:q0 + 200, c0:c0 + 3000].cpu() - ref).abs().max() / ref.abs().max()
        assert err < 1e-4, err
    # cos(a, b) == cos(b, a): the transposed problem on a block
    back = S.similarity_score(agg[:4096], text[:500])  # 500 != D: the transposing branch
    assert (back.T - sim[:500, :4096]).abs().max() < 2e-6
