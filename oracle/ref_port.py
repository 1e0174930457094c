"""ORACLE (test infrastructure, not product code) — torch-CPU port of the reference's concept-DB hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

The reference is pure Python on top of torch and cannot travel to the GPU box, so its CPU
timing there is taken on this port. It restates the reference's own op sequence — not the kernels' — so that
the time measured is the reference algorithm's:

  * aggregation            `tensor.clone().flatten(2).mean(-1).detach().cpu()` and siblings
                           (reference semanticlens/component_visualization/aggregators.py:61,87,114,141,168,195,242)
  * top-k state update     bf16 cast of the transposed aggregate, id repeat, two cats, torch.topk, gather
                           (reference .../activation_caching.py:101-141)
  * hook + id numbering    per-layer running counter (reference .../activation_caching.py:403-416)
  * sweep loop             DataLoader order, `model(images.to(device)).cpu()` (reference .../activation_based.py:341-358)
  * embed-all + gather     `embeds[sample_ids]`, id -1 aliases the last image (reference .../activation_based.py:385-433)
  * scores                 clarity / similarity / polysemanticity (reference semanticlens/scores.py:19-185)

Pinned: tests/test_oracle_port.py runs this port against the fixtures under tests/golden/ that
oracle/make_golden.py recorded from the imported reference (bit-exact, ids included — same torch.topk).
"""

from __future__ import annotations

from collections import Counter
from contextlib import contextmanager

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# aggregators (aggregators.py:38-244)
# ------------------------------------------------------------------------------------------------
def _need(t, nd):
    if t.ndim != nd:
        raise ValueError(f"Input tensor should be {nd}D.")


def aggregate_conv_mean(t):
    _need(t, 4)
    return t.clone().flatten(2).mean(-1).detach().cpu()


def aggregate_conv_max(t):
    _need(t, 4)
    return t.clone().flatten(2).amax(-1).detach().cpu()


def aggregate_transformer_mean(t):
    _need(t, 3)
    return t.clone().mean(1).detach().cpu()


def aggregate_transformer_absmean(t):
    _need(t, 3)
    return t.clone().abs().mean(1).detach().cpu()


def aggregate_transformer_max(t):
    _need(t, 3)
    return t.clone().amax(1).detach().cpu()


def aggregate_transformer_absmax(t):
    _need(t, 3)
    return t.clone().abs().amax(1).detach().cpu()


def get_aggregate_transformer_special_token(pos: int):
    def aggregate_transformer_special_token(t):
        _need(t, 3)
        return t.clone()[:, pos].detach().cpu()

    return aggregate_transformer_special_token


AGGREGATORS = {
    f.__name__: f
    for f in (
        aggregate_conv_mean,
        aggregate_conv_max,
        aggregate_transformer_mean,
        aggregate_transformer_absmean,
        aggregate_transformer_max,
        aggregate_transformer_absmax,
    )
}


# ------------------------------------------------------------------------------------------------
# ActMax / hooks (activation_caching.py:101-141, 288-315, 403-416)
# ------------------------------------------------------------------------------------------------
class ActMaxPort:
    def __init__(self, n_collect: int, n_latents: int | None = None):
        self.n_collect, self.n_latents = n_collect, n_latents
        self.activations = self.sample_ids = None
        if n_latents is not None:
            self._setup()

    def _setup(self):
        self.activations = -torch.zeros(self.n_latents, self.n_collect, dtype=torch.bfloat16)  # -0.0
        self.sample_ids = -torch.ones(self.n_latents, self.n_collect, dtype=torch.int64)

    def update(self, acts: torch.Tensor, sample_ids: torch.Tensor):
        assert acts.ndim == 2
        if self.activations is None:
            self.n_latents = acts.shape[1]
            self._setup()
        new_vals = acts.T.to(torch.bfloat16)
        new_ids = sample_ids.repeat(self.n_latents, 1)
        pool_vals = torch.cat([self.activations, new_vals], dim=1)
        pool_ids = torch.cat([self.sample_ids, new_ids], dim=1)
        self.activations, pick = torch.topk(pool_vals, k=self.n_collect, dim=1)
        self.sample_ids = torch.gather(pool_ids, 1, pick)


class HookSweepPort:
    """ActMaxCache restated: one ActMaxPort per layer, ids from a per-layer counter."""

    def __init__(self, layer_names, aggregation_fn, n_collect):
        self.layer_names = list(layer_names)
        self.fn = aggregation_fn
        self.state = {n: ActMaxPort(n_collect) for n in self.layer_names}
        self.counter = Counter()

    def _hook(self, name):
        def hook_fn(module, ins, outs):
            a = self.fn(outs)
            assert a.ndim == 2
            b = a.shape[0]
            ids = torch.arange(self.counter[name], self.counter[name] + b)
            self.counter[name] += b
            self.state[name].update(a, ids)

        return hook_fn

    @contextmanager
    def hooked(self, model):
        handles = [m.register_forward_hook(self._hook(n)) for n, m in model.named_modules() if n in self.layer_names]
        try:
            yield
        finally:
            for h in handles:
                h.remove()


@torch.no_grad()
def sweep(model, batches, layer_names, aggregation_fn, n_collect, device="cpu"):
    """The reference's `_run` loop over an iterable of (images, labels) batches."""
    hs = HookSweepPort(layer_names, aggregation_fn, n_collect)
    with hs.hooked(model):
        for images, _ in batches:
            model(images.to(device)).cpu()
    return hs.state


def concept_db(states: dict, embeds: torch.Tensor) -> dict:
    """`embeds[sample_ids]` per layer (python-negative index: -1 -> last image)."""
    return {name: embeds[st.sample_ids] for name, st in states.items()}


# ------------------------------------------------------------------------------------------------
# scores (scores.py:19-185)
# ------------------------------------------------------------------------------------------------
def clarity_score(V: torch.Tensor) -> torch.Tensor:
    k = V.shape[1]
    Vn = torch.nn.functional.normalize(V, dim=-1)
    s = Vn.mean(1).square().sum(-1)
    return ((s - 1.0 / k) / (k - 1)) * k


def similarity_score(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    if x.shape != y.shape:
        xn = torch.nn.functional.normalize(x, dim=-1)
        yn = torch.nn.functional.normalize(y, dim=-1)
        if x.shape[1] == y.shape[0]:
            return xn @ yn
        if x.shape[1] == y.shape[1]:
            return xn @ yn.T
        raise ValueError("x and y must have the same shape")
    return torch.nn.functional.cosine_similarity(x, y, dim=-1)


def redundancy_score(cones: torch.Tensor) -> torch.Tensor:
    c = torch.nn.functional.normalize(cones, dim=-1)
    sim = c @ c.transpose(-1, -2)
    sim = sim - 2 * torch.eye(sim.shape[-1], device=sim.device)
    return sim.max(-1).values.mean(-1)


def polysemanticity_score(V: torch.Tensor, replace_empty_clusters: bool = True) -> torch.Tensor:
    """scores.py:132-185: sklearn KMeans(2, n_init=10, random_state=123) per neuron (float64 inside sklearn),
    1 - clarity(centres); neurons whose smaller cluster has < 2 members get the mean-vs-sample fallback, which
    is evaluated in V's dtype and only then widened to float64."""
    from sklearn.cluster import KMeans

    fits = [KMeans(n_clusters=2, n_init=10, random_state=123).fit(e.detach().cpu()) for e in V]
    centres = torch.stack([torch.from_numpy(f.cluster_centers_) for f in fits])
    poly = 1 - clarity_score(centres)
    if replace_empty_clusters:
        counts = torch.zeros(len(fits), 2)
        for i, f in enumerate(fits):
            cnt = np.unique(f.labels_, return_counts=True)[1]
            if len(cnt) == 2:
                counts[i] = torch.from_numpy(cnt).float()
        small = counts.amin(-1) < 2
        sub = V[small]
        if sub.shape[0] > 0:
            n = min(10, sub.shape[1])
            acc = 0
            for j in range(n):
                acc = acc + clarity_score(torch.stack([sub.mean(1), sub[:, j]], dim=1))
            poly[small] = 1 - acc.double() / n
    return poly
