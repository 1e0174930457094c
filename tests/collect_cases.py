"""Seeded generators shared by oracle/make_golden.py (which records the reference's outputs for them) and the tests
(which regenerate the same inputs): only outputs and an input checksum live in tests/golden/."""

import numpy as np
import torch

LARGE = dict(k=20, batch=256, C=2048, H=7, W=7, n_batches=3, seed=20261017)


def large_maps():
    """3 batches of (256, 2048, 7, 7) fp32 post-ReLU-like maps: per-channel gain and offset so that channels differ in
    scale, some are dead (all zeros -> placeholders survive) and bf16 ties in the aggregates are common."""
    c = LARGE
    rng = np.random.default_rng(c["seed"])
    gain = rng.uniform(0.05, 4.0, size=(1, c["C"], 1, 1)).astype(np.float32)
    offset = rng.uniform(-2.5, 0.5, size=(1, c["C"], 1, 1)).astype(np.float32)
    out = []
    for _ in range(c["n_batches"]):
        x = rng.standard_normal((c["batch"], c["C"], c["H"], c["W"]), dtype=np.float32)
        out.append(np.maximum(x * gain + offset, 0.0).astype(np.float32))
    return out


class FakeTextFM:
    """Deterministic stand-in for a foundation model's text side: tokens = byte codes, embedding = sum of table rows
    weighted by position. Pure torch on the CPU, so the reference's and this repo's ``_embed_text_probes`` can both
    drive it."""

    device = "cpu"

    def __init__(self, dim: int = 16, context: int = 24):
        rng = np.random.default_rng(5)
        self.table = torch.from_numpy(rng.standard_normal((256, dim)).astype(np.float32))
        self.context = context

    def tokenize(self, texts):
        texts = [texts] if isinstance(texts, str) else list(texts)
        ids = torch.zeros((len(texts), self.context), dtype=torch.int64)
        for i, t in enumerate(texts):
            b = list(t.encode("utf-8"))[: self.context]
            ids[i, : len(b)] = torch.tensor(b, dtype=torch.int64) if b else ids[i, :0]
        return ids

    def encode_text(self, ids):
        w = torch.arange(1, ids.shape[1] + 1, dtype=torch.float32).view(1, -1, 1) / ids.shape[1]
        return (self.table[ids] * w * (ids > 0).unsqueeze(-1)).sum(1)

    def to(self, device):
        return self


# name -> (queries, templates, batch_size)
TEXT_PROBE = {
    "plain": (("a dog", "striped fur", "sky"), None, None),
    "one_template": (("a dog", "striped fur", "sky"), ("a photo of {}",), None),
    "one_query": (("wheel",), ("a photo of {}", "an image of a {}", "{} texture"), 2),
    "mixed": (("a dog", "sky"), ("a photo of {}", "{} texture", "close-up of {}"), 4),
}
