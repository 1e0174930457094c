"""ORACLE (test infrastructure, not product code) — plain torch fp32 restatement of the image tower the reference
reaches through `open_clip` (reference semanticlens/foundation_models/clip.py:53-56,103-118,137-163).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

PARITY UNPINNED by the reference itself: the arithmetic lives in the third-party dependency `open-clip-torch`
(pinned 3.0.0 in the reference's uv.lock:1568-1569, floor >=3.0.0 in pyproject.toml:20), which is neither vendored
under /root/reference nor installed in this image, and the reference's own tests at this boundary
(tests/foundation_models/test_clip.py:32-85) check shapes only, on random weights. This file restates the published
architecture of open_clip's `VisionTransformer`:

    x = conv1(img)                      kernel = stride = patch, no bias         -> (B, width, g, g) -> (B, g*g, width)
    x = cat([class_embedding, x]) + positional_embedding ; x = ln_pre(x)
    for each block:  x = x + out_proj(MHA(ln_1(x))) ; x = x + c_proj(act(c_fc(ln_2(x))))
    pooled = ln_post(x)[:, 0] ; return pooled @ proj                          (features are NOT normalised)

with nn.MultiheadAttention semantics (packed in_proj, scale = head_dim**-0.5, softmax over keys), LayerNorm eps
1e-5, act = GELU(erf) or QuickGELU (x * sigmoid(1.702 x)). tests/test_vit_oracle.py cross-checks it against the
independent implementation in HuggingFace transformers (`CLIPVisionModelWithProjection`, importable here) by weight
mapping, which is the strongest pin available offline. Weight names follow open_clip's state_dict (`visual.*`).

The eval preprocess (open_clip `image_transform`, is_train=False) for inputs that already have the model's
resolution reduces to ToTensor + Normalize(mean, std); `preprocess_u8` restates exactly that.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)


@dataclass(frozen=True)
class VitConfig:
    name: str
    image_size: int
    patch: int
    width: int
    layers: int
    heads: int
    mlp: int
    embed_dim: int
    act: str = "gelu"  # "gelu" | "quick_gelu" | "gelu_tanh"
    eps: float = 1e-5
    mean: tuple = OPENAI_MEAN
    std: tuple = OPENAI_STD

    @property
    def grid(self):
        return self.image_size // self.patch

    @property
    def tokens(self):
        return self.grid * self.grid + 1


CONFIGS = {
    "ViT-B-32": VitConfig("ViT-B-32", 224, 32, 768, 12, 12, 3072, 512),
    "ViT-B-32-quickgelu": VitConfig("ViT-B-32-quickgelu", 224, 32, 768, 12, 12, 3072, 512, act="quick_gelu"),
    "ViT-B-16": VitConfig("ViT-B-16", 224, 16, 768, 12, 12, 3072, 512),
    "ViT-L-14": VitConfig("ViT-L-14", 224, 14, 1024, 24, 16, 4096, 768),
    # small configurations for tests
    "ViT-tiny-test": VitConfig("ViT-tiny-test", 32, 8, 64, 2, 2, 128, 32),
    "ViT-small-test": VitConfig("ViT-small-test", 64, 16, 128, 3, 2, 512, 64, act="quick_gelu"),
}


def init_weights(cfg: VitConfig, seed: int = 1) -> dict[str, torch.Tensor]:
    """Random weights in open_clip's state_dict naming, drawn like open_clip initialises them (scale = width**-0.5
    for embeddings/proj, width**-0.5 attention, (2*width)**-0.5 fc, depth-scaled output projections)."""
    g = torch.Generator().manual_seed(seed)
    W, L = cfg.width, cfg.layers
    scale = W**-0.5
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    sd = {
        "visual.conv1.weight": rn(W, 3, cfg.patch, cfg.patch) * (3 * cfg.patch * cfg.patch) ** -0.5,
        "visual.class_embedding": rn(W) * scale,
        "visual.positional_embedding": rn(cfg.tokens, W) * scale,
        "visual.ln_pre.weight": 1 + 0.1 * rn(W),
        "visual.ln_pre.bias": 0.1 * rn(W),
        "visual.ln_post.weight": 1 + 0.1 * rn(W),
        "visual.ln_post.bias": 0.1 * rn(W),
        "visual.proj": rn(W, cfg.embed_dim) * scale,
    }
    proj_std = scale * (2 * L) ** -0.5
    for i in range(L):
        p = f"visual.transformer.resblocks.{i}."
        sd[p + "ln_1.weight"] = 1 + 0.1 * rn(W)
        sd[p + "ln_1.bias"] = 0.1 * rn(W)
        sd[p + "attn.in_proj_weight"] = rn(3 * W, W) * scale
        sd[p + "attn.in_proj_bias"] = 0.02 * rn(3 * W)
        sd[p + "attn.out_proj.weight"] = rn(W, W) * proj_std
        sd[p + "attn.out_proj.bias"] = 0.02 * rn(W)
        sd[p + "ln_2.weight"] = 1 + 0.1 * rn(W)
        sd[p + "ln_2.bias"] = 0.1 * rn(W)
        sd[p + "mlp.c_fc.weight"] = rn(cfg.mlp, W) * (2 * W) ** -0.5
        sd[p + "mlp.c_fc.bias"] = 0.02 * rn(cfg.mlp)
        sd[p + "mlp.c_proj.weight"] = rn(W, cfg.mlp) * proj_std
        sd[p + "mlp.c_proj.bias"] = 0.02 * rn(W)
    return sd


def _act(x, kind):
    if kind == "gelu":
        return F.gelu(x)
    if kind == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    if kind == "gelu_tanh":
        return F.gelu(x, approximate="tanh")
    raise ValueError(kind)


def mha(x, w_in, b_in, w_out, b_out, heads):
    """nn.MultiheadAttention(batch_first) self-attention forward, need_weights=False, no mask, no dropout."""
    B, T, W = x.shape
    dh = W // heads
    qkv = F.linear(x, w_in, b_in)
    q, k, v = qkv.split(W, dim=-1)
    q = q.view(B, T, heads, dh).transpose(1, 2)
    k = k.view(B, T, heads, dh).transpose(1, 2)
    v = v.view(B, T, heads, dh).transpose(1, 2)
    att = torch.softmax((q * dh**-0.5) @ k.transpose(-1, -2), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, T, W)
    return F.linear(o, w_out, b_out)


@torch.no_grad()
def encode_image(sd: dict, cfg: VitConfig, img: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """(B, 3, S, S) preprocessed images -> (B, embed_dim) un-normalised features. dtype=float64 gives the
    high-precision ground truth used to score both the fp32 oracle and the kernels."""
    w = {k: v.to(device=img.device, dtype=dtype) for k, v in sd.items()}
    x = F.conv2d(img.to(dtype), w["visual.conv1.weight"], stride=cfg.patch)
    B = x.shape[0]
    x = x.reshape(B, cfg.width, -1).permute(0, 2, 1)
    cls = w["visual.class_embedding"].expand(B, 1, -1)
    x = torch.cat([cls, x], dim=1) + w["visual.positional_embedding"]
    x = F.layer_norm(x, (cfg.width,), w["visual.ln_pre.weight"], w["visual.ln_pre.bias"], cfg.eps)
    for i in range(cfg.layers):
        p = f"visual.transformer.resblocks.{i}."
        h = F.layer_norm(x, (cfg.width,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], cfg.eps)
        x = x + mha(h, w[p + "attn.in_proj_weight"], w[p + "attn.in_proj_bias"], w[p + "attn.out_proj.weight"],
                    w[p + "attn.out_proj.bias"], cfg.heads)
        h = F.layer_norm(x, (cfg.width,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], cfg.eps)
        h = _act(F.linear(h, w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"]), cfg.act)
        x = x + F.linear(h, w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"])
    pooled = F.layer_norm(x, (cfg.width,), w["visual.ln_post.weight"], w["visual.ln_post.bias"], cfg.eps)[:, 0]
    return pooled @ w["visual.proj"]


def preprocess_u8(cfg: VitConfig, u8: torch.Tensor) -> torch.Tensor:
    """ToTensor + Normalize on (B, 3, S, S) uint8 images that already have the model's resolution."""
    mean = torch.tensor(cfg.mean, device=u8.device).view(1, 3, 1, 1)
    std = torch.tensor(cfg.std, device=u8.device).view(1, 3, 1, 1)
    return (u8.to(torch.float32).div(255) - mean) / std


class OracleTower:
    """Callable bundle used by bench.py's CPU arm: preprocess_u8 + encode_image on the host."""

    def __init__(self, cfg: VitConfig, sd: dict):
        self.cfg, self.sd = cfg, sd

    def preprocess_u8(self, u8):
        return preprocess_u8(self.cfg, u8)

    def encode_image(self, img):
        return encode_image(self.sd, self.cfg, img)


def build(name: str, seed: int = 1) -> OracleTower:
    cfg = CONFIGS[name]
    return OracleTower(cfg, init_weights(cfg, seed))


def flops_per_image(cfg: VitConfig) -> float:
    """2*MAC count of the GEMMs and attention."""
    T, W = cfg.tokens, cfg.width
    per_layer = 2 * T * W * (3 * W + W + 2 * cfg.mlp) + 2 * 2 * T * T * W
    return (T - 1) * 2 * W * 3 * cfg.patch**2 + cfg.layers * per_layer + 2 * W * cfg.embed_dim


assert math.isclose(flops_per_image(CONFIGS["ViT-B-32"]) / 1e9, 8.8, rel_tol=0.02)
