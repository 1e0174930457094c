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


# ------------------------------------------------------------------------------------------------
# SigLIP image tower (reference: foundation_models/clip.py:190-215 `SigLipV2` = open_clip "hf-hub:timm/ViT-B-16-SigLIP2",
# i.e. a timm VisionTransformer behind open_clip's TimmModel wrapper; timm 1.0.19 / open-clip-torch 3.0.0, neither
# installed). Restated from the published architecture: biased patch conv, NO class token, learned positional embedding,
# no pre-norm, pre-LN blocks (eps 1e-6, GELU-tanh), final LayerNorm over all tokens, attention-pool ("MAP") head:
# a learned latent query attends over the tokens (nn.MultiheadAttention arithmetic), out-projection, then
# x + mlp(norm(x)); the pooled vector is the feature (no extra projection). Weight names follow the open_clip state
# dict of such a model (`visual.trunk.*`). Pinned against HF transformers' SiglipVisionModel (tests/test_vit_oracle.py).
# ------------------------------------------------------------------------------------------------
SIGLIP_MEAN = (0.5, 0.5, 0.5)
SIGLIP_STD = (0.5, 0.5, 0.5)


@dataclass(frozen=True)
class SigLipConfig:
    name: str
    image_size: int
    patch: int
    width: int
    layers: int
    heads: int
    mlp: int
    act: str = "gelu_tanh"
    eps: float = 1e-6
    mean: tuple = SIGLIP_MEAN
    std: tuple = SIGLIP_STD

    @property
    def tokens(self):
        return (self.image_size // self.patch) ** 2

    @property
    def embed_dim(self):
        return self.width


SIGLIP_CONFIGS = {
    "ViT-B-16-SigLIP2": SigLipConfig("ViT-B-16-SigLIP2", 224, 16, 768, 12, 12, 3072),
    "ViT-L-16-SigLIP-256": SigLipConfig("ViT-L-16-SigLIP-256", 256, 16, 1024, 24, 16, 4096),
    "SigLIP-tiny-test": SigLipConfig("SigLIP-tiny-test", 32, 8, 128, 2, 2, 256),
}


def init_siglip_weights(cfg: SigLipConfig, seed: int = 1) -> dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    W, L = cfg.width, cfg.layers
    scale = W**-0.5
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    t = "visual.trunk."
    sd = {
        t + "patch_embed.proj.weight": rn(W, 3, cfg.patch, cfg.patch) * (3 * cfg.patch * cfg.patch) ** -0.5,
        t + "patch_embed.proj.bias": 0.02 * rn(W),
        t + "pos_embed": rn(1, cfg.tokens, W) * scale,
        t + "norm.weight": 1 + 0.1 * rn(W),
        t + "norm.bias": 0.1 * rn(W),
        t + "attn_pool.latent": rn(1, 1, W) * scale,
        t + "attn_pool.q.weight": rn(W, W) * scale,
        t + "attn_pool.q.bias": 0.02 * rn(W),
        t + "attn_pool.kv.weight": rn(2 * W, W) * scale,
        t + "attn_pool.kv.bias": 0.02 * rn(2 * W),
        t + "attn_pool.proj.weight": rn(W, W) * scale,
        t + "attn_pool.proj.bias": 0.02 * rn(W),
        t + "attn_pool.norm.weight": 1 + 0.1 * rn(W),
        t + "attn_pool.norm.bias": 0.1 * rn(W),
        t + "attn_pool.mlp.fc1.weight": rn(cfg.mlp, W) * (2 * W) ** -0.5,
        t + "attn_pool.mlp.fc1.bias": 0.02 * rn(cfg.mlp),
        t + "attn_pool.mlp.fc2.weight": rn(W, cfg.mlp) * scale * 0.5,
        t + "attn_pool.mlp.fc2.bias": 0.02 * rn(W),
    }
    proj_std = scale * (2 * L) ** -0.5
    for i in range(L):
        p = f"{t}blocks.{i}."
        sd[p + "norm1.weight"] = 1 + 0.1 * rn(W)
        sd[p + "norm1.bias"] = 0.1 * rn(W)
        sd[p + "attn.qkv.weight"] = rn(3 * W, W) * scale
        sd[p + "attn.qkv.bias"] = 0.02 * rn(3 * W)
        sd[p + "attn.proj.weight"] = rn(W, W) * proj_std
        sd[p + "attn.proj.bias"] = 0.02 * rn(W)
        sd[p + "norm2.weight"] = 1 + 0.1 * rn(W)
        sd[p + "norm2.bias"] = 0.1 * rn(W)
        sd[p + "mlp.fc1.weight"] = rn(cfg.mlp, W) * (2 * W) ** -0.5
        sd[p + "mlp.fc1.bias"] = 0.02 * rn(cfg.mlp)
        sd[p + "mlp.fc2.weight"] = rn(W, cfg.mlp) * proj_std
        sd[p + "mlp.fc2.bias"] = 0.02 * rn(W)
    return sd


@torch.no_grad()
def encode_image_siglip(sd: dict, cfg: SigLipConfig, img: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """(B, 3, S, S) preprocessed images -> (B, width) features of the SigLIP tower."""
    w = {k: v.to(device=img.device, dtype=dtype) for k, v in sd.items()}
    t = "visual.trunk."
    W, H = cfg.width, cfg.heads
    dh = W // H
    x = F.conv2d(img.to(dtype), w[t + "patch_embed.proj.weight"], w[t + "patch_embed.proj.bias"], stride=cfg.patch)
    B = x.shape[0]
    x = x.reshape(B, W, -1).permute(0, 2, 1) + w[t + "pos_embed"]
    for i in range(cfg.layers):
        p = f"{t}blocks.{i}."
        h = F.layer_norm(x, (W,), w[p + "norm1.weight"], w[p + "norm1.bias"], cfg.eps)
        x = x + mha(h, w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"], w[p + "attn.proj.weight"], w[p + "attn.proj.bias"], H)
        h = F.layer_norm(x, (W,), w[p + "norm2.weight"], w[p + "norm2.bias"], cfg.eps)
        h = _act(F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"]), cfg.act)
        x = x + F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
    x = F.layer_norm(x, (W,), w[t + "norm.weight"], w[t + "norm.bias"], cfg.eps)
    # attention pool: one latent query per image
    a = t + "attn_pool."
    q = F.linear(w[a + "latent"].expand(B, 1, W), w[a + "q.weight"], w[a + "q.bias"]).view(B, 1, H, dh).transpose(1, 2)
    kv = F.linear(x, w[a + "kv.weight"], w[a + "kv.bias"])
    k, v = kv.split(W, dim=-1)
    k = k.view(B, -1, H, dh).transpose(1, 2)
    v = v.view(B, -1, H, dh).transpose(1, 2)
    att = torch.softmax((q * dh**-0.5) @ k.transpose(-1, -2), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, 1, W)
    o = F.linear(o, w[a + "proj.weight"], w[a + "proj.bias"])
    h = F.layer_norm(o, (W,), w[a + "norm.weight"], w[a + "norm.bias"], cfg.eps)
    h = _act(F.linear(h, w[a + "mlp.fc1.weight"], w[a + "mlp.fc1.bias"]), cfg.act)
    o = o + F.linear(h, w[a + "mlp.fc2.weight"], w[a + "mlp.fc2.bias"])
    return o[:, 0]



# ------------------------------------------------------------------------------------------------
# CLIP text tower (reference: foundation_models/clip.py:120-135 -> open_clip CLIP.encode_text, third-party): token
# embedding + positional embedding, pre-LN blocks with a CAUSAL attention mask, ln_final, the feature of the end-of-text
# token (the position of the largest token id) times text_projection; un-normalised. Pinned against HF transformers'
# CLIPTextModelWithProjection (tests/test_vit_oracle.py).
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class TextConfig:
    name: str
    context: int
    vocab: int
    width: int
    layers: int
    heads: int
    embed_dim: int
    act: str = "gelu"
    eps: float = 1e-5
    # "clip": causal mask, end-of-text (argmax) pooling, bias-free text_projection matrix (W, D).
    # "siglip": open_clip's TextTransformer as the SigLIP configs set it up (no_causal_mask, pool_type "last",
    # proj_bias: text_projection is an nn.Linear) — what the reference reaches through SigLipV2.encode_text
    # (foundation_models/clip.py:190-215). Pinned against HF transformers' SiglipTextModel (tests/test_vit_oracle.py).
    arch: str = "clip"

    @property
    def mlp(self):
        return 4 * self.width


TEXT_CONFIGS = {
    "ViT-B-32": TextConfig("ViT-B-32", 77, 49408, 512, 12, 8, 512),
    "ViT-L-14": TextConfig("ViT-L-14", 77, 49408, 768, 12, 12, 768),
    "text-tiny-test": TextConfig("text-tiny-test", 12, 100, 128, 2, 2, 32),
    "ViT-B-16-SigLIP": TextConfig("ViT-B-16-SigLIP", 64, 32000, 768, 12, 12, 768, act="gelu_tanh", eps=1e-6, arch="siglip"),
    "siglip-text-tiny-test": TextConfig("siglip-text-tiny-test", 16, 100, 128, 2, 2, 64, act="gelu_tanh", eps=1e-6, arch="siglip"),
}


def init_text_weights(cfg: TextConfig, seed: int = 1) -> dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    W, L = cfg.width, cfg.layers
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    sd = {
        "token_embedding.weight": rn(cfg.vocab, W) * 0.02,
        "positional_embedding": rn(cfg.context, W) * 0.01,
        "ln_final.weight": 1 + 0.1 * rn(W),
        "ln_final.bias": 0.1 * rn(W),
        "text_projection": rn(W, cfg.embed_dim) * W**-0.5,
    }
    if cfg.arch == "siglip":
        del sd["text_projection"]
        sd["text_projection.weight"] = rn(cfg.embed_dim, W) * W**-0.5
        sd["text_projection.bias"] = 0.02 * rn(cfg.embed_dim)
    proj_std = W**-0.5 * (2 * L) ** -0.5
    for i in range(L):
        p = f"transformer.resblocks.{i}."
        sd[p + "ln_1.weight"] = 1 + 0.1 * rn(W)
        sd[p + "ln_1.bias"] = 0.1 * rn(W)
        sd[p + "attn.in_proj_weight"] = rn(3 * W, W) * W**-0.5
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


@torch.no_grad()
def encode_text(sd: dict, cfg: TextConfig, tokens: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """(B, context) int64 token ids -> (B, embed_dim) un-normalised text features."""
    w = {k: v.to(device=tokens.device, dtype=dtype) for k, v in sd.items()}
    B, T = tokens.shape
    W, H = cfg.width, cfg.heads
    dh = W // H
    x = w["token_embedding.weight"][tokens] + w["positional_embedding"][:T]
    mask = torch.full((T, T), float("-inf"), dtype=dtype, device=tokens.device).triu(1)
    if cfg.arch == "siglip":
        mask = torch.zeros_like(mask)
    for i in range(cfg.layers):
        p = f"transformer.resblocks.{i}."
        h = F.layer_norm(x, (W,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], cfg.eps)
        qkv = F.linear(h, w[p + "attn.in_proj_weight"], w[p + "attn.in_proj_bias"])
        q, k, v = (t.view(B, T, H, dh).transpose(1, 2) for t in qkv.split(W, dim=-1))
        att = torch.softmax((q * dh**-0.5) @ k.transpose(-1, -2) + mask, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, T, W)
        x = x + F.linear(o, w[p + "attn.out_proj.weight"], w[p + "attn.out_proj.bias"])
        h = F.layer_norm(x, (W,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], cfg.eps)
        h = _act(F.linear(h, w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"]), cfg.act)
        x = x + F.linear(h, w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"])
    x = F.layer_norm(x, (W,), w["ln_final.weight"], w["ln_final.bias"], cfg.eps)
    if cfg.arch == "siglip":
        return F.linear(x[:, -1], w["text_projection.weight"], w["text_projection.bias"])
    return x[torch.arange(B, device=tokens.device), tokens.argmax(dim=-1)] @ w["text_projection"]
