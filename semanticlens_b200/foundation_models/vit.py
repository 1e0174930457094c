"""Host side of the B200 ViT image tower: weights in open_clip's ``visual.*`` naming -> split planes + a
``SlbVitWeights`` struct -> ``slb_vit_forward`` (one C-ABI call per batch, no host sync, caller-owned workspace)."""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from .. import _native as N
from .. import ops

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
    def tokens(self) -> int:
        return (self.image_size // self.patch) ** 2 + 1


# open_clip model configs (model_configs/*.json of open-clip-torch 3.0.0), image tower only
CONFIGS = {
    "ViT-B-32": VitConfig("ViT-B-32", 224, 32, 768, 12, 12, 3072, 512),
    "ViT-B-32-quickgelu": VitConfig("ViT-B-32-quickgelu", 224, 32, 768, 12, 12, 3072, 512, act="quick_gelu"),
    "ViT-B-16": VitConfig("ViT-B-16", 224, 16, 768, 12, 12, 3072, 512),
    "ViT-B-16-quickgelu": VitConfig("ViT-B-16-quickgelu", 224, 16, 768, 12, 12, 3072, 512, act="quick_gelu"),
    "ViT-L-14": VitConfig("ViT-L-14", 224, 14, 1024, 24, 16, 4096, 768),
    "ViT-L-14-quickgelu": VitConfig("ViT-L-14-quickgelu", 224, 14, 1024, 24, 16, 4096, 768, act="quick_gelu"),
}

_ACT = {"gelu": N.EPI_GELU_ERF, "quick_gelu": N.EPI_QUICKGELU, "gelu_tanh": N.EPI_GELU_TANH}


def random_state_dict(cfg: VitConfig, seed: int = 1) -> dict[str, torch.Tensor]:
    """Random ``visual.*`` weights (no network here, so no pretrained download): embeddings and proj at
    width**-0.5, attention at width**-0.5, MLP at (2*width)**-0.5, output projections depth-scaled."""
    g = torch.Generator().manual_seed(seed)
    W, L = cfg.width, cfg.layers
    scale = W**-0.5

    def rn(*s):
        return torch.randn(*s, generator=g)

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


class VitTower:
    """Device-resident tower: fp32 vectors, split-plane matrices, cached workspace."""

    def __init__(self, cfg: VitConfig, state_dict: dict[str, torch.Tensor], device, plane_format: int = N.PLANE_F16):
        self.cfg = cfg
        self.plane_format = plane_format
        self.state_dict = {k: v.detach().to(torch.float32).cpu() for k, v in state_dict.items() if k.startswith("visual.")}
        missing = [k for k in random_state_dict_keys(cfg) if k not in self.state_dict]
        if missing:
            raise KeyError(f"state dict is missing {len(missing)} image-tower tensors, e.g. {missing[:3]}")
        self._device = torch.device("cpu")
        self._struct = None
        self._keep: list = []
        self._ws: torch.Tensor | None = None
        self.to(device)

    # -- placement ----------------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self._device

    def to(self, device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device != self._device or (device.type == "cuda" and self._struct is None):
            self._device = device
            self._struct, self._keep, self._ws = None, [], None
            if device.type == "cuda":
                self._upload()
        return self

    def _upload(self):
        N.load(require_device=True)
        cfg, sd, dev, fmt = self.cfg, self.state_dict, self._device, self.plane_format
        keep = self._keep

        def vec(name):
            t = sd[name].to(dev).contiguous()
            keep.append(t)
            return t.data_ptr()

        def planes(mat: torch.Tensor):
            t = ops.split_planes(mat.to(dev), fmt)
            keep.append(t)
            return t.data_ptr()

        kc = 3 * cfg.patch * cfg.patch
        kpad = N.load().slb_patch_k(cfg.patch)
        conv = torch.zeros(cfg.width, kpad)
        conv[:, :kc] = sd["visual.conv1.weight"].reshape(cfg.width, kc)
        layers = (N.SlbVitLayer * cfg.layers)()
        for i in range(cfg.layers):
            p = f"visual.transformer.resblocks.{i}."
            ly = layers[i]
            ly.ln1_g, ly.ln1_b = vec(p + "ln_1.weight"), vec(p + "ln_1.bias")
            ly.w_qkv, ly.b_qkv = planes(sd[p + "attn.in_proj_weight"]), vec(p + "attn.in_proj_bias")
            ly.w_out, ly.b_out = planes(sd[p + "attn.out_proj.weight"]), vec(p + "attn.out_proj.bias")
            ly.ln2_g, ly.ln2_b = vec(p + "ln_2.weight"), vec(p + "ln_2.bias")
            ly.w_fc, ly.b_fc = planes(sd[p + "mlp.c_fc.weight"]), vec(p + "mlp.c_fc.bias")
            ly.w_proj, ly.b_proj = planes(sd[p + "mlp.c_proj.weight"]), vec(p + "mlp.c_proj.bias")
        w = N.SlbVitWeights()
        w.image_size, w.patch, w.width, w.layers = cfg.image_size, cfg.patch, cfg.width, cfg.layers
        w.heads, w.mlp, w.embed_dim = cfg.heads, cfg.mlp, cfg.embed_dim
        w.act, w.plane_fmt, w.has_cls, w.pool, w.ln_eps = _ACT[cfg.act], fmt, 1, 0, cfg.eps
        w.conv_w, w.conv_b = planes(conv), None
        w.cls, w.pos = vec("visual.class_embedding"), vec("visual.positional_embedding")
        w.ln_pre_g, w.ln_pre_b = vec("visual.ln_pre.weight"), vec("visual.ln_pre.bias")
        w.ln_post_g, w.ln_post_b = vec("visual.ln_post.weight"), vec("visual.ln_post.bias")
        w.proj = planes(sd["visual.proj"].T.contiguous())
        w.layer = layers
        keep.append(layers)
        self._struct = w

    # -- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, img: torch.Tensor) -> torch.Tensor:
        """(B, 3, S, S) fp32 CUDA -> (B, embed_dim) fp32 CUDA, enqueued on the current stream."""
        lib = N.load(require_device=True)
        if self._struct is None:
            raise N.SlbError("the image tower lives on the CPU: call .to('cuda') first (there is no CPU fallback)")
        N.require_cuda(img, "images")
        cfg = self.cfg
        if img.ndim != 4 or tuple(img.shape[1:]) != (3, cfg.image_size, cfg.image_size):
            raise ValueError(f"expected (B, 3, {cfg.image_size}, {cfg.image_size}) images, got {tuple(img.shape)}")
        img = img.detach().to(torch.float32).contiguous()
        B = img.shape[0]
        out = torch.empty((B, cfg.embed_dim), dtype=torch.float32, device=img.device)
        if B == 0:
            return out
        need = lib.slb_vit_workspace_bytes(ctypes.byref(self._struct), B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != img.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=img.device)
        tm = ops._timer.begin() if ops._timer else None
        with torch.cuda.device(img.device):
            rc = lib.slb_vit_forward(
                ctypes.byref(self._struct), img.data_ptr(), B, out.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                N.stream_ptr(img.device),
            )
        if ops._timer:
            ops._timer.end("K4 vit_forward", tm, 0, int(3 * flops_per_image(cfg) * B))
        N.check(rc, "slb_vit_forward")
        return out


def random_state_dict_keys(cfg: VitConfig) -> list[str]:
    keys = ["visual.conv1.weight", "visual.class_embedding", "visual.positional_embedding", "visual.ln_pre.weight",
            "visual.ln_pre.bias", "visual.ln_post.weight", "visual.ln_post.bias", "visual.proj"]
    for i in range(cfg.layers):
        p = f"visual.transformer.resblocks.{i}."
        keys += [p + s for s in ("ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias",
                                 "attn.out_proj.weight", "attn.out_proj.bias", "ln_2.weight", "ln_2.bias",
                                 "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias")]
    return keys


def flops_per_image(cfg: VitConfig) -> float:
    """2*MAC of the GEMMs and attention of one image (algorithmic, single pass)."""
    T, W = cfg.tokens, cfg.width
    per_layer = 2 * T * W * (3 * W + W + 2 * cfg.mlp) + 2 * 2 * T * T * W
    return (T - 1) * 2 * W * 3 * cfg.patch**2 + cfg.layers * per_layer + 2 * W * cfg.embed_dim
