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
    arch: str = "clip"  # "clip": open_clip VisionTransformer; "siglip": timm ViT with the attention-pool head

    @property
    def has_cls(self) -> bool:
        return self.arch == "clip"

    @property
    def resize_mode(self) -> str:
        """Eval-transform geometry of open_clip's preprocess_cfg: CLIP resizes the shorter side and centre-crops,
        SigLIP squashes the whole image to (S, S)."""
        return "squash" if self.arch == "siglip" else "shortest"

    @property
    def tokens(self) -> int:
        return (self.image_size // self.patch) ** 2 + (1 if self.has_cls else 0)


# open_clip model configs (model_configs/*.json of open-clip-torch 3.0.0), image tower only
CONFIGS = {
    "ViT-B-32": VitConfig("ViT-B-32", 224, 32, 768, 12, 12, 3072, 512),
    "ViT-B-32-quickgelu": VitConfig("ViT-B-32-quickgelu", 224, 32, 768, 12, 12, 3072, 512, act="quick_gelu"),
    "ViT-B-16": VitConfig("ViT-B-16", 224, 16, 768, 12, 12, 3072, 512),
    "ViT-B-16-quickgelu": VitConfig("ViT-B-16-quickgelu", 224, 16, 768, 12, 12, 3072, 512, act="quick_gelu"),
    "ViT-L-14": VitConfig("ViT-L-14", 224, 14, 1024, 24, 16, 4096, 768),
    "ViT-L-14-quickgelu": VitConfig("ViT-L-14-quickgelu", 224, 14, 1024, 24, 16, 4096, 768, act="quick_gelu"),
}

SIGLIP_MEAN = (0.5, 0.5, 0.5)
SIGLIP_STD = (0.5, 0.5, 0.5)


def _siglip(name, size, patch, width, layers, heads, mlp):
    return VitConfig(name, size, patch, width, layers, heads, mlp, width, act="gelu_tanh", eps=1e-6, mean=SIGLIP_MEAN,
                     std=SIGLIP_STD, arch="siglip")


# SigLIP image towers (open_clip wraps timm ViTs: biased patch conv, no class token, attention-pool head, no projection)
CONFIGS.update({
    "ViT-B-16-SigLIP2": _siglip("ViT-B-16-SigLIP2", 224, 16, 768, 12, 12, 3072),
    "hf-hub:timm/ViT-B-16-SigLIP2": _siglip("ViT-B-16-SigLIP2", 224, 16, 768, 12, 12, 3072),
    "ViT-B-16-SigLIP": _siglip("ViT-B-16-SigLIP", 224, 16, 768, 12, 12, 3072),
    "ViT-L-16-SigLIP-256": _siglip("ViT-L-16-SigLIP-256", 256, 16, 1024, 24, 16, 4096),
})

_ACT = {"gelu": N.EPI_GELU_ERF, "quick_gelu": N.EPI_QUICKGELU, "gelu_tanh": N.EPI_GELU_TANH}


def random_state_dict(cfg: VitConfig, seed: int = 1) -> dict[str, torch.Tensor]:
    if cfg.arch == "siglip":
        return _random_siglip_state_dict(cfg, seed)
    return _random_clip_state_dict(cfg, seed)


def _random_siglip_state_dict(cfg: VitConfig, seed: int) -> dict[str, torch.Tensor]:
    """Random weights in the naming of an open_clip SigLIP checkpoint (``visual.trunk.*`` = the timm ViT)."""
    g = torch.Generator().manual_seed(seed)
    W, L = cfg.width, cfg.layers
    scale = W**-0.5

    def rn(*s):
        return torch.randn(*s, generator=g)

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


def _random_clip_state_dict(cfg: VitConfig, seed: int = 1) -> dict[str, torch.Tensor]:
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
        want = expected_shapes(cfg)
        missing = [k for k in want if k not in self.state_dict]
        if missing:
            raise KeyError(f"state dict is missing {len(missing)} image-tower tensors, e.g. {missing[:3]}")
        # a checkpoint of another architecture must fail here, not as an out-of-bounds read inside a GEMM
        wrong = [(k, tuple(self.state_dict[k].shape), shp) for k, shp in want.items() if tuple(self.state_dict[k].shape) != shp]
        if wrong:
            k, got, shp = wrong[0]
            raise ValueError(f"state dict does not fit '{cfg.name}': {len(wrong)} tensor(s) have the wrong shape, e.g. "
                             f"{k} is {got}, expected {shp}")
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
            t = ops.split_planes(mat.to(dev), fmt, N.WEIGHT_PLANE_SCALE)
            keep.append(t)
            return t.data_ptr()

        kc = 3 * cfg.patch * cfg.patch
        kpad = N.load().slb_patch_k(cfg.patch)
        siglip = cfg.arch == "siglip"
        t = "visual.trunk." if siglip else "visual."
        # per-block tensor names: (ln1, qkv, out, ln2, fc, proj) in open_clip's and in timm's naming
        names = (("norm1", "attn.qkv.weight", "attn.qkv.bias", "attn.proj", "norm2", "mlp.fc1", "mlp.fc2") if siglip else
                 ("ln_1", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj", "ln_2", "mlp.c_fc", "mlp.c_proj"))
        conv = torch.zeros(cfg.width, kpad)
        conv[:, :kc] = sd[t + ("patch_embed.proj.weight" if siglip else "conv1.weight")].reshape(cfg.width, kc)
        layers = (N.SlbVitLayer * cfg.layers)()
        for i in range(cfg.layers):
            p = f"{t}blocks.{i}." if siglip else f"{t}transformer.resblocks.{i}."
            ly = layers[i]
            ln1, wqkv, bqkv, out, ln2, fc, proj = names
            ly.ln1_g, ly.ln1_b = vec(p + ln1 + ".weight"), vec(p + ln1 + ".bias")
            ly.w_qkv, ly.b_qkv = planes(sd[p + wqkv]), vec(p + bqkv)
            ly.w_out, ly.b_out = planes(sd[p + out + ".weight"]), vec(p + out + ".bias")
            ly.ln2_g, ly.ln2_b = vec(p + ln2 + ".weight"), vec(p + ln2 + ".bias")
            ly.w_fc, ly.b_fc = planes(sd[p + fc + ".weight"]), vec(p + fc + ".bias")
            ly.w_proj, ly.b_proj = planes(sd[p + proj + ".weight"]), vec(p + proj + ".bias")
        w = N.SlbVitWeights()
        w.image_size, w.patch, w.width, w.layers = cfg.image_size, cfg.patch, cfg.width, cfg.layers
        w.heads, w.mlp, w.embed_dim = cfg.heads, cfg.mlp, cfg.embed_dim
        w.act, w.plane_fmt, w.ln_eps = _ACT[cfg.act], fmt, cfg.eps
        w.conv_w = planes(conv)
        if siglip:
            w.has_cls, w.pool = 0, N.POOL_MAP
            w.conv_b = vec(t + "patch_embed.proj.bias")
            w.cls = None
            pos = sd[t + "pos_embed"].reshape(cfg.tokens, cfg.width).to(dev).contiguous()
            keep.append(pos)
            w.pos = pos.data_ptr()
            w.ln_pre_g = w.ln_pre_b = None
            w.ln_post_g, w.ln_post_b = vec(t + "norm.weight"), vec(t + "norm.bias")
            w.proj = None
            a = t + "attn_pool."
            # the latent's q projection does not depend on the image: one tiny host-side matvec at load time
            mq = (sd[a + "latent"].reshape(1, cfg.width).double() @ sd[a + "q.weight"].double().T + sd[a + "q.bias"].double())
            mq = mq.float().reshape(cfg.width).to(dev).contiguous()
            keep.append(mq)
            w.map_q = mq.data_ptr()
            w.map_w_kv, w.map_b_kv = planes(sd[a + "kv.weight"]), vec(a + "kv.bias")
            w.map_w_out, w.map_b_out = planes(sd[a + "proj.weight"]), vec(a + "proj.bias")
            w.map_ln_g, w.map_ln_b = vec(a + "norm.weight"), vec(a + "norm.bias")
            w.map_w_fc, w.map_b_fc = planes(sd[a + "mlp.fc1.weight"]), vec(a + "mlp.fc1.bias")
            w.map_w_proj, w.map_b_proj = planes(sd[a + "mlp.fc2.weight"]), vec(a + "mlp.fc2.bias")
        else:
            w.has_cls, w.pool = 1, N.POOL_CLS
            w.conv_b = None
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
        with torch.cuda.device(img.device):
            rc = lib.slb_vit_forward(
                ctypes.byref(self._struct), img.data_ptr(), B, out.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                N.stream_ptr(img.device),
            )
        N.check(rc, "slb_vit_forward")
        return out


def expected_shapes(cfg: VitConfig) -> dict[str, tuple]:
    """Name -> shape of every image-tower tensor of an open_clip checkpoint for ``cfg`` (what ``_upload`` indexes)."""
    W, M, P, T, L = cfg.width, cfg.mlp, cfg.patch, cfg.tokens, cfg.layers
    if cfg.arch == "siglip":
        t = "visual.trunk."
        shapes = {
            t + "patch_embed.proj.weight": (W, 3, P, P), t + "patch_embed.proj.bias": (W,), t + "pos_embed": (1, T, W),
            t + "norm.weight": (W,), t + "norm.bias": (W,), t + "attn_pool.latent": (1, 1, W),
            t + "attn_pool.q.weight": (W, W), t + "attn_pool.q.bias": (W,),
            t + "attn_pool.kv.weight": (2 * W, W), t + "attn_pool.kv.bias": (2 * W,),
            t + "attn_pool.proj.weight": (W, W), t + "attn_pool.proj.bias": (W,),
            t + "attn_pool.norm.weight": (W,), t + "attn_pool.norm.bias": (W,),
            t + "attn_pool.mlp.fc1.weight": (M, W), t + "attn_pool.mlp.fc1.bias": (M,),
            t + "attn_pool.mlp.fc2.weight": (W, M), t + "attn_pool.mlp.fc2.bias": (W,),
        }
        block = {"norm1.weight": (W,), "norm1.bias": (W,), "attn.qkv.weight": (3 * W, W), "attn.qkv.bias": (3 * W,),
                 "attn.proj.weight": (W, W), "attn.proj.bias": (W,), "norm2.weight": (W,), "norm2.bias": (W,),
                 "mlp.fc1.weight": (M, W), "mlp.fc1.bias": (M,), "mlp.fc2.weight": (W, M), "mlp.fc2.bias": (W,)}
        prefix = t + "blocks.{}."
    else:
        shapes = {
            "visual.conv1.weight": (W, 3, P, P), "visual.class_embedding": (W,), "visual.positional_embedding": (T, W),
            "visual.ln_pre.weight": (W,), "visual.ln_pre.bias": (W,), "visual.ln_post.weight": (W,),
            "visual.ln_post.bias": (W,), "visual.proj": (W, cfg.embed_dim),
        }
        block = {"ln_1.weight": (W,), "ln_1.bias": (W,), "attn.in_proj_weight": (3 * W, W), "attn.in_proj_bias": (3 * W,),
                 "attn.out_proj.weight": (W, W), "attn.out_proj.bias": (W,), "ln_2.weight": (W,), "ln_2.bias": (W,),
                 "mlp.c_fc.weight": (M, W), "mlp.c_fc.bias": (M,), "mlp.c_proj.weight": (W, M), "mlp.c_proj.bias": (W,)}
        prefix = "visual.transformer.resblocks.{}."
    for i in range(L):
        for name, shape in block.items():
            shapes[prefix.format(i) + name] = shape
    return shapes


def random_state_dict_keys(cfg: VitConfig) -> list[str]:
    return list(expected_shapes(cfg))


def flops_per_image(cfg: VitConfig) -> float:
    """2*MAC of the GEMMs and attention of one image (algorithmic, single pass)."""
    T, W = cfg.tokens, cfg.width
    per_layer = 2 * T * W * (3 * W + W + 2 * cfg.mlp) + 2 * 2 * T * T * W
    tail = 2 * W * cfg.embed_dim if cfg.arch == "clip" else 2 * T * W * 2 * W + 2 * 2 * T * W + 2 * W * W + 4 * W * cfg.mlp
    return (T - cfg.has_cls) * 2 * W * 3 * cfg.patch**2 + cfg.layers * per_layer + tail
