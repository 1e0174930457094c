"""Host side of the B200 CLIP ModifiedResNet image tower ("RN50", "RN101"): weights in open_clip's ``visual.*`` naming
-> channels-last conv matrices as split planes + folded BatchNorm vectors + an ``SlbRnWeights`` struct ->
``slb_rn_forward`` (one C-ABI call per batch, no host sync, caller-owned workspace). SURVEY.md §8 f4; the reference
reaches this tower through ``OpenClip("RN50")`` (foundation_models/clip.py:52-62, 103-118; BASELINE configs[0])."""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from .. import _native as N
from .. import ops
from .vit import OPENAI_MEAN, OPENAI_STD

BN_EPS = 1e-5  # nn.BatchNorm2d default, as in open_clip's ModifiedResNet


@dataclass(frozen=True)
class RnConfig:
    name: str
    image_size: int
    width: int
    layers: tuple
    heads: int
    embed_dim: int
    mean: tuple = OPENAI_MEAN
    std: tuple = OPENAI_STD
    arch: str = "modified_resnet"

    @property
    def feature_dim(self) -> int:
        return 32 * self.width

    @property
    def tokens(self) -> int:
        return (self.image_size // 32) ** 2 + 1


# open_clip model configs (model_configs/RN50.json, RN101.json of open-clip-torch 3.0.0), image tower only; the
# quickgelu variants differ in the text tower only
CONFIGS = {
    "RN50": RnConfig("RN50", 224, 64, (3, 4, 6, 3), 32, 1024),
    "RN50-quickgelu": RnConfig("RN50-quickgelu", 224, 64, (3, 4, 6, 3), 32, 1024),
    "RN101": RnConfig("RN101", 224, 64, (3, 4, 23, 3), 32, 512),
    "RN101-quickgelu": RnConfig("RN101-quickgelu", 224, 64, (3, 4, 23, 3), 32, 512),
}


def block_plan(cfg: RnConfig):
    """[(state-dict prefix, inplanes, planes, stride, has_downsample)] in execution order."""
    plan, inplanes = [], cfg.width
    for li, (n, stride) in enumerate(zip(cfg.layers, (1, 2, 2, 2))):
        planes = cfg.width * 2**li
        for bi in range(n):
            s = stride if bi == 0 else 1
            plan.append((f"visual.layer{li + 1}.{bi}.", inplanes, planes, s, s > 1 or inplanes != planes * 4))
            inplanes = planes * 4
    return plan


def conv_names(cfg: RnConfig):
    """[(conv weight key, BatchNorm prefix)] in the order SlbRnWeights.convs expects."""
    names = [("visual.conv1.weight", "visual.bn1"), ("visual.conv2.weight", "visual.bn2"), ("visual.conv3.weight", "visual.bn3")]
    for p, _i, _p, _s, ds in block_plan(cfg):
        names += [(p + "conv1.weight", p + "bn1"), (p + "conv2.weight", p + "bn2"), (p + "conv3.weight", p + "bn3")]
        if ds:
            names.append((p + "downsample.0.weight", p + "downsample.1"))
    return names


def state_dict_keys(cfg: RnConfig) -> list[str]:
    keys = []
    for wname, bn in conv_names(cfg):
        keys += [wname] + [f"{bn}.{s}" for s in ("weight", "bias", "running_mean", "running_var")]
    a = "visual.attnpool."
    keys += [a + "positional_embedding"] + [f"{a}{n}_proj.{s}" for n in "qkvc" for s in ("weight", "bias")]
    return keys


def random_state_dict(cfg: RnConfig, seed: int = 1) -> dict[str, torch.Tensor]:
    """Random ``visual.*`` weights (nothing can be downloaded here): He-scaled convolutions, BatchNorm statistics near
    (0, 1), attention pool at feature_dim**-0.5."""
    g = torch.Generator().manual_seed(seed)

    def rn(*s):
        return torch.randn(*s, generator=g)

    sd: dict[str, torch.Tensor] = {}
    shapes = [(cfg.width // 2, 3, 3), (cfg.width // 2, cfg.width // 2, 3), (cfg.width, cfg.width // 2, 3)]
    for _p, inpl, pl, _s, ds in block_plan(cfg):
        shapes += [(pl, inpl, 1), (pl, pl, 3), (4 * pl, pl, 1)] + ([(4 * pl, inpl, 1)] if ds else [])
    for (wname, bn), (cout, cin, k) in zip(conv_names(cfg), shapes):
        sd[wname] = rn(cout, cin, k, k) * (2.0 / (cin * k * k)) ** 0.5
        gain = 0.5 if bn.endswith("bn3") and "layer" in bn else 0.7 if "downsample" in bn else 1.0
        sd[bn + ".weight"] = gain * (1 + 0.1 * rn(cout))
        sd[bn + ".bias"] = 0.1 * rn(cout)
        sd[bn + ".running_mean"] = 0.1 * rn(cout)
        sd[bn + ".running_var"] = 1 + 0.2 * torch.rand(cout, generator=g)
    E, a = cfg.feature_dim, "visual.attnpool."
    sd[a + "positional_embedding"] = rn(cfg.tokens, E) / E**0.5
    for nm, out in (("q_proj", E), ("k_proj", E), ("v_proj", E), ("c_proj", cfg.embed_dim)):
        sd[a + nm + ".weight"] = rn(out, E) * E**-0.5
        sd[a + nm + ".bias"] = 0.02 * rn(out)
    return sd


class RnTower:
    """Device-resident ModifiedResNet: conv matrices as split planes, folded BatchNorm vectors, cached workspace."""

    def __init__(self, cfg: RnConfig, state_dict: dict[str, torch.Tensor], device, plane_format: int = N.PLANE_F16):
        if cfg.width % 64 or cfg.image_size % 32:
            raise ValueError("the B200 ModifiedResNet needs width % 64 == 0 and image_size % 32 == 0 (RN50, RN101)")
        self.cfg = cfg
        self.plane_format = plane_format
        self.state_dict = {k: v.detach().to(torch.float32).cpu() for k, v in state_dict.items()
                           if k.startswith("visual.") and v.is_floating_point()}
        missing = [k for k in state_dict_keys(cfg) if k not in self.state_dict]
        if missing:
            raise KeyError(f"state dict is missing {len(missing)} image-tower tensors, e.g. {missing[:3]}")
        self._device = torch.device("cpu")
        self._struct = None
        self._keep: list = []
        self._ws: torch.Tensor | None = None
        self.to(device)

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
        lib = N.load(require_device=True)
        cfg, sd, dev, fmt = self.cfg, self.state_dict, self._device, self.plane_format
        keep = self._keep

        def vec(t: torch.Tensor):
            t = t.to(torch.float32).to(dev).contiguous()
            keep.append(t)
            return t.data_ptr()

        def planes(mat: torch.Tensor):
            t = ops.split_planes(mat.to(dev), fmt, N.WEIGHT_PLANE_SCALE)
            keep.append(t)
            return t.data_ptr()

        names = conv_names(cfg)
        convs = (N.SlbConvBn * len(names))()
        for c, (wname, bn) in zip(convs, names):
            wt = sd[wname]
            cout, cin, k, _ = wt.shape
            kpad = int(lib.slb_conv_k(cin, k))
            mat = torch.zeros(cout, kpad)
            mat[:, : cin * k * k] = wt.permute(0, 2, 3, 1).reshape(cout, cin * k * k)  # (cout, ky, kx, cin): channels-last taps
            # eval-mode BatchNorm as a per-channel scale and shift of the convolution output (computed in float64)
            scale = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + BN_EPS)
            shift = sd[bn + ".bias"].double() - sd[bn + ".running_mean"].double() * scale
            c.w, c.scale, c.shift = planes(mat), vec(scale), vec(shift)
            c.cin, c.cout, c.ksize = cin, cout, k
        a = "visual.attnpool."
        w = N.SlbRnWeights()
        w.image_size, w.width, w.heads, w.out_dim = cfg.image_size, cfg.width, cfg.heads, cfg.embed_dim
        for i, n in enumerate(cfg.layers):
            w.blocks[i] = n
        w.plane_fmt, w.n_convs = fmt, len(names)
        w.convs = convs
        keep.append(convs)
        w.pos = vec(sd[a + "positional_embedding"])
        w.w_q, w.b_q = planes(sd[a + "q_proj.weight"]), vec(sd[a + "q_proj.bias"])
        w.w_kv = planes(torch.cat([sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]]))
        w.b_kv = vec(torch.cat([sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]]))
        w.w_c, w.b_c = planes(sd[a + "c_proj.weight"]), vec(sd[a + "c_proj.bias"])
        self._struct = w

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
        need = lib.slb_rn_workspace_bytes(ctypes.byref(self._struct), B)
        if need == 0:
            raise N.SlbError("slb_rn_workspace_bytes rejected the configuration")
        if self._ws is None or self._ws.numel() < need or self._ws.device != img.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=img.device)
        with torch.cuda.device(img.device):
            rc = lib.slb_rn_forward(ctypes.byref(self._struct), img.data_ptr(), B, out.data_ptr(), self._ws.data_ptr(),
                                    self._ws.numel(), N.stream_ptr(img.device))
        N.check(rc, "slb_rn_forward")
        return out


def flops_per_image(cfg: RnConfig) -> float:
    """2*MAC of the convolutions and the attention pool of one image (algorithmic, single pass)."""
    S, w = cfg.image_size, cfg.width
    h = S // 2
    f = 2.0 * h * h * (27 * (w // 2) + 9 * (w // 2) * (w // 2) + 9 * (w // 2) * w)
    h //= 2
    for _p, inpl, pl, stride, ds in block_plan(cfg):
        f += 2.0 * h * h * (inpl * pl + 9 * pl * pl)
        h //= stride
        f += 2.0 * h * h * (pl * 4 * pl + (inpl * 4 * pl if ds else 0))
    E, T = 32 * w, h * h + 1
    return f + 2.0 * T * E * 2 * E + 2.0 * E * E + 4.0 * T * E + 2.0 * E * cfg.embed_dim
