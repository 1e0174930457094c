"""TEST INFRASTRUCTURE ONLY — torch-CPU restatement of the CLIP ModifiedResNet image tower.

The reference reaches this tower through ``OpenClip("RN50").encode_image`` (foundation_models/clip.py:103-118 ->
``open_clip`` ``ModifiedResNet.forward``); BASELINE.json's configs[0] embeds with it. ``open_clip`` (open-clip-torch
3.0.0, ``modified_resnet.py``) is a third-party dependency that is not vendored in /root/reference and is not installed
here, so this file restates its published architecture with the same torch primitives it calls
(``F.conv2d``, ``F.batch_norm`` in eval mode, ``F.avg_pool2d``, ``F.multi_head_attention_forward`` with separate
q/k/v projection weights):

* stem: conv3x3(3 -> w/2, stride 2) BN ReLU, conv3x3(w/2 -> w/2) BN ReLU, conv3x3(w/2 -> w) BN ReLU, AvgPool2d(2)
* layer1..4 of Bottleneck(inplanes, planes, stride): conv1x1 BN ReLU, conv3x3 BN ReLU, AvgPool2d(stride) (the
  anti-aliased stride), conv1x1 BN; shortcut = AvgPool2d(stride) -> conv1x1 -> BN when stride > 1 or the channel count
  changes; out = ReLU(main + shortcut); planes = w, 2w, 4w, 8w, strides 1, 2, 2, 2
* AttentionPool2d: tokens = [mean over positions; positions] + positional_embedding, multi-head attention with the
  mean token as the only query that is kept, output projection c_proj to the embedding dimension.

State-dict names follow open_clip (``visual.conv1.weight``, ``visual.bn1.running_mean``, ``visual.layer2.0.downsample.0.weight``,
``visual.attnpool.q_proj.weight`` ...), so real checkpoints load into both this port and the B200 tower.

PARITY UNPINNED for the wiring: no golden vector of this tower exists in the reference (its tests construct
``OpenClip`` with ViT names only) and open_clip cannot be imported here; every arithmetic step is torch's own operator.
Partial pins (tests/test_oracle_rn.py): the stride-1 bottleneck equals torchvision's ``Bottleneck`` with the same weights,
the attention pool an explicit softmax restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)
BN_EPS = 1e-5


@dataclass(frozen=True)
class RnConfig:
    name: str
    image_size: int
    width: int
    layers: tuple
    heads: int
    embed_dim: int  # output dimension of the attention pool
    mean: tuple = OPENAI_MEAN
    std: tuple = OPENAI_STD


CONFIGS = {
    "RN50": RnConfig("RN50", 224, 64, (3, 4, 6, 3), 32, 1024),
    "RN101": RnConfig("RN101", 224, 64, (3, 4, 23, 3), 32, 512),
    "RN-tiny-test": RnConfig("RN-tiny-test", 64, 64, (1, 1, 1, 1), 32, 256),
    "RN-small-test": RnConfig("RN-small-test", 96, 64, (2, 1, 2, 1), 32, 128),
}


def block_plan(cfg: RnConfig):
    """[(prefix, inplanes, planes, stride, has_downsample)] in execution order."""
    plan, inplanes = [], cfg.width
    for li, (n, stride) in enumerate(zip(cfg.layers, (1, 2, 2, 2))):
        planes = cfg.width * 2**li
        for bi in range(n):
            s = stride if bi == 0 else 1
            plan.append((f"visual.layer{li + 1}.{bi}.", inplanes, planes, s, s > 1 or inplanes != planes * 4))
            inplanes = planes * 4
    return plan


def init_weights(cfg: RnConfig, seed: int = 5) -> dict[str, torch.Tensor]:
    """Random weights with He-scaled convolutions and BatchNorm statistics near (0, 1): activations stay O(1)."""
    g = torch.Generator().manual_seed(seed)

    def rn(*s):
        return torch.randn(*s, generator=g)

    sd: dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = rn(cout, cin, k, k) * (2.0 / (cin * k * k)) ** 0.5

    def bn(name, c, gain=1.0):
        sd[name + ".weight"] = gain * (1 + 0.1 * rn(c))
        sd[name + ".bias"] = 0.1 * rn(c)
        sd[name + ".running_mean"] = 0.1 * rn(c)
        sd[name + ".running_var"] = 1 + 0.2 * torch.rand(c, generator=g)

    w = cfg.width
    conv("visual.conv1", w // 2, 3, 3), bn("visual.bn1", w // 2)
    conv("visual.conv2", w // 2, w // 2, 3), bn("visual.bn2", w // 2)
    conv("visual.conv3", w, w // 2, 3), bn("visual.bn3", w)
    for p, inplanes, planes, stride, ds in block_plan(cfg):
        conv(p + "conv1", planes, inplanes, 1), bn(p + "bn1", planes)
        conv(p + "conv2", planes, planes, 3), bn(p + "bn2", planes)
        conv(p + "conv3", planes * 4, planes, 1), bn(p + "bn3", planes * 4, gain=0.5)
        if ds:
            conv(p + "downsample.0", planes * 4, inplanes, 1), bn(p + "downsample.1", planes * 4, gain=0.7)
    E = w * 32
    n_pos = (cfg.image_size // 32) ** 2 + 1
    a = "visual.attnpool."
    sd[a + "positional_embedding"] = rn(n_pos, E) / E**0.5
    for nm, out in (("q_proj", E), ("k_proj", E), ("v_proj", E), ("c_proj", cfg.embed_dim)):
        sd[a + nm + ".weight"] = rn(out, E) * E**-0.5
        sd[a + nm + ".bias"] = 0.02 * rn(out)
    return sd


def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        False, 0.0, BN_EPS)


def _bottleneck(x, sd, p, stride, ds):
    out = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[p + "conv2.weight"], padding=1), sd, p + "bn2"))
    if stride > 1:
        out = F.avg_pool2d(out, stride)
    out = _bn(F.conv2d(out, sd[p + "conv3.weight"]), sd, p + "bn3")
    identity = x
    if ds:
        identity = F.avg_pool2d(x, stride) if stride > 1 else x
        identity = _bn(F.conv2d(identity, sd[p + "downsample.0.weight"]), sd, p + "downsample.1")
    return F.relu(out + identity)


def trunk(sd, cfg: RnConfig, img: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
    """(B, 3, S, S) preprocessed images -> (B, 32 w, S/32, S/32) feature map."""
    x = F.relu(_bn(F.conv2d(img, sd["visual.conv1.weight"], stride=2, padding=1), sd, "visual.bn1"))
    x = F.relu(_bn(F.conv2d(x, sd["visual.conv2.weight"], padding=1), sd, "visual.bn2"))
    x = F.relu(_bn(F.conv2d(x, sd["visual.conv3.weight"], padding=1), sd, "visual.bn3"))
    x = F.avg_pool2d(x, 2)
    if taps is not None:
        taps["stem"] = x
    for p, _inpl, _pl, stride, ds in block_plan(cfg):
        x = _bottleneck(x, sd, p, stride, ds)
        if taps is not None:
            taps[p] = x
    return x


def attnpool(sd, cfg: RnConfig, x: torch.Tensor) -> torch.Tensor:
    a = "visual.attnpool."
    B, C = x.shape[0], x.shape[1]
    x = x.reshape(B, C, -1).permute(2, 0, 1)  # (HW, B, C)
    x = torch.cat([x.mean(dim=0, keepdim=True), x], dim=0)
    x = x + sd[a + "positional_embedding"][:, None, :]
    out, _ = F.multi_head_attention_forward(
        query=x, key=x, value=x, embed_dim_to_check=C, num_heads=cfg.heads,
        q_proj_weight=sd[a + "q_proj.weight"], k_proj_weight=sd[a + "k_proj.weight"], v_proj_weight=sd[a + "v_proj.weight"],
        in_proj_weight=None,
        in_proj_bias=torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]]),
        bias_k=None, bias_v=None, add_zero_attn=False, dropout_p=0.0,
        out_proj_weight=sd[a + "c_proj.weight"], out_proj_bias=sd[a + "c_proj.bias"],
        use_separate_proj_weight=True, training=False, need_weights=False)
    return out[0]


@torch.no_grad()
def encode_image(sd, cfg: RnConfig, img: torch.Tensor, dtype=torch.float32, taps: dict | None = None) -> torch.Tensor:
    sd = {k: v.to(dtype) for k, v in sd.items() if k.startswith("visual.") and v.is_floating_point()}
    return attnpool(sd, cfg, trunk(sd, cfg, img.to(dtype), taps))


def flops_per_image(cfg: RnConfig) -> float:
    """2*MAC of the convolutions and the attention pool of one image."""
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
