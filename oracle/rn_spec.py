"""TEST INFRASTRUCTURE ONLY — a SECOND, separately written restatement of CLIP's ModifiedResNet, used to cross-check
oracle/rn_port.py (VERDICT r1 "pin the ModifiedResNet oracle").

oracle/rn_port.py is functional code over a flat state dict. This file instead rebuilds the tower the way the CLIP paper's
public model definition composes it (Radford et al. 2021, "Learning Transferable Visual Models ...", sec. 3.2 + the
released model file; open_clip's modified_resnet.py keeps that structure): ``nn.Module`` objects whose nesting alone
produces the checkpoint key names, and an attention pool written with explicit matrix products instead of
``F.multi_head_attention_forward``. Loading rn_port's weights with ``strict=True`` therefore checks every key name and
shape against the module structure, and equal outputs check the wiring of the stem, of the stride-1 and the stride-2
(average-pooled) bottlenecks and of the pool — two independent statements of the same published architecture.

What this does NOT pin: that the published architecture is what open_clip 3.0.0 executes. Neither open_clip nor a
fixture of it can be had here (no network, not in the wheelhouse); both restatements come from the same public
description. The three ResNet-D details that differ from torchvision are taken from that description:
  (1) a 3-convolution stem (3x3/2, 3x3, 3x3, each BN + ReLU) followed by a 2x2 average pool instead of 7x7/2 + max pool;
  (2) strided bottlenecks convolve at stride 1 and then average-pool (anti-aliasing), on the main path after the 3x3
      convolution and on the shortcut before its 1x1 convolution;
  (3) the final global average pool is replaced by a single-query multi-head attention pool over [mean token; tokens]
      with a learned positional embedding and separate q/k/v/c projections.
"""

from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1):
        super().__init__()
        # all convolutions have stride 1; an average pool follows the second one when stride > 1
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.act1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.act2 = nn.ReLU(inplace=True)
        self.avgpool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.act3 = nn.ReLU(inplace=True)
        self.downsample = None
        if stride > 1 or inplanes != planes * self.expansion:
            # the shortcut is average-pooled first, then a stride-1 1x1 convolution ("-1" carries no parameters, so the
            # checkpoint keys of this branch are downsample.0.* and downsample.1.*)
            self.downsample = nn.Sequential(OrderedDict([
                ("-1", nn.AvgPool2d(stride)),
                ("0", nn.Conv2d(inplanes, planes * self.expansion, 1, stride=1, bias=False)),
                ("1", nn.BatchNorm2d(planes * self.expansion)),
            ]))

    def forward(self, x):
        identity = x
        out = self.act1(self.bn1(self.conv1(x)))
        out = self.act2(self.bn2(self.conv2(out)))
        out = self.avgpool(out)
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return self.act3(out + identity)


class AttentionPool2d(nn.Module):
    def __init__(self, spatial: int, embed_dim: int, num_heads: int, output_dim: int):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.zeros(spatial * spatial + 1, embed_dim))
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim)
        self.num_heads = num_heads

    def forward(self, x):
        b, c, h, w = x.shape
        tokens = x.flatten(2).transpose(1, 2)  # (B, HW, C)
        tokens = torch.cat([tokens.mean(dim=1, keepdim=True), tokens], dim=1) + self.positional_embedding
        heads, dh = self.num_heads, c // self.num_heads
        q = self.q_proj(tokens[:, :1]).view(b, 1, heads, dh).transpose(1, 2) * dh**-0.5  # only the mean token queries
        k = self.k_proj(tokens).view(b, -1, heads, dh).transpose(1, 2)
        v = self.v_proj(tokens).view(b, -1, heads, dh).transpose(1, 2)
        attn = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
        pooled = (attn @ v).transpose(1, 2).reshape(b, c)
        return self.c_proj(pooled)


class ModifiedResNet(nn.Module):
    def __init__(self, layers, output_dim: int, heads: int, image_size: int = 224, width: int = 64):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width // 2, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width // 2)
        self.act1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(width // 2, width // 2, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width // 2)
        self.act2 = nn.ReLU(inplace=True)
        self.conv3 = nn.Conv2d(width // 2, width, kernel_size=3, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(width)
        self.act3 = nn.ReLU(inplace=True)
        self.avgpool = nn.AvgPool2d(2)
        self._inplanes = width
        self.layer1 = self._make_layer(width, layers[0])
        self.layer2 = self._make_layer(width * 2, layers[1], stride=2)
        self.layer3 = self._make_layer(width * 4, layers[2], stride=2)
        self.layer4 = self._make_layer(width * 8, layers[3], stride=2)
        self.attnpool = AttentionPool2d(image_size // 32, width * 32, heads, output_dim)

    def _make_layer(self, planes: int, blocks: int, stride: int = 1):
        seq = [Bottleneck(self._inplanes, planes, stride)]
        self._inplanes = planes * Bottleneck.expansion
        seq += [Bottleneck(self._inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    def stem(self, x):
        x = self.act1(self.bn1(self.conv1(x)))
        x = self.act2(self.bn2(self.conv2(x)))
        x = self.act3(self.bn3(self.conv3(x)))
        return self.avgpool(x)

    def forward(self, x):
        x = self.stem(x)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.attnpool(x)


def build(layers, output_dim, heads, image_size, width, state_dict: dict, dtype=torch.float64) -> ModifiedResNet:
    """Instantiate and load a ``visual.*`` state dict with strict key / shape checking."""
    net = ModifiedResNet(layers, output_dim, heads, image_size, width).to(dtype).eval()
    own = {k[len("visual."):]: v.to(dtype) for k, v in state_dict.items() if k.startswith("visual.")}
    missing, unexpected = net.load_state_dict(own, strict=False)
    missing = [k for k in missing if not k.endswith("num_batches_tracked")]
    if missing or unexpected:
        raise KeyError(f"state dict does not fit the published module structure: missing {missing[:4]}, unexpected {unexpected[:4]}")
    return net
