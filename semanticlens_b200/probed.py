"""Opt-in B200 forward for the PROBED model when it is a torchvision-style ResNet or VisionTransformer.

The reference sweeps the dataset through ``self.model(x)`` under forward hooks
(``ActivationComponentVisualizer._run``, activation_based.py:341-358); the model is user code and by default this package
leaves it to PyTorch (cuDNN fp32 — 88 % of the cfg-2 step). For ``torchvision.models.ResNet``
(BasicBlock or Bottleneck: the probed models of BASELINE.json configs[0..3]) ``AcceleratedResNet`` produces the same maps
with the package's own kernels: channels-last split planes, every convolution a tcgen05 ``slb_gemm_split`` with eval-mode
BatchNorm / ReLU / shortcut in its epilogue (csrc/convnet.cu, csrc/gemm_tc.cu) — fp32-grade (22-bit operands, fp32
accumulate: ~1e-5 of the largest activation after 50 layers, measured in tests/test_probed_gpu.py), not bit-identical to
cuDNN's fp32, hence opt-in: ``ActivationComponentVisualizer(..., accelerate=True)`` or ``SLB_ACCEL_FORWARD=1``.

``torchvision.models.VisionTransformer`` (BASELINE configs[2]: ViT-B/16 probed at its encoder blocks) runs on the ViT
tower's kernels the same way (``AcceleratedViT``: ``slb_vit_trunk`` block by block, hooks fired on the residual stream).

Hooks keep working: whatever forward hooks are registered on the model's modules (the ones ``ActMaxCache`` installs) are
called with the module's output as a ``(B, C, H, W)`` tensor in channels-last memory, which K1 reads in place. Supported
hook points: every ``nn.Conv2d`` (its raw output, before BatchNorm), ``maxpool``, every residual block and ``layer1..4``.
The forward stops after the last hooked module (the reference discards the logits).
"""

from __future__ import annotations

import ctypes
import os

import torch
from torch import nn

from . import _native as N
from . import ops

ALPHA = 1.0 / (N.ACT_PLANE_SCALE * N.WEIGHT_PLANE_SCALE)


def accel_requested(flag) -> bool:
    """``accelerate=`` of the visualizer: True / False, or None = the SLB_ACCEL_FORWARD environment variable."""
    if flag is None:
        return os.environ.get("SLB_ACCEL_FORWARD", "0") == "1"
    return bool(flag)


class _Conv:
    """One convolution + its BatchNorm as GEMM operands: weight planes (Cout, conv_k(Cin, k)) with columns ordered
    (ky, kx, cin) like the im2col kernels, scale = gamma / sqrt(var + eps), shift = beta - mean * scale (folded in float64)."""

    def __init__(self, conv: nn.Conv2d, bn: nn.BatchNorm2d | None, device, fmt: int):
        k = conv.kernel_size[0]
        ok = (conv.kernel_size[0] == conv.kernel_size[1] and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]
              and conv.groups == 1 and conv.dilation == (1, 1) and conv.padding_mode == "zeros" and conv.out_channels % 8 == 0)
        if not ok:
            raise NotImplementedError(f"accelerated forward: unsupported convolution {conv}")
        self.module = conv
        self.k, self.stride, self.pad = k, conv.stride[0], conv.padding[0]
        self.cin, self.cout = conv.in_channels, conv.out_channels
        w = conv.weight.detach().to(torch.float32).cpu()
        kk = ops.conv_k(self.cin, k)
        mat = torch.zeros((self.cout, kk), dtype=torch.float32)
        mat[:, : self.cin * k * k] = w.permute(0, 2, 3, 1).reshape(self.cout, -1)
        self.w = ops.split_planes(mat.to(device), fmt, N.WEIGHT_PLANE_SCALE)
        self.bias = None if conv.bias is None else conv.bias.detach().to(device=device, dtype=torch.float32).contiguous()
        if bn is None:
            scale = torch.ones(self.cout, dtype=torch.float64)
            shift = torch.zeros(self.cout, dtype=torch.float64)
        else:
            if bn.training or bn.running_mean is None:
                raise NotImplementedError("accelerated forward: BatchNorm2d must be in eval mode with running statistics")
            gamma = bn.weight.detach().double().cpu() if bn.affine else torch.ones(self.cout, dtype=torch.float64)
            beta = bn.bias.detach().double().cpu() if bn.affine else torch.zeros(self.cout, dtype=torch.float64)
            scale = gamma / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
            shift = beta - bn.running_mean.detach().double().cpu() * scale
        if conv.bias is not None:
            shift = shift + conv.bias.detach().double().cpu() * scale
        self.scale = scale.float().to(device).contiguous()
        self.shift = shift.float().to(device).contiguous()


def _fire(module: nn.Module, out: torch.Tensor) -> None:
    for hook in list(module._forward_hooks.values()):
        hook(module, (), out)


def _as_nchw(map32: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """(B*H*W, C) fp32 -> a (B, C, H, W) view in channels-last memory (what a hook receives; K1 reads it in place)."""
    return map32.view(B, H, W, map32.shape[1]).permute(0, 3, 1, 2)


class AcceleratedResNet:
    """Runs a ``torchvision.models.ResNet`` (eval mode) on the B200 kernels and calls the forward hooks registered on its
    modules. ``passes``: ``N.PASSES_SPLIT_ACC`` (default; cross terms in their own accumulator, as the CLIP ResNet tower)
    or 3 (single accumulator: the faster CTA-pair tiles, ~3x the error)."""

    def __init__(self, model: nn.Module, device=None, plane_format: str = "f16", passes: int | None = None):
        N.load(require_device=True)
        need = ("conv1", "bn1", "maxpool", "layer1", "layer2", "layer3", "layer4")
        if not all(hasattr(model, n) for n in need):
            raise NotImplementedError("accelerated forward: expected a torchvision-style ResNet (conv1, bn1, maxpool, layer1..4)")
        if model.training:
            raise NotImplementedError("accelerated forward: put the model in eval mode (BatchNorm uses running statistics)")
        mp = model.maxpool
        if not (isinstance(mp, nn.MaxPool2d) and mp.kernel_size in (3, (3, 3)) and mp.stride in (2, (2, 2))
                and mp.padding in (1, (1, 1)) and mp.dilation in (1, (1, 1)) and not mp.ceil_mode):
            raise NotImplementedError(f"accelerated forward: unsupported pooling {mp}")
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise N.SlbError("the accelerated forward runs on a CUDA device (there is no CPU fallback)")
        self.fmt = {"f16": N.PLANE_F16, "bf16": N.PLANE_BF16}[plane_format]
        self.passes = passes if passes is not None else (3 if os.environ.get("SLB_ACCEL_FAST", "0") == "1" else N.PASSES_SPLIT_ACC)
        self.explicit_im2col = os.environ.get("SLB_ACCEL_IM2COL", "0") == "1"  # debugging: materialise the im2col matrices
        dev, fmt = self.device, self.fmt
        self.stem = _Conv(model.conv1, model.bn1, dev, fmt)
        self.blocks = []  # (block module, kind, convs..., downsample | None, owning layer if it is the layer's last block)
        for lname in ("layer1", "layer2", "layer3", "layer4"):
            layer = getattr(model, lname)
            blocks = list(layer.children())
            for i, blk in enumerate(blocks):
                ds = None
                if getattr(blk, "downsample", None) is not None:
                    d = list(blk.downsample.children())
                    if len(d) != 2 or not isinstance(d[0], nn.Conv2d) or not isinstance(d[1], nn.BatchNorm2d):
                        raise NotImplementedError(f"accelerated forward: unsupported shortcut {blk.downsample}")
                    ds = _Conv(d[0], d[1], dev, fmt)
                if hasattr(blk, "conv3"):
                    convs = [_Conv(blk.conv1, blk.bn1, dev, fmt), _Conv(blk.conv2, blk.bn2, dev, fmt), _Conv(blk.conv3, blk.bn3, dev, fmt)]
                    kind = "bottleneck"
                elif hasattr(blk, "conv2"):
                    convs = [_Conv(blk.conv1, blk.bn1, dev, fmt), _Conv(blk.conv2, blk.bn2, dev, fmt)]
                    kind = "basic"
                else:
                    raise NotImplementedError(f"accelerated forward: unsupported block {type(blk).__name__}")
                self.blocks.append((blk, kind, convs, ds, layer if i == len(blocks) - 1 else None))
        # modules whose output this forward can hand to a hook
        self._tappable = {id(self.stem.module), id(model.maxpool)}
        for blk, _kind, convs, ds, layer in self.blocks:
            self._tappable.add(id(blk))
            self._tappable.update(id(c.module) for c in convs)
            if ds is not None:
                self._tappable.add(id(ds.module))
            if layer is not None:
                self._tappable.add(id(layer))

    # ------------------------------------------------------------------------------------------------------------------
    def _hooked(self, module: nn.Module) -> bool:
        return len(module._forward_hooks) > 0

    def check_hooks(self) -> None:
        """Every module carrying a forward hook must be one whose output this forward materialises."""
        for name, m in self.model.named_modules():
            if len(m._forward_hooks) and id(m) not in self._tappable:
                raise NotImplementedError(
                    f"accelerated forward: cannot expose the output of '{name}' ({type(m).__name__}); hook a Conv2d, maxpool, a "
                    "residual block or layer1..4, or run with accelerate=False")

    def _mm(self, c: _Conv, x_planes, B, H, W, **kw):
        """The convolution's contraction over channels-last planes of a (B, H, W, cin) map, epilogue arguments in ``kw``:
        1x1 / stride 1 is the GEMM itself; everything else with cin % 64 == 0 is the implicit GEMM (TMA im2col-mode loads,
        the stride in the tensor map: no im2col matrix, no subsampled copy); narrow 3x3 inputs take the explicit im2col."""
        if c.k == 1 and c.stride == 1:
            return ops.gemm_split(x_planes, c.w, alpha=ALPHA, passes=self.passes, **kw)
        if c.cin % 64 == 0 and not self.explicit_im2col:
            return ops.conv_gemm(x_planes, B, H, W, c.w, c.k, c.stride, c.pad, alpha=ALPHA, passes=self.passes, **kw)
        if c.k == 1:
            return ops.gemm_split(ops.subsample2_planes(x_planes, B, H, W), c.w, alpha=ALPHA, passes=self.passes, **kw)
        if c.k == 3 and c.pad == 1:
            return ops.gemm_split(ops.im2col3x3_strided(x_planes, B, H, W, c.stride), c.w, alpha=ALPHA, passes=self.passes, **kw)
        raise NotImplementedError(f"accelerated forward: unsupported convolution {c.module}")

    def _conv(self, c: _Conv, x_planes, B, H, W, *, relu, residual=None, want_f32=False, out_f32=None, want_planes=True):
        """One convolution over the planes of a (B, H, W, cin) map: BatchNorm + optional shortcut + optional ReLU. Returns
        (fp32 map | None, planes | None) of the (B, Ho, Wo, cout) output. A hooked convolution leaves its raw output first."""
        Ho, Wo = ops.conv_out(H, c.k, c.stride, c.pad), ops.conv_out(W, c.k, c.stride, c.pad)
        keep32 = out_f32 if out_f32 is not None else bool(want_f32)
        if self._hooked(c.module) and c.bias is None and not (residual is not None and not relu):
            # the hook wants the RAW convolution output: the GEMM leaves it (raw_f32) next to the BatchNorm / ReLU / shortcut
            # result of the same epilogue — one pass over the accumulator, no second kernel
            raw = torch.empty((B * Ho * Wo, c.cout), dtype=torch.float32, device=x_planes.device)
            epi = (N.EPI_ADD_RELU if residual is not None else N.EPI_RELU) if relu else N.EPI_NONE
            out = self._mm(c, x_planes, B, H, W, bias=c.shift, residual=residual, col_scale=c.scale, epilogue=epi, out_f32=keep32,
                           out_planes=want_planes, raw_f32=raw)
            _fire(c.module, _as_nchw(raw, B, Ho, Wo))
            return out
        if self._hooked(c.module):  # a convolution with its own bias: raw output first, then the affine + activation pass
            raw, _ = self._mm(c, x_planes, B, H, W, bias=c.bias, epilogue=N.EPI_NONE)
            _fire(c.module, _as_nchw(raw, B, Ho, Wo))
            shift = c.shift if c.bias is None else (c.shift - c.bias * c.scale)  # raw already holds the bias
            if residual is not None and residual.dtype != torch.float32:
                residual = ops.planes_to_f32(residual)  # (rare path: a hooked convolution with its own bias at a block tail)
            return ops.affine_act(raw, c.scale, shift, residual=residual, relu=relu, fmt=self.fmt, out_f32=keep32, out_planes=want_planes)
        if residual is not None and not relu:
            raise AssertionError("a shortcut add is always followed by the ReLU in a ResNet block")
        epi = (N.EPI_ADD_RELU if residual is not None else N.EPI_RELU) if relu else N.EPI_NONE
        return self._mm(c, x_planes, B, H, W, bias=c.shift, residual=residual, col_scale=c.scale, epilogue=epi, out_f32=keep32,
                        out_planes=want_planes)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, logits: bool = False):
        """x (B, 3, H, W) fp32 on the device. Fires the registered hooks; returns the logits (``logits=True``: the global
        pool and ``fc`` run in torch on the last map) or None."""
        self.check_hooks()
        m = self.model
        x = x.to(self.device, dtype=torch.float32).contiguous()
        B, _, H, W = x.shape
        last = None  # index of the last block that must run
        if not logits:
            hooked = {id(mod) for mod in m.modules() if len(mod._forward_hooks)}
            for i, (blk, _kind, convs, ds, layer) in enumerate(self.blocks):
                ids = {id(blk)} | {id(c.module) for c in convs} | ({id(ds.module)} if ds else set()) | ({id(layer)} if layer else set())
                if ids & hooked:
                    last = i
        else:
            last = len(self.blocks) - 1
        # ---- stem: k x k / stride conv from NCHW, raw map (hook point), BatchNorm + ReLU + MaxPool in one pass ----
        st = self.stem
        col = ops.im2col_nchw(x, st.k, st.stride, st.pad, self.fmt)
        H, W = ops.conv_out(H, st.k, st.stride, st.pad), ops.conv_out(W, st.k, st.stride, st.pad)
        raw, _ = ops.gemm_split(col, st.w, bias=st.bias, alpha=ALPHA, epilogue=N.EPI_NONE, passes=self.passes)
        del col
        if self._hooked(st.module):
            _fire(st.module, _as_nchw(raw, B, H, W))
        if last is None and not self._hooked(m.maxpool):
            return None
        shift = st.shift if st.bias is None else (st.shift - st.bias * st.scale)
        x32, xpl = ops.bn_relu_maxpool(raw, B, H, W, st.scale, shift, self.fmt, want_f32=self._hooked(m.maxpool))
        del raw
        H, W = ops.conv_out(H, 3, 2, 1), ops.conv_out(W, 3, 2, 1)
        if self._hooked(m.maxpool):
            _fire(m.maxpool, _as_nchw(x32, B, H, W))
        if last is None:
            return None
        # ---- residual blocks ----
        for i, (blk, kind, convs, ds, layer) in enumerate(self.blocks[: last + 1]):
            stride = convs[1].stride if kind == "bottleneck" else convs[0].stride
            Ho, Wo = ops.conv_out(H, 3, stride, 1), ops.conv_out(W, 3, stride, 1)
            # The residual stream lives as split planes (22 bits): the shortcut is the block's input planes, or its 1x1
            # (strided) projection + BatchNorm written as planes, and the tail reads it through SLB_EPI_ADD_RELU_PLANES —
            # the same bytes as an fp32 shortcut, but a block whose output nobody hooks never writes an fp32 copy of it
            # (a third of the tail convolution's HBM traffic, the largest single item of the forward).
            idf = xpl if ds is None else self._conv(ds, xpl, B, H, W, relu=False, want_f32=False, want_planes=True)[1]
            need32 = self._hooked(blk) or (layer is not None and self._hooked(layer)) or (logits and i == last)
            if kind == "bottleneck":
                if convs[0].stride != 1 or convs[2].stride != 1:
                    raise NotImplementedError("accelerated forward: the stride of a Bottleneck must sit in conv2 (torchvision v1.5)")
                _, t1 = self._conv(convs[0], xpl, B, H, W, relu=True)
                _, t2 = self._conv(convs[1], t1, B, H, W, relu=True)
                del t1
                x32, xpl = self._conv(convs[2], t2, B, Ho, Wo, relu=True, residual=idf, want_f32=need32)
                del t2
            else:
                _, t1 = self._conv(convs[0], xpl, B, H, W, relu=True)
                x32, xpl = self._conv(convs[1], t1, B, Ho, Wo, relu=True, residual=idf, want_f32=need32)
                del t1
            del idf
            H, W = Ho, Wo
            if self._hooked(blk):
                _fire(blk, _as_nchw(x32, B, H, W))
            if layer is not None and self._hooked(layer):
                _fire(layer, _as_nchw(x32, B, H, W))
        if not logits:
            return None
        feat = x32.view(B, H * W, -1).mean(dim=1)
        return m.fc(feat) if hasattr(m, "fc") else feat

    __call__ = forward


class AcceleratedViT:
    """Runs a ``torchvision.models.VisionTransformer`` (eval mode) on the ViT tower's kernels (patch GEMM, LayerNorm,
    tcgen05 attention, MLP GEMMs with bias / GELU / residual epilogues: csrc/vit_forward.cu ``slb_vit_trunk``) and calls
    the forward hooks registered on its encoder blocks with the residual stream as a ``(B, T, W)`` fp32 tensor (a view of
    the workspace: hooks must consume it before the next block runs, which stream order guarantees for K1 / K2).
    Supported hook points: ``encoder.layers.encoder_layer_<i>`` and ``encoder.layers``."""

    def __init__(self, model: nn.Module, device=None, plane_format: str = "f16"):
        lib = N.load(require_device=True)
        need = ("conv_proj", "class_token", "encoder", "image_size", "patch_size", "hidden_dim", "mlp_dim")
        if not all(hasattr(model, n) for n in need) or not hasattr(model.encoder, "layers"):
            raise NotImplementedError("accelerated forward: expected a torchvision-style VisionTransformer")
        if model.training:
            raise NotImplementedError("accelerated forward: put the model in eval mode (dropout must be off)")
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise N.SlbError("the accelerated forward runs on a CUDA device (there is no CPU fallback)")
        self.fmt = {"f16": N.PLANE_F16, "bf16": N.PLANE_BF16}[plane_format]
        dev, fmt = self.device, self.fmt
        self._keep: list = []
        self._ws = None
        blocks = list(model.encoder.layers.children())
        W, P, S = model.hidden_dim, model.patch_size, model.image_size
        heads = blocks[0].self_attention.num_heads
        if W % 64 or model.mlp_dim % 64 or P % 2 or S % P:
            raise NotImplementedError("accelerated forward: width and mlp must be multiples of 64, the patch size even")

        def vec(t):
            t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            self._keep.append(t)
            return t.data_ptr()

        def planes(mat):
            t = ops.split_planes(mat.detach().to(device=dev, dtype=torch.float32), fmt, N.WEIGHT_PLANE_SCALE)
            self._keep.append(t)
            return t.data_ptr()

        layers = (N.SlbVitLayer * len(blocks))()
        eps = None
        for ly, blk in zip(layers, blocks):
            att, mlp = blk.self_attention, blk.mlp
            lin = [m for m in mlp.children() if isinstance(m, nn.Linear)]
            acts = [m for m in mlp.children() if isinstance(m, nn.GELU)]
            ok = (isinstance(att, nn.MultiheadAttention) and att._qkv_same_embed_dim and att.in_proj_bias is not None and att.batch_first
                  and att.num_heads == heads and len(lin) == 2 and len(acts) == 1 and acts[0].approximate == "none"
                  and lin[0].bias is not None and lin[1].bias is not None)
            if not ok:
                raise NotImplementedError(f"accelerated forward: unsupported encoder block {blk}")
            eps = blk.ln_1.eps if eps is None else eps
            if blk.ln_1.eps != eps or blk.ln_2.eps != eps:
                raise NotImplementedError("accelerated forward: the LayerNorms must share one eps")
            ly.ln1_g, ly.ln1_b = vec(blk.ln_1.weight), vec(blk.ln_1.bias)
            ly.w_qkv, ly.b_qkv = planes(att.in_proj_weight), vec(att.in_proj_bias)
            ly.w_out, ly.b_out = planes(att.out_proj.weight), vec(att.out_proj.bias)
            ly.ln2_g, ly.ln2_b = vec(blk.ln_2.weight), vec(blk.ln_2.bias)
            ly.w_fc, ly.b_fc = planes(lin[0].weight), vec(lin[0].bias)
            ly.w_proj, ly.b_proj = planes(lin[1].weight), vec(lin[1].bias)
        kc, kpad = 3 * P * P, lib.slb_patch_k(P)
        conv = torch.zeros(W, kpad)
        conv[:, :kc] = model.conv_proj.weight.detach().float().cpu().reshape(W, kc)
        w = N.SlbVitWeights()
        w.image_size, w.patch, w.width, w.layers = S, P, W, len(blocks)
        w.heads, w.mlp, w.embed_dim = heads, model.mlp_dim, W
        w.act, w.plane_fmt, w.ln_eps = N.EPI_GELU_ERF, fmt, float(eps)
        w.has_cls, w.pool = 1, N.POOL_CLS
        w.conv_w = planes(conv)
        w.conv_b = vec(model.conv_proj.bias) if model.conv_proj.bias is not None else None
        w.cls = vec(model.class_token.reshape(W))
        w.pos = vec(model.encoder.pos_embedding.reshape(-1, W))
        w.ln_pre_g = w.ln_pre_b = None
        w.layer = layers
        self._keep.append(layers)
        self._struct = w
        self.blocks = blocks
        self.tokens = (S // P) ** 2 + 1
        self._tappable = {id(b) for b in blocks} | {id(model.encoder.layers)}

    def check_hooks(self) -> None:
        for name, m in self.model.named_modules():
            if len(m._forward_hooks) and id(m) not in self._tappable:
                raise NotImplementedError(
                    f"accelerated forward: cannot expose the output of '{name}' ({type(m).__name__}); hook the encoder blocks "
                    "(encoder.layers.encoder_layer_<i>) or run with accelerate=False")

    @torch.no_grad()
    def forward(self, x: torch.Tensor, features: bool = False):
        """x (B, 3, S, S) fp32 on the device. Fires the registered hooks; ``features=True`` returns the final residual stream
        (B, T, W) (a copy), else None. The trunk stops after the last hooked block."""
        self.check_hooks()
        lib = N.load(require_device=True)
        m, w = self.model, self._struct
        if x.ndim != 4 or tuple(x.shape[1:]) != (3, w.image_size, w.image_size):
            raise ValueError(f"expected (B, 3, {w.image_size}, {w.image_size}) images, got {tuple(x.shape)}")
        x = x.to(self.device, dtype=torch.float32).contiguous()
        B = x.shape[0]
        hooked = [i for i, b in enumerate(self.blocks) if len(b._forward_hooks)]
        seq_hooked = len(m.encoder.layers._forward_hooks) > 0
        last = len(self.blocks) if (features or seq_hooked) else (hooked[-1] + 1 if hooked else 0)
        if B == 0 or last == 0:
            return None
        need = lib.slb_vit_workspace_bytes(ctypes.byref(w), B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        stream_x = self._ws[: B * self.tokens * w.width * 4].view(torch.float32).view(B, self.tokens, w.width)
        cuts = sorted(set(i + 1 for i in hooked if i + 1 <= last) | {last})  # run up to each hooked block, fire, continue
        begin = 0
        with torch.cuda.device(x.device):
            for end in cuts:
                rc = lib.slb_vit_trunk(ctypes.byref(w), x.data_ptr(), B, begin, end, self._ws.data_ptr(), self._ws.numel(),
                                       N.stream_ptr(x.device))
                N.check(rc, "slb_vit_trunk")
                blk = self.blocks[end - 1]
                if len(blk._forward_hooks):
                    _fire(blk, stream_x)
                begin = end
        if seq_hooked:
            _fire(m.encoder.layers, stream_x)
        return stream_x.clone() if features else None

    __call__ = forward


def accelerated_forward(model: nn.Module, device=None):
    """The accelerated forward that fits ``model`` (a torchvision-style ResNet or VisionTransformer), or NotImplementedError."""
    if hasattr(model, "layer1") and hasattr(model, "conv1"):
        return AcceleratedResNet(model, device)
    if hasattr(model, "conv_proj") and hasattr(model, "encoder"):
        return AcceleratedViT(model, device)
    raise NotImplementedError(
        f"accelerate=True: no accelerated forward for {type(model).__name__} (torchvision-style ResNets and VisionTransformers "
        "are supported); run with accelerate=False")
