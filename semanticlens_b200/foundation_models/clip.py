"""``OpenClip`` — drop-in for the reference wrapper (semanticlens/foundation_models/clip.py:27-187) whose image
side runs on the B200 kernels.

The reference delegates to ``open_clip.create_model_and_transforms(url, **kwargs)``; ``open_clip`` is a third-party
dependency that is not vendored (and there is no network here), so this class restates what that call provides for
the image path: the model config of ``url``, the eval transform (Resize(bicubic) -> CenterCrop -> RGB -> ToTensor ->
Normalize with the OpenAI mean/std) and the ``VisionTransformer`` / ``ModifiedResNet`` forward. Weights come from ``state_dict=`` /
``checkpoint_path=`` (open_clip naming, ``visual.*``) or are randomly initialised (``load_weights=False`` in the
reference's own tests does the same). The CLIP text tower (``encode_text``, causal transformer on the same kernels) and
a BPE tokenizer in open_clip's scheme (``tokenize``; needs CLIP's merges file, ``bpe_path=``) serve
``Lens.text_probing``.
"""

from __future__ import annotations

import logging
import os

import numpy as np
import torch

from .. import _native as N
from . import rn
from . import text as text_mod
from . import vit
from .base import AbstractVLM

logger = logging.getLogger(__name__)


class OpenClip(AbstractVLM):
    """CLIP image tower on B200.

    Parameters
    ----------
    url : str
        open_clip model name: "ViT-B-32", "ViT-B-32-quickgelu", "ViT-B-16", "ViT-L-14", "ViT-B-16-SigLIP2",
        "ViT-L-16-SigLIP-256", ... (``vit.CONFIGS``) or a ModifiedResNet "RN50", "RN101" (``rn.CONFIGS``).
    device : str or torch.device
        Where the tower lives; kernels need a CUDA device.
    **kwargs
        ``load_weights`` / ``pretrained`` (accepted like open_clip's; nothing can be downloaded here),
        ``state_dict`` or ``checkpoint_path`` (open_clip-named weights), ``seed`` (random init),
        ``plane_format`` ("f16" default: 22-bit operands, "bf16": 16-bit operands with fp32 range).
    """

    def __init__(self, url, device="cpu", **kwargs):
        if url not in vit.CONFIGS and url not in rn.CONFIGS:
            raise ValueError(f"unknown or unsupported open_clip image tower '{url}' "
                             f"(built: {sorted(vit.CONFIGS) + sorted(rn.CONFIGS)})")
        self.url = url
        is_rn = url in rn.CONFIGS
        self.cfg = rn.CONFIGS[url] if is_rn else vit.CONFIGS[url]
        sd = kwargs.pop("state_dict", None)
        given_sd = sd is not None
        bpe_path = kwargs.pop("bpe_path", None)
        spm_path = kwargs.pop("spm_path", None)
        ckpt = kwargs.pop("checkpoint_path", None)
        seed = kwargs.pop("seed", 1)
        fmt = {"f16": N.PLANE_F16, "bf16": N.PLANE_BF16}[kwargs.pop("plane_format", "f16")]
        pretrained = kwargs.pop("pretrained", None)
        load_weights = kwargs.pop("load_weights", True)
        if kwargs:
            raise TypeError(f"unexpected arguments {sorted(kwargs)}")
        if sd is None and ckpt is not None:
            sd = _load_checkpoint(ckpt)
        if sd is None:
            # open_clip would download here (pretrained=<tag>, or an hf-hub: model id). Nothing can be fetched by this
            # package, and a silently random tower yields a meaningless concept DB: random init must be asked for, the
            # way the reference's own tests do (load_weights=False).
            wants_pretrained = bool(pretrained) or str(url).startswith("hf-hub:")
            if wants_pretrained and load_weights:
                raise ValueError(
                    f"OpenClip('{url}'" + (f", pretrained='{pretrained}'" if pretrained else "") + ") names pretrained "
                    "weights, which this package cannot download: pass state_dict= or checkpoint_path= (open_clip "
                    "naming), or load_weights=False for a randomly initialised tower")
            if not wants_pretrained:
                logger.warning("OpenClip('%s') without pretrained weights: randomly initialised tower (seed %d)", url, seed)
            sd = rn.random_state_dict(self.cfg, seed) if is_rn else vit.random_state_dict(self.cfg, seed)
        self.model = rn.RnTower(self.cfg, sd, device, fmt) if is_rn else vit.VitTower(self.cfg, sd, device, fmt)
        # text side: built lazily (encode_text / tokenize), only for the CLIP towers
        self._text_sd = {k: v for k, v in sd.items() if not k.startswith("visual.")} if ckpt or given_sd else None
        self._text_seed = seed
        self._bpe_path = bpe_path
        self._spm_path = spm_path
        self._text: text_mod.TextTower | None = None
        self._tokenizer: text_mod.SimpleTokenizer | None = None
        self._pin: list = [(None, None), (None, None)]  # (pinned staging buffer, event of its last H2D copy)
        self._pin_next = 0

    def __repr__(self):
        return f"{self.__class__.__name__}(url='{self.url}', model={type(self.model).__name__}[B200])"

    @property
    def resize_mode(self) -> str:
        """open_clip's ``preprocess_cfg["resize_mode"]``: "shortest" for the CLIP towers, "squash" for SigLIP (its
        ``_slpcfg`` / the model configs' preprocess_cfg: Resize((S, S), bicubic) without a crop)."""
        return getattr(self.cfg, "resize_mode", "shortest")

    @property
    def device(self):
        return self.model.device

    def to(self, device):
        self.model.to(device)
        return self.model

    # -- image side -------------------------------------------------------------------------------------
    def encode_image(self, img: torch.Tensor):
        """(B, 3, S, S) preprocessed images -> (B, D) un-normalised features on the GPU (reference :103-118)."""
        with torch.no_grad():
            return self.model.forward(img.to(self.device, non_blocking=True))

    def preprocess(self, img) -> torch.Tensor:
        """PIL image(s) / uint8 (B,3,H,W) tensor(s) -> normalised fp32 batch on ``device`` (reference :137-163).

        RGB PIL images are uploaded as they are and resized / centre-cropped on the GPU (slb_resize_bicubic_u8: Pillow's
        bicubic resample byte for byte); other PIL modes are resized in their own mode by PIL on the host like the
        reference's transform does, then converted. ToTensor + Normalize run on the GPU (K3).
        """
        from .. import ops

        dev = self.device
        if dev.type != "cuda":
            raise N.SlbError("OpenClip.preprocess runs on the GPU: move the model with .to('cuda') first")
        items = img if isinstance(img, (list, tuple)) else [img]
        if items and not isinstance(items[0], torch.Tensor) and os.environ.get("SLB_HOST_RESIZE") != "1":
            return ops.u8_to_f32_norm(self._pil_batch_on_device(items), self.cfg.mean, self.cfg.std)
        u8 = self._to_u8_batch(img)
        if not u8.is_cuda:
            if not u8.is_pinned():
                # two pinned staging buffers used alternately: a buffer is rewritten only after the asynchronous copy
                # that last read it has completed (its event), so the host never races the DMA engine
                slot = self._pin_next
                self._pin_next ^= 1
                buf, ev = self._pin[slot]
                if ev is not None:
                    ev.synchronize()
                if buf is None or buf.numel() < u8.numel():
                    buf = torch.empty(u8.numel(), dtype=torch.uint8, pin_memory=True)
                staged = buf[: u8.numel()].view(u8.shape)
                staged.copy_(u8)
                u8 = staged.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                self._pin[slot] = (buf, ev)
            else:
                u8 = u8.to(dev, non_blocking=True)
        return ops.u8_to_f32_norm(u8, self.cfg.mean, self.cfg.std)

    def _pil_batch_on_device(self, items) -> torch.Tensor:
        """PIL images -> (n, 3, S, S) u8 on the GPU; Resize(S, bicubic) + CenterCrop(S) of RGB images run there."""
        from .. import ops

        S, dev = self.cfg.image_size, self.device
        out = torch.empty((len(items), 3, S, S), dtype=torch.uint8, device=dev)
        for i, im in enumerate(items):
            if im.mode != "RGB":
                out[i].copy_(torch.from_numpy(_pil_to_chw_u8(im, S, self.resize_mode)), non_blocking=True)
                continue
            hwc = torch.from_numpy(np.array(im, dtype=np.uint8)).to(dev, non_blocking=True)
            if im.size == (S, S):
                out[i].copy_(hwc.permute(2, 0, 1))
            else:
                ops.resize_center_crop_u8(hwc, S, out=out[i], squash=self.resize_mode == "squash")
        return out

    def _to_u8_batch(self, img) -> torch.Tensor:
        S = self.cfg.image_size
        if isinstance(img, torch.Tensor):
            t = img if img.ndim == 4 else img.unsqueeze(0)
            if t.dtype != torch.uint8 or t.shape[1] != 3 or t.shape[2] != S or t.shape[3] != S:
                raise ValueError(f"tensor inputs must be uint8 (B, 3, {S}, {S}); got {t.dtype} {tuple(t.shape)}")
            return t.contiguous()
        items = img if isinstance(img, (list, tuple)) else [img]
        if items and isinstance(items[0], torch.Tensor):
            return self._to_u8_batch(torch.stack(list(items)))
        out = np.empty((len(items), 3, S, S), dtype=np.uint8)
        for i, im in enumerate(items):
            out[i] = _pil_to_chw_u8(im, S, self.resize_mode)
        return torch.from_numpy(out)

    # -- text side (Lens.text_probing) --------------------------------------------------------
    def _text_tower(self) -> "text_mod.TextTower":
        if self._text is None:
            tcfg = text_mod.TEXT_CONFIGS.get(self.url)
            if tcfg is None:
                raise NotImplementedError(f"no text tower is built for '{self.url}'")
            sd = self._text_sd
            if not sd or not ("token_embedding.weight" in sd or "text.token_embedding.weight" in sd):
                sd = text_mod.random_text_state_dict(tcfg, self._text_seed)
            self._text = text_mod.TextTower(tcfg, sd, self.device)
        return self._text.to(self.device)

    def encode_text(self, text_input: torch.Tensor):
        """(B, context_length) token ids -> (B, D) un-normalised text features on the GPU (reference :120-135)."""
        with torch.no_grad():
            return self._text_tower().forward(text_input)

    def tokenize(self, txt, context_length=None):
        """str or list[str] -> (n, context_length) int64 token ids on ``device`` (reference :165-187). Needs the
        vocabulary file of the model family: CLIP's BPE merges (``bpe_path=`` / ``SLB_CLIP_BPE``) or SigLIP's sentencepiece
        model (``spm_path=`` / ``SLB_SIGLIP_SPM``)."""
        if self._tokenizer is None:
            tcfg = text_mod.TEXT_CONFIGS.get(self.url)
            if tcfg is None:
                raise NotImplementedError(f"no tokenizer is built for '{self.url}'")
            if tcfg.arch == "siglip":
                self._tokenizer = text_mod.SentencePieceTokenizer(self._spm_path, context_length=tcfg.context,
                                                                  lower="SigLIP2" not in tcfg.name)
            else:
                self._tokenizer = text_mod.SimpleTokenizer(self._bpe_path, context_length=tcfg.context)
        return self._tokenizer(txt, context_length).to(self.device)


class SigLipV2(OpenClip):
    """SigLIP 2 ViT-B/16 image tower on B200 — drop-in for the reference's ``SigLipV2`` (clip.py:190-215:
    ``OpenClip("hf-hub:timm/ViT-B-16-SigLIP2")``): biased patch conv, no class token, attention-pool head, mean = std =
    0.5 preprocessing. Weights: ``state_dict=`` / ``checkpoint_path=`` in open_clip naming (``visual.trunk.*``) or random."""

    URL = "hf-hub:timm/ViT-B-16-SigLIP2"

    def __init__(self, device="cpu", **kwargs):
        super().__init__(url=self.URL, device=device, **kwargs)


def _pil_to_chw_u8(im, S: int, resize_mode: str = "shortest") -> np.ndarray:
    """open_clip eval transform up to (not including) ToTensor. resize_mode "shortest" (CLIP): Resize(S, bicubic) on the
    shorter side, CenterCrop(S); "squash" (SigLIP): Resize((S, S), bicubic), aspect ratio not kept, no crop. Then RGB."""
    from PIL import Image

    w, h = im.size
    if (w, h) != (S, S) and resize_mode == "squash":
        im = im.resize((S, S), Image.BICUBIC)
    elif (w, h) != (S, S):
        if w <= h:
            nw, nh = S, max(S, int(S * h / w))
        else:
            nw, nh = max(S, int(S * w / h)), S
        if (nw, nh) != (w, h):
            im = im.resize((nw, nh), Image.BICUBIC)
        left, top = int(round((nw - S) / 2.0)), int(round((nh - S) / 2.0))
        im = im.crop((left, top, left + S, top + S))
    im = im.convert("RGB")
    return np.asarray(im, dtype=np.uint8).transpose(2, 0, 1)


def _load_checkpoint(path) -> dict[str, torch.Tensor]:
    path = str(path)
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path)
    sd = torch.load(path, map_location="cpu", weights_only=True)
    return sd.get("state_dict", sd)
