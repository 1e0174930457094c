"""Host side of the B200 CLIP / SigLIP text towers and their tokenizers.

``slb_text_forward`` (csrc/vit_forward.cu) restates open_clip's ``CLIP.encode_text`` and, for the SigLIP models, the
``TextTransformer`` of ``CustomTextCLIP`` (no causal mask, last-position pooling, biased projection) — what the reference
reaches through ``OpenClip.encode_text`` / ``SigLipV2.encode_text`` (foundation_models/clip.py:120-135, 190-215) for
``Lens.text_probing`` (lens.py:166-203). The CLIP tokenizer (open_clip's ``SimpleTokenizer`` scheme) needs CLIP's BPE
merges file (``bpe_simple_vocab_16e6.txt.gz``), the SigLIP tokenizer a sentencepiece model; both ship with the upstream
packages / hubs and cannot be fetched here: pass ``bpe_path=`` / ``spm_path=`` (or set ``SLB_CLIP_BPE`` /
``SLB_SIGLIP_SPM``); without them ``tokenize`` raises and pre-tokenised ids can be fed to ``encode_text`` directly.
"""

from __future__ import annotations

import ctypes
import gzip
import html
import os
from dataclasses import dataclass
from functools import lru_cache

import torch

from .. import _native as N
from .. import ops
from .vit import _ACT


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
    arch: str = "clip"  # "clip": causal, end-of-text pooling, matrix projection; "siglip": bidirectional, last position, Linear

    @property
    def mlp(self) -> int:
        return 4 * self.width


# text_cfg of open_clip's model_configs (open-clip-torch 3.0.0) for the towers in vit.CONFIGS
TEXT_CONFIGS = {
    "ViT-B-32": TextConfig("ViT-B-32", 77, 49408, 512, 12, 8, 512),
    "ViT-B-32-quickgelu": TextConfig("ViT-B-32-quickgelu", 77, 49408, 512, 12, 8, 512, act="quick_gelu"),
    "ViT-B-16": TextConfig("ViT-B-16", 77, 49408, 512, 12, 8, 512),
    "ViT-B-16-quickgelu": TextConfig("ViT-B-16-quickgelu", 77, 49408, 512, 12, 8, 512, act="quick_gelu"),
    "ViT-L-14": TextConfig("ViT-L-14", 77, 49408, 768, 12, 12, 768),
    "ViT-L-14-quickgelu": TextConfig("ViT-L-14-quickgelu", 77, 49408, 768, 12, 12, 768, act="quick_gelu"),
    "RN50": TextConfig("RN50", 77, 49408, 512, 12, 8, 1024),
    "RN50-quickgelu": TextConfig("RN50-quickgelu", 77, 49408, 512, 12, 8, 1024, act="quick_gelu"),
    "RN101": TextConfig("RN101", 77, 49408, 512, 12, 8, 512),
    "RN101-quickgelu": TextConfig("RN101-quickgelu", 77, 49408, 512, 12, 8, 512, act="quick_gelu"),
}


def _siglip_text(name, vocab, width, layers, heads):
    return TextConfig(name, 64, vocab, width, layers, heads, width, act="gelu_tanh", eps=1e-6, arch="siglip")


# text_cfg of open_clip's SigLIP model configs: context 64, sentencepiece vocabularies (32 000 c4-en for SigLIP, 256 000
# Gemma for SigLIP 2), no_causal_mask, pool_type "last", proj_bias, LayerNorm eps 1e-6
TEXT_CONFIGS.update({
    "ViT-B-16-SigLIP": _siglip_text("ViT-B-16-SigLIP", 32000, 768, 12, 12),
    "ViT-L-16-SigLIP-256": _siglip_text("ViT-L-16-SigLIP-256", 32000, 1024, 24, 16),
    "ViT-B-16-SigLIP2": _siglip_text("ViT-B-16-SigLIP2", 256000, 768, 12, 12),
    "hf-hub:timm/ViT-B-16-SigLIP2": _siglip_text("ViT-B-16-SigLIP2", 256000, 768, 12, 12),
})


def text_state_dict_keys(cfg: TextConfig) -> list[str]:
    keys = ["token_embedding.weight", "positional_embedding", "ln_final.weight", "ln_final.bias"]
    keys += ["text_projection.weight", "text_projection.bias"] if cfg.arch == "siglip" else ["text_projection"]
    for i in range(cfg.layers):
        p = f"transformer.resblocks.{i}."
        keys += [p + s for s in ("ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias",
                                 "attn.out_proj.weight", "attn.out_proj.bias", "ln_2.weight", "ln_2.bias",
                                 "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias")]
    return keys


def random_text_state_dict(cfg: TextConfig, seed: int = 1) -> dict[str, torch.Tensor]:
    """Random text-tower weights in open_clip naming (open_clip's init: embeddings 0.02 / 0.01, attention width**-0.5,
    fc (2*width)**-0.5, depth-scaled output projections)."""
    g = torch.Generator().manual_seed(seed + 7919)
    W, L = cfg.width, cfg.layers

    def rn(*s):
        return torch.randn(*s, generator=g)

    sd = {
        "token_embedding.weight": rn(cfg.vocab, W) * 0.02,
        "positional_embedding": rn(cfg.context, W) * 0.01,
        "ln_final.weight": torch.ones(W),
        "ln_final.bias": torch.zeros(W),
        "text_projection": rn(W, cfg.embed_dim) * W**-0.5,
    }
    if cfg.arch == "siglip":
        sd["text_projection.weight"] = sd.pop("text_projection").T.contiguous()
        sd["text_projection.bias"] = torch.zeros(cfg.embed_dim)
    proj_std = W**-0.5 * (2 * L) ** -0.5
    for i in range(L):
        p = f"transformer.resblocks.{i}."
        sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = torch.ones(W), torch.zeros(W)
        sd[p + "attn.in_proj_weight"] = rn(3 * W, W) * W**-0.5
        sd[p + "attn.in_proj_bias"] = torch.zeros(3 * W)
        sd[p + "attn.out_proj.weight"] = rn(W, W) * proj_std
        sd[p + "attn.out_proj.bias"] = torch.zeros(W)
        sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = torch.ones(W), torch.zeros(W)
        sd[p + "mlp.c_fc.weight"] = rn(cfg.mlp, W) * (2 * W) ** -0.5
        sd[p + "mlp.c_fc.bias"] = torch.zeros(cfg.mlp)
        sd[p + "mlp.c_proj.weight"] = rn(W, cfg.mlp) * proj_std
        sd[p + "mlp.c_proj.bias"] = torch.zeros(W)
    return sd


class TextTower:
    """Device-resident CLIP text tower: fp32 vectors / embeddings, split-plane matrices, cached workspace."""

    def __init__(self, cfg: TextConfig, state_dict: dict[str, torch.Tensor], device):
        self.cfg = cfg
        keys = text_state_dict_keys(cfg)
        # open_clip's CustomTextCLIP (the SigLIP models) keeps its text tower under "text."
        if "text.token_embedding.weight" in state_dict:
            state_dict = {k[5:]: v for k, v in state_dict.items() if k.startswith("text.")}
        missing = [k for k in keys if k not in state_dict]
        if missing:
            raise KeyError(f"state dict is missing {len(missing)} text-tower tensors, e.g. {missing[:3]}")
        self.state_dict = {k: state_dict[k].detach().to(torch.float32).cpu() for k in keys}
        want = {"token_embedding.weight": (cfg.vocab, cfg.width), "positional_embedding": (cfg.context, cfg.width),
                "transformer.resblocks.0.attn.in_proj_weight": (3 * cfg.width, cfg.width)}
        want.update({"text_projection.weight": (cfg.embed_dim, cfg.width)} if cfg.arch == "siglip"
                    else {"text_projection": (cfg.width, cfg.embed_dim)})
        for k, shape in want.items():
            if k in self.state_dict and tuple(self.state_dict[k].shape) != shape:
                raise ValueError(f"text tower '{cfg.name}': {k} has shape {tuple(self.state_dict[k].shape)}, expected {shape}")
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
        N.load(require_device=True)
        cfg, sd, dev, keep = self.cfg, self.state_dict, self._device, self._keep

        def vec(name):
            t = sd[name].to(dev).contiguous()
            keep.append(t)
            return t.data_ptr()

        def planes(mat):
            t = ops.split_planes(mat.to(dev), N.PLANE_F16, N.WEIGHT_PLANE_SCALE)
            keep.append(t)
            return t.data_ptr()

        layers = (N.SlbVitLayer * cfg.layers)()
        for i in range(cfg.layers):
            p = f"transformer.resblocks.{i}."
            ly = layers[i]
            ly.ln1_g, ly.ln1_b = vec(p + "ln_1.weight"), vec(p + "ln_1.bias")
            ly.w_qkv, ly.b_qkv = planes(sd[p + "attn.in_proj_weight"]), vec(p + "attn.in_proj_bias")
            ly.w_out, ly.b_out = planes(sd[p + "attn.out_proj.weight"]), vec(p + "attn.out_proj.bias")
            ly.ln2_g, ly.ln2_b = vec(p + "ln_2.weight"), vec(p + "ln_2.bias")
            ly.w_fc, ly.b_fc = planes(sd[p + "mlp.c_fc.weight"]), vec(p + "mlp.c_fc.bias")
            ly.w_proj, ly.b_proj = planes(sd[p + "mlp.c_proj.weight"]), vec(p + "mlp.c_proj.bias")
        w = N.SlbTextWeights()
        w.context, w.vocab, w.width, w.layers = cfg.context, cfg.vocab, cfg.width, cfg.layers
        w.heads, w.mlp, w.embed_dim = cfg.heads, cfg.mlp, cfg.embed_dim
        w.act, w.plane_fmt, w.ln_eps = _ACT[cfg.act], N.PLANE_F16, cfg.eps
        w.tok_emb, w.pos = vec("token_embedding.weight"), vec("positional_embedding")
        w.ln_final_g, w.ln_final_b = vec("ln_final.weight"), vec("ln_final.bias")
        if cfg.arch == "siglip":
            w.proj, w.proj_b = planes(sd["text_projection.weight"].contiguous()), vec("text_projection.bias")
            w.non_causal = 1
        else:
            w.proj, w.proj_b, w.non_causal = planes(sd["text_projection"].T.contiguous()), None, 0
        w.layer = layers
        keep.append(layers)
        self._struct = w

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """(B, context) integer token ids -> (B, embed_dim) fp32 on the GPU (un-normalised, like open_clip)."""
        lib = N.load(require_device=True)
        if self._struct is None:
            raise N.SlbError("the text tower lives on the CPU: call .to('cuda') first (there is no CPU fallback)")
        cfg = self.cfg
        if tokens.ndim != 2 or tokens.shape[1] != cfg.context:
            raise ValueError(f"expected (B, {cfg.context}) token ids, got {tuple(tokens.shape)}")
        tokens = tokens.to(device=self._device, dtype=torch.int64).contiguous()
        if tokens.numel() and (int(tokens.min()) < 0 or int(tokens.max()) >= cfg.vocab):
            raise IndexError(f"token id out of range for a vocabulary of {cfg.vocab}")
        B = tokens.shape[0]
        out = torch.empty((B, cfg.embed_dim), dtype=torch.float32, device=self._device)
        if B == 0:
            return out
        # pooled position: CLIP = the end-of-text token = the largest id of each sequence (open_clip: text.argmax(dim=-1));
        # SigLIP = the last position (pool_type "last"; the tokenizer pads to the context length). Index plumbing.
        pos = tokens.argmax(dim=-1) if cfg.arch == "clip" else torch.full((B,), cfg.context - 1, device=self._device)
        eot = pos + torch.arange(B, device=self._device) * cfg.context
        need = lib.slb_text_workspace_bytes(ctypes.byref(self._struct), B)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self._device)
        with torch.cuda.device(self._device):
            rc = lib.slb_text_forward(ctypes.byref(self._struct), tokens.data_ptr(), eot.data_ptr(), B, out.data_ptr(),
                                      self._ws.data_ptr(), self._ws.numel(), N.stream_ptr(self._device))
        N.check(rc, "slb_text_forward")
        return out


# ------------------------------------------------------------------------------------------------
# tokenizer (open_clip SimpleTokenizer scheme: lower-cased, whitespace-cleaned text -> byte-level BPE -> <start> ids <end>)
# ------------------------------------------------------------------------------------------------
@lru_cache()
def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(2**8):
        if b not in bs:
            bs.append(b)
            cs.append(2**8 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def _pairs(word):
    return set(zip(word[:-1], word[1:]))


CLIP_MERGES = 49152 - 256 - 2  # merges CLIP's vocabulary uses (the file's header line excluded; later lines are unused)


class SimpleTokenizer:
    """Byte-pair encoding with CLIP's merges file. ``merges``: the list of merge pairs (for tests) or None to read
    ``bpe_path`` (the gzip'ed ``bpe_simple_vocab_16e6.txt.gz`` of open_clip / OpenAI CLIP)."""

    def __init__(self, bpe_path: str | None = None, merges: list[tuple[str, str]] | None = None, context_length: int = 77):
        import regex as re

        if merges is None:
            bpe_path = bpe_path or os.environ.get("SLB_CLIP_BPE")
            if not bpe_path or not os.path.exists(bpe_path):
                raise FileNotFoundError(
                    "CLIP's BPE merges file (bpe_simple_vocab_16e6.txt.gz, shipped with open_clip) is needed to tokenize "
                    "text; pass OpenClip(..., bpe_path=...) or set SLB_CLIP_BPE, or feed token ids to encode_text directly")
            lines = gzip.open(bpe_path).read().decode("utf-8").split("\n")
            merges = [tuple(m.split()) for m in lines[1 : CLIP_MERGES + 1] if m.strip()]
        self.byte_encoder = _bytes_to_unicode()
        vocab = list(self.byte_encoder.values())
        vocab = vocab + [v + "</w>" for v in vocab]
        vocab += ["".join(m) for m in merges]
        vocab += ["<start_of_text>", "<end_of_text>"]
        self.encoder = dict(zip(vocab, range(len(vocab))))
        self.bpe_ranks = dict(zip(merges, range(len(merges))))
        self.cache = {"<start_of_text>": "<start_of_text>", "<end_of_text>": "<end_of_text>"}
        self.pat = re.compile(r"""<start_of_text>|<end_of_text>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
                              re.IGNORECASE)
        self.sot, self.eot = self.encoder["<start_of_text>"], self.encoder["<end_of_text>"]
        self.context_length = context_length
        self.vocab_size = len(vocab)

    def bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        pairs = _pairs(word)
        if not pairs:
            return token + "</w>"
        while True:
            bigram = min(pairs, key=lambda pair: self.bpe_ranks.get(pair, float("inf")))
            if bigram not in self.bpe_ranks:
                break
            first, second = bigram
            new_word, i = [], 0
            while i < len(word):
                try:
                    j = word.index(first, i)
                    new_word.extend(word[i:j])
                    i = j
                except ValueError:
                    new_word.extend(word[i:])
                    break
                if word[i] == first and i < len(word) - 1 and word[i + 1] == second:
                    new_word.append(first + second)
                    i += 2
                else:
                    new_word.append(word[i])
                    i += 1
            word = tuple(new_word)
            if len(word) == 1:
                break
            pairs = _pairs(word)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text: str) -> list[int]:
        import regex as re

        text = html.unescape(html.unescape(text)).strip()
        text = re.sub(r"\s+", " ", text).strip().lower()
        ids: list[int] = []
        for token in re.findall(self.pat, text):
            token = "".join(self.byte_encoder[b] for b in token.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(token).split(" "))
        return ids

    def __call__(self, texts, context_length: int | None = None) -> torch.Tensor:
        """str or list[str] -> (n, context_length) int64, <start> ids <end> zero-padded; long inputs are truncated with
        the last id forced to <end> (open_clip's behaviour)."""
        if isinstance(texts, str):
            texts = [texts]
        T = context_length or self.context_length
        out = torch.zeros(len(texts), T, dtype=torch.long)
        for i, text in enumerate(texts):
            ids = [self.sot] + self.encode(text) + [self.eot]
            if len(ids) > T:
                ids = ids[:T]
                ids[-1] = self.eot
            out[i, : len(ids)] = torch.tensor(ids)
        return out


class SentencePieceTokenizer:
    """SigLIP's text tokenizer as open_clip's ``HFTokenizer`` drives it for the SigLIP configs (``clean="canonicalize"``,
    ``padding="max_length"``, ``truncation=True``): punctuation removed, lower-cased, whitespace collapsed, sentencepiece
    pieces + ``</s>``, padded with the pad id (1 = ``</s>`` in SigLIP's c4-en vocabulary) up to ``context_length``.
    ``spm_path``: the sentencepiece ``.model`` file of the checkpoint (``SLB_SIGLIP_SPM``); not available offline, so only
    the mechanics are tested (tests/test_tokenizer.py trains a toy model). ``lower=False`` for SigLIP 2's Gemma vocabulary
    keeps the case (its config lower-cases through the tokenizer kwargs instead)."""

    def __init__(self, spm_path: str | None = None, context_length: int = 64, pad_id: int = 1, lower: bool = True):
        spm_path = spm_path or os.environ.get("SLB_SIGLIP_SPM")
        if not spm_path or not os.path.exists(spm_path):
            raise FileNotFoundError(
                "SigLIP's sentencepiece model (the checkpoint's spiece.model / tokenizer.model) is needed to tokenize text; "
                "pass OpenClip(..., spm_path=...) or set SLB_SIGLIP_SPM, or feed token ids to encode_text directly")
        import sentencepiece as spm

        self.sp = spm.SentencePieceProcessor(model_file=spm_path)
        self.context_length, self.pad_id, self.lower = context_length, pad_id, lower
        eos = self.sp.eos_id()
        self.eos = eos if eos >= 0 else pad_id
        self.vocab_size = self.sp.get_piece_size()

    @staticmethod
    def canonicalize(text: str, lower: bool = True) -> str:
        import string

        text = text.translate(str.maketrans("", "", string.punctuation))
        text = " ".join(text.split())
        return (text.lower() if lower else text).strip()

    def __call__(self, texts, context_length: int | None = None) -> torch.Tensor:
        if isinstance(texts, str):
            texts = [texts]
        T = context_length or self.context_length
        out = torch.full((len(texts), T), self.pad_id, dtype=torch.long)
        for i, text in enumerate(texts):
            ids = (list(self.sp.encode(self.canonicalize(text, self.lower))) + [self.eos])[:T]
            out[i, : len(ids)] = torch.tensor(ids, dtype=torch.long)
        return out
