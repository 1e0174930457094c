"""Embed stage on the GPU through the C-ABI vs the torch oracle tower (oracle/vit_port.py): K3 preprocess,
LayerNorm, attention, and the whole ViT image tower. Tolerance for features: 1e-4 relative (max-norm), the bar
BASELINE.json's north_star states; the achieved error is far below and is asserted at 2e-5."""

import numpy as np
import pytest
import torch

from oracle import vit_port as vp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from semanticlens_b200 import ops

    return ops


ACT = 16.0  # SLB_ACT_PLANE_SCALE: planes written by kernels hold 16 * x as hi + lo


def rel_max(got, want):
    return ((got.double().cpu() - want.double().cpu()).abs().max() / want.double().abs().max()).item()


def test_k3_u8_norm_bitexact_vs_totensor_normalize(ops):
    cfg = vp.CONFIGS["ViT-B-32"]
    u8 = torch.randint(0, 256, (3, 3, 224, 224), dtype=torch.uint8)
    got = ops.u8_to_f32_norm(u8.cuda(), cfg.mean, cfg.std).cpu()
    assert torch.equal(got, vp.preprocess_u8(cfg, u8))


@pytest.mark.parametrize("B,C,H,W", [(1, 1, 4, 4), (2, 3, 16, 16), (5, 2, 48, 36), (3, 3, 256, 256), (2, 3, 384, 384), (7, 3, 20, 12)])
def test_k3_u8_norm_shapes_and_every_byte_value(ops, B, C, H, W):
    """The table look-up kernel on planes of several sizes (one or several slices per plane) and channel counts, with every
    byte value present, against torch's two divisions, bit for bit."""
    g = torch.Generator().manual_seed(H + C)
    u8 = torch.randint(0, 256, (B, C, H, W), dtype=torch.uint8, generator=g)
    u8.view(-1)[: min(256, u8.numel())] = torch.arange(min(256, u8.numel()), dtype=torch.uint8)
    mean, std = (0.48145466, 0.4578275, 0.40821073)[:C], (0.26862954, 0.26130258, 0.27577711)[:C]
    got = ops.u8_to_f32_norm(u8.cuda(), mean, std).cpu()
    want = (u8.float().div(255) - torch.tensor(mean).view(1, C, 1, 1)) / torch.tensor(std).view(1, C, 1, 1)
    assert torch.equal(got, want)


@pytest.mark.parametrize("B,S,P", [(2, 64, 32), (3, 48, 16), (2, 28, 14), (1, 32, 8), (2, 12, 4), (1, 224, 32)])
def test_patchify_bitexact_vs_unfold(ops, B, S, P):
    """The patch-embedding im2col (8 pixels per thread when P % 8 == 0, 2 otherwise) against F.unfold, bit for bit."""
    img = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(S + P))
    cols = torch.nn.functional.unfold(img, P, stride=P).transpose(1, 2).reshape(-1, 3 * P * P)  # (B*g*g, (c, py, px))
    got = ops.patchify(img.cuda(), P)
    want = torch.zeros(cols.shape[0], got.shape[2])
    want[:, : 3 * P * P] = cols
    assert torch.equal(got.cpu(), ops.split_planes(want.cuda(), 0, ACT).cpu())


@pytest.mark.parametrize("rows,cols", [(7, 64), (300, 768), (33, 1024), (5, 1280), (4, 2048), (9, 512)])
def test_layernorm_vs_torch(ops, rows, cols):
    x = torch.randn(rows, cols) * 3 + 0.5
    g, b = torch.randn(cols), torch.randn(cols)
    want = torch.nn.functional.layer_norm(x.double(), (cols,), g.double(), b.double(), 1e-5)
    got = ops.layernorm(x.cuda(), g.cuda(), b.cuda(), 1e-5)
    assert rel_max(got, want) < 2e-6
    planes = ops.layernorm(x.cuda(), g.cuda(), b.cuda(), 1e-5, fmt=0)
    assert rel_max((planes[0].double() + planes[1].double()) / ACT, want) < 2e-6


@pytest.mark.parametrize("B,T,H,dh,gain", [(2, 50, 12, 64, 1.0), (1, 257, 4, 64, 1.0), (3, 17, 2, 32, 1.0), (2, 197, 3, 64, 1.0),
                                          (1, 1, 1, 64, 1.0), (2, 64, 2, 64, 1.0), (1, 65, 2, 64, 3.0), (1, 320, 1, 64, 2.0),
                                          (70, 50, 12, 64, 4.0)])
def test_attention_vs_torch(ops, B, T, H, dh, gain):
    W = H * dh
    qkv = torch.randn(B, T, 3 * W, generator=torch.Generator().manual_seed(T * 7 + H)) * gain
    q, k, v = (t.view(B, T, H, dh).transpose(1, 2).double() for t in qkv.split(W, -1))
    want = (torch.softmax(q @ k.transpose(-1, -2) * dh**-0.5, -1) @ v).transpose(1, 2).reshape(B, T, W)
    got = ops.attention_packed(qkv.cuda(), H)
    tol = 2e-6 * max(1.0, gain)  # fp32 exp of logits with std ~ gain^2: the error of expf grows with |logit|
    assert rel_max(got, want) < tol
    planes = ops.attention_packed(qkv.cuda(), H, fmt=0)
    assert rel_max(((planes[0].double() + planes[1].double()) / ACT).view(B, T, W), want) < tol


@pytest.mark.parametrize("B,T,H,gain", [(2, 50, 12, 1.0), (1, 257, 4, 1.0), (2, 197, 3, 1.0), (1, 1, 1, 1.0), (2, 64, 2, 1.0),
                                       (1, 65, 2, 3.0), (1, 300, 1, 2.0), (70, 50, 12, 4.0), (3, 16, 2, 1.0),
                                       # tcgen05 tiles: exactly one tile, tile + 1-row tail, tile + 64-row second tile,
                                       # three key blocks, many (image, head) pairs
                                       (2, 128, 2, 1.0), (2, 129, 2, 1.0), (1, 192, 3, 2.0), (1, 255, 1, 1.0), (1, 256, 2, 1.0),
                                       (1, 320, 2, 1.0), (9, 257, 16, 1.0), (2, 577, 2, 1.0),
                                       # tails of <= 16 rows split the keys over the warps; 17 rows do not
                                       (2, 140, 2, 1.0), (3, 144, 1, 2.0), (1, 145, 2, 1.0), (2, 263, 3, 1.0),
                                       # tails of <= 4 rows run as fp32 SIMT rows (one CTA per image and head)
                                       (2, 131, 2, 1.0), (1, 260, 3, 2.0), (3, 132, 1, 1.0), (2, 133, 2, 1.0), (5, 385, 4, 1.0),
                                       # short sequences on tcgen05: 128 // T images share a tile (block-diagonal mask); odd
                                       # image counts leave the last group short, 127 rows = one image per tile
                                       (3, 50, 2, 1.0), (5, 64, 1, 2.0), (1, 100, 2, 1.0), (7, 17, 3, 1.0), (4, 127, 2, 1.0),
                                       (9, 42, 2, 3.0), (2, 77, 4, 1.0), (33, 50, 12, 1.0), (1, 50, 1, 1.0), (8, 16, 1, 1.0)])
def test_attention_from_planes_vs_torch(ops, B, T, H, gain):
    """The tower's path: attention reads q | k | v from the in_proj GEMM's split planes (cp.async + ldmatrix + mma.sync)."""
    dh = 64
    W = H * dh
    qkv = torch.randn(B, T, 3 * W, generator=torch.Generator().manual_seed(T * 7 + H)) * gain
    planes_in = ops.split_planes(qkv.view(B * T, 3 * W).cuda(), 0, ACT)
    # reference on the values the planes actually hold (22-bit operands)
    held = ((planes_in[0].double() + planes_in[1].double()) / ACT).view(B, T, 3 * W).cpu()
    q, k, v = (t.view(B, T, H, dh).transpose(1, 2) for t in held.split(W, -1))
    want = (torch.softmax(q @ k.transpose(-1, -2) * dh**-0.5, -1) @ v).transpose(1, 2).reshape(B, T, W)
    tol = 2e-6 * max(1.0, gain)
    got = ops.attention_planes(planes_in, B, H)
    assert rel_max(got, want) < tol
    planes = ops.attention_planes(planes_in, B, H, fmt=0)
    assert rel_max(((planes[0].double() + planes[1].double()) / ACT).view(B, T, W), want) < tol


def tower_for(name, seed=3, fmt=0):
    from semanticlens_b200.foundation_models import vit

    ocfg = vp.CONFIGS[name]
    cfg = vit.VitConfig(**{f: getattr(ocfg, f) for f in ("name", "image_size", "patch", "width", "layers", "heads", "mlp",
                                                           "embed_dim", "act", "eps", "mean", "std")})
    sd = vp.init_weights(ocfg, seed)
    return ocfg, sd, vit.VitTower(cfg, sd, "cuda", fmt)


@pytest.mark.parametrize("name,B", [("ViT-tiny-test", 5), ("ViT-small-test", 9), ("ViT-B-32", 4), ("ViT-B-16", 2)])
def test_vit_tower_vs_oracle(name, B):
    ocfg, sd, tower = tower_for(name)
    img = torch.randn(B, 3, ocfg.image_size, ocfg.image_size, generator=torch.Generator().manual_seed(1))
    got = tower.forward(img.cuda())
    truth = vp.encode_image(sd, ocfg, img, dtype=torch.float64)
    oracle32 = vp.encode_image(sd, ocfg, img)
    e_kernel, e_oracle = rel_max(got, truth), rel_max(oracle32, truth)
    print(f"{name}: kernels vs f64 {e_kernel:.2e}, torch-fp32 oracle vs f64 {e_oracle:.2e}")
    assert got.shape == (B, ocfg.embed_dim)
    assert rel_max(got, oracle32) < 1e-4  # the north-star tolerance
    assert e_kernel < 2e-5


@pytest.mark.parametrize("name,B", [("SigLIP-tiny-test", 5), ("ViT-B-16-SigLIP2", 3)])
def test_siglip_tower_vs_oracle(name, B):
    """SigLIP tower (no class token, biased patch conv, attention-pool head) against the torch oracle pinned to HF."""
    from semanticlens_b200.foundation_models import vit

    ocfg = vp.SIGLIP_CONFIGS[name]
    cfg = vit.VitConfig(ocfg.name, ocfg.image_size, ocfg.patch, ocfg.width, ocfg.layers, ocfg.heads, ocfg.mlp, ocfg.width,
                        act=ocfg.act, eps=ocfg.eps, mean=ocfg.mean, std=ocfg.std, arch="siglip")
    sd = vp.init_siglip_weights(ocfg, seed=4)
    tower = vit.VitTower(cfg, sd, "cuda")
    img = torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(2))
    want = vp.encode_image_siglip(sd, ocfg, img, dtype=torch.float64)
    got = tower.forward(img.cuda())
    assert got.shape == (B, cfg.width)
    assert rel_max(got, want) < 1e-4
    assert rel_max(got, want) < 2e-5  # 22-bit operands through 12 blocks: measured 5.5e-6 on ViT-B/16-SigLIP2


def test_siglip_l16_256_tower_vs_oracle():
    """cfg 3's foundation model at full size (T = 256: the pure tcgen05 attention path, no mma.sync tail), two images
    against the float64 oracle tower."""
    from semanticlens_b200.foundation_models import vit

    ocfg = vp.SIGLIP_CONFIGS["ViT-L-16-SigLIP-256"]
    cfg = vit.CONFIGS["ViT-L-16-SigLIP-256"]
    sd = vp.init_siglip_weights(ocfg, seed=6)
    tower = vit.VitTower(cfg, sd, "cuda")
    img = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(3))
    want = vp.encode_image_siglip(sd, ocfg, img, dtype=torch.float64)
    got = tower.forward(img.cuda())
    assert got.shape == (2, 1024)
    err = rel_max(got, want)
    print(f"ViT-L-16-SigLIP-256: kernels vs f64 {err:.2e}")
    assert err < 1e-4
    # batch composition must not matter (the 128-row query tiles straddle images differently at B = 3)
    img3 = torch.cat([img, img[:1]])
    got3 = tower.forward(img3.cuda())
    assert torch.equal(got3[:2], got) and torch.equal(got3[2], got[0])


def test_siglipv2_wrapper():
    from semanticlens_b200.foundation_models import SigLipV2

    fm = SigLipV2(device="cuda", load_weights=False)
    u8 = torch.randint(0, 255, (2, 3, 224, 224), dtype=torch.uint8)
    x = fm.preprocess(u8)
    assert torch.equal(x.cpu(), (u8.float() / 255.0 - 0.5) / 0.5)
    out = fm.encode_image(x)
    assert out.shape == (2, 768) and out.is_cuda and torch.isfinite(out).all()


def test_vit_tower_bf16_planes_are_16bit_grade():
    ocfg, sd, tower = tower_for("ViT-small-test", fmt=1)
    img = torch.randn(4, 3, ocfg.image_size, ocfg.image_size)
    err = rel_max(tower.forward(img.cuda()), vp.encode_image(sd, ocfg, img, dtype=torch.float64))
    assert err < 5e-4


def test_vit_tower_batch_composition_invariance():
    ocfg, sd, tower = tower_for("ViT-small-test")
    img = torch.randn(6, 3, ocfg.image_size, ocfg.image_size).cuda()
    whole = tower.forward(img)
    parts = torch.cat([tower.forward(img[:1]), tower.forward(img[1:4]), tower.forward(img[4:])])
    assert torch.equal(whole, parts)


def test_openclip_wrapper_api():
    from PIL import Image

    from semanticlens_b200.foundation_models import OpenClip

    fm = OpenClip("ViT-B-32-quickgelu", device="cuda", load_weights=False)
    assert fm.device.type == "cuda"
    x = fm.preprocess(Image.new("RGB", (64, 64), "red"))
    assert x.shape == (1, 3, 224, 224) and x.is_cuda and x.dtype == torch.float32
    xs = fm.preprocess([Image.new("RGB", (300, 224), "blue"), Image.new("RGB", (224, 224), "green")])
    assert xs.shape == (2, 3, 224, 224)
    e = fm.encode_image(xs)
    assert e.ndim == 2 and e.shape == (2, 512) and torch.isfinite(e).all()
    with pytest.raises(ValueError):
        OpenClip("not-a-model")
    t = fm.encode_text(torch.zeros(1, 77, dtype=torch.long))  # image and text features share the embedding dimension
    assert t.shape == (1, 512) and torch.isfinite(t).all()
    with pytest.raises(ValueError, match="token ids"):
        fm.encode_text(torch.zeros(1, 64, dtype=torch.long))  # a SigLIP-length sequence is not a CLIP one


def test_preprocess_matches_reference_transform_on_pil():
    """Resize(bicubic)/CenterCrop/ToTensor/Normalize of the open_clip eval transform, restated with torchvision."""
    import torchvision.transforms as T
    from PIL import Image

    from semanticlens_b200.foundation_models import OpenClip

    fm = OpenClip("ViT-B-32", device="cuda", load_weights=False)
    rng = np.random.default_rng(0)
    ims = [Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for h, w in ((224, 224), (260, 300), (500, 230))]
    tf = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224), T.ToTensor(),
                    T.Normalize(fm.cfg.mean, fm.cfg.std)])
    want = torch.stack([tf(im) for im in ims])
    assert torch.equal(fm.preprocess(ims).cpu(), want)


def test_preprocess_staging_buffers_do_not_race_the_copy_engine():
    """Unpinned host batches go through pinned staging buffers; with the GPU busy (copies lag behind the host), a reused
    buffer must not be overwritten before its H2D copy has completed."""
    from semanticlens_b200.foundation_models import OpenClip

    fm = OpenClip("ViT-B-32", device="cuda", load_weights=False, seed=1)
    g = torch.Generator().manual_seed(0)
    batches = [torch.randint(0, 255, (48, 3, 224, 224), generator=g, dtype=torch.uint8) for _ in range(6)]
    busy = torch.randn(8192, 8192, device="cuda")
    for _ in range(20):
        busy = busy @ busy * 1e-4  # keep the stream occupied so that the async copies queue up
    outs = [fm.preprocess(b) for b in batches]
    torch.cuda.synchronize()
    mean = torch.tensor(fm.cfg.mean).view(1, 3, 1, 1)
    std = torch.tensor(fm.cfg.std).view(1, 3, 1, 1)
    for b, o in zip(batches, outs):
        assert torch.equal(o.cpu(), (b.float() / 255.0 - mean) / std)


@pytest.mark.parametrize("T,H,causal", [(50, 2, False), (64, 1, False), (17, 3, False), (77, 2, True), (12, 2, True), (100, 1, False),
                                        (257, 2, False), (129, 1, False)])
def test_attention_bits_do_not_depend_on_the_batch_composition(ops, T, H, causal):
    """Short sequences share a tcgen05 tile (one slot of 16 / 32 / 64 / 128 rows per image); an image's output must be the same
    bits whichever images it shares the tile with and whichever slot it lands in (the sharded sweep relies on it)."""
    B, W = 7, H * 64
    qkv = torch.randn(B, T, 3 * W, generator=torch.Generator().manual_seed(T + 31 * H))
    planes = ops.split_planes(qkv.view(B * T, 3 * W).cuda(), 0, ACT).view(2, B, T, 3 * W)

    def run(lo, hi):
        sub = planes[:, lo:hi].reshape(2, (hi - lo) * T, 3 * W).contiguous()
        return ops.attention_planes(sub, hi - lo, H, causal=causal).view(hi - lo, T, W)

    whole = run(0, B)
    parts = torch.cat([run(0, 1), run(1, 4), run(4, 6), run(6, 7)])
    assert torch.equal(whole, parts)
    shifted = torch.cat([run(0, 3), run(3, 7)])
    assert torch.equal(whole, shifted)


@pytest.mark.parametrize("B,T,H", [(3, 77, 8), (2, 12, 2), (1, 64, 1), (2, 130, 2), (5, 33, 2), (9, 16, 1), (4, 127, 1)])
def test_causal_attention_from_planes_vs_torch(ops, B, T, H):
    dh = 64
    W = H * dh
    qkv = torch.randn(B, T, 3 * W, generator=torch.Generator().manual_seed(T + H))
    planes_in = ops.split_planes(qkv.view(B * T, 3 * W).cuda(), 0, ACT)
    held = ((planes_in[0].double() + planes_in[1].double()) / ACT).view(B, T, 3 * W).cpu()
    q, k, v = (t.view(B, T, H, dh).transpose(1, 2) for t in held.split(W, -1))
    mask = torch.full((T, T), float("-inf"), dtype=torch.float64).triu(1)
    want = (torch.softmax(q @ k.transpose(-1, -2) * dh**-0.5 + mask, -1) @ v).transpose(1, 2).reshape(B, T, W)
    got = ops.attention_planes(planes_in, B, H, causal=True)
    assert rel_max(got, want) < 2e-6


@pytest.mark.parametrize("name,B", [("text-tiny-test", 5), ("ViT-B-32", 3)])
def test_text_tower_vs_oracle(name, B):
    """CLIP text tower (token + positional embedding, causal blocks, ln_final, EOT pooling, text_projection) against the
    torch oracle pinned to HF CLIPTextModelWithProjection."""
    from semanticlens_b200.foundation_models import text as T

    ocfg = vp.TEXT_CONFIGS[name]
    cfg = T.TextConfig(ocfg.name, ocfg.context, ocfg.vocab, ocfg.width, ocfg.layers, ocfg.heads, ocfg.embed_dim, ocfg.act, ocfg.eps)
    sd = vp.init_text_weights(ocfg, seed=5)
    tower = T.TextTower(cfg, sd, "cuda")
    g = torch.Generator().manual_seed(9)
    tokens = torch.zeros(B, cfg.context, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(1, cfg.context - 2, (1,), generator=g))
        tokens[b, 0] = cfg.vocab - 2
        tokens[b, 1 : 1 + n] = torch.randint(1, cfg.vocab - 2, (n,), generator=g)
        tokens[b, 1 + n] = cfg.vocab - 1
    want = vp.encode_text(sd, ocfg, tokens, dtype=torch.float64)
    got = tower.forward(tokens)
    assert got.shape == (B, cfg.embed_dim) and got.is_cuda
    assert rel_max(got, want) < 2e-5
    with pytest.raises(IndexError):
        tower.forward(torch.full((1, cfg.context), cfg.vocab, dtype=torch.int64))


def test_text_probing_end_to_end_with_token_ids():
    """Lens.text_probing's arithmetic from token ids: encode_text -> similarity_score against an aggregated concept DB."""
    from semanticlens_b200.foundation_models import OpenClip
    from semanticlens_b200.lens import _probe

    fm = OpenClip("ViT-B-32", device="cuda", load_weights=False, seed=1)
    tokens = torch.zeros(4, 77, dtype=torch.int64)
    tokens[:, 0] = 49406
    tokens[:, 1:4] = torch.arange(12).view(4, 3) + 1000
    tokens[:, 4] = 49407
    emb = fm.encode_text(tokens)
    assert emb.shape == (4, 512) and torch.isfinite(emb).all()
    db = torch.randn(40, 512)
    sim = _probe(emb.cpu(), db)
    ref = torch.nn.functional.normalize(emb.cpu(), dim=-1) @ torch.nn.functional.normalize(db, dim=-1).T
    assert (sim - ref).abs().max() < 1e-5
    with pytest.raises(FileNotFoundError, match="BPE"):
        fm.tokenize(["a photo of a dog"])


@pytest.mark.parametrize("name,B", [("siglip-text-tiny-test", 5), ("ViT-B-16-SigLIP", 3)])
def test_siglip_text_tower_vs_oracle(name, B):
    """SigLIP text tower (no causal mask, tanh GELU, eps 1e-6, LAST-position pooling, biased projection; reference
    clip.py:190-215 -> open_clip TextTransformer) against the torch oracle pinned to HF SiglipTextModel. Weights arrive in
    open_clip's CustomTextCLIP naming ("text." prefix)."""
    from semanticlens_b200.foundation_models import text as T

    ocfg = vp.TEXT_CONFIGS[name]
    cfg = T.TextConfig(ocfg.name, ocfg.context, ocfg.vocab, ocfg.width, ocfg.layers, ocfg.heads, ocfg.embed_dim, ocfg.act, ocfg.eps,
                       arch="siglip")
    sd = vp.init_text_weights(ocfg, seed=7)
    tower = T.TextTower(cfg, {"text." + k: v for k, v in sd.items()}, "cuda")
    g = torch.Generator().manual_seed(11)
    tokens = torch.ones(B, cfg.context, dtype=torch.int64)  # padded with </s> = 1 like SigLIP's tokenizer
    for b in range(B):
        n = int(torch.randint(1, cfg.context + 1, (1,), generator=g))
        tokens[b, :n] = torch.randint(2, cfg.vocab, (n,), generator=g)
    want = vp.encode_text(sd, ocfg, tokens, dtype=torch.float64)
    got = tower.forward(tokens)
    assert got.shape == (B, cfg.embed_dim) and got.is_cuda
    assert rel_max(got, want) < 2e-5
    # a CLIP-style (causal, argmax-pooled) evaluation of the same weights must differ: the flags reach the kernels
    import dataclasses
    clip_like = vp.encode_text({**sd, "text_projection": sd["text_projection.weight"].T}, dataclasses.replace(ocfg, arch="clip"),
                               tokens, dtype=torch.float64)
    assert rel_max(got, clip_like + sd["text_projection.bias"].double()) > 1e-3


def _bpe_file(tmp_path):
    from tests.bpe_fixture import WORDS, learn_merges, write_bpe_file

    return str(write_bpe_file(tmp_path / "bpe_simple_vocab_16e6.txt.gz", learn_merges(WORDS, 60), trailing=0))


def test_text_probing_from_strings_with_templates(tmp_path):
    """Lens.text_probing driven from STRINGS on the GPU (reference lens.py:59-121, 166-203): tokenizer (merges file in the
    real format) -> text tower -> template baseline subtraction -> the reference's template-major / (q t) regrouping
    -> similarity_score. Expected values: the float64 oracle tower on the same token ids, regrouped the reference's way."""
    import semanticlens_b200 as sl
    from semanticlens_b200.foundation_models import OpenClip, text as T

    fm = OpenClip("ViT-B-32", device="cuda", load_weights=False, seed=3, bpe_path=_bpe_file(tmp_path))
    lens = sl.Lens(fm, device="cuda")
    queries = ["dog", "red car", "blue house"]
    templates = ["a photo of a {}", "an image of the {}"]
    db = {"layer4": torch.randn(37, 512, generator=torch.Generator().manual_seed(0)),
          "layer3": torch.randn(19, 512, generator=torch.Generator().manual_seed(1))}
    ocfg = vp.TEXT_CONFIGS["ViT-B-32"]
    sd = T.random_text_state_dict(T.TEXT_CONFIGS["ViT-B-32"], 3)

    def oracle_embed(prompts):
        return vp.encode_text(sd, ocfg, fm.tokenize(prompts).cpu(), dtype=torch.float64)

    def cos(a, b):
        return torch.nn.functional.normalize(a, dim=-1) @ torch.nn.functional.normalize(b.double(), dim=-1).T

    # no templates: one embedding per query
    got = lens.text_probing(queries, db)
    want = oracle_embed(queries)
    for layer in db:
        assert got[layer].shape == (3, db[layer].shape[0])
        assert (got[layer].double().cpu() - cos(want, db[layer])).abs().max() < 2e-5
    # templates: prompts are built template-major, regrouped as (query, template) blocks (the reference's quirk), the
    # embedding of each template formatted with "" is subtracted, mean over the template axis
    prompts = [t.format(q) for t in templates for q in queries]
    rows = oracle_embed(prompts).view(len(queries), len(templates), -1)
    base = oracle_embed([t.format("") for t in templates]).view(1, len(templates), -1)
    want_t = (rows - base).mean(1)
    for bs in (None, 4):
        got_t = lens.text_probing(queries, db["layer4"], templates=templates, batch_size=bs)
        assert got_t.shape == (3, 37)
        assert (got_t.double().cpu() - cos(want_t, db["layer4"])).abs().max() < 5e-5
    # the quirk is observable: the query-major reading gives different numbers
    sane = (oracle_embed([t.format(q) for q in queries for t in templates]).view(len(queries), len(templates), -1) - base).mean(1)
    assert (cos(sane, db["layer4"]) - cos(want_t, db["layer4"])).abs().max() > 1e-3
    # a single string query
    one = lens.text_probing("dog", db["layer4"], templates=templates)
    assert one.shape == (1, 37)


def test_siglip_wrapper_encodes_text():
    """SigLipV2.encode_text no longer raises: SigLIP 2 text tower (vocabulary 256 000) from token ids; tokenize needs the
    sentencepiece model and says so."""
    from semanticlens_b200.foundation_models import SigLipV2

    fm = SigLipV2(device="cuda", load_weights=False, seed=2)
    tokens = torch.ones(3, 64, dtype=torch.int64)
    tokens[:, :5] = torch.arange(15).view(3, 5) + 1000
    emb = fm.encode_text(tokens)
    assert emb.shape == (3, 768) and emb.is_cuda and torch.isfinite(emb).all()
    assert (emb[0] - emb[1]).abs().max() > 0
    with pytest.raises(FileNotFoundError, match="sentencepiece"):
        fm.tokenize(["a photo of a dog"])
