"""Pin the torch restatement of open_clip's VisionTransformer (oracle/vit_port.py) against an independent
implementation of the same published architecture: HuggingFace transformers' CLIPVisionModelWithProjection,
weights mapped name by name. (open_clip itself is neither vendored nor installed: parity is otherwise unpinned.)"""

import pytest
import torch

from oracle import vit_port as vp


def hf_model(cfg, sd):
    transformers = pytest.importorskip("transformers")
    hc = transformers.CLIPVisionConfig(
        hidden_size=cfg.width, intermediate_size=cfg.mlp, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
        image_size=cfg.image_size, patch_size=cfg.patch, projection_dim=cfg.embed_dim,
        hidden_act={"gelu": "gelu", "quick_gelu": "quick_gelu"}[cfg.act], layer_norm_eps=cfg.eps,
    )
    m = transformers.CLIPVisionModelWithProjection(hc).eval()
    W = cfg.width
    new = {
        "vision_model.embeddings.class_embedding": sd["visual.class_embedding"],
        "vision_model.embeddings.patch_embedding.weight": sd["visual.conv1.weight"],
        "vision_model.embeddings.position_embedding.weight": sd["visual.positional_embedding"],
        "vision_model.pre_layrnorm.weight": sd["visual.ln_pre.weight"],
        "vision_model.pre_layrnorm.bias": sd["visual.ln_pre.bias"],
        "vision_model.post_layernorm.weight": sd["visual.ln_post.weight"],
        "vision_model.post_layernorm.bias": sd["visual.ln_post.bias"],
        "visual_projection.weight": sd["visual.proj"].T.contiguous(),
    }
    for i in range(cfg.layers):
        p, q = f"visual.transformer.resblocks.{i}.", f"vision_model.encoder.layers.{i}."
        wi, bi = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            new[q + f"self_attn.{n}.weight"] = wi[j * W : (j + 1) * W]
            new[q + f"self_attn.{n}.bias"] = bi[j * W : (j + 1) * W]
        new[q + "self_attn.out_proj.weight"] = sd[p + "attn.out_proj.weight"]
        new[q + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.bias"]
        new[q + "layer_norm1.weight"], new[q + "layer_norm1.bias"] = sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]
        new[q + "layer_norm2.weight"], new[q + "layer_norm2.bias"] = sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]
        new[q + "mlp.fc1.weight"], new[q + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]
        new[q + "mlp.fc2.weight"], new[q + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"]
    missing, unexpected = m.load_state_dict(new, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m


@pytest.mark.parametrize("name", ["ViT-tiny-test", "ViT-small-test"])
def test_oracle_tower_matches_hf_clip(name):
    cfg = vp.CONFIGS[name]
    sd = vp.init_weights(cfg, seed=3)
    m = hf_model(cfg, sd)
    img = torch.randn(3, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        want = m(pixel_values=img).image_embeds
    got = vp.encode_image(sd, cfg, img)
    assert got.shape == (3, cfg.embed_dim)
    assert (got - want).abs().max() <= 2e-5 * want.abs().max()
    # and fp32 oracle vs its own float64 evaluation
    hi = vp.encode_image(sd, cfg, img, dtype=torch.float64)
    assert (got.double() - hi).abs().max() <= 1e-5 * hi.abs().max()


def hf_siglip(cfg, sd):
    transformers = pytest.importorskip("transformers")
    hc = transformers.SiglipVisionConfig(
        hidden_size=cfg.width, intermediate_size=cfg.mlp, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
        image_size=cfg.image_size, patch_size=cfg.patch, hidden_act="gelu_pytorch_tanh", layer_norm_eps=cfg.eps,
    )
    m = transformers.SiglipVisionModel(hc).eval()
    W, t = cfg.width, "visual.trunk."
    new = {
        "vision_model.embeddings.patch_embedding.weight": sd[t + "patch_embed.proj.weight"],
        "vision_model.embeddings.patch_embedding.bias": sd[t + "patch_embed.proj.bias"],
        "vision_model.embeddings.position_embedding.weight": sd[t + "pos_embed"][0],
        "vision_model.post_layernorm.weight": sd[t + "norm.weight"],
        "vision_model.post_layernorm.bias": sd[t + "norm.bias"],
        "vision_model.head.probe": sd[t + "attn_pool.latent"],
        # HF packs the head's projections like nn.MultiheadAttention: [q; k; v]
        "vision_model.head.attention.in_proj_weight": torch.cat([sd[t + "attn_pool.q.weight"], sd[t + "attn_pool.kv.weight"]]),
        "vision_model.head.attention.in_proj_bias": torch.cat([sd[t + "attn_pool.q.bias"], sd[t + "attn_pool.kv.bias"]]),
        "vision_model.head.attention.out_proj.weight": sd[t + "attn_pool.proj.weight"],
        "vision_model.head.attention.out_proj.bias": sd[t + "attn_pool.proj.bias"],
        "vision_model.head.layernorm.weight": sd[t + "attn_pool.norm.weight"],
        "vision_model.head.layernorm.bias": sd[t + "attn_pool.norm.bias"],
        "vision_model.head.mlp.fc1.weight": sd[t + "attn_pool.mlp.fc1.weight"],
        "vision_model.head.mlp.fc1.bias": sd[t + "attn_pool.mlp.fc1.bias"],
        "vision_model.head.mlp.fc2.weight": sd[t + "attn_pool.mlp.fc2.weight"],
        "vision_model.head.mlp.fc2.bias": sd[t + "attn_pool.mlp.fc2.bias"],
    }
    for i in range(cfg.layers):
        p, q = f"{t}blocks.{i}.", f"vision_model.encoder.layers.{i}."
        wi, bi = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            new[q + f"self_attn.{n}.weight"] = wi[j * W : (j + 1) * W]
            new[q + f"self_attn.{n}.bias"] = bi[j * W : (j + 1) * W]
        new[q + "self_attn.out_proj.weight"], new[q + "self_attn.out_proj.bias"] = sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
        new[q + "layer_norm1.weight"], new[q + "layer_norm1.bias"] = sd[p + "norm1.weight"], sd[p + "norm1.bias"]
        new[q + "layer_norm2.weight"], new[q + "layer_norm2.bias"] = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
        new[q + "mlp.fc1.weight"], new[q + "mlp.fc1.bias"] = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
        new[q + "mlp.fc2.weight"], new[q + "mlp.fc2.bias"] = sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]
    missing, unexpected = m.load_state_dict(new, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m


def test_oracle_siglip_tower_matches_hf_siglip():
    cfg = vp.SIGLIP_CONFIGS["SigLIP-tiny-test"]
    sd = vp.init_siglip_weights(cfg, seed=4)
    m = hf_siglip(cfg, sd)
    img = torch.randn(3, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = m(pixel_values=img).pooler_output
    got = vp.encode_image_siglip(sd, cfg, img)
    assert got.shape == (3, cfg.width)
    assert (got - want).abs().max() <= 2e-5 * want.abs().max()
    hi = vp.encode_image_siglip(sd, cfg, img, dtype=torch.float64)
    assert (got.double() - hi).abs().max() <= 1e-5 * hi.abs().max()


def test_oracle_text_tower_matches_hf_clip_text():
    transformers = pytest.importorskip("transformers")
    cfg = vp.TEXT_CONFIGS["text-tiny-test"]
    sd = vp.init_text_weights(cfg, seed=5)
    hc = transformers.CLIPTextConfig(
        vocab_size=cfg.vocab, hidden_size=cfg.width, intermediate_size=cfg.mlp, num_hidden_layers=cfg.layers,
        num_attention_heads=cfg.heads, max_position_embeddings=cfg.context, projection_dim=cfg.embed_dim, hidden_act="gelu",
        layer_norm_eps=cfg.eps, eos_token_id=cfg.vocab - 1, bos_token_id=cfg.vocab - 2, pad_token_id=0,
    )
    m = transformers.CLIPTextModelWithProjection(hc).eval()
    W = cfg.width
    new = {
        "text_model.embeddings.token_embedding.weight": sd["token_embedding.weight"],
        "text_model.embeddings.position_embedding.weight": sd["positional_embedding"],
        "text_model.final_layer_norm.weight": sd["ln_final.weight"],
        "text_model.final_layer_norm.bias": sd["ln_final.bias"],
        "text_projection.weight": sd["text_projection"].T.contiguous(),
    }
    for i in range(cfg.layers):
        p, q = f"transformer.resblocks.{i}.", f"text_model.encoder.layers.{i}."
        wi, bi = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            new[q + f"self_attn.{n}.weight"] = wi[j * W : (j + 1) * W]
            new[q + f"self_attn.{n}.bias"] = bi[j * W : (j + 1) * W]
        new[q + "self_attn.out_proj.weight"], new[q + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"]
        new[q + "layer_norm1.weight"], new[q + "layer_norm1.bias"] = sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]
        new[q + "layer_norm2.weight"], new[q + "layer_norm2.bias"] = sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]
        new[q + "mlp.fc1.weight"], new[q + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]
        new[q + "mlp.fc2.weight"], new[q + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"]
    missing, unexpected = m.load_state_dict(new, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    # sequences: <bos> words... <eot = largest id> then padding zeros, like the CLIP tokenizer emits them
    g = torch.Generator().manual_seed(2)
    tokens = torch.zeros(4, cfg.context, dtype=torch.int64)
    for b, n in enumerate((3, 7, 10, 1)):
        tokens[b, 0] = cfg.vocab - 2
        tokens[b, 1 : 1 + n] = torch.randint(1, cfg.vocab - 2, (n,), generator=g)
        tokens[b, 1 + n] = cfg.vocab - 1
    with torch.no_grad():
        want = m(input_ids=tokens).text_embeds
    got = vp.encode_text(sd, cfg, tokens)
    assert got.shape == (4, cfg.embed_dim)
    assert (got - want).abs().max() <= 2e-5 * want.abs().max()


def test_oracle_siglip_text_tower_matches_hf_siglip_text():
    """SigLIP text tower (bidirectional blocks, tanh GELU, eps 1e-6, LAST-position pooling, biased projection) against
    HF transformers' SiglipTextModel — the executable stand-in for open_clip's TextTransformer with the SigLIP text_cfg."""
    transformers = pytest.importorskip("transformers")
    cfg = vp.TEXT_CONFIGS["siglip-text-tiny-test"]
    sd = vp.init_text_weights(cfg, seed=6)
    hc = transformers.SiglipTextConfig(
        vocab_size=cfg.vocab, hidden_size=cfg.width, intermediate_size=cfg.mlp, num_hidden_layers=cfg.layers,
        num_attention_heads=cfg.heads, max_position_embeddings=cfg.context, projection_size=cfg.embed_dim,
        hidden_act="gelu_pytorch_tanh", layer_norm_eps=cfg.eps,
    )
    m = transformers.SiglipTextModel(hc).eval()
    W = cfg.width
    new = {
        "text_model.embeddings.token_embedding.weight": sd["token_embedding.weight"],
        "text_model.embeddings.position_embedding.weight": sd["positional_embedding"],
        "text_model.final_layer_norm.weight": sd["ln_final.weight"],
        "text_model.final_layer_norm.bias": sd["ln_final.bias"],
        "text_model.head.weight": sd["text_projection.weight"],
        "text_model.head.bias": sd["text_projection.bias"],
    }
    for i in range(cfg.layers):
        p, q = f"transformer.resblocks.{i}.", f"text_model.encoder.layers.{i}."
        wi, bi = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            new[q + f"self_attn.{n}.weight"] = wi[j * W : (j + 1) * W]
            new[q + f"self_attn.{n}.bias"] = bi[j * W : (j + 1) * W]
        new[q + "self_attn.out_proj.weight"], new[q + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"]
        new[q + "layer_norm1.weight"], new[q + "layer_norm1.bias"] = sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]
        new[q + "layer_norm2.weight"], new[q + "layer_norm2.bias"] = sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]
        new[q + "mlp.fc1.weight"], new[q + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]
        new[q + "mlp.fc2.weight"], new[q + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"]
    missing, unexpected = m.load_state_dict(new, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    # SigLIP's tokenizer pads with id 1 (= </s>) up to the context length: every position is attended, the last is pooled
    g = torch.Generator().manual_seed(3)
    tokens = torch.ones(4, cfg.context, dtype=torch.int64)
    for b, n in enumerate((3, 7, 15, 1)):
        tokens[b, :n] = torch.randint(2, cfg.vocab, (n,), generator=g)
    with torch.no_grad():
        want = m(input_ids=tokens).pooler_output
    got = vp.encode_text(sd, cfg, tokens)
    assert got.shape == (4, cfg.embed_dim)
    assert (got - want).abs().max() <= 2e-5 * want.abs().max()


def test_preprocess_u8_is_totensor_normalize():
    cfg = vp.CONFIGS["ViT-tiny-test"]
    u8 = torch.randint(0, 256, (2, 3, 32, 32), dtype=torch.uint8)
    import torchvision.transforms as T

    tf = T.Compose([T.ToTensor(), T.Normalize(cfg.mean, cfg.std)])
    from PIL import Image

    want = torch.stack([tf(Image.fromarray(im.permute(1, 2, 0).numpy())) for im in u8])
    assert torch.equal(vp.preprocess_u8(cfg, u8), want)


def test_flops_model():
    assert abs(vp.flops_per_image(vp.CONFIGS["ViT-L-14"]) / 1e9 - 162.0) < 3.0
