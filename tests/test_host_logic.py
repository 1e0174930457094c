"""Host-side logic of the foundation-model wrappers that needs no GPU: construction on the CPU like the reference's own
tests do (`OpenClip(url, device="cpu")`), the loud no-CPU-fallback error, config tables, resize geometry."""

import pytest
import torch

from oracle import resize_port as rz
from oracle import rn_port as rnp


def test_openclip_constructs_on_cpu_and_refuses_to_compute_there():
    from semanticlens_b200._native import SlbError
    from semanticlens_b200.foundation_models import OpenClip

    for url, tower in (("ViT-B-32", "VitTower"), ("RN50", "RnTower")):
        fm = OpenClip(url, device="cpu", load_weights=False)
        assert fm.device.type == "cpu" and tower in repr(fm)
        if not torch.cuda.is_available():
            with pytest.raises(SlbError, match="no CPU fallback"):
                fm.encode_image(torch.zeros(1, 3, 224, 224))
    with pytest.raises(ValueError, match="unknown or unsupported"):
        OpenClip("not-a-model")
    with pytest.raises(TypeError, match="unexpected arguments"):
        OpenClip("ViT-B-32", device="cpu", load_weights=False, jit=True)


def test_rn_tower_rejects_incomplete_or_unsupported_weights_before_touching_the_gpu():
    from semanticlens_b200.foundation_models import rn

    cfg = rn.CONFIGS["RN50"]
    sd = rn.random_state_dict(cfg, 0)
    assert sorted(sd) == sorted(rn.state_dict_keys(cfg)) == sorted(rnp.init_weights(rnp.CONFIGS["RN50"]))
    del sd["visual.attnpool.c_proj.bias"]
    with pytest.raises(KeyError, match="missing 1 image-tower"):
        rn.RnTower(cfg, sd, "cpu")
    with pytest.raises(ValueError, match="width % 64"):
        rn.RnTower(rn.RnConfig("RN50x4", 288, 80, (4, 6, 10, 6), 40, 640), {}, "cpu")


def test_resize_geometry_matches_torchvision_rule():
    from semanticlens_b200 import ops

    for w, h, S in ((300, 260, 224), (100, 80, 224), (51, 50, 32), (224, 897, 224), (1920, 1080, 224), (224, 224, 224)):
        assert ops.resized_size(w, h, S) == rz.resized_size(w, h, S)
        nw, nh = ops.resized_size(w, h, S)
        assert min(nw, nh) == S and (nw >= S and nh >= S)


def test_text_configs_cover_every_clip_image_tower():
    from semanticlens_b200.foundation_models import rn, text, vit

    for url, cfg in list(vit.CONFIGS.items()) + list(rn.CONFIGS.items()):
        tcfg = text.TEXT_CONFIGS[url]
        assert tcfg.embed_dim == cfg.embed_dim, url
        assert tcfg.arch == ("siglip" if getattr(cfg, "arch", "") == "siglip" else "clip"), url
        assert tcfg.width == 64 * tcfg.heads, url  # the tensor-core attention's head_dim
