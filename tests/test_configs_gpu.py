"""BASELINE.json configs 3 and 4 as parity-test cases (small batches of the real layer sets): every hooked map is also
captured on the side, fed to the numpy oracle (oracle/collect.py), and the sweep's top-k state must equal the oracle's
bit for bit (values and ids, canonical order)."""

import numpy as np
import pytest
import torch

from oracle import collect as oc

pytestmark = pytest.mark.gpu


def bits_of(t):
    return t.cpu().view(torch.int16).numpy().view(np.uint16)


def sweep_and_check(model, layer_names, agg_fn, op, kind, batches, k):
    from semanticlens_b200.component_visualization.activation_caching import ActMaxCache

    cache = ActMaxCache(layer_names, agg_fn, k)
    seen = {n: [] for n in layer_names}
    mods = dict(model.named_modules())
    taps = [mods[n].register_forward_hook(lambda m, i, o, n=n: seen[n].append(o.detach().float().cpu().numpy())) for n in layer_names]
    with torch.no_grad(), cache.hook_context(model):
        for x in batches:
            model(x)
    for t in taps:
        t.remove()
    for n in layer_names:
        st = oc.sweep(seen[n], op, kind, k)
        am = cache.cache[n]
        assert (bits_of(am.activations) == st.bits).all(), f"{n}: top-k values differ from the oracle"
        assert (am.sample_ids.numpy() == st.ids).all(), f"{n}: top-k ids differ from the oracle"
    return cache


def test_cfg4_resnet50_all_53_conv_layers():
    import torchvision

    from semanticlens_b200.component_visualization import aggregators as A

    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    model = torchvision.models.resnet50(weights=None).eval().cuda()
    layers = [n for n, m in model.named_modules() if isinstance(m, torch.nn.Conv2d)]
    assert len(layers) == 53 and sum(dict(model.named_modules())[n].out_channels for n in layers) == 26560
    g = torch.Generator(device="cuda").manual_seed(1)
    batches = [torch.randn(6, 3, 96, 96, device="cuda", generator=g) for _ in range(3)]
    cache = sweep_and_check(model, layers, A.aggregate_conv_mean, "mean", "conv", batches, 5)
    assert cache.cache["conv1"].activations.shape == (64, 5)


def test_cfg3_vit_b16_all_block_outputs():
    import torchvision

    from semanticlens_b200.component_visualization import aggregators as A

    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    model = torchvision.models.vit_b_16(weights=None).eval().cuda()
    layers = [f"encoder.layers.encoder_layer_{i}" for i in range(12)]
    g = torch.Generator(device="cuda").manual_seed(2)
    batches = [torch.randn(5, 3, 224, 224, device="cuda", generator=g) for _ in range(2)]
    cache = sweep_and_check(model, layers, A.aggregate_transformer_mean, "mean", "tokens", batches, 4)
    assert cache.cache[layers[-1]].activations.shape == (768, 4)


def test_vit_l14_tower_vs_oracle():
    """cfg 4's foundation model at full size, one image (the float64 oracle forward takes a few seconds on the host)."""
    from oracle import vit_port as vp
    from semanticlens_b200.foundation_models import vit

    ocfg = vp.CONFIGS["ViT-L-14"]
    cfg = vit.CONFIGS["ViT-L-14"]
    sd = vp.init_weights(ocfg, seed=3)
    tower = vit.VitTower(cfg, sd, "cuda")
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    want = vp.encode_image(sd, ocfg, img, dtype=torch.float64)
    got = tower.forward(img.cuda()).cpu()
    err = ((got.double() - want).abs().max() / want.abs().max()).item()
    assert err < 1e-4, err
