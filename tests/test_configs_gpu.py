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


class _Cfg1Images(torch.utils.data.Dataset):
    def __init__(self, n, kind):
        g = torch.Generator().manual_seed(11)
        # per-image contrast so that channel means differ between images by far more than a bf16 ulp (pure noise images
        # give near-ties everywhere, and cuDNN / oneDNN then order them differently)
        gain = torch.linspace(0.15, 1.0, n)[torch.randperm(n, generator=g)].view(n, 1, 1, 1)
        self.u8 = (torch.randint(0, 255, (n, 3, 224, 224), generator=g).float() * gain).to(torch.uint8)
        self.kind, self.name = kind, f"cfg1-{kind}-{n}"

    def __len__(self):
        return self.u8.shape[0]

    def __getitem__(self, i):
        return ((self.u8[i].float() / 255 - 0.45) / 0.23, 0) if self.kind == "model" else self.u8[i]


def test_cfg1_resnet18_layer4_rn50_embed_end_to_end(tmp_path):
    """BASELINE configs[0] at a small image count: ResNet-18 (random init) probed at layer4, OpenClip RN50 embed, through
    Lens.compute_concept_db — against the torch-CPU port of the reference path with the oracle ModifiedResNet."""
    import torchvision

    from oracle import ref_port as rp
    from oracle import rn_port as rnp
    from semanticlens_b200.component_visualization import ActivationComponentVisualizer, aggregators
    from semanticlens_b200.foundation_models import OpenClip
    from semanticlens_b200.lens import Lens

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = torchvision.models.resnet18(weights=None).eval()
    net.name = "resnet18"
    n, k = 20, 5
    ds_m, ds_f = _Cfg1Images(n, "model"), _Cfg1Images(n, "fm")
    ocfg = rnp.CONFIGS["RN50"]
    sd = rnp.init_weights(ocfg, seed=9)
    mean = torch.tensor(ocfg.mean).view(1, 3, 1, 1)
    std = torch.tensor(ocfg.std).view(1, 3, 1, 1)
    ref_states = rp.sweep(net, torch.utils.data.DataLoader(ds_m, batch_size=8), ["layer4"], rp.aggregate_conv_mean, k)
    ref_embeds = torch.cat([rnp.encode_image(sd, ocfg, (ds_f.u8[i:i + 5].float() / 255.0 - mean) / std) for i in range(0, n, 5)])
    ref_db = rp.concept_db(ref_states, ref_embeds)

    fm = OpenClip("RN50", device="cuda", state_dict=sd)
    fm.name = "rn50-fm"
    cv = ActivationComponentVisualizer(net.cuda(), ds_m, ds_f, ["layer4"], k, aggregate_fn=aggregators.aggregate_conv_mean,
                                       cache_dir=str(tmp_path))
    cv.show_progress = False
    db = Lens(fm, device="cuda").compute_concept_db(cv, batch_size=8)
    am = cv.actmax_cache.cache["layer4"]
    rv, ri = ref_states["layer4"].activations, ref_states["layer4"].sample_ids
    assert db["layer4"].shape == (512, k, 1024)
    from tests.e2e_contract import check_collect_contract

    batches = [torch.stack([ds_m[i][0] for i in range(a, min(a + 8, n))]) for a in range(0, n, 8)]
    report = check_collect_contract(
        net.cpu(), ["layer4"], batches, "mean", "conv", k,
        {"layer4": (bits_of(am.activations), am.sample_ids.numpy())}, {"layer4": (bits_of(rv), ri.numpy())})
    print(report)
    assert report["layer4"]["rows_checked_exactly"] >= 0.75 * report["layer4"]["rows"]
    agree = am.sample_ids == ri
    err = (db["layer4"][agree] - ref_db["layer4"][agree]).abs().max() / ref_db["layer4"].abs().max()
    assert err < 1e-4
