"""CPU checks of the ModifiedResNet oracle (oracle/rn_port.py) and of the host-side weight layout the B200 tower uses."""

import torch
import torch.nn.functional as F

from oracle import rn_port as rp


def test_block_plan_matches_open_clip_rn50():
    plan = rp.block_plan(rp.CONFIGS["RN50"])
    assert len(plan) == 16
    assert [p for p, *_ in plan][:4] == ["visual.layer1.0.", "visual.layer1.1.", "visual.layer1.2.", "visual.layer2.0."]
    # only the first bottleneck of a stage changes resolution / width and carries a downsample branch
    assert [(s, ds) for _p, _i, _pl, s, ds in plan if ds] == [(1, True), (2, True), (2, True), (2, True)]
    assert plan[-1][1:3] == (2048, 512)
    sd = rp.init_weights(rp.CONFIGS["RN50"])
    assert sd["visual.layer4.0.downsample.0.weight"].shape == (2048, 1024, 1, 1)
    assert sd["visual.attnpool.positional_embedding"].shape == (50, 2048)
    assert sd["visual.attnpool.c_proj.weight"].shape == (1024, 2048)


def test_attention_pool_equals_explicit_softmax():
    cfg = rp.CONFIGS["RN-tiny-test"]
    sd = {k: v.double() for k, v in rp.init_weights(cfg).items()}
    x = torch.randn(2, 2048, 2, 2, dtype=torch.float64)
    got = rp.attnpool(sd, cfg, x)
    a = "visual.attnpool."
    tok = x.reshape(2, 2048, 4).permute(0, 2, 1)
    tok = torch.cat([tok.mean(1, keepdim=True), tok], 1) + sd[a + "positional_embedding"]
    q = tok[:, :1] @ sd[a + "q_proj.weight"].T + sd[a + "q_proj.bias"]
    k = tok @ sd[a + "k_proj.weight"].T + sd[a + "k_proj.bias"]
    v = tok @ sd[a + "v_proj.weight"].T + sd[a + "v_proj.bias"]
    H, dh = cfg.heads, 2048 // cfg.heads
    qh, kh, vh = (t.view(2, -1, H, dh).transpose(1, 2) for t in (q, k, v))
    o = (torch.softmax(qh @ kh.transpose(-1, -2) * dh**-0.5, -1) @ vh).transpose(1, 2).reshape(2, 2048)
    want = o @ sd[a + "c_proj.weight"].T + sd[a + "c_proj.bias"]
    assert torch.allclose(got, want, rtol=1e-10, atol=1e-12)


def test_fp32_port_tracks_fp64():
    cfg = rp.CONFIGS["RN-small-test"]
    sd = rp.init_weights(cfg)
    img = torch.randn(2, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(1))
    o32, o64 = rp.encode_image(sd, cfg, img), rp.encode_image(sd, cfg, img, dtype=torch.float64)
    assert o32.shape == (2, cfg.embed_dim)
    assert ((o32.double() - o64).abs().max() / o64.abs().max()).item() < 5e-6


def test_channels_last_im2col_layout_reproduces_conv2d():
    """The tower multiplies im2col rows, column (ky*3 + kx)*C + c, with weights permuted to (cout, ky, kx, cin), and
    folds BatchNorm into a scale and shift: restated with torch on the CPU, that equals conv2d + batch_norm."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 8, 6, 5, generator=g, dtype=torch.float64)
    wt = torch.randn(16, 8, 3, 3, generator=g, dtype=torch.float64)
    gamma, beta, mean, var = (torch.randn(16, generator=g, dtype=torch.float64) for _ in range(4))
    var = var.abs() + 0.5
    cols = F.unfold(x, 3, padding=1).view(2, 8, 9, 30).permute(0, 3, 2, 1).reshape(60, 72)
    mat = wt.permute(0, 2, 3, 1).reshape(16, 72)
    scale = gamma / torch.sqrt(var + rp.BN_EPS)
    got = (cols @ mat.T) * scale + (beta - mean * scale)
    want = F.batch_norm(F.conv2d(x, wt, padding=1), mean, var, gamma, beta, False, 0.0, rp.BN_EPS)
    assert torch.allclose(got, want.permute(0, 2, 3, 1).reshape(60, 16), rtol=1e-10, atol=1e-10)


def test_tower_host_config_matches_oracle_config():
    from semanticlens_b200.foundation_models import rn

    for name in ("RN50", "RN101"):
        o, c = rp.CONFIGS[name], rn.CONFIGS[name]
        assert (o.image_size, o.width, o.layers, o.heads, o.embed_dim) == (c.image_size, c.width, c.layers, c.heads, c.embed_dim)
        assert sorted(rn.state_dict_keys(c)) == sorted(rp.init_weights(o))
        assert rn.block_plan(c) == rp.block_plan(o)
        assert abs(rn.flops_per_image(c) - rp.flops_per_image(o)) < 1
    assert len(rn.conv_names(rn.CONFIGS["RN50"])) == 3 + sum(3 * n + 1 for n in (3, 4, 6, 3))


def test_stride1_bottleneck_matches_torchvision_bottleneck():
    """Partial pin of the wiring on an independent implementation: for stride 1, open_clip's Bottleneck (no pooling) is
    torchvision's (conv1x1-BN-ReLU, conv3x3-BN-ReLU, conv1x1-BN, + shortcut [conv1x1-BN when the width changes], ReLU).
    The stride-2 blocks differ by design (average pool instead of a strided convolution) and stay unpinned."""
    import torchvision

    cfg = rp.CONFIGS["RN-tiny-test"]
    sd = {k: v.double() for k, v in rp.init_weights(cfg).items()}
    for prefix, inpl, pl, stride, ds in rp.block_plan(cfg)[:1]:  # layer1.0: 64 -> 256 channels, stride 1, with downsample
        assert stride == 1 and ds
        tv = torchvision.models.resnet.Bottleneck(
            inpl, pl, stride=1,
            downsample=torch.nn.Sequential(torch.nn.Conv2d(inpl, 4 * pl, 1, bias=False), torch.nn.BatchNorm2d(4 * pl))).double().eval()
        tsd = {}
        for i in (1, 2, 3):
            tsd[f"conv{i}.weight"] = sd[f"{prefix}conv{i}.weight"]
            for s in ("weight", "bias", "running_mean", "running_var"):
                tsd[f"bn{i}.{s}"] = sd[f"{prefix}bn{i}.{s}"]
        tsd["downsample.0.weight"] = sd[prefix + "downsample.0.weight"]
        for s in ("weight", "bias", "running_mean", "running_var"):
            tsd[f"downsample.1.{s}"] = sd[f"{prefix}downsample.1.{s}"]
        missing = tv.load_state_dict(tsd, strict=False)
        assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys) and not missing.unexpected_keys
        x = torch.randn(2, inpl, 9, 9, dtype=torch.float64)
        with torch.no_grad():
            want = tv(x)
        assert torch.allclose(rp._bottleneck(x, sd, prefix, stride, ds), want, rtol=1e-12, atol=1e-12)


def _spec_for(cfg, sd):
    from oracle import rn_spec

    return rn_spec.build(cfg.layers, cfg.embed_dim, cfg.heads, cfg.image_size, cfg.width, sd)


def test_second_restatement_agrees_on_keys_stem_blocks_and_pool():
    """oracle/rn_spec.py rebuilds the tower as nn.Modules from the published description; rn_port's state dict must load
    strictly (names and shapes follow from the module nesting) and every stage must agree: the stem, each bottleneck —
    including the stride-2 blocks whose stride is an average pool on both paths — and the attention pool."""
    import pytest

    for name in ("RN-tiny-test", "RN-small-test"):
        cfg = rp.CONFIGS[name]
        sd = rp.init_weights(cfg, seed=11)
        net = _spec_for(cfg, sd)
        img = torch.randn(2, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
        taps = {}
        got = rp.encode_image(sd, cfg, img, dtype=torch.float64, taps=taps)
        with torch.no_grad():
            x = net.stem(img)
            assert torch.allclose(x, taps["stem"], rtol=1e-12, atol=1e-12), f"{name}: stem"
            strides_seen = set()
            for prefix, _inpl, _pl, stride, _ds in rp.block_plan(cfg):
                layer, idx = prefix.split(".")[1:3]
                x = getattr(net, layer)[int(idx)](x)
                assert torch.allclose(x, taps[prefix], rtol=1e-11, atol=1e-11), f"{name}: {prefix} (stride {stride})"
                strides_seen.add(stride)
            assert strides_seen == {1, 2}
            want = net.attnpool(x)
        assert torch.allclose(got, want, rtol=1e-10, atol=1e-11), name
    # a key the module structure does not produce is rejected
    bad = dict(sd)
    bad["visual.layer1.0.downsample.-1.weight"] = torch.zeros(1)
    with pytest.raises(KeyError, match="published module structure"):
        _spec_for(cfg, bad)


def test_second_restatement_agrees_at_rn50_size():
    cfg = rp.CONFIGS["RN50"]
    sd = rp.init_weights(cfg, seed=12)
    img = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    got = rp.encode_image(sd, cfg, img, dtype=torch.float64)
    with torch.no_grad():
        want = _spec_for(cfg, sd)(img)
    assert got.shape == (1, 1024)
    assert ((got - want).abs().max() / want.abs().max()).item() < 1e-11
