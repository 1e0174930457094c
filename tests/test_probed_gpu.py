"""The opt-in accelerated forward of the probed model (semanticlens_b200/probed.py, csrc/convnet.cu): torchvision ResNets on
the package's convolution kernels. Primitive by primitive against torch in float64, whole networks against a float64
forward, hooked maps against the torch forward's, and the end-to-end collect contract with the flag on."""

import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACT, WSC = 16.0, 1024.0


@pytest.fixture(scope="module")
def ops():
    from semanticlens_b200 import ops as o

    return o


def rel_max(got, want):
    want = want.double().cpu()
    return float((got.double().cpu() - want).abs().max() / want.abs().max())


def held(planes, scale):
    return (planes[0].double() + planes[1].double()) / scale


def weight_planes(ops, w):
    """(Cout, Cin, k, k) -> planes (2, Cout, conv_k) with columns (ky, kx, cin)."""
    cout, cin, k, _ = w.shape
    mat = torch.zeros(cout, ops.conv_k(cin, k))
    mat[:, : cin * k * k] = w.permute(0, 2, 3, 1).reshape(cout, -1)
    return ops.split_planes(mat.cuda(), 0, WSC)


@pytest.mark.parametrize("B,C,H,W,k,s,p,cout", [(2, 3, 37, 41, 7, 2, 3, 64), (1, 3, 224, 224, 7, 2, 3, 64), (2, 3, 32, 32, 3, 2, 1, 32),
                                                (1, 5, 19, 23, 5, 1, 2, 8)])
def test_stem_conv_from_nchw(ops, B, C, H, W, k, s, p, cout):
    g = torch.Generator().manual_seed(H + k)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(cout, C, k, k, generator=g) * (C * k * k) ** -0.5
    col = ops.im2col_nchw(x.cuda(), k, s, p)
    got, _ = ops.gemm_split(col, weight_planes(ops, w), alpha=1.0 / (ACT * WSC), passes=4)
    want = F.conv2d(x.double(), w.double(), stride=s, padding=p).permute(0, 2, 3, 1).reshape(-1, cout)
    assert got.shape == want.shape
    assert rel_max(got, want) < 3e-6


@pytest.mark.parametrize("B,H,W,C,s,cout", [(2, 14, 14, 64, 1, 64), (2, 14, 14, 64, 2, 128), (1, 7, 9, 24, 2, 16), (3, 56, 56, 64, 1, 64),
                                            (1, 15, 13, 40, 1, 8)])
def test_conv3x3_strided_over_planes(ops, B, H, W, C, s, cout):
    g = torch.Generator().manual_seed(H * 3 + C + s)
    a = torch.randn(B, H, W, C, generator=g)
    w = torch.randn(cout, C, 3, 3, generator=g) * (9 * C) ** -0.5
    planes = ops.split_planes(a.view(-1, C).cuda(), 0, ACT)
    x_held = held(planes, ACT).view(B, H, W, C).cpu()
    col = ops.im2col3x3_strided(planes, B, H, W, s)
    got, _ = ops.gemm_split(col, weight_planes(ops, w), alpha=1.0 / (ACT * WSC), passes=4)
    want = F.conv2d(x_held.permute(0, 3, 1, 2), w.double(), stride=s, padding=1).permute(0, 2, 3, 1).reshape(-1, cout)
    assert got.shape == want.shape
    assert rel_max(got, want) < 3e-6
    if s == 1:  # the stride-1 case is the CLIP tower's kernel: identical bits
        assert torch.equal(col, ops.im2col3x3(planes, B, H, W))


@pytest.mark.parametrize("B,H,W,C,k,s,cout", [
    (2, 14, 14, 64, 3, 1, 64), (2, 14, 14, 64, 3, 2, 128), (3, 56, 56, 64, 3, 1, 64), (1, 7, 9, 128, 3, 1, 32), (2, 15, 13, 64, 3, 2, 72),
    (5, 7, 7, 512, 3, 1, 512), (2, 28, 28, 256, 1, 2, 512), (1, 9, 7, 64, 1, 2, 8), (2, 8, 8, 64, 1, 1, 256), (3, 5, 5, 64, 3, 1, 64),
    (2, 57, 55, 128, 3, 2, 128),
    # 32-channel maps: a k-block is one filter tap under the 64-byte swizzle (the CLIP ModifiedResNet stem)
    (2, 16, 16, 32, 3, 1, 32), (3, 112, 112, 32, 3, 1, 64), (1, 9, 7, 32, 3, 2, 8), (2, 8, 8, 32, 1, 1, 64), (2, 13, 11, 32, 5, 1, 136),
    (1, 30, 30, 32, 1, 2, 256)])
def test_implicit_gemm_convolution(ops, B, H, W, C, k, s, cout):
    """slb_conv_gemm (TMA im2col-mode A operand, no im2col matrix) against torch's conv2d in float64 on the values the planes
    hold, and bit-identical to the explicit im2col + GEMM path it replaces (64-channel chunks; a 32-channel map's taps are
    accumulated in a different order than the explicit path's 64-column blocks, so there the two agree to rounding)."""
    pad = k // 2
    g = torch.Generator().manual_seed(H * 7 + C + s + k)
    a = torch.randn(B, H, W, C, generator=g)
    w = torch.randn(cout, C, k, k, generator=g) * (k * k * C) ** -0.5
    planes = ops.split_planes(a.view(-1, C).cuda(), 0, ACT)
    x_held = held(planes, ACT).view(B, H, W, C).cpu()
    wp = weight_planes(ops, w)
    got, gotp = ops.conv_gemm(planes, B, H, W, wp, k, s, pad, alpha=1.0 / (ACT * WSC), passes=4, out_planes=True)
    want = F.conv2d(x_held.permute(0, 3, 1, 2), w.double(), stride=s, padding=pad).permute(0, 2, 3, 1).reshape(-1, cout)
    assert got.shape == want.shape
    tol = 3e-6 if k * k * C <= 1152 else 1e-5  # the tensor core truncates at every accumulate: 7e-6 at K = 4608
    assert rel_max(got, want) < tol
    assert rel_max(held(gotp, ACT), want) < tol
    if C % 64:
        if k == 3:
            ref, _ = ops.gemm_split(ops.im2col3x3_strided(planes, B, H, W, s), wp, alpha=1.0 / (ACT * WSC), passes=4)
            assert rel_max(got, ref) < 2e-6
        return
    if k == 3:
        col = ops.im2col3x3_strided(planes, B, H, W, s)
    else:
        col = ops.subsample2_planes(planes, B, H, W) if s == 2 else planes
    ref, _ = ops.gemm_split(col, wp, alpha=1.0 / (ACT * WSC), passes=4)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("M,K,N", [(300, 64, 64), (128, 576, 256), (1000, 256, 72)])
def test_raw_output_rides_along_with_the_fused_epilogue(ops, M, K, N):
    """raw_f32 = the GEMM before column scale / bias / activation / residual, bit-identical to a plain GEMM, while the fused
    outputs are bit-identical to the same call without it (a hooked nn.Conv2d sees its raw output at no extra pass)."""
    g = torch.Generator().manual_seed(M + K)
    a = ops.split_planes(torch.randn(M, K, generator=g).cuda(), 0, ACT)
    w = ops.split_planes((torch.randn(N, K, generator=g) * K**-0.5).cuda(), 0, WSC)
    sc, sh = (torch.rand(N, generator=g) + 0.5).cuda(), torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    kw = dict(alpha=1.0 / (ACT * WSC), passes=4)
    plain, _ = ops.gemm_split(a, w, **kw)
    for epi, r in ((4, None), (5, res), (0, None)):
        want32, wantp = ops.gemm_split(a, w, bias=sh, col_scale=sc, residual=r, epilogue=epi, out_planes=True, **kw)
        raw = torch.full((M, N), float("nan"), device="cuda")
        got32, gotp = ops.gemm_split(a, w, bias=sh, col_scale=sc, residual=r, epilogue=epi, out_planes=True, raw_f32=raw, **kw)
        assert torch.equal(raw, plain) and torch.equal(got32, want32) and torch.equal(gotp, wantp)


@pytest.mark.parametrize("B,H,W,C", [(2, 8, 8, 16), (1, 7, 9, 24), (3, 56, 56, 256)])
def test_subsample2_is_a_pure_move(ops, B, H, W, C):
    a = torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(C))
    planes = ops.split_planes(a.view(-1, C).cuda(), 0, ACT)
    got = ops.subsample2_planes(planes, B, H, W)
    want = planes.view(2, B, H, W, C)[:, :, ::2, ::2].reshape(2, -1, C)
    assert torch.equal(got, want)


@pytest.mark.parametrize("M,C,res,relu", [(37, 64, False, True), (128, 256, True, True), (50, 8, True, False), (9, 2048, False, False)])
def test_affine_act(ops, M, C, res, relu):
    g = torch.Generator().manual_seed(M + C)
    raw, sc, sh = torch.randn(M, C, generator=g), torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    r = torch.randn(M, C, generator=g) if res else None
    want = torch.addcmul(sh, raw, sc)  # one fused multiply-add per element, like the kernel
    want = torch.from_numpy(np.float32(raw.numpy().astype(np.float64) * sc.numpy() + sh.numpy()))  # fma: a single rounding
    if res:
        want = want + r
    if relu:
        want = want.clamp_min(0)
    r_dev = r.cuda() if res else None
    out32, planes = ops.affine_act(raw.cuda(), sc.cuda(), sh.cuda(), residual=r_dev, relu=relu, out_f32=True, out_planes=True)
    assert torch.equal(out32.cpu(), want)
    assert rel_max(held(planes, ACT), want) < 2e-6
    # in place on the shortcut buffer
    if res:
        out2, _ = ops.affine_act(raw.cuda(), sc.cuda(), sh.cuda(), residual=r_dev, relu=relu, out_f32=r_dev, out_planes=False)
        assert out2.data_ptr() == r_dev.data_ptr() and torch.equal(out2.cpu(), want)


@pytest.mark.parametrize("B,H,W,C", [(2, 112, 112, 64), (1, 9, 11, 8), (3, 16, 16, 32)])
def test_bn_relu_maxpool(ops, B, H, W, C):
    g = torch.Generator().manual_seed(H + C)
    raw, sc, sh = torch.randn(B * H * W, C, generator=g), torch.rand(C, generator=g) + 0.5, 0.3 * torch.randn(C, generator=g)
    z = torch.from_numpy(np.float32(raw.numpy().astype(np.float64) * sc.numpy() + sh.numpy())).clamp_min(0)
    want = F.max_pool2d(z.view(B, H, W, C).permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).reshape(-1, C)
    out32, planes = ops.bn_relu_maxpool(raw.cuda(), B, H, W, sc.cuda(), sh.cuda(), want_f32=True)
    assert torch.equal(out32.cpu(), want)
    assert rel_max(held(planes, ACT), want) < 2e-6


def _randomize_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.7 + 0.6 * torch.rand(m.num_features, generator=g)
            m.bias.data = 0.2 * torch.randn(m.num_features, generator=g)
            m.running_mean = 0.2 * torch.randn(m.num_features, generator=g)
            m.running_var = 0.6 + 0.8 * torch.rand(m.num_features, generator=g)
    return net


@pytest.mark.parametrize("arch,B,S", [("resnet18", 4, 224), ("resnet50", 3, 224), ("resnet34", 2, 160), ("resnet50", 2, 97)])
def test_whole_network_against_float64(arch, B, S):
    """Logits of the accelerated forward (global pool + fc in torch on its last map) against the float64 model; the torch
    fp32 forward on the GPU is measured beside it."""
    import torchvision

    from semanticlens_b200.probed import AcceleratedResNet

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = _randomize_bn(getattr(torchvision.models, arch)(weights=None), 1).eval()
    x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = copy.deepcopy(net).double()(x.double())
        net = net.cuda()
        torch32 = net(x.cuda())
    got = AcceleratedResNet(net)(x.cuda(), logits=True)
    e_accel, e_torch = rel_max(got, want), rel_max(torch32, want)
    print(f"{arch} {S}: accelerated {e_accel:.2e}, torch fp32 (cuDNN) {e_torch:.2e} of the largest logit")
    assert got.shape == want.shape
    assert e_accel < 1e-4  # the north-star tolerance
    assert e_accel < 5e-5  # measured 1.7e-5 (ResNet-18), 2.6e-5 (ResNet-50), 3.0e-5 (ResNet-34); torch fp32 on cuDNN: ~1e-6


def test_hooked_maps_match_the_torch_forward():
    """Every supported hook point hands the hook a (B, C, H, W) channels-last tensor equal (to fp32-grade accuracy) to what
    the torch forward gives the same hook; unsupported hook points are refused, not silently skipped."""
    import torchvision

    from semanticlens_b200.probed import AcceleratedResNet

    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    net = _randomize_bn(torchvision.models.resnet50(weights=None), 3).eval().cuda()
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(5)).cuda()
    names = ["conv1", "maxpool", "layer1.0.conv1", "layer1.0.conv2", "layer1.0.conv3", "layer1.0.downsample.0", "layer1.0", "layer1",
             "layer2.0.conv2", "layer2.0.downsample.0", "layer2", "layer3.5", "layer4.2.conv3", "layer4"]
    mods = dict(net.named_modules())
    seen = {}
    taps = [mods[n].register_forward_hook(lambda m, i, o, n=n: seen.setdefault(n, []).append(o.detach().clone())) for n in names]
    with torch.no_grad():
        net(x)
    ref = {n: v[0] for n, v in seen.items()}
    seen.clear()
    fwd = AcceleratedResNet(net)
    assert fwd(x) is None
    for n in names:
        got, want = seen[n][0], ref[n]
        assert got.shape == want.shape, n
        assert got.is_contiguous(memory_format=torch.channels_last) or got.shape[2] * got.shape[3] == 1, n
        # the tensor core truncates at every accumulate (one-sided), which adds up over 50 layers: 3.7e-5 at layer4.2.conv3
        assert rel_max(got, want) < 6e-5, (n, rel_max(got, want))
    for t in taps:
        t.remove()
    # nothing hooked: nothing to do (the reference discards the logits)
    assert fwd(x) is None
    h = net.bn1.register_forward_hook(lambda m, i, o: None)
    with pytest.raises(NotImplementedError, match="bn1"):
        fwd(x)
    h.remove()
    with pytest.raises(NotImplementedError, match="eval"):
        AcceleratedResNet(copy.deepcopy(net).train())


class _Images(torch.utils.data.Dataset):
    name = "probed-accel-images"

    def __init__(self, n, kind, seed=4):
        g = torch.Generator().manual_seed(seed)
        base = torch.rand(n, 3, 7, 7, generator=g)
        up = F.interpolate(base, size=224, mode="bilinear", align_corners=False)
        self.u8 = (up * 255 + torch.randint(-16, 17, up.shape, generator=g)).clamp(0, 255).to(torch.uint8)
        self.kind = kind

    def __len__(self):
        return self.u8.shape[0]

    def __getitem__(self, i):
        return ((self.u8[i].float() / 255 - 0.45) / 0.23, 0) if self.kind == "model" else self.u8[i]


def bits_of(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


@pytest.mark.parametrize("arch,layers", [("resnet18", ["layer4"]), ("resnet50", ["conv1", "layer1", "layer2", "layer3", "layer4"])])
def test_collect_contract_with_the_accelerated_forward(arch, layers, tmp_path):
    """The end-to-end collect contract of SURVEY §8(c) with accelerate=True: every bf16 candidate that differs from the
    reference port's (torch CPU) sits next to a rounding midpoint given the measured fp32 noise; everything else is exact."""
    import torchvision

    from oracle import ref_port as rp
    from semanticlens_b200.component_visualization import ActivationComponentVisualizer, aggregators
    from semanticlens_b200.probed import AcceleratedResNet
    from tests.e2e_contract import check_collect_contract

    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    net = torchvision.models.__dict__[arch](weights=None).eval()
    net.name = arch
    n, k, bs = 24, 5, 8
    ds_m, ds_f = _Images(n, "model"), _Images(n, "fm")
    ref_states = rp.sweep(net, torch.utils.data.DataLoader(ds_m, batch_size=bs), layers, rp.aggregate_conv_mean, k)
    cv = ActivationComponentVisualizer(copy.deepcopy(net).cuda(), ds_m, ds_f, layers, k, aggregate_fn=aggregators.aggregate_conv_mean,
                                       accelerate=True)
    cv.show_progress = False
    cv.run(batch_size=bs)
    assert isinstance(cv._accel_forward, AcceleratedResNet)
    gpu = {name: (bits_of(cv.actmax_cache.cache[name].activations), cv.actmax_cache.cache[name].sample_ids.numpy()) for name in layers}
    ref = {name: (bits_of(ref_states[name].activations), ref_states[name].sample_ids.numpy()) for name in layers}
    batches = [torch.stack([ds_m[i][0] for i in range(a, min(a + bs, n))]) for a in range(0, n, bs)]
    report = check_collect_contract(net, layers, batches, "mean", "conv", k, gpu, ref, gpu_forward=AcceleratedResNet,
                                    noise_ceiling=1e-4)
    print(report)
    for name in layers:
        assert report[name]["rows_checked_exactly"] >= 0.5 * report[name]["rows"], (name, report[name])


def test_vit_hooked_blocks_match_float64():
    """torchvision VisionTransformer on the ViT tower's kernels: the residual stream after every hooked encoder block
    against a float64 forward of the same model (the torch fp32 forward is measured beside it)."""
    import torchvision

    from semanticlens_b200.probed import AcceleratedViT, accelerated_forward

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = torchvision.models.vit_b_16(weights=None).eval()
    g = torch.Generator().manual_seed(7)
    for p_ in net.parameters():  # a random init with non-trivial biases / LayerNorm gains
        if p_.ndim == 1:
            p_.data += 0.1 * torch.randn(p_.shape, generator=g)
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(8))
    names = [f"encoder.layers.encoder_layer_{i}" for i in (0, 5, 11)]

    def tap(model, inp, forward=None):
        seen = {}
        mods = dict(model.named_modules())
        hs = [mods[n].register_forward_hook(lambda m, i, o, n=n: seen.__setitem__(n, o.detach().clone())) for n in names]
        with torch.no_grad():
            (forward or model)(inp)
        for h in hs:
            h.remove()
        return seen

    want = tap(copy.deepcopy(net).double(), x.double())
    net = net.cuda()
    t32 = tap(net, x.cuda())
    fwd = accelerated_forward(net)
    assert isinstance(fwd, AcceleratedViT)
    got = tap(net, x.cuda(), fwd)
    for n in names:
        assert got[n].shape == (3, 197, 768) and got[n].dtype == torch.float32
        e_accel, e_torch = rel_max(got[n], want[n]), rel_max(t32[n], want[n])
        print(f"{n}: accelerated {e_accel:.2e}, torch fp32 {e_torch:.2e} of the largest activation")
        assert e_accel < 2e-5, n
    # the trunk stops after the last hooked block; with nothing hooked there is nothing to do
    assert fwd(x.cuda()) is None
    feats = fwd(x.cuda(), features=True)
    assert rel_max(feats, want[names[-1]]) < 2e-5
    h = net.encoder.ln.register_forward_hook(lambda m, i, o: None)
    with pytest.raises(NotImplementedError, match="encoder.ln"):
        fwd(x.cuda())
    h.remove()


def test_vit_collect_with_the_accelerated_forward_matches_the_oracle():
    """cfg 3's sweep with accelerate=True: the collected state is bit-exact against the canonical oracle fed the maps the
    hooks saw, and agrees with the torch-forward sweep wherever the candidates' bf16 values agree."""
    import torchvision

    from oracle import collect as oc
    from semanticlens_b200.component_visualization import ActivationComponentVisualizer, aggregators

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = torchvision.models.vit_b_16(weights=None).eval().cuda()
    net.name = "vit_b_16"
    layers = [f"encoder.layers.encoder_layer_{i}" for i in (0, 6, 11)]
    n, k, bs = 24, 5, 8
    ds_m, ds_f = _Images(n, "model"), _Images(n, "fm")
    seen = {name: [] for name in layers}
    mods = dict(net.named_modules())
    taps = [mods[name].register_forward_hook(lambda m, i, o, name=name: seen[name].append(o.detach().float().cpu().numpy())) for name in layers]
    cv = ActivationComponentVisualizer(net, ds_m, ds_f, layers, k, aggregate_fn=aggregators.aggregate_transformer_mean, accelerate=True)
    cv.show_progress = False
    cv.run(batch_size=bs)
    for t in taps:
        t.remove()
    for name in layers:
        st = oc.sweep(seen[name], "mean", "tokens", k)
        am = cv.actmax_cache.cache[name]
        assert (bits_of(am.activations) == st.bits).all() and (am.sample_ids.numpy() == st.ids).all(), name
