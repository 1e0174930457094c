"""CLIP ModifiedResNet image tower through the C-ABI against the torch oracle (oracle/rn_port.py)."""

import pytest
import torch
import torch.nn.functional as F

from oracle import rn_port as rp

pytestmark = pytest.mark.gpu

ACT, WSC = 16.0, 1024.0


@pytest.fixture(scope="module")
def ops():
    from semanticlens_b200 import ops

    return ops


def rel_max(got, want):
    return ((got.double().cpu() - want.double().cpu()).abs().max() / want.double().abs().max()).item()


def held(planes):
    return (planes[0].double() + planes[1].double()) / ACT


def nhwc_rows(x):  # (B, C, H, W) -> (B*H*W, C)
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).contiguous()


@pytest.mark.parametrize("B,S", [(2, 32), (1, 64), (3, 18)])
def test_im2col_stem_bitexact(ops, B, S):
    img = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(S))
    cols = F.unfold(img, 3, padding=1, stride=2)  # (B, 27 [c, ky, kx], L)
    L = cols.shape[-1]
    want = cols.view(B, 3, 9, L).permute(0, 3, 2, 1).reshape(B * L, 27)  # column (ky*3 + kx)*3 + c
    want = torch.cat([want, torch.zeros(B * L, 64 - 27)], 1)
    got = ops.im2col_stem(img.cuda())
    assert torch.equal(got.cpu(), ops.split_planes(want.cuda(), 0, ACT).cpu())


@pytest.mark.parametrize("B,S,cout", [(2, 32, 32), (1, 64, 64), (3, 18, 32), (2, 224, 32)])
def test_stem_conv_direct(ops, B, S, cout):
    """slb_stem_conv3x3s2 (the stem's first convolution + BatchNorm + ReLU as fp32 FMAs) against F.conv2d in float64 on the
    weights the planes hold, and against the im2col + GEMM path it replaces."""
    from semanticlens_b200 import _native as N

    g = torch.Generator().manual_seed(S + cout)
    img = torch.randn(B, 3, S, S, generator=g)
    wt = torch.randn(cout, 3, 3, 3, generator=g) * (2.0 / 27) ** 0.5
    scale, shift = 1 + 0.1 * torch.randn(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
    mat = torch.zeros(cout, 64)
    mat[:, :27] = wt.permute(0, 2, 3, 1).reshape(cout, -1)
    wp = ops.split_planes(mat.cuda(), 0, WSC)
    w_held = ((wp[0].double() + wp[1].double()) / WSC)[:, :27].view(cout, 3, 3, 3).permute(0, 3, 1, 2).cpu()
    want = torch.relu(F.conv2d(img.double(), w_held, stride=2, padding=1) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1))
    got = ops.stem_conv3x3s2(img.cuda(), wp, scale.cuda(), shift.cuda())
    assert got.shape == (2, B * (S // 2) ** 2, cout)
    assert rel_max(held(got), nhwc_rows(want)) < 1e-6
    via_gemm, _ = ops.gemm_split(ops.im2col_stem(img.cuda()), wp, bias=shift.cuda(), col_scale=scale.cuda(), epilogue=N.EPI_RELU,
                                 alpha=1 / (ACT * WSC), passes=N.PASSES_SPLIT_ACC)
    assert rel_max(held(got), via_gemm) < 3e-6
    parts = torch.cat([ops.stem_conv3x3s2(img[:1].cuda(), wp, scale.cuda(), shift.cuda()),
                       ops.stem_conv3x3s2(img[1:].cuda(), wp, scale.cuda(), shift.cuda())], 1) if B > 1 else got
    assert torch.equal(parts, got)


@pytest.mark.parametrize("B,H,W,C", [(2, 8, 8, 32), (1, 7, 7, 64), (3, 5, 9, 8), (1, 14, 14, 256)])
def test_im2col3x3_moves_plane_bits(ops, B, H, W, C):
    x = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(C + H))
    planes = ops.split_planes(nhwc_rows(x).cuda(), 0, ACT)
    got = ops.im2col3x3(planes, B, H, W)
    K = ops.conv_k(C, 3)
    assert got.shape == (2, B * H * W, K)
    for pl in range(2):
        img = planes[pl].float().view(B, H, W, C).permute(0, 3, 1, 2)  # fp16 -> fp32 is exact
        cols = F.unfold(img, 3, padding=1)  # (B, C*9 [c, tap], HW)
        want = cols.view(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * C)
        assert torch.equal(got[pl, :, : 9 * C].float(), want)
        assert not got[pl, :, 9 * C:].any()


@pytest.mark.parametrize("B,H,W,C", [(2, 8, 8, 64), (1, 14, 6, 256), (3, 2, 2, 8)])
def test_avgpool2_planes(ops, B, H, W, C):
    x = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(H * W))
    planes = ops.split_planes(nhwc_rows(x).cuda(), 0, ACT)
    want = nhwc_rows(F.avg_pool2d(held(planes).view(B, H, W, C).permute(0, 3, 1, 2), 2))
    got = ops.avgpool2_planes(planes, B, H, W)
    assert got.shape == (2, B * (H // 2) * (W // 2), C)
    assert rel_max(held(got), want) < 5e-7


def test_pool_tokens(ops):
    B, HW, C = 3, 49, 256
    g = torch.Generator().manual_seed(0)
    x, pos = torch.randn(B, HW, C, generator=g), torch.randn(HW + 1, C, generator=g) * 0.1
    tok, qry = ops.pool_tokens(x.cuda(), pos.cuda())
    want = torch.cat([x.double().mean(1, keepdim=True), x.double()], 1) + pos.double()
    assert rel_max(held(tok).view(B, HW + 1, C), want) < 5e-7
    assert torch.equal(qry.cpu(), tok.view(2, B, HW + 1, C)[:, :, 0].cpu())


def test_gemm_relu_epilogues(ops):
    from semanticlens_b200 import _native as N

    M, Nn, K = 200, 72, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    a, w = torch.randn(M, K, device="cuda", generator=g), torch.randn(Nn, K, device="cuda", generator=g) * 0.1
    bias, res = torch.randn(Nn, device="cuda", generator=g), torch.randn(M, Nn, device="cuda", generator=g)
    cs = torch.rand(Nn, device="cuda", generator=g) + 0.5
    z = (a.double() @ w.double().T) * cs.double() + bias.double()
    ap, wp = ops.split_planes(a, 0, ACT), ops.split_planes(w, 0, WSC)
    out, _ = ops.gemm_split(ap, wp, bias=bias, col_scale=cs, residual=res, epilogue=N.EPI_RELU, alpha=1 / (ACT * WSC))
    assert rel_max(out, torch.relu(z) + res.double()) < 3e-6
    out, planes = ops.gemm_split(ap, wp, bias=bias, col_scale=cs, residual=res, epilogue=N.EPI_ADD_RELU, alpha=1 / (ACT * WSC),
                                 out_planes=True)
    want = torch.relu(z + res.double())
    assert rel_max(out, want) < 3e-6
    assert (out >= 0).all() and rel_max(held(planes), want) < 3e-6


@pytest.mark.parametrize("M,Nn,K,passes", [(200, 72, 128, 3), (1000, 256, 64, 4), (260, 2048, 512, 4), (37, 8, 576, 4)])
def test_shortcut_as_planes(ops, M, Nn, K, passes):
    """SLB_EPI_ADD_RELU_PLANES: the shortcut handed over as split planes (the residual stream of the ResNet paths) gives exactly
    what the fp32 shortcut holding the same values gives, with and without the fp32 copy of the output."""
    from semanticlens_b200 import _native as N

    g = torch.Generator(device="cuda").manual_seed(M + K)
    a, w = torch.randn(M, K, device="cuda", generator=g), torch.randn(Nn, K, device="cuda", generator=g) * 0.1
    bias, res = torch.randn(Nn, device="cuda", generator=g), torch.randn(M, Nn, device="cuda", generator=g)
    cs = torch.rand(Nn, device="cuda", generator=g) + 0.5
    ap, wp, rp_ = ops.split_planes(a, 0, ACT), ops.split_planes(w, 0, WSC), ops.split_planes(res, 0, ACT)
    res_held = ops.planes_to_f32(rp_)
    assert rel_max(res_held, res) < 1e-6 and torch.equal(res_held.double().cpu(), held(rp_).cpu())
    kw = dict(bias=bias, col_scale=cs, alpha=1 / (ACT * WSC), passes=passes)
    want32, wantp = ops.gemm_split(ap, wp, residual=res_held, epilogue=N.EPI_ADD_RELU, out_planes=True, **kw)
    got32, gotp = ops.gemm_split(ap, wp, residual=rp_, epilogue=N.EPI_ADD_RELU, out_planes=True, **kw)
    assert torch.equal(got32, want32) and torch.equal(gotp, wantp)
    none32, onlyp = ops.gemm_split(ap, wp, residual=rp_, epilogue=N.EPI_ADD_RELU, out_f32=False, out_planes=True, **kw)
    assert none32 is None and torch.equal(onlyp, wantp)
    z = (a.double() @ w.double().T) * cs.double() + bias.double()
    assert rel_max(got32, torch.relu(z + res.double())) < 3e-6


@pytest.mark.parametrize("cin,cout,k,H", [(64, 256, 1, 14), (32, 32, 3, 16), (128, 128, 3, 7)])
def test_conv_bn_relu_as_gemm(ops, cin, cout, k, H):
    """One convolution + eval BatchNorm + ReLU the way slb_rn_forward runs it, against F.conv2d in float64."""
    from semanticlens_b200 import _native as N

    B = 2
    g = torch.Generator().manual_seed(cin + k)
    x = torch.randn(B, cin, H, H, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    gamma, beta = 1 + 0.1 * torch.randn(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
    mean, var = 0.1 * torch.randn(cout, generator=g), 1 + 0.2 * torch.rand(cout, generator=g)
    want = torch.relu(F.batch_norm(F.conv2d(x.double(), wt.double(), padding=k // 2), mean.double(), var.double(), gamma.double(),
                                   beta.double(), False, 0.0, 1e-5))
    planes = ops.split_planes(nhwc_rows(x).cuda(), 0, ACT)
    a = ops.im2col3x3(planes, B, H, H) if k == 3 else planes
    mat = torch.zeros(cout, ops.conv_k(cin, k))
    mat[:, : cin * k * k] = wt.permute(0, 2, 3, 1).reshape(cout, -1)
    scale = gamma.double() / torch.sqrt(var.double() + 1e-5)
    shift = beta.double() - mean.double() * scale
    out, _ = ops.gemm_split(a, ops.split_planes(mat.cuda(), 0, WSC), bias=shift.float().cuda(), col_scale=scale.float().cuda(),
                            epilogue=N.EPI_RELU, alpha=1 / (ACT * WSC), passes=N.PASSES_SPLIT_ACC)
    assert rel_max(out, nhwc_rows(want)) < 3e-6


def rn_tower(name, seed=5):
    from semanticlens_b200.foundation_models import rn

    ocfg = rp.CONFIGS[name]
    cfg = rn.RnConfig(ocfg.name, ocfg.image_size, ocfg.width, ocfg.layers, ocfg.heads, ocfg.embed_dim)
    sd = rp.init_weights(ocfg, seed)
    return ocfg, sd, rn.RnTower(cfg, sd, "cuda")


@pytest.mark.parametrize("name,B", [("RN-tiny-test", 3), ("RN-small-test", 2), ("RN50", 2)])
def test_rn_tower_vs_oracle(name, B):
    ocfg, sd, tower = rn_tower(name)
    img = torch.randn(B, 3, ocfg.image_size, ocfg.image_size, generator=torch.Generator().manual_seed(7))
    truth = rp.encode_image(sd, ocfg, img, dtype=torch.float64)
    oracle32 = rp.encode_image(sd, ocfg, img)
    got = tower.forward(img.cuda())
    e_kernel, e_oracle = rel_max(got, truth), rel_max(oracle32, truth)
    print(f"{name}: kernels vs f64 {e_kernel:.2e}, torch-fp32 oracle vs f64 {e_oracle:.2e}")
    assert got.shape == (B, ocfg.embed_dim)
    assert rel_max(got, oracle32) < 1e-4  # the north-star tolerance
    assert e_kernel < 3e-5


def test_rn_tower_batch_composition_invariance():
    ocfg, sd, tower = rn_tower("RN-tiny-test")
    img = torch.randn(5, 3, ocfg.image_size, ocfg.image_size).cuda()
    whole = tower.forward(img)
    parts = torch.cat([tower.forward(img[:2]), tower.forward(img[2:])])
    assert torch.equal(whole, parts)


def test_openclip_rn50_wrapper():
    from semanticlens_b200.foundation_models import OpenClip

    fm = OpenClip("RN50", device="cuda", load_weights=False)
    u8 = torch.randint(0, 255, (2, 3, 224, 224), dtype=torch.uint8)
    e = fm.encode_image(fm.preprocess(u8))
    assert e.shape == (2, 1024) and e.is_cuda and torch.isfinite(e).all()
    t = fm.encode_text(torch.zeros(1, 77, dtype=torch.long))
    assert t.shape == (1, 1024)
    assert "RnTower" in repr(fm)


def test_rn_argument_errors():
    import ctypes

    from semanticlens_b200 import _native as N
    from semanticlens_b200.foundation_models import rn

    with pytest.raises(ValueError, match="width"):
        rn.RnTower(rn.RnConfig("RN50x4", 288, 80, (4, 6, 10, 6), 40, 640), {}, "cuda")
    ocfg, sd, tower = rn_tower("RN-tiny-test")
    sd2 = dict(sd)
    del sd2["visual.layer2.0.downsample.0.weight"]
    with pytest.raises(KeyError, match="missing"):
        rn.RnTower(tower.cfg, sd2, "cuda")
    with pytest.raises(ValueError, match="expected"):
        tower.forward(torch.zeros(1, 3, 32, 32, device="cuda"))
    lib = N.load(require_device=True)
    img = torch.zeros(1, 3, 64, 64, device="cuda")
    out = torch.empty(1, 256, device="cuda")
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    rc = lib.slb_rn_forward(ctypes.byref(tower._struct), img.data_ptr(), 1, out.data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert rc != 0 and b"workspace" in lib.slb_last_error()
