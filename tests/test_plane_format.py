"""The split-plane operand format of the tensor-core GEMMs (include/slb200.h), restated in numpy on the
CPU: s*x = hi + lo with both planes fp16 at one power-of-two scale s. These checks document the numerical design the
CUDA kernels implement (the kernels themselves are compared with float64 matmuls in tests/test_gemm_gpu.py)."""

import numpy as np

ACT, WGT = 16.0, 1024.0  # SLB_ACT_PLANE_SCALE, SLB_WEIGHT_PLANE_SCALE


def split(x, s):
    v = np.clip(x.astype(np.float32) * np.float32(s), -65504.0, 65504.0)
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def test_planes_hold_22_bits_and_the_scale_is_exact():
    rng = np.random.default_rng(0)
    for s, scale in ((ACT, 1.0), (WGT, 0.02), (WGT, 1.0 / 768**0.5)):
        x = (rng.standard_normal(200_000) * scale).astype(np.float32)
        hi, lo = split(x, s)
        back = (hi.astype(np.float64) + lo.astype(np.float64)) / s
        big = np.abs(x) > 2.0**-8 * scale
        assert np.max(np.abs(back[big] - x[big]) / np.abs(x[big])) < 2.0**-21  # 11 + 11 significand bits
        assert np.max(np.abs(back - x)) < 2.0**-22 * np.abs(x).max()
        # without the scale the lo plane of typical weights sits in fp16's subnormal range and loses bits
        hi1, lo1 = split(x, 1.0)
        back1 = hi1.astype(np.float64) + lo1.astype(np.float64)
        if scale < 0.1:
            assert np.mean(np.abs(back1 - x)) > 4 * np.mean(np.abs(back - x))


def test_three_plane_products_reach_fp32_grade_in_one_accumulator():
    """hi*hi + hi*lo + lo*hi at a shared scale sum in ONE place; the dropped lo*lo term is 2^-22 relative."""
    rng = np.random.default_rng(1)
    M, N, K = 64, 48, 768
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) * 0.03).astype(np.float32)
    ah, al = (p.astype(np.float64) for p in split(a, ACT))
    wh, wl = (p.astype(np.float64) for p in split(w, WGT))
    alpha = 1.0 / (ACT * WGT)
    want = a.astype(np.float64) @ w.astype(np.float64).T
    three = (ah @ wh.T + ah @ wl.T + al @ wh.T) * alpha
    one = (ah @ wh.T) * alpha
    rel = lambda g: np.max(np.abs(g - want)) / np.max(np.abs(want))
    assert rel(three) < 5e-7      # operand precision: what an exact accumulator would deliver
    assert rel(one) > 1e-4        # a single 16-bit pass is 16-bit grade
    dropped = (al @ wl.T) * alpha
    assert np.max(np.abs(dropped)) / np.max(np.abs(want)) < 2.0**-20


def test_probability_planes_of_the_attention_kernels():
    """P = exp2(s - max) in [0, 1] is split at scale 2^10: the lo plane stays a normal fp16 down to p ~ 2^-14."""
    p = np.exp2(-np.linspace(0, 20, 5000)).astype(np.float32)
    hi, lo = split(p, 1024.0)
    back = (hi.astype(np.float64) + lo.astype(np.float64)) / 1024.0
    assert np.max(np.abs(back - p)) < 2.0**-22  # absolute: rows sum to >= 1, so this is relative to the row sum
    assert np.max(np.abs(back - p)[p > 2.0**-10] / p[p > 2.0**-10]) < 2.0**-21
    tiny = p < 2.0**-14  # below the normal range of the lo plane the absolute error keeps shrinking with the hi plane's ulp
    assert np.max(np.abs(back - p)[tiny]) < 2.0**-34


def test_a_residual_stream_kept_as_planes_loses_one_22_bit_rounding_per_block():
    """SLB_EPI_ADD_RELU_PLANES (the ResNet paths): the shortcut is re-read from the previous block's output planes instead of
    an fp32 copy. hi + lo is exact in fp32 (11 + 11 bits), so the only difference to an fp32 stream is the 2^-22 relative
    truncation per block; over the 16 blocks of a ResNet-50 it stays two orders of magnitude under the 1e-4 parity bar."""
    rng = np.random.default_rng(3)
    x32 = np.abs(rng.standard_normal(100_000)).astype(np.float32)
    x64 = x32.astype(np.float64)
    xp = x32.copy()
    for _ in range(16):
        upd = (rng.standard_normal(x32.shape) * 0.5).astype(np.float32)
        x64 = np.maximum(x64 + upd, 0.0)
        x32 = np.maximum(x32 + upd, np.float32(0))
        hi, lo = split(xp, ACT)
        held = (hi.astype(np.float32) + lo.astype(np.float32)) / np.float32(ACT)  # what the epilogue reconstructs: exact
        assert np.array_equal(held.astype(np.float64), (hi.astype(np.float64) + lo.astype(np.float64)) / ACT)
        xp = np.maximum(held + upd, np.float32(0))
    scale = np.abs(x64).max()
    assert np.abs(x32 - x64).max() / scale < 1e-6
    assert np.abs(xp - x64).max() / scale < 4e-6


def test_ops_helpers_for_plane_shortcuts_on_the_host():
    """ops.planes_to_f32 / ops._plane_residual: pure host logic (which epilogue a shortcut selects, shape checks)."""
    import pytest
    import torch

    from semanticlens_b200 import _native as N
    from semanticlens_b200 import ops

    x = torch.randn(5, 8)
    v = x * ACT
    hi = v.half()
    lo = (v - hi.float()).half()
    planes = torch.stack([hi, lo])
    assert torch.allclose(ops.planes_to_f32(planes), x, rtol=0, atol=2.0**-20 * x.abs().max().item())
    assert ops._plane_residual(N.EPI_RELU, None, torch.float16, 5, 8) == (N.EPI_RELU, None)
    f32 = torch.zeros(5, 8)
    assert ops._plane_residual(N.EPI_ADD_RELU, f32, torch.float16, 5, 8) == (N.EPI_ADD_RELU, f32)
    epi, res = ops._plane_residual(N.EPI_ADD_RELU, planes, torch.float16, 5, 8)
    assert epi == N.EPI_ADD_RELU_PLANES and res is planes
    with pytest.raises(AssertionError):
        ops._plane_residual(N.EPI_RELU, planes, torch.float16, 5, 8)  # a plane shortcut only goes with add + ReLU
    with pytest.raises(AssertionError):
        ops._plane_residual(N.EPI_ADD_RELU, planes, torch.float16, 6, 8)  # wrong shape
