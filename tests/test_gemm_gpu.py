"""K4 tensor-core GEMM (split planes, tcgen05) through the C-ABI vs float64 matmul."""

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from semanticlens_b200 import ops

    return ops


def rel_err(got, want):
    return ((got.double() - want).abs().max() / want.abs().max()).item()


SHAPES = [(128, 128, 64), (128, 128, 128), (256, 384, 768), (100, 136, 192), (1000, 768, 768), (12800, 2304, 768),
          (300, 512, 3072), (5, 8, 64), (129, 72, 64), (257, 128, 64), (512, 1000, 128), (16448, 1024, 256)]


SA, SW = 16.0, 1024.0  # the towers' activation / weight plane scales (powers of two: exact)


# fp16 planes carry 22 bits; all three plane products share ONE fp32 tensor-core accumulator, so the bound is the fp32
# accumulation of 3 K addends: measured 1.2e-5 at K = 3072 (relative to the largest output), 5e-6 at K <= 768.
@pytest.mark.parametrize("fmt,tol", [(0, 2e-5), (1, 6e-5)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_split3_accuracy(ops, M, N, K, fmt, tol):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    out, _ = ops.gemm_split(ops.split_planes(a, fmt, SA), ops.split_planes(w, fmt, SW), alpha=1 / (SA * SW))
    want = a.double() @ w.double().T
    assert rel_err(out, want) < tol


@pytest.mark.parametrize("M,N,K", [(300, 512, 3072), (98, 128, 1152), (1000, 768, 768), (129, 72, 64), (16448, 1024, 256)])
def test_gemm_split_accumulator_mode(ops, M, N, K):
    """SLB_PASSES_SPLIT_ACC: cross terms in their own accumulator -> a third of the truncating adds into the large one."""
    from semanticlens_b200._native import PASSES_SPLIT_ACC

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    ap, wp = ops.split_planes(a, 0, SA), ops.split_planes(w, 0, SW)
    want = a.double() @ w.double().T
    fast, _ = ops.gemm_split(ap, wp, alpha=1 / (SA * SW))
    acc, _ = ops.gemm_split(ap, wp, alpha=1 / (SA * SW), passes=PASSES_SPLIT_ACC)
    assert rel_err(acc, want) < 8e-6
    if K >= 1024:
        assert rel_err(acc, want) < 0.6 * rel_err(fast, want)


def test_gemm_unscaled_planes_lose_only_subnormal_bits(ops):
    """Without the per-tensor scale the lo plane of small weights underflows into fp16 subnormals: still < 1e-5."""
    a = torch.randn(300, 256, device="cuda")
    w = torch.randn(136, 256, device="cuda") * 0.02
    out, _ = ops.gemm_split(ops.split_planes(a), ops.split_planes(w))
    assert rel_err(out, a.double() @ w.double().T) < 1e-5


def test_gemm_single_pass_is_16bit_grade(ops):
    a = torch.randn(256, 256, device="cuda")
    w = torch.randn(256, 256, device="cuda")
    ap, wp = ops.split_planes(a, 1), ops.split_planes(w, 1)
    out1, _ = ops.gemm_split(ap, wp, passes=1)
    want1 = ap[0].double() @ wp[0].double().T  # exactly the hi·hi product
    assert rel_err(out1, want1) < 1e-6
    assert rel_err(out1, a.double() @ w.double().T) > 1e-4


@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_epilogues(ops, epi):
    M, N, K = 300, 264, 128
    a, w = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") * 0.1
    bias, res = torch.randn(N, device="cuda"), torch.randn(M, N, device="cuda")
    rs, cs = torch.rand(M, device="cuda") + 0.5, torch.rand(N, device="cuda") + 0.5
    out, planes = ops.gemm_split(ops.split_planes(a, 0, SA), ops.split_planes(w, 0, SW), bias=bias, residual=res, row_scale=rs,
                                 col_scale=cs, epilogue=epi, alpha=1 / (SA * SW), out_planes=True)
    z = (a.double() @ w.double().T) * rs.double()[:, None] * cs.double()[None] + bias.double()
    act = {0: lambda t: t, 1: lambda t: torch.nn.functional.gelu(t),
           2: lambda t: t * torch.sigmoid(1.702 * t), 3: lambda t: torch.nn.functional.gelu(t, approximate="tanh")}[epi]
    want = act(z) + res.double()
    assert rel_err(out, want) < 3e-6
    # the planes output is the packed truncating split of the fp32 output at the activation scale (tc_common.cuh,
    # slb_split_pair_act_f16): hi = the value with its low 13 mantissa bits cleared, lo = fp16(value - hi)
    x = (out * SA).clamp(-65504.0, 65504.0)
    h = (x.view(torch.int32) & -8192).view(torch.float32)
    assert torch.equal(planes[0], h.half()) and torch.equal(planes[1], (x - h).half())
    assert rel_err((planes[0].double() + planes[1].double()) / SA, out.double()) < 1e-6


def test_gemm_residual_in_place(ops):
    M, N, K = 200, 128, 64
    a, w = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    x = torch.randn(M, N, device="cuda")
    want = x.double() + a.double() @ w.double().T
    ops.gemm_split(ops.split_planes(a, 0, SA), ops.split_planes(w, 0, SW), residual=x, out_f32=x, alpha=1 / (SA * SW))
    assert rel_err(x, want) < 2e-6


def test_gemm_argument_errors(ops):
    from semanticlens_b200._native import SlbError

    a = ops.split_planes(torch.randn(8, 48, device="cuda"))
    w = ops.split_planes(torch.randn(8, 48, device="cuda"))
    with pytest.raises(SlbError, match="multiple of 64"):
        ops.gemm_split(a, w)
