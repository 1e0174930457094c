"""Host-side wrappers: torch tensors in/out, raw pointers + the current CUDA stream across the C-ABI.

Nothing here computes; every function validates its arguments, allocates outputs with the PyTorch caching
allocator and enqueues one or two libslb200 kernels on ``torch.cuda.current_stream()`` without synchronising.
"""

from __future__ import annotations

import torch

from . import _native as N


def _dev_guard(t: torch.Tensor):
    return torch.cuda.device(t.device)


# ------------------------------------------------------------------------------------------------
# collect
# ------------------------------------------------------------------------------------------------
def describe_map(t: torch.Tensor, reduce_kind: str, op: int | None = None):
    """Classify a hooked activation map for K1 without copying when possible.

    reduce_kind: "conv" (4-D, reduce H*W) or "tokens" (3-D, reduce dim 1).
    Returns (tensor_to_keep_alive, layout, B, C, inner).
    """
    if t.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        t = t.float()  # float64 / integer maps: the kernels read fp32, fp16, bf16
    if reduce_kind != "conv" and op == N.AGG_TOKEN and not t.is_contiguous():
        t = t.contiguous()  # the token select reads one (B, T, F) row per image: a permuted view is copied once
    if reduce_kind == "conv":
        B, C, H, W = t.shape
        if t.is_contiguous():
            return t, N.LAYOUT_NCHW, B, C, H * W
        if t.is_contiguous(memory_format=torch.channels_last):
            return t, N.LAYOUT_BTF, B, C, H * W  # memory is (B, H*W, C)
        t = t.contiguous()
        return t, N.LAYOUT_NCHW, B, C, H * W
    B, T, F = t.shape
    if t.is_contiguous():
        return t, N.LAYOUT_BTF, B, F, T
    if t.transpose(1, 2).is_contiguous():
        return t, N.LAYOUT_NCHW, B, F, T  # memory is (B, F, T)
    t = t.contiguous()
    return t, N.LAYOUT_BTF, B, F, T


def agg_reduce(t: torch.Tensor, op: int, reduce_kind: str, token_pos: int = 0) -> torch.Tensor:
    """K1: (B,C,H,W) / (B,T,F) activation map -> (B, C) fp32 on the same device. No sync."""
    lib = N.load(require_device=True)
    N.require_cuda(t, "activation map")
    t = t.detach()
    keep, layout, B, C, inner = describe_map(t, reduce_kind, op)
    out = torch.empty((B, C), dtype=torch.float32, device=t.device)
    if B == 0 or C == 0:
        return out
    with _dev_guard(t):
        rc = lib.slb_agg_reduce(
            keep.data_ptr(), N.dtype_code(keep.dtype), layout, B, C, inner, op, token_pos, out.data_ptr(),
            N.stream_ptr(t.device),
        )
    N.check(rc, "slb_agg_reduce")
    return out


def topk_update(
    cand: torch.Tensor,
    state_vals: torch.Tensor,
    state_ids: torch.Tensor,
    ids: torch.Tensor | None = None,
    id_base: int = 0,
) -> None:
    """K2: merge (B, C) candidates into the (C, k) bf16 / int64 state in place. No sync."""
    lib = N.load(require_device=True)
    for name, x in (("candidates", cand), ("state values", state_vals), ("state ids", state_ids)):
        N.require_cuda(x, name)
    assert cand.ndim == 2 and state_vals.ndim == 2 and state_vals.shape == state_ids.shape
    B, C = cand.shape
    assert state_vals.shape[0] == C, (state_vals.shape, cand.shape)
    k = state_vals.shape[1]
    assert state_vals.dtype == torch.bfloat16 and state_ids.dtype == torch.int64
    assert state_vals.is_contiguous() and state_ids.is_contiguous()
    if cand.dtype not in (torch.float32, torch.bfloat16):
        cand = cand.float()  # fp16 -> fp32 is exact; the kernel rounds fp32 -> bf16 (RNE) like `.to(bfloat16)`
    cand = cand.contiguous()
    if ids is not None:
        ids = ids.to(device=cand.device, dtype=torch.int64).contiguous()
        assert ids.shape == (B,)
    max_b = 8192 - k
    if max_b <= 0:
        raise N.SlbError(f"n_collect={k} is not supported (k + batch must be <= 8192)")
    with _dev_guard(cand):
        for b0 in range(0, B, max_b):
            cb = cand[b0 : b0 + max_b]
            rc = lib.slb_topk_update(
                cb.data_ptr(), N.dtype_code(cb.dtype), cb.shape[0], C,
                None if ids is None else ids[b0 : b0 + max_b].data_ptr(), id_base + b0,
                state_vals.data_ptr(), state_ids.data_ptr(), k, N.stream_ptr(cand.device),
            )
            N.check(rc, "slb_topk_update")


def agg_topk_update(
    t: torch.Tensor,
    op: int,
    reduce_kind: str,
    token_pos: int,
    id_base: int,
    state_vals: torch.Tensor,
    state_ids: torch.Tensor,
    scratch: torch.Tensor | None = None,
) -> torch.Tensor:
    """K1 + K2 as the forward hook runs them. Returns the scratch buffer (for reuse)."""
    lib = N.load(require_device=True)
    N.require_cuda(t, "activation map")
    t = t.detach()
    keep, layout, B, C, inner = describe_map(t, reduce_kind, op)
    k = state_vals.shape[1]
    need = B * C
    if scratch is None or scratch.numel() < need or scratch.device != t.device:
        scratch = torch.empty(max(need, 1), dtype=torch.float32, device=t.device)
    if B + k > 8192:
        cand = agg_reduce(t, op, reduce_kind, token_pos)
        topk_update(cand, state_vals, state_ids, None, id_base)
        return scratch
    with _dev_guard(t):
        rc = lib.slb_agg_topk_update(
            keep.data_ptr(), N.dtype_code(keep.dtype), layout, B, C, inner, op, token_pos, id_base,
            state_vals.data_ptr(), state_ids.data_ptr(), k, scratch.data_ptr(), scratch.numel() * 4,
            N.stream_ptr(t.device),
        )
    N.check(rc, "slb_agg_topk_update")
    return scratch


def topk_merge_lists(vals: torch.Tensor, ids: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """K2 list mode: (R, C, k) per-rank states -> (C, k) merged state."""
    lib = N.load(require_device=True)
    N.require_cuda(vals, "values")
    N.require_cuda(ids, "ids")
    R, C, k = vals.shape
    assert ids.shape == vals.shape and vals.dtype == torch.bfloat16 and ids.dtype == torch.int64
    vals, ids = vals.contiguous(), ids.contiguous()
    out_v = torch.empty((C, k), dtype=torch.bfloat16, device=vals.device)
    out_i = torch.empty((C, k), dtype=torch.int64, device=vals.device)
    with _dev_guard(vals):
        rc = lib.slb_topk_merge_lists(
            vals.data_ptr(), ids.data_ptr(), R, C, k, out_v.data_ptr(), out_i.data_ptr(), N.stream_ptr(vals.device)
        )
    N.check(rc, "slb_topk_merge_lists")
    return out_v, out_i


def gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """K5: table[idx] for a 2-D fp32 table and an integer index tensor of any shape (python negatives wrap)."""
    lib = N.load(require_device=True)
    N.require_cuda(table, "table")
    assert table.ndim == 2
    table = table.detach().to(torch.float32).contiguous()  # a custom FM may hand over fp16 / bf16 embeddings
    idx_d = idx.to(device=table.device, dtype=torch.int64).contiguous()
    n, D = table.shape
    out = torch.empty((*idx.shape, D), dtype=torch.float32, device=table.device)
    if idx_d.numel() and n == 0:
        raise IndexError("index into an empty table")
    with _dev_guard(table):
        rc = lib.slb_gather_rows(
            table.data_ptr(), n, D, idx_d.data_ptr(), idx_d.numel(), out.data_ptr(), N.stream_ptr(table.device)
        )
    N.check(rc, "slb_gather_rows")
    return out


# ------------------------------------------------------------------------------------------------
# embed: split planes + tensor-core GEMM
# ------------------------------------------------------------------------------------------------
PLANE_DTYPES = {N.PLANE_F16: torch.float16, N.PLANE_BF16: torch.bfloat16}


def split_planes(x: torch.Tensor, fmt: int = N.PLANE_F16, scale: float = 1.0) -> torch.Tensor:
    """fp32 (..., K) -> (2, ..., K) 16-bit planes of ``scale * x``: hi = rn16(scale * x), lo = rn16(scale * x - hi).
    ``scale`` is a power of two that lifts the tensor into the fp16 normal range (weights: N.WEIGHT_PLANE_SCALE)."""
    lib = N.load(require_device=True)
    N.require_cuda(x, "x")
    x = x.detach().to(torch.float32).contiguous()
    planes = torch.empty((2, *x.shape), dtype=PLANE_DTYPES[fmt], device=x.device)
    if x.numel():
        with _dev_guard(x):
            rc = lib.slb_split_planes(x.data_ptr(), x.numel(), fmt, float(scale), planes.data_ptr(), N.stream_ptr(x.device))
        N.check(rc, "slb_split_planes")
    return planes


def _plane_residual(epilogue: int, residual, plane_dtype, M: int, Nn: int):
    """A shortcut handed over as split planes (2, M, N) in the GEMM's plane dtype selects SLB_EPI_ADD_RELU_PLANES (the only
    epilogue with a plane shortcut: the tail of a residual block)."""
    if residual is None or residual.dtype == torch.float32:
        return epilogue, residual
    assert epilogue in (N.EPI_ADD_RELU, N.EPI_ADD_RELU_PLANES), "a plane shortcut goes with the add + ReLU epilogue"
    assert residual.dtype == plane_dtype and tuple(residual.shape) == (2, M, Nn) and residual.is_contiguous(), residual.shape
    return N.EPI_ADD_RELU_PLANES, residual


def planes_to_f32(planes: torch.Tensor) -> torch.Tensor:
    """(2, M, N) activation planes -> the fp32 values they carry (hi + lo) / ACT_PLANE_SCALE."""
    return (planes[0].float() + planes[1].float()) / N.ACT_PLANE_SCALE


def gemm_split(
    a_planes: torch.Tensor,
    w_planes: torch.Tensor,
    *,
    bias: torch.Tensor | None = None,
    residual: torch.Tensor | None = None,
    row_scale: torch.Tensor | None = None,
    col_scale: torch.Tensor | None = None,
    epilogue: int = N.EPI_NONE,
    passes: int = 3,
    alpha: float = 1.0,
    out_f32: torch.Tensor | bool = True,
    out_planes: torch.Tensor | bool = False,
    raw_f32: torch.Tensor | None = None,
):
    """K4: act(alpha * (A @ W^T) * row_scale * col_scale + bias) + residual from split planes (2, M, K) and (2, N, K).
    ``raw_f32`` (M, N) fp32: also receives alpha * (A @ W^T) * row_scale, the value before column scale / bias / activation.

    ``alpha`` = 1 / (scale of the A planes * scale of the W planes). ``out_planes`` are written at N.ACT_PLANE_SCALE.

    ``out_f32`` / ``out_planes``: True = allocate, False = not wanted, or a preallocated tensor.
    Returns (out_f32 | None, out_planes | None).
    """
    lib = N.load(require_device=True)
    N.require_cuda(a_planes, "a_planes")
    assert a_planes.ndim == 3 and w_planes.ndim == 3 and a_planes.shape[0] == 2 and w_planes.shape[0] == 2
    assert a_planes.dtype == w_planes.dtype and a_planes.dtype in (torch.float16, torch.bfloat16)
    assert a_planes.is_contiguous() and w_planes.is_contiguous()
    fmt = N.PLANE_F16 if a_planes.dtype == torch.float16 else N.PLANE_BF16
    _, M, K = a_planes.shape
    _, Nn, K2 = w_planes.shape
    assert K == K2, (a_planes.shape, w_planes.shape)
    dev = a_planes.device
    if out_f32 is True:
        out_f32 = torch.empty((M, Nn), dtype=torch.float32, device=dev)
    elif out_f32 is False:
        out_f32 = None
    if out_planes is True:
        out_planes = torch.empty((2, M, Nn), dtype=a_planes.dtype, device=dev)
    elif out_planes is False:
        out_planes = None
    epilogue, residual = _plane_residual(epilogue, residual, a_planes.dtype, M, Nn)
    for t, shape in ((bias, (Nn,)), (residual if epilogue != N.EPI_ADD_RELU_PLANES else None, (M, Nn)), (row_scale, (M,)),
                     (col_scale, (Nn,)), (out_f32, (M, Nn))):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape, (t.shape, shape)
    if out_planes is not None:
        assert tuple(out_planes.shape) == (2, M, Nn) and out_planes.is_contiguous() and out_planes.dtype == a_planes.dtype
    args = (a_planes.data_ptr(), w_planes.data_ptr(), fmt, M, Nn, K, float(alpha), N.ptr(bias), N.ptr(residual), N.ptr(row_scale),
            N.ptr(col_scale), epilogue, passes, N.ptr(out_f32), N.ptr(out_planes))
    with _dev_guard(a_planes):
        if raw_f32 is not None:
            assert raw_f32.dtype == torch.float32 and raw_f32.is_contiguous() and tuple(raw_f32.shape) == (M, Nn)
            rc = lib.slb_gemm_split_raw(*args, raw_f32.data_ptr(), N.stream_ptr(dev))
        else:
            rc = lib.slb_gemm_split(*args, N.stream_ptr(dev))
    N.check(rc, "slb_gemm_split")
    return out_f32, out_planes


def u8_to_f32_norm(u8: torch.Tensor, mean, std) -> torch.Tensor:
    """K3: uint8 (B, C, H, W) on the GPU -> (x/255 - mean[c]) / std[c] fp32 (ToTensor + Normalize)."""
    lib = N.load(require_device=True)
    N.require_cuda(u8, "images")
    assert u8.dtype == torch.uint8 and u8.ndim == 4
    u8 = u8.contiguous()
    B, C, H, W = u8.shape
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=u8.device)
    import ctypes

    m = (ctypes.c_float * C)(*[float(v) for v in mean][:C])
    s = (ctypes.c_float * C)(*[float(v) for v in std][:C])
    with _dev_guard(u8):
        rc = lib.slb_u8_to_f32_norm(u8.data_ptr(), B, C, H * W, m, s, out.data_ptr(), N.stream_ptr(u8.device))
    N.check(rc, "slb_u8_to_f32_norm")
    return out


def resized_size(w: int, h: int, S: int) -> tuple[int, int]:
    """torchvision ``Resize(S)``: the shorter side becomes S, the longer ``int(S * long / short)``."""
    if w <= h:
        return S, max(S, int(S * h / w))
    return max(S, int(S * w / h)), S


def resize_center_crop_u8(img_hwc: torch.Tensor, S: int, out: torch.Tensor | None = None,
                          squash: bool = False) -> torch.Tensor:
    """(h, w, 3) u8 CUDA -> (3, S, S) u8: Resize(S, bicubic) + CenterCrop(S), or with ``squash`` Resize((S, S), bicubic)
    without a crop; byte-identical to Pillow / torchvision."""
    lib = N.load(require_device=True)
    N.require_cuda(img_hwc, "img_hwc")
    assert img_hwc.dtype == torch.uint8 and img_hwc.ndim == 3 and img_hwc.shape[2] == 3
    img_hwc = img_hwc.contiguous()
    h, w = int(img_hwc.shape[0]), int(img_hwc.shape[1])
    if squash:
        nw, nh, left, top = S, S, 0, 0
    else:
        nw, nh = resized_size(w, h, S)
        left, top = int(round((nw - S) / 2.0)), int(round((nh - S) / 2.0))
    if out is None:
        out = torch.empty((3, S, S), dtype=torch.uint8, device=img_hwc.device)
    assert out.is_contiguous() and tuple(out.shape) == (3, S, S) and out.device == img_hwc.device
    need = lib.slb_resize_workspace_bytes(h, w, nw, nh, left, top, S, S)
    ws = torch.empty(max(need, 1), dtype=torch.uint8, device=img_hwc.device)
    with _dev_guard(img_hwc):
        rc = lib.slb_resize_bicubic_u8(img_hwc.data_ptr(), h, w, nw, nh, left, top, S, S, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                       N.stream_ptr(img_hwc.device))
    N.check(rc, "slb_resize_bicubic_u8")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor | None, eps: float, fmt: int | None = None):
    """Row LayerNorm of a contiguous (rows, cols) fp32 tensor -> fp32 (fmt None) or split planes (2, rows, cols) at
    N.ACT_PLANE_SCALE."""
    lib = N.load(require_device=True)
    N.require_cuda(x, "x")
    assert x.ndim == 2 and x.dtype == torch.float32 and x.is_contiguous()
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=torch.float32, device=x.device) if fmt is None else None
    planes = torch.empty((2, rows, cols), dtype=PLANE_DTYPES[fmt], device=x.device) if fmt is not None else None
    with _dev_guard(x):
        rc = lib.slb_layernorm(x.data_ptr(), rows, cols, cols, gamma.data_ptr(), N.ptr(beta), eps, fmt or 0,
                               N.ptr(out), N.ptr(planes), N.stream_ptr(x.device))
    N.check(rc, "slb_layernorm")
    return out if fmt is None else planes


def attention_packed(qkv: torch.Tensor, heads: int, fmt: int | None = None):
    """softmax(QK^T/sqrt(dh))V on a packed (B, T, 3W) in_proj output -> (B, T, W) fp32 or planes (2, B*T, W)."""
    lib = N.load(require_device=True)
    N.require_cuda(qkv, "qkv")
    assert qkv.ndim == 3 and qkv.dtype == torch.float32 and qkv.is_contiguous()
    B, T, W3 = qkv.shape
    W = W3 // 3
    dh = W // heads
    out = torch.empty((B, T, W), dtype=torch.float32, device=qkv.device) if fmt is None else None
    planes = torch.empty((2, B * T, W), dtype=PLANE_DTYPES[fmt], device=qkv.device) if fmt is not None else None
    base = qkv.data_ptr()
    with _dev_guard(qkv):
        rc = lib.slb_attention_small(base, T * W3, W3, base + 4 * W, base + 8 * W, T * W3, W3, B, T, T, heads, dh,
                                     float(dh) ** -0.5, fmt or 0, N.ptr(out), N.ptr(planes), N.stream_ptr(qkv.device))
    N.check(rc, "slb_attention_small")
    return out if fmt is None else planes


def attention_planes(qkv_planes: torch.Tensor, B: int, heads: int, fmt: int | None = None, causal: bool = False):
    """Attention straight from the in_proj GEMM's split planes (2, B*T, 3W) fp16 -> (B, T, W) fp32 or planes (2, B*T, W)."""
    lib = N.load(require_device=True)
    N.require_cuda(qkv_planes, "qkv_planes")
    assert qkv_planes.ndim == 3 and qkv_planes.shape[0] == 2 and qkv_planes.dtype == torch.float16 and qkv_planes.is_contiguous()
    rows, W3 = qkv_planes.shape[1:]
    T, W = rows // B, W3 // 3
    dh = W // heads
    out = torch.empty((B, T, W), dtype=torch.float32, device=qkv_planes.device) if fmt is None else None
    planes = torch.empty((2, rows, W), dtype=PLANE_DTYPES[fmt], device=qkv_planes.device) if fmt is not None else None
    with _dev_guard(qkv_planes):
        rc = lib.slb_attention_planes(qkv_planes.data_ptr(), B, T, heads, dh, float(dh) ** -0.5, 1 if causal else 0,
                                      N.PLANE_F16, N.ptr(out),
                                      N.ptr(planes), N.stream_ptr(qkv_planes.device))
    N.check(rc, "slb_attention_planes")
    return out if fmt is None else planes


# ------------------------------------------------------------------------------------------------
# analyze: scores
# ------------------------------------------------------------------------------------------------
def patchify(img: torch.Tensor, patch: int, fmt: int = N.PLANE_F16) -> torch.Tensor:
    """(B,3,S,S) fp32 -> im2col planes (2, B*(S/P)^2, patch_k(P)) of the ViT patch-embedding convolution, column (c, py, px)."""
    lib = N.load(require_device=True)
    N.require_cuda(img, "img")
    img = img.float().contiguous()
    B, _, S, _ = img.shape
    out = torch.empty((2, B * (S // patch) ** 2, int(lib.slb_patch_k(patch))), dtype=PLANE_DTYPES[fmt], device=img.device)
    with _dev_guard(img):
        N.check(lib.slb_patchify(img.data_ptr(), B, S, patch, fmt, out.data_ptr(), N.stream_ptr(img.device)), "slb_patchify")
    return out


# ---- CLIP ModifiedResNet pieces (channels-last split planes) ---------------------------------------------------------
def conv_k(cin: int, ksize: int) -> int:
    return int(N.load().slb_conv_k(cin, ksize))


def im2col_stem(img: torch.Tensor, fmt: int = N.PLANE_F16) -> torch.Tensor:
    """(B,3,S,S) fp32 -> planes (2, B*(S/2)^2, 64) of the stem's 3x3 / stride 2 / pad 1 convolution."""
    lib = N.load(require_device=True)
    N.require_cuda(img, "img")
    img = img.float().contiguous()
    B, _, S, _ = img.shape
    out = torch.empty((2, B * (S // 2) ** 2, 64), dtype=PLANE_DTYPES[fmt], device=img.device)
    with _dev_guard(img):
        N.check(lib.slb_im2col_stem(img.data_ptr(), B, S, fmt, out.data_ptr(), N.stream_ptr(img.device)), "slb_im2col_stem")
    return out


def stem_conv3x3s2(img: torch.Tensor, w_planes: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """(B,3,S,S) fp32 -> planes (2, B*(S/2)^2, cout) of relu(conv3x3 / stride 2 / pad 1 * scale + shift), computed directly
    (fp32 FMAs) from the weight planes (2, cout, 64)."""
    lib = N.load(require_device=True)
    N.require_cuda(img, "img")
    img = img.float().contiguous()
    B, _, S, _ = img.shape
    assert w_planes.ndim == 3 and w_planes.shape[0] == 2 and w_planes.shape[2] == 64 and w_planes.is_contiguous()
    cout = w_planes.shape[1]
    fmt = N.PLANE_F16 if w_planes.dtype == torch.float16 else N.PLANE_BF16
    for t in (scale, shift):
        assert t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (cout,)
    out = torch.empty((2, B * (S // 2) ** 2, cout), dtype=w_planes.dtype, device=img.device)
    with _dev_guard(img):
        N.check(lib.slb_stem_conv3x3s2(img.data_ptr(), B, S, w_planes.data_ptr(), cout, fmt, scale.data_ptr(), shift.data_ptr(),
                                       out.data_ptr(), N.stream_ptr(img.device)), "slb_stem_conv3x3s2")
    return out


def im2col3x3(planes: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """channels-last planes (2, B*H*W, C) -> (2, B*H*W, conv_k(C, 3)) of a 3x3 / stride 1 / pad 1 convolution."""
    lib = N.load(require_device=True)
    N.require_cuda(planes, "planes")
    assert planes.ndim == 3 and planes.shape[0] == 2 and planes.shape[1] == B * H * W and planes.is_contiguous()
    C = planes.shape[2]
    out = torch.empty((2, B * H * W, conv_k(C, 3)), dtype=planes.dtype, device=planes.device)
    with _dev_guard(planes):
        N.check(lib.slb_im2col3x3(planes.data_ptr(), B, H, W, C, out.data_ptr(), N.stream_ptr(planes.device)), "slb_im2col3x3")
    return out


def avgpool2_planes(planes: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    lib = N.load(require_device=True)
    N.require_cuda(planes, "planes")
    assert planes.ndim == 3 and planes.shape[0] == 2 and planes.shape[1] == B * H * W and planes.is_contiguous()
    C = planes.shape[2]
    fmt = N.PLANE_F16 if planes.dtype == torch.float16 else N.PLANE_BF16
    out = torch.empty((2, B * (H // 2) * (W // 2), C), dtype=planes.dtype, device=planes.device)
    with _dev_guard(planes):
        N.check(lib.slb_avgpool2_planes(planes.data_ptr(), B, H, W, C, fmt, out.data_ptr(), N.stream_ptr(planes.device)),
                "slb_avgpool2_planes")
    return out


# ------------------------------------------------------------------------------------------------
# accelerated probed-model forward (probed.py): torchvision-style ResNet pieces on channels-last planes
# ------------------------------------------------------------------------------------------------
def conv_out(n: int, k: int, stride: int, pad: int) -> int:
    return (n + 2 * pad - k) // stride + 1


def im2col_nchw(img: torch.Tensor, ksize: int, stride: int, pad: int, fmt: int = N.PLANE_F16) -> torch.Tensor:
    """(B,C,H,W) fp32 -> planes (2, B*Ho*Wo, conv_k(C, ksize)) of a ksize x ksize / stride / pad convolution."""
    lib = N.load(require_device=True)
    N.require_cuda(img, "img")
    img = img.float().contiguous()
    B, C, H, W = img.shape
    Ho, Wo = conv_out(H, ksize, stride, pad), conv_out(W, ksize, stride, pad)
    out = torch.empty((2, B * Ho * Wo, conv_k(C, ksize)), dtype=PLANE_DTYPES[fmt], device=img.device)
    with _dev_guard(img):
        N.check(lib.slb_im2col_nchw(img.data_ptr(), B, C, H, W, ksize, stride, pad, fmt, out.data_ptr(), N.stream_ptr(img.device)),
                "slb_im2col_nchw")
    return out


def conv_gemm(x_planes: torch.Tensor, B: int, H: int, W: int, w_planes: torch.Tensor, ksize: int, stride: int, pad: int, *,
              bias: torch.Tensor | None = None, residual: torch.Tensor | None = None, col_scale: torch.Tensor | None = None,
              epilogue: int = N.EPI_NONE, passes: int = 3, alpha: float = 1.0, out_f32: torch.Tensor | bool = True,
              out_planes: torch.Tensor | bool = False, raw_f32: torch.Tensor | None = None):
    """Implicit-GEMM convolution over channels-last planes (2, B*H*W, C) with weights (2, Cout, ksize*ksize*C), columns
    ordered (ky, kx, c). No im2col matrix: TMA im2col-mode loads feed the tcgen05 GEMM. Returns (fp32 | None, planes | None)
    of shape (B*Ho*Wo, Cout)."""
    lib = N.load(require_device=True)
    N.require_cuda(x_planes, "x_planes")
    assert x_planes.ndim == 3 and x_planes.shape[0] == 2 and x_planes.shape[1] == B * H * W and x_planes.is_contiguous()
    C = x_planes.shape[2]
    assert w_planes.ndim == 3 and w_planes.shape[0] == 2 and w_planes.shape[2] == conv_k(C, ksize) and w_planes.is_contiguous()
    assert w_planes.dtype == x_planes.dtype
    Nn = w_planes.shape[1]
    fmt = N.PLANE_F16 if x_planes.dtype == torch.float16 else N.PLANE_BF16
    M = B * conv_out(H, ksize, stride, pad) * conv_out(W, ksize, stride, pad)
    dev = x_planes.device
    if out_f32 is True:
        out_f32 = torch.empty((M, Nn), dtype=torch.float32, device=dev)
    elif out_f32 is False:
        out_f32 = None
    if out_planes is True:
        out_planes = torch.empty((2, M, Nn), dtype=x_planes.dtype, device=dev)
    elif out_planes is False:
        out_planes = None
    epilogue, residual = _plane_residual(epilogue, residual, x_planes.dtype, M, Nn)
    for t, shape in ((bias, (Nn,)), (residual if epilogue != N.EPI_ADD_RELU_PLANES else None, (M, Nn)), (col_scale, (Nn,)),
                     (out_f32, (M, Nn))):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape, (t.shape, shape)
    args = (x_planes.data_ptr(), B, H, W, C, ksize, stride, pad, w_planes.data_ptr(), Nn, fmt, float(alpha), N.ptr(bias), N.ptr(residual),
            N.ptr(col_scale), epilogue, passes, N.ptr(out_f32), N.ptr(out_planes))
    with _dev_guard(x_planes):
        if raw_f32 is not None:
            assert raw_f32.dtype == torch.float32 and raw_f32.is_contiguous() and tuple(raw_f32.shape) == (M, Nn)
            rc = lib.slb_conv_gemm_raw(*args, raw_f32.data_ptr(), N.stream_ptr(dev))
        else:
            rc = lib.slb_conv_gemm(*args, N.stream_ptr(dev))
    N.check(rc, "slb_conv_gemm")
    return out_f32, out_planes


def im2col3x3_strided(planes: torch.Tensor, B: int, H: int, W: int, stride: int) -> torch.Tensor:
    """channels-last planes (2, B*H*W, C) -> (2, B*Ho*Wo, conv_k(C, 3)) of a 3x3 / stride 1|2 / pad 1 convolution."""
    lib = N.load(require_device=True)
    N.require_cuda(planes, "planes")
    assert planes.ndim == 3 and planes.shape[0] == 2 and planes.shape[1] == B * H * W and planes.is_contiguous()
    C = planes.shape[2]
    Ho, Wo = conv_out(H, 3, stride, 1), conv_out(W, 3, stride, 1)
    out = torch.empty((2, B * Ho * Wo, conv_k(C, 3)), dtype=planes.dtype, device=planes.device)
    with _dev_guard(planes):
        N.check(lib.slb_im2col3x3_strided(planes.data_ptr(), B, H, W, C, stride, out.data_ptr(), N.stream_ptr(planes.device)),
                "slb_im2col3x3_strided")
    return out


def subsample2_planes(planes: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    lib = N.load(require_device=True)
    N.require_cuda(planes, "planes")
    assert planes.ndim == 3 and planes.shape[0] == 2 and planes.shape[1] == B * H * W and planes.is_contiguous()
    C = planes.shape[2]
    out = torch.empty((2, B * ((H + 1) // 2) * ((W + 1) // 2), C), dtype=planes.dtype, device=planes.device)
    with _dev_guard(planes):
        N.check(lib.slb_subsample2_planes(planes.data_ptr(), B, H, W, C, out.data_ptr(), N.stream_ptr(planes.device)),
                "slb_subsample2_planes")
    return out


def affine_act(raw: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, residual: torch.Tensor | None = None, relu: bool = True,
               fmt: int = N.PLANE_F16, out_f32: torch.Tensor | bool = False, out_planes: torch.Tensor | bool = True):
    """y = act(raw * scale + shift (+ residual)) over a channels-last fp32 map (M, C). Returns (fp32 | None, planes | None)."""
    lib = N.load(require_device=True)
    N.require_cuda(raw, "raw")
    assert raw.ndim == 2 and raw.dtype == torch.float32 and raw.is_contiguous()
    M, C = raw.shape
    if out_f32 is True:
        out_f32 = torch.empty_like(raw)
    elif out_f32 is False:
        out_f32 = None
    if out_planes is True:
        out_planes = torch.empty((2, M, C), dtype=PLANE_DTYPES[fmt], device=raw.device)
    elif out_planes is False:
        out_planes = None
    with _dev_guard(raw):
        N.check(lib.slb_affine_act(raw.data_ptr(), M, C, scale.data_ptr(), shift.data_ptr(), N.ptr(residual), int(relu), fmt,
                                   N.ptr(out_f32), N.ptr(out_planes), N.stream_ptr(raw.device)), "slb_affine_act")
    return out_f32, out_planes


def bn_relu_maxpool(raw: torch.Tensor, B: int, H: int, W: int, scale: torch.Tensor, shift: torch.Tensor, fmt: int = N.PLANE_F16,
                    want_f32: bool = False):
    """BatchNorm + ReLU + MaxPool2d(3, 2, 1) over a channels-last fp32 map (B*H*W, C) -> (fp32 | None, planes) of (B*Ho*Wo, C)."""
    lib = N.load(require_device=True)
    N.require_cuda(raw, "raw")
    assert raw.ndim == 2 and raw.shape[0] == B * H * W and raw.dtype == torch.float32 and raw.is_contiguous()
    C = raw.shape[1]
    Mo = B * conv_out(H, 3, 2, 1) * conv_out(W, 3, 2, 1)
    out32 = torch.empty((Mo, C), dtype=torch.float32, device=raw.device) if want_f32 else None
    planes = torch.empty((2, Mo, C), dtype=PLANE_DTYPES[fmt], device=raw.device)
    with _dev_guard(raw):
        N.check(lib.slb_bn_relu_maxpool(raw.data_ptr(), B, H, W, C, scale.data_ptr(), shift.data_ptr(), fmt, N.ptr(out32),
                                        planes.data_ptr(), N.stream_ptr(raw.device)), "slb_bn_relu_maxpool")
    return out32, planes


def pool_tokens(x: torch.Tensor, pos: torch.Tensor, fmt: int = N.PLANE_F16):
    """x (B, HW, C) fp32, pos (HW+1, C) -> (token planes (2, B*(HW+1), C), query planes (2, B, C))."""
    lib = N.load(require_device=True)
    N.require_cuda(x, "x")
    x, pos = x.float().contiguous(), pos.float().contiguous()
    B, HW, C = x.shape
    assert tuple(pos.shape) == (HW + 1, C)
    tok = torch.empty((2, B * (HW + 1), C), dtype=PLANE_DTYPES[fmt], device=x.device)
    qry = torch.empty((2, B, C), dtype=PLANE_DTYPES[fmt], device=x.device)
    with _dev_guard(x):
        N.check(lib.slb_pool_tokens(x.data_ptr(), pos.data_ptr(), B, HW, C, fmt, tok.data_ptr(), qry.data_ptr(),
                                    N.stream_ptr(x.device)), "slb_pool_tokens")
    return tok, qry


def _pad64(d: int) -> int:
    return (d + 63) // 64 * 64


MAX_FEATURES = 2048  # the score kernels keep one row in registers: 16 x 128 floats per warp


def _pad_features(x: torch.Tensor) -> torch.Tensor:
    """Zero-pad the last dimension to a multiple of 4 (the kernels read 16-byte vectors); zero columns change neither
    norms nor dot products. Embedding widths above MAX_FEATURES are rejected with a plain ValueError."""
    D = x.shape[-1]
    if D > MAX_FEATURES:
        raise ValueError(f"the B200 score kernels support at most {MAX_FEATURES} features per row (got {D})")
    if D % 4 == 0:
        return x
    return torch.nn.functional.pad(x, (0, 4 - D % 4))


UNIT_ROW_PLANE_SCALE = 1024.0  # unit-norm rows: elements ~ 1/sqrt(D)


def normalize_split_rows(x: torch.Tensor, eps: float = 1e-12, fmt: int = N.PLANE_F16,
                         scale: float = UNIT_ROW_PLANE_SCALE) -> torch.Tensor:
    """F.normalize(x, dim=-1) of a (rows, D) fp32 tensor, emitted as GEMM operand planes (2, rows, pad64(D)) at ``scale``."""
    lib = N.load(require_device=True)
    N.require_cuda(x, "x")
    assert x.ndim == 2 and x.dtype == torch.float32
    x = _pad_features(x).contiguous()
    rows, D = x.shape
    planes = torch.empty((2, rows, _pad64(D)), dtype=PLANE_DTYPES[fmt], device=x.device)
    with _dev_guard(x):
        rc = lib.slb_normalize_split_rows(x.data_ptr(), rows, D, eps, fmt, float(scale), planes.data_ptr(), None,
                                          N.stream_ptr(x.device))
    N.check(rc, "slb_normalize_split_rows")
    return planes


def cosine_gemm(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """K6: normalize(x) @ normalize(y).T for (M, D) and (N, D) fp32 CUDA tensors -> (M, N) fp32."""
    lib = N.load(require_device=True)
    N.require_cuda(x, "x")
    N.require_cuda(y, "y")
    assert x.ndim == 2 and y.ndim == 2 and x.shape[1] == y.shape[1], (x.shape, y.shape)
    x = _pad_features(x.detach().to(torch.float32)).contiguous()
    y = _pad_features(y.detach().to(torch.float32)).contiguous()
    M, D = x.shape
    Nn = y.shape[0]
    n_pad = (Nn + 7) // 8 * 8
    if n_pad != Nn:  # the GEMM writes 16-byte vectors: pad y with zero rows, slice the result
        y = torch.cat([y, torch.zeros((n_pad - Nn, D), dtype=y.dtype, device=y.device)])
    out = torch.empty((M, n_pad), dtype=torch.float32, device=x.device)
    if M == 0 or Nn == 0:
        return out[:, :Nn]
    need = lib.slb_cosine_gemm_workspace_bytes(M, n_pad, D)
    ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    with _dev_guard(x):
        rc = lib.slb_cosine_gemm(x.data_ptr(), M, y.data_ptr(), n_pad, D, out.data_ptr(), ws.data_ptr(), need,
                                 N.stream_ptr(x.device))
    N.check(rc, "slb_cosine_gemm")
    return out if n_pad == Nn else out[:, :Nn].contiguous()


def cosine_rows(x: torch.Tensor, y: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """F.cosine_similarity(x, y, dim=-1) for equal-shape fp32 CUDA tensors."""
    lib = N.load(require_device=True)
    N.require_cuda(x, "x")
    N.require_cuda(y, "y")
    assert x.shape == y.shape
    x = x.detach().to(torch.float32).contiguous()
    y = y.detach().to(torch.float32).contiguous()
    D = x.shape[-1]
    rows = x.numel() // D if D else 0
    out = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
    if rows:
        with _dev_guard(x):
            rc = lib.slb_cosine_rows(x.data_ptr(), y.data_ptr(), rows, D, eps, out.data_ptr(), N.stream_ptr(x.device))
        N.check(rc, "slb_cosine_rows")
    return out


def clarity(V: torch.Tensor) -> torch.Tensor:
    """K7: clarity_score of a (..., k, D) fp32 CUDA tensor -> (...,) fp32."""
    lib = N.load(require_device=True)
    N.require_cuda(V, "V")
    assert V.ndim >= 2
    V = _pad_features(V.detach().to(torch.float32)).contiguous()
    k, D = V.shape[-2], V.shape[-1]
    C = V.numel() // (k * D) if k * D else 0
    out = torch.empty(V.shape[:-2], dtype=torch.float32, device=V.device)
    if C:
        with _dev_guard(V):
            rc = lib.slb_clarity(V.data_ptr(), C, k, D, out.data_ptr(), N.stream_ptr(V.device))
        N.check(rc, "slb_clarity")
    return out


def kmeanspp_draws(k: int, seed: int, n_init: int = 10):
    """The data-independent draws of sklearn's k-means++ for n_clusters=2 (host side of K8): KMeans.fit creates
    ``RandomState(seed)`` once per fit and every init consumes ``choice(k, p=uniform)`` (first centre) then
    ``uniform(size=2)`` (the two local trials); nothing else touches the stream."""
    import numpy as np

    rs = np.random.RandomState(seed)
    p = np.ones(k, dtype=np.float64)
    p = p / p.sum()
    first = np.empty(n_init, dtype=np.int64)
    rand = np.empty((n_init, 2), dtype=np.float64)
    for i in range(n_init):
        first[i] = rs.choice(k, p=p)
        rand[i] = rs.uniform(size=2)
    return first, rand


def polysem_2means(V: torch.Tensor, random_state: int = 123, replace_empty_clusters: bool = True,
                   n_init: int = 10) -> torch.Tensor:
    """K8: polysemanticity of every neuron of a (C, k, D) fp32 CUDA tensor -> (C,) float64."""
    lib = N.load(require_device=True)
    N.require_cuda(V, "V")
    assert V.ndim == 3
    V = V.detach().to(torch.float32).contiguous()
    C, k, D = V.shape
    out = torch.empty((C,), dtype=torch.float64, device=V.device)
    if C == 0:
        return out
    if k == 0 or D == 0:
        raise ValueError("polysemanticity_score needs at least one sample and one feature per neuron")
    first, rand = kmeanspp_draws(k, random_state, n_init)
    need = lib.slb_polysem_workspace_bytes(C, k)
    if need == 0:
        raise N.SlbError(f"polysemanticity kernel supports at most 256 examples per neuron (got {k})")
    ws = torch.empty(need, dtype=torch.uint8, device=V.device)
    with _dev_guard(V):
        rc = lib.slb_polysem_2means(V.data_ptr(), C, k, D, first.ctypes.data, rand.ctypes.data, n_init,
                                    1 if replace_empty_clusters else 0, out.data_ptr(), ws.data_ptr(), need,
                                    N.stream_ptr(V.device))
    N.check(rc, "slb_polysem_2means")
    return out


POLYSEM_FAST_MAX_EXAMPLES = 256  # K8's Gram-matrix kernel: one thread per example
POLYSEM_MAX_CLUSTERS = 8


def kmeanspp_draws_general(k: int, n_clusters: int, seed: int, n_init: int = 10):
    """The data-independent draws of sklearn's k-means++ for any n_clusters: per init ``choice(k, p=uniform)`` (first
    centre), then for every further centre ``uniform(size=2 + int(log(n_clusters)))`` (the local trials, _kmeans.py:226,249).
    -> first (n_init,) int64, rand (n_init, n_clusters - 1, n_local_trials) float64, n_local_trials."""
    import numpy as np

    L = 2 + int(np.log(n_clusters))
    rs = np.random.RandomState(seed)
    p = np.ones(k, dtype=np.float64)
    p = p / p.sum()
    first = np.empty(n_init, dtype=np.int64)
    rand = np.empty((n_init, n_clusters - 1, L), dtype=np.float64)
    for i in range(n_init):
        first[i] = rs.choice(k, p=p)
        for c in range(1, n_clusters):
            rand[i, c - 1] = rs.uniform(size=L)
    return first, rand, L


def polysem_kmeans(V: torch.Tensor, n_clusters: int = 2, random_state: int = 123, replace_empty_clusters: bool = True,
                   n_init: int = 10) -> torch.Tensor:
    """K8g: polysemanticity for any n_clusters in [2, 8] and any number of examples: (C, k, D) fp32 CUDA -> (C,) float64."""
    lib = N.load(require_device=True)
    N.require_cuda(V, "V")
    assert V.ndim == 3
    V = V.detach().to(torch.float32).contiguous()
    C, k, D = V.shape
    out = torch.empty((C,), dtype=torch.float64, device=V.device)
    if C == 0:
        return out
    if k == 0 or D == 0:
        raise ValueError("polysemanticity_score needs at least one sample and one feature per neuron")
    if k < n_clusters:
        raise ValueError(f"n_samples={k} should be >= n_clusters={n_clusters}.")  # sklearn's message
    if not 2 <= n_clusters <= POLYSEM_MAX_CLUSTERS:
        raise N.SlbError(f"polysemanticity kernel supports 2..{POLYSEM_MAX_CLUSTERS} clusters (got {n_clusters})")
    first, rand, L = kmeanspp_draws_general(k, n_clusters, random_state, n_init)
    need = lib.slb_polysem_kmeans_workspace_bytes(C, k, D, n_clusters, n_init)
    ws = torch.empty(need, dtype=torch.uint8, device=V.device)
    with _dev_guard(V):
        rc = lib.slb_polysem_kmeans(V.data_ptr(), C, k, D, n_clusters, first.ctypes.data, rand.ctypes.data, L, n_init,
                                    1 if replace_empty_clusters else 0, out.data_ptr(), ws.data_ptr(), need,
                                    N.stream_ptr(V.device))
    N.check(rc, "slb_polysem_kmeans")
    return out


def redundancy(cones: torch.Tensor) -> torch.Tensor:
    """K9: mean_i max_{j != i} cos(cones_i, cones_j) for a (n, D) fp32 CUDA tensor -> 0-d fp32 tensor.

    One call: normalise + split the rows, one tensor-core GEMM whose epilogue keeps only the row maxima (the n x n cosine
    matrix is never written), fixed-order mean."""
    lib = N.load(require_device=True)
    N.require_cuda(cones, "cones")
    assert cones.ndim == 2
    cones = _pad_features(cones.detach().to(torch.float32)).contiguous()
    n, D = cones.shape
    out = torch.empty((), dtype=torch.float32, device=cones.device)
    if n == 0:
        raise RuntimeError("max(): Expected reduction dim -1 to have non-zero size.")  # what torch raises in the reference
    need = lib.slb_redundancy_workspace_bytes(n, D)
    ws = torch.empty(need, dtype=torch.uint8, device=cones.device)
    with _dev_guard(cones):
        rc = lib.slb_redundancy(cones.data_ptr(), n, D, out.data_ptr(), ws.data_ptr(), need, N.stream_ptr(cones.device))
    N.check(rc, "slb_redundancy")
    return out
