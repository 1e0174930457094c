"""Host-side wrappers: torch tensors in/out, raw pointers + the current CUDA stream across the C-ABI.

Nothing here computes; every function validates its arguments, allocates outputs with the PyTorch caching
allocator and enqueues one or two libslb200 kernels on ``torch.cuda.current_stream()`` without synchronising.
"""

from __future__ import annotations

import torch

from . import _native as N


def _dev_guard(t: torch.Tensor):
    return torch.cuda.device(t.device)


class KernelTimer:
    """Optional per-kernel CUDA-event timing on the launching stream (used by bench.py for the roofline line).

    While installed with :func:`set_timer`, every wrapper brackets its kernel with two events on the current
    stream and records (label, algorithmic bytes, algorithmic flops). Nothing synchronises until :meth:`totals`.
    """

    def __init__(self):
        self.records: list[tuple[str, torch.cuda.Event, torch.cuda.Event, int, int]] = []

    def begin(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def end(self, label: str, start, nbytes: int = 0, flops: int = 0):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.records.append((label, start, ev, nbytes, flops))

    def totals(self) -> dict[str, dict[str, float]]:
        torch.cuda.synchronize()
        out: dict[str, dict[str, float]] = {}
        for label, a, b, nbytes, flops in self.records:
            d = out.setdefault(label, {"ms": 0.0, "launches": 0, "bytes": 0, "flops": 0})
            d["ms"] += a.elapsed_time(b)
            d["launches"] += 1
            d["bytes"] += nbytes
            d["flops"] += flops
        return out


_timer: KernelTimer | None = None


def set_timer(t: KernelTimer | None) -> None:
    global _timer
    _timer = t


# ------------------------------------------------------------------------------------------------
# collect
# ------------------------------------------------------------------------------------------------
def describe_map(t: torch.Tensor, reduce_kind: str):
    """Classify a hooked activation map for K1 without copying when possible.

    reduce_kind: "conv" (4-D, reduce H*W) or "tokens" (3-D, reduce dim 1).
    Returns (tensor_to_keep_alive, layout, B, C, inner).
    """
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
    keep, layout, B, C, inner = describe_map(t, reduce_kind)
    out = torch.empty((B, C), dtype=torch.float32, device=t.device)
    if B == 0 or C == 0:
        return out
    with _dev_guard(t):
        t0 = _timer.begin() if _timer else None
        rc = lib.slb_agg_reduce(
            keep.data_ptr(), N.dtype_code(keep.dtype), layout, B, C, inner, op, token_pos, out.data_ptr(),
            N.stream_ptr(t.device),
        )
        if _timer:
            _timer.end("K1 agg_reduce", t0, keep.numel() * keep.element_size())
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
    keep, layout, B, C, inner = describe_map(t, reduce_kind)
    k = state_vals.shape[1]
    need = B * C
    if scratch is None or scratch.numel() < need or scratch.device != t.device:
        scratch = torch.empty(max(need, 1), dtype=torch.float32, device=t.device)
    if B + k > 8192 or _timer is not None:
        cand = agg_reduce(t, op, reduce_kind, token_pos)
        t0 = _timer.begin() if _timer else None
        topk_update(cand, state_vals, state_ids, None, id_base)
        if _timer:
            _timer.end("K2 topk_update", t0, cand.numel() * 4)
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
    assert table.ndim == 2 and table.dtype == torch.float32
    table = table.contiguous()
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
