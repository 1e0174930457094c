"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's collect stage.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Restates, in numpy:
  * the aggregators            reference semanticlens/component_visualization/aggregators.py:38-244
  * ActMax state + update      reference semanticlens/component_visualization/activation_caching.py:101-141
  * hook id numbering          reference semanticlens/component_visualization/activation_caching.py:403-416
  * torch CPU top-k value semantics (NaN first, +0 == -0)   ATen/native/TopKImpl.h:45-96
Pinned against the reference itself: tests/golden/collect_*.npz are produced by oracle/make_golden.py, which imports
/root/reference (crp/zennit stubbed) and records its inputs and outputs, plus the reference's own known-answer test
(tests/component_visualization/test_activation_caching.py:14-30).

Two aggregate flavours are provided:
  * aggregate_exact      fp64 accumulation, rounded once to fp32 — the mathematically ideal value
  * aggregate_canonical  fp32 accumulation in the kernel's documented canonical order — bit-exact comparison
"""

from __future__ import annotations

import numpy as np

OPS = ("mean", "max", "absmean", "absmax", "token")


# ------------------------------------------------------------------------------------------------
# bf16 (c10::BFloat16 round-to-nearest-even, NaN -> 0x7FC0)
# ------------------------------------------------------------------------------------------------
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    nan = (u & np.uint32(0x7FFFFFFF)) > np.uint32(0x7F800000)
    bias = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    r = ((u.astype(np.uint64) + bias.astype(np.uint64)) >> np.uint64(16)).astype(np.uint16)
    return np.where(nan, np.uint16(0x7FC0), r)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


# ------------------------------------------------------------------------------------------------
# aggregators
# ------------------------------------------------------------------------------------------------
def _reduce_axis(kind: str) -> int:
    return 2 if kind == "conv" else 1


def aggregate_exact(x: np.ndarray, op: str, kind: str, token: int = 0) -> np.ndarray:
    """(B,C,H,W) [kind=conv] or (B,T,F) [kind=tokens] -> (B, C|F) float32 via float64 accumulation."""
    if kind == "conv":
        assert x.ndim == 4
        x = x.reshape(x.shape[0], x.shape[1], -1)
    else:
        assert x.ndim == 3
    ax = _reduce_axis(kind)
    xd = x.astype(np.float64)
    if op == "mean":
        r = xd.mean(axis=ax)
    elif op == "max":
        r = _nanmax(xd, ax)
    elif op == "absmean":
        r = np.abs(xd).mean(axis=ax)
    elif op == "absmax":
        r = _nanmax(np.abs(xd), ax)
    elif op == "token":
        assert kind == "tokens"
        r = xd[:, token]
    else:
        raise ValueError(op)
    return r.astype(np.float32)


def _nanmax(x, ax):
    with np.errstate(invalid="ignore"):
        r = x.max(axis=ax)  # numpy's max propagates NaN like torch.amax
    return r


def _fold(a, b, is_max):
    if is_max:
        with np.errstate(invalid="ignore"):
            return np.where(np.isnan(a), a, np.where(np.isnan(b), b, np.maximum(a, b)))
    return (a + b).astype(np.float32)


def aggregate_canonical(x: np.ndarray, op: str, kind: str, token: int = 0) -> np.ndarray:
    """fp32 accumulation in the kernel's canonical order (semanticlens_b200/csrc/agg_reduce.cu header)."""
    x = np.asarray(x)
    in_dtype = x.dtype
    is_max = op in ("max", "absmax")
    ident = np.float32(-np.inf) if is_max else np.float32(0.0)
    if op == "token":
        return x[:, token].astype(np.float32)
    xf = x.astype(np.float32)
    if op in ("absmean", "absmax"):
        xf = np.abs(xf)
    if kind == "conv":
        B, C = xf.shape[:2]
        rows = xf.reshape(B * C, -1)
        L = rows.shape[1]
        nblk = -(-L // 256)
        pad = np.full((rows.shape[0], nblk * 256), ident, dtype=np.float32)
        pad[:, :L] = rows
        valid = np.zeros(nblk * 256, dtype=bool)
        valid[:L] = True
        pad = pad.reshape(-1, nblk, 8, 32)
        valid = valid.reshape(nblk, 8, 32)
        acc = np.full((rows.shape[0], 8, 32), ident, dtype=np.float32)
        for i in range(nblk):  # each accumulator folds its elements in increasing e
            nxt = _fold(acc, pad[:, i], is_max)
            acc = np.where(valid[i][None], nxt, acc)
        a = [acc[:, j] for j in range(8)]
        t = _fold(
            _fold(_fold(a[0], a[1], is_max), _fold(a[2], a[3], is_max), is_max),
            _fold(_fold(a[4], a[5], is_max), _fold(a[6], a[7], is_max), is_max),
            is_max,
        )  # (rows, 32)
        for o in (16, 8, 4, 2, 1):  # xor butterfly: lane i <- lane i (+) lane i^o
            idx = np.arange(32) ^ o
            t = _fold(t, t[:, idx], is_max)
        r = t[:, 0]
        n = L
        r = r.reshape(B, C)
    else:
        # two-level order: blocks of 64 tokens folded sequentially from the identity, block partials folded in block order
        B, T, F = xf.shape
        acc = np.full((B, F), ident, dtype=np.float32)
        for t0 in range(0, T, 64):
            part = np.full((B, F), ident, dtype=np.float32)
            for t_ in range(t0, min(t0 + 64, T)):
                part = _fold(part, xf[:, t_], is_max)
            acc = _fold(acc, part, is_max)
        r, n = acc, T
    if not is_max:
        r = (r / np.float32(n)).astype(np.float32)
    if in_dtype == np.float16:
        r = r.astype(np.float16).astype(np.float32)
    return r


# ------------------------------------------------------------------------------------------------
# top-k state (canonical order)
# ------------------------------------------------------------------------------------------------
ID_MASK = (1 << 47) - 1


def topk_key(bits: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """u64 key whose descending order is (value desc, +0 == -0, NaN first; id asc; placeholders id=-1 last)."""
    bits = np.asarray(bits, dtype=np.uint16).astype(np.uint32)
    ids = np.asarray(ids, dtype=np.int64)
    mag = bits & 0x7FFF
    isnan = mag > 0x7F80
    iszero = mag == 0
    neg = (bits & 0x8000) != 0
    vkey = np.where(isnan, 0xFFFF, np.where(iszero, 0x8000, np.where(neg, (~bits) & 0xFFFF, bits | 0x8000)))
    zsign = np.where(iszero & ~isnan, bits >> 15, 0).astype(np.uint64)
    idf = np.where(ids < 0, 0, (ID_MASK - ids) & ID_MASK).astype(np.uint64)
    return (vkey.astype(np.uint64) << np.uint64(48)) | (idf << np.uint64(1)) | zsign


def topk_unkey(key: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    key = np.asarray(key, dtype=np.uint64)
    vkey = (key >> np.uint64(48)).astype(np.uint32)
    z = (key & np.uint64(1)).astype(np.uint32)
    bits = np.where(
        vkey == 0xFFFF,
        0x7FC0,
        np.where(vkey == 0x8000, z << 15, np.where((vkey & 0x8000) != 0, vkey & 0x7FFF, (~vkey) & 0xFFFF)),
    ).astype(np.uint16)
    idf = ((key >> np.uint64(1)) & np.uint64(ID_MASK)).astype(np.int64)
    ids = np.where(idf == 0, -1, ID_MASK - idf).astype(np.int64)
    return bits, ids


class ActMaxOracle:
    """ActMax (reference activation_caching.py:64-141) with the canonical tie order."""

    def __init__(self, n_collect: int, n_latents: int | None = None):
        self.n_collect, self.n_latents = n_collect, n_latents
        self.bits = self.ids = None
        if n_latents is not None:
            self._setup()

    def _setup(self):
        self.bits = np.full((self.n_latents, self.n_collect), 0x8000, dtype=np.uint16)  # -0.0  (:108)
        self.ids = np.full((self.n_latents, self.n_collect), -1, dtype=np.int64)  # (:109)

    def update(self, acts: np.ndarray, sample_ids: np.ndarray):
        acts = np.asarray(acts)
        assert acts.ndim == 2
        if self.bits is None:
            self.n_latents = acts.shape[1]
            self._setup()
        cand_bits = f32_to_bf16_bits(acts.astype(np.float32)).T  # `acts.T.to(bfloat16)`  (:133)
        cand_ids = np.broadcast_to(np.asarray(sample_ids, dtype=np.int64)[None], cand_bits.shape)  # (:134)
        keys = np.concatenate([topk_key(self.bits, self.ids), topk_key(cand_bits, cand_ids)], axis=1)  # (:137-138)
        keys = _sort_desc(keys)
        self.bits, self.ids = topk_unkey(keys[:, : self.n_collect])  # topk + gather (:140-141)

    @property
    def activations(self) -> np.ndarray:
        return bf16_bits_to_f32(self.bits)


def _sort_desc(keys: np.ndarray) -> np.ndarray:
    return np.sort(keys, axis=1)[:, ::-1]


def merge_lists(bits: np.ndarray, ids: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """(R, C, k) per-rank states -> (C, k) global state (canonical order)."""
    R, C, k = bits.shape
    keys = topk_key(bits, ids).transpose(1, 0, 2).reshape(C, R * k)
    return topk_unkey(_sort_desc(keys)[:, :k])


def sweep(maps_per_batch, op: str, kind: str, n_collect: int, token: int = 0, id_base: int = 0, exact: bool = False):
    """Hook semantics over a list of per-batch maps for ONE layer: ids = running counter (reference :410-413)."""
    st = ActMaxOracle(n_collect)
    counter = id_base
    agg = aggregate_exact if exact else aggregate_canonical
    for m in maps_per_batch:
        a = agg(np.asarray(m), op, kind, token)
        st.update(a, np.arange(counter, counter + a.shape[0]))
        counter += a.shape[0]
    return st


# ------------------------------------------------------------------------------------------------
# tie-aware parity contract
# ------------------------------------------------------------------------------------------------
def values_equal(bits_a: np.ndarray, bits_b: np.ndarray) -> np.ndarray:
    """bf16 equality with +0 == -0 and all NaNs equal."""
    a = np.asarray(bits_a, dtype=np.uint16)
    b = np.asarray(bits_b, dtype=np.uint16)
    za, zb = (a & 0x7FFF) == 0, (b & 0x7FFF) == 0
    na, nb = (a & 0x7FFF) > 0x7F80, (b & 0x7FFF) > 0x7F80
    return (a == b) | (za & zb) | (na & nb)


def check_tie_aware(bits, ids, ref_bits, ref_ids, cand_bits_by_id=None) -> list[str]:
    """Return a list of violations of the contract between a kernel state and a reference state.

    (1) values equal slot by slot; (2) ids distinct per row (except -1), every id's own bf16 aggregate equals the
    slot value when cand_bits_by_id (C, N) is given, -1 only on +-0 slots; (3) for every tie group strictly above
    the k-th value, the id sets agree.
    """
    errs = []
    bits, ids, ref_bits, ref_ids = map(np.asarray, (bits, ids, ref_bits, ref_ids))
    if bits.shape != ref_bits.shape:
        return [f"shape {bits.shape} != {ref_bits.shape}"]
    eq = values_equal(bits, ref_bits)
    if not eq.all():
        errs.append(f"{(~eq).sum()} value slots differ, first at {np.argwhere(~eq)[0].tolist()}")
    C, k = bits.shape
    for c in range(C):
        row = ids[c]
        real = row[row >= 0]
        if len(np.unique(real)) != len(real):
            errs.append(f"row {c}: duplicate ids")
        if ((row < 0) & ((bits[c] & 0x7FFF) != 0)).any():
            errs.append(f"row {c}: placeholder id on a non-zero value")
        if cand_bits_by_id is not None and len(real):
            own = cand_bits_by_id[c, real]
            if not values_equal(own, bits[c][row >= 0]).all():
                errs.append(f"row {c}: an id does not carry its slot's value")
        if k == 0:
            continue
        keyv = (topk_key(bits[c], np.zeros(k, dtype=np.int64)) >> np.uint64(48)).astype(np.int64)
        kth = keyv[-1]
        for v in np.unique(keyv):
            if v == kth:
                continue  # boundary tie group: any subset of the right size is acceptable
            mine = set(ids[c][keyv == v].tolist())
            refkey = (topk_key(ref_bits[c], np.zeros(k, dtype=np.int64)) >> np.uint64(48)).astype(np.int64)
            theirs = set(ref_ids[c][refkey == v].tolist())
            if mine != theirs:
                errs.append(f"row {c}: id set of interior tie group differs")
    return errs
