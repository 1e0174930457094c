"""TEST INFRASTRUCTURE ONLY — numpy restatement of the bicubic ``Resize`` + ``CenterCrop`` of open_clip's eval transform.

The reference applies the transform per PIL image (foundation_models/clip.py:157-160 -> open_clip ``image_transform`` ->
``torchvision.transforms.Resize(S, BICUBIC)`` + ``CenterCrop(S)``), which lands in Pillow's ``ImagingResample``
(Pillow 12.2.0 ``src/libImaging/Resample.c``; third-party, not vendored in /root/reference, installed here): a
separable resampling with the Keys bicubic kernel (a = -0.5) whose support grows with the down-scaling factor
(anti-aliasing), coefficients normalised in double precision and rounded to 22-bit fixed point, a horizontal pass into an
8-bit intermediate, then a vertical pass; both passes round with ``(acc + 2^21) >> 22`` and clamp to [0, 255].
All integer: the B200 kernel (slb_resize_bicubic_u8) must match bit for bit.

PINNED: tests/test_oracle_resize.py compares this port with Pillow itself on random images of many sizes.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""

from __future__ import annotations

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x: np.ndarray) -> np.ndarray:
    x = np.abs(x)
    near = ((1.5 * x - 2.5) * x) * x + 1.0
    far = ((((x - 5.0) * x) + 8.0) * x - 4.0) * -0.5
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def precompute_coeffs(in_size: int, out_size: int):
    """-> bounds (out_size, 2) [first tap, tap count], integer coefficients (out_size, ksize) (Resample.c precompute_coeffs
    + normalize_coeffs_8bpc)."""
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = bicubic_filter((np.arange(xmax, dtype=np.float64) + xmin - center + 0.5) * ss)
        ww = np.add.accumulate(w)[-1] if xmax else 0.0  # sequential sum, like the C loop
        if ww != 0.0:
            w = w / ww
        fixed = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS)).astype(np.int32)  # C (int) truncates
        kk[xx, :xmax] = fixed
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(src: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """Resample axis 1 of src (rows, in, bands) uint8 -> (rows, out, bands) uint8."""
    out = np.empty((src.shape[0], bounds.shape[0], src.shape[2]), dtype=np.uint8)
    s32 = src.astype(np.int32)
    for xx, (xmin, n) in enumerate(bounds):
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(s32[:, xmin:xmin + n], kk[xx, :n], axes=([1], [0]))
        out[:, xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def resize_bicubic(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """(h, w, bands) uint8 -> (out_h, out_w, bands) uint8, Pillow's two passes (horizontal first)."""
    h, w = img.shape[:2]
    tmp = _pass(img, *precompute_coeffs(w, out_w)) if out_w != w else img
    if out_h != h:
        tmp = _pass(tmp.transpose(1, 0, 2), *precompute_coeffs(h, out_h)).transpose(1, 0, 2)
    return np.ascontiguousarray(tmp)


def resized_size(w: int, h: int, S: int) -> tuple[int, int]:
    """torchvision Resize(int S): the shorter side becomes S, the longer int(S * long / short)."""
    if w <= h:
        return S, max(S, int(S * h / w))
    return max(S, int(S * w / h)), S


def resize_center_crop(img: np.ndarray, S: int) -> np.ndarray:
    """(h, w, 3) uint8 -> (3, S, S) uint8: Resize(S, bicubic) on the shorter side, CenterCrop(S), channels first."""
    h, w = img.shape[:2]
    nw, nh = resized_size(w, h, S)
    r = resize_bicubic(img, nw, nh)
    left, top = int(round((nw - S) / 2.0)), int(round((nh - S) / 2.0))
    return np.ascontiguousarray(r[top:top + S, left:left + S].transpose(2, 0, 1))
