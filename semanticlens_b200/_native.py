"""ctypes binding of libslb200.so (the C-ABI declared in include/slb200.h).

This is the only place Python touches the kernels. There is no CPU fallback: if the shared library is missing
or no CUDA device is visible, every entry point raises.
"""

from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libslb200.so"
_lib = None

# enums (mirror include/slb200.h)
DT_F32, DT_F16, DT_BF16 = 0, 1, 2
LAYOUT_NCHW, LAYOUT_BTF = 0, 1
AGG_MEAN, AGG_MAX, AGG_ABSMEAN, AGG_ABSMAX, AGG_TOKEN = 0, 1, 2, 3, 4
EPI_NONE, EPI_GELU_ERF, EPI_QUICKGELU, EPI_GELU_TANH, EPI_RELU, EPI_ADD_RELU, EPI_ADD_RELU_PLANES = 0, 1, 2, 3, 4, 5, 6
POOL_CLS, POOL_MAP = 0, 1
PASSES_SPLIT_ACC = 4  # SLB_PASSES_SPLIT_ACC
PLANE_F16, PLANE_BF16 = 0, 1
ACT_PLANE_SCALE, WEIGHT_PLANE_SCALE = 16.0, 1024.0  # SLB_ACT_PLANE_SCALE / SLB_WEIGHT_PLANE_SCALE

_DTYPES = {torch.float32: DT_F32, torch.float16: DT_F16, torch.bfloat16: DT_BF16}


class SlbError(RuntimeError):
    """A libslb200 entry point returned a negative status."""


_PROTOS = {
    "slb_version": (c_int, []),
    "slb_last_error": (c_char_p, []),
    "slb_launch_count": (c_int64, []),
    "slb_device_info": (c_int, [c_void_p, c_void_p, c_void_p]),
    "slb_profile_begin": (c_int, []),
    "slb_profile_end": (c_int, []),
    "slb_profile_summary": (c_int, [c_void_p, c_int, c_void_p]),
    "slb_agg_reduce": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_int64, c_int, c_int64, c_void_p, c_void_p]),
    "slb_topk_update": (
        c_int,
        [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p],
    ),
    "slb_agg_topk_update": (
        c_int,
        [c_void_p, c_int, c_int, c_int64, c_int64, c_int64, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int64,
         c_void_p, c_size_t, c_void_p],
    ),
    "slb_topk_merge_lists": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "slb_gather_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "slb_normalize_split_rows": (c_int, [c_void_p, c_int64, c_int64, c_float, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "slb_cosine_gemm_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "slb_cosine_gemm": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "slb_cosine_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p]),
    "slb_clarity": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "slb_polysem_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "slb_polysem_phase_clocks": (c_int, [c_void_p, c_int]),
    "slb_polysem_2means": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "slb_polysem_kmeans_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int]),
    "slb_polysem_kmeans": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
         c_void_p],
    ),
    "slb_rowmax_offdiag": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "slb_redundancy_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "slb_redundancy": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "slb_split_planes": (c_int, [c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p]),
    "slb_gemm_split": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
         c_void_p, c_void_p, c_void_p],
    ),
    "slb_u8_to_f32_norm": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slb_gemm_split_raw": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slb_patchify": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "slb_assemble_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "slb_layernorm": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "slb_attention_small": (
        c_int,
        [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
         c_float, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "slb_attention_planes": (
        c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "slb_text_workspace_bytes": (c_size_t, [c_void_p, c_int64]),
    "slb_text_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "slb_patch_k": (c_int64, [c_int64]),
    "slb_attention_trace": (c_void_p, []),
    "slb_resize_workspace_bytes": (c_size_t, [c_int64] * 8),
    "slb_resize_bicubic_u8": (c_int, [c_void_p] + [c_int64] * 8 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    "slb_conv_k": (c_int64, [c_int64, c_int64]),
    "slb_im2col_stem": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "slb_stem_conv3x3s2": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slb_im2col3x3": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "slb_avgpool2_planes": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "slb_conv_gemm": (c_int, [c_void_p] + [c_int64] * 4 + [c_int] * 3 + [c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "slb_conv_gemm_raw": (c_int, [c_void_p] + [c_int64] * 4 + [c_int] * 3 + [c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "slb_im2col_nchw": (c_int, [c_void_p] + [c_int64] * 4 + [c_int] * 4 + [c_void_p, c_void_p]),
    "slb_im2col3x3_strided": (c_int, [c_void_p] + [c_int64] * 4 + [c_int, c_void_p, c_void_p]),
    "slb_subsample2_planes": (c_int, [c_void_p] + [c_int64] * 4 + [c_void_p, c_void_p]),
    "slb_affine_act": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "slb_bn_relu_maxpool": (c_int, [c_void_p] + [c_int64] * 4 + [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "slb_pool_tokens": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "slb_rn_workspace_bytes": (c_size_t, [c_void_p, c_int64]),
    "slb_rn_forward": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "slb_vit_workspace_bytes": (c_size_t, [c_void_p, c_int64]),
    "slb_vit_forward": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "slb_vit_trunk": (c_int, [c_void_p, c_void_p, c_int64, ctypes.c_int32, ctypes.c_int32, c_void_p, c_size_t, c_void_p]),
}


class SlbKernelTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 48), ("launches", c_int64), ("ms", c_double), ("flops", c_double),
                ("bytes", c_double)]


def profile_begin() -> None:
    check(load().slb_profile_begin(), "slb_profile_begin")


def profile_end() -> dict[str, dict[str, float]]:
    """Stop the live profiling started by :func:`profile_begin` and return {kernel: {launches, ms, flops, bytes}}
    (synchronises the recorded events)."""
    lib = load()
    check(lib.slb_profile_end(), "slb_profile_end")
    arr = (SlbKernelTime * 64)()
    n = c_int(0)
    check(lib.slb_profile_summary(ctypes.byref(arr), 64, ctypes.byref(n)), "slb_profile_summary")
    return {arr[i].name.decode(): {"launches": int(arr[i].launches), "ms": arr[i].ms, "flops": arr[i].flops,
                                   "bytes": arr[i].bytes} for i in range(n.value)}


class SlbVitLayer(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_out", "b_out", "ln2_g", "ln2_b", "w_fc", "b_fc", "w_proj", "b_proj")]


class SlbVitWeights(ctypes.Structure):
    _fields_ = (
        [(n, ctypes.c_int32) for n in ("image_size", "patch", "width", "layers", "heads", "mlp", "embed_dim", "act",
                                       "plane_fmt", "has_cls", "pool")]
        + [("ln_eps", c_float)]
        + [(n, c_void_p) for n in ("conv_w", "conv_b", "cls", "pos", "ln_pre_g", "ln_pre_b", "ln_post_g", "ln_post_b",
                                   "proj")]
        + [("layer", ctypes.POINTER(SlbVitLayer))]
        + [(n, c_void_p) for n in ("map_q", "map_w_kv", "map_b_kv", "map_w_out", "map_b_out", "map_ln_g", "map_ln_b",
                                   "map_w_fc", "map_b_fc", "map_w_proj", "map_b_proj")]
    )


class SlbConvBn(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("w", "scale", "shift")] + [(n, ctypes.c_int32) for n in ("cin", "cout", "ksize", "reserved")]


class SlbRnWeights(ctypes.Structure):
    _fields_ = (
        [(n, ctypes.c_int32) for n in ("image_size", "width", "heads", "out_dim")]
        + [("blocks", ctypes.c_int32 * 4)]
        + [(n, ctypes.c_int32) for n in ("plane_fmt", "n_convs")]
        + [("convs", ctypes.POINTER(SlbConvBn))]
        + [(n, c_void_p) for n in ("pos", "w_q", "b_q", "w_kv", "b_kv", "w_c", "b_c")]
    )


class SlbTextWeights(ctypes.Structure):
    _fields_ = (
        [(n, ctypes.c_int32) for n in ("context", "vocab", "width", "layers", "heads", "mlp", "embed_dim", "act", "plane_fmt")]
        + [("ln_eps", c_float)]
        + [(n, c_void_p) for n in ("tok_emb", "pos", "ln_final_g", "ln_final_b", "proj")]
        + [("layer", ctypes.POINTER(SlbVitLayer))]
        + [("non_causal", ctypes.c_int32), ("proj_b", c_void_p)]
    )


def lib_path() -> Path:
    return _LIB_PATH


def load(require_device: bool = False):
    """Load libslb200.so (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise SlbError(
                f"{_LIB_PATH} not found: build it with `python -m semanticlens_b200.csrc.build` "
                "(there is no CPU fallback for the B200 kernels)"
            )
        lib = ctypes.CDLL(str(_LIB_PATH))
        absent = [name for name in _PROTOS if not hasattr(lib, name)]
        if absent:
            raise SlbError(f"{_LIB_PATH} is stale: it does not export {absent[:4]}{' ...' if len(absent) > 4 else ''}; "
                           "rebuild it with `python -m semanticlens_b200.csrc.build --force`")
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    if require_device and not torch.cuda.is_available():
        raise SlbError("no CUDA device visible: the semanticlens_b200 hot path runs on B200 only (no CPU fallback)")
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().slb_last_error()
        raise SlbError(f"{what or 'libslb200'} failed with status {rc}: {msg.decode() if msg else ''}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError as e:
        raise TypeError(f"unsupported activation dtype {dt} (fp32, fp16, bf16 are supported)") from e


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise SlbError(f"{name} must be a CUDA tensor (got {t.device}); the B200 path has no CPU fallback")
