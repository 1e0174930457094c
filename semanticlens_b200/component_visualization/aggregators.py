"""Aggregation functions for hooked activation maps — B200 implementation.

Same names, argument checks and error messages as the reference
(semanticlens/component_visualization/aggregators.py:38-244); the function ``__name__`` is part of the on-disk
cache file name there (aggregators.py:27,32) and therefore here too.

Called directly, each function returns the (batch, components) aggregate **on the CPU** like the reference
(this synchronises). Inside the activation sweep the hook does not call them: it reads the ``_slb_op`` /
``_slb_kind`` / ``_slb_token`` tags and runs the fused K1+K2 kernels on the device without any host sync
(``ActMaxCache._get_hook``).
"""

from __future__ import annotations

import torch

from .. import _native as N
from .. import ops

_ERROR_MESSAGE = f"(Select or implement a different aggregation function in {__file__}.)"


def _run(tensor: torch.Tensor, op: int, kind: str, token: int = 0) -> torch.Tensor:
    if isinstance(tensor, tuple):  # kept for parity; unreachable exactly as in the reference (ndim is read first)
        tensor = tensor[0]
    dev = tensor if tensor.is_cuda else tensor.cuda()  # plumbing only: the reduction itself always runs on the GPU
    return ops.agg_reduce(dev, op, kind, token).to(tensor.dtype if tensor.is_floating_point() else torch.float32).cpu()


def _tag(fn, op: int, kind: str, token: int = 0):
    fn._slb_op, fn._slb_kind, fn._slb_token = op, kind, token
    return fn


def aggregate_conv_mean(tensor: torch.Tensor) -> torch.Tensor:
    """Mean over the spatial dimensions of a (batch, channels, height, width) tensor -> (batch, channels)."""
    if tensor.ndim != 4:
        raise ValueError("Input tensor should be 4D. \n" + _ERROR_MESSAGE)
    return _run(tensor, N.AGG_MEAN, "conv")


def aggregate_conv_max(tensor: torch.Tensor) -> torch.Tensor:
    """Max over the spatial dimensions of a (batch, channels, height, width) tensor -> (batch, channels)."""
    if tensor.ndim != 4:
        raise ValueError("Input tensor should be 4D. \n" + _ERROR_MESSAGE)
    return _run(tensor, N.AGG_MAX, "conv")


def aggregate_transformer_mean(tensor: torch.Tensor) -> torch.Tensor:
    """Mean over the token dimension of a (batch, tokens, features) tensor -> (batch, features)."""
    if tensor.ndim != 3:
        raise ValueError("Input tensor should be 3D. \n" + _ERROR_MESSAGE)
    return _run(tensor, N.AGG_MEAN, "tokens")


def aggregate_transformer_absmean(tensor: torch.Tensor) -> torch.Tensor:
    """Mean of absolute values over the token dimension -> (batch, features)."""
    if tensor.ndim != 3:
        raise ValueError("Input tensor should be 3D. \n" + _ERROR_MESSAGE)
    return _run(tensor, N.AGG_ABSMEAN, "tokens")


def aggregate_transformer_max(tensor: torch.Tensor) -> torch.Tensor:
    """Max over the token dimension -> (batch, features)."""
    if tensor.ndim != 3:
        raise ValueError("Input tensor should be 3D. \n" + _ERROR_MESSAGE)
    return _run(tensor, N.AGG_MAX, "tokens")


def aggregate_transformer_absmax(tensor: torch.Tensor) -> torch.Tensor:
    """Max of absolute values over the token dimension -> (batch, features)."""
    if tensor.ndim != 3:
        raise ValueError("Input tensor should be 3D. \n" + _ERROR_MESSAGE)
    return _run(tensor, N.AGG_ABSMAX, "tokens")


_tag(aggregate_conv_mean, N.AGG_MEAN, "conv")
_tag(aggregate_conv_max, N.AGG_MAX, "conv")
_tag(aggregate_transformer_mean, N.AGG_MEAN, "tokens")
_tag(aggregate_transformer_absmean, N.AGG_ABSMEAN, "tokens")
_tag(aggregate_transformer_max, N.AGG_MAX, "tokens")
_tag(aggregate_transformer_absmax, N.AGG_ABSMAX, "tokens")


def get_aggregate_transformer_special_token(token_position: int):
    """Return a function that extracts the values at ``token_position`` of a (batch, tokens, features) tensor."""

    def aggregate_transformer_special_token(tensor: torch.Tensor) -> torch.Tensor:
        if tensor.ndim != 3:
            raise ValueError("Input tensor should be 3D. \n" + _ERROR_MESSAGE)
        if not -tensor.shape[1] <= token_position < tensor.shape[1]:
            raise IndexError(
                f"index {token_position} is out of bounds for dimension 1 with size {tensor.shape[1]}"
            )
        return _run(tensor, N.AGG_TOKEN, "tokens", token_position)

    return _tag(aggregate_transformer_special_token, N.AGG_TOKEN, "tokens", token_position)
