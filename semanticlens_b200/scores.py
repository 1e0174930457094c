"""Concept scores on B200 — drop-in for ``semanticlens.scores`` (reference scores.py:19-185).

Same signatures, shape heuristics, return dtypes and exceptions as the reference; the arithmetic runs in the libslb200
kernels (K6 cosine GEMM on the tensor cores, K7 clarity, K8 2-means polysemanticity, K9 redundancy). Tensors may live on
the CPU or on a CUDA device: CPU inputs are copied to the current CUDA device, the result comes back on the input's
device (the reference returns its result on ``V.device`` / ``x.device``). There is no CPU arithmetic path.
"""

from __future__ import annotations

import logging

import torch

from . import _native as N
from . import ops

logger = logging.getLogger(__name__)


def _to_gpu(t: torch.Tensor) -> torch.Tensor:
    if t.is_cuda:
        return t
    if not torch.cuda.is_available():
        raise N.SlbError("semanticlens_b200.scores needs a CUDA device (no CPU fallback)")
    return t.to("cuda", non_blocking=True)


@torch.inference_mode()
def clarity_score(V: torch.Tensor) -> torch.Tensor:
    """Clarity of concept examples: ``((|mean_k normalize(V)|^2 - 1/k) / (k-1)) * k`` (reference scores.py:19-47).

    V : (n_neurons, n_samples, n_features) (any leading dims) -> (n_neurons,) in [-1/(k-1), 1].
    """
    if V.ndim < 2:
        raise IndexError("clarity_score expects a tensor of shape (..., n_samples, n_features)")
    out = ops.clarity(_to_gpu(V))
    return out.to(V.device) if out.device != V.device else out


@torch.inference_mode()
def similarity_score(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Cosine similarity with the reference's shape heuristics (scores.py:85-128).

    * different shapes, ``x.shape[1] == y.shape[0]`` -> ``normalize(x) @ normalize(y)``  (checked FIRST, as upstream)
    * different shapes, ``x.shape[1] == y.shape[1]`` -> ``normalize(x) @ normalize(y).T``
    * otherwise different shapes -> ``ValueError("x and y must have the same shape")``
    * equal shapes -> row-wise ``F.cosine_similarity(x, y, dim=-1)``
    """
    dev = x.device
    if x.shape != y.shape:
        if x.ndim < 2 or y.ndim != 2:
            # the reference indexes x.shape[1] / y.shape[0] and lets matmul broadcast; matrices on the right only here
            raise ValueError("x and y must have the same shape")
        xg, yg = _to_gpu(x), _to_gpu(y)
        if yg.device != xg.device:
            yg = yg.to(xg.device)
        lead = xg.shape[:-1]  # (Q,) or, for a batched x, (A, Q): matmul broadcasting = flatten the rows, reshape back
        x2 = xg.reshape(-1, xg.shape[-1])
        if x.shape[1] == y.shape[0]:
            # normalize(x) (Q, D) @ normalize(y) (D, C): y is normalised along ITS rows, then used untransposed
            if xg.shape[-1] != y.shape[0]:
                raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x2.shape)} and {tuple(y.shape)})")
            yn = torch.nn.functional.normalize(yg.float(), dim=-1).t().contiguous()  # layout plumbing, (C, D)
            xp = ops.normalize_split_rows(x2.float())
            yp = ops.split_planes(_pad_cols(yn, xp.shape[2]), scale=ops.UNIT_ROW_PLANE_SCALE)
            n_pad = (yn.shape[0] + 7) // 8 * 8
            if n_pad != yn.shape[0]:
                yp = torch.cat([yp, torch.zeros((2, n_pad - yn.shape[0], yp.shape[2]), dtype=yp.dtype, device=yp.device)], 1)
            out, _ = ops.gemm_split(xp, yp.contiguous(), passes=3, alpha=1.0 / ops.UNIT_ROW_PLANE_SCALE**2)
            out = out[:, : yn.shape[0]]
        elif x.shape[1] == y.shape[1]:
            if xg.shape[-1] != y.shape[1]:
                raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x2.shape)} and {tuple(y.T.shape)})")
            out = ops.cosine_gemm(x2, yg)
        else:
            raise ValueError("x and y must have the same shape")
        out = out.reshape(*lead, out.shape[-1])
        return out.to(dev) if out.device != dev else out
    out = ops.cosine_rows(_to_gpu(x), _to_gpu(y).to(_to_gpu(x).device))
    return out.to(dev) if out.device != dev else out


def _pad_cols(t: torch.Tensor, kpad: int) -> torch.Tensor:
    if t.shape[1] == kpad:
        return t
    out = torch.zeros((t.shape[0], kpad), dtype=t.dtype, device=t.device)
    out[:, : t.shape[1]] = t
    return out


@torch.inference_mode()
def redundancy_score(cones: torch.Tensor) -> torch.Tensor:
    """Mean over neurons of the largest cosine similarity to another neuron (reference scores.py:51-81).

    cones : (n_neurons, n_features) -> scalar; batched (..., n, D) -> (...,) like the reference's matmul broadcast.
    """
    dev = cones.device
    g = _to_gpu(cones).float()
    lead = g.shape[:-2]
    g3 = g.reshape(-1, g.shape[-2], g.shape[-1])
    outs = [ops.redundancy(m) for m in g3]
    out = torch.stack(outs).reshape(lead) if lead else outs[0]
    return out.to(dev) if out.device != dev else out


@torch.inference_mode()
def polysemanticity_score(V: torch.Tensor, replace_empty_clusters: bool = True, random_state: int = 123,
                          n_clusters: int = 2) -> torch.Tensor:
    """1 - clarity of the k-means cluster centres of each neuron's examples (reference scores.py:132-185).

    The reference fits ``sklearn.cluster.KMeans(n_clusters, n_init=10, random_state=123)`` per neuron in a Python
    loop; K8 runs the same algorithm for the default ``n_clusters=2`` with up to 256 examples per neuron (k-means++ with sklearn's RandomState stream, Lloyd to strict convergence or
    tolerance, best of 10 by inertia) for every neuron on the GPU from the neuron's Gram matrix. Returns float64
    like the reference. Neurons whose smaller cluster has fewer than 2 members take the reference's fallback
    ``1 - mean_{i<10} clarity([mean(V), V[:, i]])`` when ``replace_empty_clusters``.
    """
    if V.ndim != 3:
        raise ValueError("polysemanticity_score expects (n_neurons, n_samples, n_features)")
    dev = V.device
    if n_clusters == 2 and V.shape[1] <= ops.POLYSEM_FAST_MAX_EXAMPLES:
        out = ops.polysem_2means(_to_gpu(V), random_state=random_state, replace_empty_clusters=replace_empty_clusters)
    else:
        # any number of clusters (2..8) / of examples: the same sklearn fit in sample space (K8g)
        out = ops.polysem_kmeans(_to_gpu(V), n_clusters=n_clusters, random_state=random_state,
                                 replace_empty_clusters=replace_empty_clusters)
    return out.to(dev) if out.device != dev else out
