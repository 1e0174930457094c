"""semanticlens_b200 — the B200 (sm_100a) implementation of SemanticLens's concept-database build path.

Public namespace mirrors ``semanticlens/__init__.py:35-47`` of the reference, so ``import semanticlens_b200 as sl``
gives ``sl.Lens``, ``sl.foundation_models``, ``sl.scores``, ``sl.utils`` and the three score functions; the component
visualizer lives in ``semanticlens_b200.component_visualization`` like upstream. Every kernel is in the in-tree
``csrc/libslb200.so`` (C-ABI: include/slb200.h); there is no CPU fallback.
"""

from . import foundation_models, scores, utils
from .lens import Lens
from .scores import clarity_score, polysemanticity_score, redundancy_score

__all__ = [
    "foundation_models",
    "scores",
    "utils",
    "Lens",
    "clarity_score",
    "polysemanticity_score",
    "redundancy_score",
]

__version__ = "0.1.0"
