"""Foundation models for the embed stage (reference: semanticlens/foundation_models/__init__.py:12-14).

``OpenClip`` keeps the reference's class name, constructor and ``AbstractVLM`` methods; its image tower is the
B200 ViT (``slb_vit_forward``: TMA-fed tcgen05 GEMMs on split planes + fp32 LayerNorm/attention kernels).
"""

from .base import AbstractVLM
from .clip import OpenClip, SigLipV2

__all__ = ["AbstractVLM", "OpenClip", "SigLipV2"]
