"""Plugin contract for vision-language foundation models (reference: semanticlens/foundation_models/base.py:12-120)."""

from __future__ import annotations

from abc import ABC, abstractmethod

import torch


class AbstractVLM(ABC):
    """What ``Lens`` and the component visualizers need from a foundation model."""

    @abstractmethod
    def encode_image(self, img_input: torch.Tensor) -> torch.Tensor:
        """(B, 3, H, W) preprocessed images -> (B, D) embeddings."""

    @abstractmethod
    def encode_text(self, text_input: torch.Tensor) -> torch.Tensor:
        """Tokenised text -> (B, D) embeddings."""

    @abstractmethod
    def preprocess(self, img) -> torch.Tensor:
        """PIL image(s) -> model-ready tensor on ``device``."""

    @abstractmethod
    def tokenize(self, txt):
        """Text -> token tensor on ``device``."""

    @property
    @abstractmethod
    def device(self):
        """Device the model lives on."""

    @abstractmethod
    def to(self, device):
        """Move the model."""
