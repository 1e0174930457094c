"""Plugin contract for component visualizers (reference: semanticlens/component_visualization/base.py:16-183)."""

from __future__ import annotations

from abc import ABC, abstractmethod

import torch


class AbstractComponentVisualizer(ABC):
    """What ``Lens`` needs from a visualizer: ``run``, ``_compute_concept_db``, ``get_max_reference``,
    ``to``, ``metadata``, ``caching``, ``storage_dir``, ``device``."""

    def __init__(self, model: torch.nn.Module, device: str | torch.device | None = None):
        self.model = model
        self.model.to(device or next(model.parameters()).device)

    @abstractmethod
    def run(self, *args, **kwargs) -> None:
        """Sweep the dataset and collect what identifies each component's concept (e.g. top-k samples)."""
        raise NotImplementedError

    @abstractmethod
    def _compute_concept_db(self, fm, **kwargs) -> dict[str, torch.Tensor]:
        """Return ``{layer: (n_components, n_samples, embed_dim)}`` using foundation model ``fm``."""
        raise NotImplementedError

    @abstractmethod
    def get_max_reference(self, layer_name) -> torch.Tensor:
        """``(n_components, n_samples)`` int64 dataset indices of the maximally activating samples."""
        raise NotImplementedError

    def to(self, device: str | torch.device):
        self.model.to(device)
        return self

    @property
    def metadata(self) -> dict[str, str]:
        raise NotImplementedError

    @property
    @abstractmethod
    def caching(self) -> bool:
        raise NotImplementedError

    @property
    @abstractmethod
    def storage_dir(self):
        raise NotImplementedError

    @property
    def device(self):
        return next(self.model.parameters()).device
