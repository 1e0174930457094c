"""Component visualizers (reference: semanticlens/component_visualization/__init__.py:16-22).

``RelevanceComponentVisualizer`` (CRP/LRP based, marked "currently broken" upstream, relevance_based.py:27) is not
part of the concept-database build path and is not provided.
"""

from .activation_based import ActivationComponentVisualizer, MissingNameWarning

__all__ = ["ActivationComponentVisualizer", "MissingNameWarning"]
