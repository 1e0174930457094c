"""ORACLE helper — import the real reference (read-only, /root/reference) in THIS container.

The reference imports `crp` and `zennit` eagerly (component_visualization/__init__.py:17 -> relevance_based.py:16-19)
although only the out-of-scope relevance visualizer uses them; they are not installed, so inert stubs are injected
into sys.modules first. Used by oracle/make_golden.py and by the optional `reference`-marked tests; never at run
time on the GPU box (the reference does not travel).
"""

from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SLB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "semanticlens"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def import_reference():
    """Return the imported reference package `semanticlens` (raises ImportError if not reachable)."""
    if not available():
        raise ImportError(f"reference not found under {REFERENCE_ROOT}")

    class _Any:  # base class / callable placeholder
        def __init__(self, *a, **k):
            pass

    _stub("crp")
    _stub("crp.concepts", ChannelConcept=_Any)
    _stub("crp.helper", load_maximization=lambda *a, **k: None)
    _stub("crp.visualization", FeatureVisualization=_Any)
    _stub("crp.image", get_crop_range=lambda *a, **k: None, imgify=lambda *a, **k: None)
    _stub("zennit")
    _stub("zennit.composites", EpsilonPlusFlat=_Any)
    _stub("zennit.core", stabilize=lambda x, *a, **k: x)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import semanticlens  # noqa: E402

    return semanticlens
