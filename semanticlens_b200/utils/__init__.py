"""Host-side helpers (reference: semanticlens/utils/__init__.py)."""

from .helper import get_fallback_name
from .log_setup import setup_colored_logging

__all__ = ["get_fallback_name", "setup_colored_logging"]
