"""Logging on the package logger; level from SEMANTICLENS_LOG_LEVEL like the reference (utils/log_setup.py:61-93)."""

from __future__ import annotations

import logging
import os

_LOGGER_NAME = "semanticlens_b200"
logging.getLogger(_LOGGER_NAME).addHandler(logging.NullHandler())


def setup_colored_logging(level: str | int | None = None) -> logging.Logger:
    """Attach one stream handler to the package logger (idempotent)."""
    level = level or os.environ.get("SEMANTICLENS_LOG_LEVEL", "INFO")
    logger = logging.getLogger(_LOGGER_NAME)
    logger.setLevel(level)
    if not any(isinstance(h, logging.StreamHandler) and not isinstance(h, logging.NullHandler) for h in logger.handlers):
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(name)s: %(message)s"))
        logger.addHandler(handler)
    return logger
