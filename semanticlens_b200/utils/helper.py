"""Naming helpers. The fallback name is part of the cache directory grammar, so it must equal the reference's
(semanticlens/utils/helper.py:11-18): ``ClassName-<sha256(str(obj)) as a decimal int>``."""

from __future__ import annotations

import hashlib


def _string_hash(s: str) -> int:
    return int(hashlib.sha256(s.encode()).hexdigest(), 16)


def get_fallback_name(obj) -> str:
    return f"{type(obj).__name__}-{_string_hash(str(obj))}"
