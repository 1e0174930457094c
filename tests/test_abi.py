"""The C-ABI shared library loads without a GPU and exports every symbol include/slb200.h declares."""

import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "slb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slb_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    from semanticlens_b200 import _native

    lib = ctypes.CDLL(str(_native.lib_path()))
    syms = declared_symbols()
    assert len(syms) >= 8
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in slb200.h but not exported: {missing}"


def test_python_prototypes_cover_header():
    from semanticlens_b200 import _native

    assert set(declared_symbols()) == set(_native._PROTOS), set(declared_symbols()) ^ set(_native._PROTOS)


def test_version_and_error_string():
    from semanticlens_b200 import _native

    lib = _native.load()
    assert lib.slb_version() == 100
    assert isinstance(lib.slb_last_error(), bytes)


def test_argument_validation_without_device():
    """Bad arguments are rejected before any CUDA call."""
    from semanticlens_b200 import _native

    lib = _native.load()
    assert lib.slb_agg_reduce(None, 0, 7, 1, 1, 1, 0, 0, None, None) == -1
    assert b"layout" in lib.slb_last_error()
    assert lib.slb_topk_update(None, 0, 4, 4, None, 0, None, None, 4, None) == -1
    assert lib.slb_topk_update(None, 0, 4, 4, None, 0, None, None, 0, None) == 0  # k = 0 is a legal no-op
