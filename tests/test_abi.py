"""CPU checks of the drop-in boundary: the library loads and exports every symbol include/uml_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "uml_b200.h").read_text()
    return sorted(set(re.findall(r"^UML_API [^;(]*?\b(uml_[a-z0-9_]+)\(", text, flags=re.M)))


@pytest.fixture(scope="module")
def built_lib():
    from unionml_b200 import _build

    return _build.build()


def test_header_declares_expected_surface():
    syms = _declared_symbols()
    for name in ("uml_engine_create", "uml_linear_load", "uml_stage_rows", "uml_linear_predict",
                 "uml_linear_predict_peers", "uml_linear_predict_host", "uml_last_error"):
        assert name in syms


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(str(built_lib))
    missing = [s for s in _declared_symbols() if not hasattr(handle, s)]
    assert not missing, missing
    handle.uml_abi_version.restype = ctypes.c_int
    assert handle.uml_abi_version() == 2


def test_python_binding_covers_the_header(built_lib):
    from unionml_b200 import _native

    assert sorted(_native.SIGNATURES) == _declared_symbols()
    _native.lib()  # argtypes/restype resolve for every symbol


def test_no_device_fails_loudly(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from unionml_b200.engine import Engine

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under unionml_b200/ may import, link or execute it."""
    for path in (ROOT / "unionml_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".h"}:
            text = path.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
            assert "liboracle" not in text and "oracle/" not in text, path


def test_pylist_helper_fills_list_from_labels():
    """csrc_host/uml_pylist.c (host glue of the `List[float]` contract): slots reference the per-class float objects,
    ranges and labels are checked, the replaced entries are released."""
    import ctypes as C
    import sys

    import numpy as np

    from unionml_b200 import _build, _native

    _build.build_pylist()
    h = _native.pylist()
    assert h is not None
    table = [float(c) for c in (3, 5, 8)]
    labels = np.array([2, 0, 1, 1, 2], dtype=np.int32)
    out = [None] * 7
    before = sys.getrefcount(table[1])
    assert h.uml_list_fill_from_labels(out, 1, C.c_void_p(labels.ctypes.data), 5, table) == 0
    after_first = sys.getrefcount(table[1])
    assert after_first == before + 2  # two rows carry class 1
    h.uml_list_fill_from_labels(out, 1, C.c_void_p(labels.ctypes.data), 5, table)  # overwrite: old references dropped
    after_second = sys.getrefcount(table[1])  # (measured outside `assert`: pytest's rewriting keeps temporaries alive)
    assert after_second == after_first
    assert out == [None, 8.0, 3.0, 5.0, 5.0, 8.0, None]
    assert [type(v) for v in out[1:6]] == [float] * 5 and out[3] is table[1]
    with pytest.raises(IndexError):
        h.uml_list_fill_from_labels(out, 4, C.c_void_p(labels.ctypes.data), 5, table)
    bad = np.array([0, 3], dtype=np.int32)
    with pytest.raises(ValueError):
        h.uml_list_fill_from_labels(out, 0, C.c_void_p(bad.ctypes.data), 2, table)
    with pytest.raises(TypeError):
        h.uml_list_fill_from_labels(tuple(out), 0, C.c_void_p(labels.ctypes.data), 1, table)
