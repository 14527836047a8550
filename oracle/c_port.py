"""Oracle (test infrastructure only): loader for the plain-C restatement ``oracle/linear_predict.c``.

``build()`` compiles it with ``gcc -O3 -fopenmp`` into ``oracle/_build/liboracle_linear.so`` (git-ignored, travels with
the repo snapshot).  Only tests and ``bench.py``'s CPU legs may call this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "linear_predict.c"
LIB = HERE / "_build" / "liboracle_linear.so"
_lib = None


def build(force: bool = False) -> Path:
    LIB.parent.mkdir(exist_ok=True)
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        subprocess.run(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        h = C.CDLL(str(LIB))
        for name in ("oracle_linear_predict_f64", "oracle_linear_predict_f32"):
            fn = getattr(h, name)
            fn.restype = None
            fn.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        h.oracle_num_threads.restype = C.c_int
        _lib = h
    return _lib


def predict_indices(X: np.ndarray, coef, intercept) -> np.ndarray:
    """Class indices per row (int32), float64 arithmetic, first-maximum tie rule."""
    coef = np.ascontiguousarray(np.atleast_2d(coef), dtype=np.float64)
    intercept = np.ascontiguousarray(np.atleast_1d(intercept), dtype=np.float64)
    X = np.asarray(X)
    if X.dtype == np.float32:
        X = np.ascontiguousarray(X)
        fn = lib().oracle_linear_predict_f32
    else:
        X = np.ascontiguousarray(X, dtype=np.float64)
        fn = lib().oracle_linear_predict_f64
    if X.ndim != 2 or X.shape[1] != coef.shape[1]:
        raise ValueError(f"X has shape {X.shape}, model expects {coef.shape[1]} features")
    out = np.empty(X.shape[0], dtype=np.int32)
    fn(X.ctypes.data, X.shape[0], X.shape[1], coef.ctypes.data, intercept.ctypes.data, coef.shape[0], out.ctypes.data)
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())
