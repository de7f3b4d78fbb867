"""Threaded CSR matvec for the CPU baseline (TEST INFRASTRUCTURE): wraps scipy CSR matrices so that
`M @ x` runs the OpenMP kernel of oracle/csrc/omp_kernels.c.  Falls back to scipy if the library
has not been built."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_omp_kernels.so")
        if not os.path.exists(path):
            try:
                from . import build_c
                build_c.build()
            except Exception:
                return None
        _lib = C.CDLL(path)
        _lib.oracle_omp_threads.restype = C.c_int
        _lib.oracle_csr_matvec.argtypes = [C.c_int64] + [C.c_void_p] * 5
    return _lib


def threads() -> int:
    l = lib()
    return int(l.oracle_omp_threads()) if l else 1


class OmpCsr:
    def __init__(self, M):
        M = sp.csr_matrix(M)
        self.shape = M.shape
        self.nnz = M.nnz
        self.indptr = np.ascontiguousarray(M.indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(M.indices, dtype=np.int32)
        self.data = np.ascontiguousarray(M.data, dtype=np.float64)
        self._scipy = M
        self._lib = lib()

    def __matmul__(self, x):
        if self._lib is None or x.ndim != 1:
            return self._scipy @ x
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.shape[0])
        self._lib.oracle_csr_matvec(self.shape[0], self.indptr.ctypes.data, self.indices.ctypes.data, self.data.ctypes.data,
                                    x.ctypes.data, y.ctypes.data)
        return y

    def diagonal(self):
        return self._scipy.diagonal()


def accelerate(obj, _seen=None):
    """Replace every scipy CSR matrix reachable from `obj` (BlockPC, SchurLower, InnerKSP, SAAMG levels)
    by an OmpCsr.  Call after set-up; only `M @ vector` is used afterwards."""
    _seen = _seen if _seen is not None else set()
    if obj is None or id(obj) in _seen or isinstance(obj, (np.ndarray, OmpCsr, str, int, float)):
        return obj
    _seen.add(id(obj))
    if isinstance(obj, (list, tuple)):
        for o in obj:
            accelerate(o, _seen)
        return obj
    if not hasattr(obj, "__dict__") or callable(obj) and not hasattr(obj, "levels") and not hasattr(obj, "k0") \
            and not hasattr(obj, "opts") and not hasattr(obj, "k_s"):
        return obj
    for k, v in list(vars(obj).items()):
        if sp.issparse(v):
            setattr(obj, k, OmpCsr(v))
        else:
            accelerate(v, _seen)
    return obj
