"""ctypes glue for oracle/csrc/cpu_solver.c (TEST INFRASTRUCTURE / CPU baseline).

The numpy oracle builds the preconditioner (hierarchies, Schur complement); `CSolver` hands the
finished operators to the C + OpenMP solve loop so that the timed CPU leg of bench.py runs the same
algorithm as `oracle.krylov.gmres` + `oracle.blockpc.BlockPC` without Python inside the iteration.
Only the 2-way block preconditioner with `preonly` inner solves and the pressure-Schur split is
supported -- the benchmarked configuration; everything else stays in numpy.
tests/test_oracle_cport.py checks it against the numpy oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "cpu_solver.c")
_OUT = os.path.join(_HERE, "_cpu_solver.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", _SRC, "-o", _OUT, "-lm"],
                       check=True)
    return _OUT


class _Csr(C.Structure):
    _fields_ = [("nrows", C.c_int64), ("ncols", C.c_int64), ("indptr", C.c_void_p), ("indices", C.c_void_p),
                ("data", C.c_void_p)]


class _Level(C.Structure):
    _fields_ = [("A", _Csr), ("P", _Csr), ("R", _Csr), ("dinv", C.c_void_p), ("lmax", C.c_double)]


class _Amg(C.Structure):
    _fields_ = [("nlevels", C.c_int), ("levels", C.POINTER(_Level)), ("coarse_inv", C.c_void_p), ("degree", C.c_int),
                ("ratio", C.c_double), ("wx", C.c_void_p), ("wb", C.c_void_p), ("wr", C.c_void_p), ("wd", C.c_void_p)]


class _BlockPC(C.Structure):
    _fields_ = [("ns", C.c_int64), ("nf", C.c_int64), ("np", C.c_int64), ("Ks", C.POINTER(_Amg)),
                ("Kf", C.POINTER(_Amg)), ("Kp", C.POINTER(_Amg)), ("Mfps", _Csr), ("Apf", _Csr), ("Kv", C.POINTER(_Amg))]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_gmres_right.restype = C.c_int
        _lib.oracle_gmres_right.argtypes = [C.POINTER(_Csr), C.POINTER(_BlockPC), C.c_void_p, C.c_void_p, C.c_double,
                                            C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        _lib.oracle_amg_alloc.argtypes = [C.POINTER(_Amg)]
        _lib.oracle_amg_free.argtypes = [C.POINTER(_Amg)]
        _lib.oracle_amg_apply.argtypes = [C.POINTER(_Amg), C.c_void_p, C.c_void_p]
        _lib.oracle_cport_threads.restype = C.c_int
    return _lib


def threads() -> int:
    return int(lib().oracle_cport_threads())


def set_threads(t: int) -> None:
    lib().oracle_cport_set_threads(int(t))


class _Keep:
    """Owns the numpy arrays a C struct points into."""

    def __init__(self):
        self.refs = []

    def csr(self, M) -> _Csr:
        M = sp.csr_matrix(M)
        M.sort_indices()
        ip = np.ascontiguousarray(M.indptr, dtype=np.int64)
        ix = np.ascontiguousarray(M.indices, dtype=np.int32)
        da = np.ascontiguousarray(M.data, dtype=np.float64)
        self.refs += [ip, ix, da]
        return _Csr(M.shape[0], M.shape[1], ip.ctypes.data, ix.ctypes.data, da.ctypes.data)

    def vec(self, v) -> int:
        v = np.ascontiguousarray(v, dtype=np.float64)
        self.refs.append(v)
        return v.ctypes.data


class CAmg:
    """An `oracle.amg.SAAMG` hierarchy handed to the C V-cycle."""

    def __init__(self, amg):
        self.keep = _Keep()
        n = len(amg.levels)
        self.levels = (_Level * n)()
        empty = sp.csr_matrix((0, 0))
        for l, L in enumerate(amg.levels):
            last = l == n - 1
            self.levels[l] = _Level(self.keep.csr(L.A), self.keep.csr(empty if last else L.P),
                                    self.keep.csr(empty if last else L.R), self.keep.vec(L.dinv), float(L.lmax))
        inv = self.keep.vec(amg.levels[-1].inv) if amg.coarse_direct else None
        self.h = _Amg(n, self.levels, inv, int(amg.deg), float(amg.ratio), None, None, None, None)
        self.n = amg.levels[0].A.shape[0]
        lib().oracle_amg_alloc(C.byref(self.h))

    def __call__(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.n)
        lib().oracle_amg_apply(C.byref(self.h), b.ctypes.data, x.ctypes.data)
        return x

    def __del__(self):
        try:
            lib().oracle_amg_free(C.byref(self.h))
        except Exception:
            pass


class CResult:
    def __init__(self, x, its, reason, history):
        self.x, self.its, self.reason, self.history = x, its, reason, history
        self.rnorm = history[-1]


class CSolver:
    """GMRES(right) + 2-way BlockPC (preonly inner solves, SchurLower first='f') in C.

    `pc` is a finished `oracle.blockpc.BlockPC`; the system is permuted to [s | f | p] once here."""

    def __init__(self, sys_, pc):
        if pc.three_way or pc.anderson is not None:
            raise ValueError("C port covers the 2-way preconditioner without Anderson acceleration")
        fp = pc.k_fp
        cc = hasattr(fp, "k_visc")              # oracle.blockpc.SchurLowerCC (velocity first by construction)
        if not cc and getattr(fp, "first", None) != "f":
            raise ValueError("C port covers the pressure-Schur split (first='f')")
        k1 = fp.k_mass if cc else fp.k1
        for k in (pc.k_s, fp.k0, k1) + ((fp.k_visc,) if cc else ()):
            if k.type != "preonly":
                raise ValueError("C port covers preonly inner solves")
        self.keep = _Keep()
        # BlockPC works on [s | fp] with fp = [f | p] in 2-way numbering (lib/IndexSet.py:43-54)
        self.perm = np.concatenate([sys_.is_s, sys_.is_fp])
        A = sp.csr_matrix(sys_.A)[self.perm][:, self.perm].tocsr()
        self.A = self.keep.csr(A)
        self.ks, self.kf, self.kp = CAmg(pc.k_s.M), CAmg(fp.k0.M), CAmg(k1.M)
        self.kv = CAmg(fp.k_visc.M) if cc else None
        self.pc = _BlockPC(len(sys_.is_s), fp.nf, fp.np_, C.pointer(self.ks.h), C.pointer(self.kf.h),
                           C.pointer(self.kp.h), self.keep.csr(pc.Mfp_s), self.keep.csr(fp.A10),
                           C.pointer(self.kv.h) if cc else None)
        self.n = A.shape[0]

    def solve(self, b, rtol=1e-8, atol=0.0, max_it=100) -> CResult:
        bp = np.ascontiguousarray(np.asarray(b, dtype=np.float64)[self.perm])
        xp = np.zeros(self.n)
        hist = np.zeros(max_it + 1)
        reason = C.c_int(0)
        its = lib().oracle_gmres_right(C.byref(self.A), C.byref(self.pc), bp.ctypes.data, xp.ctypes.data, rtol, atol,
                                       max_it, hist.ctypes.data, C.byref(reason))
        x = np.empty(self.n)
        x[self.perm] = xp
        return CResult(x, its, reason.value, hist[: its + 1].copy())
