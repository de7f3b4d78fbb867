"""AAR: alternating Anderson-Richardson (reference lib/AAR.py:7-137), device-resident.

Same constructor signature and methods; `solve(b, sol)` runs AAR::solve in libporo.so
(csrc/solver.cu) including the reference's quirks (mixed norms in err_rel, window pairing,
history never reset between solves).  The least squares uses the Gram matrix F^T F computed
by one fused pass over the window (the reference gathers whole vectors to rank 0 and calls
np.linalg.qr, lib/AAR.py:85-105).
"""
from __future__ import annotations

import ctypes as C

from .. import _capi
from .backend import _tensor, get_context


class AAR:
    def __init__(self, order, p, omega, beta, matA, x0=None, pc=None, atol=1e-12, rtol=1e-8, maxiter=1000,
                 monitor_convergence=False, ctx=None):
        self.order, self.p, self.omega, self.beta = order, p, omega, beta
        self.matA, self.pc = matA, pc
        self.atol, self.rtol, self.maxiter = atol, rtol, maxiter
        self.monitor_convergence = monitor_convergence
        self.ctx = ctx or get_context()
        if pc is None:
            raise ValueError("AAR needs a preconditioner (the reference's pc=None branch is unreachable, lib/AAR.py:33-38)")
        h = C.c_void_p()
        _capi.check(self.ctx.lib.poro_aar_create(self.ctx.h, matA.mat().handle, pc.handle, int(order), int(p),
                                                 float(omega), float(beta), float(atol), float(rtol), int(maxiter),
                                                 int(bool(monitor_convergence)), C.byref(h)))
        self.h = h
        self._keep = (matA, pc)   # borrowed by the library, see include/poro.h
        self.it = 0

    def set_up(self):
        pass

    def solve(self, b, sol):
        its = C.c_int()
        _capi.check(self.ctx.lib.poro_aar_solve(self.h, _capi._ptr(_tensor(b)), _capi._ptr(_tensor(sol)), C.byref(its)))
        self.it = its.value
        return self.it

    def getIterationNumber(self):
        return self.it

    def residual_history(self):
        n = C.c_int()
        buf = (C.c_double * (self.maxiter + 2))()
        _capi.check(self.ctx.lib.poro_aar_residual_history(self.h, buf, self.maxiter + 2, C.byref(n)))
        return list(buf)[: n.value]

    def __del__(self):
        try:
            if self.h:
                self.ctx.lib.poro_aar_destroy(self.h)
                self.h = None
        except Exception:
            pass
