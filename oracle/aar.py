"""AAR (alternating Anderson-Richardson) and the inner Anderson acceleration, restated in
numpy with the reference's control flow and quirks (TEST INFRASTRUCTURE).

lib/AAR.py:46-128   AAR.solve            lib/AAR.py:133-137  update_residual
lib/AndersonAcceleration.py:19-78        AndersonAcceleration.get_next_vector

Quirks kept on purpose: error0 is the UNpreconditioned ||b - A x0|| while
err_abs is the preconditioned ||M^-1(b - A x)||; F[0] mixes both; while the window fills,
X[i] is paired with F[i] although it belongs to F[i+1] and the newest coefficient is
dropped; the convergence test sees the residual of the previous iterate; least squares by
np.linalg.qr exactly as AAR.py:102-105.
"""
from __future__ import annotations

import numpy as np


class AAR:
    def __init__(self, order, p, omega, beta, A, pc, atol=1e-12, rtol=1e-8, maxiter=1000, monitor=False,
                 lstsq="qr"):
        self.order, self.p, self.omega, self.beta = order, p, omega, beta
        self.A, self.pc = A, pc
        self.atol, self.rtol, self.maxiter, self.monitor = atol, rtol, maxiter, monitor
        self.F, self.X = [], []                                  # never reset between solves (AAR.py:20-22)
        self.lstsq = lstsq
        self.history = []
        self.types = []

    def _alpha(self, fk):
        F = np.vstack(self.F).T                                  # AAR.py:102 (F0 == F on one rank)
        if self.lstsq == "qr":
            Q, R = np.linalg.qr(F)
            return np.linalg.solve(R, -Q.T @ fk)                 # AAR.py:103-105
        G = F.T @ F                                              # Gram variant (what the GPU path does)
        return gram_solve(G, -(F.T @ fk))

    def solve(self, b):
        xk = np.zeros_like(b)                                    # :48-50
        fk = b - self.A(xk)                                      # :55-56 (unpreconditioned)
        error0 = np.linalg.norm(fk)                              # :67
        err_abs, err_rel, it = error0, 1.0, 0
        self.history = [err_abs]
        while err_abs > self.atol and err_rel > self.rtol and it < self.maxiter:   # :73
            d_fk, d_xk = fk.copy(), xk.copy()                    # :75-76
            fk = self.pc(b - self.A(xk))                         # :77, :133-137
            d_fk = fk - d_fk                                     # :78
            self.F.append(d_fk)
            if len(self.F) > self.order:
                self.F.pop(0)
            nf = np.linalg.norm(fk)
            if nf < 1e-14:                                       # :91
                typ = ""
            elif it == 0 or self.order == 0 or (it + 1) / self.p % 1 > 0:   # :94
                typ = "R"
                xk = xk + self.omega * fk
            else:
                typ = "A"
                mk = min(self.order, it)                         # :99
                alpha = self._alpha(fk)
                xk = xk + self.beta * fk                         # :109
                for i in range(mk):                              # :110-111
                    xk = xk + alpha[i] * (self.X[i] + self.beta * self.F[i])
            d_xk = xk - d_xk                                     # :113
            self.X.append(d_xk)
            if len(self.X) > self.order:
                self.X.pop(0)
            err_abs = nf                                         # :117
            err_rel = err_abs / error0
            it += 1
            self.history.append(err_abs)
            self.types.append(typ)
        self.it = it
        return xk

    def getIterationNumber(self):
        return self.it


def gram_solve(G, rhs):
    """Solve G a = rhs for a symmetric PSD Gram matrix, robust to rank deficiency.

    Jacobi-scaled eigen-decomposition with relative truncation: identical on host and in the
    GPU path's host-side m x m solve (m <= order <= ~10).
    """
    d = np.sqrt(np.maximum(np.diag(G), 1e-300))
    Gs = G / np.outer(d, d)
    w, V = np.linalg.eigh(Gs)
    keep = w > 1e-14 * w.max()
    y = V[:, keep] @ ((V[:, keep].T @ (rhs / d)) / w[keep])
    return y / d


class AndersonAcceleration:
    """lib/AndersonAcceleration.py:19-78: Anderson(order) on successive PC outputs g_k."""

    def __init__(self, order, lstsq="qr"):
        self.order, self.k = order, 0
        self.F, self.X = [], []
        self.lstsq = lstsq

    def get_next_vector(self, gk):
        if self.k == 0:                                          # :21-34
            self.xk = np.zeros_like(gk)
            self.fk = np.zeros_like(gk)
        d_fk, d_xk = self.fk.copy(), self.xk.copy()              # :36-37
        self.fk = gk - self.xk                                   # :39-40
        mk = min(self.k, self.order)                             # :42
        if mk > 0:
            d_fk = self.fk - d_fk                                # :44
            if np.linalg.norm(d_fk) < 1e-12:                     # :45-47
                self.k -= 1
                self.xk = gk.copy()
            else:
                self.F.append(d_fk)
                if len(self.F) > self.order:
                    self.F.pop(0)
                F = np.vstack(self.F).T
                if self.lstsq == "qr":
                    Q, R = np.linalg.qr(F)
                    alpha = np.linalg.solve(R, -Q.T @ self.fk)   # :60-63
                else:
                    alpha = gram_solve(F.T @ F, -(F.T @ self.fk))
                self.xk = self.xk + self.fk                      # :67
                for i in range(mk):                              # :68-69
                    self.xk = self.xk + alpha[i] * (self.X[i] + self.F[i])
        else:
            self.xk = gk.copy()                                  # :71
        d_xk = self.xk - d_xk                                    # :73
        self.X.append(d_xk)
        if len(self.X) > self.order:
            self.X.pop(0)
        self.k += 1
        return self.xk.copy()                                    # :78 (gk overwritten)
