"""PETSc-semantics Krylov restatements in numpy (TEST INFRASTRUCTURE).

The reference's outer solve is `KSP.solve` configured at lib/Solver.py:92-102 (prefix
`global_`, `setTolerances(rtol, atol, 1e20, maxiter)`, `setGMRESRestart(maxiter)`, zero
initial guess because `setInitialGuessNonzero` is commented out at :94); inner solves are
KSPs with prefixes s_ f_ p_ fp_ diff_ (lib/Preconditioner.py:77-100).  PETSc's source is
not in the reference tree; the semantics restated here are the documented ones
(SURVEY.md §8c item 7):

* GMRES: zero guess; left PC (default) monitors ||M^-1 r||, right PC monitors ||r||;
  classical Gram-Schmidt without refinement (or with one refinement pass, `cgs2=True`);
  Givens-updated residual estimate; `its` counts Arnoldi steps; restart length m.
* CG: preconditioned CG, monitored norm preconditioned (default) / unpreconditioned / natural.
* Default convergence test: stop when rnorm <= max(rtol * rnorm_0, atol); diverge when
  rnorm >= dtol * rnorm_0; reason -3 when its >= max_it.
Reasons follow PETSc's KSPConvergedReason codes: 2 rtol, 3 atol, -3 its, -4 dtol, -8 indefinite.
"""
from __future__ import annotations

import numpy as np

CONVERGED_RTOL, CONVERGED_ATOL, CONVERGED_ITS = 2, 3, 4
DIVERGED_ITS, DIVERGED_DTOL, DIVERGED_BREAKDOWN, DIVERGED_INDEFINITE = -3, -4, -5, -8


class KSPResult:
    def __init__(self, x, its, reason, history):
        self.x, self.its, self.reason, self.history = x, its, reason, history

    @property
    def rnorm(self):
        return self.history[-1] if self.history else 0.0


def _converged(rnorm, it, state, rtol, atol, dtol):
    if it == 0:
        state["rnorm0"] = rnorm
        state["ttol"] = max(rtol * rnorm, atol)
    if rnorm != rnorm:
        return -9
    if rnorm <= state["ttol"]:
        return CONVERGED_ATOL if rnorm < atol else CONVERGED_RTOL
    if rnorm >= dtol * state["rnorm0"]:
        return DIVERGED_DTOL
    return 0


def gmres(A, b, M=None, rtol=1e-5, atol=1e-50, dtol=1e5, max_it=10000, restart=30, pc_side="left",
          cgs2=False, x0=None, flexible=False) -> KSPResult:
    """A, M: callables v -> A v, v -> M^-1 v.  flexible=True is PETSc's KSPFGMRES (right PC,
    z_j = M^-1 v_j stored, x = x0 + Z y), the correct choice when M is a nonlinear inner solve."""
    if flexible:
        pc_side = "right"
    n = len(b)
    M = M or (lambda v: v)
    x = np.zeros(n) if x0 is None else x0.copy()
    hist, state = [], {}
    its, reason = 0, 0
    right = pc_side == "right"
    first = True
    while reason == 0:
        r = b - A(x) if (x0 is not None or not first) else b.copy()
        if not right:
            r = M(r)
        beta = np.linalg.norm(r)
        if first:
            hist.append(beta)
            reason = _converged(beta, 0, state, rtol, atol, dtol)
            first = False
            if reason:
                break
        if beta == 0.0:
            reason = CONVERGED_ATOL
            break
        m = restart
        V = np.zeros((m + 1, n))
        Z = np.zeros((m, n)) if flexible else None
        H = np.zeros((m + 1, m))
        cs, sn = np.zeros(m), np.zeros(m)
        g = np.zeros(m + 1)
        g[0] = beta
        V[0] = r / beta
        j = 0
        while reason == 0 and j < m and its < max_it:
            if flexible:
                Z[j] = M(V[j])
                w = A(Z[j])
            else:
                w = A(M(V[j])) if right else M(A(V[j]))
            h = V[: j + 1] @ w
            w = w - h @ V[: j + 1]
            if cgs2:
                h2 = V[: j + 1] @ w
                w = w - h2 @ V[: j + 1]
                h = h + h2
            hn = np.linalg.norm(w)
            H[: j + 1, j] = h
            H[j + 1, j] = hn
            if hn > 0:
                V[j + 1] = w / hn
            # apply previous rotations, then the new one
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            den = np.hypot(H[j, j], H[j + 1, j])
            if den == 0.0:
                reason = DIVERGED_BREAKDOWN
                break
            cs[j], sn[j] = H[j, j] / den, H[j + 1, j] / den
            H[j, j] = den
            H[j + 1, j] = 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            its += 1
            j += 1
            res = abs(g[j])
            hist.append(res)
            reason = _converged(res, its, state, rtol, atol, dtol)
            if hn == 0.0 and reason == 0:      # happy breakdown
                reason = CONVERGED_RTOL
        if j > 0:
            y = np.linalg.solve(np.triu(H[:j, :j]), g[:j])
            if flexible:
                x = x + y @ Z[:j]
            else:
                dx = y @ V[:j]
                x = x + (M(dx) if right else dx)
        if reason == 0 and its >= max_it:
            reason = DIVERGED_ITS
    return KSPResult(x, its, reason, hist)


def cg(A, b, M=None, rtol=1e-5, atol=1e-50, dtol=1e5, max_it=10000, norm_type="preconditioned") -> KSPResult:
    n = len(b)
    M = M or (lambda v: v)
    x = np.zeros(n)
    r = b.copy()
    z = M(r)
    hist, state = [], {}

    def nrm(r, z):
        if norm_type == "preconditioned":
            return np.linalg.norm(z)
        if norm_type == "unpreconditioned":
            return np.linalg.norm(r)
        return np.sqrt(abs(r @ z))

    dp = nrm(r, z)
    hist.append(dp)
    reason = _converged(dp, 0, state, rtol, atol, dtol)
    its = 0
    p = None
    betaold = 1.0
    while reason == 0:
        beta = r @ z
        if beta == 0.0:
            reason = CONVERGED_ATOL
            break
        p = z.copy() if p is None else z + (beta / betaold) * p
        w = A(p)
        dpi = p @ w
        if dpi <= 0.0 or dpi != dpi:
            reason = DIVERGED_INDEFINITE
            break
        a = beta / dpi
        x += a * p
        r -= a * w
        z = M(r)
        betaold = beta
        its += 1
        dp = nrm(r, z)
        hist.append(dp)
        reason = _converged(dp, its, state, rtol, atol, dtol)
        if reason == 0 and its >= max_it:
            reason = DIVERGED_ITS
    return KSPResult(x, its, reason, hist)


def preonly(A, b, M=None, **kw) -> KSPResult:
    M = M or (lambda v: v)
    return KSPResult(M(b), 1, CONVERGED_ITS, [])


class InnerKSP:
    """A configured inner solve y = KSP(A, M) \\ x, counting iterations."""

    def __init__(self, A, M=None, ksp_type="preonly", **opts):
        self.A = (lambda v, A=A: A @ v) if not callable(A) else A
        self.M, self.type, self.opts = M, ksp_type, opts
        self.total_its = 0
        self.calls = 0

    def __call__(self, x):
        fn = {"preonly": preonly, "cg": cg, "gmres": gmres}[self.type]
        res = fn(self.A, x, self.M, **self.opts)
        self.total_its += res.its
        self.calls += 1
        return res.x
