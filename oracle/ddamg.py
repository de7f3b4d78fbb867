"""CPU twin of the ROW-PARTITIONED preconditioners of round 1 (TEST INFRASTRUCTURE).

What libporo.so does on R ranks (csrc/capi.cu: make_inner, csrc/amg.cu: Amg::cycle with fine_mat):
  * every AMG hierarchy is built from the rank's owned diagonal block (principal sub-matrix);
  * level-0 Chebyshev smoothing and residuals use the TRUE distributed operator;
  * restriction, coarse levels and prolongation are rank-local;
  * the selfp Schur complement is assembled from owned parts only and gets a rank-local hierarchy.
Emulated here on one process with index sets per rank, so that multi-GPU iteration counts have something to be
compared with (tests/test_oracle_ddamg.py, profiles/r1_scaling.md).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .amg import SAAMG


class DDAmg:
    """Rank-local hierarchies, global level-0 smoothing (one V-cycle)."""

    def __init__(self, A, bs, B, parts, **kw):
        self.A = sp.csr_matrix(A)
        self.parts = parts
        self.loc = [SAAMG(self.A[p][:, p], bs, None if B is None else B[p], **kw) for p in parts]
        d = self.A.diagonal()
        self.dinv = 1.0 / np.where(d != 0, d, 1.0)
        # each rank smooths its rows with ITS OWN eigenvalue estimate
        self.lmax = np.zeros(self.A.shape[0])
        for p, h in zip(parts, self.loc):
            self.lmax[p] = h.levels[0].lmax
        self.deg, self.ratio = self.loc[0].deg, self.loc[0].ratio
        self.single_level = all(len(h.levels) == 1 for h in self.loc)

    def _cheby(self, b, x, zero):
        lmax = self.lmax
        lmin = lmax / self.ratio
        th, de = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sg = th / de
        rho = 1.0 / sg
        r = b.copy() if zero else b - self.A @ x
        d = self.dinv * r / th
        for k in range(self.deg):
            x = x + d
            if k == self.deg - 1:
                break
            r = r - self.A @ d
            rn = 1.0 / (2.0 * sg - rho)
            d = rn * rho * d + (2.0 * rn / de) * (self.dinv * r)
            rho = rn
        return x

    def __call__(self, b):
        x = self._cheby(b, np.zeros_like(b), True)
        if self.single_level:                      # `chebyshev` PC or blocks below the coarse size
            return x
        r = b - self.A @ x
        for p, h in zip(self.parts, self.loc):
            if len(h.levels) > 1:
                L0 = h.levels[0]
                x[p] += L0.P @ h._cycle(1, L0.R @ r[p])
        return self._cheby(b, x, False)


class LocalSchurAmg:
    """selfp Schur complement S = A11 - A10 diag(A00)^-1 A01 assembled from OWNED parts only, rank-local AMG."""

    def __init__(self, A00, A01, A10, A11, parts0, parts1, bs=1, B=None, **kw):
        dinv = 1.0 / A00.diagonal()
        self.parts = parts1
        self.loc = []
        for p0, p1 in zip(parts0, parts1):
            S = (A11[p1][:, p1] - A10[p1][:, p0] @ sp.diags(dinv[p0]) @ A01[p0][:, p1]).tocsr()
            self.loc.append(SAAMG(S, bs, None if B is None else B[p1], **kw))

    def __call__(self, b):
        x = np.zeros_like(b)
        for p1, h in zip(self.parts, self.loc):
            x[p1] = h(b[p1])
        return x
