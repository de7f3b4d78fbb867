"""Round-2 design study on the CPU oracle: what does an fp32 preconditioner cost in outer iterations?

The V-cycles and the Chebyshev sweeps are HBM-bound on the matrix streams (12 B per nonzero in fp64 CSR, 76 B per
3x3 block in BSR); storing the HIERARCHY (level operators, P, R, D^-1) and its work vectors in fp32 would cut that to
8 B / 40 B.  The outer Krylov method stays fp64 (operator, basis, Hessenberg).  An fp32 preconditioner is a slightly
nonlinear operator, so the comparison uses right-preconditioned GMRES (what bench.py runs) and FGMRES.

    python profiles/mixed_precision_study.py N
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from oracle.amg import SAAMG, rigid_body_modes
from oracle.blockpc import BlockPC, SchurLower, krylov_solver
from oracle.krylov import gmres
from oracle.problems import swelling


class SingleAmg(SAAMG):
    """Same hierarchy (built in fp64), stored and cycled in fp32; input and output vectors fp64.
    The right-hand side is scaled to unit max-norm first so that fp32's RANGE is never the issue."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for L in self.levels:
            L.A = L.A.astype(np.float32)
            L.dinv = L.dinv.astype(np.float32)
            L.lmax = float(L.lmax)                      # a Python scalar does not promote the fp32 vectors
            if hasattr(L, "P"):
                L.P, L.R = L.P.astype(np.float32), L.R.astype(np.float32)
        if self.coarse_direct:
            self.levels[-1].inv = self.levels[-1].inv.astype(np.float32)

    def __call__(self, b):
        s = np.abs(b).max()
        if s == 0:
            return np.zeros_like(b)
        y = self._cycle(0, (b / s).astype(np.float32))
        assert y.dtype == np.float32
        return y.astype(np.float64) * s


def run(N):
    sys_, _ = swelling(3, N, "diagonal")
    B = rigid_body_modes(sys_.coords_s, 3)

    def solve(name, amg_cls, which, flexible):
        cls = lambda blk: amg_cls if blk in which else SAAMG
        mk_s = lambda M: cls("s")(M, 3, B, theta=0.04)
        mk_f = lambda M: cls("f")(M, 3, B, max_levels=1, cheby_degree=4)
        mk_p = lambda M: cls("p")(M, 1, None)
        mkfp = lambda M: SchurLower(M, sys_.nf, sys_.np_, krylov_solver("preonly", mk_f), krylov_solver("preonly", mk_p), "f")
        pc = BlockPC(sys_, {"s": krylov_solver("preonly", mk_s), "fp": mkfp})
        t = time.time()
        r = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=1e-8, atol=0.0, dtol=1e20, max_it=200, restart=200, pc_side="right",
                  flexible=flexible)
        true = np.linalg.norm(sys_.b - sys_.A @ r.x) / np.linalg.norm(sys_.b)
        print("N %2d  %-58s its %3d  estimate %.2e  true residual %.2e  (%.0f s)" % (N, name, r.its, r.rnorm / r.history[0], true,
                                                                                    time.time() - t), flush=True)

    solve("fp64 preconditioner, GMRES(right)   [bench.py]", SAAMG, "", False)
    solve("fp32 hierarchies (s, f, p), GMRES(right)", SingleAmg, "sfp", False)
    solve("fp32 hierarchies (s, f, p), FGMRES", SingleAmg, "sfp", True)
    solve("fp32 s only, FGMRES", SingleAmg, "s", True)
    solve("fp32 s and f, fp64 pressure Schur, FGMRES", SingleAmg, "sf", True)


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 12)
