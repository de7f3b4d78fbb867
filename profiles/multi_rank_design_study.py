"""Round-2 design study on the CPU twin (oracle/ddamg.py reproduces the GPU multi-rank iteration counts exactly):
which change to the row-partitioned preconditioner buys back the single-GPU iteration count?

    python profiles/multi_rank_design_study.py N R       (benchmark options: s theta 0.04, f Chebyshev(4), p V-cycle)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from oracle.amg import SAAMG, rigid_body_modes
from oracle.blockpc import BlockPC
from oracle.ddamg import DDAmg, LocalSchurAmg
from oracle.krylov import gmres
from oracle.problems import swelling
from poro_b200.partition import slab_ranges

N, R = int(sys.argv[1]), int(sys.argv[2])
s, _ = swelling(3, N, "diagonal")
h2 = 1e-2 / (2 * N)
plane_s = np.rint(s.coords_s[:, 2] / h2).astype(int)
plane_p = np.rint(s.coords_p[:, 2] / h2).astype(int)
ranges = slab_ranges(2 * N + 1, R)
ps = [np.flatnonzero((plane_s >= a) & (plane_s < b)) for a, b in ranges]
pp = [np.flatnonzero((plane_p >= a) & (plane_p < b)) for a, b in ranges]
Bg = rigid_body_modes(s.coords_s, 3)
Bl = np.zeros((s.ns, 6))
for p in ps:
    Bl[p] = rigid_body_modes(s.coords_s[p], 3)
f, p_ = np.arange(s.nf), s.nf + np.arange(s.np_)
KW_S = dict(theta=0.04)


class RASAmg(DDAmg):
    """hierarchies on [owned | overlap] blocks, correction restricted to the owned rows."""
    def __init__(self, A, bs, B, parts, parts_ext, **kw):
        self.A = sp.csr_matrix(A); self.parts = parts; self.ext = parts_ext
        self.loc = [SAAMG(self.A[pe][:, pe], bs, None if B is None else B[pe], **kw) for pe in parts_ext]
        d = self.A.diagonal(); self.dinv = 1.0 / np.where(d != 0, d, 1.0)
        self.lmax = np.zeros(self.A.shape[0])
        for pe, pown, h in zip(parts_ext, parts, self.loc):
            self.lmax[pown] = h.levels[0].lmax
        self.deg, self.ratio = self.loc[0].deg, self.loc[0].ratio
        self.single_level = False
        self.own = [np.isin(pe, pown) for pe, pown in zip(parts_ext, parts)]
    def __call__(self, b):
        x = self._cheby(b, np.zeros_like(b), True)
        r = b - self.A @ x
        for pe, m, h in zip(self.ext, self.own, self.loc):
            L0 = h.levels[0]
            corr = L0.P @ h._cycle(1, L0.R @ r[pe])
            x[pe[m]] += corr[m]
        return self._cheby(b, x, False)


def ext_parts(planes, layers):
    return [np.flatnonzero((planes >= a - layers) & (planes < b + layers)) for a, b in ranges]


def solve(name, mk_s, mk_f, mk_p_factory):
    class Schur:
        def __init__(self, M):
            self.A00, self.A01, self.A10, self.A11 = M[f][:, f], M[f][:, p_], M[p_][:, f], M[p_][:, p_]
            self.k0 = mk_f(self.A00)
            self.k1 = mk_p_factory(self)
        def __call__(self, x):
            y0 = self.k0(x[: s.nf])
            return np.concatenate([y0, self.k1(x[s.nf:] - self.A10 @ y0)])
    t = time.time()
    pc = BlockPC(s, {"s": mk_s, "fp": lambda M: Schur(M)})
    r = gmres(lambda v: s.A @ v, s.b, pc, rtol=1e-8, atol=0.0, dtol=1e20, max_it=600, restart=600, pc_side="right")
    print("N %d R %d  %-58s its %4d  (%.0f s)" % (N, R, name, r.its, time.time() - t), flush=True)


cheb_f = lambda M: DDAmg(M, 3, Bl, ps, max_levels=1, cheby_degree=4, dense_limit=0)       # global Chebyshev(4): what the GPU does
glob_S = lambda S: (S.A11 - S.A10 @ sp.diags(1.0 / S.A00.diagonal()) @ S.A01).tocsr()
p_local = lambda S: LocalSchurAmg(S.A00, S.A01, S.A10, S.A11, ps, pp)
p_globalS_localAmg = lambda S: DDAmg(glob_S(S), 1, None, pp)
p_single = lambda S: SAAMG(glob_S(S), 1, None)
s_local = lambda M: DDAmg(M, 3, Bl, ps, **KW_S)
s_single = lambda M: SAAMG(M, 3, Bg, **KW_S)
s_ras = lambda layers: (lambda M: RASAmg(M, 3, Bg, ps, ext_parts(plane_s, layers), **KW_S))

# "uncoupled" aggregation: aggregates never cross a rank boundary (each rank aggregates its own nodes with the
# existing kernels), prolongator smoothing and the Galerkin product use the distributed operator
lab_s_dof = np.zeros(s.ns, int); lab_p = np.zeros(s.np_, int)
for r_, (p0, p1) in enumerate(zip(ps, pp)):
    lab_s_dof[p0] = r_; lab_p[p1] = r_
s_uncoupled = lambda B: (lambda M: SAAMG(M, 3, B, node_labels=lab_s_dof[::3], **KW_S))
p_uncoupled = lambda S: SAAMG(glob_S(S), 1, None, node_labels=lab_p)

solve("single hierarchy everywhere (1 GPU)", s_single, cheb_f, p_single)
solve("uncoupled aggregation + distributed Galerkin (s, S_p)", s_uncoupled(Bg), cheb_f, p_uncoupled)
solve("  same, rigid-body modes about each rank's own centre", s_uncoupled(Bl), cheb_f, p_uncoupled)
solve("round 1: s local coarse, S_p from owned parts + local AMG", s_local, cheb_f, p_local)
solve("s local coarse, S_p halo-aware + global level-0 smoothing", s_local, cheb_f, p_globalS_localAmg)
solve("s single hierarchy, S_p from owned parts + local AMG", s_single, cheb_f, p_local)
solve("s RAS overlap 2 planes, S_p round 1", s_ras(2), cheb_f, p_local)
solve("s RAS overlap 2 planes, S_p halo-aware + global smoothing", s_ras(2), cheb_f, p_globalS_localAmg)
solve("s RAS overlap 4 planes, S_p halo-aware + global smoothing", s_ras(4), cheb_f, p_globalS_localAmg)
solve("s single hierarchy, S_p halo-aware + global smoothing", s_single, cheb_f, p_globalS_localAmg)
