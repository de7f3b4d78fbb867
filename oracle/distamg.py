"""Rank-level restatement of a DISTRIBUTED smoothed-aggregation set-up and V-cycle (TEST INFRASTRUCTURE).

Round 1's multi-GPU preconditioner keeps every coarse level rank-local (oracle/ddamg.py is its twin) and loses
the single-GPU iteration count.  profiles/r1_multi_rank_design_study.md shows that "uncoupled" aggregation (each
rank aggregates only its own nodes) with a distributed prolongator smoothing and Galerkin product recovers it.
This module is that algorithm written the way the ranks will execute it, so that the CUDA + NCCL version has
something to be checked against line by line:

  * every rank holds its rows of the level operator as CSR with columns [owned | ghost] (the layout of
    `poro_mat_create_csr` + `poro_halo_set`), ghost columns sorted by global id = neighbour-major;
  * the only communication primitives are
        halo_vec   -- fixed-width rows of the owned boundary nodes to the neighbours (vectors, T rows),
        halo_rows  -- variable-length sparse rows of the owned boundary nodes (rows of P and of A*P),
        allreduce  -- scalars (eigenvalue estimate, level sizes), and one allgather on the coarsest level;
  * everything else is rank-local work with kernels that exist already (strength graph, MIS(2), tentative
    prolongator, SpGEMM, transpose).

The ranks are emulated in one process: the per-rank state lives in Python lists and the primitives move data
between list entries while counting messages and bytes (`Traffic`).  tests/test_oracle_distamg.py checks the
assembled pieces against the global formulas P = (I - w D^-1 A) T, A_c = P^T A P, the cycle against a
single-process V-cycle on the same hierarchy, and the outer iteration count against the single-GPU solver.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from .amg import aggregate_mis2, hash32, strength_graph, tentative_prolongator


@dataclass
class Traffic:
    messages: int = 0
    bytes: int = 0
    allreduces: int = 0

    def add(self, nbytes):
        if nbytes:
            self.messages += 1
            self.bytes += int(nbytes)


@dataclass
class Plan:
    """Halo plan of one rank on one level (same content as partition.HaloPlan / poro_halo_set)."""
    offset: int                 # global id of the first owned row
    n_owned: int
    ghost_gid: np.ndarray       # global ids of the ghost columns, ascending (= neighbour-major, owner's local order)
    neigh: list = field(default_factory=list)        # neighbour ranks, ascending
    send_idx: dict = field(default_factory=dict)     # neighbour -> owned local indices it reads
    recv_slice: dict = field(default_factory=dict)   # neighbour -> slice of the ghost block it fills


def make_plans(offsets: np.ndarray, ghost_gids: list, traffic: Traffic | None = None) -> list:
    """The one handshake of a level: every rank tells the owners which rows it reads."""
    R = len(ghost_gids)
    plans = []
    for r in range(R):
        g = np.asarray(ghost_gids[r], np.int64)
        assert np.all(np.diff(g) > 0)
        owner = np.searchsorted(offsets, g, side="right") - 1
        assert not np.any(owner == r)
        p = Plan(int(offsets[r]), int(offsets[r + 1] - offsets[r]), g)
        for q in np.unique(owner):
            idx = np.flatnonzero(owner == q)
            p.recv_slice[int(q)] = slice(int(idx[0]), int(idx[-1]) + 1)
        plans.append(p)
    for r in range(R):
        for q, sl in plans[r].recv_slice.items():
            plans[q].send_idx[r] = plans[r].ghost_gid[sl] - offsets[q]
            if traffic is not None:
                traffic.add(8 * (sl.stop - sl.start))
    for p in plans:
        p.neigh = sorted(set(p.send_idx) | set(p.recv_slice))
    return plans


def halo_vec(plans: list, xs: list, traffic: Traffic | None = None) -> list:
    """xs[r]: (n_owned,) or (n_owned, w).  Returns the extended arrays [owned | ghost]."""
    out = []
    for r, p in enumerate(plans):
        x = xs[r]
        ghost = np.zeros((len(p.ghost_gid),) + x.shape[1:], x.dtype)
        for q, sl in p.recv_slice.items():
            msg = xs[q][plans[q].send_idx[r]]            # packed by the owner, sent, unpacked in place
            ghost[sl] = msg
            if traffic is not None:
                traffic.add(msg.nbytes)
        out.append(np.concatenate([x, ghost]))
    return out


def halo_rows(plans: list, mats: list, traffic: Traffic | None = None) -> list:
    """mats[r]: CSR of the owned rows with GLOBAL column ids.  Returns the CSR of the ghost rows of every rank
    (row i = ghost i of the plan).  Two messages per neighbour: row lengths, then packed (column, value) pairs."""
    out = []
    for r, p in enumerate(plans):
        blocks = []
        for q in sorted(p.recv_slice):
            rows = mats[q][plans[q].send_idx[r]]
            blocks.append(rows)
            if traffic is not None:
                traffic.add(4 * rows.shape[0])
                traffic.add(16 * rows.nnz)
        ncols = mats[r].shape[1]
        out.append(sp.vstack(blocks).tocsr() if blocks else sp.csr_matrix((0, ncols)))
    return out


@dataclass
class RankLevel:
    A: sp.csr_matrix            # owned rows x [owned | ghost]
    plan: Plan
    dinv: np.ndarray
    lmax: float = 1.0
    P: sp.csr_matrix | None = None      # owned fine rows x [owned coarse | ghost coarse]
    R: sp.csr_matrix | None = None      # owned coarse rows x [owned fine | ghost fine]
    P_glob: sp.csr_matrix | None = None   # the same P with global coarse columns (kept for the tests)
    T_glob: sp.csr_matrix | None = None


def distribute(A: sp.csr_matrix, offsets: np.ndarray):
    """Rows of a matrix in rank-contiguous numbering -> per-rank CSR with [owned | ghost] columns."""
    A = sp.csr_matrix(A)
    mats, ghosts = [], []
    for r in range(len(offsets) - 1):
        mats.append(A[offsets[r]: offsets[r + 1]].tocsr())
    return _localize(mats, offsets)


def _localize(mats_glob: list, offsets: np.ndarray, extra_cols: list | None = None):
    """Global column ids -> [owned | ghost]; `extra_cols[r]` are further global ids that must be ghosts."""
    loc, ghosts = [], []
    for r, M in enumerate(mats_glob):
        a, b = int(offsets[r]), int(offsets[r + 1])
        cols = M.indices
        off = (cols < a) | (cols >= b)
        need = cols[off]
        if extra_cols is not None:
            e = np.asarray(extra_cols[r], np.int64)
            need = np.concatenate([need, e[(e < a) | (e >= b)]])
        g = np.unique(need).astype(np.int64)
        new = np.where(off, (b - a) + np.searchsorted(g, cols), cols - a)
        L = sp.csr_matrix((M.data, new, M.indptr), shape=(M.shape[0], (b - a) + len(g)))
        L.sort_indices()
        loc.append(L)
        ghosts.append(g)
    return loc, ghosts


def _to_global_cols(M: sp.csr_matrix, plan: Plan, ncols_glob: int) -> sp.csr_matrix:
    gid = np.concatenate([plan.offset + np.arange(plan.n_owned, dtype=np.int64), plan.ghost_gid])
    return sp.csr_matrix((M.data, gid[M.indices], M.indptr), shape=(M.shape[0], ncols_glob))


class DistAmg:
    """One hierarchy over R emulated ranks.  `A` in rank-contiguous numbering, `offsets` (R+1,)."""

    def __init__(self, A, offsets, bs=1, B=None, theta=0.08, max_levels=10, coarse_size=400, cheby_degree=2,
                 cheby_ratio=10.0, power_its=15, dense_limit=4096, replicate_below=0):
        """replicate_below: once a level has at most that many rows it is gathered on every rank and the rest of
        the hierarchy is built and cycled redundantly (one allgather per cycle instead of latency-bound halo
        exchanges between ever more neighbours: coarse operators couple every rank with every other)."""
        A = sp.csr_matrix(A)
        self.tail = None
        offsets = np.asarray(offsets, np.int64)
        R = len(offsets) - 1
        self.R, self.deg, self.ratio = R, cheby_degree, cheby_ratio
        self.setup_traffic, self.cycle_traffic = Traffic(), Traffic()
        n = A.shape[0]
        if B is None:
            B = np.zeros((n, bs))
            for c in range(bs):
                B[c::bs, c] = 1.0
        Bs = [B[offsets[r]: offsets[r + 1]].copy() for r in range(R)]
        mats, ghosts = distribute(A, offsets)
        self.levels = []            # levels[l][r] : RankLevel
        self.offsets = []
        tr = self.setup_traffic
        while True:
            plans = make_plans(offsets, ghosts, tr)
            lev = []
            for r in range(R):
                d = mats[r].diagonal()                      # owned columns come first: the diagonal is local
                lev.append(RankLevel(mats[r], plans[r], 1.0 / np.where(d != 0, d, 1.0)))
            lmax = 1.1 * self._power_lmax(lev, offsets, power_its)
            for L in lev:
                L.lmax = lmax
            self.levels.append(lev)
            self.offsets.append(offsets)
            n = int(offsets[-1])
            if n <= coarse_size or len(self.levels) >= max_levels:
                break
            if n <= replicate_below and len(self.levels) > 1:
                from .amg import SAAMG
                Aglob = self.global_matrix(len(self.levels) - 1)       # allgather of the level's rows
                tr.add(12 * Aglob.nnz)
                self.tail = SAAMG(Aglob, bs, np.concatenate(Bs), theta=theta, max_levels=max_levels - len(self.levels) + 1,
                                  coarse_size=coarse_size, cheby_degree=cheby_degree, cheby_ratio=cheby_ratio,
                                  power_its=power_its, dense_limit=dense_limit)
                break
            # ---- rank-local: aggregation on the owned block, tentative prolongator -------------------------
            Ts, Bcs, ncs = [], [], []
            k = Bs[0].shape[1]
            for r, L in enumerate(lev):
                no = L.plan.n_owned
                d = L.A.diagonal()
                rowabs = np.asarray(abs(L.A).sum(1)).ravel()
                Bs[r][(rowabs - np.abs(d)) <= 1e-14 * np.abs(d)] = 0.0       # Dirichlet rows: full row, ghosts included
                sq = L.A[:, :no].tocsr()
                agg, n_agg = aggregate_mis2(strength_graph(sq, bs, theta))
                if n_agg > 0:
                    T, Bc = tentative_prolongator(agg, n_agg, bs, Bs[r])
                else:
                    T, Bc = sp.csr_matrix((no, 0)), np.zeros((0, k))
                Ts.append(T.tocsr()); Bcs.append(Bc); ncs.append(n_agg * k)
            tr.allreduces += 1                                               # level sizes
            if sum(ncs) == 0 or sum(ncs) >= 0.8 * n:
                break
            coff = np.concatenate([[0], np.cumsum(ncs)]).astype(np.int64)
            NC = int(coff[-1])
            # ---- P = T - w D^-1 A T : needs the T rows of the ghost nodes (fixed width: bs x k values + one id) ----
            Tg = [sp.csr_matrix((T.data, T.indices + coff[r], T.indptr), shape=(T.shape[0], NC)) for r, T in enumerate(Ts)]
            Tghost = halo_rows(plans, Tg, tr)
            omega = 4.0 / (3.0 * lmax / 1.1)
            Pg = []
            for r, L in enumerate(lev):
                Text = sp.vstack([Tg[r], Tghost[r]]).tocsr()
                Pg.append((Tg[r] - sp.diags(omega * L.dinv) @ (L.A @ Text)).tocsr())
            # ---- A_c = P^T A P : P rows of the ghosts, then A*P rows of the ghosts -----------------------
            Pghost = halo_rows(plans, Pg, tr)
            Pext = [sp.vstack([Pg[r], Pghost[r]]).tocsr() for r in range(R)]
            APg = [(lev[r].A @ Pext[r]).tocsr() for r in range(R)]
            APghost = halo_rows(plans, APg, tr)
            Acg, Rloc = [], []
            for r in range(R):
                mine = Pext[r][:, coff[r]: coff[r + 1]].T.tocsr()            # owned coarse rows x [owned | ghost] fine
                Ac = (mine @ sp.vstack([APg[r], APghost[r]]).tocsr()).tocsr()
                dc = Ac[:, coff[r]: coff[r + 1]].diagonal()
                dead = dc == 0
                if dead.any():
                    fix = sp.csr_matrix((dead.astype(float), (np.arange(len(dc)), coff[r] + np.arange(len(dc)))),
                                        shape=Ac.shape)
                    Ac = (Ac + fix).tocsr()
                Acg.append(Ac); Rloc.append(mine)
            # ---- next level: ghost coarse columns = those of A_c and of P ----------------------------------
            extra = [Pg[r].indices for r in range(R)]
            mats, ghosts = _localize(Acg, coff, extra)
            for r, L in enumerate(lev):
                g = ghosts[r]
                a, b = int(coff[r]), int(coff[r + 1])
                cols = Pg[r].indices
                off = (cols < a) | (cols >= b)
                new = np.where(off, (b - a) + np.searchsorted(g, cols), cols - a)
                L.P = sp.csr_matrix((Pg[r].data, new, Pg[r].indptr), shape=(Pg[r].shape[0], (b - a) + len(g)))
                L.R = Rloc[r]
                L.P_glob, L.T_glob = Pg[r], Tg[r]
            Bs, bs, offsets = Bcs, k, coff
        # coarsest level: gathered on every rank and inverted densely when small
        n = int(self.offsets[-1][-1])
        self.coarse_direct = n <= dense_limit and self.tail is None
        if self.coarse_direct:
            self.coarse_inv = np.linalg.inv(self.global_matrix(len(self.levels) - 1).toarray())
            tr.add(0)

    # ---- pieces -------------------------------------------------------------------------------------------
    def _matvec(self, lev, xs, traffic):
        ext = halo_vec([L.plan for L in lev], xs, traffic)
        return [L.A @ e for L, e in zip(lev, ext)]

    def _power_lmax(self, lev, offsets, its):
        tr = self.setup_traffic
        vs = []
        for r, L in enumerate(lev):
            gid = L.plan.offset + np.arange(L.plan.n_owned)
            vs.append((hash32(gid) % 2048).astype(float) / 1024.0 - 1.0)
        nrm = np.sqrt(sum(float(v @ v) for v in vs)); tr.allreduces += 1
        vs = [v / nrm for v in vs]
        lam = 1.0
        for _ in range(its):
            ws = [L.dinv * w for L, w in zip(lev, self._matvec(lev, vs, tr))]
            lam = np.sqrt(sum(float(w @ w) for w in ws)); tr.allreduces += 1
            if lam == 0:
                return 1.0
            vs = [w / lam for w in ws]
        return lam

    def global_matrix(self, l) -> sp.csr_matrix:
        n = int(self.offsets[l][-1])
        return sp.vstack([_to_global_cols(L.A, L.plan, n) for L in self.levels[l]]).tocsr()

    def global_P(self, l) -> sp.csr_matrix:
        return sp.vstack([L.P_glob for L in self.levels[l]]).tocsr()

    def global_T(self, l) -> sp.csr_matrix:
        return sp.vstack([L.T_glob for L in self.levels[l]]).tocsr()

    def complexity(self):
        nnz = [sum(L.A.nnz for L in lev) for lev in self.levels]
        return sum(nnz) / nnz[0]

    # ---- cycle --------------------------------------------------------------------------------------------
    def _cheby(self, lev, bs_, xs, zero):
        tr = self.cycle_traffic
        lmax = lev[0].lmax
        lmin = lmax / self.ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho = 1.0 / sigma
        if zero:
            rs = [b.copy() for b in bs_]
        else:
            rs = [b - w for b, w in zip(bs_, self._matvec(lev, xs, tr))]
        ds = [L.dinv * r / theta for L, r in zip(lev, rs)]
        for k in range(self.deg):
            xs = [x + d for x, d in zip(xs, ds)]
            if k == self.deg - 1:
                break
            rs = [r - w for r, w in zip(rs, self._matvec(lev, ds, tr))]
            rho_new = 1.0 / (2.0 * sigma - rho)
            ds = [rho_new * rho * d + (2.0 * rho_new / delta) * (L.dinv * r) for L, d, r in zip(lev, ds, rs)]
            rho = rho_new
        return xs

    def _cycle(self, l, bs_):
        lev = self.levels[l]
        tr = self.cycle_traffic
        zeros = [np.zeros_like(b) for b in bs_]
        if l == len(self.levels) - 1:
            if self.tail is not None:
                full = np.concatenate(bs_)                       # allgather, then the redundant sub-cycle
                tr.add(full.nbytes)
                x = self.tail(full)
                off = self.offsets[l]
                return [x[off[r]: off[r + 1]] for r in range(self.R)]
            if self.coarse_direct:
                full = np.concatenate(bs_)                       # allgather of the coarsest right-hand side
                tr.add(full.nbytes)
                off = self.offsets[l]
                return [self.coarse_inv[off[r]: off[r + 1]] @ full for r in range(self.R)]
            xs = self._cheby(lev, bs_, zeros, True)
            return xs if len(self.levels) == 1 else self._cheby(lev, bs_, xs, False)
        xs = self._cheby(lev, bs_, zeros, True)
        rs = [b - w for b, w in zip(bs_, self._matvec(lev, xs, tr))]
        rext = halo_vec([L.plan for L in lev], rs, tr)           # restriction reads the ghost residuals
        bc = [L.R @ e for L, e in zip(lev, rext)]
        xc = self._cycle(l + 1, bc)
        xcext = halo_vec([L.plan for L in self.levels[l + 1]], xc, tr)   # prolongation reads the ghost coarse values
        xs = [x + L.P @ e for x, L, e in zip(xs, lev, xcext)]
        return self._cheby(lev, bs_, xs, False)

    def __call__(self, b):
        off = self.offsets[0]
        xs = self._cycle(0, [b[off[r]: off[r + 1]] for r in range(self.R)])
        return np.concatenate(xs)


class GlobalCycle:
    """Single-process V-cycle on the hierarchy a DistAmg assembled (global matrices) -- the check for the
    distributed cycle."""

    def __init__(self, h: DistAmg):
        from .amg import SAAMG, Level
        self.h = SAAMG.__new__(SAAMG)
        self.h.deg, self.h.ratio = h.deg, h.ratio
        self.h.levels = []
        for l, lev in enumerate(h.levels):
            L = Level()
            L.A = h.global_matrix(l)
            L.dinv = np.concatenate([x.dinv for x in lev])
            L.lmax = lev[0].lmax
            if lev[0].P_glob is not None:
                L.P = h.global_P(l)
                L.R = L.P.T.tocsr()
            self.h.levels.append(L)
        self.h.coarse_direct = h.coarse_direct
        if h.coarse_direct:
            self.h.levels[-1].inv = h.coarse_inv

    def __call__(self, b):
        return self.h(b)
