"""One RANK of the distributed SA-AMG of oracle/distamg.py, over a real communicator (TEST INFRASTRUCTURE).

oracle/distamg.py emulates all ranks in one process and could, by mistake, read another rank's data outside
an exchange.  Here every rank is its own process and owns only its rows; everything else arrives through
`Comm` (torch.distributed, gloo on CPU): the protocol below is the one the CUDA + NCCL implementation has to
follow (grouped send/recv per neighbour, all-reduce of scalars, one all-gather of sizes per level).
tests/test_oracle_distamg_gloo.py runs it with world sizes 2 and 3 and compares the hierarchy and the cycle with
the single-process emulation.

Message protocol per level (n = neighbours of this rank on the level):
    handshake      all-gather of the ghost id lists' owners' view (ids each rank needs, grouped by owner)
    halo_vec       n messages of 8 * width * |send list| bytes                    (smoothing, residual, transfer)
    halo_rows      2n messages: int32 row lengths, then packed (int64 column, double value) pairs
    allreduce      scalars (norms of the power iteration, level sizes)
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .amg import aggregate_mis2, hash32, strength_graph, tentative_prolongator


class Comm:
    """The few collectives the algorithm needs, over torch.distributed."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        self.messages = 0
        self.bytes = 0

    def allreduce_sum(self, x: float) -> float:
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return float(t[0])

    def allgather_obj(self, obj):
        out = [None] * self.size
        self.dist.all_gather_object(out, obj)
        return out

    def exchange(self, sends: dict, recv_shapes: dict, dtype):
        """sends: neighbour -> numpy array; recv_shapes: neighbour -> shape.  One grouped send/recv."""
        torch, dist = self.torch, self.dist
        bufs, reqs = {}, []
        for q, shape in recv_shapes.items():
            bufs[q] = torch.zeros(shape, dtype=dtype)
            if bufs[q].numel():
                reqs.append(dist.irecv(bufs[q], q))
        keep = []
        for q, a in sends.items():
            t = torch.from_numpy(np.ascontiguousarray(a))
            keep.append(t)
            if t.numel():
                reqs.append(dist.isend(t, q))
                self.messages += 1
                self.bytes += t.numel() * t.element_size()
        for r in reqs:
            r.wait()
        return {q: b.numpy() for q, b in bufs.items()}


class Plan:
    def __init__(self, comm: Comm, offsets: np.ndarray, ghost_gid: np.ndarray):
        """Handshake: tell every owner which of its rows this rank reads."""
        r = comm.rank
        self.offset, self.n_owned = int(offsets[r]), int(offsets[r + 1] - offsets[r])
        self.ghost_gid = np.asarray(ghost_gid, np.int64)
        owner = np.searchsorted(offsets, self.ghost_gid, side="right") - 1
        self.recv = {}                                   # neighbour -> slice of the ghost block
        need = {}
        for q in np.unique(owner):
            idx = np.flatnonzero(owner == q)
            self.recv[int(q)] = slice(int(idx[0]), int(idx[-1]) + 1)
            need[int(q)] = self.ghost_gid[idx]
        everyone = comm.allgather_obj(need)              # everyone[p][q] = ids rank p reads from rank q
        self.send = {p: everyone[p][r] - self.offset for p in range(comm.size) if r in everyone[p]}


def halo_vec(comm: Comm, plan: Plan, x: np.ndarray) -> np.ndarray:
    torch = comm.torch
    width = x.shape[1:] if x.ndim > 1 else ()
    got = comm.exchange({q: x[idx] for q, idx in plan.send.items()},
                        {q: (sl.stop - sl.start,) + width for q, sl in plan.recv.items()}, torch.float64)
    ghost = np.zeros((len(plan.ghost_gid),) + width)
    for q, sl in plan.recv.items():
        ghost[sl] = got[q]
    return np.concatenate([x, ghost])


def halo_rows(comm: Comm, plan: Plan, M: sp.csr_matrix) -> sp.csr_matrix:
    """M: owned rows, GLOBAL column ids.  Returns the rows of the ghosts (same column space)."""
    torch = comm.torch
    M = sp.csr_matrix(M)
    M.sort_indices()
    packs = {q: M[idx] for q, idx in plan.send.items()}
    lens = comm.exchange({q: np.diff(P.indptr).astype(np.int32) for q, P in packs.items()},
                         {q: (sl.stop - sl.start,) for q, sl in plan.recv.items()}, torch.int32)
    cols = comm.exchange({q: P.indices.astype(np.int64) for q, P in packs.items()},
                         {q: (int(lens[q].sum()),) for q in plan.recv}, torch.int64)
    vals = comm.exchange({q: P.data for q, P in packs.items()},
                         {q: (int(lens[q].sum()),) for q in plan.recv}, torch.float64)
    blocks = []
    for q in sorted(plan.recv):
        ip = np.concatenate([[0], np.cumsum(lens[q])])
        blocks.append(sp.csr_matrix((vals[q], cols[q], ip), shape=(len(lens[q]), M.shape[1])))
    return sp.vstack(blocks).tocsr() if blocks else sp.csr_matrix((0, M.shape[1]))


def localize(M: sp.csr_matrix, a: int, b: int, extra=None):
    """Global column ids -> [owned | ghost] of the range [a, b); returns (local CSR, ghost ids)."""
    cols = M.indices
    off = (cols < a) | (cols >= b)
    need = cols[off]
    if extra is not None:
        e = np.asarray(extra, np.int64)
        need = np.concatenate([need, e[(e < a) | (e >= b)]])
    g = np.unique(need).astype(np.int64)
    new = np.where(off, (b - a) + np.searchsorted(g, cols), cols - a)
    L = sp.csr_matrix((M.data, new, M.indptr), shape=(M.shape[0], (b - a) + len(g)))
    L.sort_indices()
    return L, g


class Level:
    pass


class RankAmg:
    """This rank's part of the hierarchy.  `A_rows`: owned rows with GLOBAL columns (rank-contiguous numbering)."""

    def __init__(self, comm: Comm, A_rows, offsets, bs=1, B=None, theta=0.08, max_levels=10, coarse_size=400,
                 cheby_degree=2, cheby_ratio=10.0, power_its=15, dense_limit=4096):
        self.comm, self.deg, self.ratio = comm, cheby_degree, float(cheby_ratio)
        r = comm.rank
        offsets = np.asarray(offsets, np.int64)
        A_rows = sp.csr_matrix(A_rows)
        n_own = A_rows.shape[0]
        if B is None:
            B = np.zeros((n_own, bs))
            for c in range(bs):
                B[c::bs, c] = 1.0
        B = B.copy()
        A, ghosts = localize(A_rows, int(offsets[r]), int(offsets[r + 1]))
        self.levels = []
        while True:
            L = Level()
            L.A, L.offsets = A, offsets
            L.plan = Plan(comm, offsets, ghosts)
            d = A.diagonal()
            L.dinv = 1.0 / np.where(d != 0, d, 1.0)
            L.lmax = 1.1 * self._power_lmax(L, power_its)
            self.levels.append(L)
            n = int(offsets[-1])
            if n <= coarse_size or len(self.levels) >= max_levels:
                break
            no = L.plan.n_owned
            rowabs = np.asarray(abs(A).sum(1)).ravel()
            B[(rowabs - np.abs(d)) <= 1e-14 * np.abs(d)] = 0.0
            k = B.shape[1]
            agg, n_agg = aggregate_mis2(strength_graph(A[:, :no].tocsr(), bs, theta))
            if n_agg > 0:
                T, Bc = tentative_prolongator(agg, n_agg, bs, B)
            else:
                T, Bc = sp.csr_matrix((no, 0)), np.zeros((0, k))
            sizes = comm.allgather_obj(n_agg * k)                               # one all-gather of sizes
            if sum(sizes) == 0 or sum(sizes) >= 0.8 * n:
                break
            coff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            NC = int(coff[-1])
            T = T.tocsr()
            Tg = sp.csr_matrix((T.data, T.indices + coff[r], T.indptr), shape=(no, NC))
            Text = sp.vstack([Tg, halo_rows(comm, L.plan, Tg)]).tocsr()
            omega = 4.0 / (3.0 * L.lmax / 1.1)
            Pg = (Tg - sp.diags(omega * L.dinv) @ (A @ Text)).tocsr()
            Pext = sp.vstack([Pg, halo_rows(comm, L.plan, Pg)]).tocsr()
            APg = (A @ Pext).tocsr()
            APext = sp.vstack([APg, halo_rows(comm, L.plan, APg)]).tocsr()
            mine = Pext[:, coff[r]: coff[r + 1]].T.tocsr()
            Ac = (mine @ APext).tocsr()
            dc = Ac[:, coff[r]: coff[r + 1]].diagonal()
            dead = dc == 0
            if dead.any():
                Ac = (Ac + sp.csr_matrix((dead.astype(float), (np.arange(len(dc)), coff[r] + np.arange(len(dc)))),
                                         shape=Ac.shape)).tocsr()
            A, ghosts = localize(Ac, int(coff[r]), int(coff[r + 1]), Pg.indices)
            L.P, _ = localize(Pg, int(coff[r]), int(coff[r + 1]), ghosts)
            assert L.P.shape[1] == A.shape[1]
            L.R = mine
            L.P_glob, L.T_glob = Pg, Tg
            B, bs, offsets = Bc, k, coff
        Lc = self.levels[-1]
        n = int(Lc.offsets[-1])
        self.coarse_direct = n <= dense_limit
        if self.coarse_direct:
            rows = comm.allgather_obj(self.global_rows(len(self.levels) - 1))    # gather the coarsest operator once
            self.coarse_inv = np.linalg.inv(sp.vstack(rows).toarray())[Lc.plan.offset: Lc.plan.offset + Lc.plan.n_owned]

    def global_rows(self, l) -> sp.csr_matrix:
        L = self.levels[l]
        gid = np.concatenate([L.plan.offset + np.arange(L.plan.n_owned, dtype=np.int64), L.plan.ghost_gid])
        return sp.csr_matrix((L.A.data, gid[L.A.indices], L.A.indptr), shape=(L.A.shape[0], int(L.offsets[-1])))

    def _matvec(self, L, x):
        return L.A @ halo_vec(self.comm, L.plan, x)

    def _power_lmax(self, L, its):
        gid = L.plan.offset + np.arange(L.plan.n_owned)
        v = (hash32(gid) % 2048).astype(float) / 1024.0 - 1.0
        v = v / np.sqrt(self.comm.allreduce_sum(v @ v))
        lam = 1.0
        for _ in range(its):
            w = L.dinv * self._matvec(L, v)
            lam = np.sqrt(self.comm.allreduce_sum(w @ w))
            if lam == 0:
                return 1.0
            v = w / lam
        return lam

    def _cheby(self, L, b, x, zero):
        lmax = L.lmax
        lmin = lmax / self.ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho = 1.0 / sigma
        r = b.copy() if zero else b - self._matvec(L, x)
        d = L.dinv * r / theta
        for k in range(self.deg):
            x = x + d
            if k == self.deg - 1:
                break
            r = r - self._matvec(L, d)
            rho_new = 1.0 / (2.0 * sigma - rho)
            d = rho_new * rho * d + (2.0 * rho_new / delta) * (L.dinv * r)
            rho = rho_new
        return x

    def _cycle(self, l, b):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            if self.coarse_direct:
                full = np.concatenate(self.comm.allgather_obj(b))
                return self.coarse_inv @ full
            x = self._cheby(L, b, np.zeros_like(b), True)
            return x if len(self.levels) == 1 else self._cheby(L, b, x, False)
        x = self._cheby(L, b, np.zeros_like(b), True)
        r = b - self._matvec(L, x)
        xc = self._cycle(l + 1, L.R @ halo_vec(self.comm, L.plan, r))
        x = x + L.P @ halo_vec(self.comm, self.levels[l + 1].plan, xc)
        return self._cheby(L, b, x, False)

    def __call__(self, b_owned):
        return self._cycle(0, b_owned)


def dist_selfp_schur(comm: Comm, plan0: Plan, A10_local: sp.csr_matrix, A01_rows_glob: sp.csr_matrix, d00_owned: np.ndarray,
                     A11_rows_glob: sp.csr_matrix) -> sp.csr_matrix:
    """Exact selfp Schur complement of the OWNED rows of split 1 (the pressure in the benchmark's split order):
        S[owned, :] = A11[owned, :] - A10[owned, [owned0 | ghost0]] diag(A00)^-1 A01[[owned0 | ghost0], :]
    A10_local has local columns [owned | ghost] of split 0 with halo plan `plan0`; the A01 rows and the diagonal of the
    ghost dofs arrive by one sparse-row and one vector halo exchange.  Columns of the result are global ids of split 1.
    Round 1 drops the ghost terms (capi.cu: S from owned parts only), which is part of its iteration growth."""
    d_ext = halo_vec(comm, plan0, d00_owned)
    A01_ext = sp.vstack([A01_rows_glob, halo_rows(comm, plan0, A01_rows_glob)]).tocsr()
    return (sp.csr_matrix(A11_rows_glob) - A10_local @ sp.diags(1.0 / d_ext) @ A01_ext).tocsr()
