"""Rank-level distributed SA-AMG set-up and V-cycle (oracle/distamg.py, the blueprint for the multi-GPU
hierarchy) against the global formulas and against the single-process solver."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.amg import SAAMG, rigid_body_modes
from oracle.blockpc import BlockPC, SchurLower, krylov_solver
from oracle.distamg import DistAmg, GlobalCycle, distribute, halo_vec, make_plans
from oracle.krylov import gmres
from oracle.problems import swelling
from poro_b200.partition import slab_ranges


def _slab_perm(coords, n_cells, world, bs):
    """Rank-contiguous numbering of a node-blocked field cut into z-slabs: (perm of dofs, offsets)."""
    h2 = coords[:, 2].max() / (2 * n_cells)
    plane = np.rint(coords[::bs, 2] / h2).astype(int)
    parts = [np.flatnonzero((plane >= a) & (plane < b)) for a, b in slab_ranges(2 * n_cells + 1, world)]
    perm = np.concatenate([(p[:, None] * bs + np.arange(bs)).ravel() for p in parts])
    offsets = np.concatenate([[0], np.cumsum([len(p) * bs for p in parts])])
    return perm, offsets


@pytest.fixture(scope="module")
def solid():
    sys_, _ = swelling(3, 6, "diagonal")
    A = sp.csr_matrix(sys_.P)[sys_.is_s][:, sys_.is_s].tocsr()
    return sys_, A


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_reproduces_the_global_product(solid, world):
    sys_, A = solid
    perm, off = _slab_perm(sys_.coords_s, 6, world, 3)
    Ap = A[perm][:, perm].tocsr()
    mats, ghosts = distribute(Ap, off)
    plans = make_plans(off, ghosts)
    x = np.random.default_rng(1).standard_normal(Ap.shape[0])
    ext = halo_vec(plans, [x[off[r]: off[r + 1]] for r in range(world)])
    y = np.concatenate([M @ e for M, e in zip(mats, ext)])
    np.testing.assert_allclose(y, Ap @ x, rtol=1e-13, atol=1e-13 * np.abs(Ap @ x).max())
    for r, p in enumerate(plans):                       # send lists and ghost lists are two views of the same handshake
        for q in p.recv_slice:
            np.testing.assert_array_equal(plans[q].send_idx[r] + off[q], p.ghost_gid[p.recv_slice[q]])


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_setup_matches_global_formulas(solid, world):
    sys_, A = solid
    perm, off = _slab_perm(sys_.coords_s, 6, world, 3)
    Ap = A[perm][:, perm].tocsr()
    B = rigid_body_modes(sys_.coords_s, 3)[perm]
    h = DistAmg(Ap, off, 3, B, theta=0.04, coarse_size=100)
    assert len(h.levels) >= 3
    assert abs(h.global_matrix(0) - Ap).max() == 0
    for l in range(len(h.levels) - 1):
        Al, T, P = h.global_matrix(l), h.global_T(l), h.global_P(l)
        lev = h.levels[l]
        dinv = np.concatenate([L.dinv for L in lev])
        omega = 4.0 / (3.0 * lev[0].lmax / 1.1)
        P_ref = (T - sp.diags(omega * dinv) @ (Al @ T)).tocsr()
        assert abs(P - P_ref).max() <= 1e-13 * abs(P_ref).max()
        Ac_ref = (P.T @ Al @ P).tocsr()
        Ac = h.global_matrix(l + 1)
        dead = Ac_ref.diagonal() == 0
        Ac_ref = Ac_ref + sp.diags(dead.astype(float))
        assert abs(Ac - Ac_ref).max() <= 1e-12 * abs(Ac_ref).max()
        # aggregates never cross a rank boundary: T is block diagonal over the ranks
        offc = h.offsets[l + 1]
        offf = h.offsets[l]
        Tc = T.tocoo()
        assert np.array_equal(np.searchsorted(offf, Tc.row, side="right"), np.searchsorted(offc, Tc.col, side="right"))
        # the rank-local R is the transpose of the P columns the rank owns
        for r, L in enumerate(lev):
            Rg = sp.csr_matrix((L.R.data, np.concatenate([L.plan.offset + np.arange(L.plan.n_owned), L.plan.ghost_gid])[L.R.indices],
                                L.R.indptr), shape=(L.R.shape[0], Al.shape[0]))
            assert abs(Rg - P[:, offc[r]: offc[r + 1]].T).max() <= 1e-15


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_cycle_matches_single_process_cycle(solid, world):
    sys_, A = solid
    perm, off = _slab_perm(sys_.coords_s, 6, world, 3)
    Ap = A[perm][:, perm].tocsr()
    B = rigid_body_modes(sys_.coords_s, 3)[perm]
    for kw in (dict(coarse_size=100), dict(coarse_size=100, dense_limit=0), dict(max_levels=1, cheby_degree=4)):
        h = DistAmg(Ap, off, 3, B, theta=0.04, **kw)
        g = GlobalCycle(h)
        b = np.random.default_rng(2).standard_normal(Ap.shape[0])
        y, y_ref = h(b), g(b)
        np.testing.assert_allclose(y, y_ref, rtol=0, atol=1e-11 * np.abs(y_ref).max())
    assert h.cycle_traffic.messages > 0


def test_replicated_tail(solid):
    """Small coarse levels gathered on every rank: same quality of the cycle, fewer messages."""
    sys_, A = solid
    perm, off = _slab_perm(sys_.coords_s, 6, 3, 3)
    Ap = A[perm][:, perm].tocsr()
    B = rigid_body_modes(sys_.coords_s, 3)[perm]
    b = np.random.default_rng(3).standard_normal(Ap.shape[0])
    out = {}
    for rb in (0, 2000):
        h = DistAmg(Ap, off, 3, B, theta=0.04, coarse_size=100, replicate_below=rb)
        x = np.zeros_like(b)
        for _ in range(8):                                  # stationary iteration with the cycle as preconditioner
            x = x + h(b - Ap @ x)
        out[rb] = (np.linalg.norm(b - Ap @ x) / np.linalg.norm(b), h.cycle_traffic.messages, h)
    assert out[2000][2].tail is not None and out[0][2].tail is None
    assert out[2000][0] < 1e-2 and out[0][0] < 1e-2
    assert out[2000][0] < 3 * out[0][0]
    assert out[2000][1] < out[0][1]


def test_uncoupled_labels_equal_rank_level_setup_in_iteration_count():
    """Outer GMRES with the distributed hierarchies for s and S_p needs (about) the single-GPU iteration count;
    the round-1 rank-local coarse levels (oracle/ddamg.py) need far more."""
    n, world = 6, 3
    sys_, _ = swelling(3, n, "diagonal")
    Bg = rigid_body_modes(sys_.coords_s, 3)
    perm_s, off_s = _slab_perm(sys_.coords_s, n, world, 3)
    perm_p, off_p = _slab_perm(sys_.coords_p, n, world, 1)

    class Permuted:
        def __init__(self, M, perm, off, bs, B, **kw):
            self.perm = perm
            self.h = DistAmg(sp.csr_matrix(M)[perm][:, perm].tocsr(), off, bs, None if B is None else B[perm], **kw)

        def __call__(self, b):
            y = np.empty_like(b)
            y[self.perm] = self.h(b[self.perm])
            return y

    cheb_f = lambda M: SAAMG(M, 3, Bg, max_levels=1, cheby_degree=4)

    def solve(mk_s, mk_p):
        mkfp = lambda M: SchurLower(M, sys_.nf, sys_.np_, krylov_solver("preonly", cheb_f), krylov_solver("preonly", mk_p), "f")
        pc = BlockPC(sys_, {"s": krylov_solver("preonly", mk_s), "fp": mkfp})
        return gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=1e-8, atol=0.0, dtol=1e20, max_it=300, restart=300,
                     pc_side="right")

    single = solve(lambda M: SAAMG(M, 3, Bg, theta=0.04), lambda M: SAAMG(M, 1, None))
    dist = solve(lambda M: Permuted(M, perm_s, off_s, 3, Bg, theta=0.04), lambda M: Permuted(M, perm_p, off_p, 1, None))
    assert single.reason == 2 and dist.reason == 2
    assert abs(dist.its - single.its) <= 2, (dist.its, single.its)
    np.testing.assert_allclose(dist.x, single.x, rtol=0, atol=1e-6 * np.abs(single.x).max())
