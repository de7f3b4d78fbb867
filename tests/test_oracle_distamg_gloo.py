"""world_size 2 and 3 over gloo: every rank builds ITS part of the distributed SA-AMG hierarchy from its own rows and
real messages (oracle/distamg_rank.py) and must reproduce the single-process emulation (oracle/distamg.py) piece by
piece: level operators, prolongators, restrictions, eigenvalue estimates and the V-cycle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, kw, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        from test_oracle_distamg import _slab_perm
        from oracle.amg import rigid_body_modes
        from oracle.distamg import DistAmg
        from oracle.distamg_rank import Comm, RankAmg
        from oracle.problems import swelling
        sys_, par = swelling(3, N, "diagonal")
        A = sp.csr_matrix(sys_.P)[sys_.is_s][:, sys_.is_s].tocsr()
        perm, off = _slab_perm(sys_.coords_s, N, world, 3)
        Ap = A[perm][:, perm].tocsr()
        B = rigid_body_modes(sys_.coords_s, 3)[perm]
        a, b = int(off[rank]), int(off[rank + 1])
        comm = Comm()
        mine = RankAmg(comm, Ap[a:b], off, 3, B[a:b], theta=0.04, **kw)       # this rank sees only its rows
        ref = DistAmg(Ap, off, 3, B, theta=0.04, **kw)                        # the emulation of all ranks
        assert len(mine.levels) == len(ref.levels)
        for l, L in enumerate(mine.levels):
            E = ref.levels[l][rank]
            assert np.array_equal(L.plan.ghost_gid, E.plan.ghost_gid)
            assert sorted(L.plan.send) == sorted(E.plan.send_idx)
            for p_ in L.plan.send:
                assert np.array_equal(L.plan.send[p_], E.plan.send_idx[p_])
            assert abs(L.lmax - E.lmax) <= 1e-12 * E.lmax
            assert L.A.shape == E.A.shape and abs(L.A - E.A).max() <= 1e-12 * abs(E.A).max()
            if l < len(mine.levels) - 1:
                assert L.P.shape == E.P.shape and abs(L.P - E.P).max() <= 1e-12 * abs(E.P).max()
                assert L.R.shape == E.R.shape and abs(L.R - E.R).max() <= 1e-12 * abs(E.R).max()
        rhs = np.random.default_rng(5).standard_normal(Ap.shape[0])
        y = mine(rhs[a:b])
        y_ref = ref(rhs)[a:b]
        assert np.abs(y - y_ref).max() <= 1e-10 * np.abs(y_ref).max()
        # exact selfp Schur complement of the owned pressure rows (S_p = A_pp - A_pf diag(A_ff)^-1 A_fp)
        from oracle.distamg_rank import Plan, dist_selfp_schur, localize
        Pm = sp.csr_matrix(sys_.P)
        perm_f, off_f = _slab_perm(sys_.coords_s, N, world, 3)          # f lives on the same nodes as s
        perm_p, off_p = _slab_perm(sys_.coords_p, N, world, 1)
        Aff = Pm[sys_.is_f][:, sys_.is_f][perm_f][:, perm_f].tocsr()
        Afp = Pm[sys_.is_f][:, sys_.is_p][perm_f][:, perm_p].tocsr()
        Apf = Pm[sys_.is_p][:, sys_.is_f][perm_p][:, perm_f].tocsr()
        App = Pm[sys_.is_p][:, sys_.is_p][perm_p][:, perm_p].tocsr()
        fa, fb, pa, pb = int(off_f[rank]), int(off_f[rank + 1]), int(off_p[rank]), int(off_p[rank + 1])
        Apf_loc, ghost_f = localize(Apf[pa:pb], fa, fb)
        plan_f = Plan(comm, off_f, ghost_f)
        S_rows = dist_selfp_schur(comm, plan_f, Apf_loc, Afp[fa:fb], Aff.diagonal()[fa:fb], App[pa:pb])
        S_ref = (App - Apf @ sp.diags(1.0 / Aff.diagonal()) @ Afp).tocsr()[pa:pb]
        assert abs(S_rows - S_ref).max() <= 1e-13 * abs(S_ref).max()
        S_owned_only = (App[pa:pb] - Apf[pa:pb, fa:fb] @ sp.diags(1.0 / Aff.diagonal()[fa:fb]) @ Afp[fa:fb]).tocsr()
        assert abs(S_owned_only - S_ref).max() > 1e-3 * abs(S_ref).max()      # what round 1 assembles is NOT the Schur complement
        # the same exchange with the lumped mass + drag diagonal = the `cc` Schur matrix of the benchmarked option set
        # (capi.cu hands `mass_scale |diag(A_fs)|` to dist_selfp_schur instead of diag(P_ff))
        from oracle.blockpc import cc_from_matrices
        d_mass = cc_from_matrices(sys_, par)[0][perm_f]
        S_cc = dist_selfp_schur(comm, plan_f, Apf_loc, Afp[fa:fb], d_mass[fa:fb], App[pa:pb])
        S_cc_ref = (App - Apf @ sp.diags(1.0 / d_mass) @ Afp).tocsr()[pa:pb]
        assert abs(S_cc - S_cc_ref).max() <= 1e-13 * abs(S_cc_ref).max()
        assert abs(S_cc_ref - S_ref).max() > 1e-3 * abs(S_ref).max()           # and it is not selfp under another name
        q.put((rank, "ok", len(mine.levels), comm.messages))
    except Exception:                                        # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,world,kw", [(4, 2, dict(coarse_size=100)), (6, 3, dict(coarse_size=100)),
                                         (4, 2, dict(coarse_size=100, dense_limit=0))])
def test_rank_local_hierarchy_equals_emulation(N, world, kw):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + N + 10 * world + (7 if "dense_limit" in kw else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, kw, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, nlev, msgs in res:
        assert status == "ok", status
        assert nlev >= 2 and msgs > 0
