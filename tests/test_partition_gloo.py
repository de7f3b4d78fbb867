"""CPU, world_size 2 over gloo: the z-slab row partition and its halo plan (host logic of the
multi-GPU path).  Each rank assembles only its slab, exchanges halo values with torch.distributed
exactly as the plan prescribes, and must reproduce its rows of the global SpMV."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, pc_type, q, overlap=True):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.problems import swelling
        from poro_b200.partition import distributed_problem
        prob = distributed_problem(3, N, pc_type, rank, world, ctx=None, overlap_rows=overlap)
        s, plan = prob.sys, prob.sys.plan
        glob, _ = swelling(3, N, pc_type)
        x = np.random.default_rng(7).standard_normal(glob.n)
        xo = x[s.owned_global]
        # halo exchange as the plan prescribes (poro_halo_set semantics)
        halo = np.zeros(plan.n_halo)
        reqs, off = [], 0
        for k, nb in enumerate(plan.neigh):
            sb = torch.from_numpy(np.ascontiguousarray(xo[plan.send_idx[plan.send_ptr[k]:plan.send_ptr[k + 1]]]))
            rb = torch.zeros(int(plan.recv_count[k]), dtype=torch.float64)
            reqs.append((dist.isend(sb, int(nb)), dist.irecv(rb, int(nb)), rb, off))
            off += int(plan.recv_count[k])
        for sreq, rreq, rb, o in reqs:
            sreq.wait(); rreq.wait()
            halo[o:o + len(rb)] = rb.numpy()
        assert np.array_equal(halo, x[plan.halo_global])
        xe = np.concatenate([xo, halo])
        for Ml, Mg in ((s.A, glob.A), (s.P, glob.P)):
            ref = (Mg @ x)[s.owned_global]
            got = Ml @ xe
            assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)
        assert np.allclose(s.b, glob.b[s.owned_global], rtol=1e-12, atol=1e-18)
        # overlap rows: the rows of the halo dofs, restricted to [owned | halo] columns, equal the global ones
        if overlap:
            ext_g = np.concatenate([s.owned_global, plan.halo_global])
            ref_ext = glob.P[ext_g][:, ext_g]
            assert s.P_ext.shape == ref_ext.shape
            assert abs(s.P_ext - ref_ext).max() <= 1e-12 * abs(ref_ext).max()
        else:
            assert s.P_ext is None
        n_own = torch.tensor([len(s.owned_global)])
        dist.all_reduce(n_own)
        assert int(n_own) == glob.n == prob.n_global
        # field index sets cover the extended vector, owned entries first
        ext = len(s.owned_global) + plan.n_halo
        cover = np.sort(np.concatenate([s.is_s, s.is_f, s.is_p]))
        assert np.array_equal(cover, np.arange(ext))
        q.put((rank, "ok", len(s.owned_global), plan.n_halo))
    except Exception as e:                                   # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,pc_type,world,overlap", [(3, "diagonal", 2, False), (4, "diagonal 3-way", 2, True), (4, "diagonal", 3, True)])
def test_slab_partition_gloo(N, pc_type, world, overlap):
    """world 3: the middle rank has two neighbours (the layout every rank but the ends has at 4 and 8 GPUs)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + N + 10 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, pc_type, q, overlap)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, n_own, n_halo in res:
        assert status == "ok", status
        assert n_own > 0 and n_halo > 0


def test_slab_ranges_cover():
    from poro_b200.partition import slab_ranges
    for planes, world in ((69, 8), (7, 2), (9, 4)):
        r = slab_ranges(planes, world)
        assert r[0][0] == 0 and r[-1][1] == planes
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1
