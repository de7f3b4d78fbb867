"""Device-side generator of the structured-mesh systems (csrc/gen.cu + poro_b200/generator.py).

CPU: the analytic slab layout (owned / ghost ranges, halo plan, index sets) against the global numbering.
GPU: the generated local matrices of every rank of a 1-, 2- and 3-slab partition against the rows of the matrix the host
element assembler builds (pattern identical after dropping zeros, values to 1e-14), and a solve on generated matrices."""
import numpy as np
import pytest
import scipy.sparse as sp

from poro_b200.generator import SlabLayout


@pytest.mark.parametrize("N,world", [(3, 1), (4, 2), (5, 3), (6, 4)])
def test_slab_layout_is_a_partition_with_matching_halos(N, world):
    lays = [SlabLayout(3, N, r, world) for r in range(world)]
    n = lays[0].n_global
    owned = np.concatenate([l.owned_global() for l in lays])
    assert len(owned) == n and np.array_equal(np.sort(owned), np.arange(n))
    for r, l in enumerate(lays):
        ext = l.ext_global()
        assert len(ext) == l.n_ext and len(np.unique(ext)) == len(ext)
        is_s, is_f, is_p = l.index_sets()
        assert len(is_s) + len(is_f) + len(is_p) == l.n_ext
        d, n2 = 3, l.n2
        assert np.all(ext[is_s] < d * n2) and np.all((ext[is_f] >= d * n2) & (ext[is_f] < 2 * d * n2)) and np.all(ext[is_p] >= 2 * d * n2)
        plan = l.halo_plan()
        # what I receive from a neighbour is exactly what that neighbour sends me, in the same order
        pos = l.n_owned
        for k, q in enumerate(plan.neigh):
            pq = lays[q].halo_plan()
            kq = list(pq.neigh).index(r)
            sent = lays[q].owned_global()[pq.send_idx[pq.send_ptr[kq]:pq.send_ptr[kq + 1]]]
            got = ext[pos:pos + plan.recv_count[k]]
            assert np.array_equal(sent, got)
            pos += plan.recv_count[k]
        assert pos == l.n_ext


@pytest.mark.gpu
@pytest.mark.parametrize("N,world,pc_type", [(3, 1, "diagonal"), (4, 2, "diagonal"), (5, 3, "diagonal 3-way"), (4, 2, "undrained")])
def test_generated_matrices_equal_host_assembly(gpu_ctx, N, world, pc_type):
    from hostfem.problems import swelling
    from poro_b200.generator import generate_swelling3d
    ref, _ = swelling(3, N, pc_type)
    for rank in range(world):
        g = generate_swelling3d(gpu_ctx, N, pc_type, rank, world, init_dist=False)
        ext = g.layout.ext_global()
        rows = g.owned_global
        for name, got, want in (("A", g.A, ref.A), ("P", g.P, ref.P), ("P_diff", g.P_diff, ref.P_diff)):
            if got is None:
                assert want is None
                continue
            G = got.to_scipy()
            assert G.shape == (len(rows), len(ext))
            W = sp.csr_matrix(want)[rows].tocsc()
            # every column the owned rows reference must be in [owned | ghost]
            used = np.unique(sp.csr_matrix(want)[rows].indices)
            assert np.all(np.isin(used, ext)), name
            W = W[:, ext].tocsr()
            W.eliminate_zeros(); W.sort_indices()
            G.sort_indices()
            assert G.nnz == W.nnz, (name, rank, G.nnz, W.nnz)
            assert np.array_equal(G.indptr, W.indptr) and np.array_equal(G.indices, W.indices), (name, rank)
            assert abs(G - W).max() <= 1e-14 * abs(W).max(), (name, rank)
        np.testing.assert_allclose(g.b, ref.b[rows], rtol=1e-13, atol=1e-14 * np.abs(ref.b).max())
        assert np.array_equal(np.sort(g.bcs_sub_pressure), np.flatnonzero(np.isin(rows[g.layout.off_owned[2]:] - 2 * 3 * g.layout.n2,
                                                                                  ref.bcs_sub_pressure)))
    gpu_ctx.set_halo(0, [], [0], [], [])             # leave the shared context in its single-rank state


@pytest.mark.gpu
def test_solve_on_generated_system_matches_host_assembled_solve(gpu_ctx):
    """End to end on one rank: matrices generated in HBM, solved with the benchmarked options, against the same solve on
    host-assembled matrices."""
    import bench
    from helpers import gpu_solve, rel
    from hostfem.problems import swelling
    from poro_b200.generator import generate_swelling3d
    from poro_b200.lib.backend import DeviceVector
    from poro_b200.lib.Parser import load_petsc_options
    from poro_b200.lib.Preconditioner import Preconditioner
    from poro_b200.lib.Solver import Solver
    N = 6
    ref, par = swelling(3, N, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-10, "solver atol": 0.0, "solver maxiter": 100})
    want = gpu_solve(ref, par, bench.BENCH_OPTIONS)
    gpu_ctx.clear_options()
    load_petsc_options(gpu_ctx, bench.BENCH_OPTIONS, is_text=True)
    g = generate_swelling3d(gpu_ctx, N, "diagonal")
    imap = g.index_set()
    db, dx = DeviceVector(g.b, ctx=gpu_ctx), DeviceVector(n=len(g.b), ctx=gpu_ctx)
    pc = Preconditioner(imap, g.A, g.P, None, par, g.bcs_sub_pressure).get_pc()
    solver = Solver(g.A, db, pc, par, imap)
    solver.create_solver(g.A, db, pc)
    solver.solve(db, dx)
    assert solver.solver.reason == 2 and solver.getIterationNumber() == want["its"]
    assert rel(dx.numpy(), want["x"]) <= 1e-9
