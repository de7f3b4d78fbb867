"""GPU tests added in round 2:
  * the BENCHMARKED option set (options/petsc-options-b200 = bench.BENCH_OPTIONS) against its CPU twin
    (bench.oracle_solver): iteration counts within +-10 %, solution within 1e-8 of a direct solve, true residual;
  * time loop with solver / hierarchy reuse and warm start (lib/AbstractPhysics.py:73-81,
    lib/Poromechanics.py:70-98);
  * per-field infinity-norm monitor (lib/Solver.py:8-51), -ksp_converged_reason;
  * advisor findings: outer CG on an operator with node-blocked parts, `poro_*` tunables, `chebyshev` at every size,
    indefinite operator under CG.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from helpers import AMG_OPTIONS, EXACT_OPTIONS, gpu_solve, rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_options():
    import bench
    txt = open(os.path.join(ROOT, "options", "petsc-options-b200")).read()
    keys = lambda t: sorted(l.strip() for l in t.splitlines() if l.strip() and not l.strip().startswith("#"))
    assert keys(txt) == keys(bench.BENCH_OPTIONS), "options/petsc-options-b200 and bench.BENCH_OPTIONS differ"
    return bench.BENCH_OPTIONS


@pytest.mark.parametrize("N", [8, 16])
@pytest.mark.parametrize("which", ["bench", "selfp"])
def test_benchmarked_option_set_matches_cpu_twin(gpu_ctx, N, which):
    """bench.py's configuration (s, f: V-cycle theta 0.04; p: additive Cahouet-Chabard Schur preconditioner, split order fp) and
    the PETSc-`selfp` set it replaced (f: Chebyshev(4), p: V-cycle on selfp S_p), each against its CPU twin."""
    import bench
    from oracle.problems import swelling
    opts = _bench_options() if which == "bench" else bench.BENCH_OPTIONS_SELFP
    sys_, par = swelling(3, N, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-8, "solver atol": 0.0, "solver maxiter": 100, "solver type": "gmres"})
    runs = bench.oracle_solver(sys_, par, 100, opts)
    ro = runs["numpy/scipy, 1 thread"][0]()
    g = gpu_solve(sys_, par, opts)
    assert g["reason"] == 2 and ro.reason == 2
    assert abs(g["its"] - ro.its) <= max(1, int(round(0.1 * ro.its))), (g["its"], ro.its)      # north star: +-10 %
    res = np.linalg.norm(sys_.b - sys_.A @ g["x"]) / np.linalg.norm(sys_.b)
    assert res <= 1.0e-8, res
    # the two solves stop on the same test: their solutions agree far below the tolerance-induced error
    assert rel(g["x"], ro.x) <= 1e-6
    # same algorithm, hierarchies that differ in rounding-level details (eigenvalue estimates, summation order): the
    # residual histories track each other to a few per cent
    np.testing.assert_allclose(g["history"][:4], ro.history[:4], rtol=2e-2)
    if N == 8:
        # solution parity proper: both iterated to 1e-13 must sit within 1e-8 of the direct solve
        par12 = dict(par)
        par12["solver rtol"] = 1e-13
        g12 = gpu_solve(sys_, par12, opts)
        xd = spla.spsolve(sys_.A.tocsc(), sys_.b)
        assert g12["reason"] == 2
        assert rel(g12["x"], xd) <= 1e-8, rel(g12["x"], xd)


def test_chebyshev_pc_is_chebyshev_at_every_size(gpu_ctx):
    """`-pc_type chebyshev` never turns into a dense exact solve on small blocks (advisor, round 1)."""
    from oracle.amg import SAAMG, rigid_body_modes
    from oracle.problems import swelling
    from poro_b200.lib.backend import DeviceVector
    sys_, par = swelling(3, 3, "diagonal")            # f block: 3 * 7^3 = 1029 rows << 4096
    import bench
    g = gpu_solve(sys_, par, bench.BENCH_OPTIONS_SELFP, return_objects=True)
    cc = g["pc"].pc.getPythonContext()
    Pff = sp.csr_matrix(sys_.P)[sys_.is_f][:, sys_.is_f].tocsr()
    twin = SAAMG(Pff, 3, rigid_body_modes(sys_.coords_s, 3), max_levels=1, cheby_degree=4, dense_limit=0)
    r = np.random.default_rng(5).standard_normal(sys_.nf)
    dr, dz = DeviceVector(r, ctx=gpu_ctx), DeviceVector(n=sys_.nf, ctx=gpu_ctx)
    cc.inner_solve("fp0", dr, dz)
    gpu_ctx.sync()
    z = dz.numpy()
    assert rel(z, twin(r)) <= 1e-10
    assert rel(z, spla.spsolve(Pff.tocsc(), r)) > 1e-3        # i.e. NOT the exact inverse


def test_dense_lu_limit_option_is_honoured(gpu_ctx):
    """`-poro_dense_lu_limit` changes the kind of an `lu` block (the key used to be looked up without its dash)."""
    from oracle.problems import swelling
    sys_, par = swelling(2, 6, "diagonal")
    g1 = gpu_solve(sys_, par, EXACT_OPTIONS, return_objects=True)
    lv1 = g1["pc"].pc.getPythonContext().amg_info("s")
    g2 = gpu_solve(sys_, par, EXACT_OPTIONS + "\n-poro_dense_lu_limit 10\n", return_objects=True)
    lv2 = g2["pc"].pc.getPythonContext().amg_info("s")
    assert len(lv1) == 0 and len(lv2) >= 1            # dense inverse vs iterated AMG stand-in
    assert rel(g2["x"], g1["x"]) <= 1e-7


def test_outer_cg_with_node_blocked_parts(gpu_ctx):
    """Outer `cg` on an SPD system whose operator is cut into BSR parts + CSR remainder (block_dim > 1): the fused
    spmv+dot shortcut must not drop the parts (advisor, round 1)."""
    from oracle.problems import swelling
    from oracle.krylov import cg
    sys_, par = swelling(2, 8, "diagonal")
    # an SPD operator with the same field layout: block-diagonal of the symmetrised field blocks
    n = sys_.n
    blocks = []
    for idx in (sys_.is_s, sys_.is_f, sys_.is_p):
        M = sp.csr_matrix(sys_.A)[idx][:, idx]
        blocks.append(((M + M.T) * 0.5 + sp.identity(len(idx)) * abs(M).max()).tocsr())
    perm = np.concatenate([sys_.is_s, sys_.is_f, sys_.is_p])
    Aspd_p = sp.block_diag(blocks).tocsr()
    inv = np.argsort(perm)
    Aspd = Aspd_p[inv][:, inv].tocsr()

    class S:
        pass
    s2 = S()
    s2.__dict__.update(sys_.__dict__)
    s2.A, s2.P = Aspd, Aspd
    par = dict(par)
    par.update({"solver type": "cg", "solver rtol": 1e-10, "solver atol": 0.0, "solver maxiter": 400})
    g = gpu_solve(s2, par, EXACT_OPTIONS.replace("-global_ksp_type gmres", "-global_ksp_type cg"))
    xd = spla.spsolve(Aspd.tocsc(), sys_.b)
    assert g["reason"] in (2, 3)
    assert rel(g["x"], xd) <= 1e-8


def test_cg_reports_indefinite_operator(gpu_ctx):
    """p.Ap <= 0 stops CG with KSP_DIVERGED_INDEFINITE_MAT (-10) and leaves x finite."""
    from oracle.problems import swelling
    sys_, par = swelling(2, 6, "diagonal")

    class S:
        pass
    s2 = S()
    s2.__dict__.update(sys_.__dict__)
    s2.A = (-sp.identity(sys_.n)).tocsr()
    par = dict(par)
    par.update({"solver type": "cg", "solver maxiter": 20})
    g = gpu_solve(s2, par, EXACT_OPTIONS.replace("-global_ksp_type gmres", "-global_ksp_type cg"))
    assert g["reason"] == -10
    assert np.all(np.isfinite(g["x"]))


def test_time_loop_reuses_solver_and_warm_starts(gpu_ctx):
    """the time loop of lib/AbstractPhysics.py:73-81 -- solver, preconditioner and AMG hierarchies created
    at the first step (lib/Poromechanics.py:91-93) and reused; only the right-hand side changes.  With
    setInitialGuessNonzero (lib/Solver.py:94) the previous step's solution is the start."""
    from oracle.problems import swelling
    from poro_b200.lib.backend import DeviceVector
    sys_, par = swelling(3, 4, "diagonal", {"tf": 0.4})
    par = dict(par)
    par.update({"solver rtol": 1e-9, "solver atol": 0.0, "solver maxiter": 100})
    g = gpu_solve(sys_, par, AMG_OPTIONS, return_objects=True)
    solver, ksp = g["solver"], g["solver"].solver
    launches_setup = gpu_ctx.launch_count()
    lu = spla.splu(sys_.A.tocsc())
    x = DeviceVector(g["x"], ctx=gpu_ctx)
    its_cold, its_warm = [], []
    for t in (0.2, 0.3, 0.4):
        b = sys_.meta["rhs_at"](t)
        db = DeviceVector(b, ctx=gpu_ctx)
        xc = DeviceVector(n=sys_.n, ctx=gpu_ctx)
        ksp.setInitialGuessNonzero(False)
        solver.solve(db, xc)
        its_cold.append(solver.getIterationNumber())
        assert ksp.reason == 2
        ksp.setInitialGuessNonzero(True)
        solver.solve(db, x)                               # x holds the previous step's solution
        its_warm.append(solver.getIterationNumber())
        assert ksp.reason in (2, 3)
        xd = lu.solve(b)
        assert rel(x.numpy(), xd) <= 1e-7 and rel(xc.numpy(), xd) <= 1e-7
        res = np.linalg.norm(b - sys_.A @ x.numpy()) / np.linalg.norm(b)
        assert res <= 2e-9
    # loads grow smoothly in t (1 - exp(-t^2/0.25)): the warm start never costs iterations and saves some
    assert all(w <= c for w, c in zip(its_warm, its_cold)) and sum(its_warm) < sum(its_cold), (its_warm, its_cold)
    # nothing was set up again: a re-setup launches tens of thousands of kernels, six solves far fewer
    stats = g["pc"].pc.getPythonContext().stats()
    assert stats["calls_s"] > 0
    assert gpu_ctx.launch_count() - launches_setup < 60000


def test_field_monitor_and_converged_reason(gpu_ctx, capfd):
    """the per-field infinity-norm residual monitor of lib/Solver.py:8-51 and -ksp_converged_reason."""
    from oracle.blockpc import BlockPC, exact_solvers
    from oracle.krylov import gmres
    from oracle.problems import swelling
    sys_, par = swelling(2, 8, "diagonal")
    g = gpu_solve(sys_, par, EXACT_OPTIONS + "\n-global_ksp_monitor_fields\n-global_ksp_converged_reason\n", return_objects=True)
    gpu_ctx.sync()
    ksp = g["solver"].solver
    fh = ksp.getFieldHistory()
    assert len(fh) == ksp.its + 1
    # oracle: true residual of the GMRES iterate after every step, per-field infinity norms
    pc = BlockPC(sys_, exact_solvers())
    for it in (0, 1, ksp.its):
        if it == 0:
            r = sys_.b.copy()
        else:
            ro = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=0.0, atol=0.0, dtol=1e20, max_it=it, restart=par["solver maxiter"], pc_side="right")
            r = sys_.b - sys_.A @ ro.x
        ref = [np.abs(r[i]).max() for i in (sys_.is_s, sys_.is_f, sys_.is_p)]
        np.testing.assert_allclose(fh[it], ref, rtol=1e-5, atol=1e-12 * np.abs(sys_.b).max())
    out = capfd.readouterr().out
    assert "KSP errors:" in out and "KSP it 0:" in out
    assert "Linear global_ solve converged due to CONVERGED_RTOL iterations %d" % ksp.its in out or \
           "Linear global_ solve converged due to CONVERGED_ATOL iterations %d" % ksp.its in out


def test_field_convergence_test_stops_on_field_norms(gpu_ctx):
    """`-global_ksp_convergence_test_fields`: the reference's `converged` as the stopping test (rtol against the largest
    ||b_field||_2, lib/Solver.py:17,41)."""
    from oracle.problems import swelling
    sys_, par = swelling(2, 8, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-6, "solver atol": 0.0})
    g = gpu_solve(sys_, par, EXACT_OPTIONS + "\n-global_ksp_convergence_test_fields\n")
    assert g["reason"] == 2
    r = sys_.b - sys_.A @ g["x"]
    normalize = max(np.linalg.norm(sys_.b[i]) for i in (sys_.is_s, sys_.is_f, sys_.is_p))
    assert max(np.abs(r[i]).max() for i in (sys_.is_s, sys_.is_f, sys_.is_p)) / normalize < 1e-6


@pytest.mark.parametrize("block", ["s", "fp"])
def test_single_block_drivers(gpu_ctx, block):
    """solid.py / fluid-pressure.py (reference solid.py:111-180, fluid-pressure.py:85-136) -- the inner solver of
    one block as a stand-alone KSP, against a direct solve of that block assembled on the host."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(ROOT, "examples"))
    from _single_block import DEFAULTS, build
    from hostfem.problems import swelling
    N = 5
    ctx, cc, db, dx, host, rhs, keep = build(block, N, assemble="device", ctx=gpu_ctx, options_text=DEFAULTS[block])
    cc.inner_solve(block, db, dx)
    gpu_ctx.sync()
    its, reason, rnorm = cc.inner_result(block)
    ref, _ = swelling(3, N, "diagonal", {"ks": 1e6})
    idx = ref.is_s if block == "s" else np.concatenate([ref.is_f, ref.is_p])
    M = sp.csr_matrix(ref.A)[idx][:, idx].tocsc()
    np.testing.assert_allclose(rhs, ref.b[idx], rtol=1e-12, atol=1e-14 * np.abs(ref.b).max())
    xd = spla.spsolve(M, rhs)
    assert reason > 0 and 0 < its < 200
    x = dx.numpy()
    assert np.linalg.norm(rhs - M @ x) / np.linalg.norm(rhs) <= 1e-7
    assert rel(x, xd) <= 1e-5
    gpu_ctx.set_halo(0, [], [0], [], [])


def test_fp32_storage_of_preconditioner_matrices_is_only_a_preconditioner_change(gpu_ctx):
    """Opt-in `-poro_pc_fp32_matrices 1`: the hierarchies' operators are STORED in fp32 (arithmetic, vectors and the outer operator
    stay fp64).  The preconditioner remains a fixed linear operator, so GMRES converges to the same solution; only the iteration
    count may move by a step."""
    import bench
    from oracle.problems import swelling
    sys_, par = swelling(3, 8, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-10, "solver atol": 0.0, "solver maxiter": 100, "solver type": "gmres"})
    g64 = gpu_solve(sys_, par, bench.BENCH_OPTIONS)
    g32 = gpu_solve(sys_, par, bench.BENCH_OPTIONS + "\n-poro_pc_fp32_matrices 1\n")
    assert g64["reason"] == 2 and g32["reason"] == 2
    assert abs(g32["its"] - g64["its"]) <= 2
    res = np.linalg.norm(sys_.b - sys_.A @ g32["x"]) / np.linalg.norm(sys_.b)
    assert res <= 1.01e-10
    assert rel(g32["x"], g64["x"]) <= 1e-8


@pytest.mark.parametrize("extra", ["-poro_bsr_coop_gather 1", "-poro_bsr_l2_prefetch_chunks 4", "-poro_bsr_tma 1",
                                   "-poro_pc_overlap_blocks 0", "-poro_pc_graph 0"])
def test_optional_bsr_kernel_variants_are_the_same_operator(gpu_ctx, extra):
    """The measured-and-rejected kernel variants (profiles/r2_bsr_kernels.md) stay selectable; they must be the same linear
    operators as the default kernel: same iteration count, same solution.  Likewise the scheduling switches: the solid and
    the fluid-pressure solve as a chain instead of two parallel graph branches, and eager launches instead of the graph."""
    import bench
    from oracle.problems import swelling
    sys_, par = swelling(3, 8, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-10, "solver atol": 0.0, "solver maxiter": 100, "solver type": "gmres"})
    ref = gpu_solve(sys_, par, bench.BENCH_OPTIONS)
    alt = gpu_solve(sys_, par, bench.BENCH_OPTIONS + "\n" + extra + "\n")
    assert alt["reason"] == ref["reason"] == 2
    assert alt["its"] == ref["its"]
    assert rel(alt["x"], ref["x"]) <= 1e-10


@pytest.mark.parametrize("N", [8])
def test_cc_schur_preconditioner_matches_cpu_twin(gpu_ctx, N):
    """`-fp_pc_fieldsplit_schur_precondition cc` (additive Cahouet-Chabard form, the reference's 3-way pressure treatment inside
    the 2-way fieldsplit) + V-cycle on the velocity block: the library derives the lumped mass from A_fs and the viscous limit
    from P_pp with the two scalars lib/Preconditioner.py computes from the parameter dict; the CPU twin (oracle.blockpc.
    SchurLowerCC) gets them from the same formulas.  Same iteration count (+-10 %), solution within 1e-8 of a direct solve."""
    import bench
    from oracle.problems import swelling
    sys_, par = swelling(3, N, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-8, "solver atol": 0.0, "solver maxiter": 100, "solver type": "gmres"})
    ro = bench.oracle_solver(sys_, par, 100, bench.BENCH_OPTIONS_CC)["numpy/scipy, 1 thread"][0]()
    g = gpu_solve(sys_, par, bench.BENCH_OPTIONS_CC)
    assert g["reason"] == 2 and ro.reason == 2
    assert abs(g["its"] - ro.its) <= max(1, int(round(0.1 * ro.its))), (g["its"], ro.its)
    res = np.linalg.norm(sys_.b - sys_.A @ g["x"]) / np.linalg.norm(sys_.b)
    assert res <= 1.0e-8, res
    assert rel(g["x"], ro.x) <= 1e-6
    np.testing.assert_allclose(g["history"][:4], ro.history[:4], rtol=2e-2)
    # the Schur matrix the library assembled is the twin's S_mass
    cc = g["pc"].pc.getPythonContext() if hasattr(g["pc"], "pc") else None
    if N == 8:
        from oracle.blockpc import cc_from_matrices, submatrix
        d_mass, _ = cc_from_matrices(sys_, par)
        Pfp = sp.csr_matrix(sys_.P)
        App, Apf, Afp = submatrix(Pfp, sys_.is_p, sys_.is_p), submatrix(Pfp, sys_.is_p, sys_.is_f), submatrix(Pfp, sys_.is_f, sys_.is_p)
        S_twin = (App - Apf @ sp.diags(1.0 / d_mass) @ Afp).tocsr()
        S_gpu = cc.block("schur")
        assert abs(S_gpu - S_twin).max() <= 1e-12 * abs(S_twin).max()
        par12 = dict(par)
        par12["solver rtol"] = 1e-13
        g12 = gpu_solve(sys_, par12, bench.BENCH_OPTIONS_CC)
        xd = spla.spsolve(sys_.A.tocsc(), sys_.b)
        assert g12["reason"] == 2
        assert rel(g12["x"], xd) <= 1e-8, rel(g12["x"], xd)
