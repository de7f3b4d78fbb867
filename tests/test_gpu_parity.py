"""GPU parity tests: the CUDA path (through ctypes and the C ABI) against the CPU oracle."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from helpers import AMG_OPTIONS, EXACT_OPTIONS, gpu_solve, rel

pytestmark = pytest.mark.gpu


def _oracle_exact(sys_, par, solver_type="gmres"):
    from oracle.aar import AAR
    from oracle.blockpc import BlockPC, exact_solvers
    from oracle.krylov import gmres
    A = lambda v: sys_.A @ v
    pc = BlockPC(sys_, exact_solvers())
    if solver_type == "aar":
        s = AAR(par["AAR order"], par["AAR p"], par["AAR omega"], par["AAR beta"], A, pc, atol=par["solver atol"],
                rtol=par["solver rtol"], maxiter=par["solver maxiter"])
        x = s.solve(sys_.b)
        return x, s.it, s.history
    r = gmres(A, sys_.b, pc, rtol=par["solver rtol"], atol=par["solver atol"], dtol=1e20, max_it=par["solver maxiter"],
              restart=par["solver maxiter"], pc_side="right")
    return r.x, r.its, r.history


@pytest.mark.parametrize("dim,N", [(2, 10), (3, 3)])
def test_spmv_matches_scipy(gpu_ctx, dim, N):
    import torch
    from oracle.problems import swelling
    from poro_b200.lib.backend import DeviceMatrix, DeviceVector
    sys_, _ = swelling(dim, N)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(sys_.n)
    dA = DeviceMatrix(sys_.A, gpu_ctx)
    dx, dy = DeviceVector(x, ctx=gpu_ctx), DeviceVector(n=sys_.n, ctx=gpu_ctx)
    dA.mult(dx, dy)
    gpu_ctx.sync()
    y = sys_.A @ x
    # fp64, different summation order only
    assert rel(dy.numpy(), y) <= 1e-14


@pytest.mark.parametrize("dim,N,pc_type", [(2, 10, "diagonal"), (2, 10, "diagonal 3-way"), (2, 10, "undrained"),
                                           (3, 3, "diagonal"), (3, 3, "diagonal 3-way")])
def test_gmres_exact_blocks_matches_oracle(gpu_ctx, dim, N, pc_type):
    """petsc-options-exact semantics: same iteration count (+-0) and solution as the oracle."""
    from oracle.problems import swelling
    sys_, par = swelling(dim, N, pc_type)
    xo, its_o, hist_o = _oracle_exact(sys_, par)
    g = gpu_solve(sys_, par, EXACT_OPTIONS)
    assert g["its"] == its_o
    assert g["reason"] in (2, 3)
    # 'undrained' adds k_s (div u, div v) with k_s = 1e6: the blocks are ill-conditioned; the dense inverse is refined
    # once with the sparse block (PCDense), which brings it to the accuracy of the oracle's splu
    hist_tol = 2e-3 if "undrained" in pc_type else 1e-6        # the tail (1e-8 of the start) is sensitive to the block solves' rounding
    np.testing.assert_allclose(g["history"], hist_o, rtol=hist_tol, atol=1e-14)
    assert rel(g["x"], xo) <= 1e-8


def test_gmres_exact_with_shuffled_numbering(gpu_ctx):
    """The library's field permutation: a randomly renumbered system gives the same answer."""
    from oracle.problems import swelling
    sys_, par = swelling(2, 10, "diagonal")
    xo, its_o, _ = _oracle_exact(sys_, par)
    perm = np.random.default_rng(1).permutation(sys_.n)
    g = gpu_solve(sys_, par, EXACT_OPTIONS, perm=perm)
    assert g["its"] == its_o
    assert rel(g["x"], xo) <= 1e-8


@pytest.mark.parametrize("pc_type", ["diagonal", "diagonal 3-way"])
def test_aar_exact_blocks_matches_oracle(gpu_ctx, pc_type):
    """swelling.py --solver-type aar (config 2): AAR with its quirks, Gram least squares."""
    from oracle.problems import swelling
    sys_, par = swelling(2, 10, pc_type)
    xo, its_o, hist_o = _oracle_exact(sys_, par, "aar")
    g = gpu_solve(sys_, par, EXACT_OPTIONS, solver_type="aar")
    xd = spla.spsolve(sys_.A.tocsc(), sys_.b)
    assert abs(g["its"] - its_o) <= max(2, its_o // 10)
    assert rel(g["x"], xd) <= 1e-6
    # the first iterations are pure Richardson and must agree to rounding
    np.testing.assert_allclose(g["history"][:4], hist_o[:4], rtol=1e-8)


@pytest.mark.parametrize("dim,N", [(2, 20), (3, 6)])
def test_gmres_amg_solution_and_counts(gpu_ctx, dim, N):
    """AMG-preconditioned config: solution within 1e-8 of a direct solve, true residual below
    tolerance, iteration count within 10% of the oracle running the SAME algorithm on CPU."""
    from oracle.amg import SAAMG, rigid_body_modes
    from oracle.blockpc import BlockPC, SchurLower, krylov_solver
    from oracle.krylov import gmres
    from oracle.problems import swelling
    sys_, par = swelling(dim, N, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-10, "solver atol": 0.0, "solver maxiter": 100})
    B = rigid_body_modes(sys_.coords_s, dim)
    amg_v = lambda M: SAAMG(M, dim, B)
    amg_p = lambda M: SAAMG(M, 1, None)
    mkfp = lambda M: SchurLower(M, sys_.nf, sys_.np_, krylov_solver("preonly", amg_v), krylov_solver("preonly", amg_p), "f")
    pc = BlockPC(sys_, {"s": krylov_solver("preonly", amg_v), "fp": mkfp})
    ro = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=1e-10, atol=0.0, dtol=1e20, max_it=100, restart=100, pc_side="right")
    g = gpu_solve(sys_, par, AMG_OPTIONS)
    xd = spla.spsolve(sys_.A.tocsc(), sys_.b)
    assert g["reason"] == 2
    assert abs(g["its"] - ro.its) <= max(1, int(round(0.1 * ro.its)))
    res = np.linalg.norm(sys_.b - sys_.A @ g["x"]) / np.linalg.norm(sys_.b)
    assert res <= 2e-10
    assert rel(g["x"], xd) <= 1e-8


INEXACT_CG_OPTIONS = """
-global_ksp_type fgmres
-s_ksp_type cg
-s_ksp_norm_type unpreconditioned
-s_ksp_atol 0.0
-s_ksp_rtol 1e-1
-s_pc_type hypre
-fp_ksp_type preonly
-fp_pc_fieldsplit_type schur
-fp_pc_fieldsplit_schur_fact_type lower
-fp_pc_fieldsplit_schur_precondition selfp
-fp_pc_fieldsplit_order fp
-fp_fieldsplit_0_ksp_type cg
-fp_fieldsplit_0_ksp_norm_type unpreconditioned
-fp_fieldsplit_0_ksp_rtol 1e-2
-fp_fieldsplit_0_ksp_atol 0.0
-fp_fieldsplit_0_pc_type hypre
-fp_fieldsplit_1_ksp_type cg
-fp_fieldsplit_1_ksp_rtol 1e-2
-fp_fieldsplit_1_ksp_atol 0.0
-fp_fieldsplit_1_ksp_max_it 10
-fp_fieldsplit_1_pc_type hypre
"""


@pytest.mark.parametrize("dim,N", [(2, 16), (3, 5)])
def test_fgmres_inner_cg_amg_matches_oracle(gpu_ctx, dim, N):
    """petsc-options-inexact style: inner CG + AMG to loose tolerances (a NONLINEAR preconditioner), outer FGMRES.
    Exercises the device CG (fused SpMV+dot, device-resident alpha), all three PETSc norm types' default, and FGMRES."""
    from oracle.amg import SAAMG, rigid_body_modes
    from oracle.blockpc import BlockPC, SchurLower, krylov_solver
    from oracle.krylov import gmres
    from oracle.problems import swelling
    sys_, par = swelling(dim, N, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-9, "solver atol": 0.0, "solver maxiter": 100, "solver type": "fgmres"})
    B = rigid_body_modes(sys_.coords_s, dim)
    amg_v = lambda M: SAAMG(M, dim, B)
    amg_p = lambda M: SAAMG(M, 1, None)
    k_s = krylov_solver("cg", amg_v, rtol=1e-1, atol=0.0, norm_type="unpreconditioned")
    k_f = krylov_solver("cg", amg_v, rtol=1e-2, atol=0.0, norm_type="unpreconditioned")
    k_p = krylov_solver("cg", amg_p, rtol=1e-2, atol=0.0, max_it=10)
    pc = BlockPC(sys_, {"s": k_s, "fp": lambda M: SchurLower(M, sys_.nf, sys_.np_, k_f, k_p, "f")})
    ro = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=1e-9, atol=0.0, dtol=1e20, max_it=100, restart=100, flexible=True)
    g = gpu_solve(sys_, par, INEXACT_CG_OPTIONS)
    assert g["reason"] == 2
    assert abs(g["its"] - ro.its) <= max(1, int(round(0.1 * ro.its)))
    # flexible GMRES monitors the true residual even though the preconditioner is nonlinear
    res = np.linalg.norm(sys_.b - sys_.A @ g["x"]) / np.linalg.norm(sys_.b)
    assert res <= 2e-9
    st = g["stats"]
    # inner iteration counts per application agree with the oracle's within 20 %
    for key, k in (("s", pc.k_s), ("fp_split0", pc.k_fp.k0), ("fp_split1", pc.k_fp.k1)):
        gpu_avg = st["its_" + key] / max(st["calls_" + key], 1)
        cpu_avg = k.total_its / max(k.calls, 1)
        assert abs(gpu_avg - cpu_avg) <= 0.2 * cpu_avg + 0.5, (key, gpu_avg, cpu_avg)


AMG_3WAY_OPTIONS = AMG_OPTIONS + """
-f_ksp_type preonly
-f_pc_type hypre
-p_ksp_type preonly
-p_pc_type hypre
-diff_ksp_type preonly
-diff_pc_type hypre
"""


def test_gmres_amg_three_way_matches_oracle(gpu_ctx):
    """'diagonal 3-way' with one V-cycle per inner solve (prefixes s_ f_ p_ diff_): counts within 10 % of the oracle."""
    from oracle.amg import SAAMG, rigid_body_modes
    from oracle.blockpc import BlockPC, krylov_solver
    from oracle.krylov import gmres
    from oracle.problems import swelling
    sys_, par = swelling(2, 12, "diagonal 3-way")
    par = dict(par)
    par.update({"solver rtol": 1e-9, "solver atol": 0.0, "solver maxiter": 200})
    B = rigid_body_modes(sys_.coords_s, 2)
    v = krylov_solver("preonly", lambda M: SAAMG(M, 2, B))
    p = krylov_solver("preonly", lambda M: SAAMG(M, 1, None))
    pc = BlockPC(sys_, {"s": v, "f": v, "p": p, "diff": p})
    ro = gmres(lambda w: sys_.A @ w, sys_.b, pc, rtol=1e-9, atol=0.0, dtol=1e20, max_it=200, restart=200, pc_side="right")
    g = gpu_solve(sys_, par, AMG_3WAY_OPTIONS)
    assert g["reason"] == 2
    assert abs(g["its"] - ro.its) <= max(1, int(round(0.1 * ro.its)))
    assert np.linalg.norm(sys_.b - sys_.A @ g["x"]) <= 2e-9 * np.linalg.norm(sys_.b)


def test_large_mesh_properties(gpu_ctx):
    """Size-independent properties on a mesh too large for the oracle's direct solves (3D N=20, 424 k DoFs):
    linearity of the device operator (BSR + diagonal-BSR + CSR split) against scipy, true residual of the
    solve, monotone residual history, and linearity of the whole solve in b (linear preconditioner)."""
    from oracle.problems import swelling
    from poro_b200.lib.backend import DeviceVector
    sys_, par = swelling(3, 20, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-8, "solver atol": 0.0, "solver maxiter": 100})
    g = gpu_solve(sys_, par, AMG_OPTIONS, return_objects=True)
    ksp = g["solver"].solver
    assert ksp.reason == 2
    x = g["x"]
    nb = np.linalg.norm(sys_.b)
    assert np.linalg.norm(sys_.b - sys_.A @ x) <= 1.5e-8 * nb
    hist = np.array(ksp.getConvergenceHistory())
    assert np.all(np.diff(hist) <= 1e-12 * hist[0])                  # GMRES residuals never increase
    # operator linearity through the solver's internal product
    rng = np.random.default_rng(5)
    u, v = rng.standard_normal(sys_.n), rng.standard_normal(sys_.n)
    du, dv, dw = DeviceVector(u, ctx=gpu_ctx), DeviceVector(v, ctx=gpu_ctx), DeviceVector(2.0 * u - 3.0 * v, ctx=gpu_ctx)
    yu, yv, yw = (DeviceVector(n=sys_.n, ctx=gpu_ctx) for _ in range(3))
    ksp.mult(du, yu); ksp.mult(dv, yv); ksp.mult(dw, yw)
    gpu_ctx.sync()
    assert rel(yu.numpy(), sys_.A @ u) <= 1e-13
    assert rel(yw.numpy(), 2.0 * yu.numpy() - 3.0 * yv.numpy()) <= 1e-13
    # the solve is linear in b: same iteration count, doubled solution
    db2, dx2 = DeviceVector(2.0 * sys_.b, ctx=gpu_ctx), DeviceVector(n=sys_.n, ctx=gpu_ctx)
    its1 = ksp.its
    ksp.solve(db2, dx2)
    assert ksp.its == its1
    assert rel(dx2.numpy(), 2.0 * x) <= 1e-7


@pytest.mark.parametrize("order", [1, 2])
def test_inner_anderson_acceleration_matches_oracle(gpu_ctx, order):
    """`inner accel order` > 0: AndersonAcceleration.get_next_vector on the preconditioner output
    (lib/Preconditioner.py:248-249, lib/AndersonAcceleration.py:19-78).  The accelerated PC is stateful and
    nonlinear, so only the first applications are compared one by one: PreconditionerCC.apply on a fixed sequence
    of inputs must reproduce the oracle's sequence."""
    from oracle.aar import AndersonAcceleration
    from oracle.blockpc import BlockPC, exact_solvers
    from oracle.problems import swelling
    from poro_b200.lib.backend import DeviceVector
    sys_, par = swelling(2, 6, "diagonal")
    g = gpu_solve(sys_, par, EXACT_OPTIONS, overrides={"inner accel order": order, "solver maxiter": 1}, return_objects=True)
    pcx = g["pc"].pc.getPythonContext()
    # the 1-iteration solve above already consumed PC applications; replay the same sequence on the oracle:
    # GMRES(right) with maxiter 1 applies the PC to v0 = b/||b|| and then to the update V y
    pco = BlockPC(sys_, exact_solvers(), anderson=AndersonAcceleration(order))
    v0 = sys_.b / np.linalg.norm(sys_.b)
    z0 = pco(v0)
    w = sys_.A @ z0
    h = v0 @ w
    w = w - h * v0
    hn = np.linalg.norm(w)
    den = np.hypot(h, hn)
    y0 = (h / den) * np.linalg.norm(sys_.b) / den
    pco(y0 * v0)
    rng = np.random.default_rng(11)
    for k in range(4):
        x = rng.standard_normal(sys_.n)
        dx, dy = DeviceVector(x, ctx=gpu_ctx), DeviceVector(n=sys_.n, ctx=gpu_ctx)
        pcx.apply(None, dx, dy)
        gpu_ctx.sync()
        yo = pco(x)
        assert rel(dy.numpy(), yo) <= 1e-6, (k, rel(dy.numpy(), yo))


def test_footing_undrained_matches_oracle(gpu_ctx):
    """BASELINE config 1 (footing.py, pc type 'undrained', the block lower-triangular 2-way PC with a non-zero
    fp<-s coupling): exact blocks, same iteration count and solution as the oracle."""
    from oracle.problems import footing
    sys_, par = footing(12)
    xo, its_o, hist_o = _oracle_exact(sys_, par)
    g = gpu_solve(sys_, par, EXACT_OPTIONS)
    assert g["its"] == its_o and g["reason"] in (2, 3)
    assert rel(g["x"], xo) <= 1e-6
    assert g["pc"].pc.getPythonContext().block_info("fps")[2] > 0        # the coupling block is really there


def test_reference_split_order_inexact_file(gpu_ctx):
    """options/petsc-options-inexact: the REFERENCE's fieldsplit order (split 0 = pressure, split 1 = fluid velocity
    with the selfp Schur complement S_f solved exactly, lib/Preconditioner.py:113-114, petsc-options-inexact:78-106),
    inner CG + AMG, outer FGMRES.  Same algorithm in the oracle: iteration counts within 10 %."""
    import os
    from oracle.amg import SAAMG, rigid_body_modes
    from oracle.blockpc import LU, BlockPC, SchurLower, krylov_solver
    from oracle.krylov import gmres
    from oracle.problems import swelling
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys_, par = swelling(2, 10, "diagonal")
    par = dict(par)
    par.update({"solver rtol": 1e-8, "solver atol": 0.0, "solver maxiter": 200, "solver type": "fgmres"})
    B = rigid_body_modes(sys_.coords_s, 2)
    k_s = krylov_solver("cg", lambda M: SAAMG(M, 2, B), rtol=1e-1, atol=0.0, norm_type="unpreconditioned")
    k_p = krylov_solver("cg", lambda M: SAAMG(M, 1, None), rtol=1e-4, atol=0.0, max_it=10)
    pc = BlockPC(sys_, {"s": k_s, "fp": lambda M: SchurLower(M, sys_.nf, sys_.np_, k_p, lambda S: LU(S), "p")})
    ro = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=1e-8, atol=0.0, dtol=1e20, max_it=200, restart=200, flexible=True)
    g = gpu_solve(sys_, par, open(os.path.join(root, "options", "petsc-options-inexact")).read())
    assert g["reason"] == 2
    assert abs(g["its"] - ro.its) <= max(1, int(round(0.1 * ro.its)))
    assert np.linalg.norm(sys_.b - sys_.A @ g["x"]) <= 2e-8 * np.linalg.norm(sys_.b)


def test_gmres_left_preconditioning_default(gpu_ctx):
    """No `-global_ksp_pc_side`: PETSc's default LEFT preconditioning, which monitors ||M^-1 r|| (what the reference
    runs when no option file is given)."""
    from oracle.blockpc import BlockPC, exact_solvers
    from oracle.krylov import gmres
    from oracle.problems import swelling
    sys_, par = swelling(2, 10, "diagonal")
    left = EXACT_OPTIONS.replace("-global_ksp_pc_side right\n", "")
    pc = BlockPC(sys_, exact_solvers())
    ro = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=par["solver rtol"], atol=par["solver atol"], dtol=1e20,
               max_it=par["solver maxiter"], restart=par["solver maxiter"], pc_side="left")
    g = gpu_solve(sys_, par, left)
    assert g["its"] == ro.its
    np.testing.assert_allclose(g["history"], ro.history, rtol=1e-6, atol=1e-14)
    assert abs(g["history"][0] - np.linalg.norm(pc(sys_.b))) <= 1e-9 * g["history"][0]    # preconditioned norm
    assert rel(g["x"], ro.x) <= 1e-8


def test_cgs2_refinement_option(gpu_ctx):
    """-global_ksp_gmres_cgs_refinement_type refine_always (CGS2): same iteration count, same solution."""
    from oracle.problems import swelling
    sys_, par = swelling(2, 10, "diagonal 3-way")
    a = gpu_solve(sys_, par, EXACT_OPTIONS)
    b = gpu_solve(sys_, par, EXACT_OPTIONS + "\n-global_ksp_gmres_cgs_refinement_type refine_always\n")
    assert a["its"] == b["its"]
    assert rel(a["x"], b["x"]) <= 1e-9
