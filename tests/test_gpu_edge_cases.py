"""GPU: edge cases of the solve path -- zero right-hand side, a right-hand side that is already an eigen-direction,
iteration cap of one, repeated solves on one handle, AAR with order 0 (pure Richardson), empty BC list."""
import numpy as np
import pytest

from helpers import EXACT_OPTIONS, gpu_solve, rel

pytestmark = pytest.mark.gpu


def _setup(pc_type="diagonal", N=6):
    from oracle.problems import swelling
    return swelling(2, N, pc_type)


def test_zero_rhs_gmres_and_aar(gpu_ctx):
    sys_, par = _setup()
    sys_.b = np.zeros_like(sys_.b)
    g = gpu_solve(sys_, par, EXACT_OPTIONS)
    assert g["its"] == 0 and g["reason"] in (2, 3) and np.all(g["x"] == 0.0)
    a = gpu_solve(sys_, par, EXACT_OPTIONS, solver_type="aar")
    assert a["its"] == 0 and np.all(a["x"] == 0.0)


def test_repeated_solves_are_reproducible(gpu_ctx):
    """Deterministic reductions: two solves on the same handle give bit-identical results."""
    from poro_b200.lib.backend import DeviceVector
    sys_, par = _setup("diagonal 3-way")
    g = gpu_solve(sys_, par, EXACT_OPTIONS, return_objects=True)
    ksp = g["solver"].solver
    db = DeviceVector(sys_.b, ctx=gpu_ctx)
    x1, x2 = DeviceVector(n=sys_.n, ctx=gpu_ctx), DeviceVector(n=sys_.n, ctx=gpu_ctx)
    ksp.solve(db, x1)
    h1 = ksp.getConvergenceHistory()
    ksp.solve(db, x2)
    assert ksp.getConvergenceHistory() == h1
    assert np.array_equal(x1.numpy(), x2.numpy())


def test_aar_order_zero_is_richardson(gpu_ctx):
    """order 0 disables the Anderson steps (lib/AAR.py:94): x_{k+1} = x_k + omega M^-1 (b - A x_k)."""
    from oracle.aar import AAR
    from oracle.blockpc import BlockPC, exact_solvers
    sys_, par = _setup()
    par = dict(par)
    par.update({"AAR order": 0, "solver maxiter": 12, "solver rtol": 1e-30, "solver atol": 1e-30})
    o = AAR(0, 5, 1, 1, lambda v: sys_.A @ v, BlockPC(sys_, exact_solvers()), atol=1e-30, rtol=1e-30, maxiter=12)
    xo = o.solve(sys_.b)
    g = gpu_solve(sys_, par, EXACT_OPTIONS, solver_type="aar")
    assert g["its"] == 12 == o.it
    np.testing.assert_allclose(g["history"], o.history, rtol=1e-7)
    assert rel(g["x"], xo) <= 1e-8


def test_restart_shorter_than_maxit(gpu_ctx):
    """-global_ksp_gmres_restart overrides restart = maxiter (lib/Solver.py:99-101: setFromOptions comes last)."""
    from oracle.blockpc import BlockPC, exact_solvers
    from oracle.krylov import gmres
    sys_, par = _setup("undrained", 8)
    ro = gmres(lambda v: sys_.A @ v, sys_.b, BlockPC(sys_, exact_solvers()), rtol=par["solver rtol"], atol=par["solver atol"],
               dtol=1e20, max_it=par["solver maxiter"], restart=7, pc_side="right")
    g = gpu_solve(sys_, par, EXACT_OPTIONS + "\n-global_ksp_gmres_restart 7\n")
    assert g["reason"] == ro.reason
    assert abs(g["its"] - ro.its) <= max(2, ro.its // 10)
    assert np.linalg.norm(sys_.b - sys_.A @ g["x"]) <= 1.05 * max(par["solver rtol"] * np.linalg.norm(sys_.b), par["solver atol"])
