"""CPU: the host assembler (oracle/fem.py) against manufactured checks.  No reference fixtures
exist (the reference ships no tests and cannot run here: PARITY UNPINNED), so the assembler is
pinned by exact integrals of polynomial fields and by the dof counts of SURVEY.md 8(a)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle.fem import PoroAssembler, simplex_quadrature, unit_cube_mesh, unit_square_mesh
from oracle.problems import swelling, swelling_params


@pytest.mark.parametrize("d", [1, 2, 3])
def test_quadrature_integrates_monomials(d):
    from math import factorial
    X, W = simplex_quadrature(d, 4)
    assert abs(W.sum() - 1.0 / factorial(d)) < 1e-14
    # int x_0^a x_1^b ... = a! b! ... / (a+b+...+d)!
    for powers in [(2,) + (0,) * (d - 1), (1,) * d, (3,) + (1,) * (d - 1)]:
        exact = np.prod([factorial(p) for p in powers]) / factorial(sum(powers) + d)
        val = (W * np.prod(X ** np.array(powers), axis=1)).sum()
        assert abs(val - exact) < 1e-14


@pytest.mark.parametrize("dim,N,n_expected", [(2, 10, 1885), (2, 20, 7165), (3, 2, 3 * 125 * 2 + 27), (3, 4, 4499)])
def test_dof_counts(dim, N, n_expected):
    # ns = nf = d (2N+1)^d, np = (N+1)^d  (1885 for swelling.py's default mesh)
    sys_, _ = swelling(dim, N)
    assert sys_.n == n_expected
    assert sys_.ns == sys_.nf == dim * (2 * N + 1) ** dim
    assert sys_.np_ == (N + 1) ** dim


@pytest.mark.parametrize("dim", [2, 3])
def test_forms_on_polynomial_fields(dim):
    L = 2.0
    mesh = unit_square_mesh(4, L) if dim == 2 else unit_cube_mesh(2, L)
    asm = PoroAssembler(mesh, swelling_params(dim))
    d = dim
    M = sp.csr_matrix((asm.M, asm.pat22.cols, asm.pat22.indptr), shape=asm.pat22.shape)
    assert abs(M.sum() - L ** d) < 1e-12                         # volume
    mk = lambda data, bc=d: sp.bsr_matrix((data, asm.pat22.cols, asm.pat22.indptr), shape=(asm.n2 * d, asm.n2 * bc)).tocsr()
    Ke, Kdd = mk(asm._eps_blocks()), mk(asm._divdiv_blocks())
    X = asm.p2_coords
    # rigid body modes are in the kernel of (eps, eps)
    tr = np.zeros((asm.n2, d)); tr[:, 0] = 1
    rot = np.zeros((asm.n2, d)); rot[:, 0], rot[:, 1] = -X[:, 1], X[:, 0]
    assert np.abs(Ke @ tr.ravel()).max() < 1e-12 and np.abs(Ke @ rot.ravel()).max() < 1e-11
    # u = (x^2, x y, 0): eps:eps = 4x^2 + y^2/2 + x^2, div u = 3x   (P2 represents u exactly)
    u = np.zeros((asm.n2, d)); u[:, 0], u[:, 1] = X[:, 0] ** 2, X[:, 0] * X[:, 1]
    u = u.ravel()
    vol_rest = L ** (d - 2)
    ix2 = L ** 3 / 3 * L * vol_rest                                # int x^2 over the box
    assert abs(u @ Ke @ u - (5 * ix2 + 0.5 * ix2)) < 1e-9
    assert abs(u @ Kdd @ u - 9 * ix2) < 1e-9
    Dv = sp.bsr_matrix((asm._div_blocks(), asm.pat21.cols, asm.pat21.indptr), shape=(asm.n2 * d, asm.n1), blocksize=(d, 1)).tocsr()
    assert abs((Dv.T @ u).sum() - 3 * (L ** 2 / 2) * L ** (d - 1)) < 1e-10    # int div u = int 3x
    # P1 Laplacian of a linear function: energy = |grad|^2 * volume
    Kp = sp.csr_matrix((asm.Kp, asm.pat11.cols, asm.pat11.indptr), shape=asm.pat11.shape)
    p = mesh.coords[:, 0] * 2 + mesh.coords[:, 1]
    assert abs(p @ Kp @ p - 5 * L ** d) < 1e-10


@pytest.mark.parametrize("dim,N", [(2, 6), (3, 2)])
def test_bc_rows_and_rhs(dim, N):
    sys_, par = swelling(dim, N)
    A, P, b = sys_.A, sys_.P, sys_.b
    # DirichletBC.apply: row zeroed, unit diagonal, rhs zero; columns kept (Poromechanics.py:76-83)
    bc_rows = np.flatnonzero((np.diff(A.indptr) == 1) & (A.diagonal() == 1.0))
    assert len(bc_rows) > 0
    assert np.all(b[bc_rows] == 0.0)
    assert np.all(P.diagonal()[bc_rows] == 1.0)
    assert abs(A - A.T).max() > 0                                  # columns kept -> non-symmetric
    # total traction on the solid: -35.29 * n over the Neumann sides (swelling.py:35-40), minus BC rows
    c = -1e3 * 0.9 * (1 - np.exp(-0.04))
    assert abs(c + 35.29) < 5e-3
    x = spla.spsolve(A.tocsc(), b)
    assert np.linalg.norm(A @ x - b) <= 1e-10 * np.linalg.norm(b)


def test_pc_variants_differ_where_expected():
    sd, _ = swelling(2, 4, "diagonal")
    su, _ = swelling(2, 4, "undrained")
    s3, _ = swelling(2, 4, "diagonal 3-way")
    assert sd.P_diff is None and s3.P_diff is not None
    s, fp = sd.is_s, sd.is_fp
    # diagonal: no solid coupling in the fp rows (Assembler.py:149-160); undrained: none in the s rows
    assert abs(sd.P[fp][:, s]).sum() == 0
    assert abs(su.P[s][:, fp]).sum() == 0 and abs(su.P[fp][:, s]).sum() > 0
    # A is the same for every pc type
    assert abs(sd.A - su.A).max() == 0
    # pressure BCs only on P_diff
    p = s3.is_p
    dd = s3.P_diff[p][:, p]
    assert len(s3.bcs_sub_pressure) > 0
    assert np.all(dd.diagonal()[s3.bcs_sub_pressure] == 1.0)


def test_footing_problem():
    """footing.py (BASELINE config 1) on the uniform mesh: load resultant, BC sets, undrained PC converges."""
    from oracle.blockpc import BlockPC, exact_solvers
    from oracle.krylov import gmres
    from oracle.problems import footing
    s, par = footing(10)
    assert s.n == 1885 and s.pc_type == "undrained" and par["solver atol"] == 1e-4
    # vertical load -1e4 over the foot (width 32), nothing horizontal
    assert abs(s.b[1:s.ns:2].sum() + 32 * 1e4) < 1e-6 and abs(s.b[0:s.ns:2]).sum() == 0.0
    assert abs(s.b[s.ns:]).sum() == 0.0
    r = gmres(lambda v: s.A @ v, s.b, BlockPC(s, exact_solvers()), rtol=par["solver rtol"], atol=par["solver atol"], dtol=1e20,
              max_it=par["solver maxiter"], restart=par["solver maxiter"], pc_side="right")
    assert r.reason in (2, 3) and r.its == 26


def test_time_levels_and_rhs_of_later_steps():
    """examples/_driver.py: the time levels of AbstractPhysics.solve and the right-hand side of a later step
    (only the loads depend on t, lib/Assembler.py:267-268)."""
    import importlib.util
    import math
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("example_driver", os.path.join(root, "examples", "_driver.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    assert drv.time_levels(0.0, 0.1, 0.1) == [0.1]
    assert np.allclose(drv.time_levels(0.0, 0.3, 0.1), [0.1, 0.2, 0.3])
    from hostfem import problems
    sys_, par = problems.swelling(2, 4, "diagonal")
    rhs_at = sys_.meta["rhs_at"]
    assert np.array_equal(rhs_at(par["t0"] + par["dt"]), sys_.b)
    mag = lambda t: 1 - math.exp(-(t ** 2) / 0.25)
    np.testing.assert_allclose(rhs_at(0.2), sys_.b * (mag(0.2) / mag(0.1)), rtol=1e-12, atol=0)
    sysf, parf = problems.footing(4)
    np.testing.assert_allclose(sysf.meta["rhs_at"](0.3), 3.0 * sysf.b, rtol=1e-12, atol=0)     # load = min(t, 1) * 1e5
