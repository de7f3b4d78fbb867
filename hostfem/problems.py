"""The reference drivers' problem definitions, restated (input generation; stands in for the drivers).

swelling.py:11-106, swelling-3d.py:11-111.  Each function returns (PoroSystem, parameters)
where `parameters` carries the same keys the reference's `parameters` dict feeds to
Solver / Preconditioner (solver type, tolerances, pc type, inner *, AAR *).
"""
from __future__ import annotations

import math

import numpy as np

from .fem import PoroAssembler, unit_cube_mesh, unit_square_mesh

_SOLVER_KEYS_2D = {
    "solver atol": 1e-8, "solver rtol": 1e-6, "solver maxiter": 500, "solver monitor": False,
    "solver type": "gmres", "pc type": "diagonal", "inner ksp type": "gmres", "inner pc type": "hypre",
    "inner atol": 0, "inner rtol": 1e-6, "inner maxiter": 1000, "inner monitor": False,
    "inner accel order": 0, "AAR order": 10, "AAR p": 5, "AAR omega": 1, "AAR beta": 1,
}


def swelling_params(dim: int) -> dict:
    """Material + solver parameters of swelling.py:43-79 (2D) / swelling-3d.py:43-80 (3D)."""
    p = {"mu_f": 0.035, "rhof": 1e3, "rhos": 1e3, "phi0": 0.1, "mu_s": 4000, "lmbda": 700,
         "ks": 1e6 if dim == 2 else 1e8, "kf": 1e-7, "dt": 0.1, "t0": 0.0, "tf": 0.1,
         "fe degree solid": 2, "fe degree fluid": 2, "fe degree pressure": 1,
         "betas": -0.5, "betaf": 0.0, "betap": 1.0}
    p.update(_SOLVER_KEYS_2D)
    if dim == 3:
        p.update({"solver maxiter": 100, "inner maxiter": 100})
    return p


def _traction(scale):
    # swelling.py:35-40: Constant(-1e3 * scale * (1 - exp(-(t**2) / 0.25))) * FacetNormal(mesh)
    return lambda t: -1e3 * scale * (1 - math.exp(-(t ** 2) / 0.25))


def swelling_assembler(dim: int, N: int, overrides: dict | None = None):
    par = swelling_params(dim)
    if overrides:
        par.update(overrides)
    side_length = 1e-2
    if dim == 2:
        mesh = unit_square_mesh(N, side_length)
        asm = PoroAssembler(mesh, par)
        # swelling.py:94-101
        asm.set_bcs(bcs_s=[("x0", 0), ("y0", 1)],
                    bcs_f=[("y1", None), ("y0", None)],
                    bcs_p=["x0", "y1", "x1"])
        loads = dict(neumann_solid=["y1", "x1"], neumann_fluid=["x0"])      # swelling.py:23-24
    else:
        mesh = unit_cube_mesh(N, side_length)
        asm = PoroAssembler(mesh, par)
        # swelling-3d.py:95-106
        asm.set_bcs(bcs_s=[("x0", 0), ("y0", 1), ("z0", 2)],
                    bcs_f=[("z0", None), ("z1", None)],
                    bcs_p=["x0", "x1", "y0", "y1", "z1"])
        loads = dict(neumann_solid=["x1", "y1", "z1"], neumann_fluid=["x0", "y0"])  # swelling-3d.py:22-23
    loads.update(fs_sur=_traction(0.9), ff_sur=_traction(0.1))
    return asm, par, loads


def swelling(dim: int, N: int = 10, pc_type: str | None = None, overrides: dict | None = None):
    """One time step (t = dt = 0.1, AbstractPhysics.py:73-81) of the swelling test."""
    asm, par, loads = swelling_assembler(dim, N, overrides)
    if pc_type is not None:
        par["pc type"] = pc_type
    t = par["t0"] + par["dt"]
    sys_ = asm.system(par["pc type"], t, **loads)
    # the right-hand side of a later time step (lib/Assembler.py:267-268: only the tractions depend on t)
    sys_.meta.update(dict(problem="swelling-%dd" % dim, N=N, rhs_at=lambda tt: asm.rhs(tt, **loads)))
    return sys_, par


def footing_params() -> dict:
    """Material + solver parameters of footing.py:42-84."""
    E, nu = 3e4, 0.2
    p = {"mu_f": 1e-3, "rhof": 1e3, "rhos": 500, "phi0": 1e-3, "mu_s": E / (2 * (1 + nu)),
         "lmbda": E * nu / ((1 + nu) * (1 - 2 * nu)), "ks": 1e6, "kf": 1e-7, "dt": 0.1, "t0": 0.0, "tf": 0.1,
         "fe degree solid": 2, "fe degree fluid": 2, "fe degree pressure": 1, "betas": -0.5, "betaf": 0.0, "betap": 1.0}
    p.update(_SOLVER_KEYS_2D)
    p.update({"solver rtol": 1e-6, "solver atol": 1e-4, "solver maxiter": 500, "pc type": "undrained"})
    return p


def footing(N: int = 10, pc_type: str | None = None, overrides: dict | None = None):
    """footing.py (BASELINE config 1), one time step.  Square of side 64, solid clamped at the bottom, fluid no-slip
    under the foot (top, |x - 32| < 16), pressure BC on the rest of the boundary, vertical load (0, -1e4) under
    |x - 32| < 16 on the top side, P1-interpolated (footing.py:36-39, 93-110).

    Deviation (documented): the reference refines the top-middle cells twice with dolfin's Plaza `refine`
    (lib/MeshCreation.py:53-104), which cannot be reproduced without dolfin; this generator uses the uniform
    N x N 'right'-diagonal mesh, and marks Dirichlet dofs nodewise instead of facetwise."""
    par = footing_params()
    if overrides:
        par.update(overrides)
    if pc_type is not None:
        par["pc type"] = pc_type
    length = 64.0
    mesh = unit_square_mesh(N, length)
    asm = PoroAssembler(mesh, par)
    tol = 1e-10 * length
    on_boundary = lambda X: (np.abs(X[:, 0]) < tol) | (np.abs(X[:, 0] - length) < tol) | (np.abs(X[:, 1]) < tol) | (np.abs(X[:, 1] - length) < tol)
    foot = lambda X: (np.abs(X[:, 1] - length) < tol) & (np.abs(X[:, 0] - length / 2) < length / 4)
    foot_not = lambda X: on_boundary(X) & ~foot(X)
    asm.set_bcs(bcs_s=[("y0", None)], bcs_f=[(foot, None)], bcs_p=[foot_not])          # footing.py:97-108
    t = par["t0"] + par["dt"]

    def rhs_at(tt):
        val = min(tt, 1.0) * 1e5
        load = lambda X: np.stack([np.zeros(len(X)), np.where(np.abs(X[:, 0] - length / 2) < length / 4, -val, 0.0)], 1)
        return asm.rhs_vector_load_2d("y1", load, field_off=0)                           # Neumann solid = TOP (footing.py:22)

    sys_ = asm.system(par["pc type"], t, b=rhs_at(t))
    sys_.meta.update(dict(problem="footing", N=N, rhs_at=rhs_at))
    return sys_, par
