"""Shared helpers for the parity tests: build a problem with the oracle's assembler, run it
through the product path (ctypes -> libporo.so) and through the oracle."""
from __future__ import annotations

import numpy as np

EXACT_OPTIONS = """
-global_ksp_type gmres
-global_ksp_pc_side right
-s_ksp_type preonly
-s_pc_type lu
-f_ksp_type preonly
-f_pc_type lu
-p_ksp_type preonly
-p_pc_type lu
-diff_ksp_type preonly
-diff_pc_type lu
-fp_ksp_type preonly
-fp_pc_type lu
"""

# linear AMG preconditioner, pressure Schur complement (the configuration bench.py times)
AMG_OPTIONS = """
-global_ksp_type gmres
-global_ksp_pc_side right
-s_ksp_type preonly
-s_pc_type hypre
-fp_ksp_type preonly
-fp_pc_fieldsplit_type schur
-fp_pc_fieldsplit_schur_fact_type lower
-fp_pc_fieldsplit_schur_precondition selfp
-fp_pc_fieldsplit_order fp
-fp_fieldsplit_0_ksp_type preonly
-fp_fieldsplit_0_pc_type hypre
-fp_fieldsplit_1_ksp_type preonly
-fp_fieldsplit_1_pc_type hypre
"""


def gpu_solve(sys_, par, options_text, perm=None, solver_type=None, overrides=None, return_objects=False):
    """Solve sys_ on the GPU through the reference-shaped classes.  perm: optional random
    permutation of the global numbering (exercises the field permutation in the library)."""
    import torch
    from poro_b200.lib.backend import DeviceMatrix, DeviceVector, get_context
    from poro_b200.lib.IndexSet import IndexSet
    from poro_b200.lib.Parser import load_petsc_options
    from poro_b200.lib.Preconditioner import Preconditioner
    from poro_b200.lib.Solver import Solver

    ctx = get_context(0)
    ctx.clear_options()
    load_petsc_options(ctx, options_text, is_text=True)
    par = dict(par)
    if solver_type:
        par["solver type"] = solver_type
    if overrides:
        par.update(overrides)
    A, P, Pd, b = sys_.A, sys_.P, sys_.P_diff, sys_.b
    is_s, is_f, is_p = sys_.is_s, sys_.is_f, sys_.is_p
    coords_s, coords_p = sys_.coords_s, sys_.coords_p
    if perm is not None:
        # new index of old dof i is perm[i]
        inv = np.argsort(perm)
        A = A[inv][:, inv].tocsr()
        P = P[inv][:, inv].tocsr()
        Pd = Pd[inv][:, inv].tocsr() if Pd is not None else None
        b = b[inv]
        is_s, is_f, is_p = perm[is_s], perm[is_f], perm[is_p]
    two_way = "3-way" not in par["pc type"]
    imap = IndexSet(is_s, is_f, is_p, two_way=two_way, block_dim=sys_.dim if perm is None else 0,
                    coords_s=coords_s if perm is None else None, coords_p=coords_p if perm is None else None)
    dA, dP = DeviceMatrix(A, ctx), DeviceMatrix(P, ctx)
    dPd = DeviceMatrix(Pd, ctx) if Pd is not None else None
    db = DeviceVector(b, ctx=ctx)
    pcw = Preconditioner(imap, dA, dP, dPd, par, sys_.bcs_sub_pressure)
    pc = pcw.get_pc()
    solver = Solver(dA, db, pc, par, imap)
    solver.create_solver(dA, db, pc)
    x = DeviceVector(n=len(b), ctx=ctx)
    solver.set_up()
    solver.solve(db.vec(), x.vec())
    xs = x.numpy()
    if perm is not None:
        xs = xs[perm]
    out = dict(x=xs, its=solver.getIterationNumber(), solver=solver, pc=pcw)
    if not return_objects:
        inner = solver.solver
        out["reason"] = getattr(inner, "reason", None)
        out["history"] = inner.getConvergenceHistory() if hasattr(inner, "getConvergenceHistory") else inner.residual_history()
        out["stats"] = pc.getPythonContext().stats()
    return out


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
