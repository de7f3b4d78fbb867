"""Shared driver logic of the example scripts: the reference's L3-L5 layers (drivers, Parser, AbstractPhysics time
loop, Poromechanics.solve_time_step) restated over hostfem (assembly on the host) and poro_b200 (solve on the GPU).

    AbstractPhysics.solve           lib/AbstractPhysics.py:59-82   time loop, "Solved time" line
    Poromechanics.create_solver     lib/Poromechanics.py:58-68     Preconditioner(...).get_pc(); Solver(...).create_solver()
    Poromechanics.solve_time_step   lib/Poromechanics.py:70-98     RHS at t, solver created at the first step and reused
"""
from __future__ import annotations

import os
import sys
from time import perf_counter as time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def time_levels(t0: float, tf: float, dt: float):
    """The time levels AbstractPhysics.solve visits (lib/AbstractPhysics.py:73-75: `while t < tf: t += dt`)."""
    out, t = [], t0
    while t < tf - 1e-12:
        t += dt
        out.append(t)
    return out


def run(problem: str, default_N: int, argv=None):
    from hostfem import problems
    from poro_b200.lib.backend import DeviceMatrix, DeviceVector, get_context
    from poro_b200.lib.IndexSet import IndexSet
    from poro_b200.lib.Parser import Parser
    from poro_b200.lib.Preconditioner import Preconditioner
    from poro_b200.lib.Printing import parprint
    from poro_b200.lib.Solver import Solver

    ctx = get_context()
    parser = Parser(argv, ctx=ctx)                       # also loads --petsc-options FILE into the options database
    N = parser.options.N or default_N
    overrides = dict(parser.options_dict)
    overrides.pop("N", None)
    t0_asm = time()
    if problem == "footing":
        sys_, par = problems.footing(N, overrides.get("pc type"), overrides)
        dim = 2
    else:
        dim = 3 if problem == "swelling-3d" else 2
        sys_, par = problems.swelling(dim, N, overrides.get("pc type"), overrides)
    parprint("---- Problem dofs={}, h={}, solving with {} procs".format(sys_.n, (64.0 if problem == "footing" else 1e-2) / N, 1))
    parprint("---- [Assembler] Assembly A, P time = {}s".format(time() - t0_asm))
    two_way = "3-way" not in par["pc type"]
    index_map = IndexSet(sys_.is_s, sys_.is_f, sys_.is_p, two_way=two_way, block_dim=dim, coords_s=sys_.coords_s,
                         coords_p=sys_.coords_p)
    A, P = DeviceMatrix(sys_.A, ctx), DeviceMatrix(sys_.P, ctx)
    P_diff = DeviceMatrix(sys_.P_diff, ctx) if sys_.P_diff is not None else None
    b = DeviceVector(sys_.b, ctx=ctx)
    sol = DeviceVector(n=sys_.n, ctx=ctx)
    # Poromechanics.create_solver (first time step only)
    pcw = Preconditioner(index_map, A, P, P_diff, par, sys_.bcs_sub_pressure)
    pc = pcw.get_pc()
    solver = Solver(A, b, pc, par, index_map)
    solver.create_solver(A, b, pc)
    # AbstractPhysics.solve: every shipped driver runs exactly one step (t0 = 0, tf = dt = 0.1); with
    # --time-final the loop continues: the right-hand side depends on t only through the loads
    # (lib/Assembler.py:267-268), matrices, preconditioner and solver are those of the first step
    t, tf, dt = par["t0"], par.get("tf", par["dt"]), par["dt"]
    t0_sim = time()
    current = time()
    b_host = sys_.b
    for step, t in enumerate(time_levels(t, tf, dt)):
        if step > 0:
            b_host = sys_.meta["rhs_at"](t)
            b = DeviceVector(b_host, ctx=ctx)
        solver.set_up()
        solver.solve(b.vec(), sol.vec())
        its = solver.getIterationNumber()
        parprint("-------- Solved time t={:.2f}. {} iterations in {:.2f}s".format(t, its, time() - current))
        current = time()
    parprint("Total simulation time = {}s\n".format(time() - t0_sim))
    pcw.print_timings()
    solver.print_timings()
    x = sol.numpy()
    res = np.linalg.norm(b_host - sys_.A @ x) / max(np.linalg.norm(b_host), 1e-300)
    parprint("true relative residual = {:.3e}".format(res))
    return dict(its=its, x=x, residual=res, n=sys_.n)
