"""Option sweep on the benchmark problem: assemble once, then for each option set build the
preconditioner, solve to rtol 1e-8 and report outer iterations / time per solve / phase profile.

    python profiles/option_sweep.py [N]
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import BENCH_OPTIONS, PHASE_NAMES
from hostfem.problems import swelling
from poro_b200.lib.backend import DeviceMatrix, DeviceVector, get_context
from poro_b200.lib.IndexSet import IndexSet
from poro_b200.lib.Parser import load_petsc_options
from poro_b200.lib.Preconditioner import Preconditioner
from poro_b200.lib.Solver import Solver

N = int(sys.argv[1]) if len(sys.argv) > 1 else 34
sets = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {
    "base": "",
    "f_levels1": "-fp_fieldsplit_0_pc_amg_max_levels 1",
    "f_levels2": "-fp_fieldsplit_0_pc_amg_max_levels 2",
    "f_levels1_deg3": "-fp_fieldsplit_0_pc_amg_max_levels 1\n-fp_fieldsplit_0_pc_amg_cheby_degree 3",
    "s_levels3": "-s_pc_amg_max_levels 3",
    "f_jacobi": "-fp_fieldsplit_0_pc_type jacobi",
}
s, par = swelling(3, N, "diagonal")
par = dict(par); par.update({"solver rtol": 1e-8, "solver atol": 0.0, "solver maxiter": 100})
ctx = get_context(0)
dA, dP = DeviceMatrix(s.A, ctx), DeviceMatrix(s.P, ctx)
db = DeviceVector(s.b, ctx=ctx); dx = DeviceVector(n=s.n, ctx=ctx)
for name, extra in sets.items():
    ctx.clear_options()
    load_petsc_options(ctx, BENCH_OPTIONS + "\n" + extra + "\n", is_text=True)
    imap = IndexSet(s.is_s, s.is_f, s.is_p, two_way=True, block_dim=3, coords_s=s.coords_s, coords_p=s.coords_p)
    t0 = time.perf_counter()
    pcw = Preconditioner(imap, dA, dP, None, par, s.bcs_sub_pressure); pc = pcw.get_pc()
    solver = Solver(dA, db, pc, par, imap); solver.create_solver(dA, db, pc)
    ctx.sync(); tset = time.perf_counter() - t0
    ksp = solver.solver
    ksp.solve(db, dx)
    ctx.profile(1)
    ctx.sync(); t0 = time.perf_counter()
    ksp.solve(db, dx)
    ctx.sync(); dt = time.perf_counter() - t0
    ph = ctx.profile(0)
    x = dx.numpy()
    res = np.linalg.norm(s.b - s.A @ x) / np.linalg.norm(s.b)
    pcx = pc.getPythonContext()
    print("%-16s its %3d reason %d  %.1f ms/solve  %.2f ms/it  true res %.2e  setup %.1fs  levels s=%s f=%s" % (
        name, ksp.its, ksp.reason, 1e3 * dt, 1e3 * dt / max(ksp.its, 1), res, tset,
        [r for r, _ in pcx.amg_info("s")], [r for r, _ in pcx.amg_info("fp0")]), flush=True)
    print("    ", {PHASE_NAMES.get(k, k): round(v[0], 1) for k, v in sorted(ph.items()) if k < 8 or k >= 32}, flush=True)
    del solver, ksp, pc, pcw
