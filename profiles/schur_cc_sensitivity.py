import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle.amg import SAAMG, rigid_body_modes
from oracle.blockpc import BlockPC, SchurLowerCC, cc_from_matrices, krylov_solver
from oracle.krylov import gmres
from hostfem.problems import swelling_assembler
from oracle.problems import swelling
def problem(N, scale):
    if scale == 1.0:
        return swelling(3, N, "diagonal")
    asm, par, loads = swelling_assembler(3, N, {"mu_f": 0.035 * scale})
    par = dict(par); par["mu_f"] = 0.035 * scale
    return asm.system("diagonal", par["t0"] + par["dt"], **loads), par
N = int(sys.argv[1])
for scale in [float(a) for a in sys.argv[2:]]:
    s, par = problem(N, scale)
    B = rigid_body_modes(s.coords_s, 3)
    dm, Sv = cc_from_matrices(s, par)
    def run(name, th_s=0.04, th_f=0.04, deg_v=4, deg_s=2, deg_f=2, th_p=0.08):
        amg_s = lambda M: SAAMG(M, 3, B, theta=th_s, coarse_size=6000, dense_limit=8192, cheby_degree=deg_s)
        amg_f = lambda M: SAAMG(M, 3, B, theta=th_f, coarse_size=6000, dense_limit=8192, cheby_degree=deg_f)
        amg_p = lambda M: SAAMG(M, 1, None, theta=th_p, coarse_size=6000, dense_limit=8192)
        cheb_p = lambda M: SAAMG(M, 1, None, max_levels=1, cheby_degree=deg_v, dense_limit=0)
        hs = {}
        def keep(k, mk):
            def f(M):
                hs[k] = mk(M); return hs[k]
            return f
        pc = BlockPC(s, {"s": krylov_solver("preonly", keep("s", amg_s)), "fp": lambda M: SchurLowerCC(M, s.nf, s.np_, krylov_solver("preonly", keep("f", amg_f)), krylov_solver("preonly", amg_p), krylov_solver("preonly", cheb_p), dm, Sv)})
        r = gmres(lambda v: s.A @ v, s.b, pc, rtol=1e-8, atol=0.0, dtol=1e20, max_it=200, restart=200, pc_side="right")
        print("N %d mu x%g %-34s its %3d  complexity s %.2f f %.2f" % (N, scale, name, r.its, hs["s"].complexity(), hs["f"].complexity()), flush=True)
    run("bench (cc)")
    run("f theta .02", th_f=0.02)
    run("f theta .08", th_f=0.08)
    run("s theta .02", th_s=0.02)
    run("s, f theta .02", th_s=0.02, th_f=0.02)
    run("visc Chebyshev(2)", deg_v=2)
    run("f smoother degree 1", deg_f=1)
    run("f smoother degree 3", deg_f=3)
    run("s smoother degree 3", deg_s=3)
    run("p theta .25", th_p=0.25)
