import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle.amg import SAAMG, rigid_body_modes
from oracle.blockpc import BlockPC, SchurLower, SchurLowerCC, cc_from_matrices, krylov_solver
from oracle.krylov import gmres
from oracle.problems import swelling
for N in [int(a) for a in sys.argv[1:]]:
    s, par = swelling(3, N, "diagonal")
    B = rigid_body_modes(s.coords_s, 3)
    amg_s = lambda M: SAAMG(M, 3, B, theta=0.04, coarse_size=6000, dense_limit=8192)
    cheb_f = lambda M: SAAMG(M, 3, B, max_levels=1, cheby_degree=4, dense_limit=0)
    amg_f = lambda M: SAAMG(M, 3, B, theta=0.04, coarse_size=6000, dense_limit=8192)
    amg_p = lambda M: SAAMG(M, 1, None, coarse_size=6000, dense_limit=8192)
    cheb_p = lambda M: SAAMG(M, 1, None, max_levels=1, cheby_degree=4, dense_limit=0)
    dm, Sv = cc_from_matrices(s, par)
    def run(name, mkfp):
        pc = BlockPC(s, {"s": krylov_solver("preonly", amg_s), "fp": mkfp})
        t = time.time()
        r = gmres(lambda v: s.A @ v, s.b, pc, rtol=1e-8, atol=0.0, dtol=1e20, max_it=200, restart=200, pc_side="right")
        print("N %d %-28s its %d (%.0fs)" % (N, name, r.its, time.time() - t), flush=True)
    run("selfp + f cheb (bench)", lambda M: SchurLower(M, s.nf, s.np_, krylov_solver("preonly", cheb_f), krylov_solver("preonly", amg_p), "f"))
    run("cc + f V-cycle", lambda M: SchurLowerCC(M, s.nf, s.np_, krylov_solver("preonly", amg_f), krylov_solver("preonly", amg_p), krylov_solver("preonly", cheb_p), dm, Sv))
    run("cc + f cheb", lambda M: SchurLowerCC(M, s.nf, s.np_, krylov_solver("preonly", cheb_f), krylov_solver("preonly", amg_p), krylov_solver("preonly", cheb_p), dm, Sv))
    run("selfp + f V-cycle", lambda M: SchurLower(M, s.nf, s.np_, krylov_solver("preonly", amg_f), krylov_solver("preonly", amg_p), "f"))
