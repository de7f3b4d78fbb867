"""Robustness of the block preconditioners in the permeability kf and the solid bulk modulus ks (BASELINE config 3's
wording; the reference's paper-scripts sweep only N).  CPU oracle (exact-block iteration counts are identical on the
GPU: tests/test_gpu_parity.py), swelling.py, one time step, rtol 1e-6 / atol 1e-8 / maxiter 500.

    python profiles/robustness_params.py [N]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.amg import SAAMG, rigid_body_modes
from oracle.blockpc import BlockPC, SchurLower, exact_solvers, krylov_solver
from oracle.krylov import gmres
from oracle.problems import swelling

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
KF = [1e-5, 1e-7, 1e-9, 1e-11]
KS = [1e4, 1e6, 1e8, 1e10]


def amg_solvers(sys_):
    B = rigid_body_modes(sys_.coords_s, sys_.dim)
    amg_v = lambda M: SAAMG(M, sys_.dim, B)
    amg_p = lambda M: SAAMG(M, 1, None)
    mkfp = lambda M: SchurLower(M, sys_.nf, sys_.np_, krylov_solver("preonly", amg_v), krylov_solver("preonly", amg_p), "f")
    pre = lambda mk: krylov_solver("preonly", mk)
    return {"s": pre(amg_v), "f": pre(amg_v), "p": pre(amg_p), "diff": pre(amg_p), "fp": mkfp}


for pc_type in ("diagonal", "diagonal 3-way", "undrained"):
    for mode in ("exact", "amg"):
        print("\npc type = %s, inner = %s, N = %d: outer GMRES(right) iterations (true relative residual)" % (pc_type, mode, N))
        print("%-10s" % "kf \\ ks" + "".join("%22.0e" % ks for ks in KS))
        for kf in KF:
            row = "%-10.0e" % kf
            for ks in KS:
                sys_, par = swelling(2, N, pc_type, {"kf": kf, "ks": ks})
                solvers = exact_solvers() if mode == "exact" else amg_solvers(sys_)
                pc = BlockPC(sys_, solvers)
                r = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=1e-6, atol=1e-8, dtol=1e20, max_it=500, restart=500, pc_side="right")
                true = np.linalg.norm(sys_.b - sys_.A @ r.x) / np.linalg.norm(sys_.b)
                row += "%12d (%.1e)" % (r.its, true) if r.reason > 0 else "%12s (%.1e)" % (">500", true)
            print(row, flush=True)
