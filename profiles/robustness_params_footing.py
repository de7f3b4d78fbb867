import sys
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle.blockpc import BlockPC, exact_solvers
from oracle.krylov import gmres
from oracle.problems import footing
N=20
KF = [1e-5, 1e-7, 1e-9, 1e-11]; KS = [1e4, 1e6, 1e8, 1e10]
from hostfem.problems import footing_params
print("defaults", {k: footing_params()[k] for k in ("kf","ks","solver rtol","solver atol","solver maxiter")})
for pc_type in ("undrained", "undrained 3-way"):
    print("\nfooting.py, pc type = %s, inner = exact, N = %d: outer GMRES(right) iterations (true relative residual)" % (pc_type, N))
    print("%-10s" % "kf \\ ks" + "".join("%22.0e" % ks for ks in KS))
    for kf in KF:
        row = "%-10.0e" % kf
        for ks in KS:
            sys_, par = footing(N, pc_type, {"kf": kf, "ks": ks})
            pc = BlockPC(sys_, exact_solvers())
            r = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=par["solver rtol"], atol=par["solver atol"], dtol=1e20, max_it=500, restart=500, pc_side="right")
            true = np.linalg.norm(sys_.b - sys_.A @ r.x) / np.linalg.norm(sys_.b)
            row += "%12d (%.1e)" % (r.its, true) if r.reason > 0 else "%12s (%.1e)" % (">500", true)
        print(row, flush=True)
