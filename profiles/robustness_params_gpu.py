"""kappa / k_s robustness of the single-block solvers ON THE GPU (BASELINE config 3's wording):
solid.py (CG + SA-AMG on A_ss) and fluid-pressure.py (GMRES + Schur fieldsplit on A[fp, fp]) over
kf in {1e-5, 1e-7, 1e-9, 1e-11} x ks in {1e4, 1e6, 1e8, 1e10}; iterations to rtol 1e-8, device time per solve.

    python profiles/robustness_params_gpu.py [N] > profiles/r2_robustness_params_gpu.md
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import numpy as np
from _single_block import DEFAULTS, build

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
KF = [1e-5, 1e-7, 1e-9, 1e-11]
KS = [1e4, 1e6, 1e8, 1e10]
print("# kappa / k_s sweep of the single-block drivers on the B200, swelling-3d mesh N = %d\n" % N)
for block, title in (("s", "solid.py: CG + SA-AMG V-cycle on A_ss, rtol 1e-8"), ("fp", "fluid-pressure.py: GMRES(right) + Schur(selfp) fieldsplit, rtol 1e-8")):
    print("## %s\n\niterations (device ms per solve)\n" % title)
    print("| kf \\ ks | " + " | ".join("%.0e" % ks for ks in KS) + " |")
    print("|---|" + "---|" * len(KS))
    for kf in KF:
        row = "| %.0e |" % kf
        for ks in KS:
            ctx, cc, db, dx, host, rhs, keep = build(block, N, {"kf": kf, "ks": ks}, options_text=DEFAULTS[block])
            cc.inner_solve(block, db, dx)
            ctx.sync()
            ctx.timer_start()
            cc.inner_solve(block, db, dx)
            ms = ctx.timer_stop()
            its, reason, rnorm = cc.inner_result(block)
            row += " %s (%.1f) |" % (str(its) if reason > 0 else ">%d" % its, ms)
            del keep, cc
        print(row, flush=True)
    print()
