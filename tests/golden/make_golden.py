"""Generates the golden fixtures in this directory from the CPU oracle.

    python tests/golden/make_golden.py

The reference itself cannot be imported in the build container (petsc4py / dolfin / mpi4py are
absent), so these are ORACLE outputs, not reference outputs (parity unpinned).
Each .npz holds one assembled system in CSR form (A, P, P_diff), b, the index sets, the
pressure-BC map, and the oracle's results for the exact-block configuration
(petsc-options-exact semantics): GMRES(right) solution, iteration count, residual history, and the
AAR iteration count / solution.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.aar import AAR                                   # noqa: E402
from oracle.blockpc import BlockPC, exact_solvers           # noqa: E402
from oracle.krylov import gmres                              # noqa: E402
from oracle.problems import footing, swelling                # noqa: E402

CASES = [("swelling2d_N4_diagonal", 2, 4, "diagonal"), ("swelling2d_N4_diagonal3way", 2, 4, "diagonal 3-way"),
         ("swelling3d_N2_diagonal", 3, 2, "diagonal"), ("swelling2d_N6_undrained", 2, 6, "undrained"),
         ("footing_N8_undrained", 0, 8, "undrained")]            # dim 0 = footing.py (BASELINE config 1)


def csr_pack(prefix, M, out):
    if M is None:
        return
    out[prefix + "_indptr"], out[prefix + "_indices"], out[prefix + "_data"] = M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data
    out[prefix + "_shape"] = np.array(M.shape)


def main():
    for name, dim, N, pct in CASES:
        s, par = footing(N, pct) if dim == 0 else swelling(dim, N, pct)
        A = lambda v: s.A @ v
        r = gmres(A, s.b, BlockPC(s, exact_solvers()), rtol=par["solver rtol"], atol=par["solver atol"], dtol=1e20,
                  max_it=par["solver maxiter"], restart=par["solver maxiter"], pc_side="right")
        a = AAR(par["AAR order"], par["AAR p"], par["AAR omega"], par["AAR beta"], A, BlockPC(s, exact_solvers()),
                atol=par["solver atol"], rtol=par["solver rtol"], maxiter=par["solver maxiter"])
        xa = a.solve(s.b)
        out = dict(dim=s.dim, N=N, pc_type=pct, b=s.b, is_s=s.is_s, is_f=s.is_f, is_p=s.is_p,
                   bcs_sub_pressure=s.bcs_sub_pressure, coords_s=s.coords_s, coords_p=s.coords_p,
                   gmres_x=r.x, gmres_its=r.its, gmres_history=np.array(r.history), gmres_reason=r.reason,
                   aar_x=xa, aar_its=a.it, aar_history=np.array(a.history),
                   rtol=par["solver rtol"], atol=par["solver atol"], maxiter=par["solver maxiter"])
        csr_pack("A", s.A, out)
        csr_pack("P", s.P, out)
        csr_pack("Pd", s.P_diff, out)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "n =", s.n, "gmres its", r.its, "aar its", a.it)


if __name__ == "__main__":
    main()
