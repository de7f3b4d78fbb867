import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, scipy.sparse as sp
from oracle.amg import SAAMG, rigid_body_modes
from oracle.blockpc import BlockPC, krylov_solver, submatrix
from oracle.krylov import gmres
from hostfem.problems import swelling_assembler
def run(dim, N, variant, fkind, scale, rtol=1e-8):
    asm, par, loads = swelling_assembler(dim, N, {"mu_f": 0.035 * scale})
    sys_ = asm.system("diagonal", par["t0"] + par["dt"], **loads)
    c = asm._coeffs()
    phi0, mu_f, d = c["phi0"], c["mu_f"], dim
    mf = c["rhof"] * c["idt"] * phi0; drag = phi0 ** 2 * c["ikf"]
    cM = (mf + (1.0 + c["betaf"]) * drag)
    Mv_diag = asm._to_csr("22", asm._mass_blocks()).diagonal()
    Mp = asm._to_csr("11", asm.Mp)
    nf, npp = sys_.nf, sys_.np_
    B = rigid_body_modes(sys_.coords_s, dim)
    class Schur:
        def __init__(self, M):
            f, p = np.arange(nf), nf + np.arange(npp)
            self.A00, self.A01, self.A10, self.A11 = submatrix(M, f, f), submatrix(M, f, p), submatrix(M, p, f), submatrix(M, p, p)
            self.k0 = SAAMG(self.A00, dim, B, max_levels=1, cheby_degree=4, dense_limit=0) if fkind == "cheb" else SAAMG(self.A00, dim, B, theta=0.04, coarse_size=6000, dense_limit=8192)
            bc = asm.bc_f.ravel()
            if variant == "selfp":
                S = self.A11 - self.A10 @ sp.diags(1.0 / self.A00.diagonal()) @ self.A01
                self.kS = SAAMG(S.tocsr(), 1, None, coarse_size=6000, dense_limit=8192); self.kV = None
            else:
                Dm = np.where(bc, 1.0, cM * Mv_diag)
                Sm = self.A11 - self.A10 @ sp.diags(1.0 / Dm) @ self.A01
                self.kS = SAAMG(Sm.tocsr(), 1, None, coarse_size=6000, dense_limit=8192)
                Sv = ((phi0 * d / (2 * mu_f)) * Mp).tocsr()
                self.kV = SAAMG(Sv, 1, None, max_levels=1, cheby_degree=4, dense_limit=0)
        def __call__(self, x):
            y0 = self.k0(x[:nf])
            r = x[nf:] - self.A10 @ y0
            y1 = self.kS(r)
            if self.kV is not None:
                y1 = y1 + self.kV(r)
            return np.concatenate([y0, y1])
    pc = BlockPC(sys_, {"s": krylov_solver("preonly", lambda M: SAAMG(M, dim, B, theta=0.04, coarse_size=6000, dense_limit=8192)), "fp": lambda M: Schur(M)})
    r = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=rtol, atol=0.0, dtol=1e20, max_it=300, restart=300, pc_side="right")
    return r.its
N = int(sys.argv[1])
for scale in [float(a) for a in sys.argv[2:]]:
    t = time.time()
    res = {(v, f): run(3, N, v, f, scale) for v in ("selfp", "cc") for f in ("cheb", "vcycle")}
    print("N", N, "mu_f x", scale, "(emulates N = %.0f)" % (N * scale ** 0.5), res, "(%.0f s)" % (time.time() - t), flush=True)
