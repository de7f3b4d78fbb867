"""Round-2 design study on the CPU oracle: a Cahouet-Chabard-type pressure-Schur preconditioner for the 2-way fieldsplit.

selfp approximates P_ff^-1 by diag(P_ff)^-1, which degrades when the viscous part of P_ff (~ 1/h^2) overtakes its
mass + drag part (profiles/r1_robustness_2d.md).  The additive variant applies
    S^-1 r  ~=  S_mass^-1 r + S_visc^-1 r,   S_mass = P_pp - P_pf diag(c M_v)^-1 P_fp,   S_visc = (phi d / (2 mu_f)) M_p
(the weight is the reference's beta_CC1, lib/Assembler.py:131; c = rho_f phi / dt + (1 + beta_f) phi^2 / k_f).

    python profiles/schur_cc_study.py 2d | 2dv | 3d N [N ...]        (2dv: V-cycle instead of Chebyshev(4) on the fluid block)
"""
import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, scipy.sparse as sp
from oracle.amg import SAAMG, rigid_body_modes
from oracle.blockpc import BlockPC, krylov_solver, submatrix
from oracle.krylov import gmres
from hostfem.problems import swelling_assembler
FKIND = "cheb"
def run(dim, N, variant, rtol=1e-8, atol=0.0, overrides=None):
    asm, par, loads = swelling_assembler(dim, N, overrides)
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
            self.k0 = SAAMG(self.A00, dim, B, max_levels=1, cheby_degree=4) if FKIND == "cheb" else SAAMG(self.A00, dim, B)
            bc = asm.bc_f.ravel()
            if variant == "selfp":
                S = self.A11 - self.A10 @ sp.diags(1.0 / self.A00.diagonal()) @ self.A01
                self.kS = SAAMG(S.tocsr(), 1, None); self.kV = None
            else:
                Dm = np.where(bc, 1.0, cM * Mv_diag)
                Sm = self.A11 - self.A10 @ sp.diags(1.0 / Dm) @ self.A01
                self.kS = SAAMG(Sm.tocsr(), 1, None)
                Sv = ((phi0 * d / (2 * mu_f)) * Mp).tocsr()
                self.kV = SAAMG(Sv, 1, None, max_levels=1, cheby_degree=4)       # mass matrix: Chebyshev(4) is plenty
        def __call__(self, x):
            y0 = self.k0(x[:nf])
            r = x[nf:] - self.A10 @ y0
            y1 = self.kS(r)
            if self.kV is not None:
                y1 = y1 + self.kV(r)
            return np.concatenate([y0, y1])
    pc = BlockPC(sys_, {"s": krylov_solver("preonly", lambda M: SAAMG(M, dim, B, theta=0.04 if dim == 3 else 0.08)), "fp": lambda M: Schur(M)})
    r = gmres(lambda v: sys_.A @ v, sys_.b, pc, rtol=rtol, atol=atol, dtol=1e20, max_it=500, restart=500, pc_side="right")
    true = np.linalg.norm(sys_.b - sys_.A @ r.x) / np.linalg.norm(sys_.b)
    return r.its
if sys.argv[1] == "2d":
    for variant in ("selfp", "cc"):
        print("2D AMG", variant, [run(2, N, variant, 1e-6, 1e-8) for N in (10, 20, 40, 80)], flush=True)
else:
    for N in [int(a) for a in sys.argv[2:]]:
        t = time.time()
        print("3D AMG N", N, "selfp", run(3, N, "selfp"), "cc", run(3, N, "cc"), "(%.0f s)" % (time.time() - t), flush=True)
if sys.argv[1] == "2dv":
    FKIND = "vcycle"
    for variant in ("selfp", "cc"):
        print("2D AMG (f V-cycle)", variant, [run(2, N, variant, 1e-6, 1e-8) for N in (10, 20, 40, 80)], flush=True)
