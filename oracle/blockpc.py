"""PreconditionerCC.apply restated on scipy blocks (TEST INFRASTRUCTURE).

lib/Preconditioner.py:60-75   sub-matrix extraction (createSubMatrix with is_s/is_f/is_p/is_fp)
lib/Preconditioner.py:219-246 2-way: y_s = K_s\\x_s ; y_fp = K_fp \\ (x_fp - P_fp,s y_s)
lib/Preconditioner.py:150-218 3-way: two (p -> f -> s) sweeps, y = w1*y + w2*y_diff
lib/Preconditioner.py:282-285 flag_3_way = pc type in ('diagonal 3-way','undrained 3-way'), w = (1.0, 0.1)
lib/Preconditioner.py:102-118 + petsc-options-inexact:78-80  fp block: PCFIELDSPLIT schur / lower /
                              selfp with split 0 = pressure, split 1 = fluid velocity

Inner solves are injected as callables so the same composition serves the exact config
(scipy splu standing in for MUMPS, petsc-options-exact:11-35) and the inexact one
(our CG + smoothed-aggregation AMG, oracle/amg.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .krylov import InnerKSP


def submatrix(M: sp.csr_matrix, rows: np.ndarray, cols: np.ndarray) -> sp.csr_matrix:
    return M[rows][:, cols].tocsr()


class LU:
    def __init__(self, M):
        self.lu = spla.splu(sp.csc_matrix(M))

    def __call__(self, x):
        return self.lu.solve(x)


class SchurLowerSelfp:
    """PCFIELDSPLIT(schur, fact lower, precondition selfp) on the fp block, split 0 = p, split 1 = f.

    y_p = K0 \\ x_p ;  y_f = K1(S_f) \\ (x_f - P_fp y_p),   S_f = P_ff - P_fp diag(P_pp)^-1 P_pf
    (lib/Preconditioner.py:113-114 adds is_p first, then is_f; petsc-options-inexact:78-80).
    """

    def __init__(self, Mfp_fp: sp.csr_matrix, nf: int, npp: int, make_k0, make_k1):
        f = np.arange(nf)
        p = nf + np.arange(npp)
        self.nf, self.np_ = nf, npp
        self.App = submatrix(Mfp_fp, p, p)
        self.Afp = submatrix(Mfp_fp, f, p)
        self.Apf = submatrix(Mfp_fp, p, f)
        self.Aff = submatrix(Mfp_fp, f, f)
        dinv = sp.diags(1.0 / self.App.diagonal())
        self.S = (self.Aff - self.Afp @ dinv @ self.Apf).tocsr()
        self.k0 = make_k0(self.App)
        self.k1 = make_k1(self.S)

    def __call__(self, x):
        xf, xp = x[: self.nf], x[self.nf:]
        yp = self.k0(xp)
        yf = self.k1(xf - self.Afp @ yp)
        return np.concatenate([yf, yp])


class BlockPC:
    """y = M^-1 x with the reference's 2-way / 3-way composition."""

    def __init__(self, sys_, solvers: dict, w1=1.0, w2=0.1, anderson=None):
        """solvers: dict of factories {'s','f','p','fp','diff'}: CSR block -> callable x -> y."""
        P = sys_.P
        self.sys = sys_
        self.three_way = sys_.pc_type in ("diagonal 3-way", "undrained 3-way")
        s, f, p, fp = sys_.is_s, sys_.is_f, sys_.is_p, sys_.is_fp
        self.is_s, self.is_f, self.is_p, self.is_fp = s, f, p, fp
        self.Ms_s = submatrix(P, s, s)
        self.k_s = solvers["s"](self.Ms_s)
        self.w1, self.w2 = w1, w2
        self.anderson = anderson
        if self.three_way:
            self.Ms_f, self.Ms_p = submatrix(P, s, f), submatrix(P, s, p)
            self.Mf_f, self.Mf_p = submatrix(P, f, f), submatrix(P, f, p)
            self.Mp_p = submatrix(P, p, p)
            self.Mp_diff = submatrix(sys_.P_diff, p, p)
            self.k_f = solvers["f"](self.Mf_f)
            self.k_p = solvers["p"](self.Mp_p)
            self.k_diff = solvers["diff"](self.Mp_diff)
            self.bcs_sub_pressure = sys_.bcs_sub_pressure
        else:
            self.Mfp_s = submatrix(P, fp, s)
            self.Mfp_fp = submatrix(P, fp, fp)
            self.k_fp = solvers["fp"](self.Mfp_fp)

    def __call__(self, x):
        y = np.zeros_like(x)
        xs = x[self.is_s]
        if self.three_way:
            xf, xp = x[self.is_f], x[self.is_p]
            yp = self.k_p(xp)                                     # :170
            xpd = xp.copy()
            xpd[self.bcs_sub_pressure] = 0.0                      # :172-173
            ypd = self.k_diff(xpd)                                # :174
            yf = self.k_f(xf - self.Mf_p @ yp)                    # :180-182
            yfd = self.k_f(xf - self.Mf_p @ ypd)                  # :184-186
            ys = self.k_s(xs - (self.Ms_f @ yf + self.Ms_p @ yp))       # :192-196
            ysd = self.k_s(xs - (self.Ms_f @ yfd + self.Ms_p @ ypd))    # :198-202
            y[self.is_s] = self.w1 * ys + self.w2 * ysd           # :207-212
            y[self.is_f] = self.w1 * yf + self.w2 * yfd
            y[self.is_p] = self.w1 * yp + self.w2 * ypd
        else:
            ys = self.k_s(xs)                                     # :221
            t = x[self.is_fp] - self.Mfp_s @ ys                   # :232-233
            y[self.is_s] = ys
            y[self.is_fp] = self.k_fp(t)                          # :234
        if self.anderson is not None and self.anderson.order > 0:  # :248-249
            y = self.anderson.get_next_vector(y)
        return y


def exact_solvers():
    """petsc-options-exact: every inner KSP is preonly + LU."""
    mk = lambda M: LU(M)
    return {"s": mk, "f": mk, "p": mk, "fp": mk, "diff": mk}


def krylov_solver(ksp_type, make_pc, **opts):
    """Factory: CSR block -> InnerKSP(ksp_type, pc = make_pc(block))."""
    def mk(M):
        return InnerKSP(M, make_pc(M) if make_pc else None, ksp_type, **opts)
    return mk


class SchurLower:
    """General 2-split PCFIELDSPLIT(schur, lower, selfp) on the fp block.

    first='p' is the reference's order (SchurLowerSelfp above); first='f' eliminates the
    fluid velocity first and puts the Schur complement on the pressure:
        y_f = K0(P_ff) \\ x_f ;  y_p = K1(S_p) \\ (x_p - P_pf y_f),  S_p = P_pp - P_pf diag(P_ff)^-1 P_fp
    """

    def __init__(self, Mfp_fp, nf, npp, make_k0, make_k1, first="p"):
        f = np.arange(nf)
        p = nf + np.arange(npp)
        self.nf, self.np_, self.first = nf, npp, first
        i0, i1 = (p, f) if first == "p" else (f, p)
        self.A00 = submatrix(Mfp_fp, i0, i0)
        self.A01 = submatrix(Mfp_fp, i0, i1)
        self.A10 = submatrix(Mfp_fp, i1, i0)
        self.A11 = submatrix(Mfp_fp, i1, i1)
        self.S = (self.A11 - self.A10 @ sp.diags(1.0 / self.A00.diagonal()) @ self.A01).tocsr()
        self.k0, self.k1 = make_k0(self.A00), make_k1(self.S)

    def __call__(self, x):
        xf, xp = x[: self.nf], x[self.nf:]
        x0, x1 = (xp, xf) if self.first == "p" else (xf, xp)
        y0 = self.k0(x0)
        y1 = self.k1(x1 - self.A10 @ y0)
        yf, yp = (y1, y0) if self.first == "p" else (y0, y1)
        return np.concatenate([yf, yp])


class SchurLowerCC:
    """2-split lower Schur fieldsplit on the fp block (velocity first) with an additive Cahouet-Chabard-type
    preconditioner for the pressure Schur complement (round-2 candidate, profiles/r1_schur_cc_study.md):

        y_f = K0(P_ff) \\ x_f ;  r = x_p - P_pf y_f ;  y_p = K_mass(S_mass) \\ r + K_visc(S_visc) \\ r
        S_mass = P_pp - P_pf diag(d_mass)^-1 P_fp     (d_mass ~ diag of the mass + drag part of P_ff)
        S_visc = w_visc * M_p                          (the viscous limit of the Schur complement)

    `cc_from_matrices` derives d_mass and S_visc from the assembled A and P plus two scalars of the reference's
    parameter dict, i.e. from what crosses the library boundary."""

    def __init__(self, Mfp_fp, nf, npp, make_k0, make_kmass, make_kvisc, d_mass, S_visc):
        f = np.arange(nf)
        p = nf + np.arange(npp)
        self.nf, self.np_ = nf, npp
        self.A00 = submatrix(Mfp_fp, f, f)
        self.A01 = submatrix(Mfp_fp, f, p)
        self.A10 = submatrix(Mfp_fp, p, f)
        self.A11 = submatrix(Mfp_fp, p, p)
        self.S_mass = (self.A11 - self.A10 @ sp.diags(1.0 / d_mass) @ self.A01).tocsr()
        self.S_visc = sp.csr_matrix(S_visc)
        self.k0, self.k_mass, self.k_visc = make_k0(self.A00), make_kmass(self.S_mass), make_kvisc(self.S_visc)

    def __call__(self, x):
        y0 = self.k0(x[: self.nf])
        r = x[self.nf:] - self.A10 @ y0
        return np.concatenate([y0, self.k_mass(r) + self.k_visc(r)])


def cc_from_matrices(sys_, par):
    """(d_mass, S_visc) for SchurLowerCC from the assembled matrices and the physical parameters:
        A_fs = -(phi^2 / (k_f dt)) M_v  (rows of fluid Dirichlet dofs zeroed)  ->  diag(c M_v) = c dt k_f / phi^2 |diag(A_fs)|
        P_pp = (phi_s^2 / (k_s dt) + beta_p) M_p  (`diagonal` splitting)       ->  M_p
    with c = rho_f phi / dt + (1 + beta_f) phi^2 / k_f and w_visc = phi d / (2 mu_f) (beta_CC1, lib/Assembler.py:131)."""
    phi0, dt, kf, d = par["phi0"], par["dt"], par["kf"], sys_.dim
    phis = 1.0 - phi0
    drag = phi0 ** 2 / kf
    c = par["rhof"] * phi0 / dt + (1.0 + par["betaf"]) * drag
    Afs = submatrix(sys_.A, sys_.is_f, sys_.is_s)
    dm = np.abs(Afs.diagonal()) * (c * dt / drag)
    d_mass = np.where(dm > 0, dm, 1.0)
    beta_p = par["betap"] * phis ** 2 / (dt * (2 * par["mu_s"] / d + par["lmbda"]))
    Mp = submatrix(sys_.P, sys_.is_p, sys_.is_p) / (phis ** 2 / (par["ks"] * dt) + beta_p)
    return d_mass, (phi0 * d / (2.0 * par["mu_f"])) * Mp
