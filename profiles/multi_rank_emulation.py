"""CPU emulation (oracle AMG) of the multi-rank preconditioner variants: rank-local coarse levels, global coarsest
solve, overlapping (RAS) local hierarchies, pressure Schur complement from owned parts.  python profiles/multi_rank_emulation.py N R   (N >= 10 so that every slab has at least two AMG levels)"""
import sys, time, numpy as np, scipy.sparse as sp
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.problems import swelling
from oracle.krylov import gmres
from oracle.blockpc import *
from oracle.amg import SAAMG, rigid_body_modes

class DDAmg:
    """rank-local hierarchies on principal sub-blocks, global level-0 smoothing, optional global coarsest."""
    def __init__(self, A, bs, B, parts, global_coarsest=False, **kw):
        self.A = sp.csr_matrix(A); self.parts = parts
        self.loc = [SAAMG(self.A[p][:, p], bs, None if B is None else B[p], **kw) for p in parts]
        d = self.A.diagonal(); self.dinv = 1.0/np.where(d!=0,d,1.0)
        self.lmax = max(h.levels[0].lmax for h in self.loc)
        self.deg, self.ratio = self.loc[0].deg, self.loc[0].ratio
        self.gc = global_coarsest
        if global_coarsest:
            Zs = []
            for h in self.loc:
                Z = sp.identity(h.levels[0].A.shape[0], format='csr')
                for L in h.levels[:-1]: Z = (Z @ L.P).tocsr()
                Zs.append(Z)
            n = self.A.shape[0]; ncs = [Z.shape[1] for Z in Zs]; off = np.concatenate([[0], np.cumsum(ncs)])
            rows, cols, vals = [], [], []
            Zg = sp.lil_matrix((n, off[-1]))
            Zg = sp.bmat([[None]*len(Zs)]*0) if False else None
            blocks = []
            for r,(p,Z) in enumerate(zip(parts,Zs)):
                Zc = Z.tocoo(); blocks.append(sp.csr_matrix((Zc.data, (p[Zc.row], off[r]+Zc.col)), shape=(n, off[-1])))
            self.Z = sum(blocks).tocsr()
            AH = (self.Z.T @ self.A @ self.Z).toarray()
            self.AHinv = np.linalg.inv(AH); self.off = off
    def cheby(self, b, x, zero):
        lmax=self.lmax; lmin=lmax/self.ratio; th=.5*(lmax+lmin); de=.5*(lmax-lmin); sg=th/de; rho=1/sg
        r = b.copy() if zero else b - self.A@x
        d = self.dinv*r/th
        for k in range(self.deg):
            x = x + d
            if k==self.deg-1: break
            r = r - self.A@d; rn=1/(2*sg-rho); d = rn*rho*d + (2*rn/de)*(self.dinv*r); rho=rn
        return x
    def __call__(self, b):
        x = self.cheby(b, np.zeros_like(b), True)
        r = b - self.A@x
        if not self.gc:
            for p,h in zip(self.parts, self.loc):
                L0 = h.levels[0]
                if len(h.levels)==1: continue
                xc = h._cycle(1, L0.R @ r[p]); x[p] += L0.P @ xc
        else:
            # local cycles down to (not incl.) coarsest, global coarsest solve
            # restrict chain per rank, collecting level states
            states=[]
            rc_all = np.zeros(self.off[-1])
            for r_i,(p,h) in enumerate(zip(self.parts, self.loc)):
                xs=[None]*len(h.levels); bs_=[None]*len(h.levels)
                bs_[0]=None
                bl = h.levels[0].R @ r[p]
                for l in range(1, len(h.levels)-1):
                    L=h.levels[l]; bs_[l]=bl
                    xl = h._cheby(L, bl, np.zeros_like(bl), True); xs[l]=xl
                    bl = L.R @ (bl - L.A@xl)
                rc_all[self.off[r_i]:self.off[r_i+1]] = bl
                states.append((xs,bs_))
            xc_all = self.AHinv @ rc_all
            for r_i,(p,h) in enumerate(zip(self.parts, self.loc)):
                xs,bs_ = states[r_i]
                xl1 = xc_all[self.off[r_i]:self.off[r_i+1]]
                for l in range(len(h.levels)-2, 0, -1):
                    L=h.levels[l]; xl = xs[l] + L.P @ xl1
                    xl1 = h._cheby(L, bs_[l], xl, False)
                x[p] += h.levels[0].P @ xl1
        return self.cheby(b, x, False)

N = int(sys.argv[1]) if len(sys.argv)>1 else 10
R = int(sys.argv[2]) if len(sys.argv)>2 else 2
s,par = swelling(3,N,"diagonal")
B = rigid_body_modes(s.coords_s, 3)
z_s = s.coords_s[:,2]; z_p = s.coords_p[:,2]
edges = np.linspace(z_s.min()-1e-12, z_s.max()+1e-12, R+1)
parts_s = [np.flatnonzero((z_s>edges[r])&(z_s<=edges[r+1])) for r in range(R)]
parts_p = [np.flatnonzero((z_p>edges[r])&(z_p<=edges[r+1])) for r in range(R)]
kw = dict(theta=0.04)
def run(mk_s, mk_p, name):
    cheb_f = lambda M: SAAMG(M, 3, B, max_levels=1, cheby_degree=4, dense_limit=0)
    mkfp = lambda M: SchurLower(M, s.nf, s.np_, krylov_solver("preonly", cheb_f), krylov_solver("preonly", mk_p), "f")
    pc = BlockPC(s, {"s": krylov_solver("preonly", mk_s), "fp": mkfp})
    t=time.time()
    ro = gmres(lambda v: s.A@v, s.b, pc, rtol=1e-8, atol=0, dtol=1e20, max_it=400, restart=400, pc_side="right")
    print(N, R, name, "outer its", ro.its, ro.reason, "t %.0f"%(time.time()-t), flush=True)
g_s = lambda M: SAAMG(M,3,B,**kw); g_p = lambda M: SAAMG(M,1,None)
run(g_s, g_p, "global AMG (single GPU)")
run(lambda M: DDAmg(M,3,B,parts_s,False,**kw), g_p, "s local coarse, p global")
run(g_s, lambda M: DDAmg(M,1,None,parts_p,False), "s global, p local coarse")
run(lambda M: DDAmg(M,3,B,parts_s,False,**kw), lambda M: DDAmg(M,1,None,parts_p,False), "both local coarse")
run(lambda M: DDAmg(M,3,B,parts_s,True,**kw), lambda M: DDAmg(M,1,None,parts_p,False), "s global-coarsest, p local")
run(lambda M: DDAmg(M,3,B,parts_s,True,**kw), lambda M: DDAmg(M,1,None,parts_p,True), "both global-coarsest")

class RASAmg(DDAmg):
    def __init__(self, A, bs, B, parts, parts_ext, **kw):
        self.A = sp.csr_matrix(A); self.parts = parts; self.ext = parts_ext
        self.loc = [SAAMG(self.A[p][:, p], bs, None if B is None else B[p], **kw) for p in parts_ext]
        d = self.A.diagonal(); self.dinv = 1.0/np.where(d!=0,d,1.0)
        self.lmax = max(h.levels[0].lmax for h in self.loc); self.deg, self.ratio = self.loc[0].deg, self.loc[0].ratio
        self.own_in_ext = [np.isin(pe, p) for p,pe in zip(parts, parts_ext)]
    def __call__(self, b):
        x = self.cheby(b, np.zeros_like(b), True)
        r = b - self.A@x
        for p,pe,m,h in zip(self.parts, self.ext, self.own_in_ext, self.loc):
            L0=h.levels[0]
            xc = h._cycle(1, L0.R @ r[pe]); corr = L0.P @ xc
            x[pe[m]] += corr[m]
        return self.cheby(b, x, False)

h = z_s.max()/(2*N)
for layers in (2, 4):
    ext = [np.flatnonzero((z_s>edges[r]-layers*h-1e-12)&(z_s<=edges[r+1]+layers*h+1e-12)) for r in range(R)]
    run(lambda M: RASAmg(M,3,B,parts_s,ext,**kw), g_p, "s RAS overlap %d planes"%layers)

# --- pressure Schur complement assembled from owned parts only (what the multi-rank code does) + local AMG
z_f = z_s
parts_f = parts_s
class LocalSchurAmg:
    def __init__(self, S_global):
        # rebuild S from owned parts: S_r = App[p,p] - Apf[p, f_own] dinv Afp[f_own, p]
        fp = submatrix(s.P, s.is_fp, s.is_fp)
        f = np.arange(s.nf); p = s.nf + np.arange(s.np_)
        Aff = fp[f][:, f]; Afp = fp[f][:, p]; Apf = fp[p][:, f]; App = fp[p][:, p]
        dinv = 1.0/Aff.diagonal()
        self.parts = parts_p; self.loc=[]
        for pf, pp in zip(parts_f, parts_p):
            Sr = (App[pp][:, pp] - Apf[pp][:, pf] @ sp.diags(dinv[pf]) @ Afp[pf][:, pp]).tocsr()
            self.loc.append(SAAMG(Sr, 1, None))
    def __call__(self, b):
        x = np.zeros_like(b)
        for pp,h in zip(self.parts, self.loc): x[pp] = h(b[pp])
        return x
run(g_s, lambda M: LocalSchurAmg(M), "s global, p: S_p from owned parts + local AMG (block Jacobi)")
ext2 = [np.flatnonzero((z_s>edges[r]-2*h-1e-12)&(z_s<=edges[r+1]+2*h+1e-12)) for r in range(R)]
run(lambda M: RASAmg(M,3,B,parts_s,ext2,**kw), lambda M: LocalSchurAmg(M), "s RAS overlap 2 + p local Schur")
