"""Host assembler for the three-field poromechanics system (input generation; stands in for FEniCS).

dolfin/UFL are absent from this image, so the matrices the reference's solve
phase consumes are re-derived here from the forms of the reference:

  lib/Assembler.py:80-97    A           (a_s, a_f, a_p)
  lib/Assembler.py:100-117  P 'undrained'
  lib/Assembler.py:118-138  P, P_diff 'undrained 3-way'
  lib/Assembler.py:139-161  P 'diagonal'
  lib/Assembler.py:162-189  P, P_diff 'diagonal 3-way'
  lib/Assembler.py:212-215  P = A for 'lu'
  lib/Assembler.py:243-268  b(t): only surface tractions + volume loads are assembled
  lib/Poromechanics.py:14-18 space P2^d x P2^d x P1
  lib/Poromechanics.py:76-83 DirichletBC.apply on A, P, b (row zeroed, unit diagonal,
                             columns kept); pressure BCs only on P_diff
  lib/MeshCreation.py:11-50,169-215 UnitSquareMesh ("right" diagonal) / UnitCubeMesh
                             (6 tets per cube around the main diagonal), scaled by `length`

DoF numbering is ours (the reference's is dolfin's and not reproducible): fields
are contiguous [s | f | p]; vector fields are node-blocked (dof = dim*node + comp);
structured meshes number P2 nodes on the (2N+1)^d lattice, z slowest.  Only the
mesh, the forms and the BC sets define the system, so this is a permutation of
what dolfin would assemble.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from math import factorial

import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi


# ----------------------------------------------------------------------------------------
# meshes
# ----------------------------------------------------------------------------------------
@dataclass
class Mesh:
    dim: int
    coords: np.ndarray            # (nv, dim) vertex coordinates
    cells: np.ndarray             # (nc, dim+1) vertex ids
    length: float = 1.0
    N: int | None = None          # structured: cells per side
    vlat: np.ndarray | None = None  # structured: integer lattice coords of vertices (nv, dim)

    @property
    def structured(self):
        return self.vlat is not None


def unit_square_mesh(N: int, length: float = 1.0) -> Mesh:
    """dolfin UnitSquareMesh(N, N) ('right' diagonal), lib/MeshCreation.py:11-19."""
    i, j = np.meshgrid(np.arange(N + 1), np.arange(N + 1), indexing="xy")  # j rows (y), i cols (x)
    vlat = np.stack([i.ravel(), j.ravel()], axis=1)            # vertex id = j*(N+1)+i
    vid = lambda ii, jj: jj * (N + 1) + ii
    ci, cj = np.meshgrid(np.arange(N), np.arange(N), indexing="xy")
    ci, cj = ci.ravel(), cj.ravel()
    v0, v1, v2, v3 = vid(ci, cj), vid(ci + 1, cj), vid(ci, cj + 1), vid(ci + 1, cj + 1)
    cells = np.concatenate([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], axis=0)
    coords = vlat.astype(float) * (length / N)
    return Mesh(2, coords, cells, length, N, vlat)


def unit_cube_mesh(N: int, length: float = 1.0, k_range: tuple[int, int] | None = None) -> Mesh:
    """dolfin UnitCubeMesh(N, N, N), lib/MeshCreation.py:169-178.

    `k_range=(k0,k1)` restricts the CELLS to layers k0 <= k < k1 (used to assemble only
    one z-slab of rows per rank); vertex numbering stays global.
    """
    n1 = N + 1
    k, j, i = np.meshgrid(np.arange(n1), np.arange(n1), np.arange(n1), indexing="ij")
    vlat = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)  # vertex id = (k*n1+j)*n1+i
    vid = lambda ii, jj, kk: (kk * n1 + jj) * n1 + ii
    k0, k1 = (0, N) if k_range is None else k_range
    ck, cj, ci = np.meshgrid(np.arange(k0, k1), np.arange(N), np.arange(N), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()
    v = [vid(ci + a, cj + b, ck + c) for c in (0, 1) for b in (0, 1) for a in (0, 1)]
    # v0..v7 with v1=(i+1,j,k), v2=(i,j+1,k), v3=(i+1,j+1,k), v4..v7 the same at k+1
    tets = [(0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (0, 3, 2, 7), (0, 6, 4, 7), (0, 2, 6, 7)]
    cells = np.concatenate([np.stack([v[a], v[b], v[c], v[d]], 1) for a, b, c, d in tets], axis=0)
    coords = vlat.astype(float) * (length / N)
    return Mesh(3, coords, cells, length, N, vlat)


# ----------------------------------------------------------------------------------------
# reference element tensors
# ----------------------------------------------------------------------------------------
def simplex_quadrature(d: int, n: int = 4):
    """Conical-product Gauss-Jacobi rule on {x>=0, sum x <= 1}; exact to degree 2n-1."""
    pts1 = []
    for a in range(d - 1, -1, -1):          # weights (1-u)^(d-1), (1-v)^(d-2), ...
        t, w = roots_jacobi(n, a, 0)
        pts1.append(((t + 1) / 2, w * 0.5 ** (a + 1)))
    X, W = [], []
    for idx in itertools.product(range(n), repeat=d):
        u = [pts1[m][0][idx[m]] for m in range(d)]
        w = np.prod([pts1[m][1][idx[m]] for m in range(d)])
        x, rem = [], 1.0
        for m in range(d):
            x.append(u[m] * rem)
            rem *= (1 - u[m])
        X.append(x)
        W.append(w)
    return np.array(X), np.array(W)


def local_edges(d):
    return [(i, j) for i in range(d + 1) for j in range(i + 1, d + 1)]


def reference_tensors(d: int):
    """Mass / gradient tensors of the P2 and P1 Lagrange bases on the reference simplex."""
    X, W = simplex_quadrature(d, 4)
    lam = np.concatenate([1 - X.sum(1, keepdims=True), X], axis=1)          # (q, d+1)
    dlam = np.concatenate([-np.ones((1, d)), np.eye(d)], axis=0)              # (d+1, d)
    edges = local_edges(d)
    phi = [lam[:, i] * (2 * lam[:, i] - 1) for i in range(d + 1)] + [4 * lam[:, i] * lam[:, j] for i, j in edges]
    dphi = [(4 * lam[:, i] - 1)[:, None] * dlam[i][None, :] for i in range(d + 1)] + \
           [4 * (lam[:, i][:, None] * dlam[j][None, :] + lam[:, j][:, None] * dlam[i][None, :]) for i, j in edges]
    phi = np.stack(phi, 1)                   # (q, n2)
    dphi = np.stack(dphi, 1)                 # (q, n2, d)
    psi = lam                                # (q, d+1)
    dpsi = np.broadcast_to(dlam[None], (len(W), d + 1, d))
    T = {
        "M2": np.einsum("q,qa,qb->ab", W, phi, phi),
        "G2": np.einsum("q,qak,qbl->klab", W, dphi, dphi),
        "D21": np.einsum("q,qak,qb->kab", W, dphi, psi),
        "M1": np.einsum("q,qa,qb->ab", W, psi, psi),
        "G1": np.einsum("q,qak,qbl->klab", W, dpsi, dpsi),
        "F2": np.einsum("q,qa->a", W, phi),
    }
    return T


# ----------------------------------------------------------------------------------------
# dof numbering
# ----------------------------------------------------------------------------------------
def p2_numbering(mesh: Mesh):
    """Return (n2, cell_p2 (nc, n2loc), p2_coords (n2, d)).

    Local order: vertices 0..d then edges in lexicographic (i<j) order.
    Structured meshes: P2 node id = linear index on the (2N+1)^d lattice, z slowest.
    """
    d = mesh.dim
    edges = local_edges(d)
    if mesh.structured:
        L = 2 * mesh.N + 1
        stride = np.array([L ** m for m in range(d)])
        fv = 2 * mesh.vlat[mesh.cells]                                     # (nc, d+1, d)
        fe = np.stack([(mesh.vlat[mesh.cells[:, i]] + mesh.vlat[mesh.cells[:, j]]) for i, j in edges], 1)
        f = np.concatenate([fv, fe], axis=1)                               # (nc, n2loc, d)
        cell_p2 = (f * stride).sum(-1)
        n2 = L ** d
        grid = np.stack(np.meshgrid(*[np.arange(L)] * d, indexing="ij"), -1).reshape(-1, d)[:, ::-1]
        p2_coords = grid.astype(float) * (mesh.length / (2 * mesh.N))
        return n2, cell_p2, p2_coords
    nv = mesh.coords.shape[0]
    pairs = np.stack([np.stack([mesh.cells[:, i], mesh.cells[:, j]], 1) for i, j in edges], 1)  # (nc, ne, 2)
    pairs = np.sort(pairs, axis=2)
    key = pairs[..., 0].astype(np.int64) * nv + pairs[..., 1]
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    cell_e = inv.reshape(key.shape) + nv
    cell_p2 = np.concatenate([mesh.cells, cell_e], axis=1)
    ea, eb = uniq // nv, uniq % nv
    p2_coords = np.concatenate([mesh.coords, 0.5 * (mesh.coords[ea] + mesh.coords[eb])], axis=0)
    return nv + len(uniq), cell_p2, p2_coords


# ----------------------------------------------------------------------------------------
# scalar matrices sharing one pattern
# ----------------------------------------------------------------------------------------
class PatternAssembler:
    """Sums element matrices of several scalar forms over ONE (row,col) pattern."""

    def __init__(self, rows_loc: np.ndarray, cols_loc: np.ndarray, nrows: int, ncols: int):
        nc, na = rows_loc.shape
        nb = cols_loc.shape[1]
        r = np.repeat(rows_loc[:, :, None], nb, axis=2).ravel()
        c = np.repeat(cols_loc[:, None, :], na, axis=1).ravel()
        key = r.astype(np.int64) * ncols + c
        self.order = np.argsort(key, kind="stable")
        ks = key[self.order]
        first = np.ones(len(ks), bool)
        first[1:] = ks[1:] != ks[:-1]
        self.start = np.flatnonzero(first)
        uk = ks[self.start]
        self.rows = (uk // ncols).astype(np.int64)
        self.cols = (uk % ncols).astype(np.int32)
        self.shape = (nrows, ncols)
        self.indptr = np.zeros(nrows + 1, np.int64)
        np.cumsum(np.bincount(self.rows, minlength=nrows), out=self.indptr[1:])
        self.nnz = len(uk)

    def reduce(self, elem_vals: np.ndarray) -> np.ndarray:
        """elem_vals (nc, na, nb) -> values on the pattern (nnz,), roundoff noise zeroed."""
        v = np.add.reduceat(elem_vals.ravel()[self.order], self.start)
        m = np.abs(v).max() if len(v) else 0.0
        v[np.abs(v) < 1e-13 * m] = 0.0
        return v


def element_geometry(mesh: Mesh):
    """Per-element |det J| and T = J^{-T} (so grad = T @ grad_hat); classes of congruent cells."""
    X = mesh.coords[mesh.cells]                          # (nc, d+1, d)
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))  # columns = edge vectors
    # congruence classes (uniform meshes have 2 / 6 of them)
    Jr = np.round(J / (np.abs(J).max() + 1e-300), 10).reshape(len(J), -1)
    uniq, inv = np.unique(Jr, axis=0, return_inverse=True)
    inv = inv.ravel()
    rep = np.zeros(len(uniq), np.int64)
    rep[inv] = np.arange(len(J))
    Jc = J[rep]
    det = np.abs(np.linalg.det(Jc))
    T = np.transpose(np.linalg.inv(Jc), (0, 2, 1))
    return det, T, inv


# ----------------------------------------------------------------------------------------
# the assembled system
# ----------------------------------------------------------------------------------------
@dataclass
class PoroSystem:
    dim: int
    ns: int
    nf: int
    np_: int
    A: sp.csr_matrix
    P: sp.csr_matrix
    P_diff: sp.csr_matrix | None
    b: np.ndarray
    is_s: np.ndarray
    is_f: np.ndarray
    is_p: np.ndarray
    is_fp: np.ndarray
    bcs_sub_pressure: np.ndarray          # positions of pressure-BC dofs inside the p block
    coords_s: np.ndarray                  # (ns, dim) coordinates of each s dof (= f dof)
    coords_p: np.ndarray                  # (np, dim)
    pc_type: str = "diagonal"
    meta: dict = field(default_factory=dict)

    @property
    def n(self):
        return self.ns + self.nf + self.np_


def _bsr_from_pattern(pat: PatternAssembler, data: np.ndarray, bs_r: int, bs_c: int) -> sp.bsr_matrix:
    return sp.bsr_matrix((data, pat.cols, pat.indptr),
                         shape=(pat.shape[0] * bs_r, pat.shape[1] * bs_c), blocksize=(bs_r, bs_c))


def csr_hstack(blocks, col_offsets, ncols):
    """Concatenate CSR blocks with equal row counts side by side (column-sorted if offsets ascend)."""
    nrows = blocks[0].shape[0]
    counts = [np.diff(b.indptr).astype(np.int64) for b in blocks]
    indptr = np.zeros(nrows + 1, np.int64)
    np.cumsum(sum(counts), out=indptr[1:])
    nnz = int(indptr[-1])
    indices = np.empty(nnz, np.int32)
    data = np.empty(nnz, np.float64)
    start = indptr[:-1].copy()
    for b, c, off in zip(blocks, counts, col_offsets):
        if b.nnz:
            shift = np.repeat(start - b.indptr[:-1], c)
            dest = shift + np.arange(b.nnz, dtype=np.int64)
            indices[dest] = b.indices + off
            data[dest] = b.data
        start += c
    return sp.csr_matrix((data, indices, indptr), shape=(nrows, ncols))


class PoroAssembler:
    """Assembles the field blocks once, then composes A / P / P_diff for any pc type."""

    def __init__(self, mesh: Mesh, params: dict):
        self.mesh, self.par = mesh, params
        d = self.dim = mesh.dim
        T = reference_tensors(d)
        self.n2, self.cell_p2, self.p2_coords = p2_numbering(mesh)
        self.n1 = mesh.coords.shape[0]
        cell_p1 = mesh.cells
        det, Tm, cls = element_geometry(mesh)

        # ---- P2 x P2: mass M and gradient matrices G[i][j] = int d_i(phi_a) d_j(phi_b)
        self.pat22 = PatternAssembler(self.cell_p2, self.cell_p2, self.n2, self.n2)
        self.M = self.pat22.reduce((det[:, None, None] * T["M2"][None])[cls])
        Gc = np.einsum("e,eik,ejl,klab->eijab", det, Tm, Tm, T["G2"])
        self.G = [[self.pat22.reduce(Gc[:, i, j][cls]) for j in range(d)] for i in range(d)]
        # ---- P2 x P1: D[i] = int d_i(phi_a) psi_b
        self.pat21 = PatternAssembler(self.cell_p2, cell_p1, self.n2, self.n1)
        Dc = np.einsum("e,eik,kab->eiab", det, Tm, T["D21"])
        self.D = [self.pat21.reduce(Dc[:, i][cls]) for i in range(d)]
        # ---- P1 x P1: mass and Laplacian
        self.pat11 = PatternAssembler(cell_p1, cell_p1, self.n1, self.n1)
        self.Mp = self.pat11.reduce((det[:, None, None] * T["M1"][None])[cls])
        Kc = np.einsum("e,eik,eil,klab->eab", det, Tm, Tm, T["G1"])
        self.Kp = self.pat11.reduce(Kc[cls])
        self._T = T
        self.bc_s = np.zeros((self.n2, d), bool)
        self.bc_f = np.zeros((self.n2, d), bool)
        self.bc_p = np.zeros(self.n1, bool)

    # -------- vector-field building blocks on the P2 pattern (data arrays (nnzb, d, d))
    def _mass_blocks(self):
        d = self.dim
        out = np.zeros((self.pat22.nnz, d, d))
        for i in range(d):
            out[:, i, i] = self.M
        return out

    def _eps_blocks(self):
        """(eps(u), eps(v)) with u = phi_b e_j (trial), v = phi_a e_i (test)."""
        d = self.dim
        lap = sum(self.G[k][k] for k in range(d))
        out = np.zeros((self.pat22.nnz, d, d))
        for i in range(d):
            for j in range(d):
                out[:, i, j] = 0.5 * self.G[j][i]
            out[:, i, i] += 0.5 * lap
        return out

    def _divdiv_blocks(self):
        d = self.dim
        out = np.zeros((self.pat22.nnz, d, d))
        for i in range(d):
            for j in range(d):
                out[:, i, j] = self.G[i][j]
        return out

    def _div_blocks(self):
        """int psi_b d_i(phi_a): rows (a,i) [P2 vector], cols b [P1]; data (nnz21, d, 1)."""
        return np.stack(self.D, axis=1)[:, :, None]

    # -------- boundary conditions (DirichletBC on marked sides; structured or callable)
    def side_nodes(self, which: str, side):
        """Boolean mask of P2 ('2') or P1 ('1') nodes on `side`.

        side: 'x0','x1','y0','y1','z0','z1' (coordinate plane) or callable(coords)->mask.
        Matches dolfin's facet-marker DirichletBC: every dof on a marked boundary facet.
        """
        X = self.p2_coords if which == "2" else self.mesh.coords
        if callable(side):
            return side(X)
        ax = "xyz".index(side[0])
        val = 0.0 if side[1] == "0" else self.mesh.length
        return np.abs(X[:, ax] - val) < 1e-10 * max(self.mesh.length, 1.0)

    def set_bcs(self, bcs_s=(), bcs_f=(), bcs_p=()):
        """bcs_s / bcs_f: iterables of (side, comp or None); bcs_p: iterable of side."""
        for mask, bcs in ((self.bc_s, bcs_s), (self.bc_f, bcs_f)):
            mask[:] = False
            for side, comp in bcs:
                on = self.side_nodes("2", side)
                if comp is None:
                    mask[on, :] = True
                else:
                    mask[on, comp] = True
        self.bc_p[:] = False
        for side in bcs_p:
            self.bc_p |= self.side_nodes("1", side)

    # -------- block composition
    def _coeffs(self):
        p = self.par
        phi0 = p["phi0"]
        return dict(phi0=phi0, phis=1 - phi0, idt=1.0 / p["dt"], ikf=1.0 / p["kf"], d=self.dim,
                    mu_s=p["mu_s"], lmbda=p["lmbda"], rhos=p["rhos"], rhof=p["rhof"], mu_f=p["mu_f"],
                    ks=p["ks"], dt=p["dt"], betas=p["betas"], betaf=p["betaf"], betap=p["betap"])

    def field_blocks(self, which: str, pc_type: str = "diagonal"):
        """Return the 3x3 dict of un-BC'd field blocks for which in {'A','P','P_diff'}.

        Each value is (kind, data) with kind in {'22','21','12','11'} naming the pattern.
        """
        c = self._coeffs()
        phi0, phis, idt, ikf, d = c["phi0"], c["phis"], c["idt"], c["ikf"], c["d"]
        Mv, Ke, Kdd, Dv = self._mass_blocks(), self._eps_blocks(), self._divdiv_blocks(), self._div_blocks()
        KC = 2 * c["mu_s"] * Ke + c["lmbda"] * Kdd            # hooke(eps(u)) : eps(v)
        ms = c["rhos"] * idt ** 2 * phis
        mf = c["rhof"] * idt * phi0
        drag = phi0 ** 2 * ikf
        # --- A (Assembler.py:80-93)
        A = {
            "ss": ("22", (ms + drag * idt) * Mv + KC),
            "sf": ("22", -drag * Mv),
            "sp": ("21", -phis * Dv),
            "fs": ("22", -drag * idt * Mv),
            "ff": ("22", (mf + drag) * Mv + 2 * c["mu_f"] * phi0 * Ke),
            "fp": ("21", -phi0 * Dv),
            "ps": ("12", phis * idt * Dv),
            "pf": ("12", phi0 * Dv),
            "pp": ("11", phis ** 2 * idt / c["ks"] * self.Mp),
        }
        if which == "A" or pc_type == "lu":
            return A
        beta_p = c["betap"] * phis ** 2 / (c["dt"] * (2 * c["mu_s"] / d + c["lmbda"]))
        beta_CC1 = phi0 / (2 * c["mu_f"] / d)
        beta_CC2 = 1.0 / (c["rhof"] * idt / phi0 + ikf)
        pp0 = phis ** 2 * idt / c["ks"]
        if pc_type in ("undrained", "undrained 3-way"):
            Nn = c["ks"] / phis ** 2
            P = dict(A)
            # Assembler.py:103-106 / :121-124: no p and no vf coupling in the solid rows
            P["ss"] = ("22", (ms + drag * idt) * Mv + KC + Nn * phis ** 2 * Kdd)
            P["sf"] = ("22", 0.0 * Mv)
            P["sp"] = ("21", 0.0 * Dv)
            if pc_type == "undrained 3-way":
                P["pp"] = ("11", (pp0 + beta_CC1) * self.Mp)             # :134-135 (no div terms)
                P["ps"] = ("12", 0.0 * Dv)
                P["pf"] = ("12", 0.0 * Dv)
                Pd = dict(P)
                Pd["pp"] = ("11", pp0 * self.Mp + beta_CC2 * self.Kp)    # :137-138
                return P if which == "P" else Pd
            return P
        if pc_type in ("diagonal", "diagonal 3-way"):
            P = dict(A)
            P["ss"] = ("22", (ms + (1.0 + c["betas"]) * drag * idt) * Mv + KC)   # :142-145
            P["fs"] = ("22", 0.0 * Mv)                                           # :149-152 no us
            P["ff"] = ("22", (mf + (1.0 + c["betaf"]) * drag) * Mv + 2 * c["mu_f"] * phi0 * Ke)
            P["ps"] = ("12", 0.0 * Dv)                                           # :158-160 no us
            if pc_type == "diagonal":
                P["pp"] = ("11", (pp0 + beta_p) * self.Mp)
                return P
            P["pf"] = ("12", 0.0 * Dv)                                           # :184-185
            P["pp"] = ("11", (pp0 + beta_p + beta_CC1) * self.Mp)
            Pd = dict(P)
            Pd["pp"] = ("11", (pp0 + beta_p) * self.Mp + beta_CC2 * self.Kp)     # :187-189
            return P if which == "P" else Pd
        raise ValueError("unsupported pc type %r" % pc_type)

    def _to_csr(self, kind, data):
        d = self.dim
        if kind == "22":
            m = _bsr_from_pattern(self.pat22, data, d, d).tocsr()
        elif kind == "21":
            m = _bsr_from_pattern(self.pat21, data, d, 1).tocsr()
        elif kind == "12":
            m = _bsr_from_pattern(self.pat21, data, d, 1).tocsr().T.tocsr()
        else:
            m = sp.csr_matrix((data, self.pat11.cols, self.pat11.indptr), shape=self.pat11.shape)
        return m

    def compose(self, blocks: dict, apply_p_bc: bool = False) -> sp.csr_matrix:
        """Global field-major CSR with DirichletBC.apply semantics (Poromechanics.py:76-83)."""
        d = self.dim
        ns = nf = self.n2 * d
        npp = self.n1
        n = ns + nf + npp
        offs = [0, ns, ns + nf]
        bcrow = {"s": self.bc_s.ravel(), "f": self.bc_f.ravel(), "p": self.bc_p if apply_p_bc else np.zeros(npp, bool)}
        rows_out = []
        for fi, fr in enumerate("sfp"):
            row_blocks = []
            for fc in "sfp":
                v = blocks[fr + fc]                     # (kind, pattern data), or an already global sparse block
                m = self._to_csr(*v) if isinstance(v, tuple) else self._to_csr(v)
                bc = bcrow[fr]
                if bc.any():
                    keep = np.repeat(~bc, np.diff(m.indptr))
                    m.data *= keep
                    if fr == fc:
                        m = m + sp.csr_matrix((np.ones(bc.sum()), (np.flatnonzero(bc), np.flatnonzero(bc))), shape=m.shape)
                m.eliminate_zeros()
                m.sort_indices()
                row_blocks.append(m)
            rows_out.append(csr_hstack(row_blocks, offs, n))
        M = sp.vstack(rows_out, format="csr")
        M.indices = M.indices.astype(np.int32)
        return M

    # -------- right-hand side (Assembler.py:243-244, 250-251, 267-268)
    def boundary_facets(self):
        """Boundary facets: (cell index, local vertex ids of the facet, opposite local vertex)."""
        d = self.dim
        cells = self.mesh.cells
        nc = len(cells)
        facs, opp = [], []
        for o in range(d + 1):
            loc = [i for i in range(d + 1) if i != o]
            facs.append(np.sort(cells[:, loc], axis=1))
            opp.append(np.full(nc, o))
        F = np.concatenate(facs, 0)
        O = np.concatenate(opp, 0)
        C = np.tile(np.arange(nc), d + 1)
        nv = self.mesh.coords.shape[0]
        key = np.zeros(len(F), np.int64)
        for m in range(d):
            key = key * nv + F[:, m]
        _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        on = cnt[inv.ravel()] == 1
        return C[on], O[on]

    def rhs(self, t: float, neumann_solid=(), neumann_fluid=(), fs_sur=None, ff_sur=None):
        """b(t) for constant-magnitude normal tractions `f(t) * n` on the listed sides.

        fs_sur / ff_sur: callables t -> scalar multiplying the outward FacetNormal
        (swelling.py:35-40).  Volume loads and p_source are zero in every shipped driver.
        """
        d = self.dim
        ns = self.n2 * d
        b = np.zeros(2 * ns + self.n1)
        C, O = self.boundary_facets()
        cells = self.mesh.cells
        X = self.mesh.coords
        edges = local_edges(d)
        F2 = None
        for field_off, sides, fun in ((0, neumann_solid, fs_sur), (ns, neumann_fluid, ff_sur)):
            if fun is None or not sides:
                continue
            mag = fun(t)
            for side in sides:
                ax = "xyz".index(side[0])
                val = 0.0 if side[1] == "0" else self.mesh.length
                for o in range(d + 1):
                    sel = C[O == o]
                    loc = [i for i in range(d + 1) if i != o]
                    fx = X[cells[sel][:, loc]]                                  # (nfac, d, d)
                    on = np.all(np.abs(fx[:, :, ax] - val) < 1e-10 * max(self.mesh.length, 1.0), axis=1)
                    sel, fx = sel[on], fx[on]
                    if not len(sel):
                        continue
                    if d == 2:
                        tvec = fx[:, 1] - fx[:, 0]
                        meas = np.linalg.norm(tvec, axis=1)
                        nrm = np.stack([tvec[:, 1], -tvec[:, 0]], 1) / meas[:, None]
                    else:
                        cr = np.cross(fx[:, 1] - fx[:, 0], fx[:, 2] - fx[:, 0])
                        meas = 0.5 * np.linalg.norm(cr, axis=1)
                        nrm = cr / (2 * meas)[:, None]
                    inward = X[cells[sel, o]] - fx[:, 0]
                    flip = np.sign(-(nrm * inward).sum(1))
                    nrm = nrm * flip[:, None]
                    # integral of the P2 basis over the facet (facet is itself a P2 simplex of dim d-1)
                    if F2 is None:
                        F2 = reference_tensors(d - 1)["F2"] * factorial(d - 1) if d > 1 else None
                    w = F2                                                     # per unit measure
                    fed = local_edges(d - 1)
                    # local P2 dof ids (in the cell) of the facet's vertices and edges
                    vloc = loc
                    eloc = [d + 1 + edges.index((min(loc[i], loc[j]), max(loc[i], loc[j]))) for i, j in fed]
                    dofs = self.cell_p2[sel][:, vloc + eloc]                   # (nfac, n2 facet)
                    contrib = mag * meas[:, None, None] * w[None, :, None] * nrm[:, None, :]   # (nfac, nloc, d)
                    idx = field_off + dofs[:, :, None] * d + np.arange(d)[None, None, :]
                    np.add.at(b, idx.ravel(), contrib.ravel())
        bc = np.concatenate([self.bc_s.ravel(), self.bc_f.ravel(), np.zeros(self.n1, bool)])
        b[bc] = 0.0
        return b

    def rhs_vector_load_2d(self, side: str, g, field_off: int = 0):
        """b for a VECTOR surface load on one side of a 2D mesh, g(x) -> (n, 2), interpolated linearly between the
        facet's end vertices (a dolfin `Expression(..., degree=1)`, footing.py:38-39).  Returns the full-length vector.
        Exact edge integrals of P1 x P2: int l_a phi_a = L/6, int l_a phi_b = 0, int l_a phi_mid = L/3."""
        assert self.dim == 2
        d = 2
        ns = self.n2 * d
        b = np.zeros(2 * ns + self.n1)
        C, O = self.boundary_facets()
        cells, X = self.mesh.cells, self.mesh.coords
        edges = local_edges(2)
        ax = "xy".index(side[0])
        val = 0.0 if side[1] == "0" else self.mesh.length
        for o in range(3):
            sel = C[O == o]
            loc = [i for i in range(3) if i != o]
            fx = X[cells[sel][:, loc]]
            on = np.all(np.abs(fx[:, :, ax] - val) < 1e-10 * max(self.mesh.length, 1.0), axis=1)
            sel, fx = sel[on], fx[on]
            if not len(sel):
                continue
            L = np.linalg.norm(fx[:, 1] - fx[:, 0], axis=1)
            ga, gb = g(fx[:, 0]), g(fx[:, 1])                                   # (nfac, 2)
            eloc = d + 1 + edges.index((min(loc), max(loc)))
            da, db, dm = self.cell_p2[sel, loc[0]], self.cell_p2[sel, loc[1]], self.cell_p2[sel, eloc]
            for comp in range(2):
                np.add.at(b, field_off + da * d + comp, L * ga[:, comp] / 6.0)
                np.add.at(b, field_off + db * d + comp, L * gb[:, comp] / 6.0)
                np.add.at(b, field_off + dm * d + comp, L * (ga[:, comp] + gb[:, comp]) / 3.0)
        bc = np.concatenate([self.bc_s.ravel(), self.bc_f.ravel(), np.zeros(self.n1, bool)])
        b[bc] = 0.0
        return b

    # -------- everything
    def system(self, pc_type: str, t: float, neumann_solid=(), neumann_fluid=(), fs_sur=None, ff_sur=None, b=None) -> PoroSystem:
        d = self.dim
        ns = nf = self.n2 * d
        npp = self.n1
        three_way = "3-way" in pc_type          # Poromechanics.py:22-24
        A = self.compose(self.field_blocks("A"))
        P = self.compose(self.field_blocks("P", pc_type))
        P_diff = self.compose(self.field_blocks("P_diff", pc_type), apply_p_bc=True) if three_way else None
        if b is None:
            b = self.rhs(t, neumann_solid, neumann_fluid, fs_sur, ff_sur)
        is_s = np.arange(ns, dtype=np.int64)
        is_f = ns + np.arange(nf, dtype=np.int64)
        is_p = ns + nf + np.arange(npp, dtype=np.int64)
        is_fp = np.concatenate([is_f, is_p])
        coords_s = np.repeat(self.p2_coords, d, axis=0)
        return PoroSystem(d, ns, nf, npp, A, P, P_diff, b, is_s, is_f, is_p, is_fp,
                          np.flatnonzero(self.bc_p).astype(np.int64), coords_s, self.mesh.coords.copy(), pc_type)
