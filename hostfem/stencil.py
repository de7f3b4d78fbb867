"""Row-wise generator of the structured-mesh systems (the blueprint of a device-side matrix generator).

On dolfin's UnitSquareMesh / UnitCubeMesh every macro cell (square = 2 triangles, cube = 6 tetrahedra,
lib/MeshCreation.py:11-19, 169-178) is a translate of the first one and the coefficients of all forms are
constants (lib/Assembler.py:80-189), so the assembled matrix is the scatter-sum of ONE macro-cell matrix.
Written row-wise this needs neither a sort nor atomics: the row of a node is the sum over the (at most 2^d)
cells that contain it, and which cells exist depends only on the node's position class per axis
(P2 lattice: odd / even interior / first plane / last plane; P1 lattice: interior / first / last).  Each of
the <= 4^d (3^d) classes has ONE stencil -- a list of column offsets with their value blocks -- and every row
of the class is that stencil shifted.  A CUDA kernel that holds the class tables in constant memory and lets
one thread write one row therefore generates A, P and P_diff at write bandwidth, per rank for its own
z-slab of rows (`plane_range`), which is what the >= 50 M-DoF configurations need (the
host cannot even store those matrices).

Here the same algorithm in numpy: the macro-cell matrices come from `PoroAssembler` on a one-cell mesh, the
classes are expanded vectorised.  `tests/test_hostfem_stencil.py` checks the result against the generic
element-by-element assembler (matrices, right-hand side, boundary conditions, index sets).
Host-side input generation standing in for FEniCS -- not on the GPU solve path, not the oracle.
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.sparse as sp

from .fem import PoroAssembler, PoroSystem, unit_cube_mesh, unit_square_mesh


def _axis_classes(kind: int, N: int):
    """Position classes of one lattice axis: (coordinates, [(e, a)]): the node lies in cells c0 + e with local
    coordinate a, where c0 = x // 2 on the P2 lattice (kind 2) and c0 = x on the P1 lattice (kind 1)."""
    if kind == 2:
        L = 2 * N + 1
        out = [(np.arange(1, L, 2), [(0, 1)]), (np.array([0]), [(0, 0)]), (np.array([L - 1]), [(-1, 2)])]
        if N > 1:
            out.append((np.arange(2, L - 1, 2), [(-1, 2), (0, 0)]))
        return out
    out = [(np.array([0]), [(0, 0)]), (np.array([N]), [(-1, 1)])]
    if N > 1:
        out.append((np.arange(1, N), [(-1, 1), (0, 0)]))
    return out


class StructuredGenerator:
    def __init__(self, dim: int, N: int, length: float, params: dict):
        self.dim, self.N, self.length, self.par = dim, N, length, params
        h = length / N
        self.cell = PoroAssembler((unit_square_mesh if dim == 2 else unit_cube_mesh)(1, h), params)
        self.L = {2: 2 * N + 1, 1: N + 1}            # lattice extent per kind
        self.nloc = {2: 3, 1: 2}                      # local lattice extent inside one cell
        self.n2, self.n1 = self.L[2] ** dim, self.L[1] ** dim
        d = dim
        self.bc_s = np.zeros((self.n2, d), bool)
        self.bc_f = np.zeros((self.n2, d), bool)
        self.bc_p = np.zeros(self.n1, bool)
        self.drop_tol = 1e-13
        self.class_tables = {}                        # (which, pc_type, key) -> list of class stencils (for inspection)

    # ---- lattices ---------------------------------------------------------------------------------------
    def lattice(self, kind: int) -> np.ndarray:
        """(n, d) integer coordinates of the nodes, x fastest (node id = sum x_m L^m)."""
        L = self.L[kind]
        g = np.stack(np.meshgrid(*[np.arange(L)] * self.dim, indexing="ij"), -1).reshape(-1, self.dim)[:, ::-1]
        return g

    def coords(self, kind: int) -> np.ndarray:
        return self.lattice(kind).astype(float) * (self.length / (self.L[kind] - 1))

    def side_nodes(self, which: str, side: str) -> np.ndarray:
        kind = 2 if which == "2" else 1
        ax = "xyz".index(side[0])
        val = 0 if side[1] == "0" else self.L[kind] - 1
        return self.lattice(kind)[:, ax] == val

    set_bcs = PoroAssembler.set_bcs

    # ---- class stencils ---------------------------------------------------------------------------------
    def class_stencils(self, K: np.ndarray, kr: int, kc: int, br: int, bc: int):
        """Yields (axis class codes, axis classes, column offsets (tuples), value blocks) for every position class
        of the row lattice: the stencil of the class = sum over the cells containing the node of the macro-cell rows."""
        d, N = self.dim, self.N
        nr, nc = self.nloc[kr], self.nloc[kc]
        sc = 2 if kc == 2 else 1
        classes = _axis_classes(kr, N)
        for codes in itertools.product(range(len(classes)), repeat=d):
            combo = tuple(classes[c] for c in codes)
            st = {}
            for cells in itertools.product(*[c[1] for c in combo]):           # the cells containing the node
                a_loc = sum(cells[m][1] * nr ** m for m in range(d))
                for b in itertools.product(range(nc), repeat=d):              # their column nodes; b[m] = local coordinate
                    b_loc = sum(b[m] * nc ** m for m in range(d))
                    blk = K[a_loc * br:(a_loc + 1) * br, b_loc * bc:(b_loc + 1) * bc]
                    if not blk.any():
                        continue
                    delta = tuple(sc * cells[m][0] + b[m] for m in range(d))
                    st[delta] = st.get(delta, 0.0) + blk
            # contributions of neighbouring cells that cancel (mixed derivatives across a shared face) leave
            # round-off instead of the exact zero an element-by-element sum produces: flush it
            tiny = self.drop_tol * np.abs(K).max()
            for t in list(st):
                st[t] = np.where(np.abs(st[t]) < tiny, 0.0, st[t])
                if not st[t].any():
                    del st[t]
            deltas = sorted(st, key=lambda t: t[::-1])                        # ascending column id: last axis slowest
            vals = np.array([st[t] for t in deltas]).reshape(len(deltas), br, bc)
            yield codes, combo, deltas, vals

    def table_arrays(self, K: np.ndarray, kr: int, kc: int, br: int, bc: int):
        """The class tables in the layout of csrc/next/gen_stencil.cuh (BlockTable): class id = sum code_m * ncl^m
        with ncl = 4 (P2 rows) or 3 (P1 rows); (cls_ptr int32, off int64 column-node offsets, vals)."""
        d = self.dim
        ncl = 4 if kr == 2 else 3
        Lc = self.L[kc]
        per = {}
        for codes, combo, deltas, vals in self.class_stencils(K, kr, kc, br, bc):
            cid = sum(codes[m] * ncl ** m for m in range(d))
            per[cid] = (np.array([sum(t[m] * Lc ** m for m in range(d)) for t in deltas], np.int64), vals)
        cls_ptr = np.zeros(ncl ** d + 1, np.int32)
        offs, vs = [], []
        for c in range(ncl ** d):
            o, v = per.get(c, (np.zeros(0, np.int64), np.zeros((0, br, bc))))
            cls_ptr[c + 1] = cls_ptr[c] + len(o)
            offs.append(o); vs.append(v.reshape(len(o), br * bc))
        return cls_ptr, np.concatenate(offs), np.ascontiguousarray(np.concatenate(vs))

    # ---- one field block --------------------------------------------------------------------------------
    def expand(self, K: np.ndarray, kr: int, kc: int, br: int, bc: int, plane_range=None):
        """Global block from the macro-cell matrix K ((nloc_r^d * br) x (nloc_c^d * bc), dense).
        Returns a BSR matrix with blocks br x bc; with plane_range=(a, b) only the rows of the nodes on the
        planes a <= x_last < b of the ROW lattice are generated (global row and column numbering kept)."""
        d, N = self.dim, self.N
        Lr, Lc = self.L[kr], self.L[kc]
        nr, nc = self.nloc[kr], self.nloc[kc]
        sc = 2 if kc == 2 else 1
        n_rows = Lr ** d
        per_class = []
        counts = np.zeros(n_rows, np.int64)
        tables = []
        for codes, combo, deltas, vals in self.class_stencils(K, kr, kc, br, bc):
            tables.append((combo, deltas, vals))
            if not deltas:
                continue
            # ---- the rows of this class
            axes = [c[0] for c in combo]
            if plane_range is not None:
                axes = axes[:-1] + [axes[-1][(axes[-1] >= plane_range[0]) & (axes[-1] < plane_range[1])]]
            if any(len(a) == 0 for a in axes):
                continue
            X = np.stack(np.meshgrid(*axes[::-1], indexing="ij"), -1).reshape(-1, d)[:, ::-1]      # (rows, d), axis 0 = x
            rows = sum(X[:, m].astype(np.int64) * Lr ** m for m in range(d))
            c0 = X // 2 if kr == 2 else X
            base = sum(sc * c0[:, m].astype(np.int64) * Lc ** m for m in range(d))
            off = np.array([sum(t[m] * Lc ** m for m in range(d)) for t in deltas], np.int64)
            per_class.append((rows, base, off, vals))
            counts[rows] = len(deltas)
        indptr = np.zeros(n_rows + 1, np.int64)
        np.cumsum(counts, out=indptr[1:])
        nnzb = int(indptr[-1])
        indices = np.empty(nnzb, np.int32)
        data = np.empty((nnzb, br, bc))
        for rows, base, off, vals in per_class:
            dest = indptr[rows][:, None] + np.arange(len(off))[None, :]
            indices[dest] = (base[:, None] + off[None, :]).astype(np.int32)
            data[dest] = vals[None]
        self._last_tables = tables
        return sp.bsr_matrix((data, indices, indptr), shape=(n_rows * br, Lc ** d * bc), blocksize=(br, bc))

    def field_blocks(self, which: str, pc_type: str = "diagonal", plane_ranges=None):
        """The 3x3 dict of un-BC'd global field blocks as CSR.  plane_ranges = {'2': (a, b), '1': (a, b)} keeps
        only the rows of a z-slab (P2 planes for s and f, P1 planes for p)."""
        d = self.dim
        local = self.cell.field_blocks(which, pc_type)
        shape = {"22": (2, 2, d, d), "21": (2, 1, d, 1), "12": (1, 2, 1, d), "11": (1, 1, 1, 1)}
        out = {}
        for key, (kind, data) in local.items():
            K = self.cell._to_csr(kind, data).toarray()
            kr, kc, br, bc = shape[kind]
            pr = None if plane_ranges is None else plane_ranges[str(kr)]
            M = self.expand(K, kr, kc, br, bc, pr).tocsr()
            self.class_tables[(which, pc_type, key)] = self._last_tables
            out[key] = M
        return out

    # ---- right-hand side: the cell-face vector of a side, scattered to the boundary cells of that side ------
    def rhs(self, t, neumann_solid=(), neumann_fluid=(), fs_sur=None, ff_sur=None):
        d, N = self.dim, self.N
        ns = self.n2 * d
        b = np.zeros(2 * ns + self.n1)
        L2 = self.L[2]
        loc = np.stack(np.meshgrid(*[np.arange(3)] * d, indexing="ij"), -1).reshape(-1, d)[:, ::-1]   # local P2 lattice
        for field_off, sides, fun, kw in ((0, neumann_solid, fs_sur, "neumann_solid"),
                                          (ns, neumann_fluid, ff_sur, "neumann_fluid")):
            if fun is None or not sides:
                continue
            for side in sides:
                args = dict(neumann_solid=(), neumann_fluid=(), fs_sur=fs_sur, ff_sur=ff_sur)
                args[kw] = [side]
                bl = self.cell.rhs(t, **args)
                lo = 0 if kw == "neumann_solid" else 3 ** d * d                # fluid block of the one-cell vector
                face = bl[lo: lo + 3 ** d * d].reshape(3 ** d, d)                # per local P2 node
                ax = "xyz".index(side[0])
                cidx = [np.arange(N)] * d
                cidx[ax] = np.array([0 if side[1] == "0" else N - 1])
                C = np.stack(np.meshgrid(*cidx[::-1], indexing="ij"), -1).reshape(-1, d)[:, ::-1]     # boundary cells
                nodes = sum((2 * C[:, None, m] + loc[None, :, m]).astype(np.int64) * L2 ** m for m in range(d))
                idx = field_off + nodes[:, :, None] * d + np.arange(d)[None, None, :]
                np.add.at(b, idx.ravel(), np.broadcast_to(face[None], idx.shape).ravel())
        bc = np.concatenate([self.bc_s.ravel(), self.bc_f.ravel(), np.zeros(self.n1, bool)])
        b[bc] = 0.0
        return b

    # ---- the whole system ---------------------------------------------------------------------------------
    def compose(self, blocks: dict, apply_p_bc: bool = False) -> sp.csr_matrix:
        shell = _Shell(self)
        return PoroAssembler.compose(shell, blocks, apply_p_bc)

    def system(self, pc_type: str, t: float, neumann_solid=(), neumann_fluid=(), fs_sur=None, ff_sur=None) -> PoroSystem:
        d = self.dim
        ns = nf = self.n2 * d
        npp = self.n1
        three_way = "3-way" in pc_type
        A = self.compose(self.field_blocks("A"))
        P = self.compose(self.field_blocks("P", pc_type))
        P_diff = self.compose(self.field_blocks("P_diff", pc_type), apply_p_bc=True) if three_way else None
        b = self.rhs(t, neumann_solid, neumann_fluid, fs_sur, ff_sur)
        is_s = np.arange(ns, dtype=np.int64)
        is_f = ns + np.arange(nf, dtype=np.int64)
        is_p = ns + nf + np.arange(npp, dtype=np.int64)
        return PoroSystem(d, ns, nf, npp, A, P, P_diff, b, is_s, is_f, is_p, np.concatenate([is_f, is_p]),
                          np.flatnonzero(self.bc_p).astype(np.int64), np.repeat(self.coords(2), d, axis=0),
                          self.coords(1), pc_type)


class _Shell:
    """What PoroAssembler.compose reads from `self`, for blocks that are already global CSR matrices."""

    def __init__(self, g: StructuredGenerator):
        self.dim, self.n2, self.n1 = g.dim, g.n2, g.n1
        self.bc_s, self.bc_f, self.bc_p = g.bc_s, g.bc_f, g.bc_p

    def _to_csr(self, M):
        return M.copy()


def swelling_generator(dim: int, N: int, overrides: dict | None = None):
    """The swelling problems of hostfem.problems with the structured generator (same BCs and loads)."""
    from .problems import _traction, swelling_params
    par = swelling_params(dim)
    if overrides:
        par.update(overrides)
    gen = StructuredGenerator(dim, N, 1e-2, par)
    if dim == 2:
        gen.set_bcs(bcs_s=[("x0", 0), ("y0", 1)], bcs_f=[("y1", None), ("y0", None)], bcs_p=["x0", "y1", "x1"])
        loads = dict(neumann_solid=["y1", "x1"], neumann_fluid=["x0"])
    else:
        gen.set_bcs(bcs_s=[("x0", 0), ("y0", 1), ("z0", 2)], bcs_f=[("z0", None), ("z1", None)],
                    bcs_p=["x0", "x1", "y0", "y1", "z1"])
        loads = dict(neumann_solid=["x1", "y1", "z1"], neumann_fluid=["x0", "y0"])
    loads.update(fs_sur=_traction(0.9), ff_sur=_traction(0.1))
    return gen, par, loads


def swelling(dim: int, N: int = 10, pc_type: str | None = None, overrides: dict | None = None):
    gen, par, loads = swelling_generator(dim, N, overrides)
    if pc_type is not None:
        par["pc type"] = pc_type
    sys_ = gen.system(par["pc type"], par["t0"] + par["dt"], **loads)
    sys_.meta.update(dict(problem="swelling-%dd" % dim, N=N, generator="stencil"))
    return sys_, par
