"""Row partition of the assembled system over ranks (one rank per GPU) and its halo plan.

The reference runs row-partitioned under MPI (PETSc MPIAIJ + ParMETIS, swelling-3d.py:7,
paper-scripts/robustness_2d.sh:29); every MatMult does a VecScatter of ghost entries.  Here the
structured cube is cut into contiguous z-slabs of P2-node planes; each rank assembles only the
cells that touch its planes, keeps its rows, and renumbers columns into the local extended
vector [owned | halo].  Halo entries are ordered neighbour-major and, inside one neighbour, in
the owner's local order, which is what poro_halo_set expects.

Host logic only (numpy + torch.distributed for the one set-up exchange); testable with the
gloo backend on CPU.  The data-path collectives (halo send/recv, all-reduce) run inside
libporo.so over NCCL.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


def slab_ranges(n_planes: int, world: int):
    """Contiguous ranges [a, b) of fine-lattice planes per rank, as even as possible."""
    base, rem = divmod(n_planes, world)
    out, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < rem else 0)
        out.append((a, b))
        a = b
    return out


@dataclass
class HaloPlan:
    n_owned: int
    neigh: np.ndarray          # neighbour ranks, ascending
    send_ptr: np.ndarray       # (nneigh+1,)
    send_idx: np.ndarray       # owned local indices to send, concatenated per neighbour
    recv_count: np.ndarray     # (nneigh,)
    halo_global: np.ndarray    # global ids of the halo entries in extended order

    @property
    def n_halo(self):
        return int(self.recv_count.sum())


def build_halo_plan(owned_global: np.ndarray, needed_global: np.ndarray, owner_of, rank: int, world: int,
                    all_gather_object) -> HaloPlan:
    """owned_global: global ids of the owned dofs in LOCAL order.  needed_global: global ids of the
    off-rank columns this rank reads.  owner_of(global ids) -> ranks.  all_gather_object(obj) ->
    list over ranks (torch.distributed.all_gather_object wrapper, or a fake for 1 process)."""
    needed_global = np.unique(needed_global)
    owners = owner_of(needed_global) if len(needed_global) else np.zeros(0, np.int64)
    requests = {int(r): needed_global[owners == r] for r in np.unique(owners)}
    everyone = all_gather_object(requests)             # everyone[q][r] = ids rank q needs from rank r
    order = np.argsort(owned_global, kind="stable")
    sorted_owned = owned_global[order]

    def local_index(gids):
        pos = np.searchsorted(sorted_owned, gids)
        assert np.all(sorted_owned[pos] == gids), "request for a dof this rank does not own"
        return order[pos]

    neigh = sorted(set(requests.keys()) | {q for q in range(world) if q != rank and rank in everyone[q] and len(everyone[q][rank])})
    send_ptr, send_idx, recv_count, halo_global = [0], [], [], []
    for q in neigh:
        want = everyone[q].get(rank, np.zeros(0, np.int64)) if q != rank else np.zeros(0, np.int64)
        li = np.sort(local_index(np.asarray(want, dtype=np.int64))) if len(want) else np.zeros(0, np.int64)
        send_idx.append(li)
        send_ptr.append(send_ptr[-1] + len(li))
        # what I receive from q arrives in q's local order; q's local order of my request list is the
        # order of q's local indices, which q computes exactly as above.  I need the matching global ids:
        mine = requests.get(q, np.zeros(0, np.int64))
        recv_count.append(len(mine))
        halo_global.append(mine)                          # re-ordered below once q's local order is known
    plan = HaloPlan(len(owned_global), np.asarray(neigh, np.int32), np.asarray(send_ptr, np.int64),
                    np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, np.int32),
                    np.asarray(recv_count, np.int64),
                    np.concatenate(halo_global).astype(np.int64) if halo_global else np.zeros(0, np.int64))
    # second exchange: the owner tells each requester the global ids in the order it will send them
    sent_order = {int(q): owned_global[plan.send_idx[plan.send_ptr[k]:plan.send_ptr[k + 1]]] for k, q in enumerate(neigh)}
    everyone2 = all_gather_object(sent_order)
    hg = [np.asarray(everyone2[int(q)].get(rank, np.zeros(0, np.int64)), dtype=np.int64) for q in neigh]
    plan.halo_global = np.concatenate(hg) if hg else np.zeros(0, np.int64)
    assert len(plan.halo_global) == plan.n_halo
    return plan


def localize_columns(M_rows: sp.csr_matrix, owned_global: np.ndarray, halo_global: np.ndarray) -> sp.csr_matrix:
    """Rows already restricted to the owned dofs; map global column ids to [owned | halo] positions."""
    n_owned, n_halo = len(owned_global), len(halo_global)
    allg = np.concatenate([owned_global, halo_global])
    order = np.argsort(allg, kind="stable")
    sg = allg[order]
    pos = np.searchsorted(sg, M_rows.indices)
    assert np.all(sg[pos] == M_rows.indices), "matrix references a column that is neither owned nor in the halo"
    out = sp.csr_matrix((M_rows.data, order[pos].astype(np.int32), M_rows.indptr), shape=(M_rows.shape[0], n_owned + n_halo))
    out.sort_indices()
    return out


@dataclass
class LocalSystem:
    dim: int
    A: sp.csr_matrix
    P: sp.csr_matrix
    P_diff: sp.csr_matrix | None
    b: np.ndarray
    is_s: np.ndarray
    is_f: np.ndarray
    is_p: np.ndarray
    bcs_sub_pressure: np.ndarray
    coords_s: np.ndarray
    coords_p: np.ndarray
    owned_global: np.ndarray
    plan: HaloPlan
    meta: dict = field(default_factory=dict)
    # optional (overlap_rows=True): P with the rows of the halo dofs appended, (n_owned + n_halo) square, columns outside
    # [owned | halo] dropped -- the input a halo-aware Schur complement / overlapping hierarchy needs (round-2 plan)
    P_ext: sp.csr_matrix | None = None


@dataclass
class DistributedProblem:
    sys: LocalSystem
    par: dict
    n_global: int
    rank: int
    world: int

    def index_set(self):
        from .lib.IndexSet import IndexSet
        s = self.sys
        return IndexSet(s.is_s, s.is_f, s.is_p, two_way=True, block_dim=s.dim, coords_s=s.coords_s, coords_p=s.coords_p)


def _gather_fn(world):
    if world == 1:
        return lambda obj: [obj]
    import torch.distributed as dist

    def g(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    return g


def restrict_rows_to_columns(M_rows: sp.csr_matrix, keep_global: np.ndarray, n_cols_global: int) -> sp.csr_matrix:
    """Drop the entries of M_rows whose (global) column is not in keep_global."""
    mask = np.zeros(n_cols_global, bool)
    mask[keep_global] = True
    C = M_rows.tocoo()
    sel = mask[C.col]
    return sp.csr_matrix((C.data[sel], (C.row[sel], C.col[sel])), shape=M_rows.shape)


def distributed_problem(dim: int, N: int, pc_type: str, rank: int, world: int, ctx=None, overrides=None,
                        overlap_rows: bool = False) -> DistributedProblem:
    """Assemble this rank's z-slab of the 3D swelling problem and install the halo plan on `ctx`."""
    assert dim == 3, "slab partition is implemented for the structured cube"
    from hostfem.fem import PoroAssembler, unit_cube_mesh    # host assembler (stands in for FEniCS; not the oracle)
    from hostfem.problems import _traction, swelling_params
    par = swelling_params(3)
    if overrides:
        par.update(overrides)
    par["pc type"] = pc_type
    L = 2 * N + 1
    a, b_ = slab_ranges(L, world)[rank]
    k0 = max(0, (a - 1) // 2 if a > 0 else 0)
    k1 = min(N, (b_ - 1) // 2 + 1)
    if overlap_rows:            # one more cell layer: the rows of the halo dofs become complete as well
        k0, k1 = max(0, k0 - 1), min(N, k1 + 1)
    mesh = unit_cube_mesh(N, 1e-2, k_range=(k0, k1))
    asm = PoroAssembler(mesh, par)
    asm.set_bcs(bcs_s=[("x0", 0), ("y0", 1), ("z0", 2)], bcs_f=[("z0", None), ("z1", None)],
                bcs_p=["x0", "x1", "y0", "y1", "z1"])
    n2, n1 = asm.n2, asm.n1
    ns = nf = 3 * n2
    plane2 = np.arange(n2) // (L * L)                      # fine plane of each P2 node
    plane1 = 2 * (np.arange(n1) // ((N + 1) * (N + 1)))    # fine plane of each P1 vertex
    own2 = np.flatnonzero((plane2 >= a) & (plane2 < b_))
    own1 = np.flatnonzero((plane1 >= a) & (plane1 < b_))
    gs = (3 * own2[:, None] + np.arange(3)[None, :]).ravel()
    owned_global = np.concatenate([gs, ns + gs, ns + nf + own1]).astype(np.int64)
    ranges = slab_ranges(L, world)
    starts = np.array([r[0] for r in ranges])

    def owner_of(g):
        g = np.asarray(g)
        plane = np.where(g < ns + nf, ((g % ns) // 3) // (L * L), 2 * ((g - ns - nf) // ((N + 1) * (N + 1))))
        return np.searchsorted(starts, plane, side="right") - 1

    t = par["t0"] + par["dt"]
    A = asm.compose(asm.field_blocks("A"))[owned_global]
    P_full = asm.compose(asm.field_blocks("P", pc_type))
    P = P_full[owned_global]
    three_way = "3-way" in pc_type
    Pd = asm.compose(asm.field_blocks("P_diff", pc_type), apply_p_bc=True)[owned_global] if three_way else None
    bfull = asm.rhs(t, ["x1", "y1", "z1"], ["x0", "y0"], _traction(0.9), _traction(0.1))
    cols = np.unique(np.concatenate([M.indices for M in (A, P) + ((Pd,) if Pd is not None else ())]))
    is_owned = np.zeros(ns + nf + n1, bool)
    is_owned[owned_global] = True
    needed = cols[~is_owned[cols]]
    plan = build_halo_plan(owned_global, needed, owner_of, rank, world, _gather_fn(world))
    A_l = localize_columns(A.tocsr(), owned_global, plan.halo_global)
    P_l = localize_columns(P.tocsr(), owned_global, plan.halo_global)
    Pd_l = localize_columns(Pd.tocsr(), owned_global, plan.halo_global) if Pd is not None else None
    n_own = len(owned_global)
    ext_global = np.concatenate([owned_global, plan.halo_global])
    fld = np.where(ext_global < ns, 0, np.where(ext_global < ns + nf, 1, 2))
    ext_idx = np.arange(len(ext_global))
    is_s, is_f, is_p = ext_idx[fld == 0], ext_idx[fld == 1], ext_idx[fld == 2]
    coords_s = np.repeat(asm.p2_coords[own2], 3, axis=0)
    coords_p = mesh.coords[own1]
    bc_p_local = np.flatnonzero(asm.bc_p[own1]).astype(np.int64)
    sys_ = LocalSystem(3, A_l, P_l, Pd_l, bfull[owned_global], is_s, is_f, is_p, bc_p_local, coords_s, coords_p,
                       owned_global, plan, dict(N=N, planes=(a, b_), cells=(k0, k1)))
    if overlap_rows:
        halo_rows = restrict_rows_to_columns(P_full[plan.halo_global], ext_global, ns + nf + n1)
        halo_l = localize_columns(halo_rows, owned_global, plan.halo_global)
        sys_.P_ext = sp.vstack([P_l, halo_l], format="csr")
    if ctx is not None:
        if world > 1:
            import torch
            import torch.distributed as dist
            from . import _capi
            uid = [_capi.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.init_dist(rank, world, uid[0])
        ctx.set_halo(n_own, plan.neigh, plan.send_ptr, plan.send_idx, plan.recv_count)
    return DistributedProblem(sys_, par, ns + nf + n1, rank, world)
