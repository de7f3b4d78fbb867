"""IndexSet: field index sets of the mixed space (reference lib/IndexSet.py:29-67).

The reference derives them from dolfin dofmaps; here they are given as arrays.  In 2-way
mode the reference re-expresses `is_f` / `is_p` as positions INSIDE the fp sub-space
(lib/IndexSet.py:43-54, an O(n^2) numba membership loop); the same remap is done with
np.searchsorted.
"""
from __future__ import annotations

from time import perf_counter as time

import numpy as np

from .Printing import parprint


def get_local_fp_dofs(dofs_fp_global, dofmap_f, dofmap_p):
    """Positions of the f / p dofs inside the sorted fp list (lib/IndexSet.py:10-26)."""
    fp = np.asarray(dofs_fp_global)
    order = np.argsort(fp, kind="stable")
    pos_f = order[np.searchsorted(fp, dofmap_f, sorter=order)]
    pos_p = order[np.searchsorted(fp, dofmap_p, sorter=order)]
    return np.sort(pos_f), np.sort(pos_p)


class IndexSet:
    def __init__(self, dofmap_s, dofmap_f, dofmap_p, two_way=True, block_dim=0, coords_s=None, coords_p=None):
        t0 = time()
        self.dofmap_s = np.asarray(dofmap_s, dtype=np.int64)
        f = np.asarray(dofmap_f, dtype=np.int64)
        p = np.asarray(dofmap_p, dtype=np.int64)
        self.ns, self.nf, self.np = len(self.dofmap_s), len(f), len(p)
        self.dofmap_fp = np.sort(np.concatenate([f, p]))         # lib/IndexSet.py:37
        self.two_way = bool(two_way)
        if two_way:
            f, p = get_local_fp_dofs(self.dofmap_fp, f, p)       # lib/IndexSet.py:43-54
        self.dofmap_f, self.dofmap_p = f, p
        self.is_s, self.is_f, self.is_p, self.is_fp = self.dofmap_s, self.dofmap_f, self.dofmap_p, self.dofmap_fp
        self.block_dim = block_dim
        self.coords_s, self.coords_p = coords_s, coords_p
        parprint("---- [Indexes] computed local indices in {:.3f}s".format(time() - t0))

    def get_dimensions(self):
        return self.ns, self.nf, self.np

    def get_index_sets(self):
        return self.is_s, self.is_f, self.is_p, self.is_fp

    def install(self, ctx):
        """Hand the index sets (and optional coordinates) to the device library."""
        ctx.set_fields(self.is_s, self.is_f, self.is_p, self.is_fp, two_way_local_fp=self.two_way,
                       block_dim=self.block_dim)
        if self.coords_s is not None:
            dim = np.asarray(self.coords_s).shape[1]
            ctx.set_coords(dim, self.coords_s, self.coords_p)
