"""The row logic of the device-side matrix generator (csrc/gen_stencil.cuh, __host__ __device__; used by csrc/gen.cu) run
on the CPU through a g++ harness, against hostfem/stencil.py and the element assembler.  The kernels themselves are tested on
the GPU in tests/test_generator.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from hostfem import problems, stencil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NEXT = os.path.join(ROOT, "poroelasticity-linear-solvers_b200", "csrc")
SHAPE = {"22": (2, 2), "21": (2, 1), "12": (1, 2), "11": (1, 1)}


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gen") / "gen_harness.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + NEXT,
                    os.path.join(ROOT, "tests", "gen_stencil_harness.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    lib.gen_host_counts.restype = C.c_int64
    return lib


def _generate(lib, gen, K, kr, kc, br, bc, diag, bc_row=None, node_range=None):
    cls_ptr, off, vals = gen.table_arrays(K, kr, kc, br, bc)
    n_nodes = gen.L[kr] ** gen.dim
    node0, nrows = (0, n_nodes) if node_range is None else (node_range[0], node_range[1] - node_range[0])
    rowptr = np.zeros(nrows + 1, np.int64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    head = [gen.dim, gen.N, kr, kc, br, bc, diag, vp(cls_ptr), vp(off), vp(vals), C.c_int64(node0), C.c_int64(nrows)]
    nnzb = lib.gen_host_counts(*head, vp(rowptr))
    col = np.full(nnzb, -1, np.int32)
    val = np.full((nnzb, br, bc), np.nan)
    flags = None if bc_row is None else np.ascontiguousarray(bc_row, dtype=np.uint8)
    lib.gen_host_fill(*head, vp(rowptr), None if flags is None else vp(flags), vp(col), vp(val))
    assert not np.isnan(val).any() and (col >= 0).all()
    return sp.bsr_matrix((val, col, rowptr), shape=(nrows * br, gen.L[kc] ** gen.dim * bc), blocksize=(br, bc))


@pytest.mark.parametrize("dim,N", [(2, 1), (2, 4), (3, 1), (3, 3)])
def test_row_logic_equals_numpy_generator(harness, dim, N):
    gen, par, _ = stencil.swelling_generator(dim, N)
    for key, (kind, data) in gen.cell.field_blocks("A").items():
        K = gen.cell._to_csr(kind, data).toarray()
        kr, kc = SHAPE[kind]
        br, bc = (dim if kr == 2 else 1), (dim if kc == 2 else 1)
        ref = gen.expand(K, kr, kc, br, bc)
        got = _generate(harness, gen, K, kr, kc, br, bc, int(key[0] == key[1]))
        assert np.array_equal(ref.indptr, got.indptr) and np.array_equal(ref.indices, got.indices)
        assert np.array_equal(ref.data, got.data)


def test_dirichlet_rows_and_node_ranges(harness):
    """With the BC flags the generated blocks are the blocks of the assembled A (DirichletBC.apply semantics:
    zero row, unit diagonal, columns kept); a node range generates exactly those rows."""
    dim, N = 3, 3
    gen, par, _ = stencil.swelling_generator(dim, N)
    ref, _ = problems.swelling(dim, N, "diagonal")
    masks = {"s": gen.bc_s.ravel(), "f": gen.bc_f.ravel(), "p": np.zeros(gen.n1, bool)}     # p BCs only go to P_diff
    sets = {"s": ref.is_s, "f": ref.is_f, "p": ref.is_p}
    for key, (kind, data) in gen.cell.field_blocks("A").items():
        K = gen.cell._to_csr(kind, data).toarray()
        kr, kc = SHAPE[kind]
        br, bc = (dim if kr == 2 else 1), (dim if kc == 2 else 1)
        got = _generate(harness, gen, K, kr, kc, br, bc, int(key[0] == key[1]), masks[key[0]]).tocsr()
        got.eliminate_zeros()
        want = ref.A[sets[key[0]]][:, sets[key[1]]].tocsr()
        assert got.nnz == want.nnz
        assert abs(got - want).max() <= 1e-14 * max(abs(want).max(), 1e-300)
        n_nodes = gen.L[kr] ** dim
        a, b = n_nodes // 3, 2 * n_nodes // 3
        part = _generate(harness, gen, K, kr, kc, br, bc, int(key[0] == key[1]), masks[key[0]], (a, b)).tocsr()
        part.eliminate_zeros()
        assert abs(part - want[a * br: b * br]).max() <= 1e-14 * max(abs(want).max(), 1e-300)
