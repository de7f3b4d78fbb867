"""The row-wise structured generator (hostfem/stencil.py, blueprint of the device-side matrix generator) against
the generic element-by-element assembler: same matrices, pattern, right-hand side and boundary data."""
import numpy as np
import pytest

from hostfem import problems, stencil


def _rel(X, Y):
    return abs(X - Y).max() / abs(X).max()


@pytest.mark.parametrize("dim,N,pc_type", [(2, 1, "diagonal"), (2, 5, "diagonal"), (2, 4, "undrained 3-way"),
                                           (3, 1, "diagonal"), (3, 3, "diagonal 3-way"), (3, 4, "undrained")])
def test_generator_equals_assembler(dim, N, pc_type):
    ref, _ = problems.swelling(dim, N, pc_type)
    gen, _ = stencil.swelling(dim, N, pc_type)
    for name in ("A", "P", "P_diff"):
        X, Y = getattr(ref, name), getattr(gen, name)
        if X is None:
            assert Y is None
            continue
        assert X.shape == Y.shape and X.nnz == Y.nnz
        assert np.array_equal(X.indptr, Y.indptr) and np.array_equal(X.indices, Y.indices)
        assert _rel(X, Y) < 1e-14
    assert np.abs(ref.b - gen.b).max() <= 1e-14 * np.abs(ref.b).max()
    for name in ("is_s", "is_f", "is_p", "is_fp", "bcs_sub_pressure"):
        assert np.array_equal(getattr(ref, name), getattr(gen, name))
    assert np.abs(ref.coords_s - gen.coords_s).max() < 1e-15 and np.abs(ref.coords_p - gen.coords_p).max() < 1e-15


def test_slab_rows_are_the_rows_of_the_full_block():
    """plane_ranges generates only a z-slab of rows (what one rank needs), in global numbering."""
    gen, par, _ = stencil.swelling_generator(3, 3)
    full = gen.field_blocks("A")
    L2, L1 = gen.L[2], gen.L[1]
    slab = gen.field_blocks("A", plane_ranges={"2": (2, 5), "1": (1, 3)})
    rows2 = np.flatnonzero((np.arange(gen.n2) // (L2 * L2) >= 2) & (np.arange(gen.n2) // (L2 * L2) < 5))
    rows1 = np.flatnonzero((np.arange(gen.n1) // (L1 * L1) >= 1) & (np.arange(gen.n1) // (L1 * L1) < 3))
    for key, M in full.items():
        S = slab[key]
        if key[0] in "sf":
            rows = (3 * rows2[:, None] + np.arange(3)).ravel()
        else:
            rows = rows1
        other = np.setdiff1d(np.arange(M.shape[0]), rows)
        assert abs(S[rows] - M[rows]).max() == 0
        assert S[other].nnz == 0


def test_class_tables_are_small():
    """What the device kernel keeps resident: at most 4^d (P2 rows) / 3^d (P1 rows) stencils per block."""
    gen, par, _ = stencil.swelling_generator(3, 4)
    gen.field_blocks("A")
    t_ss = gen.class_tables[("A", "diagonal", "ss")]
    t_pp = gen.class_tables[("A", "diagonal", "pp")]
    assert len(t_ss) == 64 and len(t_pp) == 27
    assert max(len(d) for _, d, _ in t_ss) <= 125 and max(len(d) for _, d, _ in t_pp) <= 27
    interior_vertex = [d for combo, d, _ in t_ss if all(len(c[1]) == 2 for c in combo)][0]
    assert len(interior_vertex) == 65          # a P2 vertex node of the 6-tet cube couples to 65 nodes
