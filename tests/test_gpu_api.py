"""GPU: the reference-shaped API surface beyond Solver.solve -- PreconditionerCC.apply, the host-buffer
solve, extracted blocks (createSubMatrix parity), timings, options from a file, error behaviour."""
import os

import numpy as np
import pytest

from helpers import EXACT_OPTIONS, gpu_solve, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pc_type", ["diagonal", "diagonal 3-way", "undrained"])
def test_pc_apply_matches_oracle_block_pc(gpu_ctx, pc_type):
    """y = M^-1 x through PreconditionerCC.apply(pc, x, y) (lib/Preconditioner.py:141-250) with exact blocks."""
    from oracle.blockpc import BlockPC, exact_solvers
    from oracle.problems import swelling
    from poro_b200.lib.backend import DeviceVector
    sys_, par = swelling(2, 8, pc_type)
    g = gpu_solve(sys_, par, EXACT_OPTIONS, return_objects=True)
    pcx = g["pc"].pc.getPythonContext()
    x = np.random.default_rng(2).standard_normal(sys_.n)
    dx, dy = DeviceVector(x, ctx=gpu_ctx), DeviceVector(n=sys_.n, ctx=gpu_ctx)
    pcx.apply(None, dx, dy)
    gpu_ctx.sync()
    yo = BlockPC(sys_, exact_solvers())(x)
    assert rel(dy.numpy(), yo) <= (1e-7 if "undrained" in pc_type else 1e-10)
    g["pc"].print_timings()
    g["solver"].print_timings()


def test_extracted_blocks_match_scipy_submatrices(gpu_ctx):
    """MatCreateSubMatrix parity (lib/Preconditioner.py:60-75), also under a shuffled numbering."""
    from oracle.blockpc import submatrix
    from oracle.problems import swelling
    sys_, par = swelling(2, 6, "diagonal 3-way")
    perm = np.random.default_rng(4).permutation(sys_.n)
    g = gpu_solve(sys_, par, EXACT_OPTIONS, perm=perm, return_objects=True)
    pcx = g["pc"].pc.getPythonContext()
    s, f, p = sys_.is_s, sys_.is_f, sys_.is_p
    for name, rows, cols, M in (("ss", s, s, sys_.P), ("sf", s, f, sys_.P), ("sp", s, p, sys_.P), ("ff", f, f, sys_.P),
                                ("fp", f, p, sys_.P), ("pp", p, p, sys_.P), ("diff", p, p, sys_.P_diff)):
        ref = submatrix(M, rows, cols)
        got = pcx.block(name)
        assert got.shape == ref.shape
        assert abs(got - ref).max() == 0.0, name


def test_solve_host_and_history(gpu_ctx):
    import torch
    from oracle.problems import swelling
    sys_, par = swelling(2, 8, "diagonal")
    g = gpu_solve(sys_, par, EXACT_OPTIONS, return_objects=True)
    ksp = g["solver"].solver
    b = torch.from_numpy(sys_.b.copy()).pin_memory()
    x = torch.zeros_like(b).pin_memory()
    ksp.solve_host(b, x)
    assert ksp.reason in (2, 3)
    assert rel(x.numpy(), g["x"]) <= 1e-12
    h = ksp.getConvergenceHistory()
    assert len(h) == ksp.its + 1 and h[-1] == pytest.approx(ksp.rnorm)
    # plain numpy host buffers work too
    xb = np.zeros_like(sys_.b)
    ksp.solve_host(sys_.b, xb)
    assert rel(xb, g["x"]) <= 1e-12


def test_options_file_and_maxit_reason(gpu_ctx, tmp_path):
    from oracle.problems import swelling
    from poro_b200.lib.Parser import load_petsc_options
    p = tmp_path / "opts"
    p.write_text(EXACT_OPTIONS + "\n# -global_ksp_type cg\n-global_ksp_max_it 3\n")
    sys_, par = swelling(2, 6, "diagonal")
    g = gpu_solve(sys_, par, open(p).read())
    assert g["its"] == 3 and g["reason"] == -3          # non-convergence is a reason code, not an error
    opts = dict(load_petsc_options(gpu_ctx, str(p)))
    assert opts["-global_ksp_max_it"] == "3" and "-global_ksp_type" in opts and opts["-global_ksp_type"] == "gmres"


def test_shape_mismatch_is_an_error(gpu_ctx):
    from oracle.problems import swelling
    from poro_b200._capi import PoroError
    from poro_b200.lib.backend import DeviceMatrix
    from poro_b200.lib.IndexSet import IndexSet
    from poro_b200.lib.Preconditioner import Preconditioner
    sys_, par = swelling(2, 4, "diagonal")
    other, _ = swelling(2, 6, "diagonal")
    imap = IndexSet(sys_.is_s, sys_.is_f, sys_.is_p, two_way=True)
    with pytest.raises(PoroError):
        Preconditioner(imap, DeviceMatrix(other.A, gpu_ctx), DeviceMatrix(other.P, gpu_ctx), None, par, []).get_pc()
