"""CPU: the C-ABI library loads and exports every symbol include/poro.h declares; host-side logic
(options grammar, IndexSet remap) mirrors the reference."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "poro.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(poro_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from poro_b200 import _capi
    lib = _capi.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    # and every bound prototype is declared in the header
    for n in _capi.SIGNATURES:
        assert n in names, "binding without declaration: " + n


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from poro_b200 import _capi
    with pytest.raises(_capi.PoroError, match="no CPU fallback"):
        _capi.Context(0)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "poroelasticity-linear-solvers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_options_grammar_matches_reference_parser():
    from poro_b200.lib.Parser import parse_petsc_options
    text = """
    # comment
    -global_ksp_type gmres
    -s_ksp_rtol   1e-1
    -fp_ksp_gmres_modifiedgramschmidt
    -p_pc_type hypre # trailing comment makes the reference skip the WHOLE line
    -a b c
    """
    opts = dict(parse_petsc_options(text))
    assert opts["-global_ksp_type"] == "gmres"
    assert opts["-s_ksp_rtol"] == "1e-1"                       # several spaces: first and last token
    assert opts["-fp_ksp_gmres_modifiedgramschmidt"] is None   # bare flag
    assert "-p_pc_type" not in opts                             # line contains '#'
    assert opts["-a"] == "c"


def test_index_set_two_way_remap_matches_reference_loop():
    from poro_b200.lib.IndexSet import IndexSet
    rng = np.random.default_rng(0)
    n = 60
    perm = rng.permutation(n)
    s, f, p = np.sort(perm[:25]), np.sort(perm[25:50]), np.sort(perm[50:])
    imap = IndexSet(s, f, p, two_way=True)
    # the reference's membership loop (lib/IndexSet.py:10-26), restated literally
    fp = sorted(list(f) + list(p))
    ref_f = [i for i, d in enumerate(fp) if d in set(f)]
    ref_p = [i for i, d in enumerate(fp) if d in set(p)]
    assert list(imap.is_f) == ref_f and list(imap.is_p) == ref_p
    assert list(imap.is_fp) == fp and imap.get_dimensions() == (25, 25, 10)
    # 3-way keeps global numbering
    im3 = IndexSet(s, f, p, two_way=False)
    assert list(im3.is_f) == list(f)


def test_preconditioner_rejects_unknown_pc_type():
    from poro_b200.lib.Preconditioner import Preconditioner
    par = {"pc type": "bogus", "inner ksp type": "gmres", "inner pc type": "hypre", "inner rtol": 0, "inner atol": 0,
           "inner maxiter": 1, "inner accel order": 0, "inner monitor": False}
    with pytest.raises(SystemExit):
        Preconditioner(None, None, None, None, par, [])


def test_shipped_option_files_parse_and_match_bench():
    from poro_b200.lib.Parser import parse_petsc_options
    import bench
    for name in ("petsc-options-exact", "petsc-options-inexact", "petsc-options-b200"):
        opts = parse_petsc_options(open(os.path.join(ROOT, "options", name)).read())
        assert len(opts) >= 10 and all(k.startswith("-") for k, _ in opts)
    shipped = dict(parse_petsc_options(open(os.path.join(ROOT, "options", "petsc-options-b200")).read()))
    assert shipped == dict(parse_petsc_options(bench.BENCH_OPTIONS))
