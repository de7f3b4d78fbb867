"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle reproduces them.  GPU: the CUDA path reproduces them without importing the oracle."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import EXACT_OPTIONS, rel

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _csr(g, p):
    if p + "_data" not in g:
        return None
    return sp.csr_matrix((g[p + "_data"], g[p + "_indices"], g[p + "_indptr"]), shape=tuple(g[p + "_shape"]))


class _Sys:
    pass


def _load(path):
    g = np.load(path)
    s = _Sys()
    s.A, s.P, s.P_diff, s.b = _csr(g, "A"), _csr(g, "P"), _csr(g, "Pd"), g["b"]
    s.is_s, s.is_f, s.is_p = g["is_s"], g["is_f"], g["is_p"]
    s.is_fp = np.concatenate([s.is_f, s.is_p])
    s.bcs_sub_pressure, s.coords_s, s.coords_p = g["bcs_sub_pressure"], g["coords_s"], g["coords_p"]
    s.dim, s.pc_type = int(g["dim"]), str(g["pc_type"])
    s.ns, s.nf, s.np_ = len(s.is_s), len(s.is_f), len(s.is_p)
    s.n = len(s.b)
    par = {"solver rtol": float(g["rtol"]), "solver atol": float(g["atol"]), "solver maxiter": int(g["maxiter"]),
           "solver monitor": False, "solver type": "gmres", "pc type": s.pc_type, "inner ksp type": "gmres",
           "inner pc type": "hypre", "inner rtol": 1e-6, "inner atol": 0, "inner maxiter": 1000, "inner monitor": False,
           "inner accel order": 0, "AAR order": 10, "AAR p": 5, "AAR omega": 1, "AAR beta": 1}
    return g, s, par


def test_fixtures_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    from oracle.blockpc import BlockPC, exact_solvers
    from oracle.krylov import gmres
    g, s, par = _load(path)
    r = gmres(lambda v: s.A @ v, s.b, BlockPC(s, exact_solvers()), rtol=par["solver rtol"], atol=par["solver atol"],
              dtol=1e20, max_it=par["solver maxiter"], restart=par["solver maxiter"], pc_side="right")
    assert r.its == int(g["gmres_its"])
    assert rel(r.x, g["gmres_x"]) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_gpu_gmres_reproduces_golden(gpu_ctx, path):
    from helpers import gpu_solve
    g, s, par = _load(path)
    out = gpu_solve(s, par, EXACT_OPTIONS)
    loose = "undrained" in s.pc_type
    assert out["its"] == int(g["gmres_its"])
    np.testing.assert_allclose(out["history"], g["gmres_history"], rtol=5e-3 if loose else 1e-6, atol=1e-14)
    assert rel(out["x"], g["gmres_x"]) <= 1e-8          # `undrained` too: the dense exact blocks are refined once (PCDense)


@pytest.mark.gpu
@pytest.mark.parametrize("path", [f for f in FILES if "undrained" not in f], ids=lambda f: os.path.basename(f)[:-4])
def test_gpu_aar_reproduces_golden(gpu_ctx, path):
    from helpers import gpu_solve
    g, s, par = _load(path)
    out = gpu_solve(s, par, EXACT_OPTIONS, solver_type="aar")
    its = int(g["aar_its"])
    assert abs(out["its"] - its) <= max(2, its // 10)
    # both are stopped by the same 1e-6 test on an ill-scaled system: compare loosely, residuals tightly
    res = np.linalg.norm(s.b - s.A @ out["x"]) / np.linalg.norm(s.b)
    res_g = np.linalg.norm(s.b - s.A @ g["aar_x"]) / np.linalg.norm(s.b)
    assert res <= max(10 * res_g, 1e-5)
