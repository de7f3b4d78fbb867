"""CPU: the SA-AMG restatement (oracle/amg.py): deterministic aggregation, near-nullspace
reproduction, mesh-independent CG convergence on the diagonal blocks."""
import numpy as np
import pytest

from oracle.amg import SAAMG, aggregate_mis2, rigid_body_modes, strength_graph, tentative_prolongator
from oracle.blockpc import submatrix
from oracle.krylov import cg
from oracle.problems import swelling


@pytest.fixture(scope="module")
def blocks():
    s, _ = swelling(2, 12, "diagonal")
    return s, submatrix(s.P, s.is_s, s.is_s), submatrix(s.P, s.is_p, s.is_p)


def test_aggregation_is_deterministic_and_covers(blocks):
    s, Pss, _ = blocks
    S = strength_graph(Pss, 2, 0.08)
    a1, n1 = aggregate_mis2(S)
    a2, n2 = aggregate_mis2(S)
    assert n1 == n2 and np.array_equal(a1, a2)
    assert (a1 >= 0).mean() > 0.95 and a1.max() == n1 - 1
    # roots are pairwise at distance >= 3: no two aggregates' roots adjacent
    counts = np.bincount(a1[a1 >= 0])
    assert counts.min() >= 1 and counts.max() < 200


def test_tentative_prolongator_reproduces_nullspace(blocks):
    s, Pss, _ = blocks
    B = rigid_body_modes(s.coords_s, 2)
    S = strength_graph(Pss, 2, 0.08)
    agg, na = aggregate_mis2(S)
    T, Bc = tentative_prolongator(agg, na, 2, B)
    member = np.repeat(agg >= 0, 2)
    assert np.abs(T @ Bc - B)[member].max() < 1e-12          # B = T Bc on aggregated rows
    G = (T.T @ T).toarray()
    assert np.abs(G - np.diag(np.diag(G))).max() < 1e-12     # orthonormal columns per aggregate


@pytest.mark.parametrize("N", [8, 16])
def test_cg_amg_converges_fast(N):
    s, _ = swelling(2, N, "diagonal")
    Pss = submatrix(s.P, s.is_s, s.is_s)
    amg = SAAMG(Pss, 2, rigid_body_modes(s.coords_s, 2))
    d = Pss.diagonal()
    off = np.asarray(abs(Pss).sum(1)).ravel() - abs(d)
    b = np.random.default_rng(0).standard_normal(Pss.shape[0])
    b[off <= 1e-14 * abs(d)] = 0.0                           # residuals on the path vanish on Dirichlet rows
    r = cg(lambda v: Pss @ v, b, amg, rtol=1e-8, max_it=100, norm_type="unpreconditioned")
    assert r.reason == 2 and r.its <= 25
    assert amg.complexity() < 2.0
