"""CPU: the twin of the row-partitioned preconditioner (oracle/ddamg.py) -- rank-local hierarchies, global
level-0 smoothing, Schur complement from owned parts.  Its iteration counts are what the multi-GPU runs produced on
2 and 4 B200 (tests/multi_gpu_worker.py 8: 49 and 70 iterations), which pins the distributed path to an oracle."""
import numpy as np
import pytest

from oracle.amg import rigid_body_modes
from oracle.blockpc import BlockPC, submatrix
from oracle.ddamg import DDAmg, LocalSchurAmg
from oracle.krylov import gmres
from oracle.problems import swelling

# observed on the GPUs with tests/multi_gpu_worker.py (N = 8, rtol 1e-10, AMG_OPTIONS)
GPU_COUNTS = {2: 49, 4: 70}


def twin_solve(N, R, rtol=1e-10):
    from poro_b200.partition import slab_ranges
    s, _ = swelling(3, N, "diagonal")
    h2 = 1e-2 / (2 * N)
    plane_s = np.rint(s.coords_s[:, 2] / h2).astype(int)
    plane_p = np.rint(s.coords_p[:, 2] / h2).astype(int)
    ranges = slab_ranges(2 * N + 1, R)
    ps = [np.flatnonzero((plane_s >= a) & (plane_s < b)) for a, b in ranges]
    pp = [np.flatnonzero((plane_p >= a) & (plane_p < b)) for a, b in ranges]
    Bl = np.zeros((s.ns, 6))
    for p in ps:                                   # each rank builds its rigid-body modes from its own coordinates
        Bl[p] = rigid_body_modes(s.coords_s[p], 3)
    f, p_ = np.arange(s.nf), s.nf + np.arange(s.np_)

    class Schur:
        def __init__(self, M):
            self.A00, self.A01, self.A10, self.A11 = M[f][:, f], M[f][:, p_], M[p_][:, f], M[p_][:, p_]
            self.k0 = DDAmg(self.A00, 3, Bl, ps)
            self.k1 = LocalSchurAmg(self.A00, self.A01, self.A10, self.A11, ps, pp)

        def __call__(self, x):
            y0 = self.k0(x[: s.nf])
            return np.concatenate([y0, self.k1(x[s.nf:] - self.A10 @ y0)])

    pc = BlockPC(s, {"s": lambda M: DDAmg(M, 3, Bl, ps), "fp": lambda M: Schur(M)})
    r = gmres(lambda v: s.A @ v, s.b, pc, rtol=rtol, atol=0.0, dtol=1e20, max_it=300, restart=300, pc_side="right")
    return s, r


@pytest.mark.parametrize("R", [2, 4])
def test_twin_reproduces_multi_gpu_iteration_counts(R):
    s, r = twin_solve(8, R)
    assert r.reason == 2
    assert abs(r.its - GPU_COUNTS[R]) <= 2
    assert np.linalg.norm(s.b - s.A @ r.x) <= 2e-10 * np.linalg.norm(s.b)


def test_one_rank_twin_is_the_plain_amg():
    """With a single part the twin must coincide with the undistributed preconditioner."""
    from oracle.amg import SAAMG
    s, _ = swelling(2, 10, "diagonal")
    Pss = submatrix(s.P, s.is_s, s.is_s)
    B = rigid_body_modes(s.coords_s, 2)
    x = np.random.default_rng(0).standard_normal(s.ns)
    a = SAAMG(Pss, 2, B)(x)
    b = DDAmg(Pss, 2, B, [np.arange(s.ns)])(x)
    assert np.linalg.norm(a - b) <= 1e-12 * np.linalg.norm(a)
