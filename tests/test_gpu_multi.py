"""GPU, 2 ranks (run under torchrun on a 2-GPU box; skipped otherwise): the row-partitioned solve
reproduces the single-GPU solve -- same solution to 1e-8; iteration counts may differ slightly
because the AMG hierarchies are built per rank (block-Jacobi)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("cc", ["0", "1"])
def test_two_rank_solve_matches_single(cc):
    """cc = 1: the additive Cahouet-Chabard Schur preconditioner (lumped mass of the GHOST velocity dofs exchanged like
    diag(P_ff) for selfp; the Chebyshev sweep on P_pp uses the pressure halo plan)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, PORO_WORKER_CC=cc))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI-GPU OK" in r.stdout
