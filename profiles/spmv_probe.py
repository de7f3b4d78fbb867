"""SpMV probe for ncu: assembles swelling-3d N (default 34), uploads a matrix and times / profiles
y = M x through poro_mat_mult (the raw-matrix entry point; same kernels as the solver).

    python profiles/spmv_probe.py [N] [reps] [which]     which in {A, ss, ss_bsr, ff_bsr}
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from hostfem.problems import swelling
from poro_b200.lib.backend import DeviceMatrix, DeviceVector, get_context

N = int(sys.argv[1]) if len(sys.argv) > 1 else 34
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
which = sys.argv[3].split(",") if len(sys.argv) > 3 else ["A"]
s, _ = swelling(3, N, "diagonal")
ctx = get_context(0)
for w in which:
    ctx.clear_options()
    if w == "A":
        M = s.A
    elif w == "rem":
        import scipy.sparse as sp
        Z = sp.csr_matrix((s.np_, s.np_))
        M = (s.A - sp.block_diag([s.A[s.is_s][:, s.is_s], s.A[s.is_f][:, s.is_f], Z])).tocsr()
        M.eliminate_zeros()
    else:
        idx = s.is_s if w.startswith("ss") else s.is_f
        M = s.A[idx][:, idx].tocsr()
    bsr = w.endswith("_bsr")
    if bsr:
        ctx.set_option("-poro_mat_block_hint", 3)
    dM = DeviceMatrix(M, ctx)
    n = M.shape[0]
    x = DeviceVector(np.random.default_rng(0).standard_normal(n), ctx=ctx)
    y = DeviceVector(n=n, ctx=ctx)
    for _ in range(3):
        dM.mult(x, y)
    ctx.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()          # the library's stream is a blocking stream: default-stream events order with it
    for _ in range(reps):
        dM.mult(x, y)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    if bsr:
        import scipy.sparse as sp
        nnzb = sp.bsr_matrix(M, blocksize=(3, 3)).indices.size
        nbytes = 76 * nnzb + 4 * (n // 3 + 1) + 8 * n + 8 * n
    else:
        nbytes = 12 * M.nnz + 4 * (n + 1) + 8 * n + 8 * n
    ref = M @ x.numpy()
    err = np.linalg.norm(y.numpy() - ref) / np.linalg.norm(ref)
    print("%-7s N=%d n=%d nnz=%d  %.4f ms/SpMV  %.1f GB/s algorithmic (%.1f%% of 6559.4)  csr-equivalent %.1f GB/s  relerr %.1e" % (
        w, N, n, M.nnz, ms, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / 6559.4, (12 * M.nnz + 20 * n) / ms / 1e6, err), flush=True)
