"""Worker for tests/test_gpu_multi.py and for manual runs:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_worker.py [N]
Solves swelling-3d (default N=8) row-partitioned over the ranks and compares with a direct solve."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from helpers import AMG_OPTIONS

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8

from poro_b200.lib.backend import DeviceMatrix, DeviceVector, get_context
from poro_b200.lib.Parser import load_petsc_options
from poro_b200.lib.Preconditioner import Preconditioner
from poro_b200.lib.Solver import Solver
from poro_b200.partition import distributed_problem

ctx = get_context(local)
load_petsc_options(ctx, AMG_OPTIONS, is_text=True)
for opt in os.environ.get("PORO_EXTRA_OPTIONS", "").split(";"):      # e.g. "-poro_p2p 0;-poro_verbose"
    if opt.strip():
        kv = opt.split()
        ctx.set_option(kv[0], kv[1] if len(kv) > 1 else None)
CC = os.environ.get("PORO_WORKER_CC", "0") == "1"      # additive Cahouet-Chabard Schur preconditioner instead of selfp
if CC:
    ctx.set_option("-fp_pc_fieldsplit_schur_precondition", "cc")
prob = distributed_problem(3, N, "diagonal", rank, world, ctx)
s, par = prob.sys, dict(prob.par)
par.update({"solver rtol": 1e-10, "solver atol": 0.0, "solver maxiter": 100})
imap = prob.index_set()
dA, dP = DeviceMatrix(s.A, ctx), DeviceMatrix(s.P, ctx)
db, dx = DeviceVector(s.b, ctx=ctx), DeviceVector(n=len(s.b), ctx=ctx)
pc = Preconditioner(imap, dA, dP, None, par, s.bcs_sub_pressure).get_pc()
solver = Solver(dA, db, pc, par, imap)
solver.create_solver(dA, db, pc)
# distributed SpMV against the global matrix
from oracle.problems import swelling
glob, _ = swelling(3, N, "diagonal")
xg = np.random.default_rng(3).standard_normal(glob.n)
dxin, dy = DeviceVector(xg[s.owned_global], ctx=ctx), DeviceVector(n=len(s.b), ctx=ctx)
dA.mult(dxin, dy)
ctx.sync()
ref = (glob.A @ xg)[s.owned_global]
err_spmv = np.linalg.norm(dy.numpy() - ref) / np.linalg.norm(ref)
solver.solve(db, dx)
ksp = solver.solver
import scipy.sparse.linalg as spla
xd = spla.spsolve(glob.A.tocsc(), glob.b)
xl = dx.numpy()
num = torch.tensor([np.sum((xl - xd[s.owned_global]) ** 2), np.sum(xd[s.owned_global] ** 2)], device="cuda")
dist.all_reduce(num)
err = float(torch.sqrt(num[0] / num[1]))
if rank == 0:
    print("ranks", world, "N", N, "its", ksp.its, "reason", ksp.reason, "spmv err %.2e" % err_spmv, "solution err %.2e" % err, flush=True)
# iteration count of the CPU twin of the DISTRIBUTED hierarchies (oracle/distamg.py: uncoupled aggregation, distributed
# prolongator smoothing + Galerkin product, halo-aware selfp Schur complement), and of the single-GPU algorithm
twin_its = single_its = None
if rank == 0:
    import scipy.sparse as sp
    from oracle.amg import SAAMG, rigid_body_modes
    from oracle.blockpc import BlockPC, SchurLower, SchurLowerCC, cc_from_matrices, krylov_solver
    from oracle.distamg import DistAmg
    from oracle.krylov import gmres
    from poro_b200.partition import slab_ranges

    def slab_perm(coords, bs):
        h2 = coords[:, 2].max() / (2 * N)
        plane = np.rint(coords[::bs, 2] / h2).astype(int)
        parts = [np.flatnonzero((plane >= a) & (plane < b)) for a, b in slab_ranges(2 * N + 1, world)]
        perm = np.concatenate([(p[:, None] * bs + np.arange(bs)).ravel() for p in parts])
        return perm, np.concatenate([[0], np.cumsum([len(p) * bs for p in parts])])

    class Permuted:
        def __init__(self, M, perm, off, bs, B, **kw):
            self.perm = perm
            self.h = DistAmg(sp.csr_matrix(M)[perm][:, perm].tocsr(), off, bs, None if B is None else B[perm], **kw)

        def __call__(self, b):
            y = np.empty_like(b)
            y[self.perm] = self.h(b[self.perm])
            return y

    Bg = rigid_body_modes(glob.coords_s, 3)
    perm_s, off_s = slab_perm(glob.coords_s, 3)
    perm_p, off_p = slab_perm(glob.coords_p, 1)

    def outer(mk_v, mk_p):
        mkfp = lambda M: SchurLower(M, glob.nf, glob.np_, krylov_solver("preonly", mk_v), krylov_solver("preonly", mk_p), "f")
        if CC:
            d_mass, S_visc = cc_from_matrices(glob, par)
            cheb_p = lambda M: SAAMG(M, 1, None, max_levels=1, cheby_degree=4, dense_limit=0)
            mkfp = lambda M: SchurLowerCC(M, glob.nf, glob.np_, krylov_solver("preonly", mk_v), krylov_solver("preonly", mk_p),
                                          krylov_solver("preonly", cheb_p), d_mass, S_visc)
        pcg = BlockPC(glob, {"s": krylov_solver("preonly", mk_v), "fp": mkfp})
        return gmres(lambda v: glob.A @ v, glob.b, pcg, rtol=1e-10, atol=0.0, dtol=1e20, max_it=100, restart=100, pc_side="right")

    twin_its = outer(lambda M: Permuted(M, perm_s, off_s, 3, Bg), lambda M: Permuted(M, perm_p, off_p, 1, None)).its
    single_its = outer(lambda M: SAAMG(M, 3, Bg), lambda M: SAAMG(M, 1, None)).its
    print("twin (distributed hierarchy on CPU) its", twin_its, " single-process hierarchy its", single_its, flush=True)
twin_ok = True
if rank == 0:
    # same algorithm on both sides: +-10 % (the GPU builds its rigid-body modes about each rank's own centre);
    # and the row partition must not cost more than 25 % over the single-GPU count
    twin_ok = abs(ksp.its - twin_its) <= max(2, twin_its // 10) and ksp.its <= max(single_its + 2, int(1.25 * single_its))
ok = err_spmv < 1e-13 and ksp.reason == 2 and err < 1e-8 and twin_ok
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0 and int(flag) == 1:
    print("MULTI-GPU OK", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
