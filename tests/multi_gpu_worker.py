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
# iteration counts of the CPU twin of the row-partitioned preconditioner (oracle/ddamg.py, tests/test_oracle_ddamg.py)
TWIN_ITS = {(8, 2): 49, (8, 4): 70}
twin_ok = (N, world) not in TWIN_ITS or abs(ksp.its - TWIN_ITS[(N, world)]) <= max(2, TWIN_ITS[(N, world)] // 10)
ok = err_spmv < 1e-13 and ksp.reason == 2 and err < 1e-8 and twin_ok
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0 and int(flag) == 1:
    print("MULTI-GPU OK", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
