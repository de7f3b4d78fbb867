import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from helpers import AMG_OPTIONS, gpu_solve, rel
from oracle.problems import swelling
from oracle.blockpc import *
from poro_b200.lib.backend import DeviceVector, get_context
N = int(sys.argv[1])
s, par = swelling(3, N, "diagonal")
par = dict(par); par.update({"solver rtol": 1e-8, "solver atol": 0.0})
g = gpu_solve(s, par, AMG_OPTIONS, return_objects=True)
pcx = g["pc"].pc.getPythonContext()
ctx = get_context(0)
fpo = SchurLower(submatrix(s.P, s.is_fp, s.is_fp), s.nf, s.np_, lambda M: None, lambda M: None, "f")
Sg = pcx.block("schur")
print("S diff", abs(Sg - fpo.S).max() / abs(fpo.S).max(), Sg.nnz, fpo.S.nnz)
print("ss diff", abs(pcx.block("ss") - submatrix(s.P, s.is_s, s.is_s)).max())
n = s.np_
I = np.eye(n)
inv = np.zeros((n, n))
for j in range(n):
    dr, dz = DeviceVector(I[:, j], ctx=ctx), DeviceVector(n=n, ctx=ctx)
    pcx.inner_solve("fp1", dr, dz); ctx.sync()
    inv[:, j] = dz.numpy()
Sd = Sg.toarray()
R = inv @ Sd - np.eye(n)
print("||inv S - I|| max", abs(R).max(), "rows bad", np.flatnonzero(abs(R).max(1) > 1e-8)[:20], "cols bad", np.flatnonzero(abs(R).max(0) > 1e-8)[:20])
