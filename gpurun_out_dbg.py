import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from helpers import AMG_OPTIONS, gpu_solve, rel
from oracle.problems import swelling
N = int(sys.argv[1])
s, par = swelling(3, N, "diagonal")
par = dict(par); par.update({"solver rtol": 1e-8, "solver atol": 0.0})
g = gpu_solve(s, par, AMG_OPTIONS + "\n-poro_verbose\n")
print("N", N, "its", g["its"])
