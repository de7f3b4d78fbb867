"""Per-mode timing of the BSR SpMV kernels on the solid block of swelling-3d (B200).

    python profiles/bsr_modes_probe.py [N] [reps]

For every kernel variant (plain ld.global.cs kernel of bsr.cu; TMA pipeline shapes 0/1/2 of bsr_tma.cu) and every epilogue
mode the solver uses, prints the device time per launch and the algorithmic GB/s (76 B per 3x3 block + vectors)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scipy.sparse as sp
from hostfem.problems import swelling
from poro_b200 import _capi
from poro_b200.lib.backend import get_context

N = int(sys.argv[1]) if len(sys.argv) > 1 else 34
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
s, _ = swelling(3, N, "diagonal")
M = s.P[s.is_s][:, s.is_s].tocsr()
n = M.shape[0]
nnzb = sp.bsr_matrix(M, blocksize=(3, 3)).indices.size
ctx = get_context(0)
peak = 6559.4
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
MODES = {0: ("y = A x", 16), 1: ("y = z - A x", 24), 2: ("y = z + A x", 24), 3: ("Chebyshev step", 64), 4: ("w = A p, p.w", 24)}
out = {"N": N, "n": n, "nnzb": int(nnzb), "peak": peak, "rows": []}
VARIANTS = [("ld.global.cs (bsr.cu)", {"-poro_bsr_tma": 0}), ("TMA 512x2, 2 CTA/SM", {"-poro_bsr_tma": 1, "-poro_bsr_tma_cfg": 0}),
            ("TMA 256x2, 4 CTA/SM", {"-poro_bsr_tma": 1, "-poro_bsr_tma_cfg": 1}), ("TMA 256x3, 3 CTA/SM", {"-poro_bsr_tma": 1, "-poro_bsr_tma_cfg": 2}),
            ("ld.global.cs, no operand prefetch", {"-poro_bsr_tma": 0, "-poro_bsr_prefetch": 0})]
VARIANTS += [("ld.global.cs + L2 prefetch %d chunks ahead" % d, {"-poro_bsr_tma": 0, "-poro_bsr_l2_prefetch_chunks": d})
             for d in (740, 1184, 1776, 2368, 4736, 640, 888)]
VARIANTS += [("ld.global.cs + cooperative gathers", {"-poro_bsr_tma": 0, "-poro_bsr_coop_gather": 1})]      # index 12
if os.environ.get("PROBE_VARIANTS"):
    VARIANTS = [VARIANTS[int(i)] for i in os.environ["PROBE_VARIANTS"].split(",")]
if os.environ.get("PROBE_MODES"):
    MODES = {int(m): MODES[int(m)] for m in os.environ["PROBE_MODES"].split(",")}
for name, opts in VARIANTS:
    ctx.clear_options()
    ctx.set_option("-poro_mat_block_hint", 3)
    for k, v in opts.items():
        ctx.set_option(k, v)
    dM = _capi.Mat.from_scipy(ctx, M)
    for mode, (label, vec_bytes) in MODES.items():
        ms = dM.bench(mode, reps)
        nbytes = 76 * nnzb + 4 * (n // 3 + 1) + vec_bytes * n
        gbs = nbytes / ms / 1e6
        out["rows"].append({"kernel": name, "mode": label, "ms": ms, "GBs": gbs, "frac": gbs / peak})
        print("%-24s %-16s %.4f ms  %7.1f GB/s  %.3f of measured peak" % (name, label, ms, gbs, gbs / peak), flush=True)
    dM.destroy()
print(json.dumps(out))
