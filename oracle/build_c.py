"""Builds the oracle's C part (CPU baseline only): the C + OpenMP solve loop (oracle/_cpu_solver.so)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
TARGETS = [("cpu_solver.c", "_cpu_solver.so")]
OUT = os.path.join(HERE, TARGETS[0][1])


def build(force=False):
    outs = []
    for src, out in TARGETS:
        src, out = os.path.join(HERE, "csrc", src), os.path.join(HERE, out)
        if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", src, "-o", out, "-lm"],
                           check=True)
        outs.append(out)
    return outs[0]


if __name__ == "__main__":
    build(force=True)
    print([t[1] for t in TARGETS])
