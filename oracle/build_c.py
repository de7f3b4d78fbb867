"""Builds the oracle's optional OpenMP kernels (CPU baseline only) into oracle/_omp_kernels.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "omp_kernels.c")
OUT = os.path.join(HERE, "_omp_kernels.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", SRC, "-o", OUT], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
