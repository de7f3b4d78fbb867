"""solid.py of the reference (solid.py:111-180) on the GPU: the solid block alone, KSP prefix `s_`.
    python examples/solid.py [-N n] [--petsc-options FILE]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _single_block import run  # noqa: E402

if __name__ == "__main__":
    run("s", 10)
