"""fluid-pressure.py of the reference (fluid-pressure.py:85-136) on the GPU: the fluid-pressure block alone, KSP prefix `fp_`
with a (f, p) fieldsplit.
    python examples/fluid-pressure.py [-N n] [--petsc-options FILE]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _single_block import run  # noqa: E402

if __name__ == "__main__":
    run("fp", 10)
