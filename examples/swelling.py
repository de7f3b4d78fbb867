"""swelling.py of the reference, on the GPU solve phase.  Same flags as the reference's Parser:
    python examples/swelling.py [-N n] [--solver-type gmres|fgmres|cg|aar] [--pc-type "diagonal 3-way"] [--petsc-options options/petsc-options-exact] [--monitor]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _driver import run  # noqa: E402

if __name__ == "__main__":
    run("swelling", 10)
