"""Host-side problem generator: meshes, the P2^d x P2^d x P1 assembler and the driver configurations.

This package stands in for the part of the reference that the north star LEAVES ON THE HOST and
outside the solve path: FEniCS assembly (lib/Assembler.py, lib/MeshCreation.py, lib/Poromechanics.py:76-83)
and the problem definitions of the drivers (swelling.py, swelling-3d.py).  dolfin is absent from this
image, so the matrices the solve phase consumes are generated here.  It is input-data generation for
bench.py, the tests and the oracle -- it is NOT part of the CUDA product path and NOT part of the CPU
oracle of the solve phase (oracle/), which only checks and times the solver.
"""
