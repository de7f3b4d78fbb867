"""Re-export of the host-side assembler (moved to hostfem/fem.py: it is input generation, not the oracle)."""
from hostfem.fem import *          # noqa: F401,F403
from hostfem.fem import PoroAssembler, PoroSystem, simplex_quadrature, unit_cube_mesh, unit_square_mesh  # noqa: F401
