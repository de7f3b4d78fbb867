"""Re-export of the driver configurations (moved to hostfem/problems.py)."""
from hostfem.problems import *     # noqa: F401,F403
from hostfem.problems import _traction, footing, footing_params, swelling, swelling_assembler, swelling_params  # noqa: F401
