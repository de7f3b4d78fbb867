"""CPU oracle for the solve phase of nabw/poroelasticity-linear-solvers.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker or the
timed CPU baseline.

PARITY UNPINNED: the reference (pure Python over petsc4py / dolfin / hypre /
MUMPS) ships no tests, golden vectors or recorded iteration counts, and none
of its dependencies exist in this image, so the restatement below cannot be
checked against outputs of the reference itself.  What pins it instead:
manufactured-solution checks of the assembler, direct-solve cross-checks of
the Krylov restatements, and the per-function file:line citations.

Modules
-------
fem, problems   re-exports of hostfem/ (the host-side assembler and driver configs: input generation that
                stands in for FEniCS, shared by bench.py, the tests and this oracle; not part of the solve path)
krylov     PETSc-semantics GMRES / CG / preonly restatements (lib/Solver.py:92-102)
blockpc    PreconditionerCC.apply 2-way / 3-way (lib/Preconditioner.py:141-250)
aar        AAR.solve incl. its quirks (lib/AAR.py:46-137)
anderson   AndersonAcceleration.get_next_vector (lib/AndersonAcceleration.py:19-78)
amg        CPU restatement of OUR smoothed-aggregation AMG (not hypre)
ddamg      CPU twin of the row-partitioned (multi-GPU) preconditioners: rank-local hierarchies, global smoothing
distamg    rank-level restatement of the DISTRIBUTED hierarchy (uncoupled aggregation, distributed Galerkin), all ranks
           emulated in one process; distamg_rank: the same as one rank over torch.distributed (gloo)
cport      ctypes glue of csrc/cpu_solver.c: the timed solve loop of the CPU baseline in C + OpenMP
"""
