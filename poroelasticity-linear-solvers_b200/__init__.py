"""B200-native solve phase for three-field poromechanics (drop-in for the PETSc Krylov/PC path of
nabw/poroelasticity-linear-solvers).  Import as `poro_b200`.

    poro_b200._capi      ctypes binding of libporo.so (include/poro.h)
    poro_b200.lib.*      Solver / Preconditioner / AAR / AndersonAcceleration / IndexSet / Parser
                         with the reference's class, method and parameter-key surface
"""
__version__ = "0.1.0"
