"""AndersonAcceleration (reference lib/AndersonAcceleration.py:7-78).

On the GPU path the acceleration of the preconditioner output runs inside libporo.so
(csrc/solver.cu: Anderson::get_next_vector, Gram-matrix least squares); this class carries
the `order` exactly like the reference object does and is consulted by PreconditionerCC.
"""


class AndersonAcceleration:
    def __init__(self, order):
        self.order = order
        self.k = 0
