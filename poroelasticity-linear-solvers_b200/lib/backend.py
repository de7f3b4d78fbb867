"""Device-side stand-ins for the petsc4py / dolfin objects the reference's Solver and
Preconditioner receive: `PETScMatrix.mat()`, `PETScVector.vec()` (lib/Solver.py:67,95;
lib/Preconditioner.py:284-289).  PyTorch owns the device buffers; all arithmetic happens in
libporo.so through raw pointers."""
from __future__ import annotations

import os

import numpy as np

from .. import _capi

_CTX = None


def get_context(device: int | None = None) -> _capi.Context:
    """Process-wide context (the role PETSc's global state plays in the reference)."""
    global _CTX
    if _CTX is None:
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        _CTX = _capi.Context(device)
    return _CTX


def reset_context():
    global _CTX
    if _CTX is not None:
        _CTX.close()
    _CTX = None


class DeviceMatrix:
    """CSR matrix resident in HBM; `.mat()` returns itself like dolfin's PETScMatrix."""

    def __init__(self, A, ctx: _capi.Context | None = None):
        self.ctx = ctx or get_context()
        A = A.tocsr()
        self.shape = A.shape
        self.nnz = A.nnz
        self._m = _capi.Mat.from_scipy(self.ctx, A)

    def mat(self):
        return self

    @property
    def handle(self):
        return self._m.h

    def mult(self, x, y):
        self._m.mult(_tensor(x), _tensor(y))


class DeviceVector:
    """fp64 vector on the GPU (torch tensor); `.vec()` returns itself like PETScVector."""

    def __init__(self, data=None, n: int | None = None, ctx: _capi.Context | None = None):
        import torch
        self.ctx = ctx or get_context()
        dev = torch.device("cuda", self.ctx.device)
        if data is None:
            self.t = torch.zeros(n, dtype=torch.float64, device=dev)
        elif isinstance(data, torch.Tensor):
            self.t = data.to(device=dev, dtype=torch.float64).contiguous()
        else:
            self.t = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float64)).to(dev)

    def vec(self):
        return self

    def copy(self):
        return DeviceVector(self.t.clone(), ctx=self.ctx)

    def norm(self):
        return float(self.t.norm())

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __len__(self):
        return self.t.numel()


def _tensor(v):
    return v.t if isinstance(v, DeviceVector) else v
