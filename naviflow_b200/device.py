"""Device plumbing: the library context, pitched device fields, grid descriptors.

PyTorch is used for exactly two things: owning device memory (``torch.empty(..., device='cuda')``)
and giving us the CUDA stream the library launches on.  No torch op touches the hot path.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import NfBcProgram, NfGrid, check
from .host import boundary_program

_contexts = {}


def pad_ld(ny):
    """Common row pitch (doubles) of all fields of an (nx, ny) grid: >= ny+1, multiple of 16 (128 B)."""
    return ((ny + 1 + 15) // 16) * 16


class Context:
    """One nf_ctx per (process, device), bound to torch's current stream on that device."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("naviflow_b200 needs a CUDA device (there is no CPU fallback)")
        self.torch = torch
        self.device = int(device)
        torch.cuda.set_device(self.device)
        self.lib = _lib.lib()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        h = C.c_void_p()
        st = self.lib.nf_ctx_create(C.byref(h), self.device, C.c_void_p(stream))
        if st != 0:
            raise _lib.NfError(f"nf_ctx_create failed ({st}): {self.lib.nf_last_error(None).decode()}")
        self.handle = h
        self.stream = stream

    def check(self, status, what=""):
        check(self.handle, status, what)

    def sync(self):
        self.check(self.lib.nf_sync(self.handle), "nf_sync")

    def launches(self):
        return int(self.lib.nf_launch_count(self.handle))

    # ---- memory -------------------------------------------------------------------------------
    def empty(self, nx, ny):
        """Zeroed device field able to hold u (nx+1, ny), v (nx, ny+1) or p (nx, ny)."""
        return self.torch.zeros((nx + 1, pad_ld(ny)), dtype=self.torch.float64, device=f"cuda:{self.device}")

    def upload(self, arr, nx, ny, out=None):
        """Host ndarray (rows<=nx+1, cols<=ny+1) -> pitched device field."""
        a = np.ascontiguousarray(arr, dtype=np.float64)
        t = self.empty(nx, ny) if out is None else out
        t[: a.shape[0], : a.shape[1]].copy_(self.torch.from_numpy(a))
        return t

    def download(self, t, rows, cols):
        return t[:rows, :cols].cpu().numpy().copy()

    def grid(self, nx, ny, dx, dy, rho=1.0):
        return NfGrid(nx, ny, pad_ld(ny), 0, 0, nx, 0, 0, dx, dy, rho)

    def __del__(self):
        try:
            self.lib.nf_ctx_destroy(self.handle)
        except Exception:
            pass


def get_context(device=None):
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


def bc_program_struct(bc, nx, ny, nx_arg=None):
    d = boundary_program(bc, nx, ny, nx_arg)
    s = NfBcProgram()
    for k in ("u_edge", "u_corner", "v_edge", "v_corner"):
        for i in range(4):
            getattr(s, k)[i] = d[k][i]
    s.v_right_row = d["v_right_row"]
    return s


def mesh_scalars(mesh):
    nx, ny = mesh.get_dimensions()
    dx, dy = mesh.get_cell_sizes()
    length = getattr(mesh, "length", dx * (nx - 1))
    height = getattr(mesh, "height", dy * (ny - 1))
    return nx, ny, dx, dy, length, height
