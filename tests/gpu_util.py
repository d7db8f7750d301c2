"""Thin helpers for the -m gpu parity tests: call libnaviflow_b200's C-ABI on NumPy inputs."""
import ctypes as C

import numpy as np

from naviflow_b200 import _lib
from naviflow_b200._lib import NfGrid, NfLinks
from naviflow_b200.device import bc_program_struct, get_context, pad_ld, ptr


class Dev:
    def __init__(self, n, rho=1.0, nx=None, ny=None):
        self.ctx = get_context()
        self.lib = self.ctx.lib
        self.nx = nx or n
        self.ny = ny or n
        self.dx, self.dy = 1.0 / (self.nx - 1), 1.0 / (self.ny - 1)
        self.g = self.ctx.grid(self.nx, self.ny, self.dx, self.dy, rho)

    def up(self, a):
        return self.ctx.upload(a, self.nx, self.ny)

    def zeros(self):
        return self.ctx.empty(self.nx, self.ny)

    def down(self, t, rows=None, cols=None):
        return self.ctx.download(t, rows or self.nx, cols or self.ny)

    def call(self, name, *args):
        self.ctx.check(getattr(self.lib, name)(self.ctx.handle, *args), name)

    def gref(self):
        return C.byref(self.g)


def grid_for(ctx, n):
    return ctx.grid(n, n, 1.0 / (n - 1), 1.0 / (n - 1), 1.0)


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def same_nan(a, b):
    """bit-exact including NaN positions"""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(
        np.nan_to_num(a, nan=0.0), np.nan_to_num(b, nan=0.0))
