"""GPU velocity correction with the reference's ``VelocityUpdater`` interface
(solver/velocity_solver/standard.py:10-69): u = u* + d_u (p'_W - p'_P) on the interior faces, then the
velocity boundary conditions."""
from __future__ import annotations

import ctypes as C

from .device import bc_program_struct, get_context, ptr


class GpuVelocityUpdater:
    def __init__(self, device=None):
        self._device = device
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context(self._device)
        return self._ctx

    def update_velocity(self, mesh, u_star, v_star, p_prime, d_u, d_v, boundary_conditions):
        ctx = self.ctx
        nx, ny = mesh.get_dimensions()
        dx, dy = mesh.get_cell_sizes()
        g = ctx.grid(nx, ny, dx, dy, 1.0)
        prog = bc_program_struct(boundary_conditions, nx, ny)
        us, vs, pp = ctx.upload(u_star, nx, ny), ctx.upload(v_star, nx, ny), ctx.upload(p_prime, nx, ny)
        du, dv = ctx.upload(d_u, nx, ny), ctx.upload(d_v, nx, ny)
        u, v = ctx.empty(nx, ny), ctx.empty(nx, ny)
        ctx.check(ctx.lib.nf_correct_velocity(ctx.handle, C.byref(g), C.byref(prog), ptr(us), ptr(vs), ptr(pp),
                                              ptr(du), ptr(dv), ptr(u), ptr(v)), "nf_correct_velocity")
        return ctx.download(u, nx + 1, ny), ctx.download(v, nx, ny + 1)
