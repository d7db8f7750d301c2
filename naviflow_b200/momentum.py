"""GPU momentum predictor with the reference's ``MomentumSolver`` plugin interface.

Twin of ``JacobiMatrixMomentumSolver`` (solver/momentum_solver/jacobi_matrix_solver.py:11-375): power-law
link coefficients with under-relaxation and Practice-B boundary folding, ``n_jacobi_sweeps`` fixed Jacobi
sweeps started from the current velocity, ``d_u = dy/a_p`` (NaN where a_p = 0), and the relaxed residual
norm with boundaries masked.  ``return_dict=True`` (the default, as ``SimpleSolver`` calls it,
Algorithms/simple.py:121-133) returns ``(u_star, d_u, {'rel_norm', 'field'})``; ``return_dict=False`` returns
the reference class's 4-tuple.
"""
from __future__ import annotations

import ctypes as C

from ._lib import NfLinks
from .device import bc_program_struct, get_context, ptr
from .host import practice_b_sides


class GpuJacobiMomentumSolver:
    def __init__(self, discretization_scheme="power_law", n_jacobi_sweeps=1, device=None):
        if discretization_scheme != "power_law":  # jacobi_matrix_solver.py:21-24
            raise ValueError(f"Unsupported discretization scheme: {discretization_scheme}")
        self.n_jacobi_sweeps = int(n_jacobi_sweeps)
        self._device = device
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context(self._device)
        return self._ctx

    def _solve(self, is_u, mesh, fluid, u, v, p, alpha, bc):
        ctx = self.ctx
        nx, ny = mesh.get_dimensions()
        dx, dy = mesh.get_cell_sizes()
        g = ctx.grid(nx, ny, dx, dy, fluid.get_density())
        ud, vd, pd = ctx.upload(u, nx, ny), ctx.upload(v, nx, ny), ctx.upload(p, nx, ny)
        ubc, vbc = ud.clone(), vd.clone()
        prog = bc_program_struct(bc, nx, ny)
        ctx.check(ctx.lib.nf_apply_velocity_bc(ctx.handle, C.byref(g), C.byref(prog), ptr(ubc), ptr(vbc)),
                  "nf_apply_velocity_bc")
        arrays = [ctx.empty(nx, ny) for _ in range(6)]
        links = NfLinks(*[a.data_ptr() for a in arrays])
        d = ctx.empty(nx, ny)
        fn = ctx.lib.nf_momentum_links_u if is_u else ctx.lib.nf_momentum_links_v
        ctx.check(fn(ctx.handle, C.byref(g), ptr(ubc), ptr(vbc), ptr(pd), float(fluid.get_viscosity()), float(alpha),
                     practice_b_sides(bc), links, ptr(d)), "nf_momentum_links")
        x = ud if is_u else vd          # x0 = current velocity, not the BC'd copy (jacobi_matrix_solver.py:158, :196)
        tmp = ctx.empty(nx, ny)
        ctx.check(ctx.lib.nf_momentum_jacobi(ctx.handle, C.byref(g), int(is_u), links, ptr(x), ptr(tmp),
                                             self.n_jacobi_sweeps), "nf_momentum_jacobi")
        field = ctx.empty(nx, ny)
        norm = C.c_double()
        ctx.check(ctx.lib.nf_momentum_residual(ctx.handle, C.byref(g), int(is_u), links, ptr(x), ptr(field),
                                               C.byref(norm)), "nf_momentum_residual")
        rows, cols = (nx + 1, ny) if is_u else (nx, ny + 1)
        self._last_links = arrays
        return ctx.download(x, rows, cols), ctx.download(d, rows, cols), norm.value, ctx.download(field, rows, cols)

    def solve_u_momentum(self, mesh, fluid, u, v, p, relaxation_factor=0.7, boundary_conditions=None,
                         return_dict=True):
        us, du, norm, field = self._solve(True, mesh, fluid, u, v, p, relaxation_factor, boundary_conditions)
        if return_dict:
            return us, du, {"rel_norm": norm, "field": field}
        return us, du, norm, field

    def solve_v_momentum(self, mesh, fluid, u, v, p, relaxation_factor=0.7, boundary_conditions=None,
                         return_dict=True):
        vs, dv, norm, field = self._solve(False, mesh, fluid, u, v, p, relaxation_factor, boundary_conditions)
        if return_dict:
            return vs, dv, {"rel_norm": norm, "field": field}
        return vs, dv, norm, field
