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


class GpuMatrixFreeMomentumSolver:
    """Twin of ``MatrixFreeMomentumSolver`` (solver/momentum_solver/matrix_free_momentum.py:11-544) with
    ``solver_type='bicgstab'``: power-law links, a_P clamped to 1e-12 and divided by the relaxation factor, source relaxed
    with the relaxed a_P, scipy's BiCGSTAB recurrence on the relaxed system (interior rows 5-point, boundary rows
    identity) started from the current velocity and stopped at ``||r|| < max(tolerance, 1e-5 ||b||)``, BCs re-applied to
    the solution (with the reference's ``nx+1`` call), ``d = dy/a_P`` (0 where a_P vanishes) and ``rel_norm`` = the absolute
    norm of the *unrelaxed* residual over the interior.

    The reference preconditions the iteration with an ILU (``ilu_drop_tol``, ``ilu_fill_factor``) of a matrix that, as
    coded, lacks the north/south links; the device iteration is unpreconditioned, so the two agree to the stopping
    tolerance, not bit for bit (this solver cannot be pinned below ~1e-6 against itself either).  The ILU
    arguments are accepted and ignored; ``gmres`` / ``idrs`` are not implemented."""

    def __init__(self, discretization_scheme="power_law", tolerance=1e-8, max_iterations=200, solver_type="bicgstab",
                 ilu_drop_tol=1e-3, ilu_fill_factor=15, idrs_s=4, device=None):
        solver_type = str(solver_type).lower()
        if solver_type not in {"gmres", "bicgstab", "idrs"}:  # matrix_free_momentum.py:31-33
            raise ValueError("solver_type must be 'gmres', 'bicgstab', or 'idrs'")
        if solver_type != "bicgstab":
            raise NotImplementedError("the device momentum solver implements solver_type='bicgstab'")
        if discretization_scheme != "power_law":
            if discretization_scheme in ("quick", "second_order_upwind"):
                raise NotImplementedError("only the power-law scheme is implemented on the device")
            raise ValueError(f"Unsupported discretization scheme: {discretization_scheme}")
        self.tol = float(tolerance)
        self.maxiter = int(max_iterations)
        self.solver_type = solver_type
        self.ilu_drop_tol, self.ilu_fill_factor, self.idrs_s = float(ilu_drop_tol), int(ilu_fill_factor), int(idrs_s)
        self._device = device
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context(self._device)
        return self._ctx

    def _solve(self, is_u, mesh, fluid, u, v, p, alpha, bc):
        from ._lib import NfKrylovInfo
        ctx = self.ctx
        nx, ny = mesh.get_dimensions()
        dx, dy = mesh.get_cell_sizes()
        g = ctx.grid(nx, ny, dx, dy, fluid.get_density())
        ud, vd, pd = ctx.upload(u, nx, ny), ctx.upload(v, nx, ny), ctx.upload(p, nx, ny)
        ubc, vbc = ud.clone(), vd.clone()
        prog = bc_program_struct(bc, nx, ny, nx + 1)   # the reference passes nx+1 (matrix_free_momentum.py:419, :491)
        apply_bc = lambda a, b: ctx.check(ctx.lib.nf_apply_velocity_bc(ctx.handle, C.byref(g), C.byref(prog), ptr(a), ptr(b)),
                                          "nf_apply_velocity_bc")
        apply_bc(ubc, vbc)
        arrays = [ctx.empty(nx, ny) for _ in range(6)]
        links = NfLinks(*[a.data_ptr() for a in arrays])
        d, ap_un, src_un = ctx.empty(nx, ny), ctx.empty(nx, ny), ctx.empty(nx, ny)
        ctx.check(ctx.lib.nf_momentum_links_mf(ctx.handle, C.byref(g), int(is_u), ptr(ubc), ptr(vbc), ptr(pd),
                                               float(fluid.get_viscosity()), float(alpha), practice_b_sides(bc), links,
                                               ptr(d), ptr(ap_un), ptr(src_un)), "nf_momentum_links_mf")
        x = ud if is_u else vd   # x0 = the current velocity (:436, :508)
        work = ctx.torch.zeros((5 * (nx + 1), x.shape[1]), dtype=ctx.torch.float64, device=x.device)
        info = NfKrylovInfo()
        ctx.check(ctx.lib.nf_momentum_bicgstab(ctx.handle, C.byref(g), int(is_u), links, ptr(x), self.tol, 1e-5,
                                               self.maxiter, 10, ptr(work), C.byref(info)), "nf_momentum_bicgstab")
        if is_u:
            apply_bc(x, vbc)
        else:
            apply_bc(ubc, x)
        unrelaxed = NfLinks(arrays[0].data_ptr(), arrays[1].data_ptr(), arrays[2].data_ptr(), arrays[3].data_ptr(),
                            ap_un.data_ptr(), src_un.data_ptr())
        field = ctx.empty(nx, ny)
        norm = C.c_double()
        ctx.check(ctx.lib.nf_momentum_residual_unrelaxed(ctx.handle, C.byref(g), int(is_u), unrelaxed, ptr(x), ptr(field),
                                                         C.byref(norm)), "nf_momentum_residual_unrelaxed")
        rows, cols = (nx + 1, ny) if is_u else (nx, ny + 1)
        self.last_info = info
        res = {"rel_norm": norm.value, "field": ctx.download(field, rows, cols), "iterations": int(info.iterations),
               "solver_type": self.solver_type}
        return ctx.download(x, rows, cols), ctx.download(d, rows, cols), res

    def solve_u_momentum(self, mesh, fluid, u, v, p, relaxation_factor=0.7, boundary_conditions=None,
                         return_dict=True):
        us, du, res = self._solve(True, mesh, fluid, u, v, p, relaxation_factor, boundary_conditions)
        return (us, du, res) if return_dict else (us, du, res["rel_norm"])

    def solve_v_momentum(self, mesh, fluid, u, v, p, relaxation_factor=0.7, boundary_conditions=None,
                         return_dict=True):
        vs, dv, res = self._solve(False, mesh, fluid, u, v, p, relaxation_factor, boundary_conditions)
        return (vs, dv, res) if return_dict else (vs, dv, res["rel_norm"])


class _GpuExtendedStencilDiscretization:
    """Twin of the reference's higher-order discretization objects: ``calculate_u_coefficients`` /
    ``calculate_v_coefficients`` return the same dict of ten arrays (a_e, a_w, a_n, a_s, a_ee, a_ww, a_nn, a_ss, a_p,
    source), evaluated by ``nf_momentum_links_ext`` (csrc/nf_links_ext.cu).  As in the reference the velocities are used as
    passed (the caller applies the boundary conditions first) and nothing is relaxed."""
    _scheme = 0
    _KEYS = ("a_e", "a_w", "a_n", "a_s", "a_ee", "a_ww", "a_nn", "a_ss", "a_p", "source")

    def __init__(self, device=None):
        self._device = device
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context(self._device)
        return self._ctx

    def _coefficients(self, is_u, mesh, fluid, u, v, p, bc):
        from ._lib import NfLinksExt
        ctx = self.ctx
        nx, ny = mesh.get_dimensions()
        dx, dy = mesh.get_cell_sizes()
        g = ctx.grid(nx, ny, dx, dy, fluid.get_density())
        ud, vd, pd = ctx.upload(u, nx, ny), ctx.upload(v, nx, ny), ctx.upload(p, nx, ny)
        arrays = [ctx.empty(nx, ny) for _ in range(10)]
        out = NfLinksExt(*[a.data_ptr() for a in arrays])
        sides = practice_b_sides(bc) if bc is not None else 0
        ctx.check(ctx.lib.nf_momentum_links_ext(ctx.handle, C.byref(g), int(is_u), int(self._scheme), ptr(ud), ptr(vd), ptr(pd),
                                                float(fluid.get_viscosity()), int(sides), out), "nf_momentum_links_ext")
        rows, cols = (nx + 1, ny) if is_u else (nx, ny + 1)
        return {k: ctx.download(a, rows, cols) for k, a in zip(self._KEYS, arrays)}


class GpuQUICKDiscretization(_GpuExtendedStencilDiscretization):
    """``QUICKDiscretization`` (discretization/quick.py:27-219)."""
    _scheme = 1

    def calculate_u_coefficients(self, mesh, fluid, u, v, p, bc=None):
        return self._coefficients(1, mesh, fluid, u, v, p, bc)

    def calculate_v_coefficients(self, mesh, fluid, u, v, p, bc=None):
        return self._coefficients(0, mesh, fluid, u, v, p, bc)


class GpuSecondOrderUpwindDiscretization(_GpuExtendedStencilDiscretization):
    """``SecondOrderUpwindDiscretization`` (discretization/second_order_upwind.py:26-325)."""
    _scheme = 2

    def calculate_u_coefficients(self, mesh, fluid, u, v, p, bc_manager=None):
        return self._coefficients(1, mesh, fluid, u, v, p, bc_manager)

    def calculate_v_coefficients(self, mesh, fluid, u, v, p, bc_manager=None):
        return self._coefficients(0, mesh, fluid, u, v, p, bc_manager)
