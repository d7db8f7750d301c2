"""GPU pressure-correction solvers with the reference's ``PressureSolver`` plugin interface.

Each class mirrors its CPU twin's constructor arguments, ``solve`` signature, return values and
``info['rel_norm']`` convention, so it can be dropped into the reference's
``SimpleSolver`` (NumPy in / NumPy out, one H2D/D2H round trip per call) or used device-resident
inside :class:`naviflow_b200.simple.GpuSimpleSolver`.

Reference twins (paths relative to /root/reference/naviflow_oo/solver/pressure_solver):
  GpuJacobiSolver        jacobi.py:10-248
  GpuGaussSeidelSolver   gauss_seidel.py:10-393 (method_type='red_black' only)
  GpuMultiGridSolver     multigrid.py:21-750
  GpuBiCGSTABSolver      matrix_free_BiCGSTAB.py:15-343 (unpreconditioned)
  GpuCGSolver            new: scipy.sparse.linalg.cg on compute_Ap_product (the reference's CG classes
                         need pyamg)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NfKrylovInfo, NfMgConfig, NfMgInfo
from .device import get_context, mesh_scalars, pad_ld, ptr


class _GpuPressureBase:
    """Shared plumbing: H2D of (u*, v*, d_u, d_v), RHS on device, D2H of (p', residual field)."""

    def __init__(self, tolerance=1e-6, max_iterations=1000, device=None):
        self.tolerance = tolerance
        self.max_iterations = max_iterations
        self.residual_history = []
        self.inner_iterations_history = []
        self.total_inner_iterations = 0
        self.convergence_rates = []
        self._device = device
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context(self._device)
        return self._ctx

    def _stage(self, nx, ny, dx, dy, rho, u_star, v_star, d_u, d_v, b=None):
        """Uploads the inputs; returns (grid, b_dev, du_dev, dv_dev)."""
        ctx = self.ctx
        g = ctx.grid(nx, ny, dx, dy, rho)
        du = ctx.upload(d_u, nx, ny)
        dv = ctx.upload(d_v, nx, ny)
        if b is None:
            us = ctx.upload(u_star, nx, ny)
            vs = ctx.upload(v_star, nx, ny)
            bd = ctx.empty(nx, ny)
            ctx.check(ctx.lib.nf_continuity_rhs(ctx.handle, C.byref(g), ptr(us), ptr(vs), ptr(bd)), "nf_continuity_rhs")
        else:
            b2 = np.asarray(b, dtype=np.float64)
            if b2.ndim == 1:
                b2 = b2.reshape((nx, ny), order="F")
            bd = ctx.upload(b2, nx, ny)
        return g, bd, du, dv

    def _residual(self, g, x, b, du, dv):
        ctx = self.ctx
        r = ctx.empty(g.nx, g.ny)
        ctx.check(ctx.lib.nf_pressure_residual(ctx.handle, C.byref(g), ptr(x), ptr(b), ptr(du), ptr(dv), ptr(r)),
                  "nf_pressure_residual")
        return r

    def _norm(self, g, x, interior=False):
        ctx = self.ctx
        out = C.c_double()
        ctx.check(ctx.lib.nf_norm2(ctx.handle, C.byref(g), ptr(x), 1 if interior else 0, C.byref(out)), "nf_norm2")
        return out.value

    def get_solver_info(self):
        info = {"name": type(self).__name__, "inner_iterations_history": self.inner_iterations_history,
                "total_inner_iterations": self.total_inner_iterations}
        if self.convergence_rates:
            info["convergence_rate"] = sum(self.convergence_rates) / len(self.convergence_rates)
        info["solver_specific"] = {"tolerance": self.tolerance, "max_iterations": self.max_iterations}
        return info


class _StationaryBase(_GpuPressureBase):
    """Jacobi / red-black SOR share the reference's smoother protocol (gauss_seidel.py:55-57, jacobi.py:80-82)."""

    def _iterate(self, g, p, b, du, dv, n):
        raise NotImplementedError

    def solve(self, mesh=None, u_star=None, v_star=None, d_u=None, d_v=None, p_star=None, p=None, b=None,
              nx=None, ny=None, dx=None, dy=None, rho=1.0, num_iterations=None, track_residuals=True,
              return_dict=False):
        if mesh is not None:
            nx, ny = mesh.get_dimensions()
            dx, dy = mesh.get_cell_sizes()
        if num_iterations is None:
            num_iterations = self.max_iterations
        ctx = self.ctx
        g, bd, du, dv = self._stage(nx, ny, dx, dy, rho, u_star, v_star, d_u, d_v, b)
        if p is None:
            pd = ctx.empty(nx, ny)
        else:
            p2 = np.asarray(p, dtype=np.float64)
            if p2.ndim == 1:
                p2 = p2.reshape((nx, ny), order="F")
            pd = ctx.upload(p2, nx, ny)
        if track_residuals:
            self.residual_history = []
        inner = 0
        r = None
        if not track_residuals:
            self._iterate(g, pd, bd, du, dv, num_iterations)
            inner = num_iterations
        else:
            bn = self._norm(g, bd)
            for k in range(num_iterations):
                self._iterate(g, pd, bd, du, dv, 1)
                inner += 1
                r = self._residual(g, pd, bd, du, dv)
                rn = self._norm(g, r)
                self.residual_history.append(rn)
                if len(self.residual_history) >= 2 and self.residual_history[-2] > 0:
                    self.convergence_rates.append(rn / self.residual_history[-2])
                if bn > 0 and rn / bn < self.tolerance:
                    break
        self.inner_iterations_history.append(inner)
        self.total_inner_iterations += inner
        p_out = ctx.download(pd, nx, ny)
        if not return_dict:
            return p_out
        field = ctx.download(r, nx, ny) if r is not None else None
        return p_out, self._info(field, nx, ny, inner)


class GpuJacobiSolver(_StationaryBase):
    """Weighted Jacobi with the reference's boundary-doubled diagonal (jacobi.py:38-78)."""

    def __init__(self, tolerance=1e-6, max_iterations=1000, omega=1.0, device=None):
        super().__init__(tolerance, max_iterations, device)
        self.omega = omega

    def _iterate(self, g, p, b, du, dv, n):
        ctx = self.ctx
        key = (g.nx, g.ny)
        if getattr(self, "_tmp_key", None) != key:   # ping-pong scratch, kept across calls (no allocation per iteration)
            self._tmp, self._tmp_key = ctx.empty(g.nx, g.ny), key
        tmp = self._tmp
        ctx.check(ctx.lib.nf_jacobi_iterate(ctx.handle, C.byref(g), ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv),
                                            float(self.omega), int(n)), "nf_jacobi_iterate")

    def _info(self, field, nx, ny, inner):
        h = self.residual_history
        return {"rel_norm": 1.0 if not h else h[-1] / max(h[0], 1e-10),  # jacobi.py:213
                "abs_norm": h[-1] if h else 1.0, "iterations": inner, "field": field}


class GpuGaussSeidelSolver(_StationaryBase):
    """GaussSeidelSolver twin: red-black SOR (gauss_seidel.py:268-305) and the sequential 'standard' (lexicographic) /
    'symmetric' sweeps (:307-367), which run as anti-diagonal wavefronts with the loop's exact bits (nf_gs_lex.cu)."""

    def __init__(self, tolerance=1e-6, max_iterations=1000, omega=1.0, method_type="red_black", device=None):
        super().__init__(tolerance, max_iterations, device)
        if method_type not in ("red_black", "standard", "symmetric"):
            raise ValueError("method_type must be one of 'red_black', 'standard', or 'symmetric'")
        self.omega = omega
        self.method_type = method_type


    def solve(self, mesh=None, u_star=None, v_star=None, d_u=None, d_v=None, p_star=None, p=None, b=None,
              nx=None, ny=None, dx=None, dy=None, rho=1.0, num_iterations=None, track_residuals=True,
              return_dict=True):
        """Signature of ``GaussSeidelSolver.solve`` (gauss_seidel.py:55-57): unlike JacobiSolver it returns
        ``(p, info)`` by default; as a multigrid smoother it is called with ``return_dict=False`` (multigrid.py:352-354)."""
        return super().solve(mesh, u_star, v_star, d_u, d_v, p_star, p, b, nx, ny, dx, dy, rho, num_iterations,
                             track_residuals, return_dict)
    def _iterate(self, g, p, b, du, dv, n):
        ctx = self.ctx
        if self.method_type != "red_black":
            ctx.check(ctx.lib.nf_gs_lex_sweeps(ctx.handle, C.byref(g), ptr(p), ptr(b), ptr(du), ptr(dv), float(self.omega),
                                               int(n), int(self.method_type == "symmetric")), "nf_gs_lex_sweeps")
            return
        ctx.check(ctx.lib.nf_rbsor_sweeps(ctx.handle, C.byref(g), ptr(p), ptr(b), ptr(du), ptr(dv), float(self.omega),
                                          int(n)), "nf_rbsor_sweeps")

    def _info(self, field, nx, ny, inner):
        # ||r_interior|| / running maximum (gauss_seidel.py:189-200; stateful like the reference)
        cur = float(np.linalg.norm(field[1:nx - 1, 1:ny - 1])) if field is not None else 0.0
        self.p_max_l2 = max(getattr(self, "p_max_l2", cur), cur)
        return {"rel_norm": cur / self.p_max_l2 if self.p_max_l2 > 0 else 1.0, "field": field}


_CYCLE = {"v": 0, "w": 1, "fmg": 2}


class GpuMultiGridSolver(_GpuPressureBase):
    """Geometric multigrid with the reference's constructor (multigrid.py:31-37).  ``smoother`` is a
    GpuGaussSeidelSolver / GpuJacobiSolver (or the reference's own smoother object: only its class name,
    ``omega`` and ``method_type`` are read)."""

    def __init__(self, smoother, max_iterations=100, tolerance=1e-8, pre_smoothing=1, post_smoothing=1,
                 cycle_type="v", cycle_type_buildup="v", cycle_type_final=None, max_cycles_buildup=1,
                 restriction_method="restrict_full_weighting", interpolation_method="interpolate_linear",
                 coarsest_grid_size=7, device=None):
        super().__init__(tolerance, max_iterations, device)
        if coarsest_grid_size < 3:
            raise ValueError("Coarsest grid size must be at least 3")
        if coarsest_grid_size % 2 == 0:
            raise ValueError("Coarsest grid size must be odd")
        if restriction_method not in ("restrict_inject", "restrict_full_weighting"):
            raise ValueError("Restriction method must be one of: ['restrict_inject', 'restrict_full_weighting']")
        if interpolation_method not in ("interpolate_linear", "interpolate_cubic"):
            raise ValueError("Interpolation method must be one of: ['interpolate_linear', 'interpolate_cubic']")
        if cycle_type not in _CYCLE:
            raise ValueError("cycle_type must be 'v', 'w' or 'fmg'")
        name = type(smoother).__name__.lower()
        if "jacobi" in name:
            self._smoother_id = 1
        elif "gauss" in name or "seidel" in name:
            mt = getattr(smoother, "method_type", "red_black")
            if mt not in ("red_black", "standard", "symmetric"):
                raise ValueError("method_type must be one of 'red_black', 'standard', or 'symmetric'")
            self._smoother_id = {"red_black": 0, "standard": 2, "symmetric": 3}[mt]
        else:
            raise ValueError(f"unsupported smoother {type(smoother).__name__}")
        self.smoother = smoother
        self.smoother_omega = getattr(smoother, "omega", 1.0)
        self.pre_smoothing, self.post_smoothing = pre_smoothing, post_smoothing
        self.cycle_type, self.cycle_type_buildup, self.cycle_type_final = cycle_type, cycle_type_buildup, cycle_type_final
        self.max_cycles_buildup = max_cycles_buildup
        self.restriction_method, self.interpolation_method = restriction_method, interpolation_method
        self.coarsest_grid_size = coarsest_grid_size
        self.rho = 1.0
        self._mg = None
        self._mg_key = None

    def config_struct(self, length=1.0, height=1.0):
        c = NfMgConfig()
        c.smoother = self._smoother_id
        c.pre, c.post = int(self.pre_smoothing), int(self.post_smoothing)
        c.cycle_type = _CYCLE[self.cycle_type]
        c.cycle_buildup = _CYCLE.get(self.cycle_type_buildup, 0)
        c.cycle_final = -1 if self.cycle_type_final is None else _CYCLE[self.cycle_type_final]
        c.max_cycles_buildup = int(self.max_cycles_buildup)
        c.restriction = 0 if self.restriction_method == "restrict_full_weighting" else 1
        c.interpolation = 0 if self.interpolation_method == "interpolate_linear" else 1
        c.coarsest = int(self.coarsest_grid_size)
        c.max_iterations = int(self.max_iterations)
        c.omega = float(self.smoother_omega)
        c.tolerance = float(self.tolerance)
        c.length, c.height, c.rho = float(length), float(height), float(self.rho)
        return c

    def _hierarchy(self, nx, ny, length, height):
        key = (nx, ny, length, height)
        if self._mg_key != key:
            self._free()
            ctx = self.ctx
            cfg = self.config_struct(length, height)
            h = C.c_void_p()
            ctx.check(ctx.lib.nf_mg_create(ctx.handle, C.byref(h), nx, ny, pad_ld(ny), C.byref(cfg)), "nf_mg_create")
            self._mg, self._mg_key = h, key
        return self._mg

    def _free(self):
        if self._mg is not None:
            self.ctx.lib.nf_mg_destroy(self._mg)
            self._mg, self._mg_key = None, None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def solve(self, mesh, u_star, v_star, d_u, d_v, p_star, return_dict=True):
        nx, ny, dx, dy, length, height = mesh_scalars(mesh)
        ctx = self.ctx
        g, bd, du, dv = self._stage(nx, ny, dx, dy, self.rho, u_star, v_star, d_u, d_v)
        mg = self._hierarchy(nx, ny, length, height)
        x = ctx.empty(nx, ny)
        r = ctx.empty(nx, ny)
        info = NfMgInfo()
        ctx.check(ctx.lib.nf_mg_setup(mg, ptr(du), ptr(dv)), "nf_mg_setup")
        ctx.check(ctx.lib.nf_mg_solve(mg, ptr(bd), ptr(x), ptr(r), C.byref(info)), "nf_mg_solve")
        self.last_info = info
        self.residual_history.append(info.r_norm / info.b_norm if info.b_norm > 0 else info.r_norm)
        self.inner_iterations_history.append(info.cycles)
        self.total_inner_iterations += info.cycles
        p_out = ctx.download(x, nx, ny)
        if not return_dict:
            return p_out
        return p_out, {"rel_norm": info.r_norm, "field": ctx.download(r, nx, ny)}  # absolute ||r|| (multigrid.py:257)


class _KrylovBase(_GpuPressureBase):
    _fn = None
    _nwork = 0

    def __init__(self, tolerance=1e-7, max_iterations=1000, use_preconditioner=False, preconditioner="jacobi",
                 mg_pre_smoothing=2, mg_post_smoothing=2, mg_cycles=1, mg_cycle_type="v", mg_cycle_type_buildup="v",
                 mg_max_cycles_buildup=1, mg_coarsest_grid_size=7, mg_restriction_method="restrict_full_weighting",
                 mg_interpolation_method="interpolate_linear", smoother_relaxation=0.8,
                 smoother_method_type="red_black", check_every=10, device=None):
        """Constructor of MatrixFreeBiCGSTABSolver (matrix_free_BiCGSTAB.py:20-100)."""
        super().__init__(tolerance, max_iterations, device)
        self.use_preconditioner = bool(use_preconditioner)
        self.preconditioner = preconditioner
        self.check_every = check_every
        self.mg_cycles = mg_cycles
        self.mg_cycle_type = "fmg" if mg_cycle_type in ("f", "fmg") else mg_cycle_type
        self.mg_precond = None
        if self.use_preconditioner:
            if preconditioner != "multigrid":
                # the reference's 'jacobi' path calls an undefined method (matrix_free_BiCGSTAB.py:229)
                raise NotImplementedError("only preconditioner='multigrid' exists (the reference's 'jacobi' path is broken)")
            if self._fn != "nf_bicgstab_solve":
                raise NotImplementedError("the multigrid preconditioner is implemented for BiCGSTAB")
            smoother = GpuGaussSeidelSolver(omega=smoother_relaxation, method_type=smoother_method_type)
            self.omega = smoother.omega
            self.mg_precond = GpuMultiGridSolver(
                smoother=smoother, tolerance=tolerance, max_iterations=1, pre_smoothing=mg_pre_smoothing,
                post_smoothing=mg_post_smoothing, cycle_type=self.mg_cycle_type, cycle_type_buildup=mg_cycle_type_buildup,
                max_cycles_buildup=mg_max_cycles_buildup, coarsest_grid_size=mg_coarsest_grid_size,
                restriction_method=mg_restriction_method, interpolation_method=mg_interpolation_method, device=device)

    def solve(self, mesh, u_star, v_star, d_u, d_v, p_star, return_dict=True):
        nx, ny, dx, dy, length, height = mesh_scalars(mesh)
        ctx = self.ctx
        torch = ctx.torch
        g, bd, du, dv = self._stage(nx, ny, dx, dy, 1.0, u_star, v_star, d_u, d_v)
        x = ctx.empty(nx, ny)
        nwork = 7 if self.mg_precond is not None else self._nwork
        work = torch.zeros((nwork * (nx + 1), pad_ld(ny)), dtype=torch.float64, device=x.device)
        info = NfKrylovInfo()
        # scipy's default rtol = 1e-5 governs: the reference never passes rtol (matrix_free_BiCGSTAB.py:234-242)
        if self.mg_precond is not None:
            mg = self.mg_precond._hierarchy(nx, ny, length, height)
            ctx.check(ctx.lib.nf_mg_setup(mg, ptr(du), ptr(dv)), "nf_mg_setup")
            ctx.check(ctx.lib.nf_bicgstab_solve_mg(ctx.handle, C.byref(g), ptr(bd), ptr(x), ptr(du), ptr(dv),
                                                   float(self.tolerance), 1e-5, int(self.max_iterations), 1, ptr(work), mg,
                                                   int(self.mg_cycles), _CYCLE[self.mg_cycle_type], C.byref(info)),
                      "nf_bicgstab_solve_mg")
        else:
            fn = getattr(ctx.lib, self._fn)
            ctx.check(fn(ctx.handle, C.byref(g), ptr(bd), ptr(x), ptr(du), ptr(dv), float(self.tolerance), 1e-5,
                         int(self.max_iterations), int(self.check_every), ptr(work), C.byref(info)), self._fn)
        self.last_info = info
        self.inner_iterations_history.append(info.iterations)
        self.total_inner_iterations += info.iterations
        if info.info != 0:
            print(f"Warning: {type(self).__name__} did not converge, info={info.info}")
        # rel_norm = ||r_interior|| / ||b_interior|| of the true residual (matrix_free_BiCGSTAB.py:255-279)
        r = self._residual(g, x, bd, du, dv)
        rn = self._norm(g, r, interior=True)
        bn = self._norm(g, bd, interior=True)
        p_out = ctx.download(x, nx, ny)
        if not return_dict:
            return p_out
        field = ctx.download(r, nx, ny)
        field[0, :] = 0; field[-1, :] = 0; field[:, 0] = 0; field[:, -1] = 0
        return p_out, {"rel_norm": rn / bn if bn > 0 else rn, "field": field, "iterations": info.iterations}


class GpuCGSolver(_KrylovBase):
    """scipy ``cg`` on the matrix-free operator, restated on the device (operation order of
    scipy/sparse/linalg/_isolve/iterative.py)."""
    _fn = "nf_cg_solve"
    _nwork = 4


class GpuBiCGSTABSolver(_KrylovBase):
    """MatrixFreeBiCGSTABSolver twin (scipy ``bicgstab`` operation order)."""
    _fn = "nf_bicgstab_solve"
    _nwork = 5


class GpuGeoMultigridPrecondCGSolver(_GpuPressureBase):
    """Twin of ``GeoMultigridPrecondCGSolver`` (pressure_solver/geo_multigrid_cg.py:16-258): scipy's ``cg`` with M =
    ``mg_cycles`` multigrid cycles from zero, ``atol = tolerance`` (scipy's default ``rtol = 1e-5`` governs as well).  Same
    constructor and, like the reference, ``solve`` returns the bare ``p'`` array (it does not plug into the outer loops,
    which expect ``(p', info)``).  The reference's default ``mg_cycle_type='f'`` is not a cycle type its MultiGridSolver
    knows (multigrid.py:96-99): the ValueError is raised here at the same place; a V-cycle needs a ``smoother`` object."""

    def __init__(self, tolerance=1e-5, max_iterations=500, mg_pre_smoothing=3, mg_post_smoothing=2, mg_cycles=1,
                 mg_cycle_type="f", mg_cycle_type_buildup="w", mg_max_cycles_buildup=1, mg_coarsest_grid_size=7,
                 mg_restriction_method="restrict_inject", mg_interpolation_method="interpolate_cubic", smoother=None,
                 device=None):
        super().__init__(tolerance, max_iterations, device)
        self.inner_iterations = []
        self.mg_cycles = mg_cycles
        self.mg_precond = GpuMultiGridSolver(
            smoother=smoother, tolerance=tolerance * 0.1, max_iterations=1, pre_smoothing=mg_pre_smoothing,
            post_smoothing=mg_post_smoothing, cycle_type=mg_cycle_type, cycle_type_buildup=mg_cycle_type_buildup,
            max_cycles_buildup=mg_max_cycles_buildup, coarsest_grid_size=mg_coarsest_grid_size,
            restriction_method=mg_restriction_method, interpolation_method=mg_interpolation_method, device=device)
        self.omega = getattr(smoother, "omega", 1.0) if smoother else 1.0

    def solve(self, mesh, u_star, v_star, d_u, d_v, p_star):
        nx, ny, dx, dy, length, height = mesh_scalars(mesh)
        ctx = self.ctx
        torch = ctx.torch
        g, bd, du, dv = self._stage(nx, ny, dx, dy, 1.0, u_star, v_star, d_u, d_v)
        x = ctx.empty(nx, ny)
        work = torch.zeros((5 * (nx + 1), pad_ld(ny)), dtype=torch.float64, device=x.device)
        info = NfKrylovInfo()
        mg = self.mg_precond._hierarchy(nx, ny, length, height)
        ctx.check(ctx.lib.nf_mg_setup(mg, ptr(du), ptr(dv)), "nf_mg_setup")
        ctx.check(ctx.lib.nf_cg_solve_mg(ctx.handle, C.byref(g), ptr(bd), ptr(x), ptr(du), ptr(dv), float(self.tolerance), 1e-5,
                                         int(self.max_iterations), ptr(work), mg, int(self.mg_cycles),
                                         _CYCLE[self.mg_precond.cycle_type], C.byref(info)), "nf_cg_solve_mg")
        self.last_info = info
        self.inner_iterations.append(info.iterations)
        self.inner_iterations_history.append(info.iterations)
        self.total_inner_iterations += info.iterations
        if info.info != 0:
            print(f"Warning: Geo-Multigrid Preconditioned CG did not converge, info={info.info}")
        return ctx.download(x, nx, ny)

    def get_solver_info(self):
        return {"name": "GeoMultigridPrecondCGSolver", "inner_iterations_history": self.inner_iterations,
                "total_inner_iterations": sum(self.inner_iterations), "convergence_rate": None,
                "solver_specific": {"method": "conjugate_gradient", "preconditioner": "geometric_multigrid",
                                    "pre_smoothing": self.mg_precond.pre_smoothing,
                                    "post_smoothing": self.mg_precond.post_smoothing,
                                    "cycle_type": self.mg_precond.cycle_type}}
