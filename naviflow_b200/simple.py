"""GpuSimpleSolver -- the SIMPLE outer loop, device resident.

Twin of ``SimpleSolver(BaseAlgorithm)`` (solver/Algorithms/simple.py:16-269, base_algorithm.py:13-235 of
/root/reference/naviflow_oo): same constructor, ``set_boundary_condition``, ``solve`` signature, result
histories (``u_rel_norm``, ``v_rel_norm``, ``p_rel_norm``, ``total_rel_norm``) and stopping test
(``max(u_rel_norm, v_rel_norm) > tolerance``).  The whole iteration runs in libnaviflow_b200
(``nf_simple_iterate``): fields are uploaded once at ``solve()`` entry and downloaded once at exit.

The plugin objects select the device kernels:
  momentum_solver  GpuJacobiMomentumSolver(n_jacobi_sweeps)        -> fixed Jacobi sweeps
  pressure_solver  GpuMultiGridSolver | GpuJacobiSolver | GpuGaussSeidelSolver | GpuCGSolver | GpuBiCGSTABSolver
  velocity_updater GpuVelocityUpdater (or None)
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from ._lib import NfSimpleConfig, NfSimpleInfo
from .device import bc_program_struct, get_context, mesh_scalars
from .host import BoundaryConditionManager, SimulationResult, ghia_errors, practice_b_sides
from .momentum import GpuJacobiMomentumSolver, GpuMatrixFreeMomentumSolver
from .pressure import (GpuBiCGSTABSolver, GpuCGSolver, GpuGaussSeidelSolver, GpuJacobiSolver,
                       GpuMultiGridSolver)
from .profiler import Profiler
from .velocity import GpuVelocityUpdater

_FIELD = {"u": 0, "v": 1, "p": 2, "u_star": 3, "v_star": 4, "d_u": 5, "d_v": 6, "p_prime": 7, "b": 8,
          "p_res": 9, "u_res": 10, "v_res": 11}


def slab_rows(nx, world, rank):
    """Cell rows [begin, end) rank `rank` owns when nx rows are cut over `world` ranks (host-only query of the
    library's partition rule; works without a GPU)."""
    from . import _lib
    b, e = C.c_int(), C.c_int()
    st = _lib.lib().nf_slab_rows(int(nx), int(world), int(rank), C.byref(b), C.byref(e))
    if st != 0:
        raise ValueError(f"nf_slab_rows({nx}, {world}, {rank}) failed with status {st}")
    return b.value, e.value


def assemble_rows(arr, begin, end, device=None):
    """Every rank holds valid rows [begin, end) of the same full-size array (the last rank up to the end): sum the
    zero-padded pieces over the process group so that all ranks end with the complete array (in place)."""
    import torch
    import torch.distributed as dist
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(arr.shape, dtype=torch.float64, device=device)
    t[begin:end].copy_(torch.from_numpy(np.ascontiguousarray(arr[begin:end])))
    dist.all_reduce(t)
    arr[...] = t.cpu().numpy()
    return arr


class GpuSimpleSolver:
    _piso_corrections = 0          # 0: SIMPLE; GpuPisoSolver sets n_corrections
    _history_appends_per_iteration = 2  # the reference appends twice (simple.py:177, :196)
    _profile_prefix = "SIMPLE"     # file name of the saved run record (simple.py:265)

    def __init__(self, mesh, fluid, pressure_solver=None, momentum_solver=None, velocity_updater=None,
                 boundary_conditions=None, alpha_p=0.3, alpha_u=0.7, fix_lid_corners=False, device=None,
                 distributed=None, virtual_ranks=1, track_unrelaxed_residual=False):
        """``distributed``: cut the grid into row slabs over the ranks of the initialised torch.distributed
        (NCCL) process group (default: automatically when its world size is > 1).  ``virtual_ranks`` > 1 cuts
        the grid into that many slabs inside this process on this device (same code path; used by the tests)."""
        self.mesh, self.fluid = mesh, fluid
        self.pressure_solver = pressure_solver
        self.momentum_solver = momentum_solver if momentum_solver is not None else GpuJacobiMomentumSolver()
        self.velocity_updater = velocity_updater if velocity_updater is not None else GpuVelocityUpdater()
        self.alpha_p, self.alpha_u = alpha_p, alpha_u
        self.fix_lid_corners = fix_lid_corners
        if boundary_conditions is not None and hasattr(boundary_conditions, "apply_velocity_boundary_conditions"):
            self.bc_manager = boundary_conditions
        else:
            self.bc_manager = BoundaryConditionManager()
            if isinstance(boundary_conditions, dict):
                for loc, conds in boundary_conditions.items():
                    for typ, vals in conds.items():
                        self.bc_manager.set_condition(loc, typ, vals)
        self.boundary_conditions = self.bc_manager.to_dict()
        self.profiler = Profiler(self.__class__.__name__, mesh, fluid, algorithm=self)  # base_algorithm.py:60
        self.residual_history = []
        self._device = device
        self._state = None
        self._state_key = None
        self._team = None
        self._virtual_ranks = int(virtual_ranks)
        # every iteration record then carries ||S_un - A_un u*||, ||S_un - A_un v*|| (``unrelaxed_residual_history``): the
        # outer loop's convergence measure when the momentum predictor is the fixed-sweep Jacobi solver
        self.track_unrelaxed_residual = bool(track_unrelaxed_residual)
        self.unrelaxed_residual_history = []
        self._distributed = distributed
        self._p_residual_cache = None
        self._final_u_residual_field = self._final_v_residual_field = None
        self.initialize_fields()

    # ---- BaseAlgorithm interface ------------------------------------------------------------------
    _PIN_WHOLE_LIMIT = 1 << 30   # bytes: larger host fields are not page-locked as a whole (see _pin_local_rows)

    @staticmethod
    def _host_zeros(shape):
        """Host field; page-locked when a CUDA device is present so that solve()'s H2D/D2H run at PCIe speed.  Fields
        above 1 GiB are allocated pageable here and only the rows this process transfers are registered later
        (_pin_local_rows): on N ranks every rank moves 1/N of the rows."""
        try:
            import torch
            if torch.cuda.is_available() and int(np.prod(shape)) * 8 <= GpuSimpleSolver._PIN_WHOLE_LIMIT:
                t = torch.zeros(shape, dtype=torch.float64).pin_memory()
                return t.numpy()  # the array keeps the pinned tensor alive through its base
        except Exception:
            pass
        return np.zeros(shape)

    def _pin_local_rows(self):
        """cudaHostRegister of the rows [row0, row1) of u, v, p that nf_simple_upload / nf_simple_download touch for this
        process's slab (whole arrays on a single rank), for host fields that are not page-locked yet.  Registered
        ranges are remembered per array and released in _free()."""
        if self._state is None:
            return
        try:
            import torch
            rt = torch.cuda.cudart()
        except Exception:
            return
        if not hasattr(self, "_registered"):
            self._registered = {}
        b, e = self.local_rows()
        halo = 8
        for name in ("u", "v", "p"):
            arr = getattr(self, name)
            if arr.nbytes <= self._PIN_WHOLE_LIMIT and getattr(arr, "base", None) is not None:
                continue  # page-locked as a whole by _host_zeros (a view of the pinned tensor)
            if not (arr.flags.c_contiguous and arr.dtype == np.float64):
                continue
            lo = max(b - halo, 0)
            hi = min(e + halo + 1, arr.shape[0])
            row_bytes = arr.strides[0]
            start = arr.ctypes.data + lo * row_bytes
            end = arr.ctypes.data + hi * row_bytes
            start_al = start & ~4095
            end_al = (end + 4095) & ~4095
            key = (arr.ctypes.data, arr.nbytes)
            if self._registered.get(name, (None,))[0] == key:
                continue
            self._unregister(name)
            try:
                err = rt.cudaHostRegister(start_al, end_al - start_al, 0)
                if int(err) == 0:
                    self._registered[name] = (key, start_al)
            except Exception:
                pass

    def _unregister(self, name=None):
        reg = getattr(self, "_registered", None)
        if not reg:
            return
        try:
            import torch
            rt = torch.cuda.cudart()
        except Exception:
            return
        for nm in ([name] if name else list(reg)):
            ent = reg.pop(nm, None)
            if ent:
                try:
                    rt.cudaHostUnregister(ent[1])
                except Exception:
                    pass

    def initialize_fields(self):
        nx, ny = self.mesh.get_dimensions()
        self.p = self._host_zeros((nx, ny))
        self.u = self._host_zeros((nx + 1, ny))
        self.v = self._host_zeros((nx, ny + 1))
        self.apply_boundary_conditions()

    def apply_boundary_conditions(self):
        nx, ny = self.mesh.get_dimensions()
        self.u, self.v = self.bc_manager.apply_velocity_boundary_conditions(self.u, self.v, nx, ny)

    def set_boundary_condition(self, boundary, condition_type, values=None):
        self.bc_manager.set_condition(boundary, condition_type, values)
        self.boundary_conditions = self.bc_manager.to_dict()
        self.apply_boundary_conditions()

    @property
    def _final_p_residual_field(self):
        """Pressure residual field of the last iteration (simple.py:217-219), downloaded on first access."""
        if self._p_residual_cache is None and self._state is not None:
            ctx, st = self._ensure_state()
            nx, ny = self.mesh.get_dimensions()
            self._p_residual_cache = self._download(ctx, st, "p_res", nx, ny)
        return self._p_residual_cache

    @_final_p_residual_field.setter
    def _final_p_residual_field(self, value):
        self._p_residual_cache = value

    def get_max_divergence(self):
        dx, dy = self.mesh.get_cell_sizes()
        div = (self.u[1:, :] - self.u[:-1, :]) / dx + (self.v[:, 1:] - self.v[:, :-1]) / dy
        return float(np.max(np.abs(div[1:-1, 1:-1])))

    # ---- device state -----------------------------------------------------------------------------
    def _config(self):
        nx, ny, dx, dy, length, height = mesh_scalars(self.mesh)
        c = NfSimpleConfig()
        c.nx, c.ny = nx, ny
        ms = self.momentum_solver
        c.momentum_solver, c.momentum_maxiter, c.momentum_tolerance, c.n_momentum_sweeps = 0, 0, 0.0, 0
        if isinstance(ms, GpuJacobiMomentumSolver):
            c.n_momentum_sweeps = ms.n_jacobi_sweeps
        elif isinstance(ms, GpuMatrixFreeMomentumSolver):
            c.momentum_solver, c.momentum_maxiter, c.momentum_tolerance = 1, ms.maxiter, ms.tol
        else:
            raise TypeError("GpuSimpleSolver needs a GpuJacobiMomentumSolver or a GpuMatrixFreeMomentumSolver")
        ps = self.pressure_solver
        c.krylov_maxiter = 0
        c.piso_corrections = int(self._piso_corrections)
        c.simplec_divisor = 0.0
        c.track_unrelaxed_residual = 1 if self.track_unrelaxed_residual else 0
        c.pressure_iterations = 0
        c.pressure_omega = 1.0
        c.pressure_tolerance = 0.0
        if isinstance(ps, GpuMultiGridSolver):
            c.pressure_solver = 0
            c.mg = ps.config_struct(length, height)
        elif isinstance(ps, (GpuJacobiSolver, GpuGaussSeidelSolver)):
            if ps.tolerance > 0:
                raise NotImplementedError(
                    "device-resident Jacobi / SOR pressure solves run a fixed number of iterations: pass tolerance=0")
            c.pressure_solver = 1 if isinstance(ps, GpuJacobiSolver) else \
                {"red_black": 2, "standard": 5, "symmetric": 6}[ps.method_type]
            c.pressure_iterations = int(ps.max_iterations)
            c.pressure_omega = float(ps.omega)
        elif isinstance(ps, (GpuCGSolver, GpuBiCGSTABSolver)):
            c.pressure_solver = 3 if isinstance(ps, GpuCGSolver) else 4
            c.krylov_maxiter = int(ps.max_iterations)
            c.pressure_tolerance = float(ps.tolerance)
            c.krylov_check_every = int(ps.check_every)
            if getattr(ps, "mg_precond", None) is not None:
                # MatrixFreeBiCGSTABSolver(use_preconditioner=True, preconditioner='multigrid'): same recurrence with
                # M = mg_cycles multigrid cycles (matrix_free_BiCGSTAB.py:102-161); the host polls every iteration
                from .pressure import _CYCLE
                c.pressure_solver = 7
                c.mg = ps.mg_precond.config_struct(length, height)
                c.krylov_mg_cycles = int(ps.mg_cycles)
                c.krylov_mg_kind = _CYCLE[ps.mg_cycle_type]
                c.krylov_check_every = 1
        else:
            raise TypeError("pressure_solver must be one of the naviflow_b200 Gpu*Solver classes")
        c.sides = practice_b_sides(self.bc_manager)
        c.length, c.height = length, height
        c.rho, c.mu = float(self.fluid.get_density()), float(self.fluid.get_viscosity())
        c.alpha_p, c.alpha_u = float(self.alpha_p), float(self.alpha_u)
        c.bc = bc_program_struct(self.bc_manager, nx, ny)
        c.bc_mf = bc_program_struct(self.bc_manager, nx, ny, nx + 1)
        return c

    def _world(self):
        """(world, rank) of the process group the grid is cut over, (1, 0) when not distributed."""
        if self._distributed is False:
            return 1, 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                return dist.get_world_size(), dist.get_rank()
        except Exception:
            pass
        return 1, 0

    def _make_team(self, ctx):
        world, rank = self._world()
        team = C.c_void_p()
        if world > 1:
            import torch
            import torch.distributed as dist
            ident = (C.c_ubyte * 128)()
            if rank == 0:
                ctx.check(ctx.lib.nf_nccl_unique_id(ctx.handle, ident), "nf_nccl_unique_id")
            t = torch.tensor(list(ident), dtype=torch.uint8, device=f"cuda:{ctx.device}")
            dist.broadcast(t, src=0)
            ident = (C.c_ubyte * 128)(*t.cpu().tolist())
            ctx.check(ctx.lib.nf_team_create_nccl(ctx.handle, world, rank, ident, C.byref(team)), "nf_team_create_nccl")
        elif self._virtual_ranks > 1:
            ctx.check(ctx.lib.nf_team_create_virtual(ctx.handle, self._virtual_ranks, C.byref(team)),
                      "nf_team_create_virtual")
        else:
            return None
        return team

    def _ensure_state(self):
        ctx = get_context(self._device)
        cfg = self._config()
        key = bytes(cfg)
        if self._state is None or key != self._state_key:
            self._free()
            h = C.c_void_p()
            self._team = self._make_team(ctx)
            if self._team is None:
                ctx.check(ctx.lib.nf_simple_create(ctx.handle, C.byref(h), C.byref(cfg)), "nf_simple_create")
            else:
                ctx.check(ctx.lib.nf_simple_create_team(self._team, C.byref(h), C.byref(cfg)), "nf_simple_create_team")
            self._state, self._state_key = h, key
        return ctx, self._state

    def close(self):
        """Releases the device state.  COLLECTIVE on a distributed solver (the peer-memory arena is torn down by all
        ranks together, nf_team_free): every rank must call it (or drop the solver) at the same point."""
        self._free()

    def _free(self):
        self._unregister()
        if self._state is not None:
            lib = get_context(self._device).lib
            lib.nf_simple_destroy(self._state)
            if self._team is not None:
                lib.nf_team_free(self._team)
            self._state, self._state_key, self._team = None, None, None

    def uses_p2p(self):
        """True when the slab exchanges of this solver run as peer-memory kernels over NVLink (nf_p2p.cu), False when
        they go through NCCL or when the solver is not cut over processes."""
        ctx, _ = self._ensure_state()
        return self._team is not None and bool(ctx.lib.nf_team_uses_p2p(self._team))

    def local_rows(self):
        """Cell rows [begin, end) this process owns (the whole grid unless distributed over processes)."""
        ctx, st = self._ensure_state()
        b, e = C.c_int(), C.c_int()
        ctx.check(ctx.lib.nf_simple_local_rows(st, 0, C.byref(b), C.byref(e)), "nf_simple_local_rows")
        if self._virtual_ranks > 1 and self._world()[0] == 1:
            return 0, self.mesh.get_dimensions()[0]
        return b.value, e.value

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def _upload(self, ctx, st, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        ctx.check(ctx.lib.nf_simple_upload(st, _FIELD[name], a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1]),
                  "nf_simple_upload")

    def _download(self, ctx, st, name, rows, cols, out=None):
        if out is None or out.shape != (rows, cols) or out.dtype != np.float64 or not out.flags.c_contiguous:
            out = np.empty((rows, cols), dtype=np.float64)
        ctx.check(ctx.lib.nf_simple_download(st, _FIELD[name], out.ctypes.data_as(C.c_void_p), rows, cols),
                  "nf_simple_download")
        return out

    def device_field(self, name):
        """Raw device pointer of a resident field (after the first solve())."""
        ctx, st = self._ensure_state()
        return ctx.lib.nf_simple_field(st, _FIELD[name])

    def iterate_resident(self, n_iterations, tolerance=0.0, want_fields=False):
        """Runs outer iterations on the resident state without touching the host fields.  Returns the list
        of per-iteration records (dicts).  Used by solve() and by bench.py."""
        ctx, st = self._ensure_state()
        infos = (NfSimpleInfo * max(n_iterations, 1))()
        done = C.c_int(0)
        ctx.check(ctx.lib.nf_simple_iterate(st, int(n_iterations), float(tolerance), 1 if want_fields else 0, infos,
                                            C.byref(done)), "nf_simple_iterate")
        return [dict(u_rel_norm=r.u_rel_norm, v_rel_norm=r.v_rel_norm, p_rel_norm=r.p_rel_norm,
                     u_abs_res=r.u_abs_res, v_abs_res=r.v_abs_res, pressure_iterations=r.pressure_iterations,
                     u_unrelaxed_res=r.u_unrelaxed_res, v_unrelaxed_res=r.v_unrelaxed_res)
                for r in infos[: done.value]]

    def smoother_timing(self, on):
        """Switches the live CUDA-event timing of the finest-level smoother launches on/off; returns the
        (total ms, launches) accumulated since the last call."""
        ctx, st = self._ensure_state()
        ms, cnt = C.c_double(), C.c_longlong()
        ctx.check(ctx.lib.nf_simple_smoother_timing(st, 1 if on else 0, C.byref(ms), C.byref(cnt)),
                  "nf_simple_smoother_timing")
        return ms.value, cnt.value

    def phase_timing(self, on):
        """Switches the CUDA-event timing of the iteration's phases on/off; returns the accumulated
        ``{'momentum_ms', 'pressure_ms', 'correct_ms', 'iterations'}`` since the last call."""
        ctx, st = self._ensure_state()
        a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_longlong()
        ctx.check(ctx.lib.nf_simple_phase_timing(st, 1 if on else 0, C.byref(a), C.byref(b), C.byref(c), C.byref(n)),
                  "nf_simple_phase_timing")
        return {"momentum_ms": a.value, "pressure_ms": b.value, "correct_ms": c.value, "iterations": n.value}

    def push_fields(self):
        ctx, st = self._ensure_state()
        self._pin_local_rows()
        for name in ("u", "v", "p"):
            self._upload(ctx, st, name, getattr(self, name))

    def pull_fields(self, gather=True):
        ctx, st = self._ensure_state()
        nx, ny = self.mesh.get_dimensions()
        self.u = self._download(ctx, st, "u", nx + 1, ny, self.u)
        self.v = self._download(ctx, st, "v", nx, ny + 1, self.v)
        self.p = self._download(ctx, st, "p", nx, ny, self.p)
        if gather and self._world()[0] > 1:  # every rank holds its own rows: assemble the full fields everywhere
            self.u, self.v, self.p = (self._assemble(a) for a in (self.u, self.v, self.p))

    def _assemble(self, arr):
        import torch
        import torch.distributed as dist
        world, rank = self._world()
        b, e = self.local_rows()
        if rank == world - 1:
            e = arr.shape[0]
        return assemble_rows(arr, b, e, device=f"cuda:{get_context(self._device).device}")

    # ---- SimpleSolver.solve -----------------------------------------------------------------------
    def save_profiling_data(self, filename=None, profile_dir="results/profiles"):
        """base_algorithm.py:200-216"""
        return self.profiler.save(filename, profile_dir)

    def solve(self, max_iterations=1000, tolerance=1e-6, save_profile=True, profile_dir="results/profiles",
              track_infinity_norm=False, infinity_norm_interval=10, use_l2_norm=False, chunk=None, gather=True):
        """Signature and defaults of ``SimpleSolver.solve`` (simple.py:78-79), including ``save_profile=True``: the run
        record (naviflow_b200/profiler.py, the reference's HDF5 layout) goes to ``profile_dir``.  Extra keywords:
        ``chunk`` = outer iterations per device call, ``gather=False`` (distributed runs): every rank keeps only its own
        rows in ``self.u/v/p`` instead of assembling the full fields on all ranks at the end."""
        self.profiler.initialize()
        self.profiler.start()
        t0 = time.perf_counter()
        ctx, st = self._ensure_state()
        nx, ny = self.mesh.get_dimensions()
        self.push_fields()
        self.residual_history = []
        self.x_momentum_rel_norms, self.y_momentum_rel_norms, self.pressure_rel_norms = [], [], []
        self.infinity_norm_history = []
        self.pressure_iterations_history = []
        self._u_abs_res_history = []
        self.unrelaxed_residual_history = []
        iteration = 1
        total = 1.0
        if chunk is None:
            chunk = infinity_norm_interval if track_infinity_norm else (1 if tolerance > 0 else max_iterations)
        try:
            while iteration <= max_iterations and total > tolerance:
                n = min(chunk, max_iterations - iteration + 1)
                t_chunk = time.perf_counter() - t0
                recs = self.iterate_resident(n, tolerance, want_fields=False)
                if recs:
                    self.profiler.add_residual_block(
                        iteration, [max(r["u_rel_norm"], r["v_rel_norm"]) for r in recs],
                        [max(r["u_rel_norm"], r["v_rel_norm"]) if self._piso_corrections != -2 else r["u_abs_res"]
                         for r in recs], [r["p_rel_norm"] for r in recs], t_chunk, time.perf_counter() - t0)
                for r in recs:
                    total = max(r["u_rel_norm"], r["v_rel_norm"])
                    self.x_momentum_rel_norms.append(r["u_rel_norm"])
                    self.y_momentum_rel_norms.append(r["v_rel_norm"])
                    self.pressure_rel_norms.append(r["p_rel_norm"])
                    self.pressure_iterations_history.append(r["pressure_iterations"])
                    self._u_abs_res_history.append(r["u_abs_res"])
                    self.unrelaxed_residual_history.append(max(r["u_unrelaxed_res"], r["v_unrelaxed_res"]))
                    self.residual_history.extend([total] * self._history_appends_per_iteration)
                iteration += len(recs)
                if track_infinity_norm and (iteration - 1) % infinity_norm_interval == 0:
                    self.pull_fields()
                    inf, l2 = ghia_errors(self.u, self.v, self.mesh, self.fluid.get_reynolds_number())
                    self.infinity_norm_history.append(l2 if use_l2_norm else inf)
                if len(recs) < n:
                    break
        except KeyboardInterrupt:
            print("Interrupted by user.")
        self.pull_fields(gather)
        self._p_residual_cache = None  # the residual field stays on the device until somebody asks for it
        final_residual = total
        self.profiler.set_iterations(iteration - 1)
        self.profiler.set_convergence_info(tolerance=tolerance, final_residual=final_residual,
                                           residual_history=self.residual_history, converged=(final_residual < tolerance))
        self.profiler.set_pressure_solver_info(  # simple.py:233-240; the device loop counts the inner iterations itself
            solver_name=type(self.pressure_solver).__name__, inner_iterations=self.pressure_iterations_history)
        self.profiler.end()
        result = SimulationResult(self.u, self.v, self.p, self.mesh, iterations=iteration - 1,
                                  residuals=self.residual_history, reynolds=self.fluid.get_reynolds_number(),
                                  wall_time=time.perf_counter() - t0)
        result.add_history("u_rel_norm", self.x_momentum_rel_norms)
        result.add_history("v_rel_norm", self.y_momentum_rel_norms)
        result.add_history("p_rel_norm", self.pressure_rel_norms)
        result.add_history("total_rel_norm", self.residual_history)
        if self.infinity_norm_history:
            result.add_history("infinity_norm_error", self.infinity_norm_history)
        import os
        if save_profile and self._world()[1] == 0 and os.environ.get("NAVIFLOW_B200_NO_PROFILE_FILES") != "1":
            os.makedirs(profile_dir, exist_ok=True)
            filename = os.path.join(profile_dir, f"{self._profile_prefix}_Re{int(self.fluid.get_reynolds_number())}"
                                                 f"_mesh{nx}x{ny}_profile.h5")
            print(f"Saved profile to {self.save_profiling_data(filename)}")
        return result


class GpuPisoSolver(GpuSimpleSolver):
    """Twin of ``PisoSolver(BaseAlgorithm)`` (solver/Algorithms/piso.py:9-175): predictor with alpha_u, then
    ``n_corrections`` x (pressure solve, p = p* + alpha_p p' with zero-gradient edges, velocity correction); between
    corrections both momentum equations are solved again from the corrected (u, v, p) with relaxation factor 1
    (:92-104).  The recorded norms are the predictor's and the last correction's (:107-109); ``residual_history``
    gets one entry per iteration (:119).  Same device loop as GpuSimpleSolver (``nf_simple_config.piso_corrections``)."""
    _history_appends_per_iteration = 1
    _profile_prefix = "PISO"

    def __init__(self, mesh, fluid, pressure_solver=None, momentum_solver=None, velocity_updater=None,
                 boundary_conditions=None, alpha_p=0.3, alpha_u=0.7, fix_lid_corners=False, n_corrections=2, **kw):
        if int(n_corrections) < 1:
            # range(0) in piso.py:73 leaves p_res_info unbound -> UnboundLocalError at :108; refuse up front
            raise ValueError("n_corrections must be >= 1")
        self.n_corrections = int(n_corrections)
        self._piso_corrections = self.n_corrections
        super().__init__(mesh, fluid, pressure_solver, momentum_solver, velocity_updater, boundary_conditions,
                         alpha_p=alpha_p, alpha_u=alpha_u, fix_lid_corners=fix_lid_corners, **kw)


class GpuSimplerSolver(GpuSimpleSolver):
    """Twin of ``SimplerSolver(BaseAlgorithm)`` (solver/Algorithms/simpler.py:22-262) as the reference codes it: momentum
    predictor from (u, v, p); the pressure solver's answer from (u*, v*, d) is added to p unrelaxed ("p-bar", :124-128);
    both momentum equations are solved again from the same (u, v) with the new p (:131-154); pressure correction p',
    ``p += alpha_p p'`` (:162-163) and the velocity correction with p' (:165-167).  The recorded momentum norms are the first
    predictor's, ``p_rel_norm = ||p - p_old|| / sqrt(nx ny)`` (:172); ``residual_history`` gets one entry per iteration.
    ``alpha_p`` / ``alpha_u`` are keyword-only like the reference's.  Same device loop (``piso_corrections = -1``)."""
    _history_appends_per_iteration = 1
    _piso_corrections = -1
    _profile_prefix = "SIMPLER"

    def __init__(self, mesh, fluid, pressure_solver=None, momentum_solver=None, velocity_updater=None,
                 boundary_conditions=None, *, alpha_p=0.3, alpha_u=0.7, **kw):
        super().__init__(mesh, fluid, pressure_solver, momentum_solver, velocity_updater, boundary_conditions,
                         alpha_p=alpha_p, alpha_u=alpha_u, **kw)

    def solve(self, *, max_iterations=1000, tolerance=1e-6, save_profile=True, profile_dir="results/profiles",
              track_infinity_norm=False, infinity_norm_interval=10, use_l2_norm=False, chunk=None, gather=True):
        """Keyword-only like ``SimplerSolver.solve`` (simpler.py:78-88)."""
        return super().solve(max_iterations, tolerance, save_profile, profile_dir, track_infinity_norm,
                             infinity_norm_interval, use_l2_norm, chunk, gather)


class GpuSimplecSolver(GpuSimpleSolver):
    """Twin of ``SimplecSolver(BaseAlgorithm)`` (solver/Algorithms/simplec.py:11-283) AS THE REFERENCE CODES IT: momentum
    predictor with alpha_u; ``d_u, d_v`` divided by ``1 - (1 - alpha_u)`` (:126-127); pressure correction with the scaled
    ``d``; 5-point smoothing of ``p'`` with a zero boundary ring (:141-147); ``p = p* + alpha_p p'`` without zero-gradient edge
    copies (:154); velocity correction with the smoothed ``p'`` (:162-166).  Residuals are infinity norms: ``residual_history``
    = max|u - u_old|, |v - v_old| (the stopping test), ``momentum_residual_history`` = max|u* - u|, |v* - v|,
    ``pressure_residual_history`` = max|p - p_old|.  The reference's adaptive ``alpha_p *= 0.95`` (:150-153) compares the
    previous total residual with itself and never fires; it is reproduced by leaving ``alpha_p`` alone.  (The reference's
    loop no longer runs against its own solvers -- it unpacks 2-tuples, :107-117 -- the golden runs use two adapters that only
    reshape return values, oracle/make_golden.py:simplec_runs.)  Same device loop (``piso_corrections = -2``), single slab."""
    _history_appends_per_iteration = 1
    _piso_corrections = -2
    _profile_prefix = "SIMPLEC"

    def __init__(self, mesh, fluid, pressure_solver=None, momentum_solver=None, velocity_updater=None,
                 boundary_conditions=None, alpha_p=0.2, alpha_u=0.7, **kw):
        super().__init__(mesh, fluid, pressure_solver, momentum_solver, velocity_updater, boundary_conditions,
                         alpha_p=alpha_p, alpha_u=alpha_u, **kw)

    def _config(self):
        c = super()._config()
        c.simplec_divisor = 1 - (1 - self.alpha_u)  # the reference's expression, evaluated in the same order
        return c

    def solve(self, max_iterations=1000, tolerance=1e-6, save_profile=True, profile_dir="results/profiles",
              track_infinity_norm=False, infinity_norm_interval=10, use_l2_norm=False, chunk=None, gather=True):
        result = super().solve(max_iterations, tolerance, save_profile, profile_dir, track_infinity_norm,
                               infinity_norm_interval, use_l2_norm, chunk, gather)
        # the device record carries the three infinity norms (include/naviflow_b200.h: piso_corrections == -2)
        self.momentum_residual_history = list(self._u_abs_res_history)
        self.pressure_residual_history = list(self.pressure_rel_norms)
        result.momentum_residuals = self.momentum_residual_history
        result.pressure_residuals = self.pressure_residual_history
        return result
