"""Generate tests/golden/*.npz from the REAL reference (/root/reference) -- TEST INFRASTRUCTURE.

Run in the build container only:  ``python -m oracle.make_golden``.
The fixtures hold seeded inputs *and* the reference's outputs, so the GPU box (which has no
/root/reference) can check both the NumPy oracle and the CUDA path against the reference.
Also dumps the Ghia et al. tables (data, cavity_flow.py:29-124) to naviflow_b200/ghia_tables.json.
"""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

from . import reference_loader as rl

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "..", "tests", "golden")


def _quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def synth_pressure_inputs(n, seed, mu=1e-3):
    """SURVEY.md section 8d C4 synthetic inputs (seeded)."""
    rng = np.random.default_rng(seed)
    dx = dy = 1.0 / (n - 1)
    d_u = (0.7 * dy / (4 * mu)) * (1 + 0.1 * rng.random((n + 1, n)))
    d_v = (0.7 * dx / (4 * mu)) * (1 + 0.1 * rng.random((n, n + 1)))
    d_u[0, :] = np.nan; d_u[n, :] = np.nan      # as JacobiMatrixMomentumSolver leaves them
    d_v[:, 0] = np.nan; d_v[:, n] = np.nan
    us = 1e-2 * rng.standard_normal((n + 1, n)); us[0, :] = us[n, :] = 0
    vs = 1e-2 * rng.standard_normal((n, n + 1)); vs[:, 0] = vs[:, n] = 0
    x = rng.standard_normal((n, n))
    u = 0.1 * rng.standard_normal((n + 1, n))
    v = 0.1 * rng.standard_normal((n, n + 1))
    p = rng.standard_normal((n, n))
    return dict(d_u=d_u, d_v=d_v, u_star=us, v_star=vs, x=x, u=u, v=v, p=p)


def cavity_bc(R):
    bc = R.BoundaryConditionManager()
    bc.set_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        bc.set_condition(b, "wall")
    return bc


def kernel_kats(R, n, seed):
    out = synth_pressure_inputs(n, seed)
    nx = ny = n
    mesh = R.StructuredMesh(nx, ny, 1.0, 1.0)
    dx, dy = mesh.get_cell_sizes()
    fluid = R.FluidProperties(density=1.0, reynolds_number=1000, characteristic_velocity=1.0)
    bc = cavity_bc(R)
    rho = 1.0
    d_u, d_v, us, vs, x, u, v, p = (out[k] for k in ("d_u", "d_v", "u_star", "v_star", "x", "u", "v", "p"))
    # a2
    ub, vb = bc.apply_velocity_boundary_conditions(u.copy(), v.copy(), nx, ny)
    out["bc_u"], out["bc_v"] = ub, vb
    ub1, vb1 = bc.apply_velocity_boundary_conditions(u.copy(), v.copy(), nx + 1, ny)
    out["bc1_u"], out["bc1_v"] = ub1, vb1
    # a3/a4
    pl = R.PowerLawDiscretization()
    for nm, c in (("cu", pl.calculate_u_coefficients(mesh, fluid, ub, vb, p, bc)),
                  ("cv", pl.calculate_v_coefficients(mesh, fluid, ub, vb, p, bc))):
        for k, a in c.items():
            out[f"{nm}_{k}"] = a
    # a6
    ms = R.JacobiMatrixMomentumSolver(n_jacobi_sweeps=4)
    r = ms.solve_u_momentum(mesh, fluid, u, v, p, 0.7, bc)
    out["mom_u_star"], out["mom_d_u"], out["mom_u_norm"], out["mom_u_field"] = r[0], r[1], np.float64(r[2]), r[3]
    r = ms.solve_v_momentum(mesh, fluid, u, v, p, 0.7, bc)
    out["mom_v_star"], out["mom_d_v"], out["mom_v_norm"], out["mom_v_field"] = r[0], r[1], np.float64(r[2]), r[3]
    # a8/a9
    b = R.get_rhs(nx, ny, dx, dy, rho, us, vs).reshape((nx, ny), order="F")
    out["rhs"] = b
    out["Ax"] = R.compute_Ap_product(x.flatten("F"), nx, ny, dx, dy, rho, d_u, d_v).reshape((nx, ny), order="F")
    # a10/a11
    js = R.JacobiSolver(omega=0.8)
    out["jacobi_diag"] = js._get_diagonal_elements(nx, ny, dx, dy, rho, d_u, d_v)
    out["jacobi3"] = js.solve(mesh=mesh, p=x.copy(), b=b.copy(), d_u=d_u, d_v=d_v, rho=rho, num_iterations=3,
                              track_residuals=False, return_dict=False)
    gs = R.GaussSeidelSolver(omega=1.5, method_type="red_black")
    out["rbsor3"] = gs.solve(mesh=mesh, p=x.copy(), b=b.copy(), d_u=d_u, d_v=d_v, rho=rho, num_iterations=3,
                             track_residuals=False, return_dict=False)
    # a12 transfer operators
    H = R.multigrid_helpers
    out["fw"] = H.restrict_full_weighting(x)
    out["inject"] = H.restrict_inject(x)
    out["lin_from_fw"] = H.interpolate_linear(out["fw"], nx)
    out["lin_from_inject"] = H.interpolate_linear(out["inject"], nx)
    if out["fw"].shape[0] >= 4:
        out["cub_from_fw"] = H.interpolate_cubic(out["fw"], nx)
    nc = out["fw"].shape[0]
    out["rc_du_fw"], out["rc_dv_fw"] = H.restrict_coefficients(d_u, d_v, nx, ny, nc, nc, dx, dy)
    nci = out["inject"].shape[0]
    out["rc_du_inject"], out["rc_dv_inject"] = H.restrict_coefficients(d_u, d_v, nx, ny, nci, nci, dx, dy)
    # coarse direct solve (multigrid.py:268-302)
    if n <= 15:
        A = R.get_coeff_mat(nx, ny, dx, dy, rho, d_u, d_v)
        from scipy.sparse.linalg import spsolve
        out["direct"] = spsolve(A, b.flatten("F")).reshape((nx, ny), order="F")
    # a14
    vu = R.StandardVelocityUpdater()
    out["corr_u"], out["corr_v"] = vu.update_velocity(mesh, us, vs, x, d_u, d_v, bc)
    return out


def mg_kats(R, n, seed):
    inp = synth_pressure_inputs(n, seed)
    d_u, d_v, us, vs = inp["d_u"], inp["d_v"], inp["u_star"], inp["v_star"]
    mesh = R.StructuredMesh(n, n, 1.0, 1.0)
    out = dict(d_u=d_u, d_v=d_v, u_star=us, v_star=vs)
    cfgs = {
        "v_lin_fw": dict(cycle_type="v", max_iterations=3, tolerance=1e-14),
        "v_cub_fw": dict(cycle_type="v", max_iterations=2, tolerance=1e-14, interpolation_method="interpolate_cubic"),
        "w_lin_fw": dict(cycle_type="w", max_iterations=2, tolerance=1e-14),
        "fmg_cub_v": dict(cycle_type="fmg", cycle_type_final="v", max_iterations=100, tolerance=1e-3,
                          interpolation_method="interpolate_cubic"),
        "v_tol": dict(cycle_type="v", max_iterations=100, tolerance=1e-3),
    }
    if n % 2 == 1:
        cfgs["v_lin_inject"] = dict(cycle_type="v", max_iterations=2, tolerance=1e-14,
                                    restriction_method="restrict_inject")
    for name, kw in cfgs.items():
        ps = R.MultiGridSolver(smoother=R.GaussSeidelSolver(omega=1.5, method_type="red_black"),
                               pre_smoothing=3, post_smoothing=3, coarsest_grid_size=7, **kw)
        p1, i1 = _quiet(ps.solve, mesh, us, vs, d_u, d_v, None)
        out[name + "_p"] = p1
        out[name + "_relnorm"] = np.float64(i1["rel_norm"])
        out[name + "_ncycles"] = np.int64(len(ps.residual_history))
    ps = R.MultiGridSolver(smoother=R.JacobiSolver(omega=0.8), max_iterations=2, tolerance=1e-14,
                           pre_smoothing=2, post_smoothing=2)
    p1, _ = _quiet(ps.solve, mesh, us, vs, d_u, d_v, None)
    out["v_jacobi_smoother_p"] = p1
    bs = R.MatrixFreeBiCGSTABSolver(tolerance=1e-7, max_iterations=1000)
    p1, i1 = _quiet(bs.solve, mesh, us, vs, d_u, d_v, None)
    out["bicgstab_p"] = p1
    out["bicgstab_relnorm"] = np.float64(i1["rel_norm"])
    out["bicgstab_iters"] = np.int64(len(bs.residual_history))
    return out


def simple_runs(R):
    out = {}

    def run(n, Re, ps, k, N):
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        fluid = R.FluidProperties(density=1.0, reynolds_number=Re, characteristic_velocity=1.0)
        alg = R.SimpleSolver(mesh, fluid, ps, R.JacobiMatrixMomentumAdapter(n_jacobi_sweeps=k),
                             R.StandardVelocityUpdater(), alpha_p=0.3, alpha_u=0.7)
        alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
        for b in ("bottom", "left", "right"):
            alg.set_boundary_condition(b, "wall")
        _quiet(alg.solve, max_iterations=N, tolerance=0.0, save_profile=False, track_infinity_norm=False)
        return alg

    def mk(name):
        GS = R.GaussSeidelSolver
        if name == "fmg":
            return R.MultiGridSolver(smoother=GS(omega=1.5, method_type="red_black"), max_iterations=100,
                                     tolerance=1e-3, pre_smoothing=3, post_smoothing=3, cycle_type="fmg",
                                     cycle_type_buildup="v", cycle_type_final="v", max_cycles_buildup=1,
                                     restriction_method="restrict_full_weighting",
                                     interpolation_method="interpolate_cubic", coarsest_grid_size=7)
        if name == "v":
            return R.MultiGridSolver(smoother=GS(omega=1.5, method_type="red_black"), max_iterations=100,
                                     tolerance=1e-3, pre_smoothing=3, post_smoothing=3)
        if name == "jacobi":
            return R.JacobiSolver(tolerance=0.0, max_iterations=50, omega=0.8)
        if name == "rbsor":
            return GS(tolerance=0.0, max_iterations=30, omega=1.5, method_type="red_black")
        if name == "direct":
            return R.DirectPressureSolver()
        raise ValueError(name)

    for n, Re, k, N, names in ((31, 100, 5, 40, ("fmg", "v", "jacobi", "rbsor", "direct")),
                               (63, 1000, 20, 25, ("fmg", "v")),
                               (64, 1000, 3, 12, ("v",)),
                               (127, 1000, 5, 8, ("v",))):
        for name in names:
            alg = run(n, Re, mk(name), k, N)
            key = f"n{n}_Re{Re}_k{k}_N{N}_{name}"
            out[key + "_u"], out[key + "_v"], out[key + "_p"] = alg.u, alg.v, alg.p
            out[key + "_hist"] = np.array(alg.residual_history[::2])  # appended twice per iteration (simple.py:177,196)
            if n <= 63:
                cf = R.cavity_flow
                out[key + "_ghia"] = np.array([cf.calculate_infinity_norm_error(alg.u, alg.v, alg.mesh, Re),
                                               cf.calculate_l2_norm_error(alg.u, alg.v, alg.mesh, Re)])
    return out


def piso_runs(R):
    """PisoSolver with the deterministic momentum adapter."""
    out = {}
    for n, Re, k, N, nc, name in ((31, 100, 5, 15, 2, "v"), (31, 100, 5, 15, 3, "rbsor"), (63, 1000, 10, 8, 2, "v")):
        GS = R.GaussSeidelSolver
        ps = (R.MultiGridSolver(smoother=GS(omega=1.5, method_type="red_black"), max_iterations=100, tolerance=1e-3,
                                pre_smoothing=3, post_smoothing=3) if name == "v"
              else GS(tolerance=0.0, max_iterations=30, omega=1.5, method_type="red_black"))
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        fluid = R.FluidProperties(density=1.0, reynolds_number=Re, characteristic_velocity=1.0)
        alg = R.PisoSolver(mesh, fluid, ps, R.JacobiMatrixMomentumAdapter(n_jacobi_sweeps=k), R.StandardVelocityUpdater(),
                           alpha_p=0.3, alpha_u=0.7, n_corrections=nc)
        alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
        for b in ("bottom", "left", "right"):
            alg.set_boundary_condition(b, "wall")
        _quiet(alg.solve, max_iterations=N, tolerance=0.0, save_profile=False, track_infinity_norm=False)
        key = f"n{n}_Re{Re}_k{k}_N{N}_c{nc}_{name}"
        out[key + "_u"], out[key + "_v"], out[key + "_p"] = alg.u, alg.v, alg.p
        out[key + "_hist"] = np.array(alg.residual_history)
    return out


def simpler_runs(R):
    """SimplerSolver with the deterministic momentum adapter."""
    out = {}
    for n, Re, k, N, name in ((31, 100, 5, 12, "v"), (31, 100, 5, 12, "rbsor"), (63, 1000, 10, 6, "v")):
        GS = R.GaussSeidelSolver
        ps = (R.MultiGridSolver(smoother=GS(omega=1.5, method_type="red_black"), max_iterations=100, tolerance=1e-3,
                                pre_smoothing=3, post_smoothing=3) if name == "v"
              else GS(tolerance=0.0, max_iterations=30, omega=1.5, method_type="red_black"))
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        fluid = R.FluidProperties(density=1.0, reynolds_number=Re, characteristic_velocity=1.0)
        alg = R.SimplerSolver(mesh, fluid, ps, R.JacobiMatrixMomentumAdapter(n_jacobi_sweeps=k), R.StandardVelocityUpdater(),
                              alpha_p=0.3, alpha_u=0.7)
        alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
        for b in ("bottom", "left", "right"):
            alg.set_boundary_condition(b, "wall")
        res = _quiet(alg.solve, max_iterations=N, tolerance=0.0, save_profile=False, track_infinity_norm=False)
        key = f"n{n}_Re{Re}_k{k}_N{N}_{name}"
        out[key + "_u"], out[key + "_v"], out[key + "_p"] = alg.u, alg.v, alg.p
        out[key + "_hist"] = np.array(res.get_history("total_rel_norm"))
        out[key + "_phist"] = np.array(res.get_history("p_rel_norm"))
    return out


def simplec_runs(R):
    """SimplecSolver.  Its loop is stale against the reference's own solvers: it unpacks 2-tuples from
    the momentum solver (simplec.py:107-117) and indexes the pressure solver's answer as an array (:141-147).  Two adapters
    that only reshape return values make it run; every number is the reference's."""
    from naviflow_oo.solver.Algorithms.simplec import SimplecSolver

    class Momentum2(R.JacobiMatrixMomentumSolver):
        def solve_u_momentum(self, mesh, fluid, u, v, p, relaxation_factor=0.7, boundary_conditions=None):
            return super().solve_u_momentum(mesh, fluid, u, v, p, relaxation_factor, boundary_conditions)[:2]

        def solve_v_momentum(self, mesh, fluid, u, v, p, relaxation_factor=0.7, boundary_conditions=None):
            return super().solve_v_momentum(mesh, fluid, u, v, p, relaxation_factor, boundary_conditions)[:2]

    def array_only(cls):
        class ArrayOnly(cls):
            def solve(self, mesh, u_star, v_star, d_u, d_v, p_star):
                out = super().solve(mesh, u_star, v_star, d_u, d_v, p_star)
                return out[0] if isinstance(out, tuple) else out
        return ArrayOnly

    out = {}
    for n, Re, k, N, name in ((31, 100, 5, 12, "v"), (31, 100, 5, 12, "rbsor"), (63, 1000, 10, 8, "v")):
        GS = R.GaussSeidelSolver
        if name == "v":
            ps = array_only(R.MultiGridSolver)(smoother=GS(omega=1.5, method_type="red_black"), max_iterations=100,
                                               tolerance=1e-3, pre_smoothing=3, post_smoothing=3)
        else:
            ps = array_only(GS)(tolerance=0.0, max_iterations=30, omega=1.5, method_type="red_black")
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        fluid = R.FluidProperties(density=1.0, reynolds_number=Re, characteristic_velocity=1.0)
        alg = SimplecSolver(mesh, fluid, ps, Momentum2(n_jacobi_sweeps=k), R.StandardVelocityUpdater(),
                            alpha_p=0.2, alpha_u=0.7)
        alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
        for b in ("bottom", "left", "right"):
            alg.set_boundary_condition(b, "wall")
        _quiet(alg.solve, max_iterations=N, tolerance=0.0, save_profile=False, track_infinity_norm=False)
        key = f"n{n}_Re{Re}_k{k}_N{N}_{name}"
        out[key + "_u"], out[key + "_v"], out[key + "_p"] = alg.u, alg.v, alg.p
        out[key + "_total"] = np.array(alg.residual_history)
        out[key + "_momentum"] = np.array(alg.momentum_residual_history)
        out[key + "_pressure"] = np.array(alg.pressure_residual_history)
        out[key + "_alpha_p"] = np.array([alg.alpha_p])
    return out


def mf_momentum_kats(R):
    """MatrixFreeMomentumSolver (a7) on seeded fields + whole SimpleSolver runs with it (direct pressure solve)."""
    out = {}
    for n, Re, seed in ((15, 100, 715), (32, 1000, 732)):
        rng = np.random.default_rng(seed)
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        fluid = R.FluidProperties(density=1.0, reynolds_number=Re, characteristic_velocity=1.0)
        bc = R.BoundaryConditionManager()
        bc.set_condition("top", "velocity", {"u": 1.0, "v": 0.0})
        for b in ("bottom", "left", "right"):
            bc.set_condition(b, "wall")
        u = 0.3 * rng.standard_normal((n + 1, n))
        v = 0.3 * rng.standard_normal((n, n + 1))
        p = rng.standard_normal((n, n))
        ms = R.MatrixFreeMomentumSolver(tolerance=1e-8, max_iterations=200, solver_type="bicgstab")
        us, du, iu = ms.solve_u_momentum(mesh, fluid, u.copy(), v.copy(), p.copy(), relaxation_factor=0.7,
                                         boundary_conditions=bc)
        vs, dv, iv = ms.solve_v_momentum(mesh, fluid, u.copy(), v.copy(), p.copy(), relaxation_factor=0.7,
                                         boundary_conditions=bc)
        k = f"kat_n{n}_Re{Re}"
        out.update({k + "_u": u, k + "_v": v, k + "_p": p, k + "_us": us, k + "_du": du, k + "_vs": vs, k + "_dv": dv,
                    k + "_unorm": iu["rel_norm"], k + "_vnorm": iv["rel_norm"], k + "_ufield": iu["field"],
                    k + "_vfield": iv["field"], k + "_iters": np.array([iu["iterations"], iv["iterations"]])})
    for n, Re, N in ((31, 100, 30), (63, 1000, 20)):
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        fluid = R.FluidProperties(density=1.0, reynolds_number=Re, characteristic_velocity=1.0)
        alg = R.SimpleSolver(mesh, fluid, R.DirectPressureSolver(),
                             R.MatrixFreeMomentumSolver(tolerance=1e-8, max_iterations=200, solver_type="bicgstab"),
                             R.StandardVelocityUpdater(), alpha_p=0.3, alpha_u=0.7)
        alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
        for b in ("bottom", "left", "right"):
            alg.set_boundary_condition(b, "wall")
        res = _quiet(alg.solve, max_iterations=N, tolerance=0.0, save_profile=False, track_infinity_norm=False)
        k = f"run_n{n}_Re{Re}_N{N}"
        out[k + "_u"], out[k + "_v"], out[k + "_p"] = alg.u, alg.v, alg.p
        out[k + "_hist"] = np.array(res.get_history("total_rel_norm"))[::2]
    return out


def mg_lex_kats(R):
    """MultiGridSolver with the sequential Gauss-Seidel smoothers (the configuration of the reference's README table)."""
    out = {}
    for n, seed in ((31, 1131), (40, 1140)):
        inp = synth_pressure_inputs(n, seed)
        d_u, d_v, us, vs = inp["d_u"], inp["d_v"], inp["u_star"], inp["v_star"]
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        out.update({f"n{n}_d_u": d_u, f"n{n}_d_v": d_v, f"n{n}_u_star": us, f"n{n}_v_star": vs})
        for mt, kw in (("standard", dict(pre_smoothing=2, post_smoothing=2, max_iterations=2, tolerance=1e-14)),
                       ("symmetric", dict(pre_smoothing=1, post_smoothing=1, max_iterations=100, tolerance=1e-4))):
            ps = R.MultiGridSolver(smoother=R.GaussSeidelSolver(omega=1.2, method_type=mt), **kw)
            p1, i1 = _quiet(ps.solve, mesh, us, vs, d_u, d_v, None)
            out[f"n{n}_{mt}_p"] = p1
            out[f"n{n}_{mt}_relnorm"] = np.float64(i1["rel_norm"])
            out[f"n{n}_{mt}_ncycles"] = np.int64(len(ps.residual_history))
    return out


RECT_CASES = ((40, 24, 100, 5, 10, "rbsor"), (24, 40, 400, 3, 8, "jacobi"), (33, 70, 100, 4, 6, "rbsor"))


def cg_mg_kats(R):
    """GeoMultigridPrecondCGSolver (geo_multigrid_cg.py) on seeded systems.  Its defaults do not construct
    (mg_cycle_type='f' is rejected by MultiGridSolver, and the V-cycle needs a smoother object), so the runs name a cycle
    type and pass the red-black smoother."""
    from naviflow_oo.solver.pressure_solver.geo_multigrid_cg import GeoMultigridPrecondCGSolver
    out = {}
    for n, kind, cycles, seed in ((31, "v", 1, 931), (64, "v", 2, 964), (65, "w", 1, 965)):
        s = synth_pressure_inputs(n, seed)
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        sol = GeoMultigridPrecondCGSolver(tolerance=1e-7, max_iterations=200, mg_pre_smoothing=2, mg_post_smoothing=2,
                                          mg_cycles=cycles, mg_cycle_type=kind, mg_restriction_method="restrict_full_weighting",
                                          mg_interpolation_method="interpolate_linear",
                                          smoother=R.GaussSeidelSolver(omega=0.8, method_type="red_black"))
        p = _quiet(sol.solve, mesh, s["u_star"], s["v_star"], s["d_u"], s["d_v"], None)
        key = f"n{n}_{kind}{cycles}"
        for f in ("u_star", "v_star", "d_u", "d_v"):
            out[f"{key}_{f}"] = s[f]
        out[key + "_p"] = p
        out[key + "_iterations"] = np.array([sol.inner_iterations[-1]])
    return out


def rect_runs(R):
    """SimpleSolver on rectangular cell grids (nx != ny, unit square: dx != dy) with the stationary pressure solvers."""
    out = {}
    for nx, ny, Re, k, N, name in RECT_CASES:
        ps = (R.GaussSeidelSolver(tolerance=0.0, max_iterations=30, omega=1.5, method_type="red_black") if name == "rbsor"
              else R.JacobiSolver(tolerance=0.0, max_iterations=50, omega=0.8))
        mesh = R.StructuredMesh(nx, ny, 1.0, 1.0)
        fluid = R.FluidProperties(density=1.0, reynolds_number=Re, characteristic_velocity=1.0)
        alg = R.SimpleSolver(mesh, fluid, ps, R.JacobiMatrixMomentumAdapter(n_jacobi_sweeps=k), R.StandardVelocityUpdater(),
                             alpha_p=0.3, alpha_u=0.7)
        alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
        for b in ("bottom", "left", "right"):
            alg.set_boundary_condition(b, "wall")
        res = _quiet(alg.solve, max_iterations=N, tolerance=0.0, save_profile=False, track_infinity_norm=False)
        key = f"nx{nx}_ny{ny}_Re{Re}_k{k}_N{N}_{name}"
        out[key + "_u"], out[key + "_v"], out[key + "_p"] = alg.u, alg.v, alg.p
        out[key + "_hist"] = np.array(res.get_history("total_rel_norm"))[::2]
    return out


def bicgstab_mg_kats(R):
    """MatrixFreeBiCGSTABSolver with the multigrid preconditioner."""
    out = {}
    for n, kind, cycles in ((31, "v", 1), (64, "v", 2), (65, "w", 1), (63, "fmg", 1)):
        s = synth_pressure_inputs(n, 6000 + n)
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        bs = R.MatrixFreeBiCGSTABSolver(tolerance=1e-7, max_iterations=200, use_preconditioner=True,
                                        preconditioner="multigrid", mg_cycles=cycles, mg_cycle_type=kind)
        p, info = _quiet(bs.solve, mesh, s["u_star"], s["v_star"], s["d_u"], s["d_v"], None)
        k = f"n{n}_{kind}{cycles}"
        out.update({k + "_d_u": s["d_u"], k + "_d_v": s["d_v"], k + "_u_star": s["u_star"], k + "_v_star": s["v_star"],
                    k + "_p": p, k + "_relnorm": np.float64(info["rel_norm"])})
    return out


def gs_lex_kats(R):
    """GaussSeidelSolver(method_type='standard' | 'symmetric'): 3 sweeps on seeded systems."""
    out = {}
    for n, seed in ((15, 915), (33, 933), (40, 940)):
        rng = np.random.default_rng(seed)
        mesh = R.StructuredMesh(n, n, 1.0, 1.0)
        dx, dy = mesh.get_cell_sizes()
        d_u = (0.7 * dy / 4e-3) * (1 + 0.1 * rng.random((n + 1, n)))
        d_v = (0.7 * dx / 4e-3) * (1 + 0.1 * rng.random((n, n + 1)))
        b = 1e-2 * rng.standard_normal((n, n))
        b[0, 0] = 0.0
        p0 = 1e-3 * rng.standard_normal((n, n))
        out[f"n{n}_du"], out[f"n{n}_dv"], out[f"n{n}_b"], out[f"n{n}_p0"] = d_u, d_v, b, p0
        for mt in ("standard", "symmetric"):
            gs = R.GaussSeidelSolver(omega=1.5, method_type=mt)
            out[f"n{n}_{mt}"] = gs.solve(mesh=mesh, p=p0.copy(), b=b.copy(), d_u=d_u, d_v=d_v, rho=1.0, num_iterations=3,
                                         track_residuals=False, return_dict=False)
    return out


def ext_links_kats(R):
    """QUICKDiscretization / SecondOrderUpwindDiscretization outputs (quick.py, second_order_upwind.py) on seeded fields:
    square and rectangular grids, with the cavity's boundary conditions (Practice B on all four sides) and with bc=None."""
    out = {}
    cases = []
    for nx, ny, seed in ((9, 9, 901), (16, 16, 902), (31, 31, 903), (12, 20, 904), (5, 4, 905)):
        rng = np.random.default_rng(seed)
        mesh = R.StructuredMesh(nx, ny, 1.0, 0.7 if nx != ny else 1.0)
        fluid = R.FluidProperties(density=1.3, reynolds_number=400, characteristic_velocity=1.0)
        u = 0.5 * rng.standard_normal((nx + 1, ny))
        v = 0.5 * rng.standard_normal((nx, ny + 1))
        p = rng.standard_normal((nx, ny))
        bc = cavity_bc(R)
        ub, vb = bc.apply_velocity_boundary_conditions(u.copy(), v.copy(), nx, ny)
        tag = f"{nx}x{ny}"
        cases.append(tag)
        out[f"{tag}_u"], out[f"{tag}_v"], out[f"{tag}_p"] = ub, vb, p
        out[f"{tag}_dims"] = np.array([nx, ny], dtype=np.int64)
        dx, dy = mesh.get_cell_sizes()
        out[f"{tag}_scal"] = np.array([dx, dy, fluid.get_density(), fluid.get_viscosity()])
        for sname, cls in (("quick", R.QUICKDiscretization), ("sou", R.SecondOrderUpwindDiscretization)):
            for bname, b in (("bc", bc), ("nobc", None)):
                d = cls()
                for comp, c in (("u", d.calculate_u_coefficients(mesh, fluid, ub, vb, p, b)),
                                ("v", d.calculate_v_coefficients(mesh, fluid, ub, vb, p, b))):
                    for k, a in c.items():
                        out[f"{tag}_{sname}_{bname}_{comp}_{k}"] = a
    out["cases"] = np.array(cases)
    return out


def main():
    warnings.filterwarnings("ignore")
    R = rl.ref()
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1:  # regenerate single fixtures: make_golden.py simplec_runs ...
        for name in sys.argv[1:]:
            np.savez_compressed(os.path.join(GOLD, name + ".npz"), **globals()[name](R))
            print("written", name)
        return
    for n, seed in ((8, 108), (15, 115), (31, 131), (32, 132)):
        np.savez_compressed(os.path.join(GOLD, f"kernels_n{n}.npz"), **kernel_kats(R, n, seed))
    for n, seed in ((31, 231), (33, 233), (64, 264)):
        np.savez_compressed(os.path.join(GOLD, f"mg_n{n}.npz"), **mg_kats(R, n, seed))
    np.savez_compressed(os.path.join(GOLD, "simple_runs.npz"), **simple_runs(R))
    np.savez_compressed(os.path.join(GOLD, "piso_runs.npz"), **piso_runs(R))
    np.savez_compressed(os.path.join(GOLD, "simpler_runs.npz"), **simpler_runs(R))
    np.savez_compressed(os.path.join(GOLD, "simplec_runs.npz"), **simplec_runs(R))
    np.savez_compressed(os.path.join(GOLD, "mf_momentum.npz"), **mf_momentum_kats(R))
    np.savez_compressed(os.path.join(GOLD, "gs_lex.npz"), **gs_lex_kats(R))
    np.savez_compressed(os.path.join(GOLD, "mg_lex.npz"), **mg_lex_kats(R))
    np.savez_compressed(os.path.join(GOLD, "bicgstab_mg.npz"), **bicgstab_mg_kats(R))
    np.savez_compressed(os.path.join(GOLD, "rect_runs.npz"), **rect_runs(R))
    np.savez_compressed(os.path.join(GOLD, "cg_mg_kats.npz"), **cg_mg_kats(R))
    np.savez_compressed(os.path.join(GOLD, "ext_links_kats.npz"), **ext_links_kats(R))
    cf = R.cavity_flow.BenchmarkData
    tables = {}
    for Re in (100, 400, 1000, 3200, 5000, 7500, 10000):
        t = cf.get_ghia_data(Re)
        tables[str(Re)] = {k: [float(x) for x in t[k]] for k in ("x", "v", "y", "u")}
    with open(os.path.join(HERE, "..", "naviflow_b200", "ghia_tables.json"), "w") as f:
        json.dump(tables, f, indent=1)
    print("golden fixtures written to", os.path.normpath(GOLD))


if __name__ == "__main__":
    main()
