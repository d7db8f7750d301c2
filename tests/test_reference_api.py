"""The Gpu* plugin classes against the REAL reference classes (build container only: needs /root/reference).

Drop-in rule: every parameter of the reference method exists in the twin with the same name, position,
kind and default; the twin may append optional keywords (``device``, ``chunk`` ...) after them.  Checked with
``inspect.signature`` for the constructors and the plugin entry points the outer loops call
(multigrid.py:31-37,121; gauss_seidel.py:21,55; jacobi.py:80; base_momentum_solver.py:144-204; standard.py:10;
base_algorithm.py:24-66; simple.py:78; piso.py; simpler.py:78; simplec.py:19,47)."""
import inspect

import pytest

from oracle import reference_loader as RL

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not RL.available(), reason="reference tree not present")]


def _compatible(gpu_fn, ref_fn, allow_looser_defaults=()):
    sg, sr = inspect.signature(gpu_fn), inspect.signature(ref_fn)
    gp, rp = list(sg.parameters.values()), list(sr.parameters.values())
    problems = []
    for k, r in enumerate(rp):
        if k >= len(gp):
            problems.append(f"missing parameter {r.name}")
            continue
        g = gp[k]
        if g.name != r.name:
            problems.append(f"position {k}: {g.name} != {r.name}")
        if g.kind != r.kind:
            problems.append(f"{r.name}: kind {g.kind} != {r.kind}")
        if g.default != r.default and r.name not in allow_looser_defaults:
            problems.append(f"{r.name}: default {g.default!r} != {r.default!r}")
    for g in gp[len(rp):]:
        if g.default is inspect.Parameter.empty and g.kind not in (g.VAR_KEYWORD, g.VAR_POSITIONAL):
            problems.append(f"extra parameter {g.name} has no default")
    return problems


def _pairs():
    import naviflow_b200 as nb
    R = RL.ref()
    from naviflow_oo.solver.Algorithms.simplec import SimplecSolver
    return [
        (nb.GpuMultiGridSolver, R.MultiGridSolver, ["__init__", "solve", "get_solver_info"]),
        (nb.GpuGaussSeidelSolver, R.GaussSeidelSolver, ["__init__", "solve", "get_solver_info"]),
        (nb.GpuJacobiSolver, R.JacobiSolver, ["__init__", "solve", "get_solver_info"]),
        (nb.GpuBiCGSTABSolver, R.MatrixFreeBiCGSTABSolver, ["__init__", "solve", "get_solver_info"]),
        (nb.GpuGeoMultigridPrecondCGSolver,
         __import__("naviflow_oo.solver.pressure_solver.geo_multigrid_cg",
                    fromlist=["GeoMultigridPrecondCGSolver"]).GeoMultigridPrecondCGSolver,
         ["__init__", "solve", "get_solver_info"]),
        (nb.GpuJacobiMomentumSolver, R.JacobiMatrixMomentumSolver, ["__init__", "solve_u_momentum", "solve_v_momentum"]),
        (nb.GpuMatrixFreeMomentumSolver, R.MatrixFreeMomentumSolver, ["__init__", "solve_u_momentum", "solve_v_momentum"]),
        (nb.GpuVelocityUpdater, R.StandardVelocityUpdater, ["__init__", "update_velocity"]),
        (nb.GpuSimpleSolver, R.SimpleSolver, ["__init__", "solve", "set_boundary_condition", "initialize_fields",
                                              "apply_boundary_conditions", "get_max_divergence", "save_profiling_data"]),
        (nb.GpuPisoSolver, R.PisoSolver, ["__init__", "solve"]),
        (nb.GpuSimplerSolver, R.SimplerSolver, ["__init__", "solve"]),
        (nb.GpuSimplecSolver, SimplecSolver, ["__init__", "solve"]),
        (nb.GpuQUICKDiscretization, R.QUICKDiscretization, ["calculate_u_coefficients", "calculate_v_coefficients"]),
        (nb.GpuSecondOrderUpwindDiscretization, R.SecondOrderUpwindDiscretization,
         ["calculate_u_coefficients", "calculate_v_coefficients"]),
        (nb.Profiler, __import__("naviflow_oo.utils.profiler", fromlist=["Profiler"]).Profiler,
         ["__init__", "start", "end", "start_section", "end_section", "set_iterations", "set_convergence_info",
          "add_residual_data", "set_pressure_solver_info", "save"]),
    ]


def test_plugin_signatures_match_the_reference_classes():
    bad = []
    for gpu, ref, names in _pairs():
        for n in names:
            assert hasattr(gpu, n), f"{gpu.__name__} lacks {n}"
            for p in _compatible(getattr(gpu, n), getattr(ref, n)):
                bad.append(f"{gpu.__name__}.{n}: {p}")
    assert not bad, "\n".join(bad)


def test_algorithm_objects_expose_the_base_algorithm_attributes():
    """base_algorithm.py:24-66: .u .v .p .mesh .fluid .bc_manager .profiler .pressure_solver .momentum_solver
    .velocity_updater, and the profiler is named after the class like the reference's."""
    import naviflow_b200 as nb
    mesh = nb.StructuredMesh(9, 9, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=100, characteristic_velocity=1.0)
    import ctypes  # noqa: F401  (constructing the algorithm object needs no GPU)
    alg = nb.GpuSimpleSolver(mesh, fluid, None, None, None)
    for a in ("u", "v", "p", "mesh", "fluid", "bc_manager", "profiler", "pressure_solver", "momentum_solver",
              "velocity_updater", "alpha_p", "alpha_u"):
        assert hasattr(alg, a), a
    assert alg.profiler.algorithm_name == "GpuSimpleSolver"
    assert alg.u.shape == (10, 9) and alg.v.shape == (9, 10) and alg.p.shape == (9, 9)


def test_virtual_subclass_registration():
    import naviflow_b200 as nb
    RL.load()
    pairs = nb.register_with_reference()
    assert len(pairs) == 13
    from naviflow_oo.solver.pressure_solver.base_pressure_solver import PressureSolver
    from naviflow_oo.solver.Algorithms.base_algorithm import BaseAlgorithm
    assert issubclass(nb.GpuMultiGridSolver, PressureSolver) and issubclass(nb.GpuSimpleSolver, BaseAlgorithm)


def test_host_objects_drop_into_the_reference_simple_solver(golden_dir):
    """INTEGRATION.md path A from the other side: the reference's own SimpleSolver loop (simple.py:78-269) runs with THIS
    package's StructuredMesh, FluidProperties and BoundaryConditionManager in place of its own (and the reference's CPU
    solvers) and reproduces the golden run bit for bit -- the host objects are drop-in, method for method."""
    import contextlib
    import io
    import os
    import numpy as np
    import naviflow_b200 as nb
    R = RL.ref()
    n, Re, k, N = 31, 100, 5, 40
    mesh = nb.StructuredMesh(n, n, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=Re, characteristic_velocity=1.0)
    bc = nb.BoundaryConditionManager()
    ps = R.GaussSeidelSolver(tolerance=0.0, max_iterations=30, omega=1.5, method_type="red_black")
    alg = R.SimpleSolver(mesh, fluid, ps, R.JacobiMatrixMomentumAdapter(n_jacobi_sweeps=k), R.StandardVelocityUpdater(),
                         alpha_p=0.3, alpha_u=0.7)
    # base_algorithm.py:51-54 accepts only instances of the reference's own (non-abstract) manager class in the constructor,
    # so the foreign manager is attached afterwards; every later use goes through self.bc_manager
    alg.bc_manager = bc
    alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        alg.set_boundary_condition(b, "wall")
    assert alg.bc_manager is bc and set(bc.conditions) == {"top", "bottom", "left", "right"}
    with contextlib.redirect_stdout(io.StringIO()):
        alg.solve(max_iterations=N, tolerance=0.0, save_profile=False, track_infinity_norm=False)
    G = np.load(os.path.join(golden_dir, "simple_runs.npz"))
    key = f"n{n}_Re{Re}_k{k}_N{N}_rbsor"
    for f in ("u", "v", "p"):
        np.testing.assert_array_equal(getattr(alg, f), G[f"{key}_{f}"])
