"""CPU tests: the C-ABI library loads and exports every symbol include/naviflow_b200.h declares (no compute
calls without a GPU), and the host-side mirror of the reference's objects behaves like the reference
(checked against the oracle and the reference's golden vectors)."""
import os
import re

import numpy as np
import pytest

from oracle import np_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "naviflow_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nf_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_header_symbol():
    from naviflow_b200 import _lib
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes table covers the header one to one
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.nf_version() >= 100


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """sizeof of every POD of include/naviflow_b200.h as gcc lays it out == the ctypes mirror in naviflow_b200/_lib.py."""
    import ctypes as C
    import subprocess
    from naviflow_b200 import _lib
    pairs = [("nf_grid", _lib.NfGrid), ("nf_bc_program", _lib.NfBcProgram), ("nf_mg_config", _lib.NfMgConfig),
             ("nf_simple_config", _lib.NfSimpleConfig), ("nf_simple_info", _lib.NfSimpleInfo),
             ("nf_krylov_info", _lib.NfKrylovInfo), ("nf_links", _lib.NfLinks), ("nf_links_ext", _lib.NfLinksExt),
             ("nf_mg_info", _lib.NfMgInfo)]
    header = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "naviflow_b200.h")
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(void){printf("%s\\n", %s);return 0;}\n'
                   % (header, " ".join(["%zu"] * len(pairs)), ", ".join(f"sizeof({c})" for c, _ in pairs)))
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(t) for _, t in pairs]
    assert C.sizeof(_lib.NfGrid) == 8 * 4 + 3 * 8


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import naviflow_b200 as nb
    ps = nb.GpuJacobiSolver()
    with pytest.raises(RuntimeError, match="CUDA device"):
        ps.ctx


def cavity_bc():
    import naviflow_b200 as nb
    bc = nb.BoundaryConditionManager()
    bc.set_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        bc.set_condition(b, "wall")
    return bc


@pytest.mark.parametrize("n", [8, 15, 31, 32])
def test_boundary_manager_matches_reference_golden(golden_dir, n):
    g = dict(np.load(os.path.join(golden_dir, f"kernels_n{n}.npz")))
    bc = cavity_bc()
    u, v = bc.apply_velocity_boundary_conditions(g["u"].copy(), g["v"].copy(), n, n)
    np.testing.assert_array_equal(u, g["bc_u"]); np.testing.assert_array_equal(v, g["bc_v"])
    u, v = bc.apply_velocity_boundary_conditions(g["u"].copy(), g["v"].copy(), n + 1, n)
    np.testing.assert_array_equal(u, g["bc1_u"]); np.testing.assert_array_equal(v, g["bc1_v"])
    assert bc.get_boundary_types() == O.boundary_types(O.bc_conditions())
    assert list(bc.get_boundary_types()) == ["top", "bottom", "left", "right"]


@pytest.mark.parametrize("order", [("top", "bottom", "left", "right"), ("left", "right", "top", "bottom"),
                                   ("bottom", "top")])
def test_boundary_program_is_the_routine_evaluated_symbolically(order):
    """nf_bc_program (edge / corner constants) reproduces the insertion-order dependent corner values."""
    import naviflow_b200 as nb
    from naviflow_b200.host import boundary_program, practice_b_sides
    bc = nb.BoundaryConditionManager()
    entries = []
    for loc in order:
        if loc == "top":
            bc.set_condition("top", "velocity", {"u": 1.0, "v": 0.25}); entries.append(("top", "velocity", {"u": 1.0, "v": 0.25}))
        else:
            bc.set_condition(loc, "wall"); entries.append((loc, "wall", None))
    cond = O.bc_conditions(entries)
    for nx_arg_off in (0, 1):
        n = 11
        u = np.full((n + 1, n), 7.0); v = np.full((n, n + 1), 7.0)
        O.apply_velocity_bc(u, v, n + nx_arg_off, n, cond)
        prog = boundary_program(bc, n, n, n + nx_arg_off)
        got_u = [u[0, 4], u[n, 4], u[4, 0], u[4, n - 1]]
        assert prog["u_edge"] == got_u
        assert prog["u_corner"] == [u[0, 0], u[0, n - 1], u[n, 0], u[n, n - 1]]
        ve = [v[0, 4], v[n - 1, 4], v[4, 0], v[4, n]]
        for a, b in zip(prog["v_edge"], ve):
            assert (np.isnan(a) and b == 7.0) or a == b
        assert prog["v_corner"] == [v[0, 0], v[0, n], v[n - 1, 0], v[n - 1, n]]
        assert prog["v_right_row"] == (n - 1 if nx_arg_off == 0 else -1)
    mask = practice_b_sides(bc)
    assert mask == sum(bit for bit, loc in ((1, "left"), (2, "right"), (4, "bottom"), (8, "top")) if loc in order)


def test_mesh_and_fluid_quirks():
    import naviflow_b200 as nb
    m = nb.StructuredMesh(63, 63, 1.0, 1.0)
    assert m.get_cell_sizes() == (1.0 / 62, 1.0 / 62)          # dx = L/(nx-1): structured.py:27-28
    assert m.get_dimensions() == (63, 63)
    f = nb.FluidProperties(density=1.0, reynolds_number=1000, characteristic_velocity=1.0)
    assert f.get_viscosity() == 1e-3 and f.get_reynolds_number() == 1000
    with pytest.raises(ValueError):
        nb.FluidProperties()
    with pytest.raises(ValueError):
        nb.BoundaryConditionManager().set_condition("north", "wall")


def test_constructor_validation_matches_reference():
    import naviflow_b200 as nb
    gs = nb.GpuGaussSeidelSolver(omega=1.5)
    with pytest.raises(ValueError):
        nb.GpuMultiGridSolver(gs, coarsest_grid_size=2)
    with pytest.raises(ValueError):
        nb.GpuMultiGridSolver(gs, coarsest_grid_size=8)
    with pytest.raises(ValueError):
        nb.GpuMultiGridSolver(gs, restriction_method="nope")
    with pytest.raises(ValueError):
        nb.GpuMultiGridSolver(gs, interpolation_method="nope")
    with pytest.raises(ValueError):
        nb.GpuGaussSeidelSolver(method_type="zebra")
    cfg = nb.GpuMultiGridSolver(gs, pre_smoothing=3, post_smoothing=3, cycle_type="fmg", cycle_type_final="v",
                                interpolation_method="interpolate_cubic").config_struct()
    assert (cfg.smoother, cfg.pre, cfg.post, cfg.cycle_type, cfg.cycle_final, cfg.interpolation, cfg.omega) == \
           (0, 3, 3, 2, 0, 1, 1.5)


def test_ghia_errors_match_reference_golden(golden_dir):
    import naviflow_b200 as nb
    g = dict(np.load(os.path.join(golden_dir, "simple_runs.npz")))
    key = "n63_Re1000_k20_N25_fmg"
    mesh = nb.StructuredMesh(63, 63)
    inf, l2 = nb.ghia_errors(g[key + "_u"], g[key + "_v"], mesh, 1000)
    np.testing.assert_allclose([inf, l2], g[key + "_ghia"], rtol=1e-12)


def test_constructor_argument_errors_match_the_reference_classes():
    """Argument validation happens on the host before any device work (gauss_seidel.py:38-39, matrix_free_momentum.py:31-33,
    piso.py:73 -- n_corrections = 0 would leave p_res_info unbound there)."""
    import naviflow_b200 as nb
    with pytest.raises(ValueError):
        nb.GpuGaussSeidelSolver(method_type="zebra")
    for mt in ("red_black", "standard", "symmetric"):
        assert nb.GpuGaussSeidelSolver(method_type=mt).method_type == mt
    with pytest.raises(ValueError):
        nb.GpuMatrixFreeMomentumSolver(solver_type="cholesky")
    with pytest.raises(NotImplementedError):
        nb.GpuMatrixFreeMomentumSolver(solver_type="idrs")
    with pytest.raises(ValueError):
        nb.GpuMatrixFreeMomentumSolver(discretization_scheme="central")
    mesh = nb.StructuredMesh(17, 17, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=100, characteristic_velocity=1.0)
    with pytest.raises(ValueError):
        nb.GpuPisoSolver(mesh, fluid, nb.GpuJacobiSolver(), n_corrections=0)
    piso = nb.GpuPisoSolver(mesh, fluid, nb.GpuJacobiSolver(tolerance=0.0, max_iterations=5), n_corrections=3)
    assert piso.n_corrections == 3 and piso.u.shape == (18, 17) and piso.v.shape == (17, 18)


def test_simple_config_struct_carries_the_plugin_settings():
    """GpuSimpleSolver._config(): the nf_simple_config the C-ABI receives (no device needed)."""
    import naviflow_b200 as nb
    mesh = nb.StructuredMesh(33, 33, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=400, characteristic_velocity=1.0)
    alg = nb.GpuPisoSolver(mesh, fluid, nb.GpuGaussSeidelSolver(tolerance=0.0, max_iterations=7, omega=1.8, method_type="symmetric"),
                           nb.GpuMatrixFreeMomentumSolver(tolerance=1e-9, max_iterations=77), n_corrections=2, alpha_p=0.2,
                           alpha_u=0.6)
    alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        alg.set_boundary_condition(b, "wall")
    c = alg._config()
    assert (c.nx, c.ny, c.pressure_solver, c.pressure_iterations, c.piso_corrections) == (33, 33, 6, 7, 2)
    assert (c.momentum_solver, c.momentum_maxiter) == (1, 77) and c.momentum_tolerance == 1e-9
    assert c.pressure_omega == 1.8 and c.alpha_p == 0.2 and c.alpha_u == 0.6 and c.sides == 15
    assert abs(c.mu - 1.0 / 400) < 1e-18
    # the Krylov momentum twin applies the BCs with the reference's nx+1 call: the v "right" edge is then skipped
    assert c.bc.v_right_row == 32 and c.bc_mf.v_right_row == -1
    simple = nb.GpuSimpleSolver(mesh, fluid, nb.GpuJacobiSolver(tolerance=0.0, max_iterations=3),
                                nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=4))
    c = simple._config()
    assert (c.pressure_solver, c.piso_corrections, c.momentum_solver, c.n_momentum_sweeps) == (1, 0, 0, 4)
    with pytest.raises(NotImplementedError):
        nb.GpuSimpleSolver(mesh, fluid, nb.GpuJacobiSolver(tolerance=1e-3))._config()


def test_profiler_record_layout_round_trip(tmp_path):
    """naviflow_b200.Profiler keeps the reference's on-disk tree (utils/profiler.py:317-443): groups simulation/mesh_size,
    performance, convergence, system, algorithm, pressure_solver/{smoother,multigrid}, momentum_solver as attributes and
    residual_history/<column> datasets; without h5py the same tree goes to an .npz that load_profile reads back."""
    import naviflow_b200 as nb

    class PS:
        tolerance, max_iterations, cycle_type, pre_smoothing, post_smoothing, smoother_omega = 1e-3, 100, "v", 3, 3, 1.5
        smoother = object()

    class Alg:
        alpha_p, alpha_u = 0.3, 0.7
        pressure_solver = PS()
        momentum_solver = object()

    mesh = nb.StructuredMesh(17, 17, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=100, characteristic_velocity=1.0)
    prof = nb.Profiler("SimpleSolver", mesh, fluid, algorithm=Alg())
    prof.start()
    for it in range(1, 4):
        prof.add_residual_data(it, 1.0 / it, 2.0 / it, 3.0 / it, None if it < 3 else 0.05)
    prof.set_iterations(3)
    prof.set_convergence_info(1e-6, 1.0 / 3, [1.0, 0.5, 1.0 / 3])
    prof.set_pressure_solver_info("GpuMultiGridSolver", inner_iterations=[5, 6, 7])
    prof.end()
    path = prof.save(str(tmp_path / "SIMPLE_Re100_mesh17x17_profile.h5"))
    tree = nb.load_profile(path)
    assert set(tree) >= {"simulation", "performance", "convergence", "system", "algorithm", "pressure_solver",
                         "momentum_solver", "residual_history"}
    assert int(tree["simulation"]["mesh_size"]["x"]) == 17 and float(tree["simulation"]["reynolds_number"]) == 100.0
    assert int(tree["performance"]["iterations"]) == 3 and not bool(tree["convergence"]["converged"])
    assert float(tree["algorithm"]["alpha_p"]) == 0.3
    assert str(tree["pressure_solver"]["type"]) == "PS" and int(tree["pressure_solver"]["multigrid"]["pre_smoothing"]) == 3
    assert float(tree["pressure_solver"]["smoother"]["omega"]) == 1.5
    rh = tree["residual_history"]
    assert set(rh) == {"iteration", "wall_time", "cpu_time", "total_residual", "momentum_residual", "pressure_residual",
                       "infinity_norm_error"}
    np.testing.assert_allclose(rh["total_residual"], [1.0, 0.5, 1.0 / 3])
    assert np.isnan(rh["infinity_norm_error"][0]) and rh["infinity_norm_error"][2] == 0.05
    info = prof.profiling_data["pressure_solver_info"]
    assert info["total_inner_iterations"] == 18 and info["max_inner_iterations"] == 7 and info["min_inner_iterations"] == 5
