"""GPU parity, tier T0: every kernel of libnaviflow_b200 (called through the C-ABI) against
 (a) the reference's own outputs stored in tests/golden/kernels_n*.npz and
 (b) the NumPy oracle on fresh seeded inputs, including odd / even / non-trivial sizes.
Elementwise kernels must be bit-exact; the power-law links are within 1e-14 (pow rounding)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

COND = O.bc_conditions()
MU = 1e-3


def load(golden_dir, n):
    return dict(np.load(os.path.join(golden_dir, f"kernels_n{n}.npz")))


def cavity_bc():
    import naviflow_b200 as nb
    bc = nb.BoundaryConditionManager()
    bc.set_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        bc.set_condition(b, "wall")
    return bc


def cubic_workspace(d, gc):
    """caller-provided device scratch of nf_prolong_cubic (include/naviflow_b200.h: nf_workspace_bytes)"""
    import torch
    nbytes = d.lib.nf_workspace_bytes(1, d.nx, d.ny, gc.nx, gc.ld)
    assert nbytes > 0
    return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device="cuda")


def synth(n, seed):
    from oracle.make_golden import synth_pressure_inputs
    return synth_pressure_inputs(n, seed)


@pytest.mark.parametrize("n", [8, 15, 31, 32])
def test_pressure_kernels_vs_reference_golden(golden_dir, n):
    from gpu_util import Dev, ptr, same_nan
    g = load(golden_dir, n)
    d = Dev(n)
    du, dv, us, vs, x = (d.up(g[k]) for k in ("d_u", "d_v", "u_star", "v_star", "x"))
    b = d.zeros()
    d.call("nf_continuity_rhs", d.gref(), ptr(us), ptr(vs), ptr(b))
    np.testing.assert_array_equal(d.down(b), g["rhs"])
    out = d.zeros()
    d.call("nf_pressure_apply", d.gref(), ptr(x), ptr(du), ptr(dv), ptr(out))
    np.testing.assert_array_equal(d.down(out), g["Ax"])
    d.call("nf_pressure_residual", d.gref(), ptr(x), ptr(b), ptr(du), ptr(dv), ptr(out))
    np.testing.assert_array_equal(d.down(out), g["rhs"] - g["Ax"])
    d.call("nf_jacobi_diag", d.gref(), ptr(du), ptr(dv), ptr(out))
    np.testing.assert_array_equal(d.down(out), g["jacobi_diag"])
    p, tmp = d.up(g["x"]), d.zeros()
    d.call("nf_jacobi_iterate", d.gref(), ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), 0.8, 3)
    np.testing.assert_array_equal(d.down(p), g["jacobi3"])
    p = d.up(g["x"])
    d.call("nf_rbsor_sweeps", d.gref(), ptr(p), ptr(b), ptr(du), ptr(dv), 1.5, 3)
    np.testing.assert_array_equal(d.down(p), g["rbsor3"])


@pytest.mark.parametrize("n", [8, 15, 31, 32])
def test_transfer_kernels_vs_reference_golden(golden_dir, n):
    from gpu_util import Dev, grid_for, ptr, rel, same_nan
    g = load(golden_dir, n)
    d = Dev(n)
    ctx = d.ctx
    x, du, dv = d.up(g["x"]), d.up(g["d_u"]), d.up(g["d_v"])
    for name, nc, key in (("nf_restrict_fw", (n - 1) // 2, "fw"), ("nf_restrict_inject", n // 2, "inject")):
        gc = grid_for(ctx, nc)
        c = ctx.empty(nc, nc)
        d.call(name, d.gref(), ptr(x), C.byref(gc), ptr(c))
        np.testing.assert_array_equal(ctx.download(c, nc, nc), g[key])
        # bilinear prolongation of the reference's coarse array back to n (the reference's index rules)
        cd = ctx.upload(g[key], nc, nc)
        f = d.zeros()
        d.call("nf_prolong_linear", C.byref(gc), ptr(cd), d.gref(), ptr(f), 0)
        np.testing.assert_array_equal(d.down(f), g["lin_from_" + key])
        f = d.up(g["x"])
        d.call("nf_prolong_linear", C.byref(gc), ptr(cd), d.gref(), ptr(f), 1)
        np.testing.assert_array_equal(d.down(f), g["x"] + g["lin_from_" + key])
        duc, dvc = ctx.empty(nc, nc), ctx.empty(nc, nc)
        d.call("nf_restrict_coeffs", d.gref(), ptr(du), ptr(dv), C.byref(gc), ptr(duc), ptr(dvc))
        assert same_nan(ctx.download(duc, nc + 1, nc), g["rc_du_" + key])
        assert same_nan(ctx.download(dvc, nc, nc + 1), g["rc_dv_" + key])
    if "cub_from_fw" in g:
        nc = (n - 1) // 2
        gc = grid_for(ctx, nc)
        cd = ctx.upload(g["fw"], nc, nc)
        f = d.zeros()
        ws = cubic_workspace(d, gc)
        d.call("nf_prolong_cubic", C.byref(gc), ptr(cd), d.gref(), ptr(f), 0, ptr(ws), ws.numel() * 8)
        assert rel(d.down(f), g["cub_from_fw"]) < 1e-13


@pytest.mark.parametrize("n", [8, 15, 31, 32])
def test_momentum_and_correction_vs_reference_golden(golden_dir, n):
    from gpu_util import Dev, NfLinks, bc_program_struct, ptr, rel, same_nan
    from naviflow_b200.host import practice_b_sides
    g = load(golden_dir, n)
    bc = cavity_bc()
    d = Dev(n)
    prog = bc_program_struct(bc, n, n)
    u, v = d.up(g["u"]), d.up(g["v"])
    d.call("nf_apply_velocity_bc", d.gref(), C.byref(prog), ptr(u), ptr(v))
    np.testing.assert_array_equal(d.down(u, n + 1, n), g["bc_u"])
    np.testing.assert_array_equal(d.down(v, n, n + 1), g["bc_v"])
    prog1 = bc_program_struct(bc, n, n, n + 1)   # callers that pass nx+1 (matrix_free_momentum.py:419)
    u1, v1 = d.up(g["u"]), d.up(g["v"])
    d.call("nf_apply_velocity_bc", d.gref(), C.byref(prog1), ptr(u1), ptr(v1))
    np.testing.assert_array_equal(d.down(u1, n + 1, n), g["bc1_u"])
    np.testing.assert_array_equal(d.down(v1, n, n + 1), g["bc1_v"])
    # links (un-relaxed: alpha = 1) against the reference's PowerLawDiscretization output
    p = d.up(g["p"])
    for is_u, tag, shape in ((1, "cu", (n + 1, n)), (0, "cv", (n, n + 1))):
        arrs = [d.zeros() for _ in range(6)]
        links = NfLinks(*[a.data_ptr() for a in arrs])
        dd = d.zeros()
        fn = "nf_momentum_links_u" if is_u else "nf_momentum_links_v"
        d.call(fn, d.gref(), ptr(u), ptr(v), ptr(p), MU, 1.0, practice_b_sides(bc), links, ptr(dd))
        for a, key in zip(arrs, ("a_e", "a_w", "a_n", "a_s", "a_p", "source")):
            got, want = d.down(a, *shape), g[f"{tag}_{key}"]
            assert rel(got, want) < 1e-14, (tag, key, rel(got, want))
    # full predictor (4 sweeps, alpha 0.7) against JacobiMatrixMomentumSolver
    import naviflow_b200 as nb
    mesh = nb.StructuredMesh(n, n)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=1000)
    ms = nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=4)
    us, du, norm, field = ms.solve_u_momentum(mesh, fluid, g["u"], g["v"], g["p"], 0.7, bc, return_dict=False)
    assert rel(us, g["mom_u_star"]) < 1e-13 and rel(field, g["mom_u_field"]) < 1e-11
    assert np.array_equal(np.isnan(du), np.isnan(g["mom_d_u"])) and rel(np.nan_to_num(du), np.nan_to_num(g["mom_d_u"])) < 1e-14
    assert abs(norm - g["mom_u_norm"]) <= 1e-11 * g["mom_u_norm"]
    vs, dv, norm, field = ms.solve_v_momentum(mesh, fluid, g["u"], g["v"], g["p"], 0.7, bc, return_dict=False)
    assert rel(vs, g["mom_v_star"]) < 1e-13 and rel(field, g["mom_v_field"]) < 1e-11
    assert np.array_equal(np.isnan(dv), np.isnan(g["mom_d_v"])) and rel(np.nan_to_num(dv), np.nan_to_num(g["mom_d_v"])) < 1e-14
    assert abs(norm - g["mom_v_norm"]) <= 1e-11 * g["mom_v_norm"]
    # velocity correction (bit exact) through the plugin class
    cu, cv = nb.GpuVelocityUpdater().update_velocity(mesh, g["u_star"], g["v_star"], g["x"], g["d_u"], g["d_v"], bc)
    np.testing.assert_array_equal(cu, g["corr_u"])
    np.testing.assert_array_equal(cv, g["corr_v"])


@pytest.mark.parametrize("n", [7, 9, 16, 63, 64, 65, 127, 200, 257])
def test_kernels_vs_oracle_seeded(n):
    """Fresh seeded inputs at sizes the golden files do not hold (odd, even, 2^k+-1)."""
    from gpu_util import Dev, grid_for, ptr, rel
    s = synth(n, 1000 + n)
    dx = dy = 1.0 / (n - 1)
    d = Dev(n)
    ctx = d.ctx
    du, dv, us, vs, x = (d.up(s[k]) for k in ("d_u", "d_v", "u_star", "v_star", "x"))
    b_ref = O.continuity_rhs(n, n, dx, dy, 1.0, s["u_star"], s["v_star"])
    b = d.zeros()
    d.call("nf_continuity_rhs", d.gref(), ptr(us), ptr(vs), ptr(b))
    np.testing.assert_array_equal(d.down(b), b_ref)
    out = d.zeros()
    d.call("nf_pressure_apply", d.gref(), ptr(x), ptr(du), ptr(dv), ptr(out))
    np.testing.assert_array_equal(d.down(out), O.apply_A(s["x"], dx, dy, 1.0, s["d_u"], s["d_v"]))
    p, tmp = d.up(s["x"]), d.zeros()
    d.call("nf_jacobi_iterate", d.gref(), ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), 0.8, 5)
    np.testing.assert_array_equal(d.down(p), O.jacobi_iterate(s["x"], b_ref, dx, dy, 1.0, s["d_u"], s["d_v"], 0.8, 5))
    p = d.up(s["x"])
    d.call("nf_rbsor_sweeps", d.gref(), ptr(p), ptr(b), ptr(du), ptr(dv), 1.5, 4)
    np.testing.assert_array_equal(d.down(p), O.rb_sor(s["x"], b_ref, dx, dy, 1.0, s["d_u"], s["d_v"], 1.5, 4))
    # transfers
    nc = (n - 1) // 2
    gc = grid_for(ctx, max(nc, 3)) if nc >= 3 else None
    if nc >= 3:
        gc = grid_for(ctx, nc)
        c = ctx.empty(nc, nc)
        d.call("nf_restrict_fw", d.gref(), ptr(x), C.byref(gc), ptr(c))
        fw = O.restrict_full_weighting(s["x"])
        np.testing.assert_array_equal(ctx.download(c, nc, nc), fw)
        f = d.zeros()
        d.call("nf_prolong_linear", C.byref(gc), ptr(c), d.gref(), ptr(f), 0)
        np.testing.assert_array_equal(d.down(f), O.prolong_linear(fw, n))
        if nc >= 4:
            ws = cubic_workspace(d, gc)
            d.call("nf_prolong_cubic", C.byref(gc), ptr(c), d.gref(), ptr(f), 0, ptr(ws), ws.numel() * 8)
            assert rel(d.down(f), O.prolong_cubic(fw, n)) < 1e-12
    # norms / dot
    val = C.c_double()
    d.call("nf_norm2", d.gref(), ptr(x), 0, C.byref(val))
    assert abs(val.value - np.linalg.norm(s["x"])) <= 1e-13 * np.linalg.norm(s["x"])
    d.call("nf_norm2", d.gref(), ptr(x), 1, C.byref(val))
    assert abs(val.value - np.linalg.norm(s["x"][1:-1, 1:-1])) <= 1e-13 * np.linalg.norm(s["x"])
    d.call("nf_dot", d.gref(), ptr(x), ptr(b), C.byref(val))
    ref = float(np.sum(s["x"] * b_ref))
    assert abs(val.value - ref) <= 1e-12 * np.linalg.norm(s["x"]) * np.linalg.norm(b_ref)
    # pressure update + max divergence
    pp = d.up(s["p"])
    pnew = d.zeros()
    d.call("nf_update_pressure", d.gref(), ptr(x), ptr(pp), 0.3, ptr(pnew))
    np.testing.assert_array_equal(d.down(pnew), O.update_pressure(s["x"], s["p"], 0.3, COND))
    u, v = d.up(s["u"]), d.up(s["v"])
    d.call("nf_max_abs_divergence", d.gref(), ptr(u), ptr(v), C.byref(val))
    assert val.value == O.max_interior_divergence(s["u"], s["v"], dx, dy)


def test_argument_errors_are_reported():
    from gpu_util import Dev, ptr
    from naviflow_b200._lib import NfError
    d = Dev(9)
    x = d.zeros()
    with pytest.raises(NfError, match="n_sweeps"):
        d.call("nf_rbsor_sweeps", d.gref(), ptr(x), ptr(x), ptr(x), ptr(x), 1.5, -1)
    with pytest.raises(NfError, match="alias"):
        d.call("nf_pressure_apply", d.gref(), ptr(x), ptr(x), ptr(x), ptr(x))


@pytest.mark.parametrize("path", ["tiled", "tma", "stream"])
@pytest.mark.parametrize("n", [7, 8, 31, 40, 53, 64, 65, 127, 130, 257])
def test_fused_rbsor_is_bit_identical(n, path, monkeypatch):
    """Temporally blocked red-black SOR (1..7 sweeps, tile/halo edges at every size class) against the oracle
    and against the unfused colour kernels; "tma" forces the persistent TMA pipeline and "stream" the streaming
    (wavefront) kernel at every size they accept."""
    from gpu_util import Dev, ptr
    monkeypatch.setenv("NF_RBSOR_TMA", "0" if path == "tma" else "1000000")
    monkeypatch.setenv("NF_RBSOR_STREAM", "0" if path == "stream" else "1000000000")
    s = synth(n, 2000 + n)
    dx = dy = 1.0 / (n - 1)
    d = Dev(n)
    du, dv, us, vs = (d.up(s[k]) for k in ("d_u", "d_v", "u_star", "v_star"))
    b_ref = O.continuity_rhs(n, n, dx, dy, 1.0, s["u_star"], s["v_star"])
    b = d.up(b_ref)
    inv = d.zeros()
    d.call("nf_pressure_inv_diag", d.gref(), ptr(du), ptr(dv), ptr(inv))
    np.testing.assert_array_equal(d.down(inv), 1.0 / O.sor_coefficients(n, n, dx, dy, 1.0, s["d_u"], s["d_v"])[4])
    for sweeps in (0, 1, 2, 3, 4, 7):
        want = O.rb_sor(s["x"], b_ref, dx, dy, 1.0, s["d_u"], s["d_v"], 1.5, sweeps)
        for use_inv in (None, inv):
            p, tmp = d.up(s["x"]), d.zeros()
            d.call("nf_rbsor_sweeps_fused", d.gref(), ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), ptr(use_inv), 1.5, sweeps)
            np.testing.assert_array_equal(d.down(p), want, err_msg=f"n={n} sweeps={sweeps} inv={use_inv is not None}")
        p2 = d.up(s["x"])
        d.call("nf_rbsor_sweeps", d.gref(), ptr(p2), ptr(b), ptr(du), ptr(dv), 1.5, sweeps)
        np.testing.assert_array_equal(d.down(p2), want)


@pytest.mark.parametrize("n,path", [(1025, "tma"), (1025, "stream"), (2049, "stream"), (2048, "stream")])
def test_fused_rbsor_large_grid_matches_unfused(n, path, monkeypatch):
    """Production regimes (several tiles per SM for the TMA pipeline; one job per warp of a single wave with row chunks,
    edge strips and an even-sized level for the streaming kernel) against the colour-pass kernels."""
    from gpu_util import Dev, ptr
    monkeypatch.setenv("NF_RBSOR_TMA", "0" if path == "tma" else "1000000")
    monkeypatch.setenv("NF_RBSOR_STREAM", "0" if path == "stream" else "1000000000")
    s = synth(n, 77)
    d = Dev(n)
    du, dv, us, vs = (d.up(s[k]) for k in ("d_u", "d_v", "u_star", "v_star"))
    b = d.zeros()
    d.call("nf_continuity_rhs", d.gref(), ptr(us), ptr(vs), ptr(b))
    for sweeps in (3, 5):
        p, tmp, p2 = d.up(s["x"]), d.zeros(), d.up(s["x"])
        inv = d.zeros()
        d.call("nf_pressure_inv_diag", d.gref(), ptr(du), ptr(dv), ptr(inv))
        d.call("nf_rbsor_sweeps_fused", d.gref(), ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), ptr(inv), 1.5, sweeps)
        d.call("nf_rbsor_sweeps", d.gref(), ptr(p2), ptr(b), ptr(du), ptr(dv), 1.5, sweeps)
        np.testing.assert_array_equal(d.down(p), d.down(p2))


@pytest.mark.parametrize("n,tma", [(9, "-1"), (31, "-1"), (40, "-1"), (64, "-1"), (65, "-1"), (130, "-1"), (257, "-1"),
                                   (9, "0"), (40, "0"), (65, "0"), (130, "0"), (257, "0"), (1200, None)])
def test_fused_momentum_sweeps_are_bit_identical(n, tma, monkeypatch):
    """Temporally blocked Jacobi momentum sweeps + fused residual against the sweep-by-sweep kernels; both operand paths of
    the fused kernel (NF_MOMENTUM_TMA=-1: loads by the threads, =0: TMA prefetch stage on every grid, default: on large grids)."""
    from gpu_util import Dev, NfLinks, bc_program_struct, ptr
    from naviflow_b200.host import practice_b_sides
    if tma is None:
        monkeypatch.delenv("NF_MOMENTUM_TMA", raising=False)
    else:
        monkeypatch.setenv("NF_MOMENTUM_TMA", tma)
    s = synth(n, 3000 + n)
    bc = cavity_bc()
    d = Dev(n)
    prog = bc_program_struct(bc, n, n)
    u, v, p = d.up(s["u"]), d.up(s["v"]), d.up(s["p"])
    d.call("nf_apply_velocity_bc", d.gref(), C.byref(prog), ptr(u), ptr(v))
    for is_u, shape in ((1, (n + 1, n)), (0, (n, n + 1))):
        arrs = [d.zeros() for _ in range(6)]
        links = NfLinks(*[a.data_ptr() for a in arrs])
        dd = d.zeros()
        fn = "nf_momentum_links_u" if is_u else "nf_momentum_links_v"
        d.call(fn, d.gref(), ptr(u), ptr(v), ptr(p), MU, 0.7, practice_b_sides(bc), links, ptr(dd))
        x0 = s["u"] if is_u else s["v"]
        for sweeps in (1, 2, 5, 6, 7, 13):
            xa, ta, fa = d.up(x0), d.zeros(), d.zeros()
            xb, tb, fb = d.up(x0), d.zeros(), d.zeros()
            na, nbv = C.c_double(), C.c_double()
            d.call("nf_momentum_jacobi", d.gref(), is_u, links, ptr(xa), ptr(ta), sweeps)
            d.call("nf_momentum_residual", d.gref(), is_u, links, ptr(xa), ptr(fa), C.byref(na))
            d.call("nf_momentum_jacobi_fused", d.gref(), is_u, links, ptr(xb), ptr(tb), sweeps, ptr(fb), C.byref(nbv))
            np.testing.assert_array_equal(d.down(xb, *shape), d.down(xa, *shape), err_msg=f"n={n} u={is_u} k={sweeps}")
            np.testing.assert_array_equal(d.down(fb, *shape), d.down(fa, *shape))
            assert abs(na.value - nbv.value) <= 1e-12 * abs(na.value)


@pytest.mark.gpu
def test_quick_and_second_order_upwind_links_vs_reference_golden(golden_dir):
    """nf_momentum_links_ext through the plugin twins of QUICKDiscretization / SecondOrderUpwindDiscretization
    (quick.py:27-219, second_order_upwind.py:26-325) against the reference's outputs: all ten arrays of both components,
    with the cavity's boundary conditions and with bc=None, square / rectangular / minimal grids -- bit-exact."""
    import naviflow_b200 as nb
    G = np.load(os.path.join(golden_dir, "ext_links_kats.npz"))
    bc = nb.BoundaryConditionManager()
    bc.set_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        bc.set_condition(b, "wall")
    keys = ("a_e", "a_w", "a_n", "a_s", "a_ee", "a_ww", "a_nn", "a_ss", "a_p", "source")
    for tag in G["cases"]:
        tag = str(tag)
        nx, ny = (int(x) for x in G[f"{tag}_dims"])
        dx, dy, rho, mu = (float(x) for x in G[f"{tag}_scal"])
        mesh = nb.StructuredMesh(nx, ny, 1.0, 0.7 if nx != ny else 1.0)   # as oracle/make_golden.py:ext_links_kats
        assert mesh.get_cell_sizes() == (dx, dy)
        fluid = nb.FluidProperties(density=rho, viscosity=mu)
        u, v, p = G[f"{tag}_u"], G[f"{tag}_v"], G[f"{tag}_p"]
        for sch, cls in (("quick", nb.GpuQUICKDiscretization), ("sou", nb.GpuSecondOrderUpwindDiscretization)):
            d = cls()
            for bname, b in (("bc", bc), ("nobc", None)):
                for comp, got in (("u", d.calculate_u_coefficients(mesh, fluid, u, v, p, b)),
                                  ("v", d.calculate_v_coefficients(mesh, fluid, u, v, p, b))):
                    assert tuple(sorted(got)) == tuple(sorted(keys))
                    for k in keys:
                        np.testing.assert_array_equal(got[k], G[f"{tag}_{sch}_{bname}_{comp}_{k}"],
                                                      err_msg=f"{tag} {sch} {bname} {comp} {k}")
