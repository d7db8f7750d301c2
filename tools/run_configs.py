#!/usr/bin/env python
"""Runs the BASELINE.json configurations that are not the bench line and prints one JSON record per config.

  python tools/run_configs.py c1            63^2 Re=100, FMG multigrid (GS_vcycle.py settings), GPU vs CPU oracle port
  python tools/run_configs.py c2 [iters]    1025^2 Re=1000: CG pressure solve timing + long multigrid run with Ghia check
  python tools/run_configs.py c4 [n]        pressure-solver sweep (Jacobi / RB-SOR / BiCGSTAB / CG / multigrid) at 2049^2
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def cavity(nb, n, Re):
    mesh = nb.StructuredMesh(n, n, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=Re, characteristic_velocity=1.0)
    return mesh, fluid


def set_bcs(alg):
    alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        alg.set_boundary_condition(b, "wall")


def c1():
    """configs[0]: the reference's own CPU-runnable case with the deterministic momentum solver:
    1500 fixed outer iterations; golden norms measured with the real reference in the survey."""
    import torch
    import naviflow_b200 as nb
    import bench  # the CPU-baseline leg (the only place outside tests/ that runs oracle/)
    n, Re, k, N = 63, 100, 20, 1500
    mesh, fluid = cavity(nb, n, Re)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5, method_type="red_black"), max_iterations=100,
                               tolerance=1e-3, pre_smoothing=3, post_smoothing=3, cycle_type="fmg",
                               cycle_type_buildup="v", cycle_type_final="v", max_cycles_buildup=1,
                               restriction_method="restrict_full_weighting", interpolation_method="interpolate_cubic",
                               coarsest_grid_size=7)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=k), alpha_p=0.3, alpha_u=0.7)
    set_bcs(alg)
    alg.solve(max_iterations=5, tolerance=0.0, save_profile=False)  # warm up
    alg.initialize_fields()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    alg.solve(max_iterations=N, tolerance=0.0, save_profile=False)
    t_gpu = time.perf_counter() - t0
    inf, l2 = nb.ghia_errors(alg.u, alg.v, mesh, Re)
    # CPU oracle port, bounded sample: 100 iterations
    t_cpu = bench.cpu_baseline_simple_run(n, Re, k, 100, dict(omega=1.5, pre=3, post=3, cycle_type="fmg", cycle_type_final="v",
                                                               interpolation="interpolate_cubic", tolerance=1e-3))
    golden = {"u": 14.5192971946493, "v": 9.52637420908884, "p": 33.7706672709561, "ghia_inf": 0.05498, "ghia_l2": 0.01865}
    rec = {"config": "c1: 63^2 Re=100 SIMPLE, FMG(1)+V(3,3) RB-SOR 1.5 cubic, 20 Jacobi momentum sweeps, 1500 iterations",
           "gpu_s_per_iter": t_gpu / N, "gpu_iter_per_s": N / t_gpu, "cpu_port_s_per_iter": t_cpu,
           "reference_published_s_per_iter": 0.04185,
           "norm_u": float(np.linalg.norm(alg.u)), "norm_v": float(np.linalg.norm(alg.v)), "norm_p": float(np.linalg.norm(alg.p)),
           "reference_norms": golden, "ghia_inf": inf, "ghia_l2": l2,
           "rel_dev_u": abs(np.linalg.norm(alg.u) - golden["u"]) / golden["u"],
           "rel_dev_v": abs(np.linalg.norm(alg.v) - golden["v"]) / golden["v"],
           "rel_dev_p": abs(np.linalg.norm(alg.p) - golden["p"]) / golden["p"]}
    print(json.dumps(rec))


def c2(iters):
    """configs[1]: 1025^2 Re=1000.  (a) matrix-free CG pressure solve inside SIMPLE: time per outer iteration and CG
    iterations; (b) the same case run long with the multigrid pressure solve: Ghia centre-line validation."""
    import torch
    import naviflow_b200 as nb
    n, Re = 1025, 1000
    mesh, fluid = cavity(nb, n, Re)
    out = {"config": "c2: 1025^2 Re=1000"}
    for name, ps in (("cg", nb.GpuCGSolver(tolerance=1e-7, max_iterations=3000)),
                     ("bicgstab", nb.GpuBiCGSTABSolver(tolerance=1e-7, max_iterations=3000))):
        alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=5), alpha_p=0.3, alpha_u=0.7)
        set_bcs(alg)
        alg.solve(max_iterations=2, tolerance=0.0, save_profile=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = alg.solve(max_iterations=10, tolerance=0.0, save_profile=False)
        dt = time.perf_counter() - t0
        its = alg.pressure_iterations_history
        out[name] = {"s_per_outer_iter": dt / 10, "krylov_iters_per_solve": float(np.mean(its)),
                     "us_per_krylov_iter": dt / max(1, sum(its)) * 1e6, "p_rel_norm_last": res.get_history("p_rel_norm")[-1]}
        del alg
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=100, tolerance=1e-3,
                               pre_smoothing=3, post_smoothing=3)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=5), alpha_p=0.3, alpha_u=0.7)
    set_bcs(alg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hist = []
    done = 0
    while done < iters:
        chunk = min(2000, iters - done)
        alg.push_fields() if done == 0 else None
        recs = alg.iterate_resident(chunk, 0.0)
        done += chunk
        alg.pull_fields()
        inf, l2 = nb.ghia_errors(alg.u, alg.v, mesh, Re)
        hist.append({"iter": done, "u_rel_norm": recs[-1]["u_rel_norm"], "v_rel_norm": recs[-1]["v_rel_norm"],
                     "ghia_inf": inf, "ghia_l2": l2, "max_div": alg.get_max_divergence(),
                     "mg_cycles": recs[-1]["pressure_iterations"], "elapsed_s": time.perf_counter() - t0})
        print(json.dumps(hist[-1]), file=sys.stderr, flush=True)
    out["multigrid_long_run"] = {"iterations": done, "s_total": time.perf_counter() - t0, "history": hist}
    print(json.dumps(out))


def c4(n):
    """configs[3]: single pressure solve A p' = b with the seeded synthetic inputs of SURVEY 8d (C4)."""
    import ctypes as C
    import torch
    import naviflow_b200 as nb
    from naviflow_b200._lib import NfKrylovInfo, NfMgInfo
    from naviflow_b200.device import get_context, pad_ld, ptr
    import bench  # the CPU-baseline leg (the only place outside tests/ that runs oracle/)
    mu = 1e-3
    rng = np.random.default_rng(0)
    dx = dy = 1.0 / (n - 1)
    d_u = (0.7 * dy / (4 * mu)) * (1 + 0.1 * rng.random((n + 1, n)))
    d_v = (0.7 * dx / (4 * mu)) * (1 + 0.1 * rng.random((n, n + 1)))
    us = 1e-2 * rng.standard_normal((n + 1, n)); us[0, :] = us[n, :] = 0
    vs = 1e-2 * rng.standard_normal((n, n + 1)); vs[:, 0] = vs[:, n] = 0
    ctx = get_context(0)
    lib, H = ctx.lib, ctx.handle
    g = ctx.grid(n, n, dx, dy, 1.0)
    G = C.byref(g)
    du, dv, usd, vsd = (ctx.upload(a, n, n) for a in (d_u, d_v, us, vs))
    b, x, tmp, r = (ctx.empty(n, n) for _ in range(4))
    ctx.check(lib.nf_continuity_rhs(H, G, ptr(usd), ptr(vsd), ptr(b)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    val = C.c_double()

    def relres():
        ctx.check(lib.nf_pressure_residual(H, G, ptr(x), ptr(b), ptr(du), ptr(dv), ptr(r)))
        ctx.check(lib.nf_norm2(H, G, ptr(r), 0, C.byref(val))); rn = val.value
        ctx.check(lib.nf_norm2(H, G, ptr(b), 0, C.byref(val)))
        return rn / val.value

    def timed(fn):
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    rec = {"config": f"c4: pressure-solver sweep at {n}^2, synthetic seeded system (SURVEY 8d C4)", "n": n, "solvers": {}}
    cells = float(n) * n
    # stationary solvers: time per iteration over 200 iterations, residual reduction reached
    for name, call in (("jacobi_w0.8", lambda k: lib.nf_jacobi_iterate(H, G, ptr(x), ptr(tmp), ptr(b), ptr(du), ptr(dv), 0.8, k)),
                       ("rbsor_w1.5", lambda k: lib.nf_rbsor_sweeps_fused(H, G, ptr(x), ptr(tmp), ptr(b), ptr(du), ptr(dv), None, 1.5, k))):
        x.zero_()
        ctx.check(call(6))
        x.zero_()
        ms = timed(lambda: ctx.check(call(300)))
        rec["solvers"][name] = {"ms_per_iteration": ms / 300, "MLUPS": cells * 300 / ms / 1e3, "rel_residual_after_300": relres()}
    # sequential Gauss-Seidel of the reference's `04 gauss_seidel` script (lexicographic, omega 1.8) and the symmetric variant:
    # block wavefronts (nf_gs_lex.cu), 20 sweeps
    for name, sym in (("gs_lexicographic_w1.8", 0), ("gs_symmetric_w1.8", 1)):
        x.zero_()
        ctx.check(lib.nf_gs_lex_sweeps(H, G, ptr(x), ptr(b), ptr(du), ptr(dv), 1.8, 1, sym))
        x.zero_()
        ms = timed(lambda: ctx.check(lib.nf_gs_lex_sweeps(H, G, ptr(x), ptr(b), ptr(du), ptr(dv), 1.8, 20, sym)))
        rec["solvers"][name] = {"ms_per_iteration": ms / 20, "MLUPS": cells * 20 / ms / 1e3, "rel_residual_after_20": relres()}
    # Krylov
    work = torch.zeros((5 * (n + 1), pad_ld(n)), dtype=torch.float64, device=x.device)
    for name, fn in (("bicgstab", lib.nf_bicgstab_solve), ("cg", lib.nf_cg_solve)):
        info = NfKrylovInfo()
        ms = timed(lambda: ctx.check(fn(H, G, ptr(b), ptr(x), ptr(du), ptr(dv), 1e-7, 1e-5, 5000, 50, ptr(work), C.byref(info))))
        rec["solvers"][name] = {"iterations": info.iterations, "scipy_info": info.info, "ms_total": ms,
                                "ms_per_iteration": ms / max(1, info.iterations), "rel_residual": relres(),
                                "MLUPS": cells * info.iterations / ms / 1e3}
    # multigrid V(3,3)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=100, tolerance=1e-5,
                               pre_smoothing=3, post_smoothing=3)
    cfg = ps.config_struct(1.0, 1.0)
    mg = C.c_void_p()
    ctx.check(lib.nf_mg_create(H, C.byref(mg), n, n, pad_ld(n), C.byref(cfg)))
    ctx.check(lib.nf_mg_setup(mg, ptr(du), ptr(dv)))
    mi = NfMgInfo()
    ctx.check(lib.nf_mg_solve(mg, ptr(b), ptr(x), ptr(r), C.byref(mi)))
    ms = timed(lambda: ctx.check(lib.nf_mg_solve(mg, ptr(b), ptr(x), ptr(r), C.byref(mi))))
    rec["solvers"]["multigrid_v33"] = {"cycles_to_1e-5": mi.cycles, "ms_total": ms, "ms_per_cycle": ms / max(1, mi.cycles),
                                       "rel_residual": mi.r_norm / mi.b_norm, "MLUPS_cycles": cells * mi.cycles / ms / 1e3}
    lib.nf_mg_destroy(mg)
    # CPU oracle port: one iteration / sweep / cycle each (bounded)
    cpu = bench.cpu_baseline_pressure_kernels(n, dx, dy, d_u, d_v, us, vs, with_mg=(n <= 2049))
    rec["cpu_oracle_port_single_thread"] = cpu
    print(json.dumps(rec))


def ghia(n, iters, k, chunk):
    """Long run towards the steady state with the multigrid pressure solve: Ghia centre-line errors vs iteration."""
    import torch
    import naviflow_b200 as nb
    Re = 1000
    mesh, fluid = cavity(nb, n, Re)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=100, tolerance=1e-3,
                               pre_smoothing=3, post_smoothing=3)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=k), alpha_p=0.3, alpha_u=0.7)
    set_bcs(alg)
    alg.push_fields()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hist, done = [], 0
    while done < iters:
        c = min(chunk, iters - done)
        recs = alg.iterate_resident(c, 0.0)
        done += c
        alg.pull_fields()
        inf, l2 = nb.ghia_errors(alg.u, alg.v, mesh, Re)
        hist.append({"iter": done, "u_rel_norm": recs[-1]["u_rel_norm"], "v_rel_norm": recs[-1]["v_rel_norm"],
                     "ghia_inf": inf, "ghia_l2": l2, "max_div": alg.get_max_divergence(),
                     "elapsed_s": time.perf_counter() - t0})
        print(json.dumps(hist[-1]), file=sys.stderr, flush=True)
    nx = n
    rec = {"config": f"ghia: {n}^2 Re=1000 SIMPLE (alpha 0.3/0.7), {k} Jacobi momentum sweeps, multigrid V(3,3) to 1e-3",
           "iterations": done, "s_total": time.perf_counter() - t0, "history": hist,
           "u_centerline": alg.u[nx // 2, :: max(1, n // 64)].tolist(), "v_centerline": alg.v[:: max(1, n // 64), n // 2].tolist()}
    print(json.dumps(rec))


def c3(n, k, max_iters):
    """configs[2]: n^2 Re=1000 SIMPLE + multigrid V(3,3) run to the reference's stopping test
    max(u_rel_norm, v_rel_norm) <= 1e-6 (simple.py:114, :174)."""
    import torch
    import naviflow_b200 as nb
    Re = 1000
    mesh, fluid = cavity(nb, n, Re)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=100, tolerance=1e-3,
                               pre_smoothing=3, post_smoothing=3)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=k), alpha_p=0.3, alpha_u=0.7)
    set_bcs(alg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = alg.solve(max_iterations=max_iters, tolerance=1e-6, save_profile=False)
    dt = time.perf_counter() - t0
    h = res.get_history("total_rel_norm")[::2]
    inf, l2 = nb.ghia_errors(alg.u, alg.v, mesh, Re)
    rec = {"config": f"c3: {n}^2 Re=1000 SIMPLE to max(u,v rel_norm) <= 1e-6, {k} Jacobi momentum sweeps, multigrid V(3,3) to 1e-3",
           "iterations": res.iterations, "converged": bool(h[-1] <= 1e-6), "final_total_rel_norm": h[-1],
           "wall_s_including_h2d_d2h": dt, "iter_per_s": res.iterations / dt, "MLUPS": n * n * res.iterations / dt / 1e6,
           "mg_cycles_mean": float(np.mean(alg.pressure_iterations_history)),
           "residual_trajectory": {str(i): h[i - 1] for i in (1, 10, 100, 1000, 2000, 5000, 10000, 20000) if i <= len(h)},
           "p_rel_norm_last": res.get_history("p_rel_norm")[-1], "max_divergence": alg.get_max_divergence(),
           "ghia_inf_at_stop": inf, "ghia_l2_at_stop": l2,
           "note": "the stopping quantity is the momentum solver's rel_norm, here the relaxed inner residual of the "
                   "Jacobi-sweep momentum solver (SURVEY 7.3-9): it reaches 1e-6 long before the flow is steady"}
    print(json.dumps(rec))


def c3u(n, k, budget_s, max_iters):
    """configs[2] run to the ABSOLUTE UNRELAXED momentum residual (matrix_free_momentum.py:380-400) <= 1e-6 -- the
    convergence measure that does not depend on the momentum solver -- or until the wall-clock budget is
    spent; the record holds the trajectory of that norm, of the reference's own stopping quantity and of the Ghia errors."""
    import torch
    import naviflow_b200 as nb
    Re = 1000
    mesh, fluid = cavity(nb, n, Re)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=100, tolerance=1e-3,
                               pre_smoothing=3, post_smoothing=3)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=k), alpha_p=0.3, alpha_u=0.7,
                             track_unrelaxed_residual=True)
    set_bcs(alg)
    alg.push_fields()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done, traj, unrel, cycles = 0, [], 1.0, []
    chunk = 250
    while done < max_iters and (time.perf_counter() - t0) < budget_s:
        recs = alg.iterate_resident(min(chunk, max_iters - done), 0.0)
        done += len(recs)
        cycles += [r["pressure_iterations"] for r in recs]
        unrel = max(recs[-1]["u_unrelaxed_res"], recs[-1]["v_unrelaxed_res"])
        hit = next((i for i, r in enumerate(recs) if max(r["u_unrelaxed_res"], r["v_unrelaxed_res"]) <= 1e-6), None)
        alg.pull_fields()
        inf, l2 = nb.ghia_errors(alg.u, alg.v, mesh, Re)
        traj.append({"iter": done, "unrelaxed_momentum_residual": unrel,
                     "relaxed_rel_norm": max(recs[-1]["u_rel_norm"], recs[-1]["v_rel_norm"]), "ghia_inf": inf, "ghia_l2": l2,
                     "elapsed_s": time.perf_counter() - t0})
        print(json.dumps(traj[-1]), file=sys.stderr, flush=True)
        if hit is not None:
            break
    dt = time.perf_counter() - t0
    rec = {"config": f"c3u: {n}^2 Re=1000 SIMPLE to the absolute unrelaxed momentum residual <= 1e-6 (budget {budget_s} s, "
                     f"{max_iters} iterations), {k} Jacobi momentum sweeps, multigrid V(3,3) to 1e-3",
           "iterations": done, "converged_unrelaxed_1e-6": bool(unrel <= 1e-6), "final_unrelaxed_residual": unrel,
           "wall_s_including_checkpoint_downloads": dt, "iter_per_s": done / dt, "mg_cycles_mean": float(np.mean(cycles)),
           "trajectory": traj, "max_divergence": alg.get_max_divergence()}
    print(json.dumps(rec))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c1"
    if which == "c1":
        c1()
    elif which == "c2":
        c2(int(sys.argv[2]) if len(sys.argv) > 2 else 20000)
    elif which == "c4":
        c4(int(sys.argv[2]) if len(sys.argv) > 2 else 2049)
    elif which == "c3":
        c3(int(sys.argv[2]) if len(sys.argv) > 2 else 4097, int(sys.argv[3]) if len(sys.argv) > 3 else 20,
           int(sys.argv[4]) if len(sys.argv) > 4 else 20000)
    elif which == "c3u":
        c3u(int(sys.argv[2]) if len(sys.argv) > 2 else 4097, int(sys.argv[3]) if len(sys.argv) > 3 else 20,
            float(sys.argv[4]) if len(sys.argv) > 4 else 240.0, int(sys.argv[5]) if len(sys.argv) > 5 else 200000)
    elif which == "ghia":
        ghia(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
