#!/usr/bin/env python
"""bench.py -- SIMPLE outer-iteration throughput on the BASELINE.json workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid 4097]

Workload (config.workload): lid-driven cavity n x n (default 4097, BASELINE configs[2]), Re = 1000, from rest,
SIMPLE (alpha_p 0.3, alpha_u 0.7) with JacobiMatrixMomentumSolver-style fixed Jacobi momentum sweeps and the
geometric-multigrid pressure solve: V-cycles, red-black SOR smoother omega 1.5, 3 pre + 3 post sweeps, full
weighting + bilinear prolongation, coarsest 7, cycles until ||r||/||b|| < 1e-3 (at most `--mg-cycles`).
A step = one SIMPLE outer iteration.  value = cells * iterations / time (MLUPS), device resident.
e2e = the same through GpuSimpleSolver.solve() with host arrays (H2D of u,v,p and D2H of u,v,p
inside the timed region, one outer iteration per call).
--impl reference: the reference algorithm's CPU path (NumPy oracle port, oracle/np_oracle.py) on a bounded
sample grid of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

RE = 1000.0
MG = dict(omega=1.5, pre=3, post=3, tol=1e-3)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=0, help="grid size (cells per side); default 4097 on 1 GPU and "
                    "4097*sqrt(N) on N GPUs (weak scaling: 16.8 M cells per GPU)")
    ap.add_argument("--momentum-sweeps", type=int, default=5)
    ap.add_argument("--mg-cycles", type=int, default=100, help="max V-cycles per pressure solve")
    ap.add_argument("--cpu-sample-n", type=int, default=0, help="grid of the bounded CPU sample (0: 513 inside the "
                    "GPU arm's cpu_baseline, 1025 or 513 for --impl reference depending on --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-kernels", action="store_true", help="skip the port's 4097^2 kernel timings (~20 s)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_solver(n, args):
    import naviflow_b200 as nb
    mesh = nb.StructuredMesh(n, n, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=RE, characteristic_velocity=1.0)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=MG["omega"], method_type="red_black"),
                               max_iterations=args.mg_cycles, tolerance=MG["tol"], pre_smoothing=MG["pre"],
                               post_smoothing=MG["post"], cycle_type="v",
                               restriction_method="restrict_full_weighting",
                               interpolation_method="interpolate_linear", coarsest_grid_size=7)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=args.momentum_sweeps),
                             nb.GpuVelocityUpdater(), alpha_p=0.3, alpha_u=0.7)
    alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        alg.set_boundary_condition(b, "wall")
    return alg


def cpu_oracle_step_fn(n, args):
    """One SIMPLE outer iteration of the reference algorithm on the host (NumPy port), same settings."""
    from oracle import np_oracle as O
    cfg = O.MGConfig(omega=MG["omega"], pre=MG["pre"], post=MG["post"], tolerance=MG["tol"],
                     max_iterations=args.mg_cycles)
    ps = O.make_pressure_solver("mg", cfg=cfg)
    state = {"st": None}

    def step():
        st, h = O.simple_solve(n, n, RE, ps, n_sweeps=args.momentum_sweeps, max_iterations=1, tolerance=0.0,
                               state=state["st"])
        state["st"] = st
        return h
    return step


# ---- cpu_baseline leg: every use of oracle/ outside tests/ and __graft_entry__.smoke() lives in the functions of this
# section (cpu_oracle_step_fn above, run_reference below, and the two helpers the config runner tools/run_configs.py calls)
def cpu_baseline_simple_run(n, reynolds, n_sweeps, iterations, mg_kwargs):
    """Seconds per outer iteration of the oracle port's SIMPLE loop with the given multigrid settings (bounded run)."""
    from oracle import np_oracle as O
    cfg = O.MGConfig(**mg_kwargs)
    t0 = time.perf_counter()
    O.simple_solve(n, n, reynolds, O.make_pressure_solver("mg", cfg=cfg), n_sweeps=n_sweeps, max_iterations=iterations,
                   tolerance=0.0)
    return (time.perf_counter() - t0) / iterations


def cpu_baseline_pressure_kernels(n, dx, dy, d_u, d_v, us, vs, with_mg=True, with_lex=True):
    """Milliseconds of one application / iteration / sweep / cycle of the oracle port's pressure kernels on the given
    system (single thread), plus one sequential Gauss-Seidel sweep timed on a 257^2 sample and scaled per cell."""
    from oracle import np_oracle as O
    bh = O.continuity_rhs(n, n, dx, dy, 1.0, us, vs)
    cpu = {}
    t0 = time.perf_counter(); O.apply_A(bh, dx, dy, 1.0, d_u, d_v); cpu["A_p_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); O.jacobi_iterate(np.zeros_like(bh), bh, dx, dy, 1.0, d_u, d_v, 0.8, 1)
    cpu["jacobi_iteration_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); O.rb_sor(np.zeros_like(bh), bh, dx, dy, 1.0, d_u, d_v, 1.5, 1)
    cpu["rbsor_sweep_ms"] = (time.perf_counter() - t0) * 1e3
    if with_mg and with_lex:
        ns = 257
        rs = np.random.default_rng(5)
        dus = (0.7 * dy / 4e-3) * (1 + 0.1 * rs.random((ns + 1, ns)))
        dvs = (0.7 * dx / 4e-3) * (1 + 0.1 * rs.random((ns, ns + 1)))
        t0 = time.perf_counter()
        O.gs_lex(np.zeros((ns, ns)), 1e-2 * rs.standard_normal((ns, ns)), dx, dy, 1.0, dus, dvs, 1.8, 1)
        cpu["gs_lexicographic_sweep_ms_scaled_from_257"] = (time.perf_counter() - t0) * 1e3 * (n * n) / (ns * ns)
    if with_mg:
        mcfg = O.MGConfig(omega=1.5, pre=3, post=3)
        t0 = time.perf_counter(); O.mg_cycle(mcfg, np.zeros_like(bh), bh, dx, dy, d_u, d_v)
        cpu["mg_v33_cycle_ms"] = (time.perf_counter() - t0) * 1e3
    return cpu


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        t = [p.get("num_threads", 1) for p in threadpool_info()]
        return max(t) if t else 1
    except Exception:
        return 1


def reference_sample_n(args):
    """Grid of the bounded CPU sample: 1025^2 (about 8 s per outer iteration of the NumPy port) when the whole
    --steps/--warmup run then stays within a few minutes, else 513^2 (about 2 s)."""
    if args.cpu_sample_n > 0:
        return args.cpu_sample_n
    return 1025 if (args.steps + args.warmup) <= 12 else 513


def run_reference(args, rank, world):
    """Reference arm: the reference algorithm's CPU path (oracle port: /root/reference is pure Python and not
    present on the GPU box) on the host cores; rank 0 only.  `config` is this repo's arm's config (the workload
    both arms are quoted on); each timed step is one outer iteration of that workload on the bounded sample grid
    named in `sample` / `cpu_baseline.sample` -- MLUPS is per cell, so the grid size cancels to first order."""
    if rank != 0:
        return
    n = reference_sample_n(args)
    step = cpu_oracle_step_fn(n, args)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    mlups = n * n * args.steps / dt / 1e6
    sample = (f"{n}x{n} grid of the same workload (same Re, relaxation, momentum sweeps, V(3,3) settings and stopping "
              f"tolerance), {args.steps} outer iterations after {args.warmup} warm-up; NumPy/SciPy oracle port, single "
              f"threaded (host has {os.cpu_count()} cores, BLAS threads {cpu_threads()})")
    line = {
        "impl": "reference", "metric": "simple_outer_mlups", "value": mlups, "unit": "MLUPS",
        "iter_per_s": args.steps / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.n, world),
        "sample": {"n": n, "cells": n * n, "note": "bounded sample of config.workload: ms_per_step and iter_per_s are "
                                                   "per outer iteration of THIS grid, value (MLUPS) is per cell"},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n, world=1):
    """The workload both arms are quoted on (identical dict in both lines)."""
    return {"workload": f"lid-driven cavity {n}x{n} Re=1000 SIMPLE from rest, {args.momentum_sweeps} Jacobi momentum "
                        f"sweeps/component, multigrid V(3,3) RB-SOR omega 1.5 FW+bilinear coarsest 7, cycles to "
                        f"||r||/||b||<1e-3 (max {args.mg_cycles})",
            "n": n, "reynolds": RE, "alpha_p": 0.3, "alpha_u": 0.7,
            "l2_policy": "fields (134 MB each at 4097^2, ~25 live arrays) exceed the 126 MB L2; no explicit flush",
            "parallelism": "single GPU" if world == 1 else f"{world} row slabs, one rank per GPU"}


WEAK_N = {1: 4097, 2: 5793, 4: 8193, 8: 11585}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.n <= 0:
        args.n = WEAK_N.get(world, int(round(4097 * world ** 0.5)))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: naviflow_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import ctypes as C
    from naviflow_b200.device import get_context, ptr
    n = args.n
    alg = build_solver(n, args)
    ctx = get_context(local)
    alg.push_fields()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region -------------------------------------------------------------
    if args.warmup > 0:
        alg.iterate_resident(args.warmup, 0.0)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    recs = alg.iterate_resident(args.steps, 0.0)
    e1.record()
    barrier()
    launches = ctx.launches() - l0        # kernels launched (graph nodes included) inside the timed region
    # phases of the step on the following `steps` iterations (4 CUDA events per iteration between the phases' launches;
    # the V-cycle graph is still replayed): momentum predictor / pressure solve incl. RHS + hierarchy / corrections
    alg.phase_timing(True)
    precs = alg.iterate_resident(args.steps, 0.0)
    phases = alg.phase_timing(False)
    live_ms, live_launches = 0.0, 0
    if world == 1:
        # the dominant kernel inside the real step: two more outer iterations with CUDA-event pairs around every
        # finest-level smoother launch (the instrumented iterations launch kernel by kernel instead of replaying the
        # V-cycle graph, so they are kept out of the timed region above)
        alg.smoother_timing(True)
        alg.iterate_resident(2, 0.0)
        live_ms, live_launches = alg.smoother_timing(False)
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    cells = float(n) * n                  # one grid, cut into row slabs over the ranks
    mlups = cells * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (the finest-level smoother launch) ---------------------------
    peak, peak_src = measured_peaks()
    lib = ctx.lib
    nr = min(n, 4097)                      # the dominant kernel is also timed alone on a 4097^2 level (> L2)
    g = ctx.grid(nr, nr, 1.0 / (nr - 1), 1.0 / (nr - 1), 1.0)
    rng = np.random.default_rng(0)
    mk = lambda scale: ctx.upload(scale * (1 + 0.1 * rng.random((nr + 1, nr + 1))), nr, nr)
    roof = {"b": mk(1e-3), "d_u": mk(40.0 / nr), "d_v": mk(40.0 / nr)}
    fld = lambda name: ptr(roof[name])
    n_alg, n = n, nr
    scratch, scratch2 = ctx.empty(n, n), ctx.empty(n, n)
    reps = 8
    inv = ctx.empty(n, n)
    ctx.check(lib.nf_pressure_inv_diag(ctx.handle, C.byref(g), fld("d_u"), fld("d_v"), ptr(inv)))
    args_f = (ctx.handle, C.byref(g), ptr(scratch), ptr(scratch2), fld("b"), fld("d_u"), fld("d_v"), ptr(inv), 1.5)
    ctx.check(lib.nf_rbsor_sweeps_fused(*args_f, 6))
    torch.cuda.synchronize()
    e0.record()
    ctx.check(lib.nf_rbsor_sweeps_fused(*args_f, 3 * reps))       # `reps` plain 3-sweep launches
    e1.record()
    torch.cuda.synchronize()
    ms_launch = e0.elapsed_time(e1) / reps
    alg_bytes = 3 * 40.0 * n * n            # 40 B/cell per full sweep x 3 sweeps per launch
    achieved = alg_bytes / (ms_launch * 1e-3) / 1e9
    # the unfused colour-pass kernel (one launch = half a sweep = 20 B/cell) timed the same way, for reference
    e0.record()
    ctx.check(lib.nf_rbsor_sweeps(ctx.handle, C.byref(g), ptr(scratch), fld("b"), fld("d_u"), fld("d_v"), 1.5, reps))
    e1.record()
    torch.cuda.synchronize()
    ms_color = e0.elapsed_time(e1) / (2 * reps)
    # DRAM bytes per launch of each variant from this round's ncu capture (dram__bytes_read.sum + dram__bytes_write.sum)
    prof = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r2_dominant_kernel.json")) as f:
            prof = json.load(f)
    except Exception:
        pass
    variants = prof.get("variants", {}) if prof.get("n") == n else {}
    plain_traffic = variants.get("plain", {}).get("dram_bytes_per_launch")
    isolated = {"variant": "plain (3 sweeps)", "ms_per_launch": ms_launch, "achieved": achieved, "frac": achieved / peak,
                "traffic": plain_traffic,
                "dram_frac": (plain_traffic / (ms_launch * 1e-3) / 1e9 / peak) if plain_traffic else None,
                "how": f"{reps} back-to-back launches on a synthetic {n}^2 level, CUDA events"}
    traffic = plain_traffic
    live = live_launches > 0 and n_alg == n
    if live:   # the number the roofline is quoted on: launches inside the real steps
        ms_launch = live_ms / live_launches
        if os.environ.get("NF_RBSOR_EXTRA", "1") != "0":
            # inside the V-cycle the finest-level pre-smoothing launch also carries the residual + full-weighting
            # restriction (34 B/cell) and the post-smoothing launch the residual norms of the convergence test (40 B/cell)
            # on top of 3 sweeps (120 B/cell) each: (154 + 160) / 2 = 157 B/cell per launch on average
            # -- and, with the streaming smoother, the bilinear prolongation + correction of the cycle (18 B/cell, fused
            # into the post-smoother's load): (154 + 178) / 2 = 166 B/cell per launch on average (SURVEY 8d figures)
            fused_prolong = os.environ.get("NF_MG_PROLONG_FUSED", "1") != "0" and os.environ.get("NF_RBSOR_STREAM") is None
            alg_bytes = (166.0 if fused_prolong else 157.0) * n * n
            tr = [variants.get(k, {}).get("dram_bytes_per_launch") for k in ("pre_restrict", "post_norms")]
            traffic = (tr[0] + tr[1]) / 2.0 if all(tr) else None
        achieved = alg_bytes / (ms_launch * 1e-3) / 1e9
    roofline = {"kernel": "k_rbsor_stream<3> (finest level: 3 red-black SOR sweeps = 6 colour passes per launch, streaming "
                          "wavefront form; the pre-smoothing launch also carries the V-cycle's residual + restriction, the "
                          "post-smoothing launch the convergence-test norms and the prolongation + correction)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_launch,
                "launches_timed": int(live_launches) if live else reps,
                "timing": "CUDA events around every finest-level launch of 2 further outer iterations of the same run"
                          if live else isolated["how"],
                "traffic": traffic,
                "dram_frac": (traffic / (ms_launch * 1e-3) / 1e9 / peak) if traffic else None,
                "traffic_source": prof.get("source"),
                "note": "frac counts SURVEY 8d's algorithmic bytes (3 temporally blocked sweeps count 3 x 40 B/cell, so it "
                        "can exceed 1); dram_frac = ncu-measured DRAM bytes / live time / peak is the physical utilisation",
                "isolated": isolated,
                "unfused_color_pass": {"ms_per_launch": ms_color, "achieved": 20.0 * n * n / (ms_color * 1e-3) / 1e9}}

    n = n_alg
    del roof, scratch, scratch2, inv
    # ---- end to end through the public API (host arrays in, host arrays out) --------------------------
    e2e = None
    if not args.no_e2e:
        ksteps = max(1, min(args.steps, 5))
        alg.solve(max_iterations=1, tolerance=0.0, save_profile=False, gather=False)  # warm the path (pinned buffers, first-touch)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            alg.solve(max_iterations=1, tolerance=0.0, save_profile=False, gather=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        fb = 8.0 * n * (n + 1)
        e2e = {"value": cells * ksteps / dt / 1e6, "unit": "MLUPS", "steps": ksteps,
               "h2d_bytes_per_step": int(2 * fb + 8.0 * n * n), "d2h_bytes_per_step": int(2 * fb + 8.0 * n * n),
               "note": "whole-job bytes; with N ranks each rank moves its own row slab (+8 halo rows up)"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = args.cpu_sample_n if args.cpu_sample_n > 0 else 513
        step = cpu_oracle_step_fn(ns, args)
        step()
        t0 = time.perf_counter()
        k = 0
        while k < 8 and (time.perf_counter() - t0) < 15.0:
            step(); k += 1
        dt = time.perf_counter() - t0
        cpu = {"value": ns * ns * k / dt / 1e6, "unit": "MLUPS", "cores": 1, "kind": "port",
               "sample": f"{ns}x{ns} grid of the same workload, {k} outer iterations after 1 warm-up, NumPy oracle port "
                         f"(single threaded; host has {os.cpu_count()} cores)"}
        if n >= 4097 and not args.no_cpu_kernels:
            # the port's pressure kernels at the config's own size (one application each, single thread)
            nk = 4097
            rk = np.random.default_rng(1)
            dxk = 1.0 / (nk - 1)
            duk = (0.7 * dxk / 4e-3) * (1 + 0.1 * rk.random((nk + 1, nk)))
            dvk = (0.7 * dxk / 4e-3) * (1 + 0.1 * rk.random((nk, nk + 1)))
            usk = 1e-2 * rk.standard_normal((nk + 1, nk)); usk[0, :] = usk[nk, :] = 0.0
            vsk = 1e-2 * rk.standard_normal((nk, nk + 1)); vsk[:, 0] = vsk[:, nk] = 0.0
            kt = cpu_baseline_pressure_kernels(nk, dxk, dxk, duk, dvk, usk, vsk, with_mg=True, with_lex=False)
            cpu["kernels_4097"] = {key: round(val, 1) for key, val in kt.items()}
            cpu["kernels_4097"]["note"] = ("oracle port, one call each on a seeded 4097^2 system: A*p, one Jacobi iteration, "
                                           "one red-black SOR sweep, one V(3,3) cycle; milliseconds, single thread")
            del duk, dvk, usk, vsk

    pressure = None
    if phases["iterations"] > 0:
        pc = float(np.mean([r["pressure_iterations"] for r in precs]))
        p_ms = phases["pressure_ms"] / phases["iterations"]
        # V(3,3) cycle = [(3+3)*40 + 40 + 10 + 18] * 4/3 = 411 B/fine cell; per solve the coefficient
        # hierarchy (27 B/cell) and the continuity RHS (24 B/cell) once
        p_bytes = (411.0 * pc + 27.0 + 24.0) * cells
        pressure = {"ms_per_step": p_ms, "cycles_per_step": pc, "ms_per_cycle": p_ms / max(pc, 1e-9),
                    "algorithmic_bytes_per_step": p_bytes, "achieved_gbs": p_bytes / (p_ms * 1e-3) / 1e9 if p_ms > 0 else None,
                    "frac_of_measured_hbm_peak": (p_bytes / (p_ms * 1e-3) / 1e9 / (peak * world)) if p_ms > 0 else None,
                    "momentum_ms_per_step": phases["momentum_ms"] / phases["iterations"],
                    "corrections_ms_per_step": phases["correct_ms"] / phases["iterations"],
                    "how": f"CUDA events between the phases of {phases['iterations']} further outer iterations of the same "
                           f"run (rank 0's clock); algorithmic bytes per SURVEY.md 8d: 411 B/cell per V(3,3) cycle + 51 B/cell "
                           f"per solve; peak = measured HBM copy bandwidth x {world} GPU(s)"}
    if rank == 0:
        cycles = [r["pressure_iterations"] for r in recs]
        line = {
            "metric": "simple_outer_mlups", "value": mlups, "unit": "MLUPS", "iter_per_s": args.steps / (ms * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n, world),
            "transport": None if world == 1 else
                         ("halos (8 rows) and norm reductions by the ranks' own kernels over NVLink peer memory (nf_p2p.cu)"
                          if alg.uses_p2p() else "halos and norm reductions by NCCL send/recv + allreduce")
                         + ", coarse multigrid levels replicated",
            "gpu_launches": int(launches), "mg_cycles_per_step": float(np.mean(cycles)) if cycles else None,
            "final_u_rel_norm": recs[-1]["u_rel_norm"] if recs else None,
            "clocks": clocks, "roofline": roofline, "pressure_solve": pressure, "e2e": e2e, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
