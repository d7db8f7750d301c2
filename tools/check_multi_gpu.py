#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_multi_gpu.py [n]: the slab run over N GPUs equals the single-GPU run bit for
bit, with the exchanges as peer-memory kernels (default) and as NCCL send/recv (NF_P2P=0); SIMPLE and PISO; prints the
time per outer iteration of both transports."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def run(n, distributed, virtual_ranks=1, iters=3, piso=0, timed=0, solver="mg"):
    import time
    import naviflow_b200 as nb
    mesh = nb.StructuredMesh(n, n, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=1000, characteristic_velocity=1.0)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=3, tolerance=1e-30,
                               pre_smoothing=3, post_smoothing=3)
    if solver == "rbsor":
        ps = nb.GpuGaussSeidelSolver(tolerance=0.0, max_iterations=12, omega=1.5)
    elif solver == "jacobi":
        ps = nb.GpuJacobiSolver(tolerance=0.0, max_iterations=20, omega=0.8)
    elif solver == "bicgstab":
        ps = nb.GpuBiCGSTABSolver(tolerance=1e-30, max_iterations=8)
    elif solver == "cg":
        ps = nb.GpuCGSolver(tolerance=1e-30, max_iterations=8)
    cls, kw = (nb.GpuPisoSolver, dict(n_corrections=piso)) if piso else (nb.GpuSimpleSolver, {})
    alg = cls(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=5), distributed=distributed,
              virtual_ranks=virtual_ranks, **kw)
    alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        alg.set_boundary_condition(b, "wall")
    res = alg.solve(max_iterations=iters, tolerance=0.0, save_profile=False)
    alg.ms_per_iteration = None
    if timed:
        alg.iterate_resident(2)
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        t0 = time.perf_counter()
        alg.iterate_resident(timed)
        torch.cuda.synchronize()
        alg.ms_per_iteration = (time.perf_counter() - t0) * 1e3 / timed
    return alg, res


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    ref, rres = run(n, False)
    ok_all = True
    for label, env, piso in (("p2p", "1", 0), ("nccl", "0", 0), ("p2p-piso2", "1", 2)):
        os.environ["NF_P2P"] = env
        alg, res = run(n, True, piso=piso, timed=0 if piso else 10)
        if piso:
            ref, rres = run(n, False, piso=piso)
        rows = alg.local_rows()
        ok = all(np.array_equal(getattr(alg, f), getattr(ref, f)) for f in ("u", "v", "p"))
        hist_ok = np.allclose(res.get_history("total_rel_norm"), rres.get_history("total_rel_norm"), rtol=1e-12)
        ok_all = ok_all and ok and hist_ok and (alg.uses_p2p() == (env == "1"))
        print(f"[{label}] rank {rank}/{world} rows {rows} uses_p2p={alg.uses_p2p()} fields_bit_identical={ok} "
              f"history_close={hist_ok} ms_per_iteration={alg.ms_per_iteration}", flush=True)
        del alg
    # the other pressure solvers on slabs: stationary ones bit-identical, Krylov to rounding (dot-product order)
    for solver in ("rbsor", "jacobi", "cg", "bicgstab"):
        os.environ["NF_P2P"] = "1"
        ref, _ = run(n, False, solver=solver, iters=2)
        alg, _ = run(n, True, solver=solver, iters=2)
        if solver in ("rbsor", "jacobi"):
            ok = all(np.array_equal(getattr(alg, f), getattr(ref, f)) for f in ("u", "v", "p"))
            err = 0.0
        else:
            err = max(np.linalg.norm(getattr(alg, f) - getattr(ref, f)) / np.linalg.norm(getattr(ref, f)) for f in ("u", "v", "p"))
            ok = err < 1e-11
        ok_all = ok_all and ok
        print(f"[p2p-{solver}] rank {rank}/{world} ok={ok} max_rel_err={err:.2e}", flush=True)
        del alg
    ok = hist_ok = ok_all
    flag = torch.tensor([int(ok and hist_ok)], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
