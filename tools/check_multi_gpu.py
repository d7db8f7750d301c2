#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_multi_gpu.py [n]: the NCCL slab run equals the single-GPU run bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def run(n, distributed, virtual_ranks=1, iters=3):
    import naviflow_b200 as nb
    mesh = nb.StructuredMesh(n, n, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=1000, characteristic_velocity=1.0)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=3, tolerance=1e-30,
                               pre_smoothing=3, post_smoothing=3)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=5), distributed=distributed,
                             virtual_ranks=virtual_ranks)
    alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        alg.set_boundary_condition(b, "wall")
    res = alg.solve(max_iterations=iters, tolerance=0.0)
    return alg, res


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    alg, res = run(n, True)
    rows = alg.local_rows()
    ref, rres = run(n, False)
    ok = all(np.array_equal(getattr(alg, f), getattr(ref, f)) for f in ("u", "v", "p"))
    hist_ok = np.allclose(res.get_history("total_rel_norm"), rres.get_history("total_rel_norm"), rtol=1e-12)
    print(f"rank {rank}/{world} rows {rows} fields_bit_identical={ok} history_close={hist_ok}", flush=True)
    flag = torch.tensor([int(ok and hist_ok)], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
