#!/usr/bin/env python
"""torchrun --nproc-per-node 2 tools/dev_nccl_bisect.py [n]: time per outer iteration of the slab run with the NCCL transport
under the environment switches of the round-2 kernels (development aid)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def run(n, timed=5):
    import naviflow_b200 as nb
    mesh = nb.StructuredMesh(n, n, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=1000, characteristic_velocity=1.0)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=3, tolerance=1e-30,
                               pre_smoothing=3, post_smoothing=3)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=5), distributed=True)
    alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        alg.set_boundary_condition(b, "wall")
    alg.push_fields()
    alg.iterate_resident(3)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    alg.iterate_resident(timed)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / timed
    p2p = alg.uses_p2p()
    alg.close()
    return ms, p2p


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2049
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cases = [("p2p default", {"NF_P2P": "1"}),
             ("nccl default", {"NF_P2P": "0"}),
             ("nccl no-graph", {"NF_P2P": "0", "NF_MG_GRAPH": "0"}),
             ("nccl no-stream", {"NF_P2P": "0", "NF_RBSOR_STREAM": "1000000000"}),
             ("nccl no-tail", {"NF_P2P": "0", "NF_MG_TAIL": "0"}),
             ("nccl round-1 kernels", {"NF_P2P": "0", "NF_MG_TAIL": "0", "NF_RBSOR_STREAM": "1000000000"})]
    for label, env in cases:
        for k in ("NF_P2P", "NF_MG_GRAPH", "NF_RBSOR_STREAM", "NF_MG_TAIL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ms, p2p = run(n)
        if dist.get_rank() == 0:
            print(f"{label:24s} uses_p2p={p2p} {ms:9.3f} ms/iteration", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
