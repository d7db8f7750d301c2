#!/usr/bin/env python
"""python tools/virtual_slab_run.py n ranks [iterations]: the bench workload cut into `ranks` virtual slabs on ONE GPU
(same slab code path as the multi-GPU run, cudaMemcpy halos) -- used to chase slab bugs under compute-sanitizer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n, ranks = int(sys.argv[1]), int(sys.argv[2])
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    import naviflow_b200 as nb
    mesh = nb.StructuredMesh(n, n, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=1000, characteristic_velocity=1.0)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=100, tolerance=1e-3,
                               pre_smoothing=3, post_smoothing=3)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=5), virtual_ranks=ranks)
    alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for b in ("bottom", "left", "right"):
        alg.set_boundary_condition(b, "wall")
    alg.push_fields()
    recs = alg.iterate_resident(iters, 0.0)
    print("ok", n, ranks, [(r["pressure_iterations"], r["u_rel_norm"]) for r in recs], flush=True)


if __name__ == "__main__":
    main()
