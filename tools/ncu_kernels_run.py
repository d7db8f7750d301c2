#!/usr/bin/env python
"""Launches every kernel of the library once or twice at one grid size, for an ncu pass:
  ncu --metrics <list in tools/summarize_ncu.py> ... python tools/ncu_kernels_run.py 4097
(pressure kernels and Krylov iterations on a seeded system, two SIMPLE outer iterations with the multigrid pressure solve,
one lexicographic Gauss-Seidel sweep at min(n, 2049))."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import naviflow_b200 as nb  # noqa: E402
from naviflow_b200._lib import NfKrylovInfo, NfLinks  # noqa: E402
from naviflow_b200.device import get_context, pad_ld, ptr  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
    ctx = get_context(0)
    lib, H = ctx.lib, ctx.handle
    g = ctx.grid(n, n, 1.0 / (n - 1), 1.0 / (n - 1), 1.0)
    G = C.byref(g)
    rng = np.random.default_rng(0)
    mk = lambda scale=1.0: ctx.upload(scale * (1 + 0.1 * rng.random((n + 1, n + 1))), n, n)
    du, dv, b, p, tmp, x, y = mk(40.0 / n), mk(40.0 / n), mk(1e-3), mk(), mk(), mk(), mk()
    inv = mk()
    ck = ctx.check
    ck(lib.nf_pressure_inv_diag(H, G, ptr(du), ptr(dv), ptr(inv)))
    ck(lib.nf_rbsor_sweeps(H, G, ptr(p), ptr(b), ptr(du), ptr(dv), 1.5, 1))
    ck(lib.nf_rbsor_sweeps_fused(H, G, ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), ptr(inv), 1.5, 3))
    ck(lib.nf_pressure_apply(H, G, ptr(p), ptr(du), ptr(dv), ptr(tmp)))
    ck(lib.nf_pressure_residual(H, G, ptr(p), ptr(b), ptr(du), ptr(dv), ptr(tmp)))
    ck(lib.nf_jacobi_iterate(H, G, ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), 0.8, 2))
    ck(lib.nf_continuity_rhs(H, G, ptr(x), ptr(y), ptr(tmp)))
    val = C.c_double()
    ck(lib.nf_norm2(H, G, ptr(x), 0, C.byref(val)))
    nc = (n - 1) // 2
    gc = ctx.grid(nc, nc, 1.0 / (nc - 1), 1.0 / (nc - 1), 1.0)
    c = ctx.empty(nc, nc)
    ck(lib.nf_restrict_fw(H, G, ptr(x), C.byref(gc), ptr(c)))
    ck(lib.nf_prolong_linear(H, C.byref(gc), ptr(c), G, ptr(x), 1))
    cdu, cdv = ctx.empty(nc, nc), ctx.empty(nc, nc)
    ck(lib.nf_restrict_coeffs(H, G, ptr(du), ptr(dv), C.byref(gc), ptr(cdu), ptr(cdv)))
    # Krylov: a few iterations each (maxiter small, the host polls once)
    info = NfKrylovInfo()
    work = torch.zeros((5 * (n + 1), pad_ld(n)), dtype=torch.float64, device=p.device)
    xk = ctx.empty(n, n)
    ck(lib.nf_cg_solve(H, G, ptr(b), ptr(xk), ptr(du), ptr(dv), 0.0, 1e-30, 3, 25, ptr(work), C.byref(info)))
    ck(lib.nf_bicgstab_solve(H, G, ptr(b), ptr(xk), ptr(du), ptr(dv), 0.0, 1e-30, 3, 10, ptr(work), C.byref(info)))
    del work
    m = min(n, 2049)
    gm = ctx.grid(m, m, 1.0 / (m - 1), 1.0 / (m - 1), 1.0)
    mkm = lambda scale=1.0: ctx.upload(scale * (1 + 0.1 * rng.random((m + 1, m + 1))), m, m)
    pm, bm, dum, dvm = mkm(), mkm(1e-3), mkm(40.0 / m), mkm(40.0 / m)
    ck(lib.nf_gs_lex_sweeps(H, C.byref(gm), ptr(pm), ptr(bm), ptr(dum), ptr(dvm), 1.8, 1, 0))
    torch.cuda.synchronize()
    del du, dv, b, p, tmp, x, y, inv
    # the outer loop: momentum links / sweeps, continuity RHS, multigrid cycle, corrections
    mesh = nb.StructuredMesh(n, n, 1.0, 1.0)
    fluid = nb.FluidProperties(density=1.0, reynolds_number=1000.0, characteristic_velocity=1.0)
    ps = nb.GpuMultiGridSolver(smoother=nb.GpuGaussSeidelSolver(omega=1.5), max_iterations=3, tolerance=1e-30,
                               pre_smoothing=3, post_smoothing=3)
    alg = nb.GpuSimpleSolver(mesh, fluid, ps, nb.GpuJacobiMomentumSolver(n_jacobi_sweeps=5), alpha_p=0.3, alpha_u=0.7,
                             track_unrelaxed_residual=True)
    alg.set_boundary_condition("top", "velocity", {"u": 1.0, "v": 0.0})
    for bnd in ("bottom", "left", "right"):
        alg.set_boundary_condition(bnd, "wall")
    os.environ["NF_MG_GRAPH"] = "0"   # kernel by kernel (ncu profiles graph nodes too, but keep the run simple)
    alg.push_fields()
    alg.iterate_resident(2, 0.0)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
