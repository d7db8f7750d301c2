#!/usr/bin/env python
"""Development check: multigrid solves with the streaming smoother (all fused modes) against the tiled kernels.
tools/dev_stream_mg.py [sizes...]   -- bit-identity of x, cycle counts, norms; timing per cycle at the large sizes"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from naviflow_b200._lib import NfMgConfig, NfMgInfo  # noqa: E402
from naviflow_b200.device import get_context, pad_ld, ptr  # noqa: E402


def solve(ctx, n, du, dv, b, stream, cycles, tol, time_it=False):
    os.environ["NF_RBSOR_STREAM"] = os.environ.get("STREAM_MIN", "0") if stream else "1000000000"
    os.environ["NF_MG_TAIL"] = "1" if stream else "0"   # the new paths together against the round-1 kernels
    lib = ctx.lib
    c = NfMgConfig()
    c.smoother, c.pre, c.post, c.cycle_type, c.cycle_buildup, c.cycle_final = 0, 3, 3, 0, 0, -1
    c.max_cycles_buildup, c.restriction, c.interpolation, c.coarsest = 1, 0, 0, 7
    c.max_iterations, c.omega, c.tolerance = cycles, 1.5, tol
    c.length, c.height, c.rho = 1.0, 1.0, 1.0
    h = C.c_void_p()
    ctx.check(lib.nf_mg_create(ctx.handle, C.byref(h), n, n, pad_ld(n), C.byref(c)), "create")
    ctx.check(lib.nf_mg_setup(h, ptr(du), ptr(dv)), "setup")
    x, r = ctx.empty(n, n), ctx.empty(n, n)
    info = NfMgInfo()
    ctx.check(lib.nf_mg_solve(h, ptr(b), ptr(x), ptr(r), C.byref(info)), "solve")
    torch.cuda.synchronize()
    ms = None
    if time_it:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record()
            ctx.check(lib.nf_mg_solve(h, ptr(b), ptr(x), ptr(r), C.byref(info)), "solve")
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ms = best / max(info.cycles, 1)
    out = ctx.download(x, n, n), ctx.download(r, n, n), info.cycles, info.r_norm, info.b_norm, ms
    lib.nf_mg_destroy(h)
    return out


def run(n, time_it):
    ctx = get_context(0)
    rng = np.random.default_rng(n)
    dx = 1.0 / (n - 1)
    du_h = (0.7 * dx / 4e-3) * (1 + 0.1 * rng.random((n + 1, n)))
    dv_h = (0.7 * dx / 4e-3) * (1 + 0.1 * rng.random((n, n + 1)))
    du_h[0, :] = np.nan; du_h[n, :] = np.nan; dv_h[:, 0] = np.nan; dv_h[:, n] = np.nan  # as the momentum solver leaves them
    b_h = 1e-2 * rng.standard_normal((n, n)); b_h[0, 0] = 0.0
    du, dv, b = ctx.upload(du_h, n, n), ctx.upload(dv_h, n, n), ctx.upload(b_h, n, n)
    ok = True
    for cycles, tol in ((2, 0.0), (30, 1e-3)):
        xa, ra, ca, rna, bna, msa = solve(ctx, n, du, dv, b, False, cycles, tol, time_it and tol > 0)
        xb, rb, cb, rnb, bnb, msb = solve(ctx, n, du, dv, b, True, cycles, tol, time_it and tol > 0)
        same = np.array_equal(xa, xb) and np.array_equal(ra, rb)
        nrm = abs(rna - rnb) <= 1e-12 * abs(rna) and abs(bna - bnb) <= 1e-12 * abs(bna)
        print(f"n={n} cycles<={cycles} tol={tol}: x/r {'bit-identical' if same else 'DIFFER'} cycles {ca}/{cb} "
              f"r_norm {rna:.15e}/{rnb:.15e} {'ok' if nrm else 'NORM MISMATCH'}"
              + (f"  ms/cycle tiled {msa:.3f} stream {msb:.3f}" if msa else ""))
        if not same:
            bad = np.argwhere(xa != xb)
            print("   first mismatches", bad[:6].tolist(), "max diff", np.nanmax(np.abs(xa - xb)))
        ok &= same and nrm and ca == cb
    return ok


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [63, 127, 130, 257, 600, 1025, 2049]
    allok = True
    for n in sizes:
        allok &= run(n, n >= 2000)
    sys.exit(0 if allok else 1)
