#!/usr/bin/env python
"""Per-kernel timings (CUDA events, legacy default stream) at one grid size: tools/bench_kernels.py [n]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from naviflow_b200._lib import NfLinks  # noqa: E402
from naviflow_b200.device import get_context, ptr  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
    ctx = get_context(0)
    lib = ctx.lib
    g = ctx.grid(n, n, 1.0 / (n - 1), 1.0 / (n - 1), 1.0)
    rng = np.random.default_rng(0)
    mk = lambda scale=1.0: ctx.upload(scale * (1 + 0.1 * rng.random((n + 1, n + 1))), n, n)
    du, dv, b, p, tmp, x, y = mk(40.0 / n), mk(40.0 / n), mk(1e-3), mk(), mk(), mk(), mk()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = {}

    def timeit(name, fn, launches, alg_bytes_per_launch, reps=5):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / launches)
        gbs = alg_bytes_per_launch / (best * 1e-3) / 1e9
        out[name] = dict(ms_per_launch=round(best, 5), alg_GBs=round(gbs, 1), frac=round(gbs / peak, 3))
        print(f"{name:34s} {best*1e3:9.1f} us/launch  {gbs:8.1f} GB/s algorithmic  ({gbs/peak:5.2f} of measured peak)")

    cells = float(n) * n
    H = ctx.handle
    G = C.byref(g)
    timeit("rbsor colour pass (unfused)", lambda: ctx.check(lib.nf_rbsor_sweeps(H, G, ptr(p), ptr(b), ptr(du), ptr(dv), 1.5, 4)), 8, 20 * cells)
    inv = mk()
    ctx.check(lib.nf_pressure_inv_diag(H, G, ptr(du), ptr(dv), ptr(inv)))
    for ns in (1, 2, 3):
        timeit(f"rbsor fused NS={ns}", lambda ns=ns: ctx.check(lib.nf_rbsor_sweeps_fused(H, G, ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), None, 1.5, ns * 4)), 4, ns * 40 * cells)
        timeit(f"rbsor fused NS={ns} +inv", lambda ns=ns: ctx.check(lib.nf_rbsor_sweeps_fused(H, G, ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), ptr(inv), 1.5, ns * 4)), 4, ns * 40 * cells)
    timeit("A*p", lambda: ctx.check(lib.nf_pressure_apply(H, G, ptr(p), ptr(du), ptr(dv), ptr(tmp))), 1, 32 * cells)
    timeit("b - A*p", lambda: ctx.check(lib.nf_pressure_residual(H, G, ptr(p), ptr(b), ptr(du), ptr(dv), ptr(tmp))), 1, 40 * cells)
    timeit("jacobi pressure iteration", lambda: ctx.check(lib.nf_jacobi_iterate(H, G, ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), 0.8, 4)), 4, 40 * cells)
    arrs = [mk() for _ in range(6)]
    links = NfLinks(*[a.data_ptr() for a in arrs])
    timeit("momentum links u", lambda: ctx.check(lib.nf_momentum_links_u(H, G, ptr(x), ptr(y), ptr(p), 1e-3, 0.7, 15, links, ptr(tmp))), 1, 80 * cells)
    timeit("momentum jacobi sweep", lambda: ctx.check(lib.nf_momentum_jacobi(H, G, 1, links, ptr(x), ptr(tmp), 4)), 4, 72 * cells)
    val = C.c_double()
    timeit("norm2", lambda: ctx.check(lib.nf_norm2(H, G, ptr(x), 0, C.byref(val))), 1, 8 * cells)
    nc = (n - 1) // 2
    gc = ctx.grid(nc, nc, 1.0 / (nc - 1), 1.0 / (nc - 1), 1.0)
    c = ctx.empty(nc, nc)
    timeit("restrict FW", lambda: ctx.check(lib.nf_restrict_fw(H, G, ptr(x), C.byref(gc), ptr(c))), 1, 10 * cells)
    timeit("prolong linear add", lambda: ctx.check(lib.nf_prolong_linear(H, C.byref(gc), ptr(c), G, ptr(x), 1)), 1, 18 * cells)
    print(json.dumps({"n": n, "peak_GBs": peak, "kernels": out}))


if __name__ == "__main__":
    main()
