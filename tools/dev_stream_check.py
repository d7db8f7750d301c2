#!/usr/bin/env python
"""Development check of the streaming smoother: bit-identity against the colour-pass kernel and timings.
tools/dev_stream_check.py [sizes...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from naviflow_b200.device import get_context, ptr  # noqa: E402


def run(n, sweeps_list=(1, 2, 3, 4, 7), time_it=False):
    ctx = get_context(0)
    lib = ctx.lib
    g = ctx.grid(n, n, 1.0 / (n - 1), 1.0 / (n - 1), 1.0)
    rng = np.random.default_rng(n)
    mk = lambda scale=1.0: ctx.upload(scale * (1 + 0.1 * rng.random((n + 1, n + 1))), n, n)
    du, dv, b, x = mk(40.0 / n), mk(40.0 / n), mk(1e-3), mk()
    inv = mk()
    H, G = ctx.handle, C.byref(g)
    ctx.check(lib.nf_pressure_inv_diag(H, G, ptr(du), ptr(dv), ptr(inv)))
    ok = True
    for sweeps in sweeps_list:
        p_ref = x.clone()
        ctx.check(lib.nf_rbsor_sweeps(H, G, ptr(p_ref), ptr(b), ptr(du), ptr(dv), 1.5, sweeps))
        for mode in ("stream", "tma"):
            os.environ["NF_RBSOR_STREAM"] = "0" if mode == "stream" else "1000000000"
            p, tmp = x.clone(), torch.zeros_like(x)
            ctx.check(lib.nf_rbsor_sweeps_fused(H, G, ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), ptr(inv), 1.5, sweeps))
            torch.cuda.synchronize()
            a, r = ctx.download(p, n, n), ctx.download(p_ref, n, n)
            same = np.array_equal(a, r)
            if not same:
                bad = np.argwhere(a != r)
                print(f"n={n} sweeps={sweeps} {mode}: MISMATCH at {len(bad)} cells, first {bad[:5].tolist()}, "
                      f"max diff {np.nanmax(np.abs(a - r)):.3e}")
                ok = False
    print(f"n={n}: {'bit-identical' if ok else 'FAILED'}")
    if time_it:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p, tmp = x.clone(), torch.zeros_like(x)
        for mode, wpc in (("tma", 0), ("stream", 8)):
            os.environ["NF_RBSOR_STREAM"] = "0" if mode == "stream" else "1000000000"
            os.environ["NF_STREAM_WPC"] = str(wpc)
            fn = lambda: ctx.check(lib.nf_rbsor_sweeps_fused(H, G, ptr(p), ptr(tmp), ptr(b), ptr(du), ptr(dv), ptr(inv), 1.5, 12))
            fn(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 4)
            print(f"n={n} {mode} wpc={wpc}: {best*1e3:.1f} us per 3-sweep launch "
                  f"({48 * n * n / best / 1e6:.0f} GB/s compulsory DRAM, {120 * n * n / best / 1e6:.0f} GB/s algorithmic)")
    return ok


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [64, 65, 127, 130, 257, 600, 1025]
    allok = True
    for n in sizes:
        allok &= run(n, time_it=n >= 1000)
    sys.exit(0 if allok else 1)
