#!/usr/bin/env python
"""A few launches of the smoother at one size, for ncu: tools/dev_stream_prof.py n sweeps mode [wpc]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from naviflow_b200.device import get_context, ptr  # noqa: E402

n, sweeps, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
os.environ["NF_RBSOR_STREAM"] = "0" if mode == "stream" else "1000000000"
if len(sys.argv) > 4:
    os.environ["NF_STREAM_WPC"] = sys.argv[4]
ctx = get_context(0)
lib = ctx.lib
g = ctx.grid(n, n, 1.0 / (n - 1), 1.0 / (n - 1), 1.0)
rng = np.random.default_rng(n)
mk = lambda scale=1.0: ctx.upload(scale * (1 + 0.1 * rng.random((n + 1, n + 1))), n, n)
du, dv, b, x = mk(40.0 / n), mk(40.0 / n), mk(1e-3), mk()
inv, tmp = mk(), mk()
H, G = ctx.handle, C.byref(g)
ctx.check(lib.nf_pressure_inv_diag(H, G, ptr(du), ptr(dv), ptr(inv)))
for _ in range(3):
    ctx.check(lib.nf_rbsor_sweeps_fused(H, G, ptr(x), ptr(tmp), ptr(b), ptr(du), ptr(dv), ptr(inv), 1.5, sweeps))
torch.cuda.synchronize()
print("ok")
