#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/bench_exchange.py: time of one halo exchange / one scalar all-reduce of the slab
team, peer-memory kernels (nf_p2p.cu) against NCCL send/recv, over the level sizes of a 16385^2 hierarchy."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def make_team(ctx, world, rank):
    ident = (C.c_ubyte * 128)()
    if rank == 0:
        ctx.check(ctx.lib.nf_nccl_unique_id(ctx.handle, ident), "nf_nccl_unique_id")
    t = torch.tensor(list(ident), dtype=torch.uint8, device=f"cuda:{ctx.device}")
    dist.broadcast(t, src=0)
    ident = (C.c_ubyte * 128)(*t.cpu().tolist())
    team = C.c_void_p()
    ctx.check(ctx.lib.nf_team_create_nccl(ctx.handle, world, rank, ident, C.byref(team)), "nf_team_create_nccl")
    return team


def main():
    from naviflow_b200.device import get_context
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = get_context(local)
    sizes = [int(a) for a in sys.argv[1:]] or [16385, 8192, 4095, 2047, 1023]
    out = []
    for transport in ("p2p", "nccl"):
        os.environ["NF_P2P"] = "1" if transport == "p2p" else "0"
        team = make_team(ctx, world, rank)
        for n in sizes:
            for depth in (6, 8):
                ex, ar = C.c_double(), C.c_double()
                ctx.check(ctx.lib.nf_team_benchmark(team, n, n, depth, 4, 200, C.byref(ex), C.byref(ar)), "nf_team_benchmark")
                out.append(dict(transport=transport, uses_p2p=bool(ctx.lib.nf_team_uses_p2p(team)), n=n, depth=depth,
                                halo_kb=depth * ((n + 16) // 16 * 16) * 8 / 1024, us_exchange=ex.value * 1e3,
                                us_allreduce=ar.value * 1e3))
        ctx.lib.nf_team_free(team)
    if rank == 0:
        for r in out:
            print(json.dumps(r), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
