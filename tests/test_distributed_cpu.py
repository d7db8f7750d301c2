"""CPU tests of the N > 1 host-side logic with a world_size-2 gloo process group: the slab partition rule (queried
from the library without a GPU), the row assembly used by GpuSimpleSolver.solve(gather=True), and bench.py's
reference arm under a multi-rank launch (rank 0 alone prints)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nx, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from naviflow_b200.simple import assemble_rows, slab_rows
    b, e = slab_rows(nx, world, rank)
    rows = [None] * world
    dist.all_gather_object(rows, (b, e))
    # every rank derives the same partition for every rank
    assert rows == [slab_rows(nx, world, r) for r in range(world)]
    # a field whose rows are known only to their owner is assembled on every rank
    full = np.arange((nx + 1) * 5, dtype=np.float64).reshape(nx + 1, 5)
    mine = np.full_like(full, -7.0)
    hi = nx + 1 if rank == world - 1 else e
    mine[b:hi] = full[b:hi]
    got = assemble_rows(mine, b, hi, device="cpu")
    np.save(os.path.join(out_dir, f"rows_{rank}.npy"), np.array(rows))
    assert np.array_equal(got, full)
    dist.destroy_process_group()


@pytest.mark.parametrize("nx", [257, 4097, 16385])
def test_slab_partition_and_assembly_world2_gloo(nx, tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + (nx % 97)
    mp.spawn(_worker, args=(2, port, nx, str(tmp_path)), nprocs=2, join=True)
    rows = np.load(tmp_path / "rows_0.npy")
    assert rows[0][0] == 0 and rows[-1][1] == nx
    assert rows[0][1] == rows[1][0] and rows[1][0] % 16 == 0          # contiguous, aligned boundary
    assert min(r[1] - r[0] for r in rows) >= 64


def test_partition_rule_properties():
    sys.path.insert(0, ROOT)
    from naviflow_b200.simple import slab_rows
    from naviflow_b200 import _lib
    import ctypes as C
    for nx, world in ((4097, 8), (16385, 8), (1025, 4), (513, 3), (257, 4)):
        parts = [slab_rows(nx, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == nx
        for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
            assert e0 == b1 and b1 % 16 == 0
        # induced ownership on the next-coarser level: coarse row I goes to the owner of fine row 2I+1
        nxc = (nx - 1) // 2
        lib = _lib.lib()
        covered = []
        for r, (b, e) in enumerate(parts):
            cb, ce = C.c_int(), C.c_int()
            nb = parts[r + 1][0] if r + 1 < world else nx
            lib.nf_slab_coarse_rows(b, nb, 1 if r == world - 1 else 0, nxc, C.byref(cb), C.byref(ce))
            for I in range(cb.value, ce.value):
                assert b <= 2 * I + 1 < e
            covered.append((cb.value, ce.value))
        assert covered[0][0] == 0 and covered[-1][1] == nxc
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    # too few rows per rank: the grid is not cut
    assert slab_rows(100, 4, 2) == (0, 100)


def test_bench_reference_arm_under_two_ranks():
    """`bench.py --impl reference` launched with WORLD_SIZE=2: rank 0 prints the JSON line, rank 1 exits 0 silently."""
    outs = []
    for rank in (0, 1):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29650")
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "1", "--warmup", "0", "--cpu-sample-n", "65"], env=env, capture_output=True,
                           text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        outs.append(p.stdout.strip())
    line = json.loads(outs[0].splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MLUPS" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert outs[1] == ""
