"""QUICK / second-order upwind links: the oracle restatement and the HOST build of the CUDA kernel's
per-cell function (naviflow_b200/csrc/nf_links_ext.cuh, compiled here with g++) against the outputs of the reference's
QUICKDiscretization / SecondOrderUpwindDiscretization (tests/golden/ext_links_kats.npz, oracle/make_golden.py)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import np_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = O.EXT_KEYS


def cases(G):
    for tag in G["cases"]:
        tag = str(tag)
        nx, ny = (int(x) for x in G[f"{tag}_dims"])
        dx, dy, rho, mu = (float(x) for x in G[f"{tag}_scal"])
        yield tag, nx, ny, dx, dy, rho, mu, G[f"{tag}_u"], G[f"{tag}_v"], G[f"{tag}_p"]


def test_oracle_ext_links_equal_the_reference(golden_dir):
    G = np.load(os.path.join(golden_dir, "ext_links_kats.npz"))
    n = 0
    for tag, nx, ny, dx, dy, rho, mu, u, v, p in cases(G):
        for sch in ("quick", "sou"):
            for bname, sides in (("bc", 15), ("nobc", 0)):
                for comp in ("u", "v"):
                    got = O.ext_links(sch, comp == "u", nx, ny, dx, dy, rho, mu, u, v, p, sides)
                    for k in KEYS:
                        np.testing.assert_array_equal(got[k], G[f"{tag}_{sch}_{bname}_{comp}_{k}"],
                                                      err_msg=f"{tag} {sch} {bname} {comp} {k}")
                        n += 1
    assert n == 400


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("hostcc") / "libhost_links_ext.so")
    src = os.path.join(HERE, "hostcc", "host_links_ext.cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.host_links_ext.restype = None
    lib.host_links_ext.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_int] + [C.c_void_p] * 4
    return lib


def test_device_cell_function_built_for_the_host_equals_the_reference(golden_dir, hostlib):
    G = np.load(os.path.join(golden_dir, "ext_links_kats.npz"))
    for tag, nx, ny, dx, dy, rho, mu, u, v, p in cases(G):
        ld = ((ny + 1 + 15) // 16) * 16

        def pitched(a):
            b = np.full((nx + 1, ld), np.nan)     # NaN padding: a read outside the arrays' logical shape would show
            b[:a.shape[0], :a.shape[1]] = a
            return np.ascontiguousarray(b)

        U, V, P = pitched(u), pitched(v), pitched(p)
        for si, sch in ((1, "quick"), (2, "sou")):
            for bname, sides in (("bc", 15), ("nobc", 0)):
                for is_u, comp in ((1, "u"), (0, "v")):
                    out = np.zeros((10, nx + 1, ld))
                    hostlib.host_links_ext(si, is_u, nx, ny, ld, dx, dy, rho, mu, sides, U.ctypes.data, V.ctypes.data,
                                           P.ctypes.data, out.ctypes.data)
                    shape = (nx + 1, ny) if is_u else (nx, ny + 1)
                    for q, k in enumerate(KEYS):
                        np.testing.assert_array_equal(out[q, :shape[0], :shape[1]], G[f"{tag}_{sch}_{bname}_{comp}_{k}"],
                                                      err_msg=f"{tag} {sch} {bname} {comp} {k}")
