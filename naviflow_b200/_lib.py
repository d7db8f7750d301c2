"""ctypes binding of libnaviflow_b200.so (the C-ABI declared in include/naviflow_b200.h).

The library is the product: there is no CPU fallback.  Importing this module without the built
shared object raises ImportError with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnaviflow_b200.so")

NF_OK = 0


class NfError(RuntimeError):
    """Raised when a libnaviflow_b200 call returns a non-zero nf_status."""


class NfGrid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("ld", C.c_int32), ("row0", C.c_int32),
                ("gb", C.c_int32), ("ge", C.c_int32), ("row1", C.c_int32), ("pad", C.c_int32),
                ("dx", C.c_double), ("dy", C.c_double),
                ("rho", C.c_double)]


class NfBcProgram(C.Structure):
    _fields_ = [("u_edge", C.c_double * 4), ("u_corner", C.c_double * 4),
                ("v_edge", C.c_double * 4), ("v_corner", C.c_double * 4),
                ("v_right_row", C.c_int32), ("pad", C.c_int32)]


class NfMgConfig(C.Structure):
    _fields_ = [("smoother", C.c_int32), ("pre", C.c_int32), ("post", C.c_int32),
                ("cycle_type", C.c_int32), ("cycle_buildup", C.c_int32), ("cycle_final", C.c_int32),
                ("max_cycles_buildup", C.c_int32), ("restriction", C.c_int32),
                ("interpolation", C.c_int32), ("coarsest", C.c_int32), ("max_iterations", C.c_int32),
                ("pad", C.c_int32), ("omega", C.c_double), ("tolerance", C.c_double),
                ("length", C.c_double), ("height", C.c_double), ("rho", C.c_double)]


class NfMgInfo(C.Structure):
    _fields_ = [("r_norm", C.c_double), ("b_norm", C.c_double), ("cycles", C.c_int32),
                ("levels", C.c_int32)]


class NfKrylovInfo(C.Structure):
    _fields_ = [("r_norm", C.c_double), ("b_norm", C.c_double), ("iterations", C.c_int32),
                ("info", C.c_int32)]


class NfLinks(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("a_e", "a_w", "a_n", "a_s", "a_p", "src")]


class NfLinksExt(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("a_e", "a_w", "a_n", "a_s", "a_ee", "a_ww", "a_nn", "a_ss", "a_p", "src")]


class NfSimpleConfig(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("n_momentum_sweeps", C.c_int32),
                ("pressure_solver", C.c_int32), ("pressure_iterations", C.c_int32),
                ("sides", C.c_int32), ("krylov_maxiter", C.c_int32), ("piso_corrections", C.c_int32),
                ("length", C.c_double), ("height", C.c_double), ("rho", C.c_double), ("mu", C.c_double),
                ("alpha_p", C.c_double), ("alpha_u", C.c_double), ("pressure_omega", C.c_double),
                ("pressure_tolerance", C.c_double),
                ("bc", NfBcProgram), ("mg", NfMgConfig),
                ("momentum_solver", C.c_int32), ("momentum_maxiter", C.c_int32), ("momentum_tolerance", C.c_double),
                ("bc_mf", NfBcProgram),
                ("simplec_divisor", C.c_double), ("krylov_check_every", C.c_int32), ("krylov_mg_cycles", C.c_int32),
                ("krylov_mg_kind", C.c_int32), ("track_unrelaxed_residual", C.c_int32)]


class NfSimpleInfo(C.Structure):
    _fields_ = [("u_rel_norm", C.c_double), ("v_rel_norm", C.c_double), ("p_rel_norm", C.c_double),
                ("u_abs_res", C.c_double), ("v_abs_res", C.c_double),
                ("pressure_iterations", C.c_int32), ("pad", C.c_int32),
                ("u_unrelaxed_res", C.c_double), ("v_unrelaxed_res", C.c_double)]


P = C.c_void_p          # device pointer
GP = C.POINTER(NfGrid)
CTX = C.c_void_p
DBL_OUT = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/naviflow_b200.h declares
SIGNATURES = {
    "nf_ctx_create": (C.c_int, [C.POINTER(CTX), C.c_int, C.c_void_p]),
    "nf_ctx_destroy": (C.c_int, [CTX]),
    "nf_last_error": (C.c_char_p, [CTX]),
    "nf_sync": (C.c_int, [CTX]),
    "nf_version": (C.c_int, []),
    "nf_launch_count": (C.c_int64, [CTX]),
    "nf_apply_velocity_bc": (C.c_int, [CTX, GP, C.POINTER(NfBcProgram), P, P]),
    "nf_continuity_rhs": (C.c_int, [CTX, GP, P, P, P]),
    "nf_pressure_apply": (C.c_int, [CTX, GP, P, P, P, P]),
    "nf_pressure_residual": (C.c_int, [CTX, GP, P, P, P, P, P]),
    "nf_jacobi_iterate": (C.c_int, [CTX, GP, P, P, P, P, P, C.c_double, C.c_int]),
    "nf_jacobi_diag": (C.c_int, [CTX, GP, P, P, P]),
    "nf_rbsor_sweeps": (C.c_int, [CTX, GP, P, P, P, P, C.c_double, C.c_int]),
    "nf_gs_lex_sweeps": (C.c_int, [CTX, GP, P, P, P, P, C.c_double, C.c_int, C.c_int]),
    "nf_rbsor_sweeps_fused": (C.c_int, [CTX, GP, P, P, P, P, P, P, C.c_double, C.c_int]),
    "nf_pressure_inv_diag": (C.c_int, [CTX, GP, P, P, P]),
    "nf_restrict_fw": (C.c_int, [CTX, GP, P, GP, P]),
    "nf_restrict_inject": (C.c_int, [CTX, GP, P, GP, P]),
    "nf_restrict_coeffs": (C.c_int, [CTX, GP, P, P, GP, P, P]),
    "nf_prolong_linear": (C.c_int, [CTX, GP, P, GP, P, C.c_int]),
    "nf_prolong_cubic": (C.c_int, [CTX, GP, P, GP, P, C.c_int, C.c_void_p, C.c_size_t]),
    "nf_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "nf_norm2": (C.c_int, [CTX, GP, P, C.c_int, DBL_OUT]),
    "nf_dot": (C.c_int, [CTX, GP, P, P, DBL_OUT]),
    "nf_mg_create": (C.c_int, [CTX, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.POINTER(NfMgConfig)]),
    "nf_mg_destroy": (C.c_int, [C.c_void_p]),
    "nf_mg_num_levels": (C.c_int, [C.c_void_p]),
    "nf_mg_level_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nf_mg_level_array": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int]),
    "nf_mg_setup": (C.c_int, [C.c_void_p, P, P]),
    "nf_mg_solve": (C.c_int, [C.c_void_p, P, P, P, C.POINTER(NfMgInfo)]),
    "nf_mg_cycle": (C.c_int, [C.c_void_p, P, P, C.c_int]),
    "nf_cg_solve": (C.c_int, [CTX, GP, P, P, P, P, C.c_double, C.c_double, C.c_int, C.c_int, P,
                              C.POINTER(NfKrylovInfo)]),
    "nf_bicgstab_solve": (C.c_int, [CTX, GP, P, P, P, P, C.c_double, C.c_double, C.c_int, C.c_int, P,
                                    C.POINTER(NfKrylovInfo)]),
    "nf_bicgstab_solve_mg": (C.c_int, [CTX, GP, P, P, P, P, C.c_double, C.c_double, C.c_int, C.c_int, P, C.c_void_p,
                                       C.c_int, C.c_int, C.POINTER(NfKrylovInfo)]),
    "nf_cg_solve_mg": (C.c_int, [CTX, GP, P, P, P, P, C.c_double, C.c_double, C.c_int, P, C.c_void_p, C.c_int, C.c_int,
                                 C.POINTER(NfKrylovInfo)]),
    "nf_momentum_links_u": (C.c_int, [CTX, GP, P, P, P, C.c_double, C.c_double, C.c_int, NfLinks, P]),
    "nf_momentum_links_v": (C.c_int, [CTX, GP, P, P, P, C.c_double, C.c_double, C.c_int, NfLinks, P]),
    "nf_momentum_links_ext": (C.c_int, [CTX, GP, C.c_int, C.c_int, P, P, P, C.c_double, C.c_int, NfLinksExt]),
    "nf_momentum_jacobi": (C.c_int, [CTX, GP, C.c_int, NfLinks, P, P, C.c_int]),
    "nf_momentum_jacobi_fused": (C.c_int, [CTX, GP, C.c_int, NfLinks, P, P, C.c_int, P, DBL_OUT]),
    "nf_momentum_residual": (C.c_int, [CTX, GP, C.c_int, NfLinks, P, P, DBL_OUT]),
    "nf_momentum_links_mf": (C.c_int, [CTX, GP, C.c_int, P, P, P, C.c_double, C.c_double, C.c_int, NfLinks, P, P, P]),
    "nf_momentum_bicgstab": (C.c_int, [CTX, GP, C.c_int, NfLinks, P, C.c_double, C.c_double, C.c_int, C.c_int, P,
                                       C.POINTER(NfKrylovInfo)]),
    "nf_momentum_residual_unrelaxed": (C.c_int, [CTX, GP, C.c_int, NfLinks, P, P, DBL_OUT]),
    "nf_correct_velocity": (C.c_int, [CTX, GP, C.POINTER(NfBcProgram), P, P, P, P, P, P, P]),
    "nf_update_pressure": (C.c_int, [CTX, GP, P, P, C.c_double, P]),
    "nf_max_abs_divergence": (C.c_int, [CTX, GP, P, P, DBL_OUT]),
    "nf_simple_create": (C.c_int, [CTX, C.POINTER(C.c_void_p), C.POINTER(NfSimpleConfig)]),
    "nf_simple_destroy": (C.c_int, [C.c_void_p]),
    "nf_nccl_unique_id": (C.c_int, [CTX, C.c_void_p]),
    "nf_team_create_nccl": (C.c_int, [CTX, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "nf_team_create_virtual": (C.c_int, [CTX, C.c_int, C.POINTER(C.c_void_p)]),
    "nf_team_free": (C.c_int, [C.c_void_p]),
    "nf_team_uses_p2p": (C.c_int, [C.c_void_p]),
    "nf_team_benchmark": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, DBL_OUT, DBL_OUT]),
    "nf_slab_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nf_slab_coarse_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nf_simple_create_team": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(NfSimpleConfig)]),
    "nf_mg_smoother_timing": (C.c_int, [C.c_void_p, C.c_int, DBL_OUT, C.POINTER(C.c_longlong)]),
    "nf_simple_smoother_timing": (C.c_int, [C.c_void_p, C.c_int, DBL_OUT, C.POINTER(C.c_longlong)]),
    "nf_simple_phase_timing": (C.c_int, [C.c_void_p, C.c_int, DBL_OUT, DBL_OUT, DBL_OUT, C.POINTER(C.c_longlong)]),
    "nf_simple_local_rows": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nf_simple_ld": (C.c_int, [C.c_void_p]),
    "nf_simple_field": (C.c_void_p, [C.c_void_p, C.c_int]),
    "nf_simple_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "nf_simple_download": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "nf_simple_iterate": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.POINTER(NfSimpleInfo),
                                    C.POINTER(C.c_int)]),
}

_lib = None


def lib():
    """The loaded shared library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: naviflow_b200 has no CPU fallback. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C naviflow_b200/csrc`.")
        _lib = C.CDLL(LIB_PATH)
        missing = []
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(_lib, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
        if missing:
            raise ImportError(f"{LIB_PATH} does not export: {', '.join(missing)} (stale build?)")
    return _lib


def check(ctx_handle, status, what=""):
    if status != NF_OK:
        msg = lib().nf_last_error(ctx_handle)
        raise NfError(f"{what or 'libnaviflow_b200'} failed with status {status}: "
                      f"{msg.decode() if msg else ''}")
