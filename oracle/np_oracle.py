"""NumPy/SciPy restatement of the NaviFlow SIMPLE hot path -- TEST INFRASTRUCTURE.

This file is the *checker* for the CUDA path (tests/, smoke(), bench.py's
cpu_baseline / --impl reference).  ``naviflow_b200`` never imports it.

Conventions
-----------
* All fields are 2-D C-ordered ``(nx, ny)``-like arrays exactly as the reference
  holds them: ``p (nx,ny)``, ``u (nx+1,ny)``, ``v (nx,ny+1)``.  The reference's
  "flattened Fortran vector" (index ``i + j*nx``) is only a relabelling of the
  same 2-D array, so every solver here works on the 2-D array directly.
* Operation order inside each expression follows the reference so that results
  are bit-identical to it wherever the reference itself is elementwise NumPy
  (pinned by tests/test_oracle_vs_reference.py with ``assert_array_equal``).
* Citations are ``path:line`` relative to ``/root/reference/naviflow_oo``.

Parity status: pinned (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.sparse.linalg import spsolve
from scipy import interpolate as _spi


# ----------------------------------------------------------------------------
# a1  mesh  (preprocessing/mesh/structured.py:6-43)
# ----------------------------------------------------------------------------
def mesh_spacing(nx, ny, length=1.0, height=1.0):
    """dx = L/(nx-1), dy = H/(ny-1)  (structured.py:27-28; node spacing quirk)."""
    return length / (nx - 1), height / (ny - 1)


# ----------------------------------------------------------------------------
# a2  velocity boundary conditions (constructor/boundary_conditions.py:164-260)
# ----------------------------------------------------------------------------
DEFAULT_CAVITY_BCS = (("top", "velocity", {"u": 1.0, "v": 0.0}),
                      ("bottom", "wall", None),
                      ("left", "wall", None),
                      ("right", "wall", None))


def bc_conditions(entries=DEFAULT_CAVITY_BCS):
    """Ordered ``{location: {type: values}}`` as BoundaryConditionManager.set_condition
    builds it (boundary_conditions.py:96-125); insertion order is significant."""
    cond = {}
    for loc, typ, vals in entries:
        cond.setdefault(loc.lower(), {})[typ.lower()] = dict(vals or {})
    return cond


def apply_velocity_bc(u, v, nx, ny, conditions):
    """In-place restatement of apply_velocity_boundary_conditions
    (boundary_conditions.py:164-260).  ``nx, ny`` are the *caller's* values
    (some reference callers pass nx+1, which switches off the v[nx-1,:] reset)."""
    def u_right(val):
        if u.shape[0] == nx + 1:
            u[nx, :] = val
        elif u.shape[0] == nx and nx > 0:
            u[nx - 1, :] = val

    def u_top(val):
        if u.shape[1] > ny - 1 and ny > 0:
            u[:, ny - 1] = val

    def v_right(val):
        if v.shape[0] > nx - 1 and nx > 0:
            v[nx - 1, :] = val

    def v_top(val):
        if v.shape[1] == ny + 1:
            v[:, ny] = val
        elif v.shape[1] == ny and ny > 0:
            v[:, ny - 1] = val

    # default walls (:180-204)
    u[0, :] = 0.0
    u_right(0.0)
    u[:, 0] = 0.0
    u_top(0.0)
    v[0, :] = 0.0
    v_right(0.0)
    v[:, 0] = 0.0
    v_top(0.0)
    # registered conditions in insertion order (:207-258)
    for loc, conds in conditions.items():
        for typ, vals in conds.items():
            if typ == "velocity":
                uu, vv = vals.get("u", 0.0), vals.get("v", 0.0)
            elif typ == "wall":
                uu, vv = 0.0, 0.0
            else:
                continue
            if loc == "top":
                u_top(uu)
                v_top(vv)
            elif loc == "bottom":
                u[:, 0] = uu
                v[:, 0] = vv
            elif loc == "left":
                u[0, :] = uu
                v[0, :] = vv
            elif loc == "right":
                u_right(uu)
                v_right(vv)
    return u, v


def boundary_types(conditions):
    """get_boundary_types (boundary_conditions.py:266-287): first type per location,
    registered locations first (insertion order), then missing ones as 'wall'."""
    out = {}
    for loc, conds in conditions.items():
        if conds:
            out[loc] = next(iter(conds.keys()))
    for loc in ("top", "bottom", "left", "right"):
        out.setdefault(loc, "wall")
    return out


# ----------------------------------------------------------------------------
# a3/a4  power-law link coefficients (momentum_solver/discretization/power_law.py)
# ----------------------------------------------------------------------------
def power_law_A(F, D):
    """A(|P|) = max(0, 1-0.1|F/D|)^5, 0 where |D|<=1e-10, NaN->0 (power_law.py:19-44)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        pe = 0.1 * np.abs(F / D)
        base = np.maximum(0.0, 1.0 - pe)
        res = np.where(np.abs(D) > 1e-10, base ** 5, 0.0)
        res = np.nan_to_num(res, nan=0.0)
    return res


def u_coefficients(nx, ny, dx, dy, rho, mu, u, v, p, sides=("left", "right", "bottom", "top")):
    """Link coefficients of the u-momentum equation (power_law.py:46-209).
    ``sides`` = boundaries that have a registered condition (Practice-B folding)."""
    shp = (nx + 1, ny)
    a_e, a_w, a_n, a_s, a_p, src = (np.zeros(shp) for _ in range(6))
    De = mu * dy / dx
    Dw = mu * dy / dx
    Dn = mu * dx / dy
    Ds = mu * dx / dy
    A = power_law_A
    # interior i=1..nx-1, j=1..ny-2 (:89-110)
    I = slice(1, nx)
    Ip = slice(2, nx + 1)
    Im = slice(0, nx - 1)
    J = slice(1, ny - 1)
    Fe = 0.5 * rho * dy * (u[Ip, J] + u[I, J])
    Fw = 0.5 * rho * dy * (u[Im, J] + u[I, J])
    Fn = 0.5 * rho * dx * (v[I, 2:ny] + v[Im, 2:ny])
    Fs = 0.5 * rho * dx * (v[I, J] + v[Im, J])
    a_e[I, J] = De * A(Fe, De) + np.maximum(-Fe, 0)
    a_w[I, J] = Dw * A(Fw, Dw) + np.maximum(Fw, 0)
    a_n[I, J] = Dn * A(Fn, Dn) + np.maximum(-Fn, 0)
    a_s[I, J] = Ds * A(Fs, Ds) + np.maximum(Fs, 0)
    a_p[I, J] = a_e[I, J] + a_w[I, J] + a_n[I, J] + a_s[I, J] + (Fe - Fw) + (Fn - Fs)
    src[I, J] = (p[Im, J] - p[I, J]) * dy
    # bottom row j=0 (:112-125)
    j = 0
    Fe = 0.5 * rho * dy * (u[Ip, j] + u[I, j])
    Fw = 0.5 * rho * dy * (u[Im, j] + u[I, j])
    Fn = 0.5 * rho * dx * (v[I, j + 1] + v[Im, j + 1])
    a_e[I, j] = De * A(Fe, De) + np.maximum(-Fe, 0)
    a_w[I, j] = Dw * A(Fw, Dw) + np.maximum(Fw, 0)
    a_n[I, j] = Dn * A(Fn, Dn) + np.maximum(-Fn, 0)
    a_s[I, j] = 0
    a_p[I, j] = a_e[I, j] + a_w[I, j] + a_n[I, j] + (Fe - Fw) + Fn
    src[I, j] = (p[Im, j] - p[I, j]) * dy
    # top row j=ny-1 (:127-140)
    j = ny - 1
    Fe = 0.5 * rho * dy * (u[Ip, j] + u[I, j])
    Fw = 0.5 * rho * dy * (u[Im, j] + u[I, j])
    Fs = 0.5 * rho * dx * (v[I, j] + v[Im, j])
    a_e[I, j] = De * A(Fe, De) + np.maximum(-Fe, 0)
    a_w[I, j] = Dw * A(Fw, Dw) + np.maximum(Fw, 0)
    a_n[I, j] = 0
    a_s[I, j] = Ds * A(Fs, Ds) + np.maximum(Fs, 0)
    a_p[I, j] = a_e[I, j] + a_w[I, j] + a_s[I, j] + (Fe - Fw) - Fs
    src[I, j] = (p[Im, j] - p[I, j]) * dy
    # Practice B (:144-199): boundary links folded into the source, a_p unchanged
    if "left" in sides:
        src[1, :] += a_w[1, :] * u[0, :]
        a_w[1, :] = 0.0
    if "right" in sides:
        src[nx - 1, :] += a_e[nx - 1, :] * u[nx, :]
        a_e[nx - 1, :] = 0.0
    if "bottom" in sides:
        src[I, 1] += a_s[I, 1] * u[I, 0]
        a_s[I, 1] = 0.0
    if "top" in sides:
        src[I, ny - 2] += a_n[I, ny - 2] * u[I, ny - 1]
        a_n[I, ny - 2] = 0.0
    return dict(a_e=a_e, a_w=a_w, a_n=a_n, a_s=a_s, a_p=a_p, source=src)


def v_coefficients(nx, ny, dx, dy, rho, mu, u, v, p, sides=("left", "right", "bottom", "top")):
    """Link coefficients of the v-momentum equation (power_law.py:211-365)."""
    shp = (nx, ny + 1)
    a_e, a_w, a_n, a_s, a_p, src = (np.zeros(shp) for _ in range(6))
    De = mu * dy / dx
    Dw = mu * dy / dx
    Dn = mu * dx / dy
    Ds = mu * dx / dy
    A = power_law_A
    # interior i=1..nx-2, j=1..ny-1 (:255-271)
    I = slice(1, nx - 1)
    Ip = slice(2, nx)
    J = slice(1, ny)
    Jm = slice(0, ny - 1)
    Jp = slice(2, ny + 1)
    Fe = 0.5 * rho * dy * (u[Ip, J] + u[Ip, Jm])
    Fw = 0.5 * rho * dy * (u[I, J] + u[I, Jm])
    Fn = 0.5 * rho * dx * (v[I, J] + v[I, Jp])
    Fs = 0.5 * rho * dx * (v[I, Jm] + v[I, J])
    a_e[I, J] = De * A(Fe, De) + np.maximum(-Fe, 0)
    a_w[I, J] = Dw * A(Fw, Dw) + np.maximum(Fw, 0)
    a_n[I, J] = Dn * A(Fn, Dn) + np.maximum(-Fn, 0)
    a_s[I, J] = Ds * A(Fs, Ds) + np.maximum(Fs, 0)
    a_p[I, J] = a_e[I, J] + a_w[I, J] + a_n[I, J] + a_s[I, J] + (Fe - Fw) + (Fn - Fs)
    src[I, J] = (p[I, Jm] - p[I, J]) * dx
    # left column i=0 (:273-286)
    i = 0
    Fe = 0.5 * rho * dy * (u[i + 1, J] + u[i + 1, Jm])
    Fn = 0.5 * rho * dx * (v[i, Jp] + v[i, J])
    Fs = 0.5 * rho * dx * (v[i, Jm] + v[i, J])
    a_e[i, J] = De * A(Fe, De) + np.maximum(-Fe, 0)
    a_w[i, J] = 0
    a_n[i, J] = Dn * A(Fn, Dn) + np.maximum(-Fn, 0)
    a_s[i, J] = Ds * A(Fs, Ds) + np.maximum(Fs, 0)
    a_p[i, J] = a_e[i, J] + a_n[i, J] + a_s[i, J] + Fe + (Fn - Fs)
    src[i, J] = (p[i, Jm] - p[i, J]) * dx
    # right column i=nx-1 (:288-301)
    i = nx - 1
    Fw = 0.5 * rho * dy * (u[i, J] + u[i, Jm])
    Fn = 0.5 * rho * dx * (v[i, Jp] + v[i, J])
    Fs = 0.5 * rho * dx * (v[i, Jm] + v[i, J])
    a_e[i, J] = 0
    a_w[i, J] = Dw * A(Fw, Dw) + np.maximum(Fw, 0)
    a_n[i, J] = Dn * A(Fn, Dn) + np.maximum(-Fn, 0)
    a_s[i, J] = Ds * A(Fs, Ds) + np.maximum(Fs, 0)
    a_p[i, J] = a_w[i, J] + a_n[i, J] + a_s[i, J] - Fw + (Fn - Fs)
    src[i, J] = (p[i, Jm] - p[i, J]) * dx
    # Practice B (:304-355)
    if "bottom" in sides:
        src[:, 1] += a_s[:, 1] * v[:, 0]
        a_s[:, 1] = 0.0
    if "top" in sides:
        src[:, ny - 1] += a_n[:, ny - 1] * v[:, ny]
        a_n[:, ny - 1] = 0.0
    if "left" in sides:
        src[1, J] += a_w[1, J] * v[0, J]
        a_w[1, J] = 0.0
    if "right" in sides:
        src[nx - 2, J] += a_e[nx - 2, J] * v[nx - 1, J]
        a_e[nx - 2, J] = 0.0
    return dict(a_e=a_e, a_w=a_w, a_n=a_n, a_s=a_s, a_p=a_p, source=src)


# ----------------------------------------------------------------------------
# a5/a6  under-relaxation + fixed-sweep Jacobi momentum solve
#        (momentum_solver/jacobi_matrix_solver.py:153-264, 266-375)
# ----------------------------------------------------------------------------
def relax_coefficients(c, alpha, phi_bc):
    """a_p <- a_p/alpha ; S <- S + (1-alpha)*a_p_un/alpha*phi_bc (jacobi_matrix_solver.py:186-187)."""
    a_p = c["a_p"] / alpha
    src = c["source"] + (1 - alpha) * c["a_p"] / alpha * phi_bc
    return a_p, src


def _offdiag_times(a_e, a_w, a_n, a_s, x):
    """(A-D) x for the 5-point matrix of _build_sparse_matrix (:48-151), accumulated in
    the CSR column order W(idx-cols), S(idx-1), N(idx+1), E(idx+cols) that scipy's
    csr_matvec uses after tocsr() sorted the indices."""
    acc = np.zeros_like(x)
    acc[1:, :] += (-a_w[1:, :]) * x[:-1, :]
    acc[:, 1:] += (-a_s[:, 1:]) * x[:, :-1]
    acc[:, :-1] += (-a_n[:, :-1]) * x[:, 1:]
    acc[:-1, :] += (-a_e[:-1, :]) * x[1:, :]
    return acc


def momentum_jacobi(a_e, a_w, a_n, a_s, a_p, src, x0, n_sweeps):
    """n fixed Jacobi sweeps x <- D^-1 (b - (A-D) x), D^-1=0 where |a_p|<=1e-12
    (jacobi_matrix_solver.py:196-208)."""
    dinv = np.zeros_like(a_p)
    m = np.abs(a_p) > 1e-12
    dinv[m] = 1.0 / a_p[m]
    x = x0.copy()
    for _ in range(n_sweeps):
        x = dinv * (src - _offdiag_times(a_e, a_w, a_n, a_s, x))
    return x


def momentum_residual(a_e, a_w, a_n, a_s, a_p, src, x):
    """r = b - A x of the relaxed system (jacobi_matrix_solver.py:221-224).  The diagonal
    term enters the CSR row sum between S and N (sorted column order)."""
    acc = np.zeros_like(x)
    acc[1:, :] += (-a_w[1:, :]) * x[:-1, :]
    acc[:, 1:] += (-a_s[:, 1:]) * x[:, :-1]
    acc += a_p * x
    acc[:, :-1] += (-a_n[:, :-1]) * x[:, 1:]
    acc[:-1, :] += (-a_e[:-1, :]) * x[1:, :]
    return src - acc


def _masked_norm_ratio(r, b):
    r = r.copy()
    b = b.copy()
    for a in (r, b):
        a[0, :] = 0.0
        a[-1, :] = 0.0
        a[:, 0] = 0.0
        a[:, -1] = 0.0
    rn = np.linalg.norm(r)
    bn = np.linalg.norm(b)
    return rn / (bn + 1e-15)


def solve_u_momentum(nx, ny, dx, dy, rho, mu, u, v, p, alpha, conditions, n_sweeps):
    """JacobiMatrixMomentumSolver.solve_u_momentum (jacobi_matrix_solver.py:153-264).
    Returns (u_star, d_u, rel_norm, residual_field)."""
    u_bc, v_bc = apply_velocity_bc(u.copy(), v.copy(), nx, ny, conditions)
    sides = tuple(s for s in ("left", "right", "bottom", "top") if conditions.get(s))
    c = u_coefficients(nx, ny, dx, dy, rho, mu, u_bc, v_bc, p, sides)
    a_p, src = relax_coefficients(c, alpha, u_bc)
    u_star = momentum_jacobi(c["a_e"], c["a_w"], c["a_n"], c["a_s"], a_p, src, u, n_sweeps)
    d_u = np.full((nx + 1, ny), np.nan)
    m = np.abs(a_p) > 1e-12
    d_u[m] = dy / a_p[m]
    r = momentum_residual(c["a_e"], c["a_w"], c["a_n"], c["a_s"], a_p, src, u_star)
    norm = _masked_norm_ratio(r, src)
    field = r.copy()
    field[0, :] = 0.0
    field[1, :] = 0.0
    if nx > 1:
        field[nx - 1, :] = 0.0
    field[nx, :] = 0.0
    return u_star, d_u, norm, field


def solve_v_momentum(nx, ny, dx, dy, rho, mu, u, v, p, alpha, conditions, n_sweeps):
    """JacobiMatrixMomentumSolver.solve_v_momentum (jacobi_matrix_solver.py:266-375)."""
    u_bc, v_bc = apply_velocity_bc(u.copy(), v.copy(), nx, ny, conditions)
    sides = tuple(s for s in ("left", "right", "bottom", "top") if conditions.get(s))
    c = v_coefficients(nx, ny, dx, dy, rho, mu, u_bc, v_bc, p, sides)
    a_p, src = relax_coefficients(c, alpha, v_bc)
    v_star = momentum_jacobi(c["a_e"], c["a_w"], c["a_n"], c["a_s"], a_p, src, v, n_sweeps)
    d_v = np.full((nx, ny + 1), np.nan)
    m = np.abs(a_p) > 1e-12
    d_v[m] = dx / a_p[m]
    r = momentum_residual(c["a_e"], c["a_w"], c["a_n"], c["a_s"], a_p, src, v_star)
    norm = _masked_norm_ratio(r, src)
    field = r.copy()
    field[:, 0] = 0.0
    field[:, 1] = 0.0
    if ny > 1:
        field[:, ny - 1] = 0.0
    field[:, ny] = 0.0
    return v_star, d_v, norm, field


def momentum_unrelaxed_residual_norm(is_u, nx, ny, dx, dy, rho, mu, u, v, p, star, conditions):
    """||S_un - A_un star|| over the interior with the masks of MatrixFreeMomentumSolver._calculate_unrelaxed_residual
    (matrix_free_momentum.py:380-400), for a predicted field `star` of the power-law system assembled from (u, v, p): the
    convergence measure of the outer loop that does not depend on the momentum solver."""
    u_bc, v_bc = apply_velocity_bc(u.copy(), v.copy(), nx, ny, conditions)
    sides = tuple(s for s in ("left", "right", "bottom", "top") if conditions.get(s))
    c = (u_coefficients if is_u else v_coefficients)(nx, ny, dx, dy, rho, mu, u_bc, v_bc, p, sides)
    r = momentum_residual(c["a_e"], c["a_w"], c["a_n"], c["a_s"], c["a_p"], c["source"], star)
    r[0, :] = 0.0; r[-1, :] = 0.0; r[:, 0] = 0.0; r[:, -1] = 0.0
    if is_u:
        r[1, :] = 0.0; r[-2, :] = 0.0
    else:
        r[:, 1] = 0.0; r[:, -2] = 0.0
    return float(np.linalg.norm(r))


# ----------------------------------------------------------------------------
# a7  MatrixFreeMomentumSolver (solver/momentum_solver/matrix_free_momentum.py:403-544): Krylov solve of the
#     relaxed momentum system, interior rows 5-point, boundary rows identity
# ----------------------------------------------------------------------------
def mf_momentum_matvec(x, a_e, a_w, a_n, a_s, a_p):
    """_matvec_u / _matvec_v (:49-79): same expression order."""
    y = np.zeros_like(x)
    y[1:-1, 1:-1] = (a_p[1:-1, 1:-1] * x[1:-1, 1:-1] - a_e[1:-1, 1:-1] * x[2:, 1:-1] - a_w[1:-1, 1:-1] * x[:-2, 1:-1]
                     - a_n[1:-1, 1:-1] * x[1:-1, 2:] - a_s[1:-1, 1:-1] * x[1:-1, :-2])
    y[[0, -1], :] = x[[0, -1], :]
    y[:, [0, -1]] = x[:, [0, -1]]
    return y


def mf_momentum_ilu(a_e, a_w, a_n, a_s, a_p, drop_tol=1e-3, fill_factor=15):
    """_build_sparse_approx + _create_ilu_preconditioner (:82-172).  As coded, the north/south diagonals fail the
    length check (they hold rows*(cols-1) entries, not size-1) and are dropped: the ILU is built from the diagonal and
    the east/west links only."""
    from scipy.sparse import diags
    from scipy.sparse.linalg import spilu
    rows, cols = a_p.shape
    size = rows * cols
    data, offsets = [a_p.flatten()], [0]
    for arr, off in ((-a_e[:-1, :].flatten(), cols), (-a_w[1:, :].flatten(), -cols),
                     (-a_n[:, :-1].flatten(), 1), (-a_s[:, 1:].flatten(), -1)):
        if arr.size == size - abs(off):
            data.append(arr)
            offsets.append(off)
    A = diags(data, offsets, shape=(size, size), format="csr")
    ilu = spilu(A, drop_tol=drop_tol, fill_factor=fill_factor)
    return lambda z: ilu.solve(z.ravel()).reshape(z.shape)


def solve_momentum_krylov(is_u, nx, ny, dx, dy, rho, mu, u, v, p, alpha, conditions, tol=1e-8, maxiter=200,
                          precondition="ilu"):
    """MatrixFreeMomentumSolver.solve_u_momentum / solve_v_momentum with solver_type='bicgstab' (:403-544).
    precondition=None runs the same Krylov iteration without the ILU (what the device does; the answers agree to the
    stopping tolerance max(tol, 1e-5 ||b||)).  Returns (star, d, abs_unrelaxed_norm, residual_field, iterations)."""
    # the reference passes nx+1 as "nx" (:419, :491), which switches off the v[nx-1,:] reset
    u_bc, v_bc = apply_velocity_bc(u.copy(), v.copy(), nx + 1, ny, conditions)
    sides = tuple(s for s in ("left", "right", "bottom", "top") if conditions.get(s))
    c = (u_coefficients if is_u else v_coefficients)(nx, ny, dx, dy, rho, mu, u_bc, v_bc, p, sides)
    phi_bc = u_bc if is_u else v_bc
    a_p_un, src_un = c["a_p"], c["source"]
    a_p = np.where(np.abs(a_p_un) > 1e-12, a_p_un, 1e-12) / alpha
    src = src_un + (1 - alpha) * a_p * phi_bc
    links = (c["a_e"], c["a_w"], c["a_n"], c["a_s"])
    M = mf_momentum_ilu(*links, a_p) if precondition == "ilu" else None
    x0 = (u if is_u else v)
    star, info, iters = bicgstab(lambda z: mf_momentum_matvec(z, *links, a_p), src, x0=x0, atol=tol, maxiter=maxiter,
                                 M=M, order="C")
    star = star.copy()
    if is_u:
        star, _ = apply_velocity_bc(star, v_bc, nx + 1, ny, conditions)
    else:
        _, star = apply_velocity_bc(u_bc, star, nx + 1, ny, conditions)
    d = np.where(np.abs(a_p) > 1e-12, (dy if is_u else dx) / a_p, 0.0)
    r = src_un - mf_momentum_matvec(star, *links, a_p_un)   # _calculate_unrelaxed_residual (:379-400)
    r[0, :] = 0.0
    r[-1, :] = 0.0
    r[:, 0] = 0.0
    r[:, -1] = 0.0
    if is_u:
        r[1, :] = 0.0
        if nx > 1:
            r[-2, :] = 0.0
        interior = r[1:nx, 1:ny - 1]
    else:
        r[:, 1] = 0.0
        if ny > 1:
            r[:, -2] = 0.0
        interior = r[1:nx - 1, 1:ny]
    return star, d, float(np.linalg.norm(interior)), r, iters


# ----------------------------------------------------------------------------
# a8  continuity RHS (pressure_solver/helpers/rhs_construction.py:3-21)
# ----------------------------------------------------------------------------
def continuity_rhs(nx, ny, dx, dy, rho, u_star, v_star):
    """b[i,j] = rho*(u*[i,j]dy - u*[i+1,j]dy + v*[i,j]dx - v*[i,j+1]dx), b[0,0]=0; 2-D."""
    b = rho * (u_star[:-1, :] * dy - u_star[1:, :] * dy + v_star[:, :-1] * dx - v_star[:, 1:] * dx)
    b[0, 0] = 0
    return b


# ----------------------------------------------------------------------------
# a9  pressure-correction operator (helpers/matrix_free.py:6-135; coeff_matrix.py:6-121)
# ----------------------------------------------------------------------------
def pressure_coefficients(nx, ny, dx, dy, rho, d_u, d_v):
    """(aE, aW, aN, aS, diag) with the reference's Neumann folding (matrix_free.py:45-84)."""
    aE = np.zeros((nx, ny))
    aW = np.zeros((nx, ny))
    aN = np.zeros((nx, ny))
    aS = np.zeros((nx, ny))
    diag = np.zeros((nx, ny))
    aE[:-1, :] = rho * d_u[1:nx, :] * dy
    aW[1:, :] = rho * d_u[1:nx, :] * dy
    aN[:, :-1] = rho * d_v[:, 1:ny] * dx
    aS[:, 1:] = rho * d_v[:, 1:ny] * dx
    diag[0, :] += aE[0, :]
    diag[nx - 1, :] += aW[nx - 1, :]
    diag[:, 0] += aN[:, 0]
    diag[:, ny - 1] += aS[:, ny - 1]
    aE[0, :] = 0
    aW[nx - 1, :] = 0
    aN[:, 0] = 0
    aS[:, ny - 1] = 0
    diag += aE + aW + aN + aS
    return aE, aW, aN, aS, diag


def apply_A(p, dx, dy, rho, d_u, d_v, pin=True):
    """A p for 2-D p (matrix_free.py:86-133): diag*p - E - W - N - S, identity row at (0,0)."""
    nx, ny = p.shape
    aE, aW, aN, aS, diag = pressure_coefficients(nx, ny, dx, dy, rho, d_u, d_v)
    out = diag * p
    out[:-1, :] -= aE[:-1, :] * p[1:, :]
    out[1:, :] -= aW[1:, :] * p[:-1, :]
    out[:, :-1] -= aN[:, :-1] * p[:, 1:]
    out[:, 1:] -= aS[:, 1:] * p[:, :-1]
    if pin:
        out[0, 0] = p[0, 0]
    return out


def assemble_A(nx, ny, dx, dy, rho, d_u, d_v, pin=True):
    """Same operator as CSR with the reference's F-order unknown numbering k = i + j*nx
    (coeff_matrix.py:42, :114-119)."""
    aE, aW, aN, aS, diag = pressure_coefficients(nx, ny, dx, dy, rho, d_u, d_v)
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    k = (ii + jj * nx)
    rows = [k.ravel()]
    cols = [k.ravel()]
    data = [diag.ravel()]
    m = ii < nx - 1
    rows.append(k[m]); cols.append(k[m] + 1); data.append(-aE[m])
    m = ii > 0
    rows.append(k[m]); cols.append(k[m] - 1); data.append(-aW[m])
    m = jj < ny - 1
    rows.append(k[m]); cols.append(k[m] + nx); data.append(-aN[m])
    m = jj > 0
    rows.append(k[m]); cols.append(k[m] - nx); data.append(-aS[m])
    A = sparse.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(nx * ny, nx * ny)).tolil()
    if pin:
        A[0, :] = 0
        A[0, 0] = 1
    return A.tocsr()


def direct_solve(rhs2d, dx, dy, rho, d_u, d_v):
    """spsolve(get_coeff_mat, rhs) (multigrid.py:268-302, direct.py:82-85); 2-D in/out."""
    nx, ny = rhs2d.shape
    A = assemble_A(nx, ny, dx, dy, rho, d_u, d_v)
    x = spsolve(A.tocsc(), rhs2d.flatten("F"))
    return x.reshape((nx, ny), order="F")


# ----------------------------------------------------------------------------
# a10  weighted Jacobi pressure iteration (pressure_solver/jacobi.py:38-78, 157-203)
# ----------------------------------------------------------------------------
def jacobi_diag(nx, ny, dx, dy, rho, d_u, d_v):
    """Non-standard diagonal: neighbour sums, boundary rows/cols doubled (jacobi.py:38-78)."""
    diag = np.zeros((nx, ny))
    diag[:-1, :] += rho * d_u[1:nx, :] * dy
    diag[1:, :] += rho * d_u[1:nx, :] * dy
    diag[:, :-1] += rho * d_v[:, 1:ny] * dx
    diag[:, 1:] += rho * d_v[:, 1:ny] * dx
    diag[0, :] += diag[0, :]
    diag[nx - 1, :] += diag[nx - 1, :]
    diag[:, 0] += diag[:, 0]
    diag[:, ny - 1] += diag[:, ny - 1]
    diag[diag < 1e-15] = 1.0
    diag[0, 0] = 1.0
    return diag


def jacobi_iterate(p, b, dx, dy, rho, d_u, d_v, omega, n_iter):
    """n_iter iterations p <- p + omega (b - A p)/diag with the pin (jacobi.py:160-203,
    track_residuals=False path).  2-D in/out, inputs untouched."""
    nx, ny = p.shape
    p = p.copy()
    b = b.copy()
    diag = jacobi_diag(nx, ny, dx, dy, rho, d_u, d_v)
    p[0, 0] = 0.0
    b[0, 0] = 0.0
    for _ in range(n_iter):
        p[0, 0] = 0.0
        Ap = apply_A(p, dx, dy, rho, d_u, d_v)
        p = p + omega * (b - Ap) / diag
        p[0, 0] = 0.0
    return p


# ----------------------------------------------------------------------------
# a11  red-black SOR (pressure_solver/gauss_seidel.py:214-305)
# ----------------------------------------------------------------------------
def sor_coefficients(nx, ny, dx, dy, rho, d_u, d_v):
    """_precompute_coefficients (gauss_seidel.py:214-266): as a9 plus aP<1e-15 -> 1."""
    aE, aW, aN, aS, aP = pressure_coefficients(nx, ny, dx, dy, rho, d_u, d_v)
    aP[aP < 1e-15] = 1.0
    return aE, aW, aN, aS, aP


def rb_sor(p, b, dx, dy, rho, d_u, d_v, omega, n_sweeps):
    """n_sweeps red-black SOR sweeps (gauss_seidel.py:144-163, 268-305).  Red = (i+j) even
    without (0,0); black = the rest including (0,0), which is re-pinned after every sweep."""
    nx, ny = p.shape
    p = p.copy()
    aE, aW, aN, aS, aP = sor_coefficients(nx, ny, dx, dy, rho, d_u, d_v)
    p[0, 0] = 0.0
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    red = ((ii + jj) % 2 == 0)
    red[0, 0] = False
    black = ~red
    for _ in range(n_sweeps):
        inv = 1.0 / aP
        for mask in (red, black):
            e = np.zeros_like(p); w = np.zeros_like(p); n = np.zeros_like(p); s = np.zeros_like(p)
            e[:-1, :] = aE[:-1, :] * p[1:, :]
            w[1:, :] = aW[1:, :] * p[:-1, :]
            n[:, :-1] = aN[:, :-1] * p[:, 1:]
            s[:, 1:] = aS[:, 1:] * p[:, :-1]
            pn = (b + e + w + n + s) * inv
            p[mask] = p[mask] + omega * (pn[mask] - p[mask])
        p[0, 0] = 0.0
    return p


# ----------------------------------------------------------------------------
# 8f rank 3: lexicographic / symmetric Gauss-Seidel (pressure_solver/gauss_seidel.py:307-367)
# ----------------------------------------------------------------------------
def gs_lex(p, b, dx, dy, rho, d_u, d_v, omega, n_sweeps, symmetric=False):
    """GaussSeidelSolver(method_type='standard' | 'symmetric').solve(p=p, b=b, num_iterations=n_sweeps): sequential SOR
    sweeps `for j: for i:` (the symmetric variant adds the reverse sweep), pinned cell (0,0) skipped and reset to 0.
    Evaluated here by anti-diagonals i+j = const (vectorised): every update reads the new west/south and the old east/north
    values exactly as the loop order does, so the bits are the loop's."""
    nx, ny = b.shape
    aE, aW, aN, aS, aP = sor_coefficients(nx, ny, dx, dy, rho, d_u, d_v)
    inv = 1.0 / aP
    p = p.copy()
    p[0, 0] = 0.0
    P = np.zeros((nx + 2, ny + 2))   # padded copy: out-of-domain neighbours contribute exactly 0
    diags = []
    for d in range(nx + ny - 1):
        i = np.arange(max(0, d - ny + 1), min(nx - 1, d) + 1)
        j = d - i
        keep = ~((i == 0) & (j == 0))
        diags.append((i[keep], j[keep]))

    def one_pass(order):
        for i, j in order:
            if i.size == 0:
                continue
            P[1:-1, 1:-1] = p
            east = np.where(i < nx - 1, aE[i, j] * P[i + 2, j + 1], 0.0)
            west = np.where(i > 0, aW[i, j] * P[i, j + 1], 0.0)
            north = np.where(j < ny - 1, aN[i, j] * P[i + 1, j + 2], 0.0)
            south = np.where(j > 0, aS[i, j] * P[i + 1, j], 0.0)
            p_new = ((((b[i, j] + east) + west) + north) + south) * inv[i, j]
            p[i, j] = p[i, j] + omega * (p_new - p[i, j])

    for _ in range(n_sweeps):
        one_pass(diags)
        if symmetric:
            one_pass(diags[::-1])
        p[0, 0] = 0.0
    return p


# ----------------------------------------------------------------------------
# a12  multigrid transfer operators (helpers/multigrid_helpers.py)
# ----------------------------------------------------------------------------
def restrict_inject(f):
    """fine[1::2, 1::2] (multigrid_helpers.py:8-21)."""
    return f[1::2, 1::2].copy()


def restrict_full_weighting(f):
    """1/4 centre + 1/8 edges + 1/16 corners at fine odd indices, nc=(nf-1)//2
    (multigrid_helpers.py:23-70)."""
    c = f[1:-1:2, 1:-1:2]
    n = f[1:-1:2, 2::2]
    s = f[1:-1:2, :-2:2]
    e = f[2::2, 1:-1:2]
    w = f[:-2:2, 1:-1:2]
    ne = f[2::2, 2::2]
    nw = f[:-2:2, 2::2]
    se = f[2::2, :-2:2]
    sw = f[:-2:2, :-2:2]
    return c / 4.0 + (n + s + e + w) / 8.0 + (ne + nw + se + sw) / 16.0


def prolong_linear(c, m):
    """Bilinear prolongation to an m x m grid with the reference's exact index rules
    (multigrid_helpers.py:73-192): coarse k -> fine 2k+1; even points averaged; outer ring
    copied from ring 1; on an even-sized fine grid the last two rows/cols stay 0."""
    mc = c.shape[0]
    f = np.zeros((m, m))
    # coincident points
    ic = np.arange(mc)
    ok = 2 * ic + 1 < m
    src = ic[ok]
    f[np.ix_(2 * src + 1, 2 * src + 1)] = c[np.ix_(src, src)]
    if m <= 3:
        return f
    ih = np.arange(mc - 1)
    okh = 2 * ih + 2 < m
    sh = ih[okh]
    # odd rows, even cols
    f[np.ix_(2 * src + 1, 2 * sh + 2)] = 0.5 * (c[np.ix_(src, sh)] + c[np.ix_(src, sh + 1)])
    # even rows, odd cols
    f[np.ix_(2 * sh + 2, 2 * src + 1)] = 0.5 * (c[np.ix_(sh, src)] + c[np.ix_(sh + 1, src)])
    # even rows, even cols
    f[np.ix_(2 * sh + 2, 2 * sh + 2)] = 0.25 * (c[np.ix_(sh, sh)] + c[np.ix_(sh + 1, sh)]
                                                + c[np.ix_(sh, sh + 1)] + c[np.ix_(sh + 1, sh + 1)])
    f[1:-1, 0] = f[1:-1, 1]
    f[1:-1, -1] = f[1:-1, -2]
    f[0, 1:-1] = f[1, 1:-1]
    f[-1, 1:-1] = f[-2, 1:-1]
    f[0, 0] = f[1, 1]
    f[0, -1] = f[1, -2]
    f[-1, 0] = f[-2, 1]
    f[-1, -1] = f[-2, -2]
    return f


def prolong_cubic(c, m):
    """'Cubic' prolongation = FITPACK interpolating bicubic spline on linspace(0,1,.)
    coordinates (multigrid_helpers.py:333-391).  The arithmetic lives in scipy
    (RectBivariateSpline); called here exactly as the reference calls it."""
    mc = c.shape[0]
    xc = np.linspace(0, 1, mc)
    xf = np.linspace(0, 1, m)
    if mc >= 4:
        return _spi.RectBivariateSpline(xc, xc, c)(xf, xf)
    kind = "quadratic" if mc == 3 else "linear"
    fn = _spi.RegularGridInterpolator((xc, xc), c, method=kind, bounds_error=False, fill_value=None)
    XX, YY = np.meshgrid(xf, xf, indexing="ij")
    return fn(np.vstack([XX.ravel(), YY.ravel()]).T).reshape((m, m))


def notaknot_matrix(mc, m):
    """Independent restatement of the 1-D operator behind prolong_cubic: the (m x mc)
    matrix P with P @ y = not-a-knot cubic spline through (linspace(0,1,mc), y) evaluated
    at linspace(0,1,m).  prolong_cubic(c, m) == P @ c @ P.T."""
    xc = np.linspace(0, 1, mc)
    xf = np.linspace(0, 1, m)
    spl = _spi.make_interp_spline(xc, np.eye(mc), k=3)
    return spl(xf)


def restrict_coefficients(d_u, d_v, nx, ny, nxc, nyc):
    """Harmonic-mean coarsening x0.25 (multigrid_helpers.py:196-329):
    d_u^c[I,J] = H(d_u[2I,2J], d_u[2I+1,2J]) for I=1..nxc-1 (arithmetic mean unless both >0),
    d_v^c[I,J] = H(d_v[2I,2J], d_v[2I,2J+1]) for J=1..nyc-1, boundary faces copied."""
    duc = np.zeros((nxc + 1, nyc))
    dvc = np.zeros((nxc, nyc + 1))

    def hmean(d1, d2):
        out = 0.5 * (d1 + d2)
        m = (d1 > 0) & (d2 > 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            out[m] = 2.0 / (1.0 / d1[m] + 1.0 / d2[m])
        return out

    I = np.arange(1, nxc)
    J = np.arange(nyc)
    I = I[2 * I < nx]
    J = J[2 * J < ny]
    duc[np.ix_(I, J)] = hmean(d_u[np.ix_(2 * I, 2 * J)], d_u[np.ix_(2 * I + 1, 2 * J)])
    I = np.arange(nxc)
    J = np.arange(1, nyc)
    I = I[2 * I < nx]
    J = J[2 * J < ny]
    dvc[np.ix_(I, J)] = hmean(d_v[np.ix_(2 * I, 2 * J)], d_v[np.ix_(2 * I, 2 * J + 1)])
    J = np.arange(nyc)
    J = J[2 * J < ny]
    duc[0, J] = d_u[0, 2 * J]
    duc[nxc, J] = d_u[nx, 2 * J]
    I = np.arange(nxc)
    I = I[2 * I < nx]
    dvc[I, 0] = d_v[2 * I, 0]
    dvc[I, nyc] = d_v[2 * I, ny]
    duc *= 0.25
    dvc *= 0.25
    return duc, dvc


_RESTRICT = {"restrict_inject": restrict_inject, "restrict_full_weighting": restrict_full_weighting}
_PROLONG = {"interpolate_linear": prolong_linear, "interpolate_cubic": prolong_cubic}


class MGConfig:
    """Constructor arguments of MultiGridSolver (multigrid.py:31-37) + smoother choice."""

    def __init__(self, smoother="red_black", omega=1.5, pre=1, post=1, cycle_type="v",
                 cycle_type_buildup="v", cycle_type_final=None, max_cycles_buildup=1,
                 restriction="restrict_full_weighting", interpolation="interpolate_linear",
                 coarsest=7, max_iterations=100, tolerance=1e-8, length=1.0, height=1.0, rho=1.0):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def _smooth(cfg, p, b, dx, dy, d_u, d_v, n):
    if cfg.smoother == "red_black":
        return rb_sor(p, b, dx, dy, cfg.rho, d_u, d_v, cfg.omega, n)
    if cfg.smoother == "jacobi":
        return jacobi_iterate(p, b, dx, dy, cfg.rho, d_u, d_v, cfg.omega, n)
    if cfg.smoother in ("standard", "symmetric"):   # sequential Gauss-Seidel smoothers (gauss_seidel.py:307-367)
        return gs_lex(p, b, dx, dy, cfg.rho, d_u, d_v, cfg.omega, n, symmetric=(cfg.smoother == "symmetric"))
    raise ValueError(cfg.smoother)


def _coarsen(cfg, fine2d, nx, ny, d_u, d_v):
    rc = _RESTRICT[cfg.restriction](fine2d)
    nxc, nyc = rc.shape
    dxc, dyc = mesh_spacing(nxc, nyc, cfg.length, cfg.height)  # StructuredMesh(nc,nc,L,H), multigrid.py:373
    duc, dvc = restrict_coefficients(d_u, d_v, nx, ny, nxc, nyc)
    return rc, nxc, nyc, dxc, dyc, duc, dvc


def mg_cycle(cfg, p, rhs, dx, dy, d_u, d_v, kind="v"):
    """One V- (multigrid.py:304-432) or W-cycle (:434-560); 2-D arrays."""
    nx, ny = rhs.shape
    if nx <= cfg.coarsest:
        return direct_solve(rhs, dx, dy, cfg.rho, d_u, d_v)
    p = _smooth(cfg, p, rhs, dx, dy, d_u, d_v, cfg.pre)
    r = rhs - apply_A(p, dx, dy, cfg.rho, d_u, d_v)
    rc, nxc, nyc, dxc, dyc, duc, dvc = _coarsen(cfg, r, nx, ny, d_u, d_v)
    ec = np.zeros_like(rc)
    for _ in range(2 if kind == "w" else 1):
        ec = mg_cycle(cfg, ec, rc, dxc, dyc, duc, dvc, kind)
    p = p + _PROLONG[cfg.interpolation](ec, nx)
    return _smooth(cfg, p, rhs, dx, dy, d_u, d_v, cfg.post)


def mg_fmg(cfg, rhs, dx, dy, d_u, d_v):
    """Recursive FMG (multigrid.py:562-688): restrict the RHS down, direct solve, cubic
    prolongation (hard-coded :631), max_cycles_buildup cycles per level with early exit."""
    nx, ny = rhs.shape
    if nx <= cfg.coarsest:
        return direct_solve(rhs, dx, dy, cfg.rho, d_u, d_v)
    rc, nxc, nyc, dxc, dyc, duc, dvc = _coarsen(cfg, rhs, nx, ny, d_u, d_v)
    sc = mg_fmg(cfg, rc, dxc, dyc, duc, dvc)
    x = prolong_cubic(sc, nx)
    for _ in range(cfg.max_cycles_buildup):
        x = mg_cycle(cfg, x, rhs, dx, dy, d_u, d_v, cfg.cycle_type_buildup)
        if cfg.tolerance < 1.0:
            r = rhs - apply_A(x, dx, dy, cfg.rho, d_u, d_v)
            rn = np.linalg.norm(r.flatten("F"))
            bn = np.linalg.norm(rhs)
            rel = rn / bn if bn > 0 else rn
            if rel < cfg.tolerance:
                break
    return x


def mg_solve(cfg, nx, ny, dx, dy, u_star, v_star, d_u, d_v):
    """MultiGridSolver.solve (multigrid.py:121-266).  Returns (p', info) with
    info = {'rel_norm': absolute ||r||_2 (:257), 'field': r, 'cycles', 'res_history'}."""
    b = continuity_rhs(nx, ny, dx, dy, cfg.rho, u_star, v_star)
    x = np.zeros_like(b)
    hist = []
    cycles = 0

    def resid(x):
        r = b - apply_A(x, dx, dy, cfg.rho, d_u, d_v)
        rn = np.linalg.norm(r.flatten("F"), 2)
        bn = np.linalg.norm(b.flatten("F"), 2)
        return r, rn, (rn / bn if bn > 0 else rn)

    if cfg.cycle_type == "fmg":
        x = mg_fmg(cfg, b, dx, dy, d_u, d_v)
        if cfg.cycle_type_final:
            x = mg_cycle(cfg, x, b, dx, dy, d_u, d_v, cfg.cycle_type_final)
            cycles += 1
        r, rn, rel = resid(x)
        hist.append(rel)
    else:
        r, rn = None, None
        for k in range(cfg.max_iterations):
            x = mg_cycle(cfg, x, b, dx, dy, d_u, d_v, cfg.cycle_type)
            cycles += 1
            r, rn, rel = resid(x)
            hist.append(rel)
            if rel < cfg.tolerance:
                break
    return x, {"rel_norm": rn, "field": r, "cycles": cycles, "res_history": hist}


# ----------------------------------------------------------------------------
# a13 / K14-K15  Krylov solvers in scipy's operation order
#   scipy/sparse/linalg/_isolve/iterative.py (scipy 1.18.x; third-party, not under
#   /root/reference).  Call sites: pressure_solver/matrix_free_BiCGSTAB.py:234-242.
#   atol_eff = max(atol, rtol*||b||), rtol = 1e-5 (scipy default; the reference never passes it).
# ----------------------------------------------------------------------------
def _krylov_tol(b, atol, rtol):
    bn = np.linalg.norm(b)
    return max(float(atol), float(rtol) * float(bn)), bn


def cg(matvec, b, x0=None, atol=0.0, rtol=1e-5, maxiter=None, M=None, order="F"):
    """scipy.sparse.linalg.cg restated.  ``b`` is 2-D; it is flattened in ``order`` ('F' is
    what the reference's call sites hand to scipy) -- only the dot-product summation order
    depends on it.  Returns (x, info, iters)."""
    shape = b.shape
    b = b.flatten(order)
    n = b.size
    x = np.zeros(n) if x0 is None else x0.flatten(order)
    atol_eff, bn = _krylov_tol(b, atol, rtol)
    if bn == 0:
        return b.reshape(shape, order=order).copy(), 0, 0
    if maxiter is None:
        maxiter = n * 10
    mv = lambda z: matvec(z.reshape(shape, order=order)).flatten(order)
    psolve = (lambda z: z.copy()) if M is None else (
        lambda z: M(z.reshape(shape, order=order)).flatten(order))
    r = b - mv(x) if x.any() else b.copy()
    rho_prev, p = None, None
    for it in range(maxiter):
        if np.linalg.norm(r) < atol_eff:
            return x.reshape(shape, order=order), 0, it
        z = psolve(r)
        rho_cur = np.dot(r, z)
        if it > 0:
            beta = rho_cur / rho_prev
            p *= beta
            p += z
        else:
            p = np.empty_like(r)
            p[:] = z[:]
        q = mv(p)
        alpha = rho_cur / np.dot(p, q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho_cur
    return x.reshape(shape, order=order), maxiter, maxiter


def bicgstab(matvec, b, x0=None, atol=0.0, rtol=1e-5, maxiter=None, M=None, order="F"):
    """scipy.sparse.linalg.bicgstab restated (same conventions as ``cg``).
    Returns (x, info, iters)."""
    shape = b.shape
    b = b.flatten(order)
    n = b.size
    x = np.zeros(n) if x0 is None else x0.flatten(order)
    atol_eff, bn = _krylov_tol(b, atol, rtol)
    if bn == 0:
        return b.reshape(shape, order=order).copy(), 0, 0
    if maxiter is None:
        maxiter = n * 10
    mv = lambda z: matvec(z.reshape(shape, order=order)).flatten(order)
    psolve = (lambda z: z.copy()) if M is None else (
        lambda z: M(z.reshape(shape, order=order)).flatten(order))
    rhotol = np.finfo(x.dtype.char).eps ** 2
    omegatol = rhotol
    rho_prev, omega, alpha, p, v = None, None, None, None, None
    r = b - mv(x) if x.any() else b.copy()
    rtilde = r.copy()
    for it in range(maxiter):
        if np.linalg.norm(r) < atol_eff:
            return x.reshape(shape, order=order), 0, it
        rho = np.dot(rtilde, r)
        if np.abs(rho) < rhotol:
            return x.reshape(shape, order=order), -10, it
        if it > 0:
            if np.abs(omega) < omegatol:
                return x.reshape(shape, order=order), -11, it
            beta = (rho / rho_prev) * (alpha / omega)
            p -= omega * v
            p *= beta
            p += r
        else:
            s = np.empty_like(r)
            p = r.copy()
        phat = psolve(p)
        v = mv(phat)
        rv = np.dot(rtilde, v)
        if rv == 0:
            return x.reshape(shape, order=order), -11, it
        alpha = rho / rv
        r -= alpha * v
        s[:] = r[:]
        if np.linalg.norm(s) < atol_eff:
            x += alpha * phat
            return x.reshape(shape, order=order), 0, it + 1
        shat = psolve(s)
        t = mv(shat)
        omega = np.dot(t, s) / np.dot(t, t)
        x += alpha * phat
        x += omega * shat
        r -= omega * t
        rho_prev = rho
    return x.reshape(shape, order=order), maxiter, maxiter


def krylov_pressure_solve(kind, nx, ny, dx, dy, u_star, v_star, d_u, d_v, tol=1e-7,
                          maxiter=1000, rho=1.0):
    """MatrixFreeBiCGSTABSolver.solve (matrix_free_BiCGSTAB.py:163-287) and its CG twin
    (the reference's only CG classes need pyamg; oracle = scipy cg on compute_Ap_product,
    SURVEY.md section 2 row 6g).  rel_norm = ||r_int|| / ||b_int|| (:255-279; the reference
    zeroes the edges of ``rhs`` in place through a view before taking its norm)."""
    b = continuity_rhs(nx, ny, dx, dy, rho, u_star, v_star)
    mv = lambda z: apply_A(z, dx, dy, rho, d_u, d_v)
    fn = cg if kind == "cg" else bicgstab
    x, info, iters = fn(mv, b, atol=tol, maxiter=maxiter)
    bi = b.copy()
    Ax = mv(x)
    for a in (bi, Ax):
        a[0, :] = 0; a[:, 0] = 0; a[-1, :] = 0; a[:, -1] = 0
    r = bi - Ax
    return x, {"rel_norm": np.linalg.norm(r) / np.linalg.norm(bi), "field": r,
               "iterations": iters, "info": info}


def mg_preconditioner(dx, dy, d_u, d_v, kind="v", cycles=1, omega=0.8, pre=2, post=2, coarsest=7, tolerance=1e-7,
                      cycle_type_buildup="v", max_cycles_buildup=1, restriction="restrict_full_weighting",
                      interpolation="interpolate_linear"):
    """M of MatrixFreeBiCGSTABSolver(use_preconditioner=True, preconditioner='multigrid') (matrix_free_BiCGSTAB.py:102-161):
    `cycles` multigrid cycles of the given kind on A y = z from y = 0 (red-black smoother with omega =
    smoother_relaxation)."""
    cfg = MGConfig(smoother="red_black", omega=omega, pre=pre, post=post, coarsest=coarsest, tolerance=tolerance,
                   cycle_type_buildup=cycle_type_buildup, max_cycles_buildup=max_cycles_buildup, restriction=restriction,
                   interpolation=interpolation)

    def M(z):
        x = np.zeros_like(z)
        for _ in range(cycles):
            x = mg_fmg(cfg, z, dx, dy, d_u, d_v) if kind == "fmg" else mg_cycle(cfg, x, z, dx, dy, d_u, d_v, kind)
        return x
    return M


def bicgstab_mg_pressure_solve(nx, ny, dx, dy, u_star, v_star, d_u, d_v, tol=1e-7, maxiter=1000, kind="v", cycles=1, **mg):
    """MatrixFreeBiCGSTABSolver.solve with the multigrid preconditioner (:163-287).  Returns (p', info dict with rel_norm =
    ||r_interior|| / ||b_interior|| as krylov_pressure_solve computes it (:255-279), iterations)."""
    b = continuity_rhs(nx, ny, dx, dy, 1.0, u_star, v_star)
    mv = lambda z: apply_A(z, dx, dy, 1.0, d_u, d_v)
    x, info, iters = bicgstab(mv, b, atol=tol, maxiter=maxiter, M=mg_preconditioner(dx, dy, d_u, d_v, kind, cycles, **mg))
    bi = b.copy()
    Ax = mv(x)
    for a in (bi, Ax):
        a[0, :] = 0; a[:, 0] = 0; a[-1, :] = 0; a[:, -1] = 0
    r = bi - Ax
    return x, {"rel_norm": np.linalg.norm(r) / np.linalg.norm(bi), "field": r, "iterations": iters, "info": info}


# ----------------------------------------------------------------------------
# a14  velocity correction (velocity_solver/standard.py:10-69)
# ----------------------------------------------------------------------------
def cg_mg_pressure_solve(nx, ny, dx, dy, u_star, v_star, d_u, d_v, tol=1e-5, maxiter=500, kind="v", cycles=1, **mg):
    """GeoMultigridPrecondCGSolver.solve (pressure_solver/geo_multigrid_cg.py:73-197): scipy cg on the pressure matrix with
    M = `cycles` multigrid cycles from zero, x0 = 0, atol = tolerance (scipy's default rtol = 1e-5 governs as well).  The
    reference multiplies with the assembled matrix whose first row is replaced by the identity (:116-123); the matrix-free
    operator is the same matrix (SURVEY 5a = 5b).  Returns (p', iterations, info): the reference returns the bare array."""
    b = continuity_rhs(nx, ny, dx, dy, 1.0, u_star, v_star)
    mv = lambda z: apply_A(z, dx, dy, 1.0, d_u, d_v)
    M = mg_preconditioner(dx, dy, d_u, d_v, kind=kind, cycles=cycles, **mg)
    x, info, iters = cg(mv, b, atol=tol, maxiter=maxiter, M=M)
    return x, iters, info


def correct_velocity(nx, ny, u_star, v_star, p_prime, d_u, d_v, conditions):
    u = u_star.copy()
    v = v_star.copy()
    u[1:nx, 1:ny - 1] = u_star[1:nx, 1:ny - 1] + d_u[1:nx, 1:ny - 1] * (
        p_prime[0:nx - 1, 1:ny - 1] - p_prime[1:nx, 1:ny - 1])
    v[1:nx - 1, 1:ny] = v_star[1:nx - 1, 1:ny] + d_v[1:nx - 1, 1:ny] * (
        p_prime[1:nx - 1, 0:ny - 1] - p_prime[1:nx - 1, 1:ny])
    return apply_velocity_bc(u, v, nx, ny, conditions)


# ----------------------------------------------------------------------------
# a15  pressure update + Neumann copies (Algorithms/simple.py:148-150,
#      Algorithms/base_algorithm.py:161-197)
# ----------------------------------------------------------------------------
def update_pressure(p_star, p_prime, alpha_p, conditions):
    p = p_star + alpha_p * p_prime
    nx, ny = p.shape
    for loc in boundary_types(conditions):
        if loc == "left":
            p[0, :] = p[1, :]
        elif loc == "right":
            p[nx - 1, :] = p[nx - 2, :]
        elif loc == "bottom":
            p[:, 0] = p[:, 1]
        elif loc == "top":
            p[:, ny - 1] = p[:, ny - 2]
    return p


# ----------------------------------------------------------------------------
# a16  SIMPLE outer loop (Algorithms/simple.py:78-269)
# ----------------------------------------------------------------------------
class SimpleState:
    def __init__(self, nx, ny, conditions):
        self.p = np.zeros((nx, ny))
        self.u = np.zeros((nx + 1, ny))
        self.v = np.zeros((nx, ny + 1))
        apply_velocity_bc(self.u, self.v, nx, ny, conditions)  # base_algorithm.py:68-93


def make_pressure_solver(kind, **kw):
    """Returns f(nx,ny,dx,dy,u*,v*,d_u,d_v) -> (p', info) for the SIMPLE loop."""
    if kind == "mg":
        cfg = kw["cfg"]
        return lambda nx, ny, dx, dy, us, vs, du, dv: mg_solve(cfg, nx, ny, dx, dy, us, vs, du, dv)
    if kind == "direct":
        def f(nx, ny, dx, dy, us, vs, du, dv):
            b = continuity_rhs(nx, ny, dx, dy, 1.0, us, vs)
            x = direct_solve(b, dx, dy, 1.0, du, dv)
            r = b - apply_A(x, dx, dy, 1.0, du, dv)
            return x, {"rel_norm": np.linalg.norm(r) / np.linalg.norm(b), "field": r}
        return f
    if kind == "jacobi":
        omega, n_iter = kw["omega"], kw["n_iter"]
        def f(nx, ny, dx, dy, us, vs, du, dv):
            b = continuity_rhs(nx, ny, dx, dy, 1.0, us, vs)
            x = jacobi_iterate(np.zeros_like(b), b, dx, dy, 1.0, du, dv, omega, n_iter)
            r = b - apply_A(x, dx, dy, 1.0, du, dv)
            return x, {"rel_norm": np.linalg.norm(r), "field": r}
        return f
    if kind == "rb_sor":
        omega, n_iter = kw["omega"], kw["n_iter"]
        def f(nx, ny, dx, dy, us, vs, du, dv):
            b = continuity_rhs(nx, ny, dx, dy, 1.0, us, vs)
            x = rb_sor(np.zeros_like(b), b, dx, dy, 1.0, du, dv, omega, n_iter)
            r = b - apply_A(x, dx, dy, 1.0, du, dv)
            return x, {"rel_norm": np.linalg.norm(r), "field": r}
        return f
    if kind in ("cg", "bicgstab"):
        tol, maxiter = kw.get("tol", 1e-7), kw.get("maxiter", 1000)
        return lambda nx, ny, dx, dy, us, vs, du, dv: krylov_pressure_solve(
            kind, nx, ny, dx, dy, us, vs, du, dv, tol=tol, maxiter=maxiter)
    raise ValueError(kind)


def simple_solve(nx, ny, reynolds, pressure_solver, n_sweeps=20, alpha_p=0.3, alpha_u=0.7,
                 max_iterations=100, tolerance=0.0, conditions=None, rho=1.0, U=1.0, L=1.0,
                 state=None, callback=None, momentum="jacobi", momentum_tol=1e-8, momentum_maxiter=200,
                 momentum_precondition="ilu"):
    """SimpleSolver.solve (simple.py:114-212) with the deterministic momentum oracle
    (JacobiMatrixMomentumSolver, n fixed sweeps) or, momentum='krylov', MatrixFreeMomentumSolver (a7; its rel_norm is
    the absolute unrelaxed residual norm).  Returns (state, history dict)."""
    conditions = bc_conditions() if conditions is None else conditions
    dx, dy = mesh_spacing(nx, ny, L, L)
    mu = rho * U * L / reynolds  # fluid.py:41
    st = SimpleState(nx, ny, conditions) if state is None else state
    p_star = st.p.copy()
    hist = {"u_rel_norm": [], "v_rel_norm": [], "p_rel_norm": [], "total_rel_norm": []}
    it = 1
    total = 1.0
    while it <= max_iterations and total > tolerance:
        if momentum == "krylov":
            us, du, un, _, ku = solve_momentum_krylov(True, nx, ny, dx, dy, rho, mu, st.u, st.v, p_star, alpha_u, conditions,
                                                      momentum_tol, momentum_maxiter, momentum_precondition)
            vs, dv, vn, _, kv = solve_momentum_krylov(False, nx, ny, dx, dy, rho, mu, st.u, st.v, p_star, alpha_u, conditions,
                                                      momentum_tol, momentum_maxiter, momentum_precondition)
            hist.setdefault("momentum_iterations", []).append((ku, kv))
        else:
            us, du, un, _ = solve_u_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, p_star, alpha_u, conditions, n_sweeps)
            vs, dv, vn, _ = solve_v_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, p_star, alpha_u, conditions, n_sweeps)
        pp, pinfo = pressure_solver(nx, ny, dx, dy, us, vs, du, dv)
        st.p = update_pressure(p_star, pp, alpha_p, conditions)
        p_star = st.p.copy()
        st.u, st.v = correct_velocity(nx, ny, us, vs, pp, du, dv, conditions)
        total = max(un, vn)
        hist["u_rel_norm"].append(un)
        hist["v_rel_norm"].append(vn)
        hist["p_rel_norm"].append(pinfo["rel_norm"])
        hist["total_rel_norm"].append(total)
        if callback is not None:
            callback(it, st, us, vs, du, dv, pp)
        it += 1
    hist["iterations"] = it - 1
    return st, hist


# ----------------------------------------------------------------------------
# "next" row: PISO outer loop (Algorithms/piso.py:41-175)
# ----------------------------------------------------------------------------
def piso_solve(nx, ny, reynolds, pressure_solver, n_sweeps=20, alpha_p=0.3, alpha_u=0.7, n_corrections=2,
               max_iterations=100, tolerance=0.0, conditions=None, rho=1.0, U=1.0, L=1.0):
    """PisoSolver.solve (piso.py:53-135) with the deterministic momentum oracle: predictor with alpha_u, then
    n_corrections x (pressure solve, p update + Neumann copies, velocity correction); between corrections the momentum
    equations are re-solved from the corrected (u, v, p) without relaxation (:92-104).  The norms reported are the
    predictor's (:107-109)."""
    conditions = bc_conditions() if conditions is None else conditions
    dx, dy = mesh_spacing(nx, ny, L, L)
    mu = rho * U * L / reynolds
    st = SimpleState(nx, ny, conditions)
    p_star = st.p.copy()
    hist = {"u_rel_norm": [], "v_rel_norm": [], "p_rel_norm": [], "total_rel_norm": []}
    it, total = 1, 1.0
    while it <= max_iterations and total > tolerance:
        us, du, un, _ = solve_u_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, p_star, alpha_u, conditions, n_sweeps)
        vs, dv, vn, _ = solve_v_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, p_star, alpha_u, conditions, n_sweeps)
        pinfo = None
        for c in range(n_corrections):
            pp, pinfo = pressure_solver(nx, ny, dx, dy, us, vs, du, dv)
            st.p = update_pressure(p_star, pp, alpha_p, conditions)
            p_star = st.p.copy()
            st.u, st.v = correct_velocity(nx, ny, us, vs, pp, du, dv, conditions)
            us, vs = st.u.copy(), st.v.copy()
            if c < n_corrections - 1:
                us, du, _, _ = solve_u_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, st.p, 1, conditions, n_sweeps)
                vs, dv, _, _ = solve_v_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, st.p, 1, conditions, n_sweeps)
        total = max(un, vn)
        hist["u_rel_norm"].append(un)
        hist["v_rel_norm"].append(vn)
        hist["p_rel_norm"].append(pinfo["rel_norm"] if pinfo else 0.0)
        hist["total_rel_norm"].append(total)
        it += 1
    hist["iterations"] = it - 1
    return st, hist


def simpler_solve(nx, ny, reynolds, pressure_solver, n_sweeps=20, alpha_p=0.3, alpha_u=0.7, max_iterations=100,
                  tolerance=0.0, conditions=None, rho=1.0, U=1.0, L=1.0):
    """SimplerSolver.solve (Algorithms/simpler.py:78-190) with the deterministic momentum oracle: predictor; the pressure
    solver's answer from (u*, v*, d) added to p UNRELAXED (p-bar, :124-128); momentum solved again from the same (u, v)
    with the new p (:131-154); pressure correction p', p += alpha_p p' (:162-163), velocity correction with p' (:165-167).
    Norms: the first predictor's; p_rel_norm = ||p - p_old|| / sqrt(nx ny) (:172)."""
    conditions = bc_conditions() if conditions is None else conditions
    dx, dy = mesh_spacing(nx, ny, L, L)
    mu = rho * U * L / reynolds
    st = SimpleState(nx, ny, conditions)
    hist = {"u_rel_norm": [], "v_rel_norm": [], "p_rel_norm": [], "total_rel_norm": []}
    it, total = 1, 1.0
    while it <= max_iterations and total > tolerance:
        p_old = st.p.copy()
        us, du, un, _ = solve_u_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, st.p, alpha_u, conditions, n_sweeps)
        vs, dv, vn, _ = solve_v_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, st.p, alpha_u, conditions, n_sweeps)
        p_bar, _ = pressure_solver(nx, ny, dx, dy, us, vs, du, dv)
        st.p = update_pressure(st.p, p_bar, 1.0, conditions)
        us, du, _, _ = solve_u_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, st.p, alpha_u, conditions, n_sweeps)
        vs, dv, _, _ = solve_v_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, st.p, alpha_u, conditions, n_sweeps)
        pp, _ = pressure_solver(nx, ny, dx, dy, us, vs, du, dv)
        st.p = update_pressure(st.p, pp, alpha_p, conditions)
        st.u, st.v = correct_velocity(nx, ny, us, vs, pp, du, dv, conditions)
        total = max(un, vn)
        hist["u_rel_norm"].append(un)
        hist["v_rel_norm"].append(vn)
        hist["p_rel_norm"].append(float(np.linalg.norm(st.p - p_old) / (np.sqrt(nx * ny) + 1.0e-30)))
        hist["total_rel_norm"].append(total)
        it += 1
    hist["iterations"] = it - 1
    return st, hist


def simplec_solve(nx, ny, reynolds, pressure_solver, n_sweeps=20, alpha_p=0.2, alpha_u=0.7, max_iterations=100,
                  tolerance=0.0, conditions=None, rho=1.0, U=1.0, L=1.0):
    """SimplecSolver.solve (Algorithms/simplec.py:47-283) AS CODED, with the deterministic momentum oracle:
    momentum predictor; d_u, d_v divided by 1 - (1 - alpha_u) (:126-127); pressure solve with the scaled d (:130-137);
    5-point smoothing of p' with a zero boundary ring (:141-147); p = p* + alpha_p p' WITHOUT the zero-gradient edge copies
    (:154; the call at :139 acts on the array that :154 replaces); velocity correction with the smoothed p' and the scaled d
    (:162-166).  Residuals are infinity norms: momentum max|u* - u|, pressure max|p - p_old|, total max|u - u_old| (:119-122,
    :157, :169-171); the loop runs while total > tolerance.  The adaptive alpha_p (:150-153) compares the previous total
    residual with itself (max_res IS residual_history[-1] at that point) and therefore never fires."""
    conditions = bc_conditions() if conditions is None else conditions
    dx, dy = mesh_spacing(nx, ny, L, L)
    mu = rho * U * L / reynolds
    st = SimpleState(nx, ny, conditions)
    p_star = st.p.copy()
    hist = {"total": [], "momentum": [], "pressure": []}
    it, max_res = 1, 1000.0
    div = 1 - (1 - alpha_u)
    while it <= max_iterations and max_res > tolerance:
        u_old, v_old, p_old = st.u.copy(), st.v.copy(), st.p.copy()
        us, du, _, _ = solve_u_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, p_star, alpha_u, conditions, n_sweeps)
        vs, dv, _, _ = solve_v_momentum(nx, ny, dx, dy, rho, mu, st.u, st.v, p_star, alpha_u, conditions, n_sweeps)
        momentum_res = max(np.max(np.abs(us - st.u)), np.max(np.abs(vs - st.v)))
        with np.errstate(invalid="ignore", divide="ignore"):
            du_c, dv_c = du / div, dv / div
        pp, _ = pressure_solver(nx, ny, dx, dy, us, vs, du_c, dv_c)
        sm = np.zeros_like(pp)
        sm[1:-1, 1:-1] = 0.6 * pp[1:-1, 1:-1] + 0.1 * (pp[2:, 1:-1] + pp[:-2, 1:-1] + pp[1:-1, 2:] + pp[1:-1, :-2])
        pp = sm
        st.p = p_star + alpha_p * pp
        pressure_res = np.max(np.abs(st.p - p_old))
        p_star = st.p.copy()
        st.u, st.v = correct_velocity(nx, ny, us, vs, pp, du_c, dv_c, conditions)
        max_res = max(np.max(np.abs(st.u - u_old)), np.max(np.abs(st.v - v_old)))
        hist["total"].append(float(max_res))
        hist["momentum"].append(float(momentum_res))
        hist["pressure"].append(float(pressure_res))
        it += 1
    hist["iterations"] = it - 1
    return st, hist


# ----------------------------------------------------------------------------
# a17  Ghia centre-line errors (postprocessing/validation/cavity_flow.py:178-301)
# ----------------------------------------------------------------------------
def ghia_errors(u, v, nx, ny, table):
    """(inf_norm_error, l2_norm_error) against a Ghia table dict with keys x,v,y,u."""
    from scipy.interpolate import interp1d
    dx, dy = mesh_spacing(nx, ny)
    x = np.linspace(dx / 2, 1 - dx / 2, nx)
    y = np.linspace(dy / 2, 1 - dy / 2, ny)
    uc = u[nx // 2, :]
    vc = v[:, ny // 2]
    ui = interp1d(y, uc, kind="cubic", bounds_error=False, fill_value="extrapolate")(np.asarray(table["y"]))
    vi = interp1d(x, vc, kind="cubic", bounds_error=False, fill_value="extrapolate")(np.asarray(table["x"]))
    ue = ui - np.asarray(table["u"])
    ve = vi - np.asarray(table["v"])
    inf = max(np.max(np.abs(ue)), np.max(np.abs(ve)))
    l2 = np.sqrt((np.sum(ue ** 2) + np.sum(ve ** 2)) / (len(ue) + len(ve)))
    return inf, l2


def max_interior_divergence(u, v, dx, dy):
    """get_max_divergence (base_algorithm.py:134-159, cavity_flow.py:147-175)."""
    div = (u[1:, :] - u[:-1, :]) / dx + (v[:, 1:] - v[:, :-1]) / dy
    return np.max(np.abs(div[1:-1, 1:-1]))


# ---------------------------------------------------------------------------------------------------------------------
# Extended-stencil convection schemes: QUICKDiscretization (discretization/quick.py:27-219) and
# SecondOrderUpwindDiscretization (discretization/second_order_upwind.py:26-325).  TEST INFRASTRUCTURE.
# A scheme is a list of accumulation steps (target, face, flux part, coefficient, diffusion term, stencil mask); the steps of
# one target are in the reference's statement order so that every sum rounds identically.
# ---------------------------------------------------------------------------------------------------------------------
EXT_KEYS = ("a_e", "a_w", "a_n", "a_s", "a_ee", "a_ww", "a_nn", "a_ss", "a_p", "source")


def _quick_face(first, second, opposite, face, diff, mask):
    return [(first, face, "pos", 0.75, diff, mask), ("a_p", face, "pos", 0.375, None, mask),
            (second, face, "pos", -0.125, None, mask), (first, face, "neg", 0.375, diff, mask),
            ("a_p", face, "neg", 0.75, None, mask), (opposite, face, "neg", -0.125, None, mask)]


_EXT_STEPS = {
    # quick.py:62-112 (u) / :149-194 (v): the same statements for both components
    ("quick", True): (_quick_face("a_e", "a_ee", "a_w", "e", "De", "EE") + _quick_face("a_w", "a_ww", "a_e", "w", "De", "WW") +
                      _quick_face("a_n", "a_nn", "a_s", "n", "Dn", "NN") + _quick_face("a_s", "a_ss", "a_n", "s", "Dn", "SS")),
    # second_order_upwind.py:88-127
    ("sou", True): [("a_e", None, None, 0.0, "De", None), ("a_w", None, None, 0.0, "De", None),
                    ("a_n", None, None, 0.0, "Dn", None), ("a_s", None, None, 0.0, "Dn", None),
                    ("a_p", "e", "pos", 1.5, None, None), ("a_w", "e", "pos", 0.5, None, None), ("a_ww", "e", "pos", -0.5, None, None),
                    ("a_e", "e", "neg", 1.5, None, None), ("a_ee", "e", "neg", 0.5, None, None),
                    ("a_w", "w", "pos", 1.5, None, None), ("a_ww", "w", "pos", -0.5, None, None),
                    ("a_p", "w", "neg", 1.5, None, None), ("a_e", "w", "neg", 0.5, None, None),
                    ("a_p", "n", "pos", 1.5, None, None), ("a_s", "n", "pos", -0.5, None, None),
                    ("a_n", "n", "neg", 1.5, None, None), ("a_nn", "n", "neg", 0.5, None, None),
                    ("a_s", "s", "pos", 1.5, None, None), ("a_ss", "s", "pos", -0.5, None, None),
                    ("a_p", "s", "neg", 1.5, None, None), ("a_n", "s", "neg", 0.5, None, None)],
    # second_order_upwind.py:234-266
    ("sou", False): [("a_e", None, None, 0.0, "De", None), ("a_w", None, None, 0.0, "De", None),
                     ("a_n", None, None, 0.0, "Dn", None), ("a_s", None, None, 0.0, "Dn", None),
                     ("a_e", "e", "pos", 1.5, None, None), ("a_ee", "e", "pos", 0.5, None, None),
                     ("a_p", "e", "neg", 1.5, None, None), ("a_w", "e", "neg", 0.5, None, None),
                     ("a_p", "w", "pos", 1.5, None, None), ("a_e", "w", "pos", 0.5, None, None),
                     ("a_w", "w", "neg", 1.5, None, None), ("a_ww", "w", "neg", 0.5, None, None),
                     ("a_n", "n", "pos", 1.5, None, None), ("a_nn", "n", "pos", 0.5, None, None),
                     ("a_p", "n", "neg", 1.5, None, None), ("a_s", "n", "neg", 0.5, None, None),
                     ("a_p", "s", "pos", 1.5, None, None), ("a_n", "s", "pos", 0.5, None, None),
                     ("a_s", "s", "neg", 1.5, None, None), ("a_ss", "s", "neg", 0.5, None, None)],
}
_EXT_STEPS[("quick", False)] = _EXT_STEPS[("quick", True)]


def ext_links(scheme, is_u, nx, ny, dx, dy, rho, mu, u, v, p, sides):
    """The ten coefficient arrays of calculate_u_coefficients (is_u) / calculate_v_coefficients of the 'quick' / 'sou' scheme;
    sides: bit mask 1 left, 2 right, 4 bottom, 8 top of the boundaries with a registered condition (0 = bc None)."""
    shape = (nx + 1, ny) if is_u else (nx, ny + 1)
    out = {k: np.zeros(shape) for k in EXT_KEYS}
    i = np.arange(1, nx) if is_u else np.arange(1, nx - 1)
    j = np.arange(1, ny - 1) if is_u else np.arange(1, ny)
    I, J = np.meshgrid(i, j, indexing="ij")
    cdy, cdx = 0.5 * rho * dy, 0.5 * rho * dx
    if is_u:
        F = {"e": cdy * (u[I + 1, J] + u[I, J]), "w": cdy * (u[I - 1, J] + u[I, J]),
             "n": cdx * (v[I, J + 1] + v[I - 1, J + 1]), "s": cdx * (v[I, J] + v[I - 1, J])}
    else:
        F = {"e": cdy * (u[I + 1, J] + u[I + 1, J - 1]), "w": cdy * (u[I, J] + u[I, J - 1]),
             "n": cdx * (v[I, J + 1] + v[I, J]), "s": cdx * (v[I, J] + v[I, J - 1])}
    D = {"De": mu * dy / dx, "Dn": mu * dx / dy}
    has = {None: np.ones(I.shape, bool), "EE": I <= (nx - 2 if is_u else nx - 3), "WW": I >= 2,
           "NN": J <= (ny - 3 if is_u else ny - 2), "SS": J >= 2}
    blk = {k: np.zeros(I.shape) for k in EXT_KEYS}
    for dst, face, part, coef, diff, mask in _EXT_STEPS[(scheme, bool(is_u))]:
        if face is None:
            term = np.full(I.shape, D[diff])
        else:
            term = coef * (np.maximum(F[face], 0.0) if part == "pos" else np.maximum(-F[face], 0.0))
            if diff is not None:
                term = term + D[diff]
        blk[dst] = np.where(has[mask], blk[dst] + term, blk[dst])
    blk["source"] = ((p[I - 1, J] - p[I, J]) * dy) if is_u else ((p[I, J - 1] - p[I, J]) * dx)
    if scheme == "sou":
        s = blk["a_e"] + blk["a_w"]
        for k in ("a_n", "a_s", "a_ee", "a_ww", "a_nn", "a_ss"):
            s = s + blk[k]
        s = s + (F["e"] - F["w"])
        s = s + (F["n"] - F["s"])
        blk["a_p"] = blk["a_p"] + s
    for k in EXT_KEYS:
        out[k][I, J] = blk[k]
    S = out["source"]
    if is_u:   # Practice B: left, right, bottom, top (quick.py:199-209)
        if sides & 1: S[1, :] += out["a_w"][1, :] * u[0, :]; out["a_w"][1, :] = 0.0
        if sides & 2: S[nx - 1, :] += out["a_e"][nx - 1, :] * u[nx, :]; out["a_e"][nx - 1, :] = 0.0
        if sides & 4: S[1:nx, 1] += out["a_s"][1:nx, 1] * u[1:nx, 0]; out["a_s"][1:nx, 1] = 0.0
        if sides & 8: S[1:nx, ny - 2] += out["a_n"][1:nx, ny - 2] * u[1:nx, ny - 1]; out["a_n"][1:nx, ny - 2] = 0.0
    else:      # bottom, top, left, right (quick.py:211-219)
        if sides & 4: S[:, 1] += out["a_s"][:, 1] * v[:, 0]; out["a_s"][:, 1] = 0.0
        if sides & 8: S[:, ny - 1] += out["a_n"][:, ny - 1] * v[:, ny]; out["a_n"][:, ny - 1] = 0.0
        if sides & 1: S[1, 1:ny] += out["a_w"][1, 1:ny] * v[0, 1:ny]; out["a_w"][1, 1:ny] = 0.0
        if sides & 2: S[nx - 2, 1:ny] += out["a_e"][nx - 2, 1:ny] * v[nx - 1, 1:ny]; out["a_e"][nx - 2, 1:ny] = 0.0
    return out
