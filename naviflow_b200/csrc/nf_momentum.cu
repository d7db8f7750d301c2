// nf_momentum.cu -- momentum predictor and correction kernels K1-K4, K17 (fp64, HBM-bound).
//
// Reference arithmetic (paths relative to /root/reference/naviflow_oo):
//   K1  apply_velocity_boundary_conditions      constructor/boundary_conditions.py:164-260
//   K2  power-law link coefficients             solver/momentum_solver/discretization/power_law.py:19-365
//       under-relaxation, d_u / d_v             solver/momentum_solver/jacobi_matrix_solver.py:186-187, 216-219
//   K3  fixed Jacobi sweeps on the 5-point rows solver/momentum_solver/jacobi_matrix_solver.py:48-151, 196-208
//   K4  relaxed residual + masked norm          solver/momentum_solver/jacobi_matrix_solver.py:221-262
//   K17 velocity correction + BCs               solver/velocity_solver/standard.py:10-69
// Field shapes: u (nx+1, ny), v (nx, ny+1), all with the grid's row pitch.
#include "nf_common.cuh"

// rows / cols of the staggered component arrays
__device__ __host__ __forceinline__ int nf_rows(const nf_grid& g, int is_u) { return g.nx + (is_u ? 1 : 0); }
__device__ __host__ __forceinline__ int nf_cols(const nf_grid& g, int is_u) { return g.ny + (is_u ? 0 : 1); }
// row range of a component array owned by this slab: cell rows [gb,ge), plus face row nx of u on the last slab
__device__ __host__ __forceinline__ int nf_row_end(const nf_grid& g, int is_u) {
  return (is_u && g.ge == g.nx) ? g.nx + 1 : g.ge;
}

// ---------------------------------------------------------------------------------------------
// K1  boundary program: final constant per edge line / corner (NaN = leave untouched)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double nf_bc_pick(const double* edge, const double* corner, int i, int j, int last_i,
                                             int last_j, bool right_active) {
  // returns NaN when (i,j) is not a boundary cell of the program
  const bool L = (i == 0), R = (i == last_i) && right_active, B = (j == 0), T = (j == last_j);
  if (L && B) return corner[0];
  if (L && T) return corner[1];
  if (R && B) return corner[2];
  if (R && T) return corner[3];
  // corner positions of a skipped right edge still belong to the bottom / top lines
  if (L) return edge[0];
  if (R) return edge[1];
  if (B) return (i == last_i) ? corner[2] : edge[2];
  if (T) return (i == last_i) ? corner[3] : edge[3];
  return nan("");
}

__device__ __forceinline__ double nf_bc_u(const nf_bc_program& bc, const nf_grid& g, int i, int j) {
  return nf_bc_pick(bc.u_edge, bc.u_corner, i, j, g.nx, g.ny - 1, true);
}
__device__ __forceinline__ double nf_bc_v(const nf_bc_program& bc, const nf_grid& g, int i, int j) {
  return nf_bc_pick(bc.v_edge, bc.v_corner, i, j, g.nx - 1, g.ny, bc.v_right_row >= 0);
}

__global__ void k_apply_velocity_bc(nf_grid g, nf_bc_program bc, double* __restrict__ u, double* __restrict__ v) {
  // one thread per perimeter index t; handles the four edge lines of u and v
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  // rows (i = t): left/right columns of cells... edges along j (i fixed) and along i (j fixed)
  if (t <= g.ny) {  // i = 0 and i = last rows: j = t
    const int j = t;
    if (j < g.ny) {
      if (g.gb == 0) { const double a = nf_bc_u(bc, g, 0, j); if (!isnan(a)) u[nf_idx(g, 0, j)] = a; }
      if (g.ge == g.nx) { const double a = nf_bc_u(bc, g, g.nx, j); if (!isnan(a)) u[nf_idx(g, g.nx, j)] = a; }
    }
    if (g.gb == 0) { const double a = nf_bc_v(bc, g, 0, j); if (!isnan(a)) v[nf_idx(g, 0, j)] = a; }
    if (g.ge == g.nx) { const double a = nf_bc_v(bc, g, g.nx - 1, j); if (!isnan(a)) v[nf_idx(g, g.nx - 1, j)] = a; }
  }
  const int i = g.gb + t;  // j = 0 and j = last columns: i = gb + t
  if (i < nf_row_end(g, 1)) {
    double a = nf_bc_u(bc, g, i, 0); if (!isnan(a)) u[nf_idx(g, i, 0)] = a;
    a = nf_bc_u(bc, g, i, g.ny - 1); if (!isnan(a)) u[nf_idx(g, i, g.ny - 1)] = a;
  }
  if (i < g.ge) {
    double a = nf_bc_v(bc, g, i, 0); if (!isnan(a)) v[nf_idx(g, i, 0)] = a;
    a = nf_bc_v(bc, g, i, g.ny); if (!isnan(a)) v[nf_idx(g, i, g.ny)] = a;
  }
}

// ---------------------------------------------------------------------------------------------
// K2  power-law scheme.  A(|P|) = max(0, 1 - 0.1|F/D|)^5, 0 where |D| <= 1e-10 (power_law.py:19-44).
//     x^5 is evaluated in double-double so that it is the correctly rounded power (the reference calls
//     libm/SVML pow through NumPy; its last bit depends on the host's math library).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double nf_pow5(double x) {
  const double h2 = x * x;
  const double l2 = __fma_rn(x, x, -h2);
  const double h4 = h2 * h2;
  double l4 = __fma_rn(h2, h2, -h4);
  l4 = __fma_rn(2.0 * h2, l2, l4);
  const double h5 = h4 * x;
  double l5 = __fma_rn(h4, x, -h5);
  l5 = __fma_rn(l4, x, l5);
  return h5 + l5;
}

// a / b for a divisor that is constant in the kernel, with rb = RN(1/b) computed once: q = RN(a*rb),
// r = a - q*b (exact, FMA), result RN(q + r*rb).  By Markstein's theorem this is the correctly rounded quotient
// (the same bits as a / b) for every a when rb is the correctly rounded reciprocal -- 3 fp64 operations instead
// of the ~30-instruction division sequence.
__device__ __forceinline__ double nf_div_const(double a, double b, double rb) {
  const double q = a * rb;
  const double r = __fma_rn(-q, b, a);
  return __fma_rn(r, rb, q);
}

__device__ __forceinline__ double nf_powerlaw(double F, double D, double rD) {
  if (!(fabs(D) > 1e-10)) return 0.0;
  const double pe = 0.1 * fabs(nf_div_const(F, D, rD));
  const double base = fmax(0.0, 1.0 - pe);
  const double r = nf_pow5(base);
  return isnan(r) ? 0.0 : r;
}

struct LinkVals {
  double ae, aw, an, as, ap, src;
};

// u-momentum links at face (i,j), 1 <= i <= nx-1 (power_law.py:89-140) + Practice B (:144-199)
__device__ __forceinline__ LinkVals nf_links_u_cell(const nf_grid& g, const double* __restrict__ u,
                                                    const double* __restrict__ v, const double* __restrict__ p,
                                                    double mu, int sides, int i, int j) {
  LinkVals L;
  const double dx = g.dx, dy = g.dy, rho = g.rho;
  const double De = mu * dy / dx, Dn = mu * dx / dy;
  const double rDe = 1.0 / De, rDn = 1.0 / Dn;
  const size_t k = nf_idx(g, i, j);
  const size_t ld = g.ld;
  const double uc = u[k];
  const double Fe = 0.5 * rho * dy * (u[k + ld] + uc);
  const double Fw = 0.5 * rho * dy * (u[k - ld] + uc);
  L.ae = De * nf_powerlaw(Fe, De, rDe) + fmax(-Fe, 0.0);
  L.aw = De * nf_powerlaw(Fw, De, rDe) + fmax(Fw, 0.0);
  if (j == 0) {  // bottom row (:112-125)
    const double Fn = 0.5 * rho * dx * (v[k + 1] + v[k - ld + 1]);
    L.an = Dn * nf_powerlaw(Fn, Dn, rDn) + fmax(-Fn, 0.0);
    L.as = 0.0;
    L.ap = (((L.ae + L.aw) + L.an) + (Fe - Fw)) + Fn;
  } else if (j == g.ny - 1) {  // top row (:127-140)
    const double Fs = 0.5 * rho * dx * (v[k] + v[k - ld]);
    L.an = 0.0;
    L.as = Dn * nf_powerlaw(Fs, Dn, rDn) + fmax(Fs, 0.0);
    L.ap = (((L.ae + L.aw) + L.as) + (Fe - Fw)) - Fs;
  } else {  // interior (:89-110)
    const double Fn = 0.5 * rho * dx * (v[k + 1] + v[k - ld + 1]);
    const double Fs = 0.5 * rho * dx * (v[k] + v[k - ld]);
    L.an = Dn * nf_powerlaw(Fn, Dn, rDn) + fmax(-Fn, 0.0);
    L.as = Dn * nf_powerlaw(Fs, Dn, rDn) + fmax(Fs, 0.0);
    L.ap = ((((L.ae + L.aw) + L.an) + L.as) + (Fe - Fw)) + (Fn - Fs);
  }
  L.src = (p[k - ld] - p[k]) * dy;
  // Practice B: links to boundary nodes go to the source, a_p unchanged
  if ((sides & 1) && i == 1) { L.src += L.aw * u[k - ld]; L.aw = 0.0; }
  if ((sides & 2) && i == g.nx - 1) { L.src += L.ae * u[k + ld]; L.ae = 0.0; }
  if ((sides & 4) && j == 1) { L.src += L.as * u[k - 1]; L.as = 0.0; }
  if ((sides & 8) && j == g.ny - 2) { L.src += L.an * u[k + 1]; L.an = 0.0; }
  return L;
}

// v-momentum links at face (i,j), 1 <= j <= ny-1 (power_law.py:255-301) + Practice B (:304-355)
__device__ __forceinline__ LinkVals nf_links_v_cell(const nf_grid& g, const double* __restrict__ u,
                                                    const double* __restrict__ v, const double* __restrict__ p,
                                                    double mu, int sides, int i, int j) {
  LinkVals L;
  const double dx = g.dx, dy = g.dy, rho = g.rho;
  const double De = mu * dy / dx, Dn = mu * dx / dy;
  const double rDe = 1.0 / De, rDn = 1.0 / Dn;
  const size_t k = nf_idx(g, i, j);
  const size_t ld = g.ld;
  const double vc = v[k];
  double Fn, Fs;
  if (i == 0 || i == g.nx - 1) {  // boundary columns use (v[i,j+1] + v[i,j]) / (v[i,j-1] + v[i,j])  (:276-278, :291-293)
    Fn = 0.5 * rho * dx * (v[k + 1] + vc);
    Fs = 0.5 * rho * dx * (v[k - 1] + vc);
  } else {  // interior uses (v[i,j] + v[i,j+1]) / (v[i,j-1] + v[i,j])  (:259-260)
    Fn = 0.5 * rho * dx * (vc + v[k + 1]);
    Fs = 0.5 * rho * dx * (v[k - 1] + vc);
  }
  L.an = Dn * nf_powerlaw(Fn, Dn, rDn) + fmax(-Fn, 0.0);
  L.as = Dn * nf_powerlaw(Fs, Dn, rDn) + fmax(Fs, 0.0);
  if (i == 0) {
    const double Fe = 0.5 * rho * dy * (u[k + ld] + u[k + ld - 1]);
    L.ae = De * nf_powerlaw(Fe, De, rDe) + fmax(-Fe, 0.0);
    L.aw = 0.0;
    L.ap = (((L.ae + L.an) + L.as) + Fe) + (Fn - Fs);
  } else if (i == g.nx - 1) {
    const double Fw = 0.5 * rho * dy * (u[k] + u[k - 1]);
    L.ae = 0.0;
    L.aw = De * nf_powerlaw(Fw, De, rDe) + fmax(Fw, 0.0);
    L.ap = (((L.aw + L.an) + L.as) - Fw) + (Fn - Fs);
  } else {
    const double Fe = 0.5 * rho * dy * (u[k + ld] + u[k + ld - 1]);
    const double Fw = 0.5 * rho * dy * (u[k] + u[k - 1]);
    L.ae = De * nf_powerlaw(Fe, De, rDe) + fmax(-Fe, 0.0);
    L.aw = De * nf_powerlaw(Fw, De, rDe) + fmax(Fw, 0.0);
    L.ap = ((((L.ae + L.aw) + L.an) + L.as) + (Fe - Fw)) + (Fn - Fs);
  }
  L.src = (p[k - 1] - p[k]) * dx;
  if ((sides & 4) && j == 1) { L.src += L.as * v[k - 1]; L.as = 0.0; }
  if ((sides & 8) && j == g.ny - 1) { L.src += L.an * v[k + 1]; L.an = 0.0; }
  if ((sides & 1) && i == 1) { L.src += L.aw * v[k - ld]; L.aw = 0.0; }
  if ((sides & 2) && i == g.nx - 2) { L.src += L.ae * v[k + ld]; L.ae = 0.0; }
  return L;
}

// MF = 0: relaxation of JacobiMatrixMomentumSolver (a6).  MF = 1: MatrixFreeMomentumSolver (a7,
// matrix_free_momentum.py:429-430, :448-449): a_P clamped to 1e-12 before the division, source relaxed with the relaxed
// a_P, d = 0 where a_P vanishes; the unrelaxed a_P and source are kept for its residual (:379-400).
template <int IS_U, int MF>
__global__ void k_momentum_links(nf_grid g, const double* __restrict__ u, const double* __restrict__ v,
                                 const double* __restrict__ p, double mu, double alpha, int sides, nf_links out,
                                 double* __restrict__ d, double* __restrict__ ap_un, double* __restrict__ src_un) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= nf_cols(g, IS_U) || i >= nf_row_end(g, IS_U)) return;
  LinkVals L;
  const bool active = IS_U ? (i >= 1 && i <= g.nx - 1) : (j >= 1 && j <= g.ny - 1);
  if (active) {
    L = IS_U ? nf_links_u_cell(g, u, v, p, mu, sides, i, j) : nf_links_v_cell(g, u, v, p, mu, sides, i, j);
  } else {
    L.ae = L.aw = L.an = L.as = L.ap = L.src = 0.0;
  }
  const size_t k = nf_idx(g, i, j);
  const double phi = IS_U ? u[k] : v[k];
  // under-relaxation (jacobi_matrix_solver.py:186-187)
  const double ralpha = 1.0 / alpha;
  double ap_rel, src_rel;
  if (MF) {
    const double apc = (fabs(L.ap) > 1e-12) ? L.ap : 1e-12;
    ap_rel = nf_div_const(apc, alpha, ralpha);
    src_rel = L.src + ((1.0 - alpha) * ap_rel) * phi;
    ap_un[k] = L.ap;
    src_un[k] = L.src;
  } else {
    ap_rel = nf_div_const(L.ap, alpha, ralpha);
    src_rel = L.src + nf_div_const((1.0 - alpha) * L.ap, alpha, ralpha) * phi;
  }
  out.a_e[k] = L.ae;
  out.a_w[k] = L.aw;
  out.a_n[k] = L.an;
  out.a_s[k] = L.as;
  out.a_p[k] = ap_rel;
  out.src[k] = src_rel;
  // d = dy/a_p (u) or dx/a_p (v); NaN where |a_p| <= 1e-12 (:213-219)
  d[k] = (fabs(ap_rel) > 1e-12) ? ((IS_U ? g.dy : g.dx) / ap_rel) : (MF ? 0.0 : nan(""));
}

// a7 residual of the UNRELAXED system at the solution (matrix_free_momentum.py:379-400): r = S - A x with identity
// boundary rows, boundary and boundary-adjacent lines (in the component's normal direction) zeroed; sum r^2 -> out[0]
template <int IS_U>
__global__ void k_momentum_residual_unrelaxed(nf_grid g, nf_links L, const double* __restrict__ x,
                                              double* __restrict__ field, double* partials, unsigned int* ticket,
                                              double* out) {
  double acc[1] = {0.0};
  const int rows = nf_rows(g, IS_U), cols = nf_cols(g, IS_U);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cols) {
    for (int i = blockIdx.y * blockDim.y + threadIdx.y; i < rows; i += gridDim.y * blockDim.y) {
      const size_t k = nf_idx(g, i, j);
      bool zero = (i == 0 || i == rows - 1 || j == 0 || j == cols - 1);
      if (IS_U) zero = zero || i == 1 || i == rows - 2;
      else zero = zero || j == 1 || j == cols - 2;
      double r = 0.0;
      if (!zero) {
        double ax = L.a_p[k] * x[k];
        ax -= L.a_e[k] * x[k + g.ld];
        ax -= L.a_w[k] * x[k - g.ld];
        ax -= L.a_n[k] * x[k + 1];
        ax -= L.a_s[k] * x[k - 1];
        r = L.src[k] - ax;
      }
      acc[0] += r * r;
      if (field) field[k] = r;
    }
  }
  nf_block_reduce_store<1>(acc, partials, ticket, out);
}

// ---------------------------------------------------------------------------------------------
// K3  Jacobi sweep  x_new = D^-1 (b - (A-D) x), D^-1 = 0 where |a_p| <= 1e-12.  The off-diagonal sum
//     is accumulated in the CSR column order W, S, N, E the reference's sparse mat-vec uses.
// ---------------------------------------------------------------------------------------------
template <int IS_U>
__global__ void k_momentum_jacobi(nf_grid g, nf_links L, const double* __restrict__ x, double* __restrict__ xn) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  const int rows = nf_rows(g, IS_U), cols = nf_cols(g, IS_U);
  if (j >= cols || i >= nf_row_end(g, IS_U)) return;
  const size_t k = nf_idx(g, i, j);
  double acc = 0.0;
  if (i > 0) acc += (-L.a_w[k]) * x[k - g.ld];
  if (j > 0) acc += (-L.a_s[k]) * x[k - 1];
  if (j < cols - 1) acc += (-L.a_n[k]) * x[k + 1];
  if (i < rows - 1) acc += (-L.a_e[k]) * x[k + g.ld];
  const double ap = L.a_p[k];
  const double dinv = (fabs(ap) > 1e-12) ? 1.0 / ap : 0.0;
  xn[k] = dinv * (L.src[k] - acc);
}

// ---------------------------------------------------------------------------------------------
// K4  r = b - A x (diagonal term between S and N, sorted CSR order); masked norms; returned field has
//     the boundary and boundary-adjacent lines in the component's normal direction zeroed (:252-262).
// ---------------------------------------------------------------------------------------------
template <int IS_U>
__global__ void k_momentum_residual(nf_grid g, nf_links L, const double* __restrict__ x, double* __restrict__ field,
                                    double* partials, unsigned int* ticket, double* out) {
  double acc2[2] = {0.0, 0.0};
  const int rows = nf_rows(g, IS_U), cols = nf_cols(g, IS_U);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cols) {
    for (int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y; i < nf_row_end(g, IS_U); i += gridDim.y * blockDim.y) {
      const size_t k = nf_idx(g, i, j);
      double acc = 0.0;
      if (i > 0) acc += (-L.a_w[k]) * x[k - g.ld];
      if (j > 0) acc += (-L.a_s[k]) * x[k - 1];
      acc += L.a_p[k] * x[k];
      if (j < cols - 1) acc += (-L.a_n[k]) * x[k + 1];
      if (i < rows - 1) acc += (-L.a_e[k]) * x[k + g.ld];
      const double b = L.src[k];
      const double r = b - acc;
      const bool edge = (i == 0 || i == rows - 1 || j == 0 || j == cols - 1);
      if (!edge) { acc2[0] += r * r; acc2[1] += b * b; }
      bool zero;
      if (IS_U) zero = (i == 0 || i == 1 || i == g.nx - 1 || i == g.nx);
      else zero = (j == 0 || j == 1 || j == g.ny - 1 || j == g.ny);
      if (field) field[k] = zero ? 0.0 : r;
    }
  }
  nf_block_reduce_store<2>(acc2, partials, ticket, out);
}

// ---------------------------------------------------------------------------------------------
// K17 velocity correction (standard.py:43-54) fused with the boundary program (:67)
// ---------------------------------------------------------------------------------------------
__global__ void k_correct_velocity(nf_grid g, nf_bc_program bc, const double* __restrict__ us,
                                   const double* __restrict__ vs, const double* __restrict__ pp,
                                   const double* __restrict__ d_u, const double* __restrict__ d_v,
                                   double* __restrict__ u, double* __restrict__ v) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j > g.ny || i >= nf_row_end(g, 1)) return;
  const size_t k = nf_idx(g, i, j);
  if (j < g.ny) {  // u (nx+1, ny)
    double val = us[k];
    if (i >= 1 && i <= g.nx - 1 && j >= 1 && j <= g.ny - 2) val = val + d_u[k] * (pp[k - g.ld] - pp[k]);
    const double a = nf_bc_u(bc, g, i, j);
    if (!isnan(a)) val = a;
    u[k] = val;
  }
  if (i < g.ge) {  // v (nx, ny+1)
    double val = vs[k];
    if (i >= 1 && i <= g.nx - 2 && j >= 1 && j <= g.ny - 1) val = val + d_v[k] * (pp[k - 1] - pp[k]);
    const double a = nf_bc_v(bc, g, i, j);
    if (!isnan(a)) val = a;
    v[k] = val;
  }
}

// =============================================================================================
// host entry points
// =============================================================================================
int nfi_apply_velocity_bc(nf_ctx* ctx, const nf_grid* g, const nf_bc_program* bc, double* u, double* v) {
  int n = g->ny + 1;
  if (g->ge - g->gb + 1 > n) n = g->ge - g->gb + 1;
  k_apply_velocity_bc<<<(n + 127) / 128, 128, 0, ctx->stream>>>(*g, *bc, u, v);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

extern "C" int nf_apply_velocity_bc(nf_ctx* ctx, const nf_grid* g, const nf_bc_program* bc, double* u, double* v) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, bc && u && v, "NULL argument");
  return nfi_apply_velocity_bc(ctx, g, bc, u, v);
}

int nfi_momentum_links(nf_ctx* ctx, const nf_grid* g, int is_u, const double* u, const double* v, const double* p,
                       double mu, double alpha, int sides, nf_links out, double* d) {
  NfLaunch2D l = nf_launch2d(nf_row_end(*g, is_u) - g->gb, nf_cols(*g, is_u));
  if (is_u) k_momentum_links<1, 0><<<l.grid, l.block, 0, ctx->stream>>>(*g, u, v, p, mu, alpha, sides, out, d, nullptr, nullptr);
  else k_momentum_links<0, 0><<<l.grid, l.block, 0, ctx->stream>>>(*g, u, v, p, mu, alpha, sides, out, d, nullptr, nullptr);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

// a7 variant of the coefficients (see k_momentum_links<.., 1>)
int nfi_momentum_links_mf(nf_ctx* ctx, const nf_grid* g, int is_u, const double* u, const double* v, const double* p,
                          double mu, double alpha, int sides, nf_links out, double* d, double* ap_un, double* src_un) {
  NfLaunch2D l = nf_launch2d(nf_row_end(*g, is_u) - g->gb, nf_cols(*g, is_u));
  if (is_u) k_momentum_links<1, 1><<<l.grid, l.block, 0, ctx->stream>>>(*g, u, v, p, mu, alpha, sides, out, d, ap_un, src_un);
  else k_momentum_links<0, 1><<<l.grid, l.block, 0, ctx->stream>>>(*g, u, v, p, mu, alpha, sides, out, d, ap_un, src_un);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

extern "C" int nf_momentum_links_mf(nf_ctx* ctx, const nf_grid* g, int is_u, const double* u_bc, const double* v_bc,
                                    const double* p, double mu, double alpha, int sides, nf_links out, double* d,
                                    double* ap_unrelaxed, double* src_unrelaxed) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, out.a_e && out.a_w && out.a_n && out.a_s && out.a_p && out.src && d && ap_unrelaxed && src_unrelaxed,
             "NULL array");
  NF_REQUIRE(ctx, alpha > 0.0, "relaxation factor must be > 0");
  NF_REQUIRE(ctx, g->row0 == 0 && g->gb == 0 && g->ge == g->nx, "single-slab grids only");
  return nfi_momentum_links_mf(ctx, g, is_u, u_bc, v_bc, p, mu, alpha, sides, out, d, ap_unrelaxed, src_unrelaxed);
}

// The same unrelaxed residual from the RELAXED links of the Jacobi-sweep predictor (jacobi_matrix_solver.py:186-187, :213-219:
// a_P <- a_P / alpha, S <- S + (1 - alpha) (a_P / alpha) phi_old): a_P,un = alpha a_P,rel and S,un = S,rel - (1 - alpha) a_P,rel
// phi_old, masks as matrix_free_momentum.py:380-400.  A convergence diagnostic of the OUTER loop (the relaxed inner
// residual the stopping test uses says nothing about it); rows [g.gb, row_end) of a slab.
template <int IS_U>
__global__ void k_momentum_unrelaxed_from_relaxed(nf_grid g, nf_links L, const double* __restrict__ x,
                                                  const double* __restrict__ phi_old, double alpha, double* partials,
                                                  unsigned int* ticket, double* out) {
  double acc[1] = {0.0};
  const int rows = nf_rows(g, IS_U), cols = nf_cols(g, IS_U);
  const int i_end = nf_row_end(g, IS_U);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cols) {
    for (int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y; i < i_end; i += gridDim.y * blockDim.y) {
      bool zero = (i == 0 || i == rows - 1 || j == 0 || j == cols - 1);
      if (IS_U) zero = zero || i == 1 || i == rows - 2;
      else zero = zero || j == 1 || j == cols - 2;
      if (zero) continue;
      const size_t k = nf_idx(g, i, j);
      const double ap_rel = L.a_p[k];
      double ax = (alpha * ap_rel) * x[k];
      ax -= L.a_e[k] * x[k + g.ld];
      ax -= L.a_w[k] * x[k - g.ld];
      ax -= L.a_n[k] * x[k + 1];
      ax -= L.a_s[k] * x[k - 1];
      const double r = (L.src[k] - (1.0 - alpha) * ap_rel * phi_old[k]) - ax;
      acc[0] += r * r;
    }
  }
  nf_block_reduce_store<1>(acc, partials, ticket, out);
}

int nfi_momentum_unrelaxed_from_relaxed(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, const double* x,
                                        const double* phi_old, double alpha, double* out) {
  NfLaunch2D l = nf_launch_reduce(nf_row_end(*g, is_u) - g->gb, nf_cols(*g, is_u));
  if (is_u) k_momentum_unrelaxed_from_relaxed<1><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, x, phi_old, alpha, ctx->partials, ctx->ticket, out);
  else k_momentum_unrelaxed_from_relaxed<0><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, x, phi_old, alpha, ctx->partials, ctx->ticket, out);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

// sum r^2 of the unrelaxed residual -> out[0] (device); L.a_p / L.src hold the UNRELAXED a_P and source
int nfi_momentum_residual_unrelaxed(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, const double* x, double* field,
                                    double* out) {
  NfLaunch2D l = nf_launch_reduce(nf_rows(*g, is_u), nf_cols(*g, is_u));
  if (is_u) k_momentum_residual_unrelaxed<1><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, x, field, ctx->partials, ctx->ticket, out);
  else k_momentum_residual_unrelaxed<0><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, x, field, ctx->partials, ctx->ticket, out);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

extern "C" int nf_momentum_residual_unrelaxed(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L_unrelaxed, const double* x,
                                              double* field_out, double* norm_host) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, L_unrelaxed.a_e && L_unrelaxed.a_w && L_unrelaxed.a_n && L_unrelaxed.a_s && L_unrelaxed.a_p &&
                      L_unrelaxed.src && x, "NULL array");
  NF_REQUIRE(ctx, g->row0 == 0 && g->gb == 0 && g->ge == g->nx, "single-slab grids only");
  NF_TRY(nfi_momentum_residual_unrelaxed(ctx, g, is_u, L_unrelaxed, x, field_out, ctx->scalars));
  double s[1];
  NF_TRY(nf_read_scalars(ctx, 0, 1, s));
  if (norm_host) *norm_host = sqrt(s[0]);
  return NF_OK;
}

static int check_links(nf_ctx* ctx, const nf_links& L) {
  NF_REQUIRE(ctx, L.a_e && L.a_w && L.a_n && L.a_s && L.a_p && L.src, "NULL link array");
  return NF_OK;
}

extern "C" int nf_momentum_links_u(nf_ctx* ctx, const nf_grid* g, const double* u, const double* v, const double* p,
                                   double mu, double alpha, int sides, nf_links out, double* d_u) {
  NF_GRID_OK(ctx, g);
  NF_TRY(check_links(ctx, out));
  NF_REQUIRE(ctx, alpha > 0.0, "relaxation factor must be > 0");
  return nfi_momentum_links(ctx, g, 1, u, v, p, mu, alpha, sides, out, d_u);
}

extern "C" int nf_momentum_links_v(nf_ctx* ctx, const nf_grid* g, const double* u, const double* v, const double* p,
                                   double mu, double alpha, int sides, nf_links out, double* d_v) {
  NF_GRID_OK(ctx, g);
  NF_TRY(check_links(ctx, out));
  NF_REQUIRE(ctx, alpha > 0.0, "relaxation factor must be > 0");
  return nfi_momentum_links(ctx, g, 0, u, v, p, mu, alpha, sides, out, d_v);
}

int nfi_momentum_jacobi(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, double* x, double* tmp, int n_sweeps) {
  NfLaunch2D l = nf_launch2d(nf_row_end(*g, is_u) - g->gb, nf_cols(*g, is_u));
  double* src = x;
  double* dst = tmp;
  for (int s = 0; s < n_sweeps; ++s) {
    if (is_u) k_momentum_jacobi<1><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, src, dst);
    else k_momentum_jacobi<0><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, src, dst);
    NF_LAUNCH_CHECK(ctx);
    double* t = src; src = dst; dst = t;
  }
  if (src != x) {
    const size_t rows = (size_t)(nf_row_end(*g, is_u) - g->gb);
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(x + (size_t)(g->gb - g->row0) * g->ld, src + (size_t)(g->gb - g->row0) * g->ld,
                                       rows * g->ld * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return NF_OK;
}

// one sweep src -> dst over the rows [g.gb, g.ge) of the grid descriptor (slab runs pass a grown range)
int nfi_momentum_sweep(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, const double* src, double* dst) {
  NfLaunch2D l = nf_launch2d(nf_row_end(*g, is_u) - g->gb, nf_cols(*g, is_u));
  if (is_u) k_momentum_jacobi<1><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, src, dst);
  else k_momentum_jacobi<0><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, src, dst);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

extern "C" int nf_momentum_jacobi(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, double* x, double* tmp,
                                  int n_sweeps) {
  NF_GRID_OK(ctx, g);
  NF_TRY(check_links(ctx, L));
  NF_REQUIRE(ctx, n_sweeps >= 0 && x && tmp && x != tmp, "bad sweep arguments");
  return nfi_momentum_jacobi(ctx, g, is_u, L, x, tmp, n_sweeps);
}

// sums -> out[0..1] (device memory) = (sum r^2, sum b^2) over the masked interior of the rows [g.gb, g.ge)
int nfi_momentum_residual_to(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, const double* x, double* field,
                             double* out) {
  NfLaunch2D l = nf_launch_reduce(nf_row_end(*g, is_u) - g->gb, nf_cols(*g, is_u));
  if (is_u) k_momentum_residual<1><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, x, field, ctx->partials, ctx->ticket, out);
  else k_momentum_residual<0><<<l.grid, l.block, 0, ctx->stream>>>(*g, L, x, field, ctx->partials, ctx->ticket, out);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nfi_momentum_residual_dev(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, const double* x, double* field,
                              int slot) {
  return nfi_momentum_residual_to(ctx, g, is_u, L, x, field, ctx->scalars + slot);
}

extern "C" int nf_momentum_residual(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, const double* x,
                                    double* field_out, double* rel_norm_host) {
  NF_GRID_OK(ctx, g);
  NF_TRY(check_links(ctx, L));
  NF_TRY(nfi_momentum_residual_dev(ctx, g, is_u, L, x, field_out, 0));
  double s[2];
  NF_TRY(nf_read_scalars(ctx, 0, 2, s));
  if (rel_norm_host) *rel_norm_host = sqrt(s[0]) / (sqrt(s[1]) + 1e-15);
  return NF_OK;
}

int nfi_correct_velocity(nf_ctx* ctx, const nf_grid* g, const nf_bc_program* bc, const double* us, const double* vs,
                         const double* pp, const double* d_u, const double* d_v, double* u, double* v) {
  NfLaunch2D l = nf_launch2d(nf_row_end(*g, 1) - g->gb, g->ny + 1);
  k_correct_velocity<<<l.grid, l.block, 0, ctx->stream>>>(*g, *bc, us, vs, pp, d_u, d_v, u, v);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

extern "C" int nf_correct_velocity(nf_ctx* ctx, const nf_grid* g, const nf_bc_program* bc, const double* us,
                                   const double* vs, const double* pp, const double* d_u, const double* d_v,
                                   double* u, double* v) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, bc && us && vs && pp && d_u && d_v && u && v, "NULL argument");
  return nfi_correct_velocity(ctx, g, bc, us, vs, pp, d_u, d_v, u, v);
}
