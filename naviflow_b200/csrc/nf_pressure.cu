// nf_pressure.cu -- pressure-correction kernels K5-K8, K16-K18 (fp64, HBM-bound 5-point stencils).
//
// Reference arithmetic (paths relative to /root/reference/naviflow_oo):
//   K5 continuity RHS            pressure_solver/helpers/rhs_construction.py:3-21
//   K6 matrix-free A*p, b-A*p    pressure_solver/helpers/matrix_free.py:6-135
//   K7 weighted Jacobi           pressure_solver/jacobi.py:38-78, 160-203
//   K8 red-black SOR             pressure_solver/gauss_seidel.py:214-305
// Coefficients are recomputed from d_u, d_v in registers (matrix-free; 32 B/cell for A*p instead of
// 56 B/cell with five stored coefficient arrays).
#include "nf_pressure.cuh"

// ---------------------------------------------------------------------------------------------
// K5  b[i,j] = rho*(u*[i,j]*dy - u*[i+1,j]*dy + v*[i,j]*dx - v*[i,j+1]*dx); b[0,0] = 0
// ---------------------------------------------------------------------------------------------
__global__ void k_continuity_rhs(nf_grid g, const double* __restrict__ us, const double* __restrict__ vs,
                                 double* __restrict__ b) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.ny || i >= g.ge) return;
  const size_t k = nf_idx(g, i, j);
  double val = g.rho * (((us[k] * g.dy - us[k + g.ld] * g.dy) + vs[k] * g.dx) - vs[k + 1] * g.dx);
  if (i == 0 && j == 0) val = 0.0;
  b[k] = val;
}

// ---------------------------------------------------------------------------------------------
// K6  out = A p   or   out = b - A p
// ---------------------------------------------------------------------------------------------
template <bool RESIDUAL>
__global__ void k_pressure_apply(nf_grid g, const double* __restrict__ p, const double* __restrict__ b,
                                 const double* __restrict__ d_u, const double* __restrict__ d_v,
                                 double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.ny || i >= g.ge) return;
  const double ap = nf_Ap_cell(g, p, d_u, d_v, i, j);
  const size_t k = nf_idx(g, i, j);
  out[k] = RESIDUAL ? (b[k] - ap) : ap;
}

// ---------------------------------------------------------------------------------------------
// K7  weighted Jacobi.  diag = neighbour-coefficient sums with the whole boundary rows/cols doubled
//     (jacobi.py:52-70; NOT diag(A)), <1e-15 -> 1, diag[0,0]=1.  p_new = p + omega*(b-Ap)/diag.
// ---------------------------------------------------------------------------------------------
__global__ void k_jacobi_diag(nf_grid g, const double* __restrict__ d_u, const double* __restrict__ d_v,
                              double* __restrict__ diag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.ny || i >= g.ge) return;
  diag[nf_idx(g, i, j)] = nf_jacobi_diag_cell(g, d_u, d_v, i, j);
}

__global__ void k_jacobi_iter(nf_grid g, const double* __restrict__ p, const double* __restrict__ b,
                              const double* __restrict__ d_u, const double* __restrict__ d_v,
                              double* __restrict__ pout, double omega) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.ny || i >= g.ge) return;
  const size_t k = nf_idx(g, i, j);
  if (i == 0 && j == 0) { pout[k] = 0.0; return; }
  const double ap = nf_Ap_cell(g, p, d_u, d_v, i, j);
  const double diag = nf_jacobi_diag_cell(g, d_u, d_v, i, j);
  pout[k] = p[k] + omega * (b[k] - ap) / diag;
}

__global__ void k_set_value(double* p, double v) { *p = v; }

// ---------------------------------------------------------------------------------------------
// K8  red-black SOR half sweep.  colour 0: (i+j) even except (0,0); colour 1: (i+j) odd; the
//     reference's black mask also contains (0,0) but re-pins it to 0 after every sweep
//     (gauss_seidel.py:148-151, :305), so (0,0) is simply held at 0.
//     p_new = ((((b+E pE)+W pW)+N pN)+S pS) * (1/aP);  p += omega*(p_new - p)   (:285-302)
// ---------------------------------------------------------------------------------------------
__global__ void k_rbsor_color(nf_grid g, double* __restrict__ p, const double* __restrict__ b,
                              const double* __restrict__ d_u, const double* __restrict__ d_v, double omega,
                              int color) {
  const int jj = blockIdx.x * blockDim.x + threadIdx.x;  // index within the colour: j = 2*jj + off
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.ge) return;
  const int j = 2 * jj + ((i + color) & 1);
  if (j >= g.ny) return;
  const size_t k = nf_idx(g, i, j);
  if (i == 0 && j == 0) { p[k] = 0.0; return; }
  const PCoef c = nf_pcoef(g, d_u, d_v, i, j);
  double aP = c.diag;
  if (aP < 1e-15) aP = 1.0;
  const double inv = 1.0 / aP;
  double acc = b[k];
  acc += (i < g.nx - 1) ? c.e * p[k + g.ld] : 0.0;
  acc += (i > 0) ? c.w * p[k - g.ld] : 0.0;
  acc += (j < g.ny - 1) ? c.n * p[k + 1] : 0.0;
  acc += (j > 0) ? c.s * p[k - 1] : 0.0;
  const double pn = acc * inv;
  const double pc = p[k];
  p[k] = pc + omega * (pn - pc);
}

// ---------------------------------------------------------------------------------------------
// K16  p = p* + alpha_p*p', then zero-gradient copies on the four edges; corners end as the
//      diagonal interior neighbour for every registration order (base_algorithm.py:161-197).
// ---------------------------------------------------------------------------------------------
__global__ void k_update_pressure(nf_grid g, const double* __restrict__ ps, const double* __restrict__ pp,
                                  double alpha, double* __restrict__ p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.ny || i >= g.ge) return;
  int si = i, sj = j;
  if (i == 0) si = 1;
  if (i == g.nx - 1) si = g.nx - 2;
  if (j == 0) sj = 1;
  if (j == g.ny - 1) sj = g.ny - 2;
  const size_t ks = nf_idx(g, si, sj);
  p[nf_idx(g, i, j)] = ps[ks] + alpha * pp[ks];
}

// ---------------------------------------------------------------------------------------------
// K18  reductions
// ---------------------------------------------------------------------------------------------
__global__ void k_sumsq(nf_grid g, const double* __restrict__ x, int interior_only, double* partials,
                        unsigned int* ticket, double* out) {
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    const bool jin = !interior_only || (j > 0 && j < g.ny - 1);
    const int lo = interior_only ? max(g.gb, 1) : g.gb;
    const int hi = interior_only ? min(g.ge, g.nx - 1) : g.ge;
    const int st = gridDim.y * blockDim.y;
    int i = lo + blockIdx.y * blockDim.y + threadIdx.y;
    if (jin) {
      for (; i + 3 * st < hi; i += 4 * st) {  // four loads in flight per thread; summation order unchanged
        const double v0 = x[nf_idx(g, i, j)], v1 = x[nf_idx(g, i + st, j)];
        const double v2 = x[nf_idx(g, i + 2 * st, j)], v3 = x[nf_idx(g, i + 3 * st, j)];
        acc[0] += v0 * v0; acc[0] += v1 * v1; acc[0] += v2 * v2; acc[0] += v3 * v3;
      }
      for (; i < hi; i += st) {
        const double v = x[nf_idx(g, i, j)];
        acc[0] += v * v;
      }
    }
  }
  nf_block_reduce_store<1>(acc, partials, ticket, out);
}

// r = b - A p (stored when r != NULL) fused with sum r^2 (and sum b^2 when WITH_B): the multigrid convergence test
// of multigrid.py:201-240 in one pass over the level (32-40 B/cell instead of 40 + 8 + 8 in three kernels)
template <bool WITH_B>
__global__ void k_residual_norms(nf_grid g, const double* __restrict__ p, const double* __restrict__ b,
                                 const double* __restrict__ d_u, const double* __restrict__ d_v,
                                 double* __restrict__ r, double* partials, unsigned int* ticket, double* out) {
  nf_pdl_entry();
  double acc[WITH_B ? 2 : 1];
  acc[0] = 0.0;
  if (WITH_B) acc[WITH_B ? 1 : 0] = 0.0;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    const int st = gridDim.y * blockDim.y;
    int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
    for (; i + st < g.ge; i += 2 * st) {  // two independent stencils in flight; summation order unchanged
      const size_t k0 = nf_idx(g, i, j), k1 = nf_idx(g, i + st, j);
      const double b0 = b[k0], b1 = b[k1];
      const double r0 = b0 - nf_Ap_cell(g, p, d_u, d_v, i, j);
      const double r1 = b1 - nf_Ap_cell(g, p, d_u, d_v, i + st, j);
      if (r) { r[k0] = r0; r[k1] = r1; }
      acc[0] += r0 * r0;
      acc[0] += r1 * r1;
      if (WITH_B) { acc[WITH_B ? 1 : 0] += b0 * b0; acc[WITH_B ? 1 : 0] += b1 * b1; }
    }
    for (; i < g.ge; i += st) {
      const size_t k = nf_idx(g, i, j);
      const double bv = b[k];
      const double rv = bv - nf_Ap_cell(g, p, d_u, d_v, i, j);
      if (r) r[k] = rv;
      acc[0] += rv * rv;
      if (WITH_B) acc[WITH_B ? 1 : 0] += bv * bv;
    }
  }
  nf_block_reduce_store<(WITH_B ? 2 : 1)>(acc, partials, ticket, out);
}

__global__ void k_dot(nf_grid g, const double* __restrict__ x, const double* __restrict__ y, double* partials,
                      unsigned int* ticket, double* out) {
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    const int st = gridDim.y * blockDim.y;
    int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
    for (; i + st < g.ge; i += 2 * st) {
      const size_t k0 = nf_idx(g, i, j), k1 = nf_idx(g, i + st, j);
      const double a0 = x[k0], b0 = y[k0], a1 = x[k1], b1 = y[k1];
      acc[0] += a0 * b0; acc[0] += a1 * b1;
    }
    for (; i < g.ge; i += st) {
      const size_t k = nf_idx(g, i, j);
      acc[0] += x[k] * y[k];
    }
  }
  nf_block_reduce_store<1>(acc, partials, ticket, out);
}

// max |div| over interior cells (base_algorithm.py:134-159): (u[i+1,j]-u[i,j])/dx + (v[i,j+1]-v[i,j])/dy
__global__ void k_max_abs_div(nf_grid g, const double* __restrict__ u, const double* __restrict__ v,
                              unsigned long long* out_bits) {
  double m = 0.0;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > 0 && j < g.ny - 1) {
    for (int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y; i < g.ge; i += gridDim.y * blockDim.y) {
      if (i > 0 && i < g.nx - 1) {
        const size_t k = nf_idx(g, i, j);
        const double d = (u[k + g.ld] - u[k]) / g.dx + (v[k + 1] - v[k]) / g.dy;
        m = fmax(m, fabs(d));
      }
    }
  }
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
  if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0)
    atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));  // non-negative doubles order as integers
}

// =============================================================================================
// host entry points
// =============================================================================================
int nf_check_grid(nf_ctx* ctx, const nf_grid* g) {
  NF_REQUIRE(ctx, g != nullptr, "grid is NULL");
  NF_REQUIRE(ctx, g->nx >= 3 && g->ny >= 3, "nx, ny must be >= 3");
  NF_REQUIRE(ctx, g->ld >= g->ny + 1, "ld must be >= ny+1");
  NF_REQUIRE(ctx, g->gb >= 0 && g->ge <= g->nx && g->gb <= g->ge, "bad row range");
  return NF_OK;
}

extern "C" int nf_continuity_rhs(nf_ctx* ctx, const nf_grid* g, const double* us, const double* vs, double* b) {
  NF_GRID_OK(ctx, g);
  NfLaunch2D l = nf_launch2d(g->ge - g->gb, g->ny);
  k_continuity_rhs<<<l.grid, l.block, 0, ctx->stream>>>(*g, us, vs, b);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nfi_apply(nf_ctx* ctx, const nf_grid* g, const double* p, const double* d_u, const double* d_v, double* out) {
  NfLaunch2D l = nf_launch2d(g->ge - g->gb, g->ny);
  k_pressure_apply<false><<<l.grid, l.block, 0, ctx->stream>>>(*g, p, nullptr, d_u, d_v, out);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

extern "C" int nf_pressure_apply(nf_ctx* ctx, const nf_grid* g, const double* p, const double* d_u,
                                 const double* d_v, double* out) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, out != p, "out must not alias p");
  return nfi_apply(ctx, g, p, d_u, d_v, out);
}

int nfi_residual(nf_ctx* ctx, const nf_grid* g, const double* p, const double* b, const double* d_u,
                 const double* d_v, double* r) {
  NfLaunch2D l = nf_launch2d(g->ge - g->gb, g->ny);
  k_pressure_apply<true><<<l.grid, l.block, 0, ctx->stream>>>(*g, p, b, d_u, d_v, r);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

extern "C" int nf_pressure_residual(nf_ctx* ctx, const nf_grid* g, const double* p, const double* b,
                                    const double* d_u, const double* d_v, double* r) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, r != p, "r must not alias p");
  return nfi_residual(ctx, g, p, b, d_u, d_v, r);
}

extern "C" int nf_jacobi_diag(nf_ctx* ctx, const nf_grid* g, const double* d_u, const double* d_v,
                              double* diag) {
  NF_GRID_OK(ctx, g);
  NfLaunch2D l = nf_launch2d(g->ge - g->gb, g->ny);
  k_jacobi_diag<<<l.grid, l.block, 0, ctx->stream>>>(*g, d_u, d_v, diag);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nfi_jacobi(nf_ctx* ctx, const nf_grid* g, double* p, double* tmp, const double* b, const double* d_u,
               const double* d_v, double omega, int n_iter) {
  if (g->row0 == 0 && g->gb == 0) {  // p[0,0] = 0 before the first A*p (jacobi.py:160, :166)
    k_set_value<<<1, 1, 0, ctx->stream>>>(p, 0.0);
    NF_LAUNCH_CHECK(ctx);
  }
  NfLaunch2D l = nf_launch2d(g->ge - g->gb, g->ny);
  double* src = p;
  double* dst = tmp;
  for (int it = 0; it < n_iter; ++it) {
    k_jacobi_iter<<<l.grid, l.block, 0, ctx->stream>>>(*g, src, b, d_u, d_v, dst, omega);
    NF_LAUNCH_CHECK(ctx);
    double* t = src; src = dst; dst = t;
  }
  if (src != p) {
    const size_t rows = (size_t)(g->ge - g->gb);
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(p + (size_t)(g->gb - g->row0) * g->ld,
                                       src + (size_t)(g->gb - g->row0) * g->ld,
                                       rows * g->ld * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return NF_OK;
}

extern "C" int nf_jacobi_iterate(nf_ctx* ctx, const nf_grid* g, double* p, double* tmp, const double* b,
                                 const double* d_u, const double* d_v, double omega, int n_iter) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, n_iter >= 0, "n_iter < 0");
  NF_REQUIRE(ctx, tmp != p, "tmp must not alias p");
  return nfi_jacobi(ctx, g, p, tmp, b, d_u, d_v, omega, n_iter);
}

int nfi_rbsor(nf_ctx* ctx, const nf_grid* g, double* p, const double* b, const double* d_u, const double* d_v,
              double omega, int n_sweeps) {
  if (g->row0 == 0 && g->gb == 0) {  // p[0,0] = 0 before the first sweep (gauss_seidel.py:145)
    k_set_value<<<1, 1, 0, ctx->stream>>>(p, 0.0);
    NF_LAUNCH_CHECK(ctx);
  }
  NfLaunch2D l = nf_launch2d(g->ge - g->gb, (g->ny + 1) / 2);
  for (int s = 0; s < n_sweeps; ++s) {
    for (int color = 0; color < 2; ++color) {
      k_rbsor_color<<<l.grid, l.block, 0, ctx->stream>>>(*g, p, b, d_u, d_v, omega, color);
      NF_LAUNCH_CHECK(ctx);
    }
  }
  return NF_OK;
}

extern "C" int nf_rbsor_sweeps(nf_ctx* ctx, const nf_grid* g, double* p, const double* b, const double* d_u,
                               const double* d_v, double omega, int n_sweeps) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, n_sweeps >= 0, "n_sweeps < 0");
  return nfi_rbsor(ctx, g, p, b, d_u, d_v, omega, n_sweeps);
}

extern "C" int nf_update_pressure(nf_ctx* ctx, const nf_grid* g, const double* ps, const double* pp,
                                  double alpha_p, double* p) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, p != ps && p != pp, "p must not alias p_star / p_prime");
  NfLaunch2D l = nf_launch2d(g->ge - g->gb, g->ny);
  k_update_pressure<<<l.grid, l.block, 0, ctx->stream>>>(*g, ps, pp, alpha_p, p);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

// sum r^2 -> out[0], (with_b) sum b^2 -> out[1]   (out: device memory)
int nfi_residual_norms(nf_ctx* ctx, const nf_grid* g, const double* p, const double* b, const double* d_u,
                       const double* d_v, double* r, int with_b, double* out) {
  NfLaunch2D l = nf_launch_reduce(g->ge - g->gb, g->ny);
  if (with_b)
    nf_launch(k_residual_norms<true>, l.grid, l.block, 0, ctx->stream, true, *g, p, b, d_u, d_v, r, ctx->partials, ctx->ticket, out);
  else
    nf_launch(k_residual_norms<false>, l.grid, l.block, 0, ctx->stream, true, *g, p, b, d_u, d_v, r, ctx->partials, ctx->ticket, out);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nfi_sumsq_to(nf_ctx* ctx, const nf_grid* g, const double* x, int interior_only, double* out) {
  NfLaunch2D l = nf_launch_reduce(g->ge - g->gb, g->ny);
  k_sumsq<<<l.grid, l.block, 0, ctx->stream>>>(*g, x, interior_only, ctx->partials, ctx->ticket, out);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nfi_sumsq_dev(nf_ctx* ctx, const nf_grid* g, const double* x, int interior_only, int slot) {
  return nfi_sumsq_to(ctx, g, x, interior_only, ctx->scalars + slot);
}

__global__ void k_fill(double* __restrict__ x, size_t n, double v) {
  nf_pdl_entry();
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    x[k] = v;
}

int nfi_fill(nf_ctx* ctx, double* x, size_t count, double value) {
  if (count == 0) return NF_OK;
  if (value == 0.0) {
    NF_CHECK_CUDA(ctx, cudaMemsetAsync(x, 0, count * sizeof(double), ctx->stream));
    return NF_OK;
  }
  size_t blocks = (count + 255) / 256;
  if (blocks > (size_t)NF_SM_COUNT * 16) blocks = (size_t)NF_SM_COUNT * 16;
  nf_launch(k_fill, (unsigned)blocks, 256, 0, ctx->stream, true, x, count, value);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

extern "C" int nf_norm2(nf_ctx* ctx, const nf_grid* g, const double* x, int interior_only, double* out_host) {
  NF_GRID_OK(ctx, g);
  NF_TRY(nfi_sumsq_dev(ctx, g, x, interior_only, 0));
  double s;
  NF_TRY(nf_read_scalars(ctx, 0, 1, &s));
  *out_host = sqrt(s);
  return NF_OK;
}

extern "C" int nf_dot(nf_ctx* ctx, const nf_grid* g, const double* x, const double* y, double* out_host) {
  NF_GRID_OK(ctx, g);
  NfLaunch2D l = nf_launch_reduce(g->ge - g->gb, g->ny);
  k_dot<<<l.grid, l.block, 0, ctx->stream>>>(*g, x, y, ctx->partials, ctx->ticket, ctx->scalars);
  NF_LAUNCH_CHECK(ctx);
  return nf_read_scalars(ctx, 0, 1, out_host);
}

extern "C" int nf_max_abs_divergence(nf_ctx* ctx, const nf_grid* g, const double* u, const double* v,
                                     double* out_host) {
  NF_GRID_OK(ctx, g);
  NF_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->scalars + 8, 0, sizeof(double), ctx->stream));
  NfLaunch2D l = nf_launch_reduce(g->ge - g->gb, g->ny);
  k_max_abs_div<<<l.grid, l.block, 0, ctx->stream>>>(*g, u, v, (unsigned long long*)(ctx->scalars + 8));
  NF_LAUNCH_CHECK(ctx);
  return nf_read_scalars(ctx, 8, 1, out_host);
}
