// nf_transfer.cu -- multigrid transfer operators K9-K12 (fp64, HBM-bound gathers).
//
// Reference arithmetic (paths relative to /root/reference/naviflow_oo):
//   K9  restrict_full_weighting / restrict_inject   pressure_solver/helpers/multigrid_helpers.py:8-70
//   K10 restrict_coefficients                       pressure_solver/helpers/multigrid_helpers.py:196-329
//   K11 interpolate_linear                          pressure_solver/helpers/multigrid_helpers.py:73-192
//   K12 interpolate_cubic (separable not-a-knot spline, applied as a banded operator P C P^T)
//                                                   pressure_solver/helpers/multigrid_helpers.py:333-391
// The index rules are the reference's own (coarse k <-> fine 2k+1, the (nf-1)//2 coarse size, the
// untouched last rows/cols on even-sized fine grids); they are reproduced, not "cleaned up".
#include "nf_pressure.cuh"

// ---------------------------------------------------------------------------------------------
// K9  full weighting: c[I,J] = f[2I+1,2J+1]/4 + (N+S+E+W)/8 + (NE+NW+SE+SW)/16   (:63-68)
//     with N=(2I+1,2J+2) S=(2I+1,2J) E=(2I+2,2J+1) W=(2I,2J+1) NE=(2I+2,2J+2) NW=(2I,2J+2)
//     SE=(2I+2,2J) SW=(2I,2J); coarse size (nf-1)//2.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double nf_fw_cell(const nf_grid& gf, const double* __restrict__ f, int I, int J) {
  const size_t k = nf_idx(gf, 2 * I + 1, 2 * J + 1);
  const size_t ld = gf.ld;
  const double c = f[k];
  const double n = f[k + 1], s = f[k - 1], e = f[k + ld], w = f[k - ld];
  const double ne = f[k + ld + 1], nw = f[k - ld + 1], se = f[k + ld - 1], sw = f[k - ld - 1];
  return (c / 4.0 + (((n + s) + e) + w) / 8.0) + (((ne + nw) + se) + sw) / 16.0;
}

__global__ void k_restrict_fw(nf_grid gf, const double* __restrict__ f, nf_grid gc, double* __restrict__ c) {
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  const int I = gc.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (J >= gc.ny || I >= gc.ge) return;
  c[nf_idx(gc, I, J)] = nf_fw_cell(gf, f, I, J);
}

// fused: c = FW(b - A p) without materialising the fine residual in HBM.  A CTA owns CR x CC coarse cells: it
// evaluates the fine residual once per fine cell of the (2CR+1) x (2CC+1) patch into shared memory, then applies
// the full-weighting stencil from there (34 B/fine cell of HBM traffic instead of 40 + 8 + 8 + 2).
constexpr int RR_CR = 8, RR_CC = 64;
__global__ void __launch_bounds__(256)
k_residual_restrict_fw(nf_grid gf, const double* __restrict__ p, const double* __restrict__ b,
                       const double* __restrict__ d_u, const double* __restrict__ d_v, nf_grid gc,
                       double* __restrict__ c, double* __restrict__ x0) {
  nf_pdl_entry();
  __shared__ double sR[2 * RR_CR + 1][2 * RR_CC + 2];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 128 x 2
  const int I0 = gc.gb + blockIdx.y * RR_CR, J0 = blockIdx.x * RR_CC;
  const int fi0 = 2 * I0, fj0 = 2 * J0;
  for (int fr = ty; fr < 2 * RR_CR + 1; fr += 2) {
    const int i = fi0 + fr;
    for (int fc = tx; fc < 2 * RR_CC + 1; fc += 128) {
      const int j = fj0 + fc;
      double rv = 0.0;
      // fine rows 2I..2I+2 of the coarse rows this launch writes; the last tile's overshoot is neither needed nor
      // (on a slab) stored
      if (i < gf.nx && j < gf.ny && i <= 2 * gc.ge && (i == 0 || nf_row_stored(gf, i - 1)) && nf_row_stored(gf, i + 1))
        rv = b[nf_idx(gf, i, j)] - nf_Ap_cell(gf, p, d_u, d_v, i, j);
      sR[fr][fc] = rv;
    }
  }
  __syncthreads();
  for (int t = ty * 128 + tx; t < RR_CR * RR_CC; t += 256) {
    const int ci = t / RR_CC, cj = t % RR_CC;
    const int I = I0 + ci, J = J0 + cj;
    if (I >= gc.ge || J >= gc.ny) continue;
    const int a = 2 * ci, q = 2 * cj;
    const double cc = sR[a + 1][q + 1], n = sR[a + 1][q + 2], s = sR[a + 1][q], e = sR[a + 2][q + 1], w = sR[a][q + 1];
    const double ne = sR[a + 2][q + 2], nw = sR[a][q + 2], se = sR[a + 2][q], sw = sR[a][q];
    c[nf_idx(gc, I, J)] = (cc / 4.0 + (((n + s) + e) + w) / 8.0) + (((ne + nw) + se) + sw) / 16.0;
    if (x0) x0[nf_idx(gc, I, J)] = 0.0;  // the coarse iterate starts from zero (saves the fill launch)
  }
}

// injection: c[I,J] = f[2I+1,2J+1]; coarse size nf//2  (:20)
__global__ void k_restrict_inject(nf_grid gf, const double* __restrict__ f, nf_grid gc, double* __restrict__ c) {
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  const int I = gc.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (J >= gc.ny || I >= gc.ge) return;
  c[nf_idx(gc, I, J)] = f[nf_idx(gf, 2 * I + 1, 2 * J + 1)];
}

// ---------------------------------------------------------------------------------------------
// K10 coefficient coarsening (:229-327): harmonic mean of the two fine faces (arithmetic mean unless
//     both are > 0), boundary faces copied, everything x0.25; entries the reference never writes are 0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double nf_hmean(double d1, double d2) {
  if (d1 > 0.0 && d2 > 0.0) return 2.0 / (1.0 / d1 + 1.0 / d2);
  return 0.5 * (d1 + d2);
}

__global__ void k_restrict_coeffs(nf_grid gf, const double* __restrict__ d_u, const double* __restrict__ d_v,
                                  nf_grid gc, double* __restrict__ duc, double* __restrict__ dvc) {
  // thread (I,J) over the (nxc+1) x (nyc+1) index box: writes duc[I,J] (J<nyc) and dvc[I,J] (I<nxc)
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  const int I = gc.gb + blockIdx.y * blockDim.y + threadIdx.y;
  const int nxc = gc.nx, nyc = gc.ny;
  const int i_end = (gc.ge == nxc) ? nxc + 1 : gc.ge;  // the rank owning the last cell row also owns face row nxc
  if (J > nyc || I >= i_end) return;
  if (J < nyc) {  // d_u^c (nxc+1, nyc)
    double val = 0.0;
    if (2 * J < gf.ny) {
      if (I == 0) val = d_u[nf_idx(gf, 0, 2 * J)];
      else if (I == nxc) val = d_u[nf_idx(gf, gf.nx, 2 * J)];
      else if (2 * I < gf.nx) val = nf_hmean(d_u[nf_idx(gf, 2 * I, 2 * J)], d_u[nf_idx(gf, 2 * I + 1, 2 * J)]);
    }
    duc[nf_idx(gc, I, J)] = val * 0.25;
  }
  if (I < nxc) {  // d_v^c (nxc, nyc+1)
    double val = 0.0;
    if (2 * I < gf.nx) {
      if (J == 0) val = d_v[nf_idx(gf, 2 * I, 0)];
      else if (J == nyc) val = d_v[nf_idx(gf, 2 * I, gf.ny)];
      else if (2 * J < gf.ny) val = nf_hmean(d_v[nf_idx(gf, 2 * I, 2 * J)], d_v[nf_idx(gf, 2 * I, 2 * J + 1)]);
    }
    dvc[nf_idx(gc, I, J)] = val * 0.25;
  }
}

// ---------------------------------------------------------------------------------------------
// K11 bilinear prolongation (:116-186) as a gather.  Interior value V(i,j), 1<=i,j<=m-2:
//       i odd  -> coarse I=(i-1)/2 (needs I<mc);  i even -> between I=(i-2)/2 and I+1 (needs I<=mc-2)
//     points outside those rules stay 0 (last two rows/cols of an even-sized fine grid).
//     Ring cells copy ring 1: f[i,j] = V(clamp(i,1,m-2), clamp(j,1,m-2))  (:170-186).
//     m <= 3: only the coincident points, no ring copy (:127-128).
// ---------------------------------------------------------------------------------------------
// nf_prolong_linear_value: nf_pressure.cuh (shared with the multigrid tail kernel)

// Bulk of the fine grid: one thread per coarse cell (I,J) writes the 2x2 fine block (2I+1..2I+2, 2J+1..2J+2) from
// c[I..I+1][J..J+1] (each coarse value loaded once per thread, no per-cell index logic).  The thin strips the
// block rule does not cover (ring rows/cols, trailing rows/cols of even-sized grids) are done by extra CTAs of the
// same launch through the generic gather nf_prolong_linear_value.
template <bool ADD>
__global__ void __launch_bounds__(256)
k_prolong_linear(nf_grid gc, const double* __restrict__ c, nf_grid gf, double* __restrict__ f, int nI, int nJ,
                 int nby_fast, int I_lo, int I_hi) {
  nf_pdl_entry();
  if ((int)blockIdx.y < nby_fast) {
    const int J = blockIdx.x * 32 + threadIdx.x;
    const int I = I_lo + blockIdx.y * 8 + threadIdx.y;
    if (I >= I_hi || J >= nJ) return;
    const size_t k = nf_idx(gc, I, J);
    const double c00 = c[k], c01 = c[k + 1], c10 = c[k + gc.ld], c11 = c[k + gc.ld + 1];
    const double v00 = c00;
    const double v01 = 0.5 * (c00 + c01);
    const double v10 = 0.5 * (c00 + c10);
    const double v11 = 0.25 * (((c00 + c10) + c01) + c11);
    const size_t kf = nf_idx(gf, 2 * I + 1, 2 * J + 1);
    const bool r0 = (2 * I + 1 >= gf.gb) && (2 * I + 1 < gf.ge);  // slab runs: only this rank's fine rows
    const bool r1 = (2 * I + 2 >= gf.gb) && (2 * I + 2 < gf.ge);
    if (ADD) {
      if (r0) { f[kf] = f[kf] + v00; f[kf + 1] = f[kf + 1] + v01; }
      if (r1) { f[kf + gf.ld] = f[kf + gf.ld] + v10; f[kf + gf.ld + 1] = f[kf + gf.ld + 1] + v11; }
    } else {
      if (r0) { f[kf] = v00; f[kf + 1] = v01; }
      if (r1) { f[kf + gf.ld] = v10; f[kf + gf.ld + 1] = v11; }
    }
    return;
  }
  // strips: rows {0} u [2nI+1, nx) over all columns, then columns {0} u [2nJ+1, ny) over rows 1..2nI
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const long long t = ((long long)(blockIdx.y - nby_fast) * gridDim.x + blockIdx.x) * 256 + tid;
  const int srows = 1 + (gf.nx - (2 * nI + 1));
  const int scols = 1 + (gf.ny - (2 * nJ + 1));
  int i, j;
  if (t < (long long)srows * gf.ny) {
    const int a = (int)(t / gf.ny);
    j = (int)(t % gf.ny);
    i = (a == 0) ? 0 : 2 * nI + a;
  } else {
    const long long t2 = t - (long long)srows * gf.ny;
    if (t2 >= (long long)scols * (2 * nI)) return;
    const int a = (int)(t2 % scols);
    i = 1 + (int)(t2 / scols);
    j = (a == 0) ? 0 : 2 * nJ + a;
  }
  if (i < gf.gb || i >= gf.ge) return;
  const double v = nf_prolong_linear_value(gc, c, gf.nx, gf.ny, i, j);
  const size_t k = nf_idx(gf, i, j);
  if (ADD) f[k] = f[k] + v;
  else f[k] = v;
}

// ---------------------------------------------------------------------------------------------
// K12 "cubic" prolongation F = P C P^T with P (m x mc) the 1-D interpolation matrix of the reference's
//     interpolator (not-a-knot cubic spline for mc>=4), stored banded: row i holds W taps starting at
//     coarse column start[i].  Two passes: T = P C  (m x mc), then F (+)= T P^T.
// ---------------------------------------------------------------------------------------------
__global__ void k_band_rows(int m, int mc_cols, int ldc, const double* __restrict__ c, int ldt, double* __restrict__ t,
                            const double* __restrict__ band, const int* __restrict__ start, int W) {
  // t[i, J] = sum_w band[i,w] * c[start[i]+w, J]
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (J >= mc_cols || i >= m) return;
  const int s = start[i];
  const double* bw = band + (size_t)i * W;
  double acc = 0.0;
  for (int w = 0; w < W; ++w) acc += bw[w] * c[(size_t)(s + w) * ldc + J];
  t[(size_t)i * ldt + J] = acc;
}

template <bool ADD>
__global__ void k_band_cols(int m_rows, int m_cols, int ldt, const double* __restrict__ t, int ldf,
                            double* __restrict__ f, const double* __restrict__ band, const int* __restrict__ start,
                            int W) {
  // f[i, j] (+)= sum_w band[j,w] * t[i, start[j]+w]
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= m_cols || i >= m_rows) return;
  const int s = start[j];
  const double* bw = band + (size_t)j * W;
  const double* tr = t + (size_t)i * ldt + s;
  double acc = 0.0;
  for (int w = 0; w < W; ++w) acc += bw[w] * tr[w];
  const size_t k = (size_t)i * ldf + j;
  if (ADD) f[k] = f[k] + acc;
  else f[k] = acc;
}

// =============================================================================================
// host entry points
// =============================================================================================
int nfi_restrict_fw(nf_ctx* ctx, const nf_grid* gf, const double* f, const nf_grid* gc, double* c) {
  NfLaunch2D l = nf_launch2d(gc->ge - gc->gb, gc->ny);
  k_restrict_fw<<<l.grid, l.block, 0, ctx->stream>>>(*gf, f, *gc, c);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nfi_residual_restrict_fw(nf_ctx* ctx, const nf_grid* gf, const double* p, const double* b, const double* d_u,
                             const double* d_v, const nf_grid* gc, double* c, double* x0) {
  dim3 grid((gc->ny + RR_CC - 1) / RR_CC, (gc->ge - gc->gb + RR_CR - 1) / RR_CR, 1);
  nf_launch(k_residual_restrict_fw, grid, dim3(128, 2, 1), 0, ctx->stream, true, *gf, p, b, d_u, d_v, *gc, c, x0);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nfi_restrict_inject(nf_ctx* ctx, const nf_grid* gf, const double* f, const nf_grid* gc, double* c) {
  NfLaunch2D l = nf_launch2d(gc->ge - gc->gb, gc->ny);
  k_restrict_inject<<<l.grid, l.block, 0, ctx->stream>>>(*gf, f, *gc, c);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nfi_restrict_coeffs(nf_ctx* ctx, const nf_grid* gf, const double* d_u, const double* d_v, const nf_grid* gc,
                        double* duc, double* dvc) {
  NfLaunch2D l = nf_launch2d(gc->ge - gc->gb + 1, gc->ny + 1);
  k_restrict_coeffs<<<l.grid, l.block, 0, ctx->stream>>>(*gf, d_u, d_v, *gc, duc, dvc);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

// Rows / columns the block rule of the bilinear prolongation covers on a fine grid: fine rows 1 .. 2 nI, columns 1 .. 2 nJ
void nfi_prolong_block_extent(const nf_grid* gc, const nf_grid* gf, int* nI_out, int* nJ_out) {
  int nI = 0, nJ = 0;
  if (gf->nx > 3 && gf->ny > 3) {
    nI = gc->nx - 1 < (gf->nx - 2) / 2 ? gc->nx - 1 : (gf->nx - 2) / 2;
    nJ = gc->ny - 1 < (gf->ny - 2) / 2 ? gc->ny - 1 : (gf->ny - 2) / 2;
    if (nI < 0) nI = 0;
    if (nJ < 0) nJ = 0;
    if (nI == 0 || nJ == 0) nI = nJ = 0;
  }
  *nI_out = nI;
  *nJ_out = nJ;
}

// add = 2: only the thin strips outside the block rule (ring rows / columns, trailing cells of even-sized grids) are
// prolonged and added -- the streaming smoother adds the block part on the way in (nf_rbsor_stream.cu, PRL)
int nfi_prolong_linear(nf_ctx* ctx, const nf_grid* gc, const double* c, const nf_grid* gf, double* f, int add) {
  const bool strips_only = (add == 2);
  // block rule valid for coarse I in [0, nI): needs c[I+1] and fine row 2I+2 <= m-2
  int nI = 0, nJ = 0;
  if (gf->nx > 3 && gf->ny > 3) {
    nI = gc->nx - 1 < (gf->nx - 2) / 2 ? gc->nx - 1 : (gf->nx - 2) / 2;
    nJ = gc->ny - 1 < (gf->ny - 2) / 2 ? gc->ny - 1 : (gf->ny - 2) / 2;
    if (nI < 0) nI = 0;
    if (nJ < 0) nJ = 0;
    if (nI == 0 || nJ == 0) nI = nJ = 0;
  }
  // coarse rows whose 2x2 fine blocks touch this slab's fine rows [gb, ge)
  int I_lo = gf->gb >= 1 ? (gf->gb - 1) / 2 : 0;
  int I_hi = gf->ge / 2 < nI ? gf->ge / 2 : nI;
  if (I_lo > I_hi) I_lo = I_hi;
  const int gx = nJ > 0 ? (nJ + 31) / 32 : 1;
  const int nby_fast = (!strips_only && nI > 0 && I_hi > I_lo) ? (I_hi - I_lo + 7) / 8 : 0;
  const long long strip = (long long)(1 + gf->nx - (2 * nI + 1)) * gf->ny + (long long)(1 + gf->ny - (2 * nJ + 1)) * (2 * nI);
  const int nby_strip = (int)((strip + (long long)gx * 256 - 1) / ((long long)gx * 256));
  dim3 grid(gx, nby_fast + nby_strip, 1), block(32, 8, 1);
  if (add) nf_launch(k_prolong_linear<true>, grid, block, 0, ctx->stream, true, *gc, c, *gf, f, nI, nJ, nby_fast, I_lo, I_hi);
  else nf_launch(k_prolong_linear<false>, grid, block, 0, ctx->stream, true, *gc, c, *gf, f, nI, nJ, nby_fast, I_lo, I_hi);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nfi_prolong_banded(nf_ctx* ctx, const nf_grid* gc, const double* c, const nf_grid* gf, double* f, double* tmp,
                       int ldt, const double* band, const int* start, int W, int add) {
  // tmp: gf->nx rows x ldt (>= gc->ny)
  NfLaunch2D l1 = nf_launch2d(gf->nx, gc->ny);
  k_band_rows<<<l1.grid, l1.block, 0, ctx->stream>>>(gf->nx, gc->ny, gc->ld, c, ldt, tmp, band, start, W);
  NF_LAUNCH_CHECK(ctx);
  NfLaunch2D l2 = nf_launch2d(gf->nx, gf->ny);
  if (add) k_band_cols<true><<<l2.grid, l2.block, 0, ctx->stream>>>(gf->nx, gf->ny, ldt, tmp, gf->ld, f, band, start, W);
  else k_band_cols<false><<<l2.grid, l2.block, 0, ctx->stream>>>(gf->nx, gf->ny, ldt, tmp, gf->ld, f, band, start, W);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

static int check_pair(nf_ctx* ctx, const nf_grid* gf, const nf_grid* gc) {
  NF_REQUIRE(ctx, gf && gc, "grid is NULL");
  NF_REQUIRE(ctx, gf->nx >= 3 && gf->ny >= 3 && gc->nx >= 1 && gc->ny >= 1, "grid too small");
  NF_REQUIRE(ctx, gf->ld >= gf->ny + 1 && gc->ld >= gc->ny + 1, "ld must be >= ny+1");
  return NF_OK;
}

extern "C" int nf_restrict_fw(nf_ctx* ctx, const nf_grid* gf, const double* f, const nf_grid* gc, double* c) {
  NF_TRY(check_pair(ctx, gf, gc));
  NF_REQUIRE(ctx, gc->nx == (gf->nx - 1) / 2 && gc->ny == (gf->ny - 1) / 2, "coarse size must be (nf-1)//2");
  return nfi_restrict_fw(ctx, gf, f, gc, c);
}

extern "C" int nf_restrict_inject(nf_ctx* ctx, const nf_grid* gf, const double* f, const nf_grid* gc, double* c) {
  NF_TRY(check_pair(ctx, gf, gc));
  NF_REQUIRE(ctx, gc->nx == gf->nx / 2 && gc->ny == gf->ny / 2, "coarse size must be nf//2");
  return nfi_restrict_inject(ctx, gf, f, gc, c);
}

extern "C" int nf_restrict_coeffs(nf_ctx* ctx, const nf_grid* gf, const double* d_u, const double* d_v,
                                  const nf_grid* gc, double* duc, double* dvc) {
  NF_TRY(check_pair(ctx, gf, gc));
  NF_REQUIRE(ctx, 2 * gc->nx <= gf->nx && 2 * gc->ny <= gf->ny, "coarse grid larger than nf//2");
  return nfi_restrict_coeffs(ctx, gf, d_u, d_v, gc, duc, dvc);
}

extern "C" int nf_prolong_linear(nf_ctx* ctx, const nf_grid* gc, const double* c, const nf_grid* gf, double* f,
                                 int add) {
  NF_TRY(check_pair(ctx, gf, gc));
  return nfi_prolong_linear(ctx, gc, c, gf, f, add);
}
