// nf_mg.cu -- geometric multigrid driver (K13 + cycle control), fp64.
//
// Reference: pressure_solver/multigrid.py (paths relative to /root/reference/naviflow_oo)
//   solve :121-266, _solve_residual_direct :268-302, _v_cycle :304-432, _w_cycle :434-560,
//   _fmg_cycle :562-688.  The reference rebuilds the coefficient hierarchy inside every cycle call;
//   here it is built once per (d_u, d_v) by nf_mg_setup -- same numbers, no repeated work.
// The level chain, coarse mesh spacing L/(nc-1) (multigrid.py:373), pin handling and cycle order are
// the reference's.  The coarsest level (nx <= coarsest_grid_size) is solved exactly: the reference
// uses SuperLU (spsolve); here the dense matrix is inverted once per setup by Gauss-Jordan with
// partial pivoting in one thread block and applied as a mat-vec.
#include <math.h>

#include <vector>

#include "nf_pressure.cuh"

int nfi_residual_restrict_fw(nf_ctx*, const nf_grid* gf, const double* p, const double* b, const double* d_u,
                             const double* d_v, const nf_grid* gc, double* c);
int nfi_rbsor_fused(nf_ctx*, const nf_grid*, double** p, double** palt, const double* b, const double* d_u,
                    const double* d_v, const double* inv, double omega, int n_sweeps);
int nfi_inv_diag(nf_ctx*, const nf_grid*, const double* d_u, const double* d_v, double* inv);
int nfi_prolong_banded(nf_ctx*, const nf_grid* gc, const double* c, const nf_grid* gf, double* f, double* tmp,
                       int ldt, const double* band, const int* start, int W, int add);

struct MgLevel {
  nf_grid g;
  double *x = nullptr, *b = nullptr, *r = nullptr;  // r doubles as the Jacobi ping-pong buffer
  double* x2 = nullptr;                              // second solution buffer (fused SOR is double buffered)
  double* inv = nullptr;                             // 1/aP of the SOR update, rebuilt by nf_mg_setup
  double *d_u = nullptr, *d_v = nullptr;
  // banded 1-D interpolation matrix from the next-coarser level onto this level (cubic prolongation)
  double* band = nullptr;
  int* start = nullptr;
  int W = 0;
  double* ptmp = nullptr;  // (nx x ld_coarse) scratch of the separable prolongation
  size_t elems = 0;        // allocation size of p-like arrays, (nx+1)*ld
};

struct nf_mg {
  nf_ctx* ctx = nullptr;
  nf_mg_config cfg;
  std::vector<MgLevel> lv;
  double* coarse_A = nullptr;    // N x N work matrix
  double* coarse_inv = nullptr;  // N x N inverse
  int coarse_N = 0;
  bool setup_done = false;
  const double *fine_du = nullptr, *fine_dv = nullptr;
};

static inline int pad_ld(int ny) { return ((ny + 1 + 15) / 16) * 16; }

// ---------------------------------------------------------------------------------------------
// K13 coarsest-level matrix (helpers/coeff_matrix.py:6-121 with the pin row :114-119), unknown
//     numbering r = i*ny + j, and its inverse.  One block.
// ---------------------------------------------------------------------------------------------
__global__ void k_coarse_invert(nf_grid g, const double* __restrict__ d_u, const double* __restrict__ d_v,
                                double* __restrict__ A, double* __restrict__ Inv, int N) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_piv;
  for (size_t k = tid; k < (size_t)N * N; k += nt) { A[k] = 0.0; Inv[k] = 0.0; }
  __syncthreads();
  for (int r = tid; r < N; r += nt) {
    const int i = r / g.ny, j = r % g.ny;
    Inv[(size_t)r * N + r] = 1.0;
    if (r == 0) { A[0] = 1.0; continue; }
    const PCoef c = nf_pcoef(g, d_u, d_v, i, j);
    A[(size_t)r * N + r] = c.diag;
    if (i < g.nx - 1) A[(size_t)r * N + r + g.ny] = -c.e;
    if (i > 0) A[(size_t)r * N + r - g.ny] = -c.w;
    if (j < g.ny - 1) A[(size_t)r * N + r + 1] = -c.n;
    if (j > 0) A[(size_t)r * N + r - 1] = -c.s;
  }
  __syncthreads();
  for (int col = 0; col < N; ++col) {
    // pivot search: max |A[r,col]|, r >= col (ties -> smallest r)
    double best = -1.0;
    int bi = col;
    for (int r = col + tid; r < N; r += nt) {
      const double v = fabs(A[(size_t)r * N + col]);
      if (v > best) { best = v; bi = r; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, best, off);
      const int oi = __shfl_down_sync(0xffffffffu, bi, off);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      double bb = s_val[0];
      int ii = s_idx[0];
      for (int w = 1; w < (nt + 31) / 32; ++w)
        if (s_val[w] > bb || (s_val[w] == bb && s_idx[w] < ii)) { bb = s_val[w]; ii = s_idx[w]; }
      s_piv = ii;
    }
    __syncthreads();
    const int piv = s_piv;
    if (piv != col) {
      for (int c2 = tid; c2 < N; c2 += nt) {
        double t = A[(size_t)col * N + c2]; A[(size_t)col * N + c2] = A[(size_t)piv * N + c2]; A[(size_t)piv * N + c2] = t;
        t = Inv[(size_t)col * N + c2]; Inv[(size_t)col * N + c2] = Inv[(size_t)piv * N + c2]; Inv[(size_t)piv * N + c2] = t;
      }
      __syncthreads();
    }
    const double pv = A[(size_t)col * N + col];
    __syncthreads();
    for (int c2 = tid; c2 < N; c2 += nt) {
      A[(size_t)col * N + c2] = A[(size_t)col * N + c2] / pv;
      Inv[(size_t)col * N + c2] = Inv[(size_t)col * N + c2] / pv;
    }
    __syncthreads();
    // eliminate column col from every other row; thread <-> (row, column) pairs
    for (size_t k = tid; k < (size_t)N * N; k += nt) {
      const int r = (int)(k / N), c2 = (int)(k % N);
      if (r == col) continue;
      const double fct = A[(size_t)r * N + col];
      if (fct == 0.0) continue;
      if (c2 != col) A[k] -= fct * A[(size_t)col * N + c2];
      Inv[k] -= fct * Inv[(size_t)col * N + c2];
    }
    __syncthreads();
    for (int r = tid; r < N; r += nt)
      if (r != col) A[(size_t)r * N + col] = 0.0;
    __syncthreads();
  }
}

// x = Inv * b on the coarsest level (b, x pitched 2-D arrays); one warp per row
__global__ void k_coarse_apply(nf_grid g, const double* __restrict__ Inv, const double* __restrict__ b,
                               double* __restrict__ x, int N) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  double acc = 0.0;
  for (int c = lane; c < N; c += 32) acc += Inv[(size_t)warp * N + c] * b[(size_t)(c / g.ny) * g.ld + (c % g.ny)];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if (lane == 0) x[(size_t)(warp / g.ny) * g.ld + (warp % g.ny)] = acc;
}

// ---------------------------------------------------------------------------------------------
// 1-D interpolation matrix behind interpolate_cubic (multigrid_helpers.py:333-391): coordinates
// linspace(0,1,mc) -> linspace(0,1,m); mc>=4: interpolating cubic spline with not-a-knot ends (what
// FITPACK's RectBivariateSpline(s=0) builds), mc==3: parabola, mc==2: straight line (the reference
// falls back to RegularGridInterpolator 'quadratic' / 'linear' there).  Stored banded.
// ---------------------------------------------------------------------------------------------
static void build_interp_band(int mc, int m, int K, std::vector<double>& band, std::vector<int>& start, int& W) {
  W = (mc < 2 * K) ? mc : 2 * K;
  band.assign((size_t)m * W, 0.0);
  start.assign(m, 0);
  const double h = 1.0 / (mc - 1);
  std::vector<int> seg(m);
  std::vector<double> tau(m);
  for (int i = 0; i < m; ++i) {
    const double x = (m > 1) ? (double)i / (double)(m - 1) : 0.0;
    int s = (int)floor(x * (mc - 1));
    if (s > mc - 2) s = mc - 2;
    if (s < 0) s = 0;
    seg[i] = s;
    tau[i] = (x - s * h) / h;
    int st = s - W / 2 + 1;
    if (st < 0) st = 0;
    if (st > mc - W) st = mc - W;
    start[i] = st;
  }
  if (mc == 2) {
    for (int i = 0; i < m; ++i) { band[(size_t)i * W + 0] = 1.0 - tau[i] - seg[i]; band[(size_t)i * W + 1] = tau[i] + seg[i]; }
    return;
  }
  if (mc == 3) {
    for (int i = 0; i < m; ++i) {
      const double x = (double)i / (double)(m - 1);
      const double x0 = 0.0, x1 = 0.5, x2 = 1.0;
      band[(size_t)i * W + 0] = (x - x1) * (x - x2) / ((x0 - x1) * (x0 - x2));
      band[(size_t)i * W + 1] = (x - x0) * (x - x2) / ((x1 - x0) * (x1 - x2));
      band[(size_t)i * W + 2] = (x - x0) * (x - x1) / ((x2 - x0) * (x2 - x1));
    }
    return;
  }
  // fine rows grouped by segment for the scatter below
  std::vector<int> seg_first(mc, m), seg_last(mc, -1);
  for (int i = 0; i < m; ++i) {
    if (i < seg_first[seg[i]]) seg_first[seg[i]] = i;
    if (i > seg_last[seg[i]]) seg_last[seg[i]] = i;
  }
  // second derivatives M of the spline through the unit vector e_k.  Interior equations
  //   M[i-1] + 4 M[i] + M[i+1] = 6 (y[i-1]-2y[i]+y[i+1]) / h^2,  i = 1..mc-2
  // not-a-knot: M[0] = 2M[1]-M[2], M[mc-1] = 2M[mc-2]-M[mc-3]  => rows 1 and mc-2 become
  //   6 M[1] = rhs[1],  6 M[mc-2] = rhs[mc-2]   (for mc >= 5; mc == 4 couples both: handled below)
  const int n = mc - 2;  // unknowns M[1..mc-2]
  std::vector<double> lo(n), di(n), up(n), rhs(n), M(mc), cp(n), dp(n);
  for (int k = 0; k < mc; ++k) {
    for (int q = 0; q < n; ++q) {
      const int i = q + 1;
      double y0 = (i - 1 == k), y1 = (i == k), y2 = (i + 1 == k);
      rhs[q] = 6.0 * (y0 - 2.0 * y1 + y2) / (h * h);
      lo[q] = 1.0; di[q] = 4.0; up[q] = 1.0;
    }
    // fold the end conditions: row i=1: M0 + 4M1 + M2 with M0 = 2M1 - M2 -> 6 M1 + 0 M2
    di[0] = 6.0; up[0] = 0.0; lo[0] = 0.0;
    di[n - 1] = 6.0; lo[n - 1] = 0.0; up[n - 1] = 0.0;
    if (n == 2) {  // mc == 4: rows are 6M1 = r1, 6M2 = r2 (both folds apply, M0=2M1-M2, M3=2M2-M1)
      lo[1] = 0.0; up[0] = 0.0;
    }
    // Thomas
    cp[0] = up[0] / di[0];
    dp[0] = rhs[0] / di[0];
    for (int q = 1; q < n; ++q) {
      const double den = di[q] - lo[q] * cp[q - 1];
      cp[q] = up[q] / den;
      dp[q] = (rhs[q] - lo[q] * dp[q - 1]) / den;
    }
    M[n] = dp[n - 1];
    for (int q = n - 2; q >= 0; --q) M[q + 1] = dp[q] - cp[q] * M[q + 2];
    M[0] = 2.0 * M[1] - M[2];
    M[mc - 1] = 2.0 * M[mc - 2] - M[mc - 3];
    // scatter column k into the band of the fine rows whose window contains k
    int s_lo = k - W, s_hi = k + W;
    if (s_lo < 0) s_lo = 0;
    if (s_hi > mc - 2) s_hi = mc - 2;
    for (int s = s_lo; s <= s_hi; ++s) {
      if (seg_last[s] < 0) continue;
      for (int i = seg_first[s]; i <= seg_last[s]; ++i) {
        const int w = k - start[i];
        if (w < 0 || w >= W) continue;
        const double t = tau[i], u = 1.0 - t;
        const double ys = (s == k), ys1 = (s + 1 == k);
        band[(size_t)i * W + w] = u * ys + t * ys1 + (h * h / 6.0) * ((u * u * u - u) * M[s] + (t * t * t - t) * M[s + 1]);
      }
    }
  }
}

// =============================================================================================
// create / destroy / setup
// =============================================================================================
static void free_level(MgLevel& L, bool owns_coeffs) {
  if (L.x && owns_coeffs) cudaFree(L.x);
  if (L.x2) cudaFree(L.x2);
  if (L.inv) cudaFree(L.inv);
  if (L.b) cudaFree(L.b);
  if (L.r) cudaFree(L.r);
  if (owns_coeffs) {
    if (L.d_u) cudaFree(L.d_u);
    if (L.d_v) cudaFree(L.d_v);
  }
  if (L.band) cudaFree(L.band);
  if (L.start) cudaFree(L.start);
  if (L.ptmp) cudaFree(L.ptmp);
}

extern "C" int nf_mg_destroy(nf_mg* mg) {
  if (!mg) return NF_OK;
  cudaSetDevice(mg->ctx->device);
  cudaStreamSynchronize(mg->ctx->stream);
  for (size_t l = 0; l < mg->lv.size(); ++l) free_level(mg->lv[l], l > 0);
  if (mg->coarse_A) cudaFree(mg->coarse_A);
  if (mg->coarse_inv) cudaFree(mg->coarse_inv);
  delete mg;
  return NF_OK;
}

extern "C" int nf_mg_create(nf_ctx* ctx, nf_mg** out, int nx, int ny, int ld, const nf_mg_config* cfg) {
  NF_REQUIRE(ctx, out && cfg, "NULL argument");
  *out = nullptr;
  NF_REQUIRE(ctx, nx == ny, "multigrid needs a square grid (the reference's transfer operators assume it)");
  NF_REQUIRE(ctx, nx >= 3 && ld >= ny + 1, "bad grid");
  NF_REQUIRE(ctx, cfg->coarsest >= 3 && (cfg->coarsest % 2) == 1, "coarsest_grid_size must be odd and >= 3");  // multigrid.py:82-85
  NF_REQUIRE(ctx, cfg->smoother == 0 || cfg->smoother == 1, "smoother must be 0 (red-black SOR) or 1 (Jacobi)");
  NF_REQUIRE(ctx, cfg->restriction == 0 || cfg->restriction == 1, "bad restriction");
  NF_REQUIRE(ctx, cfg->interpolation == 0 || cfg->interpolation == 1, "bad interpolation");
  NF_REQUIRE(ctx, cfg->cycle_type >= 0 && cfg->cycle_type <= 2, "bad cycle_type");
  NF_REQUIRE(ctx, cfg->pre >= 0 && cfg->post >= 0, "negative smoothing count");
  nf_mg* mg = new nf_mg();
  mg->ctx = ctx;
  mg->cfg = *cfg;
  int n = nx;
  int cur_ld = ld;
  for (;;) {
    MgLevel L;
    L.g.nx = n; L.g.ny = n; L.g.ld = cur_ld; L.g.row0 = 0; L.g.gb = 0; L.g.ge = n;
    L.g.dx = cfg->length / (n - 1);  // StructuredMesh(nc, nc, L, H): structured.py:27-28
    L.g.dy = cfg->height / (n - 1);
    L.g.rho = cfg->rho;
    L.elems = (size_t)(n + 1) * cur_ld;
    mg->lv.push_back(L);
    if (n <= cfg->coarsest) break;
    const int nc = cfg->restriction == 0 ? (n - 1) / 2 : n / 2;
    if (nc < 2) break;  // cannot coarsen further (degenerate; reference would fail as well)
    n = nc;
    cur_ld = pad_ld(n);
  }
  const bool need_cubic = (cfg->interpolation == 1) || (cfg->cycle_type == 2);  // FMG hard-codes cubic (:631)
  bool ok = true;
  for (size_t l = 0; l < mg->lv.size() && ok; ++l) {
    MgLevel& L = mg->lv[l];
    const size_t bytes = L.elems * sizeof(double);
    ok = ok && cudaMalloc(&L.r, bytes) == cudaSuccess && cudaMalloc(&L.x2, bytes) == cudaSuccess;
    if (ok) cudaMemsetAsync(L.x2, 0, bytes, ctx->stream);
    if (ok && cfg->smoother == 0) {
      ok = cudaMalloc(&L.inv, bytes) == cudaSuccess;
      if (ok) cudaMemsetAsync(L.inv, 0, bytes, ctx->stream);
    }
    if (l > 0) {
      ok = ok && cudaMalloc(&L.x, bytes) == cudaSuccess && cudaMalloc(&L.b, bytes) == cudaSuccess &&
           cudaMalloc(&L.d_u, bytes) == cudaSuccess && cudaMalloc(&L.d_v, bytes) == cudaSuccess;
      if (ok) {
        cudaMemsetAsync(L.x, 0, bytes, ctx->stream);
        cudaMemsetAsync(L.b, 0, bytes, ctx->stream);
        cudaMemsetAsync(L.d_u, 0, bytes, ctx->stream);
        cudaMemsetAsync(L.d_v, 0, bytes, ctx->stream);
      }
    }
    if (ok) cudaMemsetAsync(L.r, 0, bytes, ctx->stream);
    if (ok && need_cubic && l + 1 < mg->lv.size()) {
      const int mc = mg->lv[l + 1].g.nx, m = L.g.nx;
      std::vector<double> band;
      std::vector<int> start;
      int W = 0;
      build_interp_band(mc, m, 24, band, start, W);
      L.W = W;
      ok = ok && cudaMalloc(&L.band, band.size() * sizeof(double)) == cudaSuccess &&
           cudaMalloc(&L.start, start.size() * sizeof(int)) == cudaSuccess &&
           cudaMalloc(&L.ptmp, (size_t)m * mg->lv[l + 1].g.ld * sizeof(double)) == cudaSuccess;
      if (ok) {
        ok = cudaMemcpy(L.band, band.data(), band.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(L.start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
      }
    }
  }
  const MgLevel& C = mg->lv.back();
  mg->coarse_N = C.g.nx * C.g.ny;
  if (ok && mg->coarse_N > 4096) {
    ctx->err = "coarsest level too large for the dense coarse solve (nx*ny must be <= 4096)";
    nf_mg_destroy(mg);
    return NF_ERR_UNSUPPORTED;
  }
  ok = ok && cudaMalloc(&mg->coarse_A, (size_t)mg->coarse_N * mg->coarse_N * sizeof(double)) == cudaSuccess &&
       cudaMalloc(&mg->coarse_inv, (size_t)mg->coarse_N * mg->coarse_N * sizeof(double)) == cudaSuccess;
  if (!ok) {
    ctx->err = std::string("multigrid allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    nf_mg_destroy(mg);
    return NF_ERR_ALLOC;
  }
  *out = mg;
  return NF_OK;
}

extern "C" int nf_mg_num_levels(nf_mg* mg) { return mg ? (int)mg->lv.size() : 0; }

extern "C" int nf_mg_level_shape(nf_mg* mg, int level, int* nx, int* ny, int* ld) {
  if (!mg || level < 0 || level >= (int)mg->lv.size()) return NF_ERR_ARG;
  if (nx) *nx = mg->lv[level].g.nx;
  if (ny) *ny = mg->lv[level].g.ny;
  if (ld) *ld = mg->lv[level].g.ld;
  return NF_OK;
}

// level arrays for tests / callers that want to look at the hierarchy: which = 0 d_u, 1 d_v, 2 x, 3 b, 4 r
extern "C" const double* nf_mg_level_array(nf_mg* mg, int level, int which) {
  if (!mg || level < 0 || level >= (int)mg->lv.size()) return nullptr;
  const MgLevel& L = mg->lv[level];
  switch (which) {
    case 0: return L.d_u;
    case 1: return L.d_v;
    case 2: return L.x;
    case 3: return L.b;
    case 4: return L.r;
  }
  return nullptr;
}

extern "C" int nf_mg_setup(nf_mg* mg, const double* d_u, const double* d_v) {
  if (!mg) return NF_ERR_ARG;
  nf_ctx* ctx = mg->ctx;
  NF_REQUIRE(ctx, d_u && d_v, "NULL coefficient array");
  mg->lv[0].d_u = const_cast<double*>(d_u);
  mg->lv[0].d_v = const_cast<double*>(d_v);
  for (size_t l = 0; l + 1 < mg->lv.size(); ++l)
    NF_TRY(nfi_restrict_coeffs(ctx, &mg->lv[l].g, mg->lv[l].d_u, mg->lv[l].d_v, &mg->lv[l + 1].g, mg->lv[l + 1].d_u,
                               mg->lv[l + 1].d_v));
  if (mg->cfg.smoother == 0)
    for (size_t l = 0; l < mg->lv.size(); ++l)
      NF_TRY(nfi_inv_diag(ctx, &mg->lv[l].g, mg->lv[l].d_u, mg->lv[l].d_v, mg->lv[l].inv));
  const MgLevel& C = mg->lv.back();
  if (C.g.nx <= mg->cfg.coarsest) {
    k_coarse_invert<<<1, 1024, 0, ctx->stream>>>(C.g, C.d_u, C.d_v, mg->coarse_A, mg->coarse_inv, mg->coarse_N);
    NF_LAUNCH_CHECK(ctx);
  }
  mg->setup_done = true;
  return NF_OK;
}

// =============================================================================================
// cycles
// =============================================================================================
// smoothing may move the iterate to the other buffer: (*x, *alt) are swapped accordingly
static int mg_smooth(nf_mg* mg, int l, double** x, double** alt, const double* b, int n) {
  MgLevel& L = mg->lv[l];
  if (mg->cfg.smoother == 0) return nfi_rbsor_fused(mg->ctx, &L.g, x, alt, b, L.d_u, L.d_v, L.inv, mg->cfg.omega, n);
  return nfi_jacobi(mg->ctx, &L.g, *x, L.r, b, L.d_u, L.d_v, mg->cfg.omega, n);
}

static int mg_coarse_solve(nf_mg* mg, int l, double* x, const double* b) {
  nf_ctx* ctx = mg->ctx;
  const MgLevel& L = mg->lv[l];
  const int N = mg->coarse_N;
  const int threads = 128, blocks = (N * 32 + threads - 1) / threads;
  k_coarse_apply<<<blocks, threads, 0, ctx->stream>>>(L.g, mg->coarse_inv, b, x, N);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

static int mg_prolong_add(nf_mg* mg, int l, double* x, int cubic, int add) {
  // x_l (+)= P x_{l+1}
  MgLevel& L = mg->lv[l];
  MgLevel& C = mg->lv[l + 1];
  if (!cubic) return nfi_prolong_linear(mg->ctx, &C.g, C.x, &L.g, x, add);
  return nfi_prolong_banded(mg->ctx, &C.g, C.x, &L.g, x, L.ptmp, C.g.ld, L.band, L.start, L.W, add);
}

// one V (kind 0) or W (kind 1) cycle on level l: multigrid.py:304-432 / :434-560
static int mg_cycle(nf_mg* mg, int l, double** x, double** alt, const double* b, int kind) {
  nf_ctx* ctx = mg->ctx;
  MgLevel& L = mg->lv[l];
  if (L.g.nx <= mg->cfg.coarsest || l + 1 == (int)mg->lv.size()) return mg_coarse_solve(mg, l, *x, b);
  MgLevel& C = mg->lv[l + 1];
  NF_TRY(mg_smooth(mg, l, x, alt, b, mg->cfg.pre));
  if (mg->cfg.restriction == 0) {
    NF_TRY(nfi_residual_restrict_fw(ctx, &L.g, *x, b, L.d_u, L.d_v, &C.g, C.b));
  } else {
    NF_TRY(nfi_residual(ctx, &L.g, *x, b, L.d_u, L.d_v, L.r));
    NF_TRY(nfi_restrict_inject(ctx, &L.g, L.r, &C.g, C.b));
  }
  NF_TRY(nfi_fill(ctx, C.x, (size_t)C.g.nx * C.g.ld, 0.0));
  const int reps = (kind == 1) ? 2 : 1;
  for (int rep = 0; rep < reps; ++rep) NF_TRY(mg_cycle(mg, l + 1, &C.x, &C.x2, C.b, kind));
  NF_TRY(mg_prolong_add(mg, l, *x, mg->cfg.interpolation, 1));
  NF_TRY(mg_smooth(mg, l, x, alt, b, mg->cfg.post));
  return NF_OK;
}

static int mg_restrict_rhs(nf_mg* mg, int l, const double* f) {
  MgLevel& L = mg->lv[l];
  MgLevel& C = mg->lv[l + 1];
  if (mg->cfg.restriction == 0) return nfi_restrict_fw(mg->ctx, &L.g, f, &C.g, C.b);
  return nfi_restrict_inject(mg->ctx, &L.g, f, &C.g, C.b);
}

// ||b - A x|| / ||b|| on level l (host value; multigrid.py:652-676)
// *b_norm < 0 on entry: ||b|| is not known yet and is computed in the same pass; otherwise it is kept
static int mg_rel_residual(nf_mg* mg, int l, const double* x, const double* b, double* r_norm, double* b_norm) {
  nf_ctx* ctx = mg->ctx;
  MgLevel& L = mg->lv[l];
  const int with_b = (*b_norm < 0.0) ? 1 : 0;
  NF_TRY(nfi_residual_norms(ctx, &L.g, x, b, L.d_u, L.d_v, L.r, with_b, 0));
  double s[2];
  NF_TRY(nf_read_scalars(ctx, 0, with_b ? 2 : 1, s));
  *r_norm = sqrt(s[0]);
  if (with_b) *b_norm = sqrt(s[1]);
  return NF_OK;
}

// level 0 iterates between the caller's x and the hierarchy's x2: bring the result home and restore x2
static int mg_finish_level0(nf_mg* mg, double* x, double* cur, double* alt) {
  nf_ctx* ctx = mg->ctx;
  MgLevel& L = mg->lv[0];
  if (cur != x) {  // cur is the hierarchy's own buffer, alt is the caller's array
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(x, cur, (size_t)L.g.nx * L.g.ld * sizeof(double), cudaMemcpyDeviceToDevice,
                                       ctx->stream));
    L.x2 = cur;
  } else {
    L.x2 = alt;
  }
  return NF_OK;
}

// recursive FMG (multigrid.py:562-688): RHS restricted down, exact coarsest solve, cubic prolongation
// (hard-coded :631), max_cycles_buildup cycles per level with early exit on ||r||/||b|| < tol.
static int mg_fmg(nf_mg* mg, int l, double** x, double** alt, const double* b) {
  MgLevel& L = mg->lv[l];
  if (L.g.nx <= mg->cfg.coarsest || l + 1 == (int)mg->lv.size()) return mg_coarse_solve(mg, l, *x, b);
  MgLevel& C = mg->lv[l + 1];
  NF_TRY(mg_restrict_rhs(mg, l, b));
  NF_TRY(mg_fmg(mg, l + 1, &C.x, &C.x2, C.b));
  NF_TRY(mg_prolong_add(mg, l, *x, 1, 0));
  for (int c = 0; c < mg->cfg.max_cycles_buildup; ++c) {
    NF_TRY(mg_cycle(mg, l, x, alt, b, mg->cfg.cycle_buildup));
    if (mg->cfg.tolerance < 1.0 && c + 1 < mg->cfg.max_cycles_buildup) {
      double rn, bn = -1.0;
      NF_TRY(mg_rel_residual(mg, l, *x, b, &rn, &bn));
      const double rel = bn > 0.0 ? rn / bn : rn;
      if (rel < mg->cfg.tolerance) break;
    }
  }
  return NF_OK;
}

extern "C" int nf_mg_cycle(nf_mg* mg, double* x, const double* b, int kind) {
  if (!mg) return NF_ERR_ARG;
  NF_REQUIRE(mg->ctx, mg->setup_done, "nf_mg_setup has not been called");
  NF_REQUIRE(mg->ctx, kind == 0 || kind == 1, "kind must be 0 ('v') or 1 ('w')");
  double* cur = x;
  double* alt = mg->lv[0].x2;
  NF_TRY(mg_cycle(mg, 0, &cur, &alt, b, kind));
  return mg_finish_level0(mg, x, cur, alt);
}

// MultiGridSolver.solve without get_rhs.  sync == 0 (FMG mode only): the final ||r||^2, ||b||^2 stay in the
// context's device scalars 0,1 and no host synchronisation happens.
int nfi_mg_solve(nf_mg* mg, const double* b, double* x, double* r, nf_mg_info* info, int sync) {
  nf_ctx* ctx = mg->ctx;
  NF_REQUIRE(ctx, mg->setup_done, "nf_mg_setup has not been called");
  NF_REQUIRE(ctx, b && x, "NULL argument");
  MgLevel& L = mg->lv[0];
  double* own_r = L.r;
  if (r) L.r = r;  // residual field goes straight to the caller's array
  int status = NF_OK;
  double rn = 0.0, bn = -1.0;  // bn < 0: ||b|| not computed yet
  int cycles = 0;
  double* cur = x;
  double* alt = L.x2;
  do {
    status = nfi_fill(ctx, cur, (size_t)L.g.nx * L.g.ld, 0.0);  // x0 = 0 (multigrid.py:165)
    if (status) break;
    if (mg->cfg.cycle_type == 2) {
      status = mg_fmg(mg, 0, &cur, &alt, b);
      if (status) break;
      if (mg->cfg.cycle_final >= 0) {
        status = mg_cycle(mg, 0, &cur, &alt, b, mg->cfg.cycle_final);
        if (status) break;
        cycles = 1;
      }
      if (sync) {
        status = mg_rel_residual(mg, 0, cur, b, &rn, &bn);
      } else {
        status = nfi_residual_norms(ctx, &L.g, cur, b, L.d_u, L.d_v, L.r, 1, 0);
      }
    } else {
      for (int k = 0; k < mg->cfg.max_iterations; ++k) {
        status = mg_cycle(mg, 0, &cur, &alt, b, mg->cfg.cycle_type);
        if (status) break;
        ++cycles;
        status = mg_rel_residual(mg, 0, cur, b, &rn, &bn);
        if (status) break;
        const double rel = bn > 0.0 ? rn / bn : rn;
        if (rel < mg->cfg.tolerance) break;
      }
    }
  } while (0);
  L.r = own_r;
  if (status) return status;
  NF_TRY(mg_finish_level0(mg, x, cur, alt));
  if (info) {
    info->r_norm = rn;
    info->b_norm = bn < 0.0 ? 0.0 : bn;
    info->cycles = cycles;
    info->levels = (int)mg->lv.size();
  }
  return NF_OK;
}

extern "C" int nf_mg_solve(nf_mg* mg, const double* b, double* x, double* r, nf_mg_info* info) {
  if (!mg) return NF_ERR_ARG;
  return nfi_mg_solve(mg, b, x, r, info, 1);
}

nf_mg_config* nfi_mg_config(nf_mg* mg) { return &mg->cfg; }

// standalone interpolate_cubic (tests, Python-level callers): builds the band on every call
extern "C" int nf_prolong_cubic(nf_ctx* ctx, const nf_grid* gc, const double* c, const nf_grid* gf, double* f,
                                int add) {
  NF_REQUIRE(ctx, gc && gf && c && f, "NULL argument");
  NF_REQUIRE(ctx, gc->nx == gc->ny && gf->nx == gf->ny, "square grids only");
  NF_REQUIRE(ctx, gc->nx >= 2 && gf->nx >= 2, "grid too small");
  std::vector<double> band;
  std::vector<int> start;
  int W = 0;
  build_interp_band(gc->nx, gf->nx, 24, band, start, W);
  double *dband = nullptr, *tmp = nullptr;
  int* dstart = nullptr;
  NF_CHECK_CUDA(ctx, cudaMalloc(&dband, band.size() * sizeof(double)));
  NF_CHECK_CUDA(ctx, cudaMalloc(&dstart, start.size() * sizeof(int)));
  NF_CHECK_CUDA(ctx, cudaMalloc(&tmp, (size_t)gf->nx * gc->ld * sizeof(double)));
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(dband, band.data(), band.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(dstart, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  int st = nfi_prolong_banded(ctx, gc, c, gf, f, tmp, gc->ld, dband, dstart, W, add);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(dband); cudaFree(dstart); cudaFree(tmp);
  return st;
}
