// nf_mg.cu -- geometric multigrid driver (K13 + cycle control), fp64, single GPU or row slabs.
//
// Reference: pressure_solver/multigrid.py (paths relative to /root/reference/naviflow_oo)
//   solve :121-266, _solve_residual_direct :268-302, _v_cycle :304-432, _w_cycle :434-560,
//   _fmg_cycle :562-688.  The reference rebuilds the coefficient hierarchy inside every cycle call;
//   here it is built once per (d_u, d_v) by nf_mg_setup -- same numbers, no repeated work.
// The level chain, coarse mesh spacing L/(nc-1) (multigrid.py:373), pin handling and cycle order are
// the reference's.  The coarsest level (nx <= coarsest_grid_size) is solved exactly: the reference
// uses SuperLU (spsolve); here the dense matrix is inverted once per setup by Gauss-Jordan with
// partial pivoting in one thread block and applied as a mat-vec.
//
// Every phase of a cycle is written as "for each slab this process owns: launch; then exchange halos"
// (nf_slab.cuh).  Fine levels are cut into row slabs; once a slab would get thinner than NF_MIN_SLAB_ROWS the
// level (and all coarser ones) is replicated on every rank: the restricted right-hand side is shared once
// and the coarse part of the cycle runs redundantly, so the way back up needs no communication.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "nf_mg_tail.cuh"
#include "nf_pressure.cuh"
#include "nf_slab.cuh"

#define NF_MIN_SLAB_ROWS 64   // a level is cut only while every rank keeps at least this many rows
// ... and only while it is large enough for the cut to pay: a level of <= NF_REPLICATE_BELOW rows per side is launch-latency
// bound (its kernels take the same few microseconds on the whole level as on a slab), so three halo exchanges of ~7.5 us per
// cycle cost more than the redundant arithmetic (env NF_REPLICATE_BELOW overrides; coarser levels of the finest one only)
#define NF_REPLICATE_BELOW 768
#define NF_SMOOTH_HALO NF_HALO  // halo rows kept valid around the iterate: 6 for a smoother launch (2 * 3 sweeps),
                                // +1 / +2 when the residual norms / the restriction ride on it

int nfi_residual_restrict_fw(nf_ctx*, const nf_grid* gf, const double* p, const double* b, const double* d_u,
                             const double* d_v, const nf_grid* gc, double* c, double* x0);
int nfi_inv_diag(nf_ctx*, const nf_grid*, const double* d_u, const double* d_v, double* inv);
int nfi_gs_lex(nf_ctx*, const nf_grid*, double* p, const double* b, const double* d_u, const double* d_v, double omega,
               int n_sweeps, int symmetric);
int nfi_prolong_banded(nf_ctx*, const nf_grid* gc, const double* c, const nf_grid* gf, double* f, double* tmp,
                       int ldt, const double* band, const int* start, int W, int add);

struct MgSlab {  // arrays of one (level, local rank)
  double *x = nullptr, *x2 = nullptr, *b = nullptr, *r = nullptr;  // r doubles as the Jacobi ping-pong buffer
  double *d_u = nullptr, *d_v = nullptr, *inv = nullptr;
  bool owns_x = false, owns_b = false, owns_d = false;
};

struct MgLevel {
  LevelGeom geom;
  std::vector<int> rgb, rge;   // rows of THIS level each rank restricts into (ownership induced by the finer level)
  std::vector<MgSlab> s;       // one per local rank
  // banded 1-D interpolation matrix from the next-coarser level onto this level (cubic prolongation)
  double* band = nullptr;
  int* start = nullptr;
  int W = 0;
  double* ptmp = nullptr;
};

struct nf_mg {
  nf_ctx* ctx = nullptr;
  nf_team* team = nullptr;
  bool owns_team = false;
  nf_mg_config cfg;
  std::vector<MgLevel> lv;
  std::vector<double*> coarse_A, coarse_inv;  // per local rank
  std::vector<double*> scal;                  // per local rank: 8 device doubles for the norms
  double* scal_host = nullptr;                // pinned
  int coarse_N = 0;
  int tail_level = -1;  // first level (>= 1) of the single-kernel coarse end of a V-cycle (nf_mg_tail.cu), -1: none
  bool setup_done = false;
  // CUDA graph of one whole cycle at level 0 (single slab): captured after a warm-up cycle, replayed afterwards.
  // Valid while the level-0 arrays stay the same and every level swaps x/x2 an even number of times per cycle.
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t cap_stream = nullptr;
  const void* graph_key[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // x, b, d_u, d_v, x2 of level 0
  int graph_kind = -1;
  int graph_part = 0;
  bool graph_want_norm = true;
  bool graph_swapped = false;  // the captured launches leave level 0's x / x2 exchanged (odd number of smoother launches)
  bool graph_norm_fused = false;
  long long graph_nodes = 0;
  // device-side convergence loop: one graph whose WHILE node replays {cycle, norms, test} until ||r||/||b|| < tol
  cudaGraphExec_t loop_exec = nullptr;
  const void* loop_key[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int loop_kind = -1;
  bool loop_norm_fused = false;
  long long loop_nodes = 0;
  bool use_loop = true;
  // slabs over peer memory: the scalar all-reduce of the residual norms rode on the last halo exchange of the cycle
  bool norm_reduced = false;
  bool graph_norm_reduced = false;
  int warm_cycles = 0;
  bool use_graph = true;
  // optional live timing of the finest-level smoother launches (bench.py roofline): event pairs, read at the
  // synchronisation points the cycle loop already has
  bool timing = false;
  // NF_MG_PROFILE=1: event pairs around every phase of the cycles (plain launches, no graphs), summed per (level, phase)
  // and printed to stderr when the hierarchy is destroyed -- a development aid for the slab runs, where ncu cannot go
  bool prof = false;
  std::vector<cudaEvent_t> pev;
  std::vector<int> ptag;  // per pair: level * 16 + phase
  size_t pev_used = 0;
  double pms[32][16] = {};
  long long pcnt[32][16] = {};
  long long psolves = 0, pcycles = 0;
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  double smooth_ms = 0.0;
  long long smooth_launches = 0;
};

// ---------------------------------------------------------------------------------------------
// K13 coarsest-level matrix (helpers/coeff_matrix.py:6-121 with the pin row :114-119), unknown
//     numbering r = i*ny + j, and its inverse.  One block.
// ---------------------------------------------------------------------------------------------
// use_smem: the two N x N work matrices live in dynamic shared memory (2 N^2 doubles; N = 49 for the default 7 x 7 coarsest
// level) and only the inverse is written back: the ~7 barrier phases per column each waited on global-memory round trips
// before (203 us per set-up = per outer iteration at N = 49; same operations in the same order, same bits).
__global__ void k_coarse_invert(nf_grid g, const double* __restrict__ d_u, const double* __restrict__ d_v,
                                double* __restrict__ A_glob, double* __restrict__ Inv_glob, int N, int use_smem) {
  const int tid = threadIdx.x, nt = blockDim.x;
  extern __shared__ __align__(16) double s_work[];
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_piv;
  double* A = use_smem ? s_work : A_glob;
  double* Inv = use_smem ? s_work + (size_t)N * N : Inv_glob;
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
  if (use_smem)
    for (size_t k = tid; k < (size_t)N * N; k += nt) Inv_glob[k] = Inv[k];
}

// The same inverse for N <= 52 (the default 7 x 7 coarsest level: N = 49) at a fraction of the latency: both matrices in
// static shared memory, thread <-> (row group, column), the pivot search in one warp, four barriers per column.  Per element
// the operations and their order are k_coarse_invert's (swap, divide the pivot row by the pivot, subtract fct x pivot row,
// rows with fct == 0 untouched), so Inv has the same bits; A's finished columns are scratch there and are simply not
// maintained here.  203 -> 76 us (ncu) per set-up, i.e. per outer iteration.
constexpr int CI_MAX = 52;  // 2 x 52^2 doubles of static shared memory (48 KB limit); N = 49, 25, 9 for coarsest 7, 5, 3
__global__ void __launch_bounds__(1024, 1)
k_coarse_invert_small(nf_grid g, const double* __restrict__ d_u, const double* __restrict__ d_v, double* __restrict__ Inv_glob,
                      int N) {
  __shared__ double A[CI_MAX * CI_MAX], Iv[CI_MAX * CI_MAX], rowA[CI_MAX], rowI[CI_MAX];
  __shared__ double s_pv;
  __shared__ int s_piv;
  const int tid = threadIdx.x, c2 = tid & 63, rg = tid >> 6;  // 16 row groups x 64 columns
  for (int k = tid; k < N * N; k += 1024) { A[k] = 0.0; Iv[k] = 0.0; }
  __syncthreads();
  for (int r = tid; r < N; r += 1024) {
    const int i = r / g.ny, j = r % g.ny;
    Iv[r * N + r] = 1.0;
    if (r == 0) { A[0] = 1.0; continue; }
    const PCoef c = nf_pcoef(g, d_u, d_v, i, j);
    A[r * N + r] = c.diag;
    if (i < g.nx - 1) A[r * N + r + g.ny] = -c.e;
    if (i > 0) A[r * N + r - g.ny] = -c.w;
    if (j < g.ny - 1) A[r * N + r + 1] = -c.n;
    if (j > 0) A[r * N + r - 1] = -c.s;
  }
  __syncthreads();
  for (int col = 0; col < N; ++col) {
    if (tid < 32) {  // pivot: max |A[r,col]|, r >= col, ties -> smallest r
      double best = -1.0;
      int bi = col;
      for (int r = col + tid; r < N; r += 32) {
        const double v = fabs(A[r * N + col]);
        if (v > best) { best = v; bi = r; }
      }
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (tid == 0) s_piv = bi;
    }
    __syncthreads();
    const int piv = s_piv;
    double a = 0.0, ai = 0.0;
    if (rg == 0 && c2 < N) {  // row swap in registers; the new pivot row stays in (a, ai)
      a = A[col * N + c2];
      ai = Iv[col * N + c2];
      if (piv != col) {
        const double b = A[piv * N + c2], bi = Iv[piv * N + c2];
        A[piv * N + c2] = a;
        Iv[piv * N + c2] = ai;
        a = b;
        ai = bi;
      }
      if (c2 == col) s_pv = a;
    }
    __syncthreads();
    if (rg == 0 && c2 < N) {
      const double pv = s_pv;
      a = a / pv;
      ai = ai / pv;
      A[col * N + c2] = a;
      Iv[col * N + c2] = ai;
      rowA[c2] = a;
      rowI[c2] = ai;
    }
    __syncthreads();
    if (c2 < N) {
      const double ra = rowA[c2], ri = rowI[c2];
      for (int r = rg; r < N; r += 16) {
        if (r == col) continue;
        const double fct = A[r * N + col];
        if (fct == 0.0) continue;
        if (c2 > col) A[r * N + c2] -= fct * ra;
        Iv[r * N + c2] -= fct * ri;
      }
    }
    __syncthreads();
  }
  for (int k = tid; k < N * N; k += 1024) Inv_glob[k] = Iv[k];
}

// x = Inv * b on the coarsest level (b, x pitched 2-D arrays); one warp per row
__global__ void k_coarse_apply(nf_grid g, const double* __restrict__ Inv, const double* __restrict__ b,
                               double* __restrict__ x, int N) {
  nf_pdl_entry();
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
static int nlocal(const nf_mg* mg) { return (int)mg->team->local.size(); }
static void mg_harvest_timing(nf_mg* mg);

enum { PH_SMOOTH = 0, PH_EXCH_X, PH_RESTRICT, PH_FILL, PH_EXCH_B, PH_PROLONG, PH_TAIL, PH_COARSE, PH_NORM, PH_ALLREDUCE, PH_N };
static const char* const kPhaseName[PH_N] = {"smooth", "exch_x", "restrict", "fill", "exch_b", "prolong", "tail", "coarse",
                                             "norm", "allreduce"};
struct MgProf {  // scope guard: records an event pair on the context's stream when profiling is on
  nf_mg* mg;
  bool on;
  MgProf(nf_mg* m, int level, int phase) : mg(m), on(m->prof) {
    if (!on) return;
    if (mg->pev_used + 2 > mg->pev.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      mg->pev.push_back(a); mg->pev.push_back(b);
    }
    if (mg->ptag.size() < mg->pev.size() / 2) mg->ptag.resize(mg->pev.size() / 2);
    mg->ptag[mg->pev_used / 2] = level * 16 + phase;
    cudaEventRecord(mg->pev[mg->pev_used], mg->ctx->stream);
  }
  ~MgProf() {
    if (!on) return;
    cudaEventRecord(mg->pev[mg->pev_used + 1], mg->ctx->stream);
    mg->pev_used += 2;
  }
};
static void mg_prof_harvest(nf_mg* mg) {  // after a stream synchronisation
  for (size_t e = 0; e + 1 < mg->pev_used; e += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, mg->pev[e], mg->pev[e + 1]) != cudaSuccess) continue;
    const int t = mg->ptag[e / 2];
    mg->pms[(t / 16) & 31][t % 16] += ms;
    mg->pcnt[(t / 16) & 31][t % 16]++;
  }
  mg->pev_used = 0;
}
static void mg_prof_print(nf_mg* mg) {
  if (!mg->prof || mg->pcycles == 0) return;
  fprintf(stderr, "[NF_MG_PROFILE] rank %d: %lld solves, %lld cycles; microseconds per cycle (launches per cycle)\n",
          mg->team->local.empty() ? 0 : mg->team->local[0], mg->psolves, mg->pcycles);
  double total = 0.0;
  for (size_t l = 0; l < mg->lv.size() && l < 32; ++l) {
    bool any = false;
    for (int ph = 0; ph < PH_N; ++ph) any = any || mg->pcnt[l][ph] > 0;
    if (!any) continue;
    fprintf(stderr, "  level %2zu (%5d rows%s):", l, mg->lv[l].geom.nx, mg->lv[l].geom.dist ? ", cut" : "");
    for (int ph = 0; ph < PH_N; ++ph)
      if (mg->pcnt[l][ph] > 0) {
        fprintf(stderr, " %s %.1f (%.1f)", kPhaseName[ph], 1e3 * mg->pms[l][ph] / mg->pcycles,
                (double)mg->pcnt[l][ph] / mg->pcycles);
        total += mg->pms[l][ph];
      }
    fprintf(stderr, "\n");
  }
  fprintf(stderr, "  sum of phases: %.1f us per cycle\n", 1e3 * total / mg->pcycles);
}

extern "C" int nf_mg_destroy(nf_mg* mg) {
  if (!mg) return NF_OK;
  cudaSetDevice(mg->ctx->device);
  cudaStreamSynchronize(mg->ctx->stream);
  nf_team* tm = mg->team;
  for (MgLevel& L : mg->lv) {
    for (MgSlab& S : L.s) {
      if (S.owns_x) nf_team_release(tm, S.x);
      nf_team_release(tm, S.x2);
      if (S.owns_b) nf_team_release(tm, S.b);
      nf_team_release(tm, S.r);
      if (S.owns_d) { nf_team_release(tm, S.d_u); nf_team_release(tm, S.d_v); }
      nf_team_release(tm, S.inv);
    }
    if (L.band) cudaFree(L.band);
    if (L.start) cudaFree(L.start);
    if (L.ptmp) cudaFree(L.ptmp);
  }
  for (double* p : mg->coarse_A) nf_team_release(tm, p);
  for (double* p : mg->coarse_inv) nf_team_release(tm, p);
  for (double* p : mg->scal) nf_team_release(tm, p);
  if (mg->scal_host) cudaFreeHost(mg->scal_host);
  for (cudaEvent_t e : mg->ev) cudaEventDestroy(e);
  if (mg->graph_exec) cudaGraphExecDestroy(mg->graph_exec);
  if (mg->loop_exec) cudaGraphExecDestroy(mg->loop_exec);
  mg_prof_print(mg);
  for (cudaEvent_t e : mg->pev) cudaEventDestroy(e);
  if (mg->cap_stream) cudaStreamDestroy(mg->cap_stream);
  if (mg->owns_team) nf_team_destroy(mg->team);
  delete mg;
  return NF_OK;
}

static bool dev_alloc(nf_team* team, double** p, size_t elems, size_t elems_max) {
  *p = nf_team_alloc(team, elems, elems_max);
  return *p != nullptr;
}

// level-0 geometry of a team (shared with the SIMPLE driver)
LevelGeom nf_level0_geom(const nf_team* team, int nx, int ny, int ld, double length, double height, double rho) {
  LevelGeom g;
  g.nx = nx; g.ny = ny; g.ld = ld;
  g.dx = length / (nx - 1);  // StructuredMesh(n, n, L, H): structured.py:27-28
  g.dy = height / (ny - 1);
  g.rho = rho;
  g.dist = nf_split_rows(nx, team->world, NF_MIN_SLAB_ROWS, g.gb, g.ge);
  g.halo = g.dist ? NF_HALO : 0;
  return g;
}

int nfi_mg_create(nf_team* team, nf_mg** out, int nx, int ny, int ld, const nf_mg_config* cfg) {
  nf_ctx* ctx = team->ctx;
  NF_REQUIRE(ctx, out && cfg, "NULL argument");
  *out = nullptr;
  NF_REQUIRE(ctx, nx == ny, "multigrid needs a square grid (the reference's transfer operators assume it)");
  NF_REQUIRE(ctx, nx >= 3 && ld >= ny + 1 && (ld % 2) == 0, "bad grid (ld must be even and >= ny+1)");
  NF_REQUIRE(ctx, cfg->coarsest >= 3 && (cfg->coarsest % 2) == 1, "coarsest_grid_size must be odd and >= 3");  // multigrid.py:82-85
  NF_REQUIRE(ctx, cfg->smoother >= 0 && cfg->smoother <= 3,
             "smoother must be 0 (red-black SOR), 1 (Jacobi), 2 (lexicographic SOR) or 3 (symmetric SOR)");
  NF_REQUIRE(ctx, cfg->restriction == 0 || cfg->restriction == 1, "bad restriction");
  NF_REQUIRE(ctx, cfg->interpolation == 0 || cfg->interpolation == 1, "bad interpolation");
  NF_REQUIRE(ctx, cfg->cycle_type >= 0 && cfg->cycle_type <= 2, "bad cycle_type");
  NF_REQUIRE(ctx, cfg->pre >= 0 && cfg->post >= 0, "negative smoothing count");
  nf_mg* mg = new nf_mg();
  { const char* ep = getenv("NF_MG_PROFILE"); mg->prof = ep && ep[0] == '1'; }
  mg->ctx = ctx;
  mg->team = team;
  mg->cfg = *cfg;
  const int nl = (int)team->local.size();
  // level chain
  int n = nx;
  for (;;) {
    MgLevel L;
    if (mg->lv.empty()) {
      L.geom = nf_level0_geom(team, n, n, ld, cfg->length, cfg->height, cfg->rho);
      L.rgb = L.geom.gb; L.rge = L.geom.ge;
    } else {
      const MgLevel& F = mg->lv.back();
      LevelGeom g;
      g.nx = n; g.ny = n; g.ld = nf_pad_ld(n);
      g.dx = cfg->length / (n - 1); g.dy = cfg->height / (n - 1); g.rho = cfg->rho;
      // rows each rank produces when it restricts its part of the finer level
      if (F.geom.dist) nf_coarsen_split(F.geom.gb, F.geom.ge, n, L.rgb, L.rge);
      else { L.rgb.assign(team->world, 0); L.rge.assign(team->world, n); }
      bool cut = F.geom.dist;
      if (cut)
        for (int r = 0; r < team->world; ++r)
          if (L.rge[r] - L.rgb[r] < NF_MIN_SLAB_ROWS) cut = false;
      {
        static const int below = getenv("NF_REPLICATE_BELOW") ? atoi(getenv("NF_REPLICATE_BELOW")) : NF_REPLICATE_BELOW;
        if (n <= below) cut = false;
      }
      g.dist = cut;
      g.halo = cut ? NF_HALO : 0;
      if (cut) { g.gb = L.rgb; g.ge = L.rge; }
      else { g.gb.assign(team->world, 0); g.ge.assign(team->world, n); }
      L.geom = g;
    }
    mg->lv.push_back(L);
    if (n <= cfg->coarsest) break;
    const int nc = cfg->restriction == 0 ? (n - 1) / 2 : n / 2;
    if (nc < 2) break;  // cannot coarsen further (degenerate; the reference would fail as well)
    n = nc;
  }
  const bool need_cubic = (cfg->interpolation == 1) || (cfg->cycle_type == 2);  // FMG hard-codes cubic (:631)
  if (need_cubic && mg->lv[0].geom.dist) {
    ctx->err = "cubic prolongation (interpolate_cubic / FMG) is a whole-grid spline: not available on a slab-decomposed grid";
    nf_mg_destroy(mg);
    return NF_ERR_UNSUPPORTED;
  }
  if (cfg->smoother >= 2 && (mg->lv[0].geom.dist || nl != 1)) {
    ctx->err = "the sequential Gauss-Seidel smoothers run on a single slab only";
    nf_mg_destroy(mg);
    return NF_ERR_UNSUPPORTED;
  }
  if (mg->lv[0].geom.dist) {  // peer-memory halo staging (no-op without p2p)
    // one exchange may carry a level's iterate and the next level's right-hand side (mg_publish_rhs)
    int st = nf_p2p_reserve_stage(team, (size_t)NF_HALO * (mg->lv[0].geom.ld + (mg->lv.size() > 1 ? mg->lv[1].geom.ld : 0)));
    if (st != NF_OK) { nf_mg_destroy(mg); return st; }
  }
  bool ok = true;
  for (size_t l = 0; l < mg->lv.size() && ok; ++l) {
    MgLevel& L = mg->lv[l];
    L.s.resize(nl);
    for (int k = 0; k < nl && ok; ++k) {
      MgSlab& S = L.s[k];
      const size_t e = L.geom.elems(team->local[k]), em = L.geom.max_elems();
      ok = ok && dev_alloc(team, &S.r, e, em) && dev_alloc(team, &S.x2, e, em);
      if (l > 0) {
        ok = ok && dev_alloc(team, &S.x, e, em) && dev_alloc(team, &S.b, e, em) && dev_alloc(team, &S.d_u, e, em) &&
             dev_alloc(team, &S.d_v, e, em);
        S.owns_x = S.owns_b = S.owns_d = true;
      }
      if (cfg->smoother == 0) ok = ok && dev_alloc(team, &S.inv, e, em);
    }
    if (ok && need_cubic && l + 1 < mg->lv.size()) {
      const int mc = mg->lv[l + 1].geom.nx, m = L.geom.nx;
      std::vector<double> band;
      std::vector<int> start;
      int W = 0;
      build_interp_band(mc, m, 24, band, start, W);
      L.W = W;
      ok = ok && cudaMalloc(&L.band, band.size() * sizeof(double)) == cudaSuccess &&
           cudaMalloc(&L.start, start.size() * sizeof(int)) == cudaSuccess &&
           cudaMalloc(&L.ptmp, (size_t)m * mg->lv[l + 1].geom.ld * sizeof(double)) == cudaSuccess;
      if (ok)
        ok = cudaMemcpy(L.band, band.data(), band.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(L.start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    }
  }
  const MgLevel& C = mg->lv.back();
  mg->coarse_N = C.geom.nx * C.geom.ny;
  if (ok && (mg->coarse_N > 4096 || C.geom.dist)) {
    ctx->err = "coarsest level too large for the dense coarse solve (nx*ny must be <= 4096)";
    nf_mg_destroy(mg);
    return NF_ERR_UNSUPPORTED;
  }
  mg->coarse_A.assign(nl, nullptr);
  mg->coarse_inv.assign(nl, nullptr);
  mg->scal.assign(nl, nullptr);
  for (int k = 0; k < nl && ok; ++k)
    ok = dev_alloc(team, &mg->coarse_A[k], (size_t)mg->coarse_N * mg->coarse_N, (size_t)mg->coarse_N * mg->coarse_N) &&
         dev_alloc(team, &mg->coarse_inv[k], (size_t)mg->coarse_N * mg->coarse_N, (size_t)mg->coarse_N * mg->coarse_N) &&
         dev_alloc(team, &mg->scal[k], 8, 8);
  ok = ok && cudaMallocHost(&mg->scal_host, 8 * sizeof(double)) == cudaSuccess;
  if (!ok) {
    ctx->err = std::string("multigrid allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    nf_mg_destroy(mg);
    return NF_ERR_ALLOC;
  }
  // Coarse end of the V-cycle in one kernel: the levels of <= NF_TAIL_MAX_N cells per side, when they are whole
  // (replicated) grids smoothed by red-black SOR with full weighting / bilinear transfer and the chain ends in the dense solve
  {
    const char* env = getenv("NF_MG_TAIL");
    const bool want = !(env && env[0] == '0') && cfg->smoother == 0 && cfg->restriction == 0 && cfg->interpolation == 0 &&
                      C.geom.nx <= cfg->coarsest;
    if (want) {
      int first = -1;
      for (size_t l = 1; l < mg->lv.size(); ++l)
        if (mg->lv[l].geom.nx <= NF_TAIL_MAX_N && !mg->lv[l].geom.dist) { first = (int)l; break; }
      if (first >= 0 && (int)mg->lv.size() - first >= 2 && (int)mg->lv.size() - first <= NF_TAIL_MAX_LEVELS) {
        size_t bytes = 0;
        for (size_t l = first; l < mg->lv.size(); ++l)
          bytes += sizeof(double) * ((size_t)(mg->lv[l].geom.nx + 2) * (mg->lv[l].geom.ny + 2) +
                                     7 * (size_t)mg->lv[l].geom.nx * mg->lv[l].geom.ny);
        if (bytes + 2048 <= NF_TAIL_MAX_SMEM) mg->tail_level = first;
      }
    }
  }
  *out = mg;
  return NF_OK;
}

extern "C" int nf_mg_create(nf_ctx* ctx, nf_mg** out, int nx, int ny, int ld, const nf_mg_config* cfg) {
  nf_team* team = nullptr;
  NF_TRY(nf_team_create_local(ctx, 1, &team));
  int st = nfi_mg_create(team, out, nx, ny, ld, cfg);
  if (st != NF_OK) { nf_team_destroy(team); return st; }
  (*out)->owns_team = true;
  return NF_OK;
}

extern "C" int nf_mg_num_levels(nf_mg* mg) { return mg ? (int)mg->lv.size() : 0; }

extern "C" int nf_mg_level_shape(nf_mg* mg, int level, int* nx, int* ny, int* ld) {
  if (!mg || level < 0 || level >= (int)mg->lv.size()) return NF_ERR_ARG;
  if (nx) *nx = mg->lv[level].geom.nx;
  if (ny) *ny = mg->lv[level].geom.ny;
  if (ld) *ld = mg->lv[level].geom.ld;
  return NF_OK;
}

// level arrays of the first local slab (tests / inspection): which = 0 d_u, 1 d_v, 2 x, 3 b, 4 r
extern "C" const double* nf_mg_level_array(nf_mg* mg, int level, int which) {
  if (!mg || level < 0 || level >= (int)mg->lv.size()) return nullptr;
  const MgSlab& S = mg->lv[level].s[0];
  switch (which) {
    case 0: return S.d_u;
    case 1: return S.d_v;
    case 2: return S.x;
    case 3: return S.b;
    case 4: return S.r;
  }
  return nullptr;
}

static std::vector<double*> field_of(MgLevel& L, double* MgSlab::*m) {
  std::vector<double*> v;
  for (MgSlab& S : L.s) v.push_back(S.*m);
  return v;
}

// grid a rank restricts INTO at level l+1: the coarse level's own slab grid when it is cut, else the full
// (replicated) array with the row range induced by the finer partition
static nf_grid restrict_target(const MgLevel& C, int rank) {
  nf_grid g = C.geom.grid(rank);
  g.gb = C.rgb[rank];
  g.ge = C.rge[rank];
  return g;
}

// d_u, d_v: per local slab; on a cut level they must be valid NF_HALO-1 rows beyond the owned rows
int nfi_mg_setup(nf_mg* mg, double* const* d_u, double* const* d_v) {
  nf_ctx* ctx = mg->ctx;
  nf_team* team = mg->team;
  const int nl = nlocal(mg);
  for (int k = 0; k < nl; ++k) {
    NF_REQUIRE(ctx, d_u[k] && d_v[k], "NULL coefficient array");
    mg->lv[0].s[k].d_u = d_u[k];
    mg->lv[0].s[k].d_v = d_v[k];
  }
  if (mg->lv[0].geom.dist) {  // the callers' coefficients are exact NF_HALO-1 rows out: complete the halo
    std::vector<double*> du = field_of(mg->lv[0], &MgSlab::d_u), dv = field_of(mg->lv[0], &MgSlab::d_v);
    NF_TRY(nf_team_exchange(team, mg->lv[0].geom, du.data(), NF_HALO));
    NF_TRY(nf_team_exchange(team, mg->lv[0].geom, dv.data(), NF_HALO));
  }
  for (size_t l = 0; l + 1 < mg->lv.size(); ++l) {
    MgLevel& L = mg->lv[l];
    MgLevel& C = mg->lv[l + 1];
    for (int k = 0; k < nl; ++k) {
      const int r = team->local[k];
      const nf_grid gf = L.geom.grid(r), gc = restrict_target(C, r);
      NF_TRY(nfi_restrict_coeffs(ctx, &gf, L.s[k].d_u, L.s[k].d_v, &gc, C.s[k].d_u, C.s[k].d_v));
    }
    if (L.geom.dist) {
      std::vector<double*> du = field_of(C, &MgSlab::d_u), dv = field_of(C, &MgSlab::d_v);
      if (C.geom.dist) {
        NF_TRY(nf_team_exchange(team, C.geom, du.data(), NF_HALO));
        NF_TRY(nf_team_exchange(team, C.geom, dv.data(), NF_HALO));
      } else {
        NF_TRY(nf_team_share_rows(team, C.geom.ld, C.geom.nx, C.rgb, C.rge, du.data(), 1));
        NF_TRY(nf_team_share_rows(team, C.geom.ld, C.geom.nx, C.rgb, C.rge, dv.data(), 0));
      }
    }
  }
  if (mg->cfg.smoother == 0)
    for (MgLevel& L : mg->lv)
      for (int k = 0; k < nl; ++k) {
        const nf_grid g = L.geom.grid_ext(team->local[k], NF_HALO - 1);  // row i needs d_u[i+1]
        NF_TRY(nfi_inv_diag(ctx, &g, L.s[k].d_u, L.s[k].d_v, L.s[k].inv));
      }
  MgLevel& C = mg->lv.back();
  if (C.geom.nx <= mg->cfg.coarsest)
    for (int k = 0; k < nl; ++k) {
      const char* envs = getenv("NF_COARSE_INVERT_SMALL");  // =0: the general kernel for every N (tests compare the two)
      if (mg->coarse_N <= CI_MAX && !(envs && envs[0] == '0')) {
        k_coarse_invert_small<<<1, 1024, 0, ctx->stream>>>(C.geom.grid(team->local[k]), C.s[k].d_u, C.s[k].d_v,
                                                           mg->coarse_inv[k], mg->coarse_N);
        NF_LAUNCH_CHECK(ctx);
        continue;
      }
      const size_t work = 2 * (size_t)mg->coarse_N * mg->coarse_N * sizeof(double);
      const int use_smem = work <= 200 * 1024 ? 1 : 0;
      static size_t attr = 48 * 1024;
      if (use_smem && work > attr) {
        NF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_coarse_invert, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)work));
        attr = work;
      }
      k_coarse_invert<<<1, 1024, use_smem ? work : 0, ctx->stream>>>(C.geom.grid(team->local[k]), C.s[k].d_u, C.s[k].d_v,
                                                                     mg->coarse_A[k], mg->coarse_inv[k], mg->coarse_N, use_smem);
      NF_LAUNCH_CHECK(ctx);
    }
  mg->setup_done = true;
  return NF_OK;
}

extern "C" int nf_mg_setup(nf_mg* mg, const double* d_u, const double* d_v) {
  if (!mg) return NF_ERR_ARG;
  NF_REQUIRE(mg->ctx, nlocal(mg) == 1, "nf_mg_setup is the single-slab entry point");
  double* du = const_cast<double*>(d_u);
  double* dv = const_cast<double*>(d_v);
  return nfi_mg_setup(mg, &du, &dv);
}

// =============================================================================================
// cycles.  Level-0 solution / right-hand side live in lv[0].s[k].x / .b (set by the callers).
// =============================================================================================
// n smoothing sweeps on level l; on a cut level the iterate's halo is valid on entry and on return
// extra (single slab only): work fused behind the last smoother launch, see nf_smooth_extra
// defer_exchange: the caller exchanges the iterate's halo after the LAST launch itself (merged with another field)
// red: sum these 2 scalars over the ranks in the last launch's halo exchange when the transport can (-> mg->norm_reduced)
static int mg_smooth(nf_mg* mg, int l, int n, nf_smooth_extra* extra = nullptr /* one per local slab */,
                     bool defer_exchange = false, double* red = nullptr) {
  nf_ctx* ctx = mg->ctx;
  nf_team* team = mg->team;
  MgLevel& L = mg->lv[l];
  const int nl = nlocal(mg);
  if (mg->cfg.smoother >= 2) {  // sequential SOR sweeps (gauss_seidel.py:307-367) as block wavefronts; single slab
    const nf_grid g = L.geom.grid(team->local[0]);
    return nfi_gs_lex(ctx, &g, L.s[0].x, L.s[0].b, L.s[0].d_u, L.s[0].d_v, mg->cfg.omega, n, mg->cfg.smoother == 3);
  }
  if (mg->cfg.smoother == 0) {
    int left = n;
    if (left == 0)
      for (int k = 0; k < nl; ++k) {
        const nf_grid g = L.geom.grid(team->local[k]);
        NF_TRY(nfi_rbsor_fused_x(ctx, &g, &L.s[k].x, &L.s[k].x2, L.s[k].b, L.s[k].d_u, L.s[k].d_v, L.s[k].inv,
                                 mg->cfg.omega, 0, nullptr));
      }
    while (left > 0) {
      const int ns = left >= 3 ? 3 : left;
      const bool timed = mg->timing && l == 0 && ns == 3 && nl == 1;
      if (timed) {
        if (mg->ev_used + 2 > mg->ev.size()) {
          cudaEvent_t a, b;
          cudaEventCreate(&a); cudaEventCreate(&b);
          mg->ev.push_back(a); mg->ev.push_back(b);
        }
        cudaEventRecord(mg->ev[mg->ev_used], ctx->stream);
      }
      for (int k = 0; k < nl; ++k) {
        MgProf prof_(mg, l, PH_SMOOTH);
        const nf_grid g = L.geom.grid(team->local[k]);
        // the fused prolongation rides on the FIRST launch of the call, the residual work on the LAST one
        const bool first = (left == n), last = (left == ns);
        nf_smooth_extra ex1;
        nf_smooth_extra* pe = nullptr;
        if (extra && (last || (first && extra[k].prolong_c))) {
          ex1 = extra[k];
          if (!first) ex1.prolong_c = nullptr;
          if (!last) { ex1.mode = 0; ex1.in_norm_out = nullptr; }
          pe = &ex1;
        }
        NF_TRY(nfi_rbsor_fused_x(ctx, &g, &L.s[k].x, &L.s[k].x2, L.s[k].b, L.s[k].d_u, L.s[k].d_v, L.s[k].inv,
                                 mg->cfg.omega, ns, pe));
        if (pe && last) { extra[k].fused = ex1.fused; extra[k].in_norm_fused = ex1.in_norm_fused; }
        if (pe && first) extra[k].prolong_fused = ex1.prolong_fused;
      }
      if (timed) {
        cudaEventRecord(mg->ev[mg->ev_used + 1], ctx->stream);
        mg->ev_used += 2;
      }
      if (L.geom.dist && !(defer_exchange && left == ns)) {
        MgProf prof_(mg, l, PH_EXCH_X);
        std::vector<double*> x = field_of(L, &MgSlab::x);
        int st = NF_ERR_UNSUPPORTED;
        if (red && left == ns && nl == 1 && extra && extra[0].fused && nf_p2p_active(team)) {
          const LevelGeom* gp = &L.geom;
          const int depth = NF_SMOOTH_HALO;
          st = nf_p2p_exchange_multi(team, 1, &gp, x.data(), &depth, nullptr, nullptr, red, 2);
          if (st == NF_OK) mg->norm_reduced = true;
        }
        if (st == NF_ERR_UNSUPPORTED) st = nf_team_exchange(team, L.geom, x.data(), NF_SMOOTH_HALO);
        NF_TRY(st);
      }
      left -= ns;
    }
    return NF_OK;
  }
  // weighted Jacobi: one halo row per iteration
  for (int it = 0; it < (n > 0 ? n : 1); ++it) {
    for (int k = 0; k < nl; ++k) {
      const nf_grid g = L.geom.grid(team->local[k]);
      NF_TRY(nfi_jacobi(ctx, &g, L.s[k].x, L.s[k].r, L.s[k].b, L.s[k].d_u, L.s[k].d_v, mg->cfg.omega, n > 0 ? 1 : 0));
    }
    if (L.geom.dist) {
      std::vector<double*> x = field_of(L, &MgSlab::x);
      NF_TRY(nf_team_exchange(team, L.geom, x.data(), NF_SMOOTH_HALO));
    }
  }
  return NF_OK;
}

static int mg_coarse_solve(nf_mg* mg, int l) {
  nf_ctx* ctx = mg->ctx;
  MgLevel& L = mg->lv[l];
  const int N = mg->coarse_N;
  const int threads = 128, blocks = (N * 32 + threads - 1) / threads;
  MgProf prof_(mg, l, PH_COARSE);
  for (int k = 0; k < nlocal(mg); ++k) {
    nf_launch(k_coarse_apply, blocks, threads, 0, ctx->stream, true, L.geom.grid(mg->team->local[k]), mg->coarse_inv[k], L.s[k].b,
                                                        L.s[k].x, N);
    NF_LAUNCH_CHECK(ctx);
  }
  return NF_OK;
}

// x_l (+)= P x_{l+1}; on a cut level the halo rows the smoother reads are produced locally as well
static int mg_prolong(nf_mg* mg, int l, int cubic, int add) {
  nf_team* team = mg->team;
  MgLevel& L = mg->lv[l];
  MgLevel& C = mg->lv[l + 1];
  MgProf prof_(mg, l, PH_PROLONG);
  for (int k = 0; k < nlocal(mg); ++k) {
    const int r = team->local[k];
    const nf_grid gc = C.geom.grid(r), gf = L.geom.grid_ext(r, NF_SMOOTH_HALO);
    if (!cubic) NF_TRY(nfi_prolong_linear(mg->ctx, &gc, C.s[k].x, &gf, L.s[k].x, add));
    else NF_TRY(nfi_prolong_banded(mg->ctx, &gc, C.s[k].x, &gf, L.s[k].x, L.ptmp, gc.ld, L.band, L.start, L.W, add));
  }
  return NF_OK;
}

// after a restriction into level l+1: make the new right-hand side visible where the coarse level needs it
// with_x: level l's iterate is waiting for its halo exchange as well (mg_smooth's defer_exchange) -- one launch for both over
// peer memory; zero_x: that launch also clears the halo rows of level l+1's iterate (its own rows were cleared by the
// restriction kernel)
static int mg_publish_rhs(nf_mg* mg, int l, bool with_x = false, bool zero_x = false) {
  MgLevel& L = mg->lv[l];
  MgLevel& C = mg->lv[l + 1];
  if (!L.geom.dist) return NF_OK;
  if (with_x || zero_x) {  // single local slab, cut coarse level (see mg_cycle)
    MgProf prof_(mg, l, PH_EXCH_X);
    const LevelGeom* gp[2] = {&C.geom, &L.geom};
    double* f[2] = {C.s[0].b, L.s[0].x};
    const int depth[2] = {NF_SMOOTH_HALO, NF_SMOOTH_HALO};
    int st = nf_p2p_exchange_multi(mg->team, with_x ? 2 : 1, gp, f, depth, zero_x ? &C.geom : nullptr,
                                   zero_x ? C.s[0].x : nullptr, nullptr, 0);
    if (st != NF_ERR_UNSUPPORTED) return st;
    if (with_x) {
      std::vector<double*> x = field_of(L, &MgSlab::x);
      NF_TRY(nf_team_exchange(mg->team, L.geom, x.data(), NF_SMOOTH_HALO));
    }
    if (zero_x) NF_TRY(nfi_fill(mg->ctx, C.s[0].x, C.geom.elems(mg->team->local[0]), 0.0));
  }
  MgProf prof_(mg, l + 1, PH_EXCH_B);
  std::vector<double*> b = field_of(C, &MgSlab::b);
  if (C.geom.dist) return nf_team_exchange(mg->team, C.geom, b.data(), NF_SMOOTH_HALO);
  return nf_team_share_rows(mg->team, C.geom.ld, C.geom.nx, C.rgb, C.rge, b.data(), 0);
}

// levels tail_level .. coarsest of a V-cycle in one kernel per local slab (the level's right-hand side is in place, its
// solution is written; same bits as the launch-by-launch recursion)
static int mg_tail(nf_mg* mg) {
  nf_ctx* ctx = mg->ctx;
  MgProf prof_(mg, mg->tail_level, PH_TAIL);
  for (int k = 0; k < nlocal(mg); ++k) {
    nf_tail_args a;
    a.nlev = (int)mg->lv.size() - mg->tail_level;
    for (int q = 0; q < a.nlev; ++q) {
      MgLevel& L = mg->lv[mg->tail_level + q];
      MgSlab& S = L.s[k];
      nf_tail_level& T = a.lv[q];
      T.x = S.x; T.b = S.b; T.d_u = S.d_u; T.d_v = S.d_v; T.inv = S.inv;
      T.nx = L.geom.nx; T.ny = L.geom.ny; T.ld = L.geom.ld; T.dx = L.geom.dx; T.dy = L.geom.dy;
    }
    a.N = mg->coarse_N;
    a.coarse_inv = mg->coarse_inv[k];
    a.rho = mg->cfg.rho;
    a.omega = mg->cfg.omega;
    a.pre = mg->cfg.pre;
    a.post = mg->cfg.post;
    NF_TRY(nfi_mg_tail(ctx, &a));
  }
  return NF_OK;
}

// one V (kind 0) or W (kind 1) cycle on level l: multigrid.py:304-432 / :434-560
// want_norm (level 0 of the 'v' / 'w' loop): ask the last post-smoothing launch for the residual norms
// (-> mg->scal[0][0..1]); *norm_fused reports whether that happened
// part: 0 = the whole cycle; 1 = head only (pre-smoothing, residual, restriction, coarse x = 0, RHS published);
//       2 = the rest (coarse levels, prolongation, post-smoothing).  in_norm (head of level 0): ask the pre-smoothing
//       launch for the residual norms of its INPUT iterate (-> mg->scal[k][0..1]); *in_norm_fused reports whether it did.
// from_zero: the level's iterate is known to be zero on entry (the recursion of a cycle; not the cycles FMG runs on a
// prolonged iterate) -- the precondition of the single-kernel coarse end
static int mg_cycle(nf_mg* mg, int l, int kind, bool want_norm = false, bool* norm_fused = nullptr, int part = 0,
                    bool in_norm = false, bool* in_norm_fused = nullptr, bool from_zero = false) {
  nf_ctx* ctx = mg->ctx;
  nf_team* team = mg->team;
  MgLevel& L = mg->lv[l];
  const int nl = nlocal(mg);
  if (norm_fused) *norm_fused = false;
  if (in_norm_fused) *in_norm_fused = false;
  if (l == 0) mg->norm_reduced = false;
  if (l == mg->tail_level && kind == 0 && part == 0 && !want_norm && !in_norm && from_zero) return mg_tail(mg);
  if (L.geom.nx <= mg->cfg.coarsest || l + 1 == (int)mg->lv.size()) return part == 1 ? NF_OK : mg_coarse_solve(mg, l);
  MgLevel& C = mg->lv[l + 1];
  if (part != 2) {
    std::vector<nf_smooth_extra> pre(nl);
    const bool want_pre = mg->cfg.smoother == 0 && mg->cfg.restriction == 0;
    // The coarse level starts from a zero guess (multigrid.py:374).  On an unsplit level the kernel that writes the restricted
    // right-hand side zeroes the coarse iterate cell by cell as well; on slabs the whole array (halo rows included) is filled.
    const bool tail_next = (l + 1 == mg->tail_level && kind == 0);  // the tail kernel starts from zero by itself
    // cut level over peer memory: the halo exchange of the pre-smoothed iterate and the one that publishes the coarse
    // right-hand side are one launch, which clears the halo rows of the coarse iterate on the way
    const char* envm = getenv("NF_MG_MERGED_EXCHANGE");
    const bool merged = L.geom.dist && C.geom.dist && nl == 1 && mg->cfg.smoother == 0 && mg->cfg.pre > 0 && part == 0 &&
                        nf_p2p_active(team) && !(envm && envm[0] == '0');
    const bool zero_by_restriction = !tail_next && (!L.geom.dist || merged) && mg->cfg.restriction == 0;
    if (want_pre)
      for (int k = 0; k < nl; ++k) {
        pre[k].mode = 2;
        pre[k].gc = restrict_target(C, team->local[k]);
        pre[k].coarse_b = C.s[k].b;
        if (zero_by_restriction) pre[k].coarse_x_zero = C.s[k].x;
        if (in_norm) pre[k].in_norm_out = mg->scal[k];
      }
    NF_TRY(mg_smooth(mg, l, mg->cfg.pre, want_pre ? pre.data() : nullptr, merged));
    bool x_pending = merged;
    if (merged && !(want_pre && pre[0].fused)) {
      // the restriction is a kernel of its own here and reads the pre-smoothed iterate's halo rows: exchange them now
      MgProf prof_(mg, l, PH_EXCH_X);
      std::vector<double*> x = field_of(L, &MgSlab::x);
      NF_TRY(nf_team_exchange(team, L.geom, x.data(), NF_SMOOTH_HALO));
      x_pending = false;
    }
    if (in_norm_fused) {
      bool all = want_pre && in_norm;
      for (int k = 0; k < nl; ++k) all = all && pre[k].in_norm_fused;
      *in_norm_fused = all;  // same decision on every rank: it depends on the level geometry only
    }
    for (int k = 0; k < nl; ++k) {
      const int r = team->local[k];
      const nf_grid gf = L.geom.grid(r), gc = restrict_target(C, r);
      if (pre[k].fused) {
        // coarse right-hand side already written by the smoother
      } else if (mg->cfg.restriction == 0) {
        MgProf prof_(mg, l, PH_RESTRICT);
        NF_TRY(nfi_residual_restrict_fw(ctx, &gf, L.s[k].x, L.s[k].b, L.s[k].d_u, L.s[k].d_v, &gc, C.s[k].b,
                                        zero_by_restriction ? C.s[k].x : nullptr));
      } else {
        NF_TRY(nfi_residual(ctx, &gf, L.s[k].x, L.s[k].b, L.s[k].d_u, L.s[k].d_v, L.s[k].r));
        NF_TRY(nfi_restrict_inject(ctx, &gf, L.s[k].r, &gc, C.s[k].b));
      }
      if (!tail_next && !zero_by_restriction) {
        MgProf prof_(mg, l + 1, PH_FILL);
        NF_TRY(nfi_fill(ctx, C.s[k].x, C.geom.elems(r), 0.0));
      }
    }
    NF_TRY(mg_publish_rhs(mg, l, x_pending, merged && zero_by_restriction));
    if (part == 1) return NF_OK;
  }
  const int reps = (kind == 1) ? 2 : 1;
  for (int rep = 0; rep < reps; ++rep)
    NF_TRY(mg_cycle(mg, l + 1, kind, false, nullptr, 0, false, nullptr, /*from_zero=*/rep == 0));
  std::vector<nf_smooth_extra> post(nl);
  const bool want_post = want_norm && mg->cfg.smoother == 0;
  // x += P x_coarse (multigrid.py:405-415): a pass of its own, or -- bilinear prolongation, streaming smoother -- added
  // to the iterate on the way into the first post-smoothing launch
  // (streaming kernel: block rule at load + a strips-only launch; TMA kernel: every cell at tile set-up, no launch)
  int fuse_kind = (mg->cfg.smoother == 0 && mg->cfg.interpolation == 0) ? -1 : 0;
  for (int k = 0; k < nl && fuse_kind != 0; ++k) {
    const nf_grid g = L.geom.grid(team->local[k]);
    const int kd = nfi_rbsor_can_fuse_prolong(&g, mg->cfg.post, L.s[k].inv != nullptr);
    fuse_kind = (fuse_kind == -1 || fuse_kind == kd) ? kd : 0;  // the slabs of a process must agree
  }
  if (fuse_kind < 0) fuse_kind = 0;
  const bool fuse_prolong = fuse_kind != 0;
  if (fuse_kind != 2)
    NF_TRY(mg_prolong(mg, l, mg->cfg.interpolation, fuse_kind == 1 ? 2 : 1));  // 2: only the strips outside the block rule
  if (want_post || fuse_prolong)
    for (int k = 0; k < nl; ++k) {
      if (want_post) {
        post[k].mode = 1;
        post[k].out = mg->scal[k];
      }
      if (fuse_prolong) {
        post[k].prolong_c = C.s[k].x;
        post[k].prolong_gc = C.geom.grid(team->local[k]);
      }
    }
  const char* envr = getenv("NF_MG_MERGED_ALLREDUCE");
  const bool red = want_post && l == 0 && L.geom.dist && nl == 1 && mg->cfg.post > 0 && mg->cfg.post <= 3 &&
                   !(envr && envr[0] == '0');
  NF_TRY(mg_smooth(mg, l, mg->cfg.post, (want_post || fuse_prolong) ? post.data() : nullptr, false, red ? mg->scal[0] : nullptr));
  if (norm_fused) {
    bool all = want_post;
    for (int k = 0; k < nl; ++k) all = all && post[k].fused;
    *norm_fused = all;  // same decision on every rank: it depends on the level geometry only
  }
  return NF_OK;
}

// One cycle at level 0 through a CUDA graph when possible (see nf_mg::graph_exec).  The launch sequence of a
// cycle is static: ~50 launches (10 levels) collapse into one graph launch, which removes the CPU launch cost and
// most of the inter-kernel gaps on the small levels.
// part / want_norm as in mg_cycle (part 2 with want_norm = false is the body of a "lookahead norm" cycle)
static int mg_cycle_top(nf_mg* mg, int kind, bool* norm_fused, int part = 0, bool want_norm = true) {
  nf_ctx* ctx = mg->ctx;
  MgLevel& L = mg->lv[0];
  const char* env = getenv("NF_MG_GRAPH");
  // slab runs under torchrun: the NCCL halo exchanges are captured into the graph as well (every rank replays the
  // same sequence); opt out with NF_MG_GRAPH_DIST=0
  const char* envd = getenv("NF_MG_GRAPH_DIST");
  const bool dist_ok = !L.geom.dist || (mg->team->nccl != nullptr && !(envd && envd[0] == '0'));
  const bool allowed = mg->use_graph && !(env && env[0] == '0') && nlocal(mg) == 1 && dist_ok &&
                       mg->cfg.smoother == 0 && !mg->timing && !mg->prof &&
                       ((mg->cfg.pre + 2) / 3 + (mg->cfg.post + 2) / 3) % 2 == 0;  // even number of x/x2 swaps
  if (!allowed) return mg_cycle(mg, 0, kind, want_norm, norm_fused, part);
  MgSlab& S = L.s[0];
  const void* key[5] = {S.x, S.b, S.d_u, S.d_v, S.x2};
  bool same = mg->graph_exec && mg->graph_kind == kind && mg->graph_part == part && mg->graph_want_norm == want_norm;
  for (int q = 0; q < 5 && same; ++q) same = (key[q] == mg->graph_key[q]);
  if (same) {
    NF_CHECK_CUDA(ctx, cudaGraphLaunch(mg->graph_exec, ctx->stream));
    ctx->launches += mg->graph_nodes;
    *norm_fused = mg->graph_norm_fused;
    mg->norm_reduced = mg->graph_norm_reduced;
    if (mg->graph_swapped) { double* t = S.x; S.x = S.x2; S.x2 = t; }  // what the launches did to the host's pointers
    return NF_OK;
  }
  if (mg->warm_cycles < 1) {  // first cycle ever: plain launches (function attributes, tensor-map encoder, ...)
    mg->warm_cycles++;
    return mg_cycle(mg, 0, kind, want_norm, norm_fused, part);
  }
  if (mg->graph_exec) { cudaGraphExecDestroy(mg->graph_exec); mg->graph_exec = nullptr; }
  if (!mg->cap_stream) NF_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&mg->cap_stream, cudaStreamNonBlocking));
  cudaStream_t orig = ctx->stream;
  const long long l0 = ctx->launches;
  NF_CHECK_CUDA(ctx, cudaStreamBeginCapture(mg->cap_stream, cudaStreamCaptureModeRelaxed));
  ctx->stream = mg->cap_stream;
  bool nf = false;
  int st = mg_cycle(mg, 0, kind, want_norm, &nf, part);
  ctx->stream = orig;
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(mg->cap_stream, &graph);
  if (st != NF_OK || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    mg->use_graph = false;  // fall back to plain launches for good
    ctx->launches = l0;
    if (st != NF_OK) return st;
    return mg_cycle(mg, 0, kind, want_norm, norm_fused, part);
  }
  mg->graph_nodes = ctx->launches - l0;
  mg->graph_norm_fused = nf;
  mg->graph_norm_reduced = mg->norm_reduced;
  mg->graph_swapped = (S.x != key[0]);
  ctx->launches = l0;
  ce = cudaGraphInstantiate(&mg->graph_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) {
    cudaGetLastError();
    mg->graph_exec = nullptr;
    mg->use_graph = false;
    return mg_cycle(mg, 0, kind, want_norm, norm_fused, part);
  }
  for (int q = 0; q < 5; ++q) mg->graph_key[q] = key[q];
  mg->graph_kind = kind;
  mg->graph_part = part;
  mg->graph_want_norm = want_norm;
  NF_CHECK_CUDA(ctx, cudaGraphLaunch(mg->graph_exec, ctx->stream));
  ctx->launches += mg->graph_nodes;
  *norm_fused = nf;
  return NF_OK;
}

// ||b - A x|| and ||b|| on level l over the whole grid (host values; multigrid.py:185-189, :652-676).
// *b_norm < 0 on entry: ||b|| is not known yet and is computed in the same pass; otherwise it is kept.
// sync == 0: leave sum r^2, sum b^2 in scal[k][0..1] (already reduced over the team) and do not synchronise.
// have_norms: sum r^2, sum b^2 were already left in mg->scal[0][0..1] by the post-smoother (no kernel needed)
static int mg_rel_residual(nf_mg* mg, int l, double* r_norm, double* b_norm, int sync, bool have_norms = false) {
  nf_ctx* ctx = mg->ctx;
  nf_team* team = mg->team;
  MgLevel& L = mg->lv[l];
  const int nl = nlocal(mg);
  int with_b = (*b_norm < 0.0) ? 1 : 0;
  if (have_norms) {
    with_b = 1;
    MgProf prof_(mg, l, PH_ALLREDUCE);
    if (L.geom.dist && !mg->norm_reduced) NF_TRY(nf_team_allreduce(team, mg->scal.data(), 2));
  } else {
    for (int k = 0; k < nl; ++k) {
      MgProf prof_(mg, l, PH_NORM);
      const nf_grid g = L.geom.grid(team->local[k]);
      NF_TRY(nfi_residual_norms(ctx, &g, L.s[k].x, L.s[k].b, L.s[k].d_u, L.s[k].d_v, L.s[k].r, with_b, mg->scal[k]));
    }
    MgProf prof_(mg, l, PH_ALLREDUCE);
    if (L.geom.dist) NF_TRY(nf_team_allreduce(team, mg->scal.data(), with_b ? 2 : 1));
  }
  if (!sync) return NF_OK;
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(mg->scal_host, mg->scal[0], 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (mg->timing) mg_harvest_timing(mg);
  *r_norm = sqrt(mg->scal_host[0]);
  if (with_b) *b_norm = sqrt(mg->scal_host[1]);
  return NF_OK;
}

// ---- device-side convergence loop -----------------------------------------------------------------------------------
// MultiGridSolver.solve's `for cycle in range(max_iterations): ...; if rel < tol: break` (multigrid.py:185-240) as a
// CUDA-graph conditional WHILE node: the body is one cycle at level 0 + the residual norms of its result (fused into
// the post-smoother where the level-0 kernel can, one reduction kernel otherwise; all-reduced over the slabs) + a
// one-thread kernel that counts the cycle, evaluates the reference's test in the reference's arithmetic
// (sqrt(sum r^2) / sqrt(sum b^2) < tol) and arms or disarms the node.  The host launches the graph once per solve and
// reads {sum r^2, sum b^2, cycles} once: no per-cycle D2H + synchronisation, no per-cycle graph launch.
// scal[0][4] is the cycle counter (a double so that one 40-byte copy brings everything back).
__global__ void k_mg_loop_check(cudaGraphConditionalHandle handle, double* scal, double tol, int max_it) {
  nf_pdl_entry();
  const double c = scal[4] + 1.0;
  scal[4] = c;
  const double rn = sqrt(scal[0]), bn = sqrt(scal[1]);
  const double rel = bn > 0.0 ? rn / bn : rn;
  cudaGraphSetConditional(handle, (!(rel < tol) && c < (double)max_it) ? 1u : 0u);
}

// *used = false: not available here (the caller runs the host loop).  On success the iterate is in L.s[0].x, the
// cycle count in *cycles and the norms of the last iterate in *rn / *bn.
static int mg_device_loop(nf_mg* mg, int* cycles, double* rn, double* bn, bool* norm_fused, bool* used) {
  nf_ctx* ctx = mg->ctx;
  nf_team* team = mg->team;
  MgLevel& L = mg->lv[0];
  *used = false;
  const char* env = getenv("NF_MG_DEVICE_LOOP");
  const char* envg = getenv("NF_MG_GRAPH");
  const char* envd = getenv("NF_MG_GRAPH_DIST");
  const bool force = env && env[0] == '2';
  // slabs: the peer-memory transport keeps its sequence numbers on the device (replayable, every rank takes the same
  // decision because the all-reduced norms are bit-identical); NCCL inside a WHILE body only on request (=2)
  const bool dist_ok = !L.geom.dist || (!(envd && envd[0] == '0') && team->nccl != nullptr && (nf_p2p_active(team) || force));
  const bool allowed = mg->use_loop && mg->use_graph && !(env && env[0] == '0') && !(envg && envg[0] == '0') &&
                       nlocal(mg) == 1 && dist_ok && mg->cfg.smoother == 0 && !mg->timing && !mg->prof && mg->cfg.max_iterations >= 1 &&
                       mg->warm_cycles >= 1 && (mg->cfg.cycle_type == 0 || mg->cfg.cycle_type == 1) &&
                       ((mg->cfg.pre + 2) / 3 + (mg->cfg.post + 2) / 3) % 2 == 0;
  if (!allowed) return NF_OK;
  MgSlab& S = L.s[0];
  const int kind = mg->cfg.cycle_type;
  const void* key[5] = {S.x, S.b, S.d_u, S.d_v, S.x2};
  bool same = mg->loop_exec && mg->loop_kind == kind;
  for (int q = 0; q < 5 && same; ++q) same = (key[q] == mg->loop_key[q]);
  if (!same) {
    if (mg->loop_exec) { cudaGraphExecDestroy(mg->loop_exec); mg->loop_exec = nullptr; }
    if (!mg->cap_stream) NF_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&mg->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    cudaGraphConditionalHandle handle;
    cudaGraphNode_t node;
    cudaGraphNodeParams np = {};
    bool ok = cudaGraphCreate(&graph, 0) == cudaSuccess &&
              cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
    if (ok) {
      np.type = cudaGraphNodeTypeConditional;
      np.conditional.handle = handle;
      np.conditional.type = cudaGraphCondTypeWhile;
      np.conditional.size = 1;
      ok = cudaGraphAddNode(&node, graph, nullptr, 0, &np) == cudaSuccess;
    }
    int st = NF_OK;
    bool nf = false;
    const long long l0 = ctx->launches;
    if (ok) {
      cudaGraph_t body = np.conditional.phGraph_out[0];
      ok = cudaStreamBeginCaptureToGraph(mg->cap_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) == cudaSuccess;
      if (ok) {
        cudaStream_t orig = ctx->stream;
        ctx->stream = mg->cap_stream;
        st = mg_cycle(mg, 0, kind, true, &nf);
        if (st == NF_OK && !nf) {
          const nf_grid g = L.geom.grid(team->local[0]);
          st = nfi_residual_norms(ctx, &g, S.x, S.b, S.d_u, S.d_v, S.r, 1, mg->scal[0]);
        }
        if (st == NF_OK && L.geom.dist && !(nf && mg->norm_reduced)) st = nf_team_allreduce(team, mg->scal.data(), 2);
        if (st == NF_OK) {
          nf_launch(k_mg_loop_check, 1, 1, 0, mg->cap_stream, true, handle, mg->scal[0], mg->cfg.tolerance, mg->cfg.max_iterations);
          ctx->launches++;
        }
        ctx->stream = orig;
        cudaGraph_t out = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(mg->cap_stream, &out);
        ok = (st == NF_OK) && ce == cudaSuccess && S.x == key[0];
        if (S.x != key[0]) { S.x = (double*)key[0]; S.x2 = (double*)key[4]; }
      }
    }
    mg->loop_nodes = ctx->launches - l0;
    ctx->launches = l0;
    if (ok) ok = cudaGraphInstantiate(&mg->loop_exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      mg->loop_exec = nullptr;
      mg->use_loop = false;  // host loop from now on
      return st;
    }
    for (int q = 0; q < 5; ++q) mg->loop_key[q] = key[q];
    mg->loop_kind = kind;
    mg->loop_norm_fused = nf;
  }
  NF_CHECK_CUDA(ctx, cudaMemsetAsync(mg->scal[0] + 4, 0, sizeof(double), ctx->stream));
  NF_CHECK_CUDA(ctx, cudaGraphLaunch(mg->loop_exec, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(mg->scal_host, mg->scal[0], 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *cycles = (int)mg->scal_host[4];
  ctx->launches += 1 + mg->loop_nodes * (long long)*cycles;
  *rn = sqrt(mg->scal_host[0]);
  *bn = sqrt(mg->scal_host[1]);
  *norm_fused = mg->loop_norm_fused;
  *used = true;
  return NF_OK;
}

// recursive FMG (multigrid.py:562-688): RHS restricted down, exact coarsest solve, cubic prolongation
// (hard-coded :631), max_cycles_buildup cycles per level with early exit on ||r||/||b|| < tol.
static int mg_fmg(nf_mg* mg, int l) {
  nf_ctx* ctx = mg->ctx;
  nf_team* team = mg->team;
  MgLevel& L = mg->lv[l];
  if (L.geom.nx <= mg->cfg.coarsest || l + 1 == (int)mg->lv.size()) return mg_coarse_solve(mg, l);
  MgLevel& C = mg->lv[l + 1];
  for (int k = 0; k < nlocal(mg); ++k) {
    const int r = team->local[k];
    const nf_grid gf = L.geom.grid(r), gc = restrict_target(C, r);
    if (mg->cfg.restriction == 0) NF_TRY(nfi_restrict_fw(ctx, &gf, L.s[k].b, &gc, C.s[k].b));
    else NF_TRY(nfi_restrict_inject(ctx, &gf, L.s[k].b, &gc, C.s[k].b));
  }
  NF_TRY(mg_publish_rhs(mg, l));
  NF_TRY(mg_fmg(mg, l + 1));
  NF_TRY(mg_prolong(mg, l, 1, 0));
  for (int c = 0; c < mg->cfg.max_cycles_buildup; ++c) {
    NF_TRY(mg_cycle(mg, l, mg->cfg.cycle_buildup));
    if (mg->cfg.tolerance < 1.0 && c + 1 < mg->cfg.max_cycles_buildup) {
      double rn = 0.0, bn = -1.0;
      NF_TRY(mg_rel_residual(mg, l, &rn, &bn, 1));
      const double rel = bn > 0.0 ? rn / bn : rn;
      if (rel < mg->cfg.tolerance) break;
    }
  }
  return NF_OK;
}

// Level 0 iterates between the caller's x and the hierarchy's x2: bring the result home, restore x2.
static int mg_bind_level0(nf_mg* mg, double* const* b, double* const* x, double* const* r, std::vector<double*>& keep_x2,
                          std::vector<double*>& keep_r) {
  MgLevel& L = mg->lv[0];
  keep_x2.clear(); keep_r.clear();
  for (int k = 0; k < nlocal(mg); ++k) {
    keep_x2.push_back(L.s[k].x2);
    keep_r.push_back(L.s[k].r);
    L.s[k].x = x[k];
    L.s[k].b = const_cast<double*>(b[k]);
    if (r && r[k]) L.s[k].r = r[k];  // residual field goes straight to the caller's array
  }
  return NF_OK;
}

static int mg_unbind_level0(nf_mg* mg, double* const* x, const std::vector<double*>& keep_x2,
                            const std::vector<double*>& keep_r) {
  nf_ctx* ctx = mg->ctx;
  MgLevel& L = mg->lv[0];
  for (int k = 0; k < nlocal(mg); ++k) {
    MgSlab& S = L.s[k];
    if (S.x != x[k]) {  // the result sits in the hierarchy's buffer: copy it into the caller's array
      NF_CHECK_CUDA(ctx, cudaMemcpyAsync(x[k], S.x, L.geom.elems(mg->team->local[k]) * sizeof(double),
                                         cudaMemcpyDeviceToDevice, ctx->stream));
    }
    S.x2 = keep_x2[k];
    S.x = nullptr;
    S.b = nullptr;
    S.r = keep_r[k];
  }
  return NF_OK;
}

extern "C" int nf_mg_cycle(nf_mg* mg, double* x, const double* b, int kind) {
  if (!mg) return NF_ERR_ARG;
  NF_REQUIRE(mg->ctx, mg->setup_done, "nf_mg_setup has not been called");
  NF_REQUIRE(mg->ctx, kind == 0 || kind == 1, "kind must be 0 ('v') or 1 ('w')");
  NF_REQUIRE(mg->ctx, nlocal(mg) == 1, "nf_mg_cycle is the single-slab entry point");
  std::vector<double*> k2, kr;
  double* bb = const_cast<double*>(b);
  NF_TRY(mg_bind_level0(mg, &bb, &x, nullptr, k2, kr));
  int st = mg_cycle(mg, 0, kind);
  int st2 = mg_unbind_level0(mg, &x, k2, kr);
  return st ? st : st2;
}

// MultiGridSolver.solve without get_rhs, per local slab arrays.  sync == 0 (FMG mode only): the final ||r||^2,
// ||b||^2 stay in mg->scal[k][0..1] and no host synchronisation happens.
// want_field == 0: the caller does not read the residual FIELD (info['field'] of the reference): when the norms came
// out of the smoother the extra b - A x pass at the end is skipped
int nfi_mg_solve(nf_mg* mg, double* const* b, double* const* x, double* const* r, nf_mg_info* info, int sync, int want_field) {
  nf_ctx* ctx = mg->ctx;
  NF_REQUIRE(ctx, mg->setup_done, "nf_mg_setup has not been called");
  MgLevel& L = mg->lv[0];
  std::vector<double*> k2, kr;
  NF_TRY(mg_bind_level0(mg, b, x, r, k2, kr));
  int status = NF_OK;
  double rn = 0.0, bn = -1.0;  // bn < 0: ||b|| not computed yet
  int cycles = 0;
  do {
    for (int k = 0; k < nlocal(mg) && !status; ++k)
      status = nfi_fill(ctx, L.s[k].x, L.geom.elems(mg->team->local[k]), 0.0);  // x0 = 0 (multigrid.py:165)
    if (status) break;
    if (mg->cfg.cycle_type == 2) {
      status = mg_fmg(mg, 0);
      if (status) break;
      if (mg->cfg.cycle_final >= 0) {
        status = mg_cycle(mg, 0, mg->cfg.cycle_final);
        if (status) break;
        cycles = 1;
      }
      status = mg_rel_residual(mg, 0, &rn, &bn, sync);
    } else {
      bool fused_any = false;
      // "Lookahead norm": the pre-smoothing launch of cycle k+1 evaluates ||b - A x_k|| of its input while it sets up
      // its tiles (no deeper halo, unlike a norm fused behind the post-smoother), the host reads it and either lets the
      // rest of cycle k+1 run or -- converged after k cycles, multigrid.py:185-240 -- drops the launch's output (x_k is
      // still intact in the other buffer).  One wasted pre-smoothing launch per solve buys a plain post-smoother in
      // every cycle.  Needs the fused path of the level-0 pre-smoother (one launch: 1..3 sweeps, TMA variant).
      const char* envl = getenv("NF_MG_LOOKAHEAD");
      bool lookahead = !(envl && envl[0] == '0') && mg->cfg.smoother == 0 && mg->cfg.restriction == 0 &&
                       mg->cfg.pre >= 1 && mg->cfg.pre <= 3 && mg->lv.size() > 1 && L.geom.nx > mg->cfg.coarsest;
      {
        // the streaming smoother evaluates the residual norms of its OUTPUT at no extra halo cost (no speculative launch
        // needed): classic test behind the post-smoother on the levels it serves
        const nf_grid g0 = L.geom.grid(mg->team->local[0]);
        if (nfi_rbsor_stream_enabled(&g0)) lookahead = false;
      }
      bool have_final_norm = false;
      bool on_device = false;
      status = mg_device_loop(mg, &cycles, &rn, &bn, &fused_any, &on_device);
      if (status) break;
      if (on_device) have_final_norm = true;
      for (int it = 0; it < mg->cfg.max_iterations && !on_device; ++it) {
        if (lookahead) {
          std::vector<double*> x_before(nlocal(mg)), x2_before(nlocal(mg));
          for (int k = 0; k < nlocal(mg); ++k) { x_before[k] = L.s[k].x; x2_before[k] = L.s[k].x2; }
          bool inz = false;
          status = mg_cycle(mg, 0, mg->cfg.cycle_type, false, nullptr, 1, true, &inz);
          if (status) break;
          if (inz) {
            fused_any = true;
            double rin = 0.0, bin = bn;
            status = mg_rel_residual(mg, 0, &rin, &bin, 1, true);  // all-reduce (slabs) + 16-byte D2H + sync
            if (status) break;
            bn = bin;
            if (it > 0) {
              rn = rin;
              const double rel = bn > 0.0 ? rn / bn : rn;
              if (rel < mg->cfg.tolerance) {  // converged after `it` cycles: forget the speculative pre-smoothing
                for (int k = 0; k < nlocal(mg); ++k) { L.s[k].x = x_before[k]; L.s[k].x2 = x2_before[k]; }
                have_final_norm = true;
                break;
              }
            }
            bool nfz = false;
            status = mg_cycle_top(mg, mg->cfg.cycle_type, &nfz, 2, false);
            if (status) break;
            ++cycles;
            continue;
          }
          // the level is too small for the fused path: the head ran as a plain pre-smoothing; finish this cycle the
          // classic way and stay there
          lookahead = false;
          bool nfz = false;
          status = mg_cycle(mg, 0, mg->cfg.cycle_type, true, &nfz, 2);
          if (status) break;
          ++cycles;
          status = mg_rel_residual(mg, 0, &rn, &bn, 1, nfz);
          if (status) break;
          fused_any = fused_any || nfz;
          const double rel = bn > 0.0 ? rn / bn : rn;
          if (rel < mg->cfg.tolerance) { have_final_norm = true; break; }
          continue;
        }
        bool nfz = false;
        status = mg_cycle_top(mg, mg->cfg.cycle_type, &nfz);
        if (status) break;
        ++cycles;
        status = mg_rel_residual(mg, 0, &rn, &bn, 1, nfz);
        if (status) break;
        fused_any = fused_any || nfz;
        have_final_norm = true;
        const double rel = bn > 0.0 ? rn / bn : rn;
        if (rel < mg->cfg.tolerance) break;
      }
      if (!status && lookahead && !have_final_norm && cycles > 0)  // ran out of cycles: the norm after the last one
        status = mg_rel_residual(mg, 0, &rn, &bn, 1);
      if (!status && fused_any && want_field)  // the residual field itself (info['field']) once, at the end
        for (int k = 0; k < nlocal(mg) && !status; ++k) {
          const nf_grid g0 = L.geom.grid(mg->team->local[k]);
          status = nfi_residual(ctx, &g0, L.s[k].x, L.s[k].b, L.s[k].d_u, L.s[k].d_v, L.s[k].r);
        }
    }
  } while (0);
  int st2 = mg_unbind_level0(mg, x, k2, kr);
  if (mg->prof && !status && sync) {
    cudaStreamSynchronize(ctx->stream);
    mg_prof_harvest(mg);
    mg->psolves++;
    mg->pcycles += cycles;
  }
  if (status) return status;
  if (st2) return st2;
  if (info) {
    info->r_norm = rn;
    info->b_norm = bn < 0.0 ? 0.0 : bn;
    info->cycles = cycles;
    info->levels = (int)mg->lv.size();
  }
  return NF_OK;
}

double* nfi_mg_scalars(nf_mg* mg, int k) { return mg->scal[k]; }

// Multigrid as a preconditioner (matrix_free_BiCGSTAB.py:102-161): out = `cycles` cycles (kind 0 'v', 1 'w',
// 2 'fmg') applied to A out = rhs starting from out = 0.  Single slab.
int nfi_mg_apply(nf_mg* mg, const double* rhs, double* out, int cycles, int kind) {
  nf_ctx* ctx = mg->ctx;
  NF_REQUIRE(ctx, mg->setup_done, "nf_mg_setup has not been called");
  NF_REQUIRE(ctx, nlocal(mg) == 1, "the multigrid preconditioner is a single-slab entry point");
  MgLevel& L = mg->lv[0];
  std::vector<double*> k2, kr;
  double* bb = const_cast<double*>(rhs);
  NF_TRY(mg_bind_level0(mg, &bb, &out, nullptr, k2, kr));
  int st = nfi_fill(ctx, L.s[0].x, L.geom.elems(mg->team->local[0]), 0.0);
  for (int c = 0; c < cycles && st == NF_OK; ++c) {
    if (kind == 2) st = mg_fmg(mg, 0);
    else st = mg_cycle(mg, 0, kind);
  }
  int st2 = mg_unbind_level0(mg, &out, k2, kr);
  return st ? st : st2;
}

// collect the finished event pairs (call after a stream synchronisation)
static void mg_harvest_timing(nf_mg* mg) {
  for (size_t e = 0; e + 1 < mg->ev_used; e += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, mg->ev[e], mg->ev[e + 1]) == cudaSuccess) {
      mg->smooth_ms += ms;
      mg->smooth_launches += 1;
    }
  }
  mg->ev_used = 0;
}

// live timing of the finest-level k_rbsor_* <3 sweeps> launches: enable (on = 1) / disable, read and reset
extern "C" int nf_mg_smoother_timing(nf_mg* mg, int on, double* total_ms, long long* launches) {
  if (!mg) return NF_ERR_ARG;
  NF_CHECK_CUDA(mg->ctx, cudaStreamSynchronize(mg->ctx->stream));
  mg_harvest_timing(mg);
  if (total_ms) *total_ms = mg->smooth_ms;
  if (launches) *launches = mg->smooth_launches;
  mg->smooth_ms = 0.0;
  mg->smooth_launches = 0;
  mg->timing = on != 0;
  return NF_OK;
}

extern "C" int nf_mg_solve(nf_mg* mg, const double* b, double* x, double* r, nf_mg_info* info) {
  if (!mg) return NF_ERR_ARG;
  NF_REQUIRE(mg->ctx, nlocal(mg) == 1, "nf_mg_solve is the single-slab entry point");
  NF_REQUIRE(mg->ctx, b && x, "NULL argument");
  double* bb = const_cast<double*>(b);
  return nfi_mg_solve(mg, &bb, &x, &r, info, 1, 1);
}

// Device workspace of the stand-alone entry points that need scratch memory (the header's rule: no allocation inside hot
// calls; the multigrid / Krylov / SIMPLE objects own theirs).  which = NF_WS_PROLONG_CUBIC: nf_prolong_cubic from an
// nxc-cell coarse grid onto an nx-cell fine grid (band of the 1-D interpolation matrix, its start indices, the row-interpolated
// intermediate array).
static size_t cubic_ws_layout(int nc, int nf_, int ldc, size_t* off_start, size_t* off_tmp) {
  const int W = (nc < 48) ? nc : 48;
  size_t band = (size_t)nf_ * W * sizeof(double);
  band = (band + 255) / 256 * 256;
  size_t start = ((size_t)nf_ * sizeof(int) + 255) / 256 * 256;
  if (off_start) *off_start = band;
  if (off_tmp) *off_tmp = band + start;
  return band + start + (size_t)nf_ * ldc * sizeof(double);
}

extern "C" size_t nf_workspace_bytes(int which, int nx, int ny, int nxc, int ldc) {
  (void)ny;
  if (which == NF_WS_PROLONG_CUBIC) return cubic_ws_layout(nxc, nx, ldc, nullptr, nullptr);
  return 0;
}

// standalone interpolate_cubic (tests, Python-level callers): the band is built on the host and copied into the caller's
// workspace on every call; no device allocation, no stream synchronisation
extern "C" int nf_prolong_cubic(nf_ctx* ctx, const nf_grid* gc, const double* c, const nf_grid* gf, double* f,
                                int add, void* workspace, size_t workspace_bytes) {
  NF_REQUIRE(ctx, gc && gf && c && f, "NULL argument");
  NF_REQUIRE(ctx, gc->nx == gc->ny && gf->nx == gf->ny, "square grids only");
  NF_REQUIRE(ctx, gc->nx >= 2 && gf->nx >= 2, "grid too small");
  size_t off_start = 0, off_tmp = 0;
  const size_t need = cubic_ws_layout(gc->nx, gf->nx, gc->ld, &off_start, &off_tmp);
  NF_REQUIRE(ctx, workspace && workspace_bytes >= need, "workspace too small (nf_workspace_bytes(NF_WS_PROLONG_CUBIC, ...))");
  NF_REQUIRE(ctx, ((uintptr_t)workspace % 256) == 0, "workspace must be 256-byte aligned");
  std::vector<double> band;
  std::vector<int> start;
  int W = 0;
  build_interp_band(gc->nx, gf->nx, 24, band, start, W);
  // pageable host memory: cudaMemcpyAsync stages it before it returns, so the vectors may go out of scope afterwards
  char* ws = (char*)workspace;
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(ws, band.data(), band.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(ws + off_start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice,
                                     ctx->stream));
  return nfi_prolong_banded(ctx, gc, c, gf, f, (double*)(ws + off_tmp), gc->ld, (const double*)ws, (const int*)(ws + off_start),
                            W, add);
}
