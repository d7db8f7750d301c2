// nf_mg_tail.cu -- the coarse end of a V-cycle in ONE kernel.
//
// Reference: pressure_solver/multigrid.py:304-432 (_v_cycle) on the levels of <= 31 cells per side.  On those levels every
// kernel of the launch-by-launch cycle (smoother, residual + restriction, zero fill, coarse solve, prolongation) is pure launch
// latency: five launches per level.  Here one CTA (32 x 32 threads: one thread per cell-pair slot of the largest level) keeps
// the sub-hierarchy in shared memory -- x with a zero border, b, the five link coefficients and 1/aP of every level, 75 KB
// for 31 -> 15 -> 7 -- and walks down and up with block barriers between the colour passes.  The arithmetic is the
// launch-by-launch one, expression by expression (nf_pcoef, nf_Ap_cell's order, nf_prolong_linear_value, k_coarse_apply's
// summation order), so the result is bit-identical.  (A first version that also took the 63^2 level and evaluated nf_pcoef
// inside every update ran 86 us on its single SM -- slower than the launches it replaced.)
#include "nf_pressure.cuh"
#include "nf_mg_tail.cuh"

namespace {

constexpr int TAIL_THREADS = 1024;  // 32 x 32: thread (tid / 32, tid % 32) <-> cell row / cell-pair column

// One level in shared memory.  x carries a zero border (pitch ny+2, origin at [1][1]) so that the neighbour loads need no
// conditions: the link towards a cell outside the grid is zero, exactly the 0.0 the stand-alone kernels add there.
struct TailSm {
  double *x, *b, *e, *w, *n, *s, *diag, *inv;
  int nx, ny, ldx, ld;
};

__device__ __forceinline__ void tail_smooth(const TailSm& L, double omega, int n_sweeps) {
  const int tid = threadIdx.x, i = tid >> 5, jj = tid & 31;
  // the pinned cell is held at 0 (gauss_seidel.py:145, :305), also when no sweep follows; its neighbours are black cells,
  // which read it after the first red pass
  if (tid == 0) L.x[L.ldx + 1] = 0.0;
  if (n_sweeps == 0) {
    __syncthreads();
    return;
  }
  for (int sw = 0; sw < n_sweeps; ++sw) {
    for (int color = 0; color < 2; ++color) {
      const int j = 2 * jj + ((i + color) & 1);
      if (i < L.nx && j < L.ny && !(i == 0 && j == 0)) {
        const int k = i * L.ld + j, kx = (i + 1) * L.ldx + (j + 1);
        double acc = L.b[k];  // k_rbsor_color, term by term
        acc += L.e[k] * L.x[kx + L.ldx];
        acc += L.w[k] * L.x[kx - L.ldx];
        acc += L.n[k] * L.x[kx + 1];
        acc += L.s[k] * L.x[kx - 1];
        const double pn = acc * L.inv[k];
        const double pc = L.x[kx];
        L.x[kx] = pc + omega * (pn - pc);
      }
      __syncthreads();
    }
  }
}

// b - A x at (i,j): nf_Ap_cell's expression (diag*p - E - W - N - S, identity row at the pinned cell)
__device__ __forceinline__ double tail_res(const TailSm& L, int i, int j) {
  const int k = i * L.ld + j, kx = (i + 1) * L.ldx + (j + 1);
  const double pc = L.x[kx];
  if (i == 0 && j == 0) return L.b[k] - pc;
  double o = L.diag[k] * pc;
  o -= L.e[k] * L.x[kx + L.ldx];
  o -= L.w[k] * L.x[kx - L.ldx];
  o -= L.n[k] * L.x[kx + 1];
  o -= L.s[k] * L.x[kx - 1];
  return L.b[k] - o;
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) k_mg_tail(nf_tail_args a) {
  nf_pdl_entry();
  extern __shared__ __align__(16) double sm[];
  __shared__ TailSm lv[NF_TAIL_MAX_LEVELS];
  const int tid = threadIdx.x;
  if (tid == 0) {
    double* q = sm;
    for (int l = 0; l < a.nlev; ++l) {
      const nf_tail_level& T = a.lv[l];
      TailSm& L = lv[l];
      L.nx = T.nx; L.ny = T.ny; L.ld = T.ny; L.ldx = T.ny + 2;
      const int sz = T.nx * T.ny;
      L.x = q; q += (T.nx + 2) * (T.ny + 2);
      L.b = q; q += sz;
      L.e = q; q += sz;
      L.w = q; q += sz;
      L.n = q; q += sz;
      L.s = q; q += sz;
      L.diag = q; q += sz;
      L.inv = q; q += sz;
    }
  }
  __syncthreads();
  // ---- load: link coefficients (nf_pcoef on the level's d_u, d_v) and 1/aP of every level, right-hand side of the first ----
  for (int l = 0; l < a.nlev; ++l) {
    const nf_tail_level& T = a.lv[l];
    const TailSm& L = lv[l];
    nf_grid g;
    g.nx = T.nx; g.ny = T.ny; g.ld = T.ld; g.row0 = 0; g.gb = 0; g.ge = T.nx; g.row1 = 0; g.pad = 0;
    g.dx = T.dx; g.dy = T.dy; g.rho = a.rho;
    for (int t = tid; t < (T.nx + 2) * (T.ny + 2); t += TAIL_THREADS) L.x[t] = 0.0;
    for (int t = tid; t < T.nx * T.ny; t += TAIL_THREADS) {
      const int i = t / T.ny, j = t - i * T.ny;
      const PCoef c = nf_pcoef(g, T.d_u, T.d_v, i, j);
      L.e[t] = c.e; L.w[t] = c.w; L.n[t] = c.n; L.s[t] = c.s; L.diag[t] = c.diag;
      L.inv[t] = T.inv[(size_t)i * T.ld + j];
      L.b[t] = (l == 0) ? T.b[(size_t)i * T.ld + j] : 0.0;
    }
  }
  __syncthreads();
  // ---- down: pre-smoothing from a zero guess, residual, full weighting (multigrid.py:352-372) ----
  for (int l = 0; l + 1 < a.nlev; ++l) {
    const TailSm& L = lv[l];
    const TailSm& Cl = lv[l + 1];
    tail_smooth(L, a.omega, a.pre);
    for (int t = tid; t < Cl.nx * Cl.ny; t += TAIL_THREADS) {
      const int I = t / Cl.ny, J = t - I * Cl.ny;
      const int i = 2 * I + 1, j = 2 * J + 1;
      const double cc = tail_res(L, i, j), n = tail_res(L, i, j + 1), s = tail_res(L, i, j - 1);
      const double e = tail_res(L, i + 1, j), w = tail_res(L, i - 1, j);
      const double ne = tail_res(L, i + 1, j + 1), nw = tail_res(L, i - 1, j + 1);
      const double se = tail_res(L, i + 1, j - 1), sw = tail_res(L, i - 1, j - 1);
      Cl.b[t] = (cc / 4.0 + (((n + s) + e) + w) / 8.0) + (((ne + nw) + se) + sw) / 16.0;
    }
    __syncthreads();
  }
  // ---- coarsest level: x = Inv b, one warp per row, k_coarse_apply's summation order ----
  {
    const TailSm& L = lv[a.nlev - 1];
    const int N = a.N, lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < N; r += TAIL_THREADS / 32) {
      double acc = 0.0;
      for (int c = lane; c < N; c += 32) acc += a.coarse_inv[(size_t)r * N + c] * L.b[c];
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
      if (lane == 0) L.x[(r / L.ny + 1) * L.ldx + (r % L.ny) + 1] = acc;
    }
    __syncthreads();
  }
  // ---- up: x += P x_coarse (interpolate_linear, multigrid_helpers.py:116-186), post-smoothing ----
  for (int l = a.nlev - 2; l >= 0; --l) {
    const TailSm& L = lv[l];
    const TailSm& Cl = lv[l + 1];
    nf_grid gc;  // the coarse solution as nf_prolong_linear_value addresses it: origin at the border-free cell [0][0]
    gc.nx = Cl.nx; gc.ny = Cl.ny; gc.ld = Cl.ldx; gc.row0 = 0; gc.gb = 0; gc.ge = Cl.nx; gc.row1 = 0; gc.pad = 0;
    gc.dx = gc.dy = 0.0; gc.rho = a.rho;
    const double* cx = Cl.x + Cl.ldx + 1;
    for (int t = tid; t < L.nx * L.ny; t += TAIL_THREADS) {
      const int i = t / L.ny, j = t - i * L.ny;
      const int kx = (i + 1) * L.ldx + (j + 1);
      L.x[kx] = L.x[kx] + nf_prolong_linear_value(gc, cx, L.nx, L.ny, i, j);
    }
    __syncthreads();
    tail_smooth(L, a.omega, a.post);
  }
  // ---- the first level's solution goes back to the hierarchy ----
  {
    const nf_tail_level& T = a.lv[0];
    const TailSm& L = lv[0];
    for (int t = tid; t < T.nx * T.ny; t += TAIL_THREADS) {
      const int i = t / T.ny, j = t - i * T.ny;
      T.x[(size_t)i * T.ld + j] = L.x[(i + 1) * L.ldx + (j + 1)];
    }
  }
}

}  // namespace

size_t nfi_mg_tail_smem(const nf_tail_args* a) {
  size_t d = 0;
  for (int l = 0; l < a->nlev; ++l)
    d += (size_t)(a->lv[l].nx + 2) * (a->lv[l].ny + 2) + 7 * (size_t)a->lv[l].nx * a->lv[l].ny;
  return d * sizeof(double);
}

int nfi_mg_tail(nf_ctx* ctx, const nf_tail_args* a) {
  const size_t smem = nfi_mg_tail_smem(a);
  static size_t attr = 0;
  if (smem > attr) {
    NF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_mg_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  nf_launch(k_mg_tail, 1, TAIL_THREADS, smem, ctx->stream, true, *a);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}
