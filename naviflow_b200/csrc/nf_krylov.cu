// nf_krylov.cu -- matrix-free CG / BiCGSTAB on the pressure-correction operator (K14, K15), fp64.
//
// The reference delegates the Krylov arithmetic to scipy.sparse.linalg (cg / bicgstab in
// scipy/sparse/linalg/_isolve/iterative.py, called from pressure_solver/matrix_free_BiCGSTAB.py:234-242
// with A = LinearOperator(compute_Ap_product), x0 = 0, atol = tolerance and scipy's default rtol = 1e-5).
// The recurrences below follow scipy's operation order statement by statement; only the summation order
// inside the dot products differs (deterministic block tree here, BLAS there).
//
// Every scalar (rho, alpha, beta, omega, the stopping test) lives on the device: the kernel that finishes
// a reduction also runs the scalar epilogue in its last block, and every kernel starts with
// `if (state->done) return`, so the host never synchronises inside an iteration and the result is
// identical for every polling interval (check_every).
#include <math.h>

#include "nf_pressure.cuh"
#include "nf_slab.cuh"

struct KState {
  double rho, rho_prev, alpha, omega, beta;
  double rr;        // ||r||^2 of the recurrence
  double atol;      // effective tolerance max(atol, rtol*||b||)
  double bnorm;
  int done, info, iters, half;
};

// A*f at cell (i,j) for a field given by an accessor (fused p-update + SpMV)
template <class F>
__device__ __forceinline__ double nf_Ap_cell_f(const nf_grid& g, const double* __restrict__ d_u,
                                               const double* __restrict__ d_v, int i, int j, F f) {
  const size_t k = nf_idx(g, i, j);
  const double pc = f(k);
  if (i == 0 && j == 0) return pc;
  const PCoef c = nf_pcoef(g, d_u, d_v, i, j);
  double out = c.diag * pc;
  if (i < g.nx - 1) out -= c.e * f(k + g.ld);
  if (i > 0) out -= c.w * f(k - g.ld);
  if (j < g.ny - 1) out -= c.n * f(k + 1);
  if (j > 0) out -= c.s * f(k - 1);
  return out;
}

// ---- scalar epilogues of the reductions (one thread).  Single slab: run by the last block of the reducing kernel.
// Slab-decomposed: the kernel leaves its partial sums in out[], the team all-reduces them and k_krylov_epilogue runs
// the same code (every rank holds an identical copy of the state). -------------------------------------------------
enum { EP_INIT = 0, EP_CG_PQ, EP_CG_UPDATE, EP_BI_V, EP_BI_S, EP_BI_T, EP_BI_X };

__device__ __forceinline__ void ep_init(KState* st, const double* out, double atol, double rtol) {
  const double bn = sqrt(out[0]);
  st->bnorm = bn;
  st->atol = fmax(atol, rtol * bn);
  st->rr = out[0];
  st->rho = out[0];
  st->rho_prev = 0.0;
  st->alpha = 0.0;
  st->omega = 0.0;
  st->beta = 0.0;
  st->half = 0;
  st->iters = 0;
  st->info = 0;
  // bnorm == 0: scipy returns x = b (= 0) immediately; ||r|| < atol at the top of iteration 0
  st->done = (bn == 0.0 || bn < st->atol) ? 1 : 0;
  if (!st->done && fabs(st->rho) < 4.930380657631324e-32) { st->done = 1; st->info = -10; }  // bicgstab rhotol = eps^2
}
__device__ __forceinline__ void ep_cg_pq(KState* st, const double* out) { st->alpha = st->rho / out[0]; }
__device__ __forceinline__ void ep_cg_update(KState* st, const double* out, int it) {
  const double rr = out[0];
  st->rr = rr;
  st->iters = it + 1;
  st->rho_prev = st->rho;
  st->rho = rr;
  st->beta = rr / st->rho_prev;
  if (sqrt(rr) < st->atol) { st->done = 1; st->info = 0; }
}
// preconditioned CG (scipy cg with M): the update only tests ||r||; rho = r.z and beta come from the reduction behind M
__device__ __forceinline__ void ep_pcg_update(KState* st, const double* out, int it) {
  st->rr = out[0];
  st->iters = it + 1;
  if (sqrt(out[0]) < st->atol) { st->done = 1; st->info = 0; }
}
__device__ __forceinline__ void ep_pcg_rz(KState* st, const double* out, int first) {
  st->rho_prev = st->rho;
  st->rho = out[0];
  st->beta = first ? 0.0 : st->rho / st->rho_prev;
}
__device__ __forceinline__ void ep_bi_v(KState* st, const double* out, int it) {
  const double rv = out[0];
  if (rv == 0.0) { st->done = 1; st->info = -11; st->iters = it; }
  else st->alpha = st->rho / rv;
}
__device__ __forceinline__ void ep_bi_s(KState* st, const double* out) {
  st->rr = out[0];
  st->half = (sqrt(out[0]) < st->atol) ? 1 : 0;
}
__device__ __forceinline__ void ep_bi_t(KState* st, const double* out) { st->omega = out[0] / out[1]; }
__device__ __forceinline__ void ep_bi_x(KState* st, const double* out, int it) {
  const double alpha = st->alpha, omega = st->omega;
  st->iters = it + 1;
  if (st->half) { st->done = 1; st->info = 0; return; }
  st->rr = out[0];
  st->rho_prev = st->rho;
  st->rho = out[1];
  const double eps2 = 4.930380657631324e-32;  // np.finfo(float64).eps ** 2
  if (sqrt(out[0]) < st->atol) { st->done = 1; st->info = 0; }
  else if (fabs(st->rho) < eps2) { st->done = 1; st->info = -10; }
  else if (fabs(omega) < eps2) { st->done = 1; st->info = -11; }
  else st->beta = (st->rho / st->rho_prev) * (alpha / omega);
}

__global__ void k_krylov_epilogue(int which, KState* st, const double* out, int it, double atol, double rtol) {
  if (which == EP_INIT) { ep_init(st, out, atol, rtol); return; }
  if (st->done) return;
  switch (which) {
    case EP_CG_PQ: ep_cg_pq(st, out); break;
    case EP_CG_UPDATE: ep_cg_update(st, out, it); break;
    case EP_BI_V: ep_bi_v(st, out, it); break;
    case EP_BI_S: ep_bi_s(st, out); break;
    case EP_BI_T: if (!st->half) ep_bi_t(st, out); break;
    case EP_BI_X: ep_bi_x(st, out, it); break;
  }
}

#define NF_ROWLOOP(g, i) \
  _Pragma("unroll 2") for (int i = (g).gb + blockIdx.y * blockDim.y + threadIdx.y; i < (g).ge; i += gridDim.y * blockDim.y)

// ---------------------------------------------------------------------------------------------
// shared: r = b, (rtilde = b), rr = rho = b.b, stopping tolerance
// ---------------------------------------------------------------------------------------------
__global__ void k_krylov_init(nf_grid g, const double* __restrict__ b, double* __restrict__ r,
                              double* __restrict__ rt, double* __restrict__ x, KState* st, double atol, double rtol,
                              int defer, double* partials, unsigned int* ticket, double* out) {
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    NF_ROWLOOP(g, i) {
      const size_t k = nf_idx(g, i, j);
      const double v = b[k];
      r[k] = v;
      if (rt) rt[k] = v;
      x[k] = 0.0;
      acc[0] += v * v;
    }
  }
  if (nf_block_reduce_store<1>(acc, partials, ticket, out) && !defer) ep_init(st, out, atol, rtol);
}

// ---------------------------------------------------------------------------------------------
// CG.  scipy order: z = r; rho = r.z; p = z + beta p; q = A p; alpha = rho/(p.q); x += alpha p; r -= alpha q
// ---------------------------------------------------------------------------------------------
// kernel 1: p_new = r + beta*p_old (first iteration: p_new = r), q = A p_new, pq = p_new.q ; epilogue alpha
// ext = 1 (slab-decomposed): p_new is also written on one row beyond the owned ones (r's halo is exchanged once per
// iteration, p's halo maintains itself this way); q and the dot product cover the owned rows only
__global__ void k_cg_pq(nf_grid g, const double* __restrict__ r, const double* __restrict__ p_old,
                        double* __restrict__ p_new, double* __restrict__ q, const double* __restrict__ d_u,
                        const double* __restrict__ d_v, KState* st, int first, int ext, int defer, double* partials,
                        unsigned int* ticket, double* out) {
  if (st->done) return;
  const double beta = st->beta;
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    const int lo = (ext && g.gb > 0) ? g.gb - 1 : g.gb, hi = (ext && g.ge < g.nx) ? g.ge + 1 : g.ge;
#pragma unroll 2
    for (int i = lo + blockIdx.y * blockDim.y + threadIdx.y; i < hi; i += gridDim.y * blockDim.y) {
      const size_t k = nf_idx(g, i, j);
      const bool owned = i >= g.gb && i < g.ge;
      double pc, qc = 0.0;
      if (first) {
        auto f = [&](size_t kk) { return r[kk]; };
        pc = f(k);
        if (owned) qc = nf_Ap_cell_f(g, d_u, d_v, i, j, f);
      } else {
        auto f = [&](size_t kk) { return p_old[kk] * beta + r[kk]; };  // p *= beta; p += z
        pc = f(k);
        if (owned) qc = nf_Ap_cell_f(g, d_u, d_v, i, j, f);
      }
      p_new[k] = pc;
      if (owned) {
        q[k] = qc;
        acc[0] += pc * qc;
      }
    }
  }
  if (nf_block_reduce_store<1>(acc, partials, ticket, out) && !defer) ep_cg_pq(st, out);
}

// kernel 2: x += alpha p; r -= alpha q; rr = r.r ; epilogue: stopping test, rho, beta
__global__ void k_cg_update(nf_grid g, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                            const double* __restrict__ q, KState* st, int it, int defer, double* partials,
                            unsigned int* ticket, double* out) {
  if (st->done) return;
  const double alpha = st->alpha;
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    NF_ROWLOOP(g, i) {
      const size_t k = nf_idx(g, i, j);
      x[k] = x[k] + alpha * p[k];
      const double rn = r[k] - alpha * q[k];
      r[k] = rn;
      acc[0] += rn * rn;
    }
  }
  if (nf_block_reduce_store<1>(acc, partials, ticket, out) && defer != 1) {  // defer -1: preconditioned recurrence
    if (defer == -1) ep_pcg_update(st, out, it); else ep_cg_update(st, out, it);
  }
}

// rho = r.z behind the preconditioner application (preconditioned CG); epilogue beta = rho / rho_prev
__global__ void k_pcg_rz(nf_grid g, const double* __restrict__ r, const double* __restrict__ z, KState* st, int first,
                         double* partials, unsigned int* ticket, double* out) {
  if (st->done) return;
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    NF_ROWLOOP(g, i) {
      const size_t k = nf_idx(g, i, j);
      acc[0] += r[k] * z[k];
    }
  }
  if (nf_block_reduce_store<1>(acc, partials, ticket, out)) ep_pcg_rz(st, out, first);
}

// ---------------------------------------------------------------------------------------------
// BiCGSTAB (scipy order, see header comment of nf_bicgstab_solve)
// ---------------------------------------------------------------------------------------------
// p = r + beta (p - omega v)   [p -= omega*v; p *= beta; p += r];  first iteration p = r
__global__ void k_bi_p(nf_grid g, const double* __restrict__ r, double* __restrict__ p, const double* __restrict__ v,
                       const KState* st, int first) {
  if (st->done) return;
  const double beta = st->beta, omega = st->omega;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.ny) return;
  NF_ROWLOOP(g, i) {
    const size_t k = nf_idx(g, i, j);
    p[k] = first ? r[k] : ((p[k] - omega * v[k]) * beta + r[k]);
  }
}

// v = A p; rv = rtilde.v ; epilogue alpha = rho/rv (rv == 0 -> breakdown -11)
__global__ void k_bi_v(nf_grid g, const double* __restrict__ p, double* __restrict__ v, const double* __restrict__ rt,
                       const double* __restrict__ d_u, const double* __restrict__ d_v, KState* st, int it, int defer,
                       double* partials, unsigned int* ticket, double* out) {
  if (st->done) return;
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    NF_ROWLOOP(g, i) {
      const size_t k = nf_idx(g, i, j);
      const double vc = nf_Ap_cell(g, p, d_u, d_v, i, j);
      v[k] = vc;
      acc[0] += rt[k] * vc;
    }
  }
  if (nf_block_reduce_store<1>(acc, partials, ticket, out) && !defer) ep_bi_v(st, out, it);
}

// s = r - alpha v (in place); ss = s.s ; epilogue: ||s|| < atol -> half-step exit
__global__ void k_bi_s(nf_grid g, double* __restrict__ r, const double* __restrict__ v, KState* st, int defer,
                       double* partials, unsigned int* ticket, double* out) {
  if (st->done) return;
  const double alpha = st->alpha;
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    NF_ROWLOOP(g, i) {
      const size_t k = nf_idx(g, i, j);
      const double s = r[k] - alpha * v[k];
      r[k] = s;
      acc[0] += s * s;
    }
  }
  if (nf_block_reduce_store<1>(acc, partials, ticket, out) && !defer) ep_bi_s(st, out);
}

// t = A s; ts = t.s, tt = t.t ; epilogue omega = ts/tt
__global__ void k_bi_t(nf_grid g, const double* shat, const double* s, double* __restrict__ t,
                       const double* __restrict__ d_u, const double* __restrict__ d_v, KState* st, int defer,
                       double* partials, unsigned int* ticket, double* out) {
  if (st->done || st->half) return;
  double acc[2] = {0.0, 0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    NF_ROWLOOP(g, i) {
      const size_t k = nf_idx(g, i, j);
      const double tc = nf_Ap_cell(g, shat, d_u, d_v, i, j);
      t[k] = tc;
      acc[0] += tc * s[k];
      acc[1] += tc * tc;
    }
  }
  if (nf_block_reduce_store<2>(acc, partials, ticket, out) && !defer) ep_bi_t(st, out);
}

// x += alpha p; x += omega s; r = s - omega t; rr = r.r; rho' = rtilde.r ; epilogue: top-of-loop tests of the
// next iteration (norm, rho breakdown, omega breakdown) and beta.  Half-step exit: x += alpha p only.
// phat / shat: the preconditioned directions (== p / s without a preconditioner)
__global__ void k_bi_x(nf_grid g, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                       const double* __restrict__ shat, const double* __restrict__ t, const double* __restrict__ rt,
                       KState* st, int it, int defer, double* partials, unsigned int* ticket, double* out) {
  if (st->done) return;
  const double alpha = st->alpha, omega = st->omega;
  const int half = st->half;
  double acc[2] = {0.0, 0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    NF_ROWLOOP(g, i) {
      const size_t k = nf_idx(g, i, j);
      if (half) {
        x[k] = x[k] + alpha * p[k];
      } else {
        const double s = r[k];
        const double sh = shat ? shat[k] : s;  // shat == NULL: no preconditioner (shat = s, which lives in r)
        x[k] = (x[k] + alpha * p[k]) + omega * sh;
        const double rn = s - omega * t[k];
        r[k] = rn;
        acc[0] += rn * rn;
        acc[1] += rt[k] * rn;
      }
    }
  }
  if (nf_block_reduce_store<2>(acc, partials, ticket, out) && !defer) ep_bi_x(st, out, it);
}

// =============================================================================================
// host drivers
// =============================================================================================
static int krylov_state(nf_ctx* ctx, KState** dev, KState** host) {
  // the state lives in the context's scalar block (slots 16..): enough room for KState
  static_assert(sizeof(KState) <= sizeof(double) * 16, "KState too large");
  *dev = reinterpret_cast<KState*>(ctx->scalars + 16);
  *host = reinterpret_cast<KState*>(ctx->scalars_host + 16);
  return NF_OK;
}

static int krylov_poll(nf_ctx* ctx, KState* dev, KState* host) {
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(host, dev, sizeof(KState), cudaMemcpyDeviceToHost, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NF_OK;
}

static void krylov_finish(const KState* h, int maxiter, nf_krylov_info* info) {
  if (!info) return;
  info->r_norm = sqrt(h->rr);
  info->b_norm = h->bnorm;
  if (h->done) {
    info->iterations = h->iters;
    info->info = h->info;
  } else {  // scipy: info = maxiter when the loop runs out
    info->iterations = maxiter;
    info->info = maxiter;
  }
}

extern "C" int nf_cg_solve(nf_ctx* ctx, const nf_grid* g, const double* b, double* x, const double* d_u,
                           const double* d_v, double atol, double rtol, int maxiter, int check_every, double* work,
                           nf_krylov_info* info) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, b && x && d_u && d_v && work, "NULL argument");
  NF_REQUIRE(ctx, maxiter >= 0, "maxiter < 0");
  if (check_every < 1) check_every = 1;
  const size_t n = (size_t)(g->nx + 1) * g->ld;
  double* r = work;
  double* pbuf[2] = {work + n, work + 2 * n};
  double* q = work + 3 * n;
  KState *st, *hst;
  krylov_state(ctx, &st, &hst);
  NfLaunch2D l = nf_launch_reduce(g->ge - g->gb, g->ny);
  k_krylov_init<<<l.grid, l.block, 0, ctx->stream>>>(*g, b, r, nullptr, x, st, atol, rtol, 0, ctx->partials,
                                                     ctx->ticket, ctx->scalars);
  NF_LAUNCH_CHECK(ctx);
  int cur = 0;
  int it = 0;
  for (; it < maxiter; ++it) {
    k_cg_pq<<<l.grid, l.block, 0, ctx->stream>>>(*g, r, pbuf[cur], pbuf[cur ^ 1], q, d_u, d_v, st, it == 0, 0, 0,
                                                 ctx->partials, ctx->ticket, ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    cur ^= 1;
    k_cg_update<<<l.grid, l.block, 0, ctx->stream>>>(*g, x, r, pbuf[cur], q, st, it, 0, ctx->partials, ctx->ticket,
                                                     ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    if ((it + 1) % check_every == 0) {
      NF_TRY(krylov_poll(ctx, st, hst));
      if (hst->done) break;
    }
  }
  NF_TRY(krylov_poll(ctx, st, hst));
  krylov_finish(hst, maxiter, info);
  return NF_OK;
}

int nfi_mg_apply(nf_mg* mg, const double* rhs, double* out, int cycles, int kind);

// CG with the multigrid preconditioner of GeoMultigridPrecondCGSolver (pressure_solver/geo_multigrid_cg.py:125-191):
// scipy's cg(A, b, x0 = 0, M, atol) statement by statement -- ||r|| < atol test, z = M r (mg_cycles cycles on A y = r from
// y = 0), rho = r.z, p = z + (rho / rho_prev) p, q = A p, alpha = rho / p.q, x += alpha p, r -= alpha q.  The reference
// multiplies with the assembled matrix (coeff_matrix.py), which equals the matrix-free operator to rounding (SURVEY 5a = 5b).
// The host polls the device-side stopping flag once per iteration (the preconditioner is a sequence of launches that must
// not be issued for nothing).  work: 5 same-shape arrays.
extern "C" int nf_cg_solve_mg(nf_ctx* ctx, const nf_grid* g, const double* b, double* x, const double* d_u,
                              const double* d_v, double atol, double rtol, int maxiter, double* work, nf_mg* mg,
                              int mg_cycles, int mg_kind, nf_krylov_info* info) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, b && x && d_u && d_v && work && mg, "NULL argument");
  NF_REQUIRE(ctx, maxiter >= 0 && mg_cycles >= 1 && mg_kind >= 0 && mg_kind <= 2, "bad iteration arguments");
  const size_t n = (size_t)(g->nx + 1) * g->ld;
  double* r = work;
  double* pbuf[2] = {work + n, work + 2 * n};
  double* q = work + 3 * n;
  double* z = work + 4 * n;
  KState *st, *hst;
  krylov_state(ctx, &st, &hst);
  NfLaunch2D l = nf_launch_reduce(g->ge - g->gb, g->ny);
  k_krylov_init<<<l.grid, l.block, 0, ctx->stream>>>(*g, b, r, nullptr, x, st, atol, rtol, 0, ctx->partials,
                                                     ctx->ticket, ctx->scalars);
  NF_LAUNCH_CHECK(ctx);
  NF_TRY(krylov_poll(ctx, st, hst));
  int cur = 0;
  for (int it = 0; it < maxiter && !hst->done; ++it) {
    NF_TRY(nfi_mg_apply(mg, r, z, mg_cycles, mg_kind));
    k_pcg_rz<<<l.grid, l.block, 0, ctx->stream>>>(*g, r, z, st, it == 0, ctx->partials, ctx->ticket, ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    // p = z + beta p (k_cg_pq takes z in the place of r), q = A p, alpha = rho / p.q
    k_cg_pq<<<l.grid, l.block, 0, ctx->stream>>>(*g, z, pbuf[cur], pbuf[cur ^ 1], q, d_u, d_v, st, it == 0, 0, 0,
                                                 ctx->partials, ctx->ticket, ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    cur ^= 1;
    k_cg_update<<<l.grid, l.block, 0, ctx->stream>>>(*g, x, r, pbuf[cur], q, st, it, -1, ctx->partials, ctx->ticket,
                                                     ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    NF_TRY(krylov_poll(ctx, st, hst));
  }
  krylov_finish(hst, maxiter, info);
  return NF_OK;
}


static int bicgstab_impl(nf_ctx* ctx, const nf_grid* g, const double* b, double* x, const double* d_u, const double* d_v,
                         double atol, double rtol, int maxiter, int check_every, double* work, nf_mg* mg, int mg_cycles,
                         int mg_kind, nf_krylov_info* info) {
  if (check_every < 1) check_every = 1;
  const size_t n = (size_t)(g->nx + 1) * g->ld;
  double *r = work, *rt = work + n, *p = work + 2 * n, *v = work + 3 * n, *t = work + 4 * n;
  double* phat = mg ? work + 5 * n : p;   // M p
  double* shat = mg ? work + 6 * n : nullptr;   // M s (s itself lives in r)
  KState *st, *hst;
  krylov_state(ctx, &st, &hst);
  NfLaunch2D l = nf_launch_reduce(g->ge - g->gb, g->ny);
  k_krylov_init<<<l.grid, l.block, 0, ctx->stream>>>(*g, b, r, rt, x, st, atol, rtol, 0, ctx->partials, ctx->ticket,
                                                     ctx->scalars);
  NF_LAUNCH_CHECK(ctx);
  for (int it = 0; it < maxiter; ++it) {
    k_bi_p<<<l.grid, l.block, 0, ctx->stream>>>(*g, r, p, v, st, it == 0);
    NF_LAUNCH_CHECK(ctx);
    if (mg) NF_TRY(nfi_mg_apply(mg, p, phat, mg_cycles, mg_kind));
    k_bi_v<<<l.grid, l.block, 0, ctx->stream>>>(*g, phat, v, rt, d_u, d_v, st, it, 0, ctx->partials, ctx->ticket,
                                                ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    k_bi_s<<<l.grid, l.block, 0, ctx->stream>>>(*g, r, v, st, 0, ctx->partials, ctx->ticket, ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    if (mg) NF_TRY(nfi_mg_apply(mg, r, shat, mg_cycles, mg_kind));
    k_bi_t<<<l.grid, l.block, 0, ctx->stream>>>(*g, shat ? shat : r, r, t, d_u, d_v, st, 0, ctx->partials, ctx->ticket,
                                                ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    k_bi_x<<<l.grid, l.block, 0, ctx->stream>>>(*g, x, r, phat, shat, t, rt, st, it, 0, ctx->partials, ctx->ticket,
                                                ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    if ((it + 1) % check_every == 0) {
      NF_TRY(krylov_poll(ctx, st, hst));
      if (hst->done) break;
    }
  }
  NF_TRY(krylov_poll(ctx, st, hst));
  krylov_finish(hst, maxiter, info);
  return NF_OK;
}

extern "C" int nf_bicgstab_solve(nf_ctx* ctx, const nf_grid* g, const double* b, double* x, const double* d_u,
                                 const double* d_v, double atol, double rtol, int maxiter, int check_every,
                                 double* work, nf_krylov_info* info) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, b && x && d_u && d_v && work, "NULL argument");
  NF_REQUIRE(ctx, maxiter >= 0, "maxiter < 0");
  return bicgstab_impl(ctx, g, b, x, d_u, d_v, atol, rtol, maxiter, check_every, work, nullptr, 0, 0, info);
}

// BiCGSTAB with the multigrid preconditioner of matrix_free_BiCGSTAB.py:102-161: M z = mg_cycles cycles
// (mg_kind 0 'v', 1 'w', 2 'fmg') on A y = z from y = 0.  `mg` must have been set up (nf_mg_setup) with the same
// d_u, d_v; work holds 7 same-shape arrays.
extern "C" int nf_bicgstab_solve_mg(nf_ctx* ctx, const nf_grid* g, const double* b, double* x, const double* d_u,
                                    const double* d_v, double atol, double rtol, int maxiter, int check_every,
                                    double* work, nf_mg* mg, int mg_cycles, int mg_kind, nf_krylov_info* info) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, b && x && d_u && d_v && work && mg, "NULL argument");
  NF_REQUIRE(ctx, maxiter >= 0 && mg_cycles >= 1 && mg_kind >= 0 && mg_kind <= 2, "bad iteration arguments");
  return bicgstab_impl(ctx, g, b, x, d_u, d_v, atol, rtol, maxiter, check_every, work, mg, mg_cycles, mg_kind, info);
}

// =============================================================================================
// a7  MatrixFreeMomentumSolver's Krylov solve (matrix_free_momentum.py:343-362, :434-441): scipy bicgstab on the relaxed
// momentum system with x0 = the current velocity.  Operator: interior rows 5-point with the stored links, boundary rows
// identity (:49-79).  The reference preconditions with an ILU of a matrix that (as coded) lacks the north/south links; the
// device runs the same recurrence unpreconditioned -- the answers agree to the stopping tolerance max(atol, 1e-5 ||b||),
// which is the parity this solver can be pinned to at all.
// =============================================================================================
__device__ __forceinline__ double nf_links_apply(const nf_grid& ga, const nf_links& L, const double* __restrict__ x, int i,
                                                 int j) {
  const size_t k = nf_idx(ga, i, j);
  if (i == 0 || i == ga.nx - 1 || j == 0 || j == ga.ny - 1) return x[k];
  double y = L.a_p[k] * x[k];
  y -= L.a_e[k] * x[k + ga.ld];
  y -= L.a_w[k] * x[k - ga.ld];
  y -= L.a_n[k] * x[k + 1];
  y -= L.a_s[k] * x[k - 1];
  return y;
}

// r = b - A x0, rtilde = r; out[0] = b.b, out[1] = r.r ; epilogue: tolerance, rho, top-of-loop tests of iteration 0
__global__ void k_links_init(nf_grid ga, nf_links L, const double* __restrict__ x, double* __restrict__ r,
                             double* __restrict__ rt, KState* st, double atol, double rtol, double* partials,
                             unsigned int* ticket, double* out) {
  double acc[2] = {0.0, 0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ga.ny) {
    NF_ROWLOOP(ga, i) {
      const size_t k = nf_idx(ga, i, j);
      const double b = L.src[k];
      const double rv = b - nf_links_apply(ga, L, x, i, j);
      r[k] = rv;
      rt[k] = rv;
      acc[0] += b * b;
      acc[1] += rv * rv;
    }
  }
  if (nf_block_reduce_store<2>(acc, partials, ticket, out)) {
    const double bb[1] = {out[0]};
    ep_init(st, bb, atol, rtol);
    st->rr = out[1];
    st->rho = out[1];
    st->done = (st->bnorm == 0.0 || sqrt(out[1]) < st->atol) ? 1 : 0;
    if (!st->done && fabs(st->rho) < 4.930380657631324e-32) { st->done = 1; st->info = -10; }
  }
}

__global__ void k_links_v(nf_grid ga, nf_links L, const double* __restrict__ p, double* __restrict__ v,
                          const double* __restrict__ rt, KState* st, int it, double* partials, unsigned int* ticket,
                          double* out) {
  if (st->done) return;
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ga.ny) {
    NF_ROWLOOP(ga, i) {
      const size_t k = nf_idx(ga, i, j);
      const double vc = nf_links_apply(ga, L, p, i, j);
      v[k] = vc;
      acc[0] += rt[k] * vc;
    }
  }
  if (nf_block_reduce_store<1>(acc, partials, ticket, out)) ep_bi_v(st, out, it);
}

__global__ void k_links_t(nf_grid ga, nf_links L, const double* __restrict__ s, double* __restrict__ t, KState* st,
                          double* partials, unsigned int* ticket, double* out) {
  if (st->done || st->half) return;
  double acc[2] = {0.0, 0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ga.ny) {
    NF_ROWLOOP(ga, i) {
      const size_t k = nf_idx(ga, i, j);
      const double tc = nf_links_apply(ga, L, s, i, j);
      t[k] = tc;
      acc[0] += tc * s[k];
      acc[1] += tc * tc;
    }
  }
  if (nf_block_reduce_store<2>(acc, partials, ticket, out)) ep_bi_t(st, out);
}

// x: in = x0, out = solution; work: 5 arrays of (nx+1)*ld doubles; L: RELAXED links (a_p, src relaxed)
int nfi_momentum_bicgstab(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, double* x, double atol, double rtol,
                          int maxiter, int check_every, double* work, nf_krylov_info* info) {
  if (check_every < 1) check_every = 1;
  nf_grid ga = *g;  // the component's array seen as a plain rows x cols grid
  ga.nx = g->nx + (is_u ? 1 : 0);
  ga.ny = g->ny + (is_u ? 0 : 1);
  ga.gb = 0;
  ga.ge = ga.nx;
  const size_t n = (size_t)(g->nx + 1) * g->ld;
  double *r = work, *rt = work + n, *p = work + 2 * n, *v = work + 3 * n, *t = work + 4 * n;
  KState *st, *hst;
  krylov_state(ctx, &st, &hst);
  NfLaunch2D l = nf_launch_reduce(ga.nx, ga.ny);
  k_links_init<<<l.grid, l.block, 0, ctx->stream>>>(ga, L, x, r, rt, st, atol, rtol, ctx->partials, ctx->ticket,
                                                    ctx->scalars);
  NF_LAUNCH_CHECK(ctx);
  for (int it = 0; it < maxiter; ++it) {
    k_bi_p<<<l.grid, l.block, 0, ctx->stream>>>(ga, r, p, v, st, it == 0);
    NF_LAUNCH_CHECK(ctx);
    k_links_v<<<l.grid, l.block, 0, ctx->stream>>>(ga, L, p, v, rt, st, it, ctx->partials, ctx->ticket, ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    k_bi_s<<<l.grid, l.block, 0, ctx->stream>>>(ga, r, v, st, 0, ctx->partials, ctx->ticket, ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    k_links_t<<<l.grid, l.block, 0, ctx->stream>>>(ga, L, r, t, st, ctx->partials, ctx->ticket, ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    k_bi_x<<<l.grid, l.block, 0, ctx->stream>>>(ga, x, r, p, nullptr, t, rt, st, it, 0, ctx->partials, ctx->ticket,
                                                ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    if ((it + 1) % check_every == 0) {
      NF_TRY(krylov_poll(ctx, st, hst));
      if (hst->done) break;
    }
  }
  NF_TRY(krylov_poll(ctx, st, hst));
  krylov_finish(hst, maxiter, info);
  return NF_OK;
}

extern "C" int nf_momentum_bicgstab(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, double* x, double atol,
                                    double rtol, int maxiter, int check_every, double* work, nf_krylov_info* info) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, L.a_e && L.a_w && L.a_n && L.a_s && L.a_p && L.src && x && work, "NULL argument");
  NF_REQUIRE(ctx, maxiter >= 0, "maxiter < 0");
  NF_REQUIRE(ctx, g->row0 == 0 && g->gb == 0 && g->ge == g->nx, "single-slab grids only");
  return nfi_momentum_bicgstab(ctx, g, is_u, L, x, atol, rtol, maxiter, check_every, work, info);
}

// =============================================================================================
// Slab-decomposed CG / BiCGSTAB (kind 0 / 1): the same kernels per slab; every reduction leaves its partial sums in
// the slab's scratch (state[k] + 32), the team all-reduces them (peer-memory kernel or NCCL: BASELINE north_star "dot
// products use allreduce") and k_krylov_epilogue advances the slab's copy of the state.  One halo exchange of one row
// per operator application.  The summation order of the dot products depends on the cut, so the iterates agree with
// the single-slab run to rounding, not bit for bit.
// state[k]: 64 device doubles per local slab; work[k]: 4 (CG) / 5 (BiCGSTAB) slab-sized arrays.
// =============================================================================================
int nfi_krylov_team(nf_team* team, const LevelGeom& geom, int kind, double* const* b, double* const* x, double* const* d_u,
                    double* const* d_v, double atol, double rtol, int maxiter, int check_every, double* const* work,
                    double* const* state, nf_krylov_info* info) {
  nf_ctx* ctx = team->ctx;
  const int nl = (int)team->local.size();
  if (check_every < 1) check_every = 1;
  std::vector<nf_grid> g(nl);
  std::vector<NfLaunch2D> l(nl);
  std::vector<KState*> st(nl);
  std::vector<double*> out(nl), r(nl), a1(nl), a2(nl), a3(nl), a4(nl);
  for (int k = 0; k < nl; ++k) {
    const int rk = team->local[k];
    g[k] = geom.grid(rk);
    g[k].rho = 1.0;  // the pressure solvers hard-code rho = 1 (matrix_free_BiCGSTAB.py)
    l[k] = nf_launch_reduce(g[k].ge - g[k].gb + 2, g[k].ny);
    st[k] = reinterpret_cast<KState*>(state[k]);
    out[k] = state[k] + 32;
    const size_t n = geom.elems(rk);
    r[k] = work[k]; a1[k] = work[k] + n; a2[k] = work[k] + 2 * n; a3[k] = work[k] + 3 * n;
    a4[k] = kind == 1 ? work[k] + 4 * n : nullptr;
  }
  auto reduce_epilogue = [&](int which, int count, int it) -> int {
    NF_TRY(nf_team_allreduce(team, out.data(), (size_t)count));
    for (int k = 0; k < nl; ++k) {
      k_krylov_epilogue<<<1, 1, 0, ctx->stream>>>(which, st[k], out[k], it, atol, rtol);
      NF_LAUNCH_CHECK(ctx);
    }
    return NF_OK;
  };
  auto poll = [&](KState* host) -> int {
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(host, st[0], sizeof(KState), cudaMemcpyDeviceToHost, ctx->stream));
    NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NF_OK;
  };
  KState* hst = reinterpret_cast<KState*>(ctx->scalars_host + 16);
  for (int k = 0; k < nl; ++k) {
    k_krylov_init<<<l[k].grid, l[k].block, 0, ctx->stream>>>(g[k], b[k], r[k], kind == 1 ? a1[k] : nullptr, x[k], st[k],
                                                             atol, rtol, 1, ctx->partials, ctx->ticket, out[k]);
    NF_LAUNCH_CHECK(ctx);
  }
  NF_TRY(reduce_epilogue(EP_INIT, 1, 0));
  if (kind == 0) {
    std::vector<double*> pb[2] = {a1, a2};
    std::vector<double*>& q = a3;
    NF_TRY(nf_team_exchange(team, geom, r.data(), 1));
    int cur = 0;
    for (int it = 0; it < maxiter; ++it) {
      for (int k = 0; k < nl; ++k) {
        k_cg_pq<<<l[k].grid, l[k].block, 0, ctx->stream>>>(g[k], r[k], pb[cur][k], pb[cur ^ 1][k], q[k], d_u[k], d_v[k],
                                                           st[k], it == 0, 1, 1, ctx->partials, ctx->ticket, out[k]);
        NF_LAUNCH_CHECK(ctx);
      }
      NF_TRY(reduce_epilogue(EP_CG_PQ, 1, it));
      cur ^= 1;
      for (int k = 0; k < nl; ++k) {
        k_cg_update<<<l[k].grid, l[k].block, 0, ctx->stream>>>(g[k], x[k], r[k], pb[cur][k], q[k], st[k], it, 1,
                                                               ctx->partials, ctx->ticket, out[k]);
        NF_LAUNCH_CHECK(ctx);
      }
      NF_TRY(reduce_epilogue(EP_CG_UPDATE, 1, it));
      NF_TRY(nf_team_exchange(team, geom, r.data(), 1));
      if ((it + 1) % check_every == 0) {
        NF_TRY(poll(hst));
        if (hst->done) break;
      }
    }
  } else {
    std::vector<double*>&rt = a1, &p = a2, &v = a3, &t = a4;
    for (int it = 0; it < maxiter; ++it) {
      for (int k = 0; k < nl; ++k) {
        k_bi_p<<<l[k].grid, l[k].block, 0, ctx->stream>>>(g[k], r[k], p[k], v[k], st[k], it == 0);
        NF_LAUNCH_CHECK(ctx);
      }
      NF_TRY(nf_team_exchange(team, geom, p.data(), 1));
      for (int k = 0; k < nl; ++k) {
        k_bi_v<<<l[k].grid, l[k].block, 0, ctx->stream>>>(g[k], p[k], v[k], rt[k], d_u[k], d_v[k], st[k], it, 1,
                                                          ctx->partials, ctx->ticket, out[k]);
        NF_LAUNCH_CHECK(ctx);
      }
      NF_TRY(reduce_epilogue(EP_BI_V, 1, it));
      for (int k = 0; k < nl; ++k) {
        k_bi_s<<<l[k].grid, l[k].block, 0, ctx->stream>>>(g[k], r[k], v[k], st[k], 1, ctx->partials, ctx->ticket, out[k]);
        NF_LAUNCH_CHECK(ctx);
      }
      NF_TRY(reduce_epilogue(EP_BI_S, 1, it));
      NF_TRY(nf_team_exchange(team, geom, r.data(), 1));
      for (int k = 0; k < nl; ++k) {
        k_bi_t<<<l[k].grid, l[k].block, 0, ctx->stream>>>(g[k], r[k], r[k], t[k], d_u[k], d_v[k], st[k], 1, ctx->partials,
                                                          ctx->ticket, out[k]);
        NF_LAUNCH_CHECK(ctx);
      }
      NF_TRY(reduce_epilogue(EP_BI_T, 2, it));
      for (int k = 0; k < nl; ++k) {
        k_bi_x<<<l[k].grid, l[k].block, 0, ctx->stream>>>(g[k], x[k], r[k], p[k], nullptr, t[k], rt[k], st[k], it, 1,
                                                          ctx->partials, ctx->ticket, out[k]);
        NF_LAUNCH_CHECK(ctx);
      }
      NF_TRY(reduce_epilogue(EP_BI_X, 2, it));
      if ((it + 1) % check_every == 0) {
        NF_TRY(poll(hst));
        if (hst->done) break;
      }
    }
  }
  NF_TRY(poll(hst));
  krylov_finish(hst, maxiter, info);
  return NF_OK;
}
