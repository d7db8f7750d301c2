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

#define NF_ROWLOOP(g, i) \
  _Pragma("unroll 2") for (int i = (g).gb + blockIdx.y * blockDim.y + threadIdx.y; i < (g).ge; i += gridDim.y * blockDim.y)

// ---------------------------------------------------------------------------------------------
// shared: r = b, (rtilde = b), rr = rho = b.b, stopping tolerance
// ---------------------------------------------------------------------------------------------
__global__ void k_krylov_init(nf_grid g, const double* __restrict__ b, double* __restrict__ r,
                              double* __restrict__ rt, double* __restrict__ x, KState* st, double atol, double rtol,
                              double* partials, unsigned int* ticket, double* out) {
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
  if (nf_block_reduce_store<1>(acc, partials, ticket, out)) {
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
}

// ---------------------------------------------------------------------------------------------
// CG.  scipy order: z = r; rho = r.z; p = z + beta p; q = A p; alpha = rho/(p.q); x += alpha p; r -= alpha q
// ---------------------------------------------------------------------------------------------
// kernel 1: p_new = r + beta*p_old (first iteration: p_new = r), q = A p_new, pq = p_new.q ; epilogue alpha
__global__ void k_cg_pq(nf_grid g, const double* __restrict__ r, const double* __restrict__ p_old,
                        double* __restrict__ p_new, double* __restrict__ q, const double* __restrict__ d_u,
                        const double* __restrict__ d_v, KState* st, int first, double* partials,
                        unsigned int* ticket, double* out) {
  if (st->done) return;
  const double beta = st->beta;
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny) {
    NF_ROWLOOP(g, i) {
      const size_t k = nf_idx(g, i, j);
      double pc, qc;
      if (first) {
        auto f = [&](size_t kk) { return r[kk]; };
        pc = f(k);
        qc = nf_Ap_cell_f(g, d_u, d_v, i, j, f);
      } else {
        auto f = [&](size_t kk) { return p_old[kk] * beta + r[kk]; };  // p *= beta; p += z
        pc = f(k);
        qc = nf_Ap_cell_f(g, d_u, d_v, i, j, f);
      }
      p_new[k] = pc;
      q[k] = qc;
      acc[0] += pc * qc;
    }
  }
  if (nf_block_reduce_store<1>(acc, partials, ticket, out)) st->alpha = st->rho / out[0];
}

// kernel 2: x += alpha p; r -= alpha q; rr = r.r ; epilogue: stopping test, rho, beta
__global__ void k_cg_update(nf_grid g, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                            const double* __restrict__ q, KState* st, int it, double* partials, unsigned int* ticket,
                            double* out) {
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
  if (nf_block_reduce_store<1>(acc, partials, ticket, out)) {
    const double rr = out[0];
    st->rr = rr;
    st->iters = it + 1;
    st->rho_prev = st->rho;
    st->rho = rr;
    st->beta = rr / st->rho_prev;
    if (sqrt(rr) < st->atol) { st->done = 1; st->info = 0; }
  }
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
                       const double* __restrict__ d_u, const double* __restrict__ d_v, KState* st, int it,
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
  if (nf_block_reduce_store<1>(acc, partials, ticket, out)) {
    const double rv = out[0];
    if (rv == 0.0) { st->done = 1; st->info = -11; st->iters = it; }
    else st->alpha = st->rho / rv;
  }
}

// s = r - alpha v (in place); ss = s.s ; epilogue: ||s|| < atol -> half-step exit
__global__ void k_bi_s(nf_grid g, double* __restrict__ r, const double* __restrict__ v, KState* st, double* partials,
                       unsigned int* ticket, double* out) {
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
  if (nf_block_reduce_store<1>(acc, partials, ticket, out)) {
    st->rr = out[0];
    st->half = (sqrt(out[0]) < st->atol) ? 1 : 0;
  }
}

// t = A s; ts = t.s, tt = t.t ; epilogue omega = ts/tt
__global__ void k_bi_t(nf_grid g, const double* shat, const double* s, double* __restrict__ t,
                       const double* __restrict__ d_u, const double* __restrict__ d_v, KState* st, double* partials,
                       unsigned int* ticket, double* out) {
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
  if (nf_block_reduce_store<2>(acc, partials, ticket, out)) st->omega = out[0] / out[1];
}

// x += alpha p; x += omega s; r = s - omega t; rr = r.r; rho' = rtilde.r ; epilogue: top-of-loop tests of the
// next iteration (norm, rho breakdown, omega breakdown) and beta.  Half-step exit: x += alpha p only.
// phat / shat: the preconditioned directions (== p / s without a preconditioner)
__global__ void k_bi_x(nf_grid g, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                       const double* __restrict__ shat, const double* __restrict__ t, const double* __restrict__ rt,
                       KState* st, int it, double* partials, unsigned int* ticket, double* out) {
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
  if (nf_block_reduce_store<2>(acc, partials, ticket, out)) {
    st->iters = it + 1;
    if (half) { st->done = 1; st->info = 0; return; }
    st->rr = out[0];
    st->rho_prev = st->rho;
    st->rho = out[1];
    const double eps2 = 4.930380657631324e-32;  // np.finfo(float64).eps ** 2
    if (sqrt(out[0]) < st->atol) { st->done = 1; st->info = 0; }
    else if (fabs(st->rho) < eps2) { st->done = 1; st->info = -10; }
    else if (fabs(omega) < eps2) { st->done = 1; st->info = -11; }
    else st->beta = (st->rho / st->rho_prev) * (alpha / omega);
  }
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
  k_krylov_init<<<l.grid, l.block, 0, ctx->stream>>>(*g, b, r, nullptr, x, st, atol, rtol, ctx->partials, ctx->ticket,
                                                     ctx->scalars);
  NF_LAUNCH_CHECK(ctx);
  int cur = 0;
  int it = 0;
  for (; it < maxiter; ++it) {
    k_cg_pq<<<l.grid, l.block, 0, ctx->stream>>>(*g, r, pbuf[cur], pbuf[cur ^ 1], q, d_u, d_v, st, it == 0,
                                                 ctx->partials, ctx->ticket, ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    cur ^= 1;
    k_cg_update<<<l.grid, l.block, 0, ctx->stream>>>(*g, x, r, pbuf[cur], q, st, it, ctx->partials, ctx->ticket,
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
  k_krylov_init<<<l.grid, l.block, 0, ctx->stream>>>(*g, b, r, rt, x, st, atol, rtol, ctx->partials, ctx->ticket,
                                                     ctx->scalars);
  NF_LAUNCH_CHECK(ctx);
  for (int it = 0; it < maxiter; ++it) {
    k_bi_p<<<l.grid, l.block, 0, ctx->stream>>>(*g, r, p, v, st, it == 0);
    NF_LAUNCH_CHECK(ctx);
    if (mg) NF_TRY(nfi_mg_apply(mg, p, phat, mg_cycles, mg_kind));
    k_bi_v<<<l.grid, l.block, 0, ctx->stream>>>(*g, phat, v, rt, d_u, d_v, st, it, ctx->partials, ctx->ticket,
                                                ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    k_bi_s<<<l.grid, l.block, 0, ctx->stream>>>(*g, r, v, st, ctx->partials, ctx->ticket, ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    if (mg) NF_TRY(nfi_mg_apply(mg, r, shat, mg_cycles, mg_kind));
    k_bi_t<<<l.grid, l.block, 0, ctx->stream>>>(*g, shat ? shat : r, r, t, d_u, d_v, st, ctx->partials, ctx->ticket,
                                                ctx->scalars);
    NF_LAUNCH_CHECK(ctx);
    k_bi_x<<<l.grid, l.block, 0, ctx->stream>>>(*g, x, r, phat, shat, t, rt, st, it, ctx->partials, ctx->ticket,
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
