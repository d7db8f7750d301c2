// nf_simple.cu -- device-resident SIMPLE outer loop (fp64).
//
// Reference: solver/Algorithms/simple.py:78-269 (paths relative to /root/reference/naviflow_oo) with
//   momentum predictor     solver/momentum_solver/jacobi_matrix_solver.py:153-375 (fixed Jacobi sweeps)
//   pressure correction    solver/pressure_solver/{multigrid,jacobi,gauss_seidel,matrix_free_BiCGSTAB}.py
//   p update + Neumann     simple.py:148-150, base_algorithm.py:161-197
//   velocity correction    solver/velocity_solver/standard.py:10-69
// All fields stay in HBM between outer iterations; the host sees one record of norms per iteration.
#include <math.h>

#include <vector>

#include "nf_common.cuh"

// internal entry points of the other translation units
int nfi_apply_velocity_bc(nf_ctx*, const nf_grid*, const nf_bc_program*, double* u, double* v);
int nfi_momentum_links(nf_ctx*, const nf_grid*, int is_u, const double* u, const double* v, const double* p,
                       double mu, double alpha, int sides, nf_links out, double* d);
int nfi_momentum_jacobi_pp(nf_ctx*, const nf_grid*, int is_u, nf_links L, const double* x0, double* a, double* b,
                           int n_sweeps, double** result);
int nfi_momentum_residual_dev(nf_ctx*, const nf_grid*, int is_u, nf_links L, const double* x, double* field, int slot);
int nfi_correct_velocity(nf_ctx*, const nf_grid*, const nf_bc_program*, const double* us, const double* vs,
                         const double* pp, const double* d_u, const double* d_v, double* u, double* v);
int nfi_mg_solve(nf_mg* mg, const double* b, double* x, double* r, nf_mg_info* info, int sync);

enum { F_U = 0, F_V, F_P, F_USTAR, F_VSTAR, F_DU, F_DV, F_PPRIME, F_B, F_PRES, F_URES, F_VRES, F_COUNT };

struct nf_simple {
  nf_ctx* ctx = nullptr;
  nf_simple_config cfg;
  nf_grid g;
  size_t elems = 0;
  std::vector<double*> owned;
  double *u = nullptr, *v = nullptr, *p = nullptr, *p_alt = nullptr;
  double *ua = nullptr, *ub = nullptr, *va = nullptr, *vb = nullptr;  // momentum ping-pong buffers
  double *u_star = nullptr, *v_star = nullptr;                        // point into the buffers above
  double *ubc = nullptr, *vbc = nullptr;                              // BC'd copies when the state is not clean
  double *d_u = nullptr, *d_v = nullptr, *pp = nullptr, *b = nullptr, *pres = nullptr;
  double *ures = nullptr, *vres = nullptr;
  double* tmp = nullptr;  // Jacobi pressure ping-pong / Krylov work base
  double* kwork = nullptr;
  nf_links links;
  nf_mg* mg = nullptr;
  double* hist = nullptr;       // device: 8 doubles per iteration
  double* hist_host = nullptr;  // pinned
  int hist_cap = 0;
  bool bc_clean = false;
};

static inline int pad_ld(int ny) { return ((ny + 1 + 15) / 16) * 16; }

extern "C" int nf_simple_destroy(nf_simple* s) {
  if (!s) return NF_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  if (s->mg) nf_mg_destroy(s->mg);
  for (double* ptr : s->owned) cudaFree(ptr);
  if (s->hist) cudaFree(s->hist);
  if (s->hist_host) cudaFreeHost(s->hist_host);
  delete s;
  return NF_OK;
}

static double* alloc_field(nf_simple* s) {
  double* ptr = nullptr;
  if (cudaMalloc(&ptr, s->elems * sizeof(double)) != cudaSuccess) return nullptr;
  cudaMemsetAsync(ptr, 0, s->elems * sizeof(double), s->ctx->stream);
  s->owned.push_back(ptr);
  return ptr;
}

extern "C" int nf_simple_create(nf_ctx* ctx, nf_simple** out, const nf_simple_config* cfg) {
  NF_REQUIRE(ctx, out && cfg, "NULL argument");
  *out = nullptr;
  NF_REQUIRE(ctx, cfg->nx >= 3 && cfg->ny >= 3, "nx, ny must be >= 3");
  NF_REQUIRE(ctx, cfg->pressure_solver >= 0 && cfg->pressure_solver <= 4, "unknown pressure solver");
  NF_REQUIRE(ctx, cfg->alpha_u > 0.0, "alpha_u must be > 0");
  NF_REQUIRE(ctx, cfg->n_momentum_sweeps >= 0, "n_momentum_sweeps < 0");
  nf_simple* s = new nf_simple();
  s->ctx = ctx;
  s->cfg = *cfg;
  nf_grid& g = s->g;
  g.nx = cfg->nx; g.ny = cfg->ny; g.ld = pad_ld(cfg->ny); g.row0 = 0; g.gb = 0; g.ge = cfg->nx;
  g.dx = cfg->length / (cfg->nx - 1);  // structured.py:27-28
  g.dy = cfg->height / (cfg->ny - 1);
  g.rho = cfg->rho;
  s->elems = (size_t)(g.nx + 1) * g.ld;
  double** fields[] = {&s->u, &s->v, &s->p, &s->p_alt, &s->ua, &s->ub, &s->va, &s->vb, &s->ubc, &s->vbc,
                       &s->d_u, &s->d_v, &s->pp, &s->b, &s->pres, &s->ures, &s->vres, &s->tmp,
                       &s->links.a_e, &s->links.a_w, &s->links.a_n, &s->links.a_s, &s->links.a_p, &s->links.src};
  bool ok = true;
  for (double** f : fields) {
    *f = alloc_field(s);
    if (!*f) { ok = false; break; }
  }
  if (ok && cfg->pressure_solver >= 3) {
    const int nwork = cfg->pressure_solver == 3 ? 4 : 5;
    if (cudaMalloc(&s->kwork, s->elems * nwork * sizeof(double)) == cudaSuccess) {
      s->owned.push_back(s->kwork);
      cudaMemsetAsync(s->kwork, 0, s->elems * nwork * sizeof(double), ctx->stream);
    } else ok = false;
  }
  if (!ok) {
    ctx->err = std::string("SIMPLE state allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    nf_simple_destroy(s);
    return NF_ERR_ALLOC;
  }
  s->u_star = s->ua;
  s->v_star = s->va;
  if (cfg->pressure_solver == 0) {
    nf_mg_config mc = cfg->mg;
    mc.length = cfg->length; mc.height = cfg->height; mc.rho = 1.0;  // callers hard-code rho = 1 (multigrid.py:151)
    int st = nf_mg_create(ctx, &s->mg, g.nx, g.ny, g.ld, &mc);
    if (st != NF_OK) { nf_simple_destroy(s); return st; }
  }
  // initial fields: zeros with the velocity BCs applied (base_algorithm.py:68-93)
  int st = nfi_apply_velocity_bc(ctx, &g, &s->cfg.bc, s->u, s->v);
  if (st != NF_OK) { nf_simple_destroy(s); return st; }
  s->bc_clean = true;
  *out = s;
  return NF_OK;
}

extern "C" int nf_simple_ld(nf_simple* s) { return s ? s->g.ld : 0; }

static double* field_ptr(nf_simple* s, int which) {
  switch (which) {
    case F_U: return s->u;
    case F_V: return s->v;
    case F_P: return s->p;
    case F_USTAR: return s->u_star;
    case F_VSTAR: return s->v_star;
    case F_DU: return s->d_u;
    case F_DV: return s->d_v;
    case F_PPRIME: return s->pp;
    case F_B: return s->b;
    case F_PRES: return s->pres;
    case F_URES: return s->ures;
    case F_VRES: return s->vres;
  }
  return nullptr;
}

extern "C" double* nf_simple_field(nf_simple* s, int which) { return s ? field_ptr(s, which) : nullptr; }

extern "C" int nf_simple_upload(nf_simple* s, int which, const double* host, int rows, int cols) {
  if (!s) return NF_ERR_ARG;
  nf_ctx* ctx = s->ctx;
  double* dst = field_ptr(s, which);
  NF_REQUIRE(ctx, dst && host, "bad field / NULL host pointer");
  NF_REQUIRE(ctx, rows >= 1 && rows <= s->g.nx + 1 && cols >= 1 && cols <= s->g.ld, "shape does not fit the field");
  NF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(dst, (size_t)s->g.ld * sizeof(double), host, (size_t)cols * sizeof(double),
                                       (size_t)cols * sizeof(double), rows, cudaMemcpyHostToDevice, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (which == F_U || which == F_V) s->bc_clean = false;
  return NF_OK;
}

extern "C" int nf_simple_download(nf_simple* s, int which, double* host, int rows, int cols) {
  if (!s) return NF_ERR_ARG;
  nf_ctx* ctx = s->ctx;
  const double* src = field_ptr(s, which);
  NF_REQUIRE(ctx, src && host, "bad field / NULL host pointer");
  NF_REQUIRE(ctx, rows >= 1 && rows <= s->g.nx + 1 && cols >= 1 && cols <= s->g.ld, "shape does not fit the field");
  NF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(host, (size_t)cols * sizeof(double), src, (size_t)s->g.ld * sizeof(double),
                                       (size_t)cols * sizeof(double), rows, cudaMemcpyDeviceToHost, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NF_OK;
}

// history record: [0] sum r_u^2, [1] sum b_u^2, [2] sum r_v^2, [3] sum b_v^2, [4] p norm payload a, [5] payload b,
//                 [6] pressure iterations, [7] spare
__global__ void k_store_hist(const double* __restrict__ scalars, double* __restrict__ rec, double pa, double pb,
                             double iters, int p_from_scalars) {
  const int t = threadIdx.x;
  if (t < 4) rec[t] = scalars[2 + t];
  if (t == 4) rec[4] = p_from_scalars ? scalars[0] : pa;
  if (t == 5) rec[5] = p_from_scalars ? scalars[1] : pb;
  if (t == 6) rec[6] = iters;
  if (t == 7) rec[7] = 0.0;
}

static int ensure_hist(nf_simple* s, int n) {
  nf_ctx* ctx = s->ctx;
  if (n <= s->hist_cap) return NF_OK;
  if (s->hist) cudaFree(s->hist);
  if (s->hist_host) cudaFreeHost(s->hist_host);
  s->hist = nullptr; s->hist_host = nullptr; s->hist_cap = 0;
  NF_CHECK_CUDA(ctx, cudaMalloc(&s->hist, (size_t)n * 8 * sizeof(double)));
  NF_CHECK_CUDA(ctx, cudaMallocHost(&s->hist_host, (size_t)n * 8 * sizeof(double)));
  s->hist_cap = n;
  return NF_OK;
}

static void decode_record(const nf_simple* s, const double* rec, nf_simple_info* out) {
  out->u_abs_res = sqrt(rec[0]);
  out->v_abs_res = sqrt(rec[2]);
  out->u_rel_norm = sqrt(rec[0]) / (sqrt(rec[1]) + 1e-15);  // jacobi_matrix_solver.py:246-250
  out->v_rel_norm = sqrt(rec[2]) / (sqrt(rec[3]) + 1e-15);
  switch (s->cfg.pressure_solver) {
    case 0: out->p_rel_norm = sqrt(rec[4]); break;                               // absolute ||r|| (multigrid.py:257)
    case 1: case 2: out->p_rel_norm = sqrt(rec[4]); break;                       // absolute ||b - A p'||
    default: out->p_rel_norm = rec[5] > 0.0 ? sqrt(rec[4]) / sqrt(rec[5]) : sqrt(rec[4]); break;  // ||r_int||/||b_int||
  }
  out->pressure_iterations = (int)rec[6];
  out->pad = 0;
}

// one outer iteration; leaves its history record in hist[slot]
static int simple_step(nf_simple* s, int slot, int want_fields) {
  nf_ctx* ctx = s->ctx;
  const nf_grid* g = &s->g;
  const nf_simple_config& c = s->cfg;
  // velocities with BCs applied, used for the coefficients (jacobi_matrix_solver.py:170)
  const double *ubc = s->u, *vbc = s->v;
  if (!s->bc_clean) {
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(s->ubc, s->u, s->elems * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(s->vbc, s->v, s->elems * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    NF_TRY(nfi_apply_velocity_bc(ctx, g, &c.bc, s->ubc, s->vbc));
    ubc = s->ubc; vbc = s->vbc;
  }
  // u-momentum: links, n sweeps from x0 = u, residual norm
  NF_TRY(nfi_momentum_links(ctx, g, 1, ubc, vbc, s->p, c.mu, c.alpha_u, c.sides, s->links, s->d_u));
  NF_TRY(nfi_momentum_jacobi_pp(ctx, g, 1, s->links, s->u, s->ua, s->ub, c.n_momentum_sweeps, &s->u_star));
  NF_TRY(nfi_momentum_residual_dev(ctx, g, 1, s->links, s->u_star, want_fields ? s->ures : nullptr, 2));
  // v-momentum (same u, v, p*: simple.py:128-133)
  NF_TRY(nfi_momentum_links(ctx, g, 0, ubc, vbc, s->p, c.mu, c.alpha_u, c.sides, s->links, s->d_v));
  NF_TRY(nfi_momentum_jacobi_pp(ctx, g, 0, s->links, s->v, s->va, s->vb, c.n_momentum_sweeps, &s->v_star));
  NF_TRY(nfi_momentum_residual_dev(ctx, g, 0, s->links, s->v_star, want_fields ? s->vres : nullptr, 4));
  // pressure correction
  nf_grid gp = *g;
  gp.rho = 1.0;  // every pressure solver of the reference hard-codes rho = 1.0 (multigrid.py:151, jacobi.py, ...)
  NF_TRY(nf_continuity_rhs(ctx, &gp, s->u_star, s->v_star, s->b));
  double pa = 0.0, pb = 0.0, iters = 0.0;
  int p_from_scalars = 0;
  switch (c.pressure_solver) {
    case 0: {
      NF_TRY(nf_mg_setup(s->mg, s->d_u, s->d_v));
      nf_mg_info mi;
      const int fmg = (c.mg.cycle_type == 2);
      NF_TRY(nfi_mg_solve(s->mg, s->b, s->pp, s->pres, &mi, fmg ? 0 : 1));
      if (fmg) p_from_scalars = 1;
      else { pa = mi.r_norm * mi.r_norm; pb = mi.b_norm * mi.b_norm; }
      iters = mi.cycles;
      break;
    }
    case 1:
    case 2: {
      NF_TRY(nfi_fill(ctx, s->pp, (size_t)g->nx * g->ld, 0.0));
      if (c.pressure_solver == 1)
        NF_TRY(nfi_jacobi(ctx, &gp, s->pp, s->tmp, s->b, s->d_u, s->d_v, c.pressure_omega, c.pressure_iterations));
      else
        NF_TRY(nfi_rbsor(ctx, &gp, s->pp, s->b, s->d_u, s->d_v, c.pressure_omega, c.pressure_iterations));
      NF_TRY(nfi_residual(ctx, &gp, s->pp, s->b, s->d_u, s->d_v, s->pres));
      NF_TRY(nfi_sumsq_dev(ctx, &gp, s->pres, 0, 0));
      NF_TRY(nfi_sumsq_dev(ctx, &gp, s->b, 0, 1));
      p_from_scalars = 1;
      iters = c.pressure_iterations;
      break;
    }
    default: {
      nf_krylov_info ki;
      if (c.pressure_solver == 3)
        NF_TRY(nf_cg_solve(ctx, &gp, s->b, s->pp, s->d_u, s->d_v, c.pressure_tolerance, 1e-5, c.krylov_maxiter, 25,
                           s->kwork, &ki));
      else
        NF_TRY(nf_bicgstab_solve(ctx, &gp, s->b, s->pp, s->d_u, s->d_v, c.pressure_tolerance, 1e-5, c.krylov_maxiter,
                                 10, s->kwork, &ki));
      // rel_norm = ||r_int|| / ||b_int|| of the true residual (matrix_free_BiCGSTAB.py:255-279)
      NF_TRY(nfi_residual(ctx, &gp, s->pp, s->b, s->d_u, s->d_v, s->pres));
      NF_TRY(nfi_sumsq_dev(ctx, &gp, s->pres, 1, 0));
      NF_TRY(nfi_sumsq_dev(ctx, &gp, s->b, 1, 1));
      p_from_scalars = 1;
      iters = ki.iterations;
      break;
    }
  }
  k_store_hist<<<1, 32, 0, ctx->stream>>>(ctx->scalars, s->hist + (size_t)slot * 8, pa, pb, iters, p_from_scalars);
  NF_LAUNCH_CHECK(ctx);
  // p = p* + alpha_p p' with zero-gradient edges; p* <- p
  NF_TRY(nf_update_pressure(ctx, g, s->p, s->pp, c.alpha_p, s->p_alt));
  { double* t = s->p; s->p = s->p_alt; s->p_alt = t; }
  // velocity correction + BCs
  NF_TRY(nfi_correct_velocity(ctx, g, &c.bc, s->u_star, s->v_star, s->pp, s->d_u, s->d_v, s->u, s->v));
  s->bc_clean = true;
  return NF_OK;
}

extern "C" int nf_simple_iterate(nf_simple* s, int n_iterations, double tolerance, int want_fields,
                                 nf_simple_info* info_host, int* n_done) {
  if (!s) return NF_ERR_ARG;
  nf_ctx* ctx = s->ctx;
  NF_REQUIRE(ctx, n_iterations >= 0, "n_iterations < 0");
  NF_TRY(ensure_hist(s, n_iterations > 0 ? n_iterations : 1));
  int done = 0;
  for (int it = 0; it < n_iterations; ++it) {
    NF_TRY(simple_step(s, it, want_fields));
    ++done;
    if (tolerance > 0.0) {  // stopping test of simple.py:114 needs this iteration's norms
      NF_CHECK_CUDA(ctx, cudaMemcpyAsync(s->hist_host + (size_t)it * 8, s->hist + (size_t)it * 8, 8 * sizeof(double),
                                         cudaMemcpyDeviceToHost, ctx->stream));
      NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      nf_simple_info rec;
      decode_record(s, s->hist_host + (size_t)it * 8, &rec);
      if (info_host) info_host[it] = rec;
      const double total = fmax(rec.u_rel_norm, rec.v_rel_norm);
      if (!(total > tolerance)) break;
    }
  }
  if (!(tolerance > 0.0) && done > 0) {
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(s->hist_host, s->hist, (size_t)done * 8 * sizeof(double), cudaMemcpyDeviceToHost,
                                       ctx->stream));
    NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (info_host)
      for (int it = 0; it < done; ++it) decode_record(s, s->hist_host + (size_t)it * 8, &info_host[it]);
  }
  if (n_done) *n_done = done;
  return NF_OK;
}
