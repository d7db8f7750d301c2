// nf_simple.cu -- device-resident SIMPLE outer loop (fp64), single GPU or row slabs.
//
// Reference: solver/Algorithms/simple.py:78-269 (paths relative to /root/reference/naviflow_oo) with
//   momentum predictor     solver/momentum_solver/jacobi_matrix_solver.py:153-375 (fixed Jacobi sweeps)
//   pressure correction    solver/pressure_solver/{multigrid,jacobi,gauss_seidel,matrix_free_BiCGSTAB}.py
//   p update + Neumann     simple.py:148-150, base_algorithm.py:161-197
//   velocity correction    solver/velocity_solver/standard.py:10-69
// All fields stay in HBM between outer iterations; the host sees one record of norms per iteration.
//
// Slab runs (nf_slab.cuh): every rank keeps NF_HALO = 8 halo rows.  The momentum phase needs NO communication:
// the link coefficients are evaluated NF_HALO-1 rows beyond the owned rows and each Jacobi sweep is computed on a
// region that shrinks by one row, so after up to 6 sweeps the owned rows (+1) are exact.  Per outer iteration the
// ranks exchange: the halo of b (once), the multigrid halos (nf_mg.cu), p, u, v (once each) and 6 scalars.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "nf_slab.cuh"

// internal entry points of the other translation units
int nfi_apply_velocity_bc(nf_ctx*, const nf_grid*, const nf_bc_program*, double* u, double* v);
int nfi_momentum_links(nf_ctx*, const nf_grid*, int is_u, const double* u, const double* v, const double* p,
                       double mu, double alpha, int sides, nf_links out, double* d);
int nfi_momentum_sweep(nf_ctx*, const nf_grid*, int is_u, nf_links L, const double* src, double* dst);
int nfi_momentum_residual_to(nf_ctx*, const nf_grid*, int is_u, nf_links L, const double* x, double* field, double* out);
int nfi_momentum_sweeps_fused(nf_ctx*, const nf_grid*, int is_u, nf_links L, const double* xin, double* xout, int k,
                              int with_res, int norm_b, int norm_e, double* field, double* out);
int nfi_momentum_links_mf(nf_ctx*, const nf_grid*, int is_u, const double* u, const double* v, const double* p, double mu,
                          double alpha, int sides, nf_links out, double* d, double* ap_un, double* src_un);
int nfi_momentum_bicgstab(nf_ctx*, const nf_grid*, int is_u, nf_links L, double* x, double atol, double rtol, int maxiter,
                          int check_every, double* work, nf_krylov_info* info);
int nfi_momentum_residual_unrelaxed(nf_ctx*, const nf_grid*, int is_u, nf_links L, const double* x, double* field,
                                    double* out);
int nfi_momentum_unrelaxed_from_relaxed(nf_ctx*, const nf_grid*, int is_u, nf_links L, const double* x,
                                        const double* phi_old, double alpha, double* out);
#define NF_HIST 10  // doubles per history record
int nfi_correct_velocity(nf_ctx*, const nf_grid*, const nf_bc_program*, const double* us, const double* vs,
                         const double* pp, const double* d_u, const double* d_v, double* u, double* v);
int nfi_gs_lex(nf_ctx*, const nf_grid*, double* p, const double* b, const double* d_u, const double* d_v, double omega,
               int n_sweeps, int symmetric);
int nfi_krylov_team(nf_team* team, const LevelGeom& geom, int kind, double* const* b, double* const* x, double* const* d_u,
                    double* const* d_v, double atol, double rtol, int maxiter, int check_every, double* const* work,
                    double* const* state, nf_krylov_info* info);
int nfi_mg_create(nf_team* team, nf_mg** out, int nx, int ny, int ld, const nf_mg_config* cfg);
int nfi_mg_setup(nf_mg* mg, double* const* d_u, double* const* d_v);
int nfi_mg_solve(nf_mg* mg, double* const* b, double* const* x, double* const* r, nf_mg_info* info, int sync, int want_field);
double* nfi_mg_scalars(nf_mg* mg, int k);
LevelGeom nf_level0_geom(const nf_team* team, int nx, int ny, int ld, double length, double height, double rho);

enum { F_U = 0, F_V, F_P, F_USTAR, F_VSTAR, F_DU, F_DV, F_PPRIME, F_B, F_PRES, F_URES, F_VRES, F_COUNT };

struct SimpleSlab {
  std::vector<double*> owned;
  double *u = nullptr, *v = nullptr, *p = nullptr, *p_alt = nullptr;
  double *ua = nullptr, *ub = nullptr, *va = nullptr, *vb = nullptr;  // momentum ping-pong buffers
  double *u_star = nullptr, *v_star = nullptr;                        // point into the buffers above
  double *ubc = nullptr, *vbc = nullptr;                              // BC'd copies when the state is not clean
  double *d_u = nullptr, *d_v = nullptr, *pp = nullptr, *b = nullptr, *pres = nullptr;
  double *ures = nullptr, *vres = nullptr;
  double* tmp = nullptr;    // Jacobi pressure ping-pong
  double* kwork = nullptr;  // Krylov work arrays
  double *ap_un = nullptr, *src_un = nullptr, *mwork = nullptr;  // Krylov momentum predictor (a7)
  double* p_old = nullptr;  // SIMPLER / SIMPLEC: p at the start of the iteration (p_rel_norm)
  double *u_old = nullptr, *v_old = nullptr;  // SIMPLEC: velocities at the start of the iteration (total residual)
  double* cscal = nullptr;  // SIMPLEC: 8 device doubles, infinity norms [0] |u*-u| [1] |v*-v| [2] |p-p_old| [3] |u-u_old| [4] |v-v_old|
  double* kstate = nullptr; // slab-decomposed Krylov: this slab's copy of the scalar state + reduction scratch (64 doubles)
  double* scal = nullptr;   // 8 device doubles: [0..1] pressure norms, [2..5] momentum sums
  nf_links links;
};

struct nf_simple {
  nf_ctx* ctx = nullptr;
  nf_team* team = nullptr;
  bool owns_team = false;
  nf_simple_config cfg;
  LevelGeom geom;
  std::vector<SimpleSlab> s;
  nf_mg* mg = nullptr;
  double* hist = nullptr;       // device: 8 doubles per iteration (slab 0's reduced scalars)
  double* hist_host = nullptr;  // pinned
  int hist_cap = 0;
  bool bc_clean = false;
  int want_fields = 1;  // this call of nf_simple_iterate returns residual fields (else the pressure solver skips its field pass)
  // optional phase timing (nf_simple_phase_timing): 4 events per outer iteration on the context's stream
  bool phase_timing = false;
  std::vector<cudaEvent_t> pev;
  size_t pev_used = 0;
  double phase_ms[3] = {0.0, 0.0, 0.0};  // momentum predictor(s), pressure solve(s) incl. RHS, corrections
  long long phase_iters = 0;
};

static void phase_mark(nf_simple* s) {
  if (!s->phase_timing) return;
  if (s->pev_used == s->pev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    s->pev.push_back(e);
  }
  cudaEventRecord(s->pev[s->pev_used++], s->ctx->stream);
}

static int nlocal(const nf_simple* s) { return (int)s->team->local.size(); }

extern "C" int nf_simple_destroy(nf_simple* s) {
  if (!s) return NF_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  if (s->mg) nf_mg_destroy(s->mg);
  for (SimpleSlab& S : s->s)
    for (double* ptr : S.owned) nf_team_release(s->team, ptr);
  if (s->hist) cudaFree(s->hist);
  if (s->hist_host) cudaFreeHost(s->hist_host);
  for (cudaEvent_t e : s->pev) cudaEventDestroy(e);
  if (s->owns_team) nf_team_destroy(s->team);
  delete s;
  return NF_OK;
}

static double* alloc_elems(nf_simple* s, SimpleSlab& S, size_t elems, size_t elems_max) {
  double* ptr = nf_team_alloc(s->team, elems, elems_max);
  if (ptr) S.owned.push_back(ptr);
  return ptr;
}

template <class M>
static std::vector<double*> field_of(nf_simple* s, M member) {
  std::vector<double*> v;
  for (SimpleSlab& S : s->s) v.push_back(S.*member);
  return v;
}

int nfi_simple_create(nf_team* team, nf_simple** out, const nf_simple_config* cfg) {
  nf_ctx* ctx = team->ctx;
  NF_REQUIRE(ctx, out && cfg, "NULL argument");
  *out = nullptr;
  NF_REQUIRE(ctx, cfg->nx >= 3 && cfg->ny >= 3, "nx, ny must be >= 3");
  NF_REQUIRE(ctx, cfg->pressure_solver >= 0 && cfg->pressure_solver <= 7, "unknown pressure solver");
  NF_REQUIRE(ctx, cfg->piso_corrections >= -2, "unknown algorithm (piso_corrections < -2)");
  NF_REQUIRE(ctx, cfg->piso_corrections != -2 || cfg->simplec_divisor != 0.0, "SIMPLEC needs simplec_divisor");
  NF_REQUIRE(ctx, cfg->pressure_solver != 7 || (cfg->krylov_mg_cycles >= 1 && cfg->krylov_mg_kind >= 0 &&
                                                cfg->krylov_mg_kind <= 2), "bad multigrid preconditioner arguments");
  NF_REQUIRE(ctx, cfg->alpha_u > 0.0, "alpha_u must be > 0");
  NF_REQUIRE(ctx, cfg->n_momentum_sweeps >= 0, "n_momentum_sweeps < 0");
  NF_REQUIRE(ctx, cfg->momentum_solver == 0 || cfg->momentum_solver == 1, "unknown momentum solver");
  NF_REQUIRE(ctx, cfg->momentum_solver == 0 || cfg->momentum_maxiter >= 0, "momentum_maxiter < 0");
  nf_simple* s = new nf_simple();
  s->ctx = ctx;
  s->team = team;
  s->cfg = *cfg;
  s->geom = nf_level0_geom(team, cfg->nx, cfg->ny, nf_pad_ld(cfg->ny), cfg->length, cfg->height, cfg->rho);
  if (s->geom.dist && cfg->pressure_solver >= 5) {
    ctx->err = "the sequential Gauss-Seidel sweeps and the multigrid-preconditioned BiCGSTAB run on a single slab only";
    delete s;
    return NF_ERR_UNSUPPORTED;
  }
  if ((s->geom.dist || team->local.size() != 1) && cfg->piso_corrections == -2) {
    ctx->err = "the SIMPLEC loop runs on a single slab only";
    delete s;
    return NF_ERR_UNSUPPORTED;
  }
  if (s->geom.dist && cfg->momentum_solver != 0) {
    ctx->err = "slab-decomposed runs support the Jacobi-sweep momentum predictor only";
    delete s;
    return NF_ERR_UNSUPPORTED;
  }
  const int nl = (int)team->local.size();
  s->s.resize(nl);
  if (s->geom.dist) {  // peer-memory halo staging for the deepest exchange of the finest level (no-op without p2p)
    int st = nf_p2p_reserve_stage(team, (size_t)NF_HALO * s->geom.ld);
    if (st != NF_OK) { delete s; return st; }
  }
  const size_t emax = s->geom.max_elems();
  bool ok = true;
  for (int k = 0; k < nl && ok; ++k) {
    SimpleSlab& S = s->s[k];
    const size_t e = s->geom.elems(team->local[k]);
    double** fields[] = {&S.u, &S.v, &S.p, &S.p_alt, &S.ua, &S.ub, &S.va, &S.vb, &S.ubc, &S.vbc,
                         &S.d_u, &S.d_v, &S.pp, &S.b, &S.pres, &S.ures, &S.vres, &S.tmp,
                         &S.links.a_e, &S.links.a_w, &S.links.a_n, &S.links.a_s, &S.links.a_p, &S.links.src};
    for (double** f : fields) {
      *f = alloc_elems(s, S, e, emax);
      if (!*f) { ok = false; break; }
    }
    if (ok) { S.scal = alloc_elems(s, S, 8, 8); ok = S.scal != nullptr; }
    if (ok && cfg->piso_corrections <= -1) { S.p_old = alloc_elems(s, S, e, emax); ok = S.p_old != nullptr; }
    if (ok && cfg->piso_corrections == -2) {
      S.u_old = alloc_elems(s, S, e, emax);
      S.v_old = alloc_elems(s, S, e, emax);
      S.cscal = alloc_elems(s, S, 8, 8);
      ok = S.u_old && S.v_old && S.cscal;
    }
    if (ok && cfg->momentum_solver == 1) {
      S.ap_un = alloc_elems(s, S, e, emax);
      S.src_un = alloc_elems(s, S, e, emax);
      S.mwork = alloc_elems(s, S, e * 5, emax * 5);
      ok = S.ap_un && S.src_un && S.mwork;
    }
    if (ok && (cfg->pressure_solver == 3 || cfg->pressure_solver == 4 || cfg->pressure_solver == 7)) {
      const size_t nw = cfg->pressure_solver == 3 ? 4 : (cfg->pressure_solver == 4 ? 5 : 7);
      S.kwork = alloc_elems(s, S, e * nw, emax * nw);
      ok = S.kwork != nullptr;
      if (ok) { S.kstate = alloc_elems(s, S, 64, 64); ok = S.kstate != nullptr; }
    }
    S.u_star = S.ua;
    S.v_star = S.va;
  }
  if (!ok) {
    ctx->err = std::string("SIMPLE state allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    nf_simple_destroy(s);
    return NF_ERR_ALLOC;
  }
  if (cfg->pressure_solver == 0 || cfg->pressure_solver == 7) {  // the solver itself / BiCGSTAB's preconditioner
    nf_mg_config mc = cfg->mg;
    mc.length = cfg->length; mc.height = cfg->height; mc.rho = 1.0;  // callers hard-code rho = 1 (multigrid.py:151)
    int st = nfi_mg_create(team, &s->mg, s->geom.nx, s->geom.ny, s->geom.ld, &mc);
    if (st != NF_OK) { nf_simple_destroy(s); return st; }
  }
  // initial fields: zeros with the velocity BCs applied on every stored row (base_algorithm.py:68-93)
  for (int k = 0; k < nl; ++k) {
    const nf_grid g = s->geom.grid_ext(team->local[k], NF_HALO);
    int st = nfi_apply_velocity_bc(ctx, &g, &s->cfg.bc, s->s[k].u, s->s[k].v);
    if (st != NF_OK) { nf_simple_destroy(s); return st; }
  }
  s->bc_clean = true;
  *out = s;
  return NF_OK;
}

extern "C" int nf_simple_create(nf_ctx* ctx, nf_simple** out, const nf_simple_config* cfg) {
  nf_team* team = nullptr;
  NF_TRY(nf_team_create_local(ctx, 1, &team));
  int st = nfi_simple_create(team, out, cfg);
  if (st != NF_OK) { nf_team_destroy(team); return st; }
  (*out)->owns_team = true;
  return NF_OK;
}

// slab-decomposed state on an existing team (nf_team_create_nccl / nf_team_create_virtual); the team stays
// owned by the caller
extern "C" int nf_simple_create_team(nf_team* team, nf_simple** out, const nf_simple_config* cfg) {
  if (!team) return NF_ERR_ARG;
  return nfi_simple_create(team, out, cfg);
}

extern "C" int nf_simple_ld(nf_simple* s) { return s ? s->geom.ld : 0; }

extern "C" int nf_mg_smoother_timing(nf_mg* mg, int on, double* total_ms, long long* launches);
// live CUDA-event timing of the finest-level fused smoother launches inside the outer iterations (single slab)
extern "C" int nf_simple_smoother_timing(nf_simple* s, int on, double* total_ms, long long* launches) {
  if (!s || !s->mg) return NF_ERR_ARG;
  return nf_mg_smoother_timing(s->mg, on, total_ms, launches);
}

static double* field_ptr(SimpleSlab& S, int which) {
  switch (which) {
    case F_U: return S.u;
    case F_V: return S.v;
    case F_P: return S.p;
    case F_USTAR: return S.u_star;
    case F_VSTAR: return S.v_star;
    case F_DU: return S.d_u;
    case F_DV: return S.d_v;
    case F_PPRIME: return S.pp;
    case F_B: return S.b;
    case F_PRES: return S.pres;
    case F_URES: return S.ures;
    case F_VRES: return S.vres;
  }
  return nullptr;
}

// first local slab's array (single-GPU runs: the whole field)
extern "C" double* nf_simple_field(nf_simple* s, int which) { return s ? field_ptr(s->s[0], which) : nullptr; }

// rows [*row_begin, *row_end) of the cell grid owned by local slab k of this process
extern "C" int nf_simple_local_rows(nf_simple* s, int k, int* row_begin, int* row_end) {
  if (!s || k < 0 || k >= nlocal(s)) return NF_ERR_ARG;
  const int r = s->team->local[k];
  if (row_begin) *row_begin = s->geom.gb[r];
  if (row_end) *row_end = s->geom.ge[r];
  return NF_OK;
}

// host arrays are the FULL (rows, cols) fields; every local slab takes / returns its own rows
extern "C" int nf_simple_upload(nf_simple* s, int which, const double* host, int rows, int cols) {
  if (!s) return NF_ERR_ARG;
  nf_ctx* ctx = s->ctx;
  NF_REQUIRE(ctx, host && which >= 0 && which < F_COUNT, "bad field / NULL host pointer");
  NF_REQUIRE(ctx, rows >= 1 && rows <= s->geom.nx + 1 && cols >= 1 && cols <= s->geom.ld, "shape does not fit the field");
  for (int k = 0; k < nlocal(s); ++k) {
    const int r = s->team->local[k];
    const int r0 = s->geom.row0(r);
    int r1 = s->geom.row1(r);
    if (r1 > rows) r1 = rows;
    if (r1 <= r0) continue;
    NF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(field_ptr(s->s[k], which), (size_t)s->geom.ld * sizeof(double),
                                         host + (size_t)r0 * cols, (size_t)cols * sizeof(double),
                                         (size_t)cols * sizeof(double), r1 - r0, cudaMemcpyHostToDevice, ctx->stream));
  }
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (which == F_U || which == F_V) s->bc_clean = false;
  return NF_OK;
}

// writes the rows owned by this process's slabs into the full host array (other rows are left untouched)
extern "C" int nf_simple_download(nf_simple* s, int which, double* host, int rows, int cols) {
  if (!s) return NF_ERR_ARG;
  nf_ctx* ctx = s->ctx;
  NF_REQUIRE(ctx, host && which >= 0 && which < F_COUNT, "bad field / NULL host pointer");
  NF_REQUIRE(ctx, rows >= 1 && rows <= s->geom.nx + 1 && cols >= 1 && cols <= s->geom.ld, "shape does not fit the field");
  for (int k = 0; k < nlocal(s); ++k) {
    const int r = s->team->local[k];
    const int r0 = s->geom.row0(r);
    int b = s->geom.gb[r], e = s->geom.ge[r];
    if (r == s->team->world - 1) e = rows;  // the last rank also owns face row nx of u
    if (e > rows) e = rows;
    if (e <= b) continue;
    NF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(host + (size_t)b * cols, (size_t)cols * sizeof(double),
                                         field_ptr(s->s[k], which) + (size_t)(b - r0) * s->geom.ld,
                                         (size_t)s->geom.ld * sizeof(double), (size_t)cols * sizeof(double), e - b,
                                         cudaMemcpyDeviceToHost, ctx->stream));
  }
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NF_OK;
}

// history record: [0] sum r_p^2 (or payload a), [1] sum b_p^2 (payload b), [2] sum r_u^2, [3] sum b_u^2,
//                 [4] sum r_v^2, [5] sum b_v^2, [6] pressure iterations, [7] spare (SIMPLEC: total residual),
//                 [8], [9] sum of squares of the UNRELAXED u / v momentum residual (cfg.track_unrelaxed_residual)
__global__ void k_store_hist_momentum(const double* __restrict__ scal, double* __restrict__ rec, int with_unrelaxed) {
  const int t = threadIdx.x;
  if (t >= 2 && t < 6) rec[t] = scal[t];
  if (t == 8 || t == 9) rec[t] = with_unrelaxed ? scal[t - 2] : 0.0;  // scal[6], scal[7]
}

__global__ void k_store_hist_pressure(const double* __restrict__ pscal, double* __restrict__ rec, double pa, double pb,
                                      double iters, int p_from_scalars) {
  const int t = threadIdx.x;
  if (t == 0) rec[0] = p_from_scalars ? pscal[0] : pa;
  if (t == 1) rec[1] = p_from_scalars ? pscal[1] : pb;
  if (t == 6) rec[6] = iters;
  if (t == 7) rec[7] = 0.0;
}

// sum (a - b)^2 over the cells of g -> out[0]
__global__ void k_diff_sumsq(nf_grid g, const double* __restrict__ a, const double* __restrict__ b, double* partials,
                             unsigned int* ticket, double* out) {
  double acc[1] = {0.0};
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.ny)
    for (int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y; i < g.ge; i += gridDim.y * blockDim.y) {
      const size_t k = nf_idx(g, i, j);
      const double d = a[k] - b[k];
      acc[0] += d * d;
    }
  nf_block_reduce_store<1>(acc, partials, ticket, out);
}

// ---- SIMPLEC helpers (Algorithms/simplec.py) ----
// d <- d / c over `rows` x `cols` (:126-127; NaN entries of the array borders stay NaN)
__global__ void k_scale_div(int rows, int cols, int ld, double* __restrict__ d, double c) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < rows && j < cols) d[(size_t)i * ld + j] = d[(size_t)i * ld + j] / c;
}

// *out = max(*out, max |a - b|) over rows x cols (np.max(np.abs(a - b)), :119-122, :157, :169-171).  |x| >= 0, so the
// ordering of the bit patterns is the ordering of the values: one atomicMax per block, order independent.
__global__ void k_maxabs_diff(int rows, int cols, int ld, const double* __restrict__ a, const double* __restrict__ b,
                              double* out) {
  double m = 0.0;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cols)
    for (int i = blockIdx.y * blockDim.y + threadIdx.y; i < rows; i += gridDim.y * blockDim.y)
      m = fmax(m, fabs(a[(size_t)i * ld + j] - b[(size_t)i * ld + j]));
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
  if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0)
    atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
}

// p'_smooth: 0.6 p' + 0.1 (((E + W) + N) + S) in the interior, 0 on the boundary ring (:141-147)
__global__ void k_smooth_pprime(nf_grid g, const double* __restrict__ pp, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.ny || i >= g.ge) return;
  const size_t k = nf_idx(g, i, j);
  double v = 0.0;
  if (i >= 1 && i <= g.nx - 2 && j >= 1 && j <= g.ny - 2)
    v = 0.6 * pp[k] + 0.1 * (((pp[k + g.ld] + pp[k - g.ld]) + pp[k + 1]) + pp[k - 1]);
  out[k] = v;
}

// p = p* + alpha p' without the zero-gradient edge copies (:154)
__global__ void k_axpy_pressure(nf_grid g, const double* __restrict__ ps, const double* __restrict__ pp, double alpha,
                                double* __restrict__ p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.ny || i >= g.ge) return;
  const size_t k = nf_idx(g, i, j);
  p[k] = ps[k] + alpha * pp[k];
}

// SIMPLEC record: [0] max|p - p_old|, [1] momentum residual, [6] pressure iterations, [7] total residual
__global__ void k_store_hist_simplec(const double* __restrict__ cs, double* __restrict__ rec, double iters) {
  if (threadIdx.x == 0) {
    rec[0] = cs[2];
    rec[1] = fmax(cs[0], cs[1]);
    rec[6] = iters;
    rec[7] = fmax(cs[3], cs[4]);
  }
}

static int maxabs_diff(nf_ctx* ctx, int rows, int cols, int ld, const double* a, const double* b, double* out) {
  dim3 block(128, 2, 1);
  int gy = (rows + 1) / 2;
  if (gy > 296) gy = 296;
  dim3 grid((cols + 127) / 128, gy < 1 ? 1 : gy, 1);
  k_maxabs_diff<<<grid, block, 0, ctx->stream>>>(rows, cols, ld, a, b, out);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

static int ensure_hist(nf_simple* s, int n) {
  nf_ctx* ctx = s->ctx;
  if (n <= s->hist_cap) return NF_OK;
  if (s->hist) cudaFree(s->hist);
  if (s->hist_host) cudaFreeHost(s->hist_host);
  s->hist = nullptr; s->hist_host = nullptr; s->hist_cap = 0;
  NF_CHECK_CUDA(ctx, cudaMalloc(&s->hist, (size_t)n * NF_HIST * sizeof(double)));
  NF_CHECK_CUDA(ctx, cudaMallocHost(&s->hist_host, (size_t)n * NF_HIST * sizeof(double)));
  s->hist_cap = n;
  return NF_OK;
}

static void decode_record(const nf_simple* s, const double* rec, nf_simple_info* out) {
  if (s->cfg.piso_corrections == -2) {  // SIMPLEC: infinity norms (simplec.py:119-122, :157, :169-171)
    out->u_rel_norm = out->v_rel_norm = rec[7];
    out->u_abs_res = out->v_abs_res = rec[1];
    out->p_rel_norm = rec[0];
    out->pressure_iterations = (int)rec[6];
    out->pad = 0;
    out->u_unrelaxed_res = sqrt(rec[8]);
    out->v_unrelaxed_res = sqrt(rec[9]);
    return;
  }
  out->u_abs_res = sqrt(rec[2]);
  out->v_abs_res = sqrt(rec[4]);
  if (s->cfg.momentum_solver == 1) {  // absolute unrelaxed residual norm (matrix_free_momentum.py:455, :527)
    out->u_rel_norm = sqrt(rec[2]);
    out->v_rel_norm = sqrt(rec[4]);
  } else {
    out->u_rel_norm = sqrt(rec[2]) / (sqrt(rec[3]) + 1e-15);  // jacobi_matrix_solver.py:246-250
    out->v_rel_norm = sqrt(rec[4]) / (sqrt(rec[5]) + 1e-15);
  }
  if (s->cfg.piso_corrections == -1) {  // SIMPLER: ||p - p_old|| / (sqrt(nx ny) + SMALL) (simpler.py:19, :172)
    out->p_rel_norm = sqrt(rec[0]) / (sqrt((double)s->cfg.nx * (double)s->cfg.ny) + 1.0e-30);
  } else switch (s->cfg.pressure_solver) {
    case 0: case 1: case 2: case 5: case 6:
      out->p_rel_norm = sqrt(rec[0]); break;  // absolute ||b - A p'|| (multigrid.py:257)
    default: out->p_rel_norm = rec[1] > 0.0 ? sqrt(rec[0]) / sqrt(rec[1]) : sqrt(rec[0]); break;  // ||r_int||/||b_int||
  }
  out->pressure_iterations = (int)rec[6];
  out->pad = 0;
  out->u_unrelaxed_res = sqrt(rec[8]);
  out->v_unrelaxed_res = sqrt(rec[9]);
}

// a7: MatrixFreeMomentumSolver.solve_u/v_momentum (matrix_free_momentum.py:403-544), single slab.  ubc / vbc hold the
// velocities with the BCs applied the way that class does it (caller's nx+1).
static int momentum_component_krylov(nf_simple* s, int is_u, double alpha, int want_fields) {
  nf_ctx* ctx = s->ctx;
  const nf_simple_config& c = s->cfg;
  SimpleSlab& S = s->s[0];
  const nf_grid g = s->geom.grid(s->team->local[0]);
  double* d = is_u ? S.d_u : S.d_v;
  NF_TRY(nfi_momentum_links_mf(ctx, &g, is_u, S.ubc, S.vbc, S.p, c.mu, alpha, c.sides, S.links, d, S.ap_un, S.src_un));
  double* x = is_u ? S.ua : S.va;
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(x, is_u ? S.u : S.v, s->geom.elems(s->team->local[0]) * sizeof(double),
                                     cudaMemcpyDeviceToDevice, ctx->stream));  // x0 = the raw velocity (:436, :508)
  nf_krylov_info ki;
  NF_TRY(nfi_momentum_bicgstab(ctx, &g, is_u, S.links, x, c.momentum_tolerance, 1e-5, c.momentum_maxiter, 10, S.mwork, &ki));
  // BCs on the solution; the partner argument is the BC'd copy (idempotent) (:438, :510)
  if (is_u) NF_TRY(nfi_apply_velocity_bc(ctx, &g, &c.bc_mf, x, S.vbc));
  else NF_TRY(nfi_apply_velocity_bc(ctx, &g, &c.bc_mf, S.ubc, x));
  if (is_u) S.u_star = x; else S.v_star = x;
  nf_links Lun = S.links;
  Lun.a_p = S.ap_un;
  Lun.src = S.src_un;
  NF_TRY(nfi_momentum_residual_unrelaxed(ctx, &g, is_u, Lun, x, want_fields ? (is_u ? S.ures : S.vres) : nullptr,
                                         S.scal + (is_u ? 2 : 4)));
  return NF_OK;
}

// momentum predictor of one component on every local slab (no communication, see the header comment)
// track: also leave the sum of squares of the UNRELAXED residual of the predicted component in scal[6] (u) / scal[7] (v)
static int momentum_component(nf_simple* s, int is_u, double alpha, int want_fields, int track = 0) {
  nf_ctx* ctx = s->ctx;
  nf_team* team = s->team;
  const nf_simple_config& c = s->cfg;
  const bool dist = s->geom.dist;
  const int nl = nlocal(s);
  int margin = NF_HALO - 1;  // rows beyond the owned ones on which the current iterate is exact
  for (int k = 0; k < nl; ++k) {
    SimpleSlab& S = s->s[k];
    const nf_grid g = s->geom.grid_ext(team->local[k], margin);
    const double* ubc = s->bc_clean ? S.u : S.ubc;
    const double* vbc = s->bc_clean ? S.v : S.vbc;
    NF_TRY(nfi_momentum_links(ctx, &g, is_u, ubc, vbc, S.p, c.mu, alpha, c.sides, S.links, is_u ? S.d_u : S.d_v));
  }
  // sweeps: x0 = current velocity (valid NF_HALO rows out); sweep s is exact margin-1 rows out
  std::vector<const double*> src(nl);
  std::vector<double*> a(nl), b(nl);
  for (int k = 0; k < nl; ++k) {
    SimpleSlab& S = s->s[k];
    src[k] = is_u ? S.u : S.v;
    a[k] = is_u ? S.ua : S.va;
    b[k] = is_u ? S.ub : S.vb;
  }
  if (c.n_momentum_sweeps == 0) {
    for (int k = 0; k < nl; ++k) {
      NF_CHECK_CUDA(ctx, cudaMemcpyAsync(a[k], src[k], s->geom.elems(team->local[k]) * sizeof(double),
                                         cudaMemcpyDeviceToDevice, ctx->stream));
      src[k] = a[k];
    }
  }
  // Sweeps in chunks of up to 6 per tile load (nf_momentum_fused.cu); the last chunk also evaluates the residual
  // norms.  On slabs each chunk consumes k (+1) rows of halo: the computed region shrinks instead of communicating.
  const char* envf = getenv("NF_MOMENTUM_FUSED");
  const bool fused = !(envf && envf[0] == '0');
  bool res_done = false;
  int left = c.n_momentum_sweeps;
  while (left > 0) {
    int k = fused ? (left > 6 ? 6 : left) : 1;
    const bool last = (left == k);
    const int need = k + ((fused && last) ? 1 : 0);  // halo rows this chunk consumes
    if (dist && margin - need < 1) {  // out of halo: refresh it and start shrinking again
      std::vector<double*> f(nl);
      for (int kk = 0; kk < nl; ++kk) f[kk] = const_cast<double*>(src[kk]);
      NF_TRY(nf_team_exchange(team, s->geom, f.data(), NF_HALO));
      margin = NF_HALO;
      if (margin - need < 1) k = 1;  // cannot happen with NF_HALO = 8 and k <= 6, kept for safety
    }
    const int out_margin = dist ? (last && fused ? 1 : margin - k) : 0;
    for (int kk = 0; kk < nl; ++kk) {
      SimpleSlab& S = s->s[kk];
      const int r = team->local[kk];
      const nf_grid g = s->geom.grid_ext(r, out_margin);
      const nf_grid gown = s->geom.grid(r);
      double* dst = (src[kk] == a[kk]) ? b[kk] : a[kk];
      if (fused) {
        const int nb = gown.gb, ne = (is_u && gown.ge == gown.nx) ? gown.nx + 1 : gown.ge;
        NF_TRY(nfi_momentum_sweeps_fused(ctx, &g, is_u, S.links, src[kk], dst, k, last ? 1 : 0, nb, ne,
                                         (last && want_fields) ? (is_u ? S.ures : S.vres) : nullptr,
                                         S.scal + (is_u ? 2 : 4)));
      } else {
        NF_TRY(nfi_momentum_sweep(ctx, &g, is_u, S.links, src[kk], dst));
      }
      src[kk] = dst;
    }
    margin = out_margin;
    if (fused && last) res_done = true;
    left -= k;
  }
  for (int k = 0; k < nl; ++k) {
    SimpleSlab& S = s->s[k];
    if (is_u) S.u_star = const_cast<double*>(src[k]); else S.v_star = const_cast<double*>(src[k]);
  }
  if (dist && margin < 1) {  // the continuity RHS and the residual read one row beyond the owned ones
    std::vector<double*> f(nl);
    for (int k = 0; k < nl; ++k) f[k] = const_cast<double*>(src[k]);
    NF_TRY(nf_team_exchange(team, s->geom, f.data(), 1));
  }
  if (!res_done)
    for (int k = 0; k < nl; ++k) {
      SimpleSlab& S = s->s[k];
      const nf_grid g = s->geom.grid(team->local[k]);
      NF_TRY(nfi_momentum_residual_to(ctx, &g, is_u, S.links, src[k], want_fields ? (is_u ? S.ures : S.vres) : nullptr,
                                      S.scal + (is_u ? 2 : 4)));
    }
  if (track)  // the component's links are still in place (the other component overwrites them next)
    for (int k = 0; k < nl; ++k) {
      SimpleSlab& S = s->s[k];
      const nf_grid g = s->geom.grid(team->local[k]);
      const double* phi_old = is_u ? (s->bc_clean ? S.u : S.ubc) : (s->bc_clean ? S.v : S.vbc);
      NF_TRY(nfi_momentum_unrelaxed_from_relaxed(ctx, &g, is_u, S.links, src[k], phi_old, alpha, S.scal + (is_u ? 6 : 7)));
    }
  return NF_OK;
}

// velocities with BCs applied, used for the coefficients (jacobi_matrix_solver.py:170)
static int refresh_bc_copies(nf_simple* s) {
  nf_ctx* ctx = s->ctx;
  nf_team* team = s->team;
  if (s->bc_clean) return NF_OK;
  for (int k = 0; k < nlocal(s); ++k) {
    SimpleSlab& S = s->s[k];
    const size_t bytes = s->geom.elems(team->local[k]) * sizeof(double);
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(S.ubc, S.u, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(S.vbc, S.v, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    const nf_grid g = s->geom.grid_ext(team->local[k], NF_HALO);
    NF_TRY(nfi_apply_velocity_bc(ctx, &g, &s->cfg.bc, S.ubc, S.vbc));
  }
  return NF_OK;
}

// momentum predictor of both components from the same (u, v, p): simple.py:121-133, piso.py:58-71 / :92-104.
// slot >= 0: the relaxed residual sums of this solve are the iteration's record (hist[slot][2..5])
static int momentum_predictor(nf_simple* s, double alpha, int want_fields, int slot) {
  nf_ctx* ctx = s->ctx;
  if (s->cfg.momentum_solver == 1) {
    SimpleSlab& S = s->s[0];
    const size_t bytes = s->geom.elems(s->team->local[0]) * sizeof(double);
    const nf_grid g = s->geom.grid(s->team->local[0]);
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(S.ubc, S.u, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(S.vbc, S.v, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    NF_TRY(nfi_apply_velocity_bc(ctx, &g, &s->cfg.bc_mf, S.ubc, S.vbc));
    NF_TRY(momentum_component_krylov(s, 1, alpha, want_fields));
    NF_TRY(momentum_component_krylov(s, 0, alpha, want_fields));
  } else {
    const int trk = (slot >= 0 && s->cfg.track_unrelaxed_residual != 0) ? 1 : 0;
    NF_TRY(refresh_bc_copies(s));
    NF_TRY(momentum_component(s, 1, alpha, want_fields, trk));
    NF_TRY(momentum_component(s, 0, alpha, want_fields, trk));
  }
  if (slot < 0) return NF_OK;
  const int nl = nlocal(s);
  const int track = (s->cfg.track_unrelaxed_residual != 0 && s->cfg.momentum_solver == 0) ? 1 : 0;
  if (s->geom.dist) {  // momentum sums of all slabs (scal[2..5], with the unrelaxed sums scal[6..7])
    std::vector<double*> sc(nl);
    for (int k = 0; k < nl; ++k) sc[k] = s->s[k].scal + 2;
    NF_TRY(nf_team_allreduce(s->team, sc.data(), track ? 6 : 4));
  }
  k_store_hist_momentum<<<1, 32, 0, ctx->stream>>>(s->s[0].scal, s->hist + (size_t)slot * NF_HIST, track);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

// pressure correction from (u*, v*, d_u, d_v): continuity RHS, pressure solve, p = p* + alpha_p p' with zero-gradient
// edges (p* <- p), velocity correction + BCs (simple.py:136-155, piso.py:73-90).  slot >= 0: the pressure part of the
// iteration's record goes to hist[slot]
// alpha: relaxation of the pressure update; correct: also correct the velocities; diff_record (SIMPLER): the record's
// pressure entry is sum (p - p_old)^2 instead of the pressure solver's residual
// simplec: the update part follows simplec.py:141-171 instead (smoothed p', plain p update, infinity-norm record)
static int pressure_correction(nf_simple* s, int slot, double alpha, bool correct = true, bool diff_record = false,
                               bool simplec = false) {
  nf_ctx* ctx = s->ctx;
  nf_team* team = s->team;
  const nf_simple_config& c = s->cfg;
  const int nl = nlocal(s);
  const bool dist = s->geom.dist;
  // pressure correction
  std::vector<nf_grid> gp(nl);
  for (int k = 0; k < nl; ++k) {
    SimpleSlab& S = s->s[k];
    gp[k] = s->geom.grid(team->local[k]);
    gp[k].rho = 1.0;  // every pressure solver of the reference hard-codes rho = 1.0 (multigrid.py:151, jacobi.py, ...)
    NF_TRY(nf_continuity_rhs(ctx, &gp[k], S.u_star, S.v_star, S.b));
  }
  if (dist) {
    std::vector<double*> b = field_of(s, &SimpleSlab::b);
    NF_TRY(nf_team_exchange(team, s->geom, b.data(), NF_HALO));
  }
  double pa = 0.0, pb = 0.0, iters = 0.0;
  int p_from_scalars = 0;
  const double* pscal = s->s[0].scal;
  switch (c.pressure_solver) {
    case 0: {
      std::vector<double*> du = field_of(s, &SimpleSlab::d_u), dv = field_of(s, &SimpleSlab::d_v);
      std::vector<double*> b = field_of(s, &SimpleSlab::b), x = field_of(s, &SimpleSlab::pp);
      std::vector<double*> r = field_of(s, &SimpleSlab::pres);
      NF_TRY(nfi_mg_setup(s->mg, du.data(), dv.data()));
      nf_mg_info mi;
      const int fmg = (c.mg.cycle_type == 2);
      NF_TRY(nfi_mg_solve(s->mg, b.data(), x.data(), r.data(), &mi, fmg ? 0 : 1, s->want_fields));
      if (fmg) { p_from_scalars = 1; pscal = nfi_mg_scalars(s->mg, 0); }
      else { pa = mi.r_norm * mi.r_norm; pb = mi.b_norm * mi.b_norm; }
      iters = mi.cycles;
      break;
    }
    case 1:
    case 2: {
      if (!dist) {
        SimpleSlab& S = s->s[0];
        const nf_grid* g1 = &gp[0];
        NF_TRY(nfi_fill(ctx, S.pp, (size_t)g1->nx * g1->ld, 0.0));
        if (c.pressure_solver == 1)
          NF_TRY(nfi_jacobi(ctx, g1, S.pp, S.tmp, S.b, S.d_u, S.d_v, c.pressure_omega, c.pressure_iterations));
        else
          NF_TRY(nfi_rbsor(ctx, g1, S.pp, S.b, S.d_u, S.d_v, c.pressure_omega, c.pressure_iterations));
        NF_TRY(nfi_residual_norms(ctx, g1, S.pp, S.b, S.d_u, S.d_v, S.pres, 1, S.scal));
      } else {
        // slabs: Jacobi exchanges one halo row per iteration; red-black SOR runs up to 3 sweeps per launch on a region
        // that shrinks by two rows per sweep (nf_rbsor_fused.cu, bit-identical to the colour passes) and exchanges
        // 2 x sweeps rows per launch
        for (int k = 0; k < nl; ++k) NF_TRY(nfi_fill(ctx, s->s[k].pp, s->geom.elems(team->local[k]), 0.0));
        int left = c.pressure_iterations;
        while (left > 0) {
          const int ns = c.pressure_solver == 1 ? 1 : (left >= 3 ? 3 : left);
          for (int k = 0; k < nl; ++k) {
            SimpleSlab& S = s->s[k];
            if (c.pressure_solver == 1)
              NF_TRY(nfi_jacobi(ctx, &gp[k], S.pp, S.tmp, S.b, S.d_u, S.d_v, c.pressure_omega, 1));
            else
              NF_TRY(nfi_rbsor_fused_x(ctx, &gp[k], &S.pp, &S.tmp, S.b, S.d_u, S.d_v, nullptr, c.pressure_omega, ns,
                                       nullptr));
          }
          std::vector<double*> x = field_of(s, &SimpleSlab::pp);
          NF_TRY(nf_team_exchange(team, s->geom, x.data(), c.pressure_solver == 1 ? 1 : 2 * ns));
          left -= ns;
        }
        std::vector<double*> sc(nl);
        for (int k = 0; k < nl; ++k) {
          SimpleSlab& S = s->s[k];
          NF_TRY(nfi_residual_norms(ctx, &gp[k], S.pp, S.b, S.d_u, S.d_v, S.pres, 1, S.scal));
          sc[k] = S.scal;
        }
        NF_TRY(nf_team_allreduce(team, sc.data(), 2));
      }
      p_from_scalars = 1;
      iters = c.pressure_iterations;
      break;
    }
    case 5:
    case 6: {  // sequential SOR sweeps (single slab)
      SimpleSlab& S = s->s[0];
      const nf_grid* g1 = &gp[0];
      NF_TRY(nfi_fill(ctx, S.pp, (size_t)g1->nx * g1->ld, 0.0));
      NF_TRY(nfi_gs_lex(ctx, g1, S.pp, S.b, S.d_u, S.d_v, c.pressure_omega, c.pressure_iterations, c.pressure_solver == 6));
      NF_TRY(nfi_residual_norms(ctx, g1, S.pp, S.b, S.d_u, S.d_v, S.pres, 1, S.scal));
      p_from_scalars = 1;
      iters = c.pressure_iterations;
      break;
    }
    default: {
      nf_krylov_info ki;
      const int poll = c.krylov_check_every > 0 ? c.krylov_check_every : (c.pressure_solver == 3 ? 25 : 10);
      if (!dist) {
        SimpleSlab& S = s->s[0];
        const nf_grid* g1 = &gp[0];
        if (c.pressure_solver == 3) {
          NF_TRY(nf_cg_solve(ctx, g1, S.b, S.pp, S.d_u, S.d_v, c.pressure_tolerance, 1e-5, c.krylov_maxiter, poll, S.kwork,
                             &ki));
        } else if (c.pressure_solver == 7) {  // M = multigrid cycles on the same coefficients (matrix_free_BiCGSTAB.py:102-161)
          std::vector<double*> du = field_of(s, &SimpleSlab::d_u), dv = field_of(s, &SimpleSlab::d_v);
          NF_TRY(nfi_mg_setup(s->mg, du.data(), dv.data()));
          NF_TRY(nf_bicgstab_solve_mg(ctx, g1, S.b, S.pp, S.d_u, S.d_v, c.pressure_tolerance, 1e-5, c.krylov_maxiter, poll,
                                      S.kwork, s->mg, c.krylov_mg_cycles, c.krylov_mg_kind, &ki));
        } else {
          NF_TRY(nf_bicgstab_solve(ctx, g1, S.b, S.pp, S.d_u, S.d_v, c.pressure_tolerance, 1e-5, c.krylov_maxiter, poll,
                                   S.kwork, &ki));
        }
        // rel_norm = ||r_int|| / ||b_int|| of the true residual (matrix_free_BiCGSTAB.py:255-279)
        NF_TRY(nfi_residual(ctx, g1, S.pp, S.b, S.d_u, S.d_v, S.pres));
        NF_TRY(nfi_sumsq_to(ctx, g1, S.pres, 1, S.scal));
        NF_TRY(nfi_sumsq_to(ctx, g1, S.b, 1, S.scal + 1));
      } else {
        std::vector<double*> du = field_of(s, &SimpleSlab::d_u), dv = field_of(s, &SimpleSlab::d_v);
        std::vector<double*> b = field_of(s, &SimpleSlab::b), x = field_of(s, &SimpleSlab::pp);
        std::vector<double*> w = field_of(s, &SimpleSlab::kwork), ks = field_of(s, &SimpleSlab::kstate);
        NF_TRY(nf_team_exchange(team, s->geom, du.data(), 1));  // the operator reads d_u[i+1] and the neighbours' rows
        NF_TRY(nf_team_exchange(team, s->geom, dv.data(), 1));
        NF_TRY(nfi_krylov_team(team, s->geom, c.pressure_solver == 3 ? 0 : 1, b.data(), x.data(), du.data(), dv.data(),
                               c.pressure_tolerance, 1e-5, c.krylov_maxiter, poll, w.data(),
                               ks.data(), &ki));
        NF_TRY(nf_team_exchange(team, s->geom, x.data(), 1));
        std::vector<double*> sc(nl);
        for (int k = 0; k < nl; ++k) {
          SimpleSlab& S = s->s[k];
          NF_TRY(nfi_residual(ctx, &gp[k], S.pp, S.b, S.d_u, S.d_v, S.pres));
          NF_TRY(nfi_sumsq_to(ctx, &gp[k], S.pres, 1, S.scal));
          NF_TRY(nfi_sumsq_to(ctx, &gp[k], S.b, 1, S.scal + 1));
          sc[k] = S.scal;
        }
        NF_TRY(nf_team_allreduce(team, sc.data(), 2));
      }
      p_from_scalars = 1;
      iters = ki.iterations;
      break;
    }
  }
  phase_mark(s);  // end of the pressure solve
  if (simplec) {
    SimpleSlab& S = s->s[0];
    const nf_grid g = s->geom.grid(team->local[0]);
    NfLaunch2D l = nf_launch2d(g.ge - g.gb, g.ny);
    k_smooth_pprime<<<l.grid, l.block, 0, ctx->stream>>>(g, S.pp, S.tmp);
    NF_LAUNCH_CHECK(ctx);
    { double* t = S.pp; S.pp = S.tmp; S.tmp = t; }
    k_axpy_pressure<<<l.grid, l.block, 0, ctx->stream>>>(g, S.p, S.pp, alpha, S.p_alt);
    NF_LAUNCH_CHECK(ctx);
    { double* t = S.p; S.p = S.p_alt; S.p_alt = t; }
    NF_TRY(maxabs_diff(ctx, g.nx, g.ny, g.ld, S.p, S.p_old, S.cscal + 2));
    NF_TRY(nfi_correct_velocity(ctx, &g, &c.bc, S.u_star, S.v_star, S.pp, S.d_u, S.d_v, S.u, S.v));
    NF_TRY(maxabs_diff(ctx, g.nx + 1, g.ny, g.ld, S.u, S.u_old, S.cscal + 3));
    NF_TRY(maxabs_diff(ctx, g.nx, g.ny + 1, g.ld, S.v, S.v_old, S.cscal + 4));
    if (slot >= 0) {
      k_store_hist_simplec<<<1, 32, 0, ctx->stream>>>(S.cscal, s->hist + (size_t)slot * NF_HIST, iters);
      NF_LAUNCH_CHECK(ctx);
    }
    s->bc_clean = true;
    return NF_OK;
  }
  if (slot >= 0 && !diff_record) {
    k_store_hist_pressure<<<1, 32, 0, ctx->stream>>>(pscal, s->hist + (size_t)slot * NF_HIST, pa, pb, iters, p_from_scalars);
    NF_LAUNCH_CHECK(ctx);
  }
  // p = p* + alpha p' with zero-gradient edges; p* <- p.  velocity correction + BCs.
  for (int k = 0; k < nl; ++k) {
    SimpleSlab& S = s->s[k];
    const nf_grid g = s->geom.grid(team->local[k]);
    NF_TRY(nf_update_pressure(ctx, &g, S.p, S.pp, alpha, S.p_alt));
    { double* t = S.p; S.p = S.p_alt; S.p_alt = t; }
    if (correct) NF_TRY(nfi_correct_velocity(ctx, &g, &c.bc, S.u_star, S.v_star, S.pp, S.d_u, S.d_v, S.u, S.v));
  }
  if (slot >= 0 && diff_record) {  // SIMPLER: p_rel_norm = ||p - p_old|| / sqrt(nx ny) (simpler.py:172)
    std::vector<double*> sc(nl);
    for (int k = 0; k < nl; ++k) {
      SimpleSlab& S = s->s[k];
      const nf_grid g = s->geom.grid(team->local[k]);
      NfLaunch2D l = nf_launch_reduce(g.ge - g.gb, g.ny);
      k_diff_sumsq<<<l.grid, l.block, 0, ctx->stream>>>(g, S.p, S.p_old, ctx->partials, ctx->ticket, S.scal);
      NF_LAUNCH_CHECK(ctx);
      sc[k] = S.scal;
    }
    if (dist) NF_TRY(nf_team_allreduce(team, sc.data(), 1));
    k_store_hist_pressure<<<1, 32, 0, ctx->stream>>>(s->s[0].scal, s->hist + (size_t)slot * NF_HIST, 0.0, 0.0, iters, 1);
    NF_LAUNCH_CHECK(ctx);
  }
  if (dist) {
    std::vector<double*> p = field_of(s, &SimpleSlab::p);
    NF_TRY(nf_team_exchange(team, s->geom, p.data(), NF_HALO));
    if (correct) {
      std::vector<double*> u = field_of(s, &SimpleSlab::u), v = field_of(s, &SimpleSlab::v);
      NF_TRY(nf_team_exchange(team, s->geom, u.data(), NF_HALO));
      NF_TRY(nf_team_exchange(team, s->geom, v.data(), NF_HALO));
    }
  }
  if (correct) s->bc_clean = true;
  return NF_OK;
}


// one outer iteration; leaves its history record in hist[slot].
// SIMPLE (simple.py:114-212): predictor, one pressure correction.  PISO (piso.py:53-110, cfg.piso_corrections >= 1):
// predictor with alpha_u, then n corrections; between corrections the momentum equations are solved again from the
// corrected (u, v, p) without relaxation, their norms are discarded (:92-104).
static int simple_step(nf_simple* s, int slot, int want_fields) {
  const nf_simple_config& c = s->cfg;
  s->want_fields = want_fields;
  // phase marks per (predictor, correction) pair: start, end of predictor, end of pressure solve, end of corrections
  phase_mark(s);
  if (c.piso_corrections == -1) {
    // SIMPLER as the reference codes it (simpler.py:99-167): predictor; p += p-bar (the pressure solver's answer from u*, v*,
    // unrelaxed); momentum again from the same (u, v) with the new p; p' ; p += alpha_p p'; velocity correction with p'
    nf_ctx* ctx = s->ctx;
    for (int k = 0; k < nlocal(s); ++k)
      NF_CHECK_CUDA(ctx, cudaMemcpyAsync(s->s[k].p_old, s->s[k].p, s->geom.elems(s->team->local[k]) * sizeof(double),
                                         cudaMemcpyDeviceToDevice, ctx->stream));
    NF_TRY(momentum_predictor(s, c.alpha_u, want_fields, slot));
    phase_mark(s);
    NF_TRY(pressure_correction(s, -1, 1.0, false));
    phase_mark(s);
    phase_mark(s);
    NF_TRY(momentum_predictor(s, c.alpha_u, 0, -1));
    phase_mark(s);
    NF_TRY(pressure_correction(s, slot, c.alpha_p, true, true));
    phase_mark(s);
    if (s->phase_timing) s->phase_iters++;
    return NF_OK;
  }
  if (c.piso_corrections == -2) {
    // SIMPLEC as the reference codes it (simplec.py:99-171), single slab
    nf_ctx* ctx = s->ctx;
    SimpleSlab& S = s->s[0];
    const nf_grid g = s->geom.grid(s->team->local[0]);
    const size_t bytes = s->geom.elems(s->team->local[0]) * sizeof(double);
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(S.u_old, S.u, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(S.v_old, S.v, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(S.p_old, S.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    NF_CHECK_CUDA(ctx, cudaMemsetAsync(S.cscal, 0, 8 * sizeof(double), ctx->stream));
    NF_TRY(momentum_predictor(s, c.alpha_u, want_fields, slot));
    NF_TRY(maxabs_diff(ctx, g.nx + 1, g.ny, g.ld, S.u_star, S.u, S.cscal + 0));
    NF_TRY(maxabs_diff(ctx, g.nx, g.ny + 1, g.ld, S.v_star, S.v, S.cscal + 1));
    {
      dim3 block(128, 2, 1), grid((g.ny + 1 + 127) / 128, (g.nx + 1 + 1) / 2, 1);
      k_scale_div<<<grid, block, 0, ctx->stream>>>(g.nx + 1, g.ny, g.ld, S.d_u, c.simplec_divisor);
      NF_LAUNCH_CHECK(ctx);
      k_scale_div<<<grid, block, 0, ctx->stream>>>(g.nx, g.ny + 1, g.ld, S.d_v, c.simplec_divisor);
      NF_LAUNCH_CHECK(ctx);
    }
    phase_mark(s);
    NF_TRY(pressure_correction(s, slot, c.alpha_p, true, false, true));
    phase_mark(s);
    if (s->phase_timing) s->phase_iters++;
    return NF_OK;
  }
  NF_TRY(momentum_predictor(s, c.alpha_u, want_fields, slot));
  phase_mark(s);
  const int nc = c.piso_corrections >= 1 ? c.piso_corrections : 1;
  for (int k = 0; k < nc; ++k) {
    NF_TRY(pressure_correction(s, k == nc - 1 ? slot : -1, c.alpha_p));
    phase_mark(s);
    if (k < nc - 1) {
      phase_mark(s);
      NF_TRY(momentum_predictor(s, 1.0, 0, -1));
      phase_mark(s);
    }
  }
  if (s->phase_timing) s->phase_iters++;
  return NF_OK;
}

// CUDA-event timing of the three phases of the outer iteration (momentum predictor / pressure solve incl. its RHS /
// p and velocity corrections) inside nf_simple_iterate: switches the instrumentation on or off and returns + resets the
// accumulated milliseconds and the number of outer iterations they cover.  The events sit between the phases' launches
// on the context's stream (the V-cycle graph is launched between two of them), so they do not perturb the kernels.
extern "C" int nf_simple_phase_timing(nf_simple* s, int on, double* ms_momentum, double* ms_pressure, double* ms_correct,
                                      long long* iterations) {
  if (!s) return NF_ERR_ARG;
  nf_ctx* ctx = s->ctx;
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t k = 0; k + 3 < s->pev_used; k += 4) {  // groups of 4 marks: start, predictor, solve, corrections
    float a = 0.f, b = 0.f, c = 0.f;
    cudaEventElapsedTime(&a, s->pev[k], s->pev[k + 1]);
    cudaEventElapsedTime(&b, s->pev[k + 1], s->pev[k + 2]);
    cudaEventElapsedTime(&c, s->pev[k + 2], s->pev[k + 3]);
    s->phase_ms[0] += a; s->phase_ms[1] += b; s->phase_ms[2] += c;
  }
  s->pev_used = 0;
  if (ms_momentum) *ms_momentum = s->phase_ms[0];
  if (ms_pressure) *ms_pressure = s->phase_ms[1];
  if (ms_correct) *ms_correct = s->phase_ms[2];
  if (iterations) *iterations = s->phase_iters;
  s->phase_ms[0] = s->phase_ms[1] = s->phase_ms[2] = 0.0;
  s->phase_iters = 0;
  s->phase_timing = on != 0;
  return NF_OK;
}

extern "C" int nf_simple_iterate(nf_simple* s, int n_iterations, double tolerance, int want_fields,
                                 nf_simple_info* info_host, int* n_done) {
  if (!s) return NF_ERR_ARG;
  nf_ctx* ctx = s->ctx;
  NF_REQUIRE(ctx, n_iterations >= 0, "n_iterations < 0");
  NF_TRY(ensure_hist(s, n_iterations > 0 ? n_iterations : 1));
  if (s->geom.dist && n_iterations > 0) {  // halos of the state (uploads fill them from the host copy; cheap to redo)
    std::vector<double*> p = field_of(s, &SimpleSlab::p), u = field_of(s, &SimpleSlab::u), v = field_of(s, &SimpleSlab::v);
    NF_TRY(nf_team_exchange(s->team, s->geom, p.data(), NF_HALO));
    NF_TRY(nf_team_exchange(s->team, s->geom, u.data(), NF_HALO));
    NF_TRY(nf_team_exchange(s->team, s->geom, v.data(), NF_HALO));
  }
  int done = 0;
  for (int it = 0; it < n_iterations; ++it) {
    NF_TRY(simple_step(s, it, want_fields));
    ++done;
    if (tolerance > 0.0) {  // stopping test of simple.py:114 needs this iteration's norms
      NF_CHECK_CUDA(ctx, cudaMemcpyAsync(s->hist_host + (size_t)it * NF_HIST, s->hist + (size_t)it * NF_HIST, NF_HIST * sizeof(double),
                                         cudaMemcpyDeviceToHost, ctx->stream));
      NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      nf_simple_info rec;
      decode_record(s, s->hist_host + (size_t)it * NF_HIST, &rec);
      if (info_host) info_host[it] = rec;
      const double total = fmax(rec.u_rel_norm, rec.v_rel_norm);
      if (!(total > tolerance)) break;
    }
  }
  if (!(tolerance > 0.0) && done > 0) {
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(s->hist_host, s->hist, (size_t)done * NF_HIST * sizeof(double), cudaMemcpyDeviceToHost,
                                       ctx->stream));
    NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (info_host)
      for (int it = 0; it < done; ++it) decode_record(s, s->hist_host + (size_t)it * NF_HIST, &info_host[it]);
  }
  if (n_done) *n_done = done;
  if (s->team->p2p && done > 0 && nf_p2p_error(s->team)) {
    ctx->err = "peer-memory exchange timed out: the ranks of the team did not run the same sequence of exchanges";
    return NF_ERR_CUDA;
  }
  return NF_OK;
}
