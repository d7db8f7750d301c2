/*
 * naviflow_b200.h -- C-ABI of libnaviflow_b200.so (sm_100a, fp64).
 *
 * Drop-in boundary for ONE hot path of philipnickel/NaviFlow: the SIMPLE outer loop of the 2-D
 * lid-driven cavity (momentum link-coefficient assembly + Jacobi sweeps, matrix-free
 * pressure-correction operator, Jacobi / red-black SOR / geometric multigrid / CG / BiCGSTAB
 * pressure solvers, u/v/p correction, residual norms).  The reference is pure Python
 * (NumPy/SciPy); each entry point below names the reference function it replaces
 * (paths relative to /root/reference/naviflow_oo).  The Python plugin classes in
 * naviflow_b200/ bind these symbols with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).
 *
 * Conventions
 *  - All pointers are DEVICE pointers to fp64 unless the name ends in _host.
 *  - Fields are C-ordered 2-D arrays with a common row pitch `ld` (in doubles, ld >= ny+1):
 *      p-like (nx rows, ny cols), u (nx+1 rows, ny cols), v (nx rows, ny+1 cols);
 *      element [i][j] lives at base[(i - row0)*ld + j].  Pad columns are never read as data.
 *  - `nf_grid` describes one multigrid level / one slab of it.  On a single GPU row0=0,
 *      gb=0, ge=nx.  On a slab-decomposed run row0 is the global index of the first stored
 *      row (halo included) and [gb,ge) the global cell rows this rank computes.
 *  - Every function returns 0 on success or a negative nf_status; nf_last_error() gives text.
 *  - No function allocates device memory except the *_create calls; hot calls are asynchronous
 *      on the context's stream unless they return a host scalar.
 *  - One host thread per context.
 */
#ifndef NAVIFLOW_B200_H
#define NAVIFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nf_ctx nf_ctx;       /* stream, reduction scratch, error text           */
typedef struct nf_mg nf_mg;         /* multigrid hierarchy (levels, coarse inverse)     */
typedef struct nf_simple nf_simple; /* device-resident SIMPLE state (all fields + work) */
typedef struct nf_team nf_team;     /* the ranks a grid is cut into (row slabs) + their communicator */

enum nf_status {
  NF_OK = 0,
  NF_ERR_CUDA = -1,
  NF_ERR_ARG = -2,
  NF_ERR_ALLOC = -3,
  NF_ERR_UNSUPPORTED = -4
};

typedef struct nf_grid {
  int32_t nx, ny;   /* global number of cells                      */
  int32_t ld;       /* row pitch in doubles                        */
  int32_t row0;     /* global index of the first stored row        */
  int32_t gb, ge;   /* global cell-row range [gb,ge) to compute    */
  int32_t row1;     /* one past the last stored row (u-like arrays: <= nx+1); 0 = "everything" (nx+1) */
  int32_t pad;
  double dx, dy;    /* StructuredMesh spacing: L/(nx-1), H/(ny-1)  (preprocessing/mesh/structured.py:27-28) */
  double rho;
} nf_grid;

/* Edge program of BoundaryConditionManager.apply_velocity_boundary_conditions
 * (constructor/boundary_conditions.py:164-260) after evaluating the insertion-ordered
 * conditions on the host: final value per edge line / corner, NaN = "leave untouched".
 * Index order: 0 left(i=0) 1 right(i=last) 2 bottom(j=0) 3 top(j=last);
 * corners: 0 (left,bottom) 1 (left,top) 2 (right,bottom) 3 (right,top). */
typedef struct nf_bc_program {
  double u_edge[4], u_corner[4];
  double v_edge[4], v_corner[4];
  int32_t v_right_row; /* row index the "right" v edge applies to (nx-1), or -1 if skipped
                          (callers that pass nx+1, matrix_free_momentum.py:419) */
  int32_t pad;
} nf_bc_program;

/* ---- context ------------------------------------------------------------------------ */
int nf_ctx_create(nf_ctx** out, int device, void* cuda_stream /* cudaStream_t; NULL = legacy default stream */);
int nf_ctx_destroy(nf_ctx* ctx);
const char* nf_last_error(nf_ctx* ctx);
int nf_sync(nf_ctx* ctx);
int nf_version(void);
/* number of kernel launches issued through this context since creation (bench gpu_launches) */
int64_t nf_launch_count(nf_ctx* ctx);

/* ---- K1  velocity BCs: boundary_conditions.py:164-260 -------------------------------- */
int nf_apply_velocity_bc(nf_ctx*, const nf_grid*, const nf_bc_program*, double* u, double* v);

/* ---- K5  continuity RHS: pressure_solver/helpers/rhs_construction.py:3-21 ------------ */
int nf_continuity_rhs(nf_ctx*, const nf_grid*, const double* u_star, const double* v_star, double* b);

/* ---- K6  A*p and b-A*p: pressure_solver/helpers/matrix_free.py:6-135 ------------------ */
int nf_pressure_apply(nf_ctx*, const nf_grid*, const double* p, const double* d_u, const double* d_v,
                      double* out);
int nf_pressure_residual(nf_ctx*, const nf_grid*, const double* p, const double* b, const double* d_u,
                         const double* d_v, double* r);

/* ---- K7  weighted Jacobi: pressure_solver/jacobi.py:38-78,160-203 --------------------- */
/* n_iter iterations; p is updated in place (tmp is a same-shape scratch array). p[0,0]:=0. */
int nf_jacobi_iterate(nf_ctx*, const nf_grid*, double* p, double* tmp, const double* b, const double* d_u,
                      const double* d_v, double omega, int n_iter);
int nf_jacobi_diag(nf_ctx*, const nf_grid*, const double* d_u, const double* d_v, double* diag);

/* ---- K8  red-black SOR: pressure_solver/gauss_seidel.py:214-305 ----------------------- */
int nf_rbsor_sweeps(nf_ctx*, const nf_grid*, double* p, const double* b, const double* d_u,
                    const double* d_v, double omega, int n_sweeps);
/* GaussSeidelSolver method_type 'standard' / 'symmetric' (gauss_seidel.py:307-367): sequential SOR sweeps `for j: for i:`
 * (symmetric: followed by the reverse sweep), evaluated as a two-level anti-diagonal wavefront (32 x 32 blocks, one launch
 * per block diagonal) -- same bits as the loop.  Single-slab grids. */
int nf_gs_lex_sweeps(nf_ctx*, const nf_grid*, double* p, const double* b, const double* d_u, const double* d_v,
                     double omega, int n_sweeps, int symmetric);

/* Same sweeps, temporally blocked: up to 3 full sweeps (6 colour passes) per tile load, bit-identical result.
 * tmp is a same-shape scratch array (p is double buffered between launches); arrays 16-byte aligned, ld even. */
int nf_rbsor_sweeps_fused(nf_ctx*, const nf_grid*, double* p, double* tmp, const double* b, const double* d_u,
                          const double* d_v, const double* inv /* nf_pressure_inv_diag output or NULL */,
                          double omega, int n_sweeps);
/* inv[i,j] = 1/aP of the SOR update (gauss_seidel.py:214-266), reusable for every sweep with the same d_u, d_v */
int nf_pressure_inv_diag(nf_ctx*, const nf_grid*, const double* d_u, const double* d_v, double* inv);

/* ---- K9-K12 transfer operators: pressure_solver/helpers/multigrid_helpers.py ---------- */
int nf_restrict_fw(nf_ctx*, const nf_grid* fine, const double* f, const nf_grid* coarse, double* c);     /* :23-70  */
int nf_restrict_inject(nf_ctx*, const nf_grid* fine, const double* f, const nf_grid* coarse, double* c); /* :8-21   */
int nf_restrict_coeffs(nf_ctx*, const nf_grid* fine, const double* d_u, const double* d_v,
                       const nf_grid* coarse, double* d_u_c, double* d_v_c);                             /* :196-329 */
/* fine = P*coarse (add=0) or fine += P*coarse (add=1); bilinear with the reference's index rules */
int nf_prolong_linear(nf_ctx*, const nf_grid* coarse, const double* c, const nf_grid* fine, double* f,
                      int add);                                                                          /* :73-192 */
/* interpolate_cubic (:333-391): separable not-a-knot spline on linspace(0,1,.) coordinates, square grids.
 * Builds the banded 1-D operator on the host on every call (setup cost; the multigrid driver caches it) and keeps it, with
 * the row-interpolated intermediate array, in the caller's device workspace: 256-byte aligned, at least
 * nf_workspace_bytes(NF_WS_PROLONG_CUBIC, fine nx, fine ny, coarse nx, coarse ld) bytes.  No allocation, no synchronisation. */
int nf_prolong_cubic(nf_ctx*, const nf_grid* coarse, const double* c, const nf_grid* fine, double* f,
                     int add, void* workspace, size_t workspace_bytes);
/* device scratch a stand-alone entry point needs (the *_create objects own theirs) */
enum nf_workspace_kind { NF_WS_PROLONG_CUBIC = 1 };
size_t nf_workspace_bytes(int which, int nx, int ny, int nxc, int ldc);

/* ---- reductions (K18) ------------------------------------------------------------------ */
/* sqrt(sum x^2) over the nx*ny cells (rows [gb,ge)); interior_only!=0 masks the boundary ring */
int nf_norm2(nf_ctx*, const nf_grid*, const double* x, int interior_only, double* out_host);
int nf_dot(nf_ctx*, const nf_grid*, const double* x, const double* y, double* out_host);

/* ---- K13 + multigrid driver: pressure_solver/multigrid.py:121-688 ---------------------- */
typedef struct nf_mg_config {
  int32_t smoother;        /* 0 = red-black SOR (gauss_seidel.py), 1 = weighted Jacobi (jacobi.py),
                              2 / 3 = lexicographic / symmetric SOR (gauss_seidel.py:307-367; single slab) */
  int32_t pre, post;       /* pre_smoothing, post_smoothing                                       */
  int32_t cycle_type;      /* 0 'v', 1 'w', 2 'fmg'                                               */
  int32_t cycle_buildup;   /* 0 'v', 1 'w'                                                        */
  int32_t cycle_final;     /* -1 None, 0 'v', 1 'w'                                               */
  int32_t max_cycles_buildup;
  int32_t restriction;     /* 0 full weighting, 1 injection                                       */
  int32_t interpolation;   /* 0 linear, 1 cubic (not-a-knot spline)                               */
  int32_t coarsest;        /* coarsest_grid_size                                                  */
  int32_t max_iterations;
  int32_t pad;
  double omega;
  double tolerance;
  double length, height;   /* mesh.length, mesh.height (coarse meshes: multigrid.py:373)          */
  double rho;
} nf_mg_config;

typedef struct nf_mg_info {
  double r_norm;     /* ||b - A x||_2 after the last cycle (multigrid.py:257 'rel_norm') */
  double b_norm;
  int32_t cycles;    /* V/W cycles run at the finest level ('v'/'w' mode)                 */
  int32_t levels;
} nf_mg_info;

int nf_mg_create(nf_ctx*, nf_mg** out, int nx, int ny, int ld, const nf_mg_config* cfg);
int nf_mg_destroy(nf_mg*);
int nf_mg_num_levels(nf_mg*);
int nf_mg_level_shape(nf_mg*, int level, int* nx, int* ny, int* ld);
/* device array of a level (tests / inspection): which = 0 d_u, 1 d_v, 2 x, 3 b, 4 r */
const double* nf_mg_level_array(nf_mg*, int level, int which);
/* hierarchy of restricted coefficients + coarse inverse for this (d_u, d_v) (multigrid.py:380-385) */
int nf_mg_setup(nf_mg*, const double* d_u, const double* d_v);
/* MultiGridSolver.solve without get_rhs: x (out), b (in), r (out, residual field b-Ax) */
int nf_mg_solve(nf_mg*, const double* b, double* x, double* r, nf_mg_info* info_host);
/* one cycle (kind 0 'v' / 1 'w') at the finest level on (x, b): multigrid.py:304-560 */
int nf_mg_cycle(nf_mg*, double* x, const double* b, int kind);

/* ---- K14/K15 Krylov solvers in scipy's operation order (scipy _isolve/iterative.py) ------ */
typedef struct nf_krylov_info {
  double r_norm;      /* ||r|| of the recurrence at exit                   */
  double b_norm;
  int32_t iterations;
  int32_t info;       /* scipy's info: 0 converged, >0 maxiter, <0 breakdown */
} nf_krylov_info;
/* work: 4 (cg) / 7 (bicgstab) same-shape scratch arrays, contiguous, each nx*ld doubles */
int nf_cg_solve(nf_ctx*, const nf_grid*, const double* b, double* x, const double* d_u, const double* d_v,
                double atol, double rtol, int maxiter, int check_every, double* work, nf_krylov_info* info_host);
int nf_bicgstab_solve(nf_ctx*, const nf_grid*, const double* b, double* x, const double* d_u,
                      const double* d_v, double atol, double rtol, int maxiter, int check_every, double* work,
                      nf_krylov_info* info_host);

/* ---- K2-K4 momentum: discretization/power_law.py:46-365, jacobi_matrix_solver.py:153-375 -- */
typedef struct nf_links {  /* six same-shape coefficient arrays of one momentum component */
  double *a_e, *a_w, *a_n, *a_s, *a_p, *src;
} nf_links;
/* u_bc/v_bc: velocities with BCs applied; writes relaxed a_p (=a_p/alpha), relaxed source and d
 * (= dy/a_p or dx/a_p, NaN where |a_p|<=1e-12).  sides bit mask: 1 left, 2 right, 4 bottom, 8 top
 * (boundaries with a registered condition -> Practice-B folding). */
int nf_momentum_links_u(nf_ctx*, const nf_grid*, const double* u_bc, const double* v_bc, const double* p,
                        double mu, double alpha, int sides, nf_links out, double* d_u);
int nf_momentum_links_v(nf_ctx*, const nf_grid*, const double* u_bc, const double* v_bc, const double* p,
                        double mu, double alpha, int sides, nf_links out, double* d_v);
/* n_sweeps Jacobi sweeps x <- D^-1 (b - (A-D) x); is_u selects the (nx+1,ny) / (nx,ny+1) shape.
 * Result ends in x (tmp is scratch). */
int nf_momentum_jacobi(nf_ctx*, const nf_grid*, int is_u, nf_links L, double* x, double* tmp, int n_sweeps);
/* Same sweeps, temporally blocked (up to 6 per tile load, bit-identical); with rel_norm_host != NULL the relaxed
 * residual norm (and field, may be NULL) of the result is evaluated behind the last sweep. */
int nf_momentum_jacobi_fused(nf_ctx*, const nf_grid*, int is_u, nf_links L, double* x, double* tmp, int n_sweeps,
                             double* field_out, double* rel_norm_host);
/* r = b - A x (field_out, with the reference's boundary zeroing) and ||r_masked||/(||b_masked||+1e-15) */
int nf_momentum_residual(nf_ctx*, const nf_grid*, int is_u, nf_links L, const double* x, double* field_out,
                         double* rel_norm_host);

/* ---- higher-order convection schemes: discretization/quick.py:27-219 (QUICKDiscretization),
 *      discretization/second_order_upwind.py:26-325 (SecondOrderUpwindDiscretization) ----
 * Ten same-shape coefficient arrays of one component: the 5-point links, the second-neighbour links a_ee / a_ww / a_nn /
 * a_ss, a_p and the source (pressure gradient + Practice-B boundary terms), exactly what calculate_u_coefficients /
 * calculate_v_coefficients return.  No relaxation (the reference applies none here).  Single slab. */
typedef struct nf_links_ext {
  double *a_e, *a_w, *a_n, *a_s, *a_ee, *a_ww, *a_nn, *a_ss, *a_p, *src;
} nf_links_ext;
enum nf_scheme { NF_SCHEME_QUICK = 1, NF_SCHEME_SOU = 2 };
int nf_momentum_links_ext(nf_ctx*, const nf_grid*, int is_u, int scheme, const double* u_bc, const double* v_bc,
                          const double* p, double mu, int sides, nf_links_ext out);

/* ---- a7 MatrixFreeMomentumSolver (momentum_solver/matrix_free_momentum.py:403-544): Krylov momentum predictor ----
 * Coefficients with that class's relaxation (a_P clamped to 1e-12 then /alpha, source relaxed with the relaxed a_P,
 * d = 0 where a_P vanishes); u_bc/v_bc carry the BCs as that class applies them (caller's nx+1).  The unrelaxed a_P and
 * source go to ap_unrelaxed / src_unrelaxed for nf_momentum_residual_unrelaxed. */
int nf_momentum_links_mf(nf_ctx*, const nf_grid*, int is_u, const double* u_bc, const double* v_bc, const double* p,
                         double mu, double alpha, int sides, nf_links out, double* d, double* ap_unrelaxed,
                         double* src_unrelaxed);
/* scipy bicgstab (operation order as nf_bicgstab_solve) on the relaxed system, interior rows 5-point / boundary rows
 * identity (:49-79), x0 = x on entry, stop at ||r|| < max(atol, rtol ||b||); unpreconditioned (the reference's ILU only
 * changes the iteration path).  work: 5 arrays of (nx+1)*ld doubles. */
int nf_momentum_bicgstab(nf_ctx*, const nf_grid*, int is_u, nf_links L, double* x, double atol, double rtol, int maxiter,
                         int check_every, double* work, nf_krylov_info* info);
/* residual of the unrelaxed system (:379-400): L_unrelaxed = links with a_p, src = the unrelaxed arrays; boundary and
 * boundary-adjacent lines zeroed; *norm_host = ||r|| (that class's "rel_norm") */
int nf_momentum_residual_unrelaxed(nf_ctx*, const nf_grid*, int is_u, nf_links L_unrelaxed, const double* x,
                                   double* field_out, double* norm_host);

/* ---- K16/K17 corrections: velocity_solver/standard.py:10-69, Algorithms/simple.py:148-150,
 *      Algorithms/base_algorithm.py:161-197 ------------------------------------------------- */
int nf_correct_velocity(nf_ctx*, const nf_grid*, const nf_bc_program*, const double* u_star,
                        const double* v_star, const double* p_prime, const double* d_u, const double* d_v,
                        double* u, double* v);
int nf_update_pressure(nf_ctx*, const nf_grid*, const double* p_star, const double* p_prime, double alpha_p,
                       double* p);
int nf_max_abs_divergence(nf_ctx*, const nf_grid*, const double* u, const double* v, double* out_host);

/* ---- device-resident SIMPLE / PISO / SIMPLER outer loops: Algorithms/simple.py:78-269, piso.py:41-175,
 *      simpler.py:78-262 ---- */
typedef struct nf_simple_config {
  int32_t nx, ny;
  int32_t n_momentum_sweeps;    /* JacobiMatrixMomentumSolver(n_jacobi_sweeps)                          */
  int32_t pressure_solver;      /* 0 multigrid, 1 Jacobi, 2 red-black SOR, 3 CG, 4 BiCGSTAB,
                                   5 lexicographic SOR, 6 symmetric SOR, 7 BiCGSTAB with the multigrid
                                   preconditioner (5, 6, 7: single slab)                                 */
  int32_t pressure_iterations;  /* fixed iteration count of the Jacobi / SOR pressure solvers          */
  int32_t sides;                /* boundaries with a registered condition: 1 left 2 right 4 bottom 8 top */
  int32_t krylov_maxiter;
  int32_t piso_corrections;     /* 0: SIMPLE (simple.py:114-212); n >= 1: PISO with n pressure corrections per outer
                                   iteration, momentum re-solved without relaxation in between (piso.py:73-104);
                                   -1: SIMPLER as coded in simpler.py:99-167 (p += p-bar unrelaxed, momentum again,
                                   p += alpha_p p', velocity correction; p_rel_norm = ||p - p_old|| / sqrt(nx ny));
                                   -2: SIMPLEC as coded in simplec.py:99-171 (d / simplec_divisor, 5-point smoothing of
                                   p', p += alpha_p p' without edge copies; the record holds infinity norms: u_rel_norm =
                                   v_rel_norm = max|u - u_old|, |v - v_old|; u_abs_res = max|u* - u|, |v* - v|;
                                   p_rel_norm = max|p - p_old|; single slab)                                        */
  double length, height, rho, mu;
  double alpha_p, alpha_u;      /* simple.py:23-76                                                      */
  double pressure_omega;        /* Jacobi / SOR relaxation                                              */
  double pressure_tolerance;    /* Krylov atol (matrix_free_BiCGSTAB.py:234-242)                        */
  nf_bc_program bc;             /* velocity BC program evaluated with the true (nx, ny)                 */
  nf_mg_config mg;
  int32_t momentum_solver;      /* 0: fixed Jacobi sweeps (JacobiMatrixMomentumSolver, a6); 1: BiCGSTAB on the relaxed
                                   system (MatrixFreeMomentumSolver, a7; single slab only)                             */
  int32_t momentum_maxiter;     /* a7: max_iterations (matrix_free_momentum.py:17)                                     */
  double momentum_tolerance;    /* a7: atol of the Krylov solve (:16); stop at max(atol, 1e-5 ||b||)                   */
  nf_bc_program bc_mf;          /* a7: BC program as that class applies it (caller's nx+1: :419, :491)                 */
  double simplec_divisor;       /* SIMPLEC (piso_corrections == -2): d_u, d_v are divided by 1 - (1 - alpha_u), evaluated by
                                   the caller exactly as simplec.py:126-127 does                                       */
  int32_t krylov_check_every;   /* CG / BiCGSTAB pressure solves: the host polls the device-side stopping flag every n
                                   iterations (0: 25 for CG, 10 for BiCGSTAB); the answer does not depend on it        */
  int32_t krylov_mg_cycles;     /* pressure_solver 7: multigrid cycles per preconditioner application                  */
  int32_t krylov_mg_kind;       /* pressure_solver 7: 0 'v', 1 'w', 2 'fmg' (matrix_free_BiCGSTAB.py:102-161); the
                                   preconditioner's hierarchy is described by `mg`                                     */
  int32_t track_unrelaxed_residual; /* 1: every record also carries the absolute UNRELAXED momentum residual norms of the
                                   predicted velocities (matrix_free_momentum.py:380-400 masks) -- one extra pass over the
                                   links per component; the convergence measure of the outer loop for the Jacobi-sweep
                                   predictor, whose own rel_norm is the relaxed inner residual           */
} nf_simple_config;

typedef struct nf_simple_info {   /* one record per outer iteration */
  double u_rel_norm, v_rel_norm;  /* momentum solver's rel_norm: relaxed ||r||/||b|| (jacobi_matrix_solver.py:246-250)
                                     or, momentum_solver 1, the absolute unrelaxed ||r|| (matrix_free_momentum.py:455) */
  double p_rel_norm;              /* pressure solver rel_norm (its own convention)           */
  double u_abs_res, v_abs_res;    /* sqrt(sum r^2) of the relaxed momentum residual over the interior   */
  int32_t pressure_iterations;    /* multigrid cycles / Krylov iterations used                          */
  int32_t pad;
  double u_unrelaxed_res, v_unrelaxed_res; /* cfg.track_unrelaxed_residual: ||S_un - A_un u*|| over the interior, else 0 */
} nf_simple_info;

int nf_simple_create(nf_ctx*, nf_simple** out, const nf_simple_config* cfg);

/* ---- row-slab decomposition over the GPUs of one box (no counterpart in the reference, which is a single
 *      process: SURVEY.md section 5).  One rank per GPU: nf_nccl_unique_id on rank 0, broadcast the 128 bytes
 *      (torch.distributed), nf_team_create_nccl on every rank; halos travel by ncclSend/ncclRecv over NVLink,
 *      norms by ncclAllReduce.  nf_team_create_virtual cuts the grid into ranks that all live in this process
 *      on this device (same code path, cudaMemcpy halos): the parity tests use it on a single GPU. */
int nf_nccl_unique_id(nf_ctx*, void* id_out_128_bytes);
int nf_team_create_nccl(nf_ctx*, int world, int rank, const void* id_128_bytes, nf_team** out);
int nf_team_create_virtual(nf_ctx*, int virtual_ranks, nf_team** out);
/* COLLECTIVE for teams created by nf_team_create_nccl: every rank must call it (peer mappings are closed and all ranks
 * meet before any rank frees the memory its peers had mapped) */
int nf_team_free(nf_team*);
/* 1 when the team's halo exchanges / norm reductions run as the ranks' own kernels over NVLink peer memory (cudaIpc
 * arena, default for nf_team_create_nccl with world <= 8; NF_P2P=0 or a failed cudaIpc set-up selects NCCL) */
int nf_team_uses_p2p(nf_team*);
/* collective micro-benchmark of the team's transport: mean time of `reps` halo exchanges of `depth` rows of an
 * nx x ny level and of `reps` all-reduces of n_scalars doubles (CUDA events; tools/bench_exchange.py) */
int nf_team_benchmark(nf_team*, int nx, int ny, int depth, int n_scalars, int reps, double* ms_exchange,
                      double* ms_allreduce);
/* host-only partition queries: cell rows [begin, end) of `rank` when nx rows are cut over `world` ranks (boundaries
 * on multiples of 16; the grid is not cut when a slab would have fewer than 64 rows), and the rows of the
 * next-coarser level a rank restricts into (coarse row I belongs to the owner of fine row 2I+1) */
int nf_slab_rows(int nx, int world, int rank, int* row_begin, int* row_end);
int nf_slab_coarse_rows(int fine_begin, int next_fine_begin, int is_last, int nxc, int* row_begin, int* row_end);
/* SIMPLE state cut over the team (multigrid pressure solver, bilinear prolongation); the team outlives it */
int nf_simple_create_team(nf_team*, nf_simple** out, const nf_simple_config* cfg);
/* live CUDA-event timing of the finest-level fused smoother launches (3 sweeps each) inside nf_simple_iterate /
 * nf_mg_solve: switches the instrumentation on / off and returns + resets the accumulated time and launch count */
int nf_mg_smoother_timing(nf_mg*, int on, double* total_ms, long long* launches);
int nf_simple_smoother_timing(nf_simple*, int on, double* total_ms, long long* launches);
/* CUDA-event timing of the phases of the outer iteration (momentum predictor / pressure solve incl. its RHS / p and
 * velocity corrections): switches it on / off, returns + resets the accumulated ms and the iterations they cover */
int nf_simple_phase_timing(nf_simple*, int on, double* ms_momentum, double* ms_pressure, double* ms_correct,
                           long long* iterations);
/* cell rows [*row_begin, *row_end) owned by local slab k of this process (k = 0 under torchrun) */
int nf_simple_local_rows(nf_simple*, int k, int* row_begin, int* row_end);
int nf_simple_destroy(nf_simple*);
int nf_simple_ld(nf_simple*);
/* device arrays (row pitch nf_simple_ld): which = 0 u, 1 v, 2 p, 3 u_star, 4 v_star, 5 d_u, 6 d_v,
 * 7 p_prime, 8 b, 9 pressure residual field, 10 u residual field, 11 v residual field */
double* nf_simple_field(nf_simple*, int which);
/* host <-> device copies of a field; host arrays are the FULL C-contiguous (rows, cols) fields: every local slab
 * takes its rows (halo included) on upload and writes only the rows it owns on download */
int nf_simple_upload(nf_simple*, int which, const double* host, int rows, int cols);
int nf_simple_download(nf_simple*, int which, double* host, int rows, int cols);
/* runs outer iterations until max(u_rel_norm, v_rel_norm) <= tolerance or n_iterations are done
 * (simple.py:114); writes one nf_simple_info per iteration, returns the count in *n_done.
 * tolerance <= 0: no host synchronisation inside the loop. want_fields != 0 also stores the momentum
 * residual fields. */
int nf_simple_iterate(nf_simple*, int n_iterations, double tolerance, int want_fields, nf_simple_info* info_host,
                      int* n_done);

/* BiCGSTAB with the multigrid preconditioner of matrix_free_BiCGSTAB.py:102-161: M z = mg_cycles cycles (mg_kind
 * 0 'v', 1 'w', 2 'fmg') on A y = z from y = 0.  `mg` must have been set up (nf_mg_setup) with the same d_u, d_v;
 * work: 7 same-shape scratch arrays. */
int nf_bicgstab_solve_mg(nf_ctx*, const nf_grid*, const double* b, double* x, const double* d_u, const double* d_v,
                         double atol, double rtol, int maxiter, int check_every, double* work, nf_mg* mg, int mg_cycles,
                         int mg_kind, nf_krylov_info* info_host);

/* CG with the multigrid preconditioner of GeoMultigridPrecondCGSolver (pressure_solver/geo_multigrid_cg.py:125-191): scipy's
 * cg with M = mg_cycles cycles from zero.  `mg` must have been set up with the same d_u, d_v; work: 5 same-shape arrays. */
int nf_cg_solve_mg(nf_ctx*, const nf_grid*, const double* b, double* x, const double* d_u, const double* d_v, double atol,
                   double rtol, int maxiter, double* work, nf_mg* mg, int mg_cycles, int mg_kind, nf_krylov_info* info_host);

#ifdef __cplusplus
}
#endif
#endif /* NAVIFLOW_B200_H */
