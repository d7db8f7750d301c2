// nf_common.cuh -- shared device/host helpers for libnaviflow_b200 (sm_100a, fp64).
// Compiled with -fmad=false: every expression keeps the reference's (NumPy) operation order and
// rounding, so elementwise kernels are bit-identical to the reference's CPU arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/naviflow_b200.h"

#define NF_SM_COUNT 148  // B200: 2 dies x 74 SMs

struct nf_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  // deterministic two-stage reductions: per-block partials + ticket counter + result slots
  double* partials = nullptr;   // NF_MAX_PARTIALS * NF_MAX_RED doubles
  unsigned int* ticket = nullptr;
  double* scalars = nullptr;    // device scalars (64 doubles)
  double* scalars_host = nullptr;  // pinned mirror
  int64_t launches = 0;
  std::string err;
};

#define NF_MAX_PARTIALS 16384
#define NF_MAX_RED 4
#define NF_NUM_SCALARS 64

#define NF_CHECK_CUDA(ctx, expr)                                                        \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                  \
      return NF_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define NF_REQUIRE(ctx, cond, msg)                                                      \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      (ctx)->err = std::string("argument error: ") + (msg);                             \
      return NF_ERR_ARG;                                                                \
    }                                                                                   \
  } while (0)

#define NF_LAUNCH_CHECK(ctx)                                                            \
  do {                                                                                  \
    (ctx)->launches++;                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(_e);             \
      return NF_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// The kernels of a multigrid cycle are short (a few microseconds on the coarse levels) and strictly ordered, so the gap
// between two of them -- drain, flush, grid launch, parameter fetch, CTA dispatch: ~2-3 us -- is a sizeable share of the
// cycle.  A kernel that begins with nf_pdl_entry() may be launched with nf_launch(..., pdl = true): its CTAs are
// dispatched while the previous kernel of the stream still runs (that one has released them with
// griddepcontrol.launch_dependents) and block in griddepcontrol.wait until the previous grid has completed and its
// writes are visible.  Nothing before the wait touches global memory, every thread executes it, and a kernel launched
// without the attribute (or after a copy / memset) keeps the full dependency, so the order of all memory effects is
// unchanged.  Stream capture turns the attribute into a programmatic edge of the CUDA graph.  NF_PDL=0 switches it off.
__device__ __forceinline__ void nf_pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef NF_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
bool nfi_pdl_enabled();

template <class... KArgs, class... Args>
inline cudaError_t nf_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && nfi_pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// element [i][j] of a field stored from global row g.row0 with pitch g.ld
__device__ __forceinline__ size_t nf_idx(const nf_grid& g, int i, int j) {
  return (size_t)(i - g.row0) * (size_t)g.ld + (size_t)j;
}

// Rows a descriptor's arrays hold: [g.row0, nf_stored_end(g)) (row1 == 0 means the whole grid, nx+1 rows).  Tiled
// kernels whose last tile overshoots the computed range must not load rows outside this window: on a slab the rows
// beyond it are not allocated.
__device__ __host__ __forceinline__ int nf_stored_end(const nf_grid& g) { return g.row1 > 0 ? g.row1 : g.nx + 1; }
__device__ __host__ __forceinline__ bool nf_row_stored(const nf_grid& g, int i) {
  return i >= g.row0 && i < nf_stored_end(g);
}

// 2-D launch geometry: x along j (contiguous), y along rows
struct NfLaunch2D {
  dim3 grid, block;
};
static inline NfLaunch2D nf_launch2d(int rows, int cols, int bx = 128, int by = 2) {
  NfLaunch2D l;
  l.block = dim3(bx, by, 1);
  l.grid = dim3((cols + bx - 1) / bx, (rows + by - 1) / by, 1);
  if (l.grid.x == 0) l.grid.x = 1;
  if (l.grid.y == 0) l.grid.y = 1;
  return l;
}

// ---------------------------------------------------------------------------------------------
// Deterministic grid reduction of up to NF_MAX_RED sums.
// Each block reduces its threads' values in a fixed tree (warp shuffles, then warp 0), writes one
// partial per sum, takes a ticket; the last block sums the partials in index order and stores the
// results to out[0..NRED).  Same launch geometry => bit-identical result run to run.
// ---------------------------------------------------------------------------------------------
template <int NRED>
__device__ __forceinline__ bool nf_block_reduce_store(double (&val)[NRED], double* partials,
                                                      unsigned int* ticket, double* out) {
  __shared__ double s_warp[NRED][32];
  __shared__ bool s_last;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthreads = blockDim.x * blockDim.y;
  const int lane = tid & 31, warp = tid >> 5;
  const int nwarps = (nthreads + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    double v = val[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) s_warp[k][warp] = v;
  }
  __syncthreads();
  const int nblocks = gridDim.x * gridDim.y;
  const int bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
      double v = (lane < nwarps) ? s_warp[k][lane] : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0) partials[(size_t)k * NF_MAX_PARTIALS + bid] = v;
    }
    if (lane == 0) {
      __threadfence();
      unsigned int t = atomicAdd(ticket, 1u);
      s_last = (t == (unsigned int)(nblocks - 1));
    }
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    // fixed-order final sum: thread t accumulates partials t, t+T, ...; then the same tree
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
      double v = 0.0;
      for (int b = tid; b < nblocks; b += nthreads)
        v += ((volatile double*)partials)[(size_t)k * NF_MAX_PARTIALS + b];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      __syncthreads();
      if (lane == 0) s_warp[k][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < NRED; ++k) {
        double v = (lane < nwarps) ? s_warp[k][lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) out[k] = v;
      }
      if (lane == 0) *ticket = 0u;
    }
  }
  // true on exactly one thread: thread 0 of the last block, after out[] holds the final sums
  return s_last && tid == 0;
}

// grid for reduction kernels: bounded number of blocks (<= NF_MAX_PARTIALS), grid-stride rows
static inline NfLaunch2D nf_launch_reduce(int rows, int cols) {
  NfLaunch2D l;
  l.block = dim3(128, 2, 1);
  int gx = (cols + 127) / 128;
  if (gx < 1) gx = 1;
  int gy = (rows + 1) / 2;
  if (gy < 1) gy = 1;
  int max_gy = NF_MAX_PARTIALS / gx;
  if (max_gy < 1) max_gy = 1;
  int target = (NF_SM_COUNT * 32 + gx - 1) / gx;  // ~32 blocks per SM: short per-thread row loops, 4 waves
  if (target < 1) target = 1;
  if (gy > target) gy = target;
  if (gy > max_gy) gy = max_gy;
  l.grid = dim3(gx, gy, 1);
  return l;
}

int nf_read_scalars(nf_ctx* ctx, int first, int count, double* out_host);

// ---- internal launchers (no argument validation; used by the multigrid / Krylov / SIMPLE drivers) ----
int nf_check_grid(nf_ctx* ctx, const nf_grid* g);
#define NF_GRID_OK(ctx, g)            \
  do {                                \
    int _s = nf_check_grid(ctx, g);   \
    if (_s != NF_OK) return _s;       \
  } while (0)
#define NF_TRY(expr)                  \
  do {                                \
    int _s = (expr);                  \
    if (_s != NF_OK) return _s;       \
  } while (0)

int nfi_rbsor(nf_ctx*, const nf_grid*, double* p, const double* b, const double* d_u, const double* d_v,
              double omega, int n_sweeps);
int nfi_jacobi(nf_ctx*, const nf_grid*, double* p, double* tmp, const double* b, const double* d_u,
               const double* d_v, double omega, int n_iter);
int nfi_residual(nf_ctx*, const nf_grid*, const double* p, const double* b, const double* d_u, const double* d_v,
                 double* r);
int nfi_apply(nf_ctx*, const nf_grid*, const double* p, const double* d_u, const double* d_v, double* out);
// sum of squares of x over the cells of g -> device scalar ctx->scalars[slot] (asynchronous)
int nfi_sumsq_dev(nf_ctx*, const nf_grid*, const double* x, int interior_only, int slot);
int nfi_residual_norms(nf_ctx*, const nf_grid*, const double* p, const double* b, const double* d_u, const double* d_v,
                       double* r, int with_b, double* out);
int nfi_sumsq_to(nf_ctx*, const nf_grid*, const double* x, int interior_only, double* out);
int nfi_fill(nf_ctx*, double* x, size_t count, double value);
int nfi_restrict_fw(nf_ctx*, const nf_grid* fine, const double* f, const nf_grid* coarse, double* c);
int nfi_restrict_inject(nf_ctx*, const nf_grid* fine, const double* f, const nf_grid* coarse, double* c);
int nfi_restrict_coeffs(nf_ctx*, const nf_grid* fine, const double* d_u, const double* d_v, const nf_grid* coarse,
                        double* d_u_c, double* d_v_c);
int nfi_prolong_linear(nf_ctx*, const nf_grid* coarse, const double* c, const nf_grid* fine, double* f, int add);

// Work fused behind the last launch of a smoothing call (persistent TMA kernel only; nf_rbsor_fused.cu):
//   mode 1: sum (b - A p)^2 -> out[0], sum b^2 -> out[1];  mode 2: coarse_b = FW(b - A p) on the coarse grid gc and,
//   with in_norm_out, the residual norms of the launch's INPUT iterate -> in_norm_out[0..1] (lookahead convergence test).
// *fused / *in_norm_fused tell the caller whether the work was done (otherwise it runs the stand-alone kernels).
struct nf_smooth_extra {
  int mode = 0;
  nf_grid gc;
  double* coarse_b = nullptr;
  double* coarse_x_zero = nullptr;  // mode 2, optional: coarse iterate to be zeroed cell by cell along with the restriction
                                    // (the cycle recurses from a zero guess: saves the fill launch on unsplit levels)
  double* out = nullptr;
  const double* prolong_c = nullptr;  // modes 0 / 1, optional (streaming kernel only): the launch smooths p + P(prolong_c),
  nf_grid prolong_gc;                 // the bilinear prolongation of the coarse iterate on prolong_gc (saves the prolongation pass)
  bool prolong_fused = false;
  bool fused = false;
  double* in_norm_out = nullptr;
  bool in_norm_fused = false;
};
// n_sweeps red-black SOR sweeps, double buffered (*p input / result, *palt scratch; swapped once per launch);
// inv (optional): precomputed 1/aP of the level
int nfi_rbsor_fused_x(nf_ctx*, const nf_grid*, double** p, double** palt, const double* b, const double* d_u,
                      const double* d_v, const double* inv, double omega, int n_sweeps, nf_smooth_extra* extra);

// streaming (wavefront) variant of the temporally blocked smoother (nf_rbsor_stream.cu): ns in 1..3 sweeps from pin into
// pout; needs the precomputed 1/aP and an even first row.  mode / extra as nf_smooth_extra (fused work on 3-sweep launches
// only, no in_norm).  *used = false: nothing was launched, take another path.
int nfi_rbsor_stream(nf_ctx*, const nf_grid*, const double* pin, double* pout, const double* b, const double* d_u,
                     const double* d_v, const double* inv, double omega, int ns, int mode, const nf_smooth_extra* extra,
                     bool* used);
bool nfi_rbsor_stream_enabled(const nf_grid* g);
int nfi_rbsor_can_fuse_prolong(const nf_grid* g, int n_sweeps, bool has_inv);
void nfi_prolong_block_extent(const nf_grid* gc, const nf_grid* gf, int* nI, int* nJ);
