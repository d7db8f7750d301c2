// nf_rbsor_stream.cu -- streaming (wavefront) form of the temporally blocked red-black SOR smoother (K8).
//
// Same arithmetic as gauss_seidel.py:214-305 / k_rbsor_color and bit-identical to it.  Where k_rbsor_tma blocks the
// 2*NS colour passes over a 48 x 64 region (region/tile redundancy 1.64 - 2.0, one barrier per pass), this kernel
// pipelines them along the row index i:
//
//   * A WARP is the unit of work.  It owns a strip of 64 columns (lane l: the column pair j0+2l, j0+2l+1) and marches
//     down a chunk of rows.  In step s it receives row s and applies colour pass t to row s-1-t, t = 0 .. 2NS-1, in that
//     order: pass t of row r needs rows r-1 and r+1 after pass t-1 and must see them before pass t+1 -- exactly what the
//     order gives (row r+1 got pass t-1 a moment ago in the same step, row r-1 got it two steps ago and receives pass
//     t+1 later in this step).  After step s row s-2NS is final and is stored.  Halo: 2NS columns on each side of the
//     strip (trapezoid in j only) and 2NS rows at both ends of the chunk -- redundancy 64/52 x (len+12)/len instead of
//     (48 x 64)/(36 x 52).
//   * All operands of the passes are REGISTERS of the same lane: the window of 2NS+2 rows of p, the link coefficients
//     as FACE values (the face between rows r-1 and r is aW of row r and aE of row r-1; likewise in j: 2+3 instead of 8
//     doubles per row pair), b and 1/aP.  The only exchange is one warp shuffle per update for the neighbour across the
//     pair boundary.  In one step all 2NS updates hit the same cell of the pair ((i+j) parity), so the shuffles of a step
//     are issued up front and the step is one dependent chain of 2NS updates; no shared-memory traffic, no barrier.
//   * Rows arrive through a per-warp ring of TMA boxes (cp.async.bulk.tensor.2d, 1 row x 64 columns per array, out-of-
//     domain elements zero-filled), armed on per-warp mbarriers NSTG rows ahead of their use.
//   * The window rotates through 8 register slots; the step function is instantiated for the 8 phases so every slot
//     index is a compile-time constant.
//
// Domain edges (rows 0 / nx-1, columns 0 / ny-1, the pinned cell, cells outside the domain) break the face sharing
// (the reference zeroes the link of a boundary cell towards the interior but not the reverse link,
// matrix_free.py:63-84) and are handled by a generic variant of the step (BND) that only the jobs touching an edge run.
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include <cuda.h>

#include "nf_pressure.cuh"

namespace {

constexpr int SW = 64;    // strip width in cells
constexpr int NSTG = 4;   // rows in flight per warp
constexpr int SG_P = 0, SG_DU = 512, SG_B = 1024, SG_INV = 1536, SG_DV = 2048;  // stage layout (bytes)
constexpr int SG_BYTES = 2048 + 640;                                            // d_v row: 66 doubles, padded
constexpr unsigned SG_TX = 4 * 512 + 66 * 8;

__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_row(unsigned dst, const CUtensorMap* map, int c_inner, int c_outer, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      :: "r"(dst), "l"(map), "r"(c_inner), "r"(c_outer), "r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

struct Win {  // register window of a lane, 8 rotating row slots
  double pA[8], pB[8];          // p of the column pair
  double fA[8], fB[8];          // face above the row: rho*d_u[r][j]*dy  (aW of row r, aE of row r-1)
  double g0[8], g1[8], g2[8];   // faces of the row in j: rho*d_v[r][j]*dx at j = jA, jA+1, jA+2
  double bA[8], bB[8], iA[8], iB[8];
};

struct Job {
  int ia, ib;      // rows this job finalises
  int r0;          // first row it loads (even)
  int s_last;      // last step
  int j0;          // global column of lane 0's first cell (even, may be negative)
};

// One step of the wavefront.  PH = (s - r0) & 7 (compile time), NP = number of colour passes.
template <int PH, int NP, bool BND>
__device__ __forceinline__ void stream_step(Win& W, const nf_grid& g, const Job& jb, int s, int lane, double omega,
                                            const unsigned char* stage, double* __restrict__ pout) {
  constexpr int SN = PH;              // slot of row s
  constexpr int SC = (PH + 7) & 7;    // slot of row s-1 (its coefficients arrive now)
  const int gjA = jb.j0 + 2 * lane;
  // ---- the new row: p[s], d_u[s] and the coefficients of row s-1 ----
  {
    const double2 pp = *reinterpret_cast<const double2*>(stage + SG_P + 16 * lane);
    const double2 uu = *reinterpret_cast<const double2*>(stage + SG_DU + 16 * lane);
    const double2 bb = *reinterpret_cast<const double2*>(stage + SG_B + 16 * lane);
    const double2 iv = *reinterpret_cast<const double2*>(stage + SG_INV + 16 * lane);
    const double2 vv = *reinterpret_cast<const double2*>(stage + SG_DV + 16 * lane);
    const double v2 = *reinterpret_cast<const double*>(stage + SG_DV + 16 * lane + 16);
    W.pA[SN] = pp.x; W.pB[SN] = pp.y;
    if (BND) {
      if (s == 0 && gjA == 0) W.pA[SN] = 0.0;  // pinned cell (gauss_seidel.py:145, :305)
      // cells outside the domain hold p = 0 (zero-filled by the TMA unit) and are never updated
    }
    W.fA[SN] = g.rho * uu.x * g.dy;
    W.fB[SN] = g.rho * uu.y * g.dy;
    W.g0[SC] = g.rho * vv.x * g.dx;
    W.g1[SC] = g.rho * vv.y * g.dx;
    W.g2[SC] = g.rho * v2 * g.dx;
    W.bA[SC] = bb.x; W.bB[SC] = bb.y;
    W.iA[SC] = iv.x; W.iB[SC] = iv.y;
  }
  // ---- the colour passes of this step: all on the same cell of the pair ----
  constexpr bool UPD_A = (PH & 1) != 0;  // s odd (r0 even) -> even columns
  double nb[NP];  // the neighbour across the pair boundary, state after the previous step
#pragma unroll
  for (int t = 0; t < NP; ++t) {
    const int a = (PH + 7 - t) & 7;
    nb[t] = UPD_A ? __shfl_up_sync(0xffffffffu, W.pB[a], 1) : __shfl_down_sync(0xffffffffu, W.pA[a], 1);
  }
  bool colA_in = true, colB_in = true, colA_bnd = false, colB_bnd = false;
  if (BND) {
    colA_in = gjA >= 0 && gjA < g.ny;
    colB_in = gjA + 1 >= 0 && gjA + 1 < g.ny;
    colA_bnd = gjA == 0 || gjA == g.ny - 1;
    colB_bnd = gjA + 1 == 0 || gjA + 1 == g.ny - 1;
  }
#pragma unroll
  for (int t = 0; t < NP; ++t) {
    const int a = (PH + 7 - t) & 7;    // row r = s-1-t
    const int se = (PH + 8 - t) & 7;   // row r+1
    const int sw = (PH + 6 - t) & 7;   // row r-1
    const int r = s - 1 - t;
    double aE, aW, aN, aS, pc, pE, pW, pN, pS, bc, ic;
    if (UPD_A) {
      aE = W.fA[se]; aW = W.fA[a]; aN = W.g1[a]; aS = W.g0[a];
      pc = W.pA[a]; pE = W.pA[se]; pW = W.pA[sw]; pN = W.pB[a]; pS = nb[t];
      bc = W.bA[a]; ic = W.iA[a];
    } else {
      aE = W.fB[se]; aW = W.fB[a]; aN = W.g2[a]; aS = W.g1[a];
      pc = W.pB[a]; pE = W.pB[se]; pW = W.pB[sw]; pN = nb[t]; pS = W.pA[a];
      bc = W.bB[a]; ic = W.iB[a];
    }
    bool ok = true;
    if (BND) {
      const bool row_in = r >= 0 && r < g.nx;
      const bool row_bnd = r == 0 || r == g.nx - 1;
      const bool col_bnd = UPD_A ? colA_bnd : colB_bnd;
      if (row_bnd) { aE = 0.0; aW = 0.0; }   // matrix_free.py:63-84: boundary cells keep no link across the edge direction
      if (col_bnd) { aN = 0.0; aS = 0.0; }
      ok = row_in && (UPD_A ? colA_in : colB_in) && !(UPD_A && r == 0 && gjA == 0);
      if (!ok) { aE = aW = aN = aS = 0.0; }  // keeps NaN coefficients of the array borders out of the arithmetic
    }
    double acc = bc;  // ((((b + E) + W) + N) + S) * (1/aP): gauss_seidel.py:285-299
    acc += aE * pE;
    acc += aW * pW;
    acc += aN * pN;
    acc += aS * pS;
    const double pn = acc * ic;
    const double pu = pc + omega * (pn - pc);
    if (UPD_A) { if (ok) W.pA[a] = pu; } else { if (ok) W.pB[a] = pu; }
  }
  // ---- row s-NP is final ----
  {
    constexpr int SF = (PH + 8 - NP) & 7;
    const int rf = s - NP;
    const int c = 2 * lane;
    if (rf >= jb.ia && rf < jb.ib && c >= NP && c < SW - NP && gjA < g.ny) {
      const size_t kk = nf_idx(g, rf, gjA);
      if (!BND || gjA + 1 < g.ny) *reinterpret_cast<double2*>(pout + kk) = make_double2(W.pA[SF], W.pB[SF]);
      else pout[kk] = W.pA[SF];
    }
  }
}

struct StreamMaps {
  CUtensorMap p, b, du, dv, inv;
};

// registers per thread the launch bounds leave (64 K registers per SM, allocated per warp in units of 8 per thread)
constexpr int stream_maxreg(int wpc) { return ((65536 / (32 * wpc)) / 8) * 8 > 255 ? 255 : ((65536 / (32 * wpc)) / 8) * 8; }

template <int NS, int WPC>
__global__ void __launch_bounds__(32 * WPC, 1) __maxnreg__(stream_maxreg(WPC))
k_rbsor_stream(nf_grid g, const __grid_constant__ StreamMaps maps, double* __restrict__ pout, double omega, int nstrips,
               int njobs, int chunk_len) {
  constexpr int NP = 2 * NS;
  constexpr int SCOLS = SW - 2 * NP;  // columns a strip finalises
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int job = blockIdx.x * WPC + warp;
  if (job >= njobs) return;
  unsigned char* ring = smem + (size_t)warp * (NSTG * SG_BYTES);
  const unsigned bar0 = s_u32(smem + (size_t)WPC * NSTG * SG_BYTES + warp * NSTG * 8);
  const unsigned ring_u = s_u32(ring);

  Job jb;
  {
    const int chunk = job / nstrips, strip = job - chunk * nstrips;
    jb.ia = g.gb + chunk * chunk_len;
    jb.ib = jb.ia + chunk_len < g.ge ? jb.ia + chunk_len : g.ge;
    jb.r0 = (jb.ia - NP) & ~1;
    jb.s_last = jb.ib - 1 + NP;
    jb.j0 = strip * SCOLS - NP;
  }
  const bool bnd = jb.r0 <= 0 || jb.s_last >= g.nx - 1 || jb.j0 <= 0 || jb.j0 + SW >= g.ny - 1;

  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NSTG; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0 + 8 * q) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  auto issue = [&](int s) {  // stage of step s: p[s], d_u[s]; d_v, b, 1/aP of row s-1
    const int q = (s - jb.r0) & (NSTG - 1);
    const unsigned bar = bar0 + 8 * q, dst = ring_u + q * SG_BYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(SG_TX) : "memory");
    tma_row(dst + SG_P, &maps.p, jb.j0, s - g.row0, bar);
    tma_row(dst + SG_DU, &maps.du, jb.j0, s - g.row0, bar);
    tma_row(dst + SG_B, &maps.b, jb.j0, s - 1 - g.row0, bar);
    tma_row(dst + SG_INV, &maps.inv, jb.j0, s - 1 - g.row0, bar);
    tma_row(dst + SG_DV, &maps.dv, jb.j0, s - 1 - g.row0, bar);
  };
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NSTG; ++q)
      if (jb.r0 + q <= jb.s_last) issue(jb.r0 + q);
  }

  Win W;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    W.pA[q] = W.pB[q] = 0.0;
    W.fA[q] = W.fB[q] = 0.0;
    W.g0[q] = W.g1[q] = W.g2[q] = 0.0;
    W.bA[q] = W.bB[q] = 0.0;
    W.iA[q] = W.iB[q] = 1.0;
  }

#define NF_STREAM_STEP(PH)                                                                          \
  {                                                                                                 \
    const int s = s8 + PH;                                                                          \
    if (s > jb.s_last) break;                                                                       \
    const unsigned bar = bar0 + 8 * (PH & (NSTG - 1));                                              \
    while (!mbar_try(bar, (PH >> 2) & 1)) {}                                                        \
    const unsigned char* stage = ring + (PH & (NSTG - 1)) * SG_BYTES;                               \
    if (bnd) stream_step<PH, NP, true>(W, g, jb, s, lane, omega, stage, pout);                     \
    else stream_step<PH, NP, false>(W, g, jb, s, lane, omega, stage, pout);                        \
    __syncwarp();                                                                                   \
    if (lane == 0 && s + NSTG <= jb.s_last) issue(s + NSTG);                                        \
  }

  for (int s8 = jb.r0;; s8 += 8) {
    NF_STREAM_STEP(0) NF_STREAM_STEP(1) NF_STREAM_STEP(2) NF_STREAM_STEP(3)
    NF_STREAM_STEP(4) NF_STREAM_STEP(5) NF_STREAM_STEP(6) NF_STREAM_STEP(7)
  }
#undef NF_STREAM_STEP
}

// ---- tensor maps, cached per (array, shape): re-encoding costs a driver call per array and launch ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn stream_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

struct MapKey {
  const void* base;
  int rows, cols, ld, box_cols;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.base);
    h = h * 1000003u ^ (size_t)k.rows;
    h = h * 1000003u ^ (size_t)k.cols;
    h = h * 1000003u ^ (size_t)k.ld;
    h = h * 1000003u ^ (size_t)k.box_cols;
    return h;
  }
};

bool row_map(CUtensorMap* out, const double* base, int rows, int cols, int ld, int box_cols) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  std::lock_guard<std::mutex> lock(mu);
  const MapKey key{base, rows, cols, ld, box_cols};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = stream_encoder();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 1u};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return true;
}

template <int NS, int WPC>
int launch_stream(nf_ctx* ctx, const nf_grid* g, const double* pin, double* pout, const double* b, const double* d_u,
                  const double* d_v, const double* inv, double omega, bool* used) {
  constexpr int NP = 2 * NS, SCOLS = SW - 2 * NP;
  constexpr int SMEM = WPC * NSTG * SG_BYTES + WPC * NSTG * 8;
  *used = false;
  const int rows = g->ge - g->gb;
  const int nstrips = (g->ny + SCOLS - 1) / SCOLS;
  const int slots = NF_SM_COUNT * WPC;
  int nchunks = slots / nstrips;
  if (nchunks < 1) nchunks = 1;
  int chunk_len = (rows + nchunks - 1) / nchunks;
  if (chunk_len < 8) chunk_len = 8;
  chunk_len = (chunk_len + 1) & ~1;  // even chunk starts keep the row parity of the phases
  if ((g->gb & 1) != 0) return NF_OK;  // odd origin: the caller falls back
  nchunks = (rows + chunk_len - 1) / chunk_len;
  const int njobs = nstrips * nchunks;
  const int row_end = g->row1 > 0 ? g->row1 : g->nx + 1;
  const int stored_p = (row_end < g->nx ? row_end : g->nx) - g->row0;
  const int stored_u = row_end - g->row0;
  StreamMaps m;
  if (!row_map(&m.p, pin, stored_p, g->ny, g->ld, SW) || !row_map(&m.b, b, stored_p, g->ny, g->ld, SW) ||
      !row_map(&m.du, d_u, stored_u, g->ny, g->ld, SW) || !row_map(&m.dv, d_v, stored_p, g->ny + 1, g->ld, SW + 2) ||
      !row_map(&m.inv, inv, stored_p, g->ny, g->ld, SW))
    return NF_OK;
  static bool attr_set = false;
  if (!attr_set) {
    NF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_rbsor_stream<NS, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  const int grid = (njobs + WPC - 1) / WPC;
  k_rbsor_stream<NS, WPC><<<grid, 32 * WPC, SMEM, ctx->stream>>>(*g, m, pout, omega, nstrips, njobs, chunk_len);
  NF_LAUNCH_CHECK(ctx);
  *used = true;
  return NF_OK;
}

}  // namespace

int nfi_rbsor_stream(nf_ctx* ctx, const nf_grid* g, const double* pin, double* pout, const double* b, const double* d_u,
                     const double* d_v, const double* inv, double omega, int ns, bool* used) {
  *used = false;
  if (!inv) return NF_OK;
  const int wpc = getenv("NF_STREAM_WPC") ? atoi(getenv("NF_STREAM_WPC")) : 12;
  if (ns == 3) {
    if (wpc == 8) return launch_stream<3, 8>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, used);
    if (wpc == 9) return launch_stream<3, 9>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, used);
    if (wpc == 11) return launch_stream<3, 11>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, used);
    if (wpc == 10) return launch_stream<3, 10>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, used);
    return launch_stream<3, 12>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, used);
  }
  if (ns == 2) return launch_stream<2, 12>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, used);
  if (ns == 1) return launch_stream<1, 12>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, used);
  return NF_OK;
}
