// nf_momentum_fused.cu -- temporally blocked Jacobi momentum sweeps (K3) with the residual norms (K4) fused
// behind the last sweep.  Same arithmetic as k_momentum_jacobi / k_momentum_residual (nf_momentum.cu), i.e.
// jacobi_matrix_solver.py:196-208 and :221-262 of the reference, bit for bit: a Jacobi sweep of a cell depends only on
// the previous iterate of its four neighbours, which the trapezoid scheme reproduces exactly inside the tile.
//
// One CTA (32 x 16 threads) loads a 48 x 64 region (tile + halo of K rows / an even number >= K of columns) of the
// six link arrays and x once, runs K sweeps ping-ponging x between two shared-memory planes with the per-cell
// constants (a_e, a_w, a_n, a_s, 1/a_p, b) in registers, and stores the tile: 64 B/cell of HBM traffic for K sweeps
// instead of K x 72 B/cell.  Thread (tx, ty) owns the column pair (2tx, 2tx+1) of region rows ty, ty+16, ty+32.
//
// TMA variant (large grids): the seven operand arrays of the NEXT tile travel into a 168 KB shared-memory stage
// (cp.async.bulk.tensor.2d, one 48 x 64 box per array, out-of-array elements zero-filled) while the K sweeps of the
// current tile run, so the HBM stream no longer stops during the compute phases (the plain variant alternates
// "load everything" / "compute": 0.35 of the HBM peak at 4097^2).  Same registers, same arithmetic, same bits.
#include <stdlib.h>
#include <string.h>

#include <cuda.h>

#include "nf_common.cuh"

bool nfi_tensor_map_2d(CUtensorMap* out, const double* base, int rows, int cols, int ld, int box_cols, int box_rows);

namespace {

constexpr int RW = 64, RH = 48, NYT = 16, KS = 3;

__device__ __host__ __forceinline__ int rows_of(const nf_grid& g, int is_u) { return g.nx + (is_u ? 1 : 0); }
__device__ __host__ __forceinline__ int cols_of(const nf_grid& g, int is_u) { return g.ny + (is_u ? 0 : 1); }
__device__ __host__ __forceinline__ int row_end_of(const nf_grid& g, int is_u) {
  return (is_u && g.ge == g.nx) ? g.nx + 1 : g.ge;
}

template <int K, bool WITH_RES>
struct MGeom {
  static constexpr int HR = K + (WITH_RES ? 1 : 0);
  static constexpr int HC = ((HR + 1) / 2) * 2;
  static constexpr int TR = RH - 2 * HR;
  static constexpr int TC = RW - 2 * HC;
};

struct MomMaps { CUtensorMap m[7]; };  // a_e, a_w, a_n, a_s, a_p, src, x (TMA variant)
constexpr int STAGE_BYTES = 7 * RH * RW * (int)sizeof(double);
constexpr int SX_BYTES = 2 * RH * (RW + 2) * (int)sizeof(double);

__device__ __forceinline__ unsigned mom_s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// g: rows [g.gb, row_end) are written.  norm_b / norm_e: rows whose residual enters the norms / the field.
template <int IS_U, int K, bool WITH_RES, bool TMA>
__global__ void __launch_bounds__(32 * NYT, 1)
k_momentum_fused(nf_grid g, nf_links L, const __grid_constant__ MomMaps maps, const double* __restrict__ xin,
                 double* __restrict__ xout, int tiles_x, int n_tiles, int norm_b, int norm_e, double* __restrict__ field,
                 double* partials, unsigned int* ticket, double* out) {
  using G = MGeom<K, WITH_RES>;
  constexpr int HR = G::HR, HC = G::HC, TR = G::TR, TC = G::TC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double (&sX)[2][RH][RW + 2] = *reinterpret_cast<double (*)[2][RH][RW + 2]>(smem_raw + (TMA ? STAGE_BYTES : 0));
  double (&sC)[7][RH][RW] = *reinterpret_cast<double (*)[7][RH][RW]>(smem_raw);  // TMA only
  const unsigned bar = mom_s32(smem_raw + STAGE_BYTES + SX_BYTES);               // TMA only
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c0 = 2 * tx;
  const int rows = rows_of(g, IS_U), cols = cols_of(g, IS_U);
  const int rend = row_end_of(g, IS_U);
  double nrm[2] = {0.0, 0.0};

  auto issue = [&](int tile) {  // one thread: the seven boxes of a tile -> stage
    const int ti = tile / tiles_x, tj = tile - ti * tiles_x;
    const int ci = g.gb + ti * TR - HR - g.row0, cj = tj * TC - HC;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((unsigned)STAGE_BYTES) : "memory");
#pragma unroll
    for (int q = 0; q < 7; ++q)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   :: "r"(mom_s32(&sC[q][0][0])), "l"(&maps.m[q]), "r"(cj), "r"(ci), "r"(bar) : "memory");
  };
  unsigned parity = 0;
  if (TMA) {
    if (tx == 0 && ty == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      if ((int)blockIdx.x < n_tiles) issue(blockIdx.x);
    }
    __syncthreads();
  }

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int ti = tile / tiles_x, tj = tile - ti * tiles_x;
    const int i0 = g.gb + ti * TR - HR;
    const int j0 = tj * TC - HC;
    const int gj0 = j0 + c0;
    if (TMA) {
      unsigned ok = 0;
      while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
      parity ^= 1;
    }

    double ae0[KS], aw0[KS], an0[KS], as0[KS], di0[KS], b0[KS];
    double ae1[KS], aw1[KS], an1[KS], as1[KS], di1[KS], b1[KS];
    double x0[KS], x1[KS];
    bool in0[KS], in1[KS];

#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int r = ty + NYT * k;
      const int gi = i0 + r;
      const bool rin = gi >= 0 && gi < rows && nf_row_stored(g, gi);  // rows outside the slab's storage: never needed
      in0[k] = rin && gj0 >= 0 && gj0 < cols;
      in1[k] = rin && gj0 + 1 >= 0 && gj0 + 1 < cols;
      ae0[k] = aw0[k] = an0[k] = as0[k] = di0[k] = b0[k] = 0.0;
      ae1[k] = aw1[k] = an1[k] = as1[k] = di1[k] = b1[k] = 0.0;
      x0[k] = x1[k] = 0.0;
      if (in0[k]) {  // gj0 even, pitch even: aligned pair loads; column gj0+1 <= cols <= ld-1 stays inside the row
        double2 e, w, n, s, p, b, x;
        if (TMA) {
          e = *reinterpret_cast<const double2*>(&sC[0][r][c0]);
          w = *reinterpret_cast<const double2*>(&sC[1][r][c0]);
          n = *reinterpret_cast<const double2*>(&sC[2][r][c0]);
          s = *reinterpret_cast<const double2*>(&sC[3][r][c0]);
          p = *reinterpret_cast<const double2*>(&sC[4][r][c0]);
          b = *reinterpret_cast<const double2*>(&sC[5][r][c0]);
          x = *reinterpret_cast<const double2*>(&sC[6][r][c0]);
        } else {
          const size_t kk = nf_idx(g, gi, gj0);
          e = *reinterpret_cast<const double2*>(L.a_e + kk);
          w = *reinterpret_cast<const double2*>(L.a_w + kk);
          n = *reinterpret_cast<const double2*>(L.a_n + kk);
          s = *reinterpret_cast<const double2*>(L.a_s + kk);
          p = *reinterpret_cast<const double2*>(L.a_p + kk);
          b = *reinterpret_cast<const double2*>(L.src + kk);
          x = *reinterpret_cast<const double2*>(xin + kk);
        }
        // neighbour terms exist only inside the array (jacobi_matrix_solver.py:48-151): zero the others
        ae0[k] = (gi < rows - 1) ? e.x : 0.0; aw0[k] = (gi > 0) ? w.x : 0.0;
        an0[k] = (gj0 < cols - 1) ? n.x : 0.0; as0[k] = (gj0 > 0) ? s.x : 0.0;
        di0[k] = (fabs(p.x) > 1e-12) ? 1.0 / p.x : 0.0;
        b0[k] = b.x; x0[k] = x.x;
        if (in1[k]) {
          ae1[k] = (gi < rows - 1) ? e.y : 0.0; aw1[k] = (gi > 0) ? w.y : 0.0;
          an1[k] = (gj0 + 1 < cols - 1) ? n.y : 0.0; as1[k] = s.y;
          di1[k] = (fabs(p.y) > 1e-12) ? 1.0 / p.y : 0.0;
          b1[k] = b.y; x1[k] = x.y;
        }
      }
      *reinterpret_cast<double2*>(&sX[0][r][c0]) = make_double2(x0[k], x1[k]);
    }
    __syncthreads();
    if (TMA) {  // the stage is in registers now: fetch the next tile behind the sweeps
      if (tx == 0 && ty == 0 && tile + (int)gridDim.x < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(tile + gridDim.x);
      }
    }

    int cur = 0;
#pragma unroll
    for (int s = 0; s < K; ++s) {
      double n0[KS], n1[KS];
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int r = ty + NYT * k;
        const int rm = r > 0 ? r - 1 : 0, rp = r < RH - 1 ? r + 1 : RH - 1;
        const double2 up = *reinterpret_cast<const double2*>(&sX[cur][rp][c0]);   // row i+1 ("E")
        const double2 dn = *reinterpret_cast<const double2*>(&sX[cur][rm][c0]);   // row i-1 ("W")
        const double lf = sX[cur][r][c0 > 0 ? c0 - 1 : 0];                        // column j-1 ("S") of cell 0
        const double rt = sX[cur][r][c0 + 2];                                     // column j+1 ("N") of cell 1
        // x_new = D^-1 (b - (A-D) x), off-diagonal sum in CSR column order W, S, N, E
        double a0 = 0.0;
        a0 += (-aw0[k]) * dn.x;
        a0 += (-as0[k]) * lf;
        a0 += (-an0[k]) * x1[k];
        a0 += (-ae0[k]) * up.x;
        n0[k] = di0[k] * (b0[k] - a0);
        double a1 = 0.0;
        a1 += (-aw1[k]) * dn.y;
        a1 += (-as1[k]) * x0[k];
        a1 += (-an1[k]) * rt;
        a1 += (-ae1[k]) * up.y;
        n1[k] = di1[k] * (b1[k] - a1);
      }
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int r = ty + NYT * k;
        // cells outside the array stay 0; region-edge cells are recomputed with clamped neighbours: harmless,
        // they are never read by a cell that matters
        x0[k] = in0[k] ? n0[k] : 0.0;
        x1[k] = in1[k] ? n1[k] : 0.0;
        *reinterpret_cast<double2*>(&sX[cur ^ 1][r][c0]) = make_double2(x0[k], x1[k]);
      }
      cur ^= 1;
      __syncthreads();
    }

    // store the tile
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int r = ty + NYT * k;
      const int gi = i0 + r;
      if (r < HR || r >= HR + TR || gi >= rend || gi < g.gb) continue;
      if (c0 < HC || c0 >= HC + TC || gj0 >= cols) continue;
      const size_t kk = nf_idx(g, gi, gj0);
      if (gj0 + 1 < cols) *reinterpret_cast<double2*>(xout + kk) = make_double2(x0[k], x1[k]);
      else xout[kk] = x0[k];
    }

    if (WITH_RES) {
      // r = b - A x of the relaxed system (diagonal term between S and N: sorted CSR order); neighbours of tile
      // cells are exact (they lie >= K cells inside the region)
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int r = ty + NYT * k;
        const int gi = i0 + r;
        if (r < HR || r >= HR + TR || gi >= rend || gi < norm_b || gi >= norm_e) continue;
        if (c0 < HC || c0 >= HC + TC || gj0 >= cols) continue;
        const size_t kk = nf_idx(g, gi, gj0);
        const double2 ap = *reinterpret_cast<const double2*>(L.a_p + kk);
        const double2 up = *reinterpret_cast<const double2*>(&sX[cur][r + 1][c0]);
        const double2 dn = *reinterpret_cast<const double2*>(&sX[cur][r - 1][c0]);
        const double lf = sX[cur][r][c0 - 1];
        const double rt = sX[cur][r][c0 + 2];
        double a0 = 0.0;
        a0 += (-aw0[k]) * dn.x;
        a0 += (-as0[k]) * lf;
        a0 += ap.x * x0[k];
        a0 += (-an0[k]) * x1[k];
        a0 += (-ae0[k]) * up.x;
        const double r0 = b0[k] - a0;
        const bool edge_i = (gi == 0 || gi == rows - 1);
        const bool zero_i = IS_U ? (gi == 0 || gi == 1 || gi == g.nx - 1 || gi == g.nx) : false;
        {
          const int gj = gj0;
          const bool edge = edge_i || gj == 0 || gj == cols - 1;
          if (!edge) { nrm[0] += r0 * r0; nrm[1] += b0[k] * b0[k]; }
          if (field) {
            const bool zero = IS_U ? zero_i : (gj == 0 || gj == 1 || gj == g.ny - 1 || gj == g.ny);
            field[kk] = zero ? 0.0 : r0;
          }
        }
        if (gj0 + 1 < cols) {
          double a1 = 0.0;
          a1 += (-aw1[k]) * dn.y;
          a1 += (-as1[k]) * x0[k];
          a1 += ap.y * x1[k];
          a1 += (-an1[k]) * rt;
          a1 += (-ae1[k]) * up.y;
          const double r1 = b1[k] - a1;
          const int gj = gj0 + 1;
          const bool edge = edge_i || gj == 0 || gj == cols - 1;
          if (!edge) { nrm[0] += r1 * r1; nrm[1] += b1[k] * b1[k]; }
          if (field) {
            const bool zero = IS_U ? zero_i : (gj == 0 || gj == 1 || gj == g.ny - 1 || gj == g.ny);
            field[kk + 1] = zero ? 0.0 : r1;
          }
        }
      }
    }
    __syncthreads();  // sX is rewritten by the next tile
  }
  if (WITH_RES) nf_block_reduce_store<2>(nrm, partials, ticket, out);
}

template <int IS_U, int K, bool WITH_RES>
int launch(nf_ctx* ctx, const nf_grid* g, nf_links L, const double* xin, double* xout, int norm_b, int norm_e,
           double* field, double* out) {
  using G = MGeom<K, WITH_RES>;
  const int nrows = row_end_of(*g, IS_U) - g->gb;
  const int tiles_x = (cols_of(*g, IS_U) + G::TC - 1) / G::TC, tiles_y = (nrows + G::TR - 1) / G::TR;
  const int n_tiles = tiles_x * tiles_y;
  const int grid = n_tiles < NF_SM_COUNT ? n_tiles : NF_SM_COUNT;
  MomMaps maps;
  // TMA variant: several tiles per CTA (something to prefetch), 16-byte aligned operands, encoder available
  const char* env = getenv("NF_MOMENTUM_TMA");
  const int min_tiles = env ? atoi(env) : 4 * NF_SM_COUNT;
  bool tma = n_tiles >= min_tiles && min_tiles >= 0 && !(env && env[0] == '-');
  if (tma) {
    const int srows = (nf_stored_end(*g) < rows_of(*g, IS_U) ? nf_stored_end(*g) : rows_of(*g, IS_U)) - g->row0;
    const double* arr[7] = {L.a_e, L.a_w, L.a_n, L.a_s, L.a_p, L.src, xin};
    for (int q = 0; q < 7 && tma; ++q)
      tma = ((uintptr_t)arr[q] % 16) == 0 && nfi_tensor_map_2d(&maps.m[q], arr[q], srows, cols_of(*g, IS_U), g->ld, RW, RH);
  }
  if (tma) {
    constexpr int SMEM = STAGE_BYTES + SX_BYTES + 16;
    static bool attr_set = false;
    if (!attr_set) {
      NF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_momentum_fused<IS_U, K, WITH_RES, true>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      attr_set = true;
    }
    k_momentum_fused<IS_U, K, WITH_RES, true><<<grid, dim3(32, NYT, 1), SMEM, ctx->stream>>>(
        *g, L, maps, xin, xout, tiles_x, n_tiles, norm_b, norm_e, field, ctx->partials, ctx->ticket, out);
    NF_LAUNCH_CHECK(ctx);
    return NF_OK;
  }
  memset(&maps, 0, sizeof(maps));
  constexpr int SMEM = SX_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    NF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_momentum_fused<IS_U, K, WITH_RES, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  k_momentum_fused<IS_U, K, WITH_RES, false><<<grid, dim3(32, NYT, 1), SMEM, ctx->stream>>>(
      *g, L, maps, xin, xout, tiles_x, n_tiles, norm_b, norm_e, field, ctx->partials, ctx->ticket, out);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

template <int IS_U, bool WITH_RES>
int launch_k(nf_ctx* ctx, int k, const nf_grid* g, nf_links L, const double* xin, double* xout, int norm_b, int norm_e,
             double* field, double* out) {
  switch (k) {
    case 1: return launch<IS_U, 1, WITH_RES>(ctx, g, L, xin, xout, norm_b, norm_e, field, out);
    case 2: return launch<IS_U, 2, WITH_RES>(ctx, g, L, xin, xout, norm_b, norm_e, field, out);
    case 3: return launch<IS_U, 3, WITH_RES>(ctx, g, L, xin, xout, norm_b, norm_e, field, out);
    case 4: return launch<IS_U, 4, WITH_RES>(ctx, g, L, xin, xout, norm_b, norm_e, field, out);
    case 5: return launch<IS_U, 5, WITH_RES>(ctx, g, L, xin, xout, norm_b, norm_e, field, out);
    default: return launch<IS_U, 6, WITH_RES>(ctx, g, L, xin, xout, norm_b, norm_e, field, out);
  }
}

}  // namespace

// k (1..6) Jacobi sweeps xin -> xout over the rows [g.gb, row_end) of the descriptor (reads k(+1) rows beyond).
// with_res: also the relaxed residual of the result: sums of r^2 / b^2 over the non-edge cells of rows
// [norm_b, norm_e) -> out[0..1] (device), residual field (may be NULL).
int nfi_momentum_sweeps_fused(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, const double* xin, double* xout,
                              int k, int with_res, int norm_b, int norm_e, double* field, double* out) {
  if (k < 1 || k > 6) return NF_ERR_ARG;
  if (is_u) {
    if (with_res) return launch_k<1, true>(ctx, k, g, L, xin, xout, norm_b, norm_e, field, out);
    return launch_k<1, false>(ctx, k, g, L, xin, xout, norm_b, norm_e, field, out);
  }
  if (with_res) return launch_k<0, true>(ctx, k, g, L, xin, xout, norm_b, norm_e, field, out);
  return launch_k<0, false>(ctx, k, g, L, xin, xout, norm_b, norm_e, field, out);
}

// C-ABI: n_sweeps temporally blocked Jacobi sweeps (result in x; tmp is a same-shape scratch array) and, when
// rel_norm_host != NULL, the relaxed residual norm of the result as nf_momentum_residual returns it.
extern "C" int nf_momentum_jacobi_fused(nf_ctx* ctx, const nf_grid* g, int is_u, nf_links L, double* x, double* tmp,
                                        int n_sweeps, double* field_out, double* rel_norm_host) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, L.a_e && L.a_w && L.a_n && L.a_s && L.a_p && L.src, "NULL link array");
  NF_REQUIRE(ctx, n_sweeps >= 1 && x && tmp && x != tmp, "bad sweep arguments");
  NF_REQUIRE(ctx, (g->ld % 2) == 0, "row pitch must be even");
  double* src = x;
  double* dst = tmp;
  int left = n_sweeps;
  const int nb = g->gb, ne = (is_u && g->ge == g->nx) ? g->nx + 1 : g->ge;
  while (left > 0) {
    const int k = left > 6 ? 6 : left;
    const bool last = (left == k) && rel_norm_host != nullptr;
    NF_TRY(nfi_momentum_sweeps_fused(ctx, g, is_u, L, src, dst, k, last ? 1 : 0, nb, ne, last ? field_out : nullptr,
                                     ctx->scalars));
    double* t = src; src = dst; dst = t;
    left -= k;
  }
  if (src != x) {
    const size_t rows = (size_t)(ne - g->gb);
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(x + (size_t)(g->gb - g->row0) * g->ld, src + (size_t)(g->gb - g->row0) * g->ld,
                                       rows * g->ld * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (rel_norm_host) {
    double s[2];
    NF_TRY(nf_read_scalars(ctx, 0, 2, s));
    *rel_norm_host = sqrt(s[0]) / (sqrt(s[1]) + 1e-15);
  }
  return NF_OK;
}
