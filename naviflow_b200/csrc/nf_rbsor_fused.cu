// nf_rbsor_fused.cu -- temporally blocked red-black SOR (K8): NS full sweeps (2*NS colour passes) per
// tile load.  Same arithmetic as gauss_seidel.py:214-305 / k_rbsor_color and bit-identical to it: the update
// of a cell depends only on its four opposite-colour neighbours after the previous colour pass, which the
// trapezoid scheme reproduces exactly inside the tile.
//
// Layout of the work
//   * One CTA (32 x 16 threads) owns an output tile of TR x TC cells and loads the region grown by the halo
//     H = 2*NS on every side (48 rows x 64 columns).  After colour pass t the cells at distance > t from the
//     region edge are exact, so after 2*NS passes the tile is exact.
//   * Thread (tx, ty) owns the column pair (2tx, 2tx+1) of the three consecutive region rows 3ty, 3ty+1, 3ty+2:
//     one red and one black cell per row.  Their constants (b, 1/aP, aE, aW, aN, aS) live in REGISTERS for all
//     passes; p is mirrored in shared memory for the neighbouring threads, stored split by colour so that every
//     access is lane-contiguous (conflict-free): sP[parity][row][pair].  Vertical neighbours inside the row triple
//     are read from the thread's own registers (see rbsor_passes).
//   * HBM traffic per launch: read p, b, d_u, d_v (32 B/cell x region/tile overhead, largely absorbed by L2
//     between neighbouring tiles) + write p (8 B/cell) -- versus 2*NS passes x ~40 B/cell unfused.
//   * p is double buffered in global memory (p_in -> p_out): neighbouring tiles read each other's halo.
#include <stdlib.h>

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "nf_pressure.cuh"

bool nfi_tensor_map_2d(CUtensorMap* out, const double* base, int rows, int cols, int ld, int box_cols, int box_rows);

namespace {

constexpr int RCW = 64;        // region width in cells (32 column pairs = one warp per row)
constexpr int NYT = 16;        // threads in y
constexpr int KS = 3;          // row slots per thread
constexpr int RRW = NYT * KS;  // region rows (48)

struct CellCoef {
  double e, w, n, s, inv;
};

// coefficients of cell (gi,gj) from the raw d_u[gi][gj], d_u[gi+1][gj], d_v[gi][gj], d_v[gi][gj+1] with the
// reference's Neumann folding (matrix_free.py:52-84, gauss_seidel.py:243-266: aP < 1e-15 -> 1)
template <bool HAS_INV>
__device__ __forceinline__ CellCoef cell_coef_interior(const nf_grid& g, double du_c, double du_e, double dv_c,
                                                       double dv_n, double inv_pre) {
  // every cell of the region is strictly inside the domain: no folding, no masks
  CellCoef c;
  c.e = g.rho * du_e * g.dy;
  c.w = g.rho * du_c * g.dy;
  c.n = g.rho * dv_n * g.dx;
  c.s = g.rho * dv_c * g.dx;
  if (HAS_INV) {
    c.inv = inv_pre;
  } else {
    double diag = 0.0;
    diag += ((c.e + c.w) + c.n) + c.s;
    if (diag < 1e-15) diag = 1.0;
    c.inv = 1.0 / diag;
  }
  return c;
}

template <bool HAS_INV>
__device__ __forceinline__ CellCoef cell_coef(const nf_grid& g, int gi, int gj, bool inside, double du_c, double du_e,
                                              double dv_c, double dv_n, double inv_pre) {
  CellCoef c;
  if (!inside) {
    c.e = c.w = c.n = c.s = 0.0;
    c.inv = 1.0;
    return c;
  }
  double e = (gi < g.nx - 1) ? g.rho * du_e * g.dy : 0.0;
  double w = (gi > 0) ? g.rho * du_c * g.dy : 0.0;
  double n = (gj < g.ny - 1) ? g.rho * dv_n * g.dx : 0.0;
  double s = (gj > 0) ? g.rho * dv_c * g.dx : 0.0;
  double diag = 0.0;
  if (gi == 0) diag += e;
  if (gi == g.nx - 1) diag += w;
  if (gj == 0) diag += n;
  if (gj == g.ny - 1) diag += s;
  if (gi == 0) e = 0.0;
  if (gi == g.nx - 1) w = 0.0;
  if (gj == 0) n = 0.0;
  if (gj == g.ny - 1) s = 0.0;
  c.e = e; c.w = w; c.n = n; c.s = s;
  if (HAS_INV) {
    c.inv = inv_pre;  // 1/aP precomputed once per level by k_inv_diag (same expression, same rounding)
  } else {
    diag += ((e + w) + n) + s;
    if (diag < 1e-15) diag = 1.0;
    c.inv = 1.0 / diag;
  }
  return c;
}

// The 2*NS colour passes.  S0 = parity of (global row + first column) of this thread's rows: it is the same for
// all of a thread's row slots (rows differ by 16) and uniform in a warp, so the kernel branches once on it and
// every "which cell of the pair is red" decision below folds at compile time.
// aP of cell (gi,gj) as A*p uses it (matrix_free.py:63-84), from the raw d_u / d_v values
__device__ __forceinline__ double cell_diag(const nf_grid& g, int gi, int gj, double du_c, double du_e, double dv_c,
                                            double dv_n) {
  double e = (gi < g.nx - 1) ? g.rho * du_e * g.dy : 0.0;
  double w = (gi > 0) ? g.rho * du_c * g.dy : 0.0;
  double n = (gj < g.ny - 1) ? g.rho * dv_n * g.dx : 0.0;
  double s = (gj > 0) ? g.rho * dv_c * g.dx : 0.0;
  double diag = 0.0;
  if (gi == 0) diag += e;
  if (gi == g.nx - 1) diag += w;
  if (gj == 0) diag += n;
  if (gj == g.ny - 1) diag += s;
  if (gi == 0) e = 0.0;
  if (gi == g.nx - 1) w = 0.0;
  if (gj == 0) n = 0.0;
  if (gj == g.ny - 1) s = 0.0;
  diag += ((e + w) + n) + s;
  return diag;
}

// region row of slot k of the threads with threadIdx.y == ty: three CONSECUTIVE rows per thread, so that the vertical
// neighbours of a thread's middle row (and one of each outer row) are the thread's own registers
__device__ __forceinline__ int slot_row(int ty, int k) { return KS * ty + k; }

template <int NS, int S0>
__device__ __forceinline__ void rbsor_passes(double (&sP)[2][RRW][33], int tx, int ty, double omega,
                                             double (&p0)[KS], double (&p1)[KS], const double (&b0)[KS],
                                             const double (&b1)[KS], const double (&inv0)[KS], const double (&inv1)[KS],
                                             const double (&aE0)[KS], const double (&aW0)[KS], const double (&aN0)[KS],
                                             const double (&aS0)[KS], const double (&aE1)[KS], const double (&aW1)[KS],
                                             const double (&aN1)[KS], const double (&aS1)[KS], const bool (&ok0)[KS],
                                             const bool (&ok1)[KS]) {
  // All three row slots are updated in every pass without a branch (the commit is predicated on `ok`), so the
  // compiler can interleave the three independent fp64 dependency chains.  Rows outside the shrinking trapezoid
  // are updated too; their values are never read by cells that matter.  A thread owns rows 3ty, 3ty+1, 3ty+2: of the
  // six vertical neighbours of a pass only two (row 3ty-1 and row 3ty+3) come from shared memory, the others are the
  // thread's own registers -- the cell with the same index one row up / down has the opposite colour and is not
  // updated in this pass.  S0: which cell of the pair is red ((i+j) even) in slot 0's row; it alternates with the slot.
  const int r0 = slot_row(ty, 0);
  const int rup = r0 > 0 ? r0 - 1 : 0;                    // clamped at the region edge (that row's result is discarded)
  const int rdn = r0 + KS < RRW ? r0 + KS : RRW - 1;
  const int txm = tx > 0 ? tx - 1 : 0;
  const int par = ty & 1;  // parity of slot 0's row
#pragma unroll
  for (int t = 0; t < 2 * NS; ++t) {
    const int col = t & 1;            // 0: (i+j) even ("red"), 1: odd ("black")
    const int lp = par ^ S0 ^ col;    // shared-memory plane of the cells updated in this pass (the same for all slots)
    double pnew[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int r = r0 + k;
      const int s = S0 ^ col ^ (k & 1);  // which cell of the pair has colour `col` in this row: 0 -> cell 0, 1 -> cell 1
      const double pc = s ? p1[k] : p0[k];
      const double bc = s ? b1[k] : b0[k];
      const double ic = s ? inv1[k] : inv0[k];
      const double aE = s ? aE1[k] : aE0[k];
      const double aW = s ? aW1[k] : aW0[k];
      const double aN = s ? aN1[k] : aN0[k];
      const double aS = s ? aS1[k] : aS0[k];
      // opposite-colour neighbours.  Rows r+-1: the same cell index of the neighbouring slot (register) or, at the
      // ends of the thread's row triple, of the neighbouring warp's row (shared memory).  In the row: one is the
      // thread's own partner cell (register), the other belongs to the neighbouring pair (shared memory).
      const double pE = (k < KS - 1) ? (s ? p1[k < KS - 1 ? k + 1 : k] : p0[k < KS - 1 ? k + 1 : k]) : sP[lp ^ 1][rdn][tx];
      const double pW = (k > 0) ? (s ? p1[k > 0 ? k - 1 : k] : p0[k > 0 ? k - 1 : k]) : sP[lp ^ 1][rup][tx];
      const double pN = s ? sP[lp ^ 1][r][tx + 1] : p1[k];
      const double pS = s ? p0[k] : sP[lp ^ 1][r][txm];
      double acc = bc;       // ((((b + E) + W) + N) + S) * (1/aP): gauss_seidel.py:285-299
      acc += aE * pE;
      acc += aW * pW;
      acc += aN * pN;
      acc += aS * pS;
      const double pn = acc * ic;
      pnew[k] = pc + omega * (pn - pc);
    }
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int r = r0 + k;
      const int s = S0 ^ col ^ (k & 1);
      const bool ok = s ? ok1[k] : ok0[k];
      if (ok) {
        if (s) p1[k] = pnew[k]; else p0[k] = pnew[k];
        sP[lp][r][tx] = pnew[k];
      }
    }
    __syncthreads();
  }
}

template <int NS, bool HAS_INV>
__global__ void __launch_bounds__(32 * NYT, 1)
k_rbsor_fused(nf_grid g, const double* __restrict__ pin, double* __restrict__ pout, const double* __restrict__ b,
              const double* __restrict__ d_u, const double* __restrict__ d_v, const double* __restrict__ inv,
              double omega) {
  nf_pdl_entry();
  constexpr int H = 2 * NS;
  constexpr int TR = RRW - 2 * H;
  constexpr int TC = RCW - 2 * H;
  __shared__ double sP[2][RRW][33];  // [local parity][region row][column pair]

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i0 = g.gb + blockIdx.y * TR - H;  // global row of region row 0
  const int j0 = blockIdx.x * TC - H;         // global column of region column 0 (even)
  const int gj0 = j0 + 2 * tx;                // global column of this thread's first cell
  const int c0 = 2 * tx;                      // its region column

  // register state per row slot: cell 0 = (r, 2tx), cell 1 = (r, 2tx+1)
  double p0[KS], p1[KS], b0[KS], b1[KS], inv0[KS], inv1[KS];
  double aE0[KS], aW0[KS], aN0[KS], aS0[KS], aE1[KS], aW1[KS], aN1[KS], aS1[KS];
  bool ok0[KS], ok1[KS];  // cell may be updated: in the domain, not pinned, not on the region edge

#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const int r = slot_row(ty, k);
    const int gi = i0 + r;
    // rows gi and gi+1 (d_u) must lie inside the slab's storage; the overshoot rows of the last tile are never needed
    const bool row_in = (gi >= 0 && gi < g.nx) && nf_row_stored(g, gi) && nf_row_stored(g, gi + 1);
    const bool in0 = row_in && gj0 >= 0 && gj0 < g.ny;
    const bool in1 = row_in && gj0 + 1 >= 0 && gj0 + 1 < g.ny;
    double vp0 = 0.0, vp1 = 0.0, vb0 = 0.0, vb1 = 0.0;
    double uc0 = 0.0, uc1 = 0.0, ue0 = 0.0, ue1 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0, iv0 = 1.0, iv1 = 1.0;
    if (in0) {
      // gj0 is even and the pitch is even: 16-byte aligned pair loads.  Column gj0+1 <= ny lies inside the row
      // (ld >= ny+1), so the pair load is always in bounds; its second half is ignored when in1 is false.
      const size_t kk = nf_idx(g, gi, gj0);
      const double2 pp = *reinterpret_cast<const double2*>(pin + kk);
      const double2 bb = *reinterpret_cast<const double2*>(b + kk);
      const double2 ua = *reinterpret_cast<const double2*>(d_u + kk);
      const double2 ub = *reinterpret_cast<const double2*>(d_u + kk + g.ld);
      const double2 vv = *reinterpret_cast<const double2*>(d_v + kk);
      vp0 = pp.x; vp1 = pp.y; vb0 = bb.x; vb1 = bb.y;
      uc0 = ua.x; uc1 = ua.y; ue0 = ub.x; ue1 = ub.y; w0 = vv.x; w1 = vv.y;
      if (in1) w2 = d_v[kk + 2];  // d_v[gi][gj0+2], gj0+2 <= ny
      if (HAS_INV) {
        const double2 iv = *reinterpret_cast<const double2*>(inv + kk);
        iv0 = iv.x; iv1 = iv.y;
      }
    }
    const CellCoef ca = cell_coef<HAS_INV>(g, gi, gj0, in0, uc0, ue0, w0, w1, iv0);
    const CellCoef cb = cell_coef<HAS_INV>(g, gi, gj0 + 1, in1, uc1, ue1, w1, w2, iv1);
    aE0[k] = ca.e; aW0[k] = ca.w; aN0[k] = ca.n; aS0[k] = ca.s; inv0[k] = ca.inv;
    aE1[k] = cb.e; aW1[k] = cb.w; aN1[k] = cb.n; aS1[k] = cb.s; inv1[k] = cb.inv;
    if (gi == 0 && gj0 == 0) vp0 = 0.0;  // pinned cell (0,0) is held at 0 (gauss_seidel.py:145, :305)
    p0[k] = in0 ? vp0 : 0.0;
    p1[k] = in1 ? vp1 : 0.0;
    b0[k] = vb0;
    b1[k] = vb1;
    ok0[k] = in0 && !(gi == 0 && gj0 == 0) && r >= 1 && r <= RRW - 2 && c0 >= 1;
    ok1[k] = in1 && r >= 1 && r <= RRW - 2 && (c0 + 1) <= RCW - 2;
    // local parity of cell 0 is (r + c0) & 1 = r & 1; cell 1 has the other one
    sP[r & 1][r][tx] = p0[k];
    sP[(r & 1) ^ 1][r][tx] = p1[k];
  }
  __syncthreads();

  if ((i0 + ty + gj0) & 1)
    rbsor_passes<NS, 1>(sP, tx, ty, omega, p0, p1, b0, b1, inv0, inv1, aE0, aW0, aN0, aS0, aE1, aW1, aN1, aS1, ok0, ok1);
  else
    rbsor_passes<NS, 0>(sP, tx, ty, omega, p0, p1, b0, b1, inv0, inv1, aE0, aW0, aN0, aS0, aE1, aW1, aN1, aS1, ok0, ok1);

  // write the tile (cells of the region interior that belong to this CTA)
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const int r = slot_row(ty, k);
    const int gi = i0 + r;
    if (r < H || r >= H + TR || gi >= g.ge) continue;
    if (c0 < H || c0 >= H + TC || gj0 >= g.ny) continue;
    const size_t kk = nf_idx(g, gi, gj0);
    if (gj0 + 1 < g.ny) *reinterpret_cast<double2*>(pout + kk) = make_double2(p0[k], p1[k]);
    else pout[kk] = p0[k];
  }
}

template <int NS>
int launch_fused(nf_ctx* ctx, const nf_grid* g, const double* pin, double* pout, const double* b, const double* d_u,
                 const double* d_v, const double* inv, double omega) {
  constexpr int H = 2 * NS;
  constexpr int TR = RRW - 2 * H, TC = RCW - 2 * H;
  dim3 block(32, NYT, 1);
  dim3 grid((g->ny + TC - 1) / TC, (g->ge - g->gb + TR - 1) / TR, 1);
  if (inv) nf_launch(k_rbsor_fused<NS, true>, grid, block, 0, ctx->stream, true, *g, pin, pout, b, d_u, d_v, inv, omega);
  else nf_launch(k_rbsor_fused<NS, false>, grid, block, 0, ctx->stream, true, *g, pin, pout, b, d_u, d_v, nullptr, omega);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

// ---------------------------------------------------------------------------------------------
// Persistent TMA variant (large levels).  One CTA per SM walks over the tiles in row-major order.  The raw
// inputs of a tile (p, b, d_u, d_v boxes with halo; out-of-domain elements zero-filled by the TMA unit) are
// fetched by four cp.async.bulk.tensor.2d loads into a staging buffer; as soon as the CTA has moved a tile
// from staging into registers it issues the loads of its NEXT tile, so the HBM latency of tile n+1 hides
// behind the 2*NS colour passes of tile n.
// ---------------------------------------------------------------------------------------------
constexpr int ST_P = 0;                               // staging layout (bytes), every box 128-byte aligned
constexpr int ST_B = ST_P + RRW * RCW * 8;            // p, b: [48][64]
constexpr int ST_DU = ST_B + RRW * RCW * 8;           // d_u:  [49][64]  (rows gi .. gi+48)
constexpr int ST_DV = ST_DU + (RRW + 1) * RCW * 8;    // d_v:  [48][66]  (cols gj .. gj+65)
constexpr int ST_INV = ST_DV + RRW * (RCW + 2) * 8;   // 1/aP: [48][64] (only with a precomputed inverse diagonal)
constexpr int ST_END = ST_INV + RRW * RCW * 8;
constexpr int SM_SP = ST_END;                          // sP[2][48][33]
constexpr int SM_BAR = SM_SP + 2 * RRW * 33 * 8;
constexpr int SM_TOTAL = SM_BAR + 16;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c_inner, int c_outer, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      :: "r"(dst), "l"(map), "r"(c_inner), "r"(c_outer), "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(bar), "r"(parity) : "memory");
}

// EXTRA work fused behind the last colour pass (the tile then keeps a slightly deeper halo):
//   0  nothing
//   1  residual norms: sum (b - A p)^2 and sum b^2 over the level (the multigrid convergence test after the
//      post-smoothing, multigrid.py:185-240) -- saves a 40 B/cell pass
//   2  residual + full-weighting restriction: coarse_b = FW(b - A p) (multigrid.py:362-372 after the
//      pre-smoothing) -- saves a 34 B/cell pass.  With ex.in_norm the launch also evaluates the residual norms of its
//      INPUT iterate while the tile is being set up (neighbours are in the halo anyway, no deeper halo needed): the
//      pre-smoother of cycle k+1 thereby delivers the convergence test of cycle k, and the post-smoother runs without
//      extra work (nf_mg.cu, "lookahead norm")
struct TmaExtra {
  nf_grid gc;               // coarse grid (mode 2)
  double* coarse_b;         // mode 2
  double* coarse_x0;        // mode 2, optional: coarse iterate, zeroed along with the restriction
  double* partials;         // mode 1 (and mode 2 with in_norm): per-CTA partial sums, ticket, result (2 doubles)
  unsigned int* ticket;
  double* out;
  int in_norm;              // mode 2: also sum (b - A p_in)^2 and sum b^2 of the INPUT iterate -> out[0..1]
  // Fused prolongation + correction (multigrid.py:405-415): p_in = x + P x_coarse evaluated while the tile is set up
  // (nf_prolong_linear_value per region cell from the coarse array, which is L2 resident on the levels this kernel serves);
  // replaces a k_prolong_linear<ADD> pass.  NULL = off.  Rows [prl_r0, prl_r1) only: what the stand-alone pass covers on a slab.
  const double* prl_c;
  nf_grid prl_gc;
  int prl_r0, prl_r1;
};

template <int NS, int EXTRA>
struct TileGeom {
  static constexpr int HR = (EXTRA == 0) ? 2 * NS : 2 * NS + 1;   // rows of halo below the tile
  static constexpr int HC = (EXTRA == 0) ? 2 * NS : 2 * NS + 2;   // columns of halo left of the tile (even)
  static constexpr int TR = (EXTRA == 0) ? RRW - 4 * NS : (EXTRA == 1 ? RRW - 2 * HR : ((RRW - 2 - HR - 2 * NS) / 2) * 2);
  static constexpr int TC = (EXTRA == 0) ? RCW - 4 * NS : (EXTRA == 1 ? RCW - 2 * HC : ((RCW - 2 - HC - 2 * NS) / 2) * 2);
};

constexpr int SM_SD = SM_BAR + 16;                      // sD[2][48][33]: aP of boundary tiles (EXTRA != 0)
constexpr int SM_SR = SM_SD + 2 * RRW * 33 * 8;         // sR[48][65]: fine residual of the tile (EXTRA == 2)
constexpr int SM_TOTAL_X1 = SM_SR;
constexpr int SM_SN = SM_SR + RRW * 65 * 8;             // sN[512][2]: per-thread sums of the input-residual norms (EXTRA == 2)
constexpr int SM_TOTAL_X2 = SM_SN + 32 * NYT * 2 * 8;

template <int NS, bool HAS_INV, int EXTRA>
__global__ void __launch_bounds__(32 * NYT, 1)
k_rbsor_tma(nf_grid g, const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_du, const __grid_constant__ CUtensorMap map_dv,
            const __grid_constant__ CUtensorMap map_inv, double* __restrict__ pout, double omega, int tiles_x,
            int n_tiles, TmaExtra ex) {
  nf_pdl_entry();
  constexpr unsigned TX_BYTES = HAS_INV ? ST_END : ST_INV;
  using TG = TileGeom<NS, EXTRA>;
  constexpr int HR = TG::HR, HC = TG::HC, TR = TG::TR, TC = TG::TC;
  extern __shared__ __align__(128) unsigned char smem[];
  double (&sP)[2][RRW][33] = *reinterpret_cast<double (*)[2][RRW][33]>(smem + SM_SP);
  double (&sD)[2][RRW][33] = *reinterpret_cast<double (*)[2][RRW][33]>(smem + SM_SD);
  double (&sR)[RRW][65] = *reinterpret_cast<double (*)[RRW][65]>(smem + SM_SR);
  double* sN = reinterpret_cast<double*>(smem + SM_SN);
  const double* stP = reinterpret_cast<const double*>(smem + ST_P);
  const double* stB = reinterpret_cast<const double*>(smem + ST_B);
  const double* stDU = reinterpret_cast<const double*>(smem + ST_DU);
  const double* stDV = reinterpret_cast<const double*>(smem + ST_DV);
  const double* stINV = reinterpret_cast<const double*>(smem + ST_INV);
  const unsigned bar = smem_u32(smem + SM_BAR);
  const unsigned st_base = smem_u32(smem);

  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool leader = (tx == 0 && ty == 0);
  const int c0 = 2 * tx;
  double nrm[2] = {0.0, 0.0};  // EXTRA == 1: sum r^2, sum b^2 over this CTA's tiles

  auto issue = [&](int tile) {
    const int ti = tile / tiles_x, tj = tile - ti * tiles_x;
    const int ri = g.gb + ti * TR - HR, rj = tj * TC - HC;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(TX_BYTES) : "memory");
    tma_load_2d(st_base + ST_P, &map_p, rj, ri - g.row0, bar);
    tma_load_2d(st_base + ST_B, &map_b, rj, ri - g.row0, bar);
    tma_load_2d(st_base + ST_DU, &map_du, rj, ri - g.row0, bar);
    tma_load_2d(st_base + ST_DV, &map_dv, rj, ri - g.row0, bar);
    if (HAS_INV) tma_load_2d(st_base + ST_INV, &map_inv, rj, ri - g.row0, bar);
  };

  if (EXTRA == 2) {
    sN[2 * (ty * 32 + tx)] = 0.0;
    sN[2 * (ty * 32 + tx) + 1] = 0.0;
  }
  if (leader) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int tile = blockIdx.x;
  if (leader && tile < n_tiles) issue(tile);
  unsigned phase = 0;

  for (; tile < n_tiles; tile += gridDim.x) {
    const int ti = tile / tiles_x, tj = tile - ti * tiles_x;
    const int i0 = g.gb + ti * TR - HR;
    const int j0 = tj * TC - HC;
    const int gj0 = j0 + 2 * tx;

    double p0[KS], p1[KS], b0[KS], b1[KS], inv0[KS], inv1[KS];
    double aE0[KS], aW0[KS], aN0[KS], aS0[KS], aE1[KS], aW1[KS], aN1[KS], aS1[KS];
    bool ok0[KS], ok1[KS];

    mbar_wait(bar, phase);
    phase ^= 1;
    // region strictly inside the domain (no boundary row/column, not the pinned cell): branch-free set-up
    const bool interior = (i0 >= 1) && (i0 + RRW < g.nx - 1) && (j0 >= 1) && (j0 + RCW < g.ny - 1);
    if (interior) {
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int r = slot_row(ty, k);
        double2 pp = *reinterpret_cast<const double2*>(stP + r * RCW + c0);
        if (ex.prl_c) {
          const int gi = i0 + r;
          if (gi >= ex.prl_r0 && gi < ex.prl_r1) {
            pp.x = pp.x + nf_prolong_linear_value(ex.prl_gc, ex.prl_c, g.nx, g.ny, gi, gj0);
            pp.y = pp.y + nf_prolong_linear_value(ex.prl_gc, ex.prl_c, g.nx, g.ny, gi, gj0 + 1);
          }
        }
        const double2 bb = *reinterpret_cast<const double2*>(stB + r * RCW + c0);
        const double2 ua = *reinterpret_cast<const double2*>(stDU + r * RCW + c0);
        const double2 ub = *reinterpret_cast<const double2*>(stDU + (r + 1) * RCW + c0);
        const double2 vv = *reinterpret_cast<const double2*>(stDV + r * (RCW + 2) + c0);
        const double w2 = stDV[r * (RCW + 2) + c0 + 2];
        double2 iv = make_double2(1.0, 1.0);
        if (HAS_INV) iv = *reinterpret_cast<const double2*>(stINV + r * RCW + c0);
        const CellCoef ca = cell_coef_interior<HAS_INV>(g, ua.x, ub.x, vv.x, vv.y, iv.x);
        const CellCoef cb = cell_coef_interior<HAS_INV>(g, ua.y, ub.y, vv.y, w2, iv.y);
        aE0[k] = ca.e; aW0[k] = ca.w; aN0[k] = ca.n; aS0[k] = ca.s; inv0[k] = ca.inv;
        aE1[k] = cb.e; aW1[k] = cb.w; aN1[k] = cb.n; aS1[k] = cb.s; inv1[k] = cb.inv;
        p0[k] = pp.x; p1[k] = pp.y; b0[k] = bb.x; b1[k] = bb.y;
        ok0[k] = r >= 1 && r <= RRW - 2 && c0 >= 1;
        ok1[k] = r >= 1 && r <= RRW - 2 && (c0 + 1) <= RCW - 2;
        sP[r & 1][r][tx] = p0[k];
        sP[(r & 1) ^ 1][r][tx] = p1[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int r = slot_row(ty, k);
        const int gi = i0 + r;
        const bool row_in = (gi >= 0 && gi < g.nx);
        const bool in0 = row_in && gj0 >= 0 && gj0 < g.ny;
        const bool in1 = row_in && gj0 + 1 >= 0 && gj0 + 1 < g.ny;
        double2 pp = *reinterpret_cast<const double2*>(stP + r * RCW + c0);
        if (ex.prl_c) {
          if (gi >= ex.prl_r0 && gi < ex.prl_r1) {
            if (in0) pp.x = pp.x + nf_prolong_linear_value(ex.prl_gc, ex.prl_c, g.nx, g.ny, gi, gj0);
            if (in1) pp.y = pp.y + nf_prolong_linear_value(ex.prl_gc, ex.prl_c, g.nx, g.ny, gi, gj0 + 1);
          }
        }
        const double2 bb = *reinterpret_cast<const double2*>(stB + r * RCW + c0);
        const double2 ua = *reinterpret_cast<const double2*>(stDU + r * RCW + c0);
        const double2 ub = *reinterpret_cast<const double2*>(stDU + (r + 1) * RCW + c0);
        const double2 vv = *reinterpret_cast<const double2*>(stDV + r * (RCW + 2) + c0);
        const double w2 = stDV[r * (RCW + 2) + c0 + 2];
        double2 iv = make_double2(1.0, 1.0);
        if (HAS_INV) iv = *reinterpret_cast<const double2*>(stINV + r * RCW + c0);
        CellCoef ca = cell_coef<HAS_INV>(g, gi, gj0, in0, ua.x, ub.x, vv.x, vv.y, iv.x);
        const CellCoef cb = cell_coef<HAS_INV>(g, gi, gj0 + 1, in1, ua.y, ub.y, vv.y, w2, iv.y);
        const bool pinned = (gi == 0 && gj0 == 0);
        if (EXTRA != 0) {  // aP of boundary-tile cells for the residual (identity row at the pinned cell)
          double d0 = in0 ? cell_diag(g, gi, gj0, ua.x, ub.x, vv.x, vv.y) : 1.0;
          const double d1 = in1 ? cell_diag(g, gi, gj0 + 1, ua.y, ub.y, vv.y, w2) : 1.0;
          if (pinned) { d0 = 1.0; ca.e = ca.w = ca.n = ca.s = 0.0; }
          sD[r & 1][r][tx] = d0;
          sD[(r & 1) ^ 1][r][tx] = d1;
        }
        aE0[k] = ca.e; aW0[k] = ca.w; aN0[k] = ca.n; aS0[k] = ca.s; inv0[k] = ca.inv;
        aE1[k] = cb.e; aW1[k] = cb.w; aN1[k] = cb.n; aS1[k] = cb.s; inv1[k] = cb.inv;
        double vp0 = pp.x;
        if (pinned) vp0 = 0.0;
        p0[k] = in0 ? vp0 : 0.0;
        p1[k] = in1 ? pp.y : 0.0;
        b0[k] = bb.x;
        b1[k] = bb.y;
        ok0[k] = in0 && !pinned && r >= 1 && r <= RRW - 2 && c0 >= 1;
        ok1[k] = in1 && r >= 1 && r <= RRW - 2 && (c0 + 1) <= RCW - 2;
        sP[r & 1][r][tx] = p0[k];
        sP[(r & 1) ^ 1][r][tx] = p1[k];
      }
    }
    __syncthreads();  // staging fully consumed, sP complete
    if (leader && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x);

    if (EXTRA == 2) {
      if (ex.in_norm) {
        // residual of the INPUT iterate on this tile's cells, from the registers and the initial sP (same expression
        // order as the stand-alone residual: nf_Ap_cell)
        double s_r = 0.0, s_b = 0.0;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int r = slot_row(ty, k);
          const int par = r & 1;
          const int gi = i0 + r;
          const bool rin = r >= HR && r < HR + TR && gi < g.ge;
          const bool cin0 = c0 >= HC && c0 < HC + TC && gj0 < g.ny;
          const bool cin1 = c0 >= HC && c0 < HC + TC && gj0 + 1 < g.ny;
          if (rin && cin0) {
            const int rmk = r - 1, rpk = r + 1;
            {
              const int lp = par;
              const double d = interior ? ((aE0[k] + aW0[k]) + aN0[k]) + aS0[k] : sD[lp][r][tx];
              double o = d * p0[k];
              o -= aE0[k] * sP[lp ^ 1][rpk][tx];
              o -= aW0[k] * sP[lp ^ 1][rmk][tx];
              o -= aN0[k] * p1[k];
              o -= aS0[k] * sP[lp ^ 1][r][tx > 0 ? tx - 1 : 0];
              const double res = b0[k] - o;
              s_r += res * res;
              s_b += b0[k] * b0[k];
            }
            if (cin1) {
              const int lp = par ^ 1;
              const double d = interior ? ((aE1[k] + aW1[k]) + aN1[k]) + aS1[k] : sD[lp][r][tx];
              double o = d * p1[k];
              o -= aE1[k] * sP[lp ^ 1][rpk][tx];
              o -= aW1[k] * sP[lp ^ 1][rmk][tx];
              o -= aN1[k] * sP[lp ^ 1][r][tx + 1];
              o -= aS1[k] * p0[k];
              const double res = b1[k] - o;
              s_r += res * res;
              s_b += b1[k] * b1[k];
            }
          }
        }
        sN[2 * (ty * 32 + tx)] += s_r;       // own slot, fixed tile order: deterministic
        sN[2 * (ty * 32 + tx) + 1] += s_b;
        __syncthreads();  // the first colour pass overwrites sP rows that neighbouring warps have just read
      }
    }

    if ((i0 + ty + j0) & 1)
      rbsor_passes<NS, 1>(sP, tx, ty, omega, p0, p1, b0, b1, inv0, inv1, aE0, aW0, aN0, aS0, aE1, aW1, aN1, aS1, ok0, ok1);
    else
      rbsor_passes<NS, 0>(sP, tx, ty, omega, p0, p1, b0, b1, inv0, inv1, aE0, aW0, aN0, aS0, aE1, aW1, aN1, aS1, ok0, ok1);

#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int r = slot_row(ty, k);
      const int gi = i0 + r;
      if (r < HR || r >= HR + TR || gi >= g.ge) continue;
      if (c0 < HC || c0 >= HC + TC || gj0 >= g.ny) continue;
      const size_t kk = nf_idx(g, gi, gj0);
      if (gj0 + 1 < g.ny) *reinterpret_cast<double2*>(pout + kk) = make_double2(p0[k], p1[k]);
      else pout[kk] = p0[k];
    }

    if (EXTRA != 0) {
      // residual b - A p of this thread's cells from registers + the final sP (all neighbours of the cells used
      // below are exact: they lie at least 2*NS cells inside the region)
      constexpr int RHI = (EXTRA == 2) ? 1 : 0;  // mode 2 also needs the ring row / column above the tile
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int r = slot_row(ty, k);
        const int par = r & 1;
        const int gi = i0 + r;
        const bool rin = r >= HR && r < HR + TR + RHI && gi < (EXTRA == 2 ? g.nx : g.ge);
        double res0 = 0.0, res1 = 0.0;
        if (rin) {
          const int rmk = r - 1, rpk = r + 1;  // HR >= 1 and r <= RRW-2 here
          {  // cell 0 (column 2tx): N neighbour is the partner, S neighbour belongs to pair tx-1
            const int lp = par;
            const double pc = p0[k];
            const double d = interior ? ((aE0[k] + aW0[k]) + aN0[k]) + aS0[k] : sD[lp][r][tx];
            double o = d * pc;
            o -= aE0[k] * sP[lp ^ 1][rpk][tx];
            o -= aW0[k] * sP[lp ^ 1][rmk][tx];
            o -= aN0[k] * p1[k];
            o -= aS0[k] * sP[lp ^ 1][r][tx > 0 ? tx - 1 : 0];
            res0 = b0[k] - o;
          }
          {  // cell 1 (column 2tx+1)
            const int lp = par ^ 1;
            const double pc = p1[k];
            const double d = interior ? ((aE1[k] + aW1[k]) + aN1[k]) + aS1[k] : sD[lp][r][tx];
            double o = d * pc;
            o -= aE1[k] * sP[lp ^ 1][rpk][tx];
            o -= aW1[k] * sP[lp ^ 1][rmk][tx];
            o -= aN1[k] * sP[lp ^ 1][r][tx + 1];
            o -= aS1[k] * p0[k];
            res1 = b1[k] - o;
          }
        }
        const bool cin0 = c0 >= HC && c0 < HC + TC + RHI && gj0 < g.ny;
        const bool cin1 = c0 + 1 >= HC && c0 + 1 < HC + TC + RHI && gj0 + 1 < g.ny;
        if (EXTRA == 1) {
          if (rin && cin0) { nrm[0] += res0 * res0; nrm[1] += b0[k] * b0[k]; }
          if (rin && cin1) { nrm[0] += res1 * res1; nrm[1] += b1[k] * b1[k]; }
        } else {
          sR[r][c0] = (rin && cin0) ? res0 : 0.0;
          sR[r][c0 + 1] = (rin && cin1) ? res1 : 0.0;
        }
      }
      if (EXTRA == 1) __syncthreads();  // the next tile's set-up overwrites sP rows the neighbouring warps have just read
      if (EXTRA == 2) {
        __syncthreads();
        const int it = g.gb + ti * TR, jt = tj * TC;  // tile origin (even)
        for (int t = ty * 32 + tx; t < (TR / 2) * (TC / 2); t += 32 * NYT) {
          const int ci = t / (TC / 2), cj = t - ci * (TC / 2);
          const int I = it / 2 + ci, J = jt / 2 + cj;
          if (I < ex.gc.gb || I >= ex.gc.ge || J >= ex.gc.ny) continue;
          const int a = 2 * I - i0, q = 2 * J - j0;  // region coordinates of fine (2I, 2J)
          const double cc = sR[a + 1][q + 1], n = sR[a + 1][q + 2], s = sR[a + 1][q], e = sR[a + 2][q + 1], w = sR[a][q + 1];
          const double ne = sR[a + 2][q + 2], nw = sR[a][q + 2], se = sR[a + 2][q], sw = sR[a][q];
          ex.coarse_b[nf_idx(ex.gc, I, J)] = (cc / 4.0 + (((n + s) + e) + w) / 8.0) + (((ne + nw) + se) + sw) / 16.0;
          if (ex.coarse_x0) ex.coarse_x0[nf_idx(ex.gc, I, J)] = 0.0;
        }
        // the next tile's load phase writes sP / sD only; sR is rewritten after its passes (barriers in between)
      }
    }
  }
  if (EXTRA == 1) nf_block_reduce_store<2>(nrm, ex.partials, ex.ticket, ex.out);
  if (EXTRA == 2) {
    if (ex.in_norm) {
      nrm[0] = sN[2 * (ty * 32 + tx)];
      nrm[1] = sN[2 * (ty * 32 + tx) + 1];
      nf_block_reduce_store<2>(nrm, ex.partials, ex.ticket, ex.out);
    }
  }
}

// 2-D fp64 tensor map over `rows` x `cols` valid elements with row pitch ld; box = box_rows x box_cols.  Cached per
// (array, shape) in nf_rbsor_stream.cu: encoding is a driver call per array, and a smoothing call used to pay five of them
// per launch.
bool make_map(CUtensorMap* map, const double* base, int rows, int cols, int ld, int box_rows, int box_cols) {
  return nfi_tensor_map_2d(map, base, rows, cols, ld, box_cols, box_rows);
}

template <int NS, bool HAS_INV, int EXTRA>
int launch_tma(nf_ctx* ctx, const nf_grid* g, const double* pin, double* pout, const double* b, const double* d_u,
               const double* d_v, const double* inv, double omega, const TmaExtra& ex, bool* used) {
  using TG = TileGeom<NS, EXTRA>;
  constexpr int TR = TG::TR, TC = TG::TC;
  constexpr int SMEM = EXTRA == 0 ? SM_TOTAL : (EXTRA == 1 ? SM_TOTAL_X1 : SM_TOTAL_X2);
  *used = false;
  const int tiles_x = (g->ny + TC - 1) / TC, tiles_y = (g->ge - g->gb + TR - 1) / TR;
  const int n_tiles = tiles_x * tiles_y;
  // rows the arrays hold (slab runs store [row0, row1); row1 == 0 means the whole grid)
  const int row_end = g->row1 > 0 ? g->row1 : g->nx + 1;
  const int stored_p = (row_end < g->nx ? row_end : g->nx) - g->row0;  // p-like arrays
  const int stored_u = row_end - g->row0;                               // d_u (face rows)
  CUtensorMap mp, mb, mu, mv, mi;
  if (!make_map(&mp, pin, stored_p, g->ny, g->ld, RRW, RCW) || !make_map(&mb, b, stored_p, g->ny, g->ld, RRW, RCW) ||
      !make_map(&mu, d_u, stored_u, g->ny, g->ld, RRW + 1, RCW) ||
      !make_map(&mv, d_v, stored_p, g->ny + 1, g->ld, RRW, RCW + 2) ||
      !make_map(&mi, HAS_INV ? inv : b, stored_p, g->ny, g->ld, RRW, RCW))
    return NF_OK;  // caller falls back to the plain fused kernel
  static bool attr_set = false;
  if (!attr_set) {
    NF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_rbsor_tma<NS, HAS_INV, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            SMEM));
    attr_set = true;
  }
  const int grid = n_tiles < NF_SM_COUNT ? n_tiles : NF_SM_COUNT;
  nf_launch(k_rbsor_tma<NS, HAS_INV, EXTRA>, grid, dim3(32, NYT, 1), SMEM, ctx->stream, true, *g, mp, mb, mu, mv, mi, pout, omega,
                                                                                 tiles_x, n_tiles, ex);
  NF_LAUNCH_CHECK(ctx);
  *used = true;
  return NF_OK;
}

}  // namespace

// n_sweeps red-black SOR sweeps, double buffered: *p holds the input, *palt is scratch of the same shape; on
// return *p points at the buffer holding the result (the two pointers are swapped once per launch).
template <int NS>
int launch_tma_any(nf_ctx* ctx, const nf_grid* g, const double* pin, double* pout, const double* b, const double* d_u,
                   const double* d_v, const double* inv, double omega, int extra, const TmaExtra& ex, bool* used) {
  if (inv && extra == 1) return launch_tma<NS, true, 1>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
  if (inv && extra == 2) return launch_tma<NS, true, 2>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
  if (inv) return launch_tma<NS, true, 0>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
  return launch_tma<NS, false, 0>(ctx, g, pin, pout, b, d_u, d_v, nullptr, omega, ex, used);
}

// The streaming kernel serves levels of >= NF_RBSOR_STREAM rows per side (default 1500: below that a strip's row chunks get
// so short that the 12 halo rows per chunk outweigh the lower column redundancy) whose slab starts on an even row.
bool nfi_rbsor_stream_enabled(const nf_grid* g) {
  const char* env = getenv("NF_RBSOR_STREAM");
  const int min_rows = env ? atoi(env) : 1500;
  return g->nx >= min_rows && (g->ge - g->gb) >= 16 && (g->gb % 2) == 0 && (g->ld % 2) == 0;
}

// Can a smoothing call of n_sweeps on this level take the prolongation of the coarse correction into its first launch?
// (streaming kernel, a 3-sweep first launch; NF_MG_PROLONG_FUSED=0 switches it off)
// Returns 0: no (stand-alone pass); 1: streaming kernel (block rule at load + a strips-only launch of k_prolong_linear for
// the ring / trailing cells); 2: TMA kernel (every cell at tile set-up, no launch at all).
constexpr int NF_SMOOTH_ROWS_BEYOND = 8;  // == NF_HALO (nf_slab.cuh): rows around a slab the stand-alone prolongation covers
static bool tma_smoother_enabled(const nf_grid* g) {
  const char* env = getenv("NF_RBSOR_TMA");
  // NF_RBSOR_TMA=rows: minimum level size (a huge value disables the kernel).  Default: every level of >= 32
  // rows.  Round 1 used it from 600 rows up "for the pipeline"; what pays on the small levels is the FUSION that comes with
  // it (residual + restriction / norms inside the smoother launch instead of a 13-17 us latency-bound kernel of their own):
  // 15.25 -> 14.83 ms per outer iteration at 4097^2 (levels 511, 255, 127)
  const int tma_min_rows = env ? atoi(env) : 32;
  return g->nx >= tma_min_rows && (g->ge - g->gb) >= 32;  // (the levels of <= 31 rows are the single-kernel coarse end's)
}
int nfi_rbsor_can_fuse_prolong(const nf_grid* g, int n_sweeps, bool has_inv) {
  const char* env = getenv("NF_MG_PROLONG_FUSED");
  if (env && env[0] == '0') return 0;
  if (!has_inv || n_sweeps < 1) return 0;
  if (nfi_rbsor_stream_enabled(g)) return n_sweeps >= 3 ? 1 : 0;
  if (env && env[0] == '1') return 0;  // NF_MG_PROLONG_FUSED=1: streaming kernel only (round-2 behaviour before the TMA form)
  return tma_smoother_enabled(g) ? 2 : 0;
}

// inv (optional): precomputed 1/aP of this level (nfi_inv_diag); NULL = divide inside the kernel
int nfi_rbsor_fused_x(nf_ctx* ctx, const nf_grid* g, double** p, double** palt, const double* b, const double* d_u,
                      const double* d_v, const double* inv, double omega, int n_sweeps, nf_smooth_extra* extra) {
  if (extra) { extra->fused = false; extra->in_norm_fused = false; }
  if (n_sweeps == 0) {  // the reference still pins p[0,0] = 0 (gauss_seidel.py:145)
    if (g->row0 == 0 && g->gb == 0) NF_CHECK_CUDA(ctx, cudaMemsetAsync(*p, 0, sizeof(double), ctx->stream));
    return NF_OK;
  }
  // TMA path: 16-byte aligned arrays, pitch a multiple of 2 doubles, enough rows for a tile
  const bool use_tma = tma_smoother_enabled(g);
  const char* envx = getenv("NF_RBSOR_EXTRA");
  const bool allow_extra = !(envx && envx[0] == '0');
  int left = n_sweeps;
  while (left > 0) {
    const int ns = left >= 3 ? 3 : left;
    int st = NF_OK;
    bool used = false;
    TmaExtra ex;
    ex.coarse_b = nullptr; ex.coarse_x0 = nullptr; ex.partials = ctx->partials; ex.ticket = ctx->ticket; ex.out = nullptr;
    ex.in_norm = 0;
    ex.gc = *g;
    ex.prl_c = nullptr; ex.prl_gc = *g; ex.prl_r0 = 0; ex.prl_r1 = 0;
    int mode = 0;
    // extra work rides on the last launch; it needs the precomputed 1/aP, an unsplit grid and even tile origins
    if (extra && extra->mode != 0 && allow_extra && left == ns && use_tma && inv &&
        (extra->mode == 1 || (g->gb % 2) == 0)) {
      mode = extra->mode;
      ex.gc = extra->gc;
      ex.coarse_b = extra->coarse_b;
      ex.coarse_x0 = extra->coarse_x_zero;
      ex.out = extra->out;
      if (mode == 2 && extra->in_norm_out && left == n_sweeps) {  // the launch that sees the call's input iterate
        ex.in_norm = 1;
        ex.out = extra->in_norm_out;
      }
    }
    // streaming (wavefront) kernel (nf_rbsor_stream.cu) on large levels; its fused modes ride on 3-sweep launches and do not
    // deliver the input norms (the multigrid driver then runs the classic convergence test behind the post-smoother)
    const bool want_prl = extra && extra->prolong_c != nullptr && left == n_sweeps;  // rides on the FIRST launch
    if (nfi_rbsor_stream_enabled(g) && inv) {
      int smode = 0;
      if (extra && extra->mode != 0 && allow_extra && left == ns && ns == 3 && (g->gb % 2) == 0) smode = extra->mode;
      nf_smooth_extra sx;
      if (extra) sx = *extra;
      if (!want_prl) sx.prolong_c = nullptr;
      NF_TRY(nfi_rbsor_stream(ctx, g, *p, *palt, b, d_u, d_v, inv, omega, ns, smode, extra ? &sx : nullptr, &used));
      if (used && smode != 0) extra->fused = true;
      if (used && want_prl) extra->prolong_fused = true;
      if (used) mode = 0;
    }
    if (want_prl && !used && !(use_tma && inv)) {
      ctx->err = "fused prolongation was requested for a launch neither the streaming nor the TMA smoother serves";
      return NF_ERR_UNSUPPORTED;
    }
    if (want_prl && !used) {  // TMA kernel: p_in = x + P x_coarse at tile set-up, rows the stand-alone pass covers on a slab
      ex.prl_c = extra->prolong_c;
      ex.prl_gc = extra->prolong_gc;
      const int lo = g->gb - NF_SMOOTH_ROWS_BEYOND, hi = g->ge + NF_SMOOTH_ROWS_BEYOND;
      const bool cut = !(g->gb == 0 && g->ge == g->nx);
      ex.prl_r0 = cut ? (lo > 0 ? lo : 0) : 0;
      ex.prl_r1 = cut ? (hi < g->nx ? hi : g->nx) : g->nx;
    }
    if (use_tma && !used) {  // persistent TMA pipeline: pays off once every SM gets several tiles
      if (ns == 3) st = launch_tma_any<3>(ctx, g, *p, *palt, b, d_u, d_v, inv, omega, mode, ex, &used);
      else if (ns == 2) st = launch_tma_any<2>(ctx, g, *p, *palt, b, d_u, d_v, inv, omega, mode, ex, &used);
      else st = launch_tma_any<1>(ctx, g, *p, *palt, b, d_u, d_v, inv, omega, mode, ex, &used);
      if (st != NF_OK) return st;
      if (used && mode != 0) extra->fused = true;
      if (used && mode == 2 && ex.in_norm) extra->in_norm_fused = true;
      if (want_prl) {
        if (!used) {
          ctx->err = "fused prolongation: the TMA smoother could not be launched (tensor-map encoder missing?)";
          return NF_ERR_UNSUPPORTED;
        }
        extra->prolong_fused = true;
      }
    }
    if (!used) {
      if (ns == 3) st = launch_fused<3>(ctx, g, *p, *palt, b, d_u, d_v, inv, omega);
      else if (ns == 2) st = launch_fused<2>(ctx, g, *p, *palt, b, d_u, d_v, inv, omega);
      else st = launch_fused<1>(ctx, g, *p, *palt, b, d_u, d_v, inv, omega);
      if (st != NF_OK) return st;
    }
    double* t = *p; *p = *palt; *palt = t;
    left -= ns;
  }
  return NF_OK;
}

int nfi_rbsor_fused(nf_ctx* ctx, const nf_grid* g, double** p, double** palt, const double* b, const double* d_u,
                    const double* d_v, const double* inv, double omega, int n_sweeps) {
  return nfi_rbsor_fused_x(ctx, g, p, palt, b, d_u, d_v, inv, omega, n_sweeps, nullptr);
}

// 1/aP of every cell (gauss_seidel.py:214-266: aP with the Neumann folding, < 1e-15 -> 1), once per level and
// per (d_u, d_v): 24 B/cell, saves the two fp64 divisions per cell pair in every smoother launch
__global__ void k_inv_diag(nf_grid g, const double* __restrict__ d_u, const double* __restrict__ d_v,
                           double* __restrict__ inv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.ny || i >= g.ge) return;
  const PCoef c = nf_pcoef(g, d_u, d_v, i, j);
  double aP = c.diag;
  if (aP < 1e-15) aP = 1.0;
  inv[nf_idx(g, i, j)] = 1.0 / aP;
}

int nfi_inv_diag(nf_ctx* ctx, const nf_grid* g, const double* d_u, const double* d_v, double* inv) {
  NfLaunch2D l = nf_launch2d(g->ge - g->gb, g->ny);
  k_inv_diag<<<l.grid, l.block, 0, ctx->stream>>>(*g, d_u, d_v, inv);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

// C-ABI: in-place semantics with a caller-provided scratch array
extern "C" int nf_pressure_inv_diag(nf_ctx* ctx, const nf_grid* g, const double* d_u, const double* d_v, double* inv) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, d_u && d_v && inv, "NULL argument");
  return nfi_inv_diag(ctx, g, d_u, d_v, inv);
}

extern "C" int nf_rbsor_sweeps_fused(nf_ctx* ctx, const nf_grid* g, double* p, double* tmp, const double* b,
                                     const double* d_u, const double* d_v, const double* inv, double omega,
                                     int n_sweeps) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, n_sweeps >= 0, "n_sweeps < 0");
  NF_REQUIRE(ctx, p && tmp && p != tmp, "p and tmp must be distinct arrays");
  NF_REQUIRE(ctx, (g->ld % 2) == 0, "row pitch must be even (16-byte aligned pair loads)");
  NF_REQUIRE(ctx, ((uintptr_t)p % 16) == 0 && ((uintptr_t)tmp % 16) == 0 && ((uintptr_t)b % 16) == 0 &&
                      ((uintptr_t)d_u % 16) == 0 && ((uintptr_t)d_v % 16) == 0, "arrays must be 16-byte aligned");
  double* cur = p;
  double* alt = tmp;
  NF_REQUIRE(ctx, !inv || ((uintptr_t)inv % 16) == 0, "inv must be 16-byte aligned");
  NF_TRY(nfi_rbsor_fused(ctx, g, &cur, &alt, b, d_u, d_v, inv, omega, n_sweeps));
  if (cur != p) {
    const size_t rows = (size_t)(g->ge - g->gb);
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(p + (size_t)(g->gb - g->row0) * g->ld, cur + (size_t)(g->gb - g->row0) * g->ld,
                                       rows * g->ld * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return NF_OK;
}
