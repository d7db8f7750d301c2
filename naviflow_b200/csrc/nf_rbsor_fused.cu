// nf_rbsor_fused.cu -- temporally blocked red-black SOR (K8): NS full sweeps (2*NS colour passes) per
// tile load.  Same arithmetic as gauss_seidel.py:214-305 / k_rbsor_color and bit-identical to it: the update
// of a cell depends only on its four opposite-colour neighbours after the previous colour pass, which the
// trapezoid scheme reproduces exactly inside the tile.
//
// Layout of the work
//   * One CTA (32 x 16 threads) owns an output tile of TR x TC cells and loads the region grown by the halo
//     H = 2*NS on every side (48 rows x 64 columns).  After colour pass t the cells at distance > t from the
//     region edge are exact, so after 2*NS passes the tile is exact.
//   * Thread (tx, ty) owns the column pair (2tx, 2tx+1) of region rows ty, ty+16, ty+32: one red and one
//     black cell per row.  Their constants (b, 1/aP, aE, aW, aN, aS) live in REGISTERS for all passes; only p
//     goes through shared memory, stored split by colour so that every access is lane-contiguous
//     (conflict-free): sP[parity][row][pair].
//   * HBM traffic per launch: read p, b, d_u, d_v (32 B/cell x region/tile overhead, largely absorbed by L2
//     between neighbouring tiles) + write p (8 B/cell) -- versus 2*NS passes x ~40 B/cell unfused.
//   * p is double buffered in global memory (p_in -> p_out): neighbouring tiles read each other's halo.
#include "nf_pressure.cuh"

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
__device__ __forceinline__ CellCoef cell_coef(const nf_grid& g, int gi, int gj, bool inside, double du_c, double du_e,
                                              double dv_c, double dv_n) {
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
  diag += ((e + w) + n) + s;
  if (diag < 1e-15) diag = 1.0;
  c.e = e; c.w = w; c.n = n; c.s = s;
  c.inv = 1.0 / diag;
  return c;
}

template <int NS>
__global__ void __launch_bounds__(32 * NYT, 1)
k_rbsor_fused(nf_grid g, const double* __restrict__ pin, double* __restrict__ pout, const double* __restrict__ b,
              const double* __restrict__ d_u, const double* __restrict__ d_v, double omega) {
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
    const int r = ty + NYT * k;
    const int gi = i0 + r;
    const bool row_in = (gi >= 0 && gi < g.nx);
    const bool in0 = row_in && gj0 >= 0 && gj0 < g.ny;
    const bool in1 = row_in && gj0 + 1 >= 0 && gj0 + 1 < g.ny;
    double vp0 = 0.0, vp1 = 0.0, vb0 = 0.0, vb1 = 0.0;
    double uc0 = 0.0, uc1 = 0.0, ue0 = 0.0, ue1 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0;
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
    }
    const CellCoef ca = cell_coef(g, gi, gj0, in0, uc0, ue0, w0, w1);
    const CellCoef cb = cell_coef(g, gi, gj0 + 1, in1, uc1, ue1, w1, w2);
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

#pragma unroll
  for (int t = 0; t < 2 * NS; ++t) {
    const int col = t & 1;            // 0: (i+j) even ("red"), 1: odd ("black")
    const int m = 2 * NS - 1 - t;     // pass t is only needed within m cells of the tile
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int r = ty + NYT * k;
      const int gi = i0 + r;
      const int s = ((gi + gj0) & 1) ^ col;  // which cell of the pair has colour `col`: 0 -> cell 0, 1 -> cell 1
      const int lp = (r & 1) ^ s;            // its local parity
      const bool ok = s ? ok1[k] : ok0[k];
      if (ok && r >= H - m && r < H + TR + m) {
        const double pc = s ? p1[k] : p0[k];
        const double bc = s ? b1[k] : b0[k];
        const double ic = s ? inv1[k] : inv0[k];
        const double aE = s ? aE1[k] : aE0[k];
        const double aW = s ? aW1[k] : aW0[k];
        const double aN = s ? aN1[k] : aN0[k];
        const double aS = s ? aS1[k] : aS0[k];
        // opposite-colour neighbours: rows r+-1 from shared memory; in the row, one is the thread's own partner
        // cell (register) and the other belongs to the neighbouring pair
        const double pE = sP[lp ^ 1][r + 1][tx];
        const double pW = sP[lp ^ 1][r - 1][tx];
        const double pN = s ? sP[lp ^ 1][r][tx + 1] : p1[k];
        const double pS = s ? p0[k] : sP[lp ^ 1][r][tx - 1];
        double acc = bc;       // ((((b + E) + W) + N) + S) * (1/aP): gauss_seidel.py:285-299
        acc += aE * pE;
        acc += aW * pW;
        acc += aN * pN;
        acc += aS * pS;
        const double pn = acc * ic;
        const double pnew = pc + omega * (pn - pc);
        if (s) p1[k] = pnew; else p0[k] = pnew;
        sP[lp][r][tx] = pnew;
      }
    }
    __syncthreads();
  }

  // write the tile (cells of the region interior that belong to this CTA)
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const int r = ty + NYT * k;
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
                 const double* d_v, double omega) {
  constexpr int H = 2 * NS;
  constexpr int TR = RRW - 2 * H, TC = RCW - 2 * H;
  dim3 block(32, NYT, 1);
  dim3 grid((g->ny + TC - 1) / TC, (g->ge - g->gb + TR - 1) / TR, 1);
  k_rbsor_fused<NS><<<grid, block, 0, ctx->stream>>>(*g, pin, pout, b, d_u, d_v, omega);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

}  // namespace

// n_sweeps red-black SOR sweeps, double buffered: *p holds the input, *palt is scratch of the same shape; on
// return *p points at the buffer holding the result (the two pointers are swapped once per launch).
int nfi_rbsor_fused(nf_ctx* ctx, const nf_grid* g, double** p, double** palt, const double* b, const double* d_u,
                    const double* d_v, double omega, int n_sweeps) {
  if (n_sweeps == 0) {  // the reference still pins p[0,0] = 0 (gauss_seidel.py:145)
    if (g->row0 == 0 && g->gb == 0) NF_CHECK_CUDA(ctx, cudaMemsetAsync(*p, 0, sizeof(double), ctx->stream));
    return NF_OK;
  }
  int left = n_sweeps;
  while (left > 0) {
    const int ns = left >= 3 ? 3 : left;
    int st;
    if (ns == 3) st = launch_fused<3>(ctx, g, *p, *palt, b, d_u, d_v, omega);
    else if (ns == 2) st = launch_fused<2>(ctx, g, *p, *palt, b, d_u, d_v, omega);
    else st = launch_fused<1>(ctx, g, *p, *palt, b, d_u, d_v, omega);
    if (st != NF_OK) return st;
    double* t = *p; *p = *palt; *palt = t;
    left -= ns;
  }
  return NF_OK;
}

// C-ABI: in-place semantics with a caller-provided scratch array
extern "C" int nf_rbsor_sweeps_fused(nf_ctx* ctx, const nf_grid* g, double* p, double* tmp, const double* b,
                                     const double* d_u, const double* d_v, double omega, int n_sweeps) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, n_sweeps >= 0, "n_sweeps < 0");
  NF_REQUIRE(ctx, p && tmp && p != tmp, "p and tmp must be distinct arrays");
  NF_REQUIRE(ctx, (g->ld % 2) == 0, "row pitch must be even (16-byte aligned pair loads)");
  NF_REQUIRE(ctx, ((uintptr_t)p % 16) == 0 && ((uintptr_t)tmp % 16) == 0 && ((uintptr_t)b % 16) == 0 &&
                      ((uintptr_t)d_u % 16) == 0 && ((uintptr_t)d_v % 16) == 0, "arrays must be 16-byte aligned");
  double* cur = p;
  double* alt = tmp;
  NF_TRY(nfi_rbsor_fused(ctx, g, &cur, &alt, b, d_u, d_v, omega, n_sweeps));
  if (cur != p) {
    const size_t rows = (size_t)(g->ge - g->gb);
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(p + (size_t)(g->gb - g->row0) * g->ld, cur + (size_t)(g->gb - g->row0) * g->ld,
                                       rows * g->ld * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return NF_OK;
}
