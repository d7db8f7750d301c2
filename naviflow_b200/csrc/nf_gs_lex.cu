// nf_gs_lex.cu -- lexicographic ("standard") and symmetric Gauss-Seidel / SOR sweeps of the pressure-correction
// system, fp64.
//
// Reference: GaussSeidelSolver._standard_gauss_seidel_step / _symmetric_gauss_seidel_step
// (solver/pressure_solver/gauss_seidel.py:307-367 of /root/reference/naviflow_oo): `for j: for i:` over all cells except the
// pinned (0,0), each update using the newest values -- west (i-1,j) and south (i,j-1) already updated in this sweep, east
// and north still old -- with p_new = ((((b + aE pE) + aW pW) + aN pN) + aS pS) / aP, p += omega (p_new - p); p[0,0] = 0
// after the sweep.  The symmetric variant appends the same sweep in reverse order.
//
// A sequential sweep has exactly one dependency structure: cell (i,j) needs the new (i-1,j), (i,j-1) and the old (i+1,j),
// (i,j+1).  Every order that respects it produces the same bits, in particular the anti-diagonal wavefront i+j = const.
// Two levels of it here: the grid is cut into 32 x 32 blocks; blocks on one block anti-diagonal are independent and run
// as the CTAs of one launch (launches follow the block diagonals); inside a block one warp walks the 63 cell diagonals,
// lane t owning row t, with the block's p (+ halo) and its six coefficient planes staged in shared memory.
#include "nf_pressure.cuh"

namespace {

constexpr int GB = 32;  // block edge

struct GsSmem {
  double p[GB + 2][GB + 3];  // block + halo ring
  double e[GB][GB + 1], w[GB][GB + 1], n[GB][GB + 1], s[GB][GB + 1], inv[GB][GB + 1], b[GB][GB + 1];
};

// bd: block anti-diagonal (bi + bj); backward: reverse sweep (cells and blocks in descending order)
__global__ void __launch_bounds__(256) k_gs_lex_block(nf_grid g, double* __restrict__ p, const double* __restrict__ b,
                                                      const double* __restrict__ d_u, const double* __restrict__ d_v,
                                                      double omega, int bd, int bi_min, int backward) {
  extern __shared__ __align__(16) unsigned char raw[];
  GsSmem& S = *reinterpret_cast<GsSmem*>(raw);
  const int bi = bi_min + blockIdx.x, bj = bd - bi;
  const int i0 = bi * GB, j0 = bj * GB;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  // ---- stage the block: p with halo (0 outside the domain), coefficients and right-hand side
  for (int t = tid; t < (GB + 2) * (GB + 2); t += 256) {
    const int li = t / (GB + 2), lj = t - li * (GB + 2);
    const int gi = i0 + li - 1, gj = j0 + lj - 1;
    S.p[li][lj] = (gi >= 0 && gi < g.nx && gj >= 0 && gj < g.ny) ? p[nf_idx(g, gi, gj)] : 0.0;
  }
  for (int t = tid; t < GB * GB; t += 256) {
    const int li = t / GB, lj = t - li * GB;
    const int gi = i0 + li, gj = j0 + lj;
    double ce = 0.0, cw = 0.0, cn = 0.0, cs = 0.0, ci = 0.0, cb = 0.0;
    if (gi < g.nx && gj < g.ny) {
      const PCoef c = nf_pcoef(g, d_u, d_v, gi, gj);  // gauss_seidel.py:214-266 (= matrix_free.py folding)
      double aP = c.diag;
      if (aP < 1e-15) aP = 1.0;
      ce = c.e; cw = c.w; cn = c.n; cs = c.s;
      ci = 1.0 / aP;
      cb = b[nf_idx(g, gi, gj)];
    }
    S.e[li][lj] = ce; S.w[li][lj] = cw; S.n[li][lj] = cn; S.s[li][lj] = cs; S.inv[li][lj] = ci; S.b[li][lj] = cb;
  }
  __syncthreads();
  // ---- wavefront over the block's 2*GB-1 cell diagonals (warp 0; lane t owns local row t)
  if (threadIdx.y == 0) {
    const int t = threadIdx.x;
    for (int d = 0; d < 2 * GB - 1; ++d) {
      int li = t, lj = d - t;
      if (backward) { li = GB - 1 - t; lj = GB - 1 - (d - t); }
      const int gi = i0 + li, gj = j0 + lj;
      if (lj >= 0 && lj < GB && gi < g.nx && gj < g.ny && !(gi == 0 && gj == 0)) {
        const double pc = S.p[li + 1][lj + 1];
        double acc = S.b[li][lj];
        acc += (gi < g.nx - 1) ? S.e[li][lj] * S.p[li + 2][lj + 1] : 0.0;
        acc += (gi > 0) ? S.w[li][lj] * S.p[li][lj + 1] : 0.0;
        acc += (gj < g.ny - 1) ? S.n[li][lj] * S.p[li + 1][lj + 2] : 0.0;
        acc += (gj > 0) ? S.s[li][lj] * S.p[li + 1][lj] : 0.0;
        const double pn = acc * S.inv[li][lj];
        S.p[li + 1][lj + 1] = pc + omega * (pn - pc);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // ---- write the block back
  for (int t = tid; t < GB * GB; t += 256) {
    const int li = t / GB, lj = t - li * GB;
    const int gi = i0 + li, gj = j0 + lj;
    if (gi < g.nx && gj < g.ny) p[nf_idx(g, gi, gj)] = S.p[li + 1][lj + 1];
  }
}

__global__ void k_pin_zero(double* p) { *p = 0.0; }

}  // namespace

// n_sweeps lexicographic sweeps (symmetric != 0: each followed by the reverse sweep); single-slab grids
int nfi_gs_lex(nf_ctx* ctx, const nf_grid* g, double* p, const double* b, const double* d_u, const double* d_v,
               double omega, int n_sweeps, int symmetric) {
  static bool attr_set = false;
  if (!attr_set) {
    NF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_gs_lex_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GsSmem)));
    attr_set = true;
  }
  const int nbi = (g->nx + GB - 1) / GB, nbj = (g->ny + GB - 1) / GB;
  k_pin_zero<<<1, 1, 0, ctx->stream>>>(p);  // p[0,0] = 0 before the first sweep (gauss_seidel.py:145)
  NF_LAUNCH_CHECK(ctx);
  for (int s = 0; s < n_sweeps; ++s) {
    for (int pass = 0; pass < (symmetric ? 2 : 1); ++pass) {
      for (int q = 0; q < nbi + nbj - 1; ++q) {
        const int bd = pass == 0 ? q : (nbi + nbj - 2 - q);
        const int lo = bd - (nbj - 1) > 0 ? bd - (nbj - 1) : 0;
        const int hi = bd < nbi - 1 ? bd : nbi - 1;
        k_gs_lex_block<<<hi - lo + 1, dim3(32, 8, 1), sizeof(GsSmem), ctx->stream>>>(*g, p, b, d_u, d_v, omega, bd, lo, pass);
        NF_LAUNCH_CHECK(ctx);
      }
    }
    k_pin_zero<<<1, 1, 0, ctx->stream>>>(p);  // p[0,0] = 0 after the (symmetric) iteration (:340, :367)
    NF_LAUNCH_CHECK(ctx);
  }
  return NF_OK;
}

extern "C" int nf_gs_lex_sweeps(nf_ctx* ctx, const nf_grid* g, double* p, const double* b, const double* d_u,
                                const double* d_v, double omega, int n_sweeps, int symmetric) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, p && b && d_u && d_v, "NULL argument");
  NF_REQUIRE(ctx, n_sweeps >= 0, "n_sweeps < 0");
  NF_REQUIRE(ctx, g->row0 == 0 && g->gb == 0 && g->ge == g->nx, "sequential sweeps run on single-slab grids only");
  return nfi_gs_lex(ctx, g, p, b, d_u, d_v, omega, n_sweeps, symmetric);
}
