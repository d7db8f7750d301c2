// nf_links_ext.cu -- QUICK / second-order upwind momentum links on the device (nf_links_ext.cuh has the per-cell function
// and the reference citations).  One thread per cell, ten coalesced stores; 3 input arrays read through L1/L2 (each value is
// shared by the stencils of its neighbours): 24 B read + 80 B written per cell, HBM bound.
#include "nf_common.cuh"
#include "nf_links_ext.cuh"

namespace {

template <int SCHEME, int IS_U>
__global__ void k_links_ext(nf_grid g, const double* __restrict__ u, const double* __restrict__ v,
                            const double* __restrict__ p, double mu, int sides, nf_links_ext out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = g.gb + blockIdx.y * blockDim.y + threadIdx.y;
  const int rows = g.nx + (IS_U ? 1 : 0), cols = g.ny + (IS_U ? 0 : 1);
  const int row_end = (IS_U && g.ge == g.nx) ? g.nx + 1 : g.ge;
  if (j >= cols || i >= row_end || i >= rows) return;
  NfxGrid x;
  x.nx = g.nx; x.ny = g.ny; x.ld = g.ld; x.row0 = g.row0; x.dx = g.dx; x.dy = g.dy; x.rho = g.rho;
  double o[NFX_COUNT];
  nfx_cell<SCHEME, IS_U>(x, u, v, p, mu, sides, i, j, o);
  const size_t k = nf_idx(g, i, j);
  out.a_e[k] = o[NFX_E];   out.a_w[k] = o[NFX_W];   out.a_n[k] = o[NFX_N];   out.a_s[k] = o[NFX_S];
  out.a_ee[k] = o[NFX_EE]; out.a_ww[k] = o[NFX_WW]; out.a_nn[k] = o[NFX_NN]; out.a_ss[k] = o[NFX_SS];
  out.a_p[k] = o[NFX_P];   out.src[k] = o[NFX_SRC];
}

template <int SCHEME, int IS_U>
int launch(nf_ctx* ctx, const nf_grid* g, const double* u, const double* v, const double* p, double mu, int sides,
           nf_links_ext out) {
  const int rows = ((IS_U && g->ge == g->nx) ? g->nx + 1 : g->ge) - g->gb, cols = g->ny + (IS_U ? 0 : 1);
  const dim3 block(64, 4, 1), grid((cols + 63) / 64, (rows + 3) / 4, 1);
  k_links_ext<SCHEME, IS_U><<<grid, block, 0, ctx->stream>>>(*g, u, v, p, mu, sides, out);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

}  // namespace

// C-ABI: QUICKDiscretization / SecondOrderUpwindDiscretization .calculate_u_coefficients / .calculate_v_coefficients
// (quick.py:27-196, second_order_upwind.py:46-325).  u_bc, v_bc: velocities as the caller passes them to the reference
// (boundary values in the edge lines); sides as in nf_momentum_links_u (0 = the reference's bc=None).  Single slab.
extern "C" int nf_momentum_links_ext(nf_ctx* ctx, const nf_grid* g, int is_u, int scheme, const double* u_bc, const double* v_bc,
                                     const double* p, double mu, int sides, nf_links_ext out) {
  NF_GRID_OK(ctx, g);
  NF_REQUIRE(ctx, u_bc && v_bc && p, "NULL field");
  NF_REQUIRE(ctx, out.a_e && out.a_w && out.a_n && out.a_s && out.a_ee && out.a_ww && out.a_nn && out.a_ss && out.a_p && out.src,
             "NULL coefficient array");
  NF_REQUIRE(ctx, scheme == NF_SCHEME_QUICK || scheme == NF_SCHEME_SOU, "scheme must be NF_SCHEME_QUICK or NF_SCHEME_SOU");
  NF_REQUIRE(ctx, g->row0 == 0 && g->gb == 0 && g->ge == g->nx, "single-slab grids only");
  NF_REQUIRE(ctx, g->nx >= 3 && g->ny >= 3, "grid too small for the extended stencil");
  if (scheme == NF_SCHEME_QUICK)
    return is_u ? launch<NFX_SCHEME_QUICK, 1>(ctx, g, u_bc, v_bc, p, mu, sides, out)
                : launch<NFX_SCHEME_QUICK, 0>(ctx, g, u_bc, v_bc, p, mu, sides, out);
  return is_u ? launch<NFX_SCHEME_SOU, 1>(ctx, g, u_bc, v_bc, p, mu, sides, out)
              : launch<NFX_SCHEME_SOU, 0>(ctx, g, u_bc, v_bc, p, mu, sides, out);
}
