// Host build of the per-cell function of csrc/nf_links_ext.cuh (the same source the CUDA kernel compiles): lets the CPU
// test-suite check its arithmetic against the reference's golden outputs without a GPU.  TEST INFRASTRUCTURE.
#include "../../naviflow_b200/csrc/nf_links_ext.cuh"

template <int SCHEME, int IS_U>
static void run(const NfxGrid& g, const double* u, const double* v, const double* p, double mu, int sides, double* out) {
  const int rows = g.nx + (IS_U ? 1 : 0), cols = g.ny + (IS_U ? 0 : 1);
  const long plane = (long)(g.nx + 1) * g.ld;
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) {
      double o[NFX_COUNT];
      nfx_cell<SCHEME, IS_U>(g, u, v, p, mu, sides, i, j, o);
      for (int q = 0; q < NFX_COUNT; ++q) out[q * plane + nfx_idx(g, i, j)] = o[q];
    }
}

// fields: (nx+1) x ld row-major doubles; out: NFX_COUNT planes of the same shape, order a_e a_w a_n a_s a_ee a_ww a_nn a_ss a_p src
extern "C" void host_links_ext(int scheme, int is_u, int nx, int ny, int ld, double dx, double dy, double rho, double mu,
                               int sides, const double* u, const double* v, const double* p, double* out) {
  NfxGrid g;
  g.nx = nx; g.ny = ny; g.ld = ld; g.row0 = 0; g.dx = dx; g.dy = dy; g.rho = rho;
  if (scheme == NFX_SCHEME_QUICK) {
    if (is_u) run<NFX_SCHEME_QUICK, 1>(g, u, v, p, mu, sides, out);
    else run<NFX_SCHEME_QUICK, 0>(g, u, v, p, mu, sides, out);
  } else {
    if (is_u) run<NFX_SCHEME_SOU, 1>(g, u, v, p, mu, sides, out);
    else run<NFX_SCHEME_SOU, 0>(g, u, v, p, mu, sides, out);
  }
}
