// nf_links_ext.cuh -- extended-stencil momentum links: QUICK and second-order (linear) upwind.
//
// Reference: solver/momentum_solver/discretization/quick.py:27-219 (QUICKDiscretization) and
// second_order_upwind.py:26-325 (SecondOrderUpwindDiscretization): ten coefficient arrays per component (a_e, a_w, a_n,
// a_s, a_ee, a_ww, a_nn, a_ss, a_p, source) on the staggered u (nx+1, ny) / v (nx, ny+1) grids, Practice-B folding of the
// boundary-adjacent lines.  SURVEY.md 8(f) rank 4.
//
// The reference builds every array by a sequence of masked `+=` statements over the whole grid.  Here a scheme is a TABLE
// of accumulation steps {target, face, flux part, coefficient, diffusion term, stencil mask} evaluated per cell; the steps
// of one target keep the reference's statement order, so every sum rounds like the NumPy expression (the library is built
// with -fmad=false).  The per-cell function compiles for the device (nf_links_ext.cu) and for the host: the CPU test-suite
// builds it with g++ and checks it against the reference's outputs without a GPU (tests/test_links_ext_host.py).
#pragma once

#if defined(__CUDACC__)
#define NFX_HD __host__ __device__ __forceinline__
#else
#define NFX_HD inline
#endif

enum { NFX_E = 0, NFX_W, NFX_N, NFX_S, NFX_EE, NFX_WW, NFX_NN, NFX_SS, NFX_P, NFX_SRC, NFX_COUNT };
enum { NFX_SCHEME_QUICK = 1, NFX_SCHEME_SOU = 2 };

struct NfxGrid {  // what the per-cell function needs of nf_grid
  int nx, ny, ld, row0;
  double dx, dy, rho;
};

struct NfxStep {
  int dst;      // NFX_E .. NFX_P
  int face;     // 0 e, 1 w, 2 n, 3 s
  int part;     // 0: max(F, 0), 1: max(-F, 0), 2: no convective term
  int diff;     // 0 none, 1: + De, 2: + Dn
  int mask;     // 0 always, 1: cell has an EE neighbour, 2: WW, 3: NN, 4: SS
  double coef;
};

// quick.py:62-112 (u) and :149-194 (v): the same statements for both components
#define NFX_QUICK_FACE(A1, A2, A3, FACE, D, M)                                                                           \
  {A1, FACE, 0, D, M, 0.75}, {NFX_P, FACE, 0, 0, M, 0.375}, {A2, FACE, 0, 0, M, -0.125}, {A1, FACE, 1, D, M, 0.375},         \
  {NFX_P, FACE, 1, 0, M, 0.75}, {A3, FACE, 1, 0, M, -0.125}

NFX_HD double nfx_pos(double f) { return f > 0.0 ? f : 0.0; }   // np.maximum(F, 0.0) for finite F
NFX_HD double nfx_neg(double f) { return -f > 0.0 ? -f : 0.0; }  // np.maximum(-F, 0.0)

NFX_HD long nfx_idx(const NfxGrid& g, int i, int j) { return (long)(i - g.row0) * (long)g.ld + (long)j; }

// All ten values of cell (i, j) of component IS_U (1: u at (nx+1, ny), 0: v at (nx, ny+1)); sides: 1 left, 2 right,
// 4 bottom, 8 top (boundaries with a registered condition).  Cells outside the scheme's interior block are zero, as in
// the reference (np.zeros), except for what Practice B adds to them (0 * boundary value).
template <int SCHEME, int IS_U>
NFX_HD void nfx_cell(const NfxGrid& g, const double* u, const double* v, const double* p, double mu, int sides, int i, int j,
                     double (&out)[NFX_COUNT]) {
  const int nx = g.nx, ny = g.ny;
  for (int q = 0; q < NFX_COUNT; ++q) out[q] = 0.0;
  const bool active = IS_U ? (i >= 1 && i <= nx - 1 && j >= 1 && j <= ny - 2) : (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 1);
  if (active) {
    const double cdy = 0.5 * g.rho * g.dy, cdx = 0.5 * g.rho * g.dx;
    double F[4];
    if (IS_U) {  // quick.py:56-59, second_order_upwind.py:81-84
      F[0] = cdy * (u[nfx_idx(g, i + 1, j)] + u[nfx_idx(g, i, j)]);
      F[1] = cdy * (u[nfx_idx(g, i - 1, j)] + u[nfx_idx(g, i, j)]);
      F[2] = cdx * (v[nfx_idx(g, i, j + 1)] + v[nfx_idx(g, i - 1, j + 1)]);
      F[3] = cdx * (v[nfx_idx(g, i, j)] + v[nfx_idx(g, i - 1, j)]);
    } else {     // quick.py:142-145, second_order_upwind.py:228-231
      F[0] = cdy * (u[nfx_idx(g, i + 1, j)] + u[nfx_idx(g, i + 1, j - 1)]);
      F[1] = cdy * (u[nfx_idx(g, i, j)] + u[nfx_idx(g, i, j - 1)]);
      F[2] = cdx * (v[nfx_idx(g, i, j + 1)] + v[nfx_idx(g, i, j)]);
      F[3] = cdx * (v[nfx_idx(g, i, j)] + v[nfx_idx(g, i, j - 1)]);
    }
    const double De = mu * g.dy / g.dx, Dn = mu * g.dx / g.dy;
    // which second neighbours exist (QUICK only: quick.py:63, :78, :93, :103 / :149-150, :171-172)
    const bool has[5] = {true, i <= (IS_U ? nx - 2 : nx - 3), i >= 2, j <= (IS_U ? ny - 3 : ny - 2), j >= 2};
    constexpr NfxStep quick[] = {NFX_QUICK_FACE(NFX_E, NFX_EE, NFX_W, 0, 1, 1), NFX_QUICK_FACE(NFX_W, NFX_WW, NFX_E, 1, 1, 2),
                                 NFX_QUICK_FACE(NFX_N, NFX_NN, NFX_S, 2, 2, 3), NFX_QUICK_FACE(NFX_S, NFX_SS, NFX_N, 3, 2, 4)};
    // second_order_upwind.py:88-127 (u): diffusion, then the faces E, W, N, S
    constexpr NfxStep sou_u[] = {
        {NFX_E, 0, 2, 1, 0, 0.0}, {NFX_W, 0, 2, 1, 0, 0.0}, {NFX_N, 0, 2, 2, 0, 0.0}, {NFX_S, 0, 2, 2, 0, 0.0},
        {NFX_P, 0, 0, 0, 0, 1.5}, {NFX_W, 0, 0, 0, 0, 0.5}, {NFX_WW, 0, 0, 0, 0, -0.5}, {NFX_E, 0, 1, 0, 0, 1.5}, {NFX_EE, 0, 1, 0, 0, 0.5},
        {NFX_W, 1, 0, 0, 0, 1.5}, {NFX_WW, 1, 0, 0, 0, -0.5}, {NFX_P, 1, 1, 0, 0, 1.5}, {NFX_E, 1, 1, 0, 0, 0.5},
        {NFX_P, 2, 0, 0, 0, 1.5}, {NFX_S, 2, 0, 0, 0, -0.5}, {NFX_N, 2, 1, 0, 0, 1.5}, {NFX_NN, 2, 1, 0, 0, 0.5},
        {NFX_S, 3, 0, 0, 0, 1.5}, {NFX_SS, 3, 0, 0, 0, -0.5}, {NFX_P, 3, 1, 0, 0, 1.5}, {NFX_N, 3, 1, 0, 0, 0.5}};
    // second_order_upwind.py:234-266 (v)
    constexpr NfxStep sou_v[] = {
        {NFX_E, 0, 2, 1, 0, 0.0}, {NFX_W, 0, 2, 1, 0, 0.0}, {NFX_N, 0, 2, 2, 0, 0.0}, {NFX_S, 0, 2, 2, 0, 0.0},
        {NFX_E, 0, 0, 0, 0, 1.5}, {NFX_EE, 0, 0, 0, 0, 0.5}, {NFX_P, 0, 1, 0, 0, 1.5}, {NFX_W, 0, 1, 0, 0, 0.5},
        {NFX_P, 1, 0, 0, 0, 1.5}, {NFX_E, 1, 0, 0, 0, 0.5}, {NFX_W, 1, 1, 0, 0, 1.5}, {NFX_WW, 1, 1, 0, 0, 0.5},
        {NFX_N, 2, 0, 0, 0, 1.5}, {NFX_NN, 2, 0, 0, 0, 0.5}, {NFX_P, 2, 1, 0, 0, 1.5}, {NFX_S, 2, 1, 0, 0, 0.5},
        {NFX_P, 3, 0, 0, 0, 1.5}, {NFX_N, 3, 0, 0, 0, 0.5}, {NFX_S, 3, 1, 0, 0, 1.5}, {NFX_SS, 3, 1, 0, 0, 0.5}};
    const NfxStep* steps = SCHEME == NFX_SCHEME_QUICK ? quick : (IS_U ? sou_u : sou_v);
    constexpr int n_steps = SCHEME == NFX_SCHEME_QUICK ? 24 : (IS_U ? 21 : 20);
    static_assert(sizeof(quick) / sizeof(NfxStep) == 24 && sizeof(sou_u) / sizeof(NfxStep) == 21 &&
                      sizeof(sou_v) / sizeof(NfxStep) == 20, "table sizes");
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 0; s < n_steps; ++s) {
      const NfxStep st = steps[s];
      if (!has[st.mask]) continue;
      const double D = st.diff == 1 ? De : Dn;
      double term;
      if (st.part == 2) {
        term = D;
      } else {
        term = st.coef * (st.part == 0 ? nfx_pos(F[st.face]) : nfx_neg(F[st.face]));
        if (st.diff != 0) term = term + D;
      }
      out[st.dst] = out[st.dst] + term;
    }
    // pressure gradient (quick.py:115, :196; second_order_upwind.py:130, :269)
    out[NFX_SRC] = IS_U ? (p[nfx_idx(g, i - 1, j)] - p[nfx_idx(g, i, j)]) * g.dy : (p[nfx_idx(g, i, j - 1)] - p[nfx_idx(g, i, j)]) * g.dx;
    if (SCHEME == NFX_SCHEME_SOU) {  // base diagonal with the flux imbalance (second_order_upwind.py:133-143, :272-283)
      double sum = out[NFX_E] + out[NFX_W];
      sum = sum + out[NFX_N];
      sum = sum + out[NFX_S];
      sum = sum + out[NFX_EE];
      sum = sum + out[NFX_WW];
      sum = sum + out[NFX_NN];
      sum = sum + out[NFX_SS];
      sum = sum + (F[0] - F[1]);
      sum = sum + (F[2] - F[3]);
      out[NFX_P] = out[NFX_P] + sum;
    }
  }
  // Practice B (quick.py:199-219, second_order_upwind.py:150-181, :286-309): the link towards a boundary value moves into
  // the source, in the reference's order of the sides -- u: left, right, bottom, top; v: bottom, top, left, right
  if (IS_U) {
    if ((sides & 1) && i == 1) { out[NFX_SRC] = out[NFX_SRC] + out[NFX_W] * u[nfx_idx(g, 0, j)]; out[NFX_W] = 0.0; }
    if ((sides & 2) && i == nx - 1) { out[NFX_SRC] = out[NFX_SRC] + out[NFX_E] * u[nfx_idx(g, nx, j)]; out[NFX_E] = 0.0; }
    if ((sides & 4) && j == 1 && i >= 1 && i <= nx - 1) { out[NFX_SRC] = out[NFX_SRC] + out[NFX_S] * u[nfx_idx(g, i, 0)]; out[NFX_S] = 0.0; }
    if ((sides & 8) && j == ny - 2 && i >= 1 && i <= nx - 1) { out[NFX_SRC] = out[NFX_SRC] + out[NFX_N] * u[nfx_idx(g, i, ny - 1)]; out[NFX_N] = 0.0; }
  } else {
    if ((sides & 4) && j == 1) { out[NFX_SRC] = out[NFX_SRC] + out[NFX_S] * v[nfx_idx(g, i, 0)]; out[NFX_S] = 0.0; }
    if ((sides & 8) && j == ny - 1) { out[NFX_SRC] = out[NFX_SRC] + out[NFX_N] * v[nfx_idx(g, i, ny)]; out[NFX_N] = 0.0; }
    if ((sides & 1) && i == 1 && j >= 1 && j <= ny - 1) { out[NFX_SRC] = out[NFX_SRC] + out[NFX_W] * v[nfx_idx(g, 0, j)]; out[NFX_W] = 0.0; }
    if ((sides & 2) && i == nx - 2 && j >= 1 && j <= ny - 1) { out[NFX_SRC] = out[NFX_SRC] + out[NFX_E] * v[nfx_idx(g, nx - 1, j)]; out[NFX_E] = 0.0; }
  }
}
