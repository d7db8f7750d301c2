// nf_pressure.cuh -- device helpers of the matrix-free pressure-correction operator, shared by the
// smoothers, the multigrid transfer kernels and the Krylov solvers.
#pragma once
#include "nf_common.cuh"

// The five pressure coefficients of cell (i,j) with the reference's Neumann folding
// (matrix_free.py:52-84): the boundary-direction coefficient is added to the diagonal FIRST and
// then zeroed, so boundary cells are decoupled from their interior neighbour in that direction.
struct PCoef {
  double e, w, n, s, diag;
};

__device__ __forceinline__ PCoef nf_pcoef(const nf_grid& g, const double* __restrict__ d_u,
                                          const double* __restrict__ d_v, int i, int j) {
  PCoef c;
  const size_t k = nf_idx(g, i, j);
  // aE = rho*d_u[i+1,j]*dy (i<nx-1); aW = rho*d_u[i,j]*dy (i>0); aN = rho*d_v[i,j+1]*dx; aS = rho*d_v[i,j]*dx
  c.e = (i < g.nx - 1) ? g.rho * d_u[k + g.ld] * g.dy : 0.0;
  c.w = (i > 0) ? g.rho * d_u[k] * g.dy : 0.0;
  c.n = (j < g.ny - 1) ? g.rho * d_v[k + 1] * g.dx : 0.0;
  c.s = (j > 0) ? g.rho * d_v[k] * g.dx : 0.0;
  double diag = 0.0;
  if (i == 0) { diag += c.e; }
  if (i == g.nx - 1) { diag += c.w; }
  if (j == 0) { diag += c.n; }
  if (j == g.ny - 1) { diag += c.s; }
  if (i == 0) c.e = 0.0;
  if (i == g.nx - 1) c.w = 0.0;
  if (j == 0) c.n = 0.0;
  if (j == g.ny - 1) c.s = 0.0;
  diag += ((c.e + c.w) + c.n) + c.s;
  c.diag = diag;
  return c;
}

// A*p at cell (i,j): diag*p - E*pE - W*pW - N*pN - S*pS in that order (matrix_free.py:100-121);
// identity row at the pinned cell (0,0).
__device__ __forceinline__ double nf_Ap_cell(const nf_grid& g, const double* __restrict__ p,
                                             const double* __restrict__ d_u, const double* __restrict__ d_v,
                                             int i, int j) {
  const size_t k = nf_idx(g, i, j);
  const double pc = p[k];
  if (i >= 1 && i < g.nx - 1 && j >= 1 && j < g.ny - 1) {
    // interior cell: no folding, no masks (same expression order, so the same bits as the general path)
    const double e = g.rho * d_u[k + g.ld] * g.dy, w = g.rho * d_u[k] * g.dy;
    const double n = g.rho * d_v[k + 1] * g.dx, s = g.rho * d_v[k] * g.dx;
    double out = (((e + w) + n) + s) * pc;
    out -= e * p[k + g.ld];
    out -= w * p[k - g.ld];
    out -= n * p[k + 1];
    out -= s * p[k - 1];
    return out;
  }
  if (i == 0 && j == 0) return pc;
  const PCoef c = nf_pcoef(g, d_u, d_v, i, j);
  double out = c.diag * pc;
  if (i < g.nx - 1) out -= c.e * p[k + g.ld];
  if (i > 0) out -= c.w * p[k - g.ld];
  if (j < g.ny - 1) out -= c.n * p[k + 1];
  if (j > 0) out -= c.s * p[k - 1];
  return out;
}

__device__ __forceinline__ double nf_jacobi_diag_cell(const nf_grid& g, const double* __restrict__ d_u,
                                                      const double* __restrict__ d_v, int i, int j) {
  const size_t k = nf_idx(g, i, j);
  double d = 0.0;
  if (i < g.nx - 1) d += g.rho * d_u[k + g.ld] * g.dy;
  if (i > 0) d += g.rho * d_u[k] * g.dy;
  if (j < g.ny - 1) d += g.rho * d_v[k + 1] * g.dx;
  if (j > 0) d += g.rho * d_v[k] * g.dx;
  if (i == 0) d += d;
  if (i == g.nx - 1) d += d;
  if (j == 0) d += d;
  if (j == g.ny - 1) d += d;
  if (d < 1e-15) d = 1.0;
  if (i == 0 && j == 0) d = 1.0;
  return d;
}


// Value of interpolate_linear (multigrid_helpers.py:116-186) at fine cell (i,j) of an mx x my grid from the coarse array c:
// coarse k -> fine 2k+1, even points averaged, the outer ring copied from ring 1 (clamped indices); grids of <= 3 cells only
// get the coincident points (:127-128).
__device__ __forceinline__ double nf_prolong_linear_value(const nf_grid& gc, const double* __restrict__ c,
                                                          int mx, int my, int i, int j) {
  const int mcx = gc.nx, mcy = gc.ny;
  if (mx <= 3 || my <= 3) {
    if ((i & 1) && (j & 1) && (i - 1) / 2 < mcx && (j - 1) / 2 < mcy) return c[nf_idx(gc, (i - 1) / 2, (j - 1) / 2)];
    return 0.0;
  }
  i = min(max(i, 1), mx - 2);
  j = min(max(j, 1), my - 2);
  const bool io = i & 1, jo = j & 1;
  const int I = io ? (i - 1) / 2 : (i - 2) / 2;
  const int J = jo ? (j - 1) / 2 : (j - 2) / 2;
  const bool iok = io ? (I < mcx) : (I <= mcx - 2);
  const bool jok = jo ? (J < mcy) : (J <= mcy - 2);
  if (!iok || !jok) return 0.0;
  const size_t k = nf_idx(gc, I, J);
  if (io && jo) return c[k];
  if (io && !jo) return 0.5 * (c[k] + c[k + 1]);
  if (!io && jo) return 0.5 * (c[k] + c[k + gc.ld]);
  return 0.25 * (((c[k] + c[k + gc.ld]) + c[k + 1]) + c[k + gc.ld + 1]);
}

