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
// Ring of TMA boxes per warp: NSTG stages of RB rows each (RB * NSTG == 8: the stage of a step is a compile-time function of
// its phase).  One box per array and stage: the TMA unit retires roughly one op per 46 cycles per SM whatever its size, so
// single-row boxes (5 ops per step and warp) made the unit the bottleneck.
constexpr int RB = 4, NSTG = 2;
static_assert(RB * NSTG == 8, "ring length must equal the phase count");
constexpr int SG_P = 0, SG_DU = RB * 512, SG_B = 2 * RB * 512, SG_INV = 3 * RB * 512, SG_DV = 4 * RB * 512;  // stage layout
constexpr int SG_CX = ((4 * RB * 512 + RB * 66 * 8 + 127) / 128) * 128;  // coarse iterate (fused prolongation): 3 rows x 34
// coarse rows / columns of a stage's box: 34 columns are needed; the box starts on an EVEN column (measured on B200: a
// fp64 box whose inner start coordinate is odd -- a byte offset that is not a multiple of 16 -- raises "illegal instruction",
// tools/dev/tma_box_test2.cu), hence 36
constexpr int CXR = RB / 2 + 1, CXC = SW / 2 + 4;
constexpr int SG_BYTES = ((SG_CX + CXR * CXC * 8 + 127) / 128) * 128;
constexpr unsigned SG_TX = RB * (4 * 512 + 66 * 8);
constexpr unsigned SG_TX_CX = CXR * CXC * 8;

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

// Work fused into a launch (see stream_step2 / stream_load)
struct StreamExtra {
  nf_grid gc;          // EXTRA 2: coarse grid and its right-hand side
  double* coarse_b;
  double* coarse_x0;   // EXTRA 2, optional: coarse iterate, zeroed along with the restriction
  nf_grid pgc;         // PRL: grid and array of the coarse iterate whose bilinear prolongation is added to p on the way in
  const double* pcx;
  int prl_nI, prl_nJ;  // PRL: the block rule covers fine rows 1 .. 2 nI, columns 1 .. 2 nJ; the other cells (ring, trailing
                       // cells of even-sized grids) were prolonged by the strips launch of k_prolong_linear
  double* partials;    // EXTRA 1: per-job partial sums, ticket, result (sum r^2, sum b^2)
  unsigned int* ticket;
  double* out;
};

// Domain edges.  The reference zeroes the links of a boundary cell across its edge direction but keeps the reverse links
// (matrix_free.py:63-84), which breaks the face sharing at exactly four places: row 0's aE, row nx-1's aW, column 0's aN,
// column ny-1's aS.  Everything else is settled when a row is loaded: the faces that only boundary or virtual cells use
// (d_u rows 0 and nx, d_v columns 0 and ny -- NaN in real runs: the momentum solver divides by a zero aP there) are replaced
// by 0, and cells outside the domain then keep p = 0 on their own (all their operands are zero-filled by the TMA unit), so no
// update needs a validity predicate except the pinned cell (0,0).  ROWB / COLB: the variant handles row / column edges;
// only the 8-step groups that touch rows 0 / nx-1 run ROWB, only the first and last strip run COLB.
struct ColFlags {
  bool A_first, A_last, B_last;  // the lane's cell A is column 0 / ny-1, its cell B is column ny-1 (column 0 is always an A)
  bool z0, z1, z2;               // face j = jA, jA+1, jA+2 is column 0 or ny: zero at load
};

// Loads row s of p and d_u (-> slot SN) and the coefficient rows of row s-1 (-> slot SN-1) from row RW of a stage.
// p of the second cell goes to *pB_out (the double step commits it late, see below).
// PRL: p_in = x + P x_coarse (interpolate_linear, multigrid_helpers.py:116-186; the prolongation + correction of the V-cycle,
// multigrid.py:405-415, fused into the post-smoother's load -- saves an 18 B/cell pass): the block rule of
// k_prolong_linear evaluated from the coarse rows the stage carries.  The cells the block rule does not cover (ring copy,
// trailing cells of even-sized grids) are prolonged beforehand by the strips launch of k_prolong_linear (add = 2).
template <int SN, int RW, bool ROWB, bool COLB, bool PRL>
__device__ __forceinline__ void stream_load(Win& W, const nf_grid& g, int s, int lane, const unsigned char* stage,
                                            const ColFlags& cf, double* pB_out, const StreamExtra& ex, int gjA) {
  constexpr int SC = (SN + 7) & 7;
  double2 pp = *reinterpret_cast<const double2*>(stage + SG_P + RW * 512 + 16 * lane);
  if (PRL) {
    double vA, vB;
    // the box starts at the even column <= j0/2 - 1; lane l needs coarse columns j0/2 - 1 + l and + l + 1
    const double* cx = reinterpret_cast<const double*>(stage + SG_CX) + (((gjA >> 1) - lane - 1) & 1);
    if ((RW & 1) == 0) {  // even fine row 2I+2: between coarse rows I = RW/2 and I+1 of the stage's box
      const double c00 = cx[(RW / 2) * CXC + lane], c01 = cx[(RW / 2) * CXC + lane + 1];
      const double c10 = cx[(RW / 2 + 1) * CXC + lane], c11 = cx[(RW / 2 + 1) * CXC + lane + 1];
      vA = 0.25 * (((c00 + c10) + c01) + c11);
      vB = 0.5 * (c01 + c11);
    } else {              // odd fine row 2I+1: coarse row I = (RW+1)/2
      const double c00 = cx[((RW + 1) / 2) * CXC + lane], c01 = cx[((RW + 1) / 2) * CXC + lane + 1];
      vA = 0.5 * (c00 + c01);
      vB = c01;
    }
    if (ROWB || COLB) {  // cells outside the block rule (and outside the domain) get nothing here
      const bool rin = s >= 1 && s <= 2 * ex.prl_nI;
      if (!(rin && gjA >= 1 && gjA <= 2 * ex.prl_nJ)) vA = 0.0;
      if (!(rin && gjA + 1 >= 1 && gjA + 1 <= 2 * ex.prl_nJ)) vB = 0.0;
    }
    pp.x = pp.x + vA;
    pp.y = pp.y + vB;
  }
  const double2 uu = *reinterpret_cast<const double2*>(stage + SG_DU + RW * 512 + 16 * lane);
  const double2 bb = *reinterpret_cast<const double2*>(stage + SG_B + RW * 512 + 16 * lane);
  const double2 iv = *reinterpret_cast<const double2*>(stage + SG_INV + RW * 512 + 16 * lane);
  const double2 vv = *reinterpret_cast<const double2*>(stage + SG_DV + RW * 528 + 16 * lane);
  const double v2 = *reinterpret_cast<const double*>(stage + SG_DV + RW * 528 + 16 * lane + 16);
  W.pA[SN] = pp.x;
  *pB_out = pp.y;
  if (ROWB && COLB) {
    if (s == 0 && cf.A_first) W.pA[SN] = 0.0;  // pinned cell (gauss_seidel.py:145, :305)
  }
  double fa = g.rho * uu.x * g.dy, fb = g.rho * uu.y * g.dy;
  if (ROWB) {
    if (s == 0 || s == g.nx) { fa = 0.0; fb = 0.0; }
  }
  W.fA[SN] = fa;
  W.fB[SN] = fb;
  double g0 = g.rho * vv.x * g.dx, g1 = g.rho * vv.y * g.dx, g2 = g.rho * v2 * g.dx;
  if (COLB) {
    if (cf.z0) g0 = 0.0;
    if (cf.z1) g1 = 0.0;
    if (cf.z2) g2 = 0.0;
  }
  W.g0[SC] = g0;
  W.g1[SC] = g1;
  W.g2[SC] = g2;
  W.bA[SC] = bb.x; W.bB[SC] = bb.y;
  W.iA[SC] = iv.x; W.iB[SC] = iv.y;
}

// One SOR update of the cell UPD_A ? A : B of the row in slot a (global row r); nbv = the neighbour across the pair boundary.
template <bool UPD_A, int a, bool ROWB, bool COLB>
__device__ __forceinline__ void stream_update(Win& W, const nf_grid& g, int r, double omega, double nbv, const ColFlags& cf) {
  constexpr int se = (a + 1) & 7, sw = (a + 7) & 7;  // rows r+1, r-1
  double aE, aW, aN, aS, pc, pE, pW, pN, pS, bc, ic;
  if (UPD_A) {
    aE = W.fA[se]; aW = W.fA[a]; aN = W.g1[a]; aS = W.g0[a];
    pc = W.pA[a]; pE = W.pA[se]; pW = W.pA[sw]; pN = W.pB[a]; pS = nbv;
    bc = W.bA[a]; ic = W.iA[a];
  } else {
    aE = W.fB[se]; aW = W.fB[a]; aN = W.g2[a]; aS = W.g1[a];
    pc = W.pB[a]; pE = W.pB[se]; pW = W.pB[sw]; pN = nbv; pS = W.pA[a];
    bc = W.bB[a]; ic = W.iB[a];
  }
  if (ROWB) {  // row 0 keeps no E link, row nx-1 no W link (their other vertical face is zero since the load)
    if (r == 0) aE = 0.0;
    if (r == g.nx - 1) aW = 0.0;
  }
  if (COLB) {  // column 0 keeps no N link, column ny-1 no S link (their outer faces are zero since the load)
    if (UPD_A) { if (cf.A_first) aN = 0.0; if (cf.A_last) aS = 0.0; } else { if (cf.B_last) aS = 0.0; }
  }
  double acc = bc;  // ((((b + E) + W) + N) + S) * (1/aP): gauss_seidel.py:285-299
  acc += aE * pE;
  acc += aW * pW;
  acc += aN * pN;
  acc += aS * pS;
  const double pn = acc * ic;
  const double pu = pc + omega * (pn - pc);
  if (UPD_A) {
    if (ROWB && COLB) { if (!(r == 0 && cf.A_first)) W.pA[a] = pu; }
    else W.pA[a] = pu;
  } else {
    W.pB[a] = pu;
  }
}

// b - A p of the lane's two cells of a FINAL row r (nf_Ap_cell's expression order: diag*p - E - W - N - S with the reference's
// diagonal folding at the edges, matrix_free.py:63-121).  Faces are the raw ones (the faces outside the domain are zero
// since the load); nS / nN = the in-row neighbours across the pair boundary.

template <bool ROWB, bool COLB>
__device__ __forceinline__ void stream_residual(const nf_grid& g, int r, const ColFlags& cf, double pWA, double pWB,
                                                double pcA, double pcB, double pEA, double pEB, double fWA, double fWB,
                                                double fEA, double fEB, double g0, double g1, double g2, double bA,
                                                double bB, double nS, double nN, double* rA, double* rB) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    double e = q ? fEB : fEA, w = q ? fWB : fWA, n = q ? g2 : g1, s = q ? g1 : g0;
    const double pc = q ? pcB : pcA, pE = q ? pEB : pEA, pW = q ? pWB : pWA;
    const double pN = q ? nN : pcB, pS = q ? pcA : nS, b = q ? bB : bA;
    double diag;
    if (ROWB || COLB) {
      diag = 0.0;
      const bool first = COLB && q == 0 && cf.A_first, last = COLB && (q ? cf.B_last : cf.A_last);
      if (ROWB) { if (r == 0) diag += e; if (r == g.nx - 1) diag += w; }
      if (COLB) { if (first) diag += n; if (last) diag += s; }
      if (ROWB) { if (r == 0) e = 0.0; if (r == g.nx - 1) w = 0.0; }
      if (COLB) { if (first) n = 0.0; if (last) s = 0.0; }
      diag += ((e + w) + n) + s;
    } else {
      diag = ((e + w) + n) + s;
    }
    double o = diag * pc;
    o -= e * pE;
    o -= w * pW;
    o -= n * pN;
    o -= s * pS;
    double res = b - o;
    if (ROWB && COLB) { if (q == 0 && r == 0 && cf.A_first) res = b - pc; }  // identity row of the pinned cell
    if (q) *rB = res; else *rA = res;
  }
}

// Iteration T of the interleaved chains of a double step (see stream_step2): X_T, then the shuffle Y_(T+1) will need, then
// Y_(T-1).
template <int PH, int T, int NP, bool ROWB, bool COLB>
__device__ __forceinline__ void stream_pair(Win& W, const nf_grid& g, int s, double omega, const ColFlags& cf,
                                            const double (&nbX)[NP], double (&nbY)[NP]) {
  if constexpr (T < NP) {
    constexpr int ax = (PH + 15 - T) & 7;  // row s-1-T
    stream_update<false, ax, ROWB, COLB>(W, g, s - 1 - T, omega, nbX[T], cf);
    if constexpr (T + 1 < NP) nbY[T + 1] = __shfl_up_sync(0xffffffffu, W.pB[ax], 1);
  }
  if constexpr (T >= 1) {
    constexpr int ay = (PH + 17 - T) & 7;  // row s-(T-1)
    stream_update<true, ay, ROWB, COLB>(W, g, s - (T - 1), omega, nbY[T - 1], cf);
  }
  if constexpr (T < NP) stream_pair<PH, T + 1, NP, ROWB, COLB>(W, g, s, omega, cf, nbX, nbY);
}

// Two steps of the wavefront at once: step s (even phase PH: the B cells, chain X) and step s+1 (the A cells, chain Y).
// X_t = pass t on row s-1-t, Y_t = pass t on row s-t.  Y_t needs X_(t-1) (its in-row neighbours) and Y_(t-1); X_t needs only
// X_(t-1) and values of the previous double step -- so the statements are emitted as X_0, {X_1, Y_0}, {X_2, Y_1}, ... , Y_last
// and each brace holds two independent dependency chains for the scheduler to interleave (a single step is one serial chain
// of 2NS x 9 fp64 operations; with two warps per scheduler that left the issue slots 60 % empty).
// EXTRA work on the rows that have become final (both need NP == 6; the halo is two columns / rows deeper):
//   1  sum (b - A p)^2, sum b^2 over the level (the multigrid convergence test, multigrid.py:185-240) -> acc[0..1]
//   2  coarse_b = FW(b - A p) (multigrid.py:362-372 after the pre-smoothing); acc[0..1] carries the residual of the
//      previous even row.  A double step finalises rows s-6 and s-5, so the residual rows s-7 and s-6 are complete.
template <int PH, int NP, bool ROWB, bool COLB, int EXTRA, bool PRL>
__device__ __forceinline__ void stream_step2(Win& W, const nf_grid& g, const Job& jb, int s, int lane, double omega,
                                             const unsigned char* stage, const ColFlags& cf, double* __restrict__ pout,
                                             double (&acc)[2], const StreamExtra& ex) {
  static_assert((PH & 1) == 0, "double steps start on even phases");
  static_assert(EXTRA == 0 || NP == 6, "the fused residual work rides on 3-sweep launches");
  constexpr int S0 = PH, S1 = (PH + 1) & 7;  // slots of rows s, s+1
  constexpr int HL = EXTRA ? NP + 2 : NP;     // halo columns on each side of the strip
  double pB_s1;
  // rows s-8 (p) and s-7 (p of cell A, faces) leave the window with the loads below; the residual of row s-7 needs them
  double o8A = 0.0, o8B = 0.0, o7A = 0.0, o7fA = 0.0, o7fB = 0.0;
  if (EXTRA != 0) { o8A = W.pA[S0]; o8B = W.pB[S0]; o7A = W.pA[S1]; o7fA = W.fA[S1]; o7fB = W.fB[S1]; }
  const int c = 2 * lane;
  const int gjA = jb.j0 + c;
  stream_load<S0, PH % RB, ROWB, COLB, PRL>(W, g, s, lane, stage, cf, &W.pB[S0], ex, gjA);  // slot of row s-8: free
  // pB of row s+1 shares its slot with row s-7, which chain X still reads (W neighbour of its last update): committed below
  stream_load<S1, (PH + 1) % RB, ROWB, COLB, PRL>(W, g, s + 1, lane, stage, cf, &pB_s1, ex, gjA);
  double nbX[NP], nbY[NP];
#pragma unroll
  for (int t = 0; t < NP; ++t) nbX[t] = __shfl_down_sync(0xffffffffu, W.pA[(PH + 15 - t) & 7], 1);  // rows s-1-t
  nbY[0] = __shfl_up_sync(0xffffffffu, W.pB[S0], 1);                                                // row s
  stream_pair<PH, 0, NP, ROWB, COLB>(W, g, s, omega, cf, nbX, nbY);
  if (EXTRA != 0) {
    constexpr int S2 = (PH + 2) & 7, S3 = (PH + 3) & 7;  // rows s-6, s-5
    double r7A, r7B, r6A, r6B;
    {
      const double nS = __shfl_up_sync(0xffffffffu, W.pB[S1], 1), nN = __shfl_down_sync(0xffffffffu, o7A, 1);
      stream_residual<ROWB, COLB>(g, s - 7, cf, o8A, o8B, o7A, W.pB[S1], W.pA[S2], W.pB[S2], o7fA, o7fB, W.fA[S2], W.fB[S2],
                                  W.g0[S1], W.g1[S1], W.g2[S1], W.bA[S1], W.bB[S1], nS, nN, &r7A, &r7B);
    }
    {
      const double nS = __shfl_up_sync(0xffffffffu, W.pB[S2], 1), nN = __shfl_down_sync(0xffffffffu, W.pA[S2], 1);
      stream_residual<ROWB, COLB>(g, s - 6, cf, o7A, W.pB[S1], W.pA[S2], W.pB[S2], W.pA[S3], W.pB[S3], W.fA[S2], W.fB[S2],
                                  W.fA[S3], W.fB[S3], W.g0[S2], W.g1[S2], W.g2[S2], W.bA[S2], W.bB[S2], nS, nN, &r6A, &r6B);
    }
    const bool colA = c >= HL && c < SW - HL && gjA < g.ny, colB = colA && gjA + 1 < g.ny;
    if (EXTRA == 1) {
      if (s - 7 >= jb.ia && s - 7 < jb.ib) {
        if (colA) { acc[0] += r7A * r7A; acc[1] += W.bA[S1] * W.bA[S1]; }
        if (colB) { acc[0] += r7B * r7B; acc[1] += W.bB[S1] * W.bB[S1]; }
      }
      if (s - 6 >= jb.ia && s - 6 < jb.ib) {
        if (colA) { acc[0] += r6A * r6A; acc[1] += W.bA[S2] * W.bA[S2]; }
        if (colB) { acc[0] += r6B * r6B; acc[1] += W.bB[S2] * W.bB[S2]; }
      }
    } else {
      // coarse row I = (s-8)/2: fine rows s-8 (carried), s-7, s-6; coarse column J = gjA/2: fine columns gjA .. gjA+2
      const double n7 = __shfl_down_sync(0xffffffffu, r7A, 1), n6 = __shfl_down_sync(0xffffffffu, r6A, 1);
      const double n8 = __shfl_down_sync(0xffffffffu, acc[0], 1);
      const int I = (s - 8) >> 1, J = gjA >> 1;
      if (s - 7 >= jb.ia && s - 7 < jb.ib && I >= ex.gc.gb && I < ex.gc.ge && colA && J < ex.gc.ny) {
        const double cc = r7B, n = n7, sd = r7A, e = r6B, w = acc[1], ne = n6, nw = n8, se = r6A, sw = acc[0];
        ex.coarse_b[nf_idx(ex.gc, I, J)] = (cc / 4.0 + (((n + sd) + e) + w) / 8.0) + (((ne + nw) + se) + sw) / 16.0;
        if (ex.coarse_x0) ex.coarse_x0[nf_idx(ex.gc, I, J)] = 0.0;
      }
      acc[0] = r6A;
      acc[1] = r6B;
    }
  }
  W.pB[S1] = pB_s1;
  // ---- rows s-NP and s+1-NP are final ----
  if (c >= HL && c < SW - HL && gjA < g.ny) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int rf = s + q - NP;
      const int sf = (PH + q + 16 - NP) & 7;
      if (rf >= jb.ia && rf < jb.ib) {
        const size_t kk = nf_idx(g, rf, gjA);
        if (!COLB || gjA + 1 < g.ny) *reinterpret_cast<double2*>(pout + kk) = make_double2(W.pA[sf], W.pB[sf]);
        else pout[kk] = W.pA[sf];
      }
    }
  }
}

// The march of one job.  COLB (first / last strip) is a property of the job; the row variant is chosen per group of 8 steps.
template <int NP, bool COLB, int EXTRA, bool PRL, class Issue>
__device__ __forceinline__ void stream_job(const nf_grid& g, const Job& jb, int lane, double omega, const unsigned char* ring,
                                           unsigned bar0, Issue& issue, double* __restrict__ pout, double (&acc)[2],
                                           const StreamExtra& ex) {
  Win W;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    W.pA[q] = W.pB[q] = 0.0;
    W.fA[q] = W.fB[q] = 0.0;
    W.g0[q] = W.g1[q] = W.g2[q] = 0.0;
    W.bA[q] = W.bB[q] = 0.0;
    W.iA[q] = W.iB[q] = 0.0;
  }
  ColFlags cf = {false, false, false, false, false, false};
  if (COLB) {
    const int gjA = jb.j0 + 2 * lane;
    cf.A_first = gjA == 0;
    cf.A_last = gjA == g.ny - 1;
    cf.B_last = gjA + 1 == g.ny - 1;
    cf.z0 = gjA == 0 || gjA == g.ny;
    cf.z1 = gjA + 1 == 0 || gjA + 1 == g.ny;
    cf.z2 = gjA + 2 == 0 || gjA + 2 == g.ny;
  }

#define NF_STREAM_STEP2(PH, ROWB)                                                                   \
  {                                                                                                 \
    const int s = s8 + PH;                                                                          \
    if (s > jb.s_last) break;                                                                       \
    if ((PH % RB) == 0) {                                                                           \
      const unsigned bar = bar0 + 8 * (PH / RB);                                                    \
      while (!mbar_try(bar, ring_phase)) {}                                                         \
    }                                                                                               \
    const unsigned char* stage = ring + (PH / RB) * SG_BYTES;                                       \
    stream_step2<PH, NP, ROWB, COLB, EXTRA, PRL>(W, g, jb, s, lane, omega, stage, cf, pout, acc, ex); \
    if (((PH + 1) % RB) == RB - 1) { /* the stage is consumed: refill it for the steps 8 ahead */    \
      __syncwarp();                                                                                 \
      if (lane == 0 && s + 8 - (RB - 2) <= jb.s_last) issue(s + 8 - (RB - 2));                      \
    }                                                                                               \
  }

  unsigned ring_phase = 0;
  for (int s8 = jb.r0;; s8 += 8, ring_phase ^= 1) {
    // rows these 8 steps load or update: s8-NP .. s8+7 (PRL: the last two rows of an even-sized grid follow edge rules)
    if (s8 - NP <= 0 || s8 + 7 >= g.nx - 1 - (PRL ? 2 : 0)) {
      NF_STREAM_STEP2(0, true) NF_STREAM_STEP2(2, true) NF_STREAM_STEP2(4, true) NF_STREAM_STEP2(6, true)
    } else {
      NF_STREAM_STEP2(0, false) NF_STREAM_STEP2(2, false) NF_STREAM_STEP2(4, false) NF_STREAM_STEP2(6, false)
    }
  }
#undef NF_STREAM_STEP2
}

struct StreamMaps {
  CUtensorMap p, b, du, dv, inv;
  CUtensorMap cx;  // PRL: coarse iterate, box CXR x CXC
};

// registers per thread the launch bounds leave (64 K registers per SM, allocated per warp in units of 8 per thread)
constexpr int stream_maxreg(int wpc) { return ((65536 / (32 * wpc)) / 8) * 8 > 255 ? 255 : ((65536 / (32 * wpc)) / 8) * 8; }

// How the (strip, row chunk) jobs are numbered.  The strips that touch column 0 / ny-1 run the COLB variant, which costs a
// few selects more per update; they get shorter chunks so that all warps of the single wave finish together (the kernel's
// duration is the duration of its slowest warp).
struct JobPlan {
  int n_inner, first_right;   // inner strips are 1 .. first_right-1; edge strips: 0 and first_right .. nstrips-1
  int n_edge;
  int len_inner, len_edge;    // rows per chunk (even)
  int jobs_inner, njobs;      // jobs_inner = n_inner * chunks of an inner strip
};

template <int NS, int WPC, int EXTRA, bool PRL>
__global__ void __launch_bounds__(32 * WPC, 1) __maxnreg__(stream_maxreg(WPC))
k_rbsor_stream(nf_grid g, const __grid_constant__ StreamMaps maps, double* __restrict__ pout, double omega, JobPlan plan,
               StreamExtra ex) {
  constexpr int NP = 2 * NS;
  constexpr int HL = EXTRA ? NP + 2 : NP;
  constexpr int SCOLS = SW - 2 * HL;  // columns a strip finalises
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int job = blockIdx.x * WPC + warp;
  if (job >= plan.njobs) { nf_pdl_entry(); return; }
  unsigned char* ring = smem + (size_t)warp * (NSTG * SG_BYTES);
  const unsigned bar0 = s_u32(smem + (size_t)WPC * NSTG * SG_BYTES + warp * NSTG * 8);
  const unsigned ring_u = s_u32(ring);

  Job jb;
  bool colb;
  {
    int chunk, strip, len;
    if (job < plan.jobs_inner) {
      chunk = job / plan.n_inner;
      strip = 1 + (job - chunk * plan.n_inner);
      len = plan.len_inner;
      colb = false;
    } else {
      const int q = job - plan.jobs_inner;
      chunk = q / plan.n_edge;
      const int e = q - chunk * plan.n_edge;
      strip = e == 0 ? 0 : plan.first_right + e - 1;
      len = plan.len_edge;
      colb = true;
    }
    jb.ia = g.gb + chunk * len;
    jb.ib = jb.ia + len < g.ge ? jb.ia + len : g.ge;
    jb.r0 = (jb.ia - HL) & ~1;
    jb.s_last = jb.ib - 1 + NP + EXTRA;  // EXTRA 1 needs the residual of row ib-1 (p final up to ib), EXTRA 2 that of row ib
    if (((jb.s_last - jb.r0) & 1) == 0) jb.s_last += 1;  // steps come in pairs
    jb.j0 = strip * SCOLS - HL;
  }

  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NSTG; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0 + 8 * q) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  nf_pdl_entry();  // everything above is launch-independent set-up; the first TMA load follows

  auto issue = [&](int s) {  // stage of steps s .. s+RB-1: rows s.. of p, d_u; rows s-1.. of d_v, b, 1/aP
    const int q = ((s - jb.r0) / RB) & (NSTG - 1);
    const unsigned bar = bar0 + 8 * q, dst = ring_u + q * SG_BYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(SG_TX + (PRL ? SG_TX_CX : 0u))
                 : "memory");
    if (PRL)  // coarse rows (s-2)/2 .. (s-2)/2 + RB/2 and columns j0/2 - 1 .. j0/2 + 32 (s, j0 even), from an even column
      tma_row(dst + SG_CX, &maps.cx, ((jb.j0 >> 1) - 1) & ~1, ((s - 2) >> 1) - ex.pgc.row0, bar);
    tma_row(dst + SG_P, &maps.p, jb.j0, s - g.row0, bar);
    tma_row(dst + SG_DU, &maps.du, jb.j0, s - g.row0, bar);
    tma_row(dst + SG_B, &maps.b, jb.j0, s - 1 - g.row0, bar);
    tma_row(dst + SG_INV, &maps.inv, jb.j0, s - 1 - g.row0, bar);
    tma_row(dst + SG_DV, &maps.dv, jb.j0, s - 1 - g.row0, bar);
  };
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NSTG; ++q)
      if (jb.r0 + q * RB <= jb.s_last) issue(jb.r0 + q * RB);
  }

  double acc[2] = {0.0, 0.0};
  if (colb) stream_job<NP, true, EXTRA, PRL>(g, jb, lane, omega, ring, bar0, issue, pout, acc, ex);
  else stream_job<NP, false, EXTRA, PRL>(g, jb, lane, omega, ring, bar0, issue, pout, acc, ex);
  if (EXTRA == 1) {
    // deterministic reduction: fixed tree inside the warp, one partial per job, the last job to arrive adds the partials in
    // job order
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], off);
    unsigned int t = 0;
    if (lane == 0) {
      ex.partials[job] = acc[0];
      ex.partials[NF_MAX_PARTIALS + job] = acc[1];
      __threadfence();
      t = atomicAdd(ex.ticket, 1u);
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t == (unsigned int)(plan.njobs - 1)) {
      __threadfence();
      double v0 = 0.0, v1 = 0.0;
      for (int q = lane; q < plan.njobs; q += 32) {
        v0 += ((volatile double*)ex.partials)[q];
        v1 += ((volatile double*)ex.partials)[NF_MAX_PARTIALS + q];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        v0 += __shfl_down_sync(0xffffffffu, v0, off);
        v1 += __shfl_down_sync(0xffffffffu, v1, off);
      }
      if (lane == 0) { ex.out[0] = v0; ex.out[1] = v1; *ex.ticket = 0u; }
    }
  }
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
  int rows, cols, ld, box_cols, box_rows;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols &&
           box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.base);
    h = h * 1000003u ^ (size_t)k.rows;
    h = h * 1000003u ^ (size_t)k.cols;
    h = h * 1000003u ^ (size_t)k.ld;
    h = h * 1000003u ^ (size_t)k.box_cols;
    h = h * 1000003u ^ (size_t)k.box_rows;
    return h;
  }
};

bool row_map(CUtensorMap* out, const double* base, int rows, int cols, int ld, int box_cols, int box_rows = RB) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  std::lock_guard<std::mutex> lock(mu);
  const MapKey key{base, rows, cols, ld, box_cols, box_rows};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = stream_encoder();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
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

// strips / chunks of a level: one job per warp slot of a single wave (148 SMs x WPC warps)
template <int NS, int EXTRA, bool PRL>
JobPlan make_plan(const nf_grid* g, int slots) {
  constexpr int NP = 2 * NS, HL = EXTRA ? NP + 2 : NP, SCOLS = SW - 2 * HL;
  const int rows = g->ge - g->gb;
  const int nstrips = (g->ny + SCOLS - 1) / SCOLS;
  // strip k is an edge strip when its 64 columns (plus the d_v column to their right) reach column 0 or ny-1
  // (PRL: the last two columns of an even-sized grid follow edge rules of the prolongation as well)
  int first_right = (g->ny - 1 - (PRL ? 2 : 0) - SW + HL + SCOLS - 1) / SCOLS;  // smallest k with k*SCOLS - HL + SW >= ny-1
  if (first_right < 1) first_right = 1;
  if (first_right > nstrips) first_right = nstrips;
  JobPlan P;
  P.first_right = first_right;
  P.n_inner = first_right - 1;
  P.n_edge = nstrips - P.n_inner;
  const double edge_cost = PRL ? 1.15 : 1.12;  // relative cost of a COLB step
  // the smallest chunk length (of the inner strips) whose jobs fit into one wave
  int best_li = -1, best_le = -1;
  for (int li = 8; li <= ((rows + 1) & ~1) + 2; li += 2) {
    int le = (int)((li + 2 * HL) / edge_cost) - 2 * HL;
    le &= ~1;
    if (le < 8) le = 8;
    const int ci = (rows + li - 1) / li, ce = (rows + le - 1) / le;
    if ((long long)P.n_inner * ci + (long long)P.n_edge * ce <= slots) { best_li = li; best_le = le; break; }
  }
  if (best_li < 0) { best_li = (rows + 1) & ~1; best_le = best_li; }  // more strips than slots: several waves
  P.len_inner = best_li;
  P.len_edge = best_le;
  P.jobs_inner = P.n_inner * ((rows + best_li - 1) / best_li);
  P.njobs = P.jobs_inner + P.n_edge * ((rows + best_le - 1) / best_le);
  return P;
}

template <int NS, int WPC, int EXTRA, bool PRL>
int launch_stream(nf_ctx* ctx, const nf_grid* g, const double* pin, double* pout, const double* b, const double* d_u,
                  const double* d_v, const double* inv, double omega, const StreamExtra& ex, bool* used) {
  constexpr int SMEM = WPC * NSTG * SG_BYTES + WPC * NSTG * 8;
  *used = false;
  if ((g->gb & 1) != 0) return NF_OK;  // odd origin (the phases assume even chunk starts): the caller falls back
  const JobPlan plan = make_plan<NS, EXTRA, PRL>(g, NF_SM_COUNT * WPC);
  if (plan.njobs > NF_MAX_PARTIALS) return NF_OK;
  const int row_end = g->row1 > 0 ? g->row1 : g->nx + 1;
  const int stored_p = (row_end < g->nx ? row_end : g->nx) - g->row0;
  const int stored_u = row_end - g->row0;
  StreamMaps m;
  if (!row_map(&m.p, pin, stored_p, g->ny, g->ld, SW) || !row_map(&m.b, b, stored_p, g->ny, g->ld, SW) ||
      !row_map(&m.du, d_u, stored_u, g->ny, g->ld, SW) || !row_map(&m.dv, d_v, stored_p, g->ny + 1, g->ld, SW + 2) ||
      !row_map(&m.inv, inv, stored_p, g->ny, g->ld, SW))
    return NF_OK;
  m.cx = m.p;
  if (PRL) {
    const nf_grid& c = ex.pgc;
    const int c_end = c.row1 > 0 ? c.row1 : c.nx + 1;
    const int c_rows = (c_end < c.nx ? c_end : c.nx) - c.row0;
    if (!row_map(&m.cx, ex.pcx, c_rows, c.ny, c.ld, CXC, CXR)) return NF_OK;
  }
  static bool attr_set = false;
  if (!attr_set) {
    NF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_rbsor_stream<NS, WPC, EXTRA, PRL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            SMEM));
    attr_set = true;
  }
  const int grid = (plan.njobs + WPC - 1) / WPC;
  nf_launch(k_rbsor_stream<NS, WPC, EXTRA, PRL>, grid, 32 * WPC, SMEM, ctx->stream, true, *g, m, pout, omega, plan, ex);
  NF_LAUNCH_CHECK(ctx);
  *used = true;
  return NF_OK;
}

}  // namespace

// cached 2-D fp64 tensor map (box_rows x box_cols over rows x cols valid elements, pitch ld): shared with the tiled kernels
bool nfi_tensor_map_2d(CUtensorMap* out, const double* base, int rows, int cols, int ld, int box_cols, int box_rows) {
  return row_map(out, base, rows, cols, ld, box_cols, box_rows);
}

// mode 0: plain; 1: + residual norms -> extra->out[0..1]; 2: + coarse_b = FW(b - A p) on extra->gc.  extra->prolong_c: the
// launch takes p + P(prolong_c) as its input (modes 0 and 1).  The fused work exists for 3-sweep launches; *used = false
// means nothing was launched (the caller takes another path).
int nfi_rbsor_stream(nf_ctx* ctx, const nf_grid* g, const double* pin, double* pout, const double* b, const double* d_u,
                     const double* d_v, const double* inv, double omega, int ns, int mode, const nf_smooth_extra* extra,
                     bool* used) {
  *used = false;
  if (!inv) return NF_OK;
  StreamExtra ex;
  ex.gc = *g; ex.coarse_b = nullptr; ex.coarse_x0 = nullptr; ex.partials = ctx->partials; ex.ticket = ctx->ticket;
  ex.out = nullptr;
  ex.pgc = *g; ex.pcx = nullptr;
  ex.prl_nI = ex.prl_nJ = 0;
  const bool prl = extra && extra->prolong_c != nullptr;
  if (prl) {
    if (ns != 3 || mode == 2) return NF_OK;
    ex.pgc = extra->prolong_gc;
    ex.pcx = extra->prolong_c;
    nfi_prolong_block_extent(&ex.pgc, g, &ex.prl_nI, &ex.prl_nJ);
  }
  if (mode != 0) {
    if (ns != 3 || !extra) return NF_OK;
    ex.gc = extra->gc; ex.coarse_b = extra->coarse_b; ex.coarse_x0 = extra->coarse_x_zero; ex.out = extra->out;
    if (mode == 1 && prl) return launch_stream<3, 8, 1, true>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
    if (mode == 1) return launch_stream<3, 8, 1, false>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
    if (mode == 2) return launch_stream<3, 8, 2, false>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
    return NF_OK;
  }
  if (prl) return launch_stream<3, 8, 0, true>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
  if (ns == 3) return launch_stream<3, 8, 0, false>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
  if (ns == 2) return launch_stream<2, 8, 0, false>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
  if (ns == 1) return launch_stream<1, 8, 0, false>(ctx, g, pin, pout, b, d_u, d_v, inv, omega, ex, used);
  return NF_OK;
}
