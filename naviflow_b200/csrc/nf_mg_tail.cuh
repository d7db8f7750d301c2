// nf_mg_tail.cuh -- arguments of the single-CTA coarse end of a V-cycle (nf_mg_tail.cu)
#pragma once
#include "nf_common.cuh"

#define NF_TAIL_MAX_LEVELS 8
#define NF_TAIL_MAX_N 31          // levels of at most this many cells per side go into the tail kernel
#define NF_TAIL_MAX_SMEM 232448   // dynamic shared memory one CTA can get on sm_100a (227 KB)

struct nf_tail_level {
  double* x;            // level solution (written for the first tail level only)
  const double* b;      // right-hand side (read for the first tail level only)
  const double *d_u, *d_v, *inv;
  int nx, ny, ld;
  double dx, dy;
};

struct nf_tail_args {
  nf_tail_level lv[NF_TAIL_MAX_LEVELS];
  int nlev;
  int N;                      // unknowns of the coarsest level
  const double* coarse_inv;   // its dense inverse (N x N, row major)
  double rho, omega;
  int pre, post;
};

size_t nfi_mg_tail_smem(const nf_tail_args* a);
int nfi_mg_tail(nf_ctx* ctx, const nf_tail_args* a);
