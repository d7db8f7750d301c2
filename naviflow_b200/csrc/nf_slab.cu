// nf_slab.cu -- team / halo-exchange plumbing of the row-slab decomposition (see nf_slab.cuh).
// NCCL is reached through dlopen("libnccl.so.2") so that the library binds to the copy PyTorch has already
// loaded into the process (no link-time dependency, no second NCCL instance).
#include "nf_slab.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

// ---- minimal NCCL surface (types and enums as in nccl.h 2.x; the ABI of these entry points is stable) ----------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8_ = 1, ncclInt32_ = 2, ncclFloat64_ = 8, ncclSum_ = 0, ncclMin_ = 3 };
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok = false;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      *(void**)&api.GetUniqueId = dlsym(h, "ncclGetUniqueId");
      *(void**)&api.CommInitRank = dlsym(h, "ncclCommInitRank");
      *(void**)&api.CommDestroy = dlsym(h, "ncclCommDestroy");
      *(void**)&api.GroupStart = dlsym(h, "ncclGroupStart");
      *(void**)&api.GroupEnd = dlsym(h, "ncclGroupEnd");
      *(void**)&api.Send = dlsym(h, "ncclSend");
      *(void**)&api.Recv = dlsym(h, "ncclRecv");
      *(void**)&api.AllReduce = dlsym(h, "ncclAllReduce");
      *(void**)&api.Broadcast = dlsym(h, "ncclBroadcast");
      *(void**)&api.AllGather = dlsym(h, "ncclAllGather");
      *(void**)&api.GetErrorString = dlsym(h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send &&
               api.Recv && api.AllReduce && api.Broadcast;
    }
  }
  return &api;
}

#define NF_NCCL(ctx, expr)                                                                      \
  do {                                                                                          \
    ncclResult_t _r = (expr);                                                                   \
    if (_r != 0) {                                                                              \
      NcclApi* _a = nccl_api();                                                                 \
      (ctx)->err = std::string(#expr) + ": " + (_a->GetErrorString ? _a->GetErrorString(_r) : "NCCL error"); \
      return NF_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

// ---- partitions -------------------------------------------------------------------------------------------------
bool nf_split_rows(int nx, int world, int min_rows, std::vector<int>& gb, std::vector<int>& ge) {
  gb.assign(world, 0);
  ge.assign(world, nx);
  if (world <= 1) return false;
  for (int r = 0; r < world; ++r) {
    long long b = ((long long)nx * r) / world;
    b = (b / 16) * 16;  // boundaries on multiples of 16: even tile origins on the first four levels (fused
                        // residual+restriction) and 2I+1 parents aligned with the owner of the fine rows
    gb[r] = (int)b;
  }
  for (int r = 0; r < world; ++r) ge[r] = (r + 1 < world) ? gb[r + 1] : nx;
  for (int r = 0; r < world; ++r)
    if (ge[r] - gb[r] < min_rows) {
      gb.assign(world, 0);
      ge.assign(world, nx);
      return false;
    }
  return true;
}

void nf_coarsen_split(const std::vector<int>& gbf, const std::vector<int>& gef, int nxc, std::vector<int>& gb,
                      std::vector<int>& ge) {
  const int world = (int)gbf.size();
  gb.assign(world, 0);
  ge.assign(world, 0);
  for (int r = 0; r < world; ++r) {
    // smallest I with 2I+1 >= gbf[r]
    int b = gbf[r] <= 1 ? 0 : gbf[r] / 2;  // ceil((gbf-1)/2)
    if (b > nxc) b = nxc;
    gb[r] = b;
  }
  for (int r = 0; r < world; ++r) ge[r] = (r + 1 < world) ? gb[r + 1] : nxc;
}

// ---- exchanges --------------------------------------------------------------------------------------------------
int nf_team_exchange(nf_team* team, const LevelGeom& geom, double* const* fields, int depth) {
  nf_ctx* ctx = team->ctx;
  if (!geom.dist || team->world <= 1 || depth <= 0) return NF_OK;
  static const bool skip = getenv("NF_SKIP_EXCHANGE") != nullptr;  // timing experiments only: results are wrong
  if (skip) return NF_OK;
  if (depth > geom.halo) depth = geom.halo;
  if (nf_p2p_active(team)) {  // one kernel per rank over NVLink peer memory (nf_p2p.cu)
    const int st = nf_p2p_exchange(team, geom, fields[0], depth);
    if (st != NF_ERR_UNSUPPORTED) return st;
  }
  NcclApi* api = team->nccl ? nccl_api() : nullptr;
  if (api) NF_NCCL(ctx, api->GroupStart());
  for (int r = 0; r + 1 < team->world; ++r) {
    const int B = geom.ge[r];  // == gb[r+1]
    int d = depth;
    if (d > geom.ge[r] - geom.gb[r]) d = geom.ge[r] - geom.gb[r];
    if (d > geom.ge[r + 1] - geom.gb[r + 1]) d = geom.ge[r + 1] - geom.gb[r + 1];
    const size_t count = (size_t)d * geom.ld;
    const int lo = team->slot_of(r), hi = team->slot_of(r + 1);
    // rows [B-d, B) live on r (owned) and in r+1's lower halo; rows [B, B+d) live on r+1 (owned) and in r's upper halo
    if (!api) {
      double* flo = fields[lo];
      double* fhi = fields[hi];
      NF_CHECK_CUDA(ctx, cudaMemcpyAsync(fhi + (size_t)(B - d - geom.row0(r + 1)) * geom.ld,
                                         flo + (size_t)(B - d - geom.row0(r)) * geom.ld, count * sizeof(double),
                                         cudaMemcpyDeviceToDevice, ctx->stream));
      NF_CHECK_CUDA(ctx, cudaMemcpyAsync(flo + (size_t)(B - geom.row0(r)) * geom.ld,
                                         fhi + (size_t)(B - geom.row0(r + 1)) * geom.ld, count * sizeof(double),
                                         cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      ncclComm_t comm = (ncclComm_t)team->nccl;
      if (lo >= 0) {
        double* f = fields[lo];
        NF_NCCL(ctx, api->Send(f + (size_t)(B - d - geom.row0(r)) * geom.ld, count, ncclFloat64_, r + 1, comm, ctx->stream));
        NF_NCCL(ctx, api->Recv(f + (size_t)(B - geom.row0(r)) * geom.ld, count, ncclFloat64_, r + 1, comm, ctx->stream));
      }
      if (hi >= 0) {
        double* f = fields[hi];
        NF_NCCL(ctx, api->Send(f + (size_t)(B - geom.row0(r + 1)) * geom.ld, count, ncclFloat64_, r, comm, ctx->stream));
        NF_NCCL(ctx, api->Recv(f + (size_t)(B - d - geom.row0(r + 1)) * geom.ld, count, ncclFloat64_, r, comm, ctx->stream));
      }
    }
  }
  if (api) NF_NCCL(ctx, api->GroupEnd());
  return NF_OK;
}

__global__ void k_accumulate(double* __restrict__ dst, const double* __restrict__ src, size_t n) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    dst[k] = dst[k] + src[k];
}

int nf_team_allreduce(nf_team* team, double* const* bufs, size_t count) {
  nf_ctx* ctx = team->ctx;
  if (team->world <= 1 || count == 0) return NF_OK;
  if (nf_p2p_active(team)) {
    const int st = nf_p2p_allreduce(team, bufs[0], count);
    if (st != NF_ERR_UNSUPPORTED) return st;
  }
  if (team->nccl) {
    NcclApi* api = nccl_api();
    NF_NCCL(ctx, api->AllReduce(bufs[0], bufs[0], count, ncclFloat64_, ncclSum_, (ncclComm_t)team->nccl, ctx->stream));
    return NF_OK;
  }
  // in-process ranks: rank-ordered sum into slab 0, then copy back to everyone
  const unsigned blocks = (unsigned)((count + 255) / 256 > 1024 ? 1024 : (count + 255) / 256);
  for (size_t k = 1; k < team->local.size(); ++k) {
    k_accumulate<<<blocks, 256, 0, ctx->stream>>>(bufs[0], bufs[k], count);
    NF_LAUNCH_CHECK(ctx);
  }
  for (size_t k = 1; k < team->local.size(); ++k)
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(bufs[k], bufs[0], count * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return NF_OK;
}

int nf_team_share_rows(nf_team* team, int ld, int nx, const std::vector<int>& gb, const std::vector<int>& ge,
                       double* const* arrays, int utype) {
  nf_ctx* ctx = team->ctx;
  if (team->world <= 1) return NF_OK;
  if (nf_p2p_active(team)) {
    const int st = nf_p2p_share_rows(team, ld, nx, gb, ge, arrays[0], utype);
    if (st != NF_ERR_UNSUPPORTED) return st;
  }
  NcclApi* api = team->nccl ? nccl_api() : nullptr;
  if (api) NF_NCCL(ctx, api->GroupStart());
  for (int r = 0; r < team->world; ++r) {
    int b = gb[r], e = ge[r];
    if (utype && r == team->world - 1) e = nx + 1;
    if (e <= b) continue;
    const size_t off = (size_t)b * ld, count = (size_t)(e - b) * ld;
    if (api) {
      double* a = arrays[0];
      NF_NCCL(ctx, api->Broadcast(a + off, a + off, count, ncclFloat64_, r, (ncclComm_t)team->nccl, ctx->stream));
    } else {
      const int src = team->slot_of(r);
      for (size_t k = 0; k < team->local.size(); ++k)
        if ((int)k != src)
          NF_CHECK_CUDA(ctx, cudaMemcpyAsync(arrays[k] + off, arrays[src] + off, count * sizeof(double),
                                             cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  if (api) NF_NCCL(ctx, api->GroupEnd());
  return NF_OK;
}

// ---- small NCCL collectives used while the peer-memory path is set up (nf_p2p.cu) ----------------------------------------
// in-place all-gather: rank r's bytes_per_rank bytes sit at dev_buf + r * bytes_per_rank
int nf_nccl_allgather_bytes(nf_team* team, void* dev_buf, size_t bytes_per_rank) {
  nf_ctx* ctx = team->ctx;
  NcclApi* api = nccl_api();
  NF_REQUIRE(ctx, team->nccl && api->AllGather, "no NCCL communicator / ncclAllGather");
  int rank = team->local[0];
  NF_NCCL(ctx, api->AllGather((char*)dev_buf + (size_t)rank * bytes_per_rank, dev_buf, bytes_per_rank, ncclUint8_,
                              (ncclComm_t)team->nccl, ctx->stream));
  return NF_OK;
}

// *all_ok = 1 iff ok != 0 on every rank (also a barrier of the team's streams and hosts)
int nf_nccl_all_ok(nf_team* team, int ok, int* all_ok) {
  nf_ctx* ctx = team->ctx;
  NcclApi* api = nccl_api();
  NF_REQUIRE(ctx, team->nccl != nullptr, "no NCCL communicator");
  int* d = nullptr;
  NF_CHECK_CUDA(ctx, cudaMalloc(&d, sizeof(int)));
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(d, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NF_NCCL(ctx, api->AllReduce(d, d, 1, ncclInt32_, ncclMin_, (ncclComm_t)team->nccl, ctx->stream));
  int v = 0;
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(&v, d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d);
  *all_ok = v;
  return NF_OK;
}

// ---- team objects -----------------------------------------------------------------------------------------------
int nf_team_create_local(nf_ctx* ctx, int virtual_ranks, nf_team** out) {
  NF_REQUIRE(ctx, out && virtual_ranks >= 1 && virtual_ranks <= 64, "virtual_ranks must be in 1..64");
  nf_team* t = new nf_team();
  t->ctx = ctx;
  t->world = virtual_ranks;
  for (int r = 0; r < virtual_ranks; ++r) t->local.push_back(r);
  *out = t;
  return NF_OK;
}

extern "C" int nf_nccl_unique_id(nf_ctx* ctx, void* id_out_128_bytes) {
  NcclApi* api = nccl_api();
  NF_REQUIRE(ctx, api->ok, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  NF_NCCL(ctx, api->GetUniqueId(&id));
  memcpy(id_out_128_bytes, &id, sizeof(id));
  return NF_OK;
}

extern "C" int nf_team_create_nccl(nf_ctx* ctx, int world, int rank, const void* id_128_bytes, nf_team** out) {
  NF_REQUIRE(ctx, out && world >= 1 && rank >= 0 && rank < world, "bad world / rank");
  nf_team* t = new nf_team();
  t->ctx = ctx;
  t->world = world;
  t->local.push_back(rank);
  if (world > 1) {
    NcclApi* api = nccl_api();
    if (!api->ok) { delete t; ctx->err = "libnccl.so.2 could not be loaded"; return NF_ERR_UNSUPPORTED; }
    ncclUniqueId id;
    memcpy(&id, id_128_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    ncclResult_t r = api->CommInitRank(&comm, world, id, rank);
    if (r != 0) { delete t; ctx->err = "ncclCommInitRank failed"; return NF_ERR_CUDA; }
    t->nccl = comm;
    // exchanges by the ranks' own kernels over NVLink peer memory unless NF_P2P=0 (then, or when cudaIpc is not
    // available, grouped ncclSend/ncclRecv)
    // The ranks must take the same branch (nf_p2p_enable is a sequence of collectives): agree on the switch first, so that
    // an environment that differs between the ranks selects NCCL everywhere instead of deadlocking.
    const char* env = getenv("NF_P2P");
    int want = (!(env && env[0] == '0') && api->AllGather) ? 1 : 0, all_want = 0;
    {
      int st = nf_nccl_all_ok(t, want, &all_want);
      if (st != NF_OK) { nf_team_destroy(t); return st; }
    }
    if (all_want) {
      int st = nf_p2p_enable(t, rank);
      if (st != NF_OK) { nf_team_destroy(t); return st; }
    }
  }
  *out = t;
  return NF_OK;
}

// 1: halos / norms travel through peer memory (nf_p2p.cu), 0: through NCCL or in-process copies
extern "C" int nf_team_uses_p2p(nf_team* t) { return t && nf_p2p_active(t) ? 1 : 0; }

extern "C" int nf_team_create_virtual(nf_ctx* ctx, int virtual_ranks, nf_team** out) {
  return nf_team_create_local(ctx, virtual_ranks, out);
}

int nf_team_destroy(nf_team* t) {
  if (!t) return NF_OK;
  nf_p2p_destroy(t);
  if (t->nccl) {
    NcclApi* api = nccl_api();
    if (api->ok) api->CommDestroy((ncclComm_t)t->nccl);
  }
  delete t;
  return NF_OK;
}

extern "C" int nf_team_free(nf_team* t) { return nf_team_destroy(t); }

// Host-only queries of the partition rules (no device needed; used by the CPU tests and by callers that want to
// know which rows a rank will own before creating any state).
extern "C" int nf_slab_rows(int nx, int world, int rank, int* row_begin, int* row_end) {
  if (nx < 3 || world < 1 || rank < 0 || rank >= world) return NF_ERR_ARG;
  std::vector<int> gb, ge;
  nf_split_rows(nx, world, 64, gb, ge);  // 64 = NF_MIN_SLAB_ROWS: thinner slabs -> the grid is not cut
  if (row_begin) *row_begin = gb[rank];
  if (row_end) *row_end = ge[rank];
  return NF_OK;
}

// rows of the next-coarser level (nxc cells) a rank restricts into, given its fine rows [fb, fe) and those of the
// rank above it: coarse row I belongs to the owner of fine row 2I+1
extern "C" int nf_slab_coarse_rows(int fine_begin, int next_fine_begin, int is_last, int nxc, int* row_begin,
                                   int* row_end) {
  std::vector<int> gbf = {fine_begin, next_fine_begin}, gef = {next_fine_begin, 0}, gb, ge;
  nf_coarsen_split(gbf, gef, nxc, gb, ge);
  if (row_begin) *row_begin = gb[0];
  if (row_end) *row_end = is_last ? nxc : ge[0];
  return NF_OK;
}

// Times `reps` back-to-back halo exchanges of `depth` rows (and, if n_scalars > 0, all-reduces of n_scalars doubles) of
// an nx x ny level cut over the team, with CUDA events on the context's stream; the transport is the team's
// (peer-memory kernels or NCCL).  Collective.  Used by tools/bench_exchange.py.
extern "C" int nf_team_benchmark(nf_team* team, int nx, int ny, int depth, int n_scalars, int reps, double* ms_exchange,
                                 double* ms_allreduce) {
  if (!team) return NF_ERR_ARG;
  nf_ctx* ctx = team->ctx;
  NF_REQUIRE(ctx, team->local.size() == 1 && team->world > 1, "needs a team of separate processes");
  LevelGeom g;
  g.nx = nx; g.ny = ny; g.ld = nf_pad_ld(ny);
  g.dist = nf_split_rows(nx, team->world, 64, g.gb, g.ge);
  g.halo = g.dist ? NF_HALO : 0;
  NF_REQUIRE(ctx, g.dist, "grid too small to be cut");
  NF_TRY(nf_p2p_reserve_stage(team, (size_t)NF_HALO * g.ld));
  double* f = nf_team_alloc(team, g.elems(team->local[0]), g.max_elems());
  double* sc = nf_team_alloc(team, 8, 8);
  NF_REQUIRE(ctx, f && sc, "allocation failed");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.f;
  for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms up
    NF_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    for (int k = 0; k < reps; ++k) NF_TRY(nf_team_exchange(team, g, &f, depth));
    NF_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ms, e0, e1);
  }
  if (ms_exchange) *ms_exchange = ms / reps;
  ms = 0.f;
  if (n_scalars > 0)
    for (int pass = 0; pass < 2; ++pass) {
      NF_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
      for (int k = 0; k < reps; ++k) NF_TRY(nf_team_allreduce(team, &sc, (size_t)n_scalars));
      NF_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
      NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      cudaEventElapsedTime(&ms, e0, e1);
    }
  if (ms_allreduce) *ms_allreduce = ms / reps;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  nf_team_release(team, f);
  nf_team_release(team, sc);
  return NF_OK;
}
