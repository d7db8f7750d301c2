// nf_p2p.cu -- halo exchange, scalar all-reduce and row sharing of the slab decomposition done by the ranks' own
// kernels over NVLink peer memory (one process per GPU; no counterpart in the reference, which is a single process).
//
// Why: a V-cycle on slabs performs ~18 halo exchanges of a few rows each.  As grouped ncclSend/ncclRecv every one
// costs ~15-20 us of launch + protocol latency, which is what limits the 8-GPU scaling.  Here an
// exchange is ONE kernel per rank:
//   1. push: the rank stores its boundary rows straight into a staging slot in the neighbour's memory (NVLink stores,
//      fire and forget) as 16-byte packets {data.lo, seq, data.hi, seq}: the sequence number travels WITH the data
//      (the "LL" idea of NCCL; 8-byte halves are written atomically), so no system-scope fence and no separate flag
//      round trip is needed -- measured on 2 B200: a fence + flag version cost 13-17 us per exchange, NCCL 11-17 us;
//   2. wait + unpack: every thread polls the packets the neighbour stores into OUR staging slot until they carry the
//      current sequence number and writes the payload into the halo rows of the field.
// Staging slots alternate (2 per direction), which makes the write-after-read hazard impossible without a second
// handshake: a rank writes slot s again two exchanges later, after it has received the neighbour's packets of the
// exchange in between, which the neighbour sends only after its unpack of slot s has completed (stream order).
// Sequence numbers live in DEVICE memory (ctrl->*_count) and are advanced by the kernels themselves, so the kernels
// carry no per-call host state and can be captured into the CUDA graph of a multigrid cycle and replayed.
//
// Memory: every rank allocates its fields from an arena of cudaMalloc'ed chunks whose cudaIpc handles are exchanged
// once per chunk (ncclAllGather on the team's communicator); all ranks allocate the same sizes in the same order
// (callers pass the maximum over ranks), so a local pointer translates to the peer's address by chunk + offset.
// If any step of the set-up fails on any rank the team falls back to the NCCL path on every rank.
#include "nf_slab.cuh"

#include <stdlib.h>
#include <string.h>

#define NF_P2P_MAX_WORLD 8
#define NF_P2P_RED_MAX 8

struct nf_p2p_ctrl {  // one per rank, lives in chunk 0 (peer-visible)
  unsigned long long halo_flag[2];                    // [0] raised by the lower neighbour, [1] by the upper one
  unsigned long long red_flag[NF_P2P_MAX_WORLD];      // raised by rank q
  unsigned long long ack_flag[NF_P2P_MAX_WORLD];
  unsigned long long done_flag[NF_P2P_MAX_WORLD];
  unsigned long long halo_count, red_count, share_count;  // exchanges completed (local, advanced by the kernels)
  unsigned int ticket[4];
  int error;                                           // set when a wait timed out (peer ran a different program)
  int pad;
  uint4 red_stage[2][NF_P2P_MAX_WORLD][NF_P2P_RED_MAX];  // LL packets {lo, seq, hi, seq}
};

struct P2PChunk {
  char* base = nullptr;
  size_t size = 0, used = 0;
  char* peer[NF_P2P_MAX_WORLD] = {nullptr};
};

struct nf_p2p {
  bool active = false;
  int rank = 0, world = 1;
  std::vector<P2PChunk> chunks;
  nf_p2p_ctrl* ctrl = nullptr;  // local control block
  uint4* stage = nullptr;       // local staging: [from lower | from upper][slot 0 | 1][stage_elems] packets
  size_t stage_elems = 0;
  void* dbuf = nullptr;         // handle exchange scratch
};

// ---- device helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_flag_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_flag_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long nf_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spins until *flag >= seq; gives up after 30 s (a peer that runs a different launch sequence must not hang the box;
// the drivers turn ctrl->error into NF_ERR_CUDA at their next synchronisation)
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long seq, int* error) {
  if (ld_flag_sys(flag) >= seq) return true;
  const unsigned long long t0 = nf_globaltimer();
  for (;;) {
    for (int k = 0; k < 64; ++k)
      if (ld_flag_sys(flag) >= seq) return true;
    if (nf_globaltimer() - t0 > 30000000000ull) { *error = 1; return false; }
  }
}

// ---- LL packets: 8 bytes of payload + the 32-bit sequence number twice, one 16-byte store ---------------------------
__device__ __forceinline__ void ll_store(uint4* dst, double v, unsigned int seq) {
  const unsigned int lo = (unsigned int)__double2loint(v), hi = (unsigned int)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(seq), "r"(hi), "r"(seq) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const uint4* src, unsigned int seq, double* v) {
  unsigned int a, b, c, d;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(src) : "memory");
  if (b != seq || d != seq) return false;
  *v = __hiloint2double((int)c, (int)a);
  return true;
}
// polls until the packet carries seq; gives up after 30 s (see wait_flag)
__device__ __forceinline__ double ll_load(const uint4* src, unsigned int seq, int* error) {
  double v = 0.0;
  if (ll_try_load(src, seq, &v)) return v;
  const unsigned long long t0 = nf_globaltimer();
  for (;;) {
    for (int k = 0; k < 64; ++k)
      if (ll_try_load(src, seq, &v)) return v;
    if (nf_globaltimer() - t0 > 30000000000ull) { *error = 1; return 0.0; }
  }
}

// A thread's share of a segment, U packets in flight at a time: a halo of ~1 MB is ~10 packets per thread, and polling them
// one after the other made the large exchanges latency bound (measured on 4 GPUs: 29 us for the 0.8 MB merged exchange of the
// finest level against 14 us for the small ones).
template <int U>
__device__ __forceinline__ void ll_send(uint4* remote, const double* src, size_t count, size_t g0, size_t gs, unsigned int seq) {
  for (size_t k = g0; k < count; k += (size_t)U * gs) {
    double v[U];
#pragma unroll
    for (int q = 0; q < U; ++q) v[q] = (k + q * gs < count) ? src[k + q * gs] : 0.0;
#pragma unroll
    for (int q = 0; q < U; ++q)
      if (k + q * gs < count) ll_store(remote + k + q * gs, v[q], seq);
  }
}
template <int U>
__device__ __forceinline__ void ll_recv(const uint4* local, double* dst, size_t count, size_t g0, size_t gs, unsigned int seq,
                                        int* error) {
  for (size_t k = g0; k < count; k += (size_t)U * gs) {
    double v[U];
    bool ok[U];
    bool all = true;
#pragma unroll
    for (int q = 0; q < U; ++q) {
      v[q] = 0.0;
      ok[q] = (k + q * gs >= count) || ll_try_load(local + k + q * gs, seq, &v[q]);
      all = all && ok[q];
    }
    if (!all) {
      const unsigned long long t0 = nf_globaltimer();
      for (int spin = 0;; ++spin) {
        all = true;
#pragma unroll
        for (int q = 0; q < U; ++q) {
          if (!ok[q]) ok[q] = ll_try_load(local + k + q * gs, seq, &v[q]);
          all = all && ok[q];
        }
        if (all) break;
        if ((spin & 63) == 63 && nf_globaltimer() - t0 > 30000000000ull) { *error = 1; break; }
      }
    }
#pragma unroll
    for (int q = 0; q < U; ++q)
      if (k + q * gs < count) dst[k + q * gs] = v[q];
  }
}

struct HaloSeg {
  const double* src;  // my owned boundary rows
  double* dst;        // my halo rows
  size_t count;       // doubles; 0 = nothing
};
struct HaloSide {            // one neighbour; up to two fields travel in one exchange (segment 1's packets follow segment 0's)
  HaloSeg seg[2];
  double* zero;              // optional: halo rows of a third field that are cleared (coarse iterate of a cycle)
  size_t zero_count;
  uint4* remote_stage;       // neighbour's staging area for packets coming from me (slot 0)
  const uint4* local_stage;  // my staging area for packets coming from this neighbour (slot 0)
};

struct RedPeers {
  uint4* stage[NF_P2P_MAX_WORLD];  // rank q's red_stage (slot 0, row 0)
};
struct RedArgs {  // optional scalar all-reduce riding on a halo exchange (the LAST block of the launch does it)
  double* buf;
  int count, rank, world;
  RedPeers peers;
};

// sums `count` (<= 8) doubles over the ranks in rank order: identical bits on every rank.  One block.
__device__ __forceinline__ void p2p_allreduce_block(double* buf, int count, nf_p2p_ctrl* ctrl, const RedPeers& peers, int rank,
                                                    int world) {
  __shared__ unsigned long long s_rc;
  const int tid = threadIdx.x;
  if (tid == 0) s_rc = *(volatile unsigned long long*)&ctrl->red_count;
  __syncthreads();
  const unsigned long long c = s_rc;
  const unsigned int seq = (unsigned int)(c + 1);
  const int slot = (int)(c & 1);
  if (tid < count) {
    const double v = buf[tid];
    for (int q = 0; q < world; ++q)
      if (q != rank) ll_store(peers.stage[q] + ((size_t)slot * NF_P2P_MAX_WORLD + rank) * NF_P2P_RED_MAX + tid, v, seq);
    double s = 0.0;
    for (int q = 0; q < world; ++q) s += (q == rank) ? v : ll_load(&ctrl->red_stage[slot][q][tid], seq, &ctrl->error);
    buf[tid] = s;
  }
  __syncthreads();
  if (tid == 0) *(volatile unsigned long long*)&ctrl->red_count = c + 1;
}

__global__ void __launch_bounds__(256) k_p2p_halo(HaloSide lo, HaloSide hi, nf_p2p_ctrl* ctrl, size_t stage_elems, RedArgs red) {
  nf_pdl_entry();
  __shared__ unsigned long long s_c;
  const int tid = threadIdx.x;
  unsigned int nb = gridDim.x;  // blocks that move halo rows
  if (red.count > 0) {
    --nb;
    if (blockIdx.x == nb) {
      p2p_allreduce_block(red.buf, red.count, ctrl, red.peers, red.rank, red.world);
      return;
    }
  }
  if (tid == 0) s_c = *(volatile unsigned long long*)&ctrl->halo_count;
  __syncthreads();
  const unsigned long long c = s_c;
  const unsigned int seq = (unsigned int)(c + 1);
  const size_t slot = (size_t)(c & 1) * stage_elems;
  const size_t g0 = (size_t)blockIdx.x * blockDim.x + tid, gs = (size_t)nb * blockDim.x;
  // 1. push my boundary rows into the neighbours' staging slots
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const size_t off_lo = q ? lo.seg[0].count : 0, off_hi = q ? hi.seg[0].count : 0;
    ll_send<4>(lo.remote_stage + slot + off_lo, lo.seg[q].src, lo.seg[q].count, g0, gs, seq);
    ll_send<4>(hi.remote_stage + slot + off_hi, hi.seg[q].src, hi.seg[q].count, g0, gs, seq);
  }
  for (size_t k = g0; k < lo.zero_count; k += gs) lo.zero[k] = 0.0;
  for (size_t k = g0; k < hi.zero_count; k += gs) hi.zero[k] = 0.0;
  // 2. receive the neighbours' rows
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const size_t off_lo = q ? lo.seg[0].count : 0, off_hi = q ? hi.seg[0].count : 0;
    ll_recv<4>(lo.local_stage + slot + off_lo, lo.seg[q].dst, lo.seg[q].count, g0, gs, seq, &ctrl->error);
    ll_recv<4>(hi.local_stage + slot + off_hi, hi.seg[q].dst, hi.seg[q].count, g0, gs, seq, &ctrl->error);
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&ctrl->ticket[1], 1u);
    if (t == nb - 1) {  // everybody has read `c` and finished: advance the sequence, re-arm the ticket
      ctrl->ticket[1] = 0u;
      *(volatile unsigned long long*)&ctrl->halo_count = c + 1;
    }
  }
}

__global__ void k_p2p_allreduce(double* buf, int count, nf_p2p_ctrl* ctrl, RedPeers peers, int rank, int world) {
  nf_pdl_entry();
  p2p_allreduce_block(buf, count, ctrl, peers, rank, world);
}

struct SharePeers {
  double* array[NF_P2P_MAX_WORLD];               // rank q's copy of the replicated array
  unsigned long long* ack[NF_P2P_MAX_WORLD];     // rank q's ack_flag array
  unsigned long long* done[NF_P2P_MAX_WORLD];    // rank q's done_flag array
};

// replicated array: this rank computed elements [off, off+count); store them into every peer's copy.
// Two handshakes: "entered" (the peers' earlier kernels, which may still read the array, are done) and "stored".
__global__ void __launch_bounds__(256) k_p2p_share(const double* mine, size_t off, size_t count, nf_p2p_ctrl* ctrl,
                                                   SharePeers peers, int rank, int world) {
  __shared__ unsigned long long s_c;
  const int tid = threadIdx.x;
  if (tid == 0) s_c = *(volatile unsigned long long*)&ctrl->share_count;
  __syncthreads();
  const unsigned long long c = s_c, seq = c + 1;
  if (blockIdx.x == 0 && tid < world && tid != rank) st_flag_sys(peers.ack[tid] + rank, seq);
  if (tid < world && tid != rank) wait_flag(&ctrl->ack_flag[tid], seq, &ctrl->error);
  __syncthreads();
  const size_t g0 = (size_t)blockIdx.x * blockDim.x + tid, gs = (size_t)gridDim.x * blockDim.x;
  const double2* s = reinterpret_cast<const double2*>(mine + off);
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    double2* d = reinterpret_cast<double2*>(peers.array[q] + off);
    for (size_t k = g0; k < count / 2; k += gs) d[k] = s[k];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last;
  if (tid == 0) {
    const unsigned int t = atomicAdd(&ctrl->ticket[2], 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) __threadfence_system();
  }
  __syncthreads();
  if (s_last && tid < world && tid != rank) st_flag_sys(peers.done[tid] + rank, seq);
  if (tid < world && tid != rank) wait_flag(&ctrl->done_flag[tid], seq, &ctrl->error);
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&ctrl->ticket[3], 1u);
    if (t == gridDim.x - 1) {
      ctrl->ticket[2] = 0u;
      ctrl->ticket[3] = 0u;
      *(volatile unsigned long long*)&ctrl->share_count = c + 1;
    }
  }
}

// ---- arena ------------------------------------------------------------------------------------------------------
int nf_nccl_allgather_bytes(nf_team* team, void* dev_buf, size_t bytes_per_rank);  // nf_slab.cu
int nf_nccl_all_ok(nf_team* team, int ok, int* all_ok);

static int p2p_new_chunk(nf_team* team, size_t bytes) {
  nf_ctx* ctx = team->ctx;
  nf_p2p* P = team->p2p;
  P2PChunk ch;
  ch.size = bytes;
  if (!P->active) {  // NCCL transport: a private allocation, nothing collective
    NF_CHECK_CUDA(ctx, cudaMalloc((void**)&ch.base, bytes));
    ch.peer[P->rank] = ch.base;
    P->chunks.push_back(ch);
    return NF_OK;
  }
  // Collective part.  A failure that only this rank sees (allocation, export) must not make it leave while the others
  // wait in the next collective: local errors are folded into `ok`, which all ranks agree on before anybody returns.
  int ok = 1;
  if (cudaMalloc((void**)&ch.base, bytes) != cudaSuccess) { cudaGetLastError(); ch.base = nullptr; ok = 0; }
  ch.peer[P->rank] = ch.base;
  bool alloc_failed = !ch.base;
  {
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (ok && cudaIpcGetMemHandle(&h, ch.base) != cudaSuccess) { cudaGetLastError(); ok = 0; memset(&h, 0, sizeof(h)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::vector<cudaIpcMemHandle_t> all(P->world);
    if (!P->dbuf) NF_CHECK_CUDA(ctx, cudaMalloc(&P->dbuf, 64 * NF_P2P_MAX_WORLD));  // 512 bytes, first chunk only
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync((char*)P->dbuf + 64 * P->rank, &h, 64, cudaMemcpyHostToDevice, ctx->stream));
    NF_TRY(nf_nccl_allgather_bytes(team, P->dbuf, 64));
    NF_CHECK_CUDA(ctx, cudaMemcpyAsync(all.data(), P->dbuf, 64 * (size_t)P->world, cudaMemcpyDeviceToHost, ctx->stream));
    NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int all_ok = 0;
    NF_TRY(nf_nccl_all_ok(team, ok, &all_ok));  // nobody opens handles unless everybody could allocate and export
    if (all_ok) {
      for (int q = 0; q < P->world && ok; ++q) {
        if (q == P->rank) continue;
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        ch.peer[q] = (char*)ptr;
      }
      NF_TRY(nf_nccl_all_ok(team, ok, &all_ok));
    }
    if (!all_ok) {
      for (int q = 0; q < P->world; ++q)
        if (q != P->rank && ch.peer[q]) { cudaIpcCloseMemHandle(ch.peer[q]); ch.peer[q] = nullptr; }
      P->active = false;  // every rank takes this branch together: NCCL exchanges from here on
    }
  }
  if (alloc_failed) {  // agreed on above: the other ranks fell back to NCCL, this one reports its allocation failure
    ctx->err = "peer-memory arena: cudaMalloc failed";
    return NF_ERR_ALLOC;
  }
  P->chunks.push_back(ch);
  return NF_OK;
}

static int p2p_alloc_bytes(nf_team* team, size_t bytes, void** out) {
  nf_p2p* P = team->p2p;
  bytes = (bytes + 255) / 256 * 256;
  if (P->chunks.empty() || P->chunks.back().used + bytes > P->chunks.back().size) {
    size_t want = bytes * 8;
    if (want < ((size_t)64 << 20)) want = (size_t)64 << 20;
    if (want > ((size_t)2 << 30)) want = (size_t)2 << 30;
    if (want < bytes) want = bytes;
    NF_TRY(p2p_new_chunk(team, want));
  }
  P2PChunk& ch = P->chunks.back();
  *out = ch.base + ch.used;
  ch.used += bytes;
  return NF_OK;
}

// peer q's address of a pointer that lies in the local arena (nullptr when it does not)
static char* p2p_translate(const nf_p2p* P, const void* ptr, int q) {
  const char* c = (const char*)ptr;
  for (const P2PChunk& ch : P->chunks)
    if (c >= ch.base && c < ch.base + ch.size) return ch.peer[q] ? ch.peer[q] + (c - ch.base) : nullptr;
  return nullptr;
}

bool nf_p2p_active(const nf_team* team) { return team->p2p && team->p2p->active; }

bool nf_p2p_owns(const nf_team* team, const void* ptr) {
  if (!team->p2p) return false;
  const char* c = (const char*)ptr;
  for (const P2PChunk& ch : team->p2p->chunks)
    if (c >= ch.base && c < ch.base + ch.size) return true;
  return false;
}

// collective over the team (called from nf_team_create_nccl): control block in chunk 0
int nf_p2p_enable(nf_team* team, int rank) {
  nf_ctx* ctx = team->ctx;
  if (team->world > NF_P2P_MAX_WORLD || team->world < 2 || !team->nccl) return NF_OK;
  nf_p2p* P = new nf_p2p();
  P->rank = rank;
  P->world = team->world;
  P->active = true;
  team->p2p = P;
  void* p = nullptr;
  const size_t ctrl_bytes = (sizeof(nf_p2p_ctrl) + 255) / 256 * 256;
  NF_TRY(p2p_alloc_bytes(team, ctrl_bytes, &p));
  P->ctrl = (nf_p2p_ctrl*)p;
  NF_CHECK_CUDA(ctx, cudaMemsetAsync(P->ctrl, 0, ctrl_bytes, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int all_ok = 0;
  NF_TRY(nf_nccl_all_ok(team, 1, &all_ok));  // barrier: every control block is zeroed before anybody raises a flag
  return NF_OK;
}

// staging slots (2 directions x 2 slots) for exchanges of up to halo_elems doubles; grows by abandoning the old area
int nf_p2p_reserve_stage(nf_team* team, size_t halo_elems) {
  nf_p2p* P = team->p2p;
  if (!P || halo_elems <= P->stage_elems) return NF_OK;
  nf_ctx* ctx = team->ctx;
  // the exchanges in flight use the old area: drain, and let every rank arrive before anybody switches
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int all_ok = 0;
  NF_TRY(nf_nccl_all_ok(team, 1, &all_ok));
  void* p = nullptr;
  NF_TRY(p2p_alloc_bytes(team, 4 * halo_elems * sizeof(uint4), &p));
  P->stage = (uint4*)p;
  NF_CHECK_CUDA(ctx, cudaMemsetAsync(p, 0, 4 * halo_elems * sizeof(uint4), ctx->stream));
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  NF_TRY(nf_nccl_all_ok(team, 1, &all_ok));  // nobody sends packets into a slot that is still being cleared
  P->stage_elems = halo_elems;
  return NF_OK;
}

// COLLECTIVE over the team (nf_team_free is therefore collective as well): cudaFree on an exported allocation is
// undefined while a peer still maps it, so every rank first drains its stream and closes its imported mappings, then
// all ranks meet (the NCCL communicator is still alive here), and only then the local chunks are released.
void nf_p2p_destroy(nf_team* team) {
  nf_p2p* P = team->p2p;
  if (!P) return;
  cudaStreamSynchronize(team->ctx->stream);
  for (P2PChunk& ch : P->chunks)
    for (int q = 0; q < P->world; ++q)
      if (q != P->rank && ch.peer[q]) { cudaIpcCloseMemHandle(ch.peer[q]); ch.peer[q] = nullptr; }
  if (team->nccl && P->world > 1) {
    int all_ok = 0;
    nf_nccl_all_ok(team, 1, &all_ok);  // barrier: nobody maps (or spins on) this rank's memory any more
  }
  for (P2PChunk& ch : P->chunks)
    if (ch.base) cudaFree(ch.base);
  if (P->dbuf) cudaFree(P->dbuf);
  delete P;
  team->p2p = nullptr;
}

// ---- allocation front end used by the drivers -----------------------------------------------------------------------
// elems: what this rank needs; elems_max: maximum over the ranks (all ranks must allocate identical arena blocks)
double* nf_team_alloc(nf_team* team, size_t elems, size_t elems_max) {
  nf_ctx* ctx = team->ctx;
  double* ptr = nullptr;
  if (team->p2p) {
    void* p = nullptr;
    if (p2p_alloc_bytes(team, (elems_max > elems ? elems_max : elems) * sizeof(double), &p) != NF_OK) return nullptr;
    ptr = (double*)p;
    elems = elems_max > elems ? elems_max : elems;
  } else if (cudaMalloc(&ptr, elems * sizeof(double)) != cudaSuccess) {
    return nullptr;
  }
  cudaMemsetAsync(ptr, 0, elems * sizeof(double), ctx->stream);
  return ptr;
}

void nf_team_release(nf_team* team, void* ptr) {
  if (!ptr) return;
  if (team && nf_p2p_owns(team, ptr)) return;  // arena memory goes away with the team
  cudaFree(ptr);
}

// ---- the three collectives ------------------------------------------------------------------------------------------
static int p2p_red_peers(nf_p2p* P, RedPeers* peers) {
  memset(peers, 0, sizeof(*peers));
  for (int q = 0; q < P->world; ++q) {
    char* rctrl = p2p_translate(P, P->ctrl, q);
    if (!rctrl) return NF_ERR_UNSUPPORTED;
    nf_p2p_ctrl* rc = (nf_p2p_ctrl*)rctrl;
    peers->stage[q] = &rc->red_stage[0][0][0];
  }
  return NF_OK;
}

// rows exchanged across the boundary between ranks lo and lo+1 (symmetric: both sides compute the same number)
static int p2p_depth(const LevelGeom& geom, int lo, int depth) {
  int d = depth;
  if (d > geom.halo) d = geom.halo;
  if (d > geom.ge[lo] - geom.gb[lo]) d = geom.ge[lo] - geom.gb[lo];
  if (d > geom.ge[lo + 1] - geom.gb[lo + 1]) d = geom.ge[lo + 1] - geom.gb[lo + 1];
  return d;
}

// One kernel: halo rows of up to two fields (possibly of different levels), optionally the halo rows of `zero` (same
// geometry as field 1... of `zgeom`) cleared, optionally `red_count` scalars summed over all ranks.
// returns NF_ERR_UNSUPPORTED when this call cannot take the peer path (caller falls back to separate collectives)
int nf_p2p_exchange_multi(nf_team* team, int nfields, const LevelGeom* const* geoms, double* const* fields, const int* depths,
                          const LevelGeom* zgeom, double* zero, double* red_buf, int red_count) {
  nf_ctx* ctx = team->ctx;
  nf_p2p* P = team->p2p;
  const int r = P->rank;
  if (nfields < 1 || nfields > 2 || red_count > NF_P2P_RED_MAX) return NF_ERR_UNSUPPORTED;
  HaloSide side[2];
  memset(side, 0, sizeof(side));
  size_t total = 0;
  for (int s = 0; s < 2; ++s) {  // s = 0: boundary with r-1, s = 1: boundary with r+1
    const int q = s == 0 ? r - 1 : r + 1;
    if (q < 0 || q >= team->world) continue;
    const int lo = s == 0 ? q : r;  // the boundary lies between ranks lo and lo+1
    HaloSide& H = side[s];
    size_t staged = 0;
    for (int f = 0; f < nfields; ++f) {
      const LevelGeom& geom = *geoms[f];
      const int B = geom.ge[lo];
      const int d = p2p_depth(geom, lo, depths[f]);
      const size_t count = (size_t)d * geom.ld;
      H.seg[f].count = count;
      if (count == 0) continue;
      // rows [B-d, B) are owned by lo, rows [B, B+d) by lo+1
      if (s == 0) {  // I am lo+1: send my first d owned rows, receive rows [B-d, B)
        H.seg[f].src = fields[f] + (size_t)(B - geom.row0(r)) * geom.ld;
        H.seg[f].dst = fields[f] + (size_t)(B - d - geom.row0(r)) * geom.ld;
      } else {       // I am lo: send my last d owned rows, receive rows [B, B+d)
        H.seg[f].src = fields[f] + (size_t)(B - d - geom.row0(r)) * geom.ld;
        H.seg[f].dst = fields[f] + (size_t)(B - geom.row0(r)) * geom.ld;
      }
      staged += count;
    }
    if (zero && zgeom) {
      const int B = zgeom->ge[lo];
      const int d = p2p_depth(*zgeom, lo, zgeom->halo);
      H.zero_count = (size_t)d * zgeom->ld;
      H.zero = zero + (size_t)((s == 0 ? B - d : B) - zgeom->row0(r)) * zgeom->ld;
    }
    if (staged > P->stage_elems) return NF_ERR_UNSUPPORTED;
    if (staged == 0 && H.zero_count == 0) continue;
    // the neighbour stages what comes from me in its "from upper" area when I am above it (s == 0), else "from lower"
    char* rstage = p2p_translate(P, P->stage + (size_t)(s == 0 ? 1 : 0) * 2 * P->stage_elems, q);
    if (!rstage) return NF_ERR_UNSUPPORTED;
    H.remote_stage = (uint4*)rstage;
    H.local_stage = P->stage + (size_t)s * 2 * P->stage_elems;
    total += staged + H.zero_count;
  }
  RedArgs red;
  memset(&red, 0, sizeof(red));
  if (red_count > 0 && red_buf) {
    if (p2p_red_peers(P, &red.peers) != NF_OK) return NF_ERR_UNSUPPORTED;
    red.buf = red_buf; red.count = red_count; red.rank = P->rank; red.world = P->world;
  }
  if (total == 0 && red.count == 0) return NF_OK;
  static const int per_block = getenv("NF_P2P_BLOCK_BYTES") ? atoi(getenv("NF_P2P_BLOCK_BYTES")) : 8192;
  static const int max_blocks = getenv("NF_P2P_MAX_BLOCKS") ? atoi(getenv("NF_P2P_MAX_BLOCKS")) : 2 * NF_SM_COUNT;
  int blocks = (int)((total * sizeof(double) + per_block - 1) / per_block);
  if (blocks < 1) blocks = 1;
  if (blocks > max_blocks) blocks = max_blocks;
  nf_launch(k_p2p_halo, blocks + (red.count > 0 ? 1 : 0), 256, 0, ctx->stream, true, side[0], side[1], P->ctrl, P->stage_elems, red);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nf_p2p_exchange(nf_team* team, const LevelGeom& geom, double* field, int depth) {
  const LevelGeom* g = &geom;
  return nf_p2p_exchange_multi(team, 1, &g, &field, &depth, nullptr, nullptr, nullptr, 0);
}

int nf_p2p_allreduce(nf_team* team, double* buf, size_t count) {
  nf_ctx* ctx = team->ctx;
  nf_p2p* P = team->p2p;
  if (count > NF_P2P_RED_MAX) return NF_ERR_UNSUPPORTED;
  RedPeers peers;
  if (p2p_red_peers(P, &peers) != NF_OK) return NF_ERR_UNSUPPORTED;
  nf_launch(k_p2p_allreduce, 1, 32, 0, ctx->stream, true, buf, (int)count, P->ctrl, peers, P->rank, P->world);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nf_p2p_share_rows(nf_team* team, int ld, int nx, const std::vector<int>& gb, const std::vector<int>& ge, double* array,
                      int utype) {
  nf_ctx* ctx = team->ctx;
  nf_p2p* P = team->p2p;
  SharePeers peers;
  memset(&peers, 0, sizeof(peers));
  for (int q = 0; q < P->world; ++q) {
    char* ra = p2p_translate(P, array, q);
    char* rctrl = p2p_translate(P, P->ctrl, q);
    if (!ra || !rctrl) return NF_ERR_UNSUPPORTED;
    nf_p2p_ctrl* rc = (nf_p2p_ctrl*)rctrl;
    peers.array[q] = (double*)ra;
    peers.ack[q] = rc->ack_flag;
    peers.done[q] = rc->done_flag;
  }
  const int r = P->rank;
  int b = gb[r], e = ge[r];
  if (utype && r == team->world - 1) e = nx + 1;
  const size_t off = (size_t)b * ld, count = e > b ? (size_t)(e - b) * ld : 0;
  if ((off & 1) || (count & 1)) return NF_ERR_UNSUPPORTED;
  int blocks = (int)((count * sizeof(double) * (P->world - 1) + 65535) / 65536);
  if (blocks < 1) blocks = 1;
  if (blocks > 32) blocks = 32;
  k_p2p_share<<<blocks, 256, 0, ctx->stream>>>(array, off, count, P->ctrl, peers, r, P->world);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

// 1 when a wait inside one of the kernels above timed out since the team was created (stream must be idle)
int nf_p2p_error(nf_team* team) {
  if (!team->p2p || !team->p2p->ctrl) return 0;
  int e = 0;
  cudaMemcpy(&e, &team->p2p->ctrl->error, sizeof(int), cudaMemcpyDeviceToHost);
  return e;
}
