// nf_slab.cuh -- row-slab decomposition of the (nx, ny) grids over the GPUs of one box.
//
// The reference is a single NumPy process (no distributed code), so this layer is new.
// The C-ordered arrays are cut along i (their slow index) into contiguous row slabs; every slab stores
// NF_HALO extra rows on each side.  A halo exchange is therefore a contiguous copy of `depth * ld` doubles:
//   * between processes (one rank per GPU, torchrun): ncclSend / ncclRecv over NVLink, grouped per exchange;
//   * between "virtual ranks" that live in one process on one device (used by the parity tests, which run the
//     same slab code on a single GPU): cudaMemcpyAsync on the context's stream.
// The drivers (multigrid, SIMPLE) are written once against a Team: a list of the slabs this process owns
// plus the communicator.  A single-GPU run is a Team with one slab, no halo and no communication.
#pragma once
#include <vector>

#include "nf_common.cuh"

#define NF_HALO 8  // rows of halo kept on each side of a distributed slab

struct nf_p2p;  // peer-memory exchange state (nf_p2p.cu)

struct nf_team {
  nf_ctx* ctx = nullptr;
  int world = 1;               // number of ranks the grid is cut into
  std::vector<int> local;      // global ids of the ranks living in this process (1 entry under torchrun)
  void* nccl = nullptr;        // ncclComm_t when world > 1 and the ranks are separate processes
  nf_p2p* p2p = nullptr;       // NVLink peer-memory path of the exchanges (separate processes, world <= 8)
  bool is_local(int r) const { for (int q : local) if (q == r) return true; return false; }
  int slot_of(int r) const { for (size_t k = 0; k < local.size(); ++k) if (local[k] == r) return (int)k; return -1; }
};

// geometry of one grid level for every rank of the team
struct LevelGeom {
  int nx = 0, ny = 0, ld = 0;
  int halo = 0;            // 0 when the level is not cut (single rank or replicated coarse level)
  bool dist = false;       // rows are cut across ranks
  std::vector<int> gb, ge; // owned cell rows [gb, ge) per rank (replicated: [0, nx) for everyone)
  double dx = 0, dy = 0, rho = 1.0;

  int row0(int r) const { return dist ? (gb[r] - halo > 0 ? gb[r] - halo : 0) : 0; }
  int row1(int r) const { return dist ? (ge[r] + halo + 1 < nx + 1 ? ge[r] + halo + 1 : nx + 1) : nx + 1; }
  size_t elems(int r) const { return (size_t)(row1(r) - row0(r)) * ld; }
  size_t max_elems() const {  // largest slab of the level (peer-memory arenas allocate the same size on every rank)
    size_t m = 0;
    for (size_t r = 0; r < gb.size(); ++r) m = elems((int)r) > m ? elems((int)r) : m;
    return m ? m : (size_t)(nx + 1) * ld;
  }
  nf_grid grid(int r) const {
    nf_grid g;
    g.nx = nx; g.ny = ny; g.ld = ld; g.row0 = row0(r); g.gb = gb[r]; g.ge = ge[r]; g.row1 = row1(r); g.pad = 0;
    g.dx = dx; g.dy = dy; g.rho = rho;
    return g;
  }
  // same grid with the computed row range grown by e rows on each side (clipped to the domain / the storage)
  nf_grid grid_ext(int r, int e) const {
    nf_grid g = grid(r);
    if (!dist) return g;
    g.gb = gb[r] - e > 0 ? gb[r] - e : 0;
    g.ge = ge[r] + e < nx ? ge[r] + e : nx;
    return g;
  }
};

static inline int nf_pad_ld(int ny) { return ((ny + 1 + 15) / 16) * 16; }

// even split of nx rows over `world` ranks with even boundaries; returns false when a slab would be thinner than
// min_rows (then the level is not cut)
bool nf_split_rows(int nx, int world, int min_rows, std::vector<int>& gb, std::vector<int>& ge);
// ownership of the next-coarser level induced by a fine partition: coarse row I belongs to the owner of fine
// row 2I+1 (full weighting and injection both centre on it)
void nf_coarsen_split(const std::vector<int>& gbf, const std::vector<int>& gef, int nxc, std::vector<int>& gb,
                      std::vector<int>& ge);

// fields[k] is the array of local slab k (same order as team.local).  Copies `depth` owned rows across every slab
// boundary into the neighbour's halo.  utype: the array has nx+1 rows (face rows), else nx.
int nf_team_exchange(nf_team* team, const LevelGeom& geom, double* const* fields, int depth);
// sums `count` doubles at the same offset of every rank's buffer (device memory) and leaves the total in all of them
int nf_team_allreduce(nf_team* team, double* const* bufs, size_t count);
// replicated arrays (full size on every rank): rank r has computed rows [gb[r], ge[r]) (utype: the last rank also
// row nx); copy every rank's block into all the other ranks' arrays (NCCL: in-place broadcasts, one group)
int nf_team_share_rows(nf_team* team, int ld, int nx, const std::vector<int>& gb, const std::vector<int>& ge,
                       double* const* arrays, int utype);

// communicator plumbing (C-ABI wrappers in nf_slab.cu)
int nf_team_create_local(nf_ctx* ctx, int virtual_ranks, nf_team** out);
int nf_team_destroy(nf_team* team);

// ---- device memory of a team ------------------------------------------------------------------------------------
// Fields that take part in exchanges are allocated through the team: plain cudaMalloc normally, the peer-visible arena
// (nf_p2p.cu) when the ranks are separate processes.  elems = what this rank needs, elems_max = the maximum over the
// ranks (arena blocks must be identical on all ranks).  Zero-initialised (asynchronously on the context's stream).
double* nf_team_alloc(nf_team* team, size_t elems, size_t elems_max);
void nf_team_release(nf_team* team, void* ptr);
// peer-memory path (all collective over the team)
int nf_p2p_enable(nf_team* team, int rank);
int nf_p2p_reserve_stage(nf_team* team, size_t halo_elems);  // staging for halo exchanges of up to halo_elems doubles
void nf_p2p_destroy(nf_team* team);
bool nf_p2p_active(const nf_team* team);
int nf_p2p_exchange(nf_team* team, const LevelGeom& geom, double* field, int depth);
int nf_p2p_allreduce(nf_team* team, double* buf, size_t count);
// one launch: halo rows of up to two fields (levels may differ), optionally `zero`'s halo rows (geometry zgeom) cleared and
// red_count (<= 8) scalars summed over the ranks; NF_ERR_UNSUPPORTED = not possible here, use the separate collectives
int nf_p2p_exchange_multi(nf_team* team, int nfields, const LevelGeom* const* geoms, double* const* fields, const int* depths,
                          const LevelGeom* zgeom, double* zero, double* red_buf, int red_count);
int nf_p2p_share_rows(nf_team* team, int ld, int nx, const std::vector<int>& gb, const std::vector<int>& ge, double* array,
                      int utype);
int nf_p2p_error(nf_team* team);
