// mcx_comm.cu — multi-GPU slab decomposition of the diffuse-and-react step (the reference has a single partition:
// src4/world.cpp:147,277; this is new).  One process per GPU; z-slabs of the device cell grid.
//
// Scheme: every rank holds its owned z-layers plus `halo_layers` of its neighbours' molecules and
// evaluates ALL of them.  Because a molecule's random stream is keyed by (seed, molecule id, iteration) and the
// conflict rule is deterministic, a halo molecule is evaluated on the neighbour exactly as on its owner, so
// reactions across a slab face are decided identically on both sides without any mid-iteration message.
// After the evaluation each rank keeps the records whose NEW position it owns (this also moves emigrants: the
// receiving rank computed them itself) and the only exchange per iteration is the halo refresh: every kept
// record within halo_layers of a face is sent to that neighbour (grouped ncclSend/ncclRecv over NVLink), the
// neighbour bins it into its next snapshot together with its own results.
#include "mcx_comm.h"

#include <nccl.h>
#include <unistd.h>
#include <cstring>

struct McxComm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  std::string err;
  HaloRec* send[2] = {nullptr, nullptr};
  HaloRec* recv[2] = {nullptr, nullptr};
  unsigned int cap = 0;
  unsigned int* d_counts = nullptr;   // [0..1] send counts (low, high), [2..3] receive counts
  unsigned int* h_counts = nullptr;   // pinned mirror
  unsigned long long* d_red = nullptr;
  int red_cap = 0;
  unsigned int* d_fresh = nullptr;    // [world] fresh molecule ids of every rank this iteration (all-gathered), [world]: mine
  // peer-memory halo path: the pack kernel stores straight into the neighbour's buffers over NVLink
  bool p2p = false;
  unsigned long long id_floor = 0;  // global maximum of next_id at the last refresh (before the per-rank alignment)
  char* block = nullptr;                 // my receive block: [side low|high][parity 0|1] record buffers + 4 flag words
  char* peer_base[2] = {nullptr, nullptr};   // the low / high neighbour's block as mapped here
  bool peer_ipc[2] = {false, false};
  unsigned int* d_done = nullptr;        // block counter of the pack kernel
  unsigned int xchg = 0;                 // exchanges so far (same on every rank): flag tag and buffer parity
};
static size_t side_bytes(unsigned int cap) { return 2 * sizeof(HaloRec) * (size_t)cap; }
static size_t block_bytes(unsigned int cap) { return 2 * side_bytes(cap) + 256; }
static HaloRec* block_recv(char* base, unsigned int cap, int side, int parity) {
  return (HaloRec*)(base + side * side_bytes(cap) + parity * sizeof(HaloRec) * (size_t)cap);
}
static unsigned long long* block_flag(char* base, unsigned int cap, int side, int parity) {
  return (unsigned long long*)(base + 2 * side_bytes(cap)) + (side * 2 + parity);
}

#define NCK(call)                                                                         \
  do {                                                                                    \
    ncclResult_t r_ = (call);                                                             \
    if (r_ != ncclSuccess) { c->err = std::string(#call) + ": " + ncclGetErrorString(r_); return MCX_ERR_COMM; } \
  } while (0)
#define CCK(call)                                                                         \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) { c->err = std::string(#call) + ": " + cudaGetErrorString(e_); return MCX_ERR_CUDA; } \
  } while (0)

extern "C" int mcx_comm_unique_id(void* out, uint32_t bytes) {
  if (!out || bytes < sizeof(ncclUniqueId)) return MCX_ERR_INVALID_ARG;
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return MCX_ERR_COMM;
  memcpy(out, &id, sizeof(id));
  return (int)sizeof(id);
}


// ---- peer-memory setup ------------------------------------------------------------------------------------------------
// Every rank allocates one receive block and hands it to both neighbours: as a CUDA IPC handle (one process per GPU,
// the torchrun layout) or, when the ranks are threads of one process (tests/test_multi_gpu.py), as the pointer itself
// with peer access enabled.  The handles travel over the NCCL communicator that exists anyway.
struct PeerBlob { cudaIpcMemHandle_t handle; unsigned long long pid, ptr; int device, ok; char pad[128 - sizeof(cudaIpcMemHandle_t) - 24]; };
static_assert(sizeof(PeerBlob) == 128, "PeerBlob layout");

static void setup_p2p(McxComm* c) {
  int dev = 0;
  cudaGetDevice(&dev);
  bool ok = cudaMalloc((void**)&c->block, block_bytes(c->cap)) == cudaSuccess;
  ok = ok && cudaMemset(c->block, 0, block_bytes(c->cap)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&c->d_done, sizeof(unsigned int)) == cudaSuccess;
  ok = ok && cudaMemset(c->d_done, 0, sizeof(unsigned int)) == cudaSuccess;
  PeerBlob mine;
  memset(&mine, 0, sizeof(mine));
  ok = ok && cudaIpcGetMemHandle(&mine.handle, c->block) == cudaSuccess;
  mine.pid = (unsigned long long)getpid(); mine.ptr = (unsigned long long)(uintptr_t)c->block; mine.device = dev; mine.ok = ok ? 1 : 0;
  PeerBlob* d_blob = nullptr;  // [0] mine, [1] from the low neighbour, [2] from the high neighbour
  if (cudaMalloc((void**)&d_blob, 3 * sizeof(PeerBlob)) != cudaSuccess) return;
  cudaMemset(d_blob, 0, 3 * sizeof(PeerBlob));
  cudaMemcpy(d_blob, &mine, sizeof(mine), cudaMemcpyHostToDevice);
  const bool lo = c->rank > 0, hi = c->rank < c->world - 1;
  ncclGroupStart();
  if (lo) { ncclSend(d_blob, sizeof(PeerBlob), ncclChar, c->rank - 1, c->comm, 0); ncclRecv(d_blob + 1, sizeof(PeerBlob), ncclChar, c->rank - 1, c->comm, 0); }
  if (hi) { ncclSend(d_blob, sizeof(PeerBlob), ncclChar, c->rank + 1, c->comm, 0); ncclRecv(d_blob + 2, sizeof(PeerBlob), ncclChar, c->rank + 1, c->comm, 0); }
  ncclGroupEnd();
  cudaStreamSynchronize(0);
  PeerBlob got[3];
  cudaMemcpy(got, d_blob, sizeof(got), cudaMemcpyDeviceToHost);
  cudaFree(d_blob);
  for (int side = 0; side < 2 && ok; side++) {
    if (!(side == 0 ? lo : hi)) continue;
    const PeerBlob& b = got[1 + side];
    if (!b.ok) { ok = false; break; }
    if (b.pid == mine.pid) {  // same process: the pointer is valid here once peer access is on
      int can = 0;
      cudaDeviceCanAccessPeer(&can, dev, b.device);
      cudaError_t e = can ? cudaDeviceEnablePeerAccess(b.device, 0) : cudaErrorInvalidDevice;
      if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
      ok = e == cudaSuccess;
      c->peer_base[side] = (char*)(uintptr_t)b.ptr;
    } else {
      void* q = nullptr;
      ok = cudaIpcOpenMemHandle(&q, b.handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      c->peer_base[side] = (char*)q; c->peer_ipc[side] = ok;
    }
  }
  if (!ok) cudaGetLastError();
  // all ranks take the same path
  unsigned long long v = ok ? 1ull : 0ull;
  cudaMemcpy(c->d_red, &v, sizeof(v), cudaMemcpyHostToDevice);
  ncclAllReduce(c->d_red, c->d_red, 1, ncclUint64, ncclMin, c->comm, 0);
  cudaStreamSynchronize(0);
  cudaMemcpy(&v, c->d_red, sizeof(v), cudaMemcpyDeviceToHost);
  c->p2p = v == 1ull;
}

McxComm* mcx_comm_create(const void* nccl_unique_id, uint32_t id_bytes, int rank, int world_size, unsigned int halo_capacity,
                         std::string& err) {
  if (id_bytes < sizeof(ncclUniqueId)) { err = "ncclUniqueId too short"; return nullptr; }
  if (world_size < 2 || rank < 0 || rank >= world_size) { err = "mcx_comm_init needs world_size >= 2 and 0 <= rank < world_size"; return nullptr; }
  McxComm* c = new McxComm();
  c->rank = rank; c->world = world_size; c->cap = halo_capacity;
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof(id));
  ncclResult_t r = ncclCommInitRank(&c->comm, world_size, id, rank);
  if (r != ncclSuccess) { err = std::string("ncclCommInitRank: ") + ncclGetErrorString(r); delete c; return nullptr; }
  bool ok = true;
  ok = ok && cudaMalloc((void**)&c->d_counts, 4 * sizeof(unsigned int)) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&c->h_counts, 4 * sizeof(unsigned int)) == cudaSuccess;
  c->red_cap = 1024;
  ok = ok && cudaMalloc((void**)&c->d_red, sizeof(unsigned long long) * c->red_cap) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&c->d_fresh, sizeof(unsigned int) * (size_t)(world_size + 1)) == cudaSuccess;
  if (ok) cudaMemset(c->d_fresh, 0, sizeof(unsigned int) * (size_t)(world_size + 1));
  if (ok) cudaMemset(c->d_counts, 0, 4 * sizeof(unsigned int));
  if (ok && !getenv("MCX_HALO_NCCL")) setup_p2p(c);  // falls back to NCCL send/recv when peer memory is not available
  if (ok && !c->p2p) {  // staging buffers of the NCCL path
    for (int k = 0; k < 2; k++) {
      ok = ok && cudaMalloc((void**)&c->send[k], sizeof(HaloRec) * (size_t)halo_capacity) == cudaSuccess;
      ok = ok && cudaMalloc((void**)&c->recv[k], sizeof(HaloRec) * (size_t)halo_capacity) == cudaSuccess;
    }
  }
  if (!ok) { err = "halo buffer allocation failed"; mcx_comm_destroy(c); return nullptr; }
  return c;
}

void mcx_comm_destroy(McxComm* c) {
  if (!c) return;
  for (int k = 0; k < 2; k++) { if (c->send[k]) cudaFree(c->send[k]); if (c->recv[k]) cudaFree(c->recv[k]); }
  if (c->d_counts) cudaFree(c->d_counts);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->d_red) cudaFree(c->d_red);
  if (c->d_fresh) cudaFree(c->d_fresh);
  for (int k = 0; k < 2; k++) if (c->peer_ipc[k] && c->peer_base[k]) cudaIpcCloseMemHandle(c->peer_base[k]);
  if (c->block) cudaFree(c->block);
  if (c->d_done) cudaFree(c->d_done);
  if (c->comm) ncclCommDestroy(c->comm);
  delete c;
}
const char* mcx_comm_error(McxComm* c) { return c ? c->err.c_str() : ""; }
bool mcx_comm_is_p2p(const McxComm* c) { return c && c->p2p; }
const uint32_t* mcx_comm_rank_fresh(const McxComm* c) { return c ? c->d_fresh : nullptr; }
unsigned long long mcx_comm_id_floor(const McxComm* c) { return c ? c->id_floor : 0ull; }

// halo refresh: pack -> counts -> payload -> unpack (appended behind the local results in B)
// halo refresh over peer memory: ONE kernel selects the records, stores them into the neighbours' receive buffers over
// NVLink (32 bytes per record, 48 when it carries cold fields) and publishes the counts with a release store; the
// unpack kernel of the receiving rank acquires the flag and bins the records.  No NCCL call, no host round trip: the
// host keeps enqueueing iterations.  Buffers alternate with the parity of the exchange: a rank can be at most one
// exchange ahead of its neighbour (its unpack waits for the neighbour's pack), so two buffers never collide.
static int exchange_halo_p2p(McxComm* c, DevParams& p, cudaStream_t s) {
  c->xchg++;
  const int parity = (int)(c->xchg & 1u);
  HaloP2P L;
  L.tag = c->xchg; L.cap = c->cap; L.done = c->d_done;
  for (int side = 0; side < 2; side++) {
    const bool has = side == 0 ? p.has_low != 0 : p.has_high != 0;
    // I am the high-side neighbour of my low neighbour: my records land in ITS high-side buffers, and vice versa
    L.peer_recv[side] = has ? block_recv(c->peer_base[side], c->cap, 1 - side, parity) : nullptr;
    L.peer_flag[side] = has ? block_flag(c->peer_base[side], c->cap, 1 - side, parity) : nullptr;
    L.my_recv[side] = block_recv(c->block, c->cap, side, parity);
    L.my_flag[side] = block_flag(c->block, c->cap, side, parity);
  }
  mcx_launch_halo_p2p(p, L, s);
  return MCX_OK;
}

static int exchange_halo(McxComm* c, DevParams& p, cudaStream_t s) {
  if (c->p2p) return exchange_halo_p2p(c, p, s);
  const bool lo = p.has_low != 0, hi = p.has_high != 0;
  mcx_launch_pack_halo(p, c->send[0], c->send[1], c->cap, s);
  CCK(cudaMemcpyAsync(c->d_counts, &p.ctr->n_send[0], 2 * sizeof(unsigned int), cudaMemcpyDeviceToDevice, s));
  CCK(cudaMemsetAsync(c->d_counts + 2, 0, 2 * sizeof(unsigned int), s));
  NCK(ncclGroupStart());
  if (lo) { NCK(ncclSend(c->d_counts + 0, 1, ncclUint32, c->rank - 1, c->comm, s)); NCK(ncclRecv(c->d_counts + 2, 1, ncclUint32, c->rank - 1, c->comm, s)); }
  if (hi) { NCK(ncclSend(c->d_counts + 1, 1, ncclUint32, c->rank + 1, c->comm, s)); NCK(ncclRecv(c->d_counts + 3, 1, ncclUint32, c->rank + 1, c->comm, s)); }
  NCK(ncclGroupEnd());
  CCK(cudaMemcpyAsync(c->h_counts, c->d_counts, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
  CCK(cudaStreamSynchronize(s));
  const unsigned int ns0 = c->h_counts[0], ns1 = c->h_counts[1], nr0 = c->h_counts[2], nr1 = c->h_counts[3];
  if (ns0 > c->cap || ns1 > c->cap || nr0 > c->cap || nr1 > c->cap) { c->err = "halo buffer capacity exceeded (raise max_molecules)"; return MCX_ERR_CAPACITY; }
  NCK(ncclGroupStart());
  if (lo && ns0) NCK(ncclSend(c->send[0], sizeof(HaloRec) * (size_t)ns0, ncclChar, c->rank - 1, c->comm, s));
  if (lo && nr0) NCK(ncclRecv(c->recv[0], sizeof(HaloRec) * (size_t)nr0, ncclChar, c->rank - 1, c->comm, s));
  if (hi && ns1) NCK(ncclSend(c->send[1], sizeof(HaloRec) * (size_t)ns1, ncclChar, c->rank + 1, c->comm, s));
  if (hi && nr1) NCK(ncclRecv(c->recv[1], sizeof(HaloRec) * (size_t)nr1, ncclChar, c->rank + 1, c->comm, s));
  NCK(ncclGroupEnd());
  mcx_launch_unpack_halo(p, c->recv[0], nr0, 0, s);
  mcx_launch_unpack_halo(p, c->recv[1], nr1, nr0, s);
  mcx_launch_add_received(p, nr0 + nr1, s);
  return MCX_OK;
}

// Wall::has_initialized_grid is a property of the whole mesh: a wall has its grid as soon as ANY rank holds a surface
// molecule on it (the scatter marks the walls of the records a rank sees; the neighbour search of the surface-surface
// reactions must not depend on where the slabs are cut).  One byte per wall, models with surface-surface classes only.
static int share_wall_grids(McxComm* c, DevParams& p, cudaStream_t s) {
  if (!p.surfsurf || !p.wall_has_grid || p.n_walls <= 0) return MCX_OK;
  NCK(ncclAllReduce(p.wall_has_grid, p.wall_has_grid, (size_t)p.n_walls, ncclUint8, ncclMax, c->comm, s));
  return MCX_OK;
}

int mcx_comm_iteration(McxComm* c, DevParams& p, const StepPlan& plan, cudaStream_t s) {
  mcx_launch_evaluate(p, plan, s);
  if (plan.has_fresh) {
    // fresh molecule ids are global: every rank learns how many each rank hands out this iteration (FreshEvent order:
    // the ranks own consecutive ranges of cell groups), then writes its own into the product records before the halo
    // refresh copies them to the neighbours
    CCK(cudaMemcpyAsync(c->d_fresh + c->world, &p.ctr->n_fresh_ids, sizeof(unsigned int), cudaMemcpyDeviceToDevice, s));
    NCK(ncclAllGather(c->d_fresh + c->world, c->d_fresh, 1, ncclUint32, c->comm, s));
    mcx_launch_fresh_scan(p, plan, s);
    mcx_launch_assign_ids(p, plan, s);
  }
  int rc = exchange_halo(c, p, s);
  if (rc) return rc;
  if (plan.launches) *plan.launches += 4;
  mcx_launch_sort(p, plan, s);
  if (plan.prof) cudaEventRecord(plan.prof[3], s);
  return share_wall_grids(c, p, s);
}

int mcx_comm_refresh(McxComm* c, DevParams& p, const StepPlan& plan, cudaStream_t s) {
  // fresh molecule ids start above the global maximum on every rank
  unsigned int next_id = 0;
  CCK(cudaMemcpyAsync(&next_id, &p.ctr->next_id, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
  CCK(cudaStreamSynchronize(s));
  unsigned long long v = next_id;
  CCK(cudaMemcpyAsync(c->d_red, &v, sizeof(v), cudaMemcpyHostToDevice, s));
  NCK(ncclAllReduce(c->d_red, c->d_red, 1, ncclUint64, ncclMax, c->comm, s));
  CCK(cudaMemcpyAsync(&v, c->d_red, sizeof(v), cudaMemcpyDeviceToHost, s));
  CCK(cudaStreamSynchronize(s));
  c->id_floor = v;  // every id below v may be in use, none at or above it
  next_id = (unsigned int)v;  // the same on every rank: fresh ids are handed out in one global order (k_assign_ids)
  CCK(cudaMemcpyAsync(&p.ctr->next_id, &next_id, sizeof(unsigned int), cudaMemcpyHostToDevice, s));
  mcx_launch_rebin(p, plan, s);
  int rc = exchange_halo(c, p, s);
  if (rc) return rc;
  mcx_launch_sort(p, plan, s);
  return share_wall_grids(c, p, s);
}

int mcx_comm_allreduce_u64(McxComm* c, unsigned long long* host_buf, int n, cudaStream_t s) {
  if (n > c->red_cap) { c->err = "allreduce buffer too small"; return MCX_ERR_INVALID_ARG; }
  CCK(cudaMemcpyAsync(c->d_red, host_buf, sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, s));
  NCK(ncclAllReduce(c->d_red, c->d_red, n, ncclUint64, ncclSum, c->comm, s));
  CCK(cudaMemcpyAsync(host_buf, c->d_red, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, s));
  CCK(cudaStreamSynchronize(s));
  return MCX_OK;
}
