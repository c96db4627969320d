// mcx_api.cu — the C ABI of libmcx (include/mcx.h): handle, device memory, table upload, step driver.
// There is no CPU fallback: every compute entry point needs a CUDA device and returns MCX_ERR_CUDA
// (with text) when none is usable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "mcx_internal.h"
#include "mcx_philox.h"
#include "mcx_geom.h"
#include "mcx_comm.h"


static thread_local std::string g_create_error;

struct mcx_handle {
  mcx_config cfg{};
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 148;
  DevParams p{};
  StepPlan plan{};
  std::vector<void*> allocs;       // everything cudaMalloc'ed, freed in destroy
  // host copies of tables
  std::vector<mcx_species> species;
  std::vector<mcx_rxn_class> classes;
  std::vector<mcx_pathway> pathways;
  std::vector<mcx_surf_class_rxn> surf_rules;
  std::vector<uint32_t> wall_class_host;
  uint32_t n_surf_classes = 0;
  bool has_geometry = false, has_species = false, uploaded = false;
  uint64_t iteration = 0;
  uint32_t *cs[2] = {nullptr, nullptr};
  int cs_cur = 0;
  unsigned int* scan_sums = nullptr;
  unsigned int* d_n_out = nullptr;
  // table device pointers that get replaced on re-set
  void *d_species = nullptr, *d_bimol = nullptr, *d_unimol = nullptr, *d_classes = nullptr, *d_pathways = nullptr,
       *d_surf = nullptr, *d_walls = nullptr, *d_tri = nullptr, *d_verts = nullptr, *d_wclass = nullptr,
       *d_spw_start = nullptr, *d_spw_list = nullptr, *d_fw_start = nullptr, *d_fw_list = nullptr, *d_sp_flags = nullptr, *d_grids = nullptr, *d_volsurf = nullptr,
       *d_tile_slot = nullptr, *d_exd_skip = nullptr, *d_wall_cv = nullptr, *d_rxn_count_cv = nullptr,
       *d_mol_count_cv = nullptr, *d_edges = nullptr, *d_tile_claim = nullptr;
  uint32_t n_cv = 1; uint32_t* st_cv = nullptr;
  void *d_cv_mask = nullptr;
  void *d_wall_obj = nullptr, *d_surf_rxn = nullptr, *d_surf_border = nullptr, *d_wall_border = nullptr;
  std::vector<double> wall_area_host;
  void *d_wall_rs = nullptr, *d_rxn_count_rs = nullptr, *d_mol_count_rs = nullptr; uint32_t n_rs = 0;
  uint64_t n_walls_host = 0;
  bool has_surf = false, surf_allocated = false;
  // surface-surface reactions: host copy of the mesh (the neighbour-tile table is built when a table with such classes and
  // a geometry are both there) and the device tables
  bool has_surfsurf = false;
  bool has_general = false;   // some pathway puts surface products on vacant neighbour tiles (DevPathway::general)
  std::vector<double> geom_verts; std::vector<uint32_t> geom_tri, geom_wall_object;
  void *d_surfsurf = nullptr, *d_tn_start = nullptr, *d_tn_list = nullptr, *d_wall_has_grid = nullptr;
  uint32_t *st_wall = nullptr, *st_tile = nullptr; int32_t* st_orient = nullptr; double *st_u = nullptr, *st_v = nullptr;
  McxComm* comm = nullptr;
  int ncz_global = 0;
  double *st_x = nullptr, *st_y = nullptr, *st_z = nullptr, *st_ts = nullptr, *st_tu = nullptr;
  uint32_t *st_id = nullptr, *st_sp = nullptr, *st_fl = nullptr;
  unsigned long long launches = 0;
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
};

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
      return MCX_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

template <typename T>
static int dev_alloc(mcx_handle* h, T** out, size_t count) {
  void* ptr = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CK(cudaMalloc(&ptr, bytes));
  CK(cudaMemset(ptr, 0, bytes));
  h->allocs.push_back(ptr);
  *out = (T*)ptr;
  return MCX_OK;
}
template <typename T>
static int dev_replace(mcx_handle* h, void** slot, const T* src, size_t count) {
  if (*slot) {
    cudaFree(*slot);
    h->allocs.erase(std::remove(h->allocs.begin(), h->allocs.end(), *slot), h->allocs.end());
    *slot = nullptr;
  }
  T* d = nullptr;
  int rc = dev_alloc(h, &d, count);
  if (rc) return rc;
  if (count) CK(cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice));
  *slot = d;
  return MCX_OK;
}

extern "C" {

int mcx_abi_version(void) { return MCX_ABI_VERSION; }
uint32_t mcx_grid_num_tiles(const double* v9) { return mcxg::tri_num_tiles(v9); }
void mcx_grid2uv(const double* v9, uint32_t tile, double* uv2) { mcxg::tri_grid2uv(v9, tile, uv2); }
uint32_t mcx_xyz2grid(const double* v9, const double* xyz3) { return mcxg::tri_xyz2grid(v9, xyz3); }
uint64_t mcx_walls_per_subpart(const double* origin3, double partition_edge_length, uint32_t n_subparts_per_edge,
                               double rxn_radius_3d, uint32_t use_expanded_list, const double* vertices, uint64_t n_vertices,
                               const uint32_t* tri, uint64_t n_walls, uint32_t* start_out, uint32_t* list_out, uint64_t cap) {
  (void)n_vertices;
  std::vector<DevWall> walls;
  mcxg::wall_constants(vertices, tri, n_walls, walls);
  const double sp_len = partition_edge_length / n_subparts_per_edge;
  mcxg::GridSpec g{origin3[0], origin3[1], origin3[2], sp_len, 1.0 / sp_len, rxn_radius_3d, (int)n_subparts_per_edge, use_expanded_list != 0};
  std::vector<uint32_t> start, list;
  mcxg::bin_walls(g, vertices, tri, walls, start, list);
  for (size_t i = 0; i < start.size(); i++) start_out[i] = start[i];
  for (size_t i = 0; i < list.size() && i < cap; i++) list_out[i] = list[i];
  return list.size();
}

uint64_t mcx_tile_neighbor_table(const double* vertices, uint64_t n_vertices, const uint32_t* tri, uint64_t n_walls,
                                 const uint32_t* wall_object, uint32_t* start_out, uint32_t* list_out, uint64_t cap) {
  std::vector<DevWall> walls; std::vector<DevGrid> grids; std::vector<DevEdge> edges;
  mcxg::wall_constants(vertices, tri, n_walls, walls);
  const uint64_t n_tiles = mcxg::grid_constants(vertices, tri, walls, grids);
  mcxg::edge_constants(vertices, tri, walls, wall_object, edges);
  std::vector<uint32_t> start, list;
  mcxg::tile_neighbor_table(vertices, n_vertices, tri, walls, grids, edges, start, list);
  for (uint64_t i = 0; i <= n_tiles; i++) start_out[i] = start[i];
  for (size_t i = 0; i < list.size() && i < 2 * cap; i++) list_out[i] = list[i];
  return list.size() / 2;
}

const char* mcx_last_error(const mcx_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void mcx_philox_block(uint64_t seed, uint32_t mol_id, uint64_t iteration, uint32_t block, uint32_t out[4]) {
  philox4x32_10(block, (uint32_t)iteration, (uint32_t)(iteration >> 32), mol_id, (uint32_t)seed, (uint32_t)(seed >> 32), out);
}

// cells of the local grid incl. the padding of the (y, z) row blocks (row_index(), mcx_device.cuh)
static void set_cell_count(DevParams& p) {
  const int B = 1 << p.rb_log2;
  p.nby = (p.ncy + B - 1) / B;
  const int nbz = (p.ncz + B - 1) / B;
  p.n_cells = (unsigned int)((size_t)p.ncx * ((size_t)p.nby * B) * ((size_t)nbz * B));
  p.grp_x = (unsigned int)((p.ncx + 15) / 16);
  p.n_groups = (unsigned int)((size_t)p.n_cells / (size_t)p.ncx * p.grp_x);
}

int mcx_create(const mcx_config* cfg, mcx_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return MCX_ERR_INVALID_ARG; }
  if (cfg->abi_version != MCX_ABI_VERSION) { g_create_error = "ABI version mismatch"; return MCX_ERR_INVALID_ARG; }
  if (cfg->num_subparts_per_edge == 0 || cfg->num_subparts_per_edge > 300 || !(cfg->partition_edge_length > 0) ||
      cfg->max_molecules == 0 || cfg->max_molecules > 0xFFFFFFF0ull) {
    g_create_error = "invalid configuration (subpartitions, partition size or max_molecules)";
    return MCX_ERR_INVALID_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device available (libmcx has no CPU fallback): ") + cudaGetErrorString(e);
    return MCX_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return MCX_ERR_INVALID_ARG; }
  mcx_handle* h = new mcx_handle();
  h->cfg = *cfg;
  auto fail = [&](int rc) { g_create_error = h->err; mcx_destroy(h); return rc; };
  if (cudaSetDevice(cfg->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return fail(MCX_ERR_CUDA); }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
    h->err = "stream/event creation failed"; return fail(MCX_ERR_CUDA);
  }
  DevParams& p = h->p;
  p.ox = cfg->origin[0]; p.oy = cfg->origin[1]; p.oz = cfg->origin[2];
  p.part_len = cfg->partition_edge_length;
  p.n_sp = (int)cfg->num_subparts_per_edge;
  p.sp_len = cfg->partition_edge_length / cfg->num_subparts_per_edge;  // simulation_config.cpp:48
  p.sp_rcp = 1.0 / p.sp_len;                                          // simulation_config.cpp:63
  p.R = cfg->rxn_radius_3d;
  p.use_expanded = cfg->use_expanded_list ? 1 : 0;
  p.seed = cfg->seed;
  p.rng_mode = (int)cfg->rng_mode;
  p.capacity = (unsigned int)cfg->max_molecules;
  p.max_rounds = cfg->max_resolve_rounds ? cfg->max_resolve_rounds : 8;
  if (p.max_rounds > MCX_ROUNDS_MAX) p.max_rounds = MCX_ROUNDS_MAX;
  h->iteration = cfg->initial_iteration;

  // device neighbour-cell grid over the active box
  double lo[3], hi[3];
  bool whole = true;
  for (int k = 0; k < 3; k++) whole = whole && cfg->active_llf[k] == 0 && cfg->active_urb[k] == 0;
  for (int k = 0; k < 3; k++) {
    lo[k] = whole ? cfg->origin[k] : cfg->active_llf[k];
    hi[k] = whole ? cfg->origin[k] + cfg->partition_edge_length : cfg->active_urb[k];
    if (!(hi[k] > lo[k])) { h->err = "empty active box"; return fail(MCX_ERR_INVALID_ARG); }
  }
  double vol = (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
  double edge = cfg->cell_edge;
  // default: about one molecule slot per cell — the candidate walk then touches ~1-2 records per molecule
  // (multi-GPU: max_molecules is the per-rank capacity, the active box is global)
  if (!(edge > 0)) edge = std::cbrt(vol * 1.0 / ((double)cfg->max_molecules * std::max(1, cfg->world_size)));
  {
    double clamp_r = 4.0;
    if (const char* e = getenv("MCX_CELL_CLAMP_R")) clamp_r = atof(e);  // tuning knob (profiles/)
    if (edge < clamp_r * cfg->rxn_radius_3d) edge = clamp_r * cfg->rxn_radius_3d;
  }
  // Anisotropic cells of volume edge^3: the records of one x-row of cells are contiguous in the sorted snapshot,
  // so a swept box costs one [start,end) lookup per (y,z) row whatever the x resolution.  Short x cells keep the
  // x-range tight; long y/z cells keep the box within 2x2 rows (the fast pass enumerates at most 4 rows).
  const double max_cells = 2.0e8;
  double ex, ey, ez;
  double ax = 2.25, ayz = 1.5;   // tuning knobs (profiles/): x divisor and y/z multiplier of the cell edge
  if (const char* e = getenv("MCX_CELL_AX")) ax = atof(e);
  if (const char* e = getenv("MCX_CELL_AYZ")) ayz = atof(e);
  for (;;) {
    ex = edge / ax; ey = ez = edge * ayz;
    double nc = std::ceil((hi[0] - lo[0]) / ex + 2) * std::ceil((hi[1] - lo[1]) / ey + 2) * std::ceil((hi[2] - lo[2]) / ez + 2);
    if (nc <= max_cells) break;
    edge *= 1.26;
  }
  p.cell_rcp_x = 1.0 / ex; p.cell_rcp_y = 1.0 / ey; p.cell_rcp_z = 1.0 / ez;
  p.cgx = lo[0] - 0.5 * ex; p.cgy = lo[1] - 0.5 * ey; p.cgz = lo[2] - 0.5 * ez;
  p.ncx = (int)std::ceil((hi[0] - p.cgx) / ex) + 1;
  p.ncy = (int)std::ceil((hi[1] - p.cgy) / ey) + 1;
  p.ncz = (int)std::ceil((hi[2] - p.cgz) / ez) + 1;
  // plain (y, z) row order: blocks of rows bought nothing (profiles/r02_g) and the order of the fresh molecule ids
  // (FreshEvent) must not depend on the local grid of a rank
  p.rb_log2 = 0;
  if (const char* e = getenv("MCX_ROW_BLOCK_LOG2")) p.rb_log2 = std::max(0, std::min(6, atoi(e)));  // tuning knob (profiles/)
  set_cell_count(p);
  p.own_lo = 0; p.own_hi = p.ncz; p.world = 1; p.halo_layers = 0; p.z_off = 0; p.has_low = 0; p.has_high = 0;
  h->ncz_global = p.ncz;

  const size_t cap = p.capacity;
  int rc = MCX_OK;
  rc |= dev_alloc(h, &p.recA, cap); rc |= dev_alloc(h, &p.recB, cap);
  rc |= dev_alloc(h, &p.tschedA, cap); rc |= dev_alloc(h, &p.tschedB, cap);
  rc |= dev_alloc(h, &p.tuniA, cap); rc |= dev_alloc(h, &p.tuniB, cap);
  rc |= dev_alloc(h, &p.rank, cap);
  rc |= dev_alloc(h, &p.claim, cap);
  rc |= dev_alloc(h, &p.prop_partner, cap); rc |= dev_alloc(h, &p.prop_info, cap); rc |= dev_alloc(h, &p.prop_t, cap);
  rc |= dev_alloc(h, &p.pend[0], cap); rc |= dev_alloc(h, &p.pend[1], cap);
  rc |= dev_alloc(h, &p.slow_list, cap);
  rc |= dev_alloc(h, &p.second_list, cap);
  rc |= dev_alloc(h, &p.slow2_list, cap);
  rc |= dev_alloc(h, &h->cs[0], (size_t)p.n_cells + 8); rc |= dev_alloc(h, &h->cs[1], (size_t)p.n_cells + 8);
  rc |= dev_alloc(h, &h->scan_sums, (size_t)(p.n_cells + 1) / 4096 + 16);
  p.scan_sums = h->scan_sums;
  rc |= dev_alloc(h, &p.ctr, 1);
  p.fresh_cap = (unsigned int)std::max<size_t>(1u << 16, cap / 8);
  rc |= dev_alloc(h, &p.fresh_list, p.fresh_cap);
  rc |= dev_alloc(h, &p.fresh_head, (size_t)p.n_groups + 8); rc |= dev_alloc(h, &p.fresh_pref, (size_t)p.n_groups + 8);
  p.rank_fresh = nullptr; p.my_rank = cfg->rank;
  rc |= dev_alloc(h, &h->d_n_out, 4);
  if (rc) return fail(MCX_ERR_CUDA);
  // empty wall tables until geometry arrives
  std::vector<uint32_t> zero_start((size_t)p.n_sp * p.n_sp * p.n_sp + 1, 0);
  if (dev_replace(h, &h->d_spw_start, zero_start.data(), zero_start.size())) return fail(MCX_ERR_CUDA);
  p.spw_start = (const uint32_t*)h->d_spw_start;
  p.fw_start = p.spw_start; p.fw_list = nullptr; p.fw_K = 1; p.fw_rcp = p.sp_rcp;
  std::vector<uint8_t> zero_flags(zero_start.size(), 0);
  if (dev_replace(h, &h->d_sp_flags, zero_flags.data(), zero_flags.size())) return fail(MCX_ERR_CUDA);
  p.sp_flags = (const uint8_t*)h->d_sp_flags;
  h->plan.sm_count = h->sm_count;
  p.sm_count = h->sm_count;
  h->plan.launches = &h->launches;
  *out = h;
  return MCX_OK;
}

void mcx_destroy(mcx_handle* h) {
  if (!h) return;
  if (h->comm) mcx_comm_destroy(h->comm);
  for (void* a : h->allocs) cudaFree(a);
  for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static int build_tile_neighbors(mcx_handle* h);

int mcx_set_geometry(mcx_handle* h, const double* vertices, uint64_t n_vertices, const uint32_t* tri, uint64_t n_walls,
                     const uint32_t* wall_surf_class, const uint32_t* wall_object) {
  if (!h) return MCX_ERR_INVALID_ARG;
  if ((n_walls && (!vertices || !tri)) || n_walls > 0xFFFFFFF0ull) { h->err = "bad geometry arrays"; return MCX_ERR_INVALID_ARG; }
  for (uint64_t i = 0; i < 3 * n_walls; i++)
    if (tri[i] >= n_vertices) { h->err = "triangle vertex index out of range"; return MCX_ERR_INVALID_ARG; }
  CK(cudaSetDevice(h->cfg.device));
  std::vector<DevWall> walls;
  mcxg::wall_constants(vertices, tri, n_walls, walls);
  mcxg::GridSpec g{h->p.ox, h->p.oy, h->p.oz, h->p.sp_len, h->p.sp_rcp, h->p.R, h->p.n_sp, h->p.use_expanded != 0};
  std::vector<uint32_t> start, list;
  mcxg::bin_walls(g, vertices, tri, walls, start, list);
  mcxg::wall_areas(vertices, tri, n_walls, h->wall_area_host);
  h->wall_class_host.assign(n_walls, MCX_NONE);
  if (wall_surf_class) h->wall_class_host.assign(wall_surf_class, wall_surf_class + n_walls);
  int rc = MCX_OK;
  rc |= dev_replace(h, &h->d_walls, walls.data(), walls.size());
  rc |= dev_replace(h, &h->d_tri, tri, 3 * n_walls);
  rc |= dev_replace(h, &h->d_verts, vertices, 3 * n_vertices);
  rc |= dev_replace(h, &h->d_wclass, h->wall_class_host.data(), h->wall_class_host.size());
  {
    std::vector<uint32_t> wobj(std::max<uint64_t>(n_walls, 1), 0u);
    if (wall_object) wobj.assign(wall_object, wall_object + n_walls);
    rc |= dev_replace(h, &h->d_wall_obj, wobj.data(), wobj.size());
  }
  rc |= dev_replace(h, &h->d_spw_start, start.data(), start.size());
  rc |= dev_replace(h, &h->d_spw_list, list.data(), list.size());
  // fine wall grid: K^3 cells per subpartition (K == 1: the subpartition lists themselves)
  h->p.fw_K = mcxg::fine_wall_factor(g, vertices, tri, n_walls, start);
  if (const char* e = getenv("MCX_FINE_WALL_K")) h->p.fw_K = std::max(1, std::min(16, atoi(e)));  // tuning knob (profiles/)
  h->p.fw_rcp = (double)h->p.fw_K / h->p.sp_len;
  if (h->p.fw_K > 1) {
    std::vector<uint32_t> fstart, flist;
    mcxg::bin_walls_fine(g, vertices, tri, start, list, h->p.fw_K, MCX_FW_MARGIN, fstart, flist);
    rc |= dev_replace(h, &h->d_fw_start, fstart.data(), fstart.size());
    rc |= dev_replace(h, &h->d_fw_list, flist.data(), flist.size());
  }
  // surface grids of every wall and the tile table (Grid::molecules_per_tile, one entry per tile of every wall)
  std::vector<DevGrid> grids;
  const uint64_t n_tiles = mcxg::grid_constants(vertices, tri, walls, grids);
  if (n_tiles > 0xFFFFFFF0ull) { h->err = "too many surface tiles"; return MCX_ERR_INVALID_ARG; }
  rc |= dev_replace(h, &h->d_grids, grids.data(), grids.size());
  {
    std::vector<uint32_t> vacant(std::max<uint64_t>(n_tiles, 1), MCX_NONE);
    rc |= dev_replace(h, &h->d_tile_slot, vacant.data(), vacant.size());
  }
  h->p.grids = (const DevGrid*)h->d_grids; h->p.tile_slot = (uint32_t*)h->d_tile_slot; h->p.n_tiles = (unsigned int)n_tiles;
  {
    std::vector<DevEdge> edges;
    mcxg::edge_constants(vertices, tri, walls, wall_object, edges);
    rc |= dev_replace(h, &h->d_edges, edges.data(), edges.size());
    std::vector<unsigned long long> zero(std::max<uint64_t>(n_tiles, 1), 0ull);
    rc |= dev_replace(h, &h->d_tile_claim, zero.data(), zero.size());
    h->p.edges = (const DevEdge*)h->d_edges; h->p.tile_claim = (unsigned long long*)h->d_tile_claim;
  }
  // per-subpartition wall flags for the fast diffuse pass: bit0 = holds walls, bit1 = 3x3x3 neighbourhood does
  {
    const int n = h->p.n_sp;
    std::vector<uint8_t> flags((size_t)n * n * n + 1, 0);
    auto at = [&](int x, int y, int z) { return (size_t)x + (size_t)y * n + (size_t)z * n * n; };
    for (size_t sidx = 0; sidx + 1 < start.size(); sidx++)
      if (start[sidx + 1] > start[sidx]) flags[sidx] |= 1;
    for (int z = 0; z < n; z++)
      for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
          if (!(flags[at(x, y, z)] & 1)) continue;
          for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
              for (int dx = -1; dx <= 1; dx++) {
                int X = x + dx, Y = y + dy, Z = z + dz;
                if (X < 0 || Y < 0 || Z < 0 || X >= n || Y >= n || Z >= n) continue;
                flags[at(X, Y, Z)] |= 2;
              }
        }
    rc |= dev_replace(h, &h->d_sp_flags, flags.data(), flags.size());
  }
  if (rc) return MCX_ERR_CUDA;
  DevParams& p = h->p;
  p.walls = (const DevWall*)h->d_walls; p.wall_tri = (const uint32_t*)h->d_tri; p.verts = (const double*)h->d_verts;
  p.wall_class = (const uint32_t*)h->d_wclass; p.spw_start = (const uint32_t*)h->d_spw_start;
  p.wall_obj = (const uint32_t*)h->d_wall_obj;
  p.fw_start = p.fw_K > 1 ? (const uint32_t*)h->d_fw_start : (const uint32_t*)h->d_spw_start;
  p.fw_list = p.fw_K > 1 ? (const uint32_t*)h->d_fw_list : (const uint32_t*)h->d_spw_list;
  p.spw_list = (const uint32_t*)h->d_spw_list; p.n_walls = (int)n_walls; p.sp_flags = (const uint8_t*)h->d_sp_flags;
  if (h->p.wall_cv && h->n_walls_host != n_walls) {
    // the per-wall counted-volume table belongs to the previous geometry: drop it (mcx_set_counted_volumes again)
    h->p.wall_cv = nullptr; h->p.n_cv = 1; h->n_cv = 1; h->p.cv_mask = nullptr; h->p.cv_xor = 0; h->p.cv_all = 0;
  }
  if (h->p.wall_border && h->n_walls_host != n_walls) h->p.wall_border = nullptr;   // likewise
  if (h->p.wall_rs && h->n_walls_host != n_walls) { h->p.wall_rs = nullptr; h->p.n_rs = 0; h->n_rs = 0; }  // likewise
  h->n_walls_host = n_walls;
  h->has_geometry = true;
  h->geom_verts.assign(vertices, vertices + 3 * n_vertices);
  h->geom_tri.assign(tri, tri + 3 * n_walls);
  h->geom_wall_object.clear();
  if (wall_object) h->geom_wall_object.assign(wall_object, wall_object + n_walls);
  {
    // Wall::has_initialized_grid per wall (a wall gets its grid with its first surface molecule): set by the scatter
    std::vector<uint8_t> none(std::max<uint64_t>(n_walls, 1), 0);
    if (dev_replace(h, &h->d_wall_has_grid, none.data(), none.size())) return MCX_ERR_CUDA;
    h->p.wall_has_grid = (uint8_t*)h->d_wall_has_grid;
  }
  { const int trc = build_tile_neighbors(h); if (trc != MCX_OK) return trc; }
  return MCX_OK;
}

// (re)build every table that depends on species x reactions x surface classes
// neighbour tiles of every tile for react_2D_all_neighbors (mcx_geom.cpp: tile_neighbor_table), built once per geometry
static int build_tile_neighbors(mcx_handle* h) {
  h->p.tn_start = nullptr; h->p.tn_list = nullptr;
  if (!h->has_surfsurf || !h->has_geometry || h->n_walls_host == 0) return MCX_OK;
  const uint64_t n_walls = h->n_walls_host;
  std::vector<DevWall> walls; std::vector<DevGrid> grids; std::vector<DevEdge> edges;
  mcxg::wall_constants(h->geom_verts.data(), h->geom_tri.data(), n_walls, walls);
  mcxg::grid_constants(h->geom_verts.data(), h->geom_tri.data(), walls, grids);
  mcxg::edge_constants(h->geom_verts.data(), h->geom_tri.data(), walls, h->geom_wall_object.empty() ? nullptr : h->geom_wall_object.data(), edges);
  std::vector<uint32_t> start, list;
  mcxg::tile_neighbor_table(h->geom_verts.data(), h->geom_verts.size() / 3, h->geom_tri.data(), walls, grids, edges, start, list);
  if (list.empty()) list.assign(2, MCX_NONE);
  int rc = MCX_OK;
  rc |= dev_replace(h, &h->d_tn_start, start.data(), start.size());
  rc |= dev_replace(h, &h->d_tn_list, list.data(), list.size());
  if (rc) return MCX_ERR_CUDA;
  h->p.tn_start = (const uint32_t*)h->d_tn_start; h->p.tn_list = (const uint2*)h->d_tn_list;
  return MCX_OK;
}

static int rebuild_tables(mcx_handle* h) {
  const size_t ns = h->species.size();
  if (ns == 0) return MCX_OK;
  if (ns > MCX_MAX_COUNTED) { h->err = "more than 1024 species are not supported by the device counters"; return MCX_ERR_INVALID_ARG; }
  std::vector<int> bimol(ns * ns, -1), unimol(ns, -1), volsurf(ns * ns, -1), surfsurf(ns * ns, -1);
  bool any_surfsurf = false;
  bool any_surf = false;
  for (size_t a = 0; a < ns; a++) any_surf = any_surf || !(h->species[a].flags & MCX_SP_VOL);
  if (h->classes.size() > 8191) { h->err = "more than 8191 reaction classes (proposal word)"; return MCX_ERR_INVALID_ARG; }
  std::vector<DevClass> dc(h->classes.size());
  std::vector<DevPathway> dp(h->pathways.size());
  std::vector<uint8_t> general(h->pathways.size(), 0);
  // a pathway that creates more surface products than it consumes surface reactants puts the extra ones on vacant
  // neighbour tiles (find_surf_product_positions' general branch): what it needs from the table
  auto check_general = [&](const mcx_rxn_class& rc, uint32_t q, int consumed_surf, int kept_surf) -> const char* {
    const mcx_pathway& pw = h->pathways[rc.first_pathway + q];
    if (!(pw.kept_info & MCX_KEPT_VALID)) return "a pathway with surface products on vacant neighbour tiles needs kept_info (the order of the rule's products)";
    if (consumed_surf > 0 && kept_surf > 0) return "a pathway with surface products on vacant neighbour tiles that keeps one surface reactant and consumes another is not supported";
    general[rc.first_pathway + q] = 1;
    return nullptr;
  };
  for (size_t c = 0; c < h->classes.size(); c++) {
    const mcx_rxn_class& rc = h->classes[c];
    if ((uint64_t)rc.first_pathway + (uint64_t)rc.n_pathways > (uint64_t)h->pathways.size() || rc.n_pathways == 0 || rc.n_pathways > 4095) {
      h->err = "reaction class pathway range invalid"; return MCX_ERR_INVALID_ARG;
    }
    if (rc.kind == MCX_RXN_BIMOL_VOLWALL) {
      // reactants[0]: a volume species or MCX_ALL_*; reactants[1]: a surface class (the class is reached through the
      // surface-class rules, not through the species tables)
      if (rc.reactants[0] < ns && !(h->species[rc.reactants[0]].flags & MCX_SP_VOL)) { h->err = "vol-wall class: reactant 0 must be a volume species"; return MCX_ERR_INVALID_ARG; }
    } else if (rc.reactants[0] >= ns || (rc.kind != MCX_RXN_UNIMOL && rc.reactants[1] >= ns)) {
      h->err = "reaction class references an unknown species"; return MCX_ERR_INVALID_ARG;
    }
    if (rc.n_pathways > 255) { h->err = "more than 255 pathways in one reaction class"; return MCX_ERR_INVALID_ARG; }
    dc[c] = DevClass{rc.max_fixed_p, rc.kind, rc.reactants[0], rc.reactants[1], rc.first_pathway, rc.n_pathways,
                     rc.reactant_orientation[0], rc.reactant_orientation[1], 0};
    auto is_vol = [&](uint32_t sp) { return (h->species[sp].flags & MCX_SP_VOL) != 0; };
    if (rc.kind == MCX_RXN_BIMOL_VOLVOL) {
      if (!is_vol(rc.reactants[0]) || !is_vol(rc.reactants[1])) { h->err = "vol-vol class with a surface reactant"; return MCX_ERR_INVALID_ARG; }
      bimol[rc.reactants[0] * ns + rc.reactants[1]] = (int)c;
      bimol[rc.reactants[1] * ns + rc.reactants[0]] = (int)c;
    } else if (rc.kind == MCX_RXN_BIMOL_VOLSURF) {
      if (!is_vol(rc.reactants[0]) || is_vol(rc.reactants[1])) { h->err = "vol-surf class: reactants must be (volume, surface)"; return MCX_ERR_INVALID_ARG; }
      volsurf[rc.reactants[0] * ns + rc.reactants[1]] = (int)c;
      any_surf = true;
    } else if (rc.kind == MCX_RXN_UNIMOL) unimol[rc.reactants[0]] = (int)c;
    else if (rc.kind == MCX_RXN_BIMOL_SURFSURF) {
      // two surface molecules (react_2D_all_neighbors): surface products go to the tiles the consumed reactants free
      if (is_vol(rc.reactants[0]) || is_vol(rc.reactants[1])) { h->err = "surf-surf class with a volume reactant"; return MCX_ERR_INVALID_ARG; }
      if (h->p.wall_border) { h->err = "surface-surface classes together with region borders are not supported (restricted regions of the neighbour search)"; return MCX_ERR_INVALID_ARG; }
      surfsurf[rc.reactants[0] * ns + rc.reactants[1]] = (int)c;
      surfsurf[rc.reactants[1] * ns + rc.reactants[0]] = (int)c;
      any_surf = any_surfsurf = true;
      for (uint32_t q = 0; q < rc.n_pathways; q++) {
        const mcx_pathway& pw = h->pathways[rc.first_pathway + q];
        const int keep0 = pw.keep_reactant_mask & 1u, keep1 = (pw.keep_reactant_mask >> 1) & 1u;
        int needed = 0;
        for (uint32_t k = 0; k < pw.n_products && k < MCX_MAX_PRODUCTS; k++) needed += (pw.products[k] < ns && !is_vol(pw.products[k])) ? 1 : 0;
        const int freed = 2 - keep0 - keep1, actual = (int)pw.n_products + keep0 + keep1;
        if (needed > freed) {
          if (const char* why = check_general(rc, q, freed, keep0 + keep1)) { h->err = std::string("surface-surface pathway: ") + why; return MCX_ERR_INVALID_ARG; }
          continue;
        }
        const int to_recycle = std::min(actual, freed);
        if (needed == 2 && to_recycle == 2 && actual > 2) {
          h->err = "a surface-surface pathway with two surface products on the two freed tiles and a volume product is not supported (the reference draws a vacant tile for the volume entry from an empty list: a division by zero)";
          return MCX_ERR_INVALID_ARG;
        }
        if (needed != 0 && !(needed == 1 && to_recycle == 1) && needed < to_recycle) {
          h->err = "a surface-surface pathway that frees more tiles than it has surface products next to a volume product is not supported (the reference's tile assignment does not terminate)";
          return MCX_ERR_INVALID_ARG;
        }
      }
      continue;
    }
    else if (rc.kind == MCX_RXN_BIMOL_VOLWALL) {
      any_surf = true;   // kept reactants and products remember the wall of their event: the cold surface arrays are needed
      for (uint32_t q = 0; q < rc.n_pathways; q++) {
        const mcx_pathway& pw = h->pathways[rc.first_pathway + q];
        if (pw.keep_reactant_mask & ~1u) { h->err = "vol-wall pathway: only reactant 0 can be kept (the surface always is)"; return MCX_ERR_INVALID_ARG; }
        for (uint32_t k = 0; k < pw.n_products && k < MCX_MAX_PRODUCTS; k++)
          if (pw.products[k] < ns && !is_vol(pw.products[k])) { h->err = "vol-wall pathway: surface products are not supported"; return MCX_ERR_INVALID_ARG; }
      }
      continue;
    }
    else { h->err = "unknown reaction kind"; return MCX_ERR_INVALID_ARG; }
    // supported product placement (SURVEY A.2 fast cases; the general find_surf_product_positions branch is not built)
    const bool surf_reactant = rc.kind == MCX_RXN_BIMOL_VOLSURF || (rc.kind == MCX_RXN_UNIMOL && !is_vol(rc.reactants[0]));
    for (uint32_t q = 0; q < rc.n_pathways; q++) {
      const mcx_pathway& pw = h->pathways[rc.first_pathway + q];
      uint32_t n_surf_products = 0;
      for (uint32_t k = 0; k < pw.n_products && k < MCX_MAX_PRODUCTS; k++)
        if (pw.products[k] < ns && !is_vol(pw.products[k])) n_surf_products++;
      if (!surf_reactant && n_surf_products) { h->err = "surface product without a surface reactant"; return MCX_ERR_INVALID_ARG; }
      if (surf_reactant) {
        const int surf_idx = rc.kind == MCX_RXN_BIMOL_VOLSURF ? 1 : 0;
        const bool surf_kept = (pw.keep_reactant_mask >> surf_idx) & 1u;
        if (n_surf_products > 1 || (n_surf_products == 1 && surf_kept)) {
          if (const char* why = check_general(rc, q, surf_kept ? 0 : 1, surf_kept ? 1 : 0)) { h->err = why; return MCX_ERR_INVALID_ARG; }
        }
        // a kept surface partner is not claimed by the event (several molecules may react with it in one iteration),
        // so it cannot change its orientation there: it has to carry the mark of the class on both sides
        if (rc.kind == MCX_RXN_BIMOL_VOLSURF && surf_kept && (pw.kept_info & MCX_KEPT_VALID)) {
          const uint32_t code = (pw.kept_info >> 26) & 3u;
          const int o = code == 1u ? 1 : (code == 2u ? -1 : 0);
          if (rc.reactant_orientation[1] == 0 ? o != 0 : o != rc.reactant_orientation[1]) {
            h->err = "a kept surface reactant of a volume-surface reaction that changes its orientation is not supported";
            return MCX_ERR_INVALID_ARG;
          }
        }
      }
    }
  }
  for (size_t k = 0; k < h->pathways.size(); k++) {
    const mcx_pathway& pw = h->pathways[k];
    if (pw.n_products > MCX_MAX_PRODUCTS) { h->err = "too many products"; return MCX_ERR_INVALID_ARG; }
    DevPathway d{};
    d.cum_prob = pw.cum_prob; d.n_products = pw.n_products; d.keep_mask = pw.keep_reactant_mask; d.rule_id = pw.rxn_rule_id;
    d.kept_info = pw.kept_info;
    d.general = general[k];
    for (uint32_t q = 0; q < pw.n_products; q++) {
      if (pw.products[q] >= ns) { h->err = "product references an unknown species"; return MCX_ERR_INVALID_ARG; }
      d.products[q] = pw.products[q];
      d.prod_orient[q] = pw.product_orientation[q];
    }
    dp[k] = d;
  }
  std::vector<DevSpecies> ds(ns);
  for (size_t a = 0; a < ns; a++) {
    bool any = false;
    for (size_t b = 0; b < ns; b++) any = any || bimol[a * ns + b] >= 0;
    bool vs = false;
    for (size_t b = 0; b < ns; b++) vs = vs || volsurf[a * ns + b] >= 0;
    bool ss = false;   // SPECIES_FLAG_CAN_SURFSURF
    for (size_t b = 0; b < ns; b++) ss = ss || surfsurf[a * ns + b] >= 0;
    ds[a] = DevSpecies{h->species[a].space_step, h->species[a].time_step, h->species[a].flags,
                       (any && !(h->species[a].flags & MCX_SP_CANT_INITIATE)) ? 1u : 0u, vs ? 1u : 0u, ss ? 1u : 0u};
  }
  // surface-class action table; lookup order as in find_mol_reactions_with_surf_classes
  // (rxn_utils.inl:182-244): species-specific, ALL_MOLECULES, ALL_VOLUME_MOLECULES
  uint32_t nsc = 0;
  for (const auto& r : h->surf_rules) nsc = std::max(nsc, r.surf_class + 1);
  for (uint32_t c : h->wall_class_host) if (c != MCX_NONE) nsc = std::max(nsc, c + 1);
  h->n_surf_classes = nsc;
  std::vector<uint8_t> act(std::max<size_t>(1, ns * nsc * 2), MCX_SURF_REFLECTIVE);
  std::vector<int> act_rxn(std::max<size_t>(1, ns * nsc * 2), -1);
  for (const auto& r : h->surf_rules)
    if (r.type == MCX_SURF_STANDARD && (r.rxn_class >= h->classes.size() || h->classes[r.rxn_class].kind != MCX_RXN_BIMOL_VOLWALL)) {
      h->err = "surface-class rule of type MCX_SURF_STANDARD needs a MCX_RXN_BIMOL_VOLWALL reaction class"; return MCX_ERR_INVALID_ARG;
    }
  bool absorbing = false;
  for (size_t a = 0; a < ns; a++)
    for (uint32_t c = 0; c < nsc; c++)
      for (int side = 0; side < 2; side++) {
        const int orient = side == 0 ? 1 : -1;
        const uint32_t order[3] = {(uint32_t)a, MCX_ALL_MOLECULES, MCX_ALL_VOLUME_MOLECULES};
        bool done = false;
        for (int o = 0; o < 3 && !done; o++)
          for (const auto& r : h->surf_rules)
            if (r.species == order[o] && r.surf_class == c && (r.orientation == 0 || r.orientation == orient)) {
              act[(a * nsc + c) * 2 + side] = (uint8_t)r.type;
              act_rxn[(a * nsc + c) * 2 + side] = r.type == MCX_SURF_STANDARD ? (int)r.rxn_class : -1;
              absorbing = absorbing || r.type == MCX_SURF_ABSORPTIVE;
              done = true;
              break;
            }
      }
  // region borders for surface molecules (reflect_absorb_check_wall, diffusion_utils.inl:598-628): the first REFLECTIVE or
  // ABSORPTIVE rule in the order species, ALL_MOLECULES, ALL_SURFACE_MOLECULES; nothing = it passes
  std::vector<uint8_t> border(std::max<size_t>(1, ns * nsc * 2), MCX_SURF_TRANSPARENT);
  for (size_t a = 0; a < ns; a++)
    for (uint32_t c = 0; c < nsc; c++)
      for (int side = 0; side < 2; side++) {
        const int orient = side == 0 ? 1 : -1;
        const uint32_t order[3] = {(uint32_t)a, MCX_ALL_MOLECULES, MCX_ALL_SURFACE_MOLECULES};
        bool done = false;
        for (int o = 0; o < 3 && !done; o++)
          for (const auto& r : h->surf_rules)
            if (r.species == order[o] && r.surf_class == c && (r.orientation == 0 || r.orientation == orient) &&
                (r.type == MCX_SURF_REFLECTIVE || r.type == MCX_SURF_ABSORPTIVE)) {
              border[(a * nsc + c) * 2 + side] = (uint8_t)r.type;
              done = true;
              break;
            }
      }
  // exact_disk ignores walls the moving molecule travels through (exact_disk_utils.inl:957-975): trigger_intersect
  // with ORIENTATION_NONE matches the orientation-independent classes (rxn_utils.inl:149-158); the wall is ignored
  // when there is at least one and all of them are transparent
  std::vector<uint8_t> exd_skip(std::max<size_t>(1, ns * nsc), 0);
  for (size_t a = 0; a < ns; a++)
    for (uint32_t c = 0; c < nsc; c++) {
      bool any = false, all_transparent = true;
      for (const auto& r : h->surf_rules) {
        if (r.surf_class != c || r.orientation != 0) continue;
        if (r.species != (uint32_t)a && r.species != MCX_ALL_MOLECULES && r.species != MCX_ALL_VOLUME_MOLECULES) continue;
        any = true;
        all_transparent = all_transparent && r.type == MCX_SURF_TRANSPARENT;
      }
      exd_skip[a * nsc + c] = (any && all_transparent) ? 1 : 0;
    }
  int rc = MCX_OK;
  rc |= dev_replace(h, &h->d_exd_skip, exd_skip.data(), exd_skip.size());
  rc |= dev_replace(h, &h->d_species, ds.data(), ds.size());
  rc |= dev_replace(h, &h->d_bimol, bimol.data(), bimol.size());
  rc |= dev_replace(h, &h->d_unimol, unimol.data(), unimol.size());
  rc |= dev_replace(h, &h->d_classes, dc.data(), dc.size());
  rc |= dev_replace(h, &h->d_pathways, dp.data(), dp.size());
  rc |= dev_replace(h, &h->d_surf, act.data(), act.size());
  rc |= dev_replace(h, &h->d_surf_rxn, act_rxn.data(), act_rxn.size());
  rc |= dev_replace(h, &h->d_surf_border, border.data(), border.size());
  rc |= dev_replace(h, &h->d_volsurf, volsurf.data(), volsurf.size());
  rc |= dev_replace(h, &h->d_surfsurf, surfsurf.data(), surfsurf.size());
  if (rc) return MCX_ERR_CUDA;
  DevParams& p = h->p;
  p.volsurf = (const int*)h->d_volsurf;
  p.surfsurf = any_surfsurf ? (const int*)h->d_surfsurf : nullptr;
  h->has_general = false;
  for (uint8_t g : general) h->has_general = h->has_general || g != 0;
  {  // the neighbour-tile table serves the partner search of surface-surface classes and the vacant tiles of general pathways
    const bool need_tiles = any_surfsurf || h->has_general;
    if (need_tiles != h->has_surfsurf || (need_tiles && !p.tn_start)) {
      h->has_surfsurf = need_tiles;
      const int trc = build_tile_neighbors(h);
      if (trc != MCX_OK) return trc;
    }
  }
  p.exd_skip = (const uint8_t*)h->d_exd_skip;
  h->has_surf = any_surf;

  p.species = (const DevSpecies*)h->d_species; p.bimol = (const int*)h->d_bimol; p.unimol = (const int*)h->d_unimol;
  p.classes = (const DevClass*)h->d_classes; p.pathways = (const DevPathway*)h->d_pathways;
  p.surf_rxn = (const int*)h->d_surf_rxn; p.surf_border = (const uint8_t*)h->d_surf_border;
  p.surf_action = (const uint8_t*)h->d_surf; p.n_species = (int)ns; p.n_surf_classes = (int)nsc;
  h->plan.has_claims = !h->classes.empty() || absorbing;
  h->plan.has_fresh = false;
  for (const mcx_rxn_class& rc : h->classes)
    for (uint32_t q = 0; q < rc.n_pathways; q++) {
      const mcx_pathway& pw = h->pathways[rc.first_pathway + q];
      const uint32_t n_react = (rc.kind == MCX_RXN_UNIMOL || rc.kind == MCX_RXN_BIMOL_VOLWALL) ? 1u : 2u;  // surf-surf: 2
      const uint32_t kept = (uint32_t)__builtin_popcount(pw.keep_reactant_mask & ((1u << n_react) - 1u));
      if (pw.n_products > n_react - kept) h->plan.has_fresh = true;
    }
  return MCX_OK;
}

int mcx_set_species(mcx_handle* h, const mcx_species* species, uint32_t n_species) {
  if (!h || !species || n_species == 0) { if (h) h->err = "no species"; return MCX_ERR_INVALID_ARG; }
  CK(cudaSetDevice(h->cfg.device));
  std::vector<mcx_species> previous = h->species;
  h->species.assign(species, species + n_species);
  const int rc = rebuild_tables(h);
  if (rc != MCX_OK) { h->species.swap(previous); return rc; }  // the device still holds the tables of `previous`
  h->has_species = true;
  return MCX_OK;
}

int mcx_set_reactions(mcx_handle* h, const mcx_rxn_class* classes, uint32_t n_classes, const mcx_pathway* pathways,
                      uint32_t n_pathways) {
  if (!h || (n_classes && (!classes || !pathways))) { if (h) h->err = "bad reaction arrays"; return MCX_ERR_INVALID_ARG; }
  if (!h->has_species) { h->err = "mcx_set_species must precede mcx_set_reactions"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  for (uint32_t k = 0; k < n_pathways; k++)
    if (pathways[k].rxn_rule_id >= MCX_MAX_COUNTED) { h->err = "rxn_rule_id >= 1024 not supported by the device counters"; return MCX_ERR_INVALID_ARG; }
  std::vector<mcx_rxn_class> prev_classes = h->classes;
  std::vector<mcx_pathway> prev_pathways = h->pathways;
  h->classes.assign(classes, classes + n_classes);
  h->pathways.assign(pathways, pathways + n_pathways);
  const int rc = rebuild_tables(h);
  if (rc != MCX_OK) { h->classes.swap(prev_classes); h->pathways.swap(prev_pathways); }
  return rc;
}

int mcx_set_surface_classes(mcx_handle* h, const mcx_surf_class_rxn* rules, uint32_t n_rules) {
  if (!h || (n_rules && !rules)) { if (h) h->err = "bad surface class arrays"; return MCX_ERR_INVALID_ARG; }
  if (!h->has_species) { h->err = "mcx_set_species must precede mcx_set_surface_classes"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  for (uint32_t k = 0; k < n_rules; k++)
    if (rules[k].surf_class >= 4096 || rules[k].type > MCX_SURF_STANDARD) { h->err = "bad surface class rule"; return MCX_ERR_INVALID_ARG; }
  std::vector<mcx_surf_class_rxn> previous = h->surf_rules;
  h->surf_rules.assign(rules, rules + n_rules);
  const int rc = rebuild_tables(h);
  if (rc != MCX_OK) h->surf_rules.swap(previous);
  return rc;
}

int mcx_set_counted_volumes(mcx_handle* h, uint32_t n_counted_volumes, const uint8_t* wall_cv_front, const uint8_t* wall_cv_back) {
  if (!h || !wall_cv_front || !wall_cv_back) { if (h) h->err = "null counted-volume arrays"; return MCX_ERR_INVALID_ARG; }
  if (!h->has_geometry) { h->err = "mcx_set_geometry must precede mcx_set_counted_volumes"; return MCX_ERR_STATE; }
  if (n_counted_volumes == 0 || n_counted_volumes > MCX_MAX_CV) { h->err = "1..256 counted volumes are supported"; return MCX_ERR_INVALID_ARG; }
  CK(cudaSetDevice(h->cfg.device));
  std::vector<uint16_t> cv(std::max<uint64_t>(h->n_walls_host, 1), 0);
  for (uint64_t i = 0; i < h->n_walls_host; i++) {
    if (wall_cv_front[i] >= n_counted_volumes || wall_cv_back[i] >= n_counted_volumes) { h->err = "counted volume index out of range"; return MCX_ERR_INVALID_ARG; }
    cv[i] = (uint16_t)(wall_cv_front[i] | (wall_cv_back[i] << 8));
  }
  std::vector<unsigned long long> zero_r((size_t)MCX_MAX_COUNTED * n_counted_volumes, 0), zero_m((size_t)MCX_MAX_COUNTED * n_counted_volumes, 0);
  int rc = MCX_OK;
  rc |= dev_replace(h, &h->d_wall_cv, cv.data(), cv.size());
  rc |= dev_replace(h, &h->d_rxn_count_cv, zero_r.data(), zero_r.size());
  rc |= dev_replace(h, &h->d_mol_count_cv, zero_m.data(), zero_m.size());
  if (rc) return MCX_ERR_CUDA;
  h->n_cv = n_counted_volumes;
  h->p.cv_mask = nullptr; h->p.cv_xor = 0; h->p.cv_all = 0;   // belongs to the previous table (mcx_set_counted_volume_objects again)
  h->p.wall_cv = (const uint16_t*)h->d_wall_cv; h->p.rxn_count_cv = (unsigned long long*)h->d_rxn_count_cv;
  h->p.mol_count_cv = (unsigned long long*)h->d_mol_count_cv; h->p.n_cv = n_counted_volumes;
  return MCX_OK;
}

int mcx_set_counted_volume_objects(mcx_handle* h, const uint32_t* cv_object_mask, uint32_t intersecting_objects) {
  if (!h) return MCX_ERR_INVALID_ARG;
  if (!h->p.wall_cv) { h->err = "mcx_set_counted_volumes must precede mcx_set_counted_volume_objects"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  if (!cv_object_mask) { h->p.cv_mask = nullptr; h->p.cv_xor = 0; h->p.cv_all = 0; return MCX_OK; }
  uint32_t all = 0;
  for (uint32_t k = 0; k < h->n_cv; k++) {
    all |= cv_object_mask[k];
    for (uint32_t q = 0; q < k; q++) if (cv_object_mask[q] == cv_object_mask[k]) { h->err = "two counted volumes with the same set of objects"; return MCX_ERR_INVALID_ARG; }
  }
  if (intersecting_objects & ~all) { h->err = "intersecting_objects names an object that encloses no counted volume"; return MCX_ERR_INVALID_ARG; }
  if (dev_replace(h, &h->d_cv_mask, cv_object_mask, h->n_cv)) return MCX_ERR_CUDA;
  h->p.cv_mask = (const uint32_t*)h->d_cv_mask; h->p.cv_xor = intersecting_objects; h->p.cv_all = all;
  return MCX_OK;
}

int mcx_counts_by_volume(mcx_handle* h, uint64_t* mol_counts, uint64_t* rxn_counts) {
  if (!h) return MCX_ERR_INVALID_ARG;
  if (!h->uploaded) { h->err = "nothing uploaded"; return MCX_ERR_STATE; }
  if (!h->p.wall_cv) { h->err = "mcx_set_counted_volumes was not called"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  h->p.cs_cur = h->cs[h->cs_cur]; h->p.cs_next = h->cs[h->cs_cur ^ 1]; h->p.iteration = h->iteration;
  if (mol_counts) {
    mcx_launch_count_by_volume(h->p, h->stream);
    h->launches += 1;
    CK(cudaMemcpyAsync(mol_counts, h->p.mol_count_cv, sizeof(uint64_t) * h->species.size() * h->n_cv, cudaMemcpyDeviceToHost, h->stream));
  }
  if (rxn_counts) {
    uint32_t n_rules = 0;
    for (const auto& pw : h->pathways) n_rules = std::max(n_rules, pw.rxn_rule_id + 1);
    CK(cudaMemcpyAsync(rxn_counts, h->p.rxn_count_cv, sizeof(uint64_t) * (size_t)n_rules * h->n_cv, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  if (h->comm) {  // every rank counted its own molecules and the events it owns: sum over the ranks
    auto reduce = [&](uint64_t* a, size_t n) -> int {
      for (size_t at = 0; at < n; at += 1024) {
        int rc = mcx_comm_allreduce_u64(h->comm, (unsigned long long*)(a + at), (int)std::min<size_t>(1024, n - at), h->stream);
        if (rc) { h->err = mcx_comm_error(h->comm); return rc; }
      }
      return MCX_OK;
    };
    if (mol_counts) { int rc = reduce(mol_counts, h->species.size() * h->n_cv); if (rc) return rc; }
    if (rxn_counts) {
      uint32_t n_rules = 0;
      for (const auto& pw : h->pathways) n_rules = std::max(n_rules, pw.rxn_rule_id + 1);
      int rc = reduce(rxn_counts, (size_t)n_rules * h->n_cv); if (rc) return rc;
    }
  }
  return MCX_OK;
}

int mcx_set_region_borders(mcx_handle* h, const uint8_t* wall_edge_border) {
  if (!h) return MCX_ERR_INVALID_ARG;
  if (!h->has_geometry) { h->err = "mcx_set_geometry must precede mcx_set_region_borders"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  if (!wall_edge_border) { h->p.wall_border = nullptr; return MCX_OK; }
  if (h->has_surfsurf) { h->err = "region borders together with surface-surface classes are not supported (restricted regions of the neighbour search)"; return MCX_ERR_INVALID_ARG; }
  if (dev_replace(h, &h->d_wall_border, wall_edge_border, std::max<uint64_t>(h->n_walls_host, 1))) return MCX_ERR_CUDA;
  h->p.wall_border = (const uint8_t*)h->d_wall_border;
  return MCX_OK;
}

int mcx_set_surface_regions(mcx_handle* h, uint32_t n_region_sets, const uint8_t* wall_region_set) {
  if (!h || !wall_region_set) { if (h) h->err = "null surface-region array"; return MCX_ERR_INVALID_ARG; }
  if (!h->has_geometry) { h->err = "mcx_set_geometry must precede mcx_set_surface_regions"; return MCX_ERR_STATE; }
  if (n_region_sets == 0 || n_region_sets > 256) { h->err = "1..256 surface-region sets are supported"; return MCX_ERR_INVALID_ARG; }
  CK(cudaSetDevice(h->cfg.device));
  for (uint64_t i = 0; i < h->n_walls_host; i++)
    if (wall_region_set[i] >= n_region_sets) { h->err = "surface-region set index out of range"; return MCX_ERR_INVALID_ARG; }
  std::vector<unsigned long long> zero((size_t)MCX_MAX_COUNTED * n_region_sets, 0);
  int rc = MCX_OK;
  rc |= dev_replace(h, &h->d_wall_rs, wall_region_set, std::max<uint64_t>(h->n_walls_host, 1));
  rc |= dev_replace(h, &h->d_rxn_count_rs, zero.data(), zero.size());
  rc |= dev_replace(h, &h->d_mol_count_rs, zero.data(), zero.size());
  if (rc) return MCX_ERR_CUDA;
  h->n_rs = n_region_sets;
  h->p.wall_rs = (const uint8_t*)h->d_wall_rs; h->p.rxn_count_rs = (unsigned long long*)h->d_rxn_count_rs;
  h->p.mol_count_rs = (unsigned long long*)h->d_mol_count_rs; h->p.n_rs = n_region_sets;
  return MCX_OK;
}

int mcx_counts_by_surface_region(mcx_handle* h, uint64_t* mol_counts, uint64_t* rxn_counts) {
  if (!h) return MCX_ERR_INVALID_ARG;
  if (!h->uploaded) { h->err = "nothing uploaded"; return MCX_ERR_STATE; }
  if (!h->p.wall_rs) { h->err = "mcx_set_surface_regions was not called"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  h->p.cs_cur = h->cs[h->cs_cur]; h->p.cs_next = h->cs[h->cs_cur ^ 1]; h->p.iteration = h->iteration;
  uint32_t n_rules = 0;
  for (const auto& pw : h->pathways) n_rules = std::max(n_rules, pw.rxn_rule_id + 1);
  if (mol_counts) {
    mcx_launch_count_by_surface_region(h->p, h->stream);
    h->launches += 1;
    CK(cudaMemcpyAsync(mol_counts, h->p.mol_count_rs, sizeof(uint64_t) * h->species.size() * h->n_rs, cudaMemcpyDeviceToHost, h->stream));
  }
  if (rxn_counts)
    CK(cudaMemcpyAsync(rxn_counts, h->p.rxn_count_rs, sizeof(uint64_t) * (size_t)n_rules * h->n_rs, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->comm) {  // every rank counted its own molecules and the events it owns: sum over the ranks
    auto reduce = [&](uint64_t* a, size_t n) -> int {
      for (size_t at = 0; at < n; at += 1024) {
        int rc = mcx_comm_allreduce_u64(h->comm, (unsigned long long*)(a + at), (int)std::min<size_t>(1024, n - at), h->stream);
        if (rc) { h->err = mcx_comm_error(h->comm); return rc; }
      }
      return MCX_OK;
    };
    if (mol_counts) { int rc = reduce(mol_counts, h->species.size() * h->n_rs); if (rc) return rc; }
    if (rxn_counts) { int rc = reduce(rxn_counts, (size_t)n_rules * h->n_rs); if (rc) return rc; }
  }
  return MCX_OK;
}

static void bind_iteration(mcx_handle* h) {
  h->p.cs_cur = h->cs[h->cs_cur];
  h->p.cs_next = h->cs[h->cs_cur ^ 1];
  h->p.iteration = h->iteration;
}

static int check_device_error(mcx_handle* h, Counters* host_ctr) {
  CK(cudaMemcpyAsync(host_ctr, h->p.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (host_ctr->error) {
    char buf[256];
    const char* what = host_ctr->error == MCX_ERR_ESCAPED ? "escaped the simulation area defined by partition size"
                     : host_ctr->error == MCX_ERR_CAPACITY ? "could not be stored: max_molecules exhausted"
                     : host_ctr->error == MCX_ERR_OVERFLOW ? "crossed more subpartitions than the device set holds"
                     : "raised a device error";
    snprintf(buf, sizeof(buf), "molecule (id: %u) %s", host_ctr->error_id, what);
    h->err = buf;
    return host_ctr->error;
  }
  return MCX_OK;
}

// staging SoA on the device (allocated once, sized by capacity): the ABI boundary is SoA, HBM is records
static int ensure_staging(mcx_handle* h) {
  if (h->st_x) return MCX_OK;
  const size_t cap = h->p.capacity;
  int rc = MCX_OK;
  rc |= dev_alloc(h, &h->st_x, cap); rc |= dev_alloc(h, &h->st_y, cap); rc |= dev_alloc(h, &h->st_z, cap);
  rc |= dev_alloc(h, &h->st_ts, cap); rc |= dev_alloc(h, &h->st_tu, cap);
  rc |= dev_alloc(h, &h->st_id, cap); rc |= dev_alloc(h, &h->st_sp, cap); rc |= dev_alloc(h, &h->st_fl, cap);
  rc |= dev_alloc(h, &h->st_cv, cap);
  return rc ? MCX_ERR_CUDA : MCX_OK;
}

// cold per-slot surface fields and their staging: only models with surface species pay for them
static int ensure_surface_arrays(mcx_handle* h) {
  h->p.has_surf = h->has_surf ? 1 : 0;
  if (h->has_surf && h->has_general && !h->p.prop_pmask) {  // placements of the pending proposals (place_general)
    int rcg = MCX_OK;
    rcg |= dev_alloc(h, &h->p.prop_ptile, (size_t)h->p.capacity * MCX_MAX_PRODUCTS);
    rcg |= dev_alloc(h, &h->p.prop_puv, (size_t)h->p.capacity * MCX_MAX_PRODUCTS);
    rcg |= dev_alloc(h, &h->p.prop_pmask, (size_t)h->p.capacity);
    if (rcg) return MCX_ERR_CUDA;
  }
  if (!h->has_surf || h->surf_allocated) return MCX_OK;
  const size_t cap = h->p.capacity;
  DevParams& p = h->p;
  int rc = MCX_OK;
  rc |= dev_alloc(h, &p.swallA, cap); rc |= dev_alloc(h, &p.swallB, cap);
  rc |= dev_alloc(h, &p.stileA, cap); rc |= dev_alloc(h, &p.stileB, cap);
  rc |= dev_alloc(h, &p.suvA, cap); rc |= dev_alloc(h, &p.suvB, cap);
  rc |= dev_alloc(h, &h->st_wall, cap); rc |= dev_alloc(h, &h->st_tile, cap); rc |= dev_alloc(h, &h->st_orient, cap);
  rc |= dev_alloc(h, &h->st_u, cap); rc |= dev_alloc(h, &h->st_v, cap);
  if (rc) return MCX_ERR_CUDA;
  h->surf_allocated = true;
  return MCX_OK;
}

int mcx_upload_molecules(mcx_handle* h, const mcx_mol_soa* m) {
  if (!h || !m) return MCX_ERR_INVALID_ARG;
  if (!h->has_species) { h->err = "species table missing"; return MCX_ERR_STATE; }
  if (h->has_surf && !h->has_geometry) { h->err = "surface species need mcx_set_geometry before the upload"; return MCX_ERR_STATE; }
  if (m->n > h->p.capacity) { h->err = "more molecules than max_molecules"; return MCX_ERR_CAPACITY; }
  if (m->n && (!m->x || !m->y || !m->z || !m->id || !m->species)) { h->err = "null molecule arrays"; return MCX_ERR_INVALID_ARG; }
  CK(cudaSetDevice(h->cfg.device));
  if (ensure_staging(h)) return MCX_ERR_CUDA;
  { int rcs = ensure_surface_arrays(h); if (rcs) return rcs; }
  const size_t n = m->n;
  cudaStream_t s = h->stream;
  SurfSoa sv{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (m->counted_volume && h->p.wall_cv) {
    CK(cudaMemcpyAsync(h->st_cv, m->counted_volume, n * 4, cudaMemcpyHostToDevice, s));
    sv.cv = h->st_cv;
  }
  if (h->has_surf && m->wall) {
    if (!m->tile || !m->orientation || !m->u || !m->v) { h->err = "incomplete surface molecule arrays"; return MCX_ERR_INVALID_ARG; }
    CK(cudaMemcpyAsync(h->st_wall, m->wall, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->st_tile, m->tile, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->st_orient, m->orientation, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->st_u, m->u, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->st_v, m->v, n * 8, cudaMemcpyHostToDevice, s));
    sv.wall = h->st_wall; sv.tile = h->st_tile; sv.orientation = h->st_orient; sv.u = h->st_u; sv.v = h->st_v;
  }
  CK(cudaMemcpyAsync(h->st_x, m->x, n * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->st_y, m->y, n * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->st_z, m->z, n * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->st_id, m->id, n * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->st_sp, m->species, n * 4, cudaMemcpyHostToDevice, s));
  if (m->flags) CK(cudaMemcpyAsync(h->st_fl, m->flags, n * 4, cudaMemcpyHostToDevice, s));
  if (m->diffusion_time) CK(cudaMemcpyAsync(h->st_ts, m->diffusion_time, n * 8, cudaMemcpyHostToDevice, s));
  if (m->unimol_rxn_time) CK(cudaMemcpyAsync(h->st_tu, m->unimol_rxn_time, n * 8, cudaMemcpyHostToDevice, s));
  // only the population fields start over: reaction counts, SimulationStats counters and next_id are cumulative over
  // the run (the reference's MolOrRxnCountEvent reports reactions since t = 0, and ids of dead molecules must not
  // be handed out again); k_pack_soa raises next_id above every uploaded id
  mcx_launch_reset_population(h->p, (unsigned int)n, s);
  // Wall::has_initialized_grid is NOT reset: the reference's walls keep their grids when molecules are added or taken away
  // on the host; the scatter marks the walls the uploaded population sits on (mcx_set_wall_grids restores a checkpoint)
  bind_iteration(h);
  mcx_launch_pack_soa(h->p, h->st_x, h->st_y, h->st_z, h->st_id, h->st_sp, m->flags ? h->st_fl : nullptr,
                      m->diffusion_time ? h->st_ts : nullptr, m->unimol_rxn_time ? h->st_tu : nullptr, sv, (unsigned int)n, s);
  h->launches += 1;
  mcx_launch_initial_sort(h->p, h->plan, s);
  h->cs_cur ^= 1;
  if (h->comm) {  // fetch the neighbours' boundary molecules before the first iteration
    bind_iteration(h);
    int rcc = mcx_comm_refresh(h->comm, h->p, h->plan, s);
    if (rcc) { h->err = mcx_comm_error(h->comm); return rcc; }
    h->cs_cur ^= 1;
  }
  Counters hc;
  int rc = check_device_error(h, &hc);
  if (rc) { if (rc == MCX_ERR_INVALID_ARG) h->err += " (unknown species)"; return rc; }
  CK(cudaGetLastError());
  h->uploaded = true;
  return MCX_OK;
}

// Release on the device: re-bin the snapshot (A -> B), append the new molecules behind it, sort.  DESIGN.md §3.
int mcx_release_volume_molecules(mcx_handle* h, const mcx_release* r, uint32_t* first_id_out) {
  if (!h || !r) return MCX_ERR_INVALID_ARG;
  if (!h->uploaded) { h->err = "mcx_release_volume_molecules needs a previous mcx_upload_molecules (it may be empty)"; return MCX_ERR_STATE; }
  if (h->p.rng_mode != MCX_RNG_PHILOX) { h->err = "device release needs rng_mode == MCX_RNG_PHILOX (a replay releases on the host)"; return MCX_ERR_STATE; }
  if (r->species >= h->species.size() || !(h->species[r->species].flags & MCX_SP_VOL)) { h->err = "release: not a volume species"; return MCX_ERR_INVALID_ARG; }
  if (r->shape > MCX_RELEASE_REGION) { h->err = "release: unknown shape"; return MCX_ERR_INVALID_ARG; }
  if (r->shape == MCX_RELEASE_REGION) {
    if (!h->has_geometry) { h->err = "region release needs mcx_set_geometry"; return MCX_ERR_STATE; }
    if (r->region_expr_len == 0 && (r->region_in == 0 || (r->region_in & r->region_out))) { h->err = "region release: region_in must name an object and be disjoint from region_out"; return MCX_ERR_INVALID_ARG; }
    {  // a well-formed postfix program: never pops an empty stack, leaves exactly one value
      bool ok = r->region_expr_len <= sizeof(r->region_expr);
      int depth = 0;
      for (uint32_t q = 0; ok && q < r->region_expr_len; q++) {
        const uint8_t op = r->region_expr[q];
        if (op < 32) ok = ++depth <= 24;
        else if (op == MCX_REGION_UNION || op == MCX_REGION_INTERSECT || op == MCX_REGION_DIFFERENCE) { ok = depth >= 2; depth--; }
        else ok = false;
      }
      if (!ok || (r->region_expr_len && depth != 1)) { h->err = "region release: malformed region expression"; return MCX_ERR_INVALID_ARG; }
    }
  }
  if (r->counted_volume_index >= h->n_cv) { h->err = "release: counted_volume_index out of range"; return MCX_ERR_INVALID_ARG; }
  const double it = (double)h->iteration;
  if (r->release_time != 0 && !(r->release_time >= it && r->release_time < it + 1.0)) { h->err = "release_time outside the current iteration"; return MCX_ERR_INVALID_ARG; }
  if (r->number > (uint64_t)h->p.capacity) { h->err = "release larger than max_molecules"; return MCX_ERR_CAPACITY; }
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->stream;
  Counters hc;
  int rc = check_device_error(h, &hc);
  if (rc) return rc;
  unsigned long long base = hc.next_id;
  if (h->comm) {  // the same first id on every rank: the maximum over ranks
    std::vector<unsigned long long> v((size_t)h->cfg.world_size, 0ull);
    v[(size_t)h->cfg.rank] = base;
    rc = mcx_comm_allreduce_u64(h->comm, v.data(), (int)v.size(), s);
    if (rc) { h->err = mcx_comm_error(h->comm); return rc; }
    for (unsigned long long x : v) base = std::max(base, x);  // next_id is global (k_assign_ids): equal on every rank
  }
  if (base + r->number >= 0xFFFFFFF0ull) { h->err = "molecule ids exhausted"; return MCX_ERR_OVERFLOW; }
  bind_iteration(h);
  mcx_launch_rebin(h->p, h->plan, s);
  mcx_launch_release(h->p, *r, (uint32_t)base, s);
  const unsigned int next_id = (unsigned int)(base + r->number);
  CK(cudaMemcpyAsync(&h->p.ctr->next_id, &next_id, sizeof(next_id), cudaMemcpyHostToDevice, s));
  mcx_launch_sort(h->p, h->plan, s);
  h->launches += 2;
  h->cs_cur ^= 1;
  if (h->comm) {  // hand the neighbours their halo copies of the new molecules
    bind_iteration(h);
    rc = mcx_comm_refresh(h->comm, h->p, h->plan, s);
    if (rc) { h->err = mcx_comm_error(h->comm); return rc; }
    h->cs_cur ^= 1;
  }
  rc = check_device_error(h, &hc);
  if (rc) return rc;
  CK(cudaGetLastError());
  if (first_id_out) *first_id_out = (uint32_t)base;
  return MCX_OK;
}

// ReleaseEvent::release_list for volume molecules: appended behind the re-binned snapshot like the other releases
int mcx_release_list(mcx_handle* h, uint64_t n, const uint32_t* species, const double* x, const double* y, const double* z,
                     const uint32_t* counted_volume, double release_time, uint32_t* first_id_out) {
  if (!h || (n && (!species || !x || !y || !z))) { if (h) h->err = "release list: null arrays"; return MCX_ERR_INVALID_ARG; }
  if (!h->uploaded) { h->err = "mcx_release_list needs a previous mcx_upload_molecules (it may be empty)"; return MCX_ERR_STATE; }
  const double it = (double)h->iteration;
  if (release_time != 0 && !(release_time >= it && release_time < it + 1.0)) { h->err = "release_time outside the current iteration"; return MCX_ERR_INVALID_ARG; }
  if (n > (uint64_t)h->p.capacity) { h->err = "release larger than max_molecules"; return MCX_ERR_CAPACITY; }
  CK(cudaSetDevice(h->cfg.device));
  if (ensure_staging(h)) return MCX_ERR_CUDA;
  cudaStream_t s = h->stream;
  Counters hc;
  int rc = check_device_error(h, &hc);
  if (rc) return rc;
  unsigned long long base = hc.next_id;
  if (h->comm) {
    std::vector<unsigned long long> v((size_t)h->cfg.world_size, 0ull);
    v[(size_t)h->cfg.rank] = base;
    rc = mcx_comm_allreduce_u64(h->comm, v.data(), (int)v.size(), s);
    if (rc) { h->err = mcx_comm_error(h->comm); return rc; }
    for (unsigned long long q : v) base = std::max(base, q);
  }
  if (base + n >= 0xFFFFFFF0ull) { h->err = "molecule ids exhausted"; return MCX_ERR_OVERFLOW; }
  CK(cudaMemcpyAsync(h->st_x, x, n * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->st_y, y, n * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->st_z, z, n * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->st_sp, species, n * 4, cudaMemcpyHostToDevice, s));
  const bool with_cv = counted_volume && h->p.wall_cv;
  if (with_cv) CK(cudaMemcpyAsync(h->st_cv, counted_volume, n * 4, cudaMemcpyHostToDevice, s));
  bind_iteration(h);
  mcx_launch_rebin(h->p, h->plan, s);
  mcx_launch_release_list(h->p, h->st_x, h->st_y, h->st_z, h->st_sp, with_cv ? h->st_cv : nullptr, n, release_time, (uint32_t)base, s);
  const unsigned int next_id = (unsigned int)(base + n);
  CK(cudaMemcpyAsync(&h->p.ctr->next_id, &next_id, sizeof(next_id), cudaMemcpyHostToDevice, s));
  mcx_launch_sort(h->p, h->plan, s);
  h->launches += 2;
  h->cs_cur ^= 1;
  if (h->comm) {
    bind_iteration(h);
    rc = mcx_comm_refresh(h->comm, h->p, h->plan, s);
    if (rc) { h->err = mcx_comm_error(h->comm); return rc; }
    h->cs_cur ^= 1;
  }
  rc = check_device_error(h, &hc);
  if (rc) { if (rc == MCX_ERR_INVALID_ARG) h->err += " (not a volume species, or counted volume out of range)"; return rc; }
  CK(cudaGetLastError());
  if (first_id_out) *first_id_out = (uint32_t)base;
  return MCX_OK;
}

// ReleaseEvent::release_onto_regions on the device (include/mcx.h): rounds of pick / bid / settle, then the fall-back fill
int mcx_release_surface_molecules(mcx_handle* h, const mcx_surface_release* r, uint32_t* first_id_out) {
  if (!h || !r) return MCX_ERR_INVALID_ARG;
  if (!h->uploaded) { h->err = "mcx_release_surface_molecules needs a previous mcx_upload_molecules (it may be empty)"; return MCX_ERR_STATE; }
  if (h->p.rng_mode != MCX_RNG_PHILOX) { h->err = "device release needs rng_mode == MCX_RNG_PHILOX"; return MCX_ERR_STATE; }
  if (h->comm || h->cfg.world_size > 1) { h->err = "surface release on the device needs one device (a rank knows the tiles of its own slab only)"; return MCX_ERR_STATE; }
  if (r->species >= h->species.size() || (h->species[r->species].flags & MCX_SP_VOL)) { h->err = "surface release: not a surface species"; return MCX_ERR_INVALID_ARG; }
  if (!h->has_geometry || !h->p.has_surf || !h->p.n_tiles) { h->err = "surface release needs geometry and an upload after the surface species were set"; return MCX_ERR_STATE; }
  if (!r->walls || r->n_walls == 0 || r->n_walls > 0xFFFFFFF0ull) { h->err = "surface release: empty wall list"; return MCX_ERR_INVALID_ARG; }
  if (r->orientation < -1 || r->orientation > 1) { h->err = "surface release: orientation must be -1, 0 or +1"; return MCX_ERR_INVALID_ARG; }
  const double it = (double)h->iteration;
  if (r->release_time != 0 && !(r->release_time >= it && r->release_time < it + 1.0)) { h->err = "release_time outside the current iteration"; return MCX_ERR_INVALID_ARG; }
  if (r->number > (uint64_t)h->p.capacity) { h->err = "release larger than max_molecules"; return MCX_ERR_CAPACITY; }
  std::vector<double> cum(r->n_walls), area(r->n_walls);
  double total = 0;
  {
    std::vector<uint8_t> seen(h->n_walls_host, 0);
    for (uint64_t a = 0; a < r->n_walls; a++) {
      const uint32_t wi = r->walls[a];
      if (wi >= h->n_walls_host || seen[wi]) { h->err = "surface release: wall index out of range or listed twice"; return MCX_ERR_INVALID_ARG; }
      seen[wi] = 1;
      area[a] = h->wall_area_host[wi];
      total += area[a];
      cum[a] = total;   // cumm_area_and_pwall_index_pairs
    }
  }
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->stream;
  Counters hc;
  int rc = check_device_error(h, &hc);
  if (rc) return rc;
  const unsigned long long base = hc.next_id;
  if (base + r->number >= 0xFFFFFFF0ull) { h->err = "molecule ids exhausted"; return MCX_ERR_OVERFLOW; }
  const size_t n = (size_t)r->number;
  uint32_t *d_walls = nullptr, *d_claim = nullptr, *d_choice = nullptr, *d_cwall = nullptr, *d_pa = nullptr, *d_pb = nullptr;
  double *d_cum = nullptr, *d_area = nullptr;
  unsigned int* d_ctr = nullptr;
  auto release_tmp = [&]() { cudaFree(d_walls); cudaFree(d_claim); cudaFree(d_choice); cudaFree(d_cwall); cudaFree(d_pa); cudaFree(d_pb);
                             cudaFree(d_cum); cudaFree(d_area); cudaFree(d_ctr); };
  bool ok = cudaMalloc(&d_walls, r->n_walls * 4) == cudaSuccess && cudaMalloc(&d_cum, r->n_walls * 8) == cudaSuccess &&
            cudaMalloc(&d_area, r->n_walls * 8) == cudaSuccess && cudaMalloc(&d_claim, (size_t)h->p.n_tiles * 4) == cudaSuccess &&
            cudaMalloc(&d_choice, std::max<size_t>(n, 1) * 4) == cudaSuccess && cudaMalloc(&d_cwall, std::max<size_t>(n, 1) * 4) == cudaSuccess &&
            cudaMalloc(&d_pa, std::max<size_t>(n, 1) * 4) == cudaSuccess && cudaMalloc(&d_pb, std::max<size_t>(n, 1) * 4) == cudaSuccess &&
            cudaMalloc(&d_ctr, 16) == cudaSuccess;
  if (!ok) { release_tmp(); h->err = "surface release: out of device memory"; return MCX_ERR_CUDA; }
  cudaMemcpyAsync(d_walls, r->walls, r->n_walls * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_cum, cum.data(), r->n_walls * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_area, area.data(), r->n_walls * 8, cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(d_claim, 0xFF, (size_t)h->p.n_tiles * 4, s);
  SurfRelease sr{};
  sr.walls = d_walls; sr.cum_area = d_cum; sr.area = d_area; sr.n_walls = (unsigned int)r->n_walls; sr.total_area = total;
  sr.species = r->species; sr.orientation = r->orientation; sr.randomize_pos = r->randomize_pos; sr.release_time = r->release_time;
  sr.first_id = (uint32_t)base; sr.claim = d_claim; sr.choice = d_choice; sr.choice_wall = d_cwall;
  bind_iteration(h);
  unsigned int host_n = 0;
  mcx_launch_surface_release_count_vacant(h->p, sr, d_ctr, s);
  cudaMemcpyAsync(&host_n, d_ctr, 4, cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) { release_tmp(); h->err = cudaGetErrorString(cudaGetLastError()); return MCX_ERR_CUDA; }
  if ((uint64_t)host_n < r->number) {
    release_tmp();
    h->err = "surface release: " + std::to_string(r->number) + " molecules for " + std::to_string(host_n) + " vacant tiles";
    return MCX_ERR_CAPACITY;
  }
  mcx_launch_rebin(h->p, h->plan, s);
  unsigned int n_pend = (unsigned int)n;
  const uint32_t* pend_in = nullptr;
  uint32_t* bufs[2] = {d_pa, d_pb};
  int which = 0;
  for (unsigned int round = 0; round < MCX_SURFACE_RELEASE_ROUNDS && n_pend > 0; round++) {
    mcx_launch_surface_release_round(h->p, sr, pend_in, n_pend, bufs[which], d_ctr, round, s);
    h->launches += 3;
    cudaMemcpyAsync(&host_n, d_ctr, 4, cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) { release_tmp(); h->err = cudaGetErrorString(cudaGetLastError()); return MCX_ERR_CUDA; }
    n_pend = host_n;
    pend_in = bufs[which];
    which ^= 1;
  }
  if (n_pend > 0) {  // the reference's fall-back: first vacant tiles in list order, lowest id first
    std::vector<uint32_t> left(n_pend);
    cudaMemcpyAsync(left.data(), pend_in, (size_t)n_pend * 4, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    std::sort(left.begin(), left.end());
    cudaMemcpyAsync(bufs[which], left.data(), (size_t)n_pend * 4, cudaMemcpyHostToDevice, s);
    mcx_launch_surface_release_fill(h->p, sr, bufs[which], n_pend, d_ctr, s);
    h->launches += 1;
    cudaMemcpyAsync(&host_n, d_ctr, 4, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (host_n != 0) { release_tmp(); h->err = "surface release: ran out of vacant tiles"; return MCX_ERR_CAPACITY; }
  }
  const unsigned int next_id = (unsigned int)(base + r->number);
  cudaMemcpyAsync(&h->p.ctr->next_id, &next_id, sizeof(next_id), cudaMemcpyHostToDevice, s);
  mcx_launch_sort(h->p, h->plan, s);
  h->launches += 2;
  h->cs_cur ^= 1;
  rc = check_device_error(h, &hc);
  release_tmp();
  if (rc) return rc;
  CK(cudaGetLastError());
  if (first_id_out) *first_id_out = (uint32_t)base;
  return MCX_OK;
}

int mcx_get_next_molecule_id(mcx_handle* h, uint32_t* next_id_out) {
  if (!h || !next_id_out) return MCX_ERR_INVALID_ARG;
  if (!h->uploaded) { h->err = "nothing uploaded"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  Counters hc;
  int rc = check_device_error(h, &hc);
  if (rc) return rc;
  *next_id_out = hc.next_id;
  return MCX_OK;
}
int mcx_set_next_molecule_id(mcx_handle* h, uint32_t next_id) {
  if (!h) return MCX_ERR_INVALID_ARG;
  if (!h->uploaded) { h->err = "mcx_set_next_molecule_id follows mcx_upload_molecules"; return MCX_ERR_STATE; }
  if (next_id >= 0xFFFFFFF0u) { h->err = "molecule ids exhausted"; return MCX_ERR_OVERFLOW; }
  CK(cudaSetDevice(h->cfg.device));
  Counters hc;
  int rc = check_device_error(h, &hc);
  if (rc) return rc;
  if (next_id > hc.next_id) {
    CK(cudaMemcpyAsync(&h->p.ctr->next_id, &next_id, sizeof(next_id), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return MCX_OK;
}

int mcx_get_wall_grids(mcx_handle* h, uint8_t* has_grid_out, uint64_t n_walls) {
  if (!h || !has_grid_out) { if (h) h->err = "null wall-grid array"; return MCX_ERR_INVALID_ARG; }
  if (!h->has_geometry || n_walls != h->n_walls_host) { h->err = "mcx_get_wall_grids: wall count differs from the geometry"; return MCX_ERR_INVALID_ARG; }
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  if (n_walls) CK(cudaMemcpy(has_grid_out, h->p.wall_has_grid, n_walls, cudaMemcpyDeviceToHost));
  return MCX_OK;
}

int mcx_set_wall_grids(mcx_handle* h, const uint8_t* has_grid, uint64_t n_walls) {
  if (!h || !has_grid) { if (h) h->err = "null wall-grid array"; return MCX_ERR_INVALID_ARG; }
  if (!h->has_geometry || n_walls != h->n_walls_host) { h->err = "mcx_set_wall_grids: wall count differs from the geometry"; return MCX_ERR_INVALID_ARG; }
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  // grids are never taken away: the saved flags are OR-ed into what the uploaded population has set already
  std::vector<uint8_t> cur(std::max<uint64_t>(n_walls, 1), 0);
  if (n_walls) CK(cudaMemcpy(cur.data(), h->p.wall_has_grid, n_walls, cudaMemcpyDeviceToHost));
  for (uint64_t i = 0; i < n_walls; i++) cur[i] = (cur[i] || has_grid[i]) ? 1 : 0;
  if (n_walls) CK(cudaMemcpy(h->p.wall_has_grid, cur.data(), n_walls, cudaMemcpyHostToDevice));
  return MCX_OK;
}

uint64_t mcx_num_molecules(mcx_handle* h) {
  if (!h || !h->uploaded) return 0;
  cudaSetDevice(h->cfg.device);
  Counters hc;
  if (cudaMemcpy(&hc, h->p.ctr, sizeof(hc), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  uint64_t n = 0;
  for (size_t i = 0; i < h->species.size(); i++) n += hc.species_count[i];
  return n;
}

int mcx_download_molecules(mcx_handle* h, mcx_mol_soa* out, uint64_t capacity) {
  if (!h || !out) return MCX_ERR_INVALID_ARG;
  if (!h->uploaded) { h->err = "nothing uploaded"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  if (ensure_staging(h)) return MCX_ERR_CUDA;
  cudaStream_t s = h->stream;
  bind_iteration(h);
  SurfSoaOut sv{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const bool want_surf = h->surf_allocated && out->wall && out->tile && out->orientation && out->u && out->v;
  if (want_surf) { sv.wall = h->st_wall; sv.tile = h->st_tile; sv.orientation = h->st_orient; sv.u = h->st_u; sv.v = h->st_v; }
  if (out->counted_volume) sv.cv = h->st_cv;
  mcx_launch_unpack_soa(h->p, h->st_x, h->st_y, h->st_z, h->st_id, h->st_sp, h->st_fl, h->st_ts, h->st_tu, sv, h->d_n_out, s);
  h->launches += 1;
  unsigned int live = 0;
  CK(cudaMemcpyAsync(&live, h->d_n_out, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (live > capacity) { h->err = "download capacity too small"; return MCX_ERR_CAPACITY; }
  CK(cudaMemcpyAsync(out->x, h->st_x, live * 8ull, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(out->y, h->st_y, live * 8ull, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(out->z, h->st_z, live * 8ull, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(out->id, h->st_id, live * 4ull, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(out->species, h->st_sp, live * 4ull, cudaMemcpyDeviceToHost, s));
  if (out->flags) CK(cudaMemcpyAsync(out->flags, h->st_fl, live * 4ull, cudaMemcpyDeviceToHost, s));
  if (out->diffusion_time) CK(cudaMemcpyAsync(out->diffusion_time, h->st_ts, live * 8ull, cudaMemcpyDeviceToHost, s));
  if (out->unimol_rxn_time) CK(cudaMemcpyAsync(out->unimol_rxn_time, h->st_tu, live * 8ull, cudaMemcpyDeviceToHost, s));
  if (out->counted_volume) CK(cudaMemcpyAsync(out->counted_volume, h->st_cv, live * 4ull, cudaMemcpyDeviceToHost, s));
  if (want_surf) {
    CK(cudaMemcpyAsync(out->wall, h->st_wall, live * 4ull, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->tile, h->st_tile, live * 4ull, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->orientation, h->st_orient, live * 4ull, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->u, h->st_u, live * 8ull, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->v, h->st_v, live * 8ull, cudaMemcpyDeviceToHost, s));
  } else if (out->wall) {
    for (uint64_t i = 0; i < live && i < capacity; i++) out->wall[i] = MCX_NONE;  // no surface species: all volume
  }
  CK(cudaStreamSynchronize(s));
  out->n = live;
  return MCX_OK;
}

static void fill_stats(const Counters& a, const Counters& b, uint64_t iters, float ms, size_t ns, mcx_step_stats* s) {
  if (!s) return;
  memset(s, 0, sizeof(*s));
  s->iterations = iters;
  s->molecule_steps = b.molecule_steps - a.molecule_steps;
  for (size_t i = 0; i < ns; i++) s->n_live += b.species_count[i];
  s->ray_polygon_tests = b.ray_polygon_tests - a.ray_polygon_tests;
  s->ray_polygon_colls = b.ray_polygon_colls - a.ray_polygon_colls;
  s->mol_wall_reflections = b.reflections - a.reflections;
  s->mol_wall_transparent = b.transparent - a.transparent;
  s->mol_wall_absorptions = b.absorptions - a.absorptions;
  s->vol_mol_vol_mol_collisions = b.volvol_collisions - a.volvol_collisions;
  s->bimol_rxns = b.bimol_rxns - a.bimol_rxns;
  s->unimol_rxns = b.unimol_rxns - a.unimol_rxns;
  s->wall_redos = b.redos - a.redos;
  s->resolve_retries = b.retries - a.retries;
  s->unresolved_conflicts = b.unresolved - a.unresolved;
  s->products_created = b.products - a.products;
  s->deferred_molecules = b.deferred - a.deferred;
  for (int k = 0; k < 8; k++) s->deferred_by_reason[k] = b.defer_reason[k] - a.defer_reason[k];
  s->device_ms = ms;
}

static int run_iterations(mcx_handle* h, uint32_t n_iterations, mcx_step_stats* stats_out) {
  if (!h->uploaded) { h->err = "mcx_upload_molecules must precede stepping"; return MCX_ERR_STATE; }
  if (h->cfg.world_size > 1 && !h->comm) { h->err = "world_size > 1 needs mcx_comm_init before stepping"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  const unsigned long long launches_before = h->launches;
  Counters before;
  CK(cudaMemcpyAsync(&before, h->p.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  mcx_plan_tiles(h->p, before.n_slots);
  CK(cudaEventRecord(h->ev0, h->stream));
  const uint32_t n_prof = h->profiling ? std::min<uint32_t>(n_iterations, 256) : 0;
  if (h->prof_events.size() < 5ull * n_prof) {
    size_t old = h->prof_events.size();
    h->prof_events.resize(5ull * n_prof);
    for (size_t q = old; q < h->prof_events.size(); q++) CK(cudaEventCreate(&h->prof_events[q]));
  }
  for (uint32_t k = 0; k < n_iterations; k++) {
    bind_iteration(h);
    h->plan.prof = k < n_prof ? &h->prof_events[5ull * k] : nullptr;
    if (h->comm) {
      int rc = mcx_comm_iteration(h->comm, h->p, h->plan, h->stream);
      if (rc) { h->err = mcx_comm_error(h->comm); return rc; }
    } else {
      mcx_launch_iteration(h->p, h->plan, h->stream);
    }
    h->cs_cur ^= 1;
    h->iteration++;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  Counters after;
  int rc = check_device_error(h, &after);
  CK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  fill_stats(before, after, n_iterations, ms, h->species.size(), stats_out);
  h->plan.prof = nullptr;
  if (stats_out) {
    stats_out->kernel_launches = h->launches - launches_before;
    for (uint32_t k = 0; k < n_prof && rc == MCX_OK; k++) {
      float a = 0, a2 = 0, b = 0, c = 0;
      cudaEvent_t* e = &h->prof_events[5ull * k];
      cudaEventElapsedTime(&a, e[0], e[4]); cudaEventElapsedTime(&a2, e[4], e[1]);
      cudaEventElapsedTime(&b, e[1], e[2]); cudaEventElapsedTime(&c, e[2], e[3]);
      stats_out->ms_diffuse += a; stats_out->ms_diffuse_slow += a2; stats_out->ms_resolve += b; stats_out->ms_sort += c;
    }
    stats_out->profiled_iterations = n_prof;
  }
  return rc;
}

int mcx_step(mcx_handle* h, uint32_t n_iterations, mcx_step_stats* stats_out) {
  if (!h) return MCX_ERR_INVALID_ARG;
  if (h->p.rng_mode != MCX_RNG_PHILOX) { h->err = "mcx_step needs rng_mode == MCX_RNG_PHILOX (use mcx_replay_step)"; return MCX_ERR_STATE; }
  h->p.trace = nullptr; h->p.n_trace = 0; h->plan.trace = false;
  return run_iterations(h, n_iterations, stats_out);
}

static int traced_iteration(mcx_handle* h, uint64_t n_ids, mcx_trace_rec* trace_out, mcx_step_stats* stats_out) {
  mcx_trace_rec* d_trace = nullptr;
  CK(cudaMalloc((void**)&d_trace, std::max<uint64_t>(n_ids, 1) * sizeof(mcx_trace_rec)));
  CK(cudaMemset(d_trace, 0, std::max<uint64_t>(n_ids, 1) * sizeof(mcx_trace_rec)));
  h->p.trace = d_trace; h->p.n_trace = n_ids; h->plan.trace = true;
  int rc = run_iterations(h, 1, stats_out);
  h->p.trace = nullptr; h->p.n_trace = 0; h->plan.trace = false;
  if (rc == MCX_OK && trace_out) {
    cudaError_t e = cudaMemcpy(trace_out, d_trace, n_ids * sizeof(mcx_trace_rec), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { h->err = cudaGetErrorString(e); rc = MCX_ERR_CUDA; }
  }
  cudaFree(d_trace);
  return rc;
}

int mcx_replay_step(mcx_handle* h, const uint32_t* words, uint64_t n_words, const uint64_t* offset_by_id, uint64_t n_ids,
                    mcx_trace_rec* trace_out, mcx_step_stats* stats_out) {
  if (!h || !words || !offset_by_id) return MCX_ERR_INVALID_ARG;
  if (h->p.rng_mode != MCX_RNG_TAPE) { h->err = "mcx_replay_step needs rng_mode == MCX_RNG_TAPE"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  uint32_t* d_words = nullptr; unsigned long long* d_off = nullptr;
  CK(cudaMalloc((void**)&d_words, std::max<uint64_t>(n_words, 1) * 4));
  CK(cudaMalloc((void**)&d_off, std::max<uint64_t>(n_ids, 1) * 8));
  CK(cudaMemcpy(d_words, words, n_words * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_off, offset_by_id, n_ids * 8, cudaMemcpyHostToDevice));
  h->p.tape = d_words; h->p.n_words = n_words; h->p.tape_off = d_off; h->p.n_ids = n_ids;
  int rc = traced_iteration(h, n_ids, trace_out, stats_out);
  h->p.tape = nullptr; h->p.tape_off = nullptr; h->p.n_words = 0; h->p.n_ids = 0;
  cudaFree(d_words); cudaFree(d_off);
  return rc;
}

int mcx_trace_step(mcx_handle* h, uint64_t n_ids, mcx_trace_rec* trace_out, mcx_step_stats* stats_out) {
  if (!h) return MCX_ERR_INVALID_ARG;
  if (h->p.rng_mode != MCX_RNG_PHILOX) { h->err = "mcx_trace_step needs rng_mode == MCX_RNG_PHILOX"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  return traced_iteration(h, n_ids, trace_out, stats_out);
}

int mcx_counts(mcx_handle* h, uint64_t* per_species, uint32_t n_species, uint64_t* per_rxn_rule, uint32_t n_rxn_rules) {
  if (!h) return MCX_ERR_INVALID_ARG;
  if (!h->uploaded) { h->err = "nothing uploaded"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  Counters hc;
  CK(cudaMemcpy(&hc, h->p.ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  std::vector<unsigned long long> buf(2 * MCX_MAX_COUNTED, 0);
  for (uint32_t i = 0; i < MCX_MAX_COUNTED; i++) { buf[i] = hc.species_count[i]; buf[MCX_MAX_COUNTED + i] = hc.rxn_count[i]; }
  if (h->comm) {
    for (size_t at = 0; at < buf.size(); at += 1024) {
      int rc = mcx_comm_allreduce_u64(h->comm, buf.data() + at, 1024, h->stream);
      if (rc) { h->err = mcx_comm_error(h->comm); return rc; }
    }
  }
  for (uint32_t i = 0; i < n_species && per_species; i++) per_species[i] = i < MCX_MAX_COUNTED ? buf[i] : 0;
  for (uint32_t i = 0; i < n_rxn_rules && per_rxn_rule; i++) per_rxn_rule[i] = i < MCX_MAX_COUNTED ? buf[MCX_MAX_COUNTED + i] : 0;
  return MCX_OK;
}

// Slab layout of this rank: owned z-layers of the GLOBAL cell grid plus halo layers towards each neighbour.
// halo width: cfg.halo_width, or 3x the hard reach of one step (R + 6.993 * max space_step: |gauss| < 9.89,
// src/rng.c:198,212) — one reach makes the first evaluation of every owned molecule exact, the rest covers the
// conflict chains that cross the face.
static int configure_slab(mcx_handle* h) {
  DevParams& p = h->p;
  const int world = h->cfg.world_size, rank = h->cfg.rank;
  double max_step = 0;
  for (const auto& sp : h->species) max_step = std::max(max_step, sp.space_step);
  const double reach = p.R * (1.0 + 1e-9) + 1e-9 + 6.993 * max_step;
  const double width = h->cfg.halo_width > 0 ? h->cfg.halo_width : 3.0 * reach;
  if (width < reach) { h->err = "halo_width smaller than the reach of one step (R + 6.993 * max space_step)"; return MCX_ERR_INVALID_ARG; }
  const double ez = 1.0 / p.cell_rcp_z;
  const int H = (int)std::ceil(width / ez);
  const int g = h->ncz_global;
  // Balanced slabs: every rank EVALUATES its owned layers plus a halo towards each neighbour, and the two outermost
  // ranks have one neighbour only — they own H layers more, so that all ranks evaluate (g - 2H) / world + 2H layers
  // (equal slabs left the inner ranks with 7 % more work at N = 4 and 8, and everybody waits for them in the halo
  // refresh).  Mirrored by comm.layer_range() on the host.
  auto bound = [&](int k) -> int { return k <= 0 ? 0 : (k >= world ? g : H + (int)((long long)(g - 2 * H) * k / world)); };
  if (g <= 2 * H) { h->err = "slab thinner than the halo: fewer ranks or a narrower halo_width"; return MCX_ERR_INVALID_ARG; }
  const int g_lo = bound(rank), g_hi = bound(rank + 1);
  if (g_hi - g_lo < H) { h->err = "slab thinner than the halo: fewer ranks or a narrower halo_width"; return MCX_ERR_INVALID_ARG; }
  const int z_off = std::max(0, g_lo - H), z_end = std::min(g, g_hi + H);
  p.z_off = z_off; p.ncz = z_end - z_off;
  p.own_lo = g_lo - z_off; p.own_hi = g_hi - z_off;
  p.world = world; p.halo_layers = H;
  p.has_low = rank > 0; p.has_high = rank < world - 1;
  set_cell_count(p);
  return MCX_OK;
}

int mcx_comm_init(mcx_handle* h, const void* nccl_unique_id, uint32_t id_bytes) {
  if (!h || !nccl_unique_id) return MCX_ERR_INVALID_ARG;
  if (!h->has_species) { h->err = "mcx_set_species must precede mcx_comm_init (the halo width depends on the step lengths)"; return MCX_ERR_STATE; }
  if (h->uploaded) { h->err = "mcx_comm_init must precede mcx_upload_molecules"; return MCX_ERR_STATE; }
  CK(cudaSetDevice(h->cfg.device));
  int rc = configure_slab(h);
  if (rc) return rc;
  std::string err;
  const unsigned int halo_cap = std::max<unsigned int>(1u << 16, h->p.capacity / 3);
  h->comm = mcx_comm_create(nccl_unique_id, id_bytes, h->cfg.rank, h->cfg.world_size, halo_cap, err);
  if (!h->comm) { h->err = err; return MCX_ERR_COMM; }
  h->p.rank_fresh = mcx_comm_rank_fresh(h->comm);
  return MCX_OK;
}

int mcx_slab_info_get(mcx_handle* h, mcx_slab_info* out) {
  if (!h || !out) return MCX_ERR_INVALID_ARG;
  const DevParams& p = h->p;
  out->grid_origin_z = p.cgz; out->layer_rcp = p.cell_rcp_z; out->n_layers = (uint32_t)h->ncz_global;
  out->layer_lo = (uint32_t)(p.own_lo + p.z_off); out->layer_hi = (uint32_t)(p.own_hi + p.z_off);
  out->halo_layers = (uint32_t)p.halo_layers; out->rank = h->cfg.rank; out->world_size = p.world;
  return MCX_OK;
}

int mcx_comm_halo_path(mcx_handle* h) {
  if (!h) return MCX_ERR_INVALID_ARG;
  return h->comm ? (mcx_comm_is_p2p(h->comm) ? 2 : 1) : 0;
}

int mcx_fast_pass_kind(mcx_handle* h) { return h ? h->p.tile.enabled : MCX_ERR_INVALID_ARG; }

int mcx_set_profiling(mcx_handle* h, int enabled) {
  if (!h) return MCX_ERR_INVALID_ARG;
  h->profiling = enabled != 0;
  return MCX_OK;
}

// sizeof table for the binding-side layout check (tests/test_abi.py)
int mcx_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(mcx_config);
    case 1: return (int)sizeof(mcx_species);
    case 2: return (int)sizeof(mcx_rxn_class);
    case 3: return (int)sizeof(mcx_pathway);
    case 4: return (int)sizeof(mcx_surf_class_rxn);
    case 5: return (int)sizeof(mcx_mol_soa);
    case 6: return (int)sizeof(mcx_step_stats);
    case 7: return (int)sizeof(mcx_trace_rec);
    case 8: return (int)sizeof(mcx_slab_info);
    case 9: return (int)sizeof(mcx_release);
    case 10: return (int)sizeof(mcx_surface_release);
    default: return -1;
  }
}

}  // extern "C"
