// mcx_internal.h — device/host shared layouts of libmcx (not part of the C ABI).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/mcx.h"

// ---- HBM layout ------------------------------------------------------------------------------
// Hot molecule record: 32 bytes = exactly one DRAM/L2 sector, so a neighbour gather costs one
// sector per candidate and the streaming pass is perfectly coalesced (32 lanes x 32 B = 1 KiB).
// sf = species (low 16 bits) | device flags (high 16 bits).
struct __align__(32) MolRec {
  double x, y, z;
  uint32_t id;
  uint32_t sf;
};
static_assert(sizeof(MolRec) == 32, "MolRec must be one 32-byte sector");

enum : uint32_t {
  DF_DEAD = 1u << 16,        // consumed / defunct (tombstone until the next sort drops it)
  DF_SCHED_UNIMOL = 1u << 17,  // MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN
  DF_PARTIAL = 1u << 18,     // cold t_sched[] holds a fractional diffusion_time
  DF_HAS_UNIMOL = 1u << 19,  // cold t_unimol[] holds a scheduled unimolecular time
  DF_CVI_PENDING = 1u << 20, // MCX_MOL_CVI_PENDING: the counted volume is a guess, a ray cast at the next evaluation replaces it
  DF_SURF = 1u << 21,        // surface molecule: cold swall/stile/suv hold Molecule::s (src4/molecule.h)
  DF_ORIENT_UP = 1u << 22,   // s.orientation == ORIENTATION_UP (else DOWN)
  DF_CREATED_ON_SURF = 1u << 23,  // volume product of a surface reaction: cold swall/stile hold where it was created
                                  // (DiffuseAction::where_created_this_iteration, diffuse_react_event.cpp:877-885)
  SF_SPECIES_MASK = 0xFFFFu,
  SF_CVI_SHIFT = 24,         // bits 24..31: v.counted_volume_index (rides along in the flags of every evaluation)
  SF_CVI_MASK = 0xFF000000u
};

struct DevSpecies { double space_step, time_step; uint32_t flags; uint32_t can_vol_react; uint32_t can_vol_surf; uint32_t can_surf_surf; };
struct DevClass { double max_fixed_p; uint32_t kind, r0, r1, first_pathway, n_pathways; int geom0, geom1; uint32_t pad; };
struct DevPathway { double cum_prob; uint32_t n_products, products[MCX_MAX_PRODUCTS], keep_mask, rule_id; int prod_orient[MCX_MAX_PRODUCTS]; uint32_t kept_info;
                    uint32_t general; };  // general: creates more surface products than it frees tiles (products on vacant neighbour tiles)

// per-wall surface grid (Grid::initialize, src4/wall.cpp:38-74) + the wall's first entry in the tile table
// one side of a triangle (src4/wall.h:32-90 Edge): the wall across it and the flattening transform between the uv frames
struct __align__(16) DevEdge { double cos_t, sin_t, tu, tv; uint32_t nb_wall; uint32_t forward; uint32_t pad[2]; };

struct __align__(16) DevGrid {
  double strip_width_rcp, vert2_slope, fullslope, binding_factor, vert0_u, vert0_v;
  int n_axis;
  uint32_t tile_start;
};

// per-wall constants (Wall::initialize_wall_constants, src4/wall.cpp:281-342), 128 B/wall
struct __align__(16) DevWall {
  double nx, ny, nz, dist;        // WallCollisionRejectionData (partition.h:1157-1159)
  double ux, uy, uz;              // unit_u
  double vx, vy, vz;              // unit_v
  double uv1u, uv2u, uv2v;        // uv_vert1_u, uv_vert2
  double v0x, v0y, v0z;           // vertex 0
};

#define MCX_KEPT_AT_WALL 0xFFFFFFFEu  // stile of a DF_CREATED_ON_SURF record that is a KEPT reactant, not a product (mcx_device.cuh)
#define MCX_MAX_COUNTED 1024      // species and reaction rules the device counters hold (a power of two)
#define MCX_ROUNDS_MAX 32        // upper bound of mcx_config::max_resolve_rounds
struct Counters {
  // population
  unsigned int n_slots;        // records in the current snapshot buffer (incl. tombstones/ghosts)
  unsigned int n_prod;         // products appended behind n_slots this iteration
  unsigned int n_fresh_events, n_fresh_ids;  // events of this iteration that need fresh molecule ids / how many ids
  unsigned int n_disk;         // molecules handed to the exact_disk launch of the generic pass (list in pend[1])
  unsigned int n_next;         // records binned for the next snapshot
  int error;                   // first MCX_ERR_* raised on the device
  unsigned int error_id;       // molecule id that raised it
  unsigned int next_id;        // fresh molecule ids
  unsigned int n_emigrants[2]; // multi-GPU: records leaving through the low/high slab face
  unsigned int n_slow;         // entries of slow_list this iteration
  unsigned int n_send[2];      // multi-GPU: halo records packed for the low / high neighbour
  unsigned int n_second;       // entries of second_list this iteration
  unsigned int n_slow2;        // entries of slow2_list this iteration
  // statistics (SimulationStats mirror)
  unsigned long long molecule_steps, ray_polygon_tests, ray_polygon_colls, reflections, transparent,
      absorptions, volvol_collisions, bimol_rxns, unimol_rxns, redos, retries, unresolved, products, deferred;
  unsigned long long defer_reason[8];  // why the fast pass deferred a molecule (MCX_DEFER_*)
  unsigned long long species_count[MCX_MAX_COUNTED];
  unsigned long long species_next[MCX_MAX_COUNTED];  // multi-GPU: recount of owned molecules during the scatter
  unsigned long long rxn_count[MCX_MAX_COUNTED];
  // conflict rounds: proposals entering round r (list pend[0]) and losers of round r (list pend[1])
  unsigned int n_prop[MCX_ROUNDS_MAX + 1], n_lose[MCX_ROUNDS_MAX + 1];
};
#define MCX_MAX_CV 256
#define MCX_FW_MARGIN 1e-6        // inflation of wall and query boxes of the fine wall grid, in length units

// Shared-memory tiles of the fast pass (k_diffuse_tile, mcx_tile.cuh): a thread block owns the records of TX x TY x TZ
// cells and stages them plus a halo (hx cells in x, one cell row in y and z) with one bulk copy per cell row; the
// staged records are re-binned in shared memory into a finer grid (rows split 2^lsy x 2^lsz) of fp32 positions
// relative to the tile, which the partner probe pre-filters before the exact fp64 test.  Planned on the host from the
// population density (mcx_api.cu: plan_tiles); enabled == 0: the flat gather kernel runs instead.
struct TileGeom {
  int enabled;
  int TX, TY, TZ, hx;          // owned cells per tile; x halo in cells
  int ntx, nty, ntz;           // tiles per axis of the local cell grid
  unsigned int n_tiles;
  int lsy, lsz;                // log2 of the fine sub-rows per cell row in y / z
  int nfx, nfy, nfz;           // fine grid of the staged region: TX + 2 hx, (TY + 2) << lsy, (TZ + 2) << lsz
  unsigned int cap;            // staged records that fit
  float tol_d, r2p;            // slacks of the fp32 pre-filter: along the move (absolute) and the inflated R^2
  unsigned int smem_bytes;     // dynamic shared memory of the launch
};

// An event that creates more products than it frees reactant ids (mcx_kernels.cu: k_assign_ids).  Fresh ids are handed
// out AFTER the conflict rounds, in the order (cell group of the event position, id of the initiator) — a function of the
// events only, so a run is reproducible and does not depend on the number of ranks.
struct FreshEvent { uint32_t first_slot, n, init_id, next, group; };

struct DevParams {
  // partition / subpartition grid (reference semantics)
  double ox, oy, oz, part_len, sp_len, sp_rcp, R;
  int n_sp, use_expanded;
  // device neighbour-cell grid
  double cgx, cgy, cgz, cell_rcp_x, cell_rcp_y, cell_rcp_z;  // cells are short in x (rows are contiguous), long in y/z
  int ncx, ncy, ncz;
  int rb_log2, nby;             // cell rows are stored in blocks of 2^rb_log2 x 2^rb_log2 (y, z) rows; nby = blocks per layer of blocks
  unsigned int n_cells;         // ncx * rows incl. the padding of the row blocks
  // immutable tables
  const DevWall* walls;
  const uint32_t* wall_tri;
  const double* verts;
  const uint32_t* wall_class;
  const uint32_t* spw_start;
  const uint32_t* spw_list;
  // fine wall grid: every subpartition split into fw_K^3 cells; per cell the walls of that subpartition whose
  // bounding box (+ MCX_FW_MARGIN) overlaps it, ascending like spw_list (fw_K == 1: the subpartition lists)
  const uint32_t* fw_start;
  const uint32_t* fw_list;
  int fw_K;
  double fw_rcp;                // fw_K / sp_len
  const uint8_t* sp_flags;      // per subpartition: bit0 = holds walls, bit1 = any of its 3x3x3 neighbours holds walls
  const DevSpecies* species;
  const int* bimol;
  const int* unimol;
  const DevClass* classes;
  const DevPathway* pathways;
  const uint8_t* surf_action;   // [species][surf_class][side(0 front,1 back)]
  const uint8_t* surf_border;   // same index, side = orientation of a surface molecule (0 up, 1 down): what a region border of that class does
  const uint8_t* wall_border;   // per wall: bit e = edge e is a border of a reactive region (mcx_set_region_borders); null = none
  const int* surf_rxn;          // same index: the MCX_RXN_BIMOL_VOLWALL class of a MCX_SURF_STANDARD entry
  const uint8_t* exd_skip;      // [species][surf_class]: exact_disk ignores the wall (the species travels through it)
  const uint16_t* wall_cv;      // per wall: counted volume on the front side | on the back side << 8; null = none
  unsigned long long* rxn_count_cv;  // [rule * n_cv + cv]
  unsigned long long* mol_count_cv;  // [species * n_cv + cv], filled by mcx_counts_by_volume
  unsigned int n_cv;
  const uint32_t* cv_mask;      // per counted volume: the counted objects enclosing it (mcx_set_counted_volume_objects); null = off
  uint32_t cv_xor, cv_all;      // objects whose walls toggle membership instead of naming a pair of volumes; all counted objects
  const uint32_t* wall_obj;     // per wall: geometry object index (mcx_set_geometry's wall_object; null = one object)
  const uint8_t* wall_rs;       // per wall: index of the set of counted surface regions it belongs to; null = none
  unsigned long long* rxn_count_rs;  // [rule * n_rs + region set]: reactions whose initiator was a surface molecule there
  unsigned long long* mol_count_rs;  // [species * n_rs + region set], filled by mcx_counts_by_surface_region
  unsigned int n_rs;
  int n_species, n_surf_classes, n_walls;
  // surface molecules: tile table of the current snapshot and per-slot cold fields
  const DevGrid* grids;         // per wall
  const int* volsurf;           // [volume species][surface species] -> class or -1
  uint32_t* tile_slot;          // per tile: slot in recA of the occupant, MCX_NONE = vacant (rebuilt by the scatter)
  unsigned int n_tiles;
  int has_surf;                 // any surface species / vol-surf class: the cold surface arrays exist
  uint32_t *swallA, *swallB;    // s.wall_index (or creation wall of a DF_CREATED_ON_SURF volume product)
  uint32_t *stileA, *stileB;    // s.grid_tile_index (or creation tile)
  double2 *suvA, *suvB;         // s.pos
  const DevEdge* edges;         // 3 per wall
  unsigned long long* tile_claim;  // per tile: (epoch << 32) | ~id of the best mover claiming it
  // surface-surface reactions (react_2D_all_neighbors): class table, the neighbour tiles of every tile as (wall, tile)
  // pairs in the reference's list order (mcx_geom.cpp: tile_neighbor_table), and which walls have a grid yet
  const int* surfsurf;          // [surface species][surface species] -> class or -1; null = no such class
  const uint32_t* tn_start;     // per tile (+ 1)
  const uint2* tn_list;
  uint8_t* wall_has_grid;       // Wall::has_initialized_grid: the wall has held a surface molecule since the last upload
  // products on vacant neighbour tiles (pathways flagged DevPathway::general; null without such a pathway): per slot, where
  // the created surface products of the pending proposal go — written by the evaluation, read by the conflict rounds
  uint2* prop_ptile;            // [slot * MCX_MAX_PRODUCTS + c]: wall, tile of created surface product c
  double2* prop_puv;            // ... its uv
  uint32_t* prop_pmask;         // [slot]: number of created surface products | (bit 4 + c: product c sits on a vacant tile the event claims)
  // rng
  unsigned long long seed, iteration;
  int rng_mode;
  const uint32_t* tape;
  unsigned long long n_words;
  const unsigned long long* tape_off;
  unsigned long long n_ids;
  // molecule state
  MolRec* recA;      // snapshot (sorted by cell), read by everyone
  MolRec* recB;      // results of this iteration (same slot), products appended
  double* tschedA; double* tschedB;
  double* tuniA; double* tuniB;
  uint32_t* rank;    // rank inside the destination cell, MCX_NONE = not carried over
  uint32_t* cs_cur;  // cell_start of snapshot A (n_cells + 1)
  uint32_t* cs_next; // histogram -> cell_start of the next snapshot
  unsigned int* scan_sums;  // scratch of the cell-histogram scan
  unsigned long long* claim;   // per slot: (epoch << 32) | ~priority
  uint32_t* prop_partner;      // per slot: partner slot of the pending proposal
  uint32_t* prop_info;         // per slot: kind(4) | pathway(8) | orientation bits(7) | class(13)
  double* prop_t;              // per slot: absolute event time
  uint32_t* pend[2];           // pending lists (slot indices)
  uint32_t* slow_list;         // slots the fast diffuse pass deferred to the generic evaluation
  uint32_t* second_list;       // slots k_diffuse_fast<1> evaluates (two sub-steps)
  uint32_t* slow2_list;        // slots k_diffuse_fast<1> deferred to the generic evaluation
  unsigned int capacity;
  unsigned int max_rounds;
  Counters* ctr;
  mcx_trace_rec* trace;
  unsigned long long n_trace;
  // slab decomposition (multi-GPU): the local cell grid covers the owned z-layers [own_lo, own_hi) plus halo
  // layers on each side that has a neighbour; world == 1: everything is owned
  int own_lo, own_hi, world, halo_layers;
  int z_off;                    // global z-layer index of local layer 0 (cgz stays the GLOBAL grid origin, so every
                                // rank computes the same global layer for a position before subtracting its offset)
  int has_low, has_high;        // a neighbour rank exists below / above
  int sm_count;                 // multiprocessors of this device: every grid is sized in multiples of it
  TileGeom tile;
  // fresh molecule ids (FreshEvent): event list, per cell group (16 x-cells of one cell row) the head of its chain of
  // events and the number of fresh ids (exclusive prefix after the scan)
  FreshEvent* fresh_list;
  uint32_t* fresh_head;
  uint32_t* fresh_pref;
  unsigned int fresh_cap, n_groups, grp_x;
  const uint32_t* rank_fresh;   // multi-GPU: fresh ids of every rank this iteration (all-gathered), null on one device
  int my_rank;
};

// multi-GPU halo record: what a neighbour needs to evaluate a molecule exactly like its owner does
// rec (32 B) | tsched, tuni (16 B, records with DF_PARTIAL / DF_HAS_UNIMOL) | Molecule::s of surface molecules and the
// creation wall / tile of DF_CREATED_ON_SURF volume products (32 B); the peer-memory path stores only the parts a
// record has, the NCCL fallback moves whole records
struct HaloRec { MolRec rec; double tsched, tuni; uint32_t swall, stile; uint32_t pad_[2]; double su, sv; double pad2_[2]; };
static_assert(sizeof(HaloRec) == 96, "HaloRec layout (MolRec is 32-byte aligned)");

// peer-memory halo exchange (mcx_comm.cu): where this rank's pack kernel writes and where its unpack kernel reads
struct HaloP2P {
  HaloRec* peer_recv[2];             // low / high neighbour's receive buffer for my records (peer memory, null: no neighbour)
  unsigned long long* peer_flag[2];  // its flag word: (tag << 32) | count, release-stored after the records
  const HaloRec* my_recv[2];         // my receive buffers (local memory) the neighbours write into
  unsigned long long* my_flag[2];
  unsigned int tag, cap;
  unsigned int* done;                // pack kernel block counter (zero between launches)
};

// kernels / launchers implemented in mcx_kernels.cu
struct StepPlan {
  int sm_count;
  unsigned long long* launches;  // host-side counter of kernels launched (may be null)
  cudaEvent_t* prof;             // 5 events for this iteration (null = no per-kernel timing)
  bool has_claims;   // model can produce reactions / absorptions (conflict rounds needed)
  bool has_fresh;    // some pathway creates more products than it consumes reactants (fresh molecule ids needed)
  bool trace;
};
void mcx_plan_tiles(DevParams& p, unsigned long long n_records);  // fills p.tile for a population of n_records
void mcx_launch_iteration(const DevParams& p, const StepPlan& plan, cudaStream_t s);
void mcx_launch_initial_sort(const DevParams& p, const StepPlan& plan, cudaStream_t s);
// multi-GPU pieces of an iteration (mcx_comm.cu drives them around the NCCL exchange)
void mcx_launch_evaluate(const DevParams& p, const StepPlan& plan, cudaStream_t s);   // memset + diffuse + resolve rounds
void mcx_launch_fresh_scan(const DevParams& p, const StepPlan& plan, cudaStream_t s);  // prefix of the fresh ids per cell group
void mcx_launch_assign_ids(const DevParams& p, const StepPlan& plan, cudaStream_t s);  // fresh ids into the product records
void mcx_launch_release(const DevParams& p, const mcx_release& r, uint32_t first_id, cudaStream_t s);
// surface release (mcx_release_surface_molecules): one placement round / the fall-back fill
struct SurfRelease {
  const uint32_t* walls; const double* cum_area; const double* area; unsigned int n_walls; double total_area;
  uint32_t species; int orientation; uint32_t randomize_pos; double release_time; uint32_t first_id;
  uint32_t* claim;        // per tile: lowest molecule index (k) that picked it, 0xFFFFFFFF = nobody yet
  uint32_t* choice;       // per molecule index: tile picked this round (global tile index) or MCX_NONE
  uint32_t* choice_wall;  // ... and its wall
};
void mcx_launch_surface_release_round(const DevParams& p, const SurfRelease& r, const uint32_t* pend_in, unsigned int n_pend,
                                      uint32_t* pend_out, unsigned int* n_out, unsigned int round, cudaStream_t s);
void mcx_launch_surface_release_count_vacant(const DevParams& p, const SurfRelease& r, unsigned int* n_vacant, cudaStream_t s);
void mcx_launch_surface_release_fill(const DevParams& p, const SurfRelease& r, const uint32_t* pend_sorted, unsigned int n_pend,
                                     unsigned int* n_left, cudaStream_t s);
void mcx_launch_release_list(const DevParams& p, const double* x, const double* y, const double* z, const uint32_t* species,
                             const uint32_t* cv, uint64_t n, double release_time, uint32_t first_id, cudaStream_t s);  // appends behind a re-binned population
void mcx_launch_rebin(const DevParams& p, const StepPlan& plan, cudaStream_t s);      // A -> B unchanged (halo refresh without a step)
void mcx_launch_pack_halo(const DevParams& p, HaloRec* send_low, HaloRec* send_high, unsigned int cap, cudaStream_t s);
void mcx_launch_unpack_halo(const DevParams& p, const HaloRec* recv, unsigned int n, unsigned int offset, cudaStream_t s);
void mcx_launch_add_received(const DevParams& p, unsigned int n, cudaStream_t s);
void mcx_launch_halo_p2p(const DevParams& p, const HaloP2P& link, cudaStream_t s);  // pack+store to peers, acquire+unpack
void mcx_launch_sort(const DevParams& p, const StepPlan& plan, cudaStream_t s);
void mcx_launch_count_by_volume(const DevParams& p, cudaStream_t s);
void mcx_launch_count_by_surface_region(const DevParams& p, cudaStream_t s);
void mcx_launch_reset_population(const DevParams& p, unsigned int n_slots, cudaStream_t s);  // before an upload
// device staging of the surface part of mcx_mol_soa (all null: volume molecules only)
struct SurfSoa { const uint32_t* wall; const uint32_t* tile; const int32_t* orientation; const double* u; const double* v; const uint32_t* cv; };
struct SurfSoaOut { uint32_t* wall; uint32_t* tile; int32_t* orientation; double* u; double* v; uint32_t* cv; };
void mcx_launch_pack_soa(const DevParams& p, const double* x, const double* y, const double* z,
                         const uint32_t* id, const uint32_t* species, const uint32_t* flags,
                         const double* tsched, const double* tuni, SurfSoa sv, unsigned int n, cudaStream_t s);
void mcx_launch_unpack_soa(const DevParams& p, double* x, double* y, double* z, uint32_t* id,
                           uint32_t* species, uint32_t* flags, double* tsched, double* tuni, SurfSoaOut sv,
                           unsigned int* n_out, cudaStream_t s);
