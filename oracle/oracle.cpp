// oracle/oracle.cpp — TEST INFRASTRUCTURE: CPU restatement of MCell4's per-timestep
// diffuse-and-react hot path (DiffuseReactEvent) for volume molecules.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product (libmcx.so) never does.
//
// PARITY STATUS: pinned at function level and for the whole functions of diffuse_react_event.cpp that can be cut out of
// the engine; the loop around them is unpinned (DESIGN.md section 4 has the full list).
//   PINNED bit-for-bit against the reference's own compiled code (oracle/_ref, built by `make ref` from the
//   sources where they lie; golden outputs committed in tests/golden/):
//     * RNG: ISAAC64 + Ziggurat vs src/rng.c, src/isaac64.c           (tests/test_oracle_rng.py)
//     * MCell3 originals (src/wall_util.c, react_cond.c, react_util.c, util.c, diffuse.c, grid_util.c): wall constants,
//       collide_wall incl. every REDO / jump_away_line path, collide_mol, wall_in_box, test_bimolecular, pathway search,
//       distinguishable, pb_factor, unimolecular lifetimes, ray_trace_2D   (tests/test_oracle_vs_reference*.py)
//     * MCell4's own src4/ code cut out by line range and compiled unmodified: the subpartition walk, the leaf
//       arithmetic of the volume and surface path incl. exact_disk, the neighbour-tile search, and the whole functions
//       ray_trace_vol + sort_collisions_by_time, react_2D_all_neighbors, find_surf_product_positions
//                                                                       (tests/test_oracle_vs_reference_mcell4*.py)
//   UNPINNED: the loop of diffuse_vol_molecule / diffuse_surf_molecule around those functions (what follows each
//   collision, rescheduling).  The reference tree holds no golden vectors for it and diffuse_react_event.cpp as a whole
//   cannot be built here (absent libbng/nfsim/boost); it is a line-by-line restatement of the cited
//   src4 functions, checked by analytic expectations (MSD, uniform density, mass action, exponential decay) in
//   tests/test_oracle_physics.py.
//
// Absent third-party arithmetic: libbng (github.com/mcellteam/libbng, version unpinned —
// consumed as sibling checkout, CMakeLists.txt:146-148).  Restated from MCell3 originals:
// cmp_eq(a,b,eps) := fabs(a-b) < eps (ASSUMPTION, not in tree); distinguishable() src/util.c:449-463;
// get_pathway_index_for_probability := binary_search_double over cum_probs, src/react_cond.c:80-97.
//
// Two execution modes:
//   SEQUENTIAL — reference semantics: molecules one after another, one global ISAAC64 stream,
//                reactions applied immediately, products diffused from a FIFO in the same
//                iteration (diffuse_react_event.cpp:67-161).  Used for CPU timing and ensembles.
//   SNAPSHOT   — the parallel semantics the GPU implements: every molecule is evaluated against
//                the start-of-iteration state with its own word stream (tape or Philox);
//                reactions are proposals resolved in synchronous rounds.
//
// Gaps inside the path (identical in the product): DESIGN.md section 7.
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <array>
#include <deque>
#include <string>
#include <unordered_map>
#include <map>
#include <vector>

#include "../include/mcx.h"
#include "oracle_rng.h"
#include "oracle_exact_disk.h"

namespace orc {

// ---- constants (src/mcell_structs.h:238-241, src4/defines.h:178-182) -------------------
static const double EPS = 1e-12;
static const double SQRT_EPS = 1e-6;
static const double POS_EPS = EPS, STIME_EPS = EPS;
static const double POS_SQRT2 = 1.41421356238;  // src4/defines.h:182 (truncated on purpose)
static const double TIME_INVALID = -256, TIME_FOREVER = 1e20;

struct V3 { double x, y, z; };
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 mul(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
// glm::dot for dvec3: tmp = a*b; tmp.x + tmp.y + tmp.z  (libs/glm/detail/func_geometric.inl:48-55)
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// glm::cross (func_geometric.inl:68-79)
static inline V3 cross(V3 x, V3 y) {
  return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
static inline double len3_squared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline double fabs_p(double x) { return x < 0 ? -x : x; }
static inline double max3(V3 v) { return std::max(std::max(v.x, v.y), v.z); }
static inline double abs_max_2vec(V3 a, V3 b) {  // defines.h abs_max_2vec
  V3 m = {std::max(fabs(a.x), fabs(b.x)), std::max(fabs(a.y), fabs(b.y)), std::max(fabs(a.z), fabs(b.z))};
  return max3(m);
}
static inline bool cmp_eq(double a, double b, double eps = EPS) { return fabs(a - b) < eps; }  // ASSUMPTION
static inline bool cmp_lt(double a, double b, double eps) { return a < b && !cmp_eq(a, b, eps); }
static inline bool distinguishable(double a, double b, double eps) {  // src/util.c:449-463
  double c = fabs(a - b);
  a = fabs(a);
  if (a < 1) a = 1;
  b = fabs(b);
  if (b < a) eps *= a; else eps *= b;
  return c > eps;
}
static inline void guard_zero_div(V3& v) {  // defines.h guard_zero_div
  if (v.x == 0) v.x = FLT_MIN;
  if (v.y == 0) v.y = FLT_MIN;
  if (v.z == 0) v.z = FLT_MIN;
}

// ---- data ------------------------------------------------------------------------------
struct Wall {  // src4/wall.h Wall subset; constants per Wall::initialize_wall_constants (wall.cpp:281-342)
  uint32_t vi[3];
  V3 normal, unit_u, unit_v;
  double distance_to_origin, uv_vert1_u, uv_vert2_u, uv_vert2_v, area;
  uint32_t surf_class, object;
  uint8_t cv_front = 0, cv_back = 0;  // counted volume on the normal side / on the other side
  // src4/wall.h:32-90 Edge, one per triangle side: the neighbouring wall across it (MCX_NONE: free edge) and the
  // flattening transform between the two uv frames; is_forward: this wall is Edge::forward_index
  uint32_t nb_wall[3] = {MCX_NONE, MCX_NONE, MCX_NONE};
  bool edge_forward[3] = {false, false, false};
  double edge_cos[3] = {0, 0, 0}, edge_sin[3] = {0, 0, 0}, edge_tu[3] = {0, 0, 0}, edge_tv[3] = {0, 0, 0};
};

struct Mol {  // src4/molecule.h:52-260
  V3 pos;                 // v.pos; for a surface molecule uv2xyz(s.pos)
  uint32_t id, species, flags;
  double diffusion_time, unimol_rxn_time;
  uint32_t subpart;       // v.subpart_index
  uint32_t list_slot;     // position inside its reactant list (sequential mode)
  uint32_t reg_subpart;   // v.reactant_subpart_index
  // surface part (s.*): wall == MCX_NONE for a volume molecule
  uint32_t wall = MCX_NONE, tile = MCX_NONE;
  int orient = 0;
  double u = 0, v = 0;
  // DiffuseAction::where_created_this_iteration of a volume product of a surface reaction (:877-885)
  uint32_t created_wall = MCX_NONE, created_tile = MCX_NONE;
  uint32_t cvi = 0;       // v.counted_volume_index
};

struct Grid {  // src4/wall.h:101-193, constants per Grid::initialize (wall.cpp:38-74)
  int n_axis; uint32_t n_tiles;
  double strip_width_rcp, vert2_slope, fullslope, binding_factor, vert0_u, vert0_v;
};

enum { COLL_VOLMOL = 0, COLL_WALL_FRONT = 1, COLL_WALL_BACK = 2 };
// created_tile of a KEPT reactant of a surface / wall reaction (SNAPSHOT): not a rebinding guard — the molecule carries on
// from created_wall like the reference's does within the same step (last_hit_wall_index = wall, remaining displacement
// leading away from it, diffuse_react_event.cpp:945-975, reflect_from_wall): its first displacement is mirrored
// away from that wall if it points into it, and the first trace skips the wall
static const uint32_t KEPT_AT_WALL = 0xFFFFFFFEu;
struct Collision {  // src4/collision_structs.h:29-243
  int type; double time; V3 pos; uint32_t partner_id; uint32_t partner_index; int rxn_class; uint32_t wall;
};

enum { WALL_MISS, WALL_FRONT, WALL_BACK, WALL_REDO };

static inline uint64_t hash_ev(uint64_t h, uint32_t a, uint32_t b) {
  h = (h ^ a) * 0x100000001b3ULL;
  h = (h ^ b) * 0x100000001b3ULL;
  return h;
}
enum { EV_WALL = 0x57000000u, EV_COLL = 0xC0000000u, EV_RXN = 0xAE000000u, EV_ABSORB = 0xAB000000u,
       EV_REDO = 0x4ED00000u, EV_UNIMOL = 0x11000000u, EV_TRANSP = 0x7A000000u, EV_SURFMOL = 0x5F000000u, EV_BLOCKED = 0xB10C0000u, EV_DISK = 0xD1500000u,
       EV_SURFMOVE = 0x3E000000u, EV_WALLRXN = 0x9A000000u, EV_SURFSURF = 0x55000000u };

struct Stats {
  uint64_t molecule_steps = 0, ray_polygon_tests = 0, ray_polygon_colls = 0, reflections = 0,
           transparent = 0, absorptions = 0, volvol_collisions = 0, bimol_rxns = 0, unimol_rxns = 0,
           redos = 0, retries = 0, unresolved = 0, products = 0, collide_mol_tests = 0;
};

// Result of evaluating one molecule for (the rest of) one iteration.
struct Outcome {
  int kind = MCX_OUT_NONE;        // MCX_OUT_*
  V3 pos;                         // final position, or event position
  double t_now = 0;               // diffusion_time after evaluation
  double unimol_time = TIME_INVALID;
  uint32_t flags = 0;
  // claiming events
  int rxn_class = -1, pathway = -1;
  uint32_t partner_index = MCX_NONE, partner_id = MCX_NONE;
  double t_event = 0;
  bool initiator_is_reactant0 = true;
  uint32_t orient_bits = 0;       // bit k: random orientation drawn for products[k] (1 = up)
  uint32_t cvi = 0;               // counted volume of the molecule at the end of the evaluation / at the event
  uint32_t wall = MCX_NONE, tile = MCX_NONE; double u = 0, v = 0;  // surface molecule: where it is after the evaluation
  int coll_side = 0;              // volume-surface / volume-wall reaction: +1 the initiator hit the wall's front, -1 its back
  uint32_t hit_wall = MCX_NONE;   // volume-wall reaction: the wall
  uint32_t created_wall = MCX_NONE, created_tile = MCX_NONE;  // SNAPSHOT, kept initiator of a surface reaction: rebinding guard
  bool surf_moved = false;        // SNAPSHOT, surface-surface reaction: the initiator took a new tile first (claims it too)
  // SNAPSHOT, pathway with surface products on vacant neighbour tiles: where the created surface products go
  bool pl_general = false; int pl_n = 0; uint32_t pl_wall[MCX_MAX_PRODUCTS], pl_tile[MCX_MAX_PRODUCTS], pl_vacant = 0;
  double pl_u[MCX_MAX_PRODUCTS], pl_v[MCX_MAX_PRODUCTS];
};

struct World {
  mcx_config cfg;
  double sp_len, sp_rcp;
  uint32_t n_sp;  // subparts per edge
  std::vector<V3> verts;
  std::vector<Wall> walls;
  std::vector<std::vector<uint32_t>> walls_per_subpart;  // ascending wall indices (uint_set order)
  uint32_t sample_stride = 0, sample_offset = 0;  // > 0: step_snapshot only evaluates molecules with id % stride == offset, once, and changes nothing
  std::vector<mcx_species> species;
  std::vector<mcx_rxn_class> classes;
  std::vector<mcx_pathway> pathways;
  std::vector<int> bimol;        // [a*ns+b] -> class or -1
  std::vector<int> unimol;       // [a] -> class or -1
  std::vector<uint8_t> can_vol_react;
  std::vector<int> volsurf;      // [vol*ns+surf] -> class or -1
  std::vector<uint8_t> can_vol_surf;   // SPECIES_FLAG_CAN_VOLSURF
  std::vector<int> surfsurf;     // [a*ns+b] -> MCX_RXN_BIMOL_SURFSURF class or -1 (both orders)
  std::vector<uint8_t> can_surf_surf;  // SPECIES_FLAG_CAN_SURFSURF: the species takes part in a surface-surface class
  std::vector<Grid> grids;       // per wall
  std::vector<std::vector<uint32_t>> tiles;  // per wall: molecule id per tile (Grid::molecules_per_tile); empty =
                                             // grid not initialized (wall.h:339-346)
  std::vector<uint32_t> tile_start;          // first global tile of every wall (+ total)
  mutable std::vector<std::vector<uint32_t>> vertex_walls;  // Partition::walls_using_vertex_mapping, built on first use
  mutable bool assume_all_grids = false;  // find_neighbor_tiles with create_grid_flag (product placement): every wall counts
  std::vector<mcx_surf_class_rxn> surf_rules;
  std::vector<Mol> mols;
  std::vector<uint32_t> id_to_index;  // molecule_id_to_index_mapping
  std::vector<uint32_t> sched_ids;    // schedulable_molecule_ids
  // per (species, subpart) id lists: volume_molecule_reactants_per_reactant_class (partition.h:1125-1143)
  std::unordered_map<uint64_t, std::vector<uint32_t>> lists;
  uint32_t next_id = 0;
  uint64_t iteration = 0;
  Isaac64 rng;
  Stats stats;
  std::vector<uint64_t> species_count, rxn_count;
  uint32_t n_cv = 1;                       // counted volumes (index 0 = outside all)
  std::vector<uint64_t> rxn_count_cv;      // [rule * n_cv + cv] (inc_rxn_in_volume_occured_count)
  std::vector<uint32_t> cv_mask;           // per counted volume: the counted objects that enclose it (mcx_set_counted_volume_objects); empty = off
  uint32_t cv_xor = 0, cv_all = 0;         // objects whose walls toggle membership instead of naming a pair of volumes; all counted objects
  std::vector<uint8_t> wall_border;        // per wall: bit e = edge e is a border of a reactive region (mcx_set_region_borders); empty = none
  std::vector<uint8_t> wall_rs;            // per wall: set of counted surface regions (mcx_set_surface_regions); empty = none
  uint32_t n_rs = 0;
  std::vector<uint64_t> rxn_count_rs;      // [rule * n_rs + set] (inc_rxn_on_surface_occured_count, summed per region set)
  std::vector<mcx_trace_rec> trace;  // by id, when tracing
  bool tracing = false;
  std::string err;
  // replay tape recorded in sequential mode: words per molecule id for the last iteration
  std::vector<uint32_t> tape_words; std::vector<uint64_t> tape_off; std::vector<uint32_t> tape_len;
  bool record_tape = false;

  // -- subpart index math (partition.h:248-309)
  inline bool in_this_partition(V3 p) const {
    double e = cfg.partition_edge_length;
    return p.x >= cfg.origin[0] && p.y >= cfg.origin[1] && p.z >= cfg.origin[2] &&
           p.x < cfg.origin[0] + e && p.y < cfg.origin[1] + e && p.z < cfg.origin[2] + e;
  }
  inline void subpart_3d(V3 p, int idx[3]) const {
    idx[0] = (int)((p.x - cfg.origin[0]) * sp_rcp);
    idx[1] = (int)((p.y - cfg.origin[1]) * sp_rcp);
    idx[2] = (int)((p.z - cfg.origin[2]) * sp_rcp);
  }
  inline uint32_t subpart_from_3d(const int idx[3]) const {
    return (uint32_t)(idx[0] + idx[1] * (int)n_sp + idx[2] * (int)(n_sp * n_sp));
  }
  inline uint32_t subpart_index(V3 p) const { int i[3]; subpart_3d(p, i); return subpart_from_3d(i); }
  inline void subpart_3d_from_index(uint32_t s, int idx[3]) const {
    idx[0] = s % n_sp; idx[1] = (s / n_sp) % n_sp; idx[2] = (s / (n_sp * n_sp)) % n_sp;
  }
  inline bool idx_in_range(int i) const { return i >= 0 && i < (int)n_sp; }
  inline bool is_surf(uint32_t species) const { return !(this->species[species].flags & MCX_SP_VOL); }
};

// ---- geometry set-up ---------------------------------------------------------------------
// Wall::initialize_wall_constants, src4/wall.cpp:281-342
static void init_wall_constants(const World& w, Wall& f) {
  V3 v0 = w.verts[f.vi[0]], v1 = w.verts[f.vi[1]], v2 = w.verts[f.vi[2]];
  V3 vA = v1 - v0, vB = v2 - v0, vX = cross(vA, vB);
  f.area = 0.5 * sqrt(len3_squared(vX));
  if (!distinguishable(f.area, 0, EPS)) {
    f.normal = f.unit_u = f.unit_v = {0, 0, 0};
    f.uv_vert1_u = f.uv_vert2_u = f.uv_vert2_v = f.distance_to_origin = 0;
    return;
  }
  V3 f1 = v1 - v0;
  double inv_f1_len = 1 / sqrt(len3_squared(f1));
  f.unit_u = f1 * inv_f1_len;
  V3 f2 = v2 - v0;
  f.normal = cross(f.unit_u, f2);
  double inv_norm_len = 1 / sqrt(len3_squared(f.normal));
  f.normal = f.normal * inv_norm_len;
  f.unit_v = cross(f.normal, f.unit_u);
  f.distance_to_origin = dot(v0, f.normal);
  f.uv_vert1_u = dot(f1, f.unit_u);
  f.uv_vert2_u = dot(f2, f.unit_u);
  f.uv_vert2_v = dot(f2, f.unit_v);
}

// Grid::initialize, src4/wall.cpp:38-74
static void grid_init(const World& w, const Wall& f, Grid& g) {
  g.n_axis = (int)ceil(sqrt(f.area));
  if (g.n_axis < 1) g.n_axis = 1;
  g.n_tiles = (uint32_t)(g.n_axis * g.n_axis);
  g.strip_width_rcp = 1 / (f.uv_vert2_v / ((double)g.n_axis));
  g.vert2_slope = f.uv_vert2_u / f.uv_vert2_v;
  g.fullslope = f.uv_vert1_u / f.uv_vert2_v;
  g.binding_factor = ((double)g.n_tiles) / f.area;
  V3 v0 = w.verts[f.vi[0]];
  g.vert0_u = dot(v0, f.unit_u);
  g.vert0_v = dot(v0, f.unit_v);
}
// distinguishable_vec3, src4/defines.h:766-806
static bool distinguishable_vec3(V3 a, V3 b, double eps) {
  double c = fabs(a.x), cc, d;
  d = fabs(a.y); if (d > c) c = d;
  d = fabs(a.z); if (d > c) c = d;
  d = fabs(b.x); if (d > c) c = d;
  d = fabs(b.y); if (d > c) c = d;
  d = fabs(b.z); if (d > c) c = d;
  cc = fabs(a.x - b.x);
  d = fabs(a.y - b.y); if (d > cc) cc = d;
  d = fabs(a.z - b.z); if (d > cc) cc = d;
  if (c < eps) c = eps;
  return c * eps < cc;
}
// GridUtils::xyz2grid_tile_index, src4/grid_utils.inl:48-118 (== xyz2grid, src/grid_util.c:74-128)
static uint32_t xyz2grid(const World& w, V3 v, const Wall& f, const Grid& g) {
  if (g.n_tiles == 1) return 0;
  uint32_t tile_idx_mid = g.n_tiles - 2 * (uint32_t)g.n_axis + 1, tile_idx_last = g.n_tiles - 1;
  if (!distinguishable_vec3(v, w.verts[f.vi[0]], POS_EPS)) return tile_idx_mid;
  if (!distinguishable_vec3(v, w.verts[f.vi[1]], POS_EPS)) return tile_idx_last;
  if (!distinguishable_vec3(v, w.verts[f.vi[2]], POS_EPS)) return 0;
  double i = dot(v, f.unit_u) - g.vert0_u;
  double j = dot(v, f.unit_v) - g.vert0_v;
  double striploc = j * g.strip_width_rcp;
  int strip = (int)striploc;
  double striprem = striploc - strip;
  strip = g.n_axis - strip - 1;
  double u0 = j * g.vert2_slope;
  double u1_u0 = f.uv_vert1_u - j * g.fullslope;
  double stripeloc = ((i - u0) / u1_u0) * (strip + (1 - striprem));
  int stripe = (int)stripeloc;
  double striperem = stripeloc - stripe;
  int flip = (striperem < 1 - striprem) ? 0 : 1;
  int idx = strip * strip + 2 * stripe + flip;
  if (idx < 0) idx = 0;                                   // the reference raises an internal error here
  if ((uint32_t)idx >= g.n_tiles) idx = (int)g.n_tiles - 1;
  return (uint32_t)idx;
}
// GridUtils::grid2uv, src4/grid_utils.inl:233-253
static void grid2uv(const Wall& f, const Grid& g, uint32_t index, double& u, double& v) {
  int root = (int)(sqrt((double)index));
  int rootrem = (int)index - root * root;
  int k = g.n_axis - root - 1;
  int j = rootrem / 2;
  int i = rootrem - 2 * j;
  double over3n = 1 / (double)(3 * g.n_axis);
  u = ((double)(3 * j + i + 1)) * over3n * f.uv_vert1_u + ((double)(3 * k + i + 1)) * over3n * f.uv_vert2_u;
  v = ((double)(3 * k + i + 1)) * over3n * f.uv_vert2_v;
}
// GeometryUtils::uv2xyz, src4/geometry_utils.h:29-35
static V3 uv2xyz(const World& w, const Wall& f, double u, double v) {
  V3 v0 = w.verts[f.vi[0]];
  return {u * f.unit_u.x + v * f.unit_v.x + v0.x, u * f.unit_u.y + v * f.unit_v.y + v0.y,
          u * f.unit_u.z + v * f.unit_v.z + v0.z};
}

// surface_net (src4/geometry.cpp:258-356) + Edge::reinit_edge_constants (src4/wall.cpp:134-235): pair the sides of the
// triangles of one object that join the same two points in opposite directions; the face that comes first is the
// edge's forward wall and the transform is set up from its side.  (Manifold meshes: exactly two faces per edge.)
static void init_edges(World& w, uint32_t first_wall, uint32_t n_faces) {
  struct Key { double a[6]; bool operator<(const Key& o) const { return std::lexicographical_compare(a, a + 6, o.a, o.a + 6); } };
  std::map<Key, std::pair<uint32_t, int>> open_edges;  // undirected edge -> (face, side) of the first face seen
  auto key_of = [](V3 p, V3 q) {
    Key k;
    bool swap = std::lexicographical_compare(&q.x, &q.x + 3, &p.x, &p.x + 3);
    V3 lo = swap ? q : p, hi = swap ? p : q;
    k.a[0] = lo.x; k.a[1] = lo.y; k.a[2] = lo.z; k.a[3] = hi.x; k.a[4] = hi.y; k.a[5] = hi.z;
    return k;
  };
  auto same = [](V3 p, V3 q) { return p.x == q.x && p.y == q.y && p.z == q.z; };
  for (uint32_t fi = 0; fi < n_faces; fi++) {
    Wall& wb = w.walls[first_wall + fi];
    for (int j = 0; j < 3; j++) {
      int k = j + 1 < 3 ? j + 1 : 0;
      V3 pj = w.verts[wb.vi[j]], pk = w.verts[wb.vi[k]];
      Key key = key_of(pj, pk);
      auto it = open_edges.find(key);
      if (it == open_edges.end()) { open_edges[key] = {first_wall + fi, j}; continue; }
      uint32_t f0 = it->second.first; int e0 = it->second.second;
      Wall& wf = w.walls[f0];
      // compatible_edges (geometry.cpp:50-96): traversed in opposite directions, third vertices differ
      V3 a0 = w.verts[wf.vi[e0]], a1 = w.verts[wf.vi[e0 == 2 ? 0 : e0 + 1]], a2 = w.verts[wf.vi[e0 == 0 ? 2 : e0 - 1]];
      V3 b2 = w.verts[wb.vi[j == 0 ? 2 : j - 1]];
      if (!(same(a0, pk) && same(a1, pj) && !same(a2, b2)) || f0 == first_wall + fi) continue;
      open_edges.erase(it);
      // Edge::reinit_edge_constants with forward = f0, backward = this face, edge_num_used_for_init = e0
      int i = e0, jj = i + 1 == 3 ? 0 : i + 1;
      V3 wf0 = w.verts[wf.vi[0]], wfi = w.verts[wf.vi[i]], wfj = w.verts[wf.vi[jj]], wb0 = w.verts[wb.vi[0]];
      V3 di0 = wfi - wf0;
      double Ofu = dot(di0, wf.unit_u), Ofv = dot(di0, wf.unit_v);
      V3 dj0 = wfj - wf0;
      double tfu = dot(dj0, wf.unit_u) - Ofu, tfv = dot(dj0, wf.unit_v) - Ofv;
      double d_f = 1 / sqrt(tfu * tfu + tfv * tfv);
      double efu = tfu * d_f, efv = tfv * d_f, ffu = -efv, ffv = efu;
      V3 dib = wfi - wb0;
      double Obu = dot(dib, wb.unit_u), Obv = dot(dib, wb.unit_v);
      V3 djb = wfj - wb0;
      double tbu = dot(djb, wb.unit_u) - Obu, tbv = dot(djb, wb.unit_v) - Obv;
      double d_b = 1 / sqrt(tbu * tbu + tbv * tbv);
      double ebu = tbu * d_b, ebv = tbv * d_b, fbu = -ebv, fbv = ebu;
      double m00 = efu * ebu + ffu * fbu, m01 = efv * ebu + ffv * fbu;
      double m10 = efu * ebv + ffu * fbv, m11 = efv * ebv + ffv * fbv;
      double qu = Obu, qv = Obv;
      qu -= m00 * Ofu + m01 * Ofv;
      qv -= m10 * Ofu + m11 * Ofv;
      wf.nb_wall[e0] = first_wall + fi; wf.edge_forward[e0] = true;
      wb.nb_wall[j] = f0; wb.edge_forward[j] = false;
      wf.edge_cos[e0] = wb.edge_cos[j] = m00; wf.edge_sin[e0] = wb.edge_sin[j] = m01;
      wf.edge_tu[e0] = wb.edge_tu[j] = qu; wf.edge_tv[e0] = wb.edge_tv[j] = qv;
    }
  }
}
#include "oracle_tiles.h"  // find_neighbor_tiles (grid_utils.inl:296-1801)

// distinguishable_vec2, src4/defines.h:733-764
static bool distinguishable_vec2(double au, double av, double bu, double bv, double eps) {
  double c = fabs(au), cc, d;
  d = fabs(av); if (d > c) c = d;
  d = fabs(bu); if (d > c) c = d;
  d = fabs(bv); if (d > c) c = d;
  cc = fabs(au - bu);
  d = fabs(av - bv); if (d > cc) cc = d;
  if (c < eps) c = eps;
  return c * eps < cc;
}
// GridUtils::uv2grid_tile_index, src4/grid_utils.inl:119-190
static uint32_t uv2grid(const Wall& f, const Grid& g, double u, double v) {
  if (g.n_tiles == 1) return 0;
  uint32_t tile_idx_mid = g.n_tiles - 2 * (uint32_t)g.n_axis + 1, tile_idx_last = g.n_tiles - 1;
  if (!distinguishable_vec2(u, v, 0, 0, POS_EPS)) return tile_idx_mid;
  if (!distinguishable_vec2(u, v, f.uv_vert1_u, 0, POS_EPS)) return 0;
  if (!distinguishable_vec2(u, v, f.uv_vert2_u, f.uv_vert2_v, POS_EPS)) return tile_idx_last;
  double i = u, j = v;
  double striploc = j * g.strip_width_rcp;
  int strip = (int)striploc;
  double striprem = striploc - strip;
  strip = g.n_axis - strip - 1;
  double u0 = j * g.vert2_slope;
  double u1_u0 = f.uv_vert1_u - j * g.fullslope;
  double stripeloc = ((i - u0) / u1_u0) * (strip + (1 - striprem));
  int stripe = (int)stripeloc;
  double striperem = stripeloc - stripe;
  int flip = (striperem < 1 - striprem) ? 0 : 1;
  int idx = strip * strip + 2 * stripe + flip;
  if (idx < 0 || (uint32_t)idx >= g.n_tiles) return MCX_NONE;  // the reference raises an internal error
  return (uint32_t)idx;
}
// GeometryUtils::find_edge_point, src4/geometry_utils.inl:222-291.  0,1,2: edge hit; 3: stays within the wall; 4: cannot tell
enum { EDGE_WITHIN_WALL = 3, EDGE_CANNOT_TELL = 4 };
static int find_edge_point(const Wall& here, double lu, double lv, double du, double dv, double& eu, double& ev) {
  double lxd = lu * dv - lv * du;
  double lxc1 = -lv * here.uv_vert1_u;
  double dxc1 = -dv * here.uv_vert1_u;
  double f, s, t;
  if (dxc1 < -POS_EPS || dxc1 > POS_EPS) {
    f = 1 / dxc1;
    s = -lxd * f;
    if (0 < s && s < 1 && f > 0) {
      t = -lxc1 * f;
      if (POS_EPS < t && t < 1) { eu = lu + t * du; ev = lv + t * dv; return 0; }
      else if (t > 1 + POS_EPS) return EDGE_WITHIN_WALL;
    }
  }
  double lxc2 = lu * here.uv_vert2_v - lv * here.uv_vert2_u;
  double dxc2 = du * here.uv_vert2_v - dv * here.uv_vert2_u;
  if (dxc2 < -POS_EPS || dxc2 > POS_EPS) {
    f = 1 / dxc2;
    s = 1 + lxd * f;
    if (0 < s && s < 1 && f < 0) {
      t = -lxc2 * f;
      if (POS_EPS < t && t < 1) { eu = lu + t * du; ev = lv + t * dv; return 2; }
      else if (t > 1 + POS_EPS) return EDGE_WITHIN_WALL;
    }
  }
  f = dxc2 - dxc1;
  if (f < -POS_EPS || f > POS_EPS) {
    f = 1 / f;
    s = -(lxd + dxc1) * f;
    if (0 < s && s < 1 && f > 0) {
      t = (here.uv_vert1_u * here.uv_vert2_v + lxc1 - lxc2) * f;
      if (POS_EPS < t && t < 1) { eu = lu + t * du; ev = lv + t * dv; return 1; }
      else if (t > 1 + POS_EPS) return EDGE_WITHIN_WALL;
    }
  }
  return EDGE_CANNOT_TELL;
}
// GeometryUtils::traverse_surface, src4/geometry_utils.inl:305-342
static uint32_t traverse_surface(const Wall& here, double lu, double lv, int which, double& nu, double& nv) {
  if (here.nb_wall[which] == MCX_NONE) return MCX_NONE;
  double c = here.edge_cos[which], sn = here.edge_sin[which], tu = here.edge_tu[which], tv = here.edge_tv[which];
  if (here.edge_forward[which]) {
    double ru = c * lu + sn * lv, rv = -sn * lu + c * lv;
    nu = ru + tu; nv = rv + tv;
  } else {
    double ru = lu - tu, rv = lv - tv;
    nu = c * ru - sn * rv;
    nv = sn * ru + c * rv;
  }
  return here.nb_wall[which];
}
// ray_trace_surf, src4/diffuse_react_event.cpp:1578-1725 (no region borders: species.can_interact_with_border() is
// false).  Returns the wall the move ends on (MCX_NONE: ambiguous edge hit, pick another displacement) and the
// end point in that wall's frame.
static int border_action(const World& w, uint32_t species, uint32_t surf_class, int orient);
// Region borders (species.can_interact_with_border(), :1627-1665): sm_species != MCX_NONE checks the edges that are borders of
// a reactive region — leaving it (reflect_absorb_inside_out) and entering one (reflect_absorb_outside_in, diffusion_utils.inl:
// 632-700): a REFLECTIVE class turns the molecule back at the edge, an ABSORPTIVE one takes it (*absorbed = true, MCX_NONE).
static uint32_t ray_trace_surf(const World& w, uint32_t wall_index, double pu, double pv, double du, double dv,
                               double& out_u, double& out_v, uint32_t sm_species = MCX_NONE, int sm_orient = 0, bool* absorbed = nullptr) {
  const bool borders = sm_species != MCX_NONE && !w.wall_border.empty();
  if (absorbed) *absorbed = false;
  const Wall* this_wall = &w.walls[wall_index];
  uint32_t this_index = wall_index;
  double this_u = pu, this_v = pv, disp_u = du, disp_v = dv;
  for (int guard = 0; guard < 10000; guard++) {
    double bu = 0, bv = 0;
    int edge = find_edge_point(*this_wall, this_u, this_v, disp_u, disp_v, bu, bv);
    if (edge == EDGE_CANNOT_TELL) return MCX_NONE;
    if (edge == EDGE_WITHIN_WALL) { out_u = this_u + disp_u; out_v = this_v + disp_v; return this_index; }
    double old_u = this_u, old_v = this_v;
    double nu, nv;
    bool reflect_now = false;
    if (borders && ((w.wall_border[this_index] >> edge) & 1)) {   // inside out
      const int act = border_action(w, sm_species, this_wall->surf_class, sm_orient);
      if (act == MCX_SURF_ABSORPTIVE) { if (absorbed) *absorbed = true; return MCX_NONE; }
      reflect_now = act == MCX_SURF_REFLECTIVE;
    }
    uint32_t target = reflect_now ? MCX_NONE : traverse_surface(*this_wall, old_u, old_v, edge, nu, nv);
    if (target != MCX_NONE && borders) {   // outside in: the shared edge in the neighbour's numbering
      const Wall& tw = w.walls[target];
      int te = -1;
      for (int e2 = 0; e2 < 3; e2++) if (tw.nb_wall[e2] == this_index) te = e2;
      if (te >= 0 && ((w.wall_border[target] >> te) & 1)) {
        const int act = border_action(w, sm_species, tw.surf_class, sm_orient);
        if (act == MCX_SURF_ABSORPTIVE) { if (absorbed) *absorbed = true; return MCX_NONE; }
        if (act == MCX_SURF_REFLECTIVE) target = MCX_NONE;
      }
    }
    if (target != MCX_NONE) {
      this_u = nu; this_v = nv;
      double su = old_u + disp_u, sv = old_v + disp_v;
      double tu2, tv2;
      traverse_surface(*this_wall, su, sv, edge, tu2, tv2);
      disp_u = tu2 - this_u; disp_v = tv2 - this_v;
      this_wall = &w.walls[target]; this_index = target;
      continue;
    }
    // free edge: reflect
    double ndu = disp_u - (bu - old_u), ndv = disp_v - (bv - old_v);
    if (edge == 0) ndv *= -1.0;
    else if (edge == 1) {
      double ru = -this_wall->uv_vert2_v, rv = this_wall->uv_vert2_u - this_wall->uv_vert1_u;
      double f = 1.0 / sqrt(ru * ru + rv * rv);
      ru *= f; rv *= f;
      f = 2.0 * (ndu * ru + ndv * rv);
      ndu -= f * ru; ndv -= f * rv;
    } else {
      double ru = this_wall->uv_vert2_v, rv = -this_wall->uv_vert2_u;
      double f = 1.0 / sqrt(ru * ru + rv * rv);
      ru *= f; rv *= f;
      f = 2.0 * (ndu * ru + ndv * rv);
      ndu -= f * ru; ndv -= f * rv;
    }
    this_u = bu; this_v = bv; disp_u = ndu; disp_v = ndv;
  }
  return MCX_NONE;
}

static inline bool point_in_box(V3 p, V3 llf, V3 urb) {
  return p.x >= llf.x && p.x <= urb.x && p.y >= llf.y && p.y <= urb.y && p.z >= llf.z && p.z <= urb.z;
}

// WallUtils::wall_in_box, src4/wall_utils.inl:326-504
static int wall_in_box(const World& w, const Wall& f, V3 llf, V3 urb) {
  const V3 vert[3] = {w.verts[f.vi[0]], w.verts[f.vi[1]], w.verts[f.vi[2]]};
  for (int i = 0; i < 3; i++)
    if (point_in_box(vert[i], llf, urb)) return 1;
  // any wall edge through a box face
  for (int i = 0; i < 3; i++) {
    const V3& v2 = vert[i];
    const V3& v1 = vert[i == 0 ? 2 : i - 1];
    double r, a3, a4;
    const double lo[3] = {llf.x, llf.y, llf.z}, hi[3] = {urb.x, urb.y, urb.z};
    const double p1[3] = {v1.x, v1.y, v1.z}, p2[3] = {v2.x, v2.y, v2.z};
    // axis order and the (a3,a4) pairing follow the reference: x:(y,z) y:(x,z) z:(y,x)
    const int oa[3][2] = {{1, 2}, {0, 2}, {1, 0}};
    for (int ax = 0; ax < 3; ax++) {
      for (int side = 0; side < 2; side++) {
        double pl = side == 0 ? lo[ax] : hi[ax];
        if ((p1[ax] <= pl && pl < p2[ax]) || (p1[ax] > pl && pl >= p2[ax])) {
          r = (pl - p1[ax]) / (p2[ax] - p1[ax]);
          int b = oa[ax][0], c = oa[ax][1];
          a3 = p1[b] + r * (p2[b] - p1[b]);
          a4 = p1[c] + r * (p2[c] - p1[c]);
          if (lo[b] <= a3 && a3 <= hi[b] && lo[c] <= a4 && a4 <= hi[c]) return 2 + ax * 2 + side;
        }
      }
    }
  }
  // any box edge through the wall
  static const int which_x1[12] = {0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 1};
  static const int which_y1[12] = {0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 0};
  static const int which_z1[12] = {0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0};
  static const int which_x2[12] = {0, 0, 0, 1, 1, 1, 1, 0, 0, 1, 1, 0};
  static const int which_y2[12] = {0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 1, 1};
  static const int which_z2[12] = {1, 1, 0, 0, 1, 1, 0, 0, 0, 1, 1, 0};
  static const int edge1_vt[12] = {0, 1, 3, 2, 6, 7, 5, 4, 0, 1, 3, 4};
  static const int edge2_vt[12] = {1, 3, 2, 6, 7, 5, 4, 0, 2, 5, 7, 2};
  V3 n = f.normal;
  double d = f.distance_to_origin;
  double vu_[3], vv_[3];
  V3 u = vert[1] - vert[0];
  double r_u = 1 / sqrt(len3_squared(u));
  u = u * r_u;
  V3 v = cross(n, u);
  for (int j = 0; j < 3; j++) { vu_[j] = dot(vert[j], u); vv_[j] = dot(vert[j], v); }
  V3 bb = llf, ba = llf;
  double d_box[8];
  d_box[0] = dot(bb, n);
  for (int i = 0; i < 12; i++) {
    double a1, a2;
    if (i < 7) {
      ba = bb;
      bb.x = which_x2[i] ? urb.x : llf.x;
      bb.y = which_y2[i] ? urb.y : llf.y;
      bb.z = which_z2[i] ? urb.z : llf.z;
      a2 = d_box[edge2_vt[i]] = dot(bb, n);
      a1 = d_box[edge1_vt[i]];
      if ((a1 - d < 0 && a2 - d < 0) || (a1 - d > 0 && a2 - d > 0)) continue;
    } else {
      a1 = d_box[edge1_vt[i]];
      a2 = d_box[edge2_vt[i]];
      if ((a1 - d < 0 && a2 - d < 0) || (a1 - d > 0 && a2 - d > 0)) continue;
      ba.x = which_x1[i] ? urb.x : llf.x;
      ba.y = which_y1[i] ? urb.y : llf.y;
      ba.z = which_z1[i] ? urb.z : llf.z;
      bb.x = which_x2[i] ? urb.x : llf.x;
      bb.y = which_y2[i] ? urb.y : llf.y;
      bb.z = which_z2[i] ? urb.z : llf.z;
    }
    double r = (d - a1) / (a2 - a1);
    V3 c = ba + (bb - ba) * r;
    double cu = dot(c, u), cv = dot(c, v);
    int temp = 0;
    for (int j = 0; j < 3; j++) {
      int k = j == 0 ? 2 : j - 1;
      if ((vu_[k] < cu && cu <= vu_[j]) || (vu_[k] >= cu && cu > vu_[j])) {
        double rr = (cu - vu_[k]) / (vu_[j] - vu_[k]);
        if ((vv_[k] + rr * (vv_[j] - vv_[k])) > cv) temp++;
      }
    }
    if (temp & 1) return 8 + i;
  }
  return 0;
}

// GeometryUtils::wall_subparts_collision_test (geometry_utils.inl:110-207) +
// Partition::finalize_walls (partition.cpp:91-118)
static void finalize_walls(World& w) {
  size_t ns3 = (size_t)w.n_sp * w.n_sp * w.n_sp;
  w.walls_per_subpart.assign(ns3, {});
  V3 origin = {w.cfg.origin[0], w.cfg.origin[1], w.cfg.origin[2]};
  for (uint32_t wi = 0; wi < w.walls.size(); wi++) {
    const Wall& f = w.walls[wi];
    V3 p[3] = {w.verts[f.vi[0]], w.verts[f.vi[1]], w.verts[f.vi[2]]};
    V3 llf = p[0], urb = p[0];
    for (int k = 1; k < 3; k++) {
      if (p[k].x < llf.x) llf.x = p[k].x; else if (p[k].x > urb.x) urb.x = p[k].x;
      if (p[k].y < llf.y) llf.y = p[k].y; else if (p[k].y > urb.y) urb.y = p[k].y;
      if (p[k].z < llf.z) llf.z = p[k].z; else if (p[k].z > urb.z) urb.z = p[k].z;
    }
    double leeway = 1;
    if (llf.x < -leeway) leeway = -llf.x;
    if (llf.y < -leeway) leeway = -llf.y;
    if (llf.z < -leeway) leeway = -llf.z;
    if (urb.x > leeway) leeway = urb.x;
    if (urb.y > leeway) leeway = urb.y;
    if (urb.z > leeway) leeway = urb.z;
    leeway = POS_EPS + leeway * POS_EPS;
    if (w.cfg.use_expanded_list) leeway += w.cfg.rxn_radius_3d;
    V3 lw = {leeway, leeway, leeway};
    llf = llf - lw; urb = urb + lw;
    int mn[3], mx[3];
    w.subpart_3d(llf, mn); w.subpart_3d(urb, mx);
    for (int x = mn[0]; x <= mx[0]; x++)
      for (int y = mn[1]; y <= mx[1]; y++)
        for (int z = mn[2]; z <= mx[2]; z++) {
          if (!w.idx_in_range(x) || !w.idx_in_range(y) || !w.idx_in_range(z)) continue;
          int idx[3] = {x, y, z};
          uint32_t s = w.subpart_from_3d(idx);
          V3 sl = origin + V3{(double)x, (double)y, (double)z} * w.sp_len;  // get_subpart_llf_point
          V3 su = sl + V3{w.sp_len, w.sp_len, w.sp_len};
          sl = sl - lw; su = su + lw;
          if (wall_in_box(w, f, sl, su) != 0) w.walls_per_subpart[s].push_back(wi);
        }
  }
}

// ---- reactant lists (sequential mode) -------------------------------------------------------
static inline uint64_t list_key(uint32_t species, uint32_t subpart) { return ((uint64_t)species << 32) | subpart; }
static void list_insert(World& w, Mol& m) {
  if (m.wall != MCX_NONE) return;  // only volume molecules are reactants of the per-subpartition lists
  auto& v = w.lists[list_key(m.species, m.subpart)];
  m.list_slot = (uint32_t)v.size(); m.reg_subpart = m.subpart;
  v.push_back(m.id);
}
static void list_erase(World& w, Mol& m) {
  if (m.wall != MCX_NONE) return;
  auto& v = w.lists[list_key(m.species, m.reg_subpart)];
  uint32_t last = v.back();
  v[m.list_slot] = last;
  w.mols[w.id_to_index[last]].list_slot = m.list_slot;
  v.pop_back();
}

// ---- counted volumes of intersecting objects (include/mcx.h: mcx_set_counted_volume_objects) -------------------------
static inline uint32_t cv_lookup(const World& w, uint32_t mask) {
  for (size_t k = 0; k < w.cv_mask.size(); k++) if (w.cv_mask[k] == mask) return (uint32_t)k;
  return MCX_NONE;
}
static inline bool cv_uses_xor(const World& w, uint32_t wall) {
  return !w.cv_mask.empty() && w.walls[wall].object < 32 && ((w.cv_xor >> w.walls[wall].object) & 1u);
}
// update_counted_volume_id_when_crossing_wall (collision_utils.inl:1637-1694): the volume behind a wall hit on its front,
// in front of one hit on its back; walls of intersecting objects toggle their object in the molecule's set instead
static inline uint32_t cv_cross(World& w, uint32_t cvi, uint32_t wall, bool hit_front) {
  const Wall& f = w.walls[wall];
  if (!cv_uses_xor(w, wall)) return hit_front ? f.cv_back : f.cv_front;
  const uint32_t k = cv_lookup(w, w.cv_mask[cvi] ^ (1u << f.object));
  if (k == MCX_NONE) { w.err = "counted volume set not registered (mcx_set_counted_volume_objects)"; return cvi; }
  return k;
}

// ---- model table helpers ------------------------------------------------------------------
static void build_lookups(World& w) {
  size_t ns = w.species.size();
  w.bimol.assign(ns * ns, -1);
  w.unimol.assign(ns, -1);
  w.can_vol_react.assign(ns, 0);
  w.volsurf.assign(ns * ns, -1);
  w.can_vol_surf.assign(ns, 0);
  w.surfsurf.assign(ns * ns, -1);
  w.can_surf_surf.assign(ns, 0);
  for (size_t c = 0; c < w.classes.size(); c++) {
    const mcx_rxn_class& rc = w.classes[c];
    if (rc.kind == MCX_RXN_BIMOL_SURFSURF) {
      if (rc.reactants[0] < ns && rc.reactants[1] < ns) {
        w.surfsurf[rc.reactants[0] * ns + rc.reactants[1]] = (int)c;
        w.surfsurf[rc.reactants[1] * ns + rc.reactants[0]] = (int)c;
        w.can_surf_surf[rc.reactants[0]] = w.can_surf_surf[rc.reactants[1]] = 1;
      }
      continue;
    }
    if (rc.kind == MCX_RXN_BIMOL_VOLSURF) {
      if (rc.reactants[0] < ns && rc.reactants[1] < ns) {
        w.volsurf[rc.reactants[0] * ns + rc.reactants[1]] = (int)c;
        w.can_vol_surf[rc.reactants[0]] = 1;
      }
    } else if (rc.kind == MCX_RXN_BIMOL_VOLVOL) {
      w.bimol[rc.reactants[0] * ns + rc.reactants[1]] = (int)c;
      w.bimol[rc.reactants[1] * ns + rc.reactants[0]] = (int)c;
    } else if (rc.kind == MCX_RXN_UNIMOL) {
      w.unimol[rc.reactants[0]] = (int)c;
    }
  }
  // Species::can_vol_react(): has a vol-vol reaction and may initiate it
  for (size_t a = 0; a < ns; a++) {
    bool any = false;
    for (size_t b = 0; b < ns; b++) any |= w.bimol[a * ns + b] >= 0;
    w.can_vol_react[a] = any && !(w.species[a].flags & MCX_SP_CANT_INITIATE);
  }
  uint32_t max_rule = 0;
  for (auto& p : w.pathways) max_rule = std::max(max_rule, p.rxn_rule_id + 1);
  w.rxn_count.assign(max_rule, 0);
  w.rxn_count_cv.assign((size_t)max_rule * w.n_cv, 0);
  w.species_count.assign(ns, 0);
}

// surface class lookup: first matching rule in the reference's order
// (species-specific, then ALL_MOLECULES, then ALL_VOLUME_MOLECULES; rxn_utils.inl:182-244)
static int surf_action(const World& w, uint32_t species, uint32_t surf_class, int side /*COLL_WALL_*/, int* rxn_class = nullptr) {
  if (surf_class == MCX_NONE) return MCX_SURF_REFLECTIVE;
  int orient = side == COLL_WALL_FRONT ? 1 : -1;  // FRONT -> ORIENTATION_UP (diffuse_react_event.cpp:1009)
  const uint32_t order[3] = {species, MCX_ALL_MOLECULES, MCX_ALL_VOLUME_MOLECULES};
  for (int o = 0; o < 3; o++)
    for (const auto& r : w.surf_rules)
      if (r.species == order[o] && r.surf_class == surf_class && (r.orientation == 0 || r.orientation == orient)) {
        if (rxn_class) *rxn_class = (int)r.rxn_class;
        return (int)r.type;
      }
  return MCX_SURF_REFLECTIVE;
}
// reflect_absorb_check_wall (diffusion_utils.inl:598-628): what the border of the wall's reactive region does to a surface
// molecule of this species and orientation — MCX_SURF_REFLECTIVE, MCX_SURF_ABSORPTIVE (absorptive region border), or
// MCX_SURF_TRANSPARENT (nothing: it passes).  Lookup order of find_mol_reactions_with_surf_classes: the species, ALL_MOLECULES,
// ALL_SURFACE_MOLECULES.
static int border_action(const World& w, uint32_t species, uint32_t surf_class, int orient) {
  if (surf_class == MCX_NONE) return MCX_SURF_TRANSPARENT;
  const uint32_t order[3] = {species, MCX_ALL_MOLECULES, MCX_ALL_SURFACE_MOLECULES};
  for (int o = 0; o < 3; o++)
    for (const auto& r : w.surf_rules)
      if (r.species == order[o] && r.surf_class == surf_class && (r.orientation == 0 || r.orientation == orient) &&
          (r.type == MCX_SURF_REFLECTIVE || r.type == MCX_SURF_ABSORPTIVE))
        return (int)r.type;
  return MCX_SURF_TRANSPARENT;
}

// exact_disk ignores a wall the moving molecule can travel through (exact_disk_utils.inl:957-975):
// trigger_intersect with ORIENTATION_NONE matches the classes that do not depend on orientation
// (rxn_utils.inl:149-158); the wall is ignored when there is at least one and all of them are transparent
static bool exd_passes_through(const World& w, uint32_t species, uint32_t surf_class) {
  if (surf_class == MCX_NONE) return false;
  bool any = false;
  for (const auto& r : w.surf_rules) {
    if (r.surf_class != surf_class || r.orientation != 0) continue;
    if (r.species != species && r.species != MCX_ALL_MOLECULES && r.species != MCX_ALL_VOLUME_MOLECULES) continue;
    if (r.type != MCX_SURF_TRANSPARENT) return false;
    any = true;
  }
  return any;
}

// binary_search_double, src4/rxn_utils.inl:301-320 (== src/react_cond.c:80-97)
// mult: RxnClass::get_pathway_index_for_probability's local probability factor (1 except between surface molecules)
static int pathway_for_probability(const World& w, const mcx_rxn_class& rc, double match, double mult = 1) {
  int min_idx = 0, max_idx = (int)rc.n_pathways - 1;
  const mcx_pathway* A = &w.pathways[rc.first_pathway];
  while (max_idx - min_idx > 1) {
    int mid = (max_idx + min_idx) / 2;
    if (match > A[mid].cum_prob * mult) min_idx = mid; else max_idx = mid;
  }
  if (match > A[min_idx].cum_prob * mult) return max_idx;
  return min_idx;
}

// ---- surface reactions: orientation algebra and product placement -----------------------------------
// RxnUtils::trigger_bimolecular, src4/rxn_utils.inl:58-84: does the (volume, surface) pair match the class?
static bool orientations_match(const mcx_rxn_class& rc, int orientA, int orientB) {
  int geomA = rc.reactant_orientation[0], geomB = rc.reactant_orientation[1];
  if (geomA == 0 || geomB == 0 || (geomA + geomB) * (geomA - geomB) != 0) return true;
  return orientA != 0 && orientA * orientB * geomA * geomB > 0;
}
// one random bit per product whose rule orientation is NONE, drawn from the main stream right after the pathway
// is chosen (outcome_products_random, diffuse_react_event.cpp:2618-2627); only when a surface is involved
// With pw.kept_info the draws follow the order of the rule's products, kept reactants included (bit 4 + r belongs to
// kept reactant r); without it (older tables) only the new products draw.
static inline int kept_code(const mcx_pathway& pw, int r) {  // product-side orientation of kept reactant r: 0 none, +1, -1
  const uint32_t c = (pw.kept_info >> (24 + 2 * r)) & 3u;
  return c == 1 ? 1 : (c == 2 ? -1 : 0);
}
template <class RS>
static uint32_t draw_orientation_bits(const mcx_pathway& pw, RS& rs) {
  uint32_t bits = 0;
  if (!(pw.kept_info & MCX_KEPT_VALID)) {
    for (uint32_t k = 0; k < pw.n_products; k++)
      if (pw.product_orientation[k] == 0 && (rs.next() & 1)) bits |= 1u << k;
    return bits;
  }
  for (int q = 0; q < 6; q++) {
    const uint32_t nib = (pw.kept_info >> (4 * q)) & 0xFu;
    if (nib == MCX_KEPT_ORDER_END) break;
    if (nib >= MCX_KEPT_ORDER_REACTANT) {
      const int r = (int)(nib & 1u);
      if (kept_code(pw, r) == 0 && (rs.next() & 1)) bits |= 1u << (4 + r);
    } else if (nib < pw.n_products && pw.product_orientation[nib] == 0 && (rs.next() & 1)) bits |= 1u << nib;
  }
  return bits;
}
// Product-side orientation of kept reactant r (outcome_products_random :2618-2652): the rule's mark, a drawn bit when
// the rule has none, flipped when the surface reactant of a volume-surface reaction lies the other way round than
// the rule states.  0: the table does not say (kept reactants stay as they are).
static inline int kept_orientation(const mcx_rxn_class& c, const mcx_pathway& pw, int r, uint32_t orient_bits, int surf_orient) {
  if (!(pw.kept_info & MCX_KEPT_VALID)) return 0;
  int o = kept_code(pw, r);
  if (o == 0) return ((orient_bits >> (4 + r)) & 1u) ? 1 : -1;
  if (c.kind == MCX_RXN_BIMOL_VOLSURF) {
    const int gB = c.reactant_orientation[1];
    if (gB != 0 && surf_orient != gB) o = -o;
  }
  return o;
}
struct ProductSpec {
  uint32_t species; V3 pos;
  uint32_t wall = MCX_NONE, tile = MCX_NONE; int orient = 0; double u = 0, v = 0;
  uint32_t created_wall = MCX_NONE, created_tile = MCX_NONE;
  uint32_t cvi = 0;
  bool cvi_pending = false;   // on a wall of an intersecting counted object without a volume reactant to go by: ray cast later
};
// Where and how product k of a pathway is created (outcome_products_random :2446-2933, the cases of SURVEY A.2):
//  * no surface reactant: volume product at the event position;
//  * surface product: takes the surface reactant's tile and uv (find_surf_product_positions :2145-2155);
//  * volume product of a surface reaction: event position bumped 2*16*EPS off the wall to the side its
//    orientation names (update_vol_mol_after_rxn_with_surf_mol :2291-2320; the wall test of tiny_diffuse_3D is
//    omitted: another wall within 3.2e-11 length units of the position), remembered for the rebinding guard.
// cvi: counted volume of the initiator at the event; a volume product of a surface reaction takes the volume on the
// side of the wall it is released to (outcome_products_random :2757-2760)
// init_side: the side of the wall the volume initiator (whose counted volume is cvi) was on, 0 = there is none
static ProductSpec product_spec(World& w, const mcx_rxn_class& c, const mcx_pathway& pw, uint32_t k, V3 pos,
                                uint32_t orient_bits, const Mol* surf, uint32_t cvi, int init_side = 0) {
  ProductSpec ps;
  ps.species = pw.products[k]; ps.pos = pos; ps.cvi = cvi;
  if (!surf) return ps;
  int o = pw.product_orientation[k];
  if (o == 0) o = ((orient_bits >> k) & 1) ? 1 : -1;
  else if (c.kind == MCX_RXN_BIMOL_VOLSURF) {  // :2634-2652: flip when the reactant's orientation differs from the rule's
    int gB = c.reactant_orientation[1];
    if (gB != 0 && surf->orient != gB) o = -o;
  }
  const Wall& f = w.walls[surf->wall];
  if (w.is_surf(ps.species)) {
    ps.wall = surf->wall; ps.tile = surf->tile; ps.u = surf->u; ps.v = surf->v; ps.orient = o;
    ps.pos = uv2xyz(w, f, ps.u, ps.v);
    ps.cvi = 0;
  } else {
    ps.cvi = o > 0 ? f.cv_front : f.cv_back;
    if (cv_uses_xor(w, surf->wall)) {
      if (init_side != 0) ps.cvi = (o > 0) == (init_side > 0) ? cvi : cv_cross(w, cvi, surf->wall, init_side > 0);
      else ps.cvi_pending = true;
    }
    double bump = (o > 0) ? 16 * POS_EPS : -16 * POS_EPS;
    V3 d = {(2 * bump) * f.normal.x, (2 * bump) * f.normal.y, (2 * bump) * f.normal.z};
    ps.pos = pos + d;
    ps.created_wall = surf->wall; ps.created_tile = surf->tile;
  }
  return ps;
}

struct SurfSite { uint32_t wall, tile; double u, v; int orient; uint32_t species; };
static inline SurfSite site_of(const Mol& m) { return SurfSite{m.wall, m.tile, m.u, m.v, m.orient, m.species}; }

// ---- products on vacant neighbour tiles: the general branch of find_surf_product_positions (:2060-2100, 2155-2285) ----
// A pathway that creates more surface products than it consumes surface reactants puts the extra ones on vacant tiles
// around the surface reactant (find_neighbor_tiles with create_grid_flag, every wall counts), at a random point of the
// tile (grid2uv_random, :2852-2855).  The reference's bookkeeping is restated literally, quirks included: positions are
// assigned per entry of the rule's product list (kept reactants and volume products draw a tile too and waste it), and
// the c-th CREATED surface product takes the position of the c-th ENTRY (:2815-2818).
// Deviation: a unimolecular split into two surface products leaves the product that stays on the
// reactant's tile at the reactant's uv; the reference moves it next to the other product's tile (find_closest_position).
// GridUtils::grid2uv_random, src4/grid_utils.inl:256-286: a random point of a tile
static void grid2uv_random(const Wall& f, const Grid& g, uint32_t tile_index, WordSource& rs, double& u, double& v) {
  int root = (int)(sqrt((double)tile_index));
  int rootrem = (int)tile_index - root * root;
  int k = g.n_axis - root - 1;
  int j = rootrem / 2;
  int i = rootrem - 2 * j;
  double over_n = 1 / (double)(g.n_axis);
  double u_ran = rs.dbl();
  double v_ran = 1 - sqrt(rs.dbl());
  u = ((double)(j + i) + (1 - 2 * i) * (1 - v_ran) * u_ran) * over_n * f.uv_vert1_u + ((double)(k + i) + (1 - 2 * i) * v_ran) * over_n * f.uv_vert2_u;
  v = ((double)(k + i) + (1 - 2 * i) * v_ran) * over_n * f.uv_vert2_v;
}
struct Placement {
  bool general = false;
  int n = 0;                       // created surface products
  uint32_t wall[MCX_MAX_PRODUCTS], tile[MCX_MAX_PRODUCTS];
  double u[MCX_MAX_PRODUCTS], v[MCX_MAX_PRODUCTS];
  uint32_t vacant_mask = 0;        // bit c: created surface product c sits on a tile that was vacant (claimed by the event)
};
struct RuleEntry { bool kept; uint32_t idx; bool surf; };  // one entry of the rule's product list: new product idx / kept reactant idx
static int rule_entries(const World& w, const mcx_rxn_class& c, const mcx_pathway& pw, RuleEntry out[6]) {
  int n = 0;
  if (!(pw.kept_info & MCX_KEPT_VALID)) {  // older tables: the new products in their order, kept reactants behind them
    for (uint32_t k = 0; k < pw.n_products; k++) out[n++] = RuleEntry{false, k, w.is_surf(pw.products[k])};
    for (uint32_t r = 0; r < 2 && n < 6; r++)
      if ((pw.keep_reactant_mask >> r) & 1u) out[n++] = RuleEntry{true, r, c.reactants[r] < w.species.size() && w.is_surf(c.reactants[r])};
    return n;
  }
  for (int q = 0; q < 6; q++) {
    const uint32_t nib = (pw.kept_info >> (4 * q)) & 0xFu;
    if (nib == MCX_KEPT_ORDER_END) break;
    if (nib >= MCX_KEPT_ORDER_REACTANT) { const uint32_t r = nib & 1u; out[n++] = RuleEntry{true, r, c.reactants[r] < w.species.size() && w.is_surf(c.reactants[r])}; }
    else if (nib < pw.n_products) out[n++] = RuleEntry{false, nib, w.is_surf(pw.products[nib])};
  }
  return n;
}
static inline int n_surface_reactants(const World& w, const mcx_rxn_class& c, const mcx_pathway& pw, bool kept) {
  int n = 0;
  const int nr = c.kind == MCX_RXN_UNIMOL ? 1 : 2;
  if (c.kind == MCX_RXN_BIMOL_VOLWALL) return 0;
  for (int r = 0; r < nr; r++)
    if (c.reactants[r] < w.species.size() && w.is_surf(c.reactants[r]) && (((pw.keep_reactant_mask >> r) & 1u) != 0) == kept) n++;
  return n;
}
static inline bool pathway_is_general(const World& w, const mcx_rxn_class& c, const mcx_pathway& pw) {
  if (c.kind != MCX_RXN_UNIMOL && c.kind != MCX_RXN_BIMOL_VOLSURF && c.kind != MCX_RXN_BIMOL_SURFSURF) return false;
  if (n_surface_reactants(w, c, pw, false) + n_surface_reactants(w, c, pw, true) == 0) return false;
  int created = 0;
  for (uint32_t k = 0; k < pw.n_products; k++) created += w.is_surf(pw.products[k]) ? 1 : 0;
  return created > n_surface_reactants(w, c, pw, false);
}
static const char* general_pathway_problem(const World& w, const mcx_rxn_class& c, const mcx_pathway& pw) {
  if (!(pw.kept_info & MCX_KEPT_VALID)) return "a pathway with surface products on vacant neighbour tiles needs kept_info (the order of the rule's products)";
  if (n_surface_reactants(w, c, pw, false) > 0 && n_surface_reactants(w, c, pw, true) > 0)
    return "a pathway with surface products on vacant neighbour tiles that keeps one surface reactant and consumes another is not supported";
  return nullptr;
}
// recycled: the sites of the consumed surface reactants in the order of the rule's reactants; vacant(wall, tile) tells
// whether a tile can be taken.  Returns false when the reaction is blocked (RX_BLOCKED).  Draws in the reference's order:
// tile assignment, orientations, random points.
template <class RS, class VacantFn>
static bool place_general(const World& w, const mcx_rxn_class& c, const mcx_pathway& pw, uint32_t reac_wall, uint32_t reac_tile,
                          const SurfSite* recycled, int n_recycled, RS& rs, VacantFn is_vacant, Placement& pl, uint32_t& orient_bits,
                          int* entry_kind = nullptr, WallTile* entry_pos = nullptr) {   // the last two: per-entry result for the pin test
  pl.general = true; pl.n = 0; pl.vacant_mask = 0;
  RuleEntry ent[6];
  const int n_ent = rule_entries(w, c, pw, ent);
  const int n_reactants = c.kind == MCX_RXN_UNIMOL ? 1 : 2;
  int needed = 0;
  for (uint32_t k = 0; k < pw.n_products; k++) needed += w.is_surf(pw.products[k]) ? 1 : 0;
  // vacant tiles around the surface reactant, from the back of the reference's list (:2090-2098)
  TileNeighbors nbt;
  w.assume_all_grids = true;
  find_neighbor_tiles(w, reac_wall, reac_tile, nbt);
  w.assume_all_grids = false;
  std::vector<WallTile> vacant;
  for (int i = (int)nbt.size() - 1; i >= 0; i--) if (is_vacant(nbt[i].first, nbt[i].second)) vacant.push_back(nbt[i]);
  if ((int)vacant.size() + n_recycled < needed) return false;  // :2101-2105
  int assigned[6];   // -1 nothing, 0/1 recycled site, 2 + j vacant tile j
  for (int e = 0; e < 6; e++) assigned[e] = -1;
  const int to_recycle = std::min(n_ent, n_recycled);
  int next_available = 0, guard = 0;
  const uint32_t num_players = (uint32_t)(n_ent + n_reactants);
  while (next_available < to_recycle && ++guard < 100000) {  // :2159-2191
    const uint32_t rnd = rs.next() % num_players;
    if (rnd < (uint32_t)n_reactants) continue;
    const int e = (int)rnd - n_reactants;
    if (!ent[e].surf) continue;
    if (assigned[e] >= 0) continue;
    assigned[e] = next_available++;
  }
  std::vector<uint8_t> used(vacant.size(), 0);
  for (int e = 0; e < n_ent; e++) {  // :2232-2283: every entry without a position draws a vacant tile
    if (assigned[e] >= 0) continue;
    int attempts = 0; bool found = false;
    while (!found && attempts < 10) {  // SURFACE_DIFFUSION_RETRIES
      const uint32_t rnd = rs.next() % (uint32_t)vacant.size();
      if (used[rnd]) { attempts++; continue; }
      assigned[e] = 2 + (int)rnd; used[rnd] = 1; found = true;
    }
    if (attempts >= 10) return false;
  }
  if (entry_kind)
    for (int e = 0; e < n_ent; e++) {
      entry_kind[e] = assigned[e] < 0 ? 0 : (assigned[e] >= 2 ? 2 : 1);   // 0 nothing, 1 recycled, 2 vacant
      entry_pos[e] = assigned[e] >= 2 ? vacant[assigned[e] - 2]
                                      : (assigned[e] >= 0 ? WallTile(recycled[assigned[e]].wall, recycled[assigned[e]].tile) : WallTile(MCX_NONE, MCX_NONE));
    }
  orient_bits = draw_orientation_bits(pw, rs);
  int cnt = 0;   // current_surf_product_position_index
  for (int e = 0; e < n_ent; e++) {
    if (ent[e].kept || !ent[e].surf) continue;
    const int a = assigned[cnt];
    if (a >= 2) {
      const WallTile t = vacant[a - 2];
      pl.wall[cnt] = t.first; pl.tile[cnt] = t.second; pl.vacant_mask |= 1u << cnt;
      grid2uv_random(w.walls[t.first], w.grids[t.first], t.second, rs, pl.u[cnt], pl.v[cnt]);   // :2852-2855
    } else {
      const SurfSite& r = recycled[a < 0 ? 0 : a];
      pl.wall[cnt] = r.wall; pl.tile[cnt] = r.tile; pl.u[cnt] = r.u; pl.v[cnt] = r.v;
    }
    cnt++;
  }
  pl.n = cnt;
  return true;
}
static inline void put_placement(Outcome& o, const Placement& pl) {
  o.pl_general = pl.general; o.pl_n = pl.n; o.pl_vacant = pl.vacant_mask;
  for (int k = 0; k < pl.n; k++) { o.pl_wall[k] = pl.wall[k]; o.pl_tile[k] = pl.tile[k]; o.pl_u[k] = pl.u[k]; o.pl_v[k] = pl.v[k]; }
}
static inline Placement get_placement(const Outcome& o) {
  Placement pl; pl.general = o.pl_general; pl.n = o.pl_n; pl.vacant_mask = o.pl_vacant;
  for (int k = 0; k < o.pl_n; k++) { pl.wall[k] = o.pl_wall[k]; pl.tile[k] = o.pl_tile[k]; pl.u[k] = o.pl_u[k]; pl.v[k] = o.pl_v[k]; }
  return pl;
}
// a created surface product of a general pathway goes where the placement says (c counts the created surface products)
static inline void apply_placement(const World& w, const Placement* pl, int& c, ProductSpec& ps) {
  if (!pl || !pl->general || !w.is_surf(ps.species)) return;
  ps.wall = pl->wall[c]; ps.tile = pl->tile[c]; ps.u = pl->u[c]; ps.v = pl->v[c];
  ps.pos = uv2xyz(w, w.walls[ps.wall], ps.u, ps.v);
  c++;
}

// ---- surface-surface reactions (outcome_products_random :2446-2933 for a SURFMOL_SURFMOL collision) -----------------
// Which pathways of a surface-surface class can be placed: every new surface product finds a tile a consumed reactant
// frees (find_surf_product_positions :1993-2288 without its search for vacant neighbour tiles), and the reference's
// assignment loop terminates (it hands out min(products, freed tiles) tiles to surface products only, :2155-2191)
static const char* surfsurf_pathway_problem(const World& w, const mcx_pathway& pw) {
  const int keep0 = pw.keep_reactant_mask & 1, keep1 = (pw.keep_reactant_mask >> 1) & 1;
  int needed = 0;
  for (uint32_t k = 0; k < pw.n_products; k++) needed += w.is_surf(pw.products[k]) ? 1 : 0;
  const int freed = (keep0 ? 0 : 1) + (keep1 ? 0 : 1), actual = (int)pw.n_products + keep0 + keep1;
  if (needed > freed) return "a surface-surface pathway with more new surface products than consumed reactants needs vacant neighbour tiles (find_surf_product_positions' general branch is not built)";
  const int to_recycle = std::min(actual, freed);
  if (needed == 2 && to_recycle == 2 && actual > 2)
    return "a surface-surface pathway with two surface products on the two freed tiles and a volume product: the reference draws a vacant tile for the volume entry from an empty list (diffuse_react_event.cpp:2232-2251, a division by zero)";
  if (needed != 0 && !(needed == 1 && to_recycle == 1) && needed < to_recycle)
    return "a surface-surface pathway that frees more tiles than it has surface products, next to a volume product: the reference's tile assignment (diffuse_react_event.cpp:2155-2191) does not terminate";
  return nullptr;
}
// find_surf_product_positions (:1993-2288) over recycled tiles: SURFSURF_SWAP = the first surface product takes the second
// freed tile (freed tiles in the order of the rule's reactants) and the other one, if there is one, the first.  Draws
// from the stream like the reference (rng_uint % players, :2161)
static const uint32_t SURFSURF_SWAP = 64u;  // bit 6 of the orientation bits (bits 0-3: products, 4-5: kept reactants)
template <class RS>
static uint32_t surfsurf_position_bits(const World& w, const mcx_rxn_class& c, const mcx_pathway& pw, bool init_is_r0, RS& rs) {
  const int keep0 = pw.keep_reactant_mask & 1, keep1 = (pw.keep_reactant_mask >> 1) & 1;
  int needed = 0; uint32_t first_surf = MCX_NONE;
  for (uint32_t k = 0; k < pw.n_products; k++) if (w.is_surf(pw.products[k])) { needed++; if (first_surf == MCX_NONE) first_surf = k; }
  if (needed == 0) return 0;
  const int freed = (keep0 ? 0 : 1) + (keep1 ? 0 : 1), actual = (int)pw.n_products + keep0 + keep1;
  const int to_recycle = std::min(actual, freed);
  if (needed == 1 && to_recycle == 1) {  // :2140-2154: the initiator's tile when it is consumed, else the one freed tile
    const int ri = init_is_r0 ? 0 : 1;
    const bool init_consumed = ri == 0 ? !keep0 : !keep1;
    const int idx = (init_consumed && ri == 1 && !keep0) ? 1 : 0;
    return idx ? SURFSURF_SWAP : 0u;
  }
  uint32_t bits = 0, assigned = 0;
  int next_available = 0;
  const uint32_t num_players = (uint32_t)actual + 2;
  int guard = 0;
  while (next_available < to_recycle && ++guard < 100000) {  // :2159-2191
    const uint32_t rnd = rs.next() % num_players;
    if (rnd < 2) continue;
    const uint32_t k = rnd - 2;
    if (k >= pw.n_products || !w.is_surf(pw.products[k])) continue;
    if ((assigned >> k) & 1u) continue;
    assigned |= 1u << k;
    if ((next_available == 1) == (k == first_surf)) bits |= SURFSURF_SWAP;  // each of the two assignments says the same
    next_available++;
  }
  return bits;
}
// the factor of :2640-2652: a product's rule orientation flips once for every surface reactant that lies the other way
// round than the rule states
static inline int surfsurf_match(const mcx_rxn_class& c, const SurfSite& r0, const SurfSite& r1) {
  int m = 1;
  if (c.reactant_orientation[0] != 0 && r0.orient != c.reactant_orientation[0]) m = -m;
  if (c.reactant_orientation[1] != 0 && r1.orient != c.reactant_orientation[1]) m = -m;
  return m;
}
// product-side orientation of kept reactant r of a surface-surface pathway (:2618-2652, 2689-2716); 0: the table does not say
static inline int surfsurf_kept_orientation(const mcx_rxn_class& c, const mcx_pathway& pw, int r, uint32_t bits, const SurfSite& r0,
                                            const SurfSite& r1) {
  if (!(pw.kept_info & MCX_KEPT_VALID)) return 0;
  const int o = kept_code(pw, r);
  if (o == 0) return ((bits >> (4 + r)) & 1u) ? 1 : -1;
  return o * surfsurf_match(c, r0, r1);
}
// The products of the pathway: surface products on the freed tiles at the uv of the reactant that left (REACA_UV /
// REACB_UV, :2826-2846); volume products at the position of the rule's first reactant (collision without a position,
// :2745-2750), bumped off the INITIATOR's wall to the side their orientation names and remembered with its tile (:2757-2762)
static void surfsurf_products(World& w, const mcx_rxn_class& c, const mcx_pathway& pw, const SurfSite& init, const SurfSite& partner,
                              uint32_t bits, std::vector<ProductSpec>& out) {
  const bool init_is_r0 = init.species == c.reactants[0];
  const SurfSite& r0 = init_is_r0 ? init : partner;
  const SurfSite& r1 = init_is_r0 ? partner : init;
  const SurfSite* freed[2]; int n_freed = 0;
  if (!(pw.keep_reactant_mask & 1u)) freed[n_freed++] = &r0;
  if (!(pw.keep_reactant_mask & 2u)) freed[n_freed++] = &r1;
  const int match = surfsurf_match(c, r0, r1);
  uint32_t first_surf = MCX_NONE;
  for (uint32_t k = 0; k < pw.n_products && first_surf == MCX_NONE; k++) if (w.is_surf(pw.products[k])) first_surf = k;
  for (uint32_t k = 0; k < pw.n_products; k++) {
    ProductSpec ps;
    ps.species = pw.products[k];
    int o = pw.product_orientation[k];
    if (o == 0) o = ((bits >> k) & 1u) ? 1 : -1; else o *= match;
    if (w.is_surf(ps.species)) {
      const bool swap = (bits & SURFSURF_SWAP) != 0;
      const int which = std::min((k == first_surf) == swap ? 1 : 0, n_freed - 1);
      const SurfSite& t = n_freed ? *freed[which < 0 ? 0 : which] : init;  // no freed tile: a general pathway places it (apply_placement)
      ps.wall = t.wall; ps.tile = t.tile; ps.u = t.u; ps.v = t.v; ps.orient = o;
      ps.pos = uv2xyz(w, w.walls[t.wall], t.u, t.v);
      ps.cvi = 0;
    } else {
      const Wall& f = w.walls[init.wall];
      ps.cvi = o > 0 ? f.cv_front : f.cv_back;
      if (cv_uses_xor(w, init.wall)) ps.cvi_pending = true;
      const double bump = (o > 0) ? 16 * POS_EPS : -16 * POS_EPS;
      const V3 from = uv2xyz(w, w.walls[r0.wall], r0.u, r0.v);
      ps.pos = from + V3{(2 * bump) * f.normal.x, (2 * bump) * f.normal.y, (2 * bump) * f.normal.z};
      ps.created_wall = init.wall; ps.created_tile = init.tile;
    }
    out.push_back(ps);
  }
}

// Volume product k of a reaction with a reactive surface (outcome_products_random :2739-2806 with a wall collision): at the
// hit point, bumped off the wall to the side its orientation names, counted volume of that side, remembered with the
// tile under the hit point (xyz2uv + uv2grid_tile_index, :2768-2776)
static ProductSpec wall_product_spec(World& w, const mcx_pathway& pw, uint32_t k, V3 pos, uint32_t orient_bits, uint32_t wall,
                                     uint32_t cvi_init = 0, int init_side = 0) {
  ProductSpec ps;
  ps.species = pw.products[k];
  int o = pw.product_orientation[k];
  if (o == 0) o = ((orient_bits >> k) & 1) ? 1 : -1;
  const Wall& f = w.walls[wall];
  const Grid& g = w.grids[wall];
  ps.cvi = o > 0 ? f.cv_front : f.cv_back;
  if (cv_uses_xor(w, wall) && init_side != 0) ps.cvi = (o > 0) == (init_side > 0) ? cvi_init : cv_cross(w, cvi_init, wall, init_side > 0);
  const double bump = (o > 0) ? 16 * POS_EPS : -16 * POS_EPS;
  ps.pos = pos + V3{(2 * bump) * f.normal.x, (2 * bump) * f.normal.y, (2 * bump) * f.normal.z};
  const double hu = pos.x * f.unit_u.x + pos.y * f.unit_u.y + pos.z * f.unit_u.z - g.vert0_u;   // GeometryUtils::xyz2uv
  const double hv = pos.x * f.unit_v.x + pos.y * f.unit_v.y + pos.z * f.unit_v.z - g.vert0_v;
  ps.created_wall = wall; ps.created_tile = uv2grid(f, g, hu, hv);
  return ps;
}
// does the kept volume reactant of a reaction with a reactive surface cross the wall? (RX_FLIP, :2694-2716)
static inline bool wallrxn_flips(const mcx_rxn_class& c, const mcx_pathway& pw, uint32_t orient_bits) {
  if (!(pw.keep_reactant_mask & 1u) || !(pw.kept_info & MCX_KEPT_VALID)) return false;
  int o = kept_code(pw, 0);
  if (o == 0) o = ((orient_bits >> 4) & 1u) ? 1 : -1;
  return c.reactant_orientation[0] != o;
}
// ---- the evaluation context ---------------------------------------------------------------
// pick_surf_displacement (diffusion_utils.inl:60-96): Marsaglia polar method on one 32-bit word
static inline void pick_surf_displacement(WordSource& rs, double scale, double& du, double& dv) {
  double au, av, f;
  do {
    uint32_t n = rs.next();
    au = 2 * 1.52587890625e-5 * (n & 0xFFFF) - 1;
    av = 2 * 1.52587890625e-5 * (n >> 16) - 1;
    f = au * au + av * av;
  } while ((f < POS_EPS) || (f > 1));
  const double normal_factor = sqrt(-log(f) / f);
  du = au * (normal_factor * scale); dv = av * (normal_factor * scale);
}

// the displacement left after a reflection, CollisionUtils::reflect_from_wall (collision_utils.inl:1735-1744)
static inline V3 reflected_displacement(V3 displacement, V3 normal, double t_reflect) {
  const double reflect_factor = -2.0 * dot(displacement, normal);
  return (displacement + normal * reflect_factor) * (1.0 - t_reflect);
}

struct Eval {
  World& w;
  WordSource& rs;
  bool snapshot;                     // true: read-only against frozen state, return proposals
  const std::vector<Mol>* frozen;    // snapshot mode: start-of-iteration molecules (w.mols itself)
  const std::vector<uint8_t>* dead;  // snapshot mode: consumed flags by index
  bool no_partners = false;          // forced-final pass
  const std::vector<uint8_t>* tile_claimed = nullptr;  // snapshot mode: per global tile, claimed by a mover in an earlier round
  const std::vector<uint32_t>* tile_start = nullptr;   // first global tile of every wall
  mcx_trace_rec* tr = nullptr;
  uint32_t words_base = 0;
  uint64_t h = 0xcbf29ce484222325ULL;

  Eval(World& w_, WordSource& r) : w(w_), rs(r), snapshot(false), frozen(nullptr), dead(nullptr) {}

  void ev(uint32_t a, uint32_t b) { h = hash_ev(h, a, b); }

  // CollisionUtils::collect_neighboring_subparts, collision_utils_subparts.inl:38-122
  void collect_neighboring_subparts(V3 pos, const int si[3], double rxn_radius, double sp_len,
                                    std::vector<uint32_t>& out) {
    const double part_len = w.cfg.partition_edge_length;
    V3 rel = pos - V3{w.cfg.origin[0], w.cfg.origin[1], w.cfg.origin[2]};
    V3 plus = rel + V3{rxn_radius, rxn_radius, rxn_radius};
    V3 minus = rel - V3{rxn_radius, rxn_radius, rxn_radius};
    V3 boundary = {si[0] * sp_len, si[1] * sp_len, si[2] * sp_len};
    auto ins = [&](int x, int y, int z) {
      int idx[3] = {x, y, z};
      uint32_t s = w.subpart_from_3d(idx);
      if (std::find(out.begin(), out.end(), s) == out.end()) out.push_back(s);
    };
    int xd = 0, yd = 0, zd = 0;
    if (minus.x < boundary.x && minus.x > 0.0) { ins(si[0] - 1, si[1], si[2]); xd = -1; }
    else if (plus.x > boundary.x + sp_len && plus.x < part_len) { ins(si[0] + 1, si[1], si[2]); xd = +1; }
    if (minus.y < boundary.y && minus.y > 0.0) { ins(si[0], si[1] - 1, si[2]); yd = -1; }
    else if (plus.y > boundary.y + sp_len && plus.y < part_len) { ins(si[0], si[1] + 1, si[2]); yd = +1; }
    if (minus.z < boundary.z && minus.z > 0.0) { ins(si[0], si[1], si[2] - 1); zd = -1; }
    else if (plus.z > boundary.z + sp_len && plus.z < part_len) { ins(si[0], si[1], si[2] + 1); zd = +1; }
    if (xd && yd) ins(si[0] + xd, si[1] + yd, si[2]);
    if (xd && zd) ins(si[0] + xd, si[1], si[2] + zd);
    if (yd && zd) ins(si[0], si[1] + yd, si[2] + zd);
    if (xd && yd && zd) ins(si[0] + xd, si[1] + yd, si[2] + zd);
  }

  // CollisionUtils::collect_crossed_subparts, collision_utils_subparts.inl:127-300
  uint32_t collect_crossed_subparts(V3 pos, uint32_t cur_subpart, V3 displacement, bool for_mols, bool for_walls,
                                    std::vector<uint32_t>& sp_walls, std::vector<uint32_t>& sp_mols) {
    const double sp_len = w.sp_len;
    auto ins_m = [&](uint32_t s) { if (std::find(sp_mols.begin(), sp_mols.end(), s) == sp_mols.end()) sp_mols.push_back(s); };
    if (for_walls) sp_walls.push_back(cur_subpart);
    if (for_mols) ins_m(cur_subpart);
    V3 dest = pos + displacement;
    V3 dnz = displacement; guard_zero_div(dnz);
    int dir[3] = {dnz.x > 0 ? 1 : 0, dnz.y > 0 ? 1 : 0, dnz.z > 0 ? 1 : 0};
    int src[3], dst[3];
    w.subpart_3d_from_index(cur_subpart, src);
    w.subpart_3d(dest, dst);
    double rr = w.cfg.rxn_radius_3d * POS_SQRT2;
    bool expanded = w.cfg.use_expanded_list != 0;
    if (for_mols && expanded) collect_neighboring_subparts(pos, src, rr, sp_len, sp_mols);
    uint32_t dest_subpart = w.subpart_from_3d(dst);
    if (cur_subpart != dest_subpart) {
      int add[3] = {dir[0] ? 1 : -1, dir[1] ? 1 : -1, dir[2] ? 1 : -1};
      V3 cur = pos;
      int ci[3] = {src[0], src[1], src[2]};
      uint32_t cs;
      V3 rcp = {1.0 / dnz.x, 1.0 / dnz.y, 1.0 / dnz.z};
      int guard = 0;
      do {
        V3 edges = {w.cfg.origin[0] + ci[0] * sp_len + dir[0] * sp_len,
                    w.cfg.origin[1] + ci[1] * sp_len + dir[1] * sp_len,
                    w.cfg.origin[2] + ci[2] * sp_len + dir[2] * sp_len};
        V3 diff = edges - cur;
        V3 ct = mul(diff, rcp);
        if (ct.x < ct.y && ct.x <= ct.z) {
          cur = cur + displacement * ct.x; ci[0] += add[0];
          if (!w.idx_in_range(ci[0])) break;
        } else if (ct.y <= ct.z) {
          cur = cur + displacement * ct.y; ci[1] += add[1];
          if (!w.idx_in_range(ci[1])) break;
        } else {
          cur = cur + displacement * ct.z; ci[2] += add[2];
          if (!w.idx_in_range(ci[2])) break;
        }
        cs = w.subpart_from_3d(ci);
        if (for_walls) sp_walls.push_back(cs);
        if (for_mols) ins_m(cs);
        if (for_mols && expanded) collect_neighboring_subparts(cur, ci, rr, sp_len, sp_mols);
        if (++guard > 4096) break;  // safety net (not in the reference)
      } while (cs != dest_subpart);
    }
    if (for_mols && expanded) collect_neighboring_subparts(dest, dst, rr, sp_len, sp_mols);
    return dest_subpart;
  }

  // CollisionUtils::get_displacement_up_to_partition_boundary, collision_utils.inl:48-114
  V3 displacement_up_to_partition_boundary(V3 pos, V3 displacement) {
    V3 dnz = displacement; guard_zero_div(dnz);
    double e = w.cfg.partition_edge_length;
    V3 edges = {w.cfg.origin[0] + (dnz.x > 0 ? 1.0 : 0.0) * e, w.cfg.origin[1] + (dnz.y > 0 ? 1.0 : 0.0) * e,
                w.cfg.origin[2] + (dnz.z > 0 ? 1.0 : 0.0) * e};
    V3 diff = edges - pos;
    double hit_time = 1;
    if (fabs(diff.x) < POS_EPS || fabs(diff.y) < POS_EPS || fabs(diff.z) < POS_EPS) return {0, 0, 0};
    V3 ct = {diff.x / dnz.x, diff.y / dnz.y, diff.z / dnz.z};
    if (ct.x >= 0 && ct.x < ct.y && ct.x <= ct.z) hit_time = ct.x;
    else if (ct.y >= 0 && ct.y <= ct.z) hit_time = ct.y;
    else if (ct.z >= 0) hit_time = ct.z;
    return displacement * (hit_time - STIME_EPS);
  }

  // CollisionUtils::jump_away_line, collision_utils.inl:568-603
  void jump_away_line(V3 p, double k, V3 A, V3 B, V3 n, V3& v) {
    V3 e = B - A;
    double le_1 = 1.0 / sqrt(dot(e, e));
    e = e * le_1;
    V3 f = {n.y * e.z - n.z * e.y, n.z * e.x - n.x * e.z, n.x * e.y - n.y * e.x};
    double tiny = POS_EPS * (abs_max_2vec(p, v) + 1.0) / (k * max3(V3{fabs(f.x), fabs(f.y), fabs(f.z)}));
    if ((rs.next() & 1) == 0) tiny = -tiny;
    v.x -= tiny * f.x; v.y -= tiny * f.y; v.z -= tiny * f.z;
  }

  // CollisionUtils::collide_wall, collision_utils.inl:629-812 (update_move = true)
  int collide_wall(V3 pos, uint32_t wi, V3& move, double& t, V3& hit) {
    w.stats.ray_polygon_tests++;
    const Wall& f = w.walls[wi];
    double dp = dot(f.normal, pos), dv = dot(f.normal, move), dd = dp - f.distance_to_origin, d_eps;
    if (dd > 0) {
      d_eps = POS_EPS;
      if (dd < d_eps) d_eps = 0.5 * dd;
      if (dd + dv > d_eps) return WALL_MISS;
    } else {
      d_eps = -POS_EPS;
      if (dd > d_eps) d_eps = 0.5 * dd;
      if (dd < 0 && dd + dv < d_eps) return WALL_MISS;
    }
    double a;
    if (dd == 0) {
      if (dv != 0) return WALL_MISS;
      a = (abs_max_2vec(pos, move) + 1.0) * POS_EPS;
      if ((rs.next() & 1) == 0) a = -a;
      move = move - f.normal * a;  // dd == 0.0 branch
      return WALL_REDO;
    }
    a = 1.0 / dv;
    a *= -dd;
    t = a;
    hit = pos + move * a;
    V3 v0 = w.verts[f.vi[0]];
    V3 local = hit - v0;
    double b = dot(local, f.unit_u), c = dot(local, f.unit_v), ff;
    if (f.uv_vert2_v < 0) { c = -c; ff = -f.uv_vert2_v; } else ff = f.uv_vert2_v;
    if (c > 0) {
      double g = b * ff, hh = c * f.uv_vert2_u;
      if (g > hh) {
        if (c * f.uv_vert1_u + g < hh + f.uv_vert1_u * f.uv_vert2_v) return dv > 0 ? WALL_BACK : WALL_FRONT;
        else if (!distinguishable(c * f.uv_vert1_u + g, hh + f.uv_vert1_u * f.uv_vert2_v, POS_EPS)) {
          jump_away_line(pos, a, w.verts[f.vi[1]], w.verts[f.vi[2]], f.normal, move);
          return WALL_REDO;
        } else return WALL_MISS;
      } else if (!distinguishable(g, hh, POS_EPS)) {
        jump_away_line(pos, a, w.verts[f.vi[2]], v0, f.normal, move);
        return WALL_REDO;
      } else return WALL_MISS;
    } else if (!distinguishable(c, 0.0, POS_EPS)) {
      jump_away_line(pos, a, v0, w.verts[f.vi[1]], f.normal, move);
      return WALL_REDO;
    }
    return WALL_MISS;
  }

  // CollisionUtils::get_closest_wall_collision, collision_utils.inl:819-914
  bool closest_wall_collision(V3 pos, uint32_t subpart, uint32_t last_hit_wall, V3& displacement,
                              V3& disp_up_to_wall, Collision& best) {
    int guard = 0;
  restart:
    bool found = false;
    double closest = TIME_FOREVER;
    for (uint32_t wi : w.walls_per_subpart[subpart]) {
      if (wi == last_hit_wall) continue;
      double t; V3 hit;
      int ct = collide_wall(pos, wi, displacement, t, hit);
      if (ct == WALL_REDO) {
        w.stats.redos++; ev(EV_REDO, wi);
        if (tr) tr->n_redo++;
        if (++guard > 64) return false;  // safety net (not in the reference)
        goto restart;
      } else if (ct != WALL_MISS) {
        w.stats.ray_polygon_colls++;
        if (w.subpart_index(hit) != subpart) continue;
        if (t < closest) {
          found = true; closest = t;
          best.type = ct == WALL_FRONT ? COLL_WALL_FRONT : COLL_WALL_BACK;
          best.time = t; best.pos = hit; best.wall = wi; best.partner_id = MCX_NONE; best.rxn_class = -1;
        }
      }
    }
    if (found) disp_up_to_wall = best.pos - pos;
    return found;
  }

  // CollisionUtils::collide_mol, collision_utils.inl:464-515
  inline bool collide_mol(V3 mpos, uint32_t mid, V3 disp, const Mol& c, double R, double& t, V3& cpos) {
    w.stats.collide_mol_tests++;
    V3 dir = c.pos - mpos;
    double d = dot(dir, disp);
    if (d < 0) return false;
    double movelen2 = dot(disp, disp);
    if (d > movelen2) return false;
    double dirlen2 = dot(dir, dir);
    double sigma2 = R * R;
    if (movelen2 * dirlen2 - d * d > movelen2 * sigma2) return false;
    if (mid == c.id) return false;
    t = d / movelen2;
    cpos = mpos + disp * t;
    return true;
  }

  // Point location by a ray cast (region releases; Region::is_point_inside, geometry.cpp:1048-1086, and
  // compute_counted_volume_for_pos, collision_utils.inl:1515-1566, count the walls crossed on the way to a far point):
  // one ray from pos towards -x (skewed in y and z) up to the partition boundary through the subpartitions it crosses;
  // every wall it crosses is counted once, in the subpartition that holds the crossing point.
  struct RayScan { uint32_t inside_mask = 0; uint32_t first_wall = MCX_NONE; int first_side = WALL_MISS; bool redo = false; };
  RayScan scan_ray(V3 pos) {
    RayScan out;
    const double pl = w.cfg.partition_edge_length;
    const V3 raw = {-pl, pl * (1.0 / 11.0), pl * (1.0 / 22.0)};
    const V3 disp = displacement_up_to_partition_boundary(pos, raw);
    if (disp.x == 0 && disp.y == 0 && disp.z == 0) return out;
    std::vector<uint32_t> sp_walls, sp_mols;
    collect_crossed_subparts(pos, w.subpart_index(pos), disp, false, true, sp_walls, sp_mols);
    double first_t = 2.0;
    for (uint32_t S : sp_walls)
      for (uint32_t wi : w.walls_per_subpart[S]) {
        V3 move = disp, hit; double t;
        const int ct = collide_wall(pos, wi, move, t, hit);
        if (ct == WALL_REDO) { out.redo = true; return out; }
        if ((ct == WALL_FRONT || ct == WALL_BACK) && w.subpart_index(hit) == S) {
          const uint32_t obj = w.walls[wi].object;
          if (obj < 32u) out.inside_mask ^= 1u << obj;
          if (t < first_t) { first_t = t; out.first_wall = wi; out.first_side = ct; }
        }
      }
    return out;
  }

  // ray_trace_vol, diffuse_react_event.cpp:627-780.  Returns true if a wall was hit.
  // pos/subpart are the molecule's current position; on FINISHED the caller moves it.
  bool ray_trace_vol(V3 pos, uint32_t subpart, uint32_t self_id, uint32_t species, bool can_vol_react,
                     uint32_t last_hit_wall, V3& remaining, std::vector<Collision>& colls) {
    colls.clear();
    double R = w.cfg.rxn_radius_3d;
    V3 part_disp = remaining;
    if (!w.in_this_partition(pos + remaining)) part_disp = displacement_up_to_partition_boundary(pos, remaining);
    std::vector<uint32_t> sp_walls, sp_mols;
    uint32_t last_subpart = collect_crossed_subparts(pos, subpart, part_disp, can_vol_react, true, sp_walls, sp_mols);
    V3 up_to_wall = remaining;
    bool hit_wall = false, hit_in_last = false;
    Collision closest;
    for (uint32_t s : sp_walls) {
      if (closest_wall_collision(pos, s, last_hit_wall, remaining, up_to_wall, closest)) {
        colls.push_back(closest);
        hit_wall = true;
        hit_in_last = last_subpart == s;
        break;
      }
    }
    if (can_vol_react && !no_partners) {
      if (hit_wall && !hit_in_last) {
        sp_mols.clear(); sp_walls.clear();
        collect_crossed_subparts(pos, subpart, up_to_wall, true, false, sp_walls, sp_mols);
      }
      size_t ns = w.species.size();
      for (uint32_t s : sp_mols) {
        for (size_t b = 0; b < ns; b++) {
          int rc = w.bimol[species * ns + b];
          if (rc < 0) continue;
          auto it = w.lists.find(list_key((uint32_t)b, s));
          if (it == w.lists.end()) continue;
          for (uint32_t cid : it->second) {
            uint32_t cidx = w.id_to_index[cid];
            const Mol& c = w.mols[cidx];
            if (snapshot ? (*dead)[cidx] : (c.flags & MCX_MOL_DEFUNCT)) continue;
            double t; V3 cp;
            if (collide_mol(pos, self_id, remaining, c, R, t, cp)) {
              Collision k; k.type = COLL_VOLMOL; k.time = t; k.pos = cp; k.partner_id = cid; k.partner_index = cidx;
              k.rxn_class = rc; k.wall = MCX_NONE;
              colls.push_back(k);
            }
          }
        }
      }
    }
    return hit_wall;
  }

  // ExactDiskUtils::exact_disk (exact_disk_utils.inl:840-1145) for a collision at `loc` of the molecule moving by
  // `mv` with the molecule at `target`: walls of the collision subpartition (:921-925)
  double exact_disk_factor(V3 loc, V3 mv, uint32_t species, V3 target) {
    const std::vector<uint32_t>& wl = w.walls_per_subpart[w.subpart_index(loc)];
    if (wl.empty()) return 1;
    std::vector<double> tri(9 * wl.size()), plane(4 * wl.size());
    std::vector<unsigned char> skip(wl.size(), 0);
    for (size_t k = 0; k < wl.size(); k++) {
      const Wall& f = w.walls[wl[k]];
      for (int q = 0; q < 3; q++) { V3 p = w.verts[f.vi[q]]; tri[9 * k + 3 * q] = p.x; tri[9 * k + 3 * q + 1] = p.y; tri[9 * k + 3 * q + 2] = p.z; }
      plane[4 * k] = f.normal.x; plane[4 * k + 1] = f.normal.y; plane[4 * k + 2] = f.normal.z; plane[4 * k + 3] = f.distance_to_origin;
      skip[k] = exd_passes_through(w, species, f.surf_class) ? 1 : 0;
    }
    bool same = !distinguishable_vec3(loc, target, POS_EPS);
    return orc_exd::exact_disk({loc.x, loc.y, loc.z}, {mv.x, mv.y, mv.z}, w.cfg.rxn_radius_3d, {target.x, target.y, target.z},
                               same, (int)wl.size(), tri.data(), plane.data(), skip.data());
  }

  // RxnUtils::test_bimolecular, rxn_utils.inl:336-414; local_prob_factor > 0 only between two surface molecules
  int test_bimolecular(const mcx_rxn_class& rc, double scaling, double local_prob_factor = 0) {
    double max_fixed_p = rc.max_fixed_p, prob;
    if (local_prob_factor != 0) max_fixed_p = rc.max_fixed_p * local_prob_factor;
    if (max_fixed_p < scaling) {
      prob = rs.dbl() * scaling;
      if (prob >= max_fixed_p) return -1;
    } else {
      float max_p = (float)rc.max_fixed_p;  // sic: float in the reference (rxn_utils.inl:369)
      if (local_prob_factor > 0) max_p *= local_prob_factor;
      if (max_p >= scaling) {
        prob = rs.dbl() * max_p;  // skipped-reaction accounting omitted (stats only)
      } else {
        prob = rs.dbl() * scaling;
        if (prob >= max_p) return -1;
      }
    }
    return pathway_for_probability(w, rc, prob, local_prob_factor > 0 ? local_prob_factor : 1);
  }
  // RxnUtils::test_many_bimolecular with all_neighbors_flag (rxn_utils.inl:475-580): which of the n matching classes
  // reacts (-1: none); pathway = chosen_pathway_index
  int test_many_bimolecular(const std::vector<int>& rcs, const std::vector<double>& scaling, double local_prob_factor, int& pathway) {
    const int n = (int)rcs.size();
    if (n == 1) { pathway = test_bimolecular(w.classes[rcs[0]], scaling[0], local_prob_factor); return pathway; }  // sic: returns the pathway
    std::vector<double> cum(2 * n, 0.0);  // sic: twice as long, the binary search below runs over the padded array
    cum[0] = w.classes[rcs[0]].max_fixed_p * local_prob_factor / scaling[0];
    for (int i = 1; i < n; i++) cum[i] = cum[i - 1] + w.classes[rcs[i]].max_fixed_p * local_prob_factor / scaling[i];
    double prob;
    if (cum[n - 1] > 1.0) prob = rs.dbl() * cum[n - 1];  // skipped-reaction accounting omitted (stats only)
    else {
      prob = rs.dbl();
      if (prob > cum[n - 1]) return -1;
    }
    int min_idx = 0, max_idx = 2 * n - 1;  // binary_search_double(cum, prob, cum.size() - 1, 1), rxn_utils.inl:301-320
    while (max_idx - min_idx > 1) {
      const int mid = (max_idx + min_idx) / 2;
      if (prob > cum[mid]) min_idx = mid; else max_idx = mid;
    }
    const int rxn_index = prob > cum[min_idx] ? max_idx : min_idx;
    if (rxn_index > 0) prob = prob - cum[rxn_index - 1];
    prob = prob * scaling[rxn_index];
    pathway = pathway_for_probability(w, w.classes[rcs[rxn_index]], prob, local_prob_factor);
    return rxn_index;
  }

  // RxnUtils::test_intersect, rxn_utils.inl:593-626 (a Standard reaction with a reactive surface): pathway or -1
  int test_intersect(const mcx_rxn_class& rc, double scaling) {
    const double max_prob = rc.max_fixed_p;
    double pr;
    if (max_prob > scaling) pr = rs.dbl() * max_prob;
    else { pr = rs.dbl() * scaling; if (pr > max_prob) return -1; }
    if (pr > rc.max_fixed_p) return -1;
    const double match = rs.dbl() * rc.max_fixed_p;
    return pathway_for_probability(w, rc, match);
  }

  // compute_vol_displacement + pick_vol_displacement, diffusion_utils.inl:366-432,113-119
  void compute_vol_displacement(const mcx_species& sp, double& max_time, V3& disp, double& r_rate_factor,
                                double& t_steps) {
    double steps = 1.0, rate_factor;
    t_steps = steps * sp.time_step;
    if (t_steps > max_time) { t_steps = max_time; steps = max_time / sp.time_step; }
    if (steps < EPS) { steps = EPS; t_steps = EPS * sp.time_step; }
    double scale;
    if (steps == 1.0) { scale = sp.space_step; r_rate_factor = rate_factor = 1.0; }
    else { rate_factor = sqrt(steps); r_rate_factor = 1.0 / rate_factor; scale = rate_factor * sp.space_step; }
    disp.x = scale * rs.gauss() * 0.70710678118654752440;
    disp.y = scale * rs.gauss() * 0.70710678118654752440;
    disp.z = scale * rs.gauss() * 0.70710678118654752440;
    max_time = t_steps;
  }

  // pick_unimol_rxn_class_and_set_rxn_time (diffuse_react_event.cpp:1731-1758) +
  // time_of_unimol (rxn_utils.inl:721-736)
  double pick_unimol_time(uint32_t species, double current_time) {
    int rc = w.unimol[species];
    if (rc < 0) return TIME_INVALID;
    double k_tot = w.classes[rc].max_fixed_p;
    double p = rs.dbl();
    double from_now;
    if (k_tot <= 0 || !distinguishable(p, 0, EPS)) from_now = TIME_FOREVER;
    else from_now = -log(p) / k_tot;
    return current_time + from_now;
  }
};

// ================================================================================================
// One call of diffuse_single_molecule (diffuse_react_event.cpp:201-337) incl. diffuse_vol_molecule
// (:367-618) for one molecule: one sub-step of the iteration.  `again` reports that the reference would
// push a new DiffuseAction for the same iteration (:299-308, :326-329).
//
// apply == true  (SEQUENTIAL): reactions mutate the world immediately (reference semantics).
// apply == false (SNAPSHOT):   the first claiming event (bimolecular reaction, absorption, unimolecular
//                 firing) ends the evaluation and is returned as a proposal.
// ================================================================================================
static void seq_apply_bimol(World& w, uint32_t a_index, uint32_t b_index, int rc, int pathway, V3 pos, double t,
                            uint32_t orient_bits, bool& a_destroyed, bool* flip = nullptr, int coll_side = 0, const Placement* pl = nullptr);
static void seq_apply_unimol(World& w, uint32_t index, int rc, int pathway, double t, uint32_t orient_bits, bool& destroyed,
                             const Placement* pl = nullptr);
static void seq_apply_surfsurf(World& w, uint32_t a_index, uint32_t b_index, int rc, int pathway, double t, uint32_t bits, bool& a_destroyed,
                               const Placement* pl = nullptr);
static void seq_apply_wallrxn(World& w, uint32_t index, int rc, int pathway, V3 pos, double t, uint32_t orient_bits, uint32_t wall,
                              uint32_t cvi, bool& destroyed, int coll_side);
static void seq_set_defunct(World& w, Mol& m);

struct MolState { V3 pos; uint32_t subpart; double t_now; uint32_t flags; double unimol_time;
                  uint32_t created_wall = MCX_NONE, created_tile = MCX_NONE; uint32_t cvi = 0;
                  uint32_t wall = MCX_NONE, tile = MCX_NONE; double u = 0, v = 0; };

static Outcome evaluate_substep(Eval& E, uint32_t index, MolState& s, bool apply, bool& again) {
  World& w = E.w;
  Outcome out;
  again = false;
  const uint32_t m_id = w.mols[index].id, m_species = w.mols[index].species;
  const double it = (double)w.iteration, t_end = it + 1;
  std::vector<Collision> colls;
  mcx_trace_rec* tr = E.tr;
  const mcx_species sp = w.species[m_species];
  auto fill_event = [&](Outcome& o) {
    o.t_now = s.t_now; o.flags = s.flags; o.unimol_time = s.unimol_time; o.cvi = s.cvi;
    o.wall = s.wall; o.tile = s.tile; o.u = s.u; o.v = s.v;
  };

  // a tile a product may take: vacant in the tile table; SNAPSHOT: and not claimed in an earlier conflict round, none in the forced pass
  auto tile_is_vacant = [&](uint32_t wi, uint32_t ti) {
    if (!w.tiles[wi].empty() && w.tiles[wi][ti] != MCX_NONE) return false;
    if (E.snapshot) {
      if (E.no_partners) return false;
      if ((*E.tile_claimed)[(*E.tile_start)[wi] + ti]) return false;
    }
    return true;
  };
  // -- a counted volume that is only a guess (MCX_MOL_CVI_PENDING): Partition::add_volume_molecule's ray cast
  // (partition.h:572-576 -> compute_counted_volume_using_waypoints), done when the molecule is first evaluated
  if (s.flags & MCX_MOL_CVI_PENDING) {
    s.flags &= ~MCX_MOL_CVI_PENDING;
    if (w.mols[index].wall == MCX_NONE && !w.cv_mask.empty()) {
      const Eval::RayScan sc = E.scan_ray(s.pos);
      if (!sc.redo) { const uint32_t kq = cv_lookup(w, sc.inside_mask & w.cv_all); if (kq != MCX_NONE) s.cvi = kq; }
    }
  }
  // -- unimolecular firing (diffuse_single_molecule :215-223 -> react_unimol_single_molecule :1764-1826)
  if (s.unimol_time != TIME_INVALID && s.unimol_time <= s.t_now) {
    int rc = w.unimol[m_species];
    int pathway = 0;
    if (w.classes[rc].n_pathways > 1) {  // which_unimolecular, rxn_utils.inl:774-783
      double match = E.rs.dbl() * w.classes[rc].max_fixed_p;
      pathway = pathway_for_probability(w, w.classes[rc], match);
    }
    uint32_t obits = 0;
    const mcx_pathway& upw = w.pathways[w.classes[rc].first_pathway + pathway];
    Placement pl;
    bool blocked = false;
    if (w.mols[index].wall != MCX_NONE) {  // is_orientable: the reactant is a surface molecule (:2569)
      if (pathway_is_general(w, w.classes[rc], upw)) {
        const SurfSite self{s.wall, s.tile, s.u, s.v, w.mols[index].orient, m_species};
        blocked = !place_general(w, w.classes[rc], upw, s.wall, s.tile, &self, (upw.keep_reactant_mask & 1u) ? 0 : 1, E.rs, tile_is_vacant, pl, obits);
      } else obits = draw_orientation_bits(upw, E.rs);
    }
    if (blocked) {
      // RX_BLOCKED (outcome_unimolecular :2976-2999): no room for the products; the molecule lives on and draws a new lifetime
      E.ev(EV_BLOCKED, (uint32_t)rc);
    } else {
    E.ev(EV_UNIMOL | (uint32_t)pathway, (uint32_t)rc);
    if (tr) { tr->rxn_class = rc; tr->rxn_pathway = pathway; tr->t_event = s.unimol_time; }
    if (!apply) {
      out.kind = MCX_OUT_UNIMOL; out.pos = s.pos; out.rxn_class = rc; out.pathway = pathway;
      out.t_event = s.unimol_time; out.orient_bits = obits; fill_event(out); put_placement(out, pl);
      return out;
    }
    bool destroyed = false;
    w.mols[index].pos = s.pos; w.mols[index].subpart = s.subpart; w.mols[index].cvi = s.cvi;
    seq_apply_unimol(w, index, rc, pathway, s.unimol_time, obits, destroyed, &pl);
    if (destroyed) { out.kind = MCX_OUT_UNIMOL; out.pos = s.pos; out.t_event = s.unimol_time; return out; }
    }
    s.flags |= MCX_MOL_SCHEDULE_UNIMOL;  // survivor re-draws its lifetime (outcome_unimolecular :2999)
  }
  // -- newbie lifetime (diffuse_single_molecule :232-236)
  if (s.flags & MCX_MOL_SCHEDULE_UNIMOL) {
    s.flags &= ~MCX_MOL_SCHEDULE_UNIMOL;
    s.unimol_time = E.pick_unimol_time(m_species, s.t_now);
  }
  // -- get_max_time (:164-198); barrier = end of this iteration; species.time_step == 1
  double max_time = t_end - s.t_now;
  if (s.unimol_time != TIME_INVALID && s.unimol_time < s.t_now + max_time) max_time = s.unimol_time - s.t_now;

  bool destroyed = false;
  bool surf_tile_changed = false;
  // SPECIES_FLAG_CAN_SURFSURF: diffuse_surf_molecule runs for such a molecule even when it cannot diffuse (:274-283)
  const bool can_ss = s.wall != MCX_NONE && w.can_surf_surf[m_species];
  const bool surf_diffusible = (sp.flags & MCX_SP_CAN_DIFFUSE) != 0;
  bool ss_fired = false;   // SNAPSHOT: a surface-surface reaction ended the evaluation (the outcome is completed at the end)
  if (s.wall != MCX_NONE && (surf_diffusible || can_ss)) {
    // ---- diffuse_surf_molecule (:1071-1246)
    double t_steps = sp.time_step > max_time ? max_time : sp.time_step;
    double steps;
    if (sp.time_step > max_time) {
      steps = max_time / sp.time_step;
      if (steps < EPS) { t_steps = EPS * sp.time_step; steps = EPS; }
    } else steps = 1.0;
    const double space_factor = steps == 1.0 ? sp.space_step : sp.space_step * sqrt(steps);
    const uint32_t original_wall = s.wall;
    auto available = [&](uint32_t wi, uint32_t ti) {
      if (!w.tiles[wi].empty() && w.tiles[wi][ti] != MCX_NONE) return false;
      if (E.snapshot) {
        if (E.no_partners) return false;                       // forced-final pass: no new tile claims
        if ((*E.tile_claimed)[(*E.tile_start)[wi] + ti]) return false;  // a mover of an earlier round took it
      }
      return true;
    };
    for (int find_new_position = surf_diffusible ? 11 : 0; find_new_position > 0; find_new_position--) {  // SURFACE_DIFFUSION_RETRIES + 1
      double du, dv;
      pick_surf_displacement(E.rs, space_factor, du, dv);
      double nu, nv;
      bool absorbed_at_border = false;
      uint32_t new_wall = ray_trace_surf(w, s.wall, s.u, s.v, du, dv, nu, nv, m_species, w.mols[index].orient, &absorbed_at_border);
      if (absorbed_at_border) {
        // absorptive region border (:1152-1160): outcome_unimolecular of the border's class at the start of the step, no products
        E.ev(EV_ABSORB, s.wall);
        if (tr) tr->t_event = s.t_now;
        if (!apply) { out.kind = MCX_OUT_ABSORBED; out.pos = s.pos; out.t_event = s.t_now; fill_event(out); return out; }
        w.stats.absorptions++;
        destroyed = true; out.kind = MCX_OUT_ABSORBED; out.pos = s.pos; out.t_event = s.t_now;
        seq_set_defunct(w, w.mols[index]);
        return out;
      }
      if (new_wall == MCX_NONE) continue;  // ambiguous edge hit: try again
      uint32_t new_tile = uv2grid(w.walls[new_wall], w.grids[new_wall], nu, nv);
      if (new_tile == MCX_NONE) continue;
      if (new_wall == s.wall) {  // move_sm_on_same_triangle (diffusion_utils.inl:453-485)
        if (new_tile != s.tile) {
          if (!available(new_wall, new_tile)) continue;
          surf_tile_changed = true;
        }
      } else {  // move_sm_to_new_triangle (:504-548)
        if (!available(new_wall, new_tile)) continue;
        surf_tile_changed = true;
        // reschedule the unimolecular reaction of a molecule that changed wall (:1170-1186)
        double time_until_unimol = s.unimol_time - t_steps - s.t_now;
        time_until_unimol = (time_until_unimol < 0) ? 0 : time_until_unimol;
        if (s.unimol_time == TIME_INVALID || (time_until_unimol > EPS || time_until_unimol > EPS * (s.t_now + t_steps))) {
          s.unimol_time = TIME_INVALID;
          s.flags |= MCX_MOL_SCHEDULE_UNIMOL;
        }
      }
      if (apply && surf_tile_changed) {  // Grid::reset_molecule_tile / set_molecule_tile
        if (w.tiles[new_wall].empty()) w.tiles[new_wall].assign(w.grids[new_wall].n_tiles, MCX_NONE);
        w.tiles[s.wall][s.tile] = MCX_NONE;
        w.tiles[new_wall][new_tile] = m_id;
      }
      s.wall = new_wall; s.tile = new_tile; s.u = nu; s.v = nv;
      s.pos = uv2xyz(w, w.walls[new_wall], nu, nv);
      s.subpart = w.subpart_index(s.pos);
      break;
    }
    if (apply) {
      Mol& mm = w.mols[index];
      mm.wall = s.wall; mm.tile = s.tile; mm.u = s.u; mm.v = s.v; mm.pos = s.pos;
    }
    // ---- react_2D_all_neighbors (:1250-1393): after its move the molecule tests the molecules on the tiles around its own
    if (can_ss && !(sp.flags & MCX_SP_CANT_INITIATE) && !(E.snapshot && E.no_partners)) {
      TileNeighbors nbt;
      find_neighbor_tiles(w, s.wall, s.tile, nbt);
      std::vector<int> m_rc; std::vector<double> m_factor; std::vector<uint32_t> m_index;
      const size_t ns = w.species.size();
      for (const WallTile& nt : nbt) {
        // Grid::get_molecule_on_tile; the list only holds walls with a grid and the molecule's own wall (SNAPSHOT: a
        // wall it has just moved to gets its grid at the end of the iteration; nobody else is on it yet)
        if (w.tiles[nt.first].empty()) continue;
        const uint32_t nid = w.tiles[nt.first][nt.second];
        if (nid == MCX_NONE || nid == m_id) continue;        // SNAPSHOT: the tile table still shows the mover on its old tile
        const uint32_t j = w.id_to_index[nid];
        if (E.snapshot && (*E.dead)[j]) continue;
        const Mol& nsm = w.mols[j];
        const int rc = w.surfsurf[m_species * ns + nsm.species];
        // trigger_bimolecular_orientation_from_mols (rxn_utils.inl:58-97)
        if (rc < 0 || !orientations_match(w.classes[rc], w.mols[index].orient, nsm.orient)) continue;
        m_rc.push_back(rc); m_factor.push_back(t_steps / w.grids[nt.first].binding_factor); m_index.push_back(j);
      }
      if (!nbt.empty() && !m_rc.empty()) {
        const double local_prob_factor = 3.0 / nbt.size();
        int which, pathway;
        if (m_rc.size() == 1) { pathway = E.test_bimolecular(w.classes[m_rc[0]], m_factor[0], local_prob_factor); which = 0; }
        else {
          which = E.test_many_bimolecular(m_rc, m_factor, local_prob_factor, pathway);
          pathway = 0;  // sic (TODO_PATHWAYS, :1367): the first pathway of the chosen class
        }
        if (tr) for (uint32_t j : m_index) { if (tr->n_collisions < MCX_TRACE_K) tr->partner[tr->n_collisions] = w.mols[j].id; tr->n_collisions++; }
        if (which >= 0 && pathway >= 0) {
          const int rc = m_rc[which];
          const uint32_t j = m_index[which];
          const mcx_rxn_class& cl = w.classes[rc];
          const mcx_pathway& pw = w.pathways[cl.first_pathway + pathway];
          // random draws in the reference's order: tile assignment (find_surf_product_positions), then orientations
          uint32_t bits = 0;
          Placement pl;
          bool blocked = false;
          if (pathway_is_general(w, cl, pw)) {
            const SurfSite me{s.wall, s.tile, s.u, s.v, w.mols[index].orient, m_species}, other = site_of(w.mols[j]);
            const bool me_r0 = m_species == cl.reactants[0];
            SurfSite rec[2]; int n_rec = 0;
            if (!(pw.keep_reactant_mask & 1u)) rec[n_rec++] = me_r0 ? me : other;
            if (!(pw.keep_reactant_mask & 2u)) rec[n_rec++] = me_r0 ? other : me;
            blocked = !place_general(w, cl, pw, s.wall, s.tile, rec, n_rec, E.rs, tile_is_vacant, pl, bits);
          } else {
            bits = surfsurf_position_bits(w, cl, pw, m_species == cl.reactants[0], E.rs);
            bits |= draw_orientation_bits(pw, E.rs);
          }
          const double t_rxn = s.t_now;  // collision_time = diffusion_start_time (:1343)
          if (blocked) E.ev(EV_BLOCKED, w.mols[j].id);   // RX_BLOCKED: the molecule survives (:1388-1392)
          else {
          E.ev(EV_SURFSURF | (uint32_t)pathway, (uint32_t)rc);
          E.ev(EV_RXN | (bits & 0x7Fu), w.mols[j].id);
          if (tr) { tr->rxn_class = rc; tr->rxn_pathway = pathway; tr->rxn_partner = w.mols[j].id; tr->t_event = t_rxn; }
          if (!apply) {
            out.rxn_class = rc; out.pathway = pathway; out.partner_index = j; out.partner_id = w.mols[j].id;
            out.t_event = t_rxn; out.orient_bits = bits; out.surf_moved = surf_tile_changed;
            put_placement(out, pl);
            ss_fired = true;
          } else {
            bool a_destroyed = false;
            seq_apply_surfsurf(w, index, j, rc, pathway, t_rxn, bits, a_destroyed, &pl);
            if (a_destroyed) { out.kind = MCX_OUT_REACTED; out.pos = s.pos; out.t_event = t_rxn; return out; }
          }
          }
        }
      }
    }
    // MCell3 compatibility rule at the end of diffuse_surf_molecule (:1222-1236)
    if ((!surf_diffusible || s.wall != original_wall) && s.unimol_time >= t_end) {
      s.unimol_time = TIME_INVALID;
      s.flags |= MCX_MOL_SCHEDULE_UNIMOL;
    }
    max_time = t_steps;
  } else if (sp.flags & MCX_SP_CAN_DIFFUSE) {
    // ---- diffuse_vol_molecule (:367-618)
    V3 remaining; double r_rate_factor, t_steps;
    E.compute_vol_displacement(sp, max_time, remaining, r_rate_factor, t_steps);
    uint32_t last_hit_wall = MCX_NONE;
    if (s.created_tile == KEPT_AT_WALL) {  // a kept reactant carries on from the wall of its event (see KEPT_AT_WALL)
      const Wall& kw = w.walls[s.created_wall];
      const double dd = dot(kw.normal, s.pos) - kw.distance_to_origin, dn = dot(remaining, kw.normal);
      if (dd > 0 ? dn < 0 : dn > 0) remaining = remaining + kw.normal * (-2.0 * dn);
      last_hit_wall = s.created_wall;
      s.created_wall = s.created_tile = MCX_NONE;
    }
    double elapsed = s.t_now;
    bool can_vol_react = w.can_vol_react[m_species] != 0;
    bool hit;
    int trace_guard = 0;
    do {
      hit = E.ray_trace_vol(s.pos, s.subpart, m_id, m_species, can_vol_react, last_hit_wall, remaining, colls);
      if (colls.size() > 1) {  // sort_collisions_by_time (:341-364)
        std::stable_sort(colls.begin(), colls.end(), [](const Collision& a, const Collision& b) {
          if (a.time < b.time) return true;
          if (a.time > b.time) return false;
          if (a.type == COLL_VOLMOL && b.type == COLL_VOLMOL) return a.partner_id > b.partner_id;
          return false;
        });
      }
      for (Collision& c : colls) {
        if (c.type == COLL_VOLMOL) {
          w.stats.volvol_collisions++;
          if (c.time < STIME_EPS) continue;  // is_immediate_collision, collision_utils.inl:814-816
          if (apply && (w.mols[c.partner_index].flags & MCX_MOL_DEFUNCT)) continue;
          // collide_and_react_with_vol_mol (:786-829)
          double factor = E.exact_disk_factor(c.pos, remaining, m_species, w.mols[c.partner_index].pos);
          if (factor < 0) { E.ev(EV_BLOCKED, c.partner_id); continue; }  // reaction blocked by a wall
          if (factor != 1.0) E.ev(EV_DISK, (uint32_t)llrint(factor * 1073741824.0));  // 2^-30 resolution
          double abs_t = elapsed + t_steps * c.time;
          double scaling = factor * r_rate_factor;
          E.ev(EV_COLL, c.partner_id);
          if (tr) { if (tr->n_collisions < MCX_TRACE_K) tr->partner[tr->n_collisions] = c.partner_id; tr->n_collisions++; }
          int pathway = E.test_bimolecular(w.classes[c.rxn_class], scaling);
          if (pathway < 0) continue;
          E.ev(EV_RXN | (uint32_t)pathway, (uint32_t)c.rxn_class);
          if (tr) { tr->rxn_class = c.rxn_class; tr->rxn_pathway = pathway; tr->rxn_partner = c.partner_id; tr->t_event = abs_t; }
          if (!apply) {
            out.kind = MCX_OUT_REACTED; out.pos = c.pos; out.rxn_class = c.rxn_class; out.pathway = pathway;
            out.partner_index = c.partner_index; out.partner_id = c.partner_id; out.t_event = abs_t; fill_event(out);
            return out;
          }
          bool a_destroyed = false;
          w.mols[index].pos = s.pos; w.mols[index].subpart = s.subpart; w.mols[index].cvi = s.cvi;
          seq_apply_bimol(w, index, c.partner_index, c.rxn_class, pathway, c.pos, abs_t, 0, a_destroyed);
          if (a_destroyed) { destroyed = true; out.kind = MCX_OUT_REACTED; out.pos = c.pos; out.t_event = abs_t; break; }
        } else {
          // ---- wall collision (:476-567)
          const Wall& wall = w.walls[c.wall];
          int wall_rc = -1;
          int action = surf_action(w, m_species, wall.surf_class, c.type, &wall_rc);
          if (tr) {
            if (tr->n_wall_hits < MCX_TRACE_K) { tr->wall[tr->n_wall_hits] = c.wall; tr->wall_side[tr->n_wall_hits] = c.type; }
            tr->n_wall_hits++;
          }
          // ---- collide_and_react_with_surf_mol (:845-975): the surface molecule on the tile under the hit point
          if (w.can_vol_surf[m_species] && !w.tiles[c.wall].empty()) {
            const Grid& g = w.grids[c.wall];
            uint32_t j = xyz2grid(w, c.pos, wall, g);
            uint32_t occ_id = w.tiles[c.wall][j];
            uint32_t occ_index = occ_id != MCX_NONE ? w.id_to_index[occ_id] : MCX_NONE;
            bool occupied = occ_index != MCX_NONE &&
                            !(E.snapshot ? (*E.dead)[occ_index] != 0 : (w.mols[occ_index].flags & MCX_MOL_DEFUNCT) != 0);
            if (occupied && s.created_wall == c.wall && s.created_tile == j) {
              s.created_wall = s.created_tile = MCX_NONE;  // no rebinding where it was just created; next time yes (:877-885)
              occupied = false;
            }
            if (occupied) {
              const Mol& sm = w.mols[occ_index];
              int rc = w.volsurf[m_species * w.species.size() + sm.species];
              int coll_orient = c.type == COLL_WALL_FRONT ? 1 : -1;
              if (rc >= 0 && orientations_match(w.classes[rc], coll_orient, sm.orient)) {
                double scaling = r_rate_factor / g.binding_factor;
                double abs_t = elapsed + t_steps * c.time;
                E.ev(EV_SURFMOL | (uint32_t)c.type, sm.id);
                if (tr) { if (tr->n_collisions < MCX_TRACE_K) tr->partner[tr->n_collisions] = sm.id; tr->n_collisions++; }
                int pathway = E.test_bimolecular(w.classes[rc], scaling);
                Placement pl;
                uint32_t obits = 0;
                if (pathway >= 0) {
                  const mcx_pathway& vpw = w.pathways[w.classes[rc].first_pathway + pathway];
                  if (pathway_is_general(w, w.classes[rc], vpw)) {
                    const SurfSite partner = site_of(sm);
                    if (!place_general(w, w.classes[rc], vpw, sm.wall, sm.tile, &partner, (vpw.keep_reactant_mask & 2u) ? 0 : 1, E.rs, tile_is_vacant, pl, obits)) {
                      E.ev(EV_BLOCKED, sm.id);   // RX_BLOCKED (:936-975): no reaction, the molecule goes on to the wall
                      pathway = -1;
                    }
                  } else obits = draw_orientation_bits(vpw, E.rs);
                }
                if (pathway >= 0) {
                  E.ev(EV_RXN | (uint32_t)pathway, (uint32_t)rc);
                  if (tr) { tr->rxn_class = rc; tr->rxn_pathway = pathway; tr->rxn_partner = sm.id; tr->t_event = abs_t; }
                  if (!apply) {
                    out.kind = MCX_OUT_REACTED; out.pos = c.pos; out.rxn_class = rc; out.pathway = pathway;
                    out.partner_index = occ_index; out.partner_id = sm.id; out.t_event = abs_t; out.orient_bits = obits;
                    out.coll_side = coll_orient;
                    fill_event(out); put_placement(out, pl);
                    return out;
                  }
                  bool a_destroyed = false, flip = false;
                  w.mols[index].pos = s.pos; w.mols[index].subpart = s.subpart; w.mols[index].cvi = s.cvi;
                  seq_apply_bimol(w, index, occ_index, rc, pathway, c.pos, abs_t, obits, a_destroyed, &flip, coll_orient, &pl);
                  if (a_destroyed) {  // collide_res == 1
                    destroyed = true; out.kind = MCX_OUT_REACTED; out.pos = c.pos; out.t_event = abs_t;
                    break;
                  }
                  last_hit_wall = c.wall;
                  if (flip) {
                    // RX_FLIP (:945-970): the kept initiator goes through the wall at the hit point and carries on
                    s.pos = c.pos; s.subpart = w.subpart_index(s.pos);
                    s.cvi = cv_cross(w, s.cvi, c.wall, c.type == COLL_WALL_FRONT);
                    const double t_smash = c.time;
                    remaining = remaining * (1.0 - t_smash);
                    elapsed += t_steps * t_smash;
                    t_steps *= (1.0 - t_smash);
                    if (t_steps < EPS) t_steps = EPS;
                    break;
                  }
                  // RX_A_OK with the initiator kept (:972-975): on to the wall's surface class, else it reflects
                }
              }
            }
          }
          if (action == MCX_SURF_STANDARD) {
            // collide_and_react_with_walls (:1034-1066): test_intersect (rxn_utils.inl:593-626) of the one matching class
            const mcx_rxn_class& wc = w.classes[wall_rc];
            const int pathway = E.test_intersect(wc, r_rate_factor);
            action = MCX_SURF_REFLECTIVE;   // no reaction: it reflects (:1066)
            if (pathway >= 0) {
              const mcx_pathway& pw = w.pathways[wc.first_pathway + pathway];
              const uint32_t obits = draw_orientation_bits(pw, E.rs);
              const double abs_t = elapsed + t_steps * c.time;
              E.ev(EV_WALLRXN | (uint32_t)c.type, c.wall);
              E.ev(EV_RXN | (uint32_t)pathway, (uint32_t)wall_rc);
              if (tr) { tr->rxn_class = wall_rc; tr->rxn_pathway = pathway; tr->t_event = abs_t; }
              if (!apply) {
                out.kind = MCX_OUT_WALLRXN; out.pos = c.pos; out.rxn_class = wall_rc; out.pathway = pathway; out.t_event = abs_t;
                out.orient_bits = obits; out.coll_side = c.type == COLL_WALL_FRONT ? 1 : -1; out.hit_wall = c.wall;
                fill_event(out);
                return out;
              }
              bool gone = false;
              w.mols[index].pos = s.pos; w.mols[index].subpart = s.subpart; w.mols[index].cvi = s.cvi;
              seq_apply_wallrxn(w, index, wall_rc, pathway, c.pos, abs_t, obits, c.wall, s.cvi, gone, c.type == COLL_WALL_FRONT ? 1 : -1);
              if (gone) { destroyed = true; out.kind = MCX_OUT_WALLRXN; out.pos = c.pos; out.t_event = abs_t; break; }
              // RX_FLIP -> WallRxnResult::TRANSPARENT, RX_A_OK -> REFLECT (:1048-1056)
              action = wallrxn_flips(wc, pw, obits) ? MCX_SURF_TRANSPARENT : MCX_SURF_REFLECTIVE;
            }
          }
          if (action == MCX_SURF_TRANSPARENT) {
            // cross_transparent_wall (:3007-3099), non-compartment branch
            E.ev(EV_TRANSP | (uint32_t)c.type, c.wall);
            w.stats.transparent++;
            s.pos = c.pos; s.subpart = w.subpart_index(s.pos);
            // update_counted_volume_id_when_crossing_wall (collision_utils.inl:1637-1694): a FRONT hit goes inside
            s.cvi = cv_cross(w, s.cvi, c.wall, c.type == COLL_WALL_FRONT);
            double t_smash = c.time;
            remaining = remaining * (1.0 - t_smash);
            elapsed += t_steps * t_smash;
            t_steps *= (1.0 - t_smash);
            if (t_steps < EPS) t_steps = EPS;
            last_hit_wall = c.wall;
          } else if (action == MCX_SURF_ABSORPTIVE) {
            // collide_and_react_with_walls (:991-1067) -> test_intersect (rxn_utils.inl:593-626):
            // max_fixed_p = GIGANTIC > scaling: two draws, always reacts -> outcome_intersect destroys
            double abs_t = elapsed + t_steps * c.time;
            (void)E.rs.dbl(); (void)E.rs.dbl();
            E.ev(EV_ABSORB | (uint32_t)c.type, c.wall);
            if (tr) tr->t_event = abs_t;
            if (!apply) {
              out.kind = MCX_OUT_ABSORBED; out.pos = c.pos; out.t_event = abs_t; fill_event(out);
              return out;
            }
            w.stats.absorptions++;
            destroyed = true; out.kind = MCX_OUT_ABSORBED; out.pos = c.pos; out.t_event = abs_t;
            seq_set_defunct(w, w.mols[index]);
          } else {
            // reflect_from_wall, collision_utils.inl:1711-1747
            E.ev(EV_WALL | (uint32_t)c.type, c.wall);
            w.stats.reflections++;
            elapsed += t_steps * c.time;
            s.pos = c.pos; s.subpart = w.subpart_index(s.pos);
            t_steps *= (1.0 - c.time);
            last_hit_wall = c.wall;
            remaining = reflected_displacement(remaining, wall.normal, c.time);
          }
          break;  // exactly one wall per trace, then re-trace (:566)
        }
      }
      if (!hit && !destroyed) {  // RayTraceState::FINISHED (ray_trace_vol :774-777)
        s.pos = s.pos + remaining;
        if (!w.in_this_partition(s.pos)) {  // :592-612
          w.err = "molecule " + std::to_string(m_id) + " escaped the partition";
          out.kind = MCX_OUT_NONE; out.pos = s.pos; return out;
        }
        s.subpart = w.subpart_index(s.pos);
      }
      if (++trace_guard > 100000) { w.err = "ray trace did not terminate"; break; }
    } while (hit && !destroyed);
  }
  if (destroyed) return out;

  // -- reschedule (diffuse_single_molecule :283-336)
  if ((sp.flags & MCX_SP_CAN_DIFFUSE) || can_ss) {
    s.t_now += max_time;
    if ((s.unimol_time != TIME_INVALID && s.unimol_time < t_end) || cmp_lt(s.t_now, t_end, EPS)) again = true;
    else {
      double r = round(s.t_now);
      if (cmp_eq(s.t_now, r, SQRT_EPS)) s.t_now = r;
    }
  } else {
    if (s.unimol_time != TIME_INVALID) {
      s.t_now = s.unimol_time;
      if (s.unimol_time < t_end) again = true;
    } else s.t_now = TIME_FOREVER;
  }
  out.kind = ((sp.flags & MCX_SP_CAN_DIFFUSE) || can_ss) ? MCX_OUT_MOVED : MCX_OUT_STATIC;
  out.pos = s.pos; fill_event(out);
  if (ss_fired) {
    // SNAPSHOT: the surface-surface reaction is a claiming event (the initiator, the partner it consumes, the tile it moved
    // to); a kept initiator has used up its step like any other mover and takes what is left of the iteration lazily
    out.kind = MCX_OUT_REACTED;
    out.flags = again ? (out.flags | MCX_MOL_PARTIAL) : (out.flags & ~MCX_MOL_PARTIAL);
    again = false;
    return out;
  }
  if (surf_tile_changed && !apply) {
    // SNAPSHOT: taking a new tile is a claiming event; the evaluation ends here and what is left of the iteration
    // is taken lazily next iteration (like a kept initiator)
    out.kind = MCX_OUT_SURFMOVE;
    out.flags = again ? (out.flags | MCX_MOL_PARTIAL) : (out.flags & ~MCX_MOL_PARTIAL);
    E.ev(EV_SURFMOVE | (s.tile & 0xFFFFFFu), s.wall);
    again = false;
  }
  return out;
}

static MolState load_state(const World& w, const Mol& m) {
  MolState s;
  s.pos = m.pos; s.subpart = w.subpart_index(m.pos);
  s.t_now = (m.flags & MCX_MOL_PARTIAL) ? m.diffusion_time : (double)w.iteration;
  s.flags = m.flags; s.unimol_time = m.unimol_rxn_time;
  s.created_wall = m.created_wall; s.created_tile = m.created_tile; s.cvi = m.cvi;
  s.wall = m.wall; s.tile = m.tile; s.u = m.u; s.v = m.v;
  return s;
}

// SNAPSHOT: all sub-steps of the iteration back to back (the product's kernel loop)
static Outcome evaluate_iteration(Eval& E, uint32_t index) {
  MolState s = load_state(E.w, E.w.mols[index]);
  Outcome o; bool again = false; int guard = 0;
  do {
    o = evaluate_substep(E, index, s, false, again);
    if (o.kind != MCX_OUT_MOVED && o.kind != MCX_OUT_STATIC) return o;
    s.created_wall = s.created_tile = MCX_NONE;  // the guard belongs to the first DiffuseAction only (:116-129)
  } while (again && ++guard < 1000);
  o.flags &= ~MCX_MOL_PARTIAL;
  return o;
}

// ---- sequential-mode world mutation --------------------------------------------------------------
static void seq_set_defunct(World& w, Mol& m) {  // Partition::set_molecule_as_defunct, partition.h:612-628
  if (m.flags & MCX_MOL_DEFUNCT) return;
  m.flags |= MCX_MOL_DEFUNCT;
  list_erase(w, m);
  // Grid::reset_molecule_tile: a recycled tile already belongs to the product
  if (m.wall != MCX_NONE && w.tiles[m.wall][m.tile] == m.id) w.tiles[m.wall][m.tile] = MCX_NONE;
  w.species_count[m.species]--;
}
static uint32_t seq_add_molecule(World& w, const ProductSpec& ps, double t) {  // add_volume_molecule / add_surface_molecule
  Mol n{};
  n.pos = ps.pos; n.id = w.next_id++; n.species = ps.species;
  n.flags = MCX_MOL_SCHEDULE_UNIMOL | MCX_MOL_PARTIAL | (ps.cvi_pending ? MCX_MOL_CVI_PENDING : 0u);
  n.diffusion_time = t; n.unimol_rxn_time = TIME_INVALID;
  n.subpart = w.subpart_index(ps.pos);
  n.wall = ps.wall; n.tile = ps.tile; n.orient = ps.orient; n.u = ps.u; n.v = ps.v;
  n.created_wall = ps.created_wall; n.created_tile = ps.created_tile; n.cvi = ps.cvi;
  w.mols.push_back(n);
  if (w.id_to_index.size() <= n.id) w.id_to_index.resize(n.id + 1, MCX_NONE);
  w.id_to_index[n.id] = (uint32_t)w.mols.size() - 1;
  w.sched_ids.push_back(n.id);
  list_insert(w, w.mols.back());
  if (n.wall != MCX_NONE) {
    if (w.tiles[n.wall].empty()) w.tiles[n.wall].assign(w.grids[n.wall].n_tiles, MCX_NONE);  // Wall::initialize_grid
    w.tiles[n.wall][n.tile] = n.id;  // Grid::set_molecule_tile (:2899)
  }
  w.species_count[ps.species]++;
  w.stats.products++;
  return n.id;
}
static std::vector<uint32_t>* g_new_actions = nullptr;  // new_diffuse_actions FIFO of the running step

// outcome_bimolecular / outcome_products_random (:1833-1895, :2446-2933): two volume reactants, or a volume
// initiator and the surface molecule it hit
// outcome_products_random :2513-2521: a volume initiator counts the reaction in its counted volume, a surface initiator
// on its wall
static inline void count_rxn_where(World& w, uint32_t rule, const Mol& initiator, uint32_t cvi) {
  if (initiator.wall == MCX_NONE) w.rxn_count_cv[rule * w.n_cv + cvi]++;
  else if (w.n_rs) w.rxn_count_rs[rule * w.n_rs + w.wall_rs[initiator.wall]]++;
}
static void seq_apply_bimol(World& w, uint32_t a_index, uint32_t b_index, int rc, int pathway, V3 pos, double t,
                            uint32_t orient_bits, bool& a_destroyed, bool* flip, int coll_side, const Placement* pl) {
  const mcx_rxn_class& c = w.classes[rc];
  const mcx_pathway& pw = w.pathways[c.first_pathway + pathway];
  w.rxn_count[pw.rxn_rule_id]++;
  count_rxn_where(w, pw.rxn_rule_id, w.mols[a_index], w.mols[a_index].cvi);
  w.stats.bimol_rxns++;
  // reactant ordering vs rule (:2541-2554)
  bool a_is_r0 = w.mols[a_index].species == c.reactants[0];
  bool keepA = (pw.keep_reactant_mask >> (a_is_r0 ? 0 : 1)) & 1;
  bool keepB = (pw.keep_reactant_mask >> (a_is_r0 ? 1 : 0)) & 1;
  const bool surf_rxn = w.mols[b_index].wall != MCX_NONE;
  const Mol surf_copy = w.mols[b_index];  // adding molecules may reallocate w.mols
  // tiles that are going to be reused are freed first (:2606-2615)
  if (surf_rxn && !keepB) w.tiles[surf_copy.wall][surf_copy.tile] = MCX_NONE;
  int placed = 0;
  for (uint32_t k = 0; k < pw.n_products; k++) {
    ProductSpec ps = product_spec(w, c, pw, k, pos, orient_bits, surf_rxn ? &surf_copy : nullptr, w.mols[a_index].cvi, coll_side);
    apply_placement(w, pl, placed, ps);
    uint32_t nid = seq_add_molecule(w, ps, t);
    if (cmp_lt(t, (double)w.iteration + 1, EPS) && g_new_actions) g_new_actions->push_back(nid);
  }
  if (!keepA) seq_set_defunct(w, w.mols[a_index]);
  if (!keepB) seq_set_defunct(w, w.mols[b_index]);
  a_destroyed = !keepA;
  // kept reactants of a surface reaction (:2689-2716; class order is (volume, surface)): a kept surface molecule takes
  // its product-side orientation; a kept volume initiator whose product-side orientation differs from the rule's
  // reactant-side one passes through the wall (RX_FLIP)
  if (flip) *flip = false;
  if (surf_rxn && c.kind == MCX_RXN_BIMOL_VOLSURF) {
    if (keepB) { const int o = kept_orientation(c, pw, 1, orient_bits, surf_copy.orient); if (o != 0) w.mols[b_index].orient = o; }
    if (keepA && flip) {
      const int o = kept_orientation(c, pw, 0, orient_bits, surf_copy.orient);
      *flip = o != 0 && c.reactant_orientation[0] != o;
    }
  }
}
// outcome_unimolecular (:2939-3003)
static void seq_apply_unimol(World& w, uint32_t index, int rc, int pathway, double t, uint32_t orient_bits, bool& destroyed,
                             const Placement* pl) {
  const mcx_rxn_class& c = w.classes[rc];
  const mcx_pathway& pw = w.pathways[c.first_pathway + pathway];
  w.rxn_count[pw.rxn_rule_id]++;
  count_rxn_where(w, pw.rxn_rule_id, w.mols[index], w.mols[index].cvi);
  w.stats.unimol_rxns++;
  V3 pos = w.mols[index].pos;
  bool keep = pw.keep_reactant_mask & 1;
  const bool surf_rxn = w.mols[index].wall != MCX_NONE;
  const Mol surf_copy = w.mols[index];
  if (surf_rxn && !keep) w.tiles[surf_copy.wall][surf_copy.tile] = MCX_NONE;
  int placed = 0;
  for (uint32_t k = 0; k < pw.n_products; k++) {
    ProductSpec ps = product_spec(w, c, pw, k, pos, orient_bits, surf_rxn ? &surf_copy : nullptr, w.mols[index].cvi);
    apply_placement(w, pl, placed, ps);
    uint32_t nid = seq_add_molecule(w, ps, t);
    if (cmp_lt(t, (double)w.iteration + 1, EPS) && g_new_actions) g_new_actions->push_back(nid);
  }
  if (!keep) seq_set_defunct(w, w.mols[index]);
  destroyed = !keep;
  if (surf_rxn && keep) { const int o = kept_orientation(c, pw, 0, orient_bits, surf_copy.orient); if (o != 0) w.mols[index].orient = o; }
}

// outcome_bimolecular (:1833-1895) -> outcome_products_random for two surface molecules
static void seq_apply_surfsurf(World& w, uint32_t a_index, uint32_t b_index, int rc, int pathway, double t, uint32_t bits, bool& a_destroyed,
                               const Placement* pl) {
  const mcx_rxn_class& c = w.classes[rc];
  const mcx_pathway& pw = w.pathways[c.first_pathway + pathway];
  w.rxn_count[pw.rxn_rule_id]++;
  count_rxn_where(w, pw.rxn_rule_id, w.mols[a_index], 0);
  w.stats.bimol_rxns++;
  const SurfSite init = site_of(w.mols[a_index]), partner = site_of(w.mols[b_index]);
  const bool a_is_r0 = init.species == c.reactants[0];
  const bool keepA = (pw.keep_reactant_mask >> (a_is_r0 ? 0 : 1)) & 1, keepB = (pw.keep_reactant_mask >> (a_is_r0 ? 1 : 0)) & 1;
  // tiles that are going to be reused are freed first (:2606-2615)
  if (!keepA) w.tiles[init.wall][init.tile] = MCX_NONE;
  if (!keepB) w.tiles[partner.wall][partner.tile] = MCX_NONE;
  std::vector<ProductSpec> prods;
  surfsurf_products(w, c, pw, init, partner, bits, prods);
  int placed = 0;
  for (ProductSpec& ps : prods) {
    apply_placement(w, pl, placed, ps);
    uint32_t nid = seq_add_molecule(w, ps, t);
    if (cmp_lt(t, (double)w.iteration + 1, EPS) && g_new_actions) g_new_actions->push_back(nid);
  }
  // kept reactants take their product-side orientation (:2689-2716)
  const SurfSite& r0 = a_is_r0 ? init : partner; const SurfSite& r1 = a_is_r0 ? partner : init;
  if (keepA) { const int o = surfsurf_kept_orientation(c, pw, a_is_r0 ? 0 : 1, bits, r0, r1); if (o != 0) w.mols[a_index].orient = o; }
  if (keepB) { const int o = surfsurf_kept_orientation(c, pw, a_is_r0 ? 1 : 0, bits, r0, r1); if (o != 0) w.mols[b_index].orient = o; }
  if (!keepA) seq_set_defunct(w, w.mols[a_index]);
  if (!keepB) seq_set_defunct(w, w.mols[b_index]);
  a_destroyed = !keepA;
}

// outcome_intersect (:1916-1988) of a Standard reaction with a reactive surface; the surface is always kept
static void seq_apply_wallrxn(World& w, uint32_t index, int rc, int pathway, V3 pos, double t, uint32_t orient_bits, uint32_t wall,
                              uint32_t cvi, bool& destroyed, int coll_side) {
  const mcx_rxn_class& c = w.classes[rc];
  const mcx_pathway& pw = w.pathways[c.first_pathway + pathway];
  w.rxn_count[pw.rxn_rule_id]++;
  count_rxn_where(w, pw.rxn_rule_id, w.mols[index], cvi);
  w.stats.bimol_rxns++;
  for (uint32_t k = 0; k < pw.n_products; k++) {
    ProductSpec ps = wall_product_spec(w, pw, k, pos, orient_bits, wall, cvi, coll_side);
    uint32_t nid = seq_add_molecule(w, ps, t);
    if (cmp_lt(t, (double)w.iteration + 1, EPS) && g_new_actions) g_new_actions->push_back(nid);
  }
  const bool keep = pw.keep_reactant_mask & 1u;
  if (!keep) seq_set_defunct(w, w.mols[index]);
  destroyed = !keep;
}

static void trace_begin(World& w, Eval& E, const Mol& m) {
  if (!w.tracing) return;
  if (w.trace.size() <= m.id) {
    mcx_trace_rec z{}; z.rxn_class = z.rxn_pathway = z.rxn_partner = MCX_NONE;
    w.trace.resize(m.id + 1, z);
  }
  mcx_trace_rec& t = w.trace[m.id];
  if (!E.snapshot && t.rounds > 0) {  // sequential mode: later sub-step of the same iteration accumulates
    t.rounds++; E.tr = &t; E.h = t.event_hash; E.words_base = t.n_words;
    return;
  }
  uint32_t rounds = t.rounds;
  memset(&t, 0, sizeof(t));
  t.id = m.id; t.rxn_class = t.rxn_pathway = t.rxn_partner = MCX_NONE; t.rounds = rounds + 1;
  E.tr = &t;
}
static void trace_end(World& w, Eval& E, const Outcome& o) {
  (void)w;
  if (!E.tr) return;
  E.tr->outcome = o.kind; E.tr->n_words = E.words_base + E.rs.used; E.tr->event_hash = E.h;
  E.tr->pos[0] = o.pos.x; E.tr->pos[1] = o.pos.y; E.tr->pos[2] = o.pos.z;
}

// ---- SEQUENTIAL iteration: DiffuseReactEvent::step + diffuse_molecules (:53-161) -----------------------
static void step_sequential(World& w) {
  std::vector<uint32_t> ready;  // get_molecules_ready_for_diffusion, partition.h:232-246
  double t_end = (double)w.iteration + 1;
  for (uint32_t id : w.sched_ids) {
    const Mol& m = w.mols[w.id_to_index[id]];
    if (!(m.flags & MCX_MOL_DEFUNCT) && cmp_lt((m.flags & MCX_MOL_PARTIAL) ? m.diffusion_time : (double)w.iteration, t_end, EPS))
      { ready.push_back(id); if (w.species[m.species].flags & MCX_SP_CAN_DIFFUSE) w.stats.molecule_steps++; }
  }
  std::vector<uint32_t> actions;
  g_new_actions = &actions;
  if (w.record_tape) {
    w.tape_words.clear();
    w.tape_off.assign(w.next_id, 0); w.tape_len.assign(w.next_id, 0);
  }
  auto run_one = [&](uint32_t id) {
    uint32_t index = w.id_to_index[id];
    if (w.mols[index].flags & MCX_MOL_DEFUNCT) return;
    WordSource rs; rs.kind = WordSource::ISAAC; rs.isaac = &w.rng;
    size_t tape_start = w.tape_words.size();
    if (w.record_tape) rs.record = &w.tape_words;
    Eval E(w, rs);
    trace_begin(w, E, w.mols[index]);
    MolState s = load_state(w, w.mols[index]);
    bool again = false;
    Outcome o = evaluate_substep(E, index, s, true, again);
    trace_end(w, E, o);
    if (w.record_tape && id < w.tape_off.size()) {
      if (w.tape_len[id] == 0) { w.tape_off[id] = tape_start; w.tape_len[id] = rs.used; }
      else w.tape_len[id] = 0xFFFFFFFFu;  // split step: words are not contiguous in the global stream
    }
    Mol& m = w.mols[w.id_to_index[id]];
    m.created_wall = m.created_tile = MCX_NONE;  // the pair travels with the first DiffuseAction only
    if (o.kind == MCX_OUT_MOVED || o.kind == MCX_OUT_STATIC) {
      m.pos = o.pos; m.subpart = w.subpart_index(o.pos);
      m.flags = again ? (o.flags | MCX_MOL_PARTIAL) : (o.flags & ~MCX_MOL_PARTIAL);
      m.diffusion_time = o.t_now; m.unimol_rxn_time = o.unimol_time; m.cvi = o.cvi;
      // Partition::update_molecule_reactants_map, partition.h:406-418
      if (m.subpart != m.reg_subpart) { list_erase(w, m); list_insert(w, m); }
      if (again) actions.push_back(id);  // new_diffuse_actions (:305-308)
    }
  };
  // phase 1: existing molecules in schedulable order; phase 3: FIFO of products
  // (phase 2, delayed releases, does not occur: releases happen at iteration starts)
  for (uint32_t id : ready) {
    const Mol& m = w.mols[w.id_to_index[id]];
    if (m.flags & MCX_MOL_PARTIAL) { actions.push_back(id); continue; }
    run_one(id);
  }
  for (size_t i = 0; i < actions.size(); i++) run_one(actions[i]);
  g_new_actions = nullptr;
  w.iteration++;
}

// DefragmentationEvent::step, defragmentation_event.cpp:31-114
static void defragment(World& w) {
  std::vector<Mol> keep; keep.reserve(w.mols.size());
  for (auto& m : w.mols) if (!(m.flags & MCX_MOL_DEFUNCT)) keep.push_back(m);
  w.mols.swap(keep);
  std::fill(w.id_to_index.begin(), w.id_to_index.end(), MCX_NONE);
  for (uint32_t i = 0; i < w.mols.size(); i++) w.id_to_index[w.mols[i].id] = i;
  std::vector<uint32_t> s; s.reserve(w.mols.size());
  for (uint32_t id : w.sched_ids) if (id < w.id_to_index.size() && w.id_to_index[id] != MCX_NONE) s.push_back(id);
  w.sched_ids.swap(s);
}

// ---- SNAPSHOT iteration: the parallel semantics of the product ---------------------------
struct SnapStreams { int kind; const uint32_t* words; uint64_t n_words; const uint64_t* off; uint64_t n_ids; };

static void step_snapshot(World& w, const SnapStreams& st) {
  const size_t n0 = w.mols.size();
  const uint32_t max_rounds = w.cfg.max_resolve_rounds ? w.cfg.max_resolve_rounds : 8;
  std::vector<uint8_t> dead(n0, 0);
  for (size_t i = 0; i < n0; i++) {
    dead[i] = (w.mols[i].flags & MCX_MOL_DEFUNCT) ? 1 : 0;
    if (!w.sample_stride && !dead[i] && (w.species[w.mols[i].species].flags & MCX_SP_CAN_DIFFUSE)) w.stats.molecule_steps++;
  }
  std::vector<Outcome> outs(n0);
  std::vector<uint32_t> claim(n0, MCX_NONE);
  // surface tiles: claims of the movers of the running round, and the tiles claimed in earlier rounds
  std::unordered_map<uint32_t, uint32_t> tile_claim;
  std::vector<uint8_t> tile_claimed(w.tile_start.empty() ? 1 : w.tile_start.back() + 1, 0);
  std::vector<uint32_t> tiles_of_round;
  std::vector<uint32_t> pending;
  struct NewMol { ProductSpec ps; double t; uint32_t id; };
  std::vector<NewMol> born;
  std::vector<std::pair<uint32_t, int>> orient_updates;  // kept initiators of surface-surface reactions: (index, new orientation)

  auto eval_one = [&](uint32_t i, bool forced) {
    const Mol& m = w.mols[i];
    WordSource rs;
    if (st.kind == MCX_RNG_TAPE) {
      rs.kind = WordSource::TAPE;
      uint64_t off = m.id < st.n_ids ? st.off[m.id] : st.n_words;
      rs.tape = st.words + off; rs.tape_len = st.n_words - off;
    } else {
      rs.kind = WordSource::PHILOX; rs.seed = w.cfg.seed; rs.mol_id = m.id; rs.iteration = w.iteration;
    }
    Eval E(w, rs);
    E.snapshot = true; E.dead = &dead; E.no_partners = forced;
    E.tile_claimed = &tile_claimed; E.tile_start = &w.tile_start;
    trace_begin(w, E, m);
    outs[i] = evaluate_iteration(E, i);
    trace_end(w, E, outs[i]);
  };
  auto is_claiming = [](const Outcome& o) {
    return o.kind == MCX_OUT_REACTED || o.kind == MCX_OUT_ABSORBED || o.kind == MCX_OUT_UNIMOL || o.kind == MCX_OUT_SURFMOVE ||
           o.kind == MCX_OUT_WALLRXN;
  };
  auto gtile_of = [&](const Outcome& o) { return w.tile_start[o.wall] + o.tile; };
  // does the claiming event consume the partner? (kept reactants are not claimed)
  auto partner_consumed = [&](uint32_t i) {
    const Outcome& o = outs[i];
    if (o.kind != MCX_OUT_REACTED) return false;
    const mcx_rxn_class& c = w.classes[o.rxn_class];
    const mcx_pathway& pw = w.pathways[c.first_pathway + o.pathway];
    bool a_is_r0 = w.mols[i].species == c.reactants[0];
    return !((pw.keep_reactant_mask >> (a_is_r0 ? 1 : 0)) & 1);
  };
  // a surface molecule that only takes a new tile claims itself and the tile WEAKLY: a reaction that consumes it, or
  // that needs the tile for its initiator, comes first (see mcx_kernels.cu: weak_key)
  auto prio_of = [&](uint32_t i) { return outs[i].kind == MCX_OUT_SURFMOVE ? (w.mols[i].id | 0x80000000u) : w.mols[i].id; };
  auto make_claims = [&](uint32_t i) {
    uint32_t prio = prio_of(i);
    claim[i] = std::min(claim[i], prio);
    if (partner_consumed(i)) { uint32_t j = outs[i].partner_index; claim[j] = std::min(claim[j], prio); }
    auto claim_tile = [&](uint32_t gt) {
      auto it = tile_claim.find(gt);
      if (it == tile_claim.end()) tile_claim[gt] = prio; else it->second = std::min(it->second, prio);
      tiles_of_round.push_back(gt);
    };
    if (outs[i].kind == MCX_OUT_SURFMOVE || (outs[i].kind == MCX_OUT_REACTED && outs[i].surf_moved)) claim_tile(gtile_of(outs[i]));
    // products on vacant neighbour tiles claim them
    for (int k = 0; k < outs[i].pl_n; k++)
      if ((outs[i].pl_vacant >> k) & 1u) claim_tile(w.tile_start[outs[i].pl_wall[k]] + outs[i].pl_tile[k]);
  };
  auto commit = [&](uint32_t i) {  // accepted claiming event
    Outcome& o = outs[i];
    const Mol& m = w.mols[i];
    if (o.kind == MCX_OUT_ABSORBED) { dead[i] = 1; w.stats.absorptions++; w.species_count[m.species]--; return; }
    if (o.kind == MCX_OUT_SURFMOVE) { o.kind = MCX_OUT_MOVED; return; }  // stays alive on its new tile
    if (o.kind == MCX_OUT_WALLRXN) {
      // Standard reaction with a reactive surface (outcome_intersect :1916-1988): products in front of / behind the wall;
      // the molecule is consumed, or kept — then it stays on its side or crosses (RX_FLIP), waiting 2*16*EPS off the wall,
      // guarded against an immediate surface-molecule reaction on the tile under the hit point, for the rest of its step
      const mcx_rxn_class& c = w.classes[o.rxn_class];
      const mcx_pathway& pw = w.pathways[c.first_pathway + o.pathway];
      w.rxn_count[pw.rxn_rule_id]++;
      count_rxn_where(w, pw.rxn_rule_id, m, o.cvi);
      w.stats.bimol_rxns++;
      const bool keep = pw.keep_reactant_mask & 1u;
      const V3 hit = o.pos;
      if (!keep) { dead[i] = 1; w.species_count[m.species]--; }
      for (uint32_t k = 0; k < pw.n_products; k++) {
        NewMol nm; nm.ps = wall_product_spec(w, pw, k, hit, o.orient_bits, o.hit_wall, o.cvi, o.coll_side); nm.t = o.t_event;
        nm.id = (k == 0 && !keep) ? m.id : MCX_NONE;
        born.push_back(nm);
        w.species_count[nm.ps.species]++;
        w.stats.products++;
      }
      if (keep) {
        const bool flip = wallrxn_flips(c, pw, o.orient_bits);
        const Wall& f = w.walls[o.hit_wall];
        const int side = flip ? -o.coll_side : o.coll_side;
        if (flip) o.cvi = cv_cross(w, o.cvi, o.hit_wall, o.coll_side > 0);
        const double bump = (side > 0) ? 16 * POS_EPS : -16 * POS_EPS;
        o.pos = hit + V3{(2 * bump) * f.normal.x, (2 * bump) * f.normal.y, (2 * bump) * f.normal.z};
        o.created_wall = o.hit_wall; o.created_tile = KEPT_AT_WALL;
        o.kind = MCX_OUT_MOVED; o.t_now = o.t_event; o.flags |= MCX_MOL_PARTIAL;
      }
      return;
    }
    const mcx_rxn_class& c = w.classes[o.rxn_class];
    const mcx_pathway& pw = w.pathways[c.first_pathway + o.pathway];
    w.rxn_count[pw.rxn_rule_id]++;
    if (o.kind == MCX_OUT_REACTED && c.kind == MCX_RXN_BIMOL_SURFSURF) {
      // surface-surface reaction (react_2D_all_neighbors -> outcome_bimolecular): counted on the wall the initiator moved to
      if (w.n_rs) w.rxn_count_rs[pw.rxn_rule_id * w.n_rs + w.wall_rs[o.wall]]++;
      w.stats.bimol_rxns++;
      const uint32_t j = o.partner_index;
      const SurfSite init{o.wall, o.tile, o.u, o.v, m.orient, m.species}, partner = site_of(w.mols[j]);
      const bool a_is_r0 = m.species == c.reactants[0];
      const bool keepA = (pw.keep_reactant_mask >> (a_is_r0 ? 0 : 1)) & 1, keepB = (pw.keep_reactant_mask >> (a_is_r0 ? 1 : 0)) & 1;
      uint32_t reuse[2]; int n_reuse = 0;
      if (!keepA) { dead[i] = 1; w.species_count[m.species]--; reuse[n_reuse++] = m.id; }
      if (!keepB) { dead[j] = 1; w.species_count[w.mols[j].species]--; reuse[n_reuse++] = w.mols[j].id; }
      std::vector<ProductSpec> prods;
      surfsurf_products(w, c, pw, init, partner, o.orient_bits, prods);
      const Placement pl = get_placement(o);
      int placed = 0;
      for (uint32_t k = 0; k < prods.size(); k++) {
        apply_placement(w, &pl, placed, prods[k]);
        NewMol nm; nm.ps = prods[k]; nm.t = o.t_event; nm.id = (int)k < n_reuse ? reuse[k] : MCX_NONE;
        born.push_back(nm);
        w.species_count[nm.ps.species]++;
        w.stats.products++;
      }
      if (keepA) {  // stays on the tile it moved to, its step used up; takes its product-side orientation (:2706-2709)
        o.kind = MCX_OUT_MOVED;
        const SurfSite& r0 = a_is_r0 ? init : partner; const SurfSite& r1 = a_is_r0 ? partner : init;
        const int ko = surfsurf_kept_orientation(c, pw, a_is_r0 ? 0 : 1, o.orient_bits, r0, r1);
        if (ko != 0) orient_updates.push_back({i, ko});  // visible from the next snapshot on, like its new tile
      }
      return;
    }
    count_rxn_where(w, pw.rxn_rule_id, m, o.cvi);
    bool keepA, keepB = true;
    uint32_t reuse[2]; int n_reuse = 0;
    if (o.kind == MCX_OUT_REACTED) {
      w.stats.bimol_rxns++;
      bool a_is_r0 = m.species == c.reactants[0];
      keepA = (pw.keep_reactant_mask >> (a_is_r0 ? 0 : 1)) & 1;
      keepB = (pw.keep_reactant_mask >> (a_is_r0 ? 1 : 0)) & 1;
    } else {
      w.stats.unimol_rxns++;
      keepA = pw.keep_reactant_mask & 1;
    }
    if (!keepA) { dead[i] = 1; w.species_count[m.species]--; reuse[n_reuse++] = m.id; }
    if (!keepB) { uint32_t j = o.partner_index; dead[j] = 1; w.species_count[w.mols[j].species]--; reuse[n_reuse++] = w.mols[j].id; }
    // the surface reactant of the event, if any: the partner of a volume initiator, or the initiator itself
    const Mol* surf = nullptr;
    if (o.kind == MCX_OUT_REACTED && w.mols[o.partner_index].wall != MCX_NONE) surf = &w.mols[o.partner_index];
    else if (o.kind == MCX_OUT_UNIMOL && m.wall != MCX_NONE) surf = &m;
    // product ids: consumed reactants' ids are recycled first (initiator, then partner), then fresh ids
    const Placement gpl = get_placement(o);
    int placed = 0;
    for (uint32_t k = 0; k < pw.n_products; k++) {
      NewMol nm; nm.ps = product_spec(w, c, pw, k, o.pos, o.orient_bits, surf, o.cvi, o.kind == MCX_OUT_REACTED ? o.coll_side : 0); nm.t = o.t_event;
      apply_placement(w, &gpl, placed, nm.ps);
      nm.id = (int)k < n_reuse ? reuse[k] : MCX_NONE;
      born.push_back(nm);
      w.species_count[nm.ps.species]++;
      w.stats.products++;
    }
    if (keepA) {
      // kept initiator: evaluation stopped at the event; it resumes from there next iteration
      // (catalytic initiators are rare; documented deviation: remaining sub-step is taken lazily)
      o.kind = MCX_OUT_MOVED; o.t_now = o.t_event; o.flags |= MCX_MOL_PARTIAL;
      if (c.kind == MCX_RXN_UNIMOL) { o.flags |= MCX_MOL_SCHEDULE_UNIMOL; o.unimol_time = TIME_INVALID; }
      if (c.kind == MCX_RXN_UNIMOL && surf) {  // a kept surface molecule takes its product-side orientation (:2706-2709)
        const int ko = kept_orientation(c, pw, 0, o.orient_bits, surf->orient);
        if (ko != 0) w.mols[i].orient = ko;
      }
      if (c.kind == MCX_RXN_BIMOL_VOLSURF && surf) {
        // kept volume initiator of a surface reaction (:945-975, :2694-2716): it stays on the side it came from, or
        // passes through the wall when its product-side orientation differs from the reactant-side one (RX_FLIP).
        // SNAPSHOT: like a volume product of the reaction it waits 2*16*EPS off the wall on that side, guarded
        // against rebinding on the same tile once, and takes the rest of its step next iteration
        const int ko = kept_orientation(c, pw, 0, o.orient_bits, surf->orient);
        const bool flip = ko != 0 && c.reactant_orientation[0] != ko;
        const Wall& f = w.walls[surf->wall];
        const int side = flip ? -o.coll_side : o.coll_side;
        if (flip) o.cvi = cv_cross(w, o.cvi, surf->wall, o.coll_side > 0);  // update_counted_volume_id_when_crossing_wall
        const double bump = (side > 0) ? 16 * POS_EPS : -16 * POS_EPS;
        o.pos = o.pos + V3{(2 * bump) * f.normal.x, (2 * bump) * f.normal.y, (2 * bump) * f.normal.z};
        o.created_wall = surf->wall; o.created_tile = KEPT_AT_WALL;
      }
    }
  };

  if (w.sample_stride) {  // a sample of first evaluations against the snapshot (test_gpu_fullsize.py): no claims, no commits
    for (uint32_t i = 0; i < n0; i++)
      if (!dead[i] && w.mols[i].id % w.sample_stride == w.sample_offset) eval_one(i, false);
    return;
  }
  // round 0: everyone
  for (uint32_t i = 0; i < n0; i++) {
    if (dead[i]) { outs[i].kind = MCX_OUT_NONE; continue; }
    eval_one(i, false);
    if (is_claiming(outs[i])) { pending.push_back(i); make_claims(i); }
  }
  for (uint32_t round = 0; round < max_rounds && !pending.empty(); round++) {
    // resolve: decisions use the claims as they stand; commits become visible afterwards
    std::vector<uint32_t> still, accepted;
    for (uint32_t i : pending) {
      uint32_t prio = prio_of(i);
      bool ok = claim[i] == prio;
      if (ok && partner_consumed(i)) ok = claim[outs[i].partner_index] == prio;
      if (ok && (outs[i].kind == MCX_OUT_SURFMOVE || (outs[i].kind == MCX_OUT_REACTED && outs[i].surf_moved)))
        ok = tile_claim[gtile_of(outs[i])] == prio;
      for (int k = 0; ok && k < outs[i].pl_n; k++)
        if ((outs[i].pl_vacant >> k) & 1u) ok = tile_claim[w.tile_start[outs[i].pl_wall[k]] + outs[i].pl_tile[k]] == prio;
      (ok ? accepted : still).push_back(i);
    }
    // tiles claimed in this round stay unavailable for the movers of later rounds, whoever won them
    for (uint32_t gt : tiles_of_round) tile_claimed[gt] = 1;
    tiles_of_round.clear(); tile_claim.clear();
    for (uint32_t i : accepted) {
      commit(i);
      if (!dead[i]) claim[i] = MCX_NONE;  // a survivor (kept initiator, surface mover) can be claimed again in later rounds
    }
    // reset claims of the losers, then re-evaluate them against the updated dead set
    for (uint32_t i : still) { claim[i] = MCX_NONE; if (partner_consumed(i)) claim[outs[i].partner_index] = MCX_NONE; }
    pending.clear();
    bool last = round + 1 == max_rounds;
    for (uint32_t i : still) {
      if (dead[i]) { outs[i].kind = MCX_OUT_CONSUMED; if (w.tracing) w.trace[w.mols[i].id].outcome = MCX_OUT_CONSUMED; continue; }
      w.stats.retries++;
      eval_one(i, last);
      if (last) w.stats.unresolved++;
      if (is_claiming(outs[i])) {
        if (last) commit(i);  // forced pass: only self-claims (absorption, unimolecular) can occur
        else { pending.push_back(i); make_claims(i); }
      }
    }
    // claims of re-evaluated molecules compete only among themselves and with accepted ones (consumed)
  }
  for (auto& u : orient_updates) w.mols[u.first].orient = u.second;
  // finalize survivors
  for (uint32_t i = 0; i < n0; i++) {
    if (dead[i]) {
      if (!(w.mols[i].flags & MCX_MOL_DEFUNCT)) {
        w.mols[i].flags |= MCX_MOL_DEFUNCT;
        if (w.tracing && outs[i].kind != MCX_OUT_REACTED && outs[i].kind != MCX_OUT_ABSORBED && outs[i].kind != MCX_OUT_UNIMOL &&
            outs[i].kind != MCX_OUT_WALLRXN)
          w.trace[w.mols[i].id].outcome = MCX_OUT_CONSUMED;
      }
      continue;
    }
    const Outcome& o = outs[i];
    Mol& m = w.mols[i];
    m.pos = o.pos; m.flags = o.flags; m.diffusion_time = o.t_now; m.unimol_rxn_time = o.unimol_time;
    m.subpart = w.subpart_index(m.pos);
    m.created_wall = o.created_wall; m.created_tile = o.created_tile; m.cvi = o.cvi;
    if (m.wall != MCX_NONE) { m.wall = o.wall; m.tile = o.tile; m.u = o.u; m.v = o.v; }
  }
  // compaction + products (the product's per-iteration sort does both)
  std::vector<Mol> keep; keep.reserve(w.mols.size() + born.size());
  for (auto& m : w.mols) if (!(m.flags & MCX_MOL_DEFUNCT)) keep.push_back(m);
  for (auto& nm : born) {
    Mol n{};
    n.pos = nm.ps.pos; n.species = nm.ps.species; n.id = nm.id != MCX_NONE ? nm.id : w.next_id++;
    n.flags = MCX_MOL_SCHEDULE_UNIMOL | MCX_MOL_PARTIAL | (nm.ps.cvi_pending ? MCX_MOL_CVI_PENDING : 0u);
    n.diffusion_time = nm.t; n.unimol_rxn_time = TIME_INVALID; n.subpart = w.subpart_index(nm.ps.pos);
    n.wall = nm.ps.wall; n.tile = nm.ps.tile; n.orient = nm.ps.orient; n.u = nm.ps.u; n.v = nm.ps.v;
    n.created_wall = nm.ps.created_wall; n.created_tile = nm.ps.created_tile; n.cvi = nm.ps.cvi;
    keep.push_back(n);
  }
  w.mols.swap(keep);
  // tile occupancy of the next snapshot (the product's scatter rebuilds it the same way)
  for (auto& t : w.tiles) std::fill(t.begin(), t.end(), MCX_NONE);
  for (auto& m : w.mols) if (m.wall != MCX_NONE) {
    if (w.tiles[m.wall].empty()) w.tiles[m.wall].assign(w.grids[m.wall].n_tiles, MCX_NONE);
    w.tiles[m.wall][m.tile] = m.id;
  }
  w.lists.clear();
  if (w.id_to_index.size() < w.next_id) w.id_to_index.resize(w.next_id, MCX_NONE);
  std::fill(w.id_to_index.begin(), w.id_to_index.end(), MCX_NONE);
  w.sched_ids.clear();
  for (uint32_t i = 0; i < w.mols.size(); i++) {
    w.id_to_index[w.mols[i].id] = i; w.sched_ids.push_back(w.mols[i].id);
    list_insert(w, w.mols[i]);
  }
  w.iteration++;
}

}  // namespace orc

// ================================================================================================
// C interface for ctypes (tests / bench only)
// ================================================================================================
using namespace orc;
extern "C" {

void* orc_create(const mcx_config* cfg) {
  World* w = new World();
  w->cfg = *cfg;
  w->n_sp = cfg->num_subparts_per_edge;
  w->sp_len = cfg->partition_edge_length / cfg->num_subparts_per_edge;  // simulation_config.cpp:48
  w->sp_rcp = 1.0 / w->sp_len;                                         // :63
  w->rng.init((uint32_t)cfg->seed);
  w->iteration = cfg->initial_iteration;
  w->walls_per_subpart.assign((size_t)w->n_sp * w->n_sp * w->n_sp, {});
  return w;
}
void orc_destroy(void* h) { delete (World*)h; }
const char* orc_last_error(void* h) { return ((World*)h)->err.c_str(); }

int orc_set_geometry(void* h, const double* v, uint64_t nv, const uint32_t* tri, uint64_t nw,
                     const uint32_t* surf_class, const uint32_t* object) {
  World& w = *(World*)h;
  w.verts.resize(nv);
  for (uint64_t i = 0; i < nv; i++) w.verts[i] = {v[3 * i], v[3 * i + 1], v[3 * i + 2]};
  w.walls.resize(nw);
  for (uint64_t i = 0; i < nw; i++) {
    Wall& f = w.walls[i];
    f.vi[0] = tri[3 * i]; f.vi[1] = tri[3 * i + 1]; f.vi[2] = tri[3 * i + 2];
    f.surf_class = surf_class ? surf_class[i] : MCX_NONE;
    f.object = object ? object[i] : 0;
    init_wall_constants(w, f);
  }
  w.grids.resize(nw);
  for (uint64_t i = 0; i < nw; i++) grid_init(w, w.walls[i], w.grids[i]);
  w.tiles.assign(nw, {});  // a new geometry starts without grids
  w.tiles.assign(nw, {});
  w.tile_start.assign(nw + 1, 0);
  for (uint64_t i = 0; i < nw; i++) w.tile_start[i + 1] = w.tile_start[i] + w.grids[i].n_tiles;
  // Geometry objects are contiguous runs of walls with the same object id (one object when none is given)
  for (uint64_t i = 0; i < nw;) {
    uint64_t j = i;
    while (j < nw && w.walls[j].object == w.walls[i].object) j++;
    init_edges(w, (uint32_t)i, (uint32_t)(j - i));
    i = j;
  }
  finalize_walls(w);
  return 0;
}
int orc_set_species(void* h, const mcx_species* s, uint32_t n) {
  World& w = *(World*)h; w.species.assign(s, s + n); build_lookups(w); return 0;
}
int orc_set_reactions(void* h, const mcx_rxn_class* c, uint32_t nc, const mcx_pathway* p, uint32_t np) {
  World& w = *(World*)h; w.classes.assign(c, c + nc); w.pathways.assign(p, p + np); build_lookups(w);
  for (const mcx_rxn_class& rc : w.classes)
    if (rc.kind == MCX_RXN_BIMOL_SURFSURF && !w.wall_border.empty()) {
      w.err = "surface-surface classes together with region borders are not supported (restricted regions of the neighbour search)"; return -1;
    }
  for (const mcx_rxn_class& rc : w.classes)
    for (uint32_t q = 0; q < rc.n_pathways; q++) {
      const mcx_pathway& pw = w.pathways[rc.first_pathway + q];
      if (pathway_is_general(w, rc, pw)) {
        if (const char* why = general_pathway_problem(w, rc, pw)) { w.err = why; return -1; }
      } else if (rc.kind == MCX_RXN_BIMOL_SURFSURF) {
        if (const char* why = surfsurf_pathway_problem(w, pw)) { w.err = why; return -1; }
      }
    }
  return 0;
}
int orc_set_surface_classes(void* h, const mcx_surf_class_rxn* r, uint32_t n) {
  World& w = *(World*)h; w.surf_rules.assign(r, r + n); return 0;
}
int orc_set_counted_volumes(void* h, uint32_t n_cv, const uint8_t* front, const uint8_t* back) {
  World& w = *(World*)h;
  w.n_cv = n_cv ? n_cv : 1;
  for (size_t i = 0; i < w.walls.size(); i++) { w.walls[i].cv_front = front[i]; w.walls[i].cv_back = back[i]; }
  build_lookups(w);
  return 0;
}
int orc_set_counted_volume_objects(void* h, const uint32_t* cv_object_mask, uint32_t intersecting_objects) {
  World& w = *(World*)h;
  w.cv_mask.clear(); w.cv_xor = 0; w.cv_all = 0;
  if (!cv_object_mask) return 0;
  w.cv_mask.assign(cv_object_mask, cv_object_mask + w.n_cv);
  w.cv_xor = intersecting_objects;
  for (uint32_t m_ : w.cv_mask) w.cv_all |= m_;
  return 0;
}
int orc_set_region_borders(void* h, const uint8_t* wall_edge_border) {
  World& w = *(World*)h;
  if (wall_edge_border) {
    for (uint8_t c : w.can_surf_surf)
      if (c) { w.err = "region borders together with surface-surface classes are not supported (restricted regions of the neighbour search)"; return -1; }
    w.wall_border.assign(wall_edge_border, wall_edge_border + w.walls.size());
  } else w.wall_border.clear();
  return 0;
}
int orc_set_surface_regions(void* h, uint32_t n_region_sets, const uint8_t* wall_region_set) {
  World& w = *(World*)h;
  w.n_rs = n_region_sets;
  w.wall_rs.assign(wall_region_set, wall_region_set + w.walls.size());
  uint32_t max_rule = 1;
  for (auto& pw : w.pathways) max_rule = std::max(max_rule, pw.rxn_rule_id + 1);
  w.rxn_count_rs.assign((size_t)max_rule * w.n_rs, 0);
  return 0;
}
int orc_counts_by_surface_region(void* h, uint64_t* mol_counts, uint64_t* rxn_counts) {
  World& w = *(World*)h;
  if (!w.n_rs) return MCX_ERR_STATE;
  if (mol_counts) {
    std::fill(mol_counts, mol_counts + w.species.size() * w.n_rs, 0);
    for (auto& m : w.mols) if (!(m.flags & MCX_MOL_DEFUNCT) && m.wall != MCX_NONE) mol_counts[m.species * w.n_rs + w.wall_rs[m.wall]]++;
  }
  if (rxn_counts) std::copy(w.rxn_count_rs.begin(), w.rxn_count_rs.end(), rxn_counts);
  return 0;
}
int orc_counts_by_volume(void* h, uint64_t* mol_counts, uint64_t* rxn_counts) {
  World& w = *(World*)h;
  if (mol_counts) {
    std::fill(mol_counts, mol_counts + w.species.size() * w.n_cv, 0);
    for (auto& m : w.mols) if (!(m.flags & MCX_MOL_DEFUNCT)) mol_counts[m.species * w.n_cv + m.cvi]++;
  }
  if (rxn_counts) std::copy(w.rxn_count_cv.begin(), w.rxn_count_cv.end(), rxn_counts);
  return 0;
}
int orc_upload_molecules(void* h, const mcx_mol_soa* s) {
  World& w = *(World*)h;
  w.mols.clear(); w.lists.clear(); w.sched_ids.clear(); w.id_to_index.clear();
  // the walls keep their grids (Wall::has_initialized_grid) across uploads like the reference's walls do when the host adds or
  // removes molecules; only the occupancy starts over
  if (w.tiles.size() != w.walls.size()) w.tiles.assign(w.walls.size(), {});
  for (auto& tl : w.tiles) std::fill(tl.begin(), tl.end(), MCX_NONE);
  std::fill(w.species_count.begin(), w.species_count.end(), 0);
  uint32_t max_id = 0;
  for (uint64_t i = 0; i < s->n; i++) max_id = std::max(max_id, s->id[i]);
  w.id_to_index.assign((size_t)max_id + 1, MCX_NONE);
  w.mols.reserve(s->n);
  for (uint64_t i = 0; i < s->n; i++) {
    Mol m{};
    m.pos = {s->x[i], s->y[i], s->z[i]}; m.id = s->id[i]; m.species = s->species[i];
    m.flags = s->flags ? s->flags[i] : 0;
    m.diffusion_time = s->diffusion_time ? s->diffusion_time[i] : (double)w.iteration;
    m.unimol_rxn_time = s->unimol_rxn_time ? s->unimol_rxn_time[i] : TIME_INVALID;
    m.cvi = s->counted_volume ? s->counted_volume[i] : 0;
    if (m.cvi >= w.n_cv) { w.err = "counted volume index out of range"; return MCX_ERR_INVALID_ARG; }
    if (s->wall && s->wall[i] != MCX_NONE) {  // Partition::add_surface_molecule + Grid::set_molecule_tile
      m.wall = s->wall[i]; m.tile = s->tile[i]; m.orient = s->orientation[i]; m.u = s->u[i]; m.v = s->v[i];
      if (m.wall >= w.walls.size() || m.tile >= w.grids[m.wall].n_tiles) { w.err = "bad wall / tile of a surface molecule"; return MCX_ERR_INVALID_ARG; }
      m.pos = uv2xyz(w, w.walls[m.wall], m.u, m.v);
      if (w.tiles[m.wall].empty()) w.tiles[m.wall].assign(w.grids[m.wall].n_tiles, MCX_NONE);
      if (!(m.flags & MCX_MOL_DEFUNCT)) w.tiles[m.wall][m.tile] = m.id;
    }
    if (!w.in_this_partition(m.pos)) { w.err = "molecule outside partition"; return MCX_ERR_ESCAPED; }
    m.subpart = w.subpart_index(m.pos);
    w.mols.push_back(m);
    w.id_to_index[m.id] = (uint32_t)i;
    if (!(m.flags & MCX_MOL_DEFUNCT)) { w.sched_ids.push_back(m.id); list_insert(w, w.mols.back()); w.species_count[m.species]++; }
  }
  if (s->n) w.next_id = std::max(w.next_id, max_id + 1);  // ids of molecules that died before this upload are not handed out again
  return 0;
}
// ReleaseEvent::release_ellipsoid_or_rectcuboid (src4/release_event.cpp:953-1003) with the product's random-number
// contract (include/mcx.h: mcx_release): molecule id draws from its own Philox stream of the release domain
// (iteration | 2^63) instead of the reference's one sequential stream; the arithmetic per molecule is the reference's:
//   do { pos = rng_dbl - 0.5 per axis } while (spheroidal && len3_squared(pos) >= 0.25);      :965-971
//   SPHERICAL_SHELL: r = sqrt(len3_squared(pos)) * 2; pos = r == 0 ? (0, 0, 0.5) : pos / r;      :973-980
//   location = pos * diameter + location;                                                       :982-991
// new molecules: MOLECULE_FLAG_VOL | MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN (:996-997), diffusion_time = release time.
// a well-formed postfix program: never pops an empty stack, leaves exactly one value, at most 28 bytes / 24 deep
static inline bool region_expr_ok(const mcx_release& r) {
  if (r.region_expr_len > sizeof(r.region_expr)) return false;
  int depth = 0;
  for (uint32_t q = 0; q < r.region_expr_len; q++) {
    const uint8_t op = r.region_expr[q];
    if (op < 32) { if (++depth > 24) return false; }
    else if (op == MCX_REGION_UNION || op == MCX_REGION_INTERSECT || op == MCX_REGION_DIFFERENCE) { if (depth < 2) return false; depth--; }
    else return false;
  }
  return r.region_expr_len == 0 || depth == 1;
}
// is_point_inside_region_expr_recursively (release_event.cpp:787-813) on the membership mask of one ray cast: the two
// masks, or the postfix program of mcx_release::region_expr
static inline bool region_accepts(const mcx_release& r, uint32_t inside_mask) {
  if (r.region_expr_len == 0) return (inside_mask & r.region_in) == r.region_in && (inside_mask & r.region_out) == 0u;
  uint32_t stack = 0; int depth = 0;   // a stack of booleans, top = bit 0
  for (uint32_t q = 0; q < r.region_expr_len; q++) {
    const uint8_t op = r.region_expr[q];
    if (op < 32) { stack = (stack << 1) | ((inside_mask >> op) & 1u); depth++; continue; }
    const uint32_t b = stack & 1u, a = (stack >> 1) & 1u;
    const uint32_t v = op == MCX_REGION_UNION ? (a | b) : (op == MCX_REGION_INTERSECT ? (a & b) : (a & ~b & 1u));
    stack = ((stack >> 2) << 1) | v; depth--;
  }
  return depth == 1 && (stack & 1u);
}
int orc_release_volume_molecules(void* h, const mcx_release* r, uint32_t* first_id_out) {
  World& w = *(World*)h;
  if (r->species >= w.species.size() || !(w.species[r->species].flags & MCX_SP_VOL)) { w.err = "release: not a volume species"; return MCX_ERR_INVALID_ARG; }
  if (r->shape > MCX_RELEASE_REGION) { w.err = "release: unknown shape"; return MCX_ERR_INVALID_ARG; }
  if (r->shape == MCX_RELEASE_REGION && r->region_expr_len == 0 && (r->region_in == 0 || (r->region_in & r->region_out))) { w.err = "region release: bad object masks"; return MCX_ERR_INVALID_ARG; }
  if (r->shape == MCX_RELEASE_REGION && !region_expr_ok(*r)) { w.err = "region release: malformed region expression"; return MCX_ERR_INVALID_ARG; }
  if (r->counted_volume_index >= w.n_cv) { w.err = "release: counted_volume_index out of range"; return MCX_ERR_INVALID_ARG; }
  const double it = (double)w.iteration;
  if (r->release_time != 0 && !(r->release_time >= it && r->release_time < it + 1.0)) { w.err = "release_time outside the current iteration"; return MCX_ERR_INVALID_ARG; }
  const bool spheroidal = r->shape == MCX_RELEASE_SPHERICAL || r->shape == MCX_RELEASE_SPHERICAL_SHELL;
  const double t_rel = r->release_time > it ? r->release_time : it;
  const uint32_t first = w.next_id;
  for (uint64_t k = 0; k < r->number; k++) {
    WordSource ws; ws.kind = WordSource::PHILOX; ws.seed = w.cfg.seed; ws.mol_id = first + (uint32_t)k;
    ws.iteration = w.iteration | 0x8000000000000000ull;
    V3 pos;
    do {
      pos.x = ws.dbl() - 0.5;
      pos.y = ws.dbl() - 0.5;
      pos.z = ws.dbl() - 0.5;
    } while (spheroidal && pos.x * pos.x + pos.y * pos.y + pos.z * pos.z >= 0.25);
    if (r->shape == MCX_RELEASE_SPHERICAL_SHELL) {
      const double rad = sqrt(pos.x * pos.x + pos.y * pos.y + pos.z * pos.z) * 2;
      if (rad == 0) pos = {0.0, 0.0, 0.5};
      else pos = {pos.x / rad, pos.y / rad, pos.z / rad};
    }
    Mol n{};
    n.pos = {pos.x * r->diameter[0] + r->location[0], pos.y * r->diameter[1] + r->location[1], pos.z * r->diameter[2] + r->location[2]};
    uint32_t cvi_k = r->counted_volume_index;
    if (r->shape == MCX_RELEASE_REGION) {
      // release_inside_regions (release_event.cpp:904-951): redrawn until the point lies inside the region
      Eval E(w, ws);
      int tries = 0; bool ok = false;
      for (;;) {
        const bool inb = w.in_this_partition(n.pos);
        Eval::RayScan sc;
        if (inb) sc = E.scan_ray(n.pos);
        if (inb && !sc.redo && region_accepts(*r, sc.inside_mask)) {
          cvi_k = 0;
          if (sc.first_wall != MCX_NONE) cvi_k = sc.first_side == WALL_FRONT ? w.walls[sc.first_wall].cv_front : w.walls[sc.first_wall].cv_back;
          if (!w.cv_mask.empty()) { const uint32_t kq = cv_lookup(w, sc.inside_mask & w.cv_all); if (kq != MCX_NONE) cvi_k = kq; }
          ok = true;
          break;
        }
        if (++tries >= 100000) break;
        pos.x = ws.dbl() - 0.5; pos.y = ws.dbl() - 0.5; pos.z = ws.dbl() - 0.5;
        n.pos = {pos.x * r->diameter[0] + r->location[0], pos.y * r->diameter[1] + r->location[1], pos.z * r->diameter[2] + r->location[2]};
      }
      if (!ok) { w.err = "region release: no point of the box lies inside the region"; return MCX_ERR_INVALID_ARG; }
    }
    if (!w.in_this_partition(n.pos)) { w.err = "released molecule outside partition"; return MCX_ERR_ESCAPED; }
    n.id = w.next_id++; n.species = r->species;
    n.flags = MCX_MOL_SCHEDULE_UNIMOL | (t_rel > it ? MCX_MOL_PARTIAL : 0u);
    n.diffusion_time = t_rel; n.unimol_rxn_time = TIME_INVALID;
    n.subpart = w.subpart_index(n.pos); n.cvi = cvi_k;
    w.mols.push_back(n);
    if (w.id_to_index.size() <= n.id) w.id_to_index.resize((size_t)n.id + 1, MCX_NONE);
    w.id_to_index[n.id] = (uint32_t)w.mols.size() - 1;
    w.sched_ids.push_back(n.id);
    list_insert(w, w.mols.back());
    w.species_count[n.species]++;
  }
  if (first_id_out) *first_id_out = first;
  return 0;
}
// ReleaseEvent::release_onto_regions (release_event.cpp:640-760) in the product's parallel form (include/mcx.h,
// mcx_release_surface_molecules): every molecule picks tiles from its own stream, rounds of pick / lowest index wins,
// then the reference's fall-back fill.  grid2uv_random: grid_utils.inl:256-288.
int orc_release_surface_molecules(void* h, const mcx_surface_release* r, uint32_t* first_id_out) {
  World& w = *(World*)h;
  if (r->species >= w.species.size() || (w.species[r->species].flags & MCX_SP_VOL)) { w.err = "surface release: not a surface species"; return MCX_ERR_INVALID_ARG; }
  if (!r->walls || r->n_walls == 0) { w.err = "surface release: empty wall list"; return MCX_ERR_INVALID_ARG; }
  const double it = (double)w.iteration;
  if (r->release_time != 0 && !(r->release_time >= it && r->release_time < it + 1.0)) { w.err = "release_time outside the current iteration"; return MCX_ERR_INVALID_ARG; }
  const double t_rel = r->release_time > it ? r->release_time : it;
  std::vector<double> cum(r->n_walls), area(r->n_walls);
  double total = 0;
  uint64_t vacant = 0;
  for (uint64_t a = 0; a < r->n_walls; a++) {
    const uint32_t wi = r->walls[a];
    if (wi >= w.walls.size()) { w.err = "surface release: wall index out of range"; return MCX_ERR_INVALID_ARG; }
    area[a] = w.walls[wi].area; total += area[a]; cum[a] = total;
    if (w.tiles[wi].empty()) w.tiles[wi].assign(w.grids[wi].n_tiles, MCX_NONE);
    for (uint32_t occ : w.tiles[wi]) vacant += occ == MCX_NONE;
  }
  if (vacant < r->number) { w.err = "surface release: more molecules than vacant tiles"; return MCX_ERR_CAPACITY; }
  const uint32_t first = w.next_id;
  const size_t n = (size_t)r->number;
  // claim per (wall, tile): index of the molecule that holds it
  std::vector<std::vector<uint32_t>> claim(w.walls.size());
  for (uint64_t a = 0; a < r->n_walls; a++) claim[r->walls[a]].assign(w.grids[r->walls[a]].n_tiles, MCX_NONE);
  std::vector<uint32_t> got_wall(n, MCX_NONE), got_tile(n, MCX_NONE);
  std::vector<uint32_t> pend(n), next;
  for (size_t k = 0; k < n; k++) pend[k] = (uint32_t)k;
  for (unsigned int round = 0; round < MCX_SURFACE_RELEASE_ROUNDS && !pend.empty(); round++) {
    std::vector<uint32_t> pw(pend.size(), MCX_NONE), pt(pend.size(), MCX_NONE);
    for (size_t q = 0; q < pend.size(); q++) {
      WordSource ws; ws.kind = WordSource::PHILOX; ws.seed = w.cfg.seed; ws.mol_id = first + pend[q];
      ws.iteration = w.iteration | 0x8000000000000000ull;
      for (unsigned int d = 0; d < round; d++) (void)ws.next();
      double A = ws.dbl() * total;
      size_t low = 0, hi = r->n_walls - 1, mid;   // cum_area_bisect_high (release_event.cpp:60-81)
      while (hi - low > 1) { mid = (hi + low) / 2; if (cum[mid] > A) hi = mid; else low = mid; }
      const size_t at = cum[low] > A ? low : hi;
      const uint32_t wi = r->walls[at];
      if (at != 0) A -= cum[at - 1];
      const Grid& g = w.grids[wi];
      uint32_t tile = (uint32_t)((double)(g.n_axis * g.n_axis) * (A / area[at]));
      if (tile >= g.n_tiles) tile = g.n_tiles - 1;
      if (w.tiles[wi][tile] == MCX_NONE && claim[wi][tile] == MCX_NONE) { pw[q] = wi; pt[q] = tile; }
    }
    // lowest index wins a tile
    std::vector<std::vector<uint32_t>> bid = claim;
    for (size_t q = 0; q < pend.size(); q++) if (pw[q] != MCX_NONE) bid[pw[q]][pt[q]] = std::min(bid[pw[q]][pt[q]], pend[q]);
    next.clear();
    for (size_t q = 0; q < pend.size(); q++) {
      if (pw[q] != MCX_NONE && bid[pw[q]][pt[q]] == pend[q]) { got_wall[pend[q]] = pw[q]; got_tile[pend[q]] = pt[q]; claim[pw[q]][pt[q]] = pend[q]; }
      else next.push_back(pend[q]);
    }
    pend.swap(next);
  }
  if (!pend.empty()) {
    std::sort(pend.begin(), pend.end());
    size_t q = 0;
    for (uint64_t a = 0; a < r->n_walls && q < pend.size(); a++) {
      const uint32_t wi = r->walls[a];
      for (uint32_t tile = 0; tile < w.grids[wi].n_tiles && q < pend.size(); tile++) {
        if (w.tiles[wi][tile] != MCX_NONE || claim[wi][tile] != MCX_NONE) continue;
        claim[wi][tile] = pend[q]; got_wall[pend[q]] = wi; got_tile[pend[q]] = tile; q++;
      }
    }
    if (q != pend.size()) { w.err = "surface release: ran out of vacant tiles"; return MCX_ERR_CAPACITY; }
  }
  for (size_t k = 0; k < n; k++) {
    const uint32_t wi = got_wall[k], tile = got_tile[k];
    WordSource ws; ws.kind = WordSource::PHILOX; ws.seed = w.cfg.seed; ws.mol_id = first + (uint32_t)k;
    ws.iteration = w.iteration | 0xC000000000000000ull;   // placement draws: a domain of their own
    Mol m{};
    if (r->randomize_pos) grid2uv_random(w.walls[wi], w.grids[wi], tile, ws, m.u, m.v);
    else grid2uv(w.walls[wi], w.grids[wi], tile, m.u, m.v);
    m.orient = r->orientation ? r->orientation : ((ws.next() & 1) ? 1 : -1);
    m.pos = uv2xyz(w, w.walls[wi], m.u, m.v);
    m.id = w.next_id++; m.species = r->species;
    m.flags = MCX_MOL_SCHEDULE_UNIMOL | (t_rel > it ? MCX_MOL_PARTIAL : 0u);
    m.diffusion_time = t_rel; m.unimol_rxn_time = TIME_INVALID;
    m.subpart = w.subpart_index(m.pos); m.cvi = 0;
    m.wall = wi; m.tile = tile; m.created_wall = m.created_tile = MCX_NONE;
    w.mols.push_back(m);
    if (w.id_to_index.size() <= m.id) w.id_to_index.resize((size_t)m.id + 1, MCX_NONE);
    w.id_to_index[m.id] = (uint32_t)w.mols.size() - 1;
    w.sched_ids.push_back(m.id);
    list_insert(w, w.mols.back());
    w.tiles[wi][tile] = m.id;
    w.species_count[m.species]++;
  }
  if (first_id_out) *first_id_out = first;
  return 0;
}
// ReleaseEvent::release_list (release_event.cpp:1008-1040), volume molecules
int orc_release_list(void* h, uint64_t n_list, const uint32_t* species, const double* x, const double* y, const double* z,
                     const uint32_t* counted_volume, double release_time, uint32_t* first_id_out) {
  World& w = *(World*)h;
  const double it = (double)w.iteration;
  if (release_time != 0 && !(release_time >= it && release_time < it + 1.0)) { w.err = "release_time outside the current iteration"; return MCX_ERR_INVALID_ARG; }
  const double t_rel = release_time > it ? release_time : it;
  const uint32_t first = w.next_id;
  for (uint64_t k = 0; k < n_list; k++) {
    if (species[k] >= w.species.size() || !(w.species[species[k]].flags & MCX_SP_VOL)) { w.err = "release list: not a volume species"; return MCX_ERR_INVALID_ARG; }
    Mol n{};
    n.pos = {x[k], y[k], z[k]};
    if (!w.in_this_partition(n.pos)) { w.err = "released molecule outside partition"; return MCX_ERR_ESCAPED; }
    n.id = w.next_id++; n.species = species[k];
    n.flags = MCX_MOL_SCHEDULE_UNIMOL | (t_rel > it ? MCX_MOL_PARTIAL : 0u);
    n.diffusion_time = t_rel; n.unimol_rxn_time = TIME_INVALID;
    n.subpart = w.subpart_index(n.pos); n.cvi = counted_volume ? counted_volume[k] : 0;
    n.wall = n.tile = MCX_NONE; n.created_wall = n.created_tile = MCX_NONE;
    w.mols.push_back(n);
    if (w.id_to_index.size() <= n.id) w.id_to_index.resize((size_t)n.id + 1, MCX_NONE);
    w.id_to_index[n.id] = (uint32_t)w.mols.size() - 1;
    w.sched_ids.push_back(n.id);
    list_insert(w, w.mols.back());
    w.species_count[n.species]++;
  }
  if (first_id_out) *first_id_out = first;
  return 0;
}
int orc_get_next_molecule_id(void* h, uint32_t* out) { *out = ((World*)h)->next_id; return 0; }
int orc_set_next_molecule_id(void* h, uint32_t next_id) { World& w = *(World*)h; if (next_id > w.next_id) w.next_id = next_id; return 0; }
int orc_get_wall_grids(void* h, uint8_t* out, uint64_t n_walls) {
  World& w = *(World*)h;
  for (uint64_t i = 0; i < n_walls && i < w.walls.size(); i++) out[i] = (i < w.tiles.size() && !w.tiles[i].empty()) ? 1 : 0;
  return 0;
}
int orc_set_wall_grids(void* h, const uint8_t* in, uint64_t n_walls) {
  World& w = *(World*)h;
  if (w.tiles.size() != w.walls.size()) w.tiles.resize(w.walls.size());
  for (uint64_t i = 0; i < n_walls && i < w.walls.size(); i++)
    if (in[i] && w.tiles[i].empty()) w.tiles[i].assign(w.grids[i].n_tiles, MCX_NONE);
  return 0;
}
uint64_t orc_num_molecules(void* h) {
  World& w = *(World*)h; uint64_t n = 0;
  for (auto& m : w.mols) n += !(m.flags & MCX_MOL_DEFUNCT);
  return n;
}
int orc_download_molecules(void* h, mcx_mol_soa* o, uint64_t cap) {
  World& w = *(World*)h; uint64_t n = 0;
  for (auto& m : w.mols) {
    if (m.flags & MCX_MOL_DEFUNCT) continue;
    if (n >= cap) return MCX_ERR_CAPACITY;
    o->x[n] = m.pos.x; o->y[n] = m.pos.y; o->z[n] = m.pos.z; o->id[n] = m.id; o->species[n] = m.species;
    if (o->flags) o->flags[n] = m.flags;
    if (o->diffusion_time) o->diffusion_time[n] = (m.flags & MCX_MOL_PARTIAL) ? m.diffusion_time : (double)w.iteration;
    if (o->unimol_rxn_time) o->unimol_rxn_time[n] = m.unimol_rxn_time;
    if (o->counted_volume) o->counted_volume[n] = m.cvi;
    if (o->wall) {
      o->wall[n] = m.wall; o->tile[n] = m.tile; o->orientation[n] = m.orient; o->u[n] = m.u; o->v[n] = m.v;
    }
    n++;
  }
  o->n = n;
  return 0;
}
static void fill_stats(World& w, mcx_step_stats* s, uint64_t iters, double ms) {
  if (!s) return;
  memset(s, 0, sizeof(*s));
  s->iterations = iters; s->molecule_steps = w.stats.molecule_steps; s->n_live = orc_num_molecules(&w);
  s->ray_polygon_tests = w.stats.ray_polygon_tests; s->ray_polygon_colls = w.stats.ray_polygon_colls;
  s->mol_wall_reflections = w.stats.reflections; s->mol_wall_transparent = w.stats.transparent;
  s->mol_wall_absorptions = w.stats.absorptions; s->vol_mol_vol_mol_collisions = w.stats.volvol_collisions;
  s->bimol_rxns = w.stats.bimol_rxns; s->unimol_rxns = w.stats.unimol_rxns; s->wall_redos = w.stats.redos;
  s->resolve_retries = w.stats.retries; s->unresolved_conflicts = w.stats.unresolved;
  s->products_created = w.stats.products; s->device_ms = ms;
}
// mode 0 = sequential (reference semantics, global ISAAC64), 1 = snapshot with Philox streams
int orc_step(void* h, uint32_t n_iterations, int mode, mcx_step_stats* stats) {
  World& w = *(World*)h;
  w.stats = Stats();
  auto t0 = std::chrono::steady_clock::now();
  for (uint32_t k = 0; k < n_iterations; k++) {
    if (mode == 0) {
      step_sequential(w);
      if (w.iteration % 100 == 0) defragment(w);  // DEFRAGMENTATION_PERIODICITY, defines.h
    } else {
      SnapStreams st{MCX_RNG_PHILOX, nullptr, 0, nullptr, 0};
      step_snapshot(w, st);
    }
    if (!w.err.empty()) return MCX_ERR_ESCAPED;
  }
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  fill_stats(w, stats, n_iterations, ms);
  return 0;
}
// One traced iteration.  mode 0: sequential, also records the per-molecule ISAAC tape;
// mode 1: snapshot+Philox; mode 2: snapshot replaying words[off[id]...].
int orc_trace_step(void* h, int mode, const uint32_t* words, uint64_t n_words, const uint64_t* off, uint64_t n_ids,
                   mcx_trace_rec* trace_out, uint64_t n_trace, mcx_step_stats* stats) {
  World& w = *(World*)h;
  w.stats = Stats();
  w.tracing = true; w.trace.clear();
  mcx_trace_rec z{}; z.rxn_class = z.rxn_pathway = z.rxn_partner = MCX_NONE;
  w.trace.assign(std::max<uint64_t>(n_trace, w.next_id), z);
  if (mode == 0) { w.record_tape = true; step_sequential(w); w.record_tape = false; }
  else {
    SnapStreams st{mode == 2 ? MCX_RNG_TAPE : MCX_RNG_PHILOX, words, n_words, off, n_ids};
    step_snapshot(w, st);
  }
  w.tracing = false;
  for (uint64_t i = 0; i < n_trace && i < w.trace.size(); i++) trace_out[i] = w.trace[i];
  fill_stats(w, stats, 1, 0);
  return w.err.empty() ? 0 : MCX_ERR_ESCAPED;
}
// First evaluation (round 0 of the snapshot semantics) of the molecules with id % stride == offset only, against the
// current state, which stays untouched: what lets a test compare a full-size configuration with the oracle in seconds
// (the evaluation of one molecule depends on the snapshot alone).  Philox streams.
int orc_trace_sample(void* h, uint32_t stride, uint32_t offset, mcx_trace_rec* trace_out, uint64_t n_trace) {
  World& w = *(World*)h;
  const Stats keep = w.stats;
  w.tracing = true; w.trace.clear();
  mcx_trace_rec z{}; z.rxn_class = z.rxn_pathway = z.rxn_partner = MCX_NONE;
  w.trace.assign(std::max<uint64_t>(n_trace, w.next_id), z);
  w.sample_stride = stride ? stride : 1; w.sample_offset = offset;
  SnapStreams st{MCX_RNG_PHILOX, nullptr, 0, nullptr, 0};
  step_snapshot(w, st);
  w.sample_stride = 0; w.tracing = false; w.stats = keep;
  for (uint64_t i = 0; i < n_trace && i < w.trace.size(); i++) trace_out[i] = w.trace[i];
  return w.err.empty() ? 0 : MCX_ERR_ESCAPED;
}
// tape recorded by the last sequential traced iteration
uint64_t orc_tape_size(void* h) { return ((World*)h)->tape_words.size(); }
int orc_tape_get(void* h, uint32_t* words, uint64_t* off, uint32_t* len, uint64_t n_ids) {
  World& w = *(World*)h;
  memcpy(words, w.tape_words.data(), w.tape_words.size() * 4);
  for (uint64_t i = 0; i < n_ids; i++) { off[i] = i < w.tape_off.size() ? w.tape_off[i] : 0; len[i] = i < w.tape_len.size() ? w.tape_len[i] : 0; }
  return 0;
}
int orc_counts(void* h, uint64_t* per_species, uint32_t ns, uint64_t* per_rule, uint32_t nr) {
  World& w = *(World*)h;
  for (uint32_t i = 0; i < ns && per_species; i++) per_species[i] = i < w.species_count.size() ? w.species_count[i] : 0;
  for (uint32_t i = 0; i < nr && per_rule; i++) per_rule[i] = i < w.rxn_count.size() ? w.rxn_count[i] : 0;
  return 0;
}
// geometry set-up introspection for parity of the wall tables
uint64_t orc_subpart_wall_count(void* h, uint32_t subpart) { return ((World*)h)->walls_per_subpart[subpart].size(); }
int orc_subpart_walls(void* h, uint32_t subpart, uint32_t* out) {
  auto& v = ((World*)h)->walls_per_subpart[subpart];
  std::copy(v.begin(), v.end(), out); return 0;
}
int orc_wall_constants(void* h, uint32_t wi, double out[16]) {
  const Wall& f = ((World*)h)->walls[wi];
  double t[16] = {f.normal.x, f.normal.y, f.normal.z, f.distance_to_origin, f.unit_u.x, f.unit_u.y, f.unit_u.z,
                  f.unit_v.x, f.unit_v.y, f.unit_v.z, f.uv_vert1_u, f.uv_vert2_u, f.uv_vert2_v, f.area, 0, 0};
  memcpy(out, t, sizeof(t)); return 0;
}
// RNG restatement entry points (pinned against oracle/_ref/librefrng.so)
void* orc_rng_new(uint32_t seed) { Isaac64* r = new Isaac64(); r->init(seed); return r; }
void orc_rng_free(void* r) { delete (Isaac64*)r; }
uint32_t orc_rng_uint(void* r) { return ((Isaac64*)r)->next32(); }
double orc_rng_dbl(void* r) { WordSource s; s.isaac = (Isaac64*)r; return s.dbl(); }
double orc_rng_gauss(void* r) { WordSource s; s.isaac = (Isaac64*)r; return s.gauss(); }
void orc_rng_fill_uint(void* r, uint32_t* out, long n) { for (long i = 0; i < n; i++) out[i] = ((Isaac64*)r)->next32(); }
void orc_rng_fill_gauss(void* r, double* out, long n) { WordSource s; s.isaac = (Isaac64*)r; for (long i = 0; i < n; i++) out[i] = s.gauss(); }
long long orc_rng_uses(void* r) { return ((Isaac64*)r)->uses(); }
void orc_philox_block(uint64_t seed, uint32_t id, uint64_t it, uint32_t block, uint32_t out[4]) { philox_block(seed, id, it, block, out); }
// gaussians from a tape (device Ziggurat parity)
void orc_tape_gauss(const uint32_t* words, uint64_t n_words, double* out, long n, uint32_t* used) {
  WordSource s; s.kind = WordSource::TAPE; s.tape = words; s.tape_len = n_words;
  for (long i = 0; i < n; i++) out[i] = s.gauss();
  *used = s.used;
}
// ---- unit entry points: one reference function each, for pinning against oracle/_ref/libmcell3ref.so
//      (the reference's own compiled arithmetic) in tests/test_oracle_vs_reference.py -------------------------
static inline bool grid_init_flag(const unsigned char* g, unsigned i) { return !g || g[i]; }
static void unit_world(World& w, const double* v9) {
  w.cfg = mcx_config{};
  w.cfg.partition_edge_length = 1000; w.cfg.num_subparts_per_edge = 1;
  w.n_sp = 1; w.sp_len = 1000; w.sp_rcp = 1e-3;
  w.verts = {{v9[0], v9[1], v9[2]}, {v9[3], v9[4], v9[5]}, {v9[6], v9[7], v9[8]}};
  w.walls.resize(1);
  w.walls[0].vi[0] = 0; w.walls[0].vi[1] = 1; w.walls[0].vi[2] = 2;
  w.walls[0].surf_class = MCX_NONE; w.walls[0].object = 0;
  init_wall_constants(w, w.walls[0]);
}
// a mesh without a partition: wall constants and edge pairing only (surface_net + Edge::reinit_edge_constants)
static void unit_mesh(World& w, const double* verts, unsigned nv, const unsigned* tri, unsigned nw) {
  w.cfg = mcx_config{};
  w.cfg.partition_edge_length = 1000; w.cfg.num_subparts_per_edge = 1;
  w.n_sp = 1; w.sp_len = 1000; w.sp_rcp = 1e-3;
  w.verts.resize(nv);
  for (unsigned i = 0; i < nv; i++) w.verts[i] = {verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]};
  w.walls.resize(nw);
  for (unsigned i = 0; i < nw; i++) {
    Wall& f = w.walls[i];
    f.vi[0] = tri[3 * i]; f.vi[1] = tri[3 * i + 1]; f.vi[2] = tri[3 * i + 2];
    f.surf_class = MCX_NONE; f.object = 0;
    init_wall_constants(w, f);
  }
  init_edges(w, 0, nw);
}
// find_neighbor_tiles of every tile of every wall that has a grid, as a CSR of (wall, tile) pairs in list order (same
// layout as ref4_neighbor_tile_table, oracle/ref_mcell4_tiles_shim.cpp); grid_init: per wall 0 = no grid yet, null = all
unsigned long long orc_unit_neighbor_tile_table(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                                                const unsigned char* has_grid, unsigned* start, unsigned* out_pairs,
                                                unsigned long long cap) {
  World w; unit_mesh(w, verts, n_verts, tri, n_walls);
  w.grids.resize(n_walls); w.tiles.resize(n_walls);
  for (unsigned i = 0; i < n_walls; i++) {
    grid_init(w, w.walls[i], w.grids[i]);
    if (!grid_init_flag(has_grid, i)) continue;
    w.tiles[i].assign(w.grids[i].n_tiles, MCX_NONE);
  }
  unsigned long long n = 0; unsigned gt = 0;
  for (unsigned i = 0; i < n_walls; i++) {
    if (w.tiles[i].empty()) continue;
    for (uint32_t tile = 0; tile < w.grids[i].n_tiles; tile++) {
      TileNeighbors nb;
      find_neighbor_tiles(w, i, tile, nb);
      start[gt++] = (unsigned)n;
      for (const WallTile& t : nb) {
        if (n < cap) { out_pairs[2 * n] = t.first; out_pairs[2 * n + 1] = t.second; }
        n++;
      }
    }
  }
  start[gt] = (unsigned)n;
  return n;
}
// same outputs as oracle/ref_mcell3_shim.cpp: ref3_mesh_edges / ref3_find_edge_point / ref3_traverse_surface
int orc_unit_mesh_edges(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls, int* nb_wall_out,
                        int* forward_out, double* transform_out) {
  World w; unit_mesh(w, verts, n_verts, tri, n_walls);
  for (unsigned i = 0; i < n_walls; i++)
    for (int k = 0; k < 3; k++) {
      const Wall& f = w.walls[i];
      const bool paired = f.nb_wall[k] != MCX_NONE;
      nb_wall_out[3 * i + k] = paired ? (int)f.nb_wall[k] : -1;
      forward_out[3 * i + k] = paired && f.edge_forward[k] ? 1 : 0;
      double* t = transform_out + 4 * (3 * i + k);
      t[0] = paired ? f.edge_cos[k] : 0; t[1] = paired ? f.edge_sin[k] : 0;
      t[2] = paired ? f.edge_tu[k] : 0; t[3] = paired ? f.edge_tv[k] : 0;
    }
  return 0;
}
// MCell3's codes: 0..2 edge, -1 stays inside the wall, -2 cannot tell (src/wall_util.c:579-654)
int orc_unit_find_edge_point(const double* v9, const double* loc2, const double* disp2, double* edgept2) {
  World w; unit_world(w, v9);
  double eu = 0, ev = 0;
  const int r = find_edge_point(w.walls[0], loc2[0], loc2[1], disp2[0], disp2[1], eu, ev);
  edgept2[0] = eu; edgept2[1] = ev;
  return r == 3 ? -1 : (r == 4 ? -2 : r);
}
int orc_unit_traverse_surface(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                              const unsigned* q_wall, const int* q_side, const double* q_uv, unsigned n_q, int* wall_out,
                              double* uv_out) {
  World w; unit_mesh(w, verts, n_verts, tri, n_walls);
  for (unsigned q = 0; q < n_q; q++) {
    double nu = 0, nv = 0;
    const uint32_t there = traverse_surface(w.walls[q_wall[q]], q_uv[2 * q], q_uv[2 * q + 1], q_side[q], nu, nv);
    wall_out[q] = there == MCX_NONE ? -1 : (int)there;
    uv_out[2 * q] = nu; uv_out[2 * q + 1] = nv;
  }
  return 0;
}
// ray_trace_surf for a batch of (wall, uv, displacement) queries on one mesh; same outputs as ref3_ray_trace_2d
int orc_unit_ray_trace_surf(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls, const unsigned* q_wall,
                            const double* q_uv, const double* q_disp, unsigned n_q, int* wall_out, double* uv_out) {
  World w; unit_mesh(w, verts, n_verts, tri, n_walls);
  for (unsigned q = 0; q < n_q; q++) {
    double ou = 0, ov = 0;
    const uint32_t there = ray_trace_surf(w, q_wall[q], q_uv[2 * q], q_uv[2 * q + 1], q_disp[2 * q], q_disp[2 * q + 1], ou, ov);
    wall_out[q] = there == MCX_NONE ? -1 : (int)there;
    uv_out[2 * q] = there == MCX_NONE ? 0 : ou; uv_out[2 * q + 1] = there == MCX_NONE ? 0 : ov;
  }
  return 0;
}
// CollisionUtils::collect_crossed_subparts (collision_utils_subparts.inl:127-300) for one move; same outputs as
// oracle/ref_mcell4_shim.cpp's ref4_collect_crossed_subparts (out_mols in insertion order here, a set there)
unsigned orc_unit_collect_crossed_subparts(const double* origin3, double partition_edge_length, unsigned n_subparts_per_edge,
                                           int use_expanded_list, double rxn_radius, const double* pos3, const double* disp3,
                                           int collect_for_molecules, int collect_for_walls, unsigned* out_walls,
                                           unsigned* n_walls, unsigned* out_mols, unsigned* n_mols, unsigned cap) {
  World w;
  memset(&w.cfg, 0, sizeof(w.cfg));
  for (int k = 0; k < 3; k++) w.cfg.origin[k] = origin3[k];
  w.cfg.partition_edge_length = partition_edge_length;
  w.cfg.num_subparts_per_edge = n_subparts_per_edge;
  w.cfg.use_expanded_list = use_expanded_list ? 1 : 0;
  w.cfg.rxn_radius_3d = rxn_radius;
  w.n_sp = n_subparts_per_edge;
  w.sp_len = partition_edge_length / n_subparts_per_edge;
  w.sp_rcp = 1.0 / w.sp_len;
  WordSource ws;
  Eval e(w, ws);
  const V3 pos = {pos3[0], pos3[1], pos3[2]};
  std::vector<uint32_t> sw, sm;
  const uint32_t dest = e.collect_crossed_subparts(pos, w.subpart_index(pos), V3{disp3[0], disp3[1], disp3[2]},
                                                   collect_for_molecules != 0, collect_for_walls != 0, sw, sm);
  for (size_t k = 0; k < sw.size() && k < cap; k++) out_walls[k] = sw[k];
  for (size_t k = 0; k < sm.size() && k < cap; k++) out_mols[k] = sm[k];
  *n_walls = (unsigned)sw.size(); *n_mols = (unsigned)sm.size();
  return dest;
}
void orc_unit_wall_constants(const double* v9, double* out16) {
  World w; unit_world(w, v9);
  const Wall& f = w.walls[0];
  double t[16] = {f.normal.x, f.normal.y, f.normal.z, f.distance_to_origin, f.unit_u.x, f.unit_u.y, f.unit_u.z,
                  f.unit_v.x, f.unit_v.y, f.unit_v.z, f.uv_vert1_u, f.uv_vert2_u, f.uv_vert2_v, f.area, 0, 0};
  memcpy(out16, t, sizeof(t));
}
// returns the reference's codes: COLLIDE_REDO -1, COLLIDE_MISS 0, COLLIDE_FRONT 1, COLLIDE_BACK 2
// (src/mcell_structs.h); words = the 32-bit words the RNG would deliver next
int orc_unit_collide_wall(const double* point3, double* move3, const double* v9, const uint32_t* words,
                          uint64_t n_words, double* t, double* hit3, long long* words_used) {
  World w; unit_world(w, v9);
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  Eval E(w, rs);
  V3 move = {move3[0], move3[1], move3[2]}, hit = {0, 0, 0};
  double tt = 0;
  int r = E.collide_wall(V3{point3[0], point3[1], point3[2]}, 0, move, tt, hit);
  move3[0] = move.x; move3[1] = move.y; move3[2] = move.z;
  *t = tt; hit3[0] = hit.x; hit3[1] = hit.y; hit3[2] = hit.z;
  *words_used = rs.used;
  return r == WALL_MISS ? 0 : r == WALL_FRONT ? 1 : r == WALL_BACK ? 2 : -1;
}
// get_closest_wall_collision over a whole mesh held by ONE subpartition, then reflect_from_wall (collision_utils.inl:
// 819-914, 1711-1747); same outputs as oracle/ref_mcell4_leaf_shim.cpp: ref4_closest_wall_and_reflect
int orc_unit_closest_wall_and_reflect(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                                      const double* pos3, double* move3, unsigned last_hit_wall, const uint32_t* words,
                                      uint64_t n_words, unsigned* wall, int* side, double* t, double* hit3, double* pos_after3,
                                      double* disp_after3, double* t_steps_io, long long* words_used,
                                      unsigned long long* ray_polygon_tests) {
  World w; unit_mesh(w, verts, n_verts, tri, n_walls);
  w.walls_per_subpart.assign(1, {});
  for (unsigned i = 0; i < n_walls; i++) w.walls_per_subpart[0].push_back(i);
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  Eval E(w, rs);
  V3 disp = {move3[0], move3[1], move3[2]}, up_to_wall = {0, 0, 0};
  const V3 pos = {pos3[0], pos3[1], pos3[2]};
  Collision c;
  const bool found = E.closest_wall_collision(pos, 0, last_hit_wall, disp, up_to_wall, c);
  move3[0] = disp.x; move3[1] = disp.y; move3[2] = disp.z;
  *words_used = rs.used;
  *ray_polygon_tests = w.stats.ray_polygon_tests;
  if (!found) return 0;
  *wall = c.wall; *side = c.type == COLL_WALL_FRONT ? 1 : 2; *t = c.time;
  hit3[0] = c.pos.x; hit3[1] = c.pos.y; hit3[2] = c.pos.z;
  *t_steps_io *= (1.0 - c.time);
  const V3 after = reflected_displacement(disp, w.walls[c.wall].normal, c.time);
  pos_after3[0] = c.pos.x; pos_after3[1] = c.pos.y; pos_after3[2] = c.pos.z;
  disp_after3[0] = after.x; disp_after3[1] = after.y; disp_after3[2] = after.z;
  return 1;
}
// ray_trace_vol as a whole (diffuse_react_event.cpp:627-780) for ONE molecule of the uploaded population moving by
// disp3 (in/out: REDOs change it), then sort_collisions_by_time (:341-364) the way diffuse_vol_molecule calls it.
// Same outputs as oracle/ref_mcell4_raytrace_shim.cpp: ref4_ray_trace_vol.  Returns 1 when a wall was hit, 0 for
// FINISHED, -1 for an unknown id; type 0 = molecule (what = partner id), 1 / 2 = wall front / back (what = wall).
int orc_unit_ray_trace_vol(void* h, uint32_t mol_id, double* disp3, uint32_t last_hit_wall, const uint32_t* words,
                           uint64_t n_words, int cap, int* n_coll, int* type, double* time, double* pos3, uint32_t* what,
                           long long* words_used) {
  World& w = *(World*)h;
  if (mol_id >= w.id_to_index.size() || w.id_to_index[mol_id] == MCX_NONE) return -1;
  const Mol& m = w.mols[w.id_to_index[mol_id]];
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  Eval E(w, rs);
  V3 remaining = {disp3[0], disp3[1], disp3[2]};
  std::vector<Collision> colls;
  const bool hit = E.ray_trace_vol(m.pos, m.subpart, m.id, m.species, w.can_vol_react[m.species] != 0, last_hit_wall,
                                   remaining, colls);
  if (colls.size() > 1) {
    std::stable_sort(colls.begin(), colls.end(), [](const Collision& a, const Collision& b) {
      if (a.time < b.time) return true;
      if (a.time > b.time) return false;
      if (a.type == COLL_VOLMOL && b.type == COLL_VOLMOL) return a.partner_id > b.partner_id;
      return false;
    });
  }
  disp3[0] = remaining.x; disp3[1] = remaining.y; disp3[2] = remaining.z;
  *n_coll = (int)colls.size();
  for (int k = 0; k < (int)colls.size() && k < cap; k++) {
    const Collision& c = colls[k];
    type[k] = c.type; time[k] = c.time;
    pos3[3 * k] = c.pos.x; pos3[3 * k + 1] = c.pos.y; pos3[3 * k + 2] = c.pos.z;
    what[k] = c.type == COLL_VOLMOL ? c.partner_id : c.wall;
  }
  *words_used = (long long)rs.used;
  return hit ? 1 : 0;
}
// pick_surf_displacement on a tape of words; returns the words drawn
long long orc_unit_pick_surf_displacement(double scale, const uint32_t* words, uint64_t n_words, double* out2) {
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  pick_surf_displacement(rs, scale, out2[0], out2[1]);
  return (long long)rs.used;
}
// returns COLLIDE_VOL_M (3) on hit, COLLIDE_MISS (0) otherwise
int orc_unit_collide_mol(const double* point3, const double* move3, const double* target3, double R, double* t,
                         double* hit3) {
  World w; w.cfg = mcx_config{};
  WordSource rs; Eval E(w, rs);
  Mol c{}; c.pos = {target3[0], target3[1], target3[2]}; c.id = 1;
  V3 cp = {0, 0, 0}; double tt = 0;
  bool hit = E.collide_mol(V3{point3[0], point3[1], point3[2]}, 0, V3{move3[0], move3[1], move3[2]}, c, R, tt, cp);
  *t = tt; hit3[0] = cp.x; hit3[1] = cp.y; hit3[2] = cp.z;
  return hit ? 3 : 0;
}
int orc_unit_wall_in_box(const double* v9, const double* llf3, const double* urb3) {
  World w; unit_world(w, v9);
  return wall_in_box(w, w.walls[0], V3{llf3[0], llf3[1], llf3[2]}, V3{urb3[0], urb3[1], urb3[2]}) != 0 ? 1 : 0;
}
int orc_unit_distinguishable(double a, double b, double eps) { return distinguishable(a, b, eps) ? 1 : 0; }
// returns the pathway index or -1 (no reaction)
// time_of_unimol (rxn_utils.inl:721-736) as pick_unimol_time uses it, from time 0; which_unimolecular (:774-783)
double orc_unit_time_of_unimol(double k_tot, const uint32_t* words, uint64_t n_words) {
  World w; w.cfg = mcx_config{};
  mcx_rxn_class rc{}; rc.kind = MCX_RXN_UNIMOL; rc.first_pathway = 0; rc.n_pathways = 1; rc.max_fixed_p = k_tot;
  w.classes.push_back(rc);
  w.unimol.assign(1, 0);
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  Eval E(w, rs);
  return E.pick_unimol_time(0, 0.0);
}
int orc_unit_which_unimolecular(const double* cum_probs, int n, const uint32_t* words, uint64_t n_words, long long* words_used) {
  World w; w.cfg = mcx_config{};
  w.pathways.resize(n);
  for (int i = 0; i < n; i++) { w.pathways[i] = mcx_pathway{}; w.pathways[i].cum_prob = cum_probs[i]; }
  mcx_rxn_class rc{}; rc.kind = MCX_RXN_UNIMOL; rc.first_pathway = 0; rc.n_pathways = n; rc.max_fixed_p = cum_probs[n - 1];
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  int pathway = 0;
  if (rc.n_pathways > 1) {  // the lines of step_molecule's unimolecular firing
    double match = rs.dbl() * rc.max_fixed_p;
    pathway = pathway_for_probability(w, rc, match);
  }
  *words_used = rs.used;
  return pathway;
}
int orc_unit_test_bimolecular(const double* cum_probs, int n, double scaling, const uint32_t* words, uint64_t n_words,
                              long long* words_used) {
  World w; w.cfg = mcx_config{};
  w.pathways.resize(n);
  for (int i = 0; i < n; i++) { w.pathways[i] = mcx_pathway{}; w.pathways[i].cum_prob = cum_probs[i]; }
  mcx_rxn_class rc{}; rc.kind = MCX_RXN_BIMOL_VOLVOL; rc.first_pathway = 0; rc.n_pathways = n; rc.max_fixed_p = cum_probs[n - 1];
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  Eval E(w, rs);
  int r = E.test_bimolecular(rc, scaling);
  *words_used = rs.used;
  return r;
}
// surfsurf_position_bits (the recycled branches of find_surf_product_positions for two surface reactants): SURFSURF_SWAP or 0
unsigned orc_unit_surfsurf_position_bits(unsigned keep_mask, unsigned n_products, const unsigned char* product_is_surf, int init_is_r0,
                                         const uint32_t* words, uint64_t n_words, long long* words_used) {
  World w; w.cfg = mcx_config{};
  w.species.resize(2); w.species[0] = mcx_species{}; w.species[0].flags = MCX_SP_VOL; w.species[1] = mcx_species{};
  mcx_rxn_class c{}; c.kind = MCX_RXN_BIMOL_SURFSURF; c.reactants[0] = c.reactants[1] = 1;
  mcx_pathway pw{}; pw.n_products = n_products; pw.keep_reactant_mask = keep_mask;
  for (unsigned k = 0; k < n_products; k++) pw.products[k] = product_is_surf[k] ? 1 : 0;
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  const uint32_t bits = surfsurf_position_bits(w, c, pw, init_is_r0 != 0, rs);
  *words_used = (long long)rs.used;
  return bits;
}
// place_general on a mesh with given occupied tiles; same per-entry outputs as ref4_find_surf_product_positions
// (oracle/ref_mcell4_place_shim.cpp): kind per entry of the rule's product list (0 nothing, 1 a recycled tile, 2 a vacant
// tile) with its wall and tile; returns 0, or -2 when the reaction is blocked; *words_used counts the words drawn for the
// tiles only (the orientation draws and random points that follow are not part of find_surf_product_positions)
int orc_unit_place_general(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls, const unsigned* occupied,
                           unsigned n_occupied, int rxn_kind, const unsigned char* reactant_is_surf, unsigned keep_mask, unsigned kept_info,
                           unsigned n_products, const unsigned char* product_is_surf, unsigned reac_wall, unsigned reac_tile,
                           const unsigned* recycled_wall_tile, int n_recycled, const uint32_t* words, uint64_t n_words,
                           int* entry_kind, unsigned* entry_wall, unsigned* entry_tile, long long* words_used) {
  World w; unit_mesh(w, verts, n_verts, tri, n_walls);
  w.grids.resize(n_walls); w.tiles.resize(n_walls);
  for (unsigned i = 0; i < n_walls; i++) { grid_init(w, w.walls[i], w.grids[i]); w.tiles[i].assign(w.grids[i].n_tiles, MCX_NONE); }
  for (unsigned i = 0; i < n_occupied; i++) w.tiles[occupied[2 * i]][occupied[2 * i + 1]] = 1000 + i;
  w.species.resize(2); w.species[0] = mcx_species{}; w.species[0].flags = MCX_SP_VOL; w.species[1] = mcx_species{};
  mcx_rxn_class c{}; c.kind = (uint32_t)rxn_kind; c.first_pathway = 0; c.n_pathways = 1;
  c.reactants[0] = reactant_is_surf[0] ? 1 : 0; c.reactants[1] = rxn_kind == MCX_RXN_UNIMOL ? MCX_NONE : (reactant_is_surf[1] ? 1u : 0u);
  mcx_pathway pw{}; pw.n_products = n_products; pw.keep_reactant_mask = keep_mask; pw.kept_info = kept_info;
  for (unsigned k = 0; k < n_products; k++) { pw.products[k] = product_is_surf[k] ? 1 : 0; pw.product_orientation[k] = 1; }
  SurfSite rec[2];
  for (int r = 0; r < n_recycled && r < 2; r++) rec[r] = SurfSite{recycled_wall_tile[2 * r], recycled_wall_tile[2 * r + 1], 0, 0, 1, 1};
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  auto vacant = [&](uint32_t wi, uint32_t ti) { return w.tiles[wi][ti] == MCX_NONE; };
  Placement pl; uint32_t obits = 0;
  int kinds[6] = {0, 0, 0, 0, 0, 0}; WallTile pos[6];
  // the words of the tile assignment alone: the orientation draws are skipped by giving every product an orientation, the
  // random points are counted and subtracted
  const bool ok = place_general(w, c, pw, reac_wall, reac_tile, rec, n_recycled, rs, vacant, pl, obits, kinds, pos);
  long long used = (long long)rs.used;
  if (ok) used -= 2LL * __builtin_popcount(pl.vacant_mask);   // grid2uv_random: two doubles of one word each
  *words_used = used;
  RuleEntry ent[6];
  const int n_ent = rule_entries(w, c, pw, ent);
  for (int e = 0; e < n_ent; e++) { entry_kind[e] = ok ? kinds[e] : 0; entry_wall[e] = ok ? pos[e].first : MCX_NONE; entry_tile[e] = ok ? pos[e].second : MCX_NONE; }
  return ok ? 0 : -2;
}
// test_bimolecular with a local probability factor / test_many_bimolecular (react_2D_all_neighbors); same outputs as
// ref4_test_bimolecular_lpf / ref4_test_many_bimolecular (oracle/ref_mcell4_tiles_shim.cpp)
int orc_unit_test_bimolecular_lpf(const double* cum_probs, int n, double scaling, double local_prob_factor, const uint32_t* words,
                                  uint64_t n_words, long long* words_used) {
  World w; w.cfg = mcx_config{};
  w.pathways.resize(n);
  for (int i = 0; i < n; i++) { w.pathways[i] = mcx_pathway{}; w.pathways[i].cum_prob = cum_probs[i]; }
  mcx_rxn_class rc{}; rc.kind = MCX_RXN_BIMOL_SURFSURF; rc.first_pathway = 0; rc.n_pathways = n; rc.max_fixed_p = cum_probs[n - 1];
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  Eval E(w, rs);
  int r = E.test_bimolecular(rc, scaling, local_prob_factor);
  *words_used = rs.used;
  return r;
}
int orc_unit_test_many_bimolecular(const double* cum_probs, const int* n_pathways, int n, const double* scaling, double local_prob_factor,
                                   const uint32_t* words, uint64_t n_words, int* pathway, long long* words_used) {
  World w; w.cfg = mcx_config{};
  std::vector<int> rcs; std::vector<double> sc(scaling, scaling + n);
  int q = 0;
  for (int i = 0; i < n; i++) {
    mcx_rxn_class rc{}; rc.kind = MCX_RXN_BIMOL_SURFSURF; rc.first_pathway = (uint32_t)w.pathways.size(); rc.n_pathways = (uint32_t)n_pathways[i];
    for (int k = 0; k < n_pathways[i]; k++) { mcx_pathway pw{}; pw.cum_prob = cum_probs[q++]; w.pathways.push_back(pw); }
    rc.max_fixed_p = w.pathways.back().cum_prob;
    w.classes.push_back(rc); rcs.push_back(i);
  }
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  Eval E(w, rs);
  int pw_out = -7;
  const int r = E.test_many_bimolecular(rcs, sc, local_prob_factor, pw_out);
  *pathway = pw_out; *words_used = rs.used;
  return r;
}
int orc_unit_test_intersect(const double* cum_probs, int n, double scaling, const uint32_t* words, uint64_t n_words,
                            long long* words_used) {
  World w; w.cfg = mcx_config{};
  w.pathways.resize(n);
  for (int i = 0; i < n; i++) { w.pathways[i] = mcx_pathway{}; w.pathways[i].cum_prob = cum_probs[i]; }
  mcx_rxn_class rc{}; rc.kind = MCX_RXN_BIMOL_VOLWALL; rc.first_pathway = 0; rc.n_pathways = n; rc.max_fixed_p = cum_probs[n - 1];
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  Eval E(w, rs);
  int r = E.test_intersect(rc, scaling);
  *words_used = rs.used;
  return r;
}
int orc_unit_pathway_for_probability(const double* cum_probs, int n, double match) {
  World w; w.pathways.resize(n);
  for (int i = 0; i < n; i++) { w.pathways[i] = mcx_pathway{}; w.pathways[i].cum_prob = cum_probs[i]; }
  mcx_rxn_class rc{}; rc.first_pathway = 0; rc.n_pathways = n;
  return pathway_for_probability(w, rc, match);
}

// surface grid arithmetic of one triangle (pinned against src/grid_util.c through oracle/_ref)
void orc_unit_grid_constants(const double* v9, double* out8) {
  World w; unit_world(w, v9);
  Grid g; grid_init(w, w.walls[0], g);
  double t[8] = {(double)g.n_axis, g.strip_width_rcp, g.vert2_slope, g.fullslope, g.binding_factor, g.vert0_u, g.vert0_v, (double)g.n_tiles};
  memcpy(out8, t, sizeof(t));
}
int orc_unit_uv2grid(const double* v9, const double* uv2) {
  World w; unit_world(w, v9);
  Grid g; grid_init(w, w.walls[0], g);
  const uint32_t t = uv2grid(w.walls[0], g, uv2[0], uv2[1]);
  return t == MCX_NONE ? -1 : (int)t;
}
int orc_unit_xyz2grid(const double* v9, const double* xyz3) {
  World w; unit_world(w, v9);
  Grid g; grid_init(w, w.walls[0], g);
  return (int)xyz2grid(w, V3{xyz3[0], xyz3[1], xyz3[2]}, w.walls[0], g);
}
void orc_unit_grid2uv(const double* v9, int idx, double* uv2) {
  World w; unit_world(w, v9);
  Grid g; grid_init(w, w.walls[0], g);
  grid2uv(w.walls[0], g, (uint32_t)idx, uv2[0], uv2[1]);
}
// grid2uv_random of tile idx of one triangle on a tape of words; returns the words drawn
long long orc_unit_grid2uv_random(const double* v9, int idx, const uint32_t* words, uint64_t n_words, double* uv2) {
  World w; unit_world(w, v9);
  Grid g; grid_init(w, w.walls[0], g);
  WordSource rs; rs.kind = WordSource::TAPE; rs.tape = words; rs.tape_len = n_words;
  grid2uv_random(w.walls[0], g, (uint32_t)idx, rs, uv2[0], uv2[1]);
  return (long long)rs.used;
}
void orc_unit_uv2xyz(const double* v9, const double* uv2, double* xyz3) {
  World w; unit_world(w, v9);
  V3 r = uv2xyz(w, w.walls[0], uv2[0], uv2[1]);
  xyz3[0] = r.x; xyz3[1] = r.y; xyz3[2] = r.z;
}
int orc_unit_exact_disk_max_pool(void) { return orc_exd::g_max_pool; }
// exact_disk over an explicit wall list (pinned against src/diffuse.c:1365 through oracle/_ref)
double orc_unit_exact_disk(const double* loc3, const double* mv3, double R, const double* target3, int n_walls,
                           const double* tri9) {
  std::vector<double> plane(4 * (size_t)std::max(n_walls, 1));
  for (int k = 0; k < n_walls; k++) {
    World w; unit_world(w, tri9 + 9 * k);
    const Wall& f = w.walls[0];
    plane[4 * k] = f.normal.x; plane[4 * k + 1] = f.normal.y; plane[4 * k + 2] = f.normal.z; plane[4 * k + 3] = f.distance_to_origin;
  }
  V3 loc = {loc3[0], loc3[1], loc3[2]}, tg = {target3[0], target3[1], target3[2]};
  bool same = !distinguishable_vec3(loc, tg, POS_EPS);
  return orc_exd::exact_disk({loc.x, loc.y, loc.z}, {mv3[0], mv3[1], mv3[2]}, R, {tg.x, tg.y, tg.z}, same, n_walls, tri9,
                             plane.data(), nullptr);
}
}
