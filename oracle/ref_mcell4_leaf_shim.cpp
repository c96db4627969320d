// oracle/ref_mcell4_leaf_shim.cpp — oracle/_ref build only (TEST INFRASTRUCTURE).
//
// Builds MCell4's OWN leaf arithmetic of the volume path into oracle/_ref/libmcell4leaf.so:
//   CollisionUtils::collide_mol               src4/collision_utils.inl:464-515
//   CollisionUtils::jump_away_line            :568-603
//   CollisionUtils::collide_wall              :629-812
//   CollisionUtils::is_immediate_collision,
//   CollisionUtils::get_closest_wall_collision :814-914
//   CollisionUtils::reflect_from_wall         :1711-1747
//   Wall::initialize_wall_constants           src4/wall.cpp:281-342
//   DiffusionUtils::pick_surf_displacement    src4/diffusion_utils.inl:60-93
//   Grid::initialize                          src4/wall.cpp:38-74
//   GridUtils::xyz2grid_tile_index, uv2grid_tile_index, grid2uv, grid2uv_random   src4/grid_utils.inl:48-118, 120-191, 233-253, 256-286
//   GeometryUtils::find_edge_point            src4/geometry_utils.inl:222-291
//   WallUtils::wall_in_box                    src4/wall_utils.inl:326-504
//   GeometryUtils::wall_subparts_collision_test   src4/geometry_utils.inl:110-207 (Partition::finalize_walls' distribution of
//                                             the walls over the subpartitions, partition.cpp:91-118)
//   ExactDiskUtils::exact_disk and everything it uses   src4/exact_disk_utils.inl:54-1145 (+ get_wall_bounding_box,
//                                             geometry_utils.inl:66-100); the moving species has no surface-class reactions
//   RxnUtils::test_bimolecular                src4/rxn_utils.inl:336-414 (RxnClass is libbng's, absent: a stand-in with
//                                             MCell3's cumulative-probability search, src/react_cond.c:110-171 / util.c bisect)
// The function texts are cut out of the reference files BY LINE RANGE AT BUILD TIME (oracle/Makefile: ref, into the
// git-ignored oracle/_ref/gen/) and compiled unmodified; nothing of them is stored in this repository.  The whole files
// cannot be compiled: collision_utils.inl pulls partition.h / world.h / geometry.h and with them libbng, boost and
// sparsehash, none of which exist here.  So the types those functions touch are stand-ins with the
// reference's member names (src4/geometry.h Wall / WallCollisionRejectionData, src4/molecule.h Molecule,
// src4/collision_structs.h Collision / CollisionType, src4/partition.h accessors); src4/defines.h with the reference's
// libs/glm and src/rng.h are the reference's own.
#include "bng/shared_defines.h"
#include "defines.h"
#include "rng.h"  // reference: src/rng.h (rng_state, rng_uint)

#include <vector>
#include <cstdio>
#include <cstdlib>

namespace MCell {

enum class CollisionType { INVALID, WALL_REDO, WALL_MISS, WALL_FRONT, WALL_BACK, VOLMOL_VOLMOL, SURFMOL_SURFMOL, VOLMOL_SURFMOL,
                           UNIMOLECULAR, INTERMEMBRANE_SURFMOL_SURFMOL };  // src4/collision_structs.h:29-42

class Partition;

struct WallCollisionRejectionData {  // src4/geometry.h
  Vec3 normal;
  pos_t distance_to_origin;
};

class Wall;
class Grid {  // src4/wall.h Grid: the members Grid::initialize and GridUtils use
public:
  uint num_tiles_along_axis = 0, num_tiles = 0, num_occupied = 0;
  pos_t strip_width_rcp, vert2_slope, fullslope, binding_factor;
  Vec2 vert0;
  wall_index_t wall_index;
  std::vector<molecule_id_t> molecules_per_tile;
  bool is_initialized() const { return num_tiles != 0; }
  void initialize(const Partition& p, const Wall& w);
};

class Wall : public WallCollisionRejectionData {
public:
  wall_index_t index = 0;
  Grid grid;
  bool has_initialized_grid() const { return grid.is_initialized(); }
  vertex_index_t vertex_indices[3];
  Vec3 unit_u, unit_v;
  pos_t uv_vert1_u;
  Vec2 uv_vert2;
  pos_t area;
  bool wall_constants_initialized = false;
  bool exists_in_partition() const { return true; }
  bool is_overlapped_wall() const { return false; }
  void initialize_wall_constants(const Partition& p);
};
class WallWithVertices : public Wall {
public:
  Vec3 vertices[3];
};

struct Molecule {
  molecule_id_t id;
  species_id_t species_id = 0;
  bool defunct = false;
  struct { Vec3 pos; subpart_index_t subpart_index; } v;
  bool is_defunct() const { return defunct; }
};

class Collision {  // the members get_closest_wall_collision / reflect_from_wall use (src4/collision_structs.h:60-175)
public:
  Collision() : type(CollisionType::INVALID), partition(nullptr), diffused_molecule_id(0), time(0), pos(0), colliding_wall_index(0) {}
  Collision(const CollisionType type_, Partition* partition_ptr, const molecule_id_t diffused_molecule_id_, const double time_,
            const Vec3& pos_, const wall_index_t colliding_wall_index_)
      : type(type_), partition(partition_ptr), diffused_molecule_id(diffused_molecule_id_), time(time_), pos(pos_),
        colliding_wall_index(colliding_wall_index_) {}
  CollisionType type;
  Partition* partition;
  molecule_id_t diffused_molecule_id;
  double time;
  Vec3 pos;
  wall_index_t colliding_wall_index;
};

typedef std::vector<wall_index_t> WallsInSubpart;

}  // namespace MCell
namespace BNG {
const int PATHWAY_INDEX_NO_RXN = -1;
typedef int rxn_class_pathway_index_t;   // libbng: the index of a pathway in its reaction class
class RxnContainer;
class RxnClass;
typedef std::vector<RxnClass*> RxnClassesVector;
const uint SPECIES_FLAG_CAN_VOLWALL = 1u << 3;
class Species {
public:
  uint flags = 0;
  bool has_flag(uint f) const { return (flags & f) != 0; }
};
class SpeciesContainer {
public:
  Species only;
  const Species& get(uint) const { return only; }
};
class RxnClass {  // stand-in for libbng's: what test_bimolecular and exact_disk call
public:
  bool is_transparent_type() const { return false; }
  std::vector<double> cum_probs;
  int get_num_reactions() const { return (int)cum_probs.size(); }
  void update_rxn_rates_if_needed(double) {}
  double get_max_fixed_p() const { return cum_probs.back(); }
  int get_pathway_index_for_probability(double prob, double mult) const {  // binary_search_double, src/util.c
    int min_idx = 0, max_idx = (int)cum_probs.size() - 1;
    while (max_idx - min_idx > 1) {
      const int mid = (max_idx + min_idx) / 2;
      if (prob > cum_probs[mid] * mult) min_idx = mid; else max_idx = mid;
    }
    return prob > cum_probs[min_idx] * mult ? max_idx : min_idx;
  }
};
}  // namespace BNG
namespace MCell {

struct Stats {
  mutable unsigned long long ray_polygon_tests = 0, ray_polygon_colls = 0;
  double skipped = 0;
  void inc_ray_polygon_tests() const { ray_polygon_tests++; }
  void inc_ray_polygon_colls() const { ray_polygon_colls++; }
  void inc_rxn_skipped(BNG::RxnContainer*, BNG::RxnClass*, double s) { skipped += s; }
};

struct PartitionConfig { bool use_expanded_list = true; pos_t rxn_radius_3d = 0, subpart_edge_length = 1000; };
class Partition {  // accessors of src4/partition.h used by the extracted functions
public:
  PartitionConfig config;
  void get_subpart_3d_indices(const Vec3& pos, IVec3& res) const {  // partition.h:253-262
    res.x = (int)((pos.x - origin_corner.x) * subpart_edge_length_rcp);
    res.y = (int)((pos.y - origin_corner.y) * subpart_edge_length_rcp);
    res.z = (int)((pos.z - origin_corner.z) * subpart_edge_length_rcp);
  }
  subpart_index_t get_subpart_index_from_3d_indices(const int x, const int y, const int z) const {
    return x + y * num_subparts_per_partition_edge + z * num_subparts_per_partition_edge * num_subparts_per_partition_edge;
  }
  std::vector<Vec3> vertices;
  std::vector<Wall> walls;
  WallsInSubpart all_walls;  // one subpartition holding every wall, ascending (walls_per_subpart, partition.h)
  Stats stats;
  BNG::RxnContainer* get_all_rxns() { return nullptr; }
  BNG::SpeciesContainer species;
  const BNG::SpeciesContainer& get_all_species() const { return species; }
  void get_subpart_llf_point(const subpart_index_t i, Vec3& llf) const {  // partition.h:301-305
    const uint n = num_subparts_per_partition_edge;
    llf = origin_corner + Vec3(IVec3(i % n, (i / n) % n, (i / (n * n)) % n)) * Vec3(config.subpart_edge_length);
  }
  void get_subpart_urb_point_from_llf(const Vec3& llf, Vec3& urb) const { urb = llf + Vec3(config.subpart_edge_length); }
  Vec3 origin_corner;
  pos_t subpart_edge_length_rcp;
  uint num_subparts_per_partition_edge;
  const Vec3& get_geometry_vertex(vertex_index_t i) const { return vertices[i]; }
  const Vec3& get_wall_vertex(const Wall& w, uint k) const { return vertices[w.vertex_indices[k]]; }
  const Wall& get_wall(wall_index_t i) const { return walls[i]; }
  const WallCollisionRejectionData& get_wall_collision_rejection_data(wall_index_t i) const { return walls[i]; }
  const WallsInSubpart& get_subpart_wall_indices(subpart_index_t) const { return all_walls; }
  subpart_index_t get_subpart_index(const Vec3& pos) const {  // partition.h:253-296
    const int x = (int)((pos.x - origin_corner.x) * subpart_edge_length_rcp), y = (int)((pos.y - origin_corner.y) * subpart_edge_length_rcp),
              z = (int)((pos.z - origin_corner.z) * subpart_edge_length_rcp);
    return x + y * num_subparts_per_partition_edge + z * num_subparts_per_partition_edge * num_subparts_per_partition_edge;
  }
};

#ifndef CHECK_STIME_MAX
#define CHECK_STIME_MAX(x) do { } while (0)
#endif
#define INLINE_ATTR __attribute__((always_inline))

#include "gen/mcell4_wall_constants.inl"  // Wall::initialize_wall_constants, src4/wall.cpp:281-342

namespace CollisionUtils {
#include "gen/mcell4_collision_utils.inl"  // the five functions of src4/collision_utils.inl listed above
}
#define mcell_internal_error(...) do { fprintf(stderr, __VA_ARGS__); abort(); } while (0)
#include "gen/mcell4_grid_initialize.inl"  // Grid::initialize, src4/wall.cpp:38-74
namespace WallUtils {
#include "gen/mcell4_wall_in_box.inl"  // src4/wall_utils.inl:326-504
}
namespace GeometryUtils {
#include "gen/mcell4_geometry_utils.inl"  // find_edge_point :222-291; same_side, point_in_triangle, cross2D, point_in_triangle_2D :352-443
}
namespace RxnUtils {  // only reached for species with surface-class reactions (none here)
static void trigger_intersect(Partition&, const Molecule&, int, const Wall&, bool, BNG::RxnClassesVector&) {}
}
namespace ExactDiskUtils {
#include "gen/mcell4_exact_disk.inl"  // src4/exact_disk_utils.inl:54-1145
}
namespace GridUtils {
#include "gen/mcell4_grid_utils.inl"  // xyz2grid_tile_index, uv2grid_tile_index, grid2uv
}
namespace DiffusionUtils {
#include "gen/mcell4_pick_surf_displacement.inl"  // src4/diffusion_utils.inl:60-93
}
namespace RxnUtils {
#include "gen/mcell4_test_bimolecular.inl"  // src4/rxn_utils.inl:336-414
#include "gen/mcell4_test_intersect.inl"    // src4/rxn_utils.inl:593-626
}

}  // namespace MCell

#define EXPORT extern "C" __attribute__((visibility("default")))
using namespace MCell;

namespace {
void fill(Partition& p, const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls) {
  p.origin_corner = Vec3(-500.0, -500.0, -500.0);
  p.subpart_edge_length_rcp = 1.0 / 1000.0;  // one subpartition: every hit lies in subpartition 0 like the start
  p.num_subparts_per_partition_edge = 1;
  for (unsigned i = 0; i < n_verts; i++) p.vertices.push_back(Vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
  p.walls.resize(n_walls);
  for (unsigned w = 0; w < n_walls; w++) {
    for (int k = 0; k < 3; k++) p.walls[w].vertex_indices[k] = tri[3 * w + k];
    p.walls[w].initialize_wall_constants(p);
    p.all_walls.push_back(w);
  }
}
void seed_rng(rng_state* r, unsigned seed, unsigned skip) {
  rng_init(r, seed);
  for (unsigned i = 0; i < skip; i++) (void)rng_uint(r);
}
int code(CollisionType t) {  // the MCell3 codes the other shims use: REDO -1, MISS 0, FRONT 1, BACK 2
  return t == CollisionType::WALL_REDO ? -1 : t == CollisionType::WALL_FRONT ? 1 : t == CollisionType::WALL_BACK ? 2 : 0;
}
}  // namespace

EXPORT void ref4_wall_constants(const double* v9, double* out16) {
  const unsigned tri[3] = {0, 1, 2};
  Partition p; fill(p, v9, 3, tri, 1);
  const Wall& w = p.walls[0];
  const double t[16] = {w.normal.x, w.normal.y, w.normal.z, w.distance_to_origin, w.unit_u.x, w.unit_u.y, w.unit_u.z,
                        w.unit_v.x, w.unit_v.y, w.unit_v.z, w.uv_vert1_u, w.uv_vert2.u, w.uv_vert2.v, w.area, 0, 0};
  for (int k = 0; k < 16; k++) out16[k] = t[k];
}

EXPORT int ref4_collide_wall(const double* point3, double* move3, const double* v9, unsigned seed, unsigned skip, double* t,
                             double* hit3, long long* rng_words_used) {
  const unsigned tri[3] = {0, 1, 2};
  Partition p; fill(p, v9, 3, tri, 1);
  rng_state rng; seed_rng(&rng, seed, skip);
  const long long before = rng_uses(&rng);
  Vec3 move(move3[0], move3[1], move3[2]), hit(0);
  stime_t tt = 0;
  const CollisionType r = CollisionUtils::collide_wall(p, Vec3(point3[0], point3[1], point3[2]), 0, rng, true, true, move, tt, hit);
  move3[0] = move.x; move3[1] = move.y; move3[2] = move.z;
  *t = tt; hit3[0] = hit.x; hit3[1] = hit.y; hit3[2] = hit.z;
  *rng_words_used = rng_uses(&rng) - before;
  return code(r);
}

// 3 = hit (COLLIDE_VOL_M of the MCell3 shim), 0 = miss
EXPORT int ref4_collide_mol(const double* point3, const double* move3, const double* target3, double rx_radius_3d, double* t,
                            double* hit3) {
  Molecule a, b;
  a.id = 1; b.id = 2;
  a.v.pos = Vec3(point3[0], point3[1], point3[2]);
  b.v.pos = Vec3(target3[0], target3[1], target3[2]);
  stime_t tt = 0; Vec3 h(0);
  const bool r = CollisionUtils::collide_mol(a, Vec3(move3[0], move3[1], move3[2]), b, rx_radius_3d, tt, h);
  *t = tt; hit3[0] = h.x; hit3[1] = h.y; hit3[2] = h.z;
  return r ? 3 : 0;
}

// One get_closest_wall_collision over a whole mesh (one subpartition), then reflect_from_wall when a wall was hit.
// Returns 1 when a wall was hit.  move3 is in/out (REDOs change it), out: wall, side (1 front / 2 back), time, hit point,
// the position and displacement after the reflection, words drawn, ray_polygon_tests.
EXPORT int ref4_closest_wall_and_reflect(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                                         const double* pos3, double* move3, unsigned last_hit_wall, unsigned seed, unsigned skip,
                                         unsigned* wall, int* side, double* t, double* hit3, double* pos_after3,
                                         double* disp_after3, double* t_steps_io, long long* rng_words_used,
                                         unsigned long long* ray_polygon_tests) {
  Partition p; fill(p, verts, n_verts, tri, n_walls);
  rng_state rng; seed_rng(&rng, seed, skip);
  const long long before = rng_uses(&rng);
  Molecule vm; vm.id = 7; vm.v.pos = Vec3(pos3[0], pos3[1], pos3[2]); vm.v.subpart_index = 0;
  Vec3 disp(move3[0], move3[1], move3[2]), up_to_wall(0);
  Collision c;
  const bool found = CollisionUtils::get_closest_wall_collision(p, vm, 0, last_hit_wall, rng, disp, up_to_wall, c);
  move3[0] = disp.x; move3[1] = disp.y; move3[2] = disp.z;
  *rng_words_used = rng_uses(&rng) - before;
  *ray_polygon_tests = p.stats.ray_polygon_tests;
  if (!found) return 0;
  *wall = c.colliding_wall_index; *side = code(c.type); *t = c.time;
  hit3[0] = c.pos.x; hit3[1] = c.pos.y; hit3[2] = c.pos.z;
  double t_steps = *t_steps_io;
  wall_index_t last = WALL_INDEX_INVALID;
  CollisionUtils::reflect_from_wall(p, c, vm, disp, t_steps, last);
  pos_after3[0] = vm.v.pos.x; pos_after3[1] = vm.v.pos.y; pos_after3[2] = vm.v.pos.z;
  disp_after3[0] = disp.x; disp_after3[1] = disp.y; disp_after3[2] = disp.z;
  *t_steps_io = t_steps;
  return 1;
}

EXPORT long long ref4_pick_surf_displacement(double scale, unsigned seed, unsigned skip, double* out2) {
  rng_state rng; seed_rng(&rng, seed, skip);
  const long long before = rng_uses(&rng);
  Vec2 v(0);
  DiffusionUtils::pick_surf_displacement(v, scale, rng);
  out2[0] = v.u; out2[1] = v.v;
  return rng_uses(&rng) - before;
}

// pathway index or -1; local_prob_factor = 0 (volume reactions)
EXPORT int ref4_test_bimolecular(const double* cum_probs, int n, double scaling, unsigned seed, unsigned skip, long long* rng_words_used) {
  Partition p;
  BNG::RxnClass rc;
  rc.cum_probs.assign(cum_probs, cum_probs + n);
  rng_state rng; seed_rng(&rng, seed, skip);
  const long long before = rng_uses(&rng);
  Molecule a, b;
  const int r = RxnUtils::test_bimolecular(p, &rc, rng, a, b, scaling, 0.0, 0.0);
  *rng_words_used = rng_uses(&rng) - before;
  return r;
}

// RxnUtils::test_intersect (src4/rxn_utils.inl:593-626), a Standard reaction with a reactive surface: pathway index or -1
EXPORT int ref4_test_intersect(const double* cum_probs, int n, double scaling, unsigned seed, unsigned skip, long long* rng_words_used) {
  BNG::RxnClass rc;
  rc.cum_probs.assign(cum_probs, cum_probs + n);
  rng_state rng; seed_rng(&rng, seed, skip);
  const long long before = rng_uses(&rng);
  const int r = RxnUtils::test_intersect(&rc, scaling, 0.0, rng);
  *rng_words_used = rng_uses(&rng) - before;
  return r;
}

namespace {
void fill_with_grid(Partition& p, const double* v9) {
  const unsigned tri[3] = {0, 1, 2};
  fill(p, v9, 3, tri, 1);
  p.walls[0].grid.initialize(p, p.walls[0]);
}
}  // namespace
// same outputs as ref3_grid_constants / ref3_xyz2grid / ref3_uv2grid / ref3_grid2uv / ref3_find_edge_point
EXPORT void ref4_grid_constants(const double* v9, double* out8) {
  Partition p; fill_with_grid(p, v9);
  const Grid& g = p.walls[0].grid;
  const double t[8] = {(double)g.num_tiles_along_axis, g.strip_width_rcp, g.vert2_slope, g.fullslope, g.binding_factor, g.vert0.u,
                       g.vert0.v, (double)g.num_tiles};
  for (int k = 0; k < 8; k++) out8[k] = t[k];
}
EXPORT int ref4_xyz2grid(const double* v9, const double* xyz3) {
  Partition p; fill_with_grid(p, v9);
  return (int)GridUtils::xyz2grid_tile_index(p, Vec3(xyz3[0], xyz3[1], xyz3[2]), p.walls[0]);
}
EXPORT int ref4_uv2grid(const double* v9, const double* uv2) {
  Partition p; fill_with_grid(p, v9);
  return (int)GridUtils::uv2grid_tile_index(Vec2(uv2[0], uv2[1]), p.walls[0]);
}
EXPORT void ref4_grid2uv(const double* v9, int idx, double* uv2) {
  Partition p; fill_with_grid(p, v9);
  const Vec2 r = GridUtils::grid2uv(p.walls[0], (tile_index_t)idx);
  uv2[0] = r.u; uv2[1] = r.v;
}
// GridUtils::grid2uv_random (src4/grid_utils.inl:256-286): a random point inside a tile; returns the words drawn
EXPORT long long ref4_grid2uv_random(const double* v9, int idx, unsigned seed, unsigned skip, double* uv2) {
  Partition p; fill_with_grid(p, v9);
  rng_state rng; seed_rng(&rng, seed, skip);
  const long long before = rng_uses(&rng);
  const Vec2 r = GridUtils::grid2uv_random(p.walls[0], (tile_index_t)idx, rng);
  uv2[0] = r.u; uv2[1] = r.v;
  return rng_uses(&rng) - before;
}
EXPORT int ref4_find_edge_point(const double* v9, const double* loc2, const double* disp2, double* edgept2) {
  const unsigned tri[3] = {0, 1, 2};
  Partition p; fill(p, v9, 3, tri, 1);
  Vec2 pt(0);
  const edge_index_t e = GeometryUtils::find_edge_point(p.walls[0], Vec2(loc2[0], loc2[1]), Vec2(disp2[0], disp2[1]), pt);
  // MCell3's codes: -1 stays inside (EDGE_INDEX_WITHIN_WALL), -2 cannot tell (EDGE_INDEX_CANNOT_TELL)
  const int r = e == EDGE_INDEX_WITHIN_WALL ? -1 : e == EDGE_INDEX_CANNOT_TELL ? -2 : (int)e;
  edgept2[0] = pt.u; edgept2[1] = pt.v;
  return r;
}

// same arguments as ref3_exact_disk: n walls (9 coordinates each) are the walls of the collision subpartition; the
// expanded list is on, as in every MCell4 run with volume-volume reactions
EXPORT double ref4_exact_disk(const double* loc3, const double* mv3, double R, const double* target3, int n_walls,
                              const double* tri9) {
  std::vector<double> verts(tri9, tri9 + 9 * (size_t)n_walls);
  std::vector<unsigned> tri(3 * (size_t)n_walls);
  for (size_t i = 0; i < tri.size(); i++) tri[i] = (unsigned)i;
  Partition p; fill(p, verts.data(), 3 * n_walls, tri.data(), n_walls);
  Molecule moving, target;
  moving.id = 1; target.id = 2;
  moving.v.pos = Vec3(loc3[0], loc3[1], loc3[2]); moving.v.subpart_index = 0;
  target.v.pos = Vec3(target3[0], target3[1], target3[2]); target.v.subpart_index = 0;
  Vec3 mv(mv3[0], mv3[1], mv3[2]);
  return ExactDiskUtils::exact_disk(p, Vec3(loc3[0], loc3[1], loc3[2]), mv, R, moving, target, true);
}

// Partition::finalize_walls (partition.cpp:91-118): for every wall the subpartitions wall_subparts_collision_test puts it
// into -> CSR over subpartitions with ascending wall indices (walls_per_subpart is a uint_set).  Returns the number of
// entries (list_out holds up to cap of them).
EXPORT unsigned long long ref4_walls_per_subpart(const double* origin3, double partition_edge_length, unsigned n_subparts_per_edge,
                                                 double rxn_radius_3d, int use_expanded_list, const double* verts, unsigned n_verts,
                                                 const unsigned* tri, unsigned n_walls, unsigned* start_out, unsigned* list_out,
                                                 unsigned long long cap) {
  Partition p; fill(p, verts, n_verts, tri, n_walls);
  p.origin_corner = Vec3(origin3[0], origin3[1], origin3[2]);
  p.num_subparts_per_partition_edge = n_subparts_per_edge;
  p.config.subpart_edge_length = partition_edge_length / n_subparts_per_edge;   // simulation_config.cpp:48
  p.subpart_edge_length_rcp = 1.0 / p.config.subpart_edge_length;
  p.config.use_expanded_list = use_expanded_list != 0;
  p.config.rxn_radius_3d = rxn_radius_3d;
  const size_t ns = (size_t)n_subparts_per_edge * n_subparts_per_edge * n_subparts_per_edge;
  std::vector<std::vector<unsigned>> per(ns);
  for (unsigned w = 0; w < n_walls; w++) {
    SubpartIndicesVector hit;
    GeometryUtils::wall_subparts_collision_test(p, p.walls[w], hit);
    for (subpart_index_t sidx : hit) per[sidx].push_back(w);
  }
  unsigned long long k = 0;
  for (size_t sidx = 0; sidx < ns; sidx++) {
    start_out[sidx] = (unsigned)k;
    for (unsigned w : per[sidx]) { if (k < cap) list_out[k] = w; k++; }
  }
  start_out[ns] = (unsigned)k;
  return k;
}
