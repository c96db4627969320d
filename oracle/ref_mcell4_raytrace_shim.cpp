// oracle/ref_mcell4_raytrace_shim.cpp — oracle/_ref build only (TEST INFRASTRUCTURE).
//
// Builds MCell4's OWN ray_trace_vol — the function that composes the subpartition walk, the wall test and the
// molecule test into "what does this move of a volume molecule collide with" — into oracle/_ref/libmcell4raytrace.so:
//   ray_trace_vol                                        src4/diffuse_react_event.cpp:627-780
//   sort_collisions_by_time                              src4/diffuse_react_event.cpp:341-364
//   CollisionUtils::collect_crossed_subparts,
//   collect_neighboring_subparts                         src4/collision_utils_subparts.inl (whole file, #included)
//   CollisionUtils::get_displacement_up_to_partition_boundary   src4/collision_utils.inl:48-136
//   CollisionUtils::collide_mol, collide_mol_loop_body   :464-566
//   CollisionUtils::jump_away_line, collide_wall,
//   is_immediate_collision, get_closest_wall_collision   :568-603, 629-914
//   Wall::initialize_wall_constants                      src4/wall.cpp:281-342
// The function texts are cut out of the reference files BY LINE RANGE AT BUILD TIME (oracle/Makefile: ref, into the
// git-ignored oracle/_ref/gen/) and compiled unmodified; nothing of them is stored in this repository.  The types they
// touch are stand-ins with the reference's member names (src4/partition.h accessors, src4/geometry.h Wall,
// src4/molecule.h Molecule, src4/collision_structs.h Collision); src4/defines.h with the reference's libs/glm and
// src/rng.h are the reference's own.  The per-subpartition wall lists are GIVEN by the caller (the distribution itself
// is pinned separately: ref4_walls_per_subpart of the leaf shim); the per-subpartition reactant sets are
// uint_set<molecule_id_t> like partition.h:1125-1143 keeps them (all molecules of the species that react with the mover).
#include "bng/shared_defines.h"
#include "defines.h"
#include "rng.h"  // reference: src/rng.h (rng_state, rng_uint)

#include <algorithm>
#include <cassert>
#include <map>
#include <vector>
#include <cstdio>
#include <cstdlib>

#define SRC4_DIFFUSE_REACT_EVENT_H_
#define SRC4_WORLD_H_
#define SRC4_PARTITION_H_
#define SRC4_GEOMETRY_H_
#define SRC4_GEOMETRY_UTILS_INC_

namespace BNG {
class RxnClass { public: int id = 0; };
class RxnContainer {  // libbng's container: one class per unordered species pair
public:
  std::map<std::pair<uint, uint>, RxnClass*> classes;
  RxnClass* get_bimol_rxn_class(uint a, uint b) {
    auto it = classes.find(std::make_pair(std::min(a, b), std::max(a, b)));
    return it == classes.end() ? nullptr : it->second;
  }
};
}  // namespace BNG

namespace MCell {

enum class CollisionType { INVALID, WALL_REDO, WALL_MISS, WALL_FRONT, WALL_BACK, VOLMOL_VOLMOL, SURFMOL_SURFMOL, VOLMOL_SURFMOL,
                           UNIMOLECULAR, INTERMEMBRANE_SURFMOL_SURFMOL };  // src4/collision_structs.h:29-42
enum class RayTraceState { UNDEFINED, HIT_SUBPARTITION, RAY_TRACE_HIT_WALL, FINISHED };  // src4/diffuse_react_event.h:39-44

class Partition;

struct WallCollisionRejectionData {  // src4/geometry.h
  Vec3 normal;
  pos_t distance_to_origin;
};

class Wall : public WallCollisionRejectionData {
public:
  wall_index_t index = 0;
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

class Collision {  // the members ray_trace_vol and what it calls use (src4/collision_structs.h:60-175)
public:
  Collision() : type(CollisionType::INVALID), partition(nullptr), diffused_molecule_id(0), time(0), pos(0),
                colliding_molecule_id(MOLECULE_ID_INVALID), rxn_class(nullptr), colliding_wall_index(WALL_INDEX_INVALID) {}
  Collision(const CollisionType type_, Partition* partition_ptr, const molecule_id_t diffused_molecule_id_, const double time_,
            const Vec3& pos_, const molecule_id_t colliding_molecule_id_, BNG::RxnClass* rxn_class_ptr)
      : type(type_), partition(partition_ptr), diffused_molecule_id(diffused_molecule_id_), time(time_), pos(pos_),
        colliding_molecule_id(colliding_molecule_id_), rxn_class(rxn_class_ptr), colliding_wall_index(WALL_INDEX_INVALID) {}
  Collision(const CollisionType type_, Partition* partition_ptr, const molecule_id_t diffused_molecule_id_, const double time_,
            const Vec3& pos_, const wall_index_t colliding_wall_index_)
      : type(type_), partition(partition_ptr), diffused_molecule_id(diffused_molecule_id_), time(time_), pos(pos_),
        colliding_molecule_id(MOLECULE_ID_INVALID), rxn_class(nullptr), colliding_wall_index(colliding_wall_index_) {}
  CollisionType type;
  Partition* partition;
  molecule_id_t diffused_molecule_id;
  double time;
  Vec3 pos;
  molecule_id_t colliding_molecule_id;
  BNG::RxnClass* rxn_class;
  wall_index_t colliding_wall_index;
};
typedef std::vector<Collision> CollisionsVector;  // collision_structs.h:53 (INDEXER_WA)

typedef std::vector<wall_index_t> WallsInSubpart;

struct Stats {
  mutable unsigned long long ray_polygon_tests = 0, ray_polygon_colls = 0, ray_voxel_tests = 0;
  void inc_ray_polygon_tests() const { ray_polygon_tests++; }
  void inc_ray_polygon_colls() const { ray_polygon_colls++; }
  void inc_ray_voxel_tests() const { ray_voxel_tests++; }
};

struct SimulationConfig {
  pos_t partition_edge_length;
  uint num_subparts_per_partition_edge, num_subparts_per_partition_edge_squared;
  pos_t subpart_edge_length, subpart_edge_length_rcp;
  bool use_expanded_list;
  pos_t rxn_radius_3d;
};

class Partition {  // accessors of src4/partition.h used by the extracted functions
public:
  SimulationConfig config;
  Vec3 origin_corner, opposite_corner;
  const Vec3& get_origin_corner() const { return origin_corner; }
  bool in_this_partition(const Vec3& pos) const {  // partition.h:248-251
    return glm::all(glm::greaterThanEqual(pos, origin_corner)) && glm::all(glm::lessThan(pos, opposite_corner));
  }
  bool is_subpart_index_in_range(const int index) const { return index >= 0 && index < (int)config.num_subparts_per_partition_edge; }
  void get_subpart_3d_indices(const Vec3& pos, IVec3& res) const {  // partition.h:257-262
    Vec3 relative_position = pos - origin_corner;
    res = relative_position * config.subpart_edge_length_rcp;
  }
  subpart_index_t get_subpart_index_from_3d_indices_allow_outside(const IVec3& i) const {
    return i.x + i.y * config.num_subparts_per_partition_edge + i.z * config.num_subparts_per_partition_edge_squared;
  }
  subpart_index_t get_subpart_index_from_3d_indices(const IVec3& i) const { return get_subpart_index_from_3d_indices_allow_outside(i); }
  subpart_index_t get_subpart_index_from_3d_indices(const int x, const int y, const int z) const {
    return get_subpart_index_from_3d_indices(IVec3(x, y, z));
  }
  void get_subpart_3d_indices_from_index(const subpart_index_t index, IVec3& i) const {
    const uint dim = config.num_subparts_per_partition_edge;
    i.x = index % dim; i.y = (index / dim) % dim; i.z = (index / config.num_subparts_per_partition_edge_squared) % dim;
  }
  subpart_index_t get_subpart_index(const Vec3& pos) const {  // partition.h:285-290
    IVec3 i; get_subpart_3d_indices(pos, i);
    return get_subpart_index_from_3d_indices(i);
  }

  std::vector<Vec3> vertices;
  std::vector<Wall> walls;
  std::vector<WallsInSubpart> walls_per_subpart;
  std::vector<Molecule> molecules;  // index == id
  std::map<std::pair<subpart_index_t, species_id_t>, MoleculeIdsSet> reactants;  // [subpart, species of the mover]
  MoleculeIdsSet empty_set;
  Stats stats;
  BNG::RxnContainer rxns;
  BNG::RxnContainer& get_all_rxns() { return rxns; }
  Molecule& get_m(const molecule_id_t id) { return molecules[id]; }
  const MoleculeIdsSet& get_volume_molecule_reactants(const subpart_index_t s, const species_id_t species_id) const {
    auto it = reactants.find(std::make_pair(s, species_id));
    return it == reactants.end() ? empty_set : it->second;
  }
  const Vec3& get_geometry_vertex(vertex_index_t i) const { return vertices[i]; }
  const Vec3& get_wall_vertex(const Wall& w, uint k) const { return vertices[w.vertex_indices[k]]; }
  const Wall& get_wall(wall_index_t i) const { return walls[i]; }
  const WallCollisionRejectionData& get_wall_collision_rejection_data(wall_index_t i) const { return walls[i]; }
  const WallsInSubpart& get_subpart_wall_indices(subpart_index_t s) const { return walls_per_subpart[s]; }
};

#ifndef CHECK_STIME_MAX
#define CHECK_STIME_MAX(x) do { } while (0)
#endif
#define INLINE_ATTR __attribute__((always_inline))

#include "gen/mcell4_wall_constants.inl"  // Wall::initialize_wall_constants, src4/wall.cpp:281-342

}  // namespace MCell

#include "collision_utils_subparts.inl"  // the reference's file, whole

namespace MCell {
namespace CollisionUtils {
#include "gen/mcell4_collision_utils_raytrace.inl"  // src4/collision_utils.inl:48-136, 464-603, 629-914
}
using namespace std;
#include "gen/mcell4_sort_collisions.inl"  // sort_collisions_by_time, src4/diffuse_react_event.cpp:341-364
// ray_trace_vol's NDEBUG branches call the dense-hash-set API (resize / clear_no_resize: capacity hints only); with the
// std::set that defines.h:315-317 selects under INDEXER_WA the function compiles in its debug form, asserts switched off
#undef NDEBUG
#undef assert
#define assert(x) ((void)0)
#include "gen/mcell4_ray_trace_vol.inl"    // ray_trace_vol, src4/diffuse_react_event.cpp:627-780
#define NDEBUG
}  // namespace MCell

#define EXPORT extern "C" __attribute__((visibility("default")))
using namespace MCell;

// One ray_trace_vol of molecule `mol_id` of a population in a partition of n_subparts_per_edge^3 subpartitions, then
// sort_collisions_by_time when there is more than one collision (diffuse_vol_molecule :434-436).
//   subpart_wall_off / subpart_walls: CSR of the wall indices per subpartition (ascending)
//   mol_species / mol_pos: the population (id = index); reacts[a * n_species + b] != 0: species a and b have a class
//   disp3: in/out (a REDO changes it); seed / skip: the ISAAC64 stream handed to the function
// Returns 1 = RAY_TRACE_HIT_WALL, 0 = FINISHED.  type 0 = molecule (what = its id), 1 / 2 = wall front / back (what = wall).
EXPORT int ref4_ray_trace_vol(const double* origin3, double partition_edge_length, unsigned n_subparts_per_edge, double rxn_radius_3d,
                              const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                              const unsigned* subpart_wall_off, const unsigned* subpart_walls,
                              const unsigned* mol_species, const double* mol_pos, unsigned n_mols, const unsigned char* reacts,
                              unsigned n_species, unsigned mol_id, double* disp3, unsigned last_hit_wall, unsigned seed, unsigned skip,
                              int cap, int* n_coll, int* type, double* time, double* pos3, unsigned* what, long long* rng_words_used,
                              double* pos_after3, unsigned* subpart_after) {
  Partition p;
  p.config.partition_edge_length = partition_edge_length;
  p.config.num_subparts_per_partition_edge = n_subparts_per_edge;
  p.config.num_subparts_per_partition_edge_squared = n_subparts_per_edge * n_subparts_per_edge;
  p.config.subpart_edge_length = partition_edge_length / n_subparts_per_edge;   // simulation_config.cpp:48
  p.config.subpart_edge_length_rcp = 1.0 / p.config.subpart_edge_length;         // :63
  p.config.use_expanded_list = true;
  p.config.rxn_radius_3d = rxn_radius_3d;
  p.origin_corner = Vec3(origin3[0], origin3[1], origin3[2]);
  p.opposite_corner = p.origin_corner + Vec3(partition_edge_length);
  for (unsigned i = 0; i < n_verts; i++) p.vertices.push_back(Vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
  p.walls.resize(n_walls);
  for (unsigned w = 0; w < n_walls; w++) {
    p.walls[w].index = w;
    for (int k = 0; k < 3; k++) p.walls[w].vertex_indices[k] = tri[3 * w + k];
    p.walls[w].initialize_wall_constants(p);
  }
  const unsigned n_sp = n_subparts_per_edge * n_subparts_per_edge * n_subparts_per_edge;
  p.walls_per_subpart.resize(n_sp);
  for (unsigned s = 0; s < n_sp; s++)
    for (unsigned k = subpart_wall_off[s]; k < subpart_wall_off[s + 1]; k++) p.walls_per_subpart[s].push_back(subpart_walls[k]);
  std::vector<BNG::RxnClass> classes(n_species * n_species);
  for (unsigned a = 0; a < n_species; a++)
    for (unsigned b = a; b < n_species; b++)
      if (reacts[a * n_species + b]) p.rxns.classes[std::make_pair(a, b)] = &classes[a * n_species + b];
  p.molecules.resize(n_mols);
  for (unsigned i = 0; i < n_mols; i++) {
    Molecule& m = p.molecules[i];
    m.id = i; m.species_id = mol_species[i];
    m.v.pos = Vec3(mol_pos[3 * i], mol_pos[3 * i + 1], mol_pos[3 * i + 2]);
    m.v.subpart_index = p.get_subpart_index(m.v.pos);
    for (unsigned a = 0; a < n_species; a++)  // change_vol_reactants_map_from_species, partition.h:1125-1143
      if (reacts[a * n_species + m.species_id]) p.reactants[std::make_pair(m.v.subpart_index, (species_id_t)a)].insert(i);
  }
  rng_state rng;
  rng_init(&rng, seed);
  for (unsigned i = 0; i < skip; i++) (void)rng_uint(&rng);
  const long long before = rng_uses(&rng);
  Vec3 remaining(disp3[0], disp3[1], disp3[2]);
  CollisionsVector colls;
  bool can_vol_react = false;
  for (unsigned b = 0; b < n_species; b++) can_vol_react = can_vol_react || reacts[mol_species[mol_id] * n_species + b];
  const RayTraceState st = ray_trace_vol(p, rng, mol_id, can_vol_react, last_hit_wall, remaining, colls);
  if (colls.size() > 1) sort_collisions_by_time(colls);
  *rng_words_used = rng_uses(&rng) - before;
  disp3[0] = remaining.x; disp3[1] = remaining.y; disp3[2] = remaining.z;
  *n_coll = (int)colls.size();
  for (int k = 0; k < (int)colls.size() && k < cap; k++) {
    const Collision& c = colls[k];
    const bool mol = c.type == CollisionType::VOLMOL_VOLMOL;
    type[k] = mol ? 0 : c.type == CollisionType::WALL_FRONT ? 1 : 2;
    time[k] = c.time;
    pos3[3 * k] = c.pos.x; pos3[3 * k + 1] = c.pos.y; pos3[3 * k + 2] = c.pos.z;
    what[k] = mol ? c.colliding_molecule_id : c.colliding_wall_index;
  }
  const Molecule& vm = p.molecules[mol_id];
  pos_after3[0] = vm.v.pos.x; pos_after3[1] = vm.v.pos.y; pos_after3[2] = vm.v.pos.z;
  *subpart_after = vm.v.subpart_index;
  return st == RayTraceState::RAY_TRACE_HIT_WALL ? 1 : 0;
}
