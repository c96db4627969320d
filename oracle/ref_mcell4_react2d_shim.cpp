// oracle/ref_mcell4_react2d_shim.cpp — oracle/_ref build only (TEST INFRASTRUCTURE).
//
// Builds MCell4's OWN DiffuseReactEvent::react_2D_all_neighbors (src4/diffuse_react_event.cpp:1249-1393: the reaction of a
// surface molecule with the molecules on the tiles around its own) with RxnUtils::trigger_bimolecular (src4/rxn_utils.inl:58-97)
// into oracle/_ref/libmcell4react2d.so, on top of the neighbour-tile search and the reaction tests of
// ref_mcell4_tiles_shim.cpp (same stand-ins, below).  outcome_bimolecular is a recorder: the function under test ends where
// the reaction it chose would be carried out.  No species interacts with region borders.
//
// Builds MCell4's OWN neighbour-tile search of the surface grids into oracle/_ref/libmcell4tiles.so:
//   GridUtils::is_inner_tile, is_corner_tile, grid_neighbors, tile_orientation, move_strip_up / move_strip_down,
//   find_shared_vertices_corner_tile_parent_wall, find_shared_vertices_for_neighbor_walls,
//   grid_all_neighbors_across_walls_through_vertices, bisect / bisect_high, add_more_tile_neighbors_to_list_fast,
//   grid_all_neighbors_across_walls_through_edges, grid_all_neighbors_for_inner_tile,
//   find_neighbor_tiles                              src4/grid_utils.inl:296-1801
//   GridUtils::uv2grid_tile_index, grid2xyz, grid2uv  src4/grid_utils.inl:120-191, 205-253
//   WallUtils::walls_share_full_edge, find_nbr_walls_shared_one_vertex   src4/wall_utils.inl:50-65, 79-104
//   Wall::initialize_wall_constants, Grid::initialize src4/wall.cpp:281-342, 38-74
//   RxnUtils::binary_search_double, test_bimolecular, test_many_bimolecular   src4/rxn_utils.inl:301-320, 336-414, 475-580
// The function texts are cut out of the reference files BY LINE RANGE AT BUILD TIME (oracle/Makefile: ref, into the
// git-ignored oracle/_ref/gen/) and compiled unmodified; nothing of them is stored in this repository.  The types they
// touch are stand-ins with the reference's member names (src4/wall.h Wall / Grid, src4/partition.h accessors,
// src4/diffuse_react_event.h TileNeighborVector); src4/defines.h with the reference's libs/glm is the reference's own.
// The species of the searching molecule cannot interact with region borders here (can_interact_with_border() false), so
// the restricted-region branches compile against empty stand-ins and never run; grid_neighbors' look across a wall edge
// (get_grid_neighbors_single_grid_and_index) is only reached for non-inner tiles, which find_neighbor_tiles never sends
// there: GeometryUtils::closest_interior_point aborts if it is ever called.
#include "bng/shared_defines.h"
#include "defines.h"
#include "rng.h"  // reference: src/rng.h

#include <vector>
#include <deque>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

template <class T> using small_vector = std::vector<T>;  // libbng's alias of boost::container::small_vector (absent)

namespace MCell {

class Partition;
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
  molecule_id_t get_molecule_on_tile(tile_index_t t) const { return t < molecules_per_tile.size() ? molecules_per_tile[t] : MOLECULE_ID_INVALID; }
};

class Wall {
public:
  wall_index_t index = 0, id = 0;
  Grid grid;
  bool has_initialized_grid() const { return grid.is_initialized(); }
  void initialize_grid(const Partition& p) { grid.initialize(p, *this); }
  vertex_index_t vertex_indices[3];
  wall_index_t nb_walls[3] = {WALL_INDEX_INVALID, WALL_INDEX_INVALID, WALL_INDEX_INVALID};
  Vec3 normal, unit_u, unit_v;
  pos_t distance_to_origin, uv_vert1_u;
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
  molecule_id_t id = 0;
  species_id_t species_id = 0;
  struct { Vec2 pos; wall_index_t wall_index; tile_index_t grid_tile_index; orientation_t orientation; } s;
};

class TileNeighborVector : public std::deque<WallTileIndexPair> {};  // src4/diffuse_react_event.h:54-62

}  // namespace MCell
namespace BNG {
const int PATHWAY_INDEX_NO_RXN = -1;
typedef int rxn_class_pathway_index_t;
class RxnContainer;
const uint SPECIES_FLAG_CAN_REGION_BORDER = 1u << 9;
const int PATHWAY_INDEX_LEAST_VALID = 0;
class Species {
public:
  bool has_flag(uint) const { return false; }
  bool can_interact_with_border() const { return false; }
};
class RxnClass {  // stand-in for libbng's: what test_bimolecular / test_many_bimolecular / trigger_bimolecular call
public:
  std::vector<double> cum_probs;
  int geom[2] = {0, 0};
  int index = -1;
  bool is_bimol() const { return true; }
  int get_reactant_orientation(uint i) const { return geom[i]; }
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
typedef std::vector<RxnClass*> RxnClassesVector;
class RxnContainerImpl {
public:
  std::vector<RxnClass*> table; uint n_species = 0;   // [a * n_species + b]
  RxnClass* get_bimol_rxn_class(uint a, uint b) { return table[a * n_species + b]; }
};
class BNGEngine {
public:
  RxnContainerImpl rxns;
  RxnContainerImpl& get_all_rxns() { return rxns; }
};
}  // namespace BNG
namespace MCell {

struct Stats {
  double skipped = 0;
  void inc_rxn_skipped(BNG::RxnContainer*, BNG::RxnClass*, double s) { skipped += s; }
};

class Partition {  // accessors of src4/partition.h used by the extracted functions
public:
  std::vector<Vec3> vertices;
  std::vector<Wall> walls;
  std::vector<std::vector<wall_index_t>> walls_using_vertex_mapping;  // ascending wall indices (Partition::add_wall order)
  BNG::Species species;
  BNG::BNGEngine bng_engine;
  std::vector<Molecule> molecules;   // molecule id = index
  Molecule no_molecule;
  Stats stats;
  BNG::RxnContainer* get_all_rxns() { return nullptr; }
  const Vec3& get_geometry_vertex(vertex_index_t i) const { return vertices[i]; }
  const Vec3& get_wall_vertex(const Wall& w, uint k) const { return vertices[w.vertex_indices[k]]; }
  const Wall& get_wall(wall_index_t i) const { return walls[i]; }
  Wall& get_wall(wall_index_t i) { return walls[i]; }
  Wall* get_wall_if_exists(wall_index_t i) { return i == WALL_INDEX_INVALID ? nullptr : &walls[i]; }
  const std::vector<wall_index_t>& get_walls_using_vertex(vertex_index_t v) const { return walls_using_vertex_mapping[v]; }
  Molecule& get_m(molecule_id_t id) { return molecules[id]; }
  const Molecule& get_m(molecule_id_t id) const { return id < molecules.size() ? molecules[id] : no_molecule; }
  const BNG::Species& get_species(species_id_t) const { return species; }
};

#define mcell_internal_error(...) do { fprintf(stderr, __VA_ARGS__); abort(); } while (0)
#include "gen/mcell4_wall_constants.inl"   // Wall::initialize_wall_constants, src4/wall.cpp:281-342
#include "gen/mcell4_grid_initialize.inl"  // Grid::initialize, src4/wall.cpp:38-74

namespace GeometryUtils {
#include "gen/mcell4_geometry_utils_2d.inl"  // cross2D, point_in_triangle_2D (src4/geometry_utils.inl:409-443)
static inline Vec3 uv2xyz(const Vec2&, const Wall&, const Vec3&) { abort(); }
static inline pos_t closest_interior_point(Partition&, const Vec3&, const Wall&, Vec2&) { abort(); }
}
namespace WallUtils {
#include "gen/mcell4_wall_utils_nbr.inl"  // walls_share_full_edge, find_nbr_walls_shared_one_vertex
// never run: the stand-in species cannot interact with region borders
static void find_restricted_regions_by_wall(const Partition&, const Wall&, const Molecule&, uint_set<region_index_t>&) {}
static bool wall_belongs_to_all_regions_in_region_list(const Wall&, const uint_set<region_index_t>&) { return true; }
}
namespace GridUtils {
#include "gen/mcell4_grid_utils_tiles.inl"  // src4/grid_utils.inl:120-191, 205-253, 296-1801
}
namespace RxnUtils {
#include "gen/mcell4_test_many_bimolecular.inl"  // src4/rxn_utils.inl:301-320, 336-414, 475-580
}

}  // namespace MCell


namespace MCell {
enum class CollisionType { INVALID, SURFMOL_SURFMOL };
class Collision {   // src4/collision_structs.h: what react_2D_all_neighbors constructs
public:
  Collision() {}
  Collision(CollisionType, Partition*, molecule_id_t diffused, double time_, molecule_id_t colliding, BNG::RxnClass* rxn)
      : diffused_molecule_id(diffused), colliding_molecule_id(colliding), time(time_), rxn_class(rxn) {}
  molecule_id_t diffused_molecule_id = MOLECULE_ID_INVALID, colliding_molecule_id = MOLECULE_ID_INVALID;
  double time = 0;
  BNG::RxnClass* rxn_class = nullptr;
};
namespace WallUtils {
static inline bool walls_belong_to_at_least_one_different_restricted_region(Partition&, const Wall&, const Molecule&, const Wall&, const Molecule&) { return false; }
}
namespace RxnUtils {
#include "gen/mcell4_trigger_bimolecular.inl"   // src4/rxn_utils.inl:58-97
}
using BNG::RxnClassesVector; using BNG::RxnClass; using BNG::SPECIES_FLAG_CAN_REGION_BORDER;   // diffuse_react_event.cpp: using namespace BNG
struct World { rng_state rng; };
class DiffuseReactEvent {
public:
  World* world;
  Collision last; int last_pathway = -7; double last_time = 0; int n_outcomes = 0;
  int outcome_bimolecular(Partition&, const Collision& c, const int path, const double time) {   // recorder
    last = c; last_pathway = path; last_time = time; n_outcomes++;
    return 1;  // RX_A_OK
  }
  bool react_2D_all_neighbors(Partition& p, Molecule& sm, const double time, const double diffusion_start_time);
};
#include "gen/mcell4_react_2d_all_neighbors.inl"   // src4/diffuse_react_event.cpp:1249-1393
}  // namespace MCell

#define EXPORT extern "C" __attribute__((visibility("default")))
using namespace MCell;

namespace {
void fill(Partition& p, const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls, const unsigned char* grid_init) {
  for (unsigned i = 0; i < n_verts; i++) p.vertices.push_back(Vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
  p.walls.resize(n_walls);
  p.walls_using_vertex_mapping.resize(n_verts);
  for (unsigned w = 0; w < n_walls; w++) {
    Wall& f = p.walls[w];
    f.index = f.id = w;
    for (int k = 0; k < 3; k++) { f.vertex_indices[k] = tri[3 * w + k]; p.walls_using_vertex_mapping[tri[3 * w + k]].push_back(w); }
    f.initialize_wall_constants(p);
  }
  for (unsigned w = 0; w < n_walls; w++)
    for (int k = 0; k < 3; k++) {
      const unsigned a = tri[3 * w + k], b = tri[3 * w + (k + 1) % 3];
      for (wall_index_t o : p.walls_using_vertex_mapping[a]) {
        if (o == w) continue;
        const unsigned* t = tri + 3 * o;
        if (t[0] == b || t[1] == b || t[2] == b) { p.walls[w].nb_walls[k] = o; break; }
      }
    }
  for (unsigned w = 0; w < n_walls; w++)
    if (!grid_init || grid_init[w]) p.walls[w].initialize_grid(p);
}
}  // namespace

// react_2D_all_neighbors for every molecule of a population of surface molecules, each against the same state (the
// function's outcome is recorded, not carried out) and with its own stream rng_init(seeds[i]).
//   mols: n_mols x (wall, tile, species, orientation) as int32; walls in grid_init == 0 have no grid
//   classes: table[a * n_species + b] = class index or -1; per class: geometry of reactant 0 / 1, pathways (cum_probs)
// out per molecule: partner molecule index (-1 none), class index, pathway, words drawn
EXPORT void ref4_react_2d_all_neighbors(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                                        const unsigned char* grid_init, const int* mols, unsigned n_mols, unsigned n_species,
                                        const int* class_table, unsigned n_classes, const int* class_geom, const int* class_n_pathways,
                                        const double* cum_probs, double t_steps, const unsigned* seeds, int* out4) {
  Partition p; fill(p, verts, n_verts, tri, n_walls, grid_init);
  std::vector<BNG::RxnClass> cls(n_classes);
  int q = 0;
  for (unsigned c = 0; c < n_classes; c++) {
    cls[c].index = (int)c; cls[c].geom[0] = class_geom[2 * c]; cls[c].geom[1] = class_geom[2 * c + 1];
    for (int k = 0; k < class_n_pathways[c]; k++) cls[c].cum_probs.push_back(cum_probs[q++]);
  }
  p.bng_engine.rxns.n_species = n_species;
  p.bng_engine.rxns.table.assign((size_t)n_species * n_species, nullptr);
  for (unsigned i = 0; i < n_species * n_species; i++) if (class_table[i] >= 0) p.bng_engine.rxns.table[i] = &cls[class_table[i]];
  p.molecules.resize(n_mols);
  for (unsigned i = 0; i < n_mols; i++) {
    Molecule& m = p.molecules[i];
    m.id = i; m.s.wall_index = (wall_index_t)mols[4 * i]; m.s.grid_tile_index = (tile_index_t)mols[4 * i + 1];
    m.species_id = (species_id_t)mols[4 * i + 2]; m.s.orientation = (orientation_t)mols[4 * i + 3];
    p.walls[m.s.wall_index].grid.molecules_per_tile[m.s.grid_tile_index] = i;
  }
  for (unsigned i = 0; i < n_mols; i++) {
    World world; rng_init(&world.rng, seeds[i]);
    const long long before = rng_uses(&world.rng);
    DiffuseReactEvent ev; ev.world = &world;
    ev.react_2D_all_neighbors(p, p.molecules[i], t_steps, 0.0);
    out4[4 * i + 3] = (int)(rng_uses(&world.rng) - before);
    if (ev.n_outcomes) { out4[4 * i] = (int)ev.last.colliding_molecule_id; out4[4 * i + 1] = ev.last.rxn_class->index; out4[4 * i + 2] = ev.last_pathway; }
    else { out4[4 * i] = -1; out4[4 * i + 1] = -1; out4[4 * i + 2] = -1; }
  }
}
