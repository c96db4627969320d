// oracle/ref_mcell4_place_shim.cpp — oracle/_ref build only (TEST INFRASTRUCTURE).
//
// Builds MCell4's OWN DiffuseReactEvent::find_surf_product_positions (src4/diffuse_react_event.cpp:1993-2288: where the surface
// products of a reaction go — recycled tiles, vacant neighbour tiles, RX_BLOCKED) with GridPos (src4/wall.h:381-478) into
// oracle/_ref/libmcell4place.so, on top of the neighbour-tile search of ref_mcell4_tiles_shim.cpp (same stand-ins, below).
// Region restrictions (RegionUtils::determine_molecule_region_topology / product_tile_can_be_reached) are stand-ins for a model
// without restrictive regions; libbng's RxnRule / Cplx are stand-ins with the members the function calls.
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
  bool surf = true;
  bool is_surf() const { return surf; }
  bool is_vol() const { return !surf; }
  struct { Vec2 pos; wall_index_t wall_index; tile_index_t grid_tile_index; } s;
};

class TileNeighborVector : public std::deque<WallTileIndexPair> {};  // src4/diffuse_react_event.h:54-62

}  // namespace MCell
namespace BNG {
const int PATHWAY_INDEX_NO_RXN = -1;
typedef int rxn_class_pathway_index_t;
class RxnContainer;
typedef uint compartment_id_t;
class Species {
public:
  bool surf = true;
  bool is_surf() const { return surf; }
  bool is_vol() const { return !surf; }
  bool can_interact_with_border() const { return false; }
  compartment_id_t get_primary_compartment_id() const { return 0; }
};
class Cplx {   // one product pattern of a rule
public:
  bool surf = true;
  bool is_surf() const { return surf; }
  compartment_id_t get_primary_compartment_id() const { return 0; }
};
class RxnRule {
public:
  bool unimol = false;
  std::vector<Cplx> products;
  bool is_unimol() const { return unimol; }
  bool is_intermembrane_surf_rxn() const { return false; }
  bool is_reactive_surface_rxn() const { return false; }
};
class RxnClass {  // stand-in for libbng's: what test_bimolecular / test_many_bimolecular call
public:
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
typedef std::vector<RxnClass*> RxnClassesVector;
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
  std::vector<BNG::Species> all_species;
  Molecule no_molecule;
  Stats stats;
  BNG::RxnContainer* get_all_rxns() { return nullptr; }
  const Vec3& get_geometry_vertex(vertex_index_t i) const { return vertices[i]; }
  const Vec3& get_wall_vertex(const Wall& w, uint k) const { return vertices[w.vertex_indices[k]]; }
  const Wall& get_wall(wall_index_t i) const { return walls[i]; }
  Wall& get_wall(wall_index_t i) { return walls[i]; }
  Wall* get_wall_if_exists(wall_index_t i) { return i == WALL_INDEX_INVALID ? nullptr : &walls[i]; }
  const std::vector<wall_index_t>& get_walls_using_vertex(vertex_index_t v) const { return walls_using_vertex_mapping[v]; }
  const Molecule& get_m(molecule_id_t) const { return no_molecule; }
  const BNG::Species& get_species(species_id_t id) const { return id < all_species.size() ? all_species[id] : species; }
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


#ifndef release_assert
#define release_assert(x) do { if (!(x)) { fprintf(stderr, "release_assert failed: %s\n", #x); abort(); } } while (0)
#endif
namespace MCell {
const uint UINT_INVALID = 0xFFFFFFFFu;   // libbng's shared_defines.h (absent)
#include "gen/mcell4_gridpos.inl"   // GridPosType, GridPos (src4/wall.h:381-478)
struct ProductSpeciesIdWIndices { species_id_t product_species_id; };
typedef std::vector<ProductSpeciesIdWIndices> RxnProductsVector;
class Collision {
public:
  wall_index_t colliding_wall_index = WALL_INDEX_INVALID;
  Vec3 pos;
  bool is_wall_collision() const { return false; }
};
namespace GeometryUtils { static inline Vec2 xyz2uv(const Partition&, const Vec3&, const Wall&) { abort(); } }
namespace RegionUtils {   // a model without restrictive regions
static inline int determine_molecule_region_topology(Partition&, const Molecule*, const Molecule*, bool, RegionIndicesSet&, RegionIndicesSet&,
                                                     RegionIndicesSet&, RegionIndicesSet&) { return 0; }
static inline bool product_tile_can_be_reached(Partition&, wall_index_t, bool, int, RegionIndicesSet&, RegionIndicesSet&, RegionIndicesSet&,
                                               RegionIndicesSet&) { return true; }
}
struct World { rng_state rng; };
class DiffuseReactEvent {
public:
  World* world;
  int find_surf_product_positions(Partition& p, const Collision& collision, const BNG::RxnRule* rxn, const Molecule* reacA,
                                  const bool keep_reacA, const Molecule* reacB, const bool keep_reacB, const Molecule* surf_reac,
                                  const RxnProductsVector& actual_products, GridPosVector& assigned_surf_product_positions,
                                  uint& num_surface_products, bool& surf_pos_reacA_is_used);
};
using std::min;
#include "gen/mcell4_find_surf_product_positions.inl"   // src4/diffuse_react_event.cpp:1993-2288
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

// One call of find_surf_product_positions for a reaction without a wall collision.
//   occupied: (wall, tile) pairs of the occupied tiles (n_occupied of them); walls without a grid get one first
//   entries: the rule's product list, one byte each: bit 0 surface species, (the kept ones are told by keep_a / keep_b and
//            sit where the rule has them: the function itself only looks at the species of an entry)
//   reac_a / reac_b: (is_surf, wall, tile, u, v) of the reactants in RULE order (reac_b: n_reactants == 2), surf_reac_is_b:
//            which of them is the surface reactant the neighbour tiles are taken around
// out per entry: type (GridPosType as int: 0 not initialized, 1 not assigned, 2 REACA_UV, 3 REACB_UV, 4 POS_UV, 5 RANDOM), wall, tile
// returns the function's result (0 or RX_BLOCKED = -2); *words = 32-bit words drawn
EXPORT int ref4_find_surf_product_positions(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                                            const unsigned* occupied, unsigned n_occupied, int unimol, const unsigned char* entries,
                                            unsigned n_entries, const double* reac_a5, int keep_a, const double* reac_b5, int keep_b,
                                            int surf_reac_is_b, unsigned seed, unsigned skip, int* out_type, unsigned* out_wall,
                                            unsigned* out_tile, unsigned* num_surface_products, int* reac_a_used, long long* words) {
  Partition p; fill(p, verts, n_verts, tri, n_walls, nullptr);
  for (unsigned i = 0; i < n_occupied; i++) p.walls[occupied[2 * i]].grid.molecules_per_tile[occupied[2 * i + 1]] = 1000 + i;
  p.all_species.resize(2); p.all_species[0].surf = false; p.all_species[1].surf = true;   // species id = "is surface"
  BNG::RxnRule rxn; rxn.unimol = unimol != 0;
  RxnProductsVector actual;
  for (unsigned e = 0; e < n_entries; e++) { BNG::Cplx c; c.surf = entries[e] & 1; rxn.products.push_back(c); actual.push_back(ProductSpeciesIdWIndices{(species_id_t)(entries[e] & 1)}); }
  auto mol = [](const double* r5, unsigned id) {
    Molecule m; m.id = id; m.surf = r5[0] != 0; m.species_id = m.surf ? 1 : 0;
    m.s.wall_index = (wall_index_t)r5[1]; m.s.grid_tile_index = (tile_index_t)r5[2]; m.s.pos = Vec2(r5[3], r5[4]);
    return m;
  };
  Molecule a = mol(reac_a5, 1), b;
  if (reac_b5) b = mol(reac_b5, 2);
  const Molecule* surf_reac = surf_reac_is_b ? &b : &a;
  World world; rng_init(&world.rng, seed);
  for (unsigned i = 0; i < skip; i++) (void)rng_uint(&world.rng);
  const long long before = rng_uses(&world.rng);
  DiffuseReactEvent ev; ev.world = &world;
  Collision coll;
  GridPosVector assigned;
  uint nsp = 0; bool used = false;
  const int r = ev.find_surf_product_positions(p, coll, &rxn, &a, keep_a != 0, reac_b5 ? &b : nullptr, keep_b != 0, surf_reac, actual, assigned, nsp, used);
  *words = rng_uses(&world.rng) - before;
  *num_surface_products = nsp; *reac_a_used = used ? 1 : 0;
  for (unsigned e = 0; e < n_entries; e++) {
    if (e < assigned.size()) { out_type[e] = (int)assigned[e].type; out_wall[e] = assigned[e].wall_index; out_tile[e] = assigned[e].tile_index; }
    else { out_type[e] = 0; out_wall[e] = out_tile[e] = 0xFFFFFFFFu; }
  }
  return r;
}
