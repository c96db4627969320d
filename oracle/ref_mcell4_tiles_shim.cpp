// oracle/ref_mcell4_tiles_shim.cpp — oracle/_ref build only (TEST INFRASTRUCTURE).
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
  molecule_id_t get_molecule_on_tile(tile_index_t) const { return MOLECULE_ID_INVALID; }
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
  struct { Vec2 pos; wall_index_t wall_index; tile_index_t grid_tile_index; } s;
};

class TileNeighborVector : public std::deque<WallTileIndexPair> {};  // src4/diffuse_react_event.h:54-62

}  // namespace MCell
namespace BNG {
const int PATHWAY_INDEX_NO_RXN = -1;
typedef int rxn_class_pathway_index_t;
class RxnContainer;
class Species {
public:
  bool can_interact_with_border() const { return false; }
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

#define EXPORT extern "C" __attribute__((visibility("default")))
using namespace MCell;

namespace {
// one object: the neighbour across a triangle side is the other wall that uses both of its vertex indices (what
// surface_net, src4/geometry.cpp:258-356, finds for a manifold mesh); edge k runs from vertex k to vertex k + 1
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

// GridUtils::find_neighbor_tiles(p, sm, wall, tile, create_grid_flag = false, search_for_reactant = true) as
// react_2D_all_neighbors calls it (src4/diffuse_react_event.cpp:1267): the list front to back as (wall, tile) pairs.
// grid_init: per wall, 0 = the wall has no grid yet (null = every wall has one).  Returns the length of the list.
EXPORT int ref4_find_neighbor_tiles(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                                    const unsigned char* grid_init, unsigned wall, unsigned tile, unsigned* out_pairs, unsigned cap) {
  Partition p; fill(p, verts, n_verts, tri, n_walls, grid_init);
  Molecule sm; sm.s.wall_index = wall; sm.s.grid_tile_index = tile;
  TileNeighborVector nb;
  GridUtils::find_neighbor_tiles(p, &sm, p.walls[wall], tile, false, true, nb);
  unsigned n = 0;
  for (const WallTileIndexPair& t : nb) {
    if (n < cap) { out_pairs[2 * n] = t.wall_index; out_pairs[2 * n + 1] = t.tile_index; }
    n++;
  }
  return (int)n;
}
// the same for every tile of every wall with a grid: CSR (start has total tiles + 1 entries, tiles in wall order)
EXPORT unsigned long long ref4_neighbor_tile_table(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                                                   const unsigned char* grid_init, unsigned* start, unsigned* out_pairs,
                                                   unsigned long long cap) {
  Partition p; fill(p, verts, n_verts, tri, n_walls, grid_init);
  unsigned long long n = 0; unsigned gt = 0;
  for (unsigned w = 0; w < n_walls; w++) {
    if (!p.walls[w].has_initialized_grid()) continue;
    for (unsigned tile = 0; tile < p.walls[w].grid.num_tiles; tile++) {
      Molecule sm; sm.s.wall_index = w; sm.s.grid_tile_index = tile;
      TileNeighborVector nb;
      GridUtils::find_neighbor_tiles(p, &sm, p.walls[w], tile, false, true, nb);
      start[gt++] = (unsigned)n;
      for (const WallTileIndexPair& t : nb) {
        if (n < cap) { out_pairs[2 * n] = t.wall_index; out_pairs[2 * n + 1] = t.tile_index; }
        n++;
      }
    }
  }
  start[gt] = (unsigned)n;
  return n;
}
EXPORT unsigned ref4_tiles_num_tiles(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls, unsigned* per_wall) {
  Partition p; fill(p, verts, n_verts, tri, n_walls, nullptr);
  unsigned t = 0;
  for (unsigned w = 0; w < n_walls; w++) { per_wall[w] = p.walls[w].grid.num_tiles; t += per_wall[w]; }
  return t;
}

// RxnUtils::test_many_bimolecular with all_neighbors_flag = true (react_2D_all_neighbors, diffuse_react_event.cpp:1362-1366):
// n classes given by their cumulative pathway probabilities (cum_probs, n_pathways per class), scaling per class.
// Returns the index of the chosen class or -1; *pathway = chosen_pathway_index, *words = 32-bit words drawn
EXPORT int ref4_test_many_bimolecular(const double* cum_probs, const int* n_pathways, int n, const double* scaling,
                                      double local_prob_factor, unsigned seed, unsigned skip, int* pathway, long long* words) {
  std::vector<BNG::RxnClass> cls(n);
  BNG::RxnClassesVector v;
  int q = 0;
  for (int i = 0; i < n; i++) { for (int k = 0; k < n_pathways[i]; k++) cls[i].cum_probs.push_back(cum_probs[q++]); v.push_back(&cls[i]); }
  small_vector<double> sc;
  for (int i = 0; i < n; i++) sc.push_back(scaling[i]);
  rng_state rng; rng_init(&rng, seed);
  for (unsigned i = 0; i < skip; i++) (void)rng_uint(&rng);
  const long long before = rng_uses(&rng);
  Partition p;
  BNG::rxn_class_pathway_index_t chosen = -7;
  const int r = RxnUtils::test_many_bimolecular(p, v, sc, local_prob_factor, rng, true, 0.0, chosen);
  *pathway = chosen; *words = rng_uses(&rng) - before;
  return r;
}
// RxnUtils::test_bimolecular with a local probability factor (the single-class case of react_2D_all_neighbors, :1351-1356)
EXPORT int ref4_test_bimolecular_lpf(const double* cum_probs, int n_pathways, double scaling, double local_prob_factor,
                                     unsigned seed, unsigned skip, long long* words) {
  BNG::RxnClass c;
  for (int k = 0; k < n_pathways; k++) c.cum_probs.push_back(cum_probs[k]);
  rng_state rng; rng_init(&rng, seed);
  for (unsigned i = 0; i < skip; i++) (void)rng_uint(&rng);
  const long long before = rng_uses(&rng);
  Partition p; Molecule a, b;
  const int r = RxnUtils::test_bimolecular(p, &c, rng, a, b, scaling, local_prob_factor, 0.0);
  *words = rng_uses(&rng) - before;
  return r;
}
