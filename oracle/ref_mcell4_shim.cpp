// oracle/ref_mcell4_shim.cpp — oracle/_ref build only (TEST INFRASTRUCTURE).
//
// Builds the reference's OWN subpartition walk — CollisionUtils::collect_crossed_subparts and
// collect_neighboring_subparts, src4/collision_utils_subparts.inl:38-300 — unmodified, from the source where it lies
// under /root/reference, into oracle/_ref/libmcell4ref.so (SURVEY 8c "tier-2 oracle").  MCell4 as a whole cannot be
// built here (libbng, nfsim, boost, VTK absent), so the headers that drag the engine in are switched off through
// their include guards and replaced by the few declarations the walk uses:
//   * bng/shared_defines.h            -> oracle/ref_shims4/bng/shared_defines.h (own stand-in; src4/defines.h itself,
//                                        with glm from the reference's libs/, is the reference's)
//   * Partition / SimulationConfig /  -> the minimal classes below: the index arithmetic of src4/partition.h:253-296
//     Molecule                           restated (it is what oracle.cpp restates too, and SURVEY a27 lists it)
// Nothing from the reference is copied into this repository: the .inl is #included at build time.
#include "bng/shared_defines.h"
#include "defines.h"

#define SRC4_DIFFUSE_REACT_EVENT_H_
#define SRC4_WORLD_H_
#define SRC4_PARTITION_H_
#define SRC4_GEOMETRY_H_
#define SRC4_GEOMETRY_UTILS_INC_

namespace MCell {

struct SimulationConfig {
  pos_t partition_edge_length;
  uint num_subparts_per_partition_edge, num_subparts_per_partition_edge_squared;
  pos_t subpart_edge_length, subpart_edge_length_rcp;
  bool use_expanded_list;
};

struct Molecule {
  struct { Vec3 pos; subpart_index_t subpart_index; } v;
};

class Partition {
public:
  SimulationConfig config;
  Vec3 origin_corner;
  const Vec3& get_origin_corner() const { return origin_corner; }
  bool is_subpart_index_in_range(const int index) const { return index >= 0 && index < (int)config.num_subparts_per_partition_edge; }
  void get_subpart_3d_indices(const Vec3& pos, IVec3& res) const {  // truncating conversion of (pos - origin) * rcp
    res.x = (int)((pos.x - origin_corner.x) * config.subpart_edge_length_rcp);
    res.y = (int)((pos.y - origin_corner.y) * config.subpart_edge_length_rcp);
    res.z = (int)((pos.z - origin_corner.z) * config.subpart_edge_length_rcp);
  }
  subpart_index_t get_subpart_index_from_3d_indices_allow_outside(const IVec3& i) const {
    return i.x + i.y * config.num_subparts_per_partition_edge + i.z * config.num_subparts_per_partition_edge_squared;
  }
  subpart_index_t get_subpart_index_from_3d_indices(const IVec3& i) const { return get_subpart_index_from_3d_indices_allow_outside(i); }
  subpart_index_t get_subpart_index_from_3d_indices(const int x, const int y, const int z) const {
    return get_subpart_index_from_3d_indices(IVec3(x, y, z));
  }
  void get_subpart_3d_indices_from_index(const subpart_index_t index, IVec3& i) const {
    const uint32_t dim = config.num_subparts_per_partition_edge;
    i.x = index % dim; i.y = (index / dim) % dim; i.z = (index / config.num_subparts_per_partition_edge_squared) % dim;
  }
};

}  // namespace MCell

#include "collision_utils_subparts.inl"

#define EXPORT extern "C" __attribute__((visibility("default")))

// One call of the reference's collect_crossed_subparts for a molecule at pos3 (in its subpartition) moving by disp3.
// out_walls: the ordered vector (ray_trace_vol walks it for wall hits); out_mols: the set, ascending.
// Returns the destination subpartition index.
EXPORT unsigned ref4_collect_crossed_subparts(const double* origin3, double partition_edge_length, unsigned n_subparts_per_edge,
                                              int use_expanded_list, double rxn_radius, const double* pos3, const double* disp3,
                                              int collect_for_molecules, int collect_for_walls, unsigned* out_walls,
                                              unsigned* n_walls, unsigned* out_mols, unsigned* n_mols, unsigned cap) {
  using namespace MCell;
  Partition p;
  p.config.partition_edge_length = partition_edge_length;
  p.config.num_subparts_per_partition_edge = n_subparts_per_edge;
  p.config.num_subparts_per_partition_edge_squared = n_subparts_per_edge * n_subparts_per_edge;
  p.config.subpart_edge_length = partition_edge_length / n_subparts_per_edge;   // simulation_config.cpp:48
  p.config.subpart_edge_length_rcp = 1.0 / p.config.subpart_edge_length;         // :63
  p.config.use_expanded_list = use_expanded_list != 0;
  p.origin_corner = Vec3(origin3[0], origin3[1], origin3[2]);
  Molecule vm;
  vm.v.pos = Vec3(pos3[0], pos3[1], pos3[2]);
  IVec3 si;
  p.get_subpart_3d_indices(vm.v.pos, si);
  vm.v.subpart_index = p.get_subpart_index_from_3d_indices(si);
  SubpartIndicesVector walls;
  SubpartIndicesSet mols;
  const subpart_index_t dest = CollisionUtils::collect_crossed_subparts(
      p, vm, Vec3(disp3[0], disp3[1], disp3[2]), rxn_radius, p.config.subpart_edge_length, collect_for_molecules != 0,
      collect_for_walls != 0, walls, mols);
  unsigned k = 0;
  for (subpart_index_t s : walls) { if (k < cap) out_walls[k] = s; k++; }
  *n_walls = k;
  k = 0;
  for (subpart_index_t s : mols) { if (k < cap) out_mols[k] = s; k++; }
  *n_mols = k;
  return dest;
}
