// mcx_geom.h — host-side geometry preparation (see mcx_geom.cpp)
#pragma once
#include <cstdint>
#include <vector>
#include "mcx_internal.h"

namespace mcxg {
struct GridSpec { double ox, oy, oz, sp_len, sp_rcp, R; int n_sp; bool use_expanded; };
void wall_constants(const double* verts, const uint32_t* tri, uint64_t n_walls, std::vector<DevWall>& out);
// surface grids (Grid::initialize, src4/wall.cpp:38-74); tile_start = exclusive prefix of num_tiles; returns the total
uint64_t grid_constants(const double* verts, const uint32_t* tri, const std::vector<DevWall>& walls, std::vector<DevGrid>& out);
// Wall::area (src4/wall.cpp:304) of every wall, the value Grid::initialize sizes the tile grid with
void wall_areas(const double* verts, const uint32_t* tri, uint64_t n_walls, std::vector<double>& out);
// triangle sides shared by two walls of one object (surface_net, src4/geometry.cpp:258-356) and the transform across
// them (Edge::reinit_edge_constants, src4/wall.cpp:134-235); wall_object may be null (one object)
void edge_constants(const double* verts, const uint32_t* tri, const std::vector<DevWall>& walls, const uint32_t* wall_object,
                    std::vector<DevEdge>& out);
// host helpers behind mcx_grid_num_tiles / mcx_grid2uv / mcx_xyz2grid (one triangle given by its 9 coordinates)
uint32_t tri_num_tiles(const double* v9);
void tri_grid2uv(const double* v9, uint32_t tile, double* uv2);
uint32_t tri_xyz2grid(const double* v9, const double* xyz3);
void bin_walls(const GridSpec& g, const double* verts, const uint32_t* tri, const std::vector<DevWall>& walls,
               std::vector<uint32_t>& start, std::vector<uint32_t>& list);
// fine wall grid (mcx_geom.cpp): subdivision factor K of a subpartition edge and the per-cell wall lists
int fine_wall_factor(const GridSpec& g, const double* verts, const uint32_t* tri, uint64_t n_walls, const std::vector<uint32_t>& start);
void bin_walls_fine(const GridSpec& g, const double* verts, const uint32_t* tri, const std::vector<uint32_t>& start,
                    const std::vector<uint32_t>& list, int K, double margin, std::vector<uint32_t>& fstart, std::vector<uint32_t>& flist);
// neighbour tiles of every tile (GridUtils::find_neighbor_tiles, src4/grid_utils.inl:296-1801, with every grid present):
// start has total tiles + 1 entries, list holds (wall, tile) pairs in the order react_2D_all_neighbors walks them
void tile_neighbor_table(const double* verts, uint64_t n_verts, const uint32_t* tri, const std::vector<DevWall>& walls,
                         const std::vector<DevGrid>& grids, const std::vector<DevEdge>& edges, std::vector<uint32_t>& start,
                         std::vector<uint32_t>& list);
}  // namespace mcxg
