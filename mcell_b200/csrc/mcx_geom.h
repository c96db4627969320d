// mcx_geom.h — host-side geometry preparation (see mcx_geom.cpp)
#pragma once
#include <cstdint>
#include <vector>
#include "mcx_internal.h"

namespace mcxg {
struct GridSpec { double ox, oy, oz, sp_len, sp_rcp, R; int n_sp; bool use_expanded; };
void wall_constants(const double* verts, const uint32_t* tri, uint64_t n_walls, std::vector<DevWall>& out);
void bin_walls(const GridSpec& g, const double* verts, const uint32_t* tri, const std::vector<DevWall>& walls,
               std::vector<uint32_t>& start, std::vector<uint32_t>& list);
}  // namespace mcxg
