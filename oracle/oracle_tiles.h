// oracle/oracle_tiles.h — TEST INFRASTRUCTURE (part of the CPU oracle, included by oracle.cpp inside namespace orc).
//
// Restatement of MCell4's neighbour-tile search of the surface grids, GridUtils::find_neighbor_tiles and everything it
// uses (src4/grid_utils.inl:296-1801), as react_2D_all_neighbors calls it (create_grid_flag = false,
// search_for_reactant = true; src4/diffuse_react_event.cpp:1267).  The reference pushes to the FRONT of a deque; so
// does this.  Pinned entry for entry against the reference's own compiled code (oracle/_ref/libmcell4tiles.so,
// tests/test_oracle_vs_reference_mcell4_tiles.py).  Region-border restrictions of the search (species that can interact
// with a border: find_restricted_regions_by_wall / check_if_can_move_through_border) are not restated: tables that
// combine surface-surface classes with region borders are refused.

typedef std::pair<uint32_t, uint32_t> WallTile;  // WallTileIndexPair, src4/defines.h:247-270
typedef std::deque<WallTile> TileNeighbors;      // TileNeighborVector, src4/diffuse_react_event.h:54

struct TileCoords { int root, rootrem, strip, stripe, flip; };
static inline TileCoords tile_coords(const Grid& g, uint32_t index) {  // the (strip, stripe, flip) arithmetic used throughout
  TileCoords c;
  c.root = (int)(sqrt((double)index));
  c.rootrem = (int)index - c.root * c.root;
  c.strip = g.n_axis - c.root - 1;
  c.stripe = c.rootrem / 2;
  c.flip = c.rootrem - 2 * c.stripe;
  return c;
}
// grid_utils.inl:296-319
static bool is_inner_tile(const Grid& g, uint32_t index) {
  const TileCoords c = tile_coords(g, index);
  if (c.strip == 0 || c.stripe == 0) return false;
  if (c.strip + c.stripe == g.n_axis - 1) return false;
  if (c.strip + c.stripe == g.n_axis - 2 && c.flip == 1) return false;
  return true;
}
// grid_utils.inl:329-342
static bool is_corner_tile(const Grid& g, uint32_t index) {
  if (index == 0 || index == g.n_tiles - 1) return true;
  const uint32_t tile_idx_mid = g.n_tiles - 2 * (uint32_t)g.n_axis + 1;
  return index == tile_idx_mid;
}
// grid_utils.inl:523-536: -1 when the tile above is on another wall
static int move_strip_up(const Grid& g, uint32_t index) {
  const int root = (int)(sqrt((double)index)) + 1;
  if (g.n_axis == root) return -1;
  return (int)index + 2 * root;
}
// grid_utils.inl:551-590
static int move_strip_down(const Grid& g, uint32_t index) {
  const TileCoords c = tile_coords(g, index);
  const int num_tiles_per_strip = 2 * g.n_axis - 2 * c.strip - 1;
  if (is_inner_tile(g, index)) return (int)index - num_tiles_per_strip + 1;
  if (c.strip == 0 && c.stripe > 0) {
    if (index == g.n_tiles - 1) return -1;
    return (int)index - num_tiles_per_strip + 1;
  }
  if (c.flip == 0) return -1;  // left or right border layers
  return (int)index - num_tiles_per_strip + 1;
}
// tile_orientation, grid_utils.inl:483-510
static int tile_orientation(double u, double v, const Wall& f, const Grid& g) {
  const double striploc = v * g.strip_width_rcp;
  int strip = (int)striploc;
  const double striprem = striploc - strip;
  strip = g.n_axis - strip - 1;
  const double u0 = v * g.vert2_slope;
  const double u1_u0 = f.uv_vert1_u - v * g.fullslope;
  const double stripeloc = ((u - u0) / u1_u0) * (((double)strip) + (1 - striprem));
  const int stripe = (int)(stripeloc);
  const double striperem = stripeloc - stripe;
  return (striperem < 1 - striprem) ? 0 : 1;
}
// walls that use a vertex index, ascending (Partition::walls_using_vertex_mapping, filled in Partition::add_wall order)
static const std::vector<uint32_t>& walls_using_vertex(const World& w, uint32_t v) {
  if (w.vertex_walls.size() != w.verts.size()) {
    w.vertex_walls.assign(w.verts.size(), {});
    for (uint32_t wi = 0; wi < w.walls.size(); wi++)
      for (int k = 0; k < 3; k++) w.vertex_walls[w.walls[wi].vi[k]].push_back(wi);
  }
  return w.vertex_walls[v];
}
static inline bool wall_has_grid(const World& w, uint32_t wi) { return w.assume_all_grids || !w.tiles[wi].empty(); }  // Wall::has_initialized_grid (create_grid_flag: every wall gets one)

// neighboring_wall_uses_this_vertex, grid_utils.inl:605-622
static bool neighboring_wall_uses_this_vertex(const World& w, const Wall& f, uint32_t vi) {
  for (int i = 0; i < 3; i++)
    if (f.nb_wall[i] != MCX_NONE) {
      const Wall& nw = w.walls[f.nb_wall[i]];
      for (int s = 0; s < 3; s++) if (vi == nw.vi[s]) return true;
    }
  return false;
}
// WallUtils::walls_share_full_edge, wall_utils.inl:50-65
static bool walls_share_full_edge(const World& w, const Wall& a, const Wall& b) {
  int count = 0;
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++)
      if (!distinguishable_vec3(w.verts[a.vi[i]], w.verts[b.vi[k]], POS_EPS)) count++;
  return count == 2;
}
// grid_all_neighbors_across_walls_through_vertices, grid_utils.inl:743-868 (without restricted regions)
static void neighbors_through_vertices(const World& w, const std::vector<uint32_t>& neighboring_walls, uint32_t wall, TileNeighbors& nb) {
  const Wall& f = w.walls[wall];
  const Grid& grid = w.grids[wall];
  for (uint32_t wi : neighboring_walls) {
    if (!wall_has_grid(w, wi)) continue;
    const Wall& nf = w.walls[wi];
    const Grid& ng = w.grids[wi];
    uint32_t nbr_tile;
    if (grid.n_tiles == 1) nbr_tile = 0;
    else {
      uint32_t nbr_vertex = MCX_NONE;
      for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++)
          if (f.vi[i] == nf.vi[k]) { nbr_vertex = nf.vi[k]; break; }  // sic: only the inner loop ends, the last shared vertex wins
      nbr_tile = MCX_NONE;
      if (nbr_vertex == nf.vi[0]) nbr_tile = ng.n_tiles - 2 * (uint32_t)ng.n_axis + 1;
      else if (nbr_vertex == nf.vi[1]) nbr_tile = ng.n_tiles - 1;
      else if (nbr_vertex == nf.vi[2]) nbr_tile = 0;
    }
    nb.push_front(WallTile(wi, nbr_tile));
  }
}
// bisect / bisect_high, grid_utils.inl:871-915
static int tiles_bisect(const std::vector<double>& list, int n, double val) {
  int lo = 0, hi = n, mid = 0;
  while (hi - lo > 1) { mid = (hi + lo) / 2; if (list[mid] > val) hi = mid; else lo = mid; }
  return lo;
}
static int tiles_bisect_high(const std::vector<double>& list, int n, double val) {
  int lo = 0, hi = n - 1, mid = 0;
  while (hi - lo > 1) { mid = (hi + lo) / 2; if (list[mid] > val) hi = mid; else lo = mid; }
  return list[lo] > val ? lo : hi;
}
static inline void push_front_if_not_present(TileNeighbors& nb, WallTile t) {  // grid_utils.inl:917-926
  if (std::find(nb.begin(), nb.end(), t) == nb.end()) nb.push_front(t);
}
// add_more_tile_neighbors_to_list_fast, grid_utils.inl:946-1240: the tiles of new_wall that touch the start tile along the
// shared edge start -> end (edge_index as seen from the start wall)
static void add_more_tile_neighbors(const World& w, uint32_t orig_wall, int orig_strip, int orig_stripe, int orig_flip, V3 start, V3 end,
                                    int edge_index, uint32_t new_wall, TileNeighbors& nb) {
  const Grid& og = w.grids[orig_wall];
  const Grid& ng = w.grids[new_wall];
  const Wall& of = w.walls[orig_wall];
  const Wall& nf = w.walls[new_wall];
  const int N = og.n_axis;
  double orig_pos_1 = -1, orig_pos_2 = -1;
  const int new_pos_size = ng.n_axis + 1;
  std::vector<double> new_pos(new_pos_size);
  std::vector<std::array<int, 3>> idx(new_pos_size);
  const V3 dse = start - end;
  const double edge_length = sqrt(len3_squared(dse));  // distance3
  auto two = [&](int k) { orig_pos_1 = k * edge_length / N; orig_pos_2 = (k + 1) * edge_length / N; };
  auto one = [&](int k) { orig_pos_1 = k * edge_length / N; };
  if (orig_stripe == 0) {
    if (orig_strip > 0) {
      if (orig_flip == 0) two(orig_strip); else one(orig_strip + 1);
    } else {
      if (edge_index == 0) { if (orig_flip == 0) two(orig_stripe); else one(orig_stripe + 1); }
      else if (edge_index == 1) { if (orig_flip == 0) one(orig_strip); else two(orig_strip); }
      else if (edge_index == 2) { if (orig_flip == 0) two(orig_strip); else one(orig_strip + 1); }
    }
  }
  bool check_side_flag = false;
  if (orig_strip == 0 && orig_stripe > 0) {
    if (orig_stripe == N - 1) check_side_flag = true;
    if (orig_stripe == N - 2 && orig_flip == 1) check_side_flag = true;
    if (!check_side_flag) {
      if (orig_flip == 0) two(orig_stripe); else one(orig_stripe + 1);
    } else {
      if (edge_index == 0) { if (orig_flip == 0) two(orig_stripe); else one(orig_stripe + 1); }
      else if (edge_index == 1) { if (orig_flip == 0) two(orig_strip); else one(orig_strip + 1); }
    }
  }
  if (orig_strip > 0 && orig_stripe > 0) {
    if (orig_flip == 0) two(orig_strip); else one(orig_strip + 1);
  }
  // find_shared_vertices_for_neighbor_walls, grid_utils.inl:682-728
  int shared_vert_1 = -1, shared_vert_2 = -1;
  for (int k = 0; k < 3; k++) {
    const V3 nv = w.verts[nf.vi[k]];
    if (!distinguishable_vec3(nv, w.verts[of.vi[0]], POS_EPS) || !distinguishable_vec3(nv, w.verts[of.vi[1]], POS_EPS) ||
        !distinguishable_vec3(nv, w.verts[of.vi[2]], POS_EPS)) {
      if (k == 0 || shared_vert_1 < 0) shared_vert_1 = k; else shared_vert_2 = k;
    }
  }
  int new_start_index, new_end_index;
  if (!distinguishable_vec3(start, w.verts[nf.vi[shared_vert_1]], POS_EPS)) { new_start_index = shared_vert_1; new_end_index = shared_vert_2; }
  else { new_start_index = shared_vert_2; new_end_index = shared_vert_1; }
  if (new_start_index > new_end_index) {  // invert_orig_pos
    orig_pos_1 = edge_length - orig_pos_1;
    if (orig_pos_2 > 0) orig_pos_2 = edge_length - orig_pos_2;
  }
  for (int i = 0; i < new_pos_size; i++) new_pos[i] = i * edge_length / ng.n_axis;
  int new_edge_index = 0;
  if (shared_vert_1 + shared_vert_2 == 1) new_edge_index = 0;
  else if (shared_vert_1 + shared_vert_2 == 2) new_edge_index = 2;
  else if (shared_vert_1 + shared_vert_2 == 3) new_edge_index = 1;
  // tile indices of the border layer next to the shared edge
  int last_value;
  const int mid_tile = (int)ng.n_tiles - 2 * ng.n_axis + 1;
  idx[0] = {-1, -1, new_edge_index == 1 ? (int)ng.n_tiles - 1 : mid_tile};
  last_value = idx[0][2];
  for (int i = 1; i < new_pos_size - 1; i++) {
    if (new_edge_index == 0) {
      for (int k = 0; k < 3; k++) idx[i][k] = last_value + k;
      last_value = idx[i][2];
    } else {
      for (int k = 0; k < 2; k++) idx[i][k] = new_edge_index == 1 ? last_value - k : last_value + k;
      last_value = idx[i][1];
      idx[i][2] = move_strip_down(ng, (uint32_t)last_value);
      last_value = idx[i][2];
    }
  }
  idx[new_pos_size - 1] = {last_value, -1, -1};
  int ind_high, ind_low = -1;
  if (orig_pos_1 > orig_pos_2) {
    ind_high = tiles_bisect_high(new_pos, new_pos_size, orig_pos_1);
    if (orig_pos_2 > 0) ind_low = tiles_bisect(new_pos, new_pos_size, orig_pos_2);
  } else {
    ind_high = tiles_bisect_high(new_pos, new_pos_size, orig_pos_2);
    if (orig_pos_1 > 0) ind_low = tiles_bisect(new_pos, new_pos_size, orig_pos_1);
  }
  if (ind_low >= 0) {
    for (int i = ind_low + 1; i < ind_high; i++)
      for (int k = 0; k < 3; k++) push_front_if_not_present(nb, WallTile(new_wall, (uint32_t)idx[i][k]));
  } else push_front_if_not_present(nb, WallTile(new_wall, (uint32_t)idx[ind_high][0]));
}
// grid_all_neighbors_across_walls_through_edges, grid_utils.inl:1285-1640 (search for a reactant, no restricted regions)
static void neighbors_through_edges(const World& w, uint32_t wall, uint32_t tile, TileNeighbors& nb) {
  const Wall& f = w.walls[wall];
  const Grid& g = w.grids[wall];
  const int N = g.n_axis;
  const TileCoords c = tile_coords(g, tile);
  const int strip = c.strip, stripe = c.stripe, flip = c.flip;
  const V3 v0 = w.verts[f.vi[0]], v1 = w.verts[f.vi[1]], v2 = w.verts[f.vi[2]];
  auto has = [&](int e) { return f.nb_wall[e] != MCX_NONE && wall_has_grid(w, f.nb_wall[e]); };
  auto own = [&](int t) { nb.push_front(WallTile(wall, (uint32_t)t)); };
  auto more = [&](V3 s, V3 e, int edge_index, int nbw) { add_more_tile_neighbors(w, wall, strip, stripe, flip, s, e, edge_index, f.nb_wall[nbw], nb); };
  const int ti = (int)tile;
  int temp;
  if (stripe == 0) {
    if (flip > 0) {  // inverted tile
      own(ti - 1); own(ti + 1);
      if (strip < N - 2) own(ti + 2);
      temp = move_strip_down(g, tile);
      own(temp);
      if (strip < N - 2) { own(temp + 1); own(temp + 2); }
      if (strip > 0) { temp = move_strip_up(g, tile); own(temp); own(temp - 1); own(temp + 1); }
      if (has(2)) more(v0, v2, 2, 2);
      if (strip == 0 && has(0)) more(v0, v1, 0, 0);
      if (strip == N - 2 && has(1)) more(v1, v2, 0, 1);  // sic: edge index 0 (grid_utils.inl:1408)
    } else {  // upright tile
      if (tile == 0) {
        if (g.n_tiles > 1) { temp = move_strip_up(g, tile); own(temp); own(temp - 1); own(temp + 1); }
        else if (has(0)) more(v0, v1, 0, 0);
        if (has(1)) more(v1, v2, 1, 1);
        if (has(2)) more(v0, v2, 2, 2);
      } else {
        own(ti + 1); own(ti + 2);
        temp = move_strip_down(g, tile + 1);
        own(temp);
        if (strip > 0) { temp = move_strip_up(g, tile); own(temp); own(temp - 1); own(temp + 1); own(temp + 2); }
        else {  // the top left corner
          if (has(0)) more(v0, v1, 0, 0);
          if (has(2)) more(v0, v2, 2, 2);
        }
      }
    }
  }
  if (strip == 0 && stripe > 0) {
    own(ti - 1); own(ti - 2);
    if (stripe < N - 2 || (stripe == N - 2 && flip == 0)) { own(ti + 1); own(ti + 2); }
    else if (stripe == N - 2 && flip == 1) own(ti + 1);
    if (flip > 0) {
      temp = move_strip_down(g, tile);
      own(temp); own(temp - 1); own(temp - 2);
      if (stripe < N - 2) { own(temp + 1); own(temp + 2); }
    } else {
      if (tile < g.n_tiles - 1) { temp = move_strip_down(g, tile); own(temp); own(temp - 1); own(temp + 1); }
      else { temp = move_strip_down(g, tile - 1); own(temp); }  // a corner tile
    }
    if (has(0)) more(v0, v1, 0, 0);
    if (tile == g.n_tiles - 1 || tile == g.n_tiles - 2)
      if (has(1)) more(v1, v2, 1, 1);
  }
  if (strip > 0 && stripe > 0) {  // the right border layer
    if (flip > 0) {
      own(ti - 1); own(ti - 2); own(ti + 1);
      temp = move_strip_up(g, tile); own(temp); own(temp - 1); own(temp + 1);
      temp = move_strip_down(g, tile); own(temp); own(temp - 1); own(temp - 2);
    } else {
      own(ti - 1); own(ti - 2);
      temp = move_strip_up(g, tile); own(temp); own(temp - 1); own(temp - 2); own(temp + 1);
      temp = move_strip_down(g, tile - 1); own(temp);
    }
    if (has(1)) more(v1, v2, 1, 1);
  }
}
// grid_all_neighbors_for_inner_tile, grid_utils.inl:1657-1737 (grid_neighbors :414-470 for the tile above / below)
static void neighbors_for_inner_tile(const World& w, uint32_t wall, uint32_t tile, TileNeighbors& nb) {
  const Wall& f = w.walls[wall];
  const Grid& g = w.grids[wall];
  const int ti = (int)tile;
  // grid_neighbors: k = strip counted from vertex 2, j = stripe, i = flip
  const int root = (int)(sqrt((double)tile)), rootrem = ti - root * root;
  const int k = root, j = rootrem / 2, i = rootrem - 2 * j;
  int si[3];
  si[2] = ti - 1; si[1] = ti + 1;
  si[0] = i ? 2 * j + (k - 1) * (k - 1) : 1 + 2 * j + (k + 1) * (k + 1);
  int vert_nbr = -1;
  for (int kk = 0; kk < 3; kk++)
    if (si[kk] != ti - 1 && si[kk] != ti + 1) { vert_nbr = si[kk]; break; }
  auto own = [&](int t) { nb.push_front(WallTile(wall, (uint32_t)t)); };
  own(ti - 1); own(ti - 2); own(ti + 1); own(ti + 2);
  double u, v;
  grid2uv(f, g, tile, u, v);
  int temp;
  if (tile_orientation(u, v, f, g) == 0) {
    own(vert_nbr); own(vert_nbr - 1); own(vert_nbr - 2); own(vert_nbr + 1); own(vert_nbr + 2);
    temp = move_strip_down(g, tile); own(temp); own(temp - 1); own(temp + 1);
  } else {
    temp = move_strip_up(g, tile); own(temp); own(temp - 1); own(temp + 1);
    own(vert_nbr); own(vert_nbr - 1); own(vert_nbr - 2); own(vert_nbr + 1); own(vert_nbr + 2);
  }
}
// find_neighbor_tiles, grid_utils.inl:1754-1801
static void find_neighbor_tiles(const World& w, uint32_t wall, uint32_t tile, TileNeighbors& nb) {
  const Wall& f = w.walls[wall];
  const Grid& g = w.grids[wall];
  if (is_inner_tile(g, tile)) { neighbors_for_inner_tile(w, wall, tile, nb); return; }
  if (is_corner_tile(g, tile)) {
    // find_shared_vertices_corner_tile_parent_wall (:624-670): wall vertices under the tile that a neighbouring wall uses too
    uint32_t shared_verts[3] = {MCX_NONE, MCX_NONE, MCX_NONE};
    if (tile == g.n_tiles - 2 * (uint32_t)g.n_axis + 1 && neighboring_wall_uses_this_vertex(w, f, f.vi[0])) shared_verts[0] = f.vi[0];
    if (tile == g.n_tiles - 1 && neighboring_wall_uses_this_vertex(w, f, f.vi[1])) shared_verts[1] = f.vi[1];
    if (tile == 0 && neighboring_wall_uses_this_vertex(w, f, f.vi[2])) shared_verts[2] = f.vi[2];
    // WallUtils::find_nbr_walls_shared_one_vertex (wall_utils.inl:79-104): walls that touch in that vertex only
    std::vector<uint32_t> neighboring_walls;
    for (int i = 0; i < 3; i++)
      if (shared_verts[i] != MCX_NONE)
        for (uint32_t wi : walls_using_vertex(w, shared_verts[i])) {
          if (wi == wall) continue;
          if (!walls_share_full_edge(w, f, w.walls[wi])) neighboring_walls.push_back(wi);
        }
    if (!neighboring_walls.empty()) neighbors_through_vertices(w, neighboring_walls, wall, nb);
  }
  neighbors_through_edges(w, wall, tile, nb);
}
