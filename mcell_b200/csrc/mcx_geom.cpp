// mcx_geom.cpp — host-side, init-time geometry preparation of libmcx:
//   * per-triangle constants           (replaces Wall::initialize_wall_constants, src4/wall.cpp:281-342)
//   * wall -> subpartition binning      (replaces Partition::finalize_walls, src4/partition.cpp:91-118,
//                                        GeometryUtils::wall_subparts_collision_test, geometry_utils.inl:110-207,
//                                        WallUtils::wall_in_box, wall_utils.inl:326-504)
// The binning decides which walls a molecule can ever see, so it keeps the reference's exact tests and
// floating-point expression order; the output is a CSR (spw_start / spw_list) with ascending wall indices
// per subpartition (the iteration order of the reference's uint_set<wall_index_t>).
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <map>
#include <vector>
#include "mcx_geom.h"

namespace mcxg {

struct P3 { double x, y, z; };
static inline P3 sub(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline P3 scale(P3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline double dotp(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline P3 crossp(P3 x, P3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
static inline double len2(P3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

static inline bool distinguishable(double a, double b, double eps) {
  double c = std::fabs(a - b);
  a = std::fabs(a);
  if (a < 1) a = 1;
  b = std::fabs(b);
  if (b < a) eps *= a; else eps *= b;
  return c > eps;
}

void wall_constants(const double* verts, const uint32_t* tri, uint64_t n_walls, std::vector<DevWall>& out) {
  out.resize(n_walls);
  for (uint64_t i = 0; i < n_walls; i++) {
    const double* a = verts + 3 * tri[3 * i];
    const double* b = verts + 3 * tri[3 * i + 1];
    const double* c = verts + 3 * tri[3 * i + 2];
    P3 v0 = {a[0], a[1], a[2]}, v1 = {b[0], b[1], b[2]}, v2 = {c[0], c[1], c[2]};
    DevWall w{};
    w.v0x = v0.x; w.v0y = v0.y; w.v0z = v0.z;
    P3 e1 = sub(v1, v0), e2 = sub(v2, v0);
    double area = 0.5 * std::sqrt(len2(crossp(e1, e2)));
    if (distinguishable(area, 0, 1e-12)) {
      double inv_len = 1 / std::sqrt(len2(e1));
      P3 uu = scale(e1, inv_len);
      P3 nn = crossp(uu, e2);
      double inv_n = 1 / std::sqrt(len2(nn));
      nn = scale(nn, inv_n);
      P3 vv = crossp(nn, uu);
      w.nx = nn.x; w.ny = nn.y; w.nz = nn.z;
      w.dist = dotp(v0, nn);
      w.ux = uu.x; w.uy = uu.y; w.uz = uu.z;
      w.vx = vv.x; w.vy = vv.y; w.vz = vv.z;
      w.uv1u = dotp(e1, uu);
      w.uv2u = dotp(e2, uu);
      w.uv2v = dotp(e2, vv);
    }
    out[i] = w;
  }
}

namespace {
struct Box { P3 lo, hi; };

inline bool inside(const Box& b, P3 p) {
  return p.x >= b.lo.x && p.x <= b.hi.x && p.y >= b.lo.y && p.y <= b.hi.y && p.z >= b.lo.z && p.z <= b.hi.z;
}
inline double comp(P3 p, int k) { return k == 0 ? p.x : (k == 1 ? p.y : p.z); }

// does the segment p->q pierce the box face {axis = plane} within the face rectangle?
inline bool edge_hits_face(P3 p, P3 q, const Box& b, int axis, double plane, int o1, int o2) {
  double pa = comp(p, axis), qa = comp(q, axis);
  if (!((pa <= plane && plane < qa) || (pa > plane && plane >= qa))) return false;
  double r = (plane - pa) / (qa - pa);
  double c1 = comp(p, o1) + r * (comp(q, o1) - comp(p, o1));
  double c2 = comp(p, o2) + r * (comp(q, o2) - comp(p, o2));
  return comp(b.lo, o1) <= c1 && c1 <= comp(b.hi, o1) && comp(b.lo, o2) <= c2 && c2 <= comp(b.hi, o2);
}

// triangle vs axis-aligned box, same decision sequence as wall_in_box
bool triangle_touches_box(const P3 tv[3], const DevWall& w, const Box& b) {
  for (int k = 0; k < 3; k++) if (inside(b, tv[k])) return true;
  for (int k = 0; k < 3; k++) {
    P3 q = tv[k], p = tv[k == 0 ? 2 : k - 1];
    if (edge_hits_face(p, q, b, 0, b.lo.x, 1, 2) || edge_hits_face(p, q, b, 0, b.hi.x, 1, 2)) return true;
    if (edge_hits_face(p, q, b, 1, b.lo.y, 0, 2) || edge_hits_face(p, q, b, 1, b.hi.y, 0, 2)) return true;
    if (edge_hits_face(p, q, b, 2, b.lo.z, 1, 0) || edge_hits_face(p, q, b, 2, b.hi.z, 1, 0)) return true;
  }
  // box edges against the triangle: the 12 edges in the reference's visiting order (corner bit k of
  // a/b selects hi (1) or lo (0) for x,y,z)
  static const unsigned char A[12][3] = {{0,0,0},{0,0,1},{0,1,1},{0,1,0},{1,1,0},{1,1,1},{1,0,1},{1,0,0},{0,0,0},{0,0,1},{0,1,1},{1,0,0}};
  static const unsigned char Bc[12][3] = {{0,0,1},{0,1,1},{0,1,0},{1,1,0},{1,1,1},{1,0,1},{1,0,0},{0,0,0},{0,1,0},{1,0,1},{1,1,1},{0,1,0}};
  P3 n = {w.nx, w.ny, w.nz};
  double d = w.dist;
  P3 u = sub(tv[1], tv[0]);
  double r_u = 1 / std::sqrt(len2(u));
  u = scale(u, r_u);
  P3 v = crossp(n, u);
  double tu[3], tw[3];
  for (int j = 0; j < 3; j++) { tu[j] = dotp(tv[j], u); tw[j] = dotp(tv[j], v); }
  auto corner = [&](const unsigned char s[3]) { return P3{s[0] ? b.hi.x : b.lo.x, s[1] ? b.hi.y : b.lo.y, s[2] ? b.hi.z : b.lo.z}; };
  for (int e = 0; e < 12; e++) {
    P3 ba = corner(A[e]), bb = corner(Bc[e]);
    double a1 = dotp(ba, n), a2 = dotp(bb, n);
    if ((a1 - d < 0 && a2 - d < 0) || (a1 - d > 0 && a2 - d > 0)) continue;
    double r = (d - a1) / (a2 - a1);
    P3 c = {ba.x + r * (bb.x - ba.x), ba.y + r * (bb.y - ba.y), ba.z + r * (bb.z - ba.z)};
    double cu = dotp(c, u), cv = dotp(c, v);
    int crossings = 0;
    for (int j = 0; j < 3; j++) {
      int k = j == 0 ? 2 : j - 1;
      if ((tu[k] < cu && cu <= tu[j]) || (tu[k] >= cu && cu > tu[j])) {
        double rr = (cu - tu[k]) / (tu[j] - tu[k]);
        if ((tw[k] + rr * (tw[j] - tw[k])) > cv) crossings++;
      }
    }
    if (crossings & 1) return true;
  }
  return false;
}
}  // namespace

void bin_walls(const GridSpec& g, const double* verts, const uint32_t* tri, const std::vector<DevWall>& walls,
               std::vector<uint32_t>& start, std::vector<uint32_t>& list) {
  const size_t ns3 = (size_t)g.n_sp * g.n_sp * g.n_sp;
  std::vector<std::vector<uint32_t>> per(ns3);
  auto idx = [&](double v, double o) { return (int)((v - o) * g.sp_rcp); };
  for (uint32_t wi = 0; wi < walls.size(); wi++) {
    P3 tv[3];
    for (int k = 0; k < 3; k++) { const double* q = verts + 3 * tri[3 * wi + k]; tv[k] = {q[0], q[1], q[2]}; }
    P3 lo = tv[0], hi = tv[0];
    for (int k = 1; k < 3; k++) {
      if (tv[k].x < lo.x) lo.x = tv[k].x; else if (tv[k].x > hi.x) hi.x = tv[k].x;
      if (tv[k].y < lo.y) lo.y = tv[k].y; else if (tv[k].y > hi.y) hi.y = tv[k].y;
      if (tv[k].z < lo.z) lo.z = tv[k].z; else if (tv[k].z > hi.z) hi.z = tv[k].z;
    }
    double leeway = 1;
    if (lo.x < -leeway) leeway = -lo.x;
    if (lo.y < -leeway) leeway = -lo.y;
    if (lo.z < -leeway) leeway = -lo.z;
    if (hi.x > leeway) leeway = hi.x;
    if (hi.y > leeway) leeway = hi.y;
    if (hi.z > leeway) leeway = hi.z;
    leeway = 1e-12 + leeway * 1e-12;
    if (g.use_expanded) leeway += g.R;
    lo = {lo.x - leeway, lo.y - leeway, lo.z - leeway};
    hi = {hi.x + leeway, hi.y + leeway, hi.z + leeway};
    int x0 = idx(lo.x, g.ox), y0 = idx(lo.y, g.oy), z0 = idx(lo.z, g.oz);
    int x1 = idx(hi.x, g.ox), y1 = idx(hi.y, g.oy), z1 = idx(hi.z, g.oz);
    for (int x = x0; x <= x1; x++)
      for (int y = y0; y <= y1; y++)
        for (int z = z0; z <= z1; z++) {
          if (x < 0 || y < 0 || z < 0 || x >= g.n_sp || y >= g.n_sp || z >= g.n_sp) continue;
          Box b;
          b.lo = {g.ox + x * g.sp_len, g.oy + y * g.sp_len, g.oz + z * g.sp_len};
          b.hi = {b.lo.x + g.sp_len, b.lo.y + g.sp_len, b.lo.z + g.sp_len};
          b.lo = {b.lo.x - leeway, b.lo.y - leeway, b.lo.z - leeway};
          b.hi = {b.hi.x + leeway, b.hi.y + leeway, b.hi.z + leeway};
          if (triangle_touches_box(tv, walls[wi], b)) per[(size_t)x + (size_t)y * g.n_sp + (size_t)z * g.n_sp * g.n_sp].push_back(wi);
        }
  }
  start.assign(ns3 + 1, 0);
  size_t total = 0;
  for (size_t s = 0; s < ns3; s++) { start[s] = (uint32_t)total; total += per[s].size(); }
  start[ns3] = (uint32_t)total;
  list.clear(); list.reserve(total);
  for (size_t s = 0; s < ns3; s++) list.insert(list.end(), per[s].begin(), per[s].end());
}


// ---- fine wall grid -----------------------------------------------------------------------------------------
// The reference tests a move against every wall of each subpartition it crosses (get_closest_wall_collision,
// collision_utils.inl:819-914); with the default 0.5 um subpartitions a membrane mesh puts hundreds of triangles
// into one list.  libmcx subdivides every subpartition into K^3 cells and lists, per cell, the walls OF THAT
// SUBPARTITION whose bounding box (inflated by `margin`) overlaps the cell, in the subpartition list's own
// (ascending wall index) order.  A wall a move can hit inside the subpartition has the hit point inside its bounding
// box and inside the move's bounding box, so walking the cells under the move's bounding box visits every wall that
// can matter and skipping the others changes nothing: they are COLLIDE_MISS without a random draw.
int fine_wall_factor(const GridSpec& g, const double* verts, const uint32_t* tri, uint64_t n_walls,
                     const std::vector<uint32_t>& start) {
  uint32_t longest = 0;
  for (size_t s = 0; s + 1 < start.size(); s++) longest = std::max(longest, start[s + 1] - start[s]);
  if (longest <= 24 || n_walls == 0) return 1;  // short lists: the subpartition list is the cell list
  double extent = 0;
  for (uint64_t wi = 0; wi < n_walls; wi++) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < 3; k++) {
      const double* q = verts + 3 * tri[3 * wi + k];
      for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], q[a]); hi[a] = std::max(hi[a], q[a]); }
    }
    extent += std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
  }
  extent /= (double)n_walls;
  // cells about 1.5 wall extents wide, between 1/10 of a subpartition and a whole one, and a table of at most
  // 2^25 cell starts
  int K = (int)std::floor(g.sp_len / std::max(1.5 * extent, g.sp_len / 10.0));
  K = std::max(1, std::min(10, K));
  const double ns3 = (double)g.n_sp * g.n_sp * g.n_sp;
  while (K > 1 && ns3 * K * K * K > 33554432.0) K--;
  return K;
}

void bin_walls_fine(const GridSpec& g, const double* verts, const uint32_t* tri, const std::vector<uint32_t>& start,
                    const std::vector<uint32_t>& list, int K, double margin, std::vector<uint32_t>& fstart,
                    std::vector<uint32_t>& flist) {
  const size_t ns3 = (size_t)g.n_sp * g.n_sp * g.n_sp, K3 = (size_t)K * K * K;
  const double rcp = (double)K / g.sp_len;
  fstart.assign(ns3 * K3 + 1, 0);
  auto cell_range = [&](uint32_t wi, size_t s, int lo[3], int hi[3]) {
    const int sx = (int)(s % g.n_sp), sy = (int)((s / g.n_sp) % g.n_sp), sz = (int)(s / ((size_t)g.n_sp * g.n_sp));
    const double o[3] = {g.ox + sx * g.sp_len, g.oy + sy * g.sp_len, g.oz + sz * g.sp_len};
    double blo[3] = {1e300, 1e300, 1e300}, bhi[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < 3; k++) {
      const double* q = verts + 3 * tri[3 * wi + k];
      for (int a = 0; a < 3; a++) { blo[a] = std::min(blo[a], q[a]); bhi[a] = std::max(bhi[a], q[a]); }
    }
    for (int a = 0; a < 3; a++) {
      lo[a] = std::max(0, std::min(K - 1, (int)std::floor((blo[a] - margin - o[a]) * rcp)));
      hi[a] = std::max(0, std::min(K - 1, (int)std::floor((bhi[a] + margin - o[a]) * rcp)));
    }
  };
  for (int pass = 0; pass < 2; pass++) {
    std::vector<uint32_t> cursor;
    if (pass == 1) {  // counts -> exclusive starts
      uint64_t total = 0;
      for (size_t c = 0; c < ns3 * K3; c++) { const uint32_t n = fstart[c]; fstart[c] = (uint32_t)total; total += n; }
      fstart[ns3 * K3] = (uint32_t)total;
      flist.assign(total, 0);
      cursor.assign(fstart.begin(), fstart.end() - 1);
    }
    for (size_t s = 0; s < ns3; s++)
      for (uint32_t k = start[s]; k < start[s + 1]; k++) {  // ascending wall index: every cell list is ascending too
        int lo[3], hi[3];
        cell_range(list[k], s, lo, hi);
        for (int z = lo[2]; z <= hi[2]; z++)
          for (int y = lo[1]; y <= hi[1]; y++)
            for (int x = lo[0]; x <= hi[0]; x++) {
              const size_t c = s * K3 + ((size_t)z * K + y) * K + x;
              if (pass == 0) fstart[c]++; else flist[cursor[c]++] = list[k];
            }
      }
  }
}


// ---- surface grids ------------------------------------------------------------------------------------------
static double tri_area(const double* a, const double* b, const double* c) {
  P3 v0 = {a[0], a[1], a[2]}, v1 = {b[0], b[1], b[2]}, v2 = {c[0], c[1], c[2]};
  return 0.5 * std::sqrt(len2(crossp(sub(v1, v0), sub(v2, v0))));
}
static void one_grid(const DevWall& w, double area, DevGrid& g) {
  g.n_axis = (int)std::ceil(std::sqrt(area));
  if (g.n_axis < 1) g.n_axis = 1;
  const uint32_t n_tiles = (uint32_t)(g.n_axis * g.n_axis);
  g.strip_width_rcp = 1 / (w.uv2v / ((double)g.n_axis));
  g.vert2_slope = w.uv2u / w.uv2v;
  g.fullslope = w.uv1u / w.uv2v;
  g.binding_factor = ((double)n_tiles) / area;
  P3 v0 = {w.v0x, w.v0y, w.v0z};
  g.vert0_u = dotp(v0, P3{w.ux, w.uy, w.uz});
  g.vert0_v = dotp(v0, P3{w.vx, w.vy, w.vz});
}
uint64_t grid_constants(const double* verts, const uint32_t* tri, const std::vector<DevWall>& walls, std::vector<DevGrid>& out) {
  out.resize(walls.size());
  uint64_t total = 0;
  for (size_t i = 0; i < walls.size(); i++) {
    DevGrid g{};
    one_grid(walls[i], tri_area(verts + 3 * tri[3 * i], verts + 3 * tri[3 * i + 1], verts + 3 * tri[3 * i + 2]), g);
    g.tile_start = (uint32_t)total;
    total += (uint64_t)g.n_axis * g.n_axis;
    out[i] = g;
  }
  return total;
}
void wall_areas(const double* verts, const uint32_t* tri, uint64_t n_walls, std::vector<double>& out) {
  out.resize(n_walls);
  for (uint64_t i = 0; i < n_walls; i++) out[i] = tri_area(verts + 3 * tri[3 * i], verts + 3 * tri[3 * i + 1], verts + 3 * tri[3 * i + 2]);
}
static void one_triangle(const double* v9, DevWall& w, DevGrid& g) {
  const uint32_t t[3] = {0, 1, 2};
  std::vector<DevWall> ws;
  wall_constants(v9, t, 1, ws);
  w = ws[0];
  one_grid(w, tri_area(v9, v9 + 3, v9 + 6), g);
}
uint32_t tri_num_tiles(const double* v9) {
  DevWall w; DevGrid g{};
  one_triangle(v9, w, g);
  return (uint32_t)(g.n_axis * g.n_axis);
}
// GridUtils::grid2uv, src4/grid_utils.inl:233-253
void tri_grid2uv(const double* v9, uint32_t index, double* uv2) {
  DevWall w; DevGrid g{};
  one_triangle(v9, w, g);
  int root = (int)(std::sqrt((double)index));
  int rootrem = (int)index - root * root;
  int k = g.n_axis - root - 1;
  int j = rootrem / 2;
  int i = rootrem - 2 * j;
  double over3n = 1 / (double)(3 * g.n_axis);
  uv2[0] = ((double)(3 * j + i + 1)) * over3n * w.uv1u + ((double)(3 * k + i + 1)) * over3n * w.uv2u;
  uv2[1] = ((double)(3 * k + i + 1)) * over3n * w.uv2v;
}
static bool distinguishable_vec3(P3 a, P3 b, double eps) {  // src4/defines.h:766-806
  double c = std::fabs(a.x), cc, d;
  d = std::fabs(a.y); if (d > c) c = d;
  d = std::fabs(a.z); if (d > c) c = d;
  d = std::fabs(b.x); if (d > c) c = d;
  d = std::fabs(b.y); if (d > c) c = d;
  d = std::fabs(b.z); if (d > c) c = d;
  cc = std::fabs(a.x - b.x);
  d = std::fabs(a.y - b.y); if (d > cc) cc = d;
  d = std::fabs(a.z - b.z); if (d > cc) cc = d;
  if (c < eps) c = eps;
  return c * eps < cc;
}
// GridUtils::xyz2grid_tile_index, src4/grid_utils.inl:48-118
uint32_t tri_xyz2grid(const double* v9, const double* xyz3) {
  DevWall w; DevGrid g{};
  one_triangle(v9, w, g);
  const uint32_t n_tiles = (uint32_t)(g.n_axis * g.n_axis);
  if (n_tiles == 1) return 0;
  P3 v = {xyz3[0], xyz3[1], xyz3[2]};
  if (!distinguishable_vec3(v, P3{v9[0], v9[1], v9[2]}, 1e-12)) return n_tiles - 2 * (uint32_t)g.n_axis + 1;
  if (!distinguishable_vec3(v, P3{v9[3], v9[4], v9[5]}, 1e-12)) return n_tiles - 1;
  if (!distinguishable_vec3(v, P3{v9[6], v9[7], v9[8]}, 1e-12)) return 0;
  double i = dotp(v, P3{w.ux, w.uy, w.uz}) - g.vert0_u;
  double j = dotp(v, P3{w.vx, w.vy, w.vz}) - g.vert0_v;
  double striploc = j * g.strip_width_rcp;
  int strip = (int)striploc;
  double striprem = striploc - strip;
  strip = g.n_axis - strip - 1;
  double u0 = j * g.vert2_slope;
  double u1_u0 = w.uv1u - j * g.fullslope;
  double stripeloc = ((i - u0) / u1_u0) * (strip + (1 - striprem));
  int stripe = (int)stripeloc;
  double striperem = stripeloc - stripe;
  int flip = (striperem < 1 - striprem) ? 0 : 1;
  int idx = strip * strip + 2 * stripe + flip;
  if (idx < 0) idx = 0;
  if ((uint32_t)idx >= n_tiles) idx = (int)n_tiles - 1;
  return (uint32_t)idx;
}


// ---- triangle sides ---------------------------------------------------------------------------------------------
void edge_constants(const double* verts, const uint32_t* tri, const std::vector<DevWall>& walls, const uint32_t* wall_object,
                    std::vector<DevEdge>& out) {
  const size_t nw = walls.size();
  out.assign(3 * nw, DevEdge{0, 0, 0, 0, 0xFFFFFFFFu, 0, {0, 0}});
  auto vert = [&](size_t w, int k) { const double* q = verts + 3 * tri[3 * w + k]; return P3{q[0], q[1], q[2]}; };
  auto same = [](P3 a, P3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; };
  struct Key { double a[6]; bool operator<(const Key& o) const { return std::lexicographical_compare(a, a + 6, o.a, o.a + 6); } };
  for (size_t first = 0; first < nw;) {
    size_t last = first;
    while (last < nw && (!wall_object || wall_object[last] == wall_object[first])) last++;
    std::map<Key, std::pair<size_t, int>> open_edges;  // undirected edge -> (face, side) of the first face that has it
    for (size_t fi = first; fi < last; fi++) {
      for (int j = 0; j < 3; j++) {
        const int k = j + 1 < 3 ? j + 1 : 0;
        const P3 pj = vert(fi, j), pk = vert(fi, k);
        const double a[3] = {pj.x, pj.y, pj.z}, b[3] = {pk.x, pk.y, pk.z};
        const bool swap = std::lexicographical_compare(b, b + 3, a, a + 3);
        Key key;
        for (int q = 0; q < 3; q++) { key.a[q] = swap ? b[q] : a[q]; key.a[3 + q] = swap ? a[q] : b[q]; }
        auto it = open_edges.find(key);
        if (it == open_edges.end()) { open_edges[key] = {fi, j}; continue; }
        const size_t f0 = it->second.first; const int e0 = it->second.second;
        // compatible_edges: opposite directions, different third vertices, two different faces
        const P3 a0 = vert(f0, e0), a1 = vert(f0, e0 == 2 ? 0 : e0 + 1), a2 = vert(f0, e0 == 0 ? 2 : e0 - 1);
        const P3 b2 = vert(fi, j == 0 ? 2 : j - 1);
        if (!(same(a0, pk) && same(a1, pj) && !same(a2, b2)) || f0 == fi) continue;
        open_edges.erase(it);
        const DevWall& wf = walls[f0]; const DevWall& wb = walls[fi];
        const int i = e0, jj = i + 1 == 3 ? 0 : i + 1;
        const P3 wf0 = vert(f0, 0), wfi = vert(f0, i), wfj = vert(f0, jj), wb0 = vert(fi, 0);
        const P3 fu = {wf.ux, wf.uy, wf.uz}, fv = {wf.vx, wf.vy, wf.vz}, bu = {wb.ux, wb.uy, wb.uz}, bv = {wb.vx, wb.vy, wb.vz};
        const P3 di0 = sub(wfi, wf0);
        const double Ofu = dotp(di0, fu), Ofv = dotp(di0, fv);
        const P3 dj0 = sub(wfj, wf0);
        const double tfu = dotp(dj0, fu) - Ofu, tfv = dotp(dj0, fv) - Ofv;
        const double d_f = 1 / std::sqrt(tfu * tfu + tfv * tfv);
        const double efu = tfu * d_f, efv = tfv * d_f, ffu = -efv, ffv = efu;
        const P3 dib = sub(wfi, wb0);
        const double Obu = dotp(dib, bu), Obv = dotp(dib, bv);
        const P3 djb = sub(wfj, wb0);
        const double tbu = dotp(djb, bu) - Obu, tbv = dotp(djb, bv) - Obv;
        const double d_b = 1 / std::sqrt(tbu * tbu + tbv * tbv);
        const double ebu = tbu * d_b, ebv = tbv * d_b, fbu = -ebv, fbv = ebu;
        const double m00 = efu * ebu + ffu * fbu, m01 = efv * ebu + ffv * fbu;
        const double m10 = efu * ebv + ffu * fbv, m11 = efv * ebv + ffv * fbv;
        double qu = Obu, qv = Obv;
        qu -= m00 * Ofu + m01 * Ofv;
        qv -= m10 * Ofu + m11 * Ofv;
        out[3 * f0 + e0] = DevEdge{m00, m01, qu, qv, (uint32_t)fi, 1u, {0, 0}};
        out[3 * fi + j] = DevEdge{m00, m01, qu, qv, (uint32_t)f0, 0u, {0, 0}};
      }
    }
    first = last;
  }
}


// ---- neighbour tiles (react_2D_all_neighbors' partner search) ----------------------------------------------
// The reference finds the tiles around a tile anew for every surface molecule and step (GridUtils::find_neighbor_tiles,
// src4/grid_utils.inl:1754-1801): twelve tiles of the own grid for a tile in the interior (:1657-1737), for a tile on
// the rim the own tiles next to it plus the tiles of the walls across the sides it touches whose border vertices fall
// inside its own span of the shared side (:1285-1640, 946-1240), and for a corner tile one corner tile of every wall
// that meets it in that vertex alone (:743-868).  None of this depends on the molecules, so libmcx works the lists out
// ONCE on the host, for every tile, and the device only gathers (tn_start / tn_list).  What does depend on the state
// — a neighbouring wall whose grid does not exist yet contributes nothing (:1243, :783-790) — is a per-wall filter the
// device applies to the entries (every entry names its wall; leaving a wall out does not change the others).  The
// reference builds its list by pushing to the front; the table is stored front to back, i.e. in the order
// react_2D_all_neighbors walks it.
namespace {
struct TileAddr { int strip, stripe, flip; };  // strip counted from the side v0-v1, stripe from the side v0-v2
inline TileAddr tile_addr(int n_axis, uint32_t t) {
  const int root = (int)std::sqrt((double)t), rem = (int)t - root * root;
  return TileAddr{n_axis - root - 1, rem / 2, rem & 1};
}
inline bool interior_tile(int n_axis, uint32_t t) {  // :296-319
  const TileAddr a = tile_addr(n_axis, t);
  if (a.strip == 0 || a.stripe == 0 || a.strip + a.stripe == n_axis - 1) return false;
  return !(a.strip + a.stripe == n_axis - 2 && a.flip == 1);
}
inline int strip_above(int n_axis, uint32_t t) {  // move_strip_up :523-536
  const int root = (int)std::sqrt((double)t) + 1;
  return n_axis == root ? -1 : (int)t + 2 * root;
}
inline int strip_below(int n_axis, uint32_t t) {  // move_strip_down :551-590
  const TileAddr a = tile_addr(n_axis, t);
  const int per_strip = 2 * n_axis - 2 * a.strip - 1;
  const uint32_t n_tiles = (uint32_t)(n_axis * n_axis);
  bool ok;
  if (interior_tile(n_axis, t)) ok = true;
  else if (a.strip == 0 && a.stripe > 0) ok = t != n_tiles - 1;
  else ok = a.flip != 0;
  return ok ? (int)t - per_strip + 1 : -1;
}
struct Mesh {
  const double* verts; const uint32_t* tri; const std::vector<DevWall>* walls; const std::vector<DevGrid>* grids;
  const std::vector<DevEdge>* edges; std::vector<std::vector<uint32_t>> vertex_walls;
  P3 vtx(uint32_t w, int k) const { const double* q = verts + 3 * tri[3 * w + k]; return P3{q[0], q[1], q[2]}; }
  int n_axis(uint32_t w) const { return (*grids)[w].n_axis; }
  uint32_t across(uint32_t w, int e) const { return (*edges)[3 * w + e].nb_wall; }
};
// collected in push order; the reference's list reads back to front
struct Pushed {
  std::vector<uint32_t> wt;
  void add(uint32_t w, int t) { wt.push_back(w); wt.push_back((uint32_t)t); }
  void add_once(uint32_t w, int t) {
    for (size_t i = 0; i < wt.size(); i += 2) if (wt[i] == w && wt[i + 1] == (uint32_t)t) return;
    add(w, t);
  }
};
// the tiles of wall `other` along the side it shares with wall `w` that touch the start tile (:946-1240)
void across_side(const Mesh& m, uint32_t w, TileAddr a, P3 from, P3 to, int side, uint32_t other, Pushed& out) {
  const int N = m.n_axis(w), M = m.n_axis(other);
  const double dx = from.x - to.x, dy = from.y - to.y, dz = from.z - to.z;
  const double len = std::sqrt(dx * dx + dy * dy + dz * dz);
  double p1 = -1, p2 = -1;  // the start tile's vertices on the shared side, as distances from `from`
  auto span = [&](int k) { p1 = k * len / N; p2 = (k + 1) * len / N; };
  auto point = [&](int k) { p1 = k * len / N; };
  if (a.stripe == 0) {
    if (a.strip > 0) { if (!a.flip) span(a.strip); else point(a.strip + 1); }
    else if (side == 0) { if (!a.flip) span(a.stripe); else point(a.stripe + 1); }
    else if (side == 1) { if (!a.flip) point(a.strip); else span(a.strip); }
    else { if (!a.flip) span(a.strip); else point(a.strip + 1); }
  }
  if (a.strip == 0 && a.stripe > 0) {
    const bool at_end = a.stripe == N - 1 || (a.stripe == N - 2 && a.flip == 1);
    if (!at_end || side == 0) { if (!a.flip) span(a.stripe); else point(a.stripe + 1); }
    else if (side == 1) { if (!a.flip) span(a.strip); else point(a.strip + 1); }
  }
  if (a.strip > 0 && a.stripe > 0) { if (!a.flip) span(a.strip); else point(a.strip + 1); }
  // which two vertices of the other wall lie on the shared side (:682-728), and which of them is `from`
  int s1 = -1, s2 = -1;
  for (int k = 0; k < 3; k++) {
    const P3 q = m.vtx(other, k);
    bool shared = false;
    for (int j = 0; j < 3; j++) shared |= !distinguishable_vec3(q, m.vtx(w, j), 1e-12);
    if (!shared) continue;
    if (k == 0 || s1 < 0) s1 = k; else s2 = k;
  }
  const bool from_is_s1 = !distinguishable_vec3(from, m.vtx(other, s1), 1e-12);
  const int i_from = from_is_s1 ? s1 : s2, i_to = from_is_s1 ? s2 : s1;
  if (i_from > i_to) { p1 = len - p1; if (p2 > 0) p2 = len - p2; }
  const int other_side = (s1 + s2 == 1) ? 0 : (s1 + s2 == 2) ? 2 : 1;
  // border tiles of the other wall, grouped by the border vertex they meet in
  const int n_pos = M + 1;
  std::vector<double> pos(n_pos);
  std::vector<int> grp(3 * n_pos, -1);
  for (int i = 0; i < n_pos; i++) pos[i] = i * len / M;
  const int corner0 = M * M - 2 * M + 1;
  int last = other_side == 1 ? M * M - 1 : corner0;
  grp[2] = last;
  for (int i = 1; i < n_pos - 1; i++) {
    if (other_side == 0) { for (int k = 0; k < 3; k++) grp[3 * i + k] = last + k; last += 2; }
    else {
      grp[3 * i] = last; grp[3 * i + 1] = other_side == 1 ? last - 1 : last + 1;
      last = grp[3 * i + 2] = strip_below(M, (uint32_t)grp[3 * i + 1]);
    }
  }
  grp[3 * (n_pos - 1)] = last;
  // bisect / bisect_high (:871-915)
  auto below = [&](double v) { int lo = 0, hi = n_pos; while (hi - lo > 1) { const int mid = (hi + lo) / 2; if (pos[mid] > v) hi = mid; else lo = mid; } return lo; };
  auto above = [&](double v) { int lo = 0, hi = n_pos - 1; while (hi - lo > 1) { const int mid = (hi + lo) / 2; if (pos[mid] > v) hi = mid; else lo = mid; } return pos[lo] > v ? lo : hi; };
  const double hi_pos = p1 > p2 ? p1 : p2, lo_pos = p1 > p2 ? p2 : p1;
  const int i_hi = above(hi_pos);
  const int i_lo = lo_pos > 0 ? below(lo_pos) : -1;
  if (i_lo >= 0) {
    for (int i = i_lo + 1; i < i_hi; i++)
      for (int k = 0; k < 3; k++) out.add_once(other, grp[3 * i + k]);
  } else out.add_once(other, grp[3 * i_hi]);
}
void rim_tile(const Mesh& m, uint32_t w, uint32_t t, Pushed& out) {  // :1285-1640
  const int N = m.n_axis(w), ti = (int)t, n_tiles = N * N;
  const TileAddr a = tile_addr(N, t);
  const P3 v0 = m.vtx(w, 0), v1 = m.vtx(w, 1), v2 = m.vtx(w, 2);
  auto side = [&](int e, int as_side) {
    const uint32_t o = m.across(w, e);
    if (o == MCX_NONE) return;
    const P3 from = e == 1 ? v1 : v0, to = e == 0 ? v1 : v2;
    across_side(m, w, a, from, to, as_side, o, out);
  };
  auto own = [&](int q) { out.add(w, q); };
  if (a.stripe == 0) {
    if (a.flip) {
      own(ti - 1); own(ti + 1);
      if (a.strip < N - 2) own(ti + 2);
      int q = strip_below(N, t);
      own(q);
      if (a.strip < N - 2) { own(q + 1); own(q + 2); }
      if (a.strip > 0) { q = strip_above(N, t); own(q); own(q - 1); own(q + 1); }
      side(2, 2);
      if (a.strip == 0) side(0, 0);
      if (a.strip == N - 2) side(1, 0);  // the reference passes side index 0 here (:1408)
    } else if (t == 0) {
      if (n_tiles > 1) { const int q = strip_above(N, t); own(q); own(q - 1); own(q + 1); }
      else side(0, 0);
      side(1, 1); side(2, 2);
    } else {
      own(ti + 1); own(ti + 2);
      int q = strip_below(N, t + 1);
      own(q);
      if (a.strip > 0) { q = strip_above(N, t); own(q); own(q - 1); own(q + 1); own(q + 2); }
      else { side(0, 0); side(2, 2); }
    }
  } else if (a.strip == 0) {
    own(ti - 1); own(ti - 2);
    if (a.stripe < N - 2 || (a.stripe == N - 2 && !a.flip)) { own(ti + 1); own(ti + 2); }
    else if (a.stripe == N - 2) own(ti + 1);
    if (a.flip) {
      const int q = strip_below(N, t);
      own(q); own(q - 1); own(q - 2);
      if (a.stripe < N - 2) { own(q + 1); own(q + 2); }
    } else if (ti < n_tiles - 1) { const int q = strip_below(N, t); own(q); own(q - 1); own(q + 1); }
    else own(strip_below(N, t - 1));
    side(0, 0);
    if (ti >= n_tiles - 2) side(1, 1);
  } else {
    if (a.flip) {
      own(ti - 1); own(ti - 2); own(ti + 1);
      int q = strip_above(N, t); own(q); own(q - 1); own(q + 1);
      q = strip_below(N, t); own(q); own(q - 1); own(q - 2);
    } else {
      own(ti - 1); own(ti - 2);
      const int q = strip_above(N, t); own(q); own(q - 1); own(q - 2); own(q + 1);
      own(strip_below(N, t - 1));
    }
    side(1, 1);
  }
}
void interior(const Mesh& m, uint32_t w, uint32_t t, Pushed& out) {  // :1657-1737 with grid_neighbors :414-470
  const int N = m.n_axis(w), ti = (int)t;
  const int root = (int)std::sqrt((double)t), rem = ti - root * root, j = rem / 2, i = rem & 1;
  const int cand[3] = {i ? 2 * j + (root - 1) * (root - 1) : 1 + 2 * j + (root + 1) * (root + 1), ti + 1, ti - 1};
  int other_row = -1;
  for (int k = 0; k < 3; k++) if (cand[k] != ti - 1 && cand[k] != ti + 1) { other_row = cand[k]; break; }
  auto own = [&](int q) { out.add(w, q); };
  own(ti - 1); own(ti - 2); own(ti + 1); own(ti + 2);
  // tile_orientation (:483-510) of the tile's centre (grid2uv :233-253)
  const DevWall& f = (*m.walls)[w];
  const DevGrid& g = (*m.grids)[w];
  const int k3 = N - root - 1;
  const double over3n = 1 / (double)(3 * N);
  const double cu = ((double)(3 * j + i + 1)) * over3n * f.uv1u + ((double)(3 * k3 + i + 1)) * over3n * f.uv2u;
  const double cv = ((double)(3 * k3 + i + 1)) * over3n * f.uv2v;
  const double striploc = cv * g.strip_width_rcp;
  int strip = (int)striploc;
  const double striprem = striploc - strip;
  strip = N - strip - 1;
  const double stripeloc = ((cu - cv * g.vert2_slope) / (f.uv1u - cv * g.fullslope)) * (((double)strip) + (1 - striprem));
  const double striperem = stripeloc - (int)stripeloc;
  const bool upright = striperem < 1 - striprem;
  auto five = [&]() { own(other_row); own(other_row - 1); own(other_row - 2); own(other_row + 1); own(other_row + 2); };
  if (upright) { five(); const int q = strip_below(N, t); own(q); own(q - 1); own(q + 1); }
  else { const int q = strip_above(N, t); own(q); own(q - 1); own(q + 1); five(); }
}
}  // namespace

void tile_neighbor_table(const double* verts, uint64_t n_verts, const uint32_t* tri, const std::vector<DevWall>& walls,
                         const std::vector<DevGrid>& grids, const std::vector<DevEdge>& edges, std::vector<uint32_t>& start,
                         std::vector<uint32_t>& list) {
  Mesh m{verts, tri, &walls, &grids, &edges, {}};
  m.vertex_walls.resize(n_verts);
  for (uint32_t w = 0; w < walls.size(); w++)
    for (int k = 0; k < 3; k++) m.vertex_walls[tri[3 * w + k]].push_back(w);
  start.clear(); list.clear();
  for (uint32_t w = 0; w < walls.size(); w++) {
    const int N = grids[w].n_axis;
    const uint32_t n_tiles = (uint32_t)(N * N);
    for (uint32_t t = 0; t < n_tiles; t++) {
      start.push_back((uint32_t)(list.size() / 2));
      Pushed out;
      if (interior_tile(N, t)) interior(m, w, t, out);
      else {
        const int corner_of = t == n_tiles - 2 * (uint32_t)N + 1 ? 0 : t == n_tiles - 1 ? 1 : t == 0 ? 2 : -1;  // is_corner_tile :329-342
        // a 1-tile grid is all three corners at once (:624-670 test every vertex)
        for (int c = 0; c < 3 && corner_of >= 0; c++) {
          const bool here = (c == 0 && t == n_tiles - 2 * (uint32_t)N + 1) || (c == 1 && t == n_tiles - 1) || (c == 2 && t == 0);
          if (!here) continue;
          const uint32_t v = tri[3 * w + c];
          bool used_across = false;  // neighboring_wall_uses_this_vertex :605-622
          for (int e = 0; e < 3; e++) {
            const uint32_t o = m.across(w, e);
            if (o != MCX_NONE) for (int k = 0; k < 3; k++) used_across |= tri[3 * o + k] == v;
          }
          if (!used_across) continue;
          for (uint32_t o : m.vertex_walls[v]) {  // find_nbr_walls_shared_one_vertex (wall_utils.inl:79-104)
            if (o == w) continue;
            int same = 0;  // walls_share_full_edge (wall_utils.inl:50-65)
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) same += !distinguishable_vec3(m.vtx(w, a), m.vtx(o, b), 1e-12);
            if (same == 2) continue;
            // the corner tile of that wall (:743-868): under the LAST vertex index the two walls have in common
            int t_other;
            if (n_tiles == 1) t_other = 0;
            else {
              uint32_t common = MCX_NONE;
              for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) if (tri[3 * w + a] == tri[3 * o + b]) { common = tri[3 * o + b]; break; }
              const int M = grids[o].n_axis;
              t_other = common == tri[3 * o] ? M * M - 2 * M + 1 : common == tri[3 * o + 1] ? M * M - 1 : 0;
            }
            out.add(o, t_other);
          }
        }
        rim_tile(m, w, t, out);
      }
      for (size_t q = out.wt.size(); q >= 2; q -= 2) { list.push_back(out.wt[q - 2]); list.push_back(out.wt[q - 1]); }
    }
  }
  start.push_back((uint32_t)(list.size() / 2));
}
}  // namespace mcxg
