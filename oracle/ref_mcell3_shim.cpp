// oracle/ref_mcell3_shim.cpp — oracle/_ref build only (TEST INFRASTRUCTURE).
//
// Thin extern "C" entry points around the REFERENCE's own arithmetic for the hot path, compiled from the
// sources where they lie under /root/reference/src (MCell3 originals; MCell4's src4/*.inl versions were derived
// from them and kept MCell3-identical, include/debug_config.h:36-37):
//   init_tri_wall      src/wall_util.c:1214   (== Wall::initialize_wall_constants, src4/wall.cpp:281-342)
//   collide_wall       src/wall_util.c:825    (== CollisionUtils::collide_wall, src4/collision_utils.inl:629-812)
//   jump_away_line     src/wall_util.c:771    (== src4/collision_utils.inl:568-603)
//   collide_mol        src/wall_util.c:967    (== src4/collision_utils.inl:464-515)
//   wall_in_box        src/wall_util.c:1025   (== WallUtils::wall_in_box, src4/wall_utils.inl:326-504)
//   test_bimolecular / test_intersect / binary_search_double / timeof_unimolecular   src/react_cond.c
//   compute_pb_factor  src/react_util.c:45
//   distinguishable    src/util.c:449
// Nothing from the reference is copied into this repository: this file #includes the reference translation
// unit at build time (wall_util.c must be included, not linked, because wall_in_box is static).
// Unused reference functions are dropped by -ffunction-sections/--gc-sections, so their NFsim/MDL dependencies
// never need to resolve.
#include "wall_util.c"
#include "react_util.h"

#define EXPORT extern "C" __attribute__((visibility("default")))

namespace {
struct RefWall {
  struct geom_object obj;
  struct wall w;
  struct vector3 v[3];
  RefWall(const double* vv) {
    memset(&obj, 0, sizeof(obj));
    memset(&w, 0, sizeof(w));
    for (int k = 0; k < 3; k++) { v[k].x = vv[3 * k]; v[k].y = vv[3 * k + 1]; v[k].z = vv[3 * k + 2]; }
    obj.walls = &w;
    obj.n_walls = 1;
    init_tri_wall(&obj, 0, &v[0], &v[1], &v[2]);
  }
};
void seed_rng(struct rng_state* r, unsigned seed, unsigned skip) {
  rng_init(r, seed);
  for (unsigned i = 0; i < skip; i++) (void)rng_uint(r);
}
}  // namespace

EXPORT void ref3_init_tri_wall(const double* v9, double* out16) {
  RefWall rw(v9);
  const struct wall& w = rw.w;
  double t[16] = {w.normal.x, w.normal.y, w.normal.z, w.d, w.unit_u.x, w.unit_u.y, w.unit_u.z,
                  w.unit_v.x, w.unit_v.y, w.unit_v.z, w.uv_vert1_u, w.uv_vert2.u, w.uv_vert2.v, w.area, 0, 0};
  memcpy(out16, t, sizeof(t));
}

// returns the reference's COLLIDE_REDO(-1)/MISS(0)/FRONT(1)/BACK(2) (src/mcell_structs.h:201-206); move is in/out
EXPORT int ref3_collide_wall(const double* point3, double* move3, const double* v9, unsigned seed, unsigned skip,
                             double* t, double* hit3, long long* rng_words_used) {
  RefWall rw(v9);
  struct rng_state rng;
  seed_rng(&rng, seed, skip);
  long long before = rng_uses(&rng);
  struct notifications notify;
  memset(&notify, 0, sizeof(notify));
  long long tests = 0;
  struct vector3 p = {point3[0], point3[1], point3[2]}, m = {move3[0], move3[1], move3[2]}, h = {0, 0, 0};
  double tt = 0;
  int r = collide_wall(&p, &m, &rw.w, &tt, &h, 1, &rng, &notify, &tests);
  move3[0] = m.x; move3[1] = m.y; move3[2] = m.z;
  *t = tt; hit3[0] = h.x; hit3[1] = h.y; hit3[2] = h.z;
  *rng_words_used = rng_uses(&rng) - before;
  return r;
}

EXPORT int ref3_collide_mol(const double* point3, const double* move3, const double* target3, double rx_radius_3d,
                            double* t, double* hit3) {
  struct species sp;
  memset(&sp, 0, sizeof(sp));
  struct volume_molecule vm;
  memset(&vm, 0, sizeof(vm));
  vm.properties = &sp;
  vm.pos.x = target3[0]; vm.pos.y = target3[1]; vm.pos.z = target3[2];
  struct vector3 p = {point3[0], point3[1], point3[2]}, m = {move3[0], move3[1], move3[2]}, h = {0, 0, 0};
  double tt = 0;
  int r = collide_mol(&p, &m, (struct abstract_molecule*)&vm, &tt, &h, rx_radius_3d);
  *t = tt; hit3[0] = h.x; hit3[1] = h.y; hit3[2] = h.z;
  return r;
}

EXPORT int ref3_wall_in_box(const double* v9, const double* llf3, const double* urb3) {
  RefWall rw(v9);
  struct vector3 b0 = {llf3[0], llf3[1], llf3[2]}, b1 = {urb3[0], urb3[1], urb3[2]};
  return wall_in_box(rw.w.vert, &rw.w.normal, rw.w.d, &b0, &b1);
}

EXPORT int ref3_distinguishable(double a, double b, double eps) { return distinguishable(a, b, eps); }

static void make_rxn(struct rxn* rx, double* cum_probs, int n) {
  memset(rx, 0, sizeof(*rx));
  rx->n_pathways = n;
  rx->cum_probs = cum_probs;
  rx->max_fixed_p = cum_probs[n - 1];
  rx->min_noreaction_p = cum_probs[n - 1];
}

// returns RX_NO_RX (-2... see src/mcell_structs_shared.h) or the pathway index
EXPORT int ref3_test_bimolecular(double* cum_probs, int n, double scaling, unsigned seed, unsigned skip,
                                 long long* rng_words_used) {
  struct rxn rx;
  make_rxn(&rx, cum_probs, n);
  struct rng_state rng;
  seed_rng(&rng, seed, skip);
  long long before = rng_uses(&rng);
  int r = test_bimolecular(&rx, scaling, 0.0, NULL, NULL, &rng);
  *rng_words_used = rng_uses(&rng) - before;
  return r;
}
EXPORT int ref3_test_intersect(double* cum_probs, int n, double scaling, unsigned seed, unsigned skip,
                               long long* rng_words_used) {
  struct rxn rx;
  make_rxn(&rx, cum_probs, n);
  struct rng_state rng;
  seed_rng(&rng, seed, skip);
  long long before = rng_uses(&rng);
  int r = test_intersect(&rx, scaling, &rng);
  *rng_words_used = rng_uses(&rng) - before;
  return r;
}
// timeof_unimolecular (src/react_cond.c; == RxnUtils::time_of_unimol, src4/rxn_utils.inl:721-736): the lifetime drawn
// when a molecule with a unimolecular reaction class is created; FOREVER (1e20) when k_tot <= 0
EXPORT double ref3_timeof_unimolecular(double k_tot, unsigned seed, unsigned skip) {
  struct rxn rx;
  memset(&rx, 0, sizeof(rx));
  rx.max_fixed_p = k_tot;
  struct rng_state rng;
  seed_rng(&rng, seed, skip);
  return timeof_unimolecular(&rx, NULL, &rng);
}
// which_unimolecular (src/react_cond.c; == rxn_utils.inl:774-783): pathway of a firing unimolecular class
EXPORT int ref3_which_unimolecular(double* cum_probs, int n, unsigned seed, unsigned skip, long long* rng_words_used) {
  struct rxn rx;
  make_rxn(&rx, cum_probs, n);
  struct rng_state rng;
  seed_rng(&rng, seed, skip);
  long long before = rng_uses(&rng);
  int r = which_unimolecular(&rx, NULL, &rng);
  *rng_words_used = rng_uses(&rng) - before;
  return r;
}
EXPORT int ref3_binary_search_double(double* A, double match, int max_idx, double mult) {
  return binary_search_double(A, match, max_idx, mult);
}
EXPORT int ref3_rx_no_rx(void) { return RX_NO_RX; }

// volume-volume pb_factor for two volume reactants with the given D (cm^2/s); returns pb_factor
EXPORT double ref3_compute_pb_factor_volvol(double time_unit, double length_unit, double grid_density,
                                            double rx_radius_3d, double space_step_a, double time_step_a,
                                            double space_step_b, double time_step_b, int a_cant_initiate,
                                            int b_cant_initiate) {
  struct species sa, sb;
  memset(&sa, 0, sizeof(sa)); memset(&sb, 0, sizeof(sb));
  sa.space_step = space_step_a; sa.time_step = time_step_a; sa.D = 1;
  sb.space_step = space_step_b; sb.time_step = time_step_b; sb.D = 1;
  if (a_cant_initiate) sa.flags |= CANT_INITIATE;
  if (b_cant_initiate) sb.flags |= CANT_INITIATE;
  struct species* players[2] = {&sa, &sb};
  short geom[2] = {0, 0};
  struct rxn rx;
  memset(&rx, 0, sizeof(rx));
  rx.n_reactants = 2;
  rx.players = players;
  rx.geometries = geom;
  rx.get_reactant_diffusion = rxn_get_standard_diffusion;
  rx.get_reactant_time_step = rxn_get_standard_time_step;
  rx.get_reactant_space_step = rxn_get_standard_space_step;
  struct reaction_flags rf;
  memset(&rf, 0, sizeof(rf));
  int shared = 0;
  return compute_pb_factor(time_unit, length_unit, grid_density, rx_radius_3d, &rf, &shared, &rx, 0);
}

// compute_pb_factor (src/react_util.c:45-183) for two surface molecules: time_unit * grid_density / 6, or / 3 when one of the
// two cannot initiate
EXPORT double ref3_compute_pb_factor_surfsurf(double time_unit, double length_unit, double grid_density, int a_cant_initiate,
                                              int b_cant_initiate) {
  struct species sa, sb;
  memset(&sa, 0, sizeof(sa)); memset(&sb, 0, sizeof(sb));
  sa.flags = ON_GRID; sb.flags = ON_GRID;
  sa.D = 1; sb.D = 1;
  if (a_cant_initiate) sa.flags |= CANT_INITIATE;
  if (b_cant_initiate) sb.flags |= CANT_INITIATE;
  struct species* players[2] = {&sa, &sb};
  short geom[2] = {1, 1};
  struct rxn rx;
  memset(&rx, 0, sizeof(rx));
  rx.n_reactants = 2;
  rx.players = players;
  rx.geometries = geom;
  rx.get_reactant_diffusion = rxn_get_standard_diffusion;
  rx.get_reactant_time_step = rxn_get_standard_time_step;
  rx.get_reactant_space_step = rxn_get_standard_space_step;
  struct reaction_flags rf;
  memset(&rf, 0, sizeof(rf));
  int shared = 0;
  return compute_pb_factor(time_unit, length_unit, grid_density, 0.0, &rf, &shared, &rx, 0);
}

// ---- surface grids: src/grid_util.c (== Grid::initialize src4/wall.cpp:38-74, GridUtils::xyz2grid_tile_index /
// grid2uv src4/grid_utils.inl:48-118,233-253) -------------------------------------------------------------------
#include "grid_util.h"
namespace {
struct RefGrid {
  RefWall rw;
  struct surface_grid g;
  RefGrid(const double* v9) : rw(v9) {
    memset(&g, 0, sizeof(g));
    // the allocation-free part of create_grid (src/grid_util.c:305-365)
    g.surface = &rw.w;
    g.n = (int)ceil(sqrt(rw.w.area));
    if (g.n < 1) g.n = 1;
    g.n_tiles = g.n * g.n;
    g.binding_factor = ((double)g.n_tiles) / rw.w.area;
    init_grid_geometry(&g);
    rw.w.grid = &g;
  }
};
}  // namespace
EXPORT void ref3_grid_constants(const double* v9, double* out8) {
  RefGrid rg(v9);
  double t[8] = {(double)rg.g.n, rg.g.inv_strip_wid, rg.g.vert2_slope, rg.g.fullslope, rg.g.binding_factor,
                 rg.g.vert0.u, rg.g.vert0.v, (double)rg.g.n_tiles};
  memcpy(out8, t, sizeof(t));
}
EXPORT int ref3_xyz2grid(const double* v9, const double* xyz3) {
  RefGrid rg(v9);
  struct vector3 v = {xyz3[0], xyz3[1], xyz3[2]};
  return xyz2grid(&v, &rg.g);
}
// uv2grid (src/grid_util.c:134-204 == GridUtils::uv2grid_tile_index, src4/grid_utils.inl:119-190); the point must
// lie inside the wall (the reference aborts otherwise)
EXPORT int ref3_uv2grid(const double* v9, const double* uv2) {
  RefGrid rg(v9);
  struct vector2 v = {uv2[0], uv2[1]};
  return uv2grid(&v, &rg.g);
}
EXPORT void ref3_grid2uv(const double* v9, int idx, double* uv2) {
  RefGrid rg(v9);
  struct vector2 r;
  grid2uv(&rg.g, idx, &r);
  uv2[0] = r.u; uv2[1] = r.v;
}
EXPORT void ref3_uv2xyz(const double* v9, const double* uv2, double* xyz3) {
  RefWall rw(v9);
  struct vector2 a = {uv2[0], uv2[1]};
  struct vector3 b;
  uv2xyz(&a, &rw.w, &b);
  xyz3[0] = b.x; xyz3[1] = b.y; xyz3[2] = b.z;
}

// ---- exact_disk: src/diffuse.c:1365 (== ExactDiskUtils::exact_disk, src4/exact_disk_utils.inl:840-1145) ----------
#include "diffuse.h"
#include <vector>
// n walls (9 coordinates each) form the wall list of the subvolume; the moving species has no surface-class
// reactions (CAN_VOLWALL clear) and the expanded list is on, as in every MCell4 run with volume-volume reactions
EXPORT double ref3_exact_disk(const double* loc3, const double* mv3, double R, const double* target3, int n_walls,
                              const double* tri9) {
  std::vector<RefWall*> walls;
  std::vector<struct wall_list> wl(n_walls > 0 ? n_walls : 1);
  for (int i = 0; i < n_walls; i++) walls.push_back(new RefWall(tri9 + 9 * i));
  for (int i = 0; i < n_walls; i++) { wl[i].this_wall = &walls[i]->w; wl[i].next = i + 1 < n_walls ? &wl[i + 1] : NULL; }
  struct storage st;
  memset(&st, 0, sizeof(st));
  st.exdv = create_mem(sizeof(struct exd_vertex), 64);
  struct subvolume sv;
  memset(&sv, 0, sizeof(sv));
  sv.wall_head = n_walls ? &wl[0] : NULL;
  sv.local_storage = &st;
  struct species sp;
  memset(&sp, 0, sizeof(sp));
  struct volume_molecule moving, target;
  memset(&moving, 0, sizeof(moving)); memset(&target, 0, sizeof(target));
  moving.properties = &sp; target.properties = &sp;
  target.pos.x = target3[0]; target.pos.y = target3[1]; target.pos.z = target3[2];
  struct vector3 loc = {loc3[0], loc3[1], loc3[2]}, mv = {mv3[0], mv3[1], mv3[2]};
  double r = exact_disk(NULL, &loc, &mv, R, &sv, &moving, &target, 1, NULL, NULL, NULL);
  delete_mem(st.exdv);
  for (auto* w : walls) delete w;
  return r;
}

// ---- meshes: edge pairing and flattening transforms, edge crossing of a 2-D move ---------------------------------
//   surface_net          src/wall_util.c:265   (== Geometry surface_net, src4/geometry.cpp:258-356)
//   init_edge_transform  src/wall_util.c:360   (== Edge::reinit_edge_constants, src4/wall.cpp:134-235)
//   find_edge_point      src/wall_util.c:579   (== GeometryUtils::find_edge_point, src4/geometry_utils.inl:222-291)
//   traverse_surface     src/wall_util.c:656   (== GeometryUtils::traverse_surface, src4/geometry_utils.inl:305-342)
namespace {
struct RefMesh {
  std::vector<struct vector3> verts;
  std::vector<struct wall> walls;
  std::vector<struct wall*> faces;
  struct geom_object obj;
  struct storage store;
  bool ok = false;
  RefMesh(const double* v, unsigned nv, const unsigned* tri, unsigned nw) : verts(nv), walls(nw), faces(nw) {
    memset(&obj, 0, sizeof(obj));
    memset(&store, 0, sizeof(store));
    for (unsigned i = 0; i < nv; i++) { verts[i].x = v[3 * i]; verts[i].y = v[3 * i + 1]; verts[i].z = v[3 * i + 2]; }
    memset(walls.data(), 0, sizeof(struct wall) * nw);
    obj.walls = walls.data();
    obj.n_walls = (int)nw;
    store.join = create_mem(sizeof(struct edge), 4096);
    for (unsigned i = 0; i < nw; i++) {
      init_tri_wall(&obj, (int)i, &verts[tri[3 * i]], &verts[tri[3 * i + 1]], &verts[tri[3 * i + 2]]);
      walls[i].birthplace = &store;
      faces[i] = &walls[i];
    }
    ok = store.join != NULL && surface_net(faces.data(), (int)nw) != 1;
  }
  ~RefMesh() { if (store.join) delete_mem(store.join); }
};
}  // namespace

// per (wall, side): neighbour wall (or -1), 1 when this wall is the edge's forward wall, and the transform
// (cos, sin, translate u, translate v); rows of walls without a partner on that side keep -1 / zeros
EXPORT int ref3_mesh_edges(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls, int* nb_wall_out,
                           int* forward_out, double* transform_out) {
  RefMesh m(verts, n_verts, tri, n_walls);
  if (!m.ok) return 1;
  for (unsigned i = 0; i < n_walls; i++)
    for (int k = 0; k < 3; k++) {
      const struct wall* w = &m.walls[i];
      const struct edge* e = w->edges[k];
      const bool paired = w->nb_walls[k] != NULL && e != NULL && e->backward != NULL;
      nb_wall_out[3 * i + k] = paired ? (int)(w->nb_walls[k] - m.walls.data()) : -1;
      forward_out[3 * i + k] = paired && e->forward == w ? 1 : 0;
      double* t = transform_out + 4 * (3 * i + k);
      t[0] = paired ? e->cos_theta : 0; t[1] = paired ? e->sin_theta : 0;
      t[2] = paired ? e->translate.u : 0; t[3] = paired ? e->translate.v : 0;
    }
  return 0;
}

// find_edge_point of a single triangle: returns the reference's code (0..2 edge, -1 stays inside, -2 cannot tell)
EXPORT int ref3_find_edge_point(const double* v9, const double* loc2, const double* disp2, double* edgept2) {
  RefWall rw(v9);
  struct vector2 loc = {loc2[0], loc2[1]}, disp = {disp2[0], disp2[1]}, pt = {0, 0};
  const int r = find_edge_point(&rw.w, &loc, &disp, &pt);
  edgept2[0] = pt.u; edgept2[1] = pt.v;
  return r;
}

// traverse_surface for a batch of (wall, side, uv) queries on one mesh: neighbour wall index (-1: none) and new uv
EXPORT int ref3_traverse_surface(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                                 const unsigned* q_wall, const int* q_side, const double* q_uv, unsigned n_q, int* wall_out,
                                 double* uv_out) {
  RefMesh m(verts, n_verts, tri, n_walls);
  if (!m.ok) return 1;
  for (unsigned q = 0; q < n_q; q++) {
    struct vector2 loc = {q_uv[2 * q], q_uv[2 * q + 1]}, nl = {0, 0};
    struct wall* there = traverse_surface(&m.walls[q_wall[q]], &loc, q_side[q], &nl);
    wall_out[q] = there ? (int)(there - m.walls.data()) : -1;
    if (!there) nl.u = nl.v = 0;  // free edge: the reference leaves newloc to an unset transform, callers ignore it
    uv_out[2 * q] = nl.u; uv_out[2 * q + 1] = nl.v;
  }
  return 0;
}

// Own stand-ins for four engine functions that ray_trace_2D references only on its periodic-box and region-border
// branches (src/vol_util.c, src/count_util.c are not part of this build): reaching one of them is a harness error.
#include <cstdlib>
struct subvolume* next_subvol(struct vector3*, struct vector3*, struct subvolume*, double*, double*, double*, int, int) { abort(); }
struct subvolume* find_subvolume(struct volume*, struct vector3*, struct subvolume*) { abort(); }
struct wall* find_closest_wall(struct volume*, struct vector3*, double, struct vector2*, int*, struct species*, const char*,
                               struct string_buffer*, struct string_buffer*) { abort(); }
void update_hit_data(struct hit_data**, struct wall*, struct wall*, struct surface_molecule*, struct vector2, int, int) { abort(); }

// ray_trace_2D (src/diffuse.c; == DiffuseReactEvent::ray_trace_surf, src4/diffuse_react_event.cpp:1578-1725) for a
// surface molecule of a species without region-border reactions, no periodic box: the walk across triangle edges with
// reflection at free edges.  Returns the wall the move ends on (-1: ambiguous edge hit, the caller picks another
// displacement) and the end point in that wall's frame.
EXPORT int ref3_ray_trace_2d(const double* verts, unsigned n_verts, const unsigned* tri, unsigned n_walls,
                             const unsigned* q_wall, const double* q_uv, const double* q_disp, unsigned n_q, int* wall_out,
                             double* uv_out) {
  RefMesh m(verts, n_verts, tri, n_walls);
  if (!m.ok) return 1;
  static struct volume world;   // zero-initialised: no periodic box, nothing else is read on this path
  struct species sp;
  memset(&sp, 0, sizeof(sp));
  struct periodic_image box = {0, 0, 0};
  for (unsigned q = 0; q < n_q; q++) {
    struct surface_grid grid;
    memset(&grid, 0, sizeof(grid));
    grid.surface = &m.walls[q_wall[q]];
    struct surface_molecule sm;
    memset(&sm, 0, sizeof(sm));
    sm.properties = &sp;
    sm.grid = &grid;
    sm.s_pos.u = q_uv[2 * q]; sm.s_pos.v = q_uv[2 * q + 1];
    sm.periodic_box = &box;
    struct vector2 disp = {q_disp[2 * q], q_disp[2 * q + 1]}, pos = {0, 0};
    int kill_me = 0;
    struct rxn* rx = NULL;
    struct hit_data* hd = NULL;
    struct wall* w = ray_trace_2D(&world, &sm, &disp, &pos, &kill_me, &rx, &hd);
    wall_out[q] = w ? (int)(w - m.walls.data()) : -1;
    uv_out[2 * q] = w ? pos.u : 0; uv_out[2 * q + 1] = w ? pos.v : 0;
  }
  return 0;
}

