// mcx_device.cuh — per-molecule evaluation for sm_100a: displacement sampling, subpartition DDA,
// wall ray tracing, neighbour-cell partner search, reaction tests.  fp64 throughout, compiled with
// -fmad=false so every expression rounds like the reference's -march=core2 build (no FMA).
//
// Reference functions realised here (src4/): compute_vol_displacement (diffusion_utils.inl:366-432),
// rng_gauss (src/rng.c:173-218), collect_crossed_subparts / collect_neighboring_subparts
// (collision_utils_subparts.inl:38-300), get_closest_wall_collision / collide_wall / jump_away_line
// (collision_utils.inl:568-914), collide_mol (:464-515), sort_collisions_by_time
// (diffuse_react_event.cpp:341-364), test_bimolecular (rxn_utils.inl:336-414), reflect_from_wall
// (collision_utils.inl:1711-1747), cross_transparent_wall (diffuse_react_event.cpp:3007-3099),
// diffuse_single_molecule / get_max_time (:164-337), unimolecular scheduling (:1731-1826).
#pragma once
#include <cfloat>
#include "mcx_internal.h"
#include "zig_tables.inc"

#define MCX_EPS 1e-12
#define MCX_SQRT_EPS 1e-6
#define MCX_POS_SQRT2 1.41421356238  /* src4/defines.h:182 */
#define MCX_MAX_SP_WALLS 40
#define MCX_MAX_SP_MOLS 64

__device__ const double d_zig_y[128] = MCX_ZIG_YTAB_INIT;
__device__ const double d_zig_w[128] = MCX_ZIG_WTAB_INIT;
__device__ const unsigned int d_zig_k[128] = MCX_ZIG_KTAB_INIT;

struct ZigShared { double y[128]; double w[128]; unsigned int k[128]; };

__device__ __forceinline__ void zig_load(ZigShared* z) {
  for (int i = threadIdx.x; i < 128; i += blockDim.x) { z->y[i] = d_zig_y[i]; z->w[i] = d_zig_w[i]; z->k[i] = d_zig_k[i]; }
}

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double max3d(double a, double b, double c) { return fmax(fmax(a, b), c); }
__device__ __forceinline__ double abs_max_2vec(D3 a, D3 b) {
  return max3d(fmax(fabs(a.x), fabs(b.x)), fmax(fabs(a.y), fabs(b.y)), fmax(fabs(a.z), fabs(b.z)));
}
__device__ __forceinline__ bool cmp_eq_d(double a, double b, double eps) { return fabs(a - b) < eps; }
__device__ __forceinline__ bool distinguishable_d(double a, double b, double eps) {  // src/util.c:449-463
  double c = fabs(a - b);
  a = fabs(a);
  if (a < 1) a = 1;
  b = fabs(b);
  if (b < a) eps *= a; else eps *= b;
  return c > eps;
}

#include "mcx_philox.h"

// conflict-round epochs of an iteration: round r has epoch iteration * (max_rounds + 1) + r + 1
__device__ __forceinline__ unsigned int round_epoch0(const DevParams& p) {
  return (unsigned int)(p.iteration * (unsigned long long)(p.max_rounds + 1) + 1);
}

// Rare-path helpers kept out of line: every inlined copy of exp / log / log1p or of the ten Philox rounds costs
// 100-300 SASS instructions, and the diffuse kernels instantiate the stream code at many sites (43 % of
// k_diffuse_fast's code and 29 % of k_diffuse_slow's were such copies; both kernels stalled on instruction fetch,
// profiles/r01_l).  Same functions, same roundings: results are unchanged.
__device__ __noinline__ double mcx_exp(double x) { return exp(x); }
__device__ __noinline__ double mcx_log(double x) { return log(x); }
__device__ __noinline__ double mcx_log1p(double x) { return log1p(x); }
__device__ __noinline__ uint4 philox_block_call(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  uint32_t o[4];
  philox4x32_10(c0, c1, c2, c3, k0, k1, o);
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// ---- per-molecule word stream: Philox (production) or tape slice (replay) ----------------------
struct Stream {
  const uint32_t* tape; unsigned long long tape_left;
  uint32_t k0, k1, it_lo, it_hi, id;
  uint32_t b0, b1, b2, b3;  // current Philox block (scalars, not an array: dynamic indexing would put it in local memory)
  uint32_t used;
  const ZigShared* zig;

  __device__ __forceinline__ void init(const DevParams& p, uint32_t mol_id, const ZigShared* z) {
    zig = z; used = 0; id = mol_id;
    if (p.rng_mode == MCX_RNG_TAPE) {
      unsigned long long off = mol_id < p.n_ids ? p.tape_off[mol_id] : p.n_words;
      if (off > p.n_words) off = p.n_words;
      tape = p.tape + off; tape_left = p.n_words - off;
    } else {
      tape = nullptr; tape_left = 0;
      k0 = (uint32_t)p.seed; k1 = (uint32_t)(p.seed >> 32);
      it_lo = (uint32_t)p.iteration; it_hi = (uint32_t)(p.iteration >> 32);
      // block 0 is generated here, in line (every evaluation draws from it); later blocks out of line in next()
      uint32_t o[4]; philox4x32_10(0u, it_lo, it_hi, id, k0, k1, o); b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
    }
  }
  // a stream of the release domain (iteration | 2^63): disjoint from every diffusion stream of the same molecule
  __device__ __forceinline__ void init_release(const DevParams& p, uint32_t mol_id, const ZigShared* z) {
    zig = z; used = 0; id = mol_id; tape = nullptr; tape_left = 0;
    k0 = (uint32_t)p.seed; k1 = (uint32_t)(p.seed >> 32);
    it_lo = (uint32_t)p.iteration; it_hi = (uint32_t)(p.iteration >> 32) | 0x80000000u;
    uint32_t o[4]; philox4x32_10(0u, it_lo, it_hi, id, k0, k1, o); b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
  }
  __device__ __forceinline__ uint32_t next() {
    uint32_t w;
    if (tape) w = used < tape_left ? tape[used] : 0u;
    else {
      const uint32_t k = used & 3u;
      if (k == 0u && used != 0u) { const uint4 o = philox_block_call(used >> 2, it_lo, it_hi, id, k0, k1); b0 = o.x; b1 = o.y; b2 = o.z; b3 = o.w; }
      w = k == 0u ? b0 : (k == 1u ? b1 : (k == 2u ? b2 : b3));
    }
    used++;
    return w;
  }
  __device__ __forceinline__ double dbl() { return 2.3283064365386962890625e-10 * (double)next(); }
  // src/rng.c:173-218
  __device__ double gauss() {
    double x, y, sign;
    for (;;) {
      uint32_t bits = next();
      sign = (bits & 0x80u) ? -1.0 : 1.0;
      uint32_t region = bits & 0x7fu;
      uint32_t pos_within_region = bits & 0xffffff00u;
      x = (double)pos_within_region * zig->w[region];
      if (pos_within_region < zig->k[region]) break;
      if (region != 0) {
        double yB = zig->y[region];
        double yR = zig->y[region - 1] - yB;
        y = yB + yR * dbl();
      } else {
        const double R = MCX_ZIG_R;
        x = R - mcx_log1p(-dbl()) * (1.0 / R);
        y = mcx_exp(-R * (x - 0.5 * R)) * dbl();
      }
      if (!(y >= mcx_exp(-0.5 * x * x))) break;
    }
    return sign * x;
  }
};

// ---- evaluation result ----------------------------------------------------------------------------
struct Outcome {
  int kind;               // MCX_OUT_*
  D3 pos;
  double t_now, unimol_time, t_event;
  uint32_t flags;         // device flag bits (DF_*)
  int rxn_class, pathway;
  uint32_t partner_slot, partner_id;
  uint32_t orient_bits;   // bit k: random orientation drawn for products[k] (1 = up); bit 4 + r: for kept reactant r;
                          // ORIENT_BIT_FRONT: a volume-surface reaction whose initiator hit the FRONT of the wall
  uint32_t s_wall, s_tile; double s_u, s_v;  // surface molecule that diffused: where it is now (Molecule::s)
  bool surf_moved;        // the surface fields above differ from the snapshot's
};
// surface part of a molecule's state handed to the evaluation (MCX_NONE wall: volume molecule)
struct SurfState { uint32_t wall, tile; double u, v; };

struct Tracer {
  mcx_trace_rec* tr;
  unsigned long long h;
  __device__ __forceinline__ void ev(uint32_t a, uint32_t b) {
    h = (h ^ a) * 0x100000001b3ULL;
    h = (h ^ b) * 0x100000001b3ULL;
  }
};
enum { EV_WALL = 0x57000000u, EV_COLL = 0xC0000000u, EV_RXN = 0xAE000000u, EV_ABSORB = 0xAB000000u,
       EV_REDO = 0x4ED00000u, EV_UNIMOL = 0x11000000u, EV_TRANSP = 0x7A000000u, EV_SURFMOL = 0x5F000000u, EV_BLOCKED = 0xB10C0000u, EV_DISK = 0xD1500000u,
       EV_SURFMOVE = 0x3E000000u, EV_WALLRXN = 0x9A000000u, EV_SURFSURF = 0x55000000u };

struct LocalStats {
  unsigned int ray_polygon_tests, ray_polygon_colls, reflections, transparent, volvol_collisions, redos;
};

// ---- subpartition index math (src4/partition.h:248-309) -------------------------------------------------
__device__ __forceinline__ bool in_partition(const DevParams& p, D3 q) {
  return q.x >= p.ox && q.y >= p.oy && q.z >= p.oz && q.x < p.ox + p.part_len && q.y < p.oy + p.part_len &&
         q.z < p.oz + p.part_len;
}
__device__ __forceinline__ void subpart_3d(const DevParams& p, D3 q, int idx[3]) {
  idx[0] = (int)((q.x - p.ox) * p.sp_rcp);
  idx[1] = (int)((q.y - p.oy) * p.sp_rcp);
  idx[2] = (int)((q.z - p.oz) * p.sp_rcp);
}
__device__ __forceinline__ uint32_t subpart_from_3d(const DevParams& p, int x, int y, int z) {
  return (uint32_t)(x + y * p.n_sp + z * p.n_sp * p.n_sp);
}
__device__ __forceinline__ uint32_t subpart_index(const DevParams& p, D3 q) {
  int i[3]; subpart_3d(p, q, i); return subpart_from_3d(p, i[0], i[1], i[2]);
}
__device__ __forceinline__ bool idx_in_range(const DevParams& p, int i) { return i >= 0 && i < p.n_sp; }

// ---- device neighbour-cell grid --------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(double v, double origin, double rcp, int n) {
  int c = (int)floor((v - origin) * rcp);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}
// local z-layer of a position: global layer (same expression on every rank) minus this rank's offset
__device__ __forceinline__ int cell_z(const DevParams& p, double z) {
  int c = (int)floor((z - p.cgz) * p.cell_rcp_z) - p.z_off;
  return c < 0 ? 0 : (c >= p.ncz ? p.ncz - 1 : c);
}
// Order of the cell rows in the sorted snapshot.  The records of one x-row of cells are contiguous; the rows themselves
// are stored in blocks of 2^rb x 2^rb (y, z) rows, so that the rows above and below a row (z -+ 1) lie within one block
// (a few MB) instead of a whole z-layer (12 MB at 1e8 molecules) away: with the plain (y, z) order every record was
// fetched from DRAM three times per iteration, once per z-layer that probes it (profiles/r02_a: 18.3 GB for 8.3 GB).
__device__ __forceinline__ uint32_t row_index(const DevParams& p, int cy, int cz) {
  const int b = p.rb_log2, m = (1 << b) - 1;
  return (uint32_t)((((((cz >> b) * p.nby + (cy >> b)) << b) | (cz & m)) << b) | (cy & m));
}
__device__ __forceinline__ uint32_t row_base(const DevParams& p, int cy, int cz) { return (uint32_t)p.ncx * row_index(p, cy, cz); }
__device__ __forceinline__ uint32_t cell_of(const DevParams& p, double x, double y, double z) {
  int cx = cell_coord(x, p.cgx, p.cell_rcp_x, p.ncx);
  int cy = cell_coord(y, p.cgy, p.cell_rcp_y, p.ncy);
  int cz = cell_z(p, z);
  return (uint32_t)cx + row_base(p, cy, cz);
}
// multi-GPU ownership: a position belongs to the rank whose owned z-layers contain it
__device__ __forceinline__ bool owned_z(const DevParams& p, double z) {
  const int cz = cell_z(p, z);
  return cz >= p.own_lo && cz < p.own_hi;
}

struct SpSet { uint32_t v[MCX_MAX_SP_MOLS]; int n; bool overflow; };
struct SpList { uint32_t v[MCX_MAX_SP_WALLS]; int n; };

__device__ __noinline__ void spset_insert(SpSet& s, uint32_t sp) {
  for (int i = 0; i < s.n; i++) if (s.v[i] == sp) return;
  if (s.n < MCX_MAX_SP_MOLS) s.v[s.n++] = sp; else s.overflow = true;
}
__device__ __forceinline__ bool spset_has(const SpSet& s, uint32_t sp) {
  for (int i = 0; i < s.n; i++) if (s.v[i] == sp) return true;
  return false;
}

// collect_neighboring_subparts, collision_utils_subparts.inl:38-122
__device__ void collect_neighboring_subparts(const DevParams& p, D3 pos, const int si[3], double rr, SpSet& out) {
  const double sp_len = p.sp_len, part_len = p.part_len;
  D3 rel = {pos.x - p.ox, pos.y - p.oy, pos.z - p.oz};
  D3 plus = {rel.x + rr, rel.y + rr, rel.z + rr};
  D3 minus = {rel.x - rr, rel.y - rr, rel.z - rr};
  D3 boundary = {si[0] * sp_len, si[1] * sp_len, si[2] * sp_len};
  int xd = 0, yd = 0, zd = 0;
  if (minus.x < boundary.x && minus.x > 0.0) { spset_insert(out, subpart_from_3d(p, si[0] - 1, si[1], si[2])); xd = -1; }
  else if (plus.x > boundary.x + sp_len && plus.x < part_len) { spset_insert(out, subpart_from_3d(p, si[0] + 1, si[1], si[2])); xd = 1; }
  if (minus.y < boundary.y && minus.y > 0.0) { spset_insert(out, subpart_from_3d(p, si[0], si[1] - 1, si[2])); yd = -1; }
  else if (plus.y > boundary.y + sp_len && plus.y < part_len) { spset_insert(out, subpart_from_3d(p, si[0], si[1] + 1, si[2])); yd = 1; }
  if (minus.z < boundary.z && minus.z > 0.0) { spset_insert(out, subpart_from_3d(p, si[0], si[1], si[2] - 1)); zd = -1; }
  else if (plus.z > boundary.z + sp_len && plus.z < part_len) { spset_insert(out, subpart_from_3d(p, si[0], si[1], si[2] + 1)); zd = 1; }
  if (xd && yd) spset_insert(out, subpart_from_3d(p, si[0] + xd, si[1] + yd, si[2]));
  if (xd && zd) spset_insert(out, subpart_from_3d(p, si[0] + xd, si[1], si[2] + zd));
  if (yd && zd) spset_insert(out, subpart_from_3d(p, si[0], si[1] + yd, si[2] + zd));
  if (xd && yd && zd) spset_insert(out, subpart_from_3d(p, si[0] + xd, si[1] + yd, si[2] + zd));
}

// collect_crossed_subparts, collision_utils_subparts.inl:127-300
__device__ uint32_t collect_crossed_subparts(const DevParams& p, D3 pos, uint32_t cur_subpart, D3 disp,
                                             bool for_mols, bool for_walls, SpList& spw, SpSet& spm) {
  const double sp_len = p.sp_len;
  if (for_walls) { spw.n = 0; spw.v[spw.n++] = cur_subpart; }
  if (for_mols) spset_insert(spm, cur_subpart);
  D3 dest = pos + disp;
  D3 dnz = disp;
  if (dnz.x == 0) dnz.x = FLT_MIN;
  if (dnz.y == 0) dnz.y = FLT_MIN;
  if (dnz.z == 0) dnz.z = FLT_MIN;
  int dir[3] = {dnz.x > 0 ? 1 : 0, dnz.y > 0 ? 1 : 0, dnz.z > 0 ? 1 : 0};
  int src[3], dst[3];
  src[0] = cur_subpart % p.n_sp; src[1] = (cur_subpart / p.n_sp) % p.n_sp; src[2] = (cur_subpart / (p.n_sp * p.n_sp)) % p.n_sp;
  subpart_3d(p, dest, dst);
  const double rr = p.R * MCX_POS_SQRT2;
  const bool expanded = p.use_expanded != 0;
  if (for_mols && expanded) collect_neighboring_subparts(p, pos, src, rr, spm);
  uint32_t dest_subpart = subpart_from_3d(p, dst[0], dst[1], dst[2]);
  if (cur_subpart != dest_subpart) {
    int add[3] = {dir[0] ? 1 : -1, dir[1] ? 1 : -1, dir[2] ? 1 : -1};
    D3 cur = pos;
    int ci[3] = {src[0], src[1], src[2]};
    uint32_t cs;
    D3 rcp = {1.0 / dnz.x, 1.0 / dnz.y, 1.0 / dnz.z};
    int guard = 0;
    do {
      D3 edges = {p.ox + ci[0] * sp_len + dir[0] * sp_len, p.oy + ci[1] * sp_len + dir[1] * sp_len,
                  p.oz + ci[2] * sp_len + dir[2] * sp_len};
      D3 diff = edges - cur;
      D3 ct = {diff.x * rcp.x, diff.y * rcp.y, diff.z * rcp.z};
      if (ct.x < ct.y && ct.x <= ct.z) {
        cur = cur + disp * ct.x; ci[0] += add[0];
        if (!idx_in_range(p, ci[0])) break;
      } else if (ct.y <= ct.z) {
        cur = cur + disp * ct.y; ci[1] += add[1];
        if (!idx_in_range(p, ci[1])) break;
      } else {
        cur = cur + disp * ct.z; ci[2] += add[2];
        if (!idx_in_range(p, ci[2])) break;
      }
      cs = subpart_from_3d(p, ci[0], ci[1], ci[2]);
      if (for_walls) { if (spw.n < MCX_MAX_SP_WALLS) spw.v[spw.n++] = cs; else spm.overflow = true; }
      if (for_mols) spset_insert(spm, cs);
      if (for_mols && expanded) collect_neighboring_subparts(p, cur, ci, rr, spm);
      if (++guard > 4096) break;
    } while (cs != dest_subpart);
  }
  if (for_mols && expanded) collect_neighboring_subparts(p, dest, dst, rr, spm);
  return dest_subpart;
}

// get_displacement_up_to_partition_boundary, collision_utils.inl:48-114
__device__ D3 displacement_up_to_partition_boundary(const DevParams& p, D3 pos, D3 disp) {
  D3 dnz = disp;
  if (dnz.x == 0) dnz.x = FLT_MIN;
  if (dnz.y == 0) dnz.y = FLT_MIN;
  if (dnz.z == 0) dnz.z = FLT_MIN;
  D3 edges = {p.ox + (dnz.x > 0 ? 1.0 : 0.0) * p.part_len, p.oy + (dnz.y > 0 ? 1.0 : 0.0) * p.part_len,
              p.oz + (dnz.z > 0 ? 1.0 : 0.0) * p.part_len};
  D3 diff = edges - pos;
  double hit_time = 1;
  if (fabs(diff.x) < MCX_EPS || fabs(diff.y) < MCX_EPS || fabs(diff.z) < MCX_EPS) return {0, 0, 0};
  D3 ct = {diff.x / dnz.x, diff.y / dnz.y, diff.z / dnz.z};
  if (ct.x >= 0 && ct.x < ct.y && ct.x <= ct.z) hit_time = ct.x;
  else if (ct.y >= 0 && ct.y <= ct.z) hit_time = ct.y;
  else if (ct.z >= 0) hit_time = ct.z;
  return disp * (hit_time - MCX_EPS);
}

enum { W_MISS = 0, W_FRONT = 1, W_BACK = 2, W_REDO = 3 };

__device__ __forceinline__ D3 wall_vertex(const DevParams& p, uint32_t wi, int k) {
  uint32_t v = p.wall_tri[3 * wi + k];
  return {p.verts[3 * v], p.verts[3 * v + 1], p.verts[3 * v + 2]};
}

// jump_away_line, collision_utils.inl:568-603
__device__ void jump_away_line(D3 pt, double k, D3 A, D3 B, D3 n, Stream& rs, D3& v) {
  D3 e = B - A;
  double le_1 = 1.0 / sqrt(dot3(e, e));
  e = e * le_1;
  D3 f = {n.y * e.z - n.z * e.y, n.z * e.x - n.x * e.z, n.x * e.y - n.y * e.x};
  double tiny = MCX_EPS * (abs_max_2vec(pt, v) + 1.0) / (k * max3d(fabs(f.x), fabs(f.y), fabs(f.z)));
  if ((rs.next() & 1u) == 0u) tiny = -tiny;
  v.x -= tiny * f.x; v.y -= tiny * f.y; v.z -= tiny * f.z;
}

// collide_wall, collision_utils.inl:629-812 (update_move = true)
__device__ int collide_wall(const DevParams& p, D3 pos, uint32_t wi, Stream& rs, D3& move, double& t, D3& hit) {
  const DevWall& f = p.walls[wi];
  const D3 n = {f.nx, f.ny, f.nz};
  double dp = dot3(n, pos), dv = dot3(n, move), dd = dp - f.dist, d_eps;
  if (dd > 0) {
    d_eps = MCX_EPS;
    if (dd < d_eps) d_eps = 0.5 * dd;
    if (dd + dv > d_eps) return W_MISS;
  } else {
    d_eps = -MCX_EPS;
    if (dd > d_eps) d_eps = 0.5 * dd;
    if (dd < 0 && dd + dv < d_eps) return W_MISS;
  }
  double a;
  if (dd == 0) {
    if (dv != 0) return W_MISS;
    a = (abs_max_2vec(pos, move) + 1.0) * MCX_EPS;
    if ((rs.next() & 1u) == 0u) a = -a;
    move = move - n * a;
    return W_REDO;
  }
  a = 1.0 / dv;
  a *= -dd;
  t = a;
  hit = pos + move * a;
  D3 v0 = {f.v0x, f.v0y, f.v0z};
  D3 local = hit - v0;
  double b = dot3(local, D3{f.ux, f.uy, f.uz}), c = dot3(local, D3{f.vx, f.vy, f.vz}), ff;
  if (f.uv2v < 0) { c = -c; ff = -f.uv2v; } else ff = f.uv2v;
  if (c > 0) {
    double g = b * ff, hh = c * f.uv2u;
    if (g > hh) {
      if (c * f.uv1u + g < hh + f.uv1u * f.uv2v) return dv > 0 ? W_BACK : W_FRONT;
      else if (!distinguishable_d(c * f.uv1u + g, hh + f.uv1u * f.uv2v, MCX_EPS)) {
        jump_away_line(pos, a, wall_vertex(p, wi, 1), wall_vertex(p, wi, 2), n, rs, move);
        return W_REDO;
      } else return W_MISS;
    } else if (!distinguishable_d(g, hh, MCX_EPS)) {
      jump_away_line(pos, a, wall_vertex(p, wi, 2), v0, n, rs, move);
      return W_REDO;
    } else return W_MISS;
  } else if (!distinguishable_d(c, 0.0, MCX_EPS)) {
    jump_away_line(pos, a, v0, wall_vertex(p, wi, 1), n, rs, move);
    return W_REDO;
  }
  return W_MISS;
}

// ---- counted volumes of intersecting objects (include/mcx.h: mcx_set_counted_volume_objects) ------------------------
__device__ __forceinline__ uint32_t cv_lookup(const DevParams& p, uint32_t mask) {
  for (uint32_t k = 0; k < p.n_cv; k++) if (__ldg(p.cv_mask + k) == mask) return k;
  return MCX_NONE;
}
__device__ __forceinline__ bool cv_uses_xor(const DevParams& p, uint32_t wall) {
  if (!p.cv_mask) return false;
  const uint32_t obj = p.wall_obj[wall];
  return obj < 32u && ((p.cv_xor >> obj) & 1u);
}
// update_counted_volume_id_when_crossing_wall (collision_utils.inl:1637-1694): the volume behind a wall hit on its front,
// in front of one hit on its back; a wall of an intersecting object toggles its object in the molecule's set instead.
// MCX_NONE: the resulting set is not in the table.
__device__ __forceinline__ uint32_t cv_cross(const DevParams& p, uint32_t cvi, uint32_t wall, bool hit_front) {
  if (!cv_uses_xor(p, wall)) { const uint32_t cv = __ldg(p.wall_cv + wall); return hit_front ? (cv >> 8) : (cv & 0xFFu); }
  return cv_lookup(p, __ldg(p.cv_mask + cvi) ^ (1u << p.wall_obj[wall]));
}

// ---- point location by a ray cast (region releases; the reference's Region::is_point_inside, geometry.cpp:1048-1086,
// and compute_counted_volume_for_pos, collision_utils.inl:1515-1566, count the walls crossed on the way to a far point).
// One ray from pos towards -x (skewed in y and z) up to the partition boundary, walked through the subpartitions it
// crosses with the arithmetic of collect_crossed_subparts; every wall it crosses is counted once, in the subpartition
// that holds the crossing point.  inside_mask: bit k = odd number of crossings of object k's walls (k < 32);
// first_wall / first_side: the nearest crossing (the counted volume of pos is the one on that side of that wall).
struct RayScan { uint32_t inside_mask; uint32_t first_wall; int first_side; bool redo; };
__device__ void scan_ray(const DevParams& p, D3 pos, Stream& rs, RayScan& out) {
  out.inside_mask = 0; out.first_wall = MCX_NONE; out.first_side = W_MISS; out.redo = false;
  const D3 raw = {-p.part_len, p.part_len * (1.0 / 11.0), p.part_len * (1.0 / 22.0)};
  const D3 disp = displacement_up_to_partition_boundary(p, pos, raw);
  if (disp.x == 0 && disp.y == 0 && disp.z == 0) return;  // on the partition boundary: outside everything
  double first_t = 2.0;
  auto visit = [&](uint32_t S) -> bool {
    const uint32_t w0 = p.spw_start[S], w1 = p.spw_start[S + 1];
    for (uint32_t k = w0; k < w1; k++) {
      const uint32_t wi = p.spw_list[k];
      D3 move = disp, hit; double t;
      const int ct = collide_wall(p, pos, wi, rs, move, t, hit);
      if (ct == W_REDO) { out.redo = true; return false; }
      if ((ct == W_FRONT || ct == W_BACK) && subpart_index(p, hit) == S) {
        const uint32_t obj = p.wall_obj ? p.wall_obj[wi] : 0u;
        if (obj < 32u) out.inside_mask ^= 1u << obj;
        if (t < first_t) { first_t = t; out.first_wall = wi; out.first_side = ct; }
      }
    }
    return true;
  };
  // the walk of collect_crossed_subparts (for walls), streamed
  const double sp_len = p.sp_len;
  uint32_t cur_subpart = subpart_index(p, pos);
  if (!visit(cur_subpart)) return;
  D3 dest = pos + disp;
  D3 dnz = disp;
  if (dnz.x == 0) dnz.x = FLT_MIN;
  if (dnz.y == 0) dnz.y = FLT_MIN;
  if (dnz.z == 0) dnz.z = FLT_MIN;
  int dir[3] = {dnz.x > 0 ? 1 : 0, dnz.y > 0 ? 1 : 0, dnz.z > 0 ? 1 : 0};
  int src[3], dst[3];
  src[0] = cur_subpart % p.n_sp; src[1] = (cur_subpart / p.n_sp) % p.n_sp; src[2] = (cur_subpart / (p.n_sp * p.n_sp)) % p.n_sp;
  subpart_3d(p, dest, dst);
  const uint32_t dest_subpart = subpart_from_3d(p, dst[0], dst[1], dst[2]);
  if (cur_subpart == dest_subpart) return;
  int add[3] = {dir[0] ? 1 : -1, dir[1] ? 1 : -1, dir[2] ? 1 : -1};
  D3 cur = pos;
  int ci[3] = {src[0], src[1], src[2]};
  uint32_t cs;
  D3 rcp = {1.0 / dnz.x, 1.0 / dnz.y, 1.0 / dnz.z};
  int guard = 0;
  do {
    D3 edges = {p.ox + ci[0] * sp_len + dir[0] * sp_len, p.oy + ci[1] * sp_len + dir[1] * sp_len,
                p.oz + ci[2] * sp_len + dir[2] * sp_len};
    D3 diff = edges - cur;
    D3 ct = {diff.x * rcp.x, diff.y * rcp.y, diff.z * rcp.z};
    if (ct.x < ct.y && ct.x <= ct.z) {
      cur = cur + disp * ct.x; ci[0] += add[0];
      if (!idx_in_range(p, ci[0])) break;
    } else if (ct.y <= ct.z) {
      cur = cur + disp * ct.y; ci[1] += add[1];
      if (!idx_in_range(p, ci[1])) break;
    } else {
      cur = cur + disp * ct.z; ci[2] += add[2];
      if (!idx_in_range(p, ci[2])) break;
    }
    cs = subpart_from_3d(p, ci[0], ci[1], ci[2]);
    if (!visit(cs)) return;
    if (++guard > 4096) break;
  } while (cs != dest_subpart);
}

// ---- fine wall grid (mcx_geom.cpp: bin_walls_fine) -------------------------------------------------------------------
// Cells of subpartition S under the box [lo, hi] (the caller has padded it by at least MCX_FW_MARGIN).
struct FwRange { int x0, x1, y0, y1, z0, z1; uint32_t base; };
__device__ __forceinline__ FwRange fw_range(const DevParams& p, uint32_t S, D3 lo, D3 hi) {
  const int K = p.fw_K, n = p.n_sp;
  const int sx = (int)(S % (uint32_t)n), sy = (int)((S / (uint32_t)n) % (uint32_t)n), sz = (int)(S / (uint32_t)(n * n));
  const double ox = p.ox + sx * p.sp_len, oy = p.oy + sy * p.sp_len, oz = p.oz + sz * p.sp_len;
  auto cell = [&](double v, double o) { const int c = (int)floor((v - o) * p.fw_rcp); return c < 0 ? 0 : (c >= K ? K - 1 : c); };
  FwRange r;
  r.x0 = cell(lo.x, ox); r.x1 = cell(hi.x, ox);
  r.y0 = cell(lo.y, oy); r.y1 = cell(hi.y, oy);
  r.z0 = cell(lo.z, oz); r.z1 = cell(hi.z, oz);
  r.base = S * (uint32_t)(K * K * K);
  return r;
}
__device__ __forceinline__ void segment_box(D3 pos, D3 move, double pad, D3& lo, D3& hi) {
  lo = D3{fmin(pos.x, pos.x + move.x) - pad, fmin(pos.y, pos.y + move.y) - pad, fmin(pos.z, pos.z + move.z) - pad};
  hi = D3{fmax(pos.x, pos.x + move.x) + pad, fmax(pos.y, pos.y + move.y) + pad, fmax(pos.z, pos.z + move.z) + pad};
}
// position of wall wi in the (ascending) list of its subpartition, MCX_NONE if it is not there
__device__ __forceinline__ uint32_t spw_position(const DevParams& p, uint32_t w0, uint32_t w1, uint32_t wi) {
  uint32_t a = w0, b = w1;
  while (a < b) { const uint32_t mid = (a + b) >> 1; if (__ldg(p.spw_list + mid) < wi) a = mid + 1; else b = mid; }
  return (a < w1 && __ldg(p.spw_list + a) == wi) ? a : MCX_NONE;
}
// Walk over the walls of subpartition S near a box in the order of the subpartition's own list (ascending wall index),
// each wall once: a merge of the (ascending) lists of the at most 8 cells under the box.  Larger boxes, and fw_K == 1,
// walk the whole subpartition list, which is always correct.
struct WallWalk {
  uint32_t cur[8], end[8];
  int n;
  uint32_t last, k, k1;
  bool whole, have_last;
  __device__ void init(const DevParams& p, uint32_t S, D3 lo, D3 hi) {
    whole = true; k = p.spw_start[S]; k1 = p.spw_start[S + 1]; n = 0; have_last = false; last = 0;
    if (p.fw_K == 1 || k == k1) return;
    const FwRange r = fw_range(p, S, lo, hi);
    if ((r.x1 - r.x0 + 1) * (r.y1 - r.y0 + 1) * (r.z1 - r.z0 + 1) > 8) return;
    whole = false;
    const int K = p.fw_K;
    for (int z = r.z0; z <= r.z1; z++)
      for (int y = r.y0; y <= r.y1; y++)
        for (int x = r.x0; x <= r.x1; x++) {
          const uint32_t c = r.base + (uint32_t)((z * K + y) * K + x);
          cur[n] = __ldg(p.fw_start + c); end[n] = __ldg(p.fw_start + c + 1);
          if (cur[n] < end[n]) n++;
        }
  }
  __device__ uint32_t next(const DevParams& p) {  // MCX_NONE when exhausted
    if (whole) return k < k1 ? p.spw_list[k++] : MCX_NONE;
    uint32_t best = MCX_NONE;
    for (int i = 0; i < n; i++) {
      while (cur[i] < end[i] && have_last && __ldg(p.fw_list + cur[i]) <= last) cur[i]++;
      if (cur[i] < end[i]) { const uint32_t w = __ldg(p.fw_list + cur[i]); if (w < best) best = w; }
    }
    if (best != MCX_NONE) { last = best; have_last = true; }
    return best;
  }
};

// A group of G lanes (G = 1: one lane, or an aligned group of 8 lanes of a warp) that evaluates ONE molecule together: the
// lanes run the same evaluation on the same inputs — same control flow, same random stream — and split the two loops
// that dominate it, the wall list of get_closest_wall_collision and the candidate records of the partner scan.
// One molecule per lane spent its time in chains of dependent loads, one wall or candidate after the other, with 3 of
// 32 lanes active (profiles/r01_x); the group issues 8 of those loads at once.
struct Group {
  unsigned int mask;  // lanes of the group in the warp
  int G, sub, base;   // lanes per group (1, 2, 4, 8, 16 or 32), this lane's position in the group, first lane of the group
};
__device__ __forceinline__ Group group_of(int G) {
  Group g;
  const int lane = threadIdx.x & 31;
  g.G = G;
  g.base = lane & ~(G - 1);
  g.sub = lane - g.base;
  g.mask = (G >= 32 ? 0xFFFFFFFFu : ((1u << G) - 1u)) << g.base;
  return g;
}

// first stage of collide_wall (collision_utils.inl:664-683): the move stays on one side of the wall's plane — COLLIDE_MISS
// without a random draw
__device__ __forceinline__ bool wall_plane_rejected(const DevParams& p, uint32_t wi, D3 pos, D3 move) {
  const DevWall& f = p.walls[wi];
  const D3 n = {f.nx, f.ny, f.nz};
  const double dp = dot3(n, pos), dv = dot3(n, move), dd = dp - f.dist;
  double d_eps;
  if (dd > 0) {
    d_eps = MCX_EPS;
    if (dd < d_eps) d_eps = 0.5 * dd;
    return dd + dv > d_eps;
  }
  d_eps = -MCX_EPS;
  if (dd > d_eps) d_eps = 0.5 * dd;
  return dd < 0 && dd + dv < d_eps;
}

struct WallHit { int side; double t; D3 pos; uint32_t wall; };
#define MCX_WALL_CAND_MAX 12

// get_closest_wall_collision, collision_utils.inl:819-914.  The walk leaves out walls of the list whose bounding box
// the move cannot reach (they would be COLLIDE_MISS without a random draw); ray_polygon_tests still counts the whole
// list like the reference does, from the list positions.
// G > 1: the lanes of the group first deal the list entries among themselves and drop the walls whose plane the move
// does not reach (no side effects, so neither the order nor duplicates matter); what is left — a handful of walls — goes
// through collide_wall in ascending wall order on every lane, exactly like the sequential walk (same draws, same REDOs).
__device__ bool closest_wall_collision(const DevParams& p, D3 pos, uint32_t subpart, uint32_t last_hit_wall,
                                       Stream& rs, D3& disp, D3& up_to_wall, WallHit& best, LocalStats& ls, Tracer& tc,
                                       const Group& grp) {
  const uint32_t w0 = p.spw_start[subpart], w1 = p.spw_start[subpart + 1];
  if (w0 == w1) return false;
  const bool last_in_list = last_hit_wall != MCX_NONE && spw_position(p, w0, w1, last_hit_wall) != MCX_NONE;
  int guard = 0;
restart:
  bool found = false;
  double closest = MCX_TIME_FOREVER;
  D3 lo, hi;
  segment_box(pos, disp, MCX_FW_MARGIN, lo, hi);
  WallWalk ww;
  ww.init(p, subpart, lo, hi);
  uint32_t cand[MCX_WALL_CAND_MAX];
  int nc = 0, ic = 0;
  const int G = grp.G;
  bool sequential = G == 1;
  if (G > 1) {
    // plane pre-pass over the entries of the walk's lists
    const int n_ranges = ww.whole ? 1 : ww.n;
    for (int r = 0; r < n_ranges && !sequential; r++) {
      const uint32_t e0 = ww.whole ? ww.k : ww.cur[r], e1 = ww.whole ? ww.k1 : ww.end[r];
      const uint32_t* list = ww.whole ? p.spw_list : p.fw_list;
      for (uint32_t e = e0; e < e1 && !sequential; e += G) {
        const uint32_t mine = e + (uint32_t)grp.sub;
        const uint32_t wi = mine < e1 ? __ldg(list + mine) : MCX_NONE;
        const bool keep = wi != MCX_NONE && wi != last_hit_wall && !wall_plane_rejected(p, wi, pos, disp);
        unsigned int bal = (__ballot_sync(grp.mask, keep) & grp.mask) >> grp.base;
        while (bal) {
          const int l = __ffs(bal) - 1;
          bal &= bal - 1;
          const uint32_t w = __shfl_sync(grp.mask, wi, grp.base + l);
          int at = 0;  // sorted insert, duplicates dropped (a wall can sit in several cells of the fine grid)
          while (at < nc && cand[at] < w) at++;
          if (at < nc && cand[at] == w) continue;
          if (nc == MCX_WALL_CAND_MAX) { sequential = true; break; }
          for (int q = nc; q > at; q--) cand[q] = cand[q - 1];
          cand[at] = w; nc++;
        }
      }
    }
  }
  for (;;) {
    uint32_t wi;
    if (sequential) wi = ww.next(p); else wi = ic < nc ? cand[ic++] : MCX_NONE;
    if (wi == MCX_NONE) break;
    if (wi == last_hit_wall) continue;
    double t; D3 hit;
    int ct = collide_wall(p, pos, wi, rs, disp, t, hit);
    if (ct == W_REDO) {
      // the reference had tested the list up to this wall
      ls.ray_polygon_tests += spw_position(p, w0, w1, wi) - w0 + 1 - ((last_in_list && last_hit_wall < wi) ? 1u : 0u);
      ls.redos++; tc.ev(EV_REDO, wi);
      if (tc.tr) tc.tr->n_redo++;
      if (++guard > 64) return false;
      goto restart;
    } else if (ct != W_MISS) {
      ls.ray_polygon_colls++;
      if (subpart_index(p, hit) != subpart) continue;
      if (t < closest) { found = true; closest = t; best.side = ct; best.t = t; best.pos = hit; best.wall = wi; }
    }
  }
  ls.ray_polygon_tests += w1 - w0 - (last_in_list ? 1u : 0u);
  if (found) up_to_wall = best.pos - pos;
  return found;
}

// test_bimolecular, rxn_utils.inl:336-414 (local_prob_factor == 0)
// RxnClass::get_pathway_index_for_probability: binary_search_double (rxn_utils.inl:301-320) over the cumulative pathway
// probabilities times mult (the local probability factor between two surface molecules, else 1)
__device__ __forceinline__ int pathway_for_probability(const DevParams& p, const DevClass& rc, double prob, double mult) {
  int min_idx = 0, max_idx = (int)rc.n_pathways - 1;
  const DevPathway* A = p.pathways + rc.first_pathway;
  while (max_idx - min_idx > 1) {
    int mid = (max_idx + min_idx) / 2;
    if (prob > A[mid].cum_prob * mult) min_idx = mid; else max_idx = mid;
  }
  return prob > A[min_idx].cum_prob * mult ? max_idx : min_idx;
}
// local_prob_factor > 0 only between two surface molecules (react_2D_all_neighbors: 3 / number of neighbour tiles)
__device__ int test_bimolecular(const DevParams& p, const DevClass& rc, double scaling, Stream& rs, double local_prob_factor = 0) {
  double max_fixed_p = rc.max_fixed_p, prob;
  if (local_prob_factor != 0) max_fixed_p = rc.max_fixed_p * local_prob_factor;
  if (max_fixed_p < scaling) {
    prob = rs.dbl() * scaling;
    if (prob >= max_fixed_p) return -1;
  } else {
    float max_p = (float)rc.max_fixed_p;  // sic (rxn_utils.inl:369)
    if (local_prob_factor > 0) max_p = (float)((double)max_p * local_prob_factor);  // float *= double
    if (max_p >= scaling) prob = rs.dbl() * max_p;
    else {
      prob = rs.dbl() * scaling;
      if (prob >= max_p) return -1;
    }
  }
  return pathway_for_probability(p, rc, prob, local_prob_factor > 0 ? local_prob_factor : 1.0);
}

// test_intersect, rxn_utils.inl:593-626 (a Standard reaction with a reactive surface): -1 = no reaction
__device__ int test_intersect(const DevParams& p, const DevClass& rc, double scaling, Stream& rs) {
  const double max_prob = rc.max_fixed_p;
  if (max_prob > scaling) (void)(rs.dbl() * max_prob);
  else { const double pr = rs.dbl() * scaling; if (pr > max_prob) return -1; }
  const double match = rs.dbl() * max_prob;
  int min_idx = 0, max_idx = (int)rc.n_pathways - 1;
  const DevPathway* A = p.pathways + rc.first_pathway;
  while (max_idx - min_idx > 1) {
    int mid = (max_idx + min_idx) / 2;
    if (match > A[mid].cum_prob) min_idx = mid; else max_idx = mid;
  }
  return match > A[min_idx].cum_prob ? max_idx : min_idx;
}

// A record is one 32-byte sector: sm_100 moves it with ONE 256-bit access (LDG.E.256 / STG.E.256; two 128-bit
// accesses before) — half the load instructions of the candidate probe and full-sector stores in the scatter.
__device__ __forceinline__ MolRec rec_from(double x, double y, double z, double w) {
  MolRec r; r.x = x; r.y = y; r.z = z;
  const unsigned long long m = (unsigned long long)__double_as_longlong(w);
  r.id = (uint32_t)m; r.sf = (uint32_t)(m >> 32);
  return r;
}
__device__ __forceinline__ MolRec load_rec(const MolRec* a, uint32_t i) {
  double x, y, z, w;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(a + i));
  return rec_from(x, y, z, w);
}
// snapshot records may receive DEAD flags while retry kernels run: read through L2, not the nc path
__device__ __forceinline__ MolRec load_rec_volatile(const MolRec* a, uint32_t i) {
  double x, y, z, w;
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(a + i) : "memory");
  return rec_from(x, y, z, w);
}
__device__ __forceinline__ void store_rec(MolRec* a, uint32_t i, D3 pos, uint32_t id, uint32_t sf) {
  const unsigned long long m = ((unsigned long long)sf << 32) | id;
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(a + i), "d"(pos.x), "d"(pos.y), "d"(pos.z),
               "d"(__longlong_as_double((long long)m)) : "memory");
}

// Cell range overlapped by the swept volume of a move (segment inflated by R, padded against rounding).
struct CellBox { int cx0, cx1, cy0, cy1, cz0, cz1; };
__device__ __forceinline__ CellBox swept_cells(const DevParams& p, D3 pos, D3 disp) {
  const double pad = p.R * (1.0 + 1e-9) + 1e-9;
  CellBox b;
  b.cx0 = cell_coord(fmin(pos.x, pos.x + disp.x) - pad, p.cgx, p.cell_rcp_x, p.ncx);
  b.cx1 = cell_coord(fmax(pos.x, pos.x + disp.x) + pad, p.cgx, p.cell_rcp_x, p.ncx);
  b.cy0 = cell_coord(fmin(pos.y, pos.y + disp.y) - pad, p.cgy, p.cell_rcp_y, p.ncy);
  b.cy1 = cell_coord(fmax(pos.y, pos.y + disp.y) + pad, p.cgy, p.cell_rcp_y, p.ncy);
  b.cz0 = cell_z(p, fmin(pos.z, pos.z + disp.z) - pad);
  b.cz1 = cell_z(p, fmax(pos.z, pos.z + disp.z) + pad);
  return b;
}

// Flattened walk over the candidate records of a CellBox: cells of one x-row are contiguous in the sorted
// snapshot, so a row is ONE [j, jend) range.  A single loop with a uniform body (advance-row step, then test
// step) keeps the lanes of a warp converged even though their row/candidate counts differ; the nested
// cz/cy/j loops this replaces serialised the lanes of different cells (profiles/r01_a_*: 5.5 of 32 lanes active).
template <bool SKIP_2X2>
struct CandWalkT {
  int ny, nrows, row, ry, rz, cx0, cx1, cy0, cz0;
  uint32_t j, jend;
  // enabled == false gives an empty walk (lanes that have nothing to probe stay in lock step with the warp)
  __device__ __forceinline__ void init(const CellBox& b, bool enabled = true) {
    ny = b.cy1 - b.cy0 + 1; nrows = enabled ? ny * (b.cz1 - b.cz0 + 1) : 0; row = 0; ry = 0; rz = 0;
    cx0 = b.cx0; cx1 = b.cx1; cy0 = b.cy0; cz0 = b.cz0; j = 0; jend = 0;
  }
  __device__ __forceinline__ void next_row() { row++; if (++ry == ny) { ry = 0; rz++; } }
  // returns false when exhausted; on true, `has` tells whether slot j is a candidate to test this round
  __device__ __forceinline__ bool step(const DevParams& p, bool& has, uint32_t& slot) {
    if (j >= jend) {
      if (SKIP_2X2) { while (row < nrows && ry < 2 && rz < 2) next_row(); }  // rows the 2x2 probe already covered
      if (row >= nrows) return false;
      const uint32_t base = row_base(p, cy0 + ry, cz0 + rz);
      j = __ldg(p.cs_cur + base + cx0);
      jend = __ldg(p.cs_cur + base + cx1 + 1);
      next_row();
    }
    has = j < jend;
    slot = j;
    if (has) j++;
    return true;
  }
  // up to two candidates of the current row per trip (slot1 == slot0 when only one is left)
  __device__ __forceinline__ bool step2(const DevParams& p, bool& has0, uint32_t& slot0, bool& has1, uint32_t& slot1) {
    if (j >= jend) {
      if (row >= nrows) return false;
      const uint32_t base = row_base(p, cy0 + ry, cz0 + rz);
      j = __ldg(p.cs_cur + base + cx0);
      jend = __ldg(p.cs_cur + base + cx1 + 1);
      next_row();
    }
    has0 = j < jend;
    has1 = j + 1 < jend;
    slot0 = j;
    slot1 = has1 ? j + 1 : j;
    j += has1 ? 2u : (has0 ? 1u : 0u);
    return true;
  }
};
typedef CandWalkT<false> CandWalk;

__device__ __forceinline__ bool collide_mol_hit(const MolRec& c, D3 pos, D3 disp, double movelen2, double rhs, uint32_t self_id,
                                                double& d_out) {
  D3 dir = {c.x - pos.x, c.y - pos.y, c.z - pos.z};
  double d = dot3(dir, disp);
  double dirlen2 = dot3(dir, dir);
  d_out = d;
  return !(d < 0) && !(d > movelen2) && !(movelen2 * dirlen2 - d * d > rhs) && c.id != self_id && !(c.sf & DF_DEAD);
}
// Partner scan: one pass over the neighbour cells overlapped by the swept volume (segment inflated by R).
// Finds the earliest eligible collision strictly after (t_last, id_last) in (time asc, id desc) order and
// strictly before t_limit, counting how many eligible collisions remain.
struct PartnerHit { double t; uint32_t slot, id, species; int rxn_class; bool in_own_subpart; };

template <bool VOLATILE_SNAPSHOT>
__device__ int scan_partners(const DevParams& p, D3 pos, D3 disp, uint32_t self_id, uint32_t self_species,
                             const SpSet& spm, bool need_sp_filter, double t_last, uint32_t id_last, double t_limit,
                             PartnerHit& best, const Group& grp) {
  const double movelen2 = dot3(disp, disp);
  const double rhs = movelen2 * (p.R * p.R);
  const int* bimol_row = p.bimol + self_species * p.n_species;
  int count = 0;
  best.t = MCX_TIME_FOREVER; best.id = 0; best.slot = MCX_NONE; best.species = 0; best.rxn_class = 0;
  // one candidate: collide_mol (collision_utils.inl:464-515) + eligibility + (time asc, id desc) selection
  auto consider = [&](const MolRec& c, uint32_t j) {
    double d;
    if (!collide_mol_hit(c, pos, disp, movelen2, rhs, self_id, d)) return;
    const uint32_t csp = c.sf & SF_SPECIES_MASK;
    const int rc = bimol_row[csp];
    if (rc < 0) return;
    if (need_sp_filter && !spset_has(spm, subpart_index(p, D3{c.x, c.y, c.z}))) return;
    const double t = d / movelen2;
    if (!(t < t_limit)) return;
    // strictly after (t_last, id_last): later time, or same time and smaller id
    if (t < t_last || (t == t_last && c.id >= id_last)) return;
    count++;
    if (t < best.t || (t == best.t && c.id > best.id)) {
      best.t = t; best.id = c.id; best.slot = j; best.species = csp; best.rxn_class = rc;
    }
  };
  const int G = grp.G;
  if (G == 1) {
    CandWalk cw; cw.init(swept_cells(p, pos, disp));
    bool has0, has1; uint32_t j0, j1;
    // two candidates per trip, both record loads issued before any arithmetic (the walk is latency bound)
    while (cw.step2(p, has0, j0, has1, j1)) {
      if (!has0) continue;
      const MolRec c0 = VOLATILE_SNAPSHOT ? load_rec_volatile(p.recA, j0) : load_rec(p.recA, j0);
      const MolRec c1 = VOLATILE_SNAPSHOT ? load_rec_volatile(p.recA, j1) : load_rec(p.recA, j1);
      consider(c0, j0);
      if (has1) consider(c1, j1);
    }
    return count;
  }
  // the group deals the records of every cell row among its lanes (the order of the tests does not matter: the
  // selection is a minimum over (time, -id) and the count a sum), then combines
  const CellBox b = swept_cells(p, pos, disp);
  for (int cz = b.cz0; cz <= b.cz1; cz++)
    for (int cy = b.cy0; cy <= b.cy1; cy++) {
      const uint32_t base = row_base(p, cy, cz);
      const uint32_t j0 = __ldg(p.cs_cur + base + b.cx0), j1 = __ldg(p.cs_cur + base + b.cx1 + 1);
      for (uint32_t j = j0 + (uint32_t)grp.sub; j < j1; j += G) {
        const MolRec c = VOLATILE_SNAPSHOT ? load_rec_volatile(p.recA, j) : load_rec(p.recA, j);
        consider(c, j);
      }
    }
  for (int o = 1; o < G; o <<= 1) {
    count += __shfl_xor_sync(grp.mask, count, o);
    const double ot = __shfl_xor_sync(grp.mask, best.t, o);
    const uint32_t oid = __shfl_xor_sync(grp.mask, best.id, o), oslot = __shfl_xor_sync(grp.mask, best.slot, o);
    const uint32_t osp = __shfl_xor_sync(grp.mask, best.species, o);
    const int orc = __shfl_xor_sync(grp.mask, best.rxn_class, o);
    if (oslot != MCX_NONE && (best.slot == MCX_NONE || ot < best.t || (ot == best.t && oid > best.id))) {
      best.t = ot; best.id = oid; best.slot = oslot; best.species = osp; best.rxn_class = orc;
    }
  }
  return count;
}

// Fast-path probe: every live partner with a reaction that passes collide_mol for this move (no subpartition
// filter, no ordering).  Returns their number and remembers one; 0 => the generic evaluation would find no
// collision either (its candidate set is a subset of this one).
//
// Shape: the swept box overlaps at most 3x2 (or 2x3) cell rows, else `overflow` (generic path).  The six
// row ranges are fetched up front (12 independent loads in flight), then ONE loop runs over the concatenated
// candidates, two per trip with both record loads issued before any arithmetic: no row-advance branch inside
// the loop, half the trips, and the L1/L2 latency of a record overlaps the test of the previous one.
__device__ __forceinline__ int probe_partners(const DevParams& p, bool enabled, D3 pos, D3 disp, uint32_t self_id,
                                              uint32_t self_species, PartnerHit& first, bool& overflow) {
  const double movelen2 = dot3(disp, disp);
  const double rhs = movelen2 * (p.R * p.R);
  const int* bimol_row = p.bimol + self_species * p.n_species;
  const CellBox b = swept_cells(p, pos, disp);
  const int ny = b.cy1 - b.cy0 + 1, nz = b.cz1 - b.cz0 + 1;
  // six row slots: 3 (y) x 2 (z), or 2 (y) x 3 (z) for a box that is tall in z; anything wider -> generic path
  const bool tall = nz > 2;
  overflow = enabled && (ny * nz > 6 || ny > 3 || nz > 3);
  const bool en = enabled && !overflow;
  uint32_t lo[6], cum[6];
  uint32_t total = 0;
#pragma unroll
  for (int r = 0; r < 6; r++) {
    const int ry = tall ? (r & 1) : (r % 3), rz = tall ? (r >> 1) : (r / 3);
    const bool valid = en && ry < ny && rz < nz;
    const uint32_t base = row_base(p, b.cy0 + ry, b.cz0 + rz);
    const uint32_t a = __ldg(p.cs_cur + (valid ? base + b.cx0 : 0u));
    const uint32_t e = __ldg(p.cs_cur + (valid ? base + b.cx1 + 1 : 0u));
    lo[r] = a - total;   // slot of candidate k in row r is lo[r] + k  (k counted over the concatenation)
    total += e - a;
    cum[r] = total;
  }
  int found = 0;
  first.slot = MCX_NONE; first.in_own_subpart = false; first.t = 0; first.id = 0; first.species = 0; first.rxn_class = 0;
  for (uint32_t k = 0; k < total; k += 2) {
    const uint32_t k1 = k + 1 < total ? k + 1 : k;
    const uint32_t oa = k < cum[2] ? (k < cum[0] ? lo[0] : (k < cum[1] ? lo[1] : lo[2])) : (k < cum[3] ? lo[3] : (k < cum[4] ? lo[4] : lo[5]));
    const uint32_t ob = k1 < cum[2] ? (k1 < cum[0] ? lo[0] : (k1 < cum[1] ? lo[1] : lo[2])) : (k1 < cum[3] ? lo[3] : (k1 < cum[4] ? lo[4] : lo[5]));
    const uint32_t ja = oa + k, jb = ob + k1;
    const MolRec ca = load_rec(p.recA, ja);
    const MolRec cb = load_rec(p.recA, jb);
    double da, db;
    const bool ha = collide_mol_hit(ca, pos, disp, movelen2, rhs, self_id, da);
    const bool hb = collide_mol_hit(cb, pos, disp, movelen2, rhs, self_id, db) && k1 != k;
    if (ha) {
      const uint32_t csp = ca.sf & SF_SPECIES_MASK;
      const int rc = bimol_row[csp];
      if (rc >= 0) { first.t = da / movelen2; first.slot = ja; first.id = ca.id; first.species = csp; first.rxn_class = rc; found++; }
    }
    if (hb) {
      const uint32_t csp = cb.sf & SF_SPECIES_MASK;
      const int rc = bimol_row[csp];
      if (rc >= 0) { first.t = db / movelen2; first.slot = jb; first.id = cb.id; first.species = csp; first.rxn_class = rc; found++; }
    }
  }
  if (found == 1) {  // is the partner in the molecule's own subpartition (always a collected one)?
    const MolRec c = load_rec(p.recA, first.slot);
    first.in_own_subpart = subpart_index(p, D3{c.x, c.y, c.z}) == subpart_index(p, pos);
  }
  return found;
}

// Warp-flattened variant of probe_partners (same result).  Per-lane candidate counts differ by 3x between the
// lanes of a warp (box shape x Poisson occupancy): with one lane per molecule the warp ran max-over-lanes trips at
// 8-12 active lanes (profiles/r01_d: 60 % of k_diffuse_fast's instructions).  Here the (molecule, candidate)
// pairs of the warp's 32 molecules are concatenated and dealt to the lanes 32 at a time, so the trip count is
// ceil(sum / 32): every lane tests one candidate of SOME lane's molecule per trip, the owner's move is read from
// shared memory, and the (rare) hits are reported back through shared-memory atomics.  Warp-collective.
#define MCX_FAST_MAX_HITS 4
// Layout: every table is indexed [field][owner lane], so that the 32 lanes of a warp store to 32 consecutive words
// (one wavefront of the shared-memory pipe per store) and a trip, whose 32 pairs belong to 2-3 owners, reads 2-3 words of
// different banks per load (one wavefront).  With [owner][field] rows of 32 bytes the table stores were 8-way bank
// conflicts and the 128-bit loads of the owner's move took 2.9 wavefronts each: 1.3e9 shared-memory wavefronts per launch,
// the largest share of the busiest unit of the kernel (L1TEX data pipe 64 % busy, profiles/r02_y; r3c for the change).
struct __align__(16) WarpProbe {
  double m[8][32];        // per owner lane: pos.x, pos.y, pos.z, movelen2, disp.x, disp.y, disp.z, bits(id | species << 32)
  uint32_t cum[6][32];    // inclusive row ends over the concatenation of the owner's rows
  uint32_t lo[6][32];     // slot of candidate k in row r = lo[r] + k
  uint32_t off[32];       // compacted owners: first pair index
  uint32_t owner[32];     // compacted owners: lane
  uint32_t hits[32];      // per owner lane: number of eligible collisions
  uint32_t hit_slot[32][MCX_FAST_MAX_HITS];  // per owner lane: slots of the first few of them
};
__device__ __forceinline__ int probe_partners_flat(const DevParams& p, bool enabled, D3 pos, D3 disp, uint32_t self_id,
                                                   uint32_t self_species, bool& overflow, WarpProbe* sm) {
  const int lane = threadIdx.x & 31;
  const double movelen2 = dot3(disp, disp);
  const double R2 = p.R * p.R;
  const CellBox bx = swept_cells(p, pos, disp);
  const int ny = bx.cy1 - bx.cy0 + 1, nz = bx.cz1 - bx.cz0 + 1;
  const bool tall = nz > 2;
  overflow = enabled && (ny * nz > 6 || ny > 3 || nz > 3);
  const bool en = enabled && !overflow;
  uint32_t total = 0;
#pragma unroll
  for (int r = 0; r < 6; r++) {
    const int ry = tall ? (r & 1) : (r % 3), rz = tall ? (r >> 1) : (r / 3);
    const bool valid = en && ry < ny && rz < nz;
    const uint32_t base = row_base(p, bx.cy0 + ry, bx.cz0 + rz);
    const uint32_t a = __ldg(p.cs_cur + (valid ? base + bx.cx0 : 0u));
    const uint32_t e = __ldg(p.cs_cur + (valid ? base + bx.cx1 + 1 : 0u));
    sm->lo[r][lane] = a - total;
    total += e - a;
    sm->cum[r][lane] = total;
  }
  sm->m[0][lane] = pos.x; sm->m[1][lane] = pos.y; sm->m[2][lane] = pos.z; sm->m[3][lane] = movelen2;
  sm->m[4][lane] = disp.x; sm->m[5][lane] = disp.y; sm->m[6][lane] = disp.z;
  sm->m[7][lane] = __longlong_as_double((long long)(((unsigned long long)self_species << 32) | self_id));
  sm->hits[lane] = 0;
  // exclusive prefix of the totals; owners with candidates are compacted so that their offsets strictly increase
  uint32_t incl = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  const uint32_t W = __shfl_sync(0xffffffffu, incl, 31);
  const unsigned int ne = __ballot_sync(0xffffffffu, total > 0);
  if (total > 0) { const int k = __popc(ne & ((1u << lane) - 1u)); sm->off[k] = incl - total; sm->owner[k] = lane; }
  __syncwarp();
  const uint32_t my_off = lane < __popc(ne) ? sm->off[lane] : 0xFFFFFFFFu;
  // pair q of the concatenation -> (owner lane o, candidate slot j); warp-collective (ballots), valid = q < W
  auto locate = [&](uint32_t B, uint32_t& o, uint32_t& j) -> bool {
    const uint32_t q = B + lane;
    const unsigned int below = __ballot_sync(0xffffffffu, my_off <= B);
    const unsigned int bit = (my_off > B && my_off - B < 32u) ? (1u << (my_off - B)) : 0u;
    const unsigned int mask = __reduce_or_sync(0xffffffffu, bit);
    const int k = __popc(below) - 1 + __popc(mask & ((2u << lane) - 1u));
    const bool valid = q < W;
    o = 0; j = 0;
    if (valid) {
      o = sm->owner[k];
      const uint32_t kk = q - sm->off[k];
      // row of pair kk = number of row ends at or below it (branch-free: the nested selects this replaces were
      // compiled into divergent branches, 12-20 of 32 lanes active, profiles/r01_l)
      const uint32_t r = (kk >= sm->cum[0][o]) + (kk >= sm->cum[1][o]) + (kk >= sm->cum[2][o]) + (kk >= sm->cum[3][o]) +
                         (kk >= sm->cum[4][o]);
      j = sm->lo[r][o] + kk;
    }
    return valid;
  };
  // collide_mol (collision_utils.inl:464-515) of candidate record c against owner o's move
  auto test = [&](bool valid, uint32_t o, uint32_t j, const MolRec& c) {
    if (!valid) return;
    const double ml2 = sm->m[3][o];
    const unsigned long long ids = (unsigned long long)__double_as_longlong(sm->m[7][o]);
    double d;
    if (collide_mol_hit(c, D3{sm->m[0][o], sm->m[1][o], sm->m[2][o]}, D3{sm->m[4][o], sm->m[5][o], sm->m[6][o]}, ml2, ml2 * R2,
                        (uint32_t)ids, d)) {
      const int rc = p.bimol[(uint32_t)(ids >> 32) * p.n_species + (c.sf & SF_SPECIES_MASK)];
      if (rc >= 0) { const uint32_t h = atomicAdd(&sm->hits[o], 1u); if (h < MCX_FAST_MAX_HITS) sm->hit_slot[o][h] = j; }
    }
  };
#ifdef MCX_EXPERIMENT_NO_PAIRS   // timing experiment only: the probe's set-up without its pair loop
  for (uint32_t B = W; B < W; B += 32) {
#else
  for (uint32_t B = 0; B < W; B += 32) {
#endif
    uint32_t o, j;
    const bool v = locate(B, o, j);
    if (v) { const MolRec c = load_rec(p.recA, j); test(v, o, j, c); }
  }
  __syncwarp();
  const int found = (int)sm->hits[lane];
  __syncwarp();  // the buffers are reused by the next trip of the caller's loop
  return found;    // the slots of the first MCX_FAST_MAX_HITS of them are in sm->hit_slot[lane]
}

// Direction (-1 / 0 / +1 per axis) in which collect_neighboring_subparts (collision_utils_subparts.inl:38-122) adds
// neighbours of the subpartition si around position q: the subpartitions it inserts are si + (a, b, c) with each
// component either 0 or that direction, not all 0.
__device__ __forceinline__ void neighbor_dirs(const DevParams& p, D3 q, const int si[3], double rr, int d[3]) {
  const double sp_len = p.sp_len, part_len = p.part_len;
  const D3 rel = {q.x - p.ox, q.y - p.oy, q.z - p.oz};
  const D3 plus = {rel.x + rr, rel.y + rr, rel.z + rr};
  const D3 minus = {rel.x - rr, rel.y - rr, rel.z - rr};
  const D3 boundary = {si[0] * sp_len, si[1] * sp_len, si[2] * sp_len};
  d[0] = (minus.x < boundary.x && minus.x > 0.0) ? -1 : ((plus.x > boundary.x + sp_len && plus.x < part_len) ? 1 : 0);
  d[1] = (minus.y < boundary.y && minus.y > 0.0) ? -1 : ((plus.y > boundary.y + sp_len && plus.y < part_len) ? 1 : 0);
  d[2] = (minus.z < boundary.z && minus.z > 0.0) ? -1 : ((plus.z > boundary.z + sp_len && plus.z < part_len) ? 1 : 0);
}
__device__ __forceinline__ bool in_neighbor_dirs(const int delta[3], const int d[3]) {
  return (delta[0] == 0 || delta[0] == d[0]) && (delta[1] == 0 || delta[1] == d[1]) && (delta[2] == 0 || delta[2] == d[2]);
}

// Next collision of the probe's hit list in sort_collisions_by_time order (time ascending, vol-vol ties by
// descending partner id, diffuse_react_event.cpp:341-364) strictly after (t_last, id_last); false when none is left.
// A hit counts only if its subpartition is one the reference collects for this move (ray_trace_vol :698-722):
// the molecule's own one and, when the move crosses a single subpartition face, the one it ends in (s1); for a
// move that stays inside its subpartition (stays == true) also the neighbours collect_neighboring_subparts adds
// around the start and the end point (DECIDE_FOREIGN; the expanded list is in use whenever volume-volume
// reactions exist).  decided = false: a hit in a foreign subpartition whose
// membership this function does not work out (the caller hands the molecule to the generic pass).
template <bool DECIDE_FOREIGN>
// hit_slots[h * hit_stride], h < n_hits: the snapshot slots the probe reported for this molecule.
__device__ __forceinline__ bool next_probe_hit(const DevParams& p, const uint32_t* hit_slots, int hit_stride, int n_hits, D3 pos,
                                               D3 disp, bool stays, bool single, const int s1[3], uint32_t self_species,
                                               double t_last, uint32_t id_last, PartnerHit& best, bool& decided) {
  const double movelen2 = dot3(disp, disp);
  int s0[3];
  subpart_3d(p, pos, s0);
  bool any = false;
  best.t = MCX_TIME_FOREVER; best.id = 0; best.slot = MCX_NONE; best.species = 0; best.rxn_class = 0; best.in_own_subpart = true;
  for (int h = 0; h < n_hits; h++) {
    const uint32_t j = hit_slots[h * hit_stride];
    const MolRec c = load_rec(p.recA, j);
    const D3 dir = {c.x - pos.x, c.y - pos.y, c.z - pos.z};
    const double t = dot3(dir, disp) / movelen2;
    int sc[3];
    subpart_3d(p, D3{c.x, c.y, c.z}, sc);
    const int delta[3] = {sc[0] - s0[0], sc[1] - s0[1], sc[2] - s0[2]};
    // own subpartition, or the one a single face crossing leads into (s1): crossed subpartitions are collected
    if ((delta[0] | delta[1] | delta[2]) != 0 && !(single && sc[0] == s1[0] && sc[1] == s1[1] && sc[2] == s1[2])) {
      if (!DECIDE_FOREIGN || !stays) { decided = false; continue; }
      bool member = false;
      if (p.use_expanded) {
        const double rr = p.R * MCX_POS_SQRT2;
        int d[3];
        neighbor_dirs(p, pos, s0, rr, d);
        member = in_neighbor_dirs(delta, d);
        if (!member) { neighbor_dirs(p, pos + disp, s0, rr, d); member = in_neighbor_dirs(delta, d); }
      }
      if (!member) continue;  // not a molecule of a collected subpartition: the reference does not see it
    }
    if (t < t_last || (t == t_last && c.id >= id_last)) continue;
    if (!any || t < best.t || (t == best.t && c.id > best.id)) {
      const uint32_t csp = c.sf & SF_SPECIES_MASK;
      best.t = t; best.id = c.id; best.slot = j; best.species = csp; best.rxn_class = p.bimol[self_species * p.n_species + csp];
      any = true;
    }
  }
  return any;
}

// Plane-rejection stage of collide_wall (collision_utils.inl:664-683) for every wall of one subpartition:
// true when each wall is a COLLIDE_MISS already there (no random draw, no REDO can occur).
// min_dist: smallest distance of the segment's end points to any of the planes; for a rejected wall the whole
// segment stays on one side, so no point of it is closer (used to decide whether exact_disk can see a wall).
__device__ __forceinline__ bool all_walls_plane_rejected(const DevParams& p, bool enabled, uint32_t subpart, D3 pos, D3 move,
                                                         unsigned int& n_tests, double& min_dist) {
  const uint32_t w0 = __ldg(p.spw_start + subpart), w1 = enabled ? __ldg(p.spw_start + subpart + 1) : w0;
  n_tests = w1 - w0;  // what the reference's scan of the list counts
  bool all = true;
  if (p.fw_K == 1) {  // short lists (warp-uniform branch): one flat predicated loop keeps the warp converged
#pragma unroll 1
    for (uint32_t k = w0; k < w1; k++) {
      const DevWall& f = p.walls[__ldg(p.spw_list + k)];
      const D3 n = {f.nx, f.ny, f.nz};
      const double dp = dot3(n, pos), dv = dot3(n, move), dd = dp - f.dist;
      double d_eps;
      bool miss;
      if (dd > 0) {
        d_eps = MCX_EPS;
        if (dd < d_eps) d_eps = 0.5 * dd;
        miss = dd + dv > d_eps;
      } else {
        d_eps = -MCX_EPS;
        if (dd > d_eps) d_eps = 0.5 * dd;
        miss = dd < 0 && dd + dv < d_eps;
      }
      all = all && miss;
      min_dist = fmin(min_dist, fmin(fabs(dd), fabs(dd + dv)));
    }
  } else if (w1 > w0) {
    // only walls whose bounding box comes within the interaction radius of the move can be hit or can cut the
    // interaction disk of a collision on it (exact_disk's own box test, exact_disk_utils.inl:905-925); the test has
    // no side effect, so the duplicates of walls that span several cells are harmless
    D3 lo, hi;
    segment_box(pos, move, p.R * (1.0 + 1e-9) + MCX_FW_MARGIN, lo, hi);
    const FwRange r = fw_range(p, subpart, lo, hi);
    const int K = p.fw_K;
    for (int z = r.z0; z <= r.z1; z++)
      for (int y = r.y0; y <= r.y1; y++) {
        const uint32_t row = r.base + (uint32_t)((z * K + y) * K);  // the cells of one x-run are consecutive in the table
        const uint32_t e0 = __ldg(p.fw_start + row + r.x0), e1 = __ldg(p.fw_start + row + r.x1 + 1);
#pragma unroll 1
        for (uint32_t e = e0; e < e1; e++) {
          const DevWall& f = p.walls[__ldg(p.fw_list + e)];
          const D3 n = {f.nx, f.ny, f.nz};
          const double dp = dot3(n, pos), dv = dot3(n, move), dd = dp - f.dist;
          double d_eps;
          bool miss;
          if (dd > 0) {
            d_eps = MCX_EPS;
            if (dd < d_eps) d_eps = 0.5 * dd;
            miss = dd + dv > d_eps;
          } else {
            d_eps = -MCX_EPS;
            if (dd > d_eps) d_eps = 0.5 * dd;
            miss = dd < 0 && dd + dv < d_eps;
          }
          all = all && miss;
          min_dist = fmin(min_dist, fmin(fabs(dd), fabs(dd + dv)));
        }
      }
  }
  return all;
}

// ---- surface molecules ------------------------------------------------------------------------------------
// distinguishable_vec3, src4/defines.h:766-806
__device__ __forceinline__ bool distinguishable_vec3_d(D3 a, D3 b, double eps) {
  double c = fabs(a.x), cc, d;
  d = fabs(a.y); if (d > c) c = d;
  d = fabs(a.z); if (d > c) c = d;
  d = fabs(b.x); if (d > c) c = d;
  d = fabs(b.y); if (d > c) c = d;
  d = fabs(b.z); if (d > c) c = d;
  cc = fabs(a.x - b.x);
  d = fabs(a.y - b.y); if (d > cc) cc = d;
  d = fabs(a.z - b.z); if (d > cc) cc = d;
  if (c < eps) c = eps;
  return c * eps < cc;
}
// GridUtils::xyz2grid_tile_index, src4/grid_utils.inl:48-118
__device__ uint32_t xyz2grid(const DevParams& p, D3 v, uint32_t wi) {
  const DevGrid& g = p.grids[wi];
  const DevWall& f = p.walls[wi];
  const uint32_t n_tiles = (uint32_t)(g.n_axis * g.n_axis);
  if (n_tiles == 1) return 0;
  if (!distinguishable_vec3_d(v, wall_vertex(p, wi, 0), MCX_EPS)) return n_tiles - 2 * (uint32_t)g.n_axis + 1;
  if (!distinguishable_vec3_d(v, wall_vertex(p, wi, 1), MCX_EPS)) return n_tiles - 1;
  if (!distinguishable_vec3_d(v, wall_vertex(p, wi, 2), MCX_EPS)) return 0;
  const double i = dot3(v, D3{f.ux, f.uy, f.uz}) - g.vert0_u;
  const double j = dot3(v, D3{f.vx, f.vy, f.vz}) - g.vert0_v;
  const double striploc = j * g.strip_width_rcp;
  int strip = (int)striploc;
  const double striprem = striploc - strip;
  strip = g.n_axis - strip - 1;
  const double u0 = j * g.vert2_slope;
  const double u1_u0 = f.uv1u - j * g.fullslope;
  const double stripeloc = ((i - u0) / u1_u0) * (strip + (1 - striprem));
  const int stripe = (int)stripeloc;
  const double striperem = stripeloc - stripe;
  const int flip = (striperem < 1 - striprem) ? 0 : 1;
  int idx = strip * strip + 2 * stripe + flip;
  if (idx < 0) idx = 0;
  if ((uint32_t)idx >= n_tiles) idx = (int)n_tiles - 1;
  return (uint32_t)idx;
}
// RxnUtils::trigger_bimolecular orientation test, src4/rxn_utils.inl:58-84
__device__ __forceinline__ bool orientations_match(const DevClass& rc, int orientA, int orientB) {
  const int geomA = rc.geom0, geomB = rc.geom1;
  if (geomA == 0 || geomB == 0 || (geomA + geomB) * (geomA - geomB) != 0) return true;
  return orientA != 0 && orientA * orientB * geomA * geomB > 0;
}
// one random bit per product with rule orientation NONE (outcome_products_random, diffuse_react_event.cpp:2618-2627)
// With kept_info (include/mcx.h) the draws follow the order of the rule's products, kept reactants included.
#define ORIENT_BIT_FRONT 64u
#define ORIENT_BITS_MASK 127u
__device__ __forceinline__ int kept_code(const DevPathway& pw, int r) {
  const uint32_t c = (pw.kept_info >> (24 + 2 * r)) & 3u;
  return c == 1u ? 1 : (c == 2u ? -1 : 0);
}
__device__ __forceinline__ uint32_t draw_orientation_bits(const DevPathway& pw, Stream& rs) {
  uint32_t bits = 0;
  if (!(pw.kept_info & MCX_KEPT_VALID)) {
    for (uint32_t k = 0; k < pw.n_products; k++)
      if (pw.prod_orient[k] == 0 && (rs.next() & 1u)) bits |= 1u << k;
    return bits;
  }
  for (int q = 0; q < 6; q++) {
    const uint32_t nib = (pw.kept_info >> (4 * q)) & 0xFu;
    if (nib == MCX_KEPT_ORDER_END) break;
    if (nib >= MCX_KEPT_ORDER_REACTANT) {
      const int r = (int)(nib & 1u);
      if (kept_code(pw, r) == 0 && (rs.next() & 1u)) bits |= 1u << (4 + r);
    } else if (nib < pw.n_products && pw.prod_orient[nib] == 0 && (rs.next() & 1u)) bits |= 1u << nib;
  }
  return bits;
}
// product-side orientation of kept reactant r (outcome_products_random :2618-2652); 0: the table does not say
__device__ __forceinline__ int kept_orientation(const DevClass& c, const DevPathway& pw, int r, uint32_t orient_bits, int surf_orient) {
  if (!(pw.kept_info & MCX_KEPT_VALID)) return 0;
  int o = kept_code(pw, r);
  if (o == 0) return ((orient_bits >> (4 + r)) & 1u) ? 1 : -1;
  if (c.kind == MCX_RXN_BIMOL_VOLSURF && c.geom1 != 0 && surf_orient != c.geom1) o = -o;
  return o;
}

// ---- surface-surface reactions (react_2D_all_neighbors, diffuse_react_event.cpp:1250-1393) --------------------------
#define SURFSURF_SWAP 64u   // bit 6 of the orientation bits: the first surface product takes the SECOND freed tile
#define SURFSURF_MAX_MATCHES 32
// find_surf_product_positions (:1993-2288) over the tiles the consumed reactants free (in the order of the rule's
// reactants): the one surface product of a pathway takes the initiator's tile when the initiator is consumed, else
// the one freed tile (:2140-2154); two surface products draw for the two tiles like the reference (:2155-2191)
__device__ __forceinline__ uint32_t surfsurf_position_bits(const DevParams& p, const DevPathway& pw, bool init_is_r0, Stream& rs) {
  const int keep0 = pw.keep_mask & 1u, keep1 = (pw.keep_mask >> 1) & 1u;
  int needed = 0; uint32_t first_surf = MCX_NONE;
  for (uint32_t k = 0; k < pw.n_products; k++)
    if (!(p.species[pw.products[k]].flags & MCX_SP_VOL)) { needed++; if (first_surf == MCX_NONE) first_surf = k; }
  if (needed == 0) return 0;
  const int freed = 2 - keep0 - keep1, actual = (int)pw.n_products + keep0 + keep1;
  const int to_recycle = actual < freed ? actual : freed;
  if (needed == 1 && to_recycle == 1) {
    const int ri = init_is_r0 ? 0 : 1;
    const bool init_consumed = ri == 0 ? !keep0 : !keep1;
    return (init_consumed && ri == 1 && !keep0) ? SURFSURF_SWAP : 0u;
  }
  uint32_t bits = 0, assigned = 0;
  int next_available = 0;
  const uint32_t num_players = (uint32_t)actual + 2u;
  for (int guard = 0; next_available < to_recycle && guard < 100000; guard++) {
    const uint32_t rnd = rs.next() % num_players;
    if (rnd < 2) continue;
    const uint32_t k = rnd - 2;
    if (k >= pw.n_products || (p.species[pw.products[k]].flags & MCX_SP_VOL)) continue;
    if ((assigned >> k) & 1u) continue;
    assigned |= 1u << k;
    if ((next_available == 1) == (k == first_surf)) bits |= SURFSURF_SWAP;
    next_available++;
  }
  return bits;
}
// :2640-2652: a rule orientation flips once for every surface reactant that lies the other way round than the rule states
__device__ __forceinline__ int surfsurf_match(const DevClass& c, int orient_r0, int orient_r1) {
  int m = 1;
  if (c.geom0 != 0 && orient_r0 != c.geom0) m = -m;
  if (c.geom1 != 0 && orient_r1 != c.geom1) m = -m;
  return m;
}

// ---- surface diffusion -------------------------------------------------------------------------------------------
// distinguishable_vec2, src4/defines.h:733-764
__device__ __forceinline__ bool distinguishable_vec2_d(double au, double av, double bu, double bv, double eps) {
  double c = fabs(au), cc, d;
  d = fabs(av); if (d > c) c = d;
  d = fabs(bu); if (d > c) c = d;
  d = fabs(bv); if (d > c) c = d;
  cc = fabs(au - bu);
  d = fabs(av - bv); if (d > cc) cc = d;
  if (c < eps) c = eps;
  return c * eps < cc;
}
// GridUtils::uv2grid_tile_index, src4/grid_utils.inl:119-190 (MCX_NONE where the reference raises an internal error)
__device__ uint32_t uv2grid(const DevParams& p, uint32_t wi, double u, double v) {
  const DevGrid& g = p.grids[wi];
  const DevWall& f = p.walls[wi];
  const uint32_t n_tiles = (uint32_t)(g.n_axis * g.n_axis);
  if (n_tiles == 1) return 0;
  if (!distinguishable_vec2_d(u, v, 0, 0, MCX_EPS)) return n_tiles - 2 * (uint32_t)g.n_axis + 1;
  if (!distinguishable_vec2_d(u, v, f.uv1u, 0, MCX_EPS)) return 0;
  if (!distinguishable_vec2_d(u, v, f.uv2u, f.uv2v, MCX_EPS)) return n_tiles - 1;
  const double striploc = v * g.strip_width_rcp;
  int strip = (int)striploc;
  const double striprem = striploc - strip;
  strip = g.n_axis - strip - 1;
  const double u0 = v * g.vert2_slope;
  const double u1_u0 = f.uv1u - v * g.fullslope;
  const double stripeloc = ((u - u0) / u1_u0) * (strip + (1 - striprem));
  const int stripe = (int)stripeloc;
  const double striperem = stripeloc - stripe;
  const int flip = (striperem < 1 - striprem) ? 0 : 1;
  const int idx = strip * strip + 2 * stripe + flip;
  if (idx < 0 || (uint32_t)idx >= n_tiles) return MCX_NONE;
  return (uint32_t)idx;
}
// GeometryUtils::find_edge_point, src4/geometry_utils.inl:222-291.  0,1,2: side hit; 3: stays within the wall; 4: cannot tell
__device__ int find_edge_point(const DevWall& here, double lu, double lv, double du, double dv, double& eu, double& ev) {
  const double lxd = lu * dv - lv * du;
  const double lxc1 = -lv * here.uv1u;
  const double dxc1 = -dv * here.uv1u;
  double f, s, t;
  if (dxc1 < -MCX_EPS || dxc1 > MCX_EPS) {
    f = 1 / dxc1;
    s = -lxd * f;
    if (0 < s && s < 1 && f > 0) {
      t = -lxc1 * f;
      if (MCX_EPS < t && t < 1) { eu = lu + t * du; ev = lv + t * dv; return 0; }
      else if (t > 1 + MCX_EPS) return 3;
    }
  }
  const double lxc2 = lu * here.uv2v - lv * here.uv2u;
  const double dxc2 = du * here.uv2v - dv * here.uv2u;
  if (dxc2 < -MCX_EPS || dxc2 > MCX_EPS) {
    f = 1 / dxc2;
    s = 1 + lxd * f;
    if (0 < s && s < 1 && f < 0) {
      t = -lxc2 * f;
      if (MCX_EPS < t && t < 1) { eu = lu + t * du; ev = lv + t * dv; return 2; }
      else if (t > 1 + MCX_EPS) return 3;
    }
  }
  f = dxc2 - dxc1;
  if (f < -MCX_EPS || f > MCX_EPS) {
    f = 1 / f;
    s = -(lxd + dxc1) * f;
    if (0 < s && s < 1 && f > 0) {
      t = (here.uv1u * here.uv2v + lxc1 - lxc2) * f;
      if (MCX_EPS < t && t < 1) { eu = lu + t * du; ev = lv + t * dv; return 1; }
      else if (t > 1 + MCX_EPS) return 3;
    }
  }
  return 4;
}
// GeometryUtils::traverse_surface, src4/geometry_utils.inl:305-342
__device__ __forceinline__ uint32_t traverse_surface(const DevParams& p, uint32_t wi, double lu, double lv, int which,
                                                     double& nu, double& nv) {
  const DevEdge e = p.edges[3 * wi + which];
  if (e.nb_wall == MCX_NONE) return MCX_NONE;
  if (e.forward) {
    const double ru = e.cos_t * lu + e.sin_t * lv, rv = -e.sin_t * lu + e.cos_t * lv;
    nu = ru + e.tu; nv = rv + e.tv;
  } else {
    const double ru = lu - e.tu, rv = lv - e.tv;
    nu = e.cos_t * ru - e.sin_t * rv;
    nv = e.sin_t * ru + e.cos_t * rv;
  }
  return e.nb_wall;
}
// ray_trace_surf, src4/diffuse_react_event.cpp:1578-1725 (no region borders).  Returns the wall the move ends on
// (MCX_NONE: ambiguous side hit) and the end point in that wall's frame.
// Region borders (species.can_interact_with_border(), :1627-1665; reflect_absorb_inside_out / outside_in, diffusion_utils.inl:
// 598-700): an edge that is a border of a reactive region — of the wall the molecule leaves or of the one it enters —
// turns the molecule back (REFLECTIVE class), takes it (ABSORPTIVE: absorbed = true, MCX_NONE) or lets it pass.
__device__ uint32_t ray_trace_surf(const DevParams& p, uint32_t wall_index, double pu, double pv, double du, double dv,
                                   double& out_u, double& out_v, uint32_t sm_species, int sm_orient, bool& absorbed) {
  uint32_t this_index = wall_index;
  double this_u = pu, this_v = pv, disp_u = du, disp_v = dv;
  absorbed = false;
  const bool borders = p.wall_border != nullptr;
  const uint32_t act_base = sm_species * (uint32_t)p.n_surf_classes;
  const uint32_t act_side = sm_orient > 0 ? 0u : 1u;
  for (int guard = 0; guard < 10000; guard++) {
    const DevWall& tw = p.walls[this_index];
    double bu = 0, bv = 0;
    const int edge = find_edge_point(tw, this_u, this_v, disp_u, disp_v, bu, bv);
    if (edge == 4) return MCX_NONE;
    if (edge == 3) { out_u = this_u + disp_u; out_v = this_v + disp_v; return this_index; }
    const double old_u = this_u, old_v = this_v;
    double nu, nv;
    bool reflect_now = false;
    if (borders && ((p.wall_border[this_index] >> edge) & 1u)) {   // inside out
      const uint32_t wc = p.wall_class[this_index];
      const int act = wc == MCX_NONE ? MCX_SURF_TRANSPARENT : p.surf_border[(act_base + wc) * 2 + act_side];
      if (act == MCX_SURF_ABSORPTIVE) { absorbed = true; return MCX_NONE; }
      reflect_now = act == MCX_SURF_REFLECTIVE;
    }
    uint32_t target = reflect_now ? MCX_NONE : traverse_surface(p, this_index, old_u, old_v, edge, nu, nv);
    if (target != MCX_NONE && borders) {   // outside in: the shared edge in the neighbour's numbering
      int te = -1;
      for (int e2 = 0; e2 < 3; e2++) if (p.edges[3 * target + e2].nb_wall == this_index) te = e2;
      if (te >= 0 && ((p.wall_border[target] >> te) & 1u)) {
        const uint32_t wc = p.wall_class[target];
        const int act = wc == MCX_NONE ? MCX_SURF_TRANSPARENT : p.surf_border[(act_base + wc) * 2 + act_side];
        if (act == MCX_SURF_ABSORPTIVE) { absorbed = true; return MCX_NONE; }
        if (act == MCX_SURF_REFLECTIVE) target = MCX_NONE;
      }
    }
    if (target != MCX_NONE) {
      this_u = nu; this_v = nv;
      double tu2, tv2;
      traverse_surface(p, this_index, old_u + disp_u, old_v + disp_v, edge, tu2, tv2);
      disp_u = tu2 - this_u; disp_v = tv2 - this_v;
      this_index = target;
      continue;
    }
    double ndu = disp_u - (bu - old_u), ndv = disp_v - (bv - old_v);  // free side: reflect
    if (edge == 0) ndv *= -1.0;
    else {
      double ru = edge == 1 ? -tw.uv2v : tw.uv2v, rv = edge == 1 ? tw.uv2u - tw.uv1u : -tw.uv2u;
      double f = 1.0 / sqrt(ru * ru + rv * rv);
      ru *= f; rv *= f;
      f = 2.0 * (ndu * ru + ndv * rv);
      ndu -= f * ru; ndv -= f * rv;
    }
    this_u = bu; this_v = bv; disp_u = ndu; disp_v = ndv;
  }
  return MCX_NONE;
}

#include "mcx_exact_disk.cuh"

// ===================================================================================================
// One molecule, (the rest of) one iteration.  `forced` = last conflict round: no partner search.
// ===================================================================================================
// Control flow note: no early `return` / `goto` — the outcome is carried in `decided` and every loop has a single
// exit condition, so that the lanes of a warp (32 different molecules) re-converge at the loop headers and run
// the common stages (DDA, wall tests, partner scan) together.
// created_wall / created_tile: where a DF_CREATED_ON_SURF volume product was created (MCX_NONE otherwise).
// WITH_DISK == false: the evaluation stops with err = MCX_INTERNAL_NEEDS_DISK at the first collision whose
// interaction disk is cut by a wall; the caller hands the molecule to the WITH_DISK == true instantiation.  Keeping
// exact_disk (3.5 KB of per-thread pool, 300 more bytes of spills) out of the common generic pass saves a third of
// its run time (profiles/r01_g).
#define MCX_INTERNAL_NEEDS_DISK 1000
// SURF == false compiles the surface-molecule code out (models without surface species: the launcher picks the
// instantiation from DevParams::has_surf), which gives the registers back to the volume path.
// grp: the lanes that evaluate this molecule together (Group above); all of them pass identical arguments.
// GridUtils::grid2uv (grid_utils.inl:233-253) and grid2uv_random (:256-288)
__device__ __forceinline__ void tile_uv(const DevParams& p, uint32_t wi, uint32_t tile, bool random, Stream& rs, double& u, double& v) {
  const DevWall& f = p.walls[wi];
  const DevGrid& g = p.grids[wi];
  const int root = (int)(sqrt((double)tile));
  const int rootrem = (int)tile - root * root;
  const int k = g.n_axis - root - 1;
  const int j = rootrem / 2;
  const int i = rootrem - 2 * j;
  if (!random) {
    const double over3n = 1 / (double)(3 * g.n_axis);
    u = ((double)(3 * j + i + 1)) * over3n * f.uv1u + ((double)(3 * k + i + 1)) * over3n * f.uv2u;
    v = ((double)(3 * k + i + 1)) * over3n * f.uv2v;
    return;
  }
  const double over_n = 1 / (double)(g.n_axis);
  const double u_ran = rs.dbl();
  const double v_ran = 1 - sqrt(rs.dbl());
  u = ((double)(j + i) + (1 - 2 * i) * (1 - v_ran) * u_ran) * over_n * f.uv1u + ((double)(k + i) + (1 - 2 * i) * v_ran) * over_n * f.uv2u;
  v = ((double)(k + i) + (1 - 2 * i) * v_ran) * over_n * f.uv2v;
}

// ---- products on vacant neighbour tiles: the general branch of find_surf_product_positions (:2060-2100, 2155-2285) ----
// A pathway that creates more surface products than it frees tiles (DevPathway::general) puts the extra ones on vacant
// tiles around the surface reactant — the row of the neighbour-tile table, every wall counting (create_grid_flag) — at a
// random point of the tile (grid2uv_random).  The reference's bookkeeping is kept literally: positions are assigned per
// entry of the rule's product list (kept reactants and volume products draw a tile too and waste it), the c-th CREATED
// surface product takes the position of the c-th ENTRY (:2815-2818).  Draws in the reference's order: recycled tiles,
// vacant tiles, orientations, random points.  The result goes to the proposal arrays of `slot` (all lanes of a group
// write the same values).  Returns false when the reaction is blocked (RX_BLOCKED).
// rec_*: the sites of the consumed surface reactants in the order of the rule's reactants.
__device__ __noinline__ bool place_general(const DevParams& p, const DevClass& cl, const DevPathway& pw, uint32_t slot,
                                           uint32_t reac_wall, uint32_t reac_tile, int n_rec, const uint32_t* rec_wall,
                                           const uint32_t* rec_tile, const double2* rec_uv, bool forced, unsigned int epoch_first,
                                           unsigned int epoch, bool retry, Stream& rs, uint32_t& orient_bits) {
  // entries of the rule's product list (kept_info nibbles): bit 0 surface, bit 1 kept
  int n_ent = 0; uint8_t ent[6];
  for (int q = 0; q < 6; q++) {
    const uint32_t nib = (pw.kept_info >> (4 * q)) & 0xFu;
    if (nib == MCX_KEPT_ORDER_END) break;
    if (nib >= MCX_KEPT_ORDER_REACTANT) {
      const uint32_t sp = (nib & 1u) ? cl.r1 : cl.r0;
      ent[n_ent++] = (uint8_t)(2u | ((sp < (uint32_t)p.n_species && !(p.species[sp].flags & MCX_SP_VOL)) ? 1u : 0u));
    } else if (nib < pw.n_products) ent[n_ent++] = (uint8_t)((p.species[pw.products[nib]].flags & MCX_SP_VOL) ? 0u : 1u);
  }
  const int n_reactants = cl.kind == MCX_RXN_UNIMOL ? 1 : 2;
  int needed = 0;
  for (uint32_t k = 0; k < pw.n_products; k++) needed += (p.species[pw.products[k]].flags & MCX_SP_VOL) ? 0 : 1;
  // vacant tiles around the surface reactant, from the back of the reference's list (:2090-2098)
  const uint32_t gt0 = p.grids[reac_wall].tile_start + reac_tile;
  const uint32_t qb = __ldg(p.tn_start + gt0), qe = __ldg(p.tn_start + gt0 + 1);
  uint2 vacant[SURFSURF_MAX_MATCHES];
  int n_vac = 0;
  for (uint32_t q = qe; q > qb; q--) {
    const uint2 wt = __ldg(p.tn_list + (q - 1));
    const uint32_t gt = p.grids[wt.x].tile_start + wt.y;
    const unsigned int ce = (unsigned int)(retry ? (__ldcg(p.tile_claim + gt) >> 32) : 0ull);
    const bool ok = !forced && p.tile_slot[gt] == MCX_NONE && !(retry && ce >= epoch_first && ce < epoch);
    if (ok && n_vac < SURFSURF_MAX_MATCHES) vacant[n_vac++] = wt;
  }
  if (n_vac + n_rec < needed) return false;
  int assigned[6];   // -1 nothing, 0/1 recycled site, 2 + j vacant tile j
  for (int e = 0; e < 6; e++) assigned[e] = -1;
  const int to_recycle = n_ent < n_rec ? n_ent : n_rec;
  int next_available = 0;
  const uint32_t num_players = (uint32_t)(n_ent + n_reactants);
  for (int guard = 0; next_available < to_recycle && guard < 100000; guard++) {  // :2159-2191
    const uint32_t rnd = rs.next() % num_players;
    if (rnd < (uint32_t)n_reactants) continue;
    const int e = (int)rnd - n_reactants;
    if (!(ent[e] & 1u)) continue;
    if (assigned[e] >= 0) continue;
    assigned[e] = next_available++;
  }
  uint32_t used = 0;
  for (int e = 0; e < n_ent; e++) {  // :2232-2283: every entry without a position draws a vacant tile
    if (assigned[e] >= 0) continue;
    int attempts = 0; bool found = false;
    while (!found && attempts < 10) {
      const uint32_t rnd = rs.next() % (uint32_t)n_vac;
      if ((used >> rnd) & 1u) { attempts++; continue; }
      assigned[e] = 2 + (int)rnd; used |= 1u << rnd; found = true;
    }
    if (attempts >= 10) return false;
  }
  orient_bits = draw_orientation_bits(pw, rs);
  int cnt = 0; uint32_t vac_mask = 0;
  for (int e = 0; e < n_ent; e++) {
    if (ent[e] != 1u) continue;   // created surface products only
    const int a = assigned[cnt];
    uint2 wt; double2 uv;
    if (a >= 2) {
      wt = vacant[a - 2]; vac_mask |= 1u << cnt;
      tile_uv(p, wt.x, wt.y, true, rs, uv.x, uv.y);   // grid2uv_random (:2852-2855)
    } else {
      const int r = a < 0 ? 0 : a;
      wt = make_uint2(rec_wall[r], rec_tile[r]); uv = rec_uv[r];
    }
    p.prop_ptile[slot * MCX_MAX_PRODUCTS + cnt] = wt;
    p.prop_puv[slot * MCX_MAX_PRODUCTS + cnt] = uv;
    cnt++;
  }
  p.prop_pmask[slot] = (uint32_t)cnt | (vac_mask << 4);
  return true;
}

// ---- react_2D_all_neighbors (diffuse_react_event.cpp:1250-1393): the molecules on the tiles around the tile of a surface
// molecule (after its move); the lists are static (tn_start / tn_list, built on the host), walls without a grid are
// left out here.  Out of line: its candidate arrays stay out of the stack frame of the common path.  Returns true when
// a reaction fired (out: class, pathway, partner, time, orientation / tile bits).
__device__ __noinline__ bool react_2d_all_neighbors(const DevParams& p, uint32_t self_id, uint32_t slot, uint32_t species, uint32_t flags,
                                                    const SurfState& ss, double t_steps, double t_now, unsigned int epoch_first,
                                                    unsigned int epoch, bool retry, Stream& rs, Tracer& tc, Outcome& out, int& err) {
  bool fired = false;
  const uint32_t gt0 = p.grids[ss.wall].tile_start + ss.tile;
  const uint32_t qb = __ldg(p.tn_start + gt0), qe = __ldg(p.tn_start + gt0 + 1);
  int m_rc[SURFSURF_MAX_MATCHES]; double m_factor[SURFSURF_MAX_MATCHES]; uint32_t m_slot[SURFSURF_MAX_MATCHES], m_id[SURFSURF_MAX_MATCHES];
  int n_match = 0; uint32_t n_nb = 0;
  const int my_orient = (flags & DF_ORIENT_UP) ? 1 : -1;
  for (uint32_t q = qb; q < qe; q++) {
    const uint2 wt = __ldg(p.tn_list + q);
    // Wall::has_initialized_grid (grid_utils.inl:1243, 783-790); the molecule's own wall always counts (a wall it
    // has just moved to gets its grid with it)
    if (wt.x != ss.wall && !p.wall_has_grid[wt.x]) continue;
    n_nb++;
    const DevGrid& ng = p.grids[wt.x];
    const uint32_t occ = p.tile_slot[ng.tile_start + wt.y];
    if (occ == MCX_NONE) continue;
    const MolRec nsm = load_rec_volatile(p.recA, occ);
    if (nsm.id == self_id || (nsm.sf & DF_DEAD)) continue;  // the tile table still shows the mover on its old tile
    const int rc = p.surfsurf[species * p.n_species + (nsm.sf & SF_SPECIES_MASK)];
    if (rc < 0 || !orientations_match(p.classes[rc], my_orient, (nsm.sf & DF_ORIENT_UP) ? 1 : -1)) continue;
    if (n_match >= SURFSURF_MAX_MATCHES) { err = MCX_ERR_STATE; break; }
    m_rc[n_match] = rc; m_factor[n_match] = t_steps / ng.binding_factor; m_slot[n_match] = occ; m_id[n_match] = nsm.id;
    n_match++;
  }
  if (n_nb != 0 && n_match != 0) {
    const double local_prob_factor = 3.0 / (double)n_nb;
    int which = 0, pathway;
    if (n_match == 1) pathway = test_bimolecular(p, p.classes[m_rc[0]], m_factor[0], rs, local_prob_factor);
    else {
      // RxnUtils::test_many_bimolecular with all_neighbors_flag (rxn_utils.inl:475-580)
      double cum[SURFSURF_MAX_MATCHES];
      cum[0] = p.classes[m_rc[0]].max_fixed_p * local_prob_factor / m_factor[0];
      for (int i = 1; i < n_match; i++) cum[i] = cum[i - 1] + p.classes[m_rc[i]].max_fixed_p * local_prob_factor / m_factor[i];
      double prob;
      which = -1;
      bool none = false;
      if (cum[n_match - 1] > 1.0) prob = rs.dbl() * cum[n_match - 1];
      else { prob = rs.dbl(); none = prob > cum[n_match - 1]; }
      if (!none) {
        // binary_search_double over the reference's zero-padded array of 2 n entries (:559)
        int min_idx = 0, max_idx = 2 * n_match - 1;
        while (max_idx - min_idx > 1) {
          const int mid = (max_idx + min_idx) / 2;
          if (prob > (mid < n_match ? cum[mid] : 0.0)) min_idx = mid; else max_idx = mid;
        }
        which = prob > (min_idx < n_match ? cum[min_idx] : 0.0) ? max_idx : min_idx;
        if (which >= n_match) which = n_match - 1;
      }
      pathway = 0;  // sic (TODO_PATHWAYS, :1367): the first pathway of the chosen class
    }
    if (tc.tr) for (int i = 0; i < n_match; i++) { if (tc.tr->n_collisions < MCX_TRACE_K) tc.tr->partner[tc.tr->n_collisions] = m_id[i]; tc.tr->n_collisions++; }
    if (which >= 0 && pathway >= 0) {
      const int rc = m_rc[which];
      const DevClass& cl = p.classes[rc];
      const DevPathway& pw = p.pathways[cl.first_pathway + pathway];
      // random draws in the reference's order: tile assignment (find_surf_product_positions), then orientations
      uint32_t bits = 0;
      bool blocked = false;
      if (pw.general) {
        const uint32_t ps = m_slot[which];
        const bool me_r0 = species == cl.r0;
        uint32_t rw[2], rt[2]; double2 ruv[2]; int n_rec = 0;
        for (int r = 0; r < 2; r++) {
          if ((pw.keep_mask >> r) & 1u) continue;
          const bool mine = (r == 0) == me_r0;
          rw[n_rec] = mine ? ss.wall : p.swallA[ps]; rt[n_rec] = mine ? ss.tile : p.stileA[ps];
          ruv[n_rec] = mine ? make_double2(ss.u, ss.v) : p.suvA[ps];
          n_rec++;
        }
        blocked = !place_general(p, cl, pw, slot, ss.wall, ss.tile, n_rec, rw, rt, ruv, false, epoch_first, epoch, retry, rs, bits);
      } else {
        bits = surfsurf_position_bits(p, pw, species == cl.r0, rs);
        bits |= draw_orientation_bits(pw, rs);
      }
      if (blocked) { tc.ev(EV_BLOCKED, m_id[which]); return false; }   // RX_BLOCKED: the molecule survives (:1388-1392)
      tc.ev(EV_SURFSURF | (uint32_t)pathway, (uint32_t)rc);
      tc.ev(EV_RXN | (bits & 0x7Fu), m_id[which]);
      if (tc.tr) { tc.tr->rxn_class = rc; tc.tr->rxn_pathway = pathway; tc.tr->rxn_partner = m_id[which]; tc.tr->t_event = t_now; }
      out.rxn_class = rc; out.pathway = pathway; out.partner_slot = m_slot[which]; out.partner_id = m_id[which];
      out.t_event = t_now;  // collision_time = diffusion_start_time (:1343)
      out.orient_bits = bits;
      fired = true;
    }
  }
  return fired;
}

template <bool RETRY, bool WITH_DISK, bool SURF>
__device__ void evaluate_iteration(const DevParams& p, const MolRec& m, uint32_t slot, double t_sched, double t_unimol_in,
                                   uint32_t created_wall, uint32_t created_tile, SurfState ss, unsigned int epoch,
                                   Stream& rs, bool forced, Outcome& out, LocalStats& ls, Tracer& tc, int& err, const Group& grp) {
  const uint32_t species = m.sf & SF_SPECIES_MASK;
  const DevSpecies sp = p.species[species];
  const double it = (double)p.iteration, t_end = it + 1;
  uint32_t flags = (m.sf & ~SF_SPECIES_MASK) & ~DF_CREATED_ON_SURF;  // the guard below lives for this iteration only
  D3 pos = {m.x, m.y, m.z};
  uint32_t subpart = subpart_index(p, pos);
  double t_now = (flags & DF_PARTIAL) ? t_sched : it;
  double unimol_time = (flags & DF_HAS_UNIMOL) ? t_unimol_in : MCX_TIME_INVALID;
  const bool can_diffuse = (sp.flags & MCX_SP_CAN_DIFFUSE) != 0;
  const bool can_vol_react = sp.can_vol_react != 0 && !forced;
  // SPECIES_FLAG_CAN_SURFSURF: diffuse_surf_molecule runs for such a molecule even when it cannot diffuse (:274-283)
  const bool can_ss = SURF && (flags & DF_SURF) && sp.can_surf_surf != 0 && p.surfsurf != nullptr;
  bool ss_fired = false;
  out.rxn_class = -1; out.pathway = -1; out.partner_slot = MCX_NONE; out.partner_id = MCX_NONE; out.t_event = 0;
  out.kind = MCX_OUT_NONE; out.orient_bits = 0;
  out.surf_moved = false; out.s_wall = ss.wall; out.s_tile = ss.tile; out.s_u = ss.u; out.s_v = ss.v;
  bool surf_tile_changed = false;
  bool decided = false;  // a claiming event or an error ended the evaluation; `out` is complete
  // a counted volume that is only a guess: Partition::add_volume_molecule's ray cast (partition.h:572-576 ->
  // compute_counted_volume_using_waypoints), done when the molecule is first evaluated
  if (flags & DF_CVI_PENDING) {
    flags &= ~DF_CVI_PENDING;
    if (!(flags & DF_SURF) && p.cv_mask) {
      RayScan sc;
      scan_ray(p, pos, rs, sc);
      if (!sc.redo) { const uint32_t kq = cv_lookup(p, sc.inside_mask & p.cv_all); if (kq != MCX_NONE) flags = (flags & ~SF_CVI_MASK) | (kq << SF_CVI_SHIFT); }
    }
  }

  bool again = true;
  for (int sub_guard = 0; again && !decided && sub_guard < 1000; sub_guard++) {
    again = false;
    // -- unimolecular firing (diffuse_react_event.cpp:215-223, :1764-1826)
    if (unimol_time != MCX_TIME_INVALID && unimol_time <= t_now) {
      int rc = p.unimol[species];
      const DevClass& cl = p.classes[rc];
      int pathway = 0;
      if (cl.n_pathways > 1) {  // which_unimolecular, rxn_utils.inl:774-783
        double match = rs.dbl() * cl.max_fixed_p;
        int min_idx = 0, max_idx = (int)cl.n_pathways - 1;
        const DevPathway* A = p.pathways + cl.first_pathway;
        while (max_idx - min_idx > 1) {
          int mid = (max_idx + min_idx) / 2;
          if (match > A[mid].cum_prob) min_idx = mid; else max_idx = mid;
        }
        pathway = match > A[min_idx].cum_prob ? max_idx : min_idx;
      }
      bool blocked = false;
      if (SURF && (flags & DF_SURF)) {
        const DevPathway& upw = p.pathways[cl.first_pathway + pathway];
        if (upw.general) {
          const double2 self_uv = make_double2(ss.u, ss.v);
          blocked = !place_general(p, cl, upw, slot, ss.wall, ss.tile, (upw.keep_mask & 1u) ? 0 : 1, &ss.wall, &ss.tile, &self_uv, forced,
                                   round_epoch0(p), epoch, RETRY, rs, out.orient_bits);
        } else out.orient_bits = draw_orientation_bits(upw, rs);
      }
      if (blocked) {
        // RX_BLOCKED (outcome_unimolecular :2976-2999): no room for the products; the molecule lives on and draws a new lifetime
        tc.ev(EV_BLOCKED, (uint32_t)rc);
        flags |= DF_SCHED_UNIMOL;
      } else {
        tc.ev(EV_UNIMOL | (uint32_t)pathway, (uint32_t)rc);
        if (tc.tr) { tc.tr->rxn_class = rc; tc.tr->rxn_pathway = pathway; tc.tr->t_event = unimol_time; }
        out.kind = MCX_OUT_UNIMOL; out.pos = pos; out.rxn_class = rc; out.pathway = pathway; out.t_event = unimol_time;
        out.t_now = t_now; out.flags = flags; out.unimol_time = unimol_time;
        decided = true;
      }
    }
    if (!decided) {
      // -- newbie lifetime (:232-236 -> pick_unimol_rxn_class_and_set_rxn_time :1731-1758, time_of_unimol)
      if (flags & DF_SCHED_UNIMOL) {
        flags &= ~DF_SCHED_UNIMOL;
        int rc = p.unimol[species];
        if (rc < 0) unimol_time = MCX_TIME_INVALID;
        else {
          double k_tot = p.classes[rc].max_fixed_p;
          double pr = rs.dbl();
          double from_now = (k_tot <= 0 || !distinguishable_d(pr, 0, MCX_EPS)) ? MCX_TIME_FOREVER : -mcx_log(pr) / k_tot;
          unimol_time = t_now + from_now;
        }
      }
      // -- get_max_time (:164-198)
      double max_time = t_end - t_now;
      if (unimol_time != MCX_TIME_INVALID && unimol_time < t_now + max_time) max_time = unimol_time - t_now;

      if (SURF && (flags & DF_SURF) && (can_diffuse || can_ss)) {
        // ---- diffuse_surf_molecule (:1071-1246)
        double t_steps = sp.time_step > max_time ? max_time : sp.time_step;
        double steps;
        if (sp.time_step > max_time) {
          steps = max_time / sp.time_step;
          if (steps < MCX_EPS) { t_steps = MCX_EPS * sp.time_step; steps = MCX_EPS; }
        } else steps = 1.0;
        const double space_factor = steps == 1.0 ? sp.space_step : sp.space_step * sqrt(steps);
        const uint32_t original_wall = ss.wall;
        const unsigned int epoch_first = round_epoch0(p);
        bool placed = false;
        for (int find_new_position = can_diffuse ? 11 : 0; find_new_position > 0 && !placed; find_new_position--) {  // SURFACE_DIFFUSION_RETRIES + 1
          // pick_surf_displacement (diffusion_utils.inl:60-96)
          double au, av, f;
          do {
            const uint32_t n = rs.next();
            au = 2 * 1.52587890625e-5 * (n & 0xFFFFu) - 1;
            av = 2 * 1.52587890625e-5 * (n >> 16) - 1;
            f = au * au + av * av;
          } while ((f < MCX_EPS) || (f > 1));
          const double normal_factor = sqrt(-mcx_log(f) / f);
          const double du = au * (normal_factor * space_factor), dv = av * (normal_factor * space_factor);
          double nu, nv;
          bool absorbed_at_border;
          const uint32_t new_wall = ray_trace_surf(p, ss.wall, ss.u, ss.v, du, dv, nu, nv, species, (flags & DF_ORIENT_UP) ? 1 : -1,
                                                   absorbed_at_border);
          if (absorbed_at_border) {  // absorptive region border (:1152-1160): destroyed at the start of the step, no products
            tc.ev(EV_ABSORB, ss.wall);
            if (tc.tr) tc.tr->t_event = t_now;
            out.kind = MCX_OUT_ABSORBED; out.pos = pos; out.t_event = t_now; out.t_now = t_now; out.flags = flags;
            out.unimol_time = unimol_time;
            decided = true; placed = true;
          }
          bool ok = !absorbed_at_border && new_wall != MCX_NONE;
          uint32_t new_tile = MCX_NONE;
          if (ok) { new_tile = uv2grid(p, new_wall, nu, nv); ok = new_tile != MCX_NONE; }
          bool changes_tile = false;
          if (ok && (new_wall != ss.wall || new_tile != ss.tile)) {
            // move_sm_on_same_triangle / move_sm_to_new_triangle (diffusion_utils.inl:453-548): the tile must be vacant
            // in the snapshot and not claimed by a mover of an earlier conflict round; none in the forced pass
            const uint32_t gt = p.grids[new_wall].tile_start + new_tile;
            const unsigned int ce = (unsigned int)(RETRY ? (__ldcg(p.tile_claim + gt) >> 32) : 0ull);
            ok = !forced && p.tile_slot[gt] == MCX_NONE && !(RETRY && ce >= epoch_first && ce < epoch);
            changes_tile = true;
          }
          if (ok) {
            if (changes_tile) surf_tile_changed = true;
            if (new_wall != ss.wall) {  // reschedule the unimolecular reaction of a molecule that changed wall (:1170-1186)
              double time_until_unimol = unimol_time - t_steps - t_now;
              time_until_unimol = (time_until_unimol < 0) ? 0 : time_until_unimol;
              if (unimol_time == MCX_TIME_INVALID || (time_until_unimol > MCX_EPS || time_until_unimol > MCX_EPS * (t_now + t_steps))) {
                unimol_time = MCX_TIME_INVALID;
                flags |= DF_SCHED_UNIMOL;
              }
            }
            ss.wall = new_wall; ss.tile = new_tile; ss.u = nu; ss.v = nv;
            const DevWall& fw = p.walls[new_wall];
            pos = D3{nu * fw.ux + nv * fw.vx + fw.v0x, nu * fw.uy + nv * fw.vy + fw.v0y, nu * fw.uz + nv * fw.vz + fw.v0z};
            subpart = subpart_index(p, pos);
            out.surf_moved = true;
            placed = true;
          }
        }
        // ---- react_2D_all_neighbors (:1250-1393): after its move the molecule tests the molecules on the tiles around its own
        if (can_ss && !decided && !forced && !(sp.flags & MCX_SP_CANT_INITIATE) && p.tn_start)
          ss_fired = react_2d_all_neighbors(p, m.id, slot, species, flags, ss, t_steps, t_now, round_epoch0(p), epoch, RETRY, rs, tc, out, err);
        if ((!can_diffuse || ss.wall != original_wall) && unimol_time >= t_end) {  // MCell3 compatibility rule (:1222-1236)
          unimol_time = MCX_TIME_INVALID;
          flags |= DF_SCHED_UNIMOL;
        }
        max_time = t_steps;
        out.s_wall = ss.wall; out.s_tile = ss.tile; out.s_u = ss.u; out.s_v = ss.v;
      } else if (can_diffuse) {
        // ---- compute_vol_displacement (diffusion_utils.inl:366-432)
        double steps = 1.0, t_steps = steps * sp.time_step, r_rate_factor, scale;
        if (t_steps > max_time) { t_steps = max_time; steps = max_time / sp.time_step; }
        if (steps < MCX_EPS) { steps = MCX_EPS; t_steps = MCX_EPS * sp.time_step; }
        if (steps == 1.0) { scale = sp.space_step; r_rate_factor = 1.0; }
        else { double rate_factor = sqrt(steps); r_rate_factor = 1.0 / rate_factor; scale = rate_factor * sp.space_step; }
        D3 remaining;
#pragma unroll 1
        for (int axis = 0; axis < 3; axis++) {  // one copy of the Ziggurat code (instruction-cache footprint)
          const double g = scale * rs.gauss() * 0.70710678118654752440;
          if (axis == 0) remaining.x = g; else if (axis == 1) remaining.y = g; else remaining.z = g;
        }
        max_time = t_steps;

        uint32_t last_hit_wall = MCX_NONE;
        if (created_tile == MCX_KEPT_AT_WALL) {
          // a KEPT reactant of a surface / wall reaction carries on from the wall of its event like the reference's does
          // within the same step (last_hit_wall_index = wall, the remaining displacement leads away from it,
          // diffuse_react_event.cpp:945-975, reflect_from_wall): the first displacement is mirrored away from that wall
          // if it points into it, and the first trace skips the wall
          const DevWall& kw = p.walls[created_wall];
          const D3 kn = {kw.nx, kw.ny, kw.nz};
          const double dd = dot3(kn, pos) - kw.dist, dn = dot3(remaining, kn);
          if (dd > 0 ? dn < 0 : dn > 0) remaining = remaining + kn * (-2.0 * dn);
          last_hit_wall = created_wall;
          created_wall = created_tile = MCX_NONE;
        }
        double elapsed = t_now;
        // ---- diffuse_vol_molecule loop (:423-572)
        bool tracing = true;
        for (int trace_guard = 0; tracing; trace_guard++) {
          if (trace_guard > 100000) { err = MCX_ERR_STATE; tracing = false; }
          // ---- ray_trace_vol (:627-780)
          D3 part_disp = remaining;
          if (!in_partition(p, pos + remaining)) part_disp = displacement_up_to_partition_boundary(p, pos, remaining);
          SpList spw; SpSet spm; spm.n = 0; spm.overflow = false;
          uint32_t last_subpart = collect_crossed_subparts(p, pos, subpart, part_disp, can_vol_react, true, spw, spm);
          D3 up_to_wall = remaining;
          bool hit = false, hit_in_last = false;
          WallHit wh;
          for (int k = 0; k < spw.n && !hit; k++) {
            if (closest_wall_collision(p, pos, spw.v[k], last_hit_wall, rs, remaining, up_to_wall, wh, ls, tc, grp)) {
              hit = true; hit_in_last = last_subpart == spw.v[k];
            }
          }
          bool reacted = false, aborted = false;
          if (can_vol_react) {
            if (hit && !hit_in_last) {
              spm.n = 0;
              collect_crossed_subparts(p, pos, subpart, up_to_wall, true, false, spw, spm);
            }
            if (spm.overflow) err = MCX_ERR_OVERFLOW;
            // the subpartition filter can be skipped only when the swept box lies strictly inside the own subpart
            bool filter = true;
            if (spm.n == 1) {
              int si[3] = {(int)(subpart % p.n_sp), (int)((subpart / p.n_sp) % p.n_sp), (int)(subpart / (p.n_sp * p.n_sp))};
              const double pad = p.R * 1.000001 + 1e-6;
              double lx = p.ox + si[0] * p.sp_len, ly = p.oy + si[1] * p.sp_len, lz = p.oz + si[2] * p.sp_len;
              D3 e = pos + remaining;
              filter = !(fmin(pos.x, e.x) - pad > lx && fmax(pos.x, e.x) + pad < lx + p.sp_len &&
                         fmin(pos.y, e.y) - pad > ly && fmax(pos.y, e.y) + pad < ly + p.sp_len &&
                         fmin(pos.z, e.z) - pad > lz && fmax(pos.z, e.z) + pad < lz + p.sp_len);
            }
            // sort_collisions_by_time (:341-364) realised as repeated selection of the next collision
            const double t_limit = hit ? wh.t : MCX_TIME_FOREVER;
            double t_last = -1.0; uint32_t id_last = 0;
            bool scanning = true;
            while (scanning) {
              PartnerHit ph;
              int cnt = scan_partners<RETRY>(p, pos, remaining, m.id, species, spm, filter, t_last, id_last, t_limit, ph, grp);
              if (cnt == 0) scanning = false;
              else {
                t_last = ph.t; id_last = ph.id;
                ls.volvol_collisions++;
                // collide_and_react_with_vol_mol (:786-829): the walls of the collision subpartition may hide part
                // of the interaction disk or block the reaction altogether
                double factor = 1.0;
                if (!(ph.t < MCX_EPS) && (__ldg(p.sp_flags + subpart_index(p, pos + remaining * ph.t)) & 1)) {
                  if (WITH_DISK) {
                    const MolRec tg = RETRY ? load_rec_volatile(p.recA, ph.slot) : load_rec(p.recA, ph.slot);
                    factor = exact_disk(p, pos + remaining * ph.t, remaining, species, D3{tg.x, tg.y, tg.z}, err);
                    if (factor < 0) tc.ev(EV_BLOCKED, ph.id);
                    else if (factor != 1.0) tc.ev(EV_DISK, (uint32_t)llrint(factor * 1073741824.0));  // 2^-30 resolution
                  } else if (exd_any_wall_in_reach(p, pos + remaining * ph.t, remaining)) {
                    aborted = true; scanning = false; factor = -1.0;
                  }
                }
                if (!(ph.t < MCX_EPS) && !(factor < 0)) {  // is_immediate_collision; blocked by a wall
                  double abs_t = elapsed + t_steps * ph.t;
                  double scaling = factor * r_rate_factor;
                  tc.ev(EV_COLL, ph.id);
                  if (tc.tr) { if (tc.tr->n_collisions < MCX_TRACE_K) tc.tr->partner[tc.tr->n_collisions] = ph.id; tc.tr->n_collisions++; }
                  int pathway = test_bimolecular(p, p.classes[ph.rxn_class], scaling, rs);
                  if (pathway >= 0) {
                    tc.ev(EV_RXN | (uint32_t)pathway, (uint32_t)ph.rxn_class);
                    if (tc.tr) { tc.tr->rxn_class = ph.rxn_class; tc.tr->rxn_pathway = pathway; tc.tr->rxn_partner = ph.id; tc.tr->t_event = abs_t; }
                    out.kind = MCX_OUT_REACTED; out.pos = pos + remaining * ph.t;
                    out.rxn_class = ph.rxn_class; out.pathway = pathway; out.partner_slot = ph.slot; out.partner_id = ph.id;
                    out.t_event = abs_t; out.t_now = t_now; out.flags = flags; out.unimol_time = unimol_time;
                    reacted = true;
                    scanning = false;
                  }
                }
                if (cnt == 1) scanning = false;
              }
            }
          }
          if (aborted) { err = MCX_INTERNAL_NEEDS_DISK; out.kind = MCX_OUT_NONE; decided = true; tracing = false; }
          else if (reacted) { decided = true; tracing = false; }
          else if (!hit) {  // RayTraceState::FINISHED
            pos = pos + remaining;
            if (!in_partition(p, pos)) { err = MCX_ERR_ESCAPED; out.kind = MCX_OUT_NONE; out.pos = pos; decided = true; }
            else subpart = subpart_index(p, pos);
            tracing = false;
          } else {
            // ---- wall collision (:476-567)
            const int side = wh.side;  // W_FRONT / W_BACK
            const uint32_t wclass = p.wall_class[wh.wall];
            int action = MCX_SURF_REFLECTIVE;
            const uint32_t action_at = (species * p.n_surf_classes + wclass) * 2 + (side == W_FRONT ? 0 : 1);
            if (wclass != MCX_NONE) action = p.surf_action[action_at];
            if (tc.tr) {
              if (tc.tr->n_wall_hits < MCX_TRACE_K) { tc.tr->wall[tc.tr->n_wall_hits] = wh.wall; tc.tr->wall_side[tc.tr->n_wall_hits] = side; }
              tc.tr->n_wall_hits++;
            }
            // ---- collide_and_react_with_surf_mol (:845-975): the surface molecule on the tile under the hit point
            bool surf_reacted = false;
            if (SURF && sp.can_vol_surf && p.n_tiles) {
              const DevGrid& g = p.grids[wh.wall];
              const uint32_t j = xyz2grid(p, wh.pos, wh.wall);
              const uint32_t occ = p.tile_slot[g.tile_start + j];
              bool occupied = false;
              MolRec sm;
              if (occ != MCX_NONE) {
                sm = RETRY ? load_rec_volatile(p.recA, occ) : load_rec(p.recA, occ);
                occupied = !(sm.sf & DF_DEAD);
              }
              if (occupied && created_wall == wh.wall && created_tile == j) {
                created_wall = created_tile = MCX_NONE;  // no rebinding where it was just created; next time yes
                occupied = false;
              }
              if (occupied) {
                const uint32_t ssp = sm.sf & SF_SPECIES_MASK;
                const int rc = p.volsurf[species * p.n_species + ssp];
                const int coll_orient = side == W_FRONT ? 1 : -1;
                const int surf_orient = (sm.sf & DF_ORIENT_UP) ? 1 : -1;
                if (rc >= 0 && orientations_match(p.classes[rc], coll_orient, surf_orient)) {
                  const double scaling = r_rate_factor / g.binding_factor;
                  const double abs_t = elapsed + t_steps * wh.t;
                  tc.ev(EV_SURFMOL | (uint32_t)side, sm.id);
                  if (tc.tr) { if (tc.tr->n_collisions < MCX_TRACE_K) tc.tr->partner[tc.tr->n_collisions] = sm.id; tc.tr->n_collisions++; }
                  int pathway = test_bimolecular(p, p.classes[rc], scaling, rs);
                  if (pathway >= 0) {
                    const DevPathway& vpw = p.pathways[p.classes[rc].first_pathway + pathway];
                    if (vpw.general) {
                      const uint32_t pwall = p.swallA[occ], ptile = p.stileA[occ];
                      const double2 puv = p.suvA[occ];
                      uint32_t ob = 0;
                      if (!place_general(p, p.classes[rc], vpw, slot, pwall, ptile, (vpw.keep_mask & 2u) ? 0 : 1, &pwall, &ptile, &puv, forced,
                                         round_epoch0(p), epoch, RETRY, rs, ob)) {
                        tc.ev(EV_BLOCKED, sm.id);   // RX_BLOCKED (:936-975): no reaction, the molecule goes on to the wall
                        pathway = -1;
                      }
                      out.orient_bits = ob | (coll_orient > 0 ? ORIENT_BIT_FRONT : 0u);
                    } else
                    out.orient_bits = draw_orientation_bits(vpw, rs) | (coll_orient > 0 ? ORIENT_BIT_FRONT : 0u);
                  }
                  if (pathway >= 0) {
                    tc.ev(EV_RXN | (uint32_t)pathway, (uint32_t)rc);
                    if (tc.tr) { tc.tr->rxn_class = rc; tc.tr->rxn_pathway = pathway; tc.tr->rxn_partner = sm.id; tc.tr->t_event = abs_t; }
                    out.kind = MCX_OUT_REACTED; out.pos = wh.pos;
                    out.rxn_class = rc; out.pathway = pathway; out.partner_slot = occ; out.partner_id = sm.id;
                    out.t_event = abs_t; out.t_now = t_now; out.flags = flags; out.unimol_time = unimol_time;
                    surf_reacted = true; decided = true; tracing = false;
                  }
                }
              }
            }
            if (!surf_reacted && action == MCX_SURF_STANDARD) {
              // collide_and_react_with_walls (:1034-1066): test_intersect of the one matching class; a reaction is a
              // claiming event (the proposal's partner word carries the wall), no reaction reflects
              const int wrc = p.surf_rxn[action_at];
              const int pathway = test_intersect(p, p.classes[wrc], r_rate_factor, rs);
              action = MCX_SURF_REFLECTIVE;
              if (pathway >= 0) {
                const double abs_t = elapsed + t_steps * wh.t;
                out.orient_bits = draw_orientation_bits(p.pathways[p.classes[wrc].first_pathway + pathway], rs) |
                                  (side == W_FRONT ? ORIENT_BIT_FRONT : 0u);
                tc.ev(EV_WALLRXN | (uint32_t)side, wh.wall);
                tc.ev(EV_RXN | (uint32_t)pathway, (uint32_t)wrc);
                if (tc.tr) { tc.tr->rxn_class = wrc; tc.tr->rxn_pathway = pathway; tc.tr->t_event = abs_t; }
                out.kind = MCX_OUT_WALLRXN; out.pos = wh.pos; out.rxn_class = wrc; out.pathway = pathway;
                out.partner_slot = wh.wall; out.partner_id = MCX_NONE;
                out.t_event = abs_t; out.t_now = t_now; out.flags = flags; out.unimol_time = unimol_time;
                surf_reacted = true; decided = true; tracing = false;
              }
            }
            if (surf_reacted) {
            } else if (action == MCX_SURF_TRANSPARENT) {  // cross_transparent_wall (:3007-3099)
              tc.ev(EV_TRANSP | (uint32_t)side, wh.wall);
              ls.transparent++;
              pos = wh.pos; subpart = subpart_index(p, pos);
              if (p.wall_cv) {  // update_counted_volume_id_when_crossing_wall (collision_utils.inl:1637-1694)
                const uint32_t nc = cv_cross(p, flags >> SF_CVI_SHIFT, wh.wall, side == W_FRONT);
                if (nc == MCX_NONE) err = MCX_ERR_STATE;   // a set of counted objects the host did not list
                else flags = (flags & ~SF_CVI_MASK) | (nc << SF_CVI_SHIFT);
              }
              remaining = remaining * (1.0 - wh.t);
              elapsed += t_steps * wh.t;
              t_steps *= (1.0 - wh.t);
              if (t_steps < MCX_EPS) t_steps = MCX_EPS;
              last_hit_wall = wh.wall;
            } else if (action == MCX_SURF_ABSORPTIVE) {  // test_intersect: two draws (rxn_utils.inl:593-626)
              double abs_t = elapsed + t_steps * wh.t;
              (void)rs.dbl(); (void)rs.dbl();
              tc.ev(EV_ABSORB | (uint32_t)side, wh.wall);
              if (tc.tr) tc.tr->t_event = abs_t;
              out.kind = MCX_OUT_ABSORBED; out.pos = wh.pos; out.t_event = abs_t; out.t_now = t_now; out.flags = flags;
              out.unimol_time = unimol_time;
              decided = true; tracing = false;
            } else {  // reflect_from_wall, collision_utils.inl:1711-1747
              tc.ev(EV_WALL | (uint32_t)side, wh.wall);
              ls.reflections++;
              elapsed += t_steps * wh.t;
              pos = wh.pos; subpart = subpart_index(p, pos);
              t_steps *= (1.0 - wh.t);
              last_hit_wall = wh.wall;
              const DevWall& f = p.walls[wh.wall];
              D3 n = {f.nx, f.ny, f.nz};
              double reflect_factor = -2.0 * dot3(remaining, n);
              remaining = (remaining + n * reflect_factor) * (1.0 - wh.t);
            }
          }
        }
      }
      created_wall = created_tile = MCX_NONE;  // the guard belongs to the first DiffuseAction only (:116-129)
      if (!decided) {
        // -- reschedule (:283-336)
        if (can_diffuse || can_ss) {
          t_now += max_time;
          if ((unimol_time != MCX_TIME_INVALID && unimol_time < t_end) || (t_now < t_end && !cmp_eq_d(t_now, t_end, MCX_EPS))) again = true;
          else {
            double r = round(t_now);
            if (cmp_eq_d(t_now, r, MCX_SQRT_EPS)) t_now = r;
          }
        } else {
          if (unimol_time != MCX_TIME_INVALID) {
            t_now = unimol_time;
            if (unimol_time < t_end) again = true;
          } else t_now = MCX_TIME_FOREVER;
        }
        if (SURF && ss_fired) {
          // the surface-surface reaction is a claiming event (the initiator, the partner it consumes, the tile it moved to);
          // a kept initiator has used up its step like any other mover
          out.kind = MCX_OUT_REACTED; out.pos = pos; out.t_now = t_now; out.unimol_time = unimol_time;
          out.flags = again ? (flags | DF_PARTIAL) : (flags & ~DF_PARTIAL);
          decided = true; again = false;
        } else if (SURF && surf_tile_changed) {
          // taking a new tile is a claiming event: the evaluation ends here, what is left of the iteration is taken
          // lazily next iteration (like a kept initiator)
          out.kind = MCX_OUT_SURFMOVE; out.pos = pos; out.t_now = t_now; out.unimol_time = unimol_time;
          out.flags = again ? (flags | DF_PARTIAL) : (flags & ~DF_PARTIAL);
          tc.ev(EV_SURFMOVE | (ss.tile & 0xFFFFFFu), ss.wall);
          decided = true; again = false;
        }
      }
    }
  }
  if (!decided) {
    out.kind = (can_diffuse || can_ss) ? MCX_OUT_MOVED : MCX_OUT_STATIC;
    out.pos = pos; out.t_now = t_now; out.flags = flags & ~DF_PARTIAL; out.unimol_time = unimol_time;
  }
}
