// oracle_exact_disk.h — TEST INFRASTRUCTURE (part of the CPU oracle, see oracle.cpp header).
//
// Restatement of ExactDiskUtils::exact_disk (src4/exact_disk_utils.inl:54-1145; == exact_disk, src/diffuse.c:1365,
// against which it is pinned through oracle/_ref/libmcell3ref.so): the fraction of the interaction disk (radius R,
// perpendicular to the motion, centred at the collision point) that walls of the collision subpartition do not
// hide, or -1 when a wall lies between the moving molecule and its target.  The reference's heap-allocated
// linked lists are index-linked entries of one vector here; traversal orders are the reference's.
//
// Not restated: find_boundaries_occluding_disk (:275-505) — it only runs when use_expanded_list is false, and the
// converter turns the expanded list on whenever a volume-volume reaction exists (mcell4_converter.cpp:84-87), which
// is the only case exact_disk is called in.
#pragma once
#include <cmath>
#include <vector>

namespace orc_exd {

static const double EXD_EPS = 1e-12, EXD_SQRT_EPS = 1e-6, EXD_PI = 3.14159265358979323846;
enum { ROLE_UNDEF = 0, ROLE_HEAD, ROLE_TAIL, ROLE_CROSS, ROLE_SPAN, ROLE_OTHER };
struct Vtx { double u = 0, v = 0, r2 = 0, zeta = 0; int next = -1, e = -1, span = -1, role = ROLE_UNDEF; };

static inline bool exd_distinguishable(double a, double b, double eps) {  // src/util.c:449-463
  double c = fabs(a - b);
  a = fabs(a);
  if (a < 1) a = 1;
  b = fabs(b);
  if (b < a) eps *= a; else eps *= b;
  return c > eps;
}

// exd_zetize, exact_disk_utils.inl:54-80
static double zetize(double y, double x) {
  if (y >= 0) {
    if (x >= 0) { if (x < y) return 1 - 0.5 * x / y; else return 0.5 * y / x; }
    else { if (-x < y) return 1 - 0.5 * x / y; else return 2 + 0.5 * y / x; }
  } else {
    if (x <= 0) { if (y < x) return 3 - 0.5 * x / y; else return 2 + 0.5 * y / x; }
    else { if (x < -y) return 3 - 0.5 * x / y; else return 4 + 0.5 * y / x; }
  }
}

struct P3 { double x, y, z; };
static inline double d3(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// exd_coordize, exact_disk_utils.inl:95-145
static void coordize(P3 mv, P3& m, P3& u, P3& v) {
  double a = 1 / sqrt(d3(mv, mv));
  m = {a * mv.x, a * mv.y, a * mv.z};
  double mx2 = m.x * m.x, my2 = m.y * m.y, mz2 = m.z * m.z;
  if (mx2 > my2) {
    if (mx2 > mz2) {
      if (my2 > mz2) { u = {m.y, -m.x, 0}; a = 1 - mz2; v = {m.z * m.x, m.z * m.y, -a}; }
      else { u = {m.z, 0, -m.x}; a = 1 - my2; v = {-m.y * m.x, a, -m.y * m.z}; }
    } else { u = {-m.z, 0, m.x}; a = 1 - my2; v = {m.y * m.x, -a, m.y * m.z}; }
  } else {
    if (my2 > mz2) {
      if (mx2 > mz2) { u = {-m.y, m.x, 0}; a = 1 - mz2; v = {-m.z * m.x, -m.z * m.y, a}; }
      else { u = {0, m.z, -m.y}; a = 1 - mx2; v = {-a, m.x * m.y, m.x * m.z}; }
    } else { u = {0, -m.z, m.y}; a = 1 - mx2; v = {a, -m.x * m.y, -m.x * m.z}; }
  }
  a = 1 / sqrt(a);
  u = {u.x * a, u.y * a, u.z * a};
  v = {v.x * a, v.y * a, v.z * a};
}

static int g_max_pool = 0;  // largest number of entries one call needed (sizing of the device pool)
struct Disk {
  std::vector<Vtx> V;
  int head = -1, n_edges = 0;
  int add() { V.emplace_back(); return (int)V.size() - 1; }
};

static inline double exd_span(const Vtx& v1, const Vtx& v2, const Vtx& p) {  // calculate_exd_span :507-510
  return (v1.u - p.u) * (v2.v - p.v) - (v2.u - p.u) * (v1.v - p.v);
}
static inline double time_span(const Vtx& v1, const Vtx& v2, const Vtx& p) {  // calculate_time_span :513-516
  return (p.u * v1.v - p.v * v1.u) / (p.v * (v2.u - v1.u) - p.u * (v2.v - v1.v));
}

// One wall's chord of the m = 0 plane, given the wall vertices in (m, u, v) coordinates relative to the collision
// point (exact_disk :975-1085).  Returns -1: target occluded, 0: wall skipped, 1: edge added.
static int add_wall_edge(Disk& D, const P3 vm[3], const Vtx& sm, double R2) {
  auto isect = [](P3 p0, P3 p1, Vtx& out) {  // compute_intersect_w_m0 :211-222 (x = m, y = u, z = v)
    double t = p0.x / (p0.x - p1.x);
    out.u = p0.y + t * (p1.y - p0.y);
    out.v = p0.z + t * (p1.z - p0.z);
  };
  Vtx pa, pb;
  if ((vm[0].x < 0) == (vm[1].x < 0)) {
    if ((vm[2].x < 0) == (vm[1].x < 0)) return 0;
    isect(vm[0], vm[2], pa); isect(vm[1], vm[2], pb);
  } else if ((vm[0].x < 0) == (vm[2].x < 0)) {
    isect(vm[0], vm[1], pa); isect(vm[2], vm[1], pb);
  } else {
    isect(vm[1], vm[0], pa); isect(vm[2], vm[0], pb);
  }
  pa.r2 = pa.u * pa.u + pa.v * pa.v;
  pb.r2 = pb.u * pb.u + pb.v * pb.v;
  if (pa.r2 < EXD_EPS * R2 || pb.r2 < EXD_EPS * R2) return -1;
  if (!exd_distinguishable(pa.u * pb.v, pb.u * pa.v, EXD_EPS) && pa.u * pb.u + pa.v * pb.v < 0) return -1;
  // test_intersect_line_with_circle :225-277
  double t = 0, s = 1;
  if (pa.r2 > R2 || pb.r2 > R2) {
    double pa_pb = pa.u * pb.u + pa.v * pb.v;
    if (!exd_distinguishable(pa.r2 + pb.r2, 2 * pa_pb, EXD_EPS)) {
      if (sm.r2 < pa.r2 && sm.r2 < pb.r2 && exd_distinguishable(sm.r2, pa.r2, EXD_EPS) &&
          exd_distinguishable(sm.r2, pa.r2, EXD_EPS))
        return 0;
      if (!exd_distinguishable(sm.u * pa.v, sm.v * pa.u, EXD_SQRT_EPS) ||
          !exd_distinguishable(sm.u * pb.v, sm.v * pb.u, EXD_SQRT_EPS))
        return -1;
      return 0;
    }
    double a = 1 / (pa.r2 + pb.r2 - 2 * pa_pb);
    double b = (pa_pb - pa.r2) * a;
    double c = (R2 - pa.r2) * a;
    double d = b * b + c;
    if (d <= 0) return 0;
    d = sqrt(d);
    t = -b - d;
    if (t >= 1) return 0;
    if (t < 0) t = 0;
    s = -b + d;
    if (s <= 0) return 0;
    if (s > 1) s = 1;
  }
  // construct_final_endpoints :280-315
  int ia = D.add(), ib = D.add();
  {
    Vtx& A = D.V[ia]; Vtx& B = D.V[ib];
    if (t > 0) { A.u = pa.u + t * (pb.u - pa.u); A.v = pa.v + t * (pb.v - pa.v); A.r2 = A.u * A.u + A.v * A.v; A.zeta = zetize(A.v, A.u); }
    else { A.u = pa.u; A.v = pa.v; A.r2 = pa.r2; A.zeta = zetize(pa.v, pa.u); }
    if (s < 1) { B.u = pa.u + s * (pb.u - pa.u); B.v = pa.v + s * (pb.v - pa.v); B.r2 = B.u * B.u + B.v * B.v; B.zeta = zetize(B.v, B.u); }
    else { B.u = pb.u; B.v = pb.v; B.r2 = pb.r2; B.zeta = zetize(pb.v, pb.u); }
  }
  double a = D.V[ib].zeta - D.V[ia].zeta;
  if (a < 0) a += 4;
  if (a >= 2) { std::swap(ia, ib); a = 4 - a; }
  double b = sm.zeta - D.V[ia].zeta;
  if (b < 0) b += 4;
  if (b < a) {  // blocked reaction: the line is between origin and target
    double au = D.V[ia].u - sm.u, av = D.V[ia].v - sm.v, bu = D.V[ib].u - sm.u, bv = D.V[ib].v - sm.v;
    double c = au * bv - av * bu;
    if (c < 0 || !exd_distinguishable(au * bv, av * bu, EXD_EPS)) return -1;
  }
  D.V[ia].role = ROLE_HEAD; D.V[ib].role = ROLE_TAIL;
  D.V[ia].e = ib; D.V[ib].e = -1;
  D.V[ib].next = D.head; D.V[ia].next = ib; D.head = ia;
  D.n_edges++;
  return 1;
}

// calculate_area_for_multiple_edges, exact_disk_utils.inl:517-818
static double area_multiple_edges(Disk& D, double R2) {
  std::vector<Vtx>& V = D.V;
  // two scratch vertices play the reference's stack objects pa / pb
  const int PA = D.add(), PB = D.add();
  int vp = V[D.head].next, ppa = D.head, ppb = D.head;
  V[ppa].next = -1; V[ppa].span = -1;
  while (vp != -1) {  // insertion sort by zeta
    V[vp].span = -1;
    int vq = V[vp].next;
    if (V[vp].zeta < V[ppa].zeta) { V[vp].next = ppa; ppa = vp; }
    else {
      int pqa;
      for (pqa = ppa; V[pqa].next != -1; pqa = V[pqa].next)
        if (V[vp].zeta < V[V[pqa].next].zeta) break;
      V[vp].next = V[pqa].next;
      V[pqa].next = vp;
      if (V[vp].next == -1) ppb = vp;
    }
    vp = vq;
  }
  int vertex_head = ppa;
  V[ppb].next = ppa;  // circular

  // insert points where lines cross
  ppb = -1;
  for (ppa = vertex_head; ppa != vertex_head || ppb == -1; ppa = V[ppa].next) {
    if (V[ppa].role != ROLE_HEAD) continue;
    ppb = V[ppa].e;
    for (int pqa = V[ppa].next; pqa != ppb; pqa = V[pqa].next) {
      if (V[pqa].role != ROLE_HEAD) continue;
      int pqb = V[pqa].e;
      double pau = V[ppb].u - V[ppa].u, pav = V[ppb].v - V[ppa].v;
      double pbu = V[pqb].u - V[pqa].u, pbv = V[pqb].v - V[pqa].v;
      double r = pbu * pav - pau * pbv;
      if (r * r < EXD_EPS * (pau * pau + pav * pav) * (pbu * pbu + pbv * pbv)) {  // parallel: combine
        V[pqa].e = -1; V[pqa].role = ROLE_OTHER;
        double a = V[pqb].zeta - V[ppb].zeta;
        if (a < 0) a += 4;
        if (a > 2) V[pqb].role = ROLE_OTHER;
        else { V[ppa].e = pqb; V[ppb].role = ROLE_OTHER; ppb = pqb; pqa = ppa; }
        continue;
      }
      double s = (V[ppa].u - V[pqa].u) * pav - (V[ppa].v - V[pqa].v) * pau;
      if (s * r <= EXD_EPS * R2 * R2) continue;
      double t = s / r;
      if (t >= 1 - EXD_EPS) continue;
      if (pau * pau > pav * pav) {
        s = (V[pqa].u - V[ppa].u + t * pbu) * pau;
        if (s <= EXD_EPS * R2 || s >= pau * pau * (1 - EXD_EPS)) continue;
      } else {
        s = (V[pqa].v - V[ppa].v + t * pbv) * pav;
        if (s <= EXD_EPS * R2 || s >= pav * pav * (1 - EXD_EPS)) continue;
      }
      int vq = D.add();  // (V may reallocate: no references held across this call)
      V[vq].u = V[pqa].u + t * pbu;
      V[vq].v = V[pqa].v + t * pbv;
      V[vq].r2 = V[vq].u * V[vq].u + V[vq].v * V[vq].v;
      V[vq].zeta = zetize(V[vq].v, V[vq].u);
      V[vq].e = ppb; V[vq].span = -1; V[vq].role = ROLE_CROSS;
      for (vp = ppa; vp != ppb; vp = V[vp].next) {
        double a = V[vq].zeta - V[V[vp].next].zeta;
        if (a > 2) a -= 4; else if (a < -2) a += 4;
        if (a < 0) break;
      }
      V[vq].next = V[vp].next;
      V[vp].next = vq;
      if (V[vq].zeta < V[vertex_head].zeta) vertex_head = vq;
    }
  }

  // collapse nearby points in zeta and R
  int vq;
  for (vp = vertex_head, vq = -1; vq != vertex_head; vp = vq) {
    for (vq = V[vp].next; vq != vertex_head; vq = V[vq].next) {
      if (V[vq].zeta - V[vp].zeta < EXD_EPS) {
        V[vq].zeta = V[vp].zeta;
        if (-EXD_EPS < V[vq].r2 - V[vp].r2 && EXD_EPS > V[vq].r2 - V[vp].r2) V[vq].r2 = V[vp].r2;
      } else break;
    }
  }

  // register all spanning line segments
  vq = -1;
  for (vp = vertex_head; vp != vertex_head || vq == -1; vp = V[vp].next) {
    if (V[vp].role != ROLE_HEAD) continue;
    for (vq = V[vp].next; vq != V[vp].e; vq = V[vq].next) {
      if (!exd_distinguishable(V[vq].zeta, V[vp].zeta, EXD_EPS)) continue;
      if (!exd_distinguishable(V[vq].zeta, V[V[vp].e].zeta, EXD_EPS)) break;
      if (V[vq].role == ROLE_OTHER) continue;
      int vr = D.add();
      V[vr].next = V[vq].span;
      V[vq].span = vr;
      V[vr].e = vp;
      V[vr].zeta = V[vq].zeta;
      V[vr].role = ROLE_SPAN;
    }
  }

  // walk around and accumulate the visible area
  double A = 0, zeta = 0, last_zeta = -1;
  int vs = -1;
  for (vp = vertex_head; zeta < 4 - EXD_EPS; vp = V[vp].next) {
    if (V[vp].role == ROLE_OTHER) continue;
    if (!exd_distinguishable(V[vp].zeta, last_zeta, EXD_EPS)) continue;
    last_zeta = V[vp].zeta;
    int vr = (vs == PA) ? PB : PA;
    V[vr].u = V[vp].u; V[vr].v = V[vp].v; V[vr].zeta = V[vp].zeta;
    if (V[vp].role == ROLE_TAIL) { V[vr].r2 = R2 * (1 + EXD_EPS); V[vr].e = -1; }
    else { V[vr].r2 = V[vp].r2; V[vr].e = V[vp].e; }
    for (vq = V[vp].next; !exd_distinguishable(V[vq].zeta, last_zeta, EXD_EPS); vq = V[vq].next) {
      if (V[vq].role == ROLE_HEAD) {
        if (V[vq].r2 < V[vp].r2 || V[vr].e == -1) {
          V[vr].u = V[vq].u; V[vr].v = V[vq].v; V[vr].r2 = V[vq].r2; V[vr].e = V[vq].e;
        } else if (!exd_distinguishable(V[vq].r2, V[vr].r2, EXD_EPS)) {
          double b = exd_span(V[vr], V[V[vr].e], V[V[vq].e]);
          if (b > 0) V[vr].e = V[vq].e;
        }
      }
    }
    for (vq = V[vp].span; vq != -1; vq = V[vq].next) {
      int qa = V[vq].e, qb = V[qa].e;
      double b = exd_span(V[qa], V[qb], V[vr]);
      double c = b * b;
      if (c < R2 * R2 * EXD_EPS) {  // span crosses the point
        if (V[vr].e == -1) { V[vr].r2 = V[vr].u * V[vr].u + V[vr].v * V[vr].v; V[vr].e = qb; }
        else { b = exd_span(V[vr], V[V[vr].e], V[qb]); if (b > 0) V[vr].e = qb; }
      } else if (b < 0 || V[vr].e == -1) {  // span is inside the point or spans a tail
        double t = time_span(V[qa], V[qb], V[vp]);
        V[vr].u = V[qa].u + t * (V[qb].u - V[qa].u);
        V[vr].v = V[qa].v + t * (V[qb].v - V[qa].v);
        V[vr].r2 = V[vr].u * V[vr].u + V[vr].v * V[vr].v;
        V[vr].e = qb;
      }
    }
    if (vs == -1) vs = vr;
    else {
      double c = V[vr].zeta - V[vs].zeta;
      if (c < 0) c += 4;
      if (c > EXD_EPS) {
        zeta += c;
        if (V[vs].e == -1 || (V[V[vs].e].zeta - V[vs].zeta) * (V[V[vs].e].zeta - V[vs].zeta) < EXD_EPS * EXD_EPS) {
          if (c >= 2) { V[vs].u = -V[vs].u; V[vs].v = -V[vs].v; A += 0.5 * EXD_PI * R2; }
          double a = V[vs].u * V[vr].u + V[vs].v * V[vr].v;
          double b = V[vs].u * V[vr].v - V[vs].v * V[vr].u;
          double s;
          if (a <= 0) s = atan(-a / b) + 0.5 * EXD_PI; else s = atan(b / a);
          A += 0.5 * s * R2;
        } else {
          const Vtx& E = V[V[vs].e];
          if (!exd_distinguishable(E.zeta, V[vr].zeta, EXD_EPS)) A += 0.5 * (V[vs].u * E.v - V[vs].v * E.u);
          else {
            double t = time_span(V[vs], E, V[vr]);
            double b2 = V[vs].u + (E.u - V[vs].u) * t;
            double c2 = V[vs].v + (E.v - V[vs].v) * t;
            A += 0.5 * (V[vs].u * c2 - V[vs].v * b2);
          }
        }
        vs = vr;
      } else if (V[vr].e != -1) vs = vr;
    }
  }
  return A;
}

// The geometry part of exact_disk (:840-1145) over an explicit wall list.
//   loc, mv, R, target: as in the reference;  walls: n triangles, 9 doubles each + {nx, ny, nz, d};
//   skip[w] != 0: the moving species passes through wall w (all its surface-class reactions are transparent).
static double exact_disk(P3 loc, P3 mv, double R, P3 target, bool target_is_loc, int n_walls, const double* tri9,
                         const double* plane4, const unsigned char* skip) {
  Disk D;
  double R2 = R * R;
  double m2_i = 1 / d3(mv, mv);
  P3 m, u, v;
  coordize(mv, m, u, v);
  P3 Lmuv = {d3(loc, m), d3(loc, u), d3(loc, v)};
  Vtx sm;
  if (target_is_loc) sm.u = sm.v = sm.r2 = sm.zeta = 0;
  else {
    P3 td = {target.x - loc.x, target.y - loc.y, target.z - loc.z};
    sm.u = d3(td, u); sm.v = d3(td, v);
    sm.r2 = sm.u * sm.u + sm.v * sm.v;
    sm.zeta = zetize(sm.v, sm.u);
  }
  for (int w = 0; w < n_walls; w++) {
    const double* t = tri9 + 9 * w;
    P3 wv[3] = {{t[0], t[1], t[2]}, {t[3], t[4], t[5]}, {t[6], t[7], t[8]}};
    P3 n = {plane4[4 * w], plane4[4 * w + 1], plane4[4 * w + 2]};
    double l_n = d3(loc, n);
    double d = plane4[4 * w + 3] - l_n;
    double m_n = d3(mv, n);
    if (d * d >= R2 * (1 - m2_i * m_n * m_n)) continue;
    // wall bounding box vs disk bounding box
    P3 llf = wv[0], urb = wv[0];
    for (int k = 1; k < 3; k++) {
      if (wv[k].x < llf.x) llf.x = wv[k].x; else if (wv[k].x > urb.x) urb.x = wv[k].x;
      if (wv[k].y < llf.y) llf.y = wv[k].y; else if (wv[k].y > urb.y) urb.y = wv[k].y;
      if (wv[k].z < llf.z) llf.z = wv[k].z; else if (wv[k].z > urb.z) urb.z = wv[k].z;
    }
    double a, b;
    b = R2 * (1 - mv.x * mv.x * m2_i);
    a = llf.x - loc.x; if (a > 0 && a * a >= b) continue;
    a = loc.x - urb.x; if (a > 0 && a * a >= b) continue;
    b = R2 * (1 - mv.y * mv.y * m2_i);
    a = llf.y - loc.y; if (a > 0 && a * a >= b) continue;
    a = loc.y - urb.y; if (a > 0 && a * a >= b) continue;
    b = R2 * (1 - mv.z * mv.z * m2_i);
    a = llf.z - loc.z; if (a > 0 && a * a >= b) continue;
    a = loc.z - urb.z; if (a > 0 && a * a >= b) continue;
    if (skip && skip[w]) continue;
    P3 vm[3];
    for (int k = 0; k < 3; k++) vm[k] = {d3(wv[k], m) - Lmuv.x, d3(wv[k], u) - Lmuv.y, d3(wv[k], v) - Lmuv.z};
    int r = add_wall_edge(D, vm, sm, R2);
    if (r < 0) return -1;
  }
  if (D.n_edges == 0) return 1;
  if (D.n_edges == 1) {  // :1099-1115
    const Vtx& A0 = D.V[D.head]; const Vtx& B0 = D.V[A0.e];
    double ares = A0.u * B0.u + A0.v * B0.v;
    double bres = A0.u * B0.v - A0.v * B0.u;
    double sres;
    if (ares <= 0) sres = atan(-ares / bres) + 0.5 * EXD_PI; else sres = atan(bres / ares);
    return (0.5 * bres + R2 * (EXD_PI - 0.5 * sres)) / (EXD_PI * R2);
  }
  double A = area_multiple_edges(D, R2);
  if ((int)D.V.size() > g_max_pool) g_max_pool = (int)D.V.size();
  return A / (EXD_PI * R2);
}

}  // namespace orc_exd
