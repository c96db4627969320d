// mcx_exact_disk.cuh — ExactDiskUtils::exact_disk (src4/exact_disk_utils.inl:54-1145) for sm_100a.
//
// Fraction of the interaction disk (radius R, perpendicular to the motion, centred at the collision point) that
// the walls of the collision subpartition leave visible, or -1 when a wall lies between the moving molecule and
// its target.  The reference builds heap-allocated linked lists of chord end points, crossings and "span"
// markers; here they are entries of one fixed pool per thread (structure of arrays in local memory, 16-bit
// links), so the traversal orders — and with them every floating-point result — are the reference's.  The pool is
// touched only when at least one wall really cuts the disk: for the bulk of the collisions every wall is rejected
// by the distance / bounding-box tests and the function returns 1 from registers.
//
// find_boundaries_occluding_disk (:275-505) is not needed: it runs only when use_expanded_list is off, and the
// expanded list is on whenever a volume-volume reaction exists (mcell4_converter.cpp:84-87).
#pragma once

#define EXD_POOL 72          /* 8 walls cutting one disk need 61 entries (measured on random soups, oracle/) */
#define EXD_PI 3.14159265358979323846

enum { EXD_UNDEF = 0, EXD_HEAD, EXD_TAIL, EXD_CROSS, EXD_SPAN, EXD_OTHER };

struct ExdPool {
  double u[EXD_POOL], v[EXD_POOL], r2[EXD_POOL], zeta[EXD_POOL];
  short next[EXD_POOL], e[EXD_POOL], span[EXD_POOL];
  signed char role[EXD_POOL];
  int n, head, n_edges;
  bool overflow;
  __device__ __forceinline__ int add() {
    if (n >= EXD_POOL) { overflow = true; return EXD_POOL - 1; }
    const int k = n++;
    u[k] = v[k] = r2[k] = zeta[k] = 0; next[k] = e[k] = span[k] = -1; role[k] = EXD_UNDEF;
    return k;
  }
};

// exd_zetize, exact_disk_utils.inl:54-80
__device__ __forceinline__ double exd_zetize(double y, double x) {
  if (y >= 0) {
    if (x >= 0) { if (x < y) return 1 - 0.5 * x / y; else return 0.5 * y / x; }
    else { if (-x < y) return 1 - 0.5 * x / y; else return 2 + 0.5 * y / x; }
  } else {
    if (x <= 0) { if (y < x) return 3 - 0.5 * x / y; else return 2 + 0.5 * y / x; }
    else { if (x < -y) return 3 - 0.5 * x / y; else return 4 + 0.5 * y / x; }
  }
}

// exd_coordize, exact_disk_utils.inl:95-145
__device__ void exd_coordize(D3 mv, D3& m, D3& u, D3& v) {
  double a = 1 / sqrt(dot3(mv, mv));
  m = D3{a * mv.x, a * mv.y, a * mv.z};
  const double mx2 = m.x * m.x, my2 = m.y * m.y, mz2 = m.z * m.z;
  if (mx2 > my2) {
    if (mx2 > mz2) {
      if (my2 > mz2) { u = D3{m.y, -m.x, 0}; a = 1 - mz2; v = D3{m.z * m.x, m.z * m.y, -a}; }
      else { u = D3{m.z, 0, -m.x}; a = 1 - my2; v = D3{-m.y * m.x, a, -m.y * m.z}; }
    } else { u = D3{-m.z, 0, m.x}; a = 1 - my2; v = D3{m.y * m.x, -a, m.y * m.z}; }
  } else {
    if (my2 > mz2) {
      if (mx2 > mz2) { u = D3{-m.y, m.x, 0}; a = 1 - mz2; v = D3{-m.z * m.x, -m.z * m.y, a}; }
      else { u = D3{0, m.z, -m.y}; a = 1 - mx2; v = D3{-a, m.x * m.y, m.x * m.z}; }
    } else { u = D3{0, -m.z, m.y}; a = 1 - mx2; v = D3{a, -m.x * m.y, -m.x * m.z}; }
  }
  a = 1 / sqrt(a);
  u = u * a;
  v = v * a;
}

struct ExdPoint { double u, v, r2, zeta; };

// (v1 - p) x (v2 - p) and the crossing time of p's ray with the segment v1 -> v2 (:507-516)
#define EXD_SPAN_CALC(P, i1, i2, ip) \
  (((P).u[i1] - (P).u[ip]) * ((P).v[i2] - (P).v[ip]) - ((P).u[i2] - (P).u[ip]) * ((P).v[i1] - (P).v[ip]))
#define EXD_TIME_CALC(P, i1, i2, ip) \
  (((P).u[ip] * (P).v[i1] - (P).v[ip] * (P).u[i1]) / \
   ((P).v[ip] * ((P).u[i2] - (P).u[i1]) - (P).u[ip] * ((P).v[i2] - (P).v[i1])))

// One wall's chord of the m = 0 plane (exact_disk :975-1085).  vm: wall vertices in (m, u, v) coordinates relative
// to the collision point (x = m, y = u, z = v).  -1: target occluded, 0: wall skipped, 1: edge added.
__device__ int exd_add_wall_edge(ExdPool& P, const D3 vm[3], const ExdPoint& sm, double R2) {
  ExdPoint pa, pb;
  {
    int i0, i1, j0, j1;  // pa = isect(vm[i0], vm[i1]); pb = isect(vm[j0], vm[j1])  (compute_intersect_w_m0 :211-222)
    if ((vm[0].x < 0) == (vm[1].x < 0)) {
      if ((vm[2].x < 0) == (vm[1].x < 0)) return 0;
      i0 = 0; i1 = 2; j0 = 1; j1 = 2;
    } else if ((vm[0].x < 0) == (vm[2].x < 0)) { i0 = 0; i1 = 1; j0 = 2; j1 = 1; }
    else { i0 = 1; i1 = 0; j0 = 2; j1 = 0; }
    double t = vm[i0].x / (vm[i0].x - vm[i1].x);
    pa.u = vm[i0].y + t * (vm[i1].y - vm[i0].y);
    pa.v = vm[i0].z + t * (vm[i1].z - vm[i0].z);
    t = vm[j0].x / (vm[j0].x - vm[j1].x);
    pb.u = vm[j0].y + t * (vm[j1].y - vm[j0].y);
    pb.v = vm[j0].z + t * (vm[j1].z - vm[j0].z);
  }
  pa.r2 = pa.u * pa.u + pa.v * pa.v;
  pb.r2 = pb.u * pb.u + pb.v * pb.v;
  if (pa.r2 < MCX_EPS * R2 || pb.r2 < MCX_EPS * R2) return -1;
  if (!distinguishable_d(pa.u * pb.v, pb.u * pa.v, MCX_EPS) && pa.u * pb.u + pa.v * pb.v < 0) return -1;
  // test_intersect_line_with_circle :225-277
  double t = 0, s = 1;
  if (pa.r2 > R2 || pb.r2 > R2) {
    const double pa_pb = pa.u * pb.u + pa.v * pb.v;
    if (!distinguishable_d(pa.r2 + pb.r2, 2 * pa_pb, MCX_EPS)) {
      if (sm.r2 < pa.r2 && sm.r2 < pb.r2 && distinguishable_d(sm.r2, pa.r2, MCX_EPS) && distinguishable_d(sm.r2, pa.r2, MCX_EPS))
        return 0;
      if (!distinguishable_d(sm.u * pa.v, sm.v * pa.u, MCX_SQRT_EPS) || !distinguishable_d(sm.u * pb.v, sm.v * pb.u, MCX_SQRT_EPS))
        return -1;
      return 0;
    }
    const double a = 1 / (pa.r2 + pb.r2 - 2 * pa_pb);
    const double b = (pa_pb - pa.r2) * a;
    const double c = (R2 - pa.r2) * a;
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
  int ia = P.add(), ib = P.add();
  if (t > 0) {
    P.u[ia] = pa.u + t * (pb.u - pa.u); P.v[ia] = pa.v + t * (pb.v - pa.v);
    P.r2[ia] = P.u[ia] * P.u[ia] + P.v[ia] * P.v[ia]; P.zeta[ia] = exd_zetize(P.v[ia], P.u[ia]);
  } else { P.u[ia] = pa.u; P.v[ia] = pa.v; P.r2[ia] = pa.r2; P.zeta[ia] = exd_zetize(pa.v, pa.u); }
  if (s < 1) {
    P.u[ib] = pa.u + s * (pb.u - pa.u); P.v[ib] = pa.v + s * (pb.v - pa.v);
    P.r2[ib] = P.u[ib] * P.u[ib] + P.v[ib] * P.v[ib]; P.zeta[ib] = exd_zetize(P.v[ib], P.u[ib]);
  } else { P.u[ib] = pb.u; P.v[ib] = pb.v; P.r2[ib] = pb.r2; P.zeta[ib] = exd_zetize(pb.v, pb.u); }
  double a = P.zeta[ib] - P.zeta[ia];
  if (a < 0) a += 4;
  if (a >= 2) { const int tmp = ia; ia = ib; ib = tmp; a = 4 - a; }
  double b = sm.zeta - P.zeta[ia];
  if (b < 0) b += 4;
  if (b < a) {  // the line is between origin and target: blocked
    const double au = P.u[ia] - sm.u, av = P.v[ia] - sm.v, bu = P.u[ib] - sm.u, bv = P.v[ib] - sm.v;
    const double c = au * bv - av * bu;
    if (c < 0 || !distinguishable_d(au * bv, av * bu, MCX_EPS)) return -1;
  }
  P.role[ia] = EXD_HEAD; P.role[ib] = EXD_TAIL;
  P.e[ia] = (short)ib; P.e[ib] = -1;
  P.next[ib] = (short)P.head; P.next[ia] = (short)ib; P.head = ia;
  P.n_edges++;
  return 1;
}

// calculate_area_for_multiple_edges, exact_disk_utils.inl:517-818
__device__ __noinline__ double exd_area_multiple_edges(ExdPool& P, double R2) {
  const int PA = P.add(), PB = P.add();  // the reference's stack vertices pa / pb
  int guard = 0;
  const int GUARD_MAX = 20000;  // (not in the reference) no loop below may spin on a damaged list
  int vp = P.next[P.head], ppa = P.head, ppb = P.head;
  P.next[ppa] = -1; P.span[ppa] = -1;
  while (vp != -1 && ++guard < GUARD_MAX) {  // insertion sort by zeta
    P.span[vp] = -1;
    const int vq = P.next[vp];
    if (P.zeta[vp] < P.zeta[ppa]) { P.next[vp] = (short)ppa; ppa = vp; }
    else {
      int pqa;
      for (pqa = ppa; P.next[pqa] != -1; pqa = P.next[pqa])
        if (P.zeta[vp] < P.zeta[P.next[pqa]]) break;
      P.next[vp] = P.next[pqa];
      P.next[pqa] = (short)vp;
      if (P.next[vp] == -1) ppb = vp;
    }
    vp = vq;
  }
  int vertex_head = ppa;
  P.next[ppb] = (short)ppa;  // circular

  // insert points where lines cross
  ppb = -1;
  for (ppa = vertex_head; (ppa != vertex_head || ppb == -1) && ++guard < GUARD_MAX; ppa = P.next[ppa]) {
    if (P.role[ppa] != EXD_HEAD) continue;
    ppb = P.e[ppa];
    for (int pqa = P.next[ppa]; pqa != ppb && ++guard < GUARD_MAX; pqa = P.next[pqa]) {
      if (P.role[pqa] != EXD_HEAD) continue;
      const int pqb = P.e[pqa];
      const double pau = P.u[ppb] - P.u[ppa], pav = P.v[ppb] - P.v[ppa];
      const double pbu = P.u[pqb] - P.u[pqa], pbv = P.v[pqb] - P.v[pqa];
      const double r = pbu * pav - pau * pbv;
      if (r * r < MCX_EPS * (pau * pau + pav * pav) * (pbu * pbu + pbv * pbv)) {  // parallel: combine
        P.e[pqa] = -1; P.role[pqa] = EXD_OTHER;
        double a = P.zeta[pqb] - P.zeta[ppb];
        if (a < 0) a += 4;
        if (a > 2) P.role[pqb] = EXD_OTHER;
        else { P.e[ppa] = (short)pqb; P.role[ppb] = EXD_OTHER; ppb = pqb; pqa = ppa; }
        continue;
      }
      double s = (P.u[ppa] - P.u[pqa]) * pav - (P.v[ppa] - P.v[pqa]) * pau;
      if (s * r <= MCX_EPS * R2 * R2) continue;
      const double t = s / r;
      if (t >= 1 - MCX_EPS) continue;
      if (pau * pau > pav * pav) {
        s = (P.u[pqa] - P.u[ppa] + t * pbu) * pau;
        if (s <= MCX_EPS * R2 || s >= pau * pau * (1 - MCX_EPS)) continue;
      } else {
        s = (P.v[pqa] - P.v[ppa] + t * pbv) * pav;
        if (s <= MCX_EPS * R2 || s >= pav * pav * (1 - MCX_EPS)) continue;
      }
      const int vq = P.add();
      if (P.overflow) return 0;
      P.u[vq] = P.u[pqa] + t * pbu;
      P.v[vq] = P.v[pqa] + t * pbv;
      P.r2[vq] = P.u[vq] * P.u[vq] + P.v[vq] * P.v[vq];
      P.zeta[vq] = exd_zetize(P.v[vq], P.u[vq]);
      P.e[vq] = (short)ppb; P.span[vq] = -1; P.role[vq] = EXD_CROSS;
      for (vp = ppa; vp != ppb; vp = P.next[vp]) {
        double a = P.zeta[vq] - P.zeta[P.next[vp]];
        if (a > 2) a -= 4; else if (a < -2) a += 4;
        if (a < 0) break;
      }
      P.next[vq] = P.next[vp];
      P.next[vp] = (short)vq;
      if (P.zeta[vq] < P.zeta[vertex_head]) vertex_head = vq;
    }
  }

  // collapse nearby points in zeta and R
  int vq;
  for (vp = vertex_head, vq = -1; vq != vertex_head && ++guard < GUARD_MAX; vp = vq) {
    for (vq = P.next[vp]; vq != vertex_head; vq = P.next[vq]) {
      if (P.zeta[vq] - P.zeta[vp] < MCX_EPS) {
        P.zeta[vq] = P.zeta[vp];
        if (-MCX_EPS < P.r2[vq] - P.r2[vp] && MCX_EPS > P.r2[vq] - P.r2[vp]) P.r2[vq] = P.r2[vp];
      } else break;
    }
  }

  // register all spanning line segments
  vq = -1;
  for (vp = vertex_head; (vp != vertex_head || vq == -1) && ++guard < GUARD_MAX; vp = P.next[vp]) {
    if (P.role[vp] != EXD_HEAD) continue;
    for (vq = P.next[vp]; vq != P.e[vp] && ++guard < GUARD_MAX; vq = P.next[vq]) {
      if (!distinguishable_d(P.zeta[vq], P.zeta[vp], MCX_EPS)) continue;
      if (!distinguishable_d(P.zeta[vq], P.zeta[P.e[vp]], MCX_EPS)) break;
      if (P.role[vq] == EXD_OTHER) continue;
      const int vr = P.add();
      if (P.overflow) return 0;
      P.next[vr] = P.span[vq];
      P.span[vq] = (short)vr;
      P.e[vr] = (short)vp;
      P.zeta[vr] = P.zeta[vq];
      P.role[vr] = EXD_SPAN;
    }
  }

  // walk around and accumulate the visible area
  double A = 0, zeta = 0, last_zeta = -1;
  int vs = -1;
  for (vp = vertex_head; zeta < 4 - MCX_EPS && ++guard < GUARD_MAX; vp = P.next[vp]) {
    if (P.role[vp] == EXD_OTHER) continue;
    if (!distinguishable_d(P.zeta[vp], last_zeta, MCX_EPS)) continue;
    last_zeta = P.zeta[vp];
    const int vr = (vs == PA) ? PB : PA;
    P.u[vr] = P.u[vp]; P.v[vr] = P.v[vp]; P.zeta[vr] = P.zeta[vp];
    if (P.role[vp] == EXD_TAIL) { P.r2[vr] = R2 * (1 + MCX_EPS); P.e[vr] = -1; }
    else { P.r2[vr] = P.r2[vp]; P.e[vr] = P.e[vp]; }
    for (vq = P.next[vp]; !distinguishable_d(P.zeta[vq], last_zeta, MCX_EPS) && ++guard < GUARD_MAX; vq = P.next[vq]) {
      if (P.role[vq] == EXD_HEAD) {
        if (P.r2[vq] < P.r2[vp] || P.e[vr] == -1) {
          P.u[vr] = P.u[vq]; P.v[vr] = P.v[vq]; P.r2[vr] = P.r2[vq]; P.e[vr] = P.e[vq];
        } else if (!distinguishable_d(P.r2[vq], P.r2[vr], MCX_EPS)) {
          const double b = EXD_SPAN_CALC(P, vr, P.e[vr], P.e[vq]);
          if (b > 0) P.e[vr] = P.e[vq];
        }
      }
    }
    for (vq = P.span[vp]; vq != -1; vq = P.next[vq]) {
      const int qa = P.e[vq], qb = P.e[qa];
      double b = EXD_SPAN_CALC(P, qa, qb, vr);
      const double c = b * b;
      if (c < R2 * R2 * MCX_EPS) {  // span crosses the point
        if (P.e[vr] == -1) { P.r2[vr] = P.u[vr] * P.u[vr] + P.v[vr] * P.v[vr]; P.e[vr] = (short)qb; }
        else { b = EXD_SPAN_CALC(P, vr, P.e[vr], qb); if (b > 0) P.e[vr] = (short)qb; }
      } else if (b < 0 || P.e[vr] == -1) {  // span is inside the point or spans a tail
        const double t = EXD_TIME_CALC(P, qa, qb, vp);
        P.u[vr] = P.u[qa] + t * (P.u[qb] - P.u[qa]);
        P.v[vr] = P.v[qa] + t * (P.v[qb] - P.v[qa]);
        P.r2[vr] = P.u[vr] * P.u[vr] + P.v[vr] * P.v[vr];
        P.e[vr] = (short)qb;
      }
    }
    if (vs == -1) vs = vr;
    else {
      double c = P.zeta[vr] - P.zeta[vs];
      if (c < 0) c += 4;
      if (c > MCX_EPS) {
        zeta += c;
        const int ve = P.e[vs];
        if (ve == -1 || (P.zeta[ve] - P.zeta[vs]) * (P.zeta[ve] - P.zeta[vs]) < MCX_EPS * MCX_EPS) {
          if (c >= 2) { P.u[vs] = -P.u[vs]; P.v[vs] = -P.v[vs]; A += 0.5 * EXD_PI * R2; }
          const double a = P.u[vs] * P.u[vr] + P.v[vs] * P.v[vr];
          const double b = P.u[vs] * P.v[vr] - P.v[vs] * P.u[vr];
          double s;
          if (a <= 0) s = atan(-a / b) + 0.5 * EXD_PI; else s = atan(b / a);
          A += 0.5 * s * R2;
        } else {
          if (!distinguishable_d(P.zeta[ve], P.zeta[vr], MCX_EPS)) A += 0.5 * (P.u[vs] * P.v[ve] - P.v[vs] * P.u[ve]);
          else {
            const double t = EXD_TIME_CALC(P, vs, ve, vr);
            const double b2 = P.u[vs] + (P.u[ve] - P.u[vs]) * t;
            const double c2 = P.v[vs] + (P.v[ve] - P.v[vs]) * t;
            A += 0.5 * (P.u[vs] * c2 - P.v[vs] * b2);
          }
        }
        vs = vr;
      } else if (P.e[vr] != -1) vs = vr;
    }
  }
  if (guard >= GUARD_MAX) P.overflow = true;
  return A;
}

// First rejection test of exact_disk (:927-936) over the walls of the collision subpartition: false means every
// wall is farther from the collision point than the disk reaches, i.e. exact_disk would return 1.
__device__ __forceinline__ bool exd_any_wall_in_reach(const DevParams& p, D3 loc, D3 mv) {
  const uint32_t sub = subpart_index(p, loc);
  if (__ldg(p.spw_start + sub) == __ldg(p.spw_start + sub + 1)) return false;
  const double R2 = p.R * p.R;
  const double m2_i = 1 / dot3(mv, mv);
  bool any = false;
  // walls whose bounding box is farther than R from loc fail exact_disk's box test (:905-925) whatever their plane:
  // only the cells under the box loc +- R are looked at (no side effects: duplicates are harmless)
  const double pad = p.R * (1.0 + 1e-9) + MCX_FW_MARGIN;
  const FwRange r = fw_range(p, sub, D3{loc.x - pad, loc.y - pad, loc.z - pad}, D3{loc.x + pad, loc.y + pad, loc.z + pad});
  const int K = p.fw_K;
  for (int z = r.z0; z <= r.z1; z++)
    for (int y = r.y0; y <= r.y1; y++) {
      const uint32_t row = r.base + (uint32_t)((z * K + y) * K);
      const uint32_t e0 = __ldg(p.fw_start + row + r.x0), e1 = __ldg(p.fw_start + row + r.x1 + 1);
      for (uint32_t e = e0; e < e1; e++) {
        const DevWall& f = p.walls[__ldg(p.fw_list + e)];
        const D3 n = {f.nx, f.ny, f.nz};
        const double d = f.dist - dot3(loc, n);
        const double m_n = dot3(mv, n);
        any = any || !(d * d >= R2 * (1 - m2_i * m_n * m_n));
      }
    }
  return any;
}

// exact_disk, exact_disk_utils.inl:840-1145.  err receives MCX_ERR_OVERFLOW when the pool is exhausted.
__device__ __noinline__ double exact_disk(const DevParams& p, D3 loc, D3 mv, uint32_t species, D3 target, int& err) {
  const uint32_t sub = subpart_index(p, loc);
  const uint32_t w0 = p.spw_start[sub], w1 = p.spw_start[sub + 1];
  if (w0 == w1) return 1;
  if (!exd_any_wall_in_reach(p, loc, mv)) return 1;  // registers only: the pool below is never touched
  const double R2 = p.R * p.R;
  const double m2_i = 1 / dot3(mv, mv);

  ExdPool P;
  P.n = 0; P.head = -1; P.n_edges = 0; P.overflow = false;
  D3 m, u, v;
  exd_coordize(mv, m, u, v);
  const D3 Lmuv = {dot3(loc, m), dot3(loc, u), dot3(loc, v)};
  ExdPoint sm;
  if (!distinguishable_vec3_d(loc, target, MCX_EPS)) sm.u = sm.v = sm.r2 = sm.zeta = 0;
  else {
    const D3 td = target - loc;
    sm.u = dot3(td, u); sm.v = dot3(td, v);
    sm.r2 = sm.u * sm.u + sm.v * sm.v;
    sm.zeta = exd_zetize(sm.v, sm.u);
  }
  // the walls of the collision subpartition in list order; those outside the box loc +- R fail the box test below
  WallWalk ww;
  {
    const double pad = p.R * (1.0 + 1e-9) + MCX_FW_MARGIN;
    ww.init(p, sub, D3{loc.x - pad, loc.y - pad, loc.z - pad}, D3{loc.x + pad, loc.y + pad, loc.z + pad});
  }
  for (uint32_t wi = ww.next(p); wi != MCX_NONE; wi = ww.next(p)) {
    const DevWall& f = p.walls[wi];
    const D3 n = {f.nx, f.ny, f.nz};
    const double l_n = dot3(loc, n);
    const double d = f.dist - l_n;
    const double m_n = dot3(mv, n);
    if (d * d >= R2 * (1 - m2_i * m_n * m_n)) continue;
    const D3 wv[3] = {wall_vertex(p, wi, 0), wall_vertex(p, wi, 1), wall_vertex(p, wi, 2)};
    D3 llf = wv[0], urb = wv[0];
#pragma unroll
    for (int q = 1; q < 3; q++) {
      if (wv[q].x < llf.x) llf.x = wv[q].x; else if (wv[q].x > urb.x) urb.x = wv[q].x;
      if (wv[q].y < llf.y) llf.y = wv[q].y; else if (wv[q].y > urb.y) urb.y = wv[q].y;
      if (wv[q].z < llf.z) llf.z = wv[q].z; else if (wv[q].z > urb.z) urb.z = wv[q].z;
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
    // walls the moving molecule travels through are ignored (:957-975)
    const uint32_t wclass = p.wall_class[wi];
    if (wclass != MCX_NONE && p.exd_skip[species * p.n_surf_classes + wclass]) continue;
    D3 vm[3];
#pragma unroll
    for (int q = 0; q < 3; q++) vm[q] = D3{dot3(wv[q], m) - Lmuv.x, dot3(wv[q], u) - Lmuv.y, dot3(wv[q], v) - Lmuv.z};
    if (P.n + 2 > EXD_POOL) { err = MCX_ERR_OVERFLOW; return 1; }
    const int r = exd_add_wall_edge(P, vm, sm, R2);
    if (r < 0) return -1;
  }
  if (P.n_edges == 0) return 1;
  if (P.n_edges == 1) {  // :1099-1115
    const int ia = P.head, ib = P.e[ia];
    const double ares = P.u[ia] * P.u[ib] + P.v[ia] * P.v[ib];
    const double bres = P.u[ia] * P.v[ib] - P.v[ia] * P.u[ib];
    double sres;
    if (ares <= 0) sres = atan(-ares / bres) + 0.5 * EXD_PI; else sres = atan(bres / ares);
    return (0.5 * bres + R2 * (EXD_PI - 0.5 * sres)) / (EXD_PI * R2);
  }
  const double A = exd_area_multiple_edges(P, R2);
  if (P.overflow) { err = MCX_ERR_OVERFLOW; return 1; }
  return A / (EXD_PI * R2);
}
