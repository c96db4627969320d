// mcx_kernels.cu — the per-iteration kernel pipeline of libmcx (sm_100a).
//
//   k_diffuse_fast every live molecule: the common no-wall / no-partner whole step, converged control flow;
//                  writes the result to B and bins it for the next snapshot (histogram atomics -> rank)
//   k_diffuse_slow the deferred rest (walls, collisions, split steps): generic evaluation; claiming events
//                  become proposals
//   k_resolve/k_retry  synchronous conflict-resolution rounds over the (small) pending list
//   k_scan_*       exclusive scan of the cell histogram -> cell_start of the next snapshot
//   k_scatter      counting-sort scatter B -> A (drops consumed molecules: folds
//                  SortMolsBySubpartEvent + DefragmentationEvent into every iteration)
//
// All population sizes live in device memory (Counters); kernels are grid-stride over them so that an
// iteration needs no host round trip.
#include <cstdlib>
#include <cstdio>
#include <cooperative_groups.h>
#include "mcx_device.cuh"
#include "mcx_tile.cuh"

#define TPB 256

__device__ __forceinline__ unsigned long long claim_key(unsigned int epoch, uint32_t id) {
  return ((unsigned long long)epoch << 32) | (unsigned long long)(~id);
}
// A surface molecule that only takes a new tile claims itself and the tile WEAKLY: any reaction that consumes it, or that
// needs the tile for its initiator, comes first (ids stay below 2^31, so bit 31 of ~id is set in every strong key).
// Otherwise a reaction whose partner happens to move in the same iteration would be dropped half of the time — the
// partner's own move would win the partner — and the re-evaluated initiator draws a new, independent reaction test.
__device__ __forceinline__ unsigned long long weak_key(unsigned long long key) { return key & ~0x80000000ull; }
// ---- warp-aggregated atomics: lanes of the (possibly divergent) warp that target the same address combine
// into one atomic.  Counters and list cursors are hit by every committing thread; without this the L2 atomic
// unit serialises them (profiles/r01_a: k_resolve 0.5 ms for 1.4e5 proposals).
__device__ __forceinline__ void agg_add(unsigned long long* addr, unsigned int v) {
  const unsigned int peers = __match_any_sync(__activemask(), (unsigned long long)addr);
  const unsigned int total = __reduce_add_sync(peers, v);
  if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(addr, (unsigned long long)total);
}
__device__ __forceinline__ void agg_sub(unsigned long long* addr, unsigned int v) {
  const unsigned int peers = __match_any_sync(__activemask(), (unsigned long long)addr);
  const unsigned int total = __reduce_add_sync(peers, v);
  if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(addr, (unsigned long long)(-(long long)total));
}
// reserve `v` consecutive indices at *addr for this lane; returns the first
__device__ __forceinline__ unsigned int agg_reserve(unsigned int* addr, unsigned int v) {
  const unsigned int peers = __match_any_sync(__activemask(), (unsigned long long)addr);
  const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
  // exclusive prefix of v over the peers below this lane
  unsigned int before = 0, total = 0;
  for (unsigned int rest = peers; rest; rest &= rest - 1) {
    const int src = __ffs(rest) - 1;
    const unsigned int x = __shfl_sync(peers, v, src);
    if (src < lane) before += x;
    total += x;
  }
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(addr, total);
  base = __shfl_sync(peers, base, leader);
  return base + before;
}


// ---- per-block tallies of the counters every committed event touches ------------------------------------------------
// k_resolve commits ~1.3e6 events per iteration at 1e8 molecules and each one used to issue 6-8 warp-aggregated
// global atomics (a __match_any_sync each) on a handful of addresses.  The block counts in shared memory instead
// and flushes once; commit_event falls back to the global atomics when no tally is given (forced commits of k_retry).
struct BlockTally {
  int species[MCX_MAX_COUNTED];            // signed change of the per-species population
  unsigned int rxn[MCX_MAX_COUNTED];       // per reaction rule
  unsigned int bimol, unimol, products, absorptions;
};
__device__ __forceinline__ void tally_clear(BlockTally* t) {
  for (int k = threadIdx.x; k < MCX_MAX_COUNTED; k += blockDim.x) { t->species[k] = 0; t->rxn[k] = 0; }
  if (threadIdx.x == 0) { t->bimol = 0; t->unimol = 0; t->products = 0; t->absorptions = 0; }
  __syncthreads();
}
__device__ __forceinline__ void tally_flush(const BlockTally* t, Counters* c) {
  __syncthreads();
  for (int k = threadIdx.x; k < MCX_MAX_COUNTED; k += blockDim.x) {
    if (t->species[k]) atomicAdd(&c->species_count[k], (unsigned long long)(long long)t->species[k]);
    if (t->rxn[k]) atomicAdd(&c->rxn_count[k], (unsigned long long)t->rxn[k]);
  }
  if (threadIdx.x == 0) {
    if (t->bimol) atomicAdd(&c->bimol_rxns, (unsigned long long)t->bimol);
    if (t->unimol) atomicAdd(&c->unimol_rxns, (unsigned long long)t->unimol);
    if (t->products) atomicAdd(&c->products, (unsigned long long)t->products);
    if (t->absorptions) atomicAdd(&c->absorptions, (unsigned long long)t->absorptions);
  }
}
// one shared atomic per converged group of lanes for a counter they all hit
__device__ __forceinline__ void tally_inc(unsigned int* ctr) {
  const unsigned int m = __activemask();
  if ((threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(ctr, (unsigned int)__popc(m));
}


// ---- per-warp staging of slot indices appended to a global list --------------------------------------------------
// Every warp of k_diffuse_fast used to reserve its deferred slots with one RETURNING atomicAdd on a single
// address (Counters::n_slow): 2.4e5 same-address round trips per iteration at 1e7 molecules, 55 % of the
// kernel's stall samples (profiles/r01_d).  Slots are now staged in a 64-entry shared-memory buffer per warp and
// flushed with one atomic per >32 entries.  push()/flush() are warp-collective: all 32 lanes must call them.
#define WL_CAP 64
struct WarpList {
  uint32_t* buf;       // this warp's WL_CAP entries of shared memory
  unsigned int n;      // staged entries (warp-uniform)
  unsigned int total;  // entries appended so far (warp-uniform)
  __device__ __forceinline__ void init(uint32_t* b) { buf = b; n = 0; total = 0; }
  __device__ __forceinline__ void flush(unsigned int* ctr, uint32_t* list) {
    if (n == 0) return;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    unsigned int at = 0;
    if (lane == 0) at = atomicAdd(ctr, n);
    at = __shfl_sync(0xffffffffu, at, 0);
#pragma unroll 1
    for (unsigned int k = lane; k < n; k += 32) list[at + k] = buf[k];
    __syncwarp();
    total += n;
    n = 0;
  }
  __device__ __forceinline__ void push(bool pred, uint32_t v, unsigned int* ctr, uint32_t* list) {
    const unsigned int bal = __ballot_sync(0xffffffffu, pred);
    if (bal == 0) return;
    const int lane = threadIdx.x & 31;
    if (pred) buf[n + __popc(bal & ((1u << lane) - 1u))] = v;
    n += __popc(bal);
    if (n > WL_CAP - 32) flush(ctr, list);
  }
};

__device__ __forceinline__ unsigned int round_epoch(const DevParams& p, unsigned int round) {
  return (unsigned int)(p.iteration * (unsigned long long)(p.max_rounds + 1) + round + 1);
}

__device__ __forceinline__ void flush_stats(const DevParams& p, const LocalStats& ls, unsigned int msteps) {
  // warp-aggregate, one atomic per warp per counter
  unsigned int v[7] = {ls.ray_polygon_tests, ls.ray_polygon_colls, ls.reflections, ls.transparent,
                       ls.volvol_collisions, ls.redos, msteps};
#pragma unroll
  for (int k = 0; k < 7; k++) {
    unsigned int s = v[k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    v[k] = s;
  }
  if ((threadIdx.x & 31) == 0) {
    Counters* c = p.ctr;
    if (v[0]) atomicAdd(&c->ray_polygon_tests, (unsigned long long)v[0]);
    if (v[1]) atomicAdd(&c->ray_polygon_colls, (unsigned long long)v[1]);
    if (v[2]) atomicAdd(&c->reflections, (unsigned long long)v[2]);
    if (v[3]) atomicAdd(&c->transparent, (unsigned long long)v[3]);
    if (v[4]) atomicAdd(&c->volvol_collisions, (unsigned long long)v[4]);
    if (v[5]) atomicAdd(&c->redos, (unsigned long long)v[5]);
    if (v[6]) atomicAdd(&c->molecule_steps, (unsigned long long)v[6]);
  }
}

__device__ __forceinline__ void raise_error(const DevParams& p, int err, uint32_t id) {
  if (atomicCAS(&p.ctr->error, 0, err) == 0) p.ctr->error_id = id;
}

// result record of a molecule that stays alive: write to B and bin it for the next snapshot
// surf: where a surface molecule's cold fields come from — nullptr: unchanged (copy A -> B); an Outcome that moved:
// its new wall / tile / uv; SURF_FIELDS_IN_B: the proposal already wrote them to B
#define SURF_FIELDS_IN_B ((const Outcome*)1)
__device__ __forceinline__ void finalize_alive(const DevParams& p, uint32_t slot, D3 pos, uint32_t id, uint32_t species,
                                               uint32_t flags, double t_now, double unimol_time, const Outcome* surf = nullptr) {
  if (!owned_z(p, pos.z)) { p.rank[slot] = MCX_NONE; return; }  // multi-GPU: the rank owning the new position keeps it
  uint32_t sf = species | (flags & ~(DF_HAS_UNIMOL | DF_DEAD));
  if (unimol_time != MCX_TIME_INVALID) { sf |= DF_HAS_UNIMOL; p.tuniB[slot] = unimol_time; }
  if (sf & DF_PARTIAL) p.tschedB[slot] = t_now;
  if (sf & DF_SURF) {
    if (surf == nullptr || (surf != SURF_FIELDS_IN_B && !surf->surf_moved)) {
      p.swallB[slot] = p.swallA[slot]; p.stileB[slot] = p.stileA[slot]; p.suvB[slot] = p.suvA[slot];
    } else if (surf != SURF_FIELDS_IN_B) {
      p.swallB[slot] = surf->s_wall; p.stileB[slot] = surf->s_tile; p.suvB[slot] = make_double2(surf->s_u, surf->s_v);
    }
  }
  store_rec(p.recB, slot, pos, id, sf);
  uint32_t cell = cell_of(p, pos.x, pos.y, pos.z);
  p.rank[slot] = atomicAdd(&p.cs_next[cell], 1u);
}

__device__ __forceinline__ bool partner_is_consumed(const DevParams& p, int kind, int rxn_class, int pathway,
                                                    uint32_t self_species) {
  if (kind != MCX_OUT_REACTED) return false;
  const DevClass& c = p.classes[rxn_class];
  const DevPathway& pw = p.pathways[c.first_pathway + pathway];
  bool a_is_r0 = self_species == c.r0;
  return !((pw.keep_mask >> (a_is_r0 ? 1 : 0)) & 1u);
}

// does the claiming event place surface products on vacant neighbour tiles (DevPathway::general)?
__device__ __forceinline__ bool event_is_general(const DevParams& p, int kind, int rxn_class, int pathway) {
  if (!p.prop_pmask || (kind != MCX_OUT_REACTED && kind != MCX_OUT_UNIMOL)) return false;
  const DevClass& c = p.classes[rxn_class];
  return p.pathways[c.first_pathway + pathway].general != 0;
}

// cell group of a position: 16 x-cells of one cell row (the unit of the fresh-id order, FreshEvent)
__device__ __forceinline__ uint32_t group_of(const DevParams& p, D3 q) {
  const int cx = cell_coord(q.x, p.cgx, p.cell_rcp_x, p.ncx), cy = cell_coord(q.y, p.cgy, p.cell_rcp_y, p.ncy), cz = cell_z(p, q.z);
  return row_index(p, cy, cz) * p.grp_x + (uint32_t)(cx >> 4);
}

// accepted claiming event: consume reactants, create products, count
__device__ void commit_event(const DevParams& p, uint32_t slot, int kind, int rxn_class, int pathway, uint32_t partner_slot,
                             double t_event, D3 pos, uint32_t id, uint32_t species, uint32_t flags, double t_now,
                             double unimol_time, uint32_t orient_bits, BlockTally* bt = nullptr) {
  Counters* c = p.ctr;
  // multi-GPU: halo molecules are evaluated redundantly (identically on both sides); an event is counted and its
  // products are created by the rank owning the event position; reactants are marked DEAD everywhere
  const bool own_event = owned_z(p, pos.z);
  const bool track = p.world == 1;  // incremental species counts (multi-GPU recounts during the scatter)
  if (kind == MCX_OUT_SURFMOVE) {  // the mover won its new tile: it stays alive there (cold fields already in B)
    finalize_alive(p, slot, pos, id, species, flags, t_now, unimol_time, SURF_FIELDS_IN_B);
    return;
  }
  if (kind == MCX_OUT_ABSORBED) {
    atomicOr(&p.recA[slot].sf, DF_DEAD);
    if (own_event) { if (bt) tally_inc(&bt->absorptions); else agg_add(&c->absorptions, 1u); }
    if (track) { if (bt) atomicSub(&bt->species[species], 1); else agg_sub(&c->species_count[species], 1u); }
    return;
  }
  const DevClass& cl = p.classes[rxn_class];
  const DevPathway& pw = p.pathways[cl.first_pathway + pathway];
  if (kind == MCX_OUT_WALLRXN) {
    // Standard reaction with a reactive surface (outcome_intersect :1916-1988; partner_slot carries the wall): volume
    // products at the hit point, bumped 2*16*EPS off the wall to the side their orientation names, with the counted
    // volume of that side and the tile under the hit point remembered; the molecule is consumed, or kept — then it
    // waits on its own side or behind the wall (RX_FLIP) for the rest of its step
    const uint32_t wi = partner_slot;
    const DevWall& fw = p.walls[wi];
    const DevGrid& g = p.grids[wi];
    if (own_event) { if (bt) atomicAdd(&bt->rxn[pw.rule_id & (MCX_MAX_COUNTED - 1u)], 1u); else agg_add(&c->rxn_count[pw.rule_id & (MCX_MAX_COUNTED - 1u)], 1u); }
    if (own_event && p.wall_cv) agg_add(&p.rxn_count_cv[(pw.rule_id & (MCX_MAX_COUNTED - 1u)) * p.n_cv + (flags >> SF_CVI_SHIFT)], 1u);
    if (own_event) { if (bt) tally_inc(&bt->bimol); else agg_add(&c->bimol_rxns, 1u); }
    const bool keep = pw.keep_mask & 1u;
    const double hu = pos.x * fw.ux + pos.y * fw.uy + pos.z * fw.uz - g.vert0_u;   // GeometryUtils::xyz2uv
    const double hv = pos.x * fw.vx + pos.y * fw.vy + pos.z * fw.vz - g.vert0_v;
    const uint32_t hit_tile = uv2grid(p, wi, hu, hv);
    if (!keep) {
      atomicOr(&p.recA[slot].sf, DF_DEAD);
      if (track) { if (bt) atomicSub(&bt->species[species], 1); else agg_sub(&c->species_count[species], 1u); }
    }
    const uint32_t n_new = own_event ? pw.n_products : 0u;
    const uint32_t n_reuse = keep ? 0u : 1u;
    const uint32_t first_slot = n_new ? c->n_slots + agg_reserve(&c->n_prod, n_new) : 0u;
    if (n_new > n_reuse && first_slot + n_new <= p.capacity) {
      const uint32_t e = agg_reserve(&c->n_fresh_events, 1u);
      if (e >= p.fresh_cap) raise_error(p, MCX_ERR_CAPACITY, id);
      else {
        const uint32_t grp = group_of(p, pos);
        const uint32_t nf = n_new - n_reuse;
        FreshEvent ev; ev.first_slot = first_slot + n_reuse; ev.n = nf; ev.init_id = id; ev.group = grp;
        ev.next = atomicExch(&p.fresh_head[grp], e);
        p.fresh_list[e] = ev;
        atomicAdd(&p.fresh_pref[grp], nf);
        atomicAdd(&c->n_fresh_ids, nf);
      }
    }
    for (uint32_t k = 0; k < n_new; k++) {
      const uint32_t ns = first_slot + k;
      if (ns >= p.capacity) { raise_error(p, MCX_ERR_CAPACITY, id); return; }
      int o = pw.prod_orient[k];
      if (o == 0) o = ((orient_bits >> k) & 1u) ? 1 : -1;
      const double bump = (o > 0) ? 16 * MCX_EPS : -16 * MCX_EPS;
      const D3 ppos = {pos.x + (2 * bump) * fw.nx, pos.y + (2 * bump) * fw.ny, pos.z + (2 * bump) * fw.nz};
      uint32_t pflags = DF_SCHED_UNIMOL | DF_PARTIAL;
      if (p.has_surf) { pflags |= DF_CREATED_ON_SURF; p.swallB[ns] = wi; p.stileB[ns] = hit_tile; }
      if (p.wall_cv) {
        const uint32_t cv = __ldg(p.wall_cv + wi);
        uint32_t pc = o > 0 ? (cv & 0xFFu) : (cv >> 8);
        if (cv_uses_xor(p, wi)) {   // the initiator's set, its object toggled for a product behind the wall
          const bool front = (orient_bits & ORIENT_BIT_FRONT) != 0;
          pc = (o > 0) == front ? (flags >> SF_CVI_SHIFT) : cv_cross(p, flags >> SF_CVI_SHIFT, wi, front);
          if (pc == MCX_NONE) { raise_error(p, MCX_ERR_STATE, id); pc = flags >> SF_CVI_SHIFT; }
        }
        pflags |= pc << SF_CVI_SHIFT;
      }
      const uint32_t psp = pw.products[k];
      p.tschedB[ns] = t_event;
      store_rec(p.recB, ns, ppos, (k == 0 && !keep) ? id : MCX_NONE, psp | pflags);
      p.rank[ns] = atomicAdd(&p.cs_next[cell_of(p, ppos.x, ppos.y, ppos.z)], 1u);
      if (track) { if (bt) atomicAdd(&bt->species[psp], 1); else agg_add(&c->species_count[psp], 1u); }
      if (bt) tally_inc(&bt->products); else agg_add(&c->products, 1u);
    }
    if (keep) {
      int ko = 0;
      if (pw.kept_info & MCX_KEPT_VALID) { ko = kept_code(pw, 0); if (ko == 0) ko = ((orient_bits >> 4) & 1u) ? 1 : -1; }
      const bool flip = ko != 0 && cl.geom0 != ko;
      const int coll_side = (orient_bits & ORIENT_BIT_FRONT) ? 1 : -1;
      const int side = flip ? -coll_side : coll_side;
      uint32_t f = flags | DF_PARTIAL;
      if (flip && p.wall_cv) {
        uint32_t nc = cv_cross(p, flags >> SF_CVI_SHIFT, wi, coll_side > 0);
        if (nc == MCX_NONE) { raise_error(p, MCX_ERR_STATE, id); nc = flags >> SF_CVI_SHIFT; }
        f = (f & ~SF_CVI_MASK) | (nc << SF_CVI_SHIFT);
      }
      const double bump = (side > 0) ? 16 * MCX_EPS : -16 * MCX_EPS;
      const D3 kpos = {pos.x + (2 * bump) * fw.nx, pos.y + (2 * bump) * fw.ny, pos.z + (2 * bump) * fw.nz};
      if (p.has_surf) { f |= DF_CREATED_ON_SURF; p.swallB[slot] = wi; p.stileB[slot] = MCX_KEPT_AT_WALL; }
      finalize_alive(p, slot, kpos, id, species, f, t_event, unimol_time);
    }
    return;
  }
  if (kind == MCX_OUT_REACTED && cl.kind == MCX_RXN_BIMOL_SURFSURF) {
    // Two surface molecules (react_2D_all_neighbors -> outcome_bimolecular -> outcome_products_random): the initiator
    // sits where its move took it (surface fields already in B), the partner on its tile of the snapshot.  Surface
    // products take the tiles the consumed reactants free, at the uv of the reactant that left (:2826-2846); volume
    // products start at the position of the rule's first reactant, bumped off the INITIATOR's wall to the side their
    // orientation names and remembered with its tile (:2745-2762)
    const uint32_t iw = p.swallB[slot], itile = p.stileB[slot];
    const double2 iuv = p.suvB[slot];
    if (own_event) { if (bt) atomicAdd(&bt->rxn[pw.rule_id & (MCX_MAX_COUNTED - 1u)], 1u); else agg_add(&c->rxn_count[pw.rule_id & (MCX_MAX_COUNTED - 1u)], 1u); }
    if (own_event && p.wall_rs) agg_add(&p.rxn_count_rs[(pw.rule_id & (MCX_MAX_COUNTED - 1u)) * p.n_rs + __ldg(p.wall_rs + iw)], 1u);
    if (own_event) { if (bt) tally_inc(&bt->bimol); else agg_add(&c->bimol_rxns, 1u); }
    const MolRec pr = load_rec_volatile(p.recA, partner_slot);
    const uint32_t pw_wall = p.swallA[partner_slot], pw_tile = p.stileA[partner_slot];
    const double2 puv = p.suvA[partner_slot];
    const bool a_is_r0 = species == cl.r0;
    const bool keepA = (pw.keep_mask >> (a_is_r0 ? 0 : 1)) & 1u, keepB = (pw.keep_mask >> (a_is_r0 ? 1 : 0)) & 1u;
    const int oi = (flags & DF_ORIENT_UP) ? 1 : -1, op = (pr.sf & DF_ORIENT_UP) ? 1 : -1;
    const int match = surfsurf_match(cl, a_is_r0 ? oi : op, a_is_r0 ? op : oi);
    uint32_t reuse[2]; int n_reuse = 0;
    if (!keepA) {
      atomicOr(&p.recA[slot].sf, DF_DEAD);
      if (track) { if (bt) atomicSub(&bt->species[species], 1); else agg_sub(&c->species_count[species], 1u); }
      reuse[n_reuse++] = id;
    }
    if (!keepB) {
      const uint32_t old = atomicOr(&p.recA[partner_slot].sf, DF_DEAD);
      atomicOr(&p.recB[partner_slot].sf, DF_DEAD);
      if (track) { if (bt) atomicSub(&bt->species[old & SF_SPECIES_MASK], 1); else agg_sub(&c->species_count[old & SF_SPECIES_MASK], 1u); }
      reuse[n_reuse++] = pr.id;
      if (p.trace && pr.id < p.n_trace) p.trace[pr.id].outcome = MCX_OUT_CONSUMED;
    }
    // freed tiles in the order of the rule's reactants: 0 = initiator's, 1 = partner's site
    int freed[2]; int n_freed = 0;
    const bool keep0 = pw.keep_mask & 1u, keep1 = (pw.keep_mask & 2u) != 0;
    if (!keep0) freed[n_freed++] = a_is_r0 ? 0 : 1;
    if (!keep1) freed[n_freed++] = a_is_r0 ? 1 : 0;
    uint32_t first_surf = MCX_NONE;
    for (uint32_t k = 0; k < pw.n_products && first_surf == MCX_NONE; k++) if (!(p.species[pw.products[k]].flags & MCX_SP_VOL)) first_surf = k;
    const uint32_t n_new = own_event ? pw.n_products : 0u;
    const uint32_t first_slot = n_new ? c->n_slots + agg_reserve(&c->n_prod, n_new) : 0u;
    if (n_new > (uint32_t)n_reuse && first_slot + n_new <= p.capacity) {
      const uint32_t e = agg_reserve(&c->n_fresh_events, 1u);
      if (e >= p.fresh_cap) raise_error(p, MCX_ERR_CAPACITY, id);
      else {
        const uint32_t g = group_of(p, pos);
        const uint32_t nf = n_new - (uint32_t)n_reuse;
        FreshEvent ev; ev.first_slot = first_slot + (uint32_t)n_reuse; ev.n = nf; ev.init_id = id; ev.group = g;
        ev.next = atomicExch(&p.fresh_head[g], e);
        p.fresh_list[e] = ev;
        atomicAdd(&p.fresh_pref[g], nf);
        atomicAdd(&c->n_fresh_ids, nf);
      }
    }
    const DevWall& fi = p.walls[iw];
    uint32_t n_placed = 0;
    for (uint32_t k = 0; k < n_new; k++) {
      const uint32_t ns = first_slot + k;
      if (ns >= p.capacity) { raise_error(p, MCX_ERR_CAPACITY, id); return; }
      const uint32_t psp = pw.products[k];
      int o = pw.prod_orient[k];
      if (o == 0) o = ((orient_bits >> k) & 1u) ? 1 : -1; else o *= match;
      uint32_t pflags = DF_SCHED_UNIMOL | DF_PARTIAL;
      D3 ppos;
      if (!(p.species[psp].flags & MCX_SP_VOL)) {
        const bool swap = (orient_bits & SURFSURF_SWAP) != 0;
        int which = ((k == first_surf) == swap) ? 1 : 0;
        if (which > n_freed - 1) which = n_freed - 1;
        const bool at_init = n_freed == 0 || freed[which < 0 ? 0 : which] == 0;
        uint32_t tw = at_init ? iw : pw_wall, tt = at_init ? itile : pw_tile;
        double2 tuv = at_init ? iuv : puv;
        if (pw.general && p.prop_pmask) {  // where place_general put the created surface product
          const uint2 wt = p.prop_ptile[slot * MCX_MAX_PRODUCTS + n_placed];
          tw = wt.x; tt = wt.y; tuv = p.prop_puv[slot * MCX_MAX_PRODUCTS + n_placed];
          n_placed++;
        }
        p.swallB[ns] = tw; p.stileB[ns] = tt; p.suvB[ns] = tuv;
        const DevWall& f = p.walls[tw];
        ppos = D3{tuv.x * f.ux + tuv.y * f.vx + f.v0x, tuv.x * f.uy + tuv.y * f.vy + f.v0y, tuv.x * f.uz + tuv.y * f.vz + f.v0z};
        pflags |= DF_SURF | (o > 0 ? DF_ORIENT_UP : 0u);
      } else {
        const uint32_t rw = a_is_r0 ? iw : pw_wall;       // the rule's first reactant
        const double2 ruv = a_is_r0 ? iuv : puv;
        const DevWall& f = p.walls[rw];
        const D3 from = {ruv.x * f.ux + ruv.y * f.vx + f.v0x, ruv.x * f.uy + ruv.y * f.vy + f.v0y, ruv.x * f.uz + ruv.y * f.vz + f.v0z};
        const double bump = (o > 0) ? 16 * MCX_EPS : -16 * MCX_EPS;
        ppos = D3{from.x + (2 * bump) * fi.nx, from.y + (2 * bump) * fi.ny, from.z + (2 * bump) * fi.nz};
        pflags |= DF_CREATED_ON_SURF;
        p.swallB[ns] = iw; p.stileB[ns] = itile;
        if (p.wall_cv) {
          const uint32_t cv = __ldg(p.wall_cv + iw);
          const uint32_t pc = o > 0 ? (cv & 0xFFu) : (cv >> 8);
          if (cv_uses_xor(p, iw)) pflags |= DF_CVI_PENDING;  // no volume reactant to go by: a ray cast at its first evaluation
          pflags |= pc << SF_CVI_SHIFT;
        }
      }
      p.tschedB[ns] = t_event;
      store_rec(p.recB, ns, ppos, (int)k < n_reuse ? reuse[k] : MCX_NONE, psp | pflags);
      p.rank[ns] = atomicAdd(&p.cs_next[cell_of(p, ppos.x, ppos.y, ppos.z)], 1u);
      if (track) { if (bt) atomicAdd(&bt->species[psp], 1); else agg_add(&c->species_count[psp], 1u); }
      if (bt) tally_inc(&bt->products); else agg_add(&c->products, 1u);
    }
    if (keepA) {  // stays on the tile it moved to, its step used up; takes its product-side orientation (:2706-2709)
      uint32_t f = flags;
      if (pw.kept_info & MCX_KEPT_VALID) {
        const int r = a_is_r0 ? 0 : 1;
        int ko = kept_code(pw, r);
        if (ko == 0) ko = ((orient_bits >> (4 + r)) & 1u) ? 1 : -1; else ko *= match;
        f = (f & ~DF_ORIENT_UP) | (ko > 0 ? DF_ORIENT_UP : 0u);
      }
      finalize_alive(p, slot, pos, id, species, f, t_now, unimol_time, SURF_FIELDS_IN_B);
    }
    return;
  }
  if (own_event) { if (bt) atomicAdd(&bt->rxn[pw.rule_id & (MCX_MAX_COUNTED - 1u)], 1u); else agg_add(&c->rxn_count[pw.rule_id & (MCX_MAX_COUNTED - 1u)], 1u); }
  // outcome_products_random :2513-2521: a volume initiator counts the reaction in its counted volume, a surface
  // initiator on its wall (here: in the wall's set of counted surface regions)
  if (own_event && p.wall_cv && !(flags & DF_SURF)) agg_add(&p.rxn_count_cv[(pw.rule_id & (MCX_MAX_COUNTED - 1u)) * p.n_cv + (flags >> SF_CVI_SHIFT)], 1u);
  if (own_event && p.wall_rs && (flags & DF_SURF)) agg_add(&p.rxn_count_rs[(pw.rule_id & (MCX_MAX_COUNTED - 1u)) * p.n_rs + __ldg(p.wall_rs + p.swallA[slot])], 1u);
  bool keepA, keepB = true;
  uint32_t reuse[2]; int n_reuse = 0;
  if (kind == MCX_OUT_REACTED) {
    if (own_event) { if (bt) tally_inc(&bt->bimol); else agg_add(&c->bimol_rxns, 1u); }
    bool a_is_r0 = species == cl.r0;
    keepA = (pw.keep_mask >> (a_is_r0 ? 0 : 1)) & 1u;
    keepB = (pw.keep_mask >> (a_is_r0 ? 1 : 0)) & 1u;
  } else {
    if (own_event) { if (bt) tally_inc(&bt->unimol); else agg_add(&c->unimol_rxns, 1u); }
    keepA = pw.keep_mask & 1u;
  }
  if (!keepA) {
    atomicOr(&p.recA[slot].sf, DF_DEAD);
    if (track) { if (bt) atomicSub(&bt->species[species], 1); else agg_sub(&c->species_count[species], 1u); }
    reuse[n_reuse++] = id;
  }
  if (!keepB) {
    uint32_t old = atomicOr(&p.recA[partner_slot].sf, DF_DEAD);
    atomicOr(&p.recB[partner_slot].sf, DF_DEAD);
    if (track) { if (bt) atomicSub(&bt->species[old & SF_SPECIES_MASK], 1); else agg_sub(&c->species_count[old & SF_SPECIES_MASK], 1u); }
    uint32_t pid = p.recA[partner_slot].id;
    reuse[n_reuse++] = pid;
    if (p.trace && pid < p.n_trace) p.trace[pid].outcome = MCX_OUT_CONSUMED;
  }
  // the surface reactant of the event, if any: the partner of a volume initiator, or the initiator itself
  // (outcome_products_random :2446-2933, cases of SURVEY A.2: a surface product recycles its tile and uv; a volume
  // product is bumped 2*16*EPS off the wall to the side its orientation names and remembers where it was created)
  uint32_t surf_slot = MCX_NONE, surf_sf = 0;
  if (p.has_surf) {
    if (kind == MCX_OUT_REACTED) { surf_sf = p.recA[partner_slot].sf; if (surf_sf & DF_SURF) surf_slot = partner_slot; }
    else if (flags & DF_SURF) { surf_slot = slot; surf_sf = flags; }
  }
  const uint32_t n_new = own_event ? pw.n_products : 0u;
  const uint32_t first_slot = n_new ? c->n_slots + agg_reserve(&c->n_prod, n_new) : 0u;
  const D3 event_pos = pos;
  // products beyond the reactant ids this event frees take fresh ids, assigned after the conflict rounds (k_assign_ids)
  if (n_new > (uint32_t)n_reuse && first_slot + n_new <= p.capacity) {
    const uint32_t e = agg_reserve(&c->n_fresh_events, 1u);
    if (e >= p.fresh_cap) raise_error(p, MCX_ERR_CAPACITY, id);
    else {
      const uint32_t g = group_of(p, event_pos);
      const uint32_t nf = n_new - (uint32_t)n_reuse;
      FreshEvent ev; ev.first_slot = first_slot + (uint32_t)n_reuse; ev.n = nf; ev.init_id = id; ev.group = g;
      ev.next = atomicExch(&p.fresh_head[g], e);
      p.fresh_list[e] = ev;
      atomicAdd(&p.fresh_pref[g], nf);
      atomicAdd(&c->n_fresh_ids, nf);
    }
  }
  uint32_t n_placed_g = 0;
  for (uint32_t k = 0; k < n_new; k++) {
    uint32_t ns = first_slot + k;
    if (ns >= p.capacity) { raise_error(p, MCX_ERR_CAPACITY, id); return; }
    uint32_t nid = (int)k < n_reuse ? reuse[k] : MCX_NONE;  // fresh id: filled in by k_assign_ids
    uint32_t psp = pw.products[k];
    uint32_t pflags = DF_SCHED_UNIMOL | DF_PARTIAL | (flags & SF_CVI_MASK);  // products inherit the counted volume
    pos = event_pos;
    if (surf_slot != MCX_NONE) {
      int o = pw.prod_orient[k];
      if (o == 0) o = ((orient_bits >> k) & 1u) ? 1 : -1;
      else if (cl.kind == MCX_RXN_BIMOL_VOLSURF && cl.geom1 != 0 && ((surf_sf & DF_ORIENT_UP) ? 1 : -1) != cl.geom1) o = -o;
      const uint32_t wi = p.swallA[surf_slot];
      p.swallB[ns] = wi; p.stileB[ns] = p.stileA[surf_slot];
      if (!(p.species[psp].flags & MCX_SP_VOL)) {
        pflags = (pflags & ~SF_CVI_MASK) | DF_SURF | (o > 0 ? DF_ORIENT_UP : 0u);
        if (pw.general && p.prop_pmask) {  // where place_general put the created surface product
          const uint2 wt = p.prop_ptile[slot * MCX_MAX_PRODUCTS + n_placed_g];
          const double2 tuv = p.prop_puv[slot * MCX_MAX_PRODUCTS + n_placed_g];
          n_placed_g++;
          p.swallB[ns] = wt.x; p.stileB[ns] = wt.y; p.suvB[ns] = tuv;
          const DevWall& f = p.walls[wt.x];
          pos = D3{tuv.x * f.ux + tuv.y * f.vx + f.v0x, tuv.x * f.uy + tuv.y * f.vy + f.v0y, tuv.x * f.uz + tuv.y * f.vz + f.v0z};
        } else {
        p.suvB[ns] = p.suvA[surf_slot];
        const MolRec sr = load_rec_volatile(p.recA, surf_slot);
        pos = D3{sr.x, sr.y, sr.z};
        }
      } else {
        const DevWall& f = p.walls[wi];
        const double bump = (o > 0) ? 16 * MCX_EPS : -16 * MCX_EPS;
        pos = D3{event_pos.x + (2 * bump) * f.nx, event_pos.y + (2 * bump) * f.ny, event_pos.z + (2 * bump) * f.nz};
        pflags |= DF_CREATED_ON_SURF;
        if (p.wall_cv) {  // released to the front (up) or to the back side of the wall
          const uint32_t cv = __ldg(p.wall_cv + wi);
          uint32_t pc = o > 0 ? (cv & 0xFFu) : (cv >> 8);
          if (cv_uses_xor(p, wi)) {
            if (kind == MCX_OUT_REACTED) {  // the volume initiator's set, its object toggled for a product behind the wall
              const bool front = (orient_bits & ORIENT_BIT_FRONT) != 0;
              pc = (o > 0) == front ? (flags >> SF_CVI_SHIFT) : cv_cross(p, flags >> SF_CVI_SHIFT, wi, front);
              if (pc == MCX_NONE) { raise_error(p, MCX_ERR_STATE, id); pc = flags >> SF_CVI_SHIFT; }
            } else pflags |= DF_CVI_PENDING;  // no volume reactant to go by: a ray cast at its first evaluation
          }
          pflags = (pflags & ~SF_CVI_MASK) | (pc << SF_CVI_SHIFT);
        } else pflags &= ~SF_CVI_MASK;
      }
    }
    p.tschedB[ns] = t_event;
    store_rec(p.recB, ns, pos, nid, psp | pflags);
    uint32_t cell = cell_of(p, pos.x, pos.y, pos.z);
    p.rank[ns] = atomicAdd(&p.cs_next[cell], 1u);
    if (track) { if (bt) atomicAdd(&bt->species[psp], 1); else agg_add(&c->species_count[psp], 1u); }
    if (bt) tally_inc(&bt->products); else agg_add(&c->products, 1u);
  }
  if (keepA) {
    // kept initiator stops at the event and takes the rest of its step lazily next iteration
    uint32_t f = flags | DF_PARTIAL;
    double ut = unimol_time;
    D3 kept_pos = event_pos;
    if (cl.kind == MCX_RXN_UNIMOL) {
      f |= DF_SCHED_UNIMOL; ut = MCX_TIME_INVALID;
      if (surf_slot != MCX_NONE) {  // a kept surface molecule takes its product-side orientation (:2706-2709)
        const int ko = kept_orientation(cl, pw, 0, orient_bits, (flags & DF_ORIENT_UP) ? 1 : -1);
        if (ko != 0) f = (f & ~DF_ORIENT_UP) | (ko > 0 ? DF_ORIENT_UP : 0u);
      }
    } else if (cl.kind == MCX_RXN_BIMOL_VOLSURF && surf_slot != MCX_NONE) {
      // kept volume initiator of a surface reaction (:945-975, :2694-2716): it stays on the side it came from, or passes
      // through the wall when its product-side orientation differs from the reactant-side one (RX_FLIP).  Like a
      // volume product of the reaction it waits 2*16*EPS off the wall on that side, guarded against rebinding on
      // the same tile once (the rest of its step is taken next iteration)
      const int ko = kept_orientation(cl, pw, 0, orient_bits, (surf_sf & DF_ORIENT_UP) ? 1 : -1);
      const bool flip = ko != 0 && cl.geom0 != ko;
      const int coll_side = (orient_bits & ORIENT_BIT_FRONT) ? 1 : -1;
      const int side = flip ? -coll_side : coll_side;
      const uint32_t wi = p.swallA[surf_slot];
      const DevWall& fw = p.walls[wi];
      if (flip && p.wall_cv) {  // update_counted_volume_id_when_crossing_wall: a FRONT hit goes to the back side
        uint32_t nc = cv_cross(p, flags >> SF_CVI_SHIFT, wi, coll_side > 0);
        if (nc == MCX_NONE) { raise_error(p, MCX_ERR_STATE, id); nc = flags >> SF_CVI_SHIFT; }
        f = (f & ~SF_CVI_MASK) | (nc << SF_CVI_SHIFT);
      }
      const double bump = (side > 0) ? 16 * MCX_EPS : -16 * MCX_EPS;
      kept_pos = D3{event_pos.x + (2 * bump) * fw.nx, event_pos.y + (2 * bump) * fw.ny, event_pos.z + (2 * bump) * fw.nz};
      f |= DF_CREATED_ON_SURF;
      p.swallB[slot] = wi; p.stileB[slot] = MCX_KEPT_AT_WALL;
    }
    finalize_alive(p, slot, kept_pos, id, species, f, t_event, ut,
                   (cl.kind == MCX_RXN_UNIMOL && (flags & DF_SURF)) ? SURF_FIELDS_IN_B : nullptr);
  }
}

// n_list: counter of the proposal list (pend[0]) of the round the proposal joins; nullptr: the caller appends the slot
// itself (WarpList)
__device__ __forceinline__ void write_proposal(const DevParams& p, uint32_t slot, const Outcome& o, uint32_t id,
                                               uint32_t species, unsigned int epoch, unsigned int* n_list) {
  store_rec(p.recB, slot, o.pos, id, species | (o.flags & ~DF_DEAD));
  p.tschedB[slot] = o.t_now;
  p.tuniB[slot] = o.unimol_time;
  p.prop_partner[slot] = o.partner_slot;
  p.prop_info[slot] = (uint32_t)o.kind | (((uint32_t)o.pathway & 0xFFu) << 4) | ((o.orient_bits & ORIENT_BITS_MASK) << 12) |
                      ((uint32_t)o.rxn_class << 19);
  p.prop_t[slot] = o.t_event;
  p.rank[slot] = MCX_NONE;
  unsigned long long key = claim_key(epoch, id);
  if (o.kind == MCX_OUT_SURFMOVE) key = weak_key(key);
  atomicMax(&p.claim[slot], key);
  if (partner_is_consumed(p, o.kind, o.rxn_class, o.pathway, species)) atomicMax(&p.claim[o.partner_slot], key);
  if (o.kind == MCX_OUT_SURFMOVE) {
    p.swallB[slot] = o.s_wall; p.stileB[slot] = o.s_tile; p.suvB[slot] = make_double2(o.s_u, o.s_v);
    atomicMax(&p.tile_claim[p.grids[o.s_wall].tile_start + o.s_tile], key);
  } else if (o.kind == MCX_OUT_UNIMOL && (o.flags & DF_SURF)) {
    // a surface molecule may have moved inside its tile earlier in the iteration (a sub-step that ended at its lifetime):
    // a kept reactant stays where it is now, not where the snapshot has it
    p.swallB[slot] = o.s_wall; p.stileB[slot] = o.s_tile; p.suvB[slot] = make_double2(o.s_u, o.s_v);
  } else if (o.kind == MCX_OUT_REACTED && p.classes[o.rxn_class].kind == MCX_RXN_BIMOL_SURFSURF) {
    // surface-surface reaction: where the initiator is after its move; the tile is claimed when it is a new one
    p.swallB[slot] = o.s_wall; p.stileB[slot] = o.s_tile; p.suvB[slot] = make_double2(o.s_u, o.s_v);
    if (o.s_wall != p.swallA[slot] || o.s_tile != p.stileA[slot]) atomicMax(&p.tile_claim[p.grids[o.s_wall].tile_start + o.s_tile], key);
  }
  if (event_is_general(p, o.kind, o.rxn_class, o.pathway)) {  // products on vacant neighbour tiles claim them (place_general wrote them)
    const uint32_t pm = p.prop_pmask[slot];
    for (uint32_t c = 0; c < (pm & 15u); c++)
      if ((pm >> (4 + c)) & 1u) {
        const uint2 wt = p.prop_ptile[slot * MCX_MAX_PRODUCTS + c];
        atomicMax(&p.tile_claim[p.grids[wt.x].tile_start + wt.y], key);
      }
  }
  if (n_list) {
    uint32_t k = agg_reserve(n_list, 1u);
    p.pend[0][k] = slot;
  }
}

__device__ __forceinline__ void trace_begin(const DevParams& p, Tracer& tc, uint32_t id) {
  tc.h = 0xcbf29ce484222325ULL; tc.tr = nullptr;
  if (p.trace && id < p.n_trace) {
    mcx_trace_rec* t = p.trace + id;
    uint32_t rounds = t->rounds;
    mcx_trace_rec z;
    memset(&z, 0, sizeof(z));
    z.id = id; z.rxn_class = z.rxn_pathway = z.rxn_partner = MCX_NONE; z.rounds = rounds + 1;
    *t = z;
    tc.tr = t;
  }
}
__device__ __forceinline__ void trace_end(Tracer& tc, const Outcome& o, const Stream& rs) {
  if (!tc.tr) return;
  tc.tr->outcome = o.kind; tc.tr->n_words = rs.used; tc.tr->event_hash = tc.h;
  tc.tr->pos[0] = o.pos.x; tc.tr->pos[1] = o.pos.y; tc.tr->pos[2] = o.pos.z;
}

// ---------------------------------------------------------------------------------------------------
// k_diffuse_fast: every live molecule.  The common case — a whole time step, no pending unimolecular event,
// no wall in reach, no partner in reach — is finished here with converged control flow: 3 Gaussians, one
// subpartition flag lookup, one flattened candidate walk, store + bin.  Everything else (wall hits, collisions,
// split steps, newborn molecules, subpartition-set filtering) is appended to slow_list and evaluated by
// k_diffuse_slow with the generic code, restarting the molecule's random stream from its first word, so both
// kernels together compute exactly evaluate_iteration() for every molecule.
//
// Why the fast test is exact (reference semantics):
//  * walls are only tested in subpartitions the segment crosses (ray_trace_vol :698); those lie in the index
//    block spanned by the start and end subpartitions, so "no wall in the start subpartition" (segment stays
//    inside) or "no wall in its 3x3x3 neighbourhood" (|index step| <= 1) means no wall collision;
//  * partners are the molecules of collected subpartitions that pass collide_mol; probing ALL molecules that
//    pass collide_mol (no subpartition filter) is a superset, so an empty probe means no collision.
//
// The body is written flat: every lane of the warp executes the same instruction stream with predicates
// (`simple` narrows as tests fail; lanes that dropped out run empty loops), because nested `if (simple) {...}`
// blocks made the compiler split the warp into independently scheduled groups that each ran the whole
// candidate walk (profiles/r01_b: 7 of 32 lanes active).
#ifndef MCX_FAST_MINBLOCKS
#define MCX_FAST_MINBLOCKS 4
#endif
// PASS 0 walks every slot of the snapshot and finishes the whole-step molecules.  Molecules that start the iteration
// at a fractional diffusion_time (newborn products, kept initiators: DF_PARTIAL / DF_SCHED_UNIMOL) take two
// sub-steps — one whole step, then the rest of the iteration (diffuse_single_molecule's reschedule loop, :283-336);
// they were 47 % of what the generic pass had to evaluate, at two partner scans each.  PASS 0 appends them to
// second_list and PASS 1 runs the same flat body over that list, twice per molecule, with full warps of them.
// Whatever turns out not to be simple in either pass goes to slow_list and is restarted from the snapshot there.
// The per-molecule body of the fast passes (see k_diffuse_fast below).  Warp-collective: all 32 lanes call it
// together (lanes without a molecule pass in_range == false and a valid slot).  Probe is the partner probe of the
// launch: the warp-flattened gather walk (FlatProbe) or the shared-memory tile (TileProbe, mcx_tile.cuh).
struct FastCtx {
  const ZigShared* zig;
  WarpList slow_wl, prop_wl, second_wl;
  unsigned int* n_slow_ctr;
  uint32_t* slow_out;
  unsigned int* reason_row;   // this warp's deferral-reason counters (shared memory)
  unsigned int msteps, n_tests, n_coll;
};
struct FlatProbe {
  static constexpr int HIT_STRIDE = 1;
  WarpProbe* sm;
  __device__ __forceinline__ int run(const DevParams& p, bool probing, D3 pos, D3 disp, uint32_t id, uint32_t species, uint32_t,
                                     bool& overflow, bool& outside) {
    outside = false;
    return probe_partners_flat(p, probing, pos, disp, id, species, overflow, sm);
  }
  __device__ __forceinline__ const uint32_t* hit_slots() const { return &sm->hit_slot[threadIdx.x & 31][0]; }
};

template <int PASS, class Probe>
__device__ __forceinline__ void fast_molecule(const DevParams& p, FastCtx& cx, Probe& probe, const bool in_range, const unsigned int i) {
  const double it = (double)p.iteration, t_end = it + 1.0;
    const MolRec m = load_rec(p.recA, i);
    const bool live = in_range && !(m.sf & DF_DEAD);
    const uint32_t species = m.sf & SF_SPECIES_MASK;
    uint32_t flags = m.sf & ~SF_SPECIES_MASK;
    const DevSpecies sp = p.species[species];
    const bool vol_diffuser = live && !(m.sf & (DF_SURF | DF_CREATED_ON_SURF | DF_CVI_PENDING)) && (sp.flags & MCX_SP_CAN_DIFFUSE) && sp.time_step == 1.0;
    const bool fractional = (m.sf & (DF_PARTIAL | DF_SCHED_UNIMOL)) != 0;
    bool to_second = PASS == 0 && vol_diffuser && fractional;
    bool simple = vol_diffuser && (PASS == 1 || !fractional);
    // cold fields: a predicated index keeps the loads unconditional (no warp split) without touching the arrays
    // for molecules that have nothing there
    const bool has_uni = (m.sf & DF_HAS_UNIMOL) != 0;
    const bool idle_candidate = PASS == 0 && live && !(sp.flags & MCX_SP_CAN_DIFFUSE) && !fractional && !(m.sf & DF_CVI_PENDING) &&
                                !sp.can_surf_surf;  // surface molecules with surface-surface classes test their neighbours every step
    const double t_uni_raw = __ldg(p.tuniA + (((simple || idle_candidate) && has_uni) ? i : 0u));
    double t_uni = has_uni ? t_uni_raw : MCX_TIME_INVALID;
    double t_now = it;
    if (PASS == 1) {
      const double t_sched_raw = __ldg(p.tschedA + ((m.sf & DF_PARTIAL) ? i : 0u));
      if (m.sf & DF_PARTIAL) t_now = t_sched_raw;
    }
    // a non-diffusing molecule with nothing scheduled inside this iteration (receptors, pumps): the generic
    // evaluation would return MCX_OUT_STATIC without drawing a number (diffuse_react_event.cpp:318-335)
    const bool idle = idle_candidate && !(has_uni && t_uni < t_end);
    Stream rs; rs.init(p, m.id, cx.zig);
    Tracer tc; tc.h = 0xcbf29ce484222325ULL; tc.tr = nullptr;
    if (PASS == 1) {
      // newbie lifetime (:232-236 -> pick_unimol_rxn_class_and_set_rxn_time :1731-1758): one draw; a unimolecular
      // reaction that is due comes first (:215-223, the firing below)
      if (simple && (flags & DF_SCHED_UNIMOL) && !(t_uni != MCX_TIME_INVALID && t_uni <= t_now)) {
        flags &= ~DF_SCHED_UNIMOL;
        const int rc = p.unimol[species];
        if (rc < 0) t_uni = MCX_TIME_INVALID;
        else {
          const double k_tot = p.classes[rc].max_fixed_p;
          const double pr = rs.dbl();
          const double from_now = (k_tot <= 0 || !distinguishable_d(pr, 0, MCX_EPS)) ? MCX_TIME_FOREVER : -mcx_log(pr) / k_tot;
          t_uni = t_now + from_now;
        }
      }
      if (simple) trace_begin(p, tc, m.id);
    } else {
      // a lifetime ending inside this iteration splits the step and fires (get_max_time :164-198): PASS 1
      const bool timed = simple && t_uni != MCX_TIME_INVALID && t_uni < t_end;
      if (timed) { to_second = true; simple = false; }
    }
    int reason = simple ? -1 : MCX_DEFER_TIMING;
    const bool own_start = owned_z(p, m.z);
    D3 pos = {m.x, m.y, m.z};
    unsigned int my_tests = 0, my_coll = 0;
    bool running = simple, proposed = false;  // running: this lane's evaluation is simple so far and not finished
    bool slow = live && !idle && !to_second && !simple;

#pragma unroll 1
    for (int sub = 0; sub < (PASS == 0 ? 1 : 2); sub++) {
      if (PASS == 1) {
        // unimolecular firing (diffuse_react_event.cpp:215-223, :1764-1826) ends the evaluation with a claiming event
        const bool fire = running && t_uni != MCX_TIME_INVALID && t_uni <= t_now;
        if (fire) {
          const int rc = p.unimol[species];
          const DevClass& cl = p.classes[rc];
          int pathway = 0;
          if (cl.n_pathways > 1) {  // which_unimolecular, rxn_utils.inl:774-783
            const double match = rs.dbl() * cl.max_fixed_p;
            int min_idx = 0, max_idx = (int)cl.n_pathways - 1;
            const DevPathway* A = p.pathways + cl.first_pathway;
            while (max_idx - min_idx > 1) {
              const int mid = (max_idx + min_idx) / 2;
              if (match > A[mid].cum_prob) min_idx = mid; else max_idx = mid;
            }
            pathway = match > A[min_idx].cum_prob ? max_idx : min_idx;
          }
          tc.ev(EV_UNIMOL | (uint32_t)pathway, (uint32_t)rc);
          if (tc.tr) { tc.tr->rxn_class = rc; tc.tr->rxn_pathway = pathway; tc.tr->t_event = t_uni; }
          Outcome o; o.kind = MCX_OUT_UNIMOL; o.pos = pos; o.rxn_class = rc; o.pathway = pathway; o.t_event = t_uni;
          o.t_now = t_now; o.flags = flags; o.unimol_time = t_uni; o.partner_slot = MCX_NONE; o.partner_id = MCX_NONE; o.orient_bits = 0;
          trace_end(tc, o, rs);
          write_proposal(p, i, o, m.id, species, round_epoch(p, 0), nullptr);
          proposed = true;
          if (own_start) { cx.msteps++; cx.n_tests += my_tests; cx.n_coll += my_coll; }
          running = false;
        }
      }
      simple = running;
      // compute_vol_displacement (diffusion_utils.inl:366-432) with time_step == 1; drawn by every lane
      double t_steps = 1.0, scale = sp.space_step, r_rate_factor = 1.0, t_new = t_end;
      bool again = false;
      if (PASS == 1) {
        // get_max_time (:164-198): up to the end of the iteration or to the scheduled unimolecular reaction
        double max_time = t_end - t_now;
        if (t_uni != MCX_TIME_INVALID && t_uni < t_now + max_time) max_time = t_uni - t_now;
        double steps = 1.0;
        if (t_steps > max_time) { t_steps = max_time; steps = max_time / sp.time_step; }
        simple = simple && !(steps < MCX_EPS);
        if (steps != 1.0) { const double rate_factor = sqrt(steps); r_rate_factor = 1.0 / rate_factor; scale = rate_factor * sp.space_step; }
        // reschedule (:283-336): does the molecule need another sub-step (or fire) after this one?
        t_new = t_now + t_steps;
        again = (t_uni != MCX_TIME_INVALID && t_uni < t_end) || (t_new < t_end && !cmp_eq_d(t_new, t_end, MCX_EPS));
        simple = simple && !(again && sub == 1);
        if (!simple && reason < 0) reason = MCX_DEFER_TIMING;
      }
      // the three draws share ONE copy of the Ziggurat code (a rolled loop with predicated moves): the inlined copies
      // were a quarter of the kernel's instructions and 18 % of its stall samples waited for instruction fetch —
      // same-box A/B at 1e8 molecules: 17.0 -> 16.15 ms (profiles/r01_w_*); rolling the two wall loops the same way
      // cost 0.9 ms and was not kept
      D3 disp = {0.0, 0.0, 0.0};
#pragma unroll 1
      for (int axis = 0; axis < 3; axis++) {
        const double g = scale * rs.gauss() * 0.70710678118654752440;
        if (axis == 0) disp.x = g; else if (axis == 1) disp.y = g; else disp.z = g;
      }
      const D3 dest = pos + disp;
      simple = simple && in_partition(p, dest);
      if (!simple && reason < 0) reason = MCX_DEFER_GEOMETRY;
      int s0[3], s1[3];
      subpart_3d(p, pos, s0);
      subpart_3d(p, dest, s1);
      const int dx = s1[0] - s0[0], dy = s1[1] - s0[1], dz = s1[2] - s0[2];
      const uint32_t own = simple ? subpart_from_3d(p, s0[0], s0[1], s0[2]) : 0u;
      const uint8_t f = __ldg(p.sp_flags + own);
      const bool same = (dx | dy | dz) == 0;
      const bool near = dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1 && dz >= -1 && dz <= 1;
      // one subpartition face crossed: the DDA visits exactly {start, end} (collision_utils_subparts.inl:127-300)
      const bool single = near && ((dx != 0) + (dy != 0) + (dz != 0)) == 1;
      const uint32_t dest_sp = (simple && single) ? subpart_from_3d(p, s1[0], s1[1], s1[2]) : 0u;
      const uint8_t fd = __ldg(p.sp_flags + dest_sp);
      unsigned int n_wall_tests = 0, n_wall_tests_dest = 0;
      const bool walls_own = simple && (same || single) && (f & 1);
      const bool walls_dest = simple && single && (fd & 1);
      double wall_dist = 1e300;
      const bool rejected_own = all_walls_plane_rejected(p, walls_own, own, pos, disp, n_wall_tests, wall_dist);
      const bool rejected_dest = all_walls_plane_rejected(p, walls_dest, dest_sp, pos, disp, n_wall_tests_dest, wall_dist);
      n_wall_tests += n_wall_tests_dest;
      simple = simple && ((same || single) ? ((!walls_own || rejected_own) && (!walls_dest || rejected_dest)) : (near && !(f & 2)));
      if (!simple && reason < 0) reason = (same || single) ? MCX_DEFER_WALL : MCX_DEFER_GEOMETRY;

      // partner probe: none -> move; one hit in the molecule's own subpartition (always a collected one) -> evaluated
      // below; 2..MCX_FAST_MAX_HITS hits -> PASS 1, which evaluates them in collision order; anything else -> generic
#ifdef MCX_EXPERIMENT_NO_PROBE   // timing experiment only (profiles/): what the fast pass costs without its partner probe
      const bool probing = false;
#else
      const bool probing = simple && sp.can_vol_react;
#endif
      // outside (tile probe only): the swept box leaves the staged tile — PASS 1 probes it with the gather walk
      bool overflow, outside;
#ifdef MCX_EXPERIMENT_IGNORE_HITS   // timing experiment only: the probe runs, what it finds is dropped (cost of the hit handling)
      const int n_hits = probe.run(p, probing, pos, disp, m.id, species, i, overflow, outside) & 0;
#else
      const int n_hits = probe.run(p, probing, pos, disp, m.id, species, i, overflow, outside);
#endif
      // a collision next to walls needs exact_disk (walls of the collision subpartition): generic path
      // (every wall plane of the crossed subpartitions farther than R from the whole segment: the factor is 1)
      const bool disk_walls = n_hits > 0 && wall_dist < p.R;
      const bool several = PASS == 0 && simple && !overflow && !disk_walls && n_hits > 1 && n_hits <= MCX_FAST_MAX_HITS;
      simple = simple && !overflow && n_hits <= (PASS == 0 ? 1 : MCX_FAST_MAX_HITS) && !disk_walls;
      if (!simple && reason < 0)
        reason = overflow ? MCX_DEFER_PROBE_SHAPE : (disk_walls ? MCX_DEFER_DISK : MCX_DEFER_MULTI_HIT);
      // the first collision; a hit in a foreign subpartition counts only if the reference collects that
      // subpartition for this move: PASS 1 works that out for moves that stay inside their subpartition
      PartnerHit ph;
      bool decided = true;
      bool have = simple && n_hits > 0 &&
                  next_probe_hit<PASS == 1>(p, probe.hit_slots(), Probe::HIT_STRIDE, n_hits, pos, disp, same, single, s1, species, -1.0, 0u, ph, decided);
      const bool foreign = PASS == 0 && simple && !decided && same;
      simple = simple && decided;
      if (!simple && reason < 0) reason = MCX_DEFER_FOREIGN_HIT;
      const bool elsewhere = PASS == 0 && simple && outside;
      simple = simple && !outside;
      if (several || foreign || elsewhere) { to_second = true; slow = false; running = false; }

      if (simple) {
        if (PASS == 0) trace_begin(p, tc, m.id);
        Outcome o; o.kind = MCX_OUT_MOVED; o.pos = dest;
        unsigned int colls = 0;
        // sort_collisions_by_time realised as repeated selection, like the generic pass (evaluate_iteration);
        // PASS 0 has at most one
        while (have) {
          colls++;
          if (!(ph.t < MCX_EPS)) {  // is_immediate_collision (collision_utils.inl:814-816)
            // collide_and_react_with_vol_mol (:786-829) with scaling = factor(1) * r_rate_factor
            tc.ev(EV_COLL, ph.id);
            if (tc.tr) { if (tc.tr->n_collisions < MCX_TRACE_K) tc.tr->partner[tc.tr->n_collisions] = ph.id; tc.tr->n_collisions++; }
            const int pathway = test_bimolecular(p, p.classes[ph.rxn_class], r_rate_factor, rs);
            if (pathway >= 0) {
              const double abs_t = t_now + t_steps * ph.t;
              tc.ev(EV_RXN | (uint32_t)pathway, (uint32_t)ph.rxn_class);
              if (tc.tr) { tc.tr->rxn_class = ph.rxn_class; tc.tr->rxn_pathway = pathway; tc.tr->rxn_partner = ph.id; tc.tr->t_event = abs_t; }
              o.kind = MCX_OUT_REACTED; o.pos = pos + disp * ph.t;
              o.rxn_class = ph.rxn_class; o.pathway = pathway; o.partner_slot = ph.slot; o.partner_id = ph.id;
              o.t_event = abs_t; o.t_now = t_now; o.flags = flags; o.unimol_time = t_uni; o.orient_bits = 0;
              break;
            }
          }
          if (PASS == 0 || (int)colls >= n_hits) break;
          const double t_last = ph.t; const uint32_t id_last = ph.id;
          have = next_probe_hit<PASS == 1>(p, probe.hit_slots(), Probe::HIT_STRIDE, n_hits, pos, disp, same, single, s1, species, t_last, id_last, ph, decided);
        }
        my_tests += n_wall_tests; my_coll += colls;
        if (PASS == 1 && o.kind == MCX_OUT_MOVED && again) {  // first sub-step done
          pos = dest; t_now = t_new;
        } else {
          trace_end(tc, o, rs);
          if (o.kind == MCX_OUT_MOVED) {
            if (PASS == 1) { const double r = round(t_new); if (cmp_eq_d(t_new, r, MCX_SQRT_EPS)) t_new = r; }
            finalize_alive(p, i, dest, m.id, species, flags & ~DF_PARTIAL, t_new, t_uni);
          } else { write_proposal(p, i, o, m.id, species, round_epoch(p, 0), nullptr); proposed = true; }
          if (own_start) { cx.msteps++; cx.n_tests += my_tests; cx.n_coll += my_coll; }
          running = false;
        }
      } else if (running) {
        running = false;
        slow = true;
      }
    }
    if (PASS == 0 && in_range && !live) p.rank[i] = MCX_NONE;
    if (idle) {
      if (p.trace && m.id < p.n_trace) {
        trace_begin(p, tc, m.id);
        Outcome o; o.kind = MCX_OUT_STATIC; o.pos = pos;
        trace_end(tc, o, rs);
        tc.tr->n_words = 0;
      }
      finalize_alive(p, i, pos, m.id, species, flags, 0.0, t_uni);
    }
    if (PASS == 1 && slow && tc.tr) tc.tr->rounds--;  // the generic pass starts it over (and counts the evaluation)
    // staged appends (loop bounds are warp-uniform: all 32 lanes arrive here)
    if (slow) atomicAdd(&cx.reason_row[reason & 7], 1u);
    cx.slow_wl.push(slow, i, cx.n_slow_ctr, cx.slow_out);
    cx.prop_wl.push(proposed, i, &p.ctr->n_prop[0], p.pend[0]);
    if (PASS == 0) cx.second_wl.push(to_second, i, &p.ctr->n_second, p.second_list);
}

// end of a fast-pass kernel: flush the staged list appends and the warp's statistics
template <int PASS>
__device__ __forceinline__ void fast_finish(const DevParams& p, FastCtx& cx, int lane) {
  cx.slow_wl.flush(cx.n_slow_ctr, cx.slow_out);
  cx.prop_wl.flush(&p.ctr->n_prop[0], p.pend[0]);
  if (PASS == 0) cx.second_wl.flush(&p.ctr->n_second, p.second_list);
  __syncwarp();
  if (lane == 0 && cx.slow_wl.total) atomicAdd(&p.ctr->deferred, (unsigned long long)cx.slow_wl.total);
  if (lane < 8 && cx.reason_row[lane]) atomicAdd(&p.ctr->defer_reason[lane], (unsigned long long)cx.reason_row[lane]);
  LocalStats ls = {cx.n_tests, 0, 0, 0, cx.n_coll, 0};
  flush_stats(p, ls, cx.msteps);
}

template <int PASS>
__global__ void __launch_bounds__(TPB, MCX_FAST_MINBLOCKS) k_diffuse_fast(const __grid_constant__ DevParams p) {
  __shared__ ZigShared zig;
  __shared__ uint32_t s_slow[TPB / 32][WL_CAP], s_prop[TPB / 32][WL_CAP], s_second[PASS == 0 ? TPB / 32 : 1][WL_CAP];
  __shared__ unsigned int s_reason[TPB / 32][8];
  __shared__ WarpProbe s_probe[TPB / 32];
  // an empty launch costs its blocks nothing but this test (small models are launch bound: profiles/r02_i_configs)
  if ((PASS == 0 ? p.ctr->n_slots : p.ctr->n_second) <= blockIdx.x * blockDim.x) return;
  zig_load(&zig);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < 8) s_reason[warp][lane] = 0;
  __syncthreads();
  FastCtx cx;
  cx.zig = &zig;
  cx.slow_wl.init(s_slow[warp]);
  cx.prop_wl.init(s_prop[warp]);
  cx.second_wl.init(s_second[PASS == 0 ? warp : 0]);
  const unsigned int n = PASS == 0 ? p.ctr->n_slots : p.ctr->n_second;
  // PASS 1's own deferrals go to a list of their own (a small second launch of the generic pass takes them)
  cx.n_slow_ctr = PASS == 0 ? &p.ctr->n_slow : &p.ctr->n_slow2;
  cx.slow_out = PASS == 0 ? p.slow_list : p.slow2_list;
  cx.reason_row = s_reason[warp];
  cx.msteps = 0; cx.n_tests = 0; cx.n_coll = 0;
  FlatProbe probe{&s_probe[warp]};
  for (unsigned int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const unsigned int k = base + threadIdx.x;
    const bool in_range = k < n;
    const unsigned int kk = in_range ? k : base;  // a valid entry for every lane
    const unsigned int i = PASS == 0 ? kk : __ldg(p.second_list + kk);
    fast_molecule<PASS>(p, cx, probe, in_range, i);
  }
  fast_finish<PASS>(p, cx, lane);
}

// dynamic shared memory a tile block may use: what the multiprocessor has (227 KB, 1 KB reserved per block) shared by
// the resident blocks, minus the static arrays of the kernel
#define TILE_SMEM_DYN_MAX ((TILE_TPB == 1024 ? 232448 : 115712) - 28 * TILE_TPB - 2048)
// PASS 0 over shared-memory tiles (mcx_tile.cuh): one persistent block per multiprocessor walks the tiles of the cell
// grid; the same per-molecule body as k_diffuse_fast<0>, with the partner probe served from the staged tile.
__global__ void __launch_bounds__(TILE_TPB, TILE_BLOCKS_PER_SM) k_diffuse_tile(const __grid_constant__ DevParams p) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ ZigShared zig;
  __shared__ uint32_t s_slow[TILE_TPB / 32][WL_CAP], s_prop[TILE_TPB / 32][WL_CAP], s_second[TILE_TPB / 32][WL_CAP];
  __shared__ unsigned int s_reason[TILE_TPB / 32][8];
  TileSmem ts;
  tile_smem_layout(p.tile, tile_smem, ts);
  zig_load(&zig);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < 8) s_reason[warp][lane] = 0;
  if (threadIdx.x < 32) {
    uint32_t mask = 0;
    if ((int)threadIdx.x < p.n_species && p.n_species <= 32)
      for (int sp = 0; sp < p.n_species; sp++) if (p.bimol[threadIdx.x * p.n_species + sp] >= 0) mask |= 1u << sp;
    ts.rmask[threadIdx.x] = mask;
  }
  if (threadIdx.x == 0) { mbar_init(ts.bar, 1); fence_proxy_async(); }
  __syncthreads();
  FastCtx cx;
  cx.zig = &zig;
  cx.slow_wl.init(s_slow[warp]);
  cx.prop_wl.init(s_prop[warp]);
  cx.second_wl.init(s_second[warp]);
  cx.n_slow_ctr = &p.ctr->n_slow;
  cx.slow_out = p.slow_list;
  cx.reason_row = s_reason[warp];
  cx.msteps = 0; cx.n_tests = 0; cx.n_coll = 0;
  const unsigned int n_tiles = p.tile.n_tiles;
  const int n_rows = (p.tile.TY + 2) * (p.tile.TZ + 2);
  unsigned int parity = 0;
  int cur = 0;
  if (warp == 0 && blockIdx.x < n_tiles) tile_issue(p, ts, blockIdx.x, &ts.tab[0]);
  for (unsigned int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, cur ^= 1) {
    const TileTab* tab = &ts.tab[cur];
    mbar_wait(ts.bar, parity);  // the tile's records are in shared memory, its tables are visible (release / acquire)
    parity ^= 1u;
    if (!tab->overflow) tile_bin(p, ts, tab); else __syncthreads();
    // the staging buffer is free again: the next tile's bulk copies run under this tile's evaluation
    const unsigned int next = tile + gridDim.x;
    if (warp == 0 && next < n_tiles) { fence_proxy_async(); tile_issue(p, ts, next, &ts.tab[cur ^ 1]); }
    const unsigned int n_owned = tab->n_owned;
    for (unsigned int kb = warp * 32; kb < n_owned; kb += TILE_TPB) {  // warp-uniform bounds
      const unsigned int k = kb + lane;
      const bool in_range = k < n_owned;
      const unsigned int kk = in_range ? k : kb;
      int lo = 0, hi = n_rows;  // row of owned molecule kk: last row whose owned prefix is <= kk
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tab->own_pref[mid] <= kk) lo = mid; else hi = mid; }
      const uint32_t own_idx = tab->own_lo[lo] + (kk - tab->own_pref[lo]);
      const uint32_t slot = tab->row_src[lo] + (own_idx - tab->row_off[lo]);
      TileProbe probe{&ts, tab, own_idx};
      fast_molecule<0>(p, cx, probe, in_range, slot);
    }
    __syncthreads();  // the fine grid and the tables of this tile are dead
  }
  fast_finish<0>(p, cx, lane);
}

// Tile geometry for a population of n_records in the local cell grid (called whenever stepping starts).  TY = TZ = 4
// owned cell rows and one halo row each side; the x extent from the staging capacity, the size of the fine-cell table
// and ~0.9 owned molecules per thread; fp32 slacks from the staged extent (see TileProbe::run).
void mcx_plan_tiles(DevParams& p, unsigned long long n_records) {
  TileGeom g{};
  p.tile = g;
  // opt-in: measured slower than the gather walk so far (profiles/r02_c..f)
  const char* on = getenv("MCX_TILE");
  if (!on || atoi(on) == 0 || n_records == 0 || p.n_cells == 0) return;
  const double ex = 1.0 / p.cell_rcp_x, ey = 1.0 / p.cell_rcp_y, ez = 1.0 / p.cell_rcp_z;
  const double lambda = (double)n_records / (double)p.n_cells;
  g.TY = p.ncy < 4 ? p.ncy : 4; g.TZ = p.ncz < 4 ? p.ncz : 4;
  g.lsy = g.lsz = 1;
  if (const char* e = getenv("MCX_TILE_LS")) g.lsy = g.lsz = atoi(e);
  g.hx = (int)ceil(fmax(ey, ez) / ex);
  if (g.hx < 1) g.hx = 1;
  if (g.hx > 16) return;
  g.nfy = (g.TY + 2) << g.lsy; g.nfz = (g.TZ + 2) << g.lsz;
  g.cap = TILE_TPB == 1024 ? 3328 : 1664;
  if (const char* e = getenv("MCX_TILE_CAP")) g.cap = (unsigned int)atoi(e);
  if (g.cap > TILE_TPB * TILE_REC_PER_THREAD) g.cap = TILE_TPB * TILE_REC_PER_THREAD;
  const int rows = (g.TY + 2) * (g.TZ + 2);
  double sx = TILE_NF_MAX / (double)(g.nfy * g.nfz);
  sx = fmin(sx, 0.82 * g.cap / (rows * lambda));
  double tx = fmin(sx - 2 * g.hx, 0.9 * TILE_TPB / (g.TY * g.TZ * lambda));
  if (const char* e = getenv("MCX_TILE_TX")) tx = atof(e);
  g.TX = (int)floor(tx);
  if (g.TX > p.ncx) g.TX = p.ncx;
  if (g.TX < 4) return;
  g.nfx = g.TX + 2 * g.hx;
  if ((long long)g.nfx * g.nfy * g.nfz > TILE_NF_MAX) return;
  g.ntx = (p.ncx + g.TX - 1) / g.TX; g.nty = (p.ncy + g.TY - 1) / g.TY; g.ntz = (p.ncz + g.TZ - 1) / g.TZ;
  const unsigned long long nt = (unsigned long long)g.ntx * g.nty * g.ntz;
  if (nt > 0xFFFFFFF0ull) return;
  g.n_tiles = (unsigned int)nt;
  // fp32 slacks: coordinates relative to the tile are below S, so each carries an absolute error below eps_c
  const double S = fmax(g.nfx * ex, fmax((g.TY + 2) * ey, (g.TZ + 2) * ez)) * 1.01, D = S * 1.7320508;
  const double eps_c = S * ldexp(1.0, -23);
  const double tol_d = 4.0 * (2.0 * eps_c * 1.7320508 * D + D * D * ldexp(1.0, -22));
  const double bound_v = 4.0 * D * 2.0 * eps_c * 1.7320508 + 8.0 * D * D * ldexp(1.0, -24);
  const double r2 = p.R * p.R, r2p = r2 * 1.0625 + 2.0 * bound_v;
  if (!(r2p < 2.0 * r2)) return;  // boxes too large for an fp32 pre-filter: the gather walk runs
  g.tol_d = (float)tol_d; g.r2p = (float)r2p;
  TileSmem ts;
  g.smem_bytes = tile_smem_layout(g, nullptr, ts);
  if (g.smem_bytes > TILE_SMEM_DYN_MAX) return;
  g.enabled = 1;
  p.tile = g;
}

// The generic evaluation is one long path per molecule.  A group of G lanes evaluates one molecule together (Group,
// mcx_device.cuh): identical control flow on all of them, the wall lists and candidate records dealt among them, lane 0
// of the group writes the result.  G is chosen per launch: as many lanes per molecule as the launch has threads for —
// big lists keep one molecule per lane (the redundant part of a group's work would cost throughput there: 1e8
// molecules, 5.7e5 deferred: 4.0 ms with G = 1, 6.4 ms with G = 8, profiles/r02_k), short lists get short chains.
#ifndef MCX_SLOW_MINBLOCKS
#define MCX_SLOW_MINBLOCKS 4
#endif
#ifndef MCX_GROUP_FILL
#define MCX_GROUP_FILL 1   // lanes per molecule are doubled while n * G <= MCX_GROUP_FILL * threads of the launch
#endif
__device__ __forceinline__ int lanes_per_molecule(unsigned int n) {
  const unsigned long long threads = (unsigned long long)gridDim.x * blockDim.x * MCX_GROUP_FILL;
  int g = 1;
  while (g < 8 && (unsigned long long)n * (g * 2) <= threads) g *= 2;
  return g;
}

template <bool WITH_DISK, bool SURF>
__global__ void __launch_bounds__(TPB, MCX_SLOW_MINBLOCKS) k_diffuse_slow(const __grid_constant__ DevParams p, int second) {
  __shared__ ZigShared zig;
  // second: the deferrals of k_diffuse_fast<1> instead of those of k_diffuse_fast<0>
  const unsigned int n = WITH_DISK ? p.ctr->n_disk : (second ? p.ctr->n_slow2 : p.ctr->n_slow);
  if (n == 0) return;
  zig_load(&zig);
  __syncthreads();
  const uint32_t* list = WITH_DISK ? p.pend[1] : (second ? p.slow2_list : p.slow_list);
  const unsigned int epoch = round_epoch(p, 0);
  LocalStats ls = {0, 0, 0, 0, 0, 0};
  unsigned int msteps = 0;
  const int G = lanes_per_molecule(n);
  const Group grp = group_of(G);
  const bool lead = grp.sub == 0;
  // all lanes of a warp iterate together so the warp-level stat flush sees full warps
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x; base < (unsigned long long)n * G;
       base += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long k = (base + threadIdx.x) / G;
    if (k >= n) continue;  // whole groups leave together
    // the list is in snapshot order: neighbours in space — the long evaluations next to a complicated piece of mesh —
    // sit in the same warp, which then runs them back to back; a stride permutation spreads them over the launch
    // (config 4: 6.40 -> 6.08 ms for the generic launches, profiles/r02_o)
    if (G == 1) { const unsigned long long P = (n % 1000003ull) ? 1000003ull : 999983ull; k = (k * P) % n; }
    const unsigned int i = list[k];
    MolRec m = load_rec(p.recA, i);
    const uint32_t species = m.sf & SF_SPECIES_MASK;
    double t_sched = (m.sf & DF_PARTIAL) ? p.tschedA[i] : 0.0;
    double t_uni = (m.sf & DF_HAS_UNIMOL) ? p.tuniA[i] : MCX_TIME_INVALID;
    Stream rs; rs.init(p, m.id, &zig);
    Tracer tc; tc.h = 0xcbf29ce484222325ULL; tc.tr = nullptr;
    if (lead) trace_begin(p, tc, m.id);
    Outcome o; int err = 0;
    const bool own_start = owned_z(p, m.z);
    LocalStats mls = {0, 0, 0, 0, 0, 0};  // statistics of redundantly evaluated halo molecules are not counted
    const bool guard = (m.sf & DF_CREATED_ON_SURF) != 0;
    SurfState ss = {MCX_NONE, MCX_NONE, 0.0, 0.0};
    if (SURF && (m.sf & DF_SURF)) { const double2 uv = p.suvA[i]; ss.wall = p.swallA[i]; ss.tile = p.stileA[i]; ss.u = uv.x; ss.v = uv.y; }
    evaluate_iteration<false, WITH_DISK, SURF>(p, m, i, t_sched, t_uni, (SURF && guard) ? p.swallA[i] : MCX_NONE,
                                               (SURF && guard) ? p.stileA[i] : MCX_NONE, ss, epoch, rs, false, o, mls, tc, err, grp);
    if (!lead) continue;
    if (!WITH_DISK && err == MCX_INTERNAL_NEEDS_DISK) {  // re-evaluated from scratch by the WITH_DISK launch
      if (tc.tr) tc.tr->rounds--;
      p.pend[1][agg_reserve(&p.ctr->n_disk, 1u)] = i;
      continue;
    }
    if (own_start) {
      if (p.species[species].flags & MCX_SP_CAN_DIFFUSE) msteps++;
      ls.ray_polygon_tests += mls.ray_polygon_tests; ls.ray_polygon_colls += mls.ray_polygon_colls;
      ls.reflections += mls.reflections; ls.transparent += mls.transparent;
      ls.volvol_collisions += mls.volvol_collisions; ls.redos += mls.redos;
    }
    trace_end(tc, o, rs);
    if (err && (own_start || err != MCX_ERR_ESCAPED)) raise_error(p, err, m.id);
    if (o.kind == MCX_OUT_MOVED || o.kind == MCX_OUT_STATIC) finalize_alive(p, i, o.pos, m.id, species, o.flags, o.t_now, o.unimol_time, &o);
    else if (o.kind == MCX_OUT_NONE) p.rank[i] = MCX_NONE;
    else write_proposal(p, i, o, m.id, species, epoch, &p.ctr->n_prop[0]);
  }
  flush_stats(p, ls, msteps);
}

// Conflict rounds.  Round r decides the proposals of list 0 (n_prop[r] entries) on the claims as they stand: winners
// commit, losers go to list 1 (n_lose[r]); the losers are re-evaluated against the updated DEAD flags and their new
// proposals form list 0 of round r + 1 (n_prop[r + 1]).  Every round has its own two counters, so nothing has to be
// reset between the phases (the four launches per round of round 1 became two phases of one kernel).
__device__ __forceinline__ void resolve_round(const DevParams& p, unsigned int round, BlockTally* tally) {
  tally_clear(tally);
  const unsigned int n = *(volatile unsigned int*)&p.ctr->n_prop[round];
  const unsigned int epoch = round_epoch(p, round);
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    uint32_t slot = __ldcg(p.pend[0] + k);
    MolRec e = load_rec_volatile(p.recB, slot);  // event position + identity
    uint32_t species = e.sf & SF_SPECIES_MASK;
    uint32_t info = __ldcg(p.prop_info + slot);
    int kind = info & 15, pathway = (info >> 4) & 0xFF, rxn_class = info >> 19;
    const uint32_t orient_bits = (info >> 12) & ORIENT_BITS_MASK;
    uint32_t partner = __ldcg(p.prop_partner + slot);
    unsigned long long key = claim_key(epoch, e.id);
    if (kind == MCX_OUT_SURFMOVE) key = weak_key(key);
    bool ok = __ldcg(p.claim + slot) == key;
    if (ok && partner_is_consumed(p, kind, rxn_class, pathway, species)) ok = __ldcg(p.claim + partner) == key;
    if (ok && kind == MCX_OUT_SURFMOVE) ok = __ldcg(p.tile_claim + p.grids[__ldcg(p.swallB + slot)].tile_start + __ldcg(p.stileB + slot)) == key;
    if (ok && kind == MCX_OUT_REACTED && p.classes[rxn_class].kind == MCX_RXN_BIMOL_SURFSURF) {  // the initiator took a new tile first
      const uint32_t nw = __ldcg(p.swallB + slot), nt = __ldcg(p.stileB + slot);
      if (nw != p.swallA[slot] || nt != p.stileA[slot]) ok = __ldcg(p.tile_claim + p.grids[nw].tile_start + nt) == key;
    }
    if (ok && event_is_general(p, kind, rxn_class, pathway)) {
      const uint32_t pm = __ldcg(p.prop_pmask + slot);
      for (uint32_t c = 0; ok && c < (pm & 15u); c++)
        if ((pm >> (4 + c)) & 1u) {
          const uint2 wt = p.prop_ptile[slot * MCX_MAX_PRODUCTS + c];
          ok = __ldcg(p.tile_claim + p.grids[wt.x].tile_start + wt.y) == key;
        }
    }
    if (ok) {
      commit_event(p, slot, kind, rxn_class, pathway, partner, __ldcg(p.prop_t + slot), D3{e.x, e.y, e.z}, e.id, species,
                   e.sf & ~SF_SPECIES_MASK, __ldcg(p.tschedB + slot), __ldcg(p.tuniB + slot), orient_bits, tally);
    } else {
      uint32_t q = agg_reserve(&p.ctr->n_lose[round], 1u);
      p.pend[1][q] = slot;
    }
  }
  tally_flush(tally, p.ctr);
}

template <bool SURF>
__device__ __forceinline__ void retry_round(const DevParams& p, unsigned int round, int forced, const ZigShared* zig) {
  const unsigned int n = *(volatile unsigned int*)&p.ctr->n_lose[round];
  const unsigned int epoch = round_epoch(p, round + 1);
  LocalStats ls = {0, 0, 0, 0, 0, 0};
  const int G = lanes_per_molecule(n);
  const Group grp = group_of(G);
  const bool lead = grp.sub == 0;
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x; base < (unsigned long long)n * G;
       base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long k = (base + threadIdx.x) / G;
    if (k >= n) continue;  // whole groups leave together
    uint32_t i = __ldcg(p.pend[1] + k);
    MolRec m = load_rec_volatile(p.recA, i);
    if (m.sf & DF_DEAD) {  // consumed as somebody's partner in this round
      if (lead) {
        p.rank[i] = MCX_NONE;
        if (p.trace && m.id < p.n_trace) p.trace[m.id].outcome = MCX_OUT_CONSUMED;
      }
      continue;
    }
    const bool own_start = owned_z(p, m.z);  // statistics of redundantly evaluated halo molecules are not counted
    if (lead && own_start) agg_add(&p.ctr->retries, 1u);
    if (lead && forced && own_start) agg_add(&p.ctr->unresolved, 1u);
    const uint32_t species = m.sf & SF_SPECIES_MASK;
    double t_sched = (m.sf & DF_PARTIAL) ? p.tschedA[i] : 0.0;
    double t_uni = (m.sf & DF_HAS_UNIMOL) ? p.tuniA[i] : MCX_TIME_INVALID;
    Stream rs; rs.init(p, m.id, zig);
    Tracer tc; tc.h = 0xcbf29ce484222325ULL; tc.tr = nullptr;
    if (lead) trace_begin(p, tc, m.id);
    Outcome o; int err = 0;
    LocalStats halo_ls = {0, 0, 0, 0, 0, 0};
    const bool guard = (m.sf & DF_CREATED_ON_SURF) != 0;
    SurfState ss = {MCX_NONE, MCX_NONE, 0.0, 0.0};
    if (SURF && (m.sf & DF_SURF)) { const double2 uv = p.suvA[i]; ss.wall = p.swallA[i]; ss.tile = p.stileA[i]; ss.u = uv.x; ss.v = uv.y; }
    evaluate_iteration<true, true, SURF>(p, m, i, t_sched, t_uni, (SURF && guard) ? p.swallA[i] : MCX_NONE,
                                         (SURF && guard) ? p.stileA[i] : MCX_NONE, ss, epoch, rs, forced != 0, o,
                                         (own_start && lead) ? ls : halo_ls, tc, err, grp);
    if (!lead) continue;
    trace_end(tc, o, rs);
    if (err) raise_error(p, err, m.id);
    if (o.kind == MCX_OUT_MOVED || o.kind == MCX_OUT_STATIC) finalize_alive(p, i, o.pos, m.id, species, o.flags, o.t_now, o.unimol_time, &o);
    else if (o.kind == MCX_OUT_NONE) p.rank[i] = MCX_NONE;
    else if (forced) {  // only self-claims can occur without partners: commit directly
      p.rank[i] = MCX_NONE;
      commit_event(p, i, o.kind, o.rxn_class, o.pathway, o.partner_slot, o.t_event, o.pos, m.id, species, o.flags, o.t_now,
                   o.unimol_time, o.orient_bits);
    } else write_proposal(p, i, o, m.id, species, epoch, &p.ctr->n_prop[round + 1]);
  }
  flush_stats(p, ls, 0);
}

// round 0 holds ~99 % of the proposals (1.3e6 at 1e8 molecules): its own launch with a grid sized for them
__global__ void __launch_bounds__(TPB) k_resolve0(const __grid_constant__ DevParams p) {
  __shared__ BlockTally tally;
  if (p.ctr->n_prop[0] <= blockIdx.x * blockDim.x) return;
  resolve_round(p, 0, &tally);
}
// everything after it in ONE cooperative launch: retry(r), grid barrier, resolve(r + 1), grid barrier, ... until a
// round leaves no losers — the launches of an iteration no longer depend on max_resolve_rounds and an iteration
// without conflicts pays one nearly empty kernel (32 launches of mostly empty kernels before, profiles/r02_a_launches)
template <bool SURF>
__global__ void __launch_bounds__(TPB, 2) k_rounds(const __grid_constant__ DevParams p) {
  __shared__ ZigShared zig;
  __shared__ BlockTally tally;
  if (*(volatile unsigned int*)&p.ctr->n_lose[0] == 0) return;  // no conflicts at all: every block leaves at once
  zig_load(&zig);
  __syncthreads();
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  for (unsigned int r = 0; r < p.max_rounds; r++) {
    if (*(volatile unsigned int*)&p.ctr->n_lose[r] == 0) break;  // same value in every block: final since the last barrier
    retry_round<SURF>(p, r, r + 1 == p.max_rounds ? 1 : 0, &zig);
    if (r + 1 == p.max_rounds) break;
    __threadfence();
    grid.sync();
    if (*(volatile unsigned int*)&p.ctr->n_prop[r + 1] == 0) break;
    resolve_round(p, r + 1, &tally);
    __threadfence();
    grid.sync();
  }
}

// ---- exclusive scan of the cell histogram (3 phases, 4096 cells per block) ---------------------------
#define SCAN_TPB 1024
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_TPB * SCAN_ITEMS)

__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* total, unsigned int* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int x = v;
  for (int o = 1; o < 32; o <<= 1) { unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) smem[warp] = x;
  __syncthreads();
  if (warp == 0) {
    unsigned int s = lane < (blockDim.x >> 5) ? smem[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) { unsigned int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    smem[lane] = s;
  }
  __syncthreads();
  unsigned int prefix = warp ? smem[warp - 1] : 0;
  *total = smem[(blockDim.x >> 5) - 1];
  __syncthreads();
  return prefix + x - v;
}

__global__ void __launch_bounds__(SCAN_TPB) k_scan_reduce(const uint32_t* __restrict__ cnt, unsigned int n, unsigned int* __restrict__ sums) {
  __shared__ unsigned int smem[32];
  unsigned int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  unsigned int s = 0;
  if (base + SCAN_ITEMS <= n) { uint4 v = *reinterpret_cast<const uint4*>(cnt + base); s = v.x + v.y + v.z + v.w; }
  else for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += cnt[base + k];
  unsigned int total;
  block_exclusive_scan(s, &total, smem);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_sums(unsigned int* sums, unsigned int nblocks, uint32_t* out_total, Counters* ctr) {
  __shared__ unsigned int smem[32];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned int base = 0; base < nblocks; base += SCAN_TPB) {
    unsigned int i = base + threadIdx.x;
    unsigned int v = i < nblocks ? sums[i] : 0;
    unsigned int total;
    unsigned int ex = block_exclusive_scan(v, &total, smem);
    if (i < nblocks) sums[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) { *out_total = carry; if (ctr) ctr->n_next = carry; }
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_apply(uint32_t* __restrict__ cnt, unsigned int n, const unsigned int* __restrict__ sums) {
  __shared__ unsigned int smem[32];
  unsigned int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  unsigned int v[SCAN_ITEMS];
  unsigned int s = 0;
  if (base + SCAN_ITEMS <= n) { uint4 q = *reinterpret_cast<const uint4*>(cnt + base); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
  else for (int k = 0; k < SCAN_ITEMS; k++) v[k] = base + k < n ? cnt[base + k] : 0;
  for (int k = 0; k < SCAN_ITEMS; k++) s += v[k];
  unsigned int total;
  unsigned int ex = block_exclusive_scan(s, &total, smem) + sums[blockIdx.x];
  unsigned int o[SCAN_ITEMS];
  for (int k = 0; k < SCAN_ITEMS; k++) { o[k] = ex; ex += v[k]; }
  if (base + SCAN_ITEMS <= n) *reinterpret_cast<uint4*>(cnt + base) = make_uint4(o[0], o[1], o[2], o[3]);
  else for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) cnt[base + k] = o[k];
}

// ---- counting-sort scatter B -> A ------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_scatter(const __grid_constant__ DevParams p) {
  // multi-GPU: the owned population is recounted here, per block in shared memory (one global atomic per record on
  // the same four words serialised the whole kernel in the L2 atomic unit: +1.5 ms at 5e7 records, profiles/r01_r)
  __shared__ unsigned int s_species[MCX_MAX_COUNTED];
  if (p.world > 1) { for (int k = threadIdx.x; k < MCX_MAX_COUNTED; k += blockDim.x) s_species[k] = 0; __syncthreads(); }
  const unsigned int n = p.ctr->n_slots + p.ctr->n_prod;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t r = p.rank[i];
    if (r == MCX_NONE) continue;
    double rx, ry, rz, rw;  // one 256-bit load, one 256-bit store per record
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(rx), "=d"(ry), "=d"(rz), "=d"(rw) : "l"(p.recB + i) : "memory");
    const double2 hi = make_double2(rz, rw);
    uint32_t sf = (uint32_t)((unsigned long long)__double_as_longlong(rw) >> 32);
    uint32_t cell = cell_of(p, rx, ry, rz);
    uint32_t dst = p.cs_next[cell] + r;
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p.recA + dst), "d"(rx), "d"(ry), "d"(rz), "d"(rw) : "memory");
    if (sf & DF_PARTIAL) p.tschedA[dst] = p.tschedB[i];
    if (sf & DF_HAS_UNIMOL) p.tuniA[dst] = p.tuniB[i];
    if (sf & (DF_SURF | DF_CREATED_ON_SURF)) {
      const uint32_t wi = p.swallB[i], ti = p.stileB[i];
      p.swallA[dst] = wi; p.stileA[dst] = ti;
      if (sf & DF_SURF) {
        p.suvA[dst] = p.suvB[i];
        if (!(sf & DF_DEAD)) p.tile_slot[p.grids[wi].tile_start + ti] = dst;  // Grid::molecules_per_tile of the next snapshot
        if (p.surfsurf && !p.wall_has_grid[wi]) p.wall_has_grid[wi] = 1;       // Wall::has_initialized_grid from now on
      }
    }
    // multi-GPU: recount the owned population (halo copies are not this rank's molecules)
    if (p.world > 1 && !(sf & DF_DEAD) && owned_z(p, hi.x)) atomicAdd(&s_species[(sf & SF_SPECIES_MASK) & (MCX_MAX_COUNTED - 1u)], 1u);
  }
  if (p.world > 1) {
    __syncthreads();
    for (int k = threadIdx.x; k < 256; k += blockDim.x)
      if (s_species[k]) atomicAdd(&p.ctr->species_next[k], (unsigned long long)s_species[k]);
  }
}
// Fresh molecule ids: event e of cell group g gets next_id + (fresh ids of the lower ranks) + (fresh ids of the lower
// groups: fresh_pref after the scan) + (fresh ids of the events of g whose initiator has a smaller id: chain of g).
__global__ void __launch_bounds__(TPB) k_assign_ids(const __grid_constant__ DevParams p) {
  const unsigned int n = min(p.ctr->n_fresh_events, p.fresh_cap);
  unsigned int base = p.ctr->next_id;
  if (p.rank_fresh) for (int r = 0; r < p.my_rank; r++) base += p.rank_fresh[r];
  for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const FreshEvent ev = p.fresh_list[e];
    unsigned int before = p.fresh_pref[ev.group];
    for (uint32_t q = p.fresh_head[ev.group]; q != MCX_NONE; q = p.fresh_list[q].next)
      if (p.fresh_list[q].init_id < ev.init_id) before += p.fresh_list[q].n;
    for (uint32_t k = 0; k < ev.n; k++) p.recB[ev.first_slot + k].id = base + before + k;
  }
}
__global__ void k_end_iteration(const __grid_constant__ DevParams p) {
  Counters* c = p.ctr;
  if (threadIdx.x == 0) {
    c->n_slots = c->n_next;
    unsigned int fresh = c->n_fresh_ids;
    if (p.rank_fresh) { fresh = 0; for (int r = 0; r < p.world; r++) fresh += p.rank_fresh[r]; }
    c->next_id += fresh; c->n_fresh_events = 0; c->n_fresh_ids = 0;
    c->n_prod = 0; c->n_disk = 0; c->n_slow = 0; c->n_second = 0; c->n_slow2 = 0; c->n_send[0] = 0; c->n_send[1] = 0;
  }
  if (threadIdx.x <= MCX_ROUNDS_MAX) { c->n_prop[threadIdx.x] = 0; c->n_lose[threadIdx.x] = 0; }
  if (p.world > 1) for (int k = threadIdx.x; k < MCX_MAX_COUNTED; k += blockDim.x) { c->species_count[k] = c->species_next[k]; c->species_next[k] = 0; }
}

// a new upload replaces the population, nothing else: cumulative reaction counts, statistics and next_id stay
__global__ void k_reset_population(const __grid_constant__ DevParams p, unsigned int n_slots) {
  Counters* c = p.ctr;
  if (threadIdx.x == 0) {
    c->n_slots = n_slots; c->n_prod = 0; c->n_disk = 0; c->n_next = 0; c->n_fresh_events = 0; c->n_fresh_ids = 0; c->error = 0; c->error_id = 0;
    c->n_emigrants[0] = 0; c->n_emigrants[1] = 0; c->n_slow = 0; c->n_send[0] = 0; c->n_send[1] = 0; c->n_second = 0; c->n_slow2 = 0;
  }
  for (int k = threadIdx.x; k < MCX_MAX_COUNTED; k += blockDim.x) { c->species_count[k] = 0; c->species_next[k] = 0; }
  if (threadIdx.x <= MCX_ROUNDS_MAX) { c->n_prop[threadIdx.x] = 0; c->n_lose[threadIdx.x] = 0; }
}
void mcx_launch_reset_population(const DevParams& p, unsigned int n_slots, cudaStream_t s) { k_reset_population<<<1, 256, 0, s>>>(p, n_slots); }

// initial binning of uploaded records (they sit in B, slots [0, n_slots))
__global__ void __launch_bounds__(TPB) k_bin_initial(const __grid_constant__ DevParams p) {
  const unsigned int n = p.ctr->n_slots;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    MolRec m = load_rec_volatile(p.recB, i);
    if (m.sf & DF_DEAD) { p.rank[i] = MCX_NONE; continue; }
    if (!in_partition(p, D3{m.x, m.y, m.z})) { raise_error(p, MCX_ERR_ESCAPED, m.id); p.rank[i] = MCX_NONE; continue; }
    if (!owned_z(p, m.z)) { p.rank[i] = MCX_NONE; continue; }  // multi-GPU: another rank's slab
    p.rank[i] = atomicAdd(&p.cs_next[cell_of(p, m.x, m.y, m.z)], 1u);
    // one atomic per (warp, species) instead of one per molecule on a handful of addresses
    const uint32_t spc = m.sf & SF_SPECIES_MASK;
    const unsigned int peers = __match_any_sync(__activemask(), spc);
    if (p.world == 1 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&p.ctr->species_count[spc], (unsigned long long)__popc(peers));
  }
}

// MolOrRxnCountEvent::compute_counts restricted to volumes (mol_or_rxn_count_event.cpp:607-716): molecules per
// (species, counted volume), one pass over the snapshot with per-warp aggregation of equal keys
__global__ void __launch_bounds__(TPB) k_count_by_volume(const __grid_constant__ DevParams p) {
  const unsigned int n = p.ctr->n_slots;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const MolRec m = load_rec_volatile(p.recA, i);
    if ((m.sf & DF_DEAD) || !owned_z(p, m.z)) continue;
    const unsigned int key = (m.sf & SF_SPECIES_MASK) * p.n_cv + (m.sf >> SF_CVI_SHIFT);
    const unsigned int peers = __match_any_sync(__activemask(), key);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&p.mol_count_cv[key], (unsigned long long)__popc(peers));
  }
}
// CountType::PresentOnSurfaceRegion (mol_or_rxn_count_event.cpp:528-534): surface molecules per (species, set of
// counted surface regions of their wall)
__global__ void __launch_bounds__(TPB) k_count_by_surface_region(const __grid_constant__ DevParams p) {
  const unsigned int n = p.ctr->n_slots;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const MolRec m = load_rec_volatile(p.recA, i);
    if ((m.sf & DF_DEAD) || !(m.sf & DF_SURF) || !owned_z(p, m.z)) continue;
    const unsigned int key = (m.sf & SF_SPECIES_MASK) * p.n_rs + __ldg(p.wall_rs + p.swallA[i]);
    const unsigned int peers = __match_any_sync(__activemask(), key);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&p.mol_count_rs[key], (unsigned long long)__popc(peers));
  }
}
void mcx_launch_count_by_surface_region(const DevParams& p, cudaStream_t s) {
  cudaMemsetAsync(p.mol_count_rs, 0, sizeof(unsigned long long) * (size_t)p.n_species * p.n_rs, s);
  if (p.has_surf) k_count_by_surface_region<<<p.sm_count * 8, TPB, 0, s>>>(p);
}
void mcx_launch_count_by_volume(const DevParams& p, cudaStream_t s) {
  cudaMemsetAsync(p.mol_count_cv, 0, sizeof(unsigned long long) * (size_t)p.n_species * p.n_cv, s);
  k_count_by_volume<<<p.sm_count * 8, TPB, 0, s>>>(p);
}

// is_point_inside_region_expr_recursively (release_event.cpp:787-813) on the membership mask of one ray cast: the two
// masks, or the postfix program of mcx_release::region_expr
__device__ __forceinline__ bool region_accepts(const mcx_release& r, uint32_t inside_mask) {
  if (r.region_expr_len == 0) return (inside_mask & r.region_in) == r.region_in && (inside_mask & r.region_out) == 0u;
  uint32_t stack = 0; int depth = 0;   // a stack of booleans, top = bit 0
  for (uint32_t q = 0; q < r.region_expr_len; q++) {
    const uint8_t op = r.region_expr[q];
    if (op < 32) { stack = (stack << 1) | ((inside_mask >> op) & 1u); depth++; continue; }
    const uint32_t b = stack & 1u, a = (stack >> 1) & 1u;
    const uint32_t v = op == MCX_REGION_UNION ? (a | b) : (op == MCX_REGION_INTERSECT ? (a & b) : (a & ~b & 1u));
    stack = ((stack >> 2) << 1) | v; depth--;
  }
  return depth == 1 && (stack & 1u);
}
// ---- release on the device (ReleaseEvent::release_ellipsoid_or_rectcuboid, src4/release_event.cpp:953-1003) --------
// One thread per new molecule: its own Philox stream (release domain), the reference's rejection loop and scaling,
// appended behind the re-binned population like a product.  Multi-GPU: every rank walks all ids, keeps its own slab.
__global__ void __launch_bounds__(TPB) k_release(const __grid_constant__ DevParams p, const mcx_release r, uint32_t first_id) {
  Counters* c = p.ctr;
  const bool spheroidal = r.shape == MCX_RELEASE_SPHERICAL || r.shape == MCX_RELEASE_SPHERICAL_SHELL;
  const double t_rel = r.release_time > (double)p.iteration ? r.release_time : (double)p.iteration;
  const uint32_t base_flags = DF_SCHED_UNIMOL | (t_rel > (double)p.iteration ? DF_PARTIAL : 0u) |
                              ((r.counted_volume_index << SF_CVI_SHIFT) & SF_CVI_MASK);
  for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < r.number;
       k += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t id = first_id + (uint32_t)k;
    Stream rs; rs.init_release(p, id, nullptr);
    D3 q;
    do {  // "Pick values in unit square, toss if not in unit circle"
      q.x = rs.dbl() - 0.5;
      q.y = rs.dbl() - 0.5;
      q.z = rs.dbl() - 0.5;
    } while (spheroidal && dot3(q, q) >= 0.25);
    if (r.shape == MCX_RELEASE_SPHERICAL_SHELL) {
      const double rad = sqrt(dot3(q, q)) * 2;
      if (rad == 0) q = D3{0.0, 0.0, 0.5};
      else q = D3{q.x / rad, q.y / rad, q.z / rad};
    }
    D3 pos = {q.x * r.diameter[0] + r.location[0], q.y * r.diameter[1] + r.location[1], q.z * r.diameter[2] + r.location[2]};
    uint32_t flags_k = base_flags;
    if (r.shape == MCX_RELEASE_REGION) {
      // ReleaseEvent::release_inside_regions (release_event.cpp:904-951): uniform in the bounding box, kept when the
      // point lies inside the region expression — here inside every object of region_in and outside every object
      // of region_out, found by one ray cast (scan_ray); the molecule's counted volume comes from the same ray
      int tries = 0;
      bool ok = false;
      for (;;) {
        bool inb = in_partition(p, pos);
        RayScan sc;
        if (inb) scan_ray(p, pos, rs, sc);
        if (inb && !sc.redo && region_accepts(r, sc.inside_mask)) {
          if (p.wall_cv) {
            uint32_t cvi = 0;
            if (sc.first_wall != MCX_NONE) { const uint32_t cv = __ldg(p.wall_cv + sc.first_wall); cvi = sc.first_side == W_FRONT ? (cv & 0xFFu) : (cv >> 8); }
            if (p.cv_mask) { const uint32_t kq = cv_lookup(p, sc.inside_mask & p.cv_all); if (kq != MCX_NONE) cvi = kq; }
            flags_k = (flags_k & ~SF_CVI_MASK) | (cvi << SF_CVI_SHIFT);
          }
          ok = true;
          break;
        }
        if (++tries >= 100000) break;
        q.x = rs.dbl() - 0.5; q.y = rs.dbl() - 0.5; q.z = rs.dbl() - 0.5;
        pos = D3{q.x * r.diameter[0] + r.location[0], q.y * r.diameter[1] + r.location[1], q.z * r.diameter[2] + r.location[2]};
      }
      if (!ok) { raise_error(p, MCX_ERR_INVALID_ARG, id); continue; }  // the region holds (almost) nothing of its box
    }
    if (!in_partition(p, pos)) { raise_error(p, MCX_ERR_ESCAPED, id); continue; }
    if (!owned_z(p, pos.z)) continue;
    const uint32_t ns = c->n_slots + agg_reserve(&c->n_prod, 1u);
    if (ns >= p.capacity) { raise_error(p, MCX_ERR_CAPACITY, id); continue; }
    if (base_flags & DF_PARTIAL) p.tschedB[ns] = t_rel;
    store_rec(p.recB, ns, pos, id, r.species | flags_k);
    p.rank[ns] = atomicAdd(&p.cs_next[cell_of(p, pos.x, pos.y, pos.z)], 1u);
    if (p.world == 1) agg_add(&c->species_count[r.species], 1u);
  }
}
// ---- surface molecules onto regions (ReleaseEvent::release_onto_regions, release_event.cpp:640-760; include/mcx.h) ----
// the new surface molecule on (wall, tile): appended behind the re-binned population like a product
__device__ void place_surface_molecule(const DevParams& p, const SurfRelease& r, uint32_t k, uint32_t wi, uint32_t tile) {
  Counters* c = p.ctr;
  const uint32_t id = r.first_id + k;
  Stream rs; rs.init_release(p, id, nullptr);
  rs.it_hi |= 0x40000000u;   // placement draws: a domain of their own (the tile picks use the release domain)
  { uint32_t o[4]; philox4x32_10(0u, rs.it_lo, rs.it_hi, rs.id, rs.k0, rs.k1, o); rs.b0 = o[0]; rs.b1 = o[1]; rs.b2 = o[2]; rs.b3 = o[3]; }
  double u, v;
  tile_uv(p, wi, tile, r.randomize_pos != 0, rs, u, v);
  int orient = r.orientation;
  if (orient == 0) orient = (rs.next() & 1u) ? 1 : -1;
  const DevWall& f = p.walls[wi];
  const D3 pos = {u * f.ux + v * f.vx + f.v0x, u * f.uy + v * f.vy + f.v0y, u * f.uz + v * f.vz + f.v0z};
  const double t_rel = r.release_time > (double)p.iteration ? r.release_time : (double)p.iteration;
  const uint32_t ns = c->n_slots + agg_reserve(&c->n_prod, 1u);
  if (ns >= p.capacity) { raise_error(p, MCX_ERR_CAPACITY, id); return; }
  const uint32_t flags = DF_SURF | DF_SCHED_UNIMOL | (orient > 0 ? DF_ORIENT_UP : 0u) | (t_rel > (double)p.iteration ? DF_PARTIAL : 0u);
  if (flags & DF_PARTIAL) p.tschedB[ns] = t_rel;
  p.swallB[ns] = wi; p.stileB[ns] = tile; p.suvB[ns] = make_double2(u, v);
  store_rec(p.recB, ns, pos, id, r.species | flags);
  p.rank[ns] = atomicAdd(&p.cs_next[cell_of(p, pos.x, pos.y, pos.z)], 1u);
  agg_add(&c->species_count[r.species], 1u);
}
// phase 0: every molecule without a tile picks one (draw number `round` of its stream) and bids for it with its index
__global__ void __launch_bounds__(TPB) k_srel_pick(const __grid_constant__ DevParams p, const SurfRelease r, const uint32_t* pend, unsigned int n_pend,
                                                   unsigned int round) {
  for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pend; q += gridDim.x * blockDim.x) {
    const uint32_t k = pend ? pend[q] : q;
    Stream rs; rs.init_release(p, r.first_id + k, nullptr);
    for (unsigned int d = 0; d < round; d++) (void)rs.next();
    double A = rs.dbl() * r.total_area;
    // cum_area_bisect_high (release_event.cpp:60-81)
    size_t low = 0, hi = r.n_walls - 1, mid;
    while (hi - low > 1) { mid = (hi + low) / 2; if (r.cum_area[mid] > A) hi = mid; else low = mid; }
    const size_t at = r.cum_area[low] > A ? low : hi;
    const uint32_t wi = r.walls[at];
    const double area = r.area[at];
    if (at != 0) A -= r.cum_area[at - 1];
    const DevGrid& g = p.grids[wi];
    const uint32_t nt = (uint32_t)(g.n_axis * g.n_axis);
    uint32_t tile = (uint32_t)((double)(g.n_axis * g.n_axis) * (A / area));
    if (tile >= nt) tile = nt - 1;
    const uint32_t gt = g.tile_start + tile;
    uint32_t pick = MCX_NONE;
    if (p.tile_slot[gt] == MCX_NONE && *(volatile uint32_t*)&r.claim[gt] == MCX_NONE) pick = gt;  // vacant in the snapshot and in earlier rounds
    r.choice[k] = pick; r.choice_wall[k] = wi;
  }
}
__global__ void __launch_bounds__(TPB) k_srel_bid(const __grid_constant__ DevParams p, const SurfRelease r, const uint32_t* pend, unsigned int n_pend) {
  for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pend; q += gridDim.x * blockDim.x) {
    const uint32_t k = pend ? pend[q] : q;
    const uint32_t gt = r.choice[k];
    if (gt != MCX_NONE) atomicMin(&r.claim[gt], k);
  }
}
// phase 1: the lowest bidder of a tile is placed there, everybody else goes on to the next round
__global__ void __launch_bounds__(TPB) k_srel_settle(const __grid_constant__ DevParams p, const SurfRelease r, const uint32_t* pend, unsigned int n_pend,
                                                     uint32_t* pend_out, unsigned int* n_out) {
  for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pend; q += gridDim.x * blockDim.x) {
    const uint32_t k = pend ? pend[q] : q;
    const uint32_t gt = r.choice[k];
    if (gt != MCX_NONE && r.claim[gt] == k) {
      const uint32_t wi = r.choice_wall[k];
      place_surface_molecule(p, r, k, wi, gt - p.grids[wi].tile_start);
    } else pend_out[agg_reserve(n_out, 1u)] = k;
  }
}
void mcx_launch_surface_release_round(const DevParams& p, const SurfRelease& r, const uint32_t* pend_in, unsigned int n_pend,
                                      uint32_t* pend_out, unsigned int* n_out, unsigned int round, cudaStream_t s) {
  unsigned int grid = (n_pend + TPB - 1) / TPB;
  if (grid == 0) grid = 1;
  if (grid > (unsigned int)p.sm_count * 16u) grid = (unsigned int)p.sm_count * 16u;
  cudaMemsetAsync(n_out, 0, sizeof(unsigned int), s);
  k_srel_pick<<<grid, TPB, 0, s>>>(p, r, pend_in, n_pend, round);
  k_srel_bid<<<grid, TPB, 0, s>>>(p, r, pend_in, n_pend);
  k_srel_settle<<<grid, TPB, 0, s>>>(p, r, pend_in, n_pend, pend_out, n_out);
}
// fall-back (release_event.cpp:702-744): the first vacant tiles in wall-list order, lowest index first
__global__ void k_srel_fill(const __grid_constant__ DevParams p, const SurfRelease r, const uint32_t* pend_sorted, unsigned int n_pend,
                            unsigned int* n_left) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned int q = 0;
  for (unsigned int a = 0; a < r.n_walls && q < n_pend; a++) {
    const uint32_t wi = r.walls[a];
    const DevGrid& g = p.grids[wi];
    const uint32_t nt = (uint32_t)(g.n_axis * g.n_axis);
    for (uint32_t tile = 0; tile < nt && q < n_pend; tile++) {
      const uint32_t gt = g.tile_start + tile;
      if (p.tile_slot[gt] != MCX_NONE || r.claim[gt] != MCX_NONE) continue;
      r.claim[gt] = pend_sorted[q];
      place_surface_molecule(p, r, pend_sorted[q], wi, tile);
      q++;
    }
  }
  *n_left = n_pend - q;
}
__global__ void __launch_bounds__(TPB) k_srel_count_vacant(const __grid_constant__ DevParams p, const SurfRelease r, unsigned int* n_vacant) {
  unsigned int mine = 0;
  for (unsigned int a = blockIdx.x; a < r.n_walls; a += gridDim.x) {
    const DevGrid& g = p.grids[r.walls[a]];
    const uint32_t nt = (uint32_t)(g.n_axis * g.n_axis);
    for (uint32_t tile = threadIdx.x; tile < nt; tile += blockDim.x) mine += p.tile_slot[g.tile_start + tile] == MCX_NONE ? 1u : 0u;
  }
  if (mine) atomicAdd(n_vacant, mine);
}
void mcx_launch_surface_release_count_vacant(const DevParams& p, const SurfRelease& r, unsigned int* n_vacant, cudaStream_t s) {
  cudaMemsetAsync(n_vacant, 0, sizeof(unsigned int), s);
  k_srel_count_vacant<<<std::max(1u, std::min(r.n_walls, (unsigned int)p.sm_count * 8u)), TPB, 0, s>>>(p, r, n_vacant);
}
void mcx_launch_surface_release_fill(const DevParams& p, const SurfRelease& r, const uint32_t* pend_sorted, unsigned int n_pend,
                                     unsigned int* n_left, cudaStream_t s) {
  k_srel_fill<<<1, 32, 0, s>>>(p, r, pend_sorted, n_pend, n_left);
}

// ReleaseEvent::release_list (release_event.cpp:1008-1040), volume molecules: one molecule per listed position
__global__ void __launch_bounds__(TPB) k_release_list(const __grid_constant__ DevParams p, const double* x, const double* y, const double* z,
                                                      const uint32_t* species, const uint32_t* cv, unsigned long long n, double release_time,
                                                      uint32_t first_id) {
  Counters* c = p.ctr;
  const double t_rel = release_time > (double)p.iteration ? release_time : (double)p.iteration;
  for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t id = first_id + (uint32_t)k;
    const uint32_t sp = species[k];
    if (sp >= (uint32_t)p.n_species || !(p.species[sp].flags & MCX_SP_VOL) || (cv && cv[k] >= p.n_cv)) { raise_error(p, MCX_ERR_INVALID_ARG, id); continue; }
    const D3 pos = {x[k], y[k], z[k]};
    if (!in_partition(p, pos)) { raise_error(p, MCX_ERR_ESCAPED, id); continue; }
    if (!owned_z(p, pos.z)) continue;
    const uint32_t ns = c->n_slots + agg_reserve(&c->n_prod, 1u);
    if (ns >= p.capacity) { raise_error(p, MCX_ERR_CAPACITY, id); continue; }
    const uint32_t flags = DF_SCHED_UNIMOL | (t_rel > (double)p.iteration ? DF_PARTIAL : 0u) | (cv ? (cv[k] << SF_CVI_SHIFT) & SF_CVI_MASK : 0u);
    if (flags & DF_PARTIAL) p.tschedB[ns] = t_rel;
    store_rec(p.recB, ns, pos, id, sp | flags);
    p.rank[ns] = atomicAdd(&p.cs_next[cell_of(p, pos.x, pos.y, pos.z)], 1u);
    if (p.world == 1) agg_add(&c->species_count[sp], 1u);
  }
}
void mcx_launch_release_list(const DevParams& p, const double* x, const double* y, const double* z, const uint32_t* species,
                             const uint32_t* cv, uint64_t n, double release_time, uint32_t first_id, cudaStream_t s) {
  unsigned long long grid = (n + TPB - 1) / TPB;
  if (grid == 0) grid = 1;
  if (grid > (unsigned long long)p.sm_count * 16ull) grid = (unsigned long long)p.sm_count * 16ull;
  k_release_list<<<(unsigned int)grid, TPB, 0, s>>>(p, x, y, z, species, cv, n, release_time, first_id);
}
void mcx_launch_release(const DevParams& p, const mcx_release& r, uint32_t first_id, cudaStream_t s) {
  unsigned long long grid = (r.number + TPB - 1) / TPB;
  if (grid == 0) grid = 1;
  if (grid > (unsigned long long)p.sm_count * 16ull) grid = (unsigned long long)p.sm_count * 16ull;
  k_release<<<(unsigned int)grid, TPB, 0, s>>>(p, r, first_id);
}

// ---- multi-GPU halo refresh (driven by mcx_comm.cu) --------------------------------------------------------------
// A -> B unchanged: lets the halo refresh run without an evaluation step (after an upload)
__global__ void __launch_bounds__(TPB) k_rebin(const __grid_constant__ DevParams p) {
  const unsigned int n = p.ctr->n_slots;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    MolRec m = load_rec_volatile(p.recA, i);
    if ((m.sf & DF_DEAD) || !owned_z(p, m.z)) { p.rank[i] = MCX_NONE; continue; }
    store_rec(p.recB, i, D3{m.x, m.y, m.z}, m.id, m.sf);
    if (m.sf & DF_PARTIAL) p.tschedB[i] = p.tschedA[i];
    if (m.sf & DF_HAS_UNIMOL) p.tuniB[i] = p.tuniA[i];
    if (m.sf & (DF_SURF | DF_CREATED_ON_SURF)) { p.swallB[i] = p.swallA[i]; p.stileB[i] = p.stileA[i]; if (m.sf & DF_SURF) p.suvB[i] = p.suvA[i]; }
    p.rank[i] = atomicAdd(&p.cs_next[cell_of(p, m.x, m.y, m.z)], 1u);
  }
}
// every record this rank keeps (new position owned) that lies within halo_layers of a slab face is copied to the
// neighbour, with its cold fields: the neighbour evaluates it next iteration exactly like the owner does
__global__ void __launch_bounds__(TPB) k_pack_halo(const __grid_constant__ DevParams p, HaloRec* send_low, HaloRec* send_high,
                                                   unsigned int cap) {
  const unsigned int n = p.ctr->n_slots + p.ctr->n_prod;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (p.rank[i] == MCX_NONE) continue;
    MolRec m = load_rec_volatile(p.recB, i);
    if (m.sf & DF_DEAD) continue;  // tombstone of a consumed partner: dropped by everybody next iteration
    const int cz = cell_z(p, m.z);
    const bool to_low = p.has_low && cz < p.own_lo + p.halo_layers;
    const bool to_high = p.has_high && cz >= p.own_hi - p.halo_layers;
    if (!to_low && !to_high) continue;
    HaloRec h;
    h.rec = m;
    h.tsched = (m.sf & DF_PARTIAL) ? p.tschedB[i] : 0.0;
    h.tuni = (m.sf & DF_HAS_UNIMOL) ? p.tuniB[i] : MCX_TIME_INVALID;
    h.swall = h.stile = MCX_NONE; h.su = h.sv = 0.0;
    if (m.sf & (DF_SURF | DF_CREATED_ON_SURF)) {
      h.swall = p.swallB[i]; h.stile = p.stileB[i];
      if (m.sf & DF_SURF) { const double2 uv = p.suvB[i]; h.su = uv.x; h.sv = uv.y; }
    }
    if (to_low) {
      unsigned int k = agg_reserve(&p.ctr->n_send[0], 1u);
      if (k < cap) send_low[k] = h; else raise_error(p, MCX_ERR_CAPACITY, m.id);
    }
    if (to_high) {
      unsigned int k = agg_reserve(&p.ctr->n_send[1], 1u);
      if (k < cap) send_high[k] = h; else raise_error(p, MCX_ERR_CAPACITY, m.id);
    }
  }
}
// received halo records are appended behind the local results and binned like products
__global__ void __launch_bounds__(TPB) k_unpack_halo(const __grid_constant__ DevParams p, const HaloRec* recv, unsigned int n,
                                                     unsigned int offset) {
  const unsigned int base = p.ctr->n_slots + p.ctr->n_prod + offset;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const unsigned int i = base + k;
    if (i >= p.capacity) { raise_error(p, MCX_ERR_CAPACITY, recv[k].rec.id); continue; }
    const HaloRec h = recv[k];
    store_rec(p.recB, i, D3{h.rec.x, h.rec.y, h.rec.z}, h.rec.id, h.rec.sf);
    if (h.rec.sf & DF_PARTIAL) p.tschedB[i] = h.tsched;
    if (h.rec.sf & DF_HAS_UNIMOL) p.tuniB[i] = h.tuni;
    if (h.rec.sf & (DF_SURF | DF_CREATED_ON_SURF)) { p.swallB[i] = h.swall; p.stileB[i] = h.stile; if (h.rec.sf & DF_SURF) p.suvB[i] = make_double2(h.su, h.sv); }
    p.rank[i] = atomicAdd(&p.cs_next[cell_of(p, h.rec.x, h.rec.y, h.rec.z)], 1u);
  }
}
__global__ void k_add_received(const __grid_constant__ DevParams p, unsigned int n) { p.ctr->n_prod += n; }

// ---- halo refresh over peer memory (mcx_comm.cu: exchange_halo_p2p) ---------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* q) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(q) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* q, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(q), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// k_pack_halo's selection, but every selected record is stored straight into the neighbour's receive buffer (remote
// stores over NVLink: one 256-bit store, plus one 128-bit store when the record carries cold fields).  The block that
// finishes last publishes the two counts: a system-scope release store of (tag << 32 | count) into the neighbour's
// flag word, ordered after every block's records by the fence + block-counter chain.
__global__ void __launch_bounds__(TPB) k_halo_pack_p2p(const __grid_constant__ DevParams p, const HaloP2P L) {
  // Slots in the neighbour's buffer are reserved once per block and trip: the halo zones are contiguous runs of the
  // sorted snapshot, so one warp-aggregated atomic per warp still put 3e5 returning atomics per side on ONE address
  // (the L2 atomic unit serialises them: ~1.5 ms of the 3.3 ms sort + halo phase at N = 4, profiles/r01_mg4b).
  __shared__ unsigned int s_cnt[2], s_base[2];
  const unsigned int n = p.ctr->n_slots + p.ctr->n_prod;
  const int lane = threadIdx.x & 31;
  for (unsigned int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const unsigned int i = base + threadIdx.x;
    bool to_side[2] = {false, false};
    MolRec m = {};
    if (i < n && p.rank[i] != MCX_NONE) {
      m = load_rec_volatile(p.recB, i);
      if (!(m.sf & DF_DEAD)) {  // a tombstone of a consumed partner is dropped by everybody next iteration
        const int cz = cell_z(p, m.z);
        to_side[0] = p.has_low && cz < p.own_lo + p.halo_layers;
        to_side[1] = p.has_high && cz >= p.own_hi - p.halo_layers;
      }
    }
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned int off[2] = {0, 0};
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const unsigned int mask = __ballot_sync(0xffffffffu, to_side[side]);
      unsigned int wbase = 0;
      if (lane == 0 && mask) wbase = atomicAdd(&s_cnt[side], (unsigned int)__popc(mask));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      off[side] = wbase + __popc(mask & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&p.ctr->n_send[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (to_side[0] || to_side[1]) {
      const bool cold = (m.sf & (DF_PARTIAL | DF_HAS_UNIMOL)) != 0;
      const double2 tt = cold ? make_double2((m.sf & DF_PARTIAL) ? p.tschedB[i] : 0.0, (m.sf & DF_HAS_UNIMOL) ? p.tuniB[i] : MCX_TIME_INVALID)
                              : make_double2(0.0, 0.0);
      const bool surf = (m.sf & (DF_SURF | DF_CREATED_ON_SURF)) != 0;  // Molecule::s, or where a volume product was created
      uint2 wt = make_uint2(MCX_NONE, MCX_NONE);
      double2 uv = make_double2(0.0, 0.0);
      if (surf) { wt = make_uint2(p.swallB[i], p.stileB[i]); if (m.sf & DF_SURF) uv = p.suvB[i]; }
#pragma unroll
      for (int side = 0; side < 2; side++) {
        if (!to_side[side]) continue;
        const unsigned int k = s_base[side] + off[side];
        if (k >= L.cap) { raise_error(p, MCX_ERR_CAPACITY, m.id); continue; }
        HaloRec* dst = L.peer_recv[side] + k;
        store_rec(&dst->rec, 0, D3{m.x, m.y, m.z}, m.id, m.sf);
        if (cold) *reinterpret_cast<double2*>(&dst->tsched) = tt;
        if (surf) { *reinterpret_cast<uint2*>(&dst->swall) = wt; *reinterpret_cast<double2*>(&dst->su) = uv; }
      }
    }
    __syncthreads();  // s_cnt / s_base are reused by the next trip
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = atomicAdd(L.done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence_system();
    for (int side = 0; side < 2; side++) {
      if (!L.peer_flag[side]) continue;
      const unsigned int cnt = min(atomicAdd(&p.ctr->n_send[side], 0u), L.cap);
      st_release_sys(L.peer_flag[side], ((unsigned long long)L.tag << 32) | cnt);
    }
    *L.done = 0;
  }
}
// the receiving side: wait for both neighbours' counts (acquire), then append their records behind the local results
// and bin them like products; the last block adds them to the population
__global__ void __launch_bounds__(TPB) k_halo_unpack_p2p(const __grid_constant__ DevParams p, const HaloP2P L) {
  __shared__ unsigned int s_n[2];
  if (threadIdx.x < 2) {
    const int side = threadIdx.x;
    unsigned int cnt = 0;
    if (side == 0 ? p.has_low : p.has_high) {
      const unsigned long long t0 = global_ns();
      for (;;) {
        const unsigned long long v = ld_acquire_sys(L.my_flag[side]);
        if ((unsigned int)(v >> 32) == L.tag) { cnt = (unsigned int)v; break; }
        if (global_ns() - t0 > 20000000000ull) { raise_error(p, MCX_ERR_COMM, 0); break; }  // 20 s: the neighbour is gone
        __nanosleep(200);
      }
    }
    s_n[side] = cnt;
  }
  __syncthreads();
  const unsigned int n0 = s_n[0], n1 = s_n[1];
  const unsigned int base = p.ctr->n_slots + p.ctr->n_prod;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n0 + n1; k += gridDim.x * blockDim.x) {
    const HaloRec* src = k < n0 ? L.my_recv[0] + k : L.my_recv[1] + (k - n0);
    const unsigned int i = base + k;
    const MolRec m = load_rec_volatile(&src->rec, 0);
    if (i >= p.capacity) { raise_error(p, MCX_ERR_CAPACITY, m.id); continue; }
    store_rec(p.recB, i, D3{m.x, m.y, m.z}, m.id, m.sf);
    if (m.sf & (DF_PARTIAL | DF_HAS_UNIMOL)) {
      const double2 tt = __ldcg(reinterpret_cast<const double2*>(&src->tsched));
      if (m.sf & DF_PARTIAL) p.tschedB[i] = tt.x;
      if (m.sf & DF_HAS_UNIMOL) p.tuniB[i] = tt.y;
    }
    if (m.sf & (DF_SURF | DF_CREATED_ON_SURF)) {
      const uint2 wt = __ldcg(reinterpret_cast<const uint2*>(&src->swall));
      p.swallB[i] = wt.x; p.stileB[i] = wt.y;
      if (m.sf & DF_SURF) p.suvB[i] = __ldcg(reinterpret_cast<const double2*>(&src->su));
    }
    p.rank[i] = atomicAdd(&p.cs_next[cell_of(p, m.x, m.y, m.z)], 1u);
  }
}
__global__ void k_halo_add_p2p(const __grid_constant__ DevParams p, const HaloP2P L) {
  unsigned int n = 0;
  if (p.has_low) n += (unsigned int)ld_acquire_sys(L.my_flag[0]);
  if (p.has_high) n += (unsigned int)ld_acquire_sys(L.my_flag[1]);
  p.ctr->n_prod += n;
}

// ---- SoA <-> record conversion at the ABI boundary ---------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_pack_soa(const __grid_constant__ DevParams p, const double* x, const double* y, const double* z,
                                                  const uint32_t* id, const uint32_t* species, const uint32_t* flags,
                                                  const double* tsched, const double* tuni, SurfSoa sv, unsigned int n) {
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t hf = flags ? flags[i] : 0;
    uint32_t sf = species[i] & SF_SPECIES_MASK;
    if (species[i] >= (uint32_t)p.n_species) { raise_error(p, MCX_ERR_INVALID_ARG, id[i]); continue; }
    D3 pos = {x[i], y[i], z[i]};
    const bool is_surf = !(p.species[species[i]].flags & MCX_SP_VOL);
    if (is_surf) {  // Partition::add_surface_molecule: position from the wall's uv frame (uv2xyz, geometry_utils.h:29-35)
      const uint32_t wi = sv.wall ? sv.wall[i] : MCX_NONE;
      if (!p.has_surf || wi >= (uint32_t)p.n_walls || sv.tile[i] >= (uint32_t)(p.grids[wi].n_axis * p.grids[wi].n_axis)) {
        raise_error(p, MCX_ERR_INVALID_ARG, id[i]); continue;
      }
      const DevWall& f = p.walls[wi];
      const double u = sv.u[i], v = sv.v[i];
      pos = D3{u * f.ux + v * f.vx + f.v0x, u * f.uy + v * f.vy + f.v0y, u * f.uz + v * f.vz + f.v0z};
      sf |= DF_SURF | (sv.orientation[i] > 0 ? DF_ORIENT_UP : 0u);
      p.swallB[i] = wi; p.stileB[i] = sv.tile[i]; p.suvB[i] = make_double2(u, v);
    }
    atomicMax(&p.ctr->next_id, id[i] + 1u);
    if (sv.cv && !is_surf) {
      if (sv.cv[i] >= p.n_cv) { raise_error(p, MCX_ERR_INVALID_ARG, id[i]); continue; }
      sf |= sv.cv[i] << SF_CVI_SHIFT;
    }
    if (hf & MCX_MOL_DEFUNCT) sf |= DF_DEAD;
    if (hf & MCX_MOL_SCHEDULE_UNIMOL) sf |= DF_SCHED_UNIMOL;
    if ((hf & MCX_MOL_CVI_PENDING) && !is_surf) sf |= DF_CVI_PENDING;
    if ((hf & MCX_MOL_PARTIAL) && tsched) { sf |= DF_PARTIAL; p.tschedB[i] = tsched[i]; }
    if (tuni && tuni[i] != MCX_TIME_INVALID) { sf |= DF_HAS_UNIMOL; p.tuniB[i] = tuni[i]; }
    store_rec(p.recB, i, pos, id[i], sf);
  }
}
__global__ void __launch_bounds__(TPB) k_unpack_soa(const __grid_constant__ DevParams p, double* x, double* y, double* z, uint32_t* id,
                                                    uint32_t* species, uint32_t* flags, double* tsched, double* tuni,
                                                    SurfSoaOut sv, unsigned int* n_out) {
  const unsigned int n = p.ctr->n_slots;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    MolRec m = load_rec_volatile(p.recA, i);
    if ((m.sf & DF_DEAD) || !owned_z(p, m.z)) continue;  // halo copies belong to the neighbour rank
    unsigned int k = atomicAdd(n_out, 1u);
    x[k] = m.x; y[k] = m.y; z[k] = m.z; id[k] = m.id; species[k] = m.sf & SF_SPECIES_MASK;
    uint32_t hf = 0;
    if (m.sf & DF_SCHED_UNIMOL) hf |= MCX_MOL_SCHEDULE_UNIMOL;
    if (m.sf & DF_PARTIAL) hf |= MCX_MOL_PARTIAL;
    if (m.sf & DF_CVI_PENDING) hf |= MCX_MOL_CVI_PENDING;
    if (flags) flags[k] = hf;
    if (tsched) tsched[k] = (m.sf & DF_PARTIAL) ? p.tschedA[i] : (double)p.iteration;
    if (tuni) tuni[k] = (m.sf & DF_HAS_UNIMOL) ? p.tuniA[i] : MCX_TIME_INVALID;
    if (sv.cv) sv.cv[k] = m.sf >> SF_CVI_SHIFT;
    if (sv.wall) {
      const bool is_surf = (m.sf & DF_SURF) != 0;
      sv.wall[k] = is_surf ? p.swallA[i] : MCX_NONE;
      sv.tile[k] = is_surf ? p.stileA[i] : MCX_NONE;
      sv.orientation[k] = is_surf ? ((m.sf & DF_ORIENT_UP) ? 1 : -1) : 0;
      const double2 uv = is_surf ? p.suvA[i] : make_double2(0.0, 0.0);
      sv.u[k] = uv.x; sv.v[k] = uv.y;
    }
  }
}

// ---- launchers -----------------------------------------------------------------------------------------------
static inline void count_launches(const StepPlan& plan, unsigned int n) { if (plan.launches) *plan.launches += n; }

void mcx_launch_sort(const DevParams& p, const StepPlan& plan, cudaStream_t s) {
  count_launches(plan, 5);
  const unsigned int n = p.n_cells + 1;  // last entry receives the total
  const unsigned int nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
  k_scan_reduce<<<nblocks, SCAN_TPB, 0, s>>>(p.cs_next, n, p.scan_sums);
  k_scan_sums<<<1, SCAN_TPB, 0, s>>>(p.scan_sums, nblocks, p.scan_sums + nblocks, p.ctr);
  k_scan_apply<<<nblocks, SCAN_TPB, 0, s>>>(p.cs_next, n, p.scan_sums);
  if (p.has_surf && p.n_tiles) cudaMemsetAsync(p.tile_slot, 0xFF, sizeof(uint32_t) * (size_t)p.n_tiles, s);
  k_scatter<<<plan.sm_count * 8, TPB, 0, s>>>(p);
  k_end_iteration<<<1, 256, 0, s>>>(p);
}

// cell histogram reset + fast/slow diffuse + conflict rounds: results sit in B with their ranks
void mcx_launch_evaluate(const DevParams& p, const StepPlan& plan, cudaStream_t s) {
  cudaMemsetAsync(p.cs_next, 0, sizeof(uint32_t) * (size_t)(p.n_cells + 1), s);
  if (plan.has_fresh) {
    cudaMemsetAsync(p.fresh_pref, 0, sizeof(uint32_t) * (size_t)(p.n_groups + 1), s);
    cudaMemsetAsync(p.fresh_head, 0xFF, sizeof(uint32_t) * (size_t)p.n_groups, s);
  }
  static const bool carveout_set = [] {  // tuning knob (profiles/): shared-memory carveout of the fast pass in percent
    if (const char* e = getenv("MCX_FAST_CARVEOUT")) {
      cudaFuncSetAttribute(k_diffuse_fast<0>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
      cudaFuncSetAttribute(k_diffuse_fast<1>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
    }
    return true;
  }();
  (void)carveout_set;
  if (plan.prof) cudaEventRecord(plan.prof[0], s);
  if (p.tile.enabled) {
    static const bool smem_set = [] {
      cudaFuncSetAttribute(k_diffuse_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_DYN_MAX);
      return true;
    }();
    (void)smem_set;
    k_diffuse_tile<<<plan.sm_count * TILE_BLOCKS_PER_SM, TILE_TPB, p.tile.smem_bytes, s>>>(p);
  } else {
    k_diffuse_fast<0><<<plan.sm_count * 2 * MCX_FAST_MINBLOCKS, TPB, 0, s>>>(p);
  }
  if (plan.prof) cudaEventRecord(plan.prof[4], s);
  k_diffuse_fast<1><<<plan.sm_count * MCX_FAST_MINBLOCKS, TPB, 0, s>>>(p);
  const int g_slow = plan.sm_count * 2 * MCX_SLOW_MINBLOCKS;
  if (p.has_surf) k_diffuse_slow<false, true><<<g_slow, TPB, 0, s>>>(p, 0);
  else k_diffuse_slow<false, false><<<g_slow, TPB, 0, s>>>(p, 0);
  if (p.has_surf) {
    k_diffuse_slow<false, true><<<g_slow, TPB, 0, s>>>(p, 1);
    k_diffuse_slow<true, true><<<g_slow, TPB, 0, s>>>(p, 0);
  } else {
    k_diffuse_slow<false, false><<<g_slow, TPB, 0, s>>>(p, 1);
    k_diffuse_slow<true, false><<<g_slow, TPB, 0, s>>>(p, 0);
  }
  if (plan.prof) cudaEventRecord(plan.prof[1], s);
  count_launches(plan, 5);
  if (plan.has_claims) {
    count_launches(plan, 2);
    k_resolve0<<<plan.sm_count * 8, TPB, 0, s>>>(p);
    // cooperative launch: all blocks resident (two per multiprocessor, the kernel's launch bounds)
    void* args[] = {(void*)&p};
    if (p.has_surf) cudaLaunchCooperativeKernel((const void*)k_rounds<true>, dim3(plan.sm_count * 2), dim3(TPB), args, 0, s);
    else cudaLaunchCooperativeKernel((const void*)k_rounds<false>, dim3(plan.sm_count * 2), dim3(TPB), args, 0, s);
  }
  if (plan.prof) cudaEventRecord(plan.prof[2], s);
}

void mcx_launch_fresh_scan(const DevParams& p, const StepPlan& plan, cudaStream_t s) {
  if (!plan.has_fresh) return;
  count_launches(plan, 3);
  const unsigned int n = p.n_groups + 1;
  const unsigned int nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
  k_scan_reduce<<<nblocks, SCAN_TPB, 0, s>>>(p.fresh_pref, n, p.scan_sums);
  k_scan_sums<<<1, SCAN_TPB, 0, s>>>(p.scan_sums, nblocks, p.scan_sums + nblocks, nullptr);
  k_scan_apply<<<nblocks, SCAN_TPB, 0, s>>>(p.fresh_pref, n, p.scan_sums);
}
void mcx_launch_assign_ids(const DevParams& p, const StepPlan& plan, cudaStream_t s) {
  if (!plan.has_fresh) return;
  count_launches(plan, 1);
  k_assign_ids<<<plan.sm_count * 2, TPB, 0, s>>>(p);
}

void mcx_launch_iteration(const DevParams& p, const StepPlan& plan, cudaStream_t s) {
  mcx_launch_evaluate(p, plan, s);
  mcx_launch_fresh_scan(p, plan, s);
  mcx_launch_assign_ids(p, plan, s);
  mcx_launch_sort(p, plan, s);
  if (plan.prof) cudaEventRecord(plan.prof[3], s);
}

void mcx_launch_initial_sort(const DevParams& p, const StepPlan& plan, cudaStream_t s) {
  cudaMemsetAsync(p.cs_next, 0, sizeof(uint32_t) * (size_t)(p.n_cells + 1), s);
  k_bin_initial<<<plan.sm_count * 8, TPB, 0, s>>>(p);
  count_launches(plan, 1);
  mcx_launch_sort(p, plan, s);
}

void mcx_launch_rebin(const DevParams& p, const StepPlan& plan, cudaStream_t s) {
  cudaMemsetAsync(p.cs_next, 0, sizeof(uint32_t) * (size_t)(p.n_cells + 1), s);
  k_rebin<<<plan.sm_count * 8, TPB, 0, s>>>(p);
  count_launches(plan, 1);
}
void mcx_launch_pack_halo(const DevParams& p, HaloRec* send_low, HaloRec* send_high, unsigned int cap, cudaStream_t s) {
  k_pack_halo<<<p.sm_count * 8, TPB, 0, s>>>(p, send_low, send_high, cap);
}
void mcx_launch_unpack_halo(const DevParams& p, const HaloRec* recv, unsigned int n, unsigned int offset, cudaStream_t s) {
  if (n == 0) return;
  unsigned int grid = (n + TPB - 1) / TPB;
  if (grid > (unsigned int)p.sm_count * 16u) grid = (unsigned int)p.sm_count * 16u;
  k_unpack_halo<<<grid, TPB, 0, s>>>(p, recv, n, offset);
}
void mcx_launch_add_received(const DevParams& p, unsigned int n, cudaStream_t s) { k_add_received<<<1, 1, 0, s>>>(p, n); }
void mcx_launch_halo_p2p(const DevParams& p, const HaloP2P& link, cudaStream_t s) {
  k_halo_pack_p2p<<<p.sm_count * 8, TPB, 0, s>>>(p, link);
  k_halo_unpack_p2p<<<p.sm_count * 4, TPB, 0, s>>>(p, link);
  k_halo_add_p2p<<<1, 1, 0, s>>>(p, link);
}

void mcx_launch_pack_soa(const DevParams& p, const double* x, const double* y, const double* z, const uint32_t* id,
                         const uint32_t* species, const uint32_t* flags, const double* tsched, const double* tuni,
                         SurfSoa sv, unsigned int n, cudaStream_t s) {
  unsigned int grid = (n + TPB - 1) / TPB;
  if (grid == 0) grid = 1;
  if (grid > 65535u * 16u) grid = 65535u * 16u;
  k_pack_soa<<<grid, TPB, 0, s>>>(p, x, y, z, id, species, flags, tsched, tuni, sv, n);
}
void mcx_launch_unpack_soa(const DevParams& p, double* x, double* y, double* z, uint32_t* id, uint32_t* species,
                           uint32_t* flags, double* tsched, double* tuni, SurfSoaOut sv, unsigned int* n_out, cudaStream_t s) {
  cudaMemsetAsync(n_out, 0, sizeof(unsigned int), s);
  k_unpack_soa<<<p.sm_count * 8, TPB, 0, s>>>(p, x, y, z, id, species, flags, tsched, tuni, sv, n_out);
}
