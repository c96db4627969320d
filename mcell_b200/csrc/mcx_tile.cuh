// mcx_tile.cuh — shared-memory tiles for the partner probe of the fast pass (k_diffuse_tile in mcx_kernels.cu).
//
// What the reference does per molecule — collide_mol against every molecule of the collected subpartitions
// (src4/collision_utils.inl:464-552, diffuse_react_event.cpp:627-780) — the flat fast pass realises as a gather of
// candidate records from the cell-sorted snapshot (probe_partners_flat): 13 records per molecule from L1/L2, the wait
// for them a quarter of the kernel's stall samples (profiles/r02_a_*).  Here a thread block owns a box of cells:
//
//   1. one thread-issued bulk copy (cp.async.bulk, TMA) per cell row brings the records of the box plus a halo into
//      shared memory — rows are contiguous runs of the sorted snapshot, so a tile is (TY + 2) x (TZ + 2) plain 1-D
//      copies signalled on one mbarrier; the copies of the NEXT tile are issued as soon as the staging buffer is free,
//      so they run under the evaluation of the current tile;
//   2. the staged records are re-binned in shared memory (counting sort) into a grid finer than the global one
//      (cell rows split 2^lsy x 2^lsz) as 16-byte entries: fp32 position relative to the tile, staging index, species;
//   3. every owned molecule walks the fine cells under its swept box and pre-filters the entries in fp32 with
//      one-sided slack (longer, fatter capsule: no true partner can fail it); the few survivors (0.2 per molecule) take
//      the exact fp64 collide_mol on the snapshot record, so the hit set — and with it every decision downstream — is
//      bit for bit the flat probe's.
//
// A molecule whose swept box leaves the staged region (long moves, 1 % of them) and the molecules of a tile whose
// records do not fit the staging buffer go to the second pass, which probes with the gather walk.
#pragma once
#include "mcx_device.cuh"

#ifndef TILE_TPB
#define TILE_TPB 1024          // threads of a tile block; 512: two resident blocks per multiprocessor
#endif
#define TILE_BLOCKS_PER_SM (1024 / TILE_TPB)
#define TILE_ROWS_MAX 64          // (TY + 2) * (TZ + 2)
#define TILE_NF_MAX 7168          // fine cells of a staged tile (uint16 starts)
#define TILE_Q 4                  // survivors of the pre-filter per molecule (= MCX_FAST_MAX_HITS)
#define TILE_REC_PER_THREAD 4     // cap <= TILE_TPB * TILE_REC_PER_THREAD

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) --------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
  uint32_t done;
  do {  // try_wait suspends the thread in hardware for a while before it returns false
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  } while (!done);
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- per-tile tables (two copies: the current tile and the one being prefetched) -------------------------------------
struct TileTab {
  uint32_t row_src[TILE_ROWS_MAX];       // snapshot slot of the row's first staged record
  uint32_t row_off[TILE_ROWS_MAX + 1];   // staging index of it (prefix of the row counts); [n_rows] = total
  uint32_t own_lo[TILE_ROWS_MAX];        // staging index of the row's first OWNED record (rows of the halo: none)
  uint32_t own_pref[TILE_ROWS_MAX + 1];  // prefix of the owned counts
  uint32_t total, n_owned, overflow;
  int xs0, fy_off, fz_off;               // staged origin: first x cell, first fine row in y / z (may be negative)
  float pad_;
  double ox, oy, oz;                     // position the fp32 entries are relative to
};

struct TileSmem {
  MolRec* raw;                 // staged records as the bulk copies deliver them
  float4* binned;              // fine-cell sorted entries: x, y, z relative to (ox, oy, oz); w = staging index | species << 16
  uint16_t* fstart;            // per fine cell: count, then first entry (exclusive scan); nf + 1 entries used
  uint32_t* q;                 // [TILE_Q][TILE_TPB] survivors of the pre-filter, then the confirmed hits' slots
  TileTab* tab;                // [2]
  unsigned long long* bar;     // mbarrier of the staging buffer
  unsigned int* scan_part;     // [32] block-scan partials
  uint32_t* rmask;             // [32] per species: bit s = has a volume-volume reaction class with species s
};

__host__ __device__ __forceinline__ unsigned int tile_smem_layout(const TileGeom& g, unsigned char* base, TileSmem& t) {
  size_t o = 0;
  auto take = [&](size_t bytes, size_t align) { o = (o + align - 1) / align * align; void* r = base ? base + o : nullptr; o += bytes; return r; };
  t.raw = (MolRec*)take(sizeof(MolRec) * (size_t)g.cap, 128);
  t.binned = (float4*)take(sizeof(float4) * (size_t)g.cap, 16);
  t.fstart = (uint16_t*)take(sizeof(uint16_t) * ((size_t)g.nfx * g.nfy * g.nfz + 8), 16);
  t.q = (uint32_t*)take(sizeof(uint32_t) * TILE_Q * TILE_TPB, 16);
  t.tab = (TileTab*)take(sizeof(TileTab) * 2, 16);
  t.bar = (unsigned long long*)take(8, 8);
  t.scan_part = (unsigned int*)take(4 * 32, 4);
  t.rmask = (uint32_t*)take(4 * 32, 4);
  return (unsigned int)o;
}

// tile index -> tables + bulk copies of its rows; executed by one whole warp
__device__ __forceinline__ void tile_issue(const DevParams& p, const TileSmem& ts, unsigned int tile, TileTab* tab) {
  const TileGeom& g = p.tile;
  const int lane = threadIdx.x & 31;
  const int tx = (int)(tile % (unsigned int)g.ntx), ty = (int)((tile / (unsigned int)g.ntx) % (unsigned int)g.nty),
            tz = (int)(tile / ((unsigned int)g.ntx * (unsigned int)g.nty));
  const int X0 = tx * g.TX, X1 = min(X0 + g.TX, p.ncx);
  const int xs0 = max(X0 - g.hx, 0), xs1 = min(X1 + g.hx, p.ncx);
  const int gy0 = ty * g.TY - 1, gz0 = tz * g.TZ - 1;
  const int sy = g.TY + 2, n_rows = sy * (g.TZ + 2);
  unsigned int carry = 0, ocarry = 0;
  for (int r0 = 0; r0 < n_rows; r0 += 32) {
    const int r = r0 + lane;
    const int ry = r % sy, rz = r / sy;
    const int gy = gy0 + ry, gz = gz0 + rz;
    const bool valid = r < n_rows && gy >= 0 && gy < p.ncy && gz >= 0 && gz < p.ncz;
    const bool owned = valid && ry >= 1 && ry <= g.TY && rz >= 1 && rz <= g.TZ;
    const uint32_t base = valid ? row_base(p, gy, gz) : 0u;
    const uint32_t a = valid ? __ldg(p.cs_cur + base + xs0) : 0u, e = valid ? __ldg(p.cs_cur + base + xs1) : 0u;
    const uint32_t oa = owned ? __ldg(p.cs_cur + base + X0) : a, oe = owned ? __ldg(p.cs_cur + base + X1) : a;
    const unsigned int cnt = e - a, ocnt = oe - oa;
    unsigned int inc = cnt, oinc = ocnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int y = __shfl_up_sync(0xffffffffu, inc, o), oy = __shfl_up_sync(0xffffffffu, oinc, o);
      if (lane >= o) { inc += y; oinc += oy; }
    }
    if (r < n_rows) {
      tab->row_src[r] = a;
      tab->row_off[r] = carry + inc - cnt;
      tab->own_lo[r] = carry + inc - cnt + (oa - a);
      tab->own_pref[r] = ocarry + oinc - ocnt;
    }
    carry += __shfl_sync(0xffffffffu, inc, 31);
    ocarry += __shfl_sync(0xffffffffu, oinc, 31);
  }
  const bool overflow = carry > g.cap;
  if (lane == 0) {
    tab->row_off[n_rows] = carry; tab->own_pref[n_rows] = ocarry;
    tab->total = carry; tab->n_owned = ocarry; tab->overflow = overflow ? 1u : 0u;
    tab->xs0 = xs0; tab->fy_off = gy0 * (1 << g.lsy); tab->fz_off = (gz0 + p.z_off) * (1 << g.lsz);
    tab->ox = p.cgx + (double)xs0 / p.cell_rcp_x; tab->oy = p.cgy + (double)gy0 / p.cell_rcp_y;
    tab->oz = p.cgz + (double)(gz0 + p.z_off) / p.cell_rcp_z;
  }
  __syncwarp();
  // the phase of the barrier completes when this arrival and all announced bytes are in
  if (lane == 0) mbar_arrive_expect_tx(ts.bar, overflow ? 0u : carry * (unsigned int)sizeof(MolRec));
  if (!overflow) {
    for (int r = lane; r < n_rows; r += 32) {
      const unsigned int cnt = tab->row_off[r + 1] - tab->row_off[r];
      if (cnt) bulk_g2s(ts.raw + tab->row_off[r], p.recA + tab->row_src[r], cnt * (unsigned int)sizeof(MolRec), ts.bar);
    }
  }
}

// fine cell coordinate along one axis: floor(t * 2^ls) of the global cell coordinate t = (v - origin) * rcp — the
// scaling by a power of two is exact, so fine >> ls is the global cell of cell_coord() and the staged rows hold every
// record whose fine cell lies in the staged fine range
__device__ __forceinline__ int fine_coord(double v, double origin, double rcp, int n_cells, int ls) {
  const int c = (int)floor(((v - origin) * rcp) * (double)(1 << ls));
  const int hi = (n_cells << ls) - 1;
  return c < 0 ? 0 : (c > hi ? hi : c);
}
__device__ __forceinline__ int fine_z(const DevParams& p, double z) {  // like cell_z(): global layers, minus this rank's offset below
  const int ls = p.tile.lsz;
  const int c = (int)floor(((z - p.cgz) * p.cell_rcp_z) * (double)(1 << ls));
  const int lo = p.z_off * (1 << ls), hi = (p.z_off + p.ncz) * (1 << ls) - 1;
  return c < lo ? lo : (c > hi ? hi : c);
}

// steps 2 of the header: counting sort of the staged records into the fine grid; all threads of the block
__device__ __forceinline__ void tile_bin(const DevParams& p, const TileSmem& ts, const TileTab* tab) {
  const TileGeom& g = p.tile;
  const unsigned int total = tab->total;
  const int nf = g.nfx * g.nfy * g.nfz;
  uint32_t* cnt32 = reinterpret_cast<uint32_t*>(ts.fstart);
  for (int w = threadIdx.x; w < (nf + 2) / 2 + 1; w += TILE_TPB) cnt32[w] = 0u;
  __syncthreads();
  uint32_t keep[TILE_REC_PER_THREAD];
#pragma unroll
  for (int q = 0; q < TILE_REC_PER_THREAD; q++) {
    const unsigned int i = threadIdx.x + q * TILE_TPB;
    keep[q] = MCX_NONE;
    if (i < total) {
      const double4 r = *reinterpret_cast<const double4*>(ts.raw + i);
      const uint32_t sf = (uint32_t)((unsigned long long)__double_as_longlong(r.w) >> 32);
      if (!(sf & DF_DEAD)) {
        int fx = cell_coord(r.x, p.cgx, p.cell_rcp_x, p.ncx) - tab->xs0;
        int fy = fine_coord(r.y, p.cgy, p.cell_rcp_y, p.ncy, g.lsy) - tab->fy_off;
        int fz = fine_z(p, r.z) - tab->fz_off;
        fx = min(max(fx, 0), g.nfx - 1); fy = min(max(fy, 0), g.nfy - 1); fz = min(max(fz, 0), g.nfz - 1);
        const uint32_t f = (uint32_t)((fz * g.nfy + fy) * g.nfx + fx);
        const uint32_t old = atomicAdd(&cnt32[f >> 1], (f & 1u) ? 0x10000u : 1u);
        keep[q] = f | (((f & 1u) ? (old >> 16) : (old & 0xFFFFu)) << 16);
      }
    }
  }
  __syncthreads();
  // exclusive scan of the nf counts, in place (entry nf receives the total)
  {
    const int per = (nf + 1 + TILE_TPB - 1) / TILE_TPB;
    const int b = threadIdx.x * per, e = min(b + per, nf + 1);
    unsigned int s = 0;
    for (int k = b; k < e; k++) s += ts.fstart[k];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) ts.scan_part[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned int v = ts.scan_part[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += y; }
      ts.scan_part[lane] = v;
    }
    __syncthreads();
    unsigned int run = (warp ? ts.scan_part[warp - 1] : 0u) + x - s;
    for (int k = b; k < e; k++) { const unsigned int c = ts.fstart[k]; ts.fstart[k] = (uint16_t)run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < TILE_REC_PER_THREAD; q++) {
    if (keep[q] == MCX_NONE) continue;
    const unsigned int i = threadIdx.x + q * TILE_TPB;
    const double4 r = *reinterpret_cast<const double4*>(ts.raw + i);
    const uint32_t sf = (uint32_t)((unsigned long long)__double_as_longlong(r.w) >> 32);
    const unsigned int dst = (unsigned int)ts.fstart[keep[q] & 0xFFFFu] + (keep[q] >> 16);
    ts.binned[dst] = make_float4((float)(r.x - tab->ox), (float)(r.y - tab->oy), (float)(r.z - tab->oz),
                                 __uint_as_float(i | ((sf & SF_SPECIES_MASK) << 16)));
  }
  __syncthreads();
}

// The probe of one owned molecule (step 3 of the header).  own_idx: its staging index; react_mask: bit s set when the
// molecule has a volume-volume reaction class with species s (all ones when the model has more than 32 species).
struct TileProbe {
  static constexpr int HIT_STRIDE = TILE_TPB;
  const TileSmem* ts;
  const TileTab* tab;
  uint32_t own_idx;
  __device__ __forceinline__ const uint32_t* hit_slots() const { return ts->q + threadIdx.x; }

  __device__ __forceinline__ int run(const DevParams& p, bool probing, D3 pos, D3 disp, uint32_t self_id, uint32_t self_species,
                                     uint32_t, bool& overflow, bool& outside) {
    const TileGeom& g = p.tile;
    overflow = false;
    // fine cells under the swept box (segment inflated by R, padded like swept_cells())
    const double pad = p.R * (1.0 + 1e-9) + 1e-9;
    const int fx0 = cell_coord(fmin(pos.x, pos.x + disp.x) - pad, p.cgx, p.cell_rcp_x, p.ncx) - tab->xs0;
    const int fx1 = cell_coord(fmax(pos.x, pos.x + disp.x) + pad, p.cgx, p.cell_rcp_x, p.ncx) - tab->xs0;
    const int fy0 = fine_coord(fmin(pos.y, pos.y + disp.y) - pad, p.cgy, p.cell_rcp_y, p.ncy, g.lsy) - tab->fy_off;
    const int fy1 = fine_coord(fmax(pos.y, pos.y + disp.y) + pad, p.cgy, p.cell_rcp_y, p.ncy, g.lsy) - tab->fy_off;
    const int fz0 = fine_z(p, fmin(pos.z, pos.z + disp.z) - pad) - tab->fz_off;
    const int fz1 = fine_z(p, fmax(pos.z, pos.z + disp.z) + pad) - tab->fz_off;
    outside = probing && (tab->overflow || fx0 < 0 || fy0 < 0 || fz0 < 0 || fx1 >= g.nfx || fy1 >= g.nfy || fz1 >= g.nfz);
    const bool en = probing && !outside;
    // the move in fp32, relative to the tile
    const float px = (float)(pos.x - tab->ox), py = (float)(pos.y - tab->oy), pz = (float)(pos.z - tab->oz);
    const float vx = (float)disp.x, vy = (float)disp.y, vz = (float)disp.z;
    const float m = vx * vx + vy * vy + vz * vz;
    const float d_lo = -g.tol_d, d_hi = m + g.tol_d, rhs = m * g.r2p;
    const uint32_t react_mask = p.n_species <= 32 ? ts->rmask[self_species & 31u] : 0xFFFFFFFFu;
    const int ny = fy1 - fy0 + 1, nrows = en ? ny * (fz1 - fz0 + 1) : 0;
    const int span = fx1 - fx0 + 1;
    uint32_t* const q = ts->q + threadIdx.x;
    // One flat loop: a trip first moves to the next fine row when the current one is used up, then tests one entry —
    // both steps predicated, so the lanes of a warp stay together and a lane needs sum(max(entries of row, 1)) trips
    // (written as two alternative branches the loop ran with 6 of 32 lanes active, profiles/r02_d).
    int nq = 0, row = 0, ry = 0;
    int base = (fz0 * g.nfy + fy0) * g.nfx + fx0;  // first fine cell of the next row
    const int wrap = (g.nfy - ny) * g.nfx;
    unsigned int j = 0, jend = 0;
    bool q_over = false;
    for (;;) {
      if (j >= jend) {
        if (row >= nrows) break;
        j = ts->fstart[base]; jend = ts->fstart[base + span];
        row++; base += g.nfx;
        if (++ry == ny) { ry = 0; base += wrap; }
      }
      const bool has = j < jend;
      const float4 c = ts->binned[has ? j : 0u];
      j += has ? 1u : 0u;
      const float dx = c.x - px, dy = c.y - py, dz = c.z - pz;
      const float d = dx * vx + dy * vy + dz * vz;
      const float dd = dx * dx + dy * dy + dz * dz;
      const uint32_t meta = __float_as_uint(c.w);
      const bool pass = has && d >= d_lo && d <= d_hi && m * dd - d * d <= rhs && ((react_mask >> ((meta >> 16) & 31u)) & 1u) &&
                        (meta & 0xFFFFu) != own_idx;
      if (pass) {
        if (nq < TILE_Q) q[nq * TILE_TPB] = meta & 0xFFFFu; else q_over = true;
        nq += nq < TILE_Q ? 1 : 0;
      }
    }
    outside = outside || q_over;
    // exact collide_mol (collision_utils.inl:464-515) of the survivors, on the snapshot records
    const double movelen2 = dot3(disp, disp), rhs64 = movelen2 * (p.R * p.R);
    int found = 0;
    const int n_rows = (g.TY + 2) * (g.TZ + 2);
    for (int k = 0; k < nq; k++) {
      const uint32_t idx = q[k * TILE_TPB];
      int lo = 0, hi = n_rows;  // row of a staging index: last row whose first index is <= idx
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tab->row_off[mid] <= idx) lo = mid; else hi = mid; }
      const uint32_t slot = tab->row_src[lo] + (idx - tab->row_off[lo]);
      const MolRec c = load_rec(p.recA, slot);
      double d;
      if (collide_mol_hit(c, pos, disp, movelen2, rhs64, self_id, d) &&
          __ldg(p.bimol + self_species * p.n_species + (c.sf & SF_SPECIES_MASK)) >= 0)
        q[found++ * TILE_TPB] = slot;
    }
    return q_over ? 0 : found;
  }
};
