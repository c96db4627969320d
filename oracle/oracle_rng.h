// oracle/oracle_rng.h — TEST INFRASTRUCTURE (CPU oracle; never linked into the product).
//
// Restatement of the random number generators on MCell4's hot path:
//   * ISAAC64 (Bob Jenkins, public domain) as configured by the reference:
//       src/isaac64.h:25-75 (RANDSIZL=8, 32-bit words consumed from the END of the
//       512-word block backwards), src/isaac64.c:22-155 (mix, generate, init).
//   * rng_dbl = word * 2^-32                     src/isaac64.h:35,71-75
//   * rng_gauss: 128-strip Ziggurat on 32-bit words   src/rng.c:173-218
//     (tables = oracle/zig_tables.inc, dumped from the reference build by gen_zig_tables.py)
//   * Philox4x32-10 (Salmon et al., SC'11 — "Parallel random numbers: as easy as 1,2,3"):
//     the product's per-molecule counter-based stream; restated here independently.
// Pinned by tests/test_oracle_rng.py against oracle/_ref/librefrng.so (the reference's own
// rng.c compiled here) and against the SURVEY §A.3 known-answer vectors in tests/golden/.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "zig_tables.inc"

namespace orc {

static const double ZIG_Y[128] = MCX_ZIG_YTAB_INIT;
static const double ZIG_W[128] = MCX_ZIG_WTAB_INIT;
static const uint64_t ZIG_K[128] = MCX_ZIG_KTAB_INIT;
static const double ZIG_R = MCX_ZIG_R;

// ---------------------------------------------------------------- ISAAC64
struct Isaac64 {
  static const int SIZL = 8, SIZ = 1 << SIZL, WORDS = 2 * SIZ;
  uint32_t randcnt;
  uint64_t aa, bb, cc;
  uint64_t rsl[SIZ];
  uint64_t mm[SIZ];
  uint64_t blocks;

  static inline void mix(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d, uint64_t& e,
                         uint64_t& f, uint64_t& g, uint64_t& h) {
    a -= e; f ^= h >> 9;  h += a;
    b -= f; g ^= a << 9;  a += b;
    c -= g; h ^= b >> 23; b += c;
    d -= h; a ^= c << 15; c += d;
    e -= a; b ^= d >> 14; d += e;
    f -= b; c ^= e << 20; e += f;
    g -= c; d ^= f >> 17; f += g;
    h -= d; e ^= g << 14; g += h;
  }
  inline uint64_t ind(uint64_t x) const { return mm[(x >> 3) & (SIZ - 1)]; }

  void generate() {
    uint64_t a = aa, b = bb + (++cc), x, y;
    for (int i = 0; i < SIZ; i++) {
      int i2 = (i + SIZ / 2) & (SIZ - 1);
      x = mm[i];
      switch (i & 3) {
        case 0: a = ~(a ^ (a << 21)) + mm[i2]; break;
        case 1: a = (a ^ (a >> 5)) + mm[i2]; break;
        case 2: a = (a ^ (a << 12)) + mm[i2]; break;
        default: a = (a ^ (a >> 33)) + mm[i2]; break;
      }
      mm[i] = y = ind(x) + a + b;
      rsl[i] = b = ind(y >> SIZL) + x;
    }
    bb = b; aa = a; ++blocks;
  }

  void init(uint32_t seed) {
    blocks = 0; aa = bb = cc = 0;
    uint64_t a, b, c, d, e, f, g, h;
    a = b = c = d = e = f = g = h = 0x9e3779b97f4a7c13ULL;
    for (int i = 0; i < SIZ; i++) rsl[i] = 0;
    rsl[0] = seed;
    for (int i = 0; i < 4; i++) mix(a, b, c, d, e, f, g, h);
    for (int pass = 0; pass < 2; pass++) {
      const uint64_t* src = pass == 0 ? rsl : mm;
      for (int i = 0; i < SIZ; i += 8) {
        a += src[i]; b += src[i + 1]; c += src[i + 2]; d += src[i + 3];
        e += src[i + 4]; f += src[i + 5]; g += src[i + 6]; h += src[i + 7];
        mix(a, b, c, d, e, f, g, h);
        mm[i] = a; mm[i + 1] = b; mm[i + 2] = c; mm[i + 3] = d;
        mm[i + 4] = e; mm[i + 5] = f; mm[i + 6] = g; mm[i + 7] = h;
      }
    }
    generate();
    randcnt = WORDS;
  }

  // 32-bit view of rsl[] (little endian), consumed downwards (src/isaac64.h:60-63)
  inline uint32_t next32() {
    if (randcnt == 0) { generate(); randcnt = WORDS; }
    randcnt -= 1;
    uint64_t q = rsl[randcnt >> 1];
    return (randcnt & 1) ? (uint32_t)(q >> 32) : (uint32_t)q;
  }
  long long uses() const { return (long long)WORDS * ((long long)blocks - 1) + (WORDS - (long long)randcnt); }
};

// ---------------------------------------------------------------- Philox4x32-10
static inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// stream layout shared with the product (include/mcx.h: mcx_philox_block)
static inline void philox_block(uint64_t seed, uint32_t mol_id, uint64_t iteration, uint32_t block, uint32_t out[4]) {
  uint32_t ctr[4] = {block, (uint32_t)iteration, (uint32_t)(iteration >> 32), mol_id};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  philox4x32_10(ctr, key, out);
}

// ---------------------------------------------------------------- word sources
// A molecule's evaluation draws 32-bit words from one of: the global ISAAC64 stream
// (sequential/reference semantics), a recorded tape slice, or its own Philox stream.
struct WordSource {
  enum Kind { ISAAC, TAPE, PHILOX } kind = ISAAC;
  Isaac64* isaac = nullptr;
  const uint32_t* tape = nullptr; uint64_t tape_len = 0;
  uint64_t seed = 0; uint32_t mol_id = 0; uint64_t iteration = 0;
  uint32_t buf[4]; uint32_t buf_block = 0xFFFFFFFFu;
  uint32_t used = 0;                   // words drawn so far from this source
  std::vector<uint32_t>* record = nullptr;  // optional: copy of every word drawn
  bool tape_overrun = false;

  inline uint32_t next() {
    uint32_t w;
    if (kind == ISAAC) w = isaac->next32();
    else if (kind == TAPE) {
      if (used < tape_len) w = tape[used]; else { w = 0; tape_overrun = true; }
    } else {
      uint32_t blk = used >> 2;
      if (blk != buf_block) { philox_block(seed, mol_id, iteration, blk, buf); buf_block = blk; }
      w = buf[used & 3];
    }
    used++;
    if (record) record->push_back(w);
    return w;
  }
  inline double dbl() { return 2.3283064365386962890625e-10 * (double)next(); }  // DBL32

  // src/rng.c:173-218
  double gauss() {
    double x, y, sign;
    do {
      uint64_t bits = next();
      sign = (bits & 0x80) ? -1.0 : 1.0;
      uint64_t region = bits & 0x7f;
      uint64_t pos_within_region = bits & 0xffffff00u;
      x = (double)pos_within_region * ZIG_W[region];
      if (pos_within_region < ZIG_K[region]) break;
      if (region != 0) {
        double yB = ZIG_Y[region];
        double yR = ZIG_Y[region - 1] - yB;
        y = yB + yR * dbl();
      } else {
        x = ZIG_R - log1p(-dbl()) * (1.0 / ZIG_R);
        y = exp(-ZIG_R * (x - 0.5 * ZIG_R)) * dbl();
      }
    } while (y >= exp(-0.5 * x * x));
    return sign * x;
  }
};

}  // namespace orc
