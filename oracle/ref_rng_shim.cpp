// oracle/_ref build only: thin extern "C" shim around the REFERENCE's own RNG
// (/root/reference/src/rng.{h,c}, isaac64.{h,c}), compiled where the sources lie.
// TEST INFRASTRUCTURE. Nothing from the reference is copied into the repo; this
// file only #includes the reference headers at build time (see Makefile: ref).
#include <cstdlib>
#include <cstring>
#include "rng.h"   // reference: src/rng.h (pulls isaac64.h/.c and rng.c)

extern "C" {
void* ref_rng_new(unsigned seed) {
  rng_state* r = (rng_state*)malloc(sizeof(rng_state));
  rng_init(r, seed);
  return r;
}
void ref_rng_free(void* r) { free(r); }
unsigned ref_rng_uint(void* r) { return rng_uint((rng_state*)r); }
double ref_rng_dbl(void* r) { return rng_dbl((rng_state*)r); }
double ref_rng_gauss(void* r) { return rng_gauss((rng_state*)r); }
long long ref_rng_uses(void* r) { return rng_uses((rng_state*)r); }
void ref_rng_fill_uint(void* r, unsigned* out, long n) { for (long i = 0; i < n; i++) out[i] = rng_uint((rng_state*)r); }
void ref_rng_fill_gauss(void* r, double* out, long n) { for (long i = 0; i < n; i++) out[i] = rng_gauss((rng_state*)r); }
// Ziggurat tables exactly as the reference compiled them (src/rng.c:58-167)
void ref_zig_tables(double* ytab, unsigned long long* ktab, double* wtab) {
  for (int i = 0; i < 128; i++) { ytab[i] = YTAB[i]; ktab[i] = KTAB[i]; wtab[i] = WTAB[i]; }
}
double ref_zig_r() { return SCALE_FACTOR; }
}
