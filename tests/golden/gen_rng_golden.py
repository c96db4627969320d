#!/usr/bin/env python3
"""Generate tests/golden/rng_kat.json from the REFERENCE's own RNG (src/rng.c + src/isaac64.c compiled
unmodified into oracle/_ref/librefrng.so by `make -C oracle ref`).  Run in the build container only
(/root/reference must exist); the JSON is committed and checked on every box.

The short vectors are the ones listed in SURVEY.md §A.3; the long ones cross several 512-word ISAAC64
blocks and exercise every Ziggurat branch (strip rejection and the tail)."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

O.build()
R = O.ref_rng_lib()
assert R is not None, "oracle/_ref/librefrng.so missing: run in the build container"

out = {"source": "reference src/rng.c + src/isaac64.c compiled by oracle/Makefile:ref", "seeds": {}}
for seed in (1, 12345, 0, 4294967295):
    e = {}
    r = C.c_void_p(R.ref_rng_new(seed)); e["uint"] = [int(R.ref_rng_uint(r)) for _ in range(6)]; R.ref_rng_free(r)
    r = C.c_void_p(R.ref_rng_new(seed)); e["dbl"] = [float(R.ref_rng_dbl(r)).hex() for _ in range(4)]; R.ref_rng_free(r)
    r = C.c_void_p(R.ref_rng_new(seed)); e["gauss"] = [float(R.ref_rng_gauss(r)).hex() for _ in range(4)]; R.ref_rng_free(r)
    r = C.c_void_p(R.ref_rng_new(seed))
    g = np.zeros(1000000)
    R.ref_rng_fill_gauss(r, C.c_void_p(g.ctypes.data), len(g))
    e["gauss_1e6_sum"] = float(g.sum()).hex()
    e["gauss_1e6_sumsq"] = float((g * g).sum()).hex()
    e["gauss_1e6_words_used"] = int(R.ref_rng_uses(r))
    e["gauss_1e6_sha256"] = hashlib.sha256(g.tobytes()).hexdigest()
    e["gauss_1e6_absmax"] = float(np.abs(g).max()).hex()
    R.ref_rng_free(r)
    r = C.c_void_p(R.ref_rng_new(seed))
    u = np.zeros(5000, np.uint32)
    R.ref_rng_fill_uint(r, C.c_void_p(u.ctypes.data), len(u))
    e["uint_5000_sha256"] = hashlib.sha256(u.tobytes()).hexdigest()
    e["uint_at"] = {str(i): int(u[i]) for i in (511, 512, 513, 1023, 1024, 4999)}
    R.ref_rng_free(r)
    out["seeds"][str(seed)] = e

Y = (C.c_double * 128)(); K = (C.c_ulonglong * 128)(); W = (C.c_double * 128)()
R.ref_zig_tables(Y, K, W)
R.ref_zig_r.restype = C.c_double
out["ziggurat"] = {"R": float(R.ref_zig_r()).hex(),
                   "ytab_sha256": hashlib.sha256(np.array(Y[:]).tobytes()).hexdigest(),
                   "wtab_sha256": hashlib.sha256(np.array(W[:]).tobytes()).hexdigest(),
                   "ktab_sha256": hashlib.sha256(np.array(K[:], dtype=np.uint64).tobytes()).hexdigest(),
                   "ytab_0_1_127": [float(Y[0]).hex(), float(Y[1]).hex(), float(Y[127]).hex()],
                   "ktab_0_1_127": [int(K[0]), int(K[1]), int(K[127])]}
json.dump(out, open(os.path.join(HERE, "rng_kat.json"), "w"), indent=1)
print("wrote rng_kat.json")
