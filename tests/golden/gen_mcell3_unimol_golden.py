#!/usr/bin/env python3
"""Generate tests/golden/mcell3_unimol_vectors.npz: timeof_unimolecular and which_unimolecular of the reference's
src/react_cond.c (oracle/_ref/libmcell3ref.so) on deterministic (rate, seed, skip) cases.  Build container only."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_py as O  # noqa: E402

O.build()
R3 = O.ref_mcell3_lib()
R3.ref3_timeof_unimolecular.restype = C.c_double
R3.ref3_timeof_unimolecular.argtypes = [C.c_double, C.c_uint, C.c_uint]


def cases():
    rng = np.random.default_rng(21)
    k = np.concatenate([[0.0, -1.0, 1e-300, 1e300], 10.0 ** rng.uniform(-9, 3, 196)])
    return [(float(k[i]), 50 + i % 7, i % 11) for i in range(len(k))]


def pathway_cases():
    rng = np.random.default_rng(22)
    out = []
    for i in range(150):
        n = int(rng.integers(1, 7))
        cum = np.cumsum(10.0 ** rng.uniform(-6, 0, n))
        out.append((np.ascontiguousarray(cum), 60 + i % 5, i % 9))
    return out


if __name__ == "__main__":
    t = np.array([R3.ref3_timeof_unimolecular(k, s, sk) for k, s, sk in cases()])
    pw = []
    for cum, s, sk in pathway_cases():
        used = C.c_longlong(0)
        r = R3.ref3_which_unimolecular(C.c_void_p(cum.ctypes.data), len(cum), s, sk, C.byref(used))
        pw.append((r, used.value))
    np.savez_compressed(os.path.join(HERE, "mcell3_unimol_vectors.npz"), lifetime=t, pathway=np.array(pw, np.int64))
    print("lifetimes", len(t), "FOREVER", int((t >= 1e20).sum()), "pathways", len(pw))
