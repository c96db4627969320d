#!/usr/bin/env python3
"""Generate tests/golden/mcell4_grid2uv_random_vectors.npz: outputs of MCell4's OWN compiled GridUtils::grid2uv_random
(src4/grid_utils.inl:256-286, cut out by line range and compiled unmodified into oracle/_ref/libmcell4leaf.so) on the
triangles of mcell3_cases.py, every (few) tiles of each, words from the reference RNG after rng_init(seed) + skip.  Run
in the build container only; the .npz is committed and checked on every box."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import mcell3_cases as mc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def cases():
    """(triangle index, tile index, seed, skip) — deterministic"""
    G3 = np.load(os.path.join(HERE, "mcell3_ref_vectors.npz"))
    out = []
    for ti in range(len(mc.triangles())):
        n_tiles = int(G3["grid_consts"][ti, 7])
        if n_tiles <= 0:
            continue
        step = max(1, n_tiles // 7)
        for k, tile in enumerate(range(0, n_tiles, step)):
            out.append((ti, tile, 11 + ti % 5, (3 * ti + k) % 17))
    return out


if __name__ == "__main__":
    O.build()
    R4 = O.ref_mcell4_leaf_lib()
    assert R4 is not None
    R4.ref4_grid2uv_random.restype = C.c_longlong
    R4.ref4_grid2uv_random.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_void_p]
    tris = mc.triangles()
    cs = cases()
    out = np.zeros((len(cs), 3))
    for i, (ti, tile, seed, skip) in enumerate(cs):
        uv = np.zeros(2)
        used = R4.ref4_grid2uv_random(C.c_void_p(tris[ti].ctypes.data), tile, seed, skip, C.c_void_p(uv.ctypes.data))
        out[i] = [uv[0], uv[1], used]
    np.savez_compressed(os.path.join(HERE, "mcell4_grid2uv_random_vectors.npz"), out=out)
    print("wrote %d cases" % len(cs))
