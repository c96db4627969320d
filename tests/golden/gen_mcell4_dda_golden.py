#!/usr/bin/env python3
"""Generate tests/golden/mcell4_dda_vectors.npz: outputs of the REFERENCE's own compiled subpartition walk
(oracle/_ref/libmcell4ref.so = src4/collision_utils_subparts.inl built unmodified by `make -C oracle ref`,
oracle/ref_mcell4_shim.cpp) on the cases of mcell4_dda_cases.py.  Build container only; the .npz is committed."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import mcell4_dda_cases as dc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

O.build()
R4 = O.ref_mcell4_lib()
assert R4 is not None
CAP = 512
cases = dc.moves()
dest = np.zeros(len(cases), np.uint32)
walls = np.full((len(cases), 128), 0xFFFFFFFF, np.uint32)
mols = np.full((len(cases), 128), 0xFFFFFFFF, np.uint32)
nw = np.zeros(len(cases), np.uint32)
nm = np.zeros(len(cases), np.uint32)
for i, (gi, pos, disp, fm, fw) in enumerate(cases):
    d, w, m = O.ref4_collect(R4, dc.GRIDS[gi], pos, disp, fm, fw, CAP)
    assert len(w) <= 128 and len(m) <= 128
    dest[i] = d; nw[i] = len(w); nm[i] = len(m); walls[i, :len(w)] = w; mols[i, :len(m)] = m
np.savez_compressed(os.path.join(HERE, "mcell4_dda_vectors.npz"), dest=dest, walls=walls, mols=mols, n_walls=nw, n_mols=nm)
print("cases", len(cases), "max walls", nw.max(), "max mols", nm.max(), "multi-subpart", int((nw > 1).sum()))
