#!/usr/bin/env python3
"""Generate tests/golden/mcell4_test_intersect_vectors.npz: outputs of MCell4's OWN compiled RxnUtils::test_intersect
(src4/rxn_utils.inl:593-626, cut out by line range, compiled unmodified into oracle/_ref/libmcell4leaf.so; the reaction
class behind it is libbng's — absent — and stands in with MCell3's pathway search) on the reaction cases of
mcell3_cases.py.  Run in the build container only; the .npz is committed and checked on every box."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import mcell3_cases as mc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

if __name__ == "__main__":
    O.build()
    R4 = O.ref_mcell4_leaf_lib()
    assert R4 is not None
    R4.ref4_test_intersect.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint, C.c_uint, C.c_void_p]
    cases = mc.rxn_cases()
    out = np.zeros((len(cases), 2))
    for i, (cum, scaling, seed, skip) in enumerate(cases):
        cum = np.ascontiguousarray(cum, dtype=np.float64)
        used = C.c_longlong(0)
        r = R4.ref4_test_intersect(C.c_void_p(cum.ctypes.data), len(cum), scaling, seed, skip, C.byref(used))
        out[i] = [r, used.value]
    np.savez_compressed(os.path.join(HERE, "mcell4_test_intersect_vectors.npz"), out=out)
    print("wrote %d cases, %d reacted" % (len(cases), int((out[:, 0] >= 0).sum())))
