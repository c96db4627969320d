"""compute-sanitizer target: a small surface-surface model (react_2D_all_neighbors, product placement on freed tiles, weak
mover claims) stepped a few iterations on the device and checked against the oracle's counts.
    compute-sanitizer --tool memcheck python tools/memcheck_surface_surface.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as cm  # noqa: E402
from mcell_b200 import Engine  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

for static_b in (False, True):
    t, mols = cm.surface_reactions(n_a=500, n_b=500, n_e=100, radius_um=0.12, subdivisions=2, seed=3, static_b=static_b)
    e, o = Engine(t), O.Oracle(t)
    e.upload(mols)
    o.upload(mols)
    for it in range(5):
        e.step(1)
        o.step(1, 1)
        assert (np.asarray(e.counts()[0]) == np.asarray(o.counts()[0])).all(), it
        assert (np.asarray(e.counts()[1]) == np.asarray(o.counts()[1])).all(), it
    print("static_b", static_b, "species", e.counts()[0][:6], "rules", e.counts()[1][:5])
    e.close()
print("ok")
