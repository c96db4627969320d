"""CPU tier: PINS the oracle's subpartition walk against MCell4's OWN compiled code.

tests/golden/mcell4_dda_vectors.npz holds the outputs of oracle/_ref/libmcell4ref.so — the reference's
src4/collision_utils_subparts.inl (collect_crossed_subparts + collect_neighboring_subparts, :38-300) compiled
UNMODIFIED against src4/defines.h and libs/glm (oracle/ref_mcell4_shim.cpp, oracle/Makefile: ref) — on the cases of
tests/golden/mcell4_dda_cases.py: ordinary steps, moves across many subpartitions, starts within the interaction radius
of faces / edges / corners, zero displacement components (guard_zero_div), ends exactly on a boundary, and the
partition border.  The oracle must give the same destination, the same ORDERED subpartition list for wall tests and the
same set for molecule tests.  Where the compiled reference is present it is additionally driven live on fresh cases."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import mcell4_dda_cases as dc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "mcell4_dda_vectors.npz"))


def test_subpartition_walk_matches_compiled_mcell4_golden():
    cases = dc.moves()
    assert len(cases) == len(G["dest"])
    multi = with_neighbours = 0
    for i, (gi, pos, disp, fm, fw) in enumerate(cases):
        d, w, m = O.orc_collect(dc.GRIDS[gi], pos, disp, fm, fw)
        nw, nm = int(G["n_walls"][i]), int(G["n_mols"][i])
        assert d == int(G["dest"][i]), i
        assert len(w) == nw and (w == G["walls"][i, :nw]).all(), i          # order matters: walls are tested in it
        assert len(m) == nm and (np.sort(m) == G["mols"][i, :nm]).all(), i   # the reference keeps a set
        assert len(np.unique(m)) == len(m), i
        multi += nw > 1
        with_neighbours += nm > nw
    assert multi > 300 and with_neighbours > 300


def test_subpartition_walk_matches_compiled_mcell4_live():
    R4 = O.ref_mcell4_lib()
    if R4 is None:
        pytest.skip("oracle/_ref/libmcell4ref.so not built here")
    checked = 0
    for seed in (7, 8):
        for gi, pos, disp, fm, fw in dc.moves(seed=seed, n_per_grid=400):
            a = O.ref4_collect(R4, dc.GRIDS[gi], pos, disp, fm, fw)
            b = O.orc_collect(dc.GRIDS[gi], pos, disp, fm, fw)
            assert a[0] == b[0] and (a[1] == b[1]).all() and (a[2] == np.sort(b[2])).all(), (seed, gi, pos, disp)
            checked += 1
    assert checked > 2000
