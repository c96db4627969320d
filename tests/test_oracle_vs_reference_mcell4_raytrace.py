"""ray_trace_vol AS A WHOLE against MCell4's own compiled function ("the
whole-step composition ... remains restated-and-unpinned").

oracle/_ref/libmcell4raytrace.so holds `ray_trace_vol` (src4/diffuse_react_event.cpp:627-780) and
`sort_collisions_by_time` (:341-364) cut out by line range and compiled unmodified over MCell4's own subpartition walk
(collision_utils_subparts.inl, whole), `get_displacement_up_to_partition_boundary`, `collide_mol` + `collide_mol_loop_body`,
`collide_wall` / `jump_away_line` / `get_closest_wall_collision` (collision_utils.inl:48-136, 464-603, 629-914) —
oracle/ref_mcell4_raytrace_shim.cpp.  What is compared is what the composition decides: which subpartitions' walls a
move tests and in which order, that the first hit ends the wall search, when the molecule set is collected again for
the shortened move (wall hit outside the last subpartition), that molecule times are taken on the FULL displacement,
which reactant sets are visited, the REDO restarts (words drawn, displacement changed), the cut at the partition
boundary, and the order the collisions are then evaluated in (time, ties by descending molecule id).  Everything bit
for bit.  The goldens travel to every box; the live comparison runs where oracle/_ref was built."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.dirname(HERE))
import mcell4_raytrace_cases as rc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from test_oracle_vs_reference import ref_words  # noqa: E402

NONE = 0xFFFFFFFF


@pytest.fixture(scope="module")
def scene():
    # liboracle.so up to date (a no-op when it is); the reference side only where it is missing and can be built
    subprocess.run(["make", "-s", "-C", os.path.join(os.path.dirname(HERE), "oracle"), "liboracle.so"], check=True)
    if os.path.isdir("/root/reference/src") and O.ref_mcell4_raytrace_lib() is None:
        O.build()
    t, mols = rc.scene()
    return O.RayTraceScene(t, mols, rc.CAP)


def _same(o, r, i):
    assert o["hit"] == r["hit"] and o["n"] == r["n"], (i, o, r)
    assert np.array_equal(o["type"], r["type"]) and np.array_equal(o["what"], r["what"]), (i, o, r)   # order included
    assert np.array_equal(o["time"], r["time"]) and np.array_equal(o["pos"], r["pos"]), (i, o, r)
    assert np.array_equal(o["disp"], r["disp"]) and o["words"] == r["words"], (i, o, r)


def test_ray_trace_vol_as_a_whole_equals_compiled_mcell4_goldens(scene):
    S = scene
    with np.load(os.path.join(HERE, "golden", "mcell4_raytrace_vectors.npz")) as z:
        g = {k: z[k] for k in z.files}
    cases = rc.moves(S.t, S.mols)
    assert len(cases) == len(g["hit"])
    cfg = S.t.cfg
    lo = S.origin
    hi = S.origin + cfg.partition_edge_length
    rcp = 1.0 / (cfg.partition_edge_length / cfg.num_subparts_per_edge)
    n_sp = int(cfg.num_subparts_per_edge)
    beyond_partition = 0
    for i, (mid, d, use_last, seed, skip) in enumerate(cases):
        last = S.wall_near(mid) if use_last else NONE
        o = S.oracle(mid, d, last, ref_words(seed, skip + 64)[skip:])
        k = int(g["n"][i])
        r = dict(hit=int(g["hit"][i]), n=k, type=g["type"][i, :k], what=g["what"][i, :k], time=g["time"][i, :k],
                 pos=g["pos"][i, :3 * k], disp=g["disp"][i], words=int(g["words"][i]))
        _same(o, r, i)
        end = S.pos[mid] + d
        beyond_partition += bool((end < lo).any() or (end >= hi).any())
        if not r["hit"]:   # RayTraceState::FINISHED moves the molecule (:774-777): the oracle's caller does the same
            after = S.pos[mid] + o["disp"]
            assert np.array_equal(after, g["pos_after"][i]), i
            idx = ((after - lo) * rcp).astype(np.int64)
            # (mod 2^32: a move aimed exactly through a corner of the box can slip out between its walls, in the reference
            # and in the oracle alike; diffuse_vol_molecule then reports the escaped molecule, :590-612)
            assert int(idx[0] + idx[1] * n_sp + idx[2] * n_sp * n_sp) & 0xFFFFFFFF == int(g["subpart_after"][i]), i
    # the cases cover what they are meant to cover
    assert g["hit"].sum() > 500 and (g["type"] == 0).sum() > 3000 and (g["n"] > 1).sum() > 1000 and (g["words"] > 0).sum() > 100
    assert beyond_partition > 50


def test_ray_trace_vol_as_a_whole_equals_compiled_mcell4_live(scene):
    R = O.ref_mcell4_raytrace_lib()
    if R is None:
        pytest.skip("oracle/_ref/libmcell4raytrace.so not built here")
    S = scene
    hits = several = 0
    for i, (mid, d, use_last, seed, skip) in enumerate(rc.moves(S.t, S.mols, n_cases=1200, seed=4242)):
        last = S.wall_near(mid) if use_last else NONE
        r = S.reference(R, mid, d, last, seed, skip)
        o = S.oracle(mid, d, last, ref_words(seed, skip + 64)[skip:])
        _same(o, r, i)
        hits += r["hit"]
        several += r["n"] > 1
    assert hits > 200 and several > 400


@pytest.mark.parametrize("radius,subpart,subdiv", [(0.014, 0.042, 1), (0.004, 0.021, 3)])
def test_ray_trace_vol_other_granularities_live(radius, subpart, subdiv):
    """the same comparison where the interaction radius is the largest the converter admits (mcell4_converter.cpp:121-127: neighbouring subpartitions collected on
    almost every move) and where the subpartitions are small against the move (long walks, 1280 + 12 walls)"""
    R = O.ref_mcell4_raytrace_lib()
    if R is None:
        pytest.skip("oracle/_ref/libmcell4raytrace.so not built here")
    t, mols = rc.scene(seed=57, n=9000, subdivisions=subdiv, interaction_radius=radius, subpartition_dimension=subpart)
    S = O.RayTraceScene(t, mols, 256)
    hits = several = 0
    for i, (mid, d, use_last, seed, skip) in enumerate(rc.moves(t, mols, n_cases=500, seed=777)):
        last = S.wall_near(mid) if use_last else NONE
        r = S.reference(R, mid, d, last, seed, skip)
        assert r["n"] <= 256
        o = S.oracle(mid, d, last, ref_words(seed, skip + 64)[skip:])
        _same(o, r, i)
        hits += r["hit"]
        several += r["n"] > 1
    assert hits > 50 and several > 15, (hits, several)
