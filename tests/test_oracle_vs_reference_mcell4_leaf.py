"""CPU tier: PINS the oracle's volume-path leaf arithmetic against MCell4's OWN compiled code.

tests/golden/mcell4_leaf_vectors.npz holds the outputs of oracle/_ref/libmcell4leaf.so — the reference's
CollisionUtils::collide_mol, jump_away_line, collide_wall, get_closest_wall_collision, reflect_from_wall
(src4/collision_utils.inl:464-515, 568-603, 629-812, 819-914, 1711-1747) and Wall::initialize_wall_constants
(src4/wall.cpp:281-342), cut out of the reference files by line range at build time and compiled unmodified
(oracle/ref_mcell4_leaf_shim.cpp, oracle/Makefile: ref) — on the cases of tests/golden/mcell3_cases.py and
mcell4_leaf_cases.py.  Three statements: the oracle reproduces MCell4 bit for bit; MCell4 equals its MCell3 original
on the shared cases (so the MCell3 pins of test_oracle_vs_reference.py speak for MCell4 too); and, where the compiled
code is present, the same on fresh random cases."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import mcell3_cases as mc  # noqa: E402
import mcell4_leaf_cases as lc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from test_oracle_vs_reference import ref_words, same, vp  # noqa: E402

G4 = np.load(os.path.join(HERE, "golden", "mcell4_leaf_vectors.npz"))
G3 = np.load(os.path.join(HERE, "golden", "mcell3_ref_vectors.npz"))


def test_mcell4_leaf_functions_equal_their_mcell3_originals_on_the_shared_cases():
    assert np.array_equal(G4["wall_constants"], G3["wall_constants"])
    assert np.array_equal(G4["ray_out"], G3["ray_out"], equal_nan=True)
    assert np.array_equal(G4["mol_out"], G3["mol_out"], equal_nan=True)
    assert set(G4["ray_out"][:, 0]) == {-1.0, 0.0, 1.0, 2.0}


def test_wall_constants_collide_wall_collide_mol_match_compiled_mcell4():
    U = O.unit_lib()
    tris = mc.triangles()
    out = np.zeros(16)
    for i in range(len(tris)):
        U.orc_unit_wall_constants(vp(tris[i]), vp(out))
        assert (out == G4["wall_constants"][i]).all(), i
    words = ref_words(77, 64)
    ref = G4["ray_out"]
    for i, (ti, p, m) in enumerate(mc.wall_rays(tris)):
        p = np.ascontiguousarray(p, dtype=np.float64); m = np.ascontiguousarray(m, dtype=np.float64).copy()
        t = C.c_double(0); hit = np.zeros(3); used = C.c_longlong(0)
        tape = words[i % 13:]
        code = U.orc_unit_collide_wall(vp(p), vp(m), vp(tris[ti]), vp(tape), len(tape), C.byref(t), vp(hit), C.byref(used))
        assert code == int(ref[i, 0]) and (m == ref[i, 5:8]).all() and used.value == int(ref[i, 8]), i
        if code in (1, 2):
            assert t.value == ref[i, 1] and (hit == ref[i, 2:5]).all(), i
    p, mv, tg, R = mc.mol_pairs()
    ref = G4["mol_out"]
    for i in range(len(p)):
        t = C.c_double(0); hit = np.zeros(3)
        code = U.orc_unit_collide_mol(vp(p[i]), vp(mv[i]), vp(tg[i]), R, C.byref(t), vp(hit))
        assert (code == 3) == bool(ref[i, 0]), i
        if code == 3:
            assert same(t.value, ref[i, 1]) and same(hit, ref[i, 2:5]), i


def _check_mesh_cases(cases, ref_rows):
    meshes = lc.meshes()
    n_hit = n_redo = 0
    for i, (mi, pos, move, last, seed, skip) in enumerate(cases):
        words = ref_words(seed, skip + 64)[skip:]
        got = np.array(O.orc_closest_wall_and_reflect(meshes[mi], pos, move, last, words))
        ref = np.asarray(ref_rows[i])
        # found, wall, side, t, hit, position and displacement after the reflection, remaining time, the (possibly
        # perturbed) displacement, words drawn: bit for bit.  ray_polygon_tests (last column): the oracle counts one test per
        # collide_wall call like the reference does
        assert np.array_equal(got, ref, equal_nan=True), (i, mi, got, ref)
        n_hit += ref[0] > 0
        n_redo += ref[17] > 0
    return n_hit, n_redo


def test_closest_wall_collision_and_reflection_match_compiled_mcell4_golden():
    n_hit, n_redo = _check_mesh_cases(lc.mesh_rays(), G4["mesh_out"])
    assert n_hit > 1000 and n_redo > 500   # the REDO / jump_away_line restarts are exercised


def test_leaf_functions_match_compiled_mcell4_live():
    R4 = O.ref_mcell4_leaf_lib()
    if R4 is None:
        pytest.skip("oracle/_ref/libmcell4leaf.so not built here")
    cases = lc.mesh_rays(seed=909, per_mesh=400)
    meshes = lc.meshes()
    rows = [O.ref4_closest_wall_and_reflect(R4, meshes[mi], pos, move, last, seed, skip) for mi, pos, move, last, seed, skip in cases]
    n_hit, n_redo = _check_mesh_cases(cases, rows)
    assert n_hit > 500 and n_redo > 200
    U = O.unit_lib()
    rng = np.random.default_rng(4)
    tris = mc.triangles(n=40, seed=77)
    words = ref_words(5, 32)
    for k in range(6000):
        t9 = tris[int(rng.integers(0, len(tris)))]
        a, b = rng.uniform(-0.2, 1.2, 2)
        target = t9[0:3] + a * (t9[3:6] - t9[0:3]) + b * (t9[6:9] - t9[0:3])
        d = rng.normal(size=3) * rng.uniform(0.01, 4)
        p = np.ascontiguousarray(target - d * rng.uniform(0, 1.4)); m1 = d.copy(); m2 = d.copy()
        t1, t2 = C.c_double(0), C.c_double(0); h1, h2 = np.zeros(3), np.zeros(3)
        u1, u2 = C.c_longlong(0), C.c_longlong(0)
        c1 = R4.ref4_collide_wall(vp(p), vp(m1), vp(t9), 5, 0, C.byref(t1), vp(h1), C.byref(u1))
        c2 = U.orc_unit_collide_wall(vp(p), vp(m2), vp(t9), vp(words), len(words), C.byref(t2), vp(h2), C.byref(u2))
        assert c1 == c2 and (m1 == m2).all() and u1.value == u2.value
        if c1 in (1, 2):
            assert t1.value == t2.value and (h1 == h2).all()


def test_pick_surf_displacement_and_test_bimolecular_match_compiled_mcell4():
    """pick_surf_displacement is MCell4's own rounding (a * (normal_factor * scale), diffusion_utils.inl:93; MCell3 computes
    (a * normal_factor) * scale), so only this pin covers it.  test_bimolecular keeps MCell4's float max_p
    (rxn_utils.inl:369); the pathway search behind it is libbng's (absent): the stand-in uses MCell3's."""
    L = O.lib()
    L.orc_unit_pick_surf_displacement.restype = C.c_longlong
    L.orc_unit_pick_surf_displacement.argtypes = [C.c_double, C.c_void_p, C.c_uint64, C.c_void_p]
    ref = G4["surf_out"]
    several = 0
    for i, (scale, seed, skip) in enumerate(lc.surf_displacement_cases()):
        tape = ref_words(seed, skip + 16)[skip:]
        o2 = np.zeros(2)
        used = L.orc_unit_pick_surf_displacement(scale, vp(tape), len(tape), vp(o2))
        assert o2[0] == ref[i, 0] and o2[1] == ref[i, 1] and used == int(ref[i, 2]), i
        several += used > 1
    assert several > 300   # the rejection loop of the polar method is exercised
    U = O.unit_lib()
    ref = G4["bimol_out"]
    reacted = 0
    for i, (cum, scaling, seed, skip) in enumerate(mc.rxn_cases()):
        cum = np.ascontiguousarray(cum, dtype=np.float64)
        tape = ref_words(seed, skip + 8)[skip:]
        used = C.c_longlong(0)
        r = U.orc_unit_test_bimolecular(vp(cum), len(cum), scaling, vp(tape), len(tape), C.byref(used))
        assert r == int(ref[i, 0]) and used.value == int(ref[i, 1]) == 1, (i, r, ref[i])
        reacted += r >= 0
    assert reacted > 10


def test_surface_grid_and_edge_point_functions_of_mcell4_equal_mcell3_and_the_oracle():
    """Grid::initialize, xyz2grid_tile_index, grid2uv, uv2grid_tile_index (src4/wall.cpp:38-74, grid_utils.inl:48-253) and
    find_edge_point (geometry_utils.inl:222-291) compiled from MCell4's own files give the outputs of their MCell3
    originals on the cases of the MCell3 goldens — which test_oracle_vs_reference.py / _surface.py hold the oracle to —
    and the oracle reproduces them directly."""
    import mcell3_surface_cases as sc
    G3s = np.load(os.path.join(HERE, "golden", "mcell3_surface_vectors.npz"))
    assert np.array_equal(G4["grid_consts"], G3["grid_consts"])
    assert np.array_equal(G4["grid_idx"], G3["grid_idx"]) and np.array_equal(G4["grid_uv"], G3["grid_uv"])
    assert np.array_equal(G4["fep_code"], G3s["fep_code"]) and np.array_equal(G4["fep_pt"], G3s["fep_pt"])
    assert np.array_equal(G4["uv2grid"], G3s["uv2grid"])
    U = O.unit_lib()
    tris = mc.triangles()
    out = np.zeros(8)
    for i in range(len(tris)):
        if G4["grid_consts"][i, 7] > 0:
            U.orc_unit_grid_constants(vp(tris[i]), vp(out))
            assert (out == G4["grid_consts"][i]).all(), i
    for k, (ti, pt) in enumerate(mc.grid_points(tris)):
        pt = np.ascontiguousarray(pt, np.float64)
        assert U.orc_unit_xyz2grid(vp(tris[ti]), vp(pt)) == int(G4["grid_idx"][k]), k
    L = O.lib()
    stris = sc.triangles()
    for i, (ti, loc, disp) in enumerate(sc.edge_moves(stris)):
        code, pt = O.find_edge_point(L.orc_unit_find_edge_point, stris[ti], loc, disp)
        assert code == int(G4["fep_code"][i]) and (pt == G4["fep_pt"][i]).all(), i
    assert set(G4["fep_code"]) == {-2, -1, 0, 1, 2}


def test_exact_disk_of_mcell4_equals_mcell3_and_the_oracle():
    """ExactDiskUtils::exact_disk with everything it calls (src4/exact_disk_utils.inl:54-1145), compiled from MCell4's own
    file: same fraction of the interaction disk as the MCell3 original on the 3000 cases of the MCell3 golden (next to box
    faces / edges / corners, triangle soups, blocked targets) and as the oracle; live on fresh cases where present."""
    assert np.array_equal(G4["disk_out"], G3["disk_out"])
    U = O.unit_lib()
    cases, R = mc.disk_cases()
    for i, (loc, mv, tg, walls) in enumerate(cases):
        assert same(U.orc_unit_exact_disk(vp(loc), vp(mv), R, vp(tg), len(walls), vp(walls)), G4["disk_out"][i]), i
    assert ((G4["disk_out"] > 0) & (G4["disk_out"] < 1)).sum() > 500 and (G4["disk_out"] < 0).sum() > 50
    R4 = O.ref_mcell4_leaf_lib()
    if R4 is None:
        return
    R4.ref4_exact_disk.restype = C.c_double
    R4.ref4_exact_disk.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
    cases, R = mc.disk_cases(n=3000, seed=777)
    for loc, mv, tg, walls in cases:
        a = U.orc_unit_exact_disk(vp(loc), vp(mv), R, vp(tg), len(walls), vp(walls))
        b = R4.ref4_exact_disk(vp(loc), vp(mv), R, vp(tg), len(walls), vp(walls))
        assert same(a, b), (a, b)


def _csr_cases():
    """(name, origin, partition edge [lu], subpartitions per edge, R [lu], expanded list, vertices [lu], triangles)"""
    from mcell_b200.model import create_box, create_icosphere
    R = 0.5641895835477563
    out = []
    v, f = create_box(1.0)
    out.append(("box 1 um, default subpartitions", -500.0, 1000.0, 20, R, 1, v / 0.01, f))
    out.append(("box 1 um, 0.1 um subpartitions, plain list", -500.0, 1000.0, 100, R, 0, v / 0.01, f))
    v, f = create_icosphere(0.5, 4)
    out.append(("icosphere 1280 walls, default", -500.0, 1000.0, 20, R, 1, v / 0.01, f))
    out.append(("icosphere 1280 walls, 0.125 um", -500.0, 1000.0, 80, R, 1, v / 0.01 + np.array([7.3, -11.9, 3.1]), f))
    v, f = create_icosphere(0.37, 3)
    out.append(("icosphere 320 walls on subpartition planes", -500.0, 1000.0, 40, R, 1, v / 0.01 + np.array([25.0, 25.0, 0.0]), f))
    return out


def test_product_wall_lists_match_mcell4_wall_distribution():
    """The subpartition wall lists libmcx walks (host code csrc/mcx_geom.cpp: bin_walls, exported as mcx_walls_per_subpart)
    and the oracle's walls_per_subpart against Partition::finalize_walls' own distribution — wall_subparts_collision_test +
    wall_in_box compiled from MCell4's files: the same CSR, entry for entry."""
    R4 = O.ref_mcell4_leaf_lib()
    if R4 is None:
        pytest.skip("oracle/_ref/libmcell4leaf.so not built here")
    import mcell_b200
    L = mcell_b200.load_library()
    L.mcx_walls_per_subpart.restype = C.c_uint64
    R4.ref4_walls_per_subpart.restype = C.c_ulonglong
    total = 0
    for name, origin, edge, n, R, expanded, v, f in _csr_cases():
        v = np.ascontiguousarray(v, np.float64); f = np.ascontiguousarray(f, np.uint32)
        o3 = np.array([origin] * 3)
        ns = n ** 3
        cap = 64 * len(f) + 400000
        sa, la = np.zeros(ns + 1, np.uint32), np.zeros(cap, np.uint32)
        sb, lb = np.zeros(ns + 1, np.uint32), np.zeros(cap, np.uint32)
        na = R4.ref4_walls_per_subpart(vp(o3), C.c_double(edge), C.c_uint(n), C.c_double(R), C.c_int(expanded), vp(v), C.c_uint(len(v)),
                                       vp(f), C.c_uint(len(f)), vp(sa), vp(la), C.c_ulonglong(cap))
        nb = L.mcx_walls_per_subpart(vp(o3), C.c_double(edge), C.c_uint32(n), C.c_double(R), C.c_uint32(expanded), vp(v),
                                     C.c_uint64(len(v)), vp(f), C.c_uint64(len(f)), vp(sb), vp(lb), C.c_uint64(cap))
        assert na == nb and na <= cap, (name, na, nb)
        assert np.array_equal(sa, sb) and np.array_equal(la[:na], lb[:nb]), name
        assert na > len(f)   # walls straddle subpartitions
        total += na
    assert total > 5000


def test_oracle_wall_lists_match_mcell4_wall_distribution():
    R4 = O.ref_mcell4_leaf_lib()
    if R4 is None:
        pytest.skip("oracle/_ref/libmcell4leaf.so not built here")
    import common as cm
    R4.ref4_walls_per_subpart.restype = C.c_ulonglong
    for t in (cm.sphere_classes(n=10)[0], cm.free_diffusion_box(n=10)[0]):
        o = O.Oracle(t)
        v = np.ascontiguousarray(t.vertices, np.float64); f = np.ascontiguousarray(t.tri, np.uint32)
        n = int(t.cfg.num_subparts_per_edge); ns = n ** 3
        o3 = np.array(list(t.cfg.origin), np.float64)
        cap = 64 * len(f) + 400000
        sa, la = np.zeros(ns + 1, np.uint32), np.zeros(cap, np.uint32)
        na = R4.ref4_walls_per_subpart(vp(o3), C.c_double(t.cfg.partition_edge_length), C.c_uint(n), C.c_double(t.cfg.rxn_radius_3d),
                                       C.c_int(int(t.cfg.use_expanded_list)), vp(v), C.c_uint(len(v)), vp(f), C.c_uint(len(f)),
                                       vp(sa), vp(la), C.c_ulonglong(cap))
        assert 0 < na <= cap
        for s in np.flatnonzero(np.diff(sa.astype(np.int64)) > 0):
            assert np.array_equal(o.subpart_walls(int(s)), la[sa[s]:sa[s + 1]]), s


def test_grid2uv_random_matches_compiled_mcell4():
    """GridUtils::grid2uv_random (src4/grid_utils.inl:256-286), the position of a released surface molecule inside its tile
    (mcx_release_surface_molecules, place_single_molecule_onto_grid with config.randomize_smol_pos): the oracle's
    restatement against MCell4's own compiled function, golden vectors on every box and live where oracle/_ref is."""
    import gen_mcell4_grid2uv_random_golden as gg
    L = O.lib()
    L.orc_unit_grid2uv_random.restype = C.c_longlong
    L.orc_unit_grid2uv_random.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    ref = np.load(os.path.join(HERE, "golden", "mcell4_grid2uv_random_vectors.npz"))["out"]
    tris = mc.triangles()
    cases = gg.cases()
    assert len(cases) == len(ref) > 2000
    R4 = O.ref_mcell4_leaf_lib()
    if R4 is not None:
        R4.ref4_grid2uv_random.restype = C.c_longlong
        R4.ref4_grid2uv_random.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_void_p]
    for i, (ti, tile, seed, skip) in enumerate(cases):
        tape = ref_words(seed, skip + 4)[skip:]
        uv = np.zeros(2)
        used = L.orc_unit_grid2uv_random(vp(tris[ti]), tile, vp(tape), len(tape), vp(uv))
        assert uv[0] == ref[i, 0] and uv[1] == ref[i, 1] and used == int(ref[i, 2]) == 2, i
        if R4 is not None and i % 9 == 0:   # live, other seeds
            uv_r, uv_o = np.zeros(2), np.zeros(2)
            tape2 = ref_words(seed + 100, skip + 4)[skip:]
            assert R4.ref4_grid2uv_random(vp(tris[ti]), tile, seed + 100, skip, vp(uv_r)) == 2
            L.orc_unit_grid2uv_random(vp(tris[ti]), tile, vp(tape2), len(tape2), vp(uv_o))
            assert (uv_r == uv_o).all(), i


def test_test_intersect_matches_compiled_mcell4():
    """RxnUtils::test_intersect (src4/rxn_utils.inl:593-626), the test of a finite-rate reaction with a surface class
    (SURVEY a18): pathway and number of words drawn against MCell4's own compiled function — reactions that happen
    (two draws), that do not (one), the GIGANTIC rate of an absorptive class."""
    L = O.lib()
    L.orc_unit_test_intersect.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_uint64, C.c_void_p]
    ref = np.load(os.path.join(HERE, "golden", "mcell4_test_intersect_vectors.npz"))["out"]
    R4 = O.ref_mcell4_leaf_lib()
    if R4 is not None:
        R4.ref4_test_intersect.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint, C.c_uint, C.c_void_p]
    reacted = missed = 0
    for i, (cum, scaling, seed, skip) in enumerate(mc.rxn_cases()):
        cum = np.ascontiguousarray(cum, dtype=np.float64)
        tape = ref_words(seed, skip + 8)[skip:]
        used = C.c_longlong(0)
        r = L.orc_unit_test_intersect(vp(cum), len(cum), scaling, vp(tape), len(tape), C.byref(used))
        assert r == int(ref[i, 0]) and used.value == int(ref[i, 1]), (i, r, used.value, ref[i])
        assert used.value == (2 if r >= 0 else 1)
        reacted += r >= 0
        missed += r < 0
        if R4 is not None and i % 5 == 0:   # live, other seeds
            u1, u2 = C.c_longlong(0), C.c_longlong(0)
            tape2 = ref_words(seed + 7, skip + 8)[skip:]
            a = R4.ref4_test_intersect(vp(cum), len(cum), scaling, seed + 7, skip, C.byref(u1))
            b = L.orc_unit_test_intersect(vp(cum), len(cum), scaling, vp(tape2), len(tape2), C.byref(u2))
            assert a == b and u1.value == u2.value, i
    assert reacted > 50 and missed > 50
