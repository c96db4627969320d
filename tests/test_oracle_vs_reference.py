"""CPU tier: PINS the oracle's restatement of the hot path's arithmetic against the REFERENCE'S OWN compiled code.

tests/golden/mcell3_ref_vectors.npz holds outputs of oracle/_ref/libmcell3ref.so (the reference's
src/wall_util.c, react_cond.c, react_util.c, util.c compiled unmodified, see oracle/Makefile:ref and
tests/golden/gen_mcell3_golden.py) on the deterministic cases of tests/golden/mcell3_cases.py.  The oracle must
reproduce them BIT FOR BIT (same -O3 -march=core2 -ffp-contract=off flags).  Where the compiled reference is
present (build container, and the GPU box via the travelling .so) it is additionally driven live on fresh
random cases."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import mcell3_cases as mc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "mcell3_ref_vectors.npz"))


def vp(a):
    return C.c_void_p(a.ctypes.data)


def same(a, b):
    """bit-for-bit equality that also accepts NaN == NaN (zero-length moves divide 0/0 on both sides)."""
    return np.array_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), equal_nan=True)


def ref_words(seed, n):
    """Words the reference RNG delivers after rng_init(seed) — from the oracle's ISAAC64 restatement, which
    test_oracle_rng.py pins to the reference's rng.c."""
    L = O.lib()
    r = C.c_void_p(L.orc_rng_new(C.c_uint32(seed)))
    w = np.zeros(n, np.uint32)
    L.orc_rng_fill_uint(r, vp(w), n)
    L.orc_rng_free(r)
    return w


def test_wall_constants_bit_exact():
    U = O.unit_lib()
    tris = mc.triangles()
    out = np.zeros(16)
    for i in range(len(tris)):
        U.orc_unit_wall_constants(vp(tris[i]), vp(out))
        assert (out == G["wall_constants"][i]).all(), i


def test_collide_wall_bit_exact_including_redo_paths():
    U = O.unit_lib()
    tris = mc.triangles()
    rays = mc.wall_rays(tris)
    words = ref_words(77, 64)
    ref = G["ray_out"]
    assert len(rays) == len(ref)
    seen = set()
    for i, (ti, p, m) in enumerate(rays):
        p = np.ascontiguousarray(p, dtype=np.float64); m = np.ascontiguousarray(m, dtype=np.float64).copy()
        t = C.c_double(0); hit = np.zeros(3); used = C.c_longlong(0)
        tape = words[i % 13:]
        code = U.orc_unit_collide_wall(vp(p), vp(m), vp(tris[ti]), vp(tape), len(tape), C.byref(t), vp(hit), C.byref(used))
        assert code == int(ref[i, 0]), (i, code, ref[i, 0])
        seen.add(code)
        if code in (1, 2):
            assert t.value == ref[i, 1] and (hit == ref[i, 2:5]).all(), i
        assert (m == ref[i, 5:8]).all(), i           # move is perturbed on REDO, untouched otherwise
        assert used.value == int(ref[i, 8]), i
    assert seen == {-1, 0, 1, 2}


def test_collide_mol_bit_exact():
    U = O.unit_lib()
    p, mv, tg, R = mc.mol_pairs()
    ref = G["mol_out"]
    for i in range(len(p)):
        t = C.c_double(0); hit = np.zeros(3)
        code = U.orc_unit_collide_mol(vp(p[i]), vp(mv[i]), vp(tg[i]), R, C.byref(t), vp(hit))
        assert (code == 3) == bool(ref[i, 0]), i
        if code == 3:
            assert same(t.value, ref[i, 1]) and same(hit, ref[i, 2:5]), i
    assert 500 < ref[:, 0].sum() < 3500


def test_wall_in_box_exact():
    U = O.unit_lib()
    tris = mc.triangles()
    bx = mc.boxes(tris)
    ref = G["box_out"]
    for i, (ti, lo, hi) in enumerate(bx):
        lo = np.ascontiguousarray(lo, dtype=np.float64); hi = np.ascontiguousarray(hi, dtype=np.float64)
        assert U.orc_unit_wall_in_box(vp(tris[ti]), vp(lo), vp(hi)) == int(ref[i]), i


def test_test_bimolecular_and_intersect_draws():
    """Pathway choice and number of random words.  The absorptive surface class (rate GIGANTIC) must consume
    exactly two words and always react — the oracle/product hard-code that (diffuse_vol_molecule wall branch)."""
    U = O.unit_lib()
    ref = G["rxn_out"]
    cases = mc.rxn_cases()
    for i, (cum, scaling, seed, skip) in enumerate(cases):
        cum = np.ascontiguousarray(cum, dtype=np.float64)
        with np.errstate(over="ignore"):
            f32 = float(np.float32(cum[-1]))
        if f32 != cum[-1] and abs(cum[-1] - scaling) < 1e-6:
            continue  # MCell4 compares a float-truncated max_p here (rxn_utils.inl:369); measure-zero case
        tape = ref_words(seed, skip + 8)[skip:]
        used = C.c_longlong(0)
        r = U.orc_unit_test_bimolecular(vp(cum), len(cum), scaling, vp(tape), len(tape), C.byref(used))
        assert r == int(ref[i, 0]), (i, r, ref[i])
        assert used.value == int(ref[i, 1]) == 1
    k = len(cases) - 2  # absorptive
    assert ref[k, 2] == 0 and ref[k, 3] == 2


def test_unimolecular_lifetime_and_pathway_draws():
    """timeof_unimolecular / which_unimolecular (src/react_cond.c == rxn_utils.inl:721-736, 774-783): the lifetime of a
    newborn molecule bit for bit (incl. FOREVER for k <= 0), the pathway of a firing class and its word count."""
    import gen_mcell3_unimol_golden as gu
    Gu = np.load(os.path.join(HERE, "golden", "mcell3_unimol_vectors.npz"))
    L = O.lib()
    L.orc_unit_time_of_unimol.restype = C.c_double
    L.orc_unit_time_of_unimol.argtypes = [C.c_double, C.c_void_p, C.c_uint64]
    for i, (k, seed, skip) in enumerate(gu.cases()):
        tape = ref_words(seed, skip + 4)[skip:]
        assert L.orc_unit_time_of_unimol(k, vp(tape), len(tape)) == Gu["lifetime"][i], (i, k)
    assert (Gu["lifetime"][:2] == 1e20).all() and (Gu["lifetime"][4:] < 1e20).all()
    zero = np.zeros(4, np.uint32)          # a zero word: p == 0 is not distinguishable from 0 -> FOREVER
    assert L.orc_unit_time_of_unimol(5.0, vp(zero), 4) == 1e20
    for i, (cum, seed, skip) in enumerate(gu.pathway_cases()):
        tape = ref_words(seed, skip + 4)[skip:]
        used = C.c_longlong(0)
        r = L.orc_unit_which_unimolecular(vp(cum), len(cum), vp(tape), len(tape), C.byref(used))
        assert (r, used.value) == tuple(int(x) for x in Gu["pathway"][i]), i
    assert Gu["pathway"][:, 0].max() >= 4


def test_distinguishable_and_pb_factor():
    U = O.unit_lib()
    for (a, b, e), want in zip(G["dist_cases"], G["dist_out"]):
        assert U.orc_unit_distinguishable(a, b, e) == int(want)
    # table builder (mcell_b200/model.py) against the reference's compute_pb_factor
    from mcell_b200.model import Model, Config
    m = Model(Config())
    m.add_species("A", 1e-6); m.add_species("B", 1e-6); m.add_species("H", 0.5e-6); m.add_species("T", 1e-6, target_only=True)
    m.add_reaction_rule(["A", "B"], [], 1.0)
    m.add_reaction_rule(["A", "H"], [], 1.0)
    m.add_reaction_rule(["T", "B"], [], 1.0)
    t = m.build(max_molecules=4)
    got = [t.pathways[t.classes[c].first_pathway].cum_prob for c in range(3)]
    want = G["pb_factor"]
    assert got[0] == pytest.approx(want[0], rel=1e-15)
    assert got[1] == pytest.approx(want[1], rel=1e-15)
    assert got[2] == pytest.approx(want[2], rel=1e-15)


# ------------------------------------------------------------------ live against the compiled reference
R3 = O.ref_mcell3_lib()
def test_exact_disk_bit_exact():
    """exact_disk (src4/exact_disk_utils.inl:840-1145 == src/diffuse.c:1365): occlusion factor of the interaction disk
    next to box faces / edges / corners and inside random triangle soups, incl. blocked targets."""
    U = O.unit_lib()
    cases, R = mc.disk_cases()
    ref = G["disk_out"]
    assert len(cases) == len(ref)
    for i, (loc, mv, tg, walls) in enumerate(cases):
        got = U.orc_unit_exact_disk(vp(loc), vp(mv), R, vp(tg), len(walls), vp(walls))
        assert same(got, ref[i]), (i, got, ref[i])
    assert (ref < 0).sum() > 300 and ((ref > 0) & (ref < 1)).sum() > 800 and (ref == 1).sum() > 300


def test_surface_grid_bit_exact():
    """Grid::initialize, xyz2grid_tile_index, grid2uv, uv2xyz (src4/wall.cpp:38-74, grid_utils.inl:48-118,233-253 ==
    src/grid_util.c) on random and box triangles, incl. points on the three vertices."""
    U = O.unit_lib()
    tris = mc.triangles()
    out = np.zeros(8)
    n_checked = 0
    for i in range(len(tris)):
        if G["grid_consts"][i, 7] == 0:
            continue                                             # degenerate triangle: no grid
        U.orc_unit_grid_constants(vp(tris[i]), vp(out))
        assert (out == G["grid_consts"][i]).all(), i
        n_checked += 1
    assert n_checked > 300
    gp = mc.grid_points(tris)
    uv, xyz = np.zeros(2), np.zeros(3)
    for k, (ti, pt) in enumerate(gp):
        idx = U.orc_unit_xyz2grid(vp(tris[ti]), vp(pt))
        assert idx == G["grid_idx"][k], (k, ti)
        U.orc_unit_grid2uv(vp(tris[ti]), idx, vp(uv))
        assert (uv == G["grid_uv"][k]).all(), k
        U.orc_unit_uv2xyz(vp(tris[ti]), vp(uv), vp(xyz))
        assert (xyz == G["grid_xyz"][k]).all(), k


def test_libmcx_host_grid_helpers_match_reference():
    """The product's host helpers for surface placement (mcx_grid_num_tiles / mcx_grid2uv / mcx_xyz2grid)."""
    from mcell_b200 import engine
    L = engine.load_library()
    L.mcx_grid_num_tiles.argtypes = [C.c_void_p]; L.mcx_grid_num_tiles.restype = C.c_uint32
    L.mcx_grid2uv.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]; L.mcx_grid2uv.restype = None
    L.mcx_xyz2grid.argtypes = [C.c_void_p, C.c_void_p]; L.mcx_xyz2grid.restype = C.c_uint32
    tris = mc.triangles()
    gp = mc.grid_points(tris)
    uv = np.zeros(2)
    for k, (ti, pt) in enumerate(gp):
        assert L.mcx_grid_num_tiles(vp(tris[ti])) == int(G["grid_consts"][ti, 7])
        idx = L.mcx_xyz2grid(vp(tris[ti]), vp(pt))
        assert idx == G["grid_idx"][k], k
        L.mcx_grid2uv(vp(tris[ti]), idx, vp(uv))
        assert (uv == G["grid_uv"][k]).all(), k


needs_ref = pytest.mark.skipif(R3 is None, reason="oracle/_ref/libmcell3ref.so not present on this box")


@needs_ref
def test_golden_file_is_current():
    """The committed vectors are what the compiled reference produces now (guards a stale fixture)."""
    tris = mc.triangles()
    out = np.zeros(16)
    for i in (0, 7, 150, len(tris) - 1):
        R3.ref3_init_tri_wall(vp(tris[i]), vp(out))
        assert (out == G["wall_constants"][i]).all()


@needs_ref
def test_live_random_rays_against_compiled_reference():
    U = O.unit_lib()
    rng = np.random.default_rng(2024)
    tris = mc.triangles(n=60, seed=99)
    words = ref_words(5, 32)
    n_hit = 0
    for k in range(20000):
        ti = int(rng.integers(0, len(tris)))
        t9 = tris[ti]
        a, b = rng.uniform(-0.2, 1.2, 2)
        target = t9[0:3] + a * (t9[3:6] - t9[0:3]) + b * (t9[6:9] - t9[0:3])
        d = rng.normal(size=3) * rng.uniform(0.01, 4)
        p = np.ascontiguousarray(target - d * rng.uniform(0, 1.4)); m1 = d.copy(); m2 = d.copy()
        t1, t2 = C.c_double(0), C.c_double(0); h1, h2 = np.zeros(3), np.zeros(3)
        u1, u2 = C.c_longlong(0), C.c_longlong(0)
        c1 = R3.ref3_collide_wall(vp(p), vp(m1), vp(t9), 5, 0, C.byref(t1), vp(h1), C.byref(u1))
        c2 = U.orc_unit_collide_wall(vp(p), vp(m2), vp(t9), vp(words), len(words), C.byref(t2), vp(h2), C.byref(u2))
        assert c1 == c2 and (m1 == m2).all() and u1.value == u2.value
        if c1 in (1, 2):
            n_hit += 1
            assert t1.value == t2.value and (h1 == h2).all()
    assert n_hit > 3000


@needs_ref
def test_live_random_collide_mol_and_boxes():
    U = O.unit_lib()
    p, mv, tg, R = mc.mol_pairs(n=20000, seed=321)
    for i in range(len(p)):
        t1, t2 = C.c_double(0), C.c_double(0); h1, h2 = np.zeros(3), np.zeros(3)
        c1 = R3.ref3_collide_mol(vp(p[i]), vp(mv[i]), vp(tg[i]), R, C.byref(t1), vp(h1))
        c2 = U.orc_unit_collide_mol(vp(p[i]), vp(mv[i]), vp(tg[i]), R, C.byref(t2), vp(h2))
        assert c1 == c2
        if c1 == 3:
            assert same(t1.value, t2.value) and same(h1, h2)
    tris = mc.triangles(n=80, seed=5)
    for i, (ti, lo, hi) in enumerate(mc.boxes(tris, per_tri=30, seed=6)):
        lo = np.ascontiguousarray(lo, dtype=np.float64); hi = np.ascontiguousarray(hi, dtype=np.float64)
        assert (R3.ref3_wall_in_box(vp(tris[ti]), vp(lo), vp(hi)) != 0) == bool(U.orc_unit_wall_in_box(vp(tris[ti]), vp(lo), vp(hi)))


@needs_ref
def test_live_random_exact_disk_against_compiled_reference():
    U = O.unit_lib()
    cases, R = mc.disk_cases(n=6000, seed=4242)
    n_partial = 0
    for loc, mv, tg, walls in cases:
        a = U.orc_unit_exact_disk(vp(loc), vp(mv), R, vp(tg), len(walls), vp(walls))
        b = R3.ref3_exact_disk(vp(loc), vp(mv), R, vp(tg), len(walls), vp(walls))
        assert same(a, b), (a, b)
        n_partial += 0 < b < 1
    assert n_partial > 1500
