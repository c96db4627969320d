"""CPU tier: PINS the neighbour-tile search and the reaction test of surface-surface reactions against
MCell4's OWN compiled code.

tests/golden/mcell4_tiles_vectors.npz holds the outputs of oracle/_ref/libmcell4tiles.so — the reference's
GridUtils::find_neighbor_tiles and everything under it (src4/grid_utils.inl:296-1801), RxnUtils::test_bimolecular with a
local probability factor and test_many_bimolecular (src4/rxn_utils.inl:336-414, 475-580), cut out of the reference files
by line range at build time and compiled unmodified (oracle/ref_mcell4_tiles_shim.cpp, oracle/Makefile: ref) — on the
cases of tests/golden/mcell4_tiles_cases.py.  Three statements: the oracle's restatement (oracle/oracle_tiles.h)
reproduces MCell4 entry for entry, with every grid present and with grids missing; so does the PRODUCT's table
(mcx_tile_neighbor_table, the host code behind the device's tn_start / tn_list) once the entries of walls without a grid
are left out the way the device leaves them out; and, where the compiled code is present, the same on fresh cases."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import mcell4_tiles_cases as tc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from test_oracle_vs_reference import ref_words, vp  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "mcell4_tiles_vectors.npz"))


def _vp(a):
    return None if a is None else vp(a)


def oracle_table(V, T, mask, per):
    L = O.lib()
    L.orc_unit_neighbor_tile_table.restype = C.c_ulonglong
    nt = int(per.sum())
    start = np.zeros(nt + 1, np.uint32)
    out = np.zeros(2 * 48 * nt, np.uint32)
    n = L.orc_unit_neighbor_tile_table(vp(V), len(V), vp(T), len(T), _vp(mask), vp(start), vp(out), C.c_ulonglong(48 * nt))
    ntl = int(per[mask.astype(bool)].sum()) if mask is not None else nt
    return start[:ntl + 1].copy(), out[:2 * n].copy()


def product_table(V, T, per):
    from mcell_b200.engine import load_library
    L = load_library()
    L.mcx_tile_neighbor_table.restype = C.c_uint64
    L.mcx_tile_neighbor_table.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    nt = int(per.sum())
    start = np.zeros(nt + 1, np.uint32)
    out = np.zeros(2 * 48 * nt, np.uint32)
    n = L.mcx_tile_neighbor_table(vp(V), len(V), vp(T), len(T), None, vp(start), vp(out), 48 * nt)
    return start, out[:2 * n].reshape(-1, 2)


def filtered(start, pairs, mask, per):
    """What the device does with the static table (mcx_device.cuh: react_2D block): tiles of walls without a grid do not
    exist, entries that name such a wall are skipped."""
    first = np.concatenate([[0], np.cumsum(per)])
    s_out, p_out = [0], []
    for w in range(len(per)):
        if mask is not None and not mask[w]:
            continue
        for t in range(int(per[w])):
            g = int(first[w]) + t
            e = pairs[start[g]:start[g + 1]]
            if mask is not None:
                e = e[mask[e[:, 0]].astype(bool)]
            p_out.append(e)
            s_out.append(s_out[-1] + len(e))
    return np.array(s_out, np.uint32), (np.concatenate(p_out) if p_out else np.zeros((0, 2), np.uint32)).reshape(-1)


def test_oracle_find_neighbor_tiles_equals_compiled_mcell4():
    n_tiles = n_entries = 0
    sizes = set()
    for k, (V, T) in enumerate(tc.meshes()):
        per = G["per_%d" % k]
        sizes |= set(per.tolist())
        for q, mask in enumerate(tc.grid_masks(len(T), k)):
            start, pairs = oracle_table(V, T, mask, per)
            assert np.array_equal(start, G["start_%d_%d" % (k, q)]), (k, q)
            assert np.array_equal(pairs, G["pairs_%d_%d" % (k, q)]), (k, q)
            n_tiles += len(start) - 1; n_entries += len(pairs) // 2
    assert n_tiles > 15000 and n_entries > 120000 and {1, 4, 9, 16, 25, 36} <= sizes


def test_product_neighbor_tile_table_equals_compiled_mcell4():
    """The table libmcx builds once per geometry, filtered per wall like the device filters it, is MCell4's list for every
    tile — in the reference's order, which decides which partner a molecule with several candidates reacts with."""
    longest = 0
    for k, (V, T) in enumerate(tc.meshes()):
        per = G["per_%d" % k]
        start, pairs = product_table(V, T, per)
        longest = max(longest, int(np.diff(start).max()))
        for q, mask in enumerate(tc.grid_masks(len(T), k)):
            s, p = filtered(start, pairs, mask, per)
            assert np.array_equal(s, G["start_%d_%d" % (k, q)]), (k, q)
            assert np.array_equal(p, G["pairs_%d_%d" % (k, q)]), (k, q)
    assert 12 <= longest <= 32   # the device holds up to 32 matching neighbours (SURFSURF_MAX_MATCHES)


def test_test_bimolecular_with_local_prob_factor_and_test_many_bimolecular_match_compiled_mcell4():
    L = O.lib()
    L.orc_unit_test_bimolecular_lpf.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_uint64, C.c_void_p]
    ref = G["lpf_out"]
    reacted = 0
    for i, (cum, scaling, lpf, seed, skip) in enumerate(tc.lpf_cases()):
        tape = ref_words(seed, skip + 8)[skip:]
        used = C.c_longlong(0)
        r = L.orc_unit_test_bimolecular_lpf(vp(np.ascontiguousarray(cum)), len(cum), scaling, lpf, vp(tape), len(tape), C.byref(used))
        assert r == int(ref[i, 0]) and used.value == int(ref[i, 1]) == 1, (i, r, ref[i])
        reacted += r >= 0
    assert 40 < reacted < 360
    L.orc_unit_test_many_bimolecular.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    ref = G["many_out"]
    chosen = set()
    for i, (cums, scaling, lpf, seed, skip) in enumerate(tc.many_cases()):
        flat = np.ascontiguousarray(np.concatenate(cums)); npw = np.array([len(c) for c in cums], np.int32)
        tape = ref_words(seed, skip + 8)[skip:]
        pw = C.c_int(0); used = C.c_longlong(0)
        r = L.orc_unit_test_many_bimolecular(vp(flat), vp(npw), len(cums), vp(np.ascontiguousarray(scaling)), lpf, vp(tape), len(tape),
                                             C.byref(pw), C.byref(used))
        assert r == int(ref[i, 0]) and used.value == int(ref[i, 2]) == 1, (i, r, ref[i])
        if r >= 0:
            assert pw.value == int(ref[i, 1]), (i, pw.value, ref[i])
        chosen.add(r)
    assert {-1, 0, 1, 2} <= chosen


def test_table_builder_surface_surface_classes_use_the_reference_pb_factor():
    """mcell_b200/model.py builds MCX_RXN_BIMOL_SURFSURF classes with MCell3's compute_pb_factor for two surface molecules
    (src/react_util.c:84-100, compiled in libmcell3ref.so: time_unit * grid_density / 6, / 3 when one reactant is
    TARGET_ONLY), rule order of the reactants, orientations, kept reactants and the order of the rule's products."""
    from mcell_b200 import abi
    from mcell_b200.model import Model, Config, create_icosphere
    for tu, gd, a_t, b_t, want in G["pb_surfsurf"]:
        m = Model(Config(time_step=float(tu), surface_grid_density=float(gd)))
        m.add_species("A", 1e-7, surface=True, target_only=bool(a_t))
        m.add_species("B", 1e-7, surface=True, target_only=bool(b_t))
        m.add_species("C", 1e-7, surface=True)
        m.add_species("V", 1e-6)
        m.add_reaction_rule(["B,", "A'"], ["C'", "V,"], 7.0)
        m.add_reaction_rule(["B,", "A'"], ["B,", "C"], 3.0)
        sv, sf = create_icosphere(0.1, 1)
        m.add_geometry_object(sv, sf)
        t = m.build(max_molecules=8)
        c = t.classes[0]
        assert c.kind == abi.MCX_RXN_BIMOL_SURFSURF and (c.reactants[0], c.reactants[1]) == (1, 0)
        assert (c.reactant_orientation[0], c.reactant_orientation[1]) == (-1, 1) and c.n_pathways == 2
        p0, p1 = t.pathways[c.first_pathway], t.pathways[c.first_pathway + 1]
        assert p0.cum_prob == pytest.approx(7.0 * want, rel=1e-15) and p1.cum_prob == pytest.approx(10.0 * want, rel=1e-15)
        assert c.max_fixed_p == p1.cum_prob
        assert p0.keep_reactant_mask == 0 and p0.n_products == 2 and (p0.products[0], p0.products[1]) == (2, 3)
        assert p1.keep_reactant_mask == 1 and p1.n_products == 1 and p1.products[0] == 2 and p1.product_orientation[0] == 0
        # rule order of the products of pathway 1: kept reactant 0 (B, down), then product 0
        assert p1.kept_info & abi.MCX_KEPT_VALID
        assert (p1.kept_info & 0xF) == abi.MCX_KEPT_ORDER_REACTANT and ((p1.kept_info >> 4) & 0xF) == 0
        assert ((p1.kept_info >> 24) & 3) == 2


def _orc_place(case, ms):
    """the oracle's place_general on one case of tc.place_cases(), in the layout of gen_mcell4_tiles_golden.ref_place"""
    from mcell_b200 import abi
    k, occ, si, sites, surf_reac, seed, skip = case
    V, T = ms[k]
    kind, surf_flags, entries = tc.PLACE_SHAPES[si]
    new_at = [q for q, e in enumerate(entries) if e[0] != "K"]
    keep_mask = sum(1 << r for r in range(2) if ("K%d" % r) in entries)
    info = abi.MCX_KEPT_VALID | (1 << 24) | (1 << 26)        # kept reactants carry an orientation: no draws for them
    for q in range(6):
        if q >= len(entries):
            nib = abi.MCX_KEPT_ORDER_END
        elif entries[q][0] == "K":
            nib = abi.MCX_KEPT_ORDER_REACTANT + int(entries[q][1])
        else:
            nib = new_at.index(q)
        info |= nib << (4 * q)
    prod_surf = np.array([1 if entries[q] == "S" else 0 for q in new_at] + [0] * 4, np.uint8)[:4]
    rsurf = np.array(list(surf_flags) + [0], np.uint8)[:2]
    rec = [sites[r] for r in range(len(sites)) if surf_flags[r] and not (keep_mask >> r) & 1]
    rec_arr = np.array([x for s_ in rec for x in s_] + [0, 0, 0, 0], np.uint32)
    rw, rt = sites[surf_reac]
    tape = ref_words(seed, skip + 80)[skip:]
    L = O.lib()
    L.orc_unit_place_general.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_int, C.c_void_p, C.c_uint, C.c_uint,
                                         C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
    ek = np.zeros(6, np.int32); ew = np.zeros(6, np.uint32); et = np.zeros(6, np.uint32)
    used = C.c_longlong(0)
    rc = L.orc_unit_place_general(vp(V), len(V), vp(T), len(T), vp(occ), len(occ), kind, vp(rsurf), keep_mask, info, len(new_at), vp(prod_surf),
                                  rw, rt, vp(rec_arr), len(rec), vp(tape), len(tape), vp(ek), vp(ew), vp(et), C.byref(used))
    row = [rc, used.value]
    for e in range(6):
        kk = int(ek[e]) if (rc == 0 and e < len(entries)) else 0
        row += [kk, int(ew[e]) if kk else -1, int(et[e]) if kk else -1]
    return row


def test_product_placement_on_vacant_tiles_equals_compiled_mcell4():
    """find_surf_product_positions (src4/diffuse_react_event.cpp:1993-2288), MCell4's own function compiled unmodified
    (oracle/_ref/libmcell4place.so), against the oracle's place_general on 359 cases: unimolecular, volume-surface and
    surface-surface rule shapes with kept reactants and volume products in every position of the product list, on six
    meshes at 15-98 % occupancy — which entry gets a recycled tile and which vacant tile each of the others draws (wall and
    tile), RX_BLOCKED both ways (too few vacant tiles; ten failed attempts), and the number of words drawn."""
    ms = tc.meshes()
    ref = G["place_out"]
    cases = tc.place_cases()
    assert len(cases) == len(ref)
    blocked_after_draws = 0
    for i, case in enumerate(cases):
        row = _orc_place(case, ms)
        assert row == ref[i].tolist(), (i, tc.PLACE_SHAPES[case[2]], row, ref[i].tolist())
        blocked_after_draws += row[0] == -2 and row[1] > 0
    assert (ref[:, 0] == 0).sum() > 150 and (ref[:, 0] == -2).sum() > 50 and blocked_after_draws > 3


def test_recycled_tiles_of_surface_surface_pathways_equal_compiled_mcell4():
    """The recycled branches of find_surf_product_positions for two surface reactants (the initiator's tile for a single
    surface product, :2140-2154; the random hand-out of two freed tiles, :2155-2191), compiled MCell4 against the oracle's
    surfsurf_position_bits: which freed tile the c-th created surface product takes, and the words drawn."""
    L = O.lib()
    L.orc_unit_surfsurf_position_bits.restype = C.c_uint
    L.orc_unit_surfsurf_position_bits.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    ref = G["recycle_out"]
    sites = [(0, 1), (3, 2)]          # reactant 0, reactant 1 (gen_mcell4_tiles_golden.py)
    swapped = drew = 0
    for i, (si, init_b, seed, skip) in enumerate(tc.recycle_cases()):
        entries = tc.RECYCLE_SHAPES[si]
        keep_mask = sum(1 << r for r in range(2) if ("K%d" % r) in entries)
        new = [e for e in entries if e[0] != "K"]
        prod_surf = np.array([1 if e == "S" else 0 for e in new] + [0] * 4, np.uint8)[:4]
        tape = ref_words(seed, skip + 80)[skip:]
        used = C.c_longlong(0)
        bits = L.orc_unit_surfsurf_position_bits(keep_mask, len(new), vp(prod_surf), 0 if init_b else 1, vp(tape), len(tape), C.byref(used))
        assert ref[i, 0] == 0 and used.value == ref[i, 1], (i, entries, used.value, ref[i, 1])
        freed = [sites[r] for r in range(2) if not (keep_mask >> r) & 1]
        n_created = int(prod_surf.sum())
        swap = 1 if bits & 64 else 0
        want = [freed[min(swap if c == 0 else 1 - swap, len(freed) - 1)] for c in range(n_created)]
        got = [(int(ref[i, 2 + 3 * c + 1]), int(ref[i, 2 + 3 * c + 2])) for c in range(n_created)]
        assert got == want and all(ref[i, 2 + 3 * c] == 1 for c in range(n_created)), (i, entries, init_b, got, want)
        swapped += swap; drew += used.value > 0
    assert swapped > 20 and drew > 10


def test_react_2d_all_neighbors_as_a_whole_equals_compiled_mcell4():
    """DiffuseReactEvent::react_2D_all_neighbors (src4/diffuse_react_event.cpp:1249-1393) compiled unmodified with
    trigger_bimolecular (rxn_utils.inl:58-97), the neighbour-tile search and the reaction tests under it
    (oracle/_ref/libmcell4react2d.so; outcome_bimolecular is a recorder) against the ORACLE'S WHOLE STEP in the product's
    semantics: static surface molecules of four species (an evaluation is the neighbour test alone), every molecule
    replaying the ISAAC64 stream the reference function was given.  For every molecule evaluated once: the partner it
    reacts with, the class and the pathway are the reference's, and so is the number of words up to the decision — i.e. the
    neighbour list and its order, walls without a grid, the orientation test, time / binding_factor, the local
    probability factor, test_bimolecular vs test_many_bimolecular and the first-pathway quirk all agree."""
    from mcell_b200 import abi
    checked = reacted = many = 0
    for k, (t, mols, seeds) in enumerate(tc.react2d_models()):
        ref = G["react2d_%d" % k]
        n = mols.n
        per = 24
        words = np.concatenate([ref_words(int(sd), per) for sd in seeds]).astype(np.uint32)
        off = (np.arange(n, dtype=np.uint64) * per)
        o = O.Oracle(t)
        o.upload(mols)
        tr, st = o.trace_step(2, n, words, off)
        once = np.flatnonzero(tr["rounds"] == 1)
        assert len(once) > 0.6 * n
        got_partner = np.where(tr["rxn_partner"][once] == abi.MCX_NONE, -1, tr["rxn_partner"][once].astype(np.int64))
        got_class = np.where(tr["rxn_class"][once] == abi.MCX_NONE, -1, tr["rxn_class"][once].astype(np.int64))
        got_path = np.where(tr["rxn_pathway"][once] == abi.MCX_NONE, -1, tr["rxn_pathway"][once].astype(np.int64))
        assert (got_partner == ref[once, 0]).all(), (k, np.flatnonzero(got_partner != ref[once, 0])[:5])
        assert (got_class == ref[once, 1]).all() and (got_path == ref[once, 2]).all(), k
        quiet = once[ref[once, 0] < 0]                     # no reaction: the words drawn are the test's alone
        assert (tr["n_words"][quiet] == ref[quiet, 3]).all(), k
        assert (tr["n_words"][once] >= ref[once, 3]).all(), k
        checked += len(once); reacted += int((ref[once, 0] >= 0).sum()); many += int((tr["n_collisions"][once] > 1).sum())
    assert checked > 5000 and reacted > 1500 and many > 1500


def test_live_mcell4_placement_and_react_2d_on_fresh_cases():
    """The same two comparisons against the compiled reference functions themselves, on cases that are not in the goldens
    (other seeds, other occupancies); skipped where oracle/_ref is not built."""
    ref_dir = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
    if not (os.path.exists(os.path.join(ref_dir, "libmcell4place.so")) and os.path.exists(os.path.join(ref_dir, "libmcell4react2d.so"))):
        pytest.skip("oracle/_ref/libmcell4place.so / libmcell4react2d.so are not built here (need the reference tree)")
    import gen_mcell4_tiles_golden as gen
    LP = C.CDLL(os.path.join(ref_dir, "libmcell4place.so"))
    ms = tc.meshes()
    rng = np.random.default_rng(1234)
    n_ok = 0
    for case in tc.place_cases(n_per_mesh=25):
        k, occ, si, sites, surf_reac, seed, skip = case
        case = (k, occ, si, sites, surf_reac, int(rng.integers(1000, 9000)), int(rng.integers(0, 30)))   # other streams
        want = gen.ref_place(LP, ms, case)
        assert _orc_place(case, ms) == want, (tc.PLACE_SHAPES[si], want)
        n_ok += want[0] == 0
    assert n_ok > 40
    from mcell_b200 import abi
    t, mols, _ = tc.react2d_models()[1]
    seeds = rng.integers(200000, 900000, mols.n).astype(np.uint32)
    ref = gen.ref_react2d(t, mols, seeds)
    words = np.concatenate([ref_words(int(sd), 24) for sd in seeds]).astype(np.uint32)
    o = O.Oracle(t)
    o.upload(mols)
    tr, _ = o.trace_step(2, mols.n, words, np.arange(mols.n, dtype=np.uint64) * 24)
    once = np.flatnonzero(tr["rounds"] == 1)
    got = np.where(tr["rxn_partner"][once] == abi.MCX_NONE, -1, tr["rxn_partner"][once].astype(np.int64))
    assert (got == ref[once, 0]).all() and (ref[once, 0] >= 0).sum() > 500


def test_live_mcell4_on_fresh_meshes():
    path = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libmcell4tiles.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libmcell4tiles.so is not built here (needs the reference tree)")
    R = C.CDLL(path)
    R.ref4_neighbor_tile_table.restype = C.c_ulonglong
    from mcell_b200.model import create_icosphere
    rng = np.random.default_rng(77)
    for sub, scale in ((1, 80), (2, 140), (3, 420)):
        v, t = create_icosphere(0.05, sub)
        V = np.ascontiguousarray(np.asarray(v, np.float64) * scale * (1 + 0.3 * rng.random((len(v), 3))))
        T = np.ascontiguousarray(np.asarray(t), np.uint32)
        T = np.ascontiguousarray(np.stack([np.roll(row, int(rng.integers(0, 3))) for row in T]))   # rotate vertex orders
        per = np.zeros(len(T), np.uint32)
        nt = R.ref4_tiles_num_tiles(vp(V), len(V), vp(T), len(T), vp(per))
        for mask in (None, (rng.random(len(T)) < 0.6).astype(np.uint8)):
            start = np.zeros(nt + 1, np.uint32); out = np.zeros(2 * 48 * nt, np.uint32)
            n = R.ref4_neighbor_tile_table(vp(V), len(V), vp(T), len(T), _vp(mask), vp(start), vp(out), C.c_ulonglong(48 * nt))
            ntl = int(per[mask.astype(bool)].sum()) if mask is not None else nt
            s_o, p_o = oracle_table(V, T, mask, per)
            assert np.array_equal(s_o, start[:ntl + 1]) and np.array_equal(p_o, out[:2 * n])
            s_p, p_p = product_table(V, T, per)
            s_f, p_f = filtered(s_p, p_p, mask, per)
            assert np.array_equal(s_f, start[:ntl + 1]) and np.array_equal(p_f, out[:2 * n])
