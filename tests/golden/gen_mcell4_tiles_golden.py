"""Generates tests/golden/mcell4_tiles_vectors.npz from oracle/_ref/libmcell4tiles.so — MCell4's own find_neighbor_tiles
(src4/grid_utils.inl:296-1801), test_bimolecular with a local probability factor and test_many_bimolecular
(src4/rxn_utils.inl:336-414, 475-580), cut out of the reference files by line range and compiled unmodified
(oracle/ref_mcell4_tiles_shim.cpp, `make -C oracle ref`).  Run in the build container (needs /root/reference):
    python tests/golden/gen_mcell4_tiles_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
import mcell4_tiles_cases as tc  # noqa: E402


def vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def ref_table(L, V, T, mask):
    per = np.zeros(len(T), np.uint32)
    nt = L.ref4_tiles_num_tiles(vp(V), len(V), vp(T), len(T), vp(per))
    start = np.zeros(nt + 1, np.uint32)
    out = np.zeros(2 * 48 * nt, np.uint32)
    n = L.ref4_neighbor_tile_table(vp(V), len(V), vp(T), len(T), vp(mask), vp(start), vp(out), C.c_ulonglong(48 * nt))
    ntl = int(per[mask.astype(bool)].sum()) if mask is not None else nt
    return per, start[:ntl + 1].copy(), out[:2 * n].copy()


def ref_place(LP, ms, case, shape=None):
    """-> [rc, words, (kind, wall, tile) x 6] with kind 0 nothing / 1 recycled / 2 vacant"""
    k, occ, si, sites, surf_reac, seed, skip = case
    V, T = ms[k]
    kind, surf_flags, entries = shape or tc.PLACE_SHAPES[si]
    ent = np.array([1 if (e == "S" or (e[0] == "K" and surf_flags[int(e[1])])) else 0 for e in entries], np.uint8)
    keep = [("K%d" % r) in entries for r in range(2)]
    def r5(site):
        return np.array([0, 0, 0, 0, 0], np.float64) if site is None else np.array([1, site[0], site[1], 0.1, 0.1], np.float64)
    a5 = r5(sites[0]); b5 = r5(sites[1]) if len(sites) > 1 else None
    ot = np.zeros(6, np.int32); ow = np.zeros(6, np.uint32); otl = np.zeros(6, np.uint32)
    nsp = C.c_uint(0); used = C.c_int(0); words = C.c_longlong(0)
    LP.ref4_find_surf_product_positions.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_int, C.c_void_p, C.c_uint,
                                                    C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = LP.ref4_find_surf_product_positions(vp(V), len(V), vp(T), len(T), vp(occ), len(occ), 1 if kind == 1 else 0, vp(ent), len(ent),
                                             vp(a5), int(keep[0]), vp(b5), int(keep[1]), surf_reac, seed, skip, vp(ot), vp(ow), vp(otl),
                                             C.byref(nsp), C.byref(used), C.byref(words))
    row = [rc, words.value]
    for e in range(6):
        kk = 0
        if rc == 0 and e < len(ent):
            kk = 1 if ot[e] in (2, 3) else (2 if ot[e] == 5 else 0)
        row += [kk, int(ow[e]) if kk else -1, int(otl[e]) if kk else -1]
    return row


def ref_react2d(t, mols, seeds):
    """per molecule: partner id (-1 none), class, pathway, words drawn — from the compiled reference function"""
    LR = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libmcell4react2d.so"))
    V = np.ascontiguousarray(t.vertices, np.float64); T = np.ascontiguousarray(t.tri, np.uint32)
    n = mols.n
    has_grid = np.zeros(len(T), np.uint8); has_grid[mols.wall[:n]] = 1
    m4 = np.ascontiguousarray(np.stack([mols.wall[:n], mols.tile[:n], mols.species[:n], mols.orientation[:n]], 1).astype(np.int32))
    ns = t.n_species
    table = np.full(ns * ns, -1, np.int32); geom = []; npw = []; cum = []
    for c in range(t.n_classes):
        rc = t.classes[c]
        table[rc.reactants[0] * ns + rc.reactants[1]] = c; table[rc.reactants[1] * ns + rc.reactants[0]] = c
        geom += [rc.reactant_orientation[0], rc.reactant_orientation[1]]; npw.append(rc.n_pathways)
        cum += [t.pathways[rc.first_pathway + q].cum_prob for q in range(rc.n_pathways)]
    geom = np.array(geom, np.int32); npw = np.array(npw, np.int32); cum = np.array(cum, np.float64)
    out4 = np.zeros(4 * n, np.int32)
    LR.ref4_react_2d_all_neighbors.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p,
                                               C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    LR.ref4_react_2d_all_neighbors(vp(V), len(V), vp(T), len(T), vp(has_grid), vp(m4), n, ns, vp(table), t.n_classes, vp(geom), vp(npw), vp(cum),
                                   1.0, vp(np.ascontiguousarray(seeds)), vp(out4))
    return out4.reshape(n, 4)


def main():
    L = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libmcell4tiles.so"))
    L.ref4_neighbor_tile_table.restype = C.c_ulonglong
    out = {}
    for k, (V, T) in enumerate(tc.meshes()):
        for q, mask in enumerate(tc.grid_masks(len(T), k)):
            per, start, pairs = ref_table(L, V, T, mask)
            out["per_%d" % k] = per
            out["start_%d_%d" % (k, q)] = start
            out["pairs_%d_%d" % (k, q)] = pairs
    L.ref4_test_bimolecular_lpf.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_uint, C.c_uint, C.c_void_p]
    rows = []
    for cum, scaling, lpf, seed, skip in tc.lpf_cases():
        used = C.c_longlong(0)
        r = L.ref4_test_bimolecular_lpf(vp(np.ascontiguousarray(cum)), len(cum), scaling, lpf, seed, skip, C.byref(used))
        rows.append((r, used.value))
    out["lpf_out"] = np.array(rows, np.int64)
    L.ref4_test_many_bimolecular.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
    rows = []
    for cums, scaling, lpf, seed, skip in tc.many_cases():
        flat = np.ascontiguousarray(np.concatenate(cums)); npw = np.array([len(c) for c in cums], np.int32)
        pw = C.c_int(0); used = C.c_longlong(0)
        r = L.ref4_test_many_bimolecular(vp(flat), vp(npw), len(cums), vp(np.ascontiguousarray(scaling)), lpf, seed, skip, C.byref(pw), C.byref(used))
        rows.append((r, pw.value, used.value))
    out["many_out"] = np.array(rows, np.int64)
    # compute_pb_factor of MCell3 (src/react_util.c:84-100, libmcell3ref.so) for two surface molecules: (time_unit,
    # grid_density, a TARGET_ONLY, b TARGET_ONLY) -> factor
    L3 = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libmcell3ref.so"))
    L3.ref3_compute_pb_factor_surfsurf.restype = C.c_double
    L3.ref3_compute_pb_factor_surfsurf.argtypes = [C.c_double] * 3 + [C.c_int] * 2
    out["pb_surfsurf"] = np.array([[tu, gd, a, b, L3.ref3_compute_pb_factor_surfsurf(tu, 1 / np.sqrt(gd), gd, a, b)]
                                   for tu, gd in ((1e-6, 1e4), (5e-7, 1.5e4)) for a, b in ((0, 0), (1, 0), (0, 1))])
    # find_surf_product_positions (src4/diffuse_react_event.cpp:1993-2288, libmcell4place.so)
    LP = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libmcell4place.so"))
    rows = []
    ms = tc.meshes()
    for case in tc.place_cases():
        rows.append(ref_place(LP, ms, case))
    out["place_out"] = np.array(rows, np.int64)
    # the recycled branches of the same function for two surface reactants
    V, T = ms[2]
    rows = []
    for si, init_b, seed, skip in tc.recycle_cases():
        entries = tc.RECYCLE_SHAPES[si]
        case = (2, np.array([[0, 1], [3, 2]], np.uint32), None, [(0, 1), (3, 2)], init_b, seed, skip)
        rows.append(ref_place(LP, ms, case, shape=(5, (1, 1), entries)))
    out["recycle_out"] = np.array(rows, np.int64)
    # react_2D_all_neighbors as a whole (src4/diffuse_react_event.cpp:1249-1393, libmcell4react2d.so)
    for k, (t, mols, seeds) in enumerate(tc.react2d_models()):
        out["react2d_%d" % k] = ref_react2d(t, mols, seeds)
    np.savez_compressed(os.path.join(HERE, "mcell4_tiles_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
