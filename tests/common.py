"""Scenario builders shared by the CPU and GPU test tiers (same seeded inputs for oracle and product)."""
import numpy as np

from mcell_b200 import abi
from mcell_b200.model import Model, Config, MolArrays, create_box, create_icosphere, release_uniform_box


def free_diffusion_box(n=20000, edge_um=1.0, seed=1, D=1e-6, rng_mode=abi.MCX_RNG_PHILOX, cap_factor=2):
    """BASELINE config 1 shape: one species in a reflective cube."""
    m = Model(Config(seed=seed))
    m.add_species("A", D)
    v, f = create_box(edge_um)
    m.add_geometry_object(v, f)
    t = m.build(max_molecules=int(n * cap_factor) + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n, edge_um, t.length_unit, margin=1e-4)
    return t, MolArrays.from_positions(pos, 0)


def reactive_box(n=20000, edge_um=0.5, seed=1, p_target=0.3, rng_mode=abi.MCX_RNG_PHILOX, density_scale=1.0,
                 subpartition_dimension=0.5, cap_factor=2, products=("C",), max_resolve_rounds=0, cell_edge=0.0):
    """BASELINE config 2 shape: A + B -> C in a reflective box (rate chosen so max_fixed_p = p_target)."""
    m = Model(Config(seed=seed, subpartition_dimension=subpartition_dimension))
    m.add_species("A", 1e-6)
    m.add_species("B", 1e-6)
    m.add_species("C", 0.5e-6)
    pb = _pb_factor(m, 0, 1)
    m.add_reaction_rule(["A", "B"], list(products), p_target / pb)
    v, f = create_box(edge_um)
    m.add_geometry_object(v, f)
    t = m.build(max_molecules=int(n * cap_factor) + 64, rng_mode=rng_mode, max_resolve_rounds=max_resolve_rounds,
                cell_edge=cell_edge)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n, edge_um, t.length_unit, margin=1e-4)
    species = (np.arange(n) % 2).astype(np.uint32)
    return t, MolArrays.from_positions(pos, species)


def _pb_factor(m, a, b):
    import math
    from mcell_b200.model import N_AV, MY_PI
    lu, ts = m.length_unit, m.config.time_step
    eff = (m.space_step(m.species[a].diffusion_constant_3d) + m.space_step(m.species[b].diffusion_constant_3d)) * lu / ts
    R = m.rxn_radius_um
    return 1.0 / (2.0 * math.sqrt(MY_PI) * R * R * eff) * 1.0e15 / N_AV


def sphere_classes(n=20000, radius_um=0.25, subdivisions=3, seed=1, rng_mode=abi.MCX_RNG_PHILOX, box_um=0.8):
    """BASELINE config 3 shape without receptors: ligand L around/inside an icosphere whose faces are
    absorptive (class 0), transparent (class 1) or reflective, inside a reflective bounding box."""
    m = Model(Config(seed=seed))
    m.add_species("L", 1e-6)
    m.add_species("M", 2e-6)
    sv, sf = create_icosphere(radius_um, subdivisions)
    cls = np.full(len(sf), abi.MCX_NONE, np.uint32)
    cz = sv[sf].mean(axis=1)[:, 2]
    cls[cz > 0.08] = 0      # top cap absorbs L only
    cls[cz < -0.08] = 1     # bottom cap transparent for everybody
    m.add_geometry_object(sv, sf, cls)
    bv, bf = create_box(box_um)
    m.add_geometry_object(bv, bf)
    m.add_surface_property(0, abi.MCX_SURF_ABSORPTIVE, species="L")
    m.add_surface_property(1, abi.MCX_SURF_TRANSPARENT, species=None)
    t = m.build(max_molecules=2 * n + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n, box_um, t.length_unit, margin=1e-3)
    species = (np.arange(n) % 2).astype(np.uint32)
    return t, MolArrays.from_positions(pos, species)


def reversible_box(n=20000, edge_um=0.5, seed=1, rng_mode=abi.MCX_RNG_PHILOX, k_off=2e5, p_target=0.4):
    """Ca + CB <-> CaCB (config 4 chemistry without the mesh): bimolecular + unimolecular."""
    m = Model(Config(seed=seed))
    m.add_species("Ca", 2e-6)
    m.add_species("CB", 0.3e-6)
    m.add_species("CaCB", 0.3e-6)
    pb = _pb_factor(m, 0, 1)
    m.add_reaction_rule(["Ca", "CB"], ["CaCB"], p_target / pb)
    m.add_reaction_rule(["CaCB"], ["Ca", "CB"], k_off)
    v, f = create_box(edge_um)
    m.add_geometry_object(v, f)
    t = m.build(max_molecules=3 * n + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n, edge_um, t.length_unit, margin=1e-4)
    species = (np.arange(n) % 3).astype(np.uint32)
    return t, MolArrays.from_positions(pos, species, schedule_unimol=True)


def isaac_slices(seed, n_ids, words_per_mol):
    """Per-molecule tapes cut from one ISAAC64 stream of the reference RNG restatement."""
    import ctypes as C
    from oracle import oracle_py as O
    L = O.lib()
    r = C.c_void_p(L.orc_rng_new(C.c_uint32(seed)))
    words = np.zeros(n_ids * words_per_mol, np.uint32)
    L.orc_rng_fill_uint(r, C.c_void_p(words.ctypes.data), len(words))
    L.orc_rng_free(r)
    off = (np.arange(n_ids, dtype=np.uint64) * np.uint64(words_per_mol))
    return words, off


def rel_close(a, b, tol=1e-12):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))


TRACE_INT_FIELDS = ("outcome", "n_words", "n_wall_hits", "n_collisions", "n_redo", "wall", "wall_side", "partner",
                    "rxn_class", "rxn_pathway", "rxn_partner", "event_hash")


def compare_traces(tr_a, tr_b, ids, check_rounds=False):
    """Bit-exact on every integer field, 1e-12 relative on positions/times. Returns list of mismatch strings."""
    bad = []
    a, b = tr_a[ids], tr_b[ids]
    for f in TRACE_INT_FIELDS + (("rounds",) if check_rounds else ()):
        eq = a[f] == b[f]
        if eq.ndim > 1:
            eq = eq.all(axis=1)
        if not eq.all():
            k = np.flatnonzero(~eq)[:5]
            bad.append("%s differs for ids %s: %s vs %s" % (f, ids[k], a[f][k], b[f][k]))
    ok = rel_close(a["pos"], b["pos"]).all(axis=1)
    if not ok.all():
        k = np.flatnonzero(~ok)[:5]
        bad.append("pos differs for ids %s: %s vs %s" % (ids[k], a["pos"][k], b["pos"][k]))
    ok = rel_close(a["t_event"], b["t_event"])
    if not ok.all():
        k = np.flatnonzero(~ok)[:5]
        bad.append("t_event differs for ids %s" % ids[k])
    return bad
