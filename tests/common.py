"""Scenario builders shared by the CPU and GPU test tiers (same seeded inputs for oracle and product)."""
import numpy as np

from mcell_b200 import abi
from mcell_b200.model import (Model, Config, MolArrays, create_box, create_icosphere, release_uniform_box, release_on_walls,
                              counted_volume_of)


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
                 subpartition_dimension=0.5, cap_factor=2, products=("C",), max_resolve_rounds=0, cell_edge=0.0, D=1e-6):
    """BASELINE config 2 shape: A + B -> C in a reflective box (rate chosen so max_fixed_p = p_target)."""
    m = Model(Config(seed=seed, subpartition_dimension=subpartition_dimension))
    m.add_species("A", D)
    m.add_species("B", D)
    m.add_species("C", 0.5 * D)
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


def ligand_receptor_sphere(n_lig=6000, n_rec=1500, n_pump=600, radius_um=0.25, subdivisions=3, seed=1, box_um=0.8,
                           rng_mode=abi.MCX_RNG_PHILOX, p_bind=0.5, k_off=1e5, k_pump=2e5, release_products=True, regions=False,
                           max_molecules=None):
    """BASELINE config 3/4 surface chemistry on an icosphere inside a reflective box:
       L' + R' -> LR'        ligand binds receptors from the outside (front) only
       LR'     -> L' + R'    unbinding releases the ligand on the outside
       Ca, + P' -> CaP'      pumps take calcium from the inside (back) ...
       CaP'    -> P' + Ca'   ... and release it outside."""
    import math
    from mcell_b200.model import N_AV, MY_PI
    m = Model(Config(seed=seed))
    L = m.add_species("L", 1e-6)
    Ca = m.add_species("Ca", 2e-6)
    R = m.add_species("R", 0.0, surface=True)
    LR = m.add_species("LR", 0.0, surface=True)
    P = m.add_species("P", 0.0, surface=True)
    CaP = m.add_species("CaP", 0.0, surface=True)

    def k_for(p, D):  # vol-surf pb_factor with both orientations in one class (src/react_util.c:145-157)
        pb = 2.0 * 1.0e11 * m.config.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * m.config.time_step / D)
        return p / pb

    # release_products=False: the bound ligand / pumped calcium is degraded, so every reaction has one product and
    # molecule ids stay deterministic (fresh ids of second products come from device atomics)
    m.add_reaction_rule(["L'", "R'"], ["LR'"], k_for(p_bind, 1e-6))
    m.add_reaction_rule(["LR'"], ["L'", "R'"] if release_products else ["R'"], k_off)
    m.add_reaction_rule(["Ca,", "P'"], ["CaP'"], k_for(p_bind, 2e-6))
    m.add_reaction_rule(["CaP'"], ["P'", "Ca'"] if release_products else ["P'"], k_pump)
    sv, sf = create_icosphere(radius_um, subdivisions)
    m.add_geometry_object(sv, sf)
    bv, bf = create_box(box_um)
    m.add_geometry_object(bv, bf)
    if regions:   # two overlapping counted surface regions of the sphere: walls fall into the sets {}, {north}, {band}, {north, band}
        cz = np.asarray(sv)[np.asarray(sf)].mean(axis=1)[:, 2]
        m.add_surface_region("north", 0, np.flatnonzero(cz > 0))
        m.add_surface_region("band", 0, np.flatnonzero(np.abs(cz) < 0.4 * radius_um))
    n_total = n_lig + n_rec + n_pump
    t = m.build(max_molecules=max_molecules or 2 * n_total + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n_lig, box_um, t.length_unit, margin=1e-3)
    vol = MolArrays.from_positions(pos, (np.arange(n_lig) % 2).astype(np.uint32) * Ca + (1 - np.arange(n_lig) % 2).astype(np.uint32) * L,
                                   schedule_unimol=True)
    sphere_walls = np.arange(len(sf), dtype=np.uint32)
    surf = release_on_walls(rng, t, sphere_walls, n_rec + n_pump, R, orientation=1, first_id=n_lig)
    surf.species[n_rec:] = P
    return t, MolArrays.concat([vol, surf])


def transporter_sphere(n_vol=20000, n_trans=2500, n_enz=1500, radius_um=0.25, subdivisions=3, seed=1, box_um=0.8,
                       rng_mode=abi.MCX_RNG_PHILOX, p_react=0.5, enzyme=True):
    """Kept volume reactants of surface reactions (diffuse_react_event.cpp:945-975, 2689-2716) on a counted
    icosphere inside a reflective box:
       A' + T' -> A, + T'          transporter: A hits T from the front and passes through the wall (RX_FLIP)
       S' + E' -> S' + E' + Pr'    enzyme: S and E are kept, the product is released in front of the wall; S reflects"""
    import math
    from mcell_b200.model import N_AV, MY_PI
    m = Model(Config(seed=seed))
    A = m.add_species("A", 1e-6)
    S = m.add_species("S", 1e-6)
    Pr = m.add_species("Pr", 2e-6)
    T = m.add_species("T", 0.0, surface=True)
    E = m.add_species("E", 0.0, surface=True)

    def k_for(p, D):
        pb = 2.0 * 1.0e11 * m.config.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * m.config.time_step / D)
        return p / pb

    m.add_reaction_rule(["A'", "T'"], ["A,", "T'"], k_for(p_react, 1e-6))
    if enzyme:   # its product takes a fresh id; without it molecule ids stay deterministic
        m.add_reaction_rule(["S'", "E'"], ["S'", "E'", "Pr'"], k_for(p_react, 1e-6))
    sv, sf = create_icosphere(radius_um, subdivisions)
    m.add_geometry_object(sv, sf, counted=True)
    bv, bf = create_box(box_um)
    m.add_geometry_object(bv, bf, counted=True)
    n_total = n_vol + n_trans + n_enz
    t = m.build(max_molecules=3 * n_total + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n_vol, box_um, t.length_unit, margin=1e-3)
    vol = MolArrays.from_positions(pos, (np.arange(n_vol) % 2).astype(np.uint32) * S + (1 - np.arange(n_vol) % 2).astype(np.uint32) * A,
                                   schedule_unimol=True)
    vol.counted_volume[:] = counted_volume_of(t, pos)
    sphere_walls = np.arange(len(sf), dtype=np.uint32)
    surf = release_on_walls(rng, t, sphere_walls, n_trans + n_enz, T, orientation=1, first_id=n_vol)
    surf.species[n_trans:] = E
    return t, MolArrays.concat([vol, surf])


def permeable_sphere(n=20000, radius_um=0.25, subdivisions=3, seed=1, box_um=0.8, rng_mode=abi.MCX_RNG_PHILOX,
                     p_in=0.3, p_out=0.15, products=True):
    """Finite-rate reactions with a surface class (collide_and_react_with_walls -> test_intersect ->
    outcome_intersect) on a counted icosphere of surface class 0 inside a reflective box:
       A' @ sc -> A,          A crosses inwards with probability p_in per hit from outside (RX_FLIP)
       A, @ sc -> A'          ... and outwards with p_out per hit from inside
       B' @ sc -> C' + D,     (products=True) B is consumed; C appears outside, D inside
       E' @ sc -> E' + F,     E is kept (reflects), F appears inside"""
    import math
    from mcell_b200.model import N_AV, MY_PI
    m = Model(Config(seed=seed))
    for name in ("A", "B", "C", "D", "E", "F"):
        m.add_species(name, 1e-6)

    def k_for(p):   # marked reactant: doubled factor (src/react_util.c:145-157)
        pb = 2.0 * 1.0e11 * m.config.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * m.config.time_step / 1e-6)
        return p / pb

    m.add_surface_class_reaction(0, "A'", ["A,"], k_for(p_in))
    m.add_surface_class_reaction(0, "A,", ["A'"], k_for(p_out))
    if products:
        m.add_surface_class_reaction(0, "B'", ["C'", "D,"], k_for(0.4))
        m.add_surface_class_reaction(0, "E'", ["E'", "F,"], k_for(0.25))
    sv, sf = create_icosphere(radius_um, subdivisions)
    m.add_geometry_object(sv, sf, surf_class=0, counted=True)
    bv, bf = create_box(box_um)
    m.add_geometry_object(bv, bf, counted=True)
    t = m.build(max_molecules=3 * n + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n, box_um, t.length_unit, margin=1e-3)
    sp = np.zeros(n, np.uint32)
    if products:
        sp = (np.arange(n) % 3).astype(np.uint32) * 0 + np.where(np.arange(n) % 3 == 1, 1, 0).astype(np.uint32) + \
            np.where(np.arange(n) % 3 == 2, 4, 0).astype(np.uint32)     # A, B, E in equal parts
    mols = MolArrays.from_positions(pos, sp, schedule_unimol=True)
    mols.counted_volume[:] = counted_volume_of(t, pos)
    return t, mols


def diffusing_receptors(n_rec=3000, n_lig=8000, radius_um=0.25, subdivisions=3, seed=1, box_um=0.8, D_surf=1e-7,
                        rng_mode=abi.MCX_RNG_PHILOX, p_bind=0.5, k_off=1e5, with_ligand=True, border=None):
    """Surface diffusion: receptors R diffuse on an icosphere (diffuse_surf_molecule, ray_trace_surf
    across triangle edges, one molecule per tile), ligand L binds them from outside, LR' -> R' keeps ids deterministic."""
    import math
    from mcell_b200.model import N_AV, MY_PI
    m = Model(Config(seed=seed))
    L = m.add_species("L", 1e-6)
    R = m.add_species("R", D_surf, surface=True)
    LR = m.add_species("LR", D_surf * 0.5, surface=True)
    pb = 2.0 * 1.0e11 * m.config.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * m.config.time_step / 1e-6)
    if with_ligand:
        m.add_reaction_rule(["L'", "R'"], ["LR'"], p_bind / pb)
        m.add_reaction_rule(["LR'"], ["R'"], k_off)
    sv, sf = create_icosphere(radius_um, subdivisions)
    m.add_geometry_object(sv, sf)
    bv, bf = create_box(box_um)
    m.add_geometry_object(bv, bf)
    if border is not None:
        # a reactive region "cap" (surface class 1) whose outline R cannot cross: abi.MCX_SURF_REFLECTIVE turns it back,
        # abi.MCX_SURF_ABSORPTIVE (absorptive region border) takes it; LR is not affected
        cz = np.asarray(sv)[np.asarray(sf)].mean(axis=1)[:, 2]
        m.add_surface_region("cap", 0, np.flatnonzero(cz > 0.3 * radius_um), surf_class=1)
        m.add_surface_property(1, border, species="R")
    t = m.build(max_molecules=2 * (n_rec + n_lig) + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n_lig, box_um, t.length_unit, margin=1e-3)
    vol = MolArrays.from_positions(pos, L, schedule_unimol=True)
    surf = release_on_walls(rng, t, np.arange(len(sf), dtype=np.uint32), n_rec, R, orientation=1, first_id=n_lig)
    return t, MolArrays.concat([vol, surf])


def surface_reactions(n_a=1500, n_b=1500, n_e=300, radius_um=0.25, subdivisions=3, seed=1, box_um=0.8, D_surf=2e-7,
                      rng_mode=abi.MCX_RNG_PHILOX, p=0.3, k_off=2e5, static_b=False, max_molecules=None):
    """Surface-surface reactions (react_2D_all_neighbors): A' + B' -> C' (both consumed, C on the initiator's
    tile), C' -> A' + V' would need a second tile and is not part of it: C' -> A' keeps ids deterministic; A' + E' -> D' + E'
    (catalytic: E kept), D' + D' -> A' + B' (same species, two surface products over the two freed tiles: the random tile
    assignment), B' + E, -> V, (orientation classes that never match the release: E is released up).  static_b: B cannot
    diffuse but still initiates reactions (SPECIES_FLAG_CAN_SURFSURF molecules are diffused every step)."""
    m = Model(Config(seed=seed))
    A = m.add_species("A", D_surf, surface=True)
    B = m.add_species("B", 0.0 if static_b else D_surf * 0.5, surface=True)
    Cc = m.add_species("C", D_surf * 0.5, surface=True)
    D = m.add_species("D", D_surf, surface=True)
    E = m.add_species("E", D_surf * 0.25, surface=True)
    V = m.add_species("V", 1e-6)
    pb = m.config.time_step * m.config.surface_grid_density / 6.0
    m.add_reaction_rule(["A'", "B'"], ["C'"], p / pb)
    m.add_reaction_rule(["C'"], ["A'"], k_off)
    m.add_reaction_rule(["A'", "E'"], ["D'", "E'"], 0.5 * p / pb)
    m.add_reaction_rule(["D'", "D'"], ["A'", "B'"], 2 * p / pb)
    m.add_reaction_rule(["B'", "E,"], ["V,"], p / pb)
    sv, sf = create_icosphere(radius_um, subdivisions)
    m.add_geometry_object(sv, sf)
    bv, bf = create_box(box_um)
    m.add_geometry_object(bv, bf)
    n = n_a + n_b + n_e
    t = m.build(max_molecules=max_molecules or 2 * n + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    walls = np.arange(len(sf), dtype=np.uint32)
    # one draw of distinct tiles for all three species (release_on_walls keeps the tiles of one call distinct)
    allm = release_on_walls(rng, t, walls, n, A, orientation=1, first_id=0)
    # release_on_walls hands the molecules out wall by wall: mix the species over the sphere
    sp = np.full(n, A, dtype=allm.species.dtype)
    sp[n_a:n_a + n_b] = B
    sp[n_a + n_b:] = E
    allm.species[:] = rng.permutation(sp)
    return t, allm


def vacant_tile_products(n_r=900, n_a=600, n_lig=6000, radius_um=0.25, subdivisions=3, seed=1, box_um=0.8, D_surf=2e-7,
                         rng_mode=abi.MCX_RNG_PHILOX, max_molecules=None):
    """Surface products on vacant neighbour tiles (find_surf_product_positions' general branch):
    R' -> R' + A' (a kept reactant emits a surface molecule), L' + R' -> P' + A' (two surface products, one freed tile),
    P' -> A' + A' (a split: one product stays, one takes a neighbour tile), A' + A' -> P' + A' + W' would need ... no:
    A' + A' -> R' (keeps the population bounded), A' -> V, (leaves the surface)."""
    import math
    from mcell_b200.model import N_AV, MY_PI
    m = Model(Config(seed=seed))
    L = m.add_species("L", 1e-6)
    R = m.add_species("R", D_surf, surface=True)
    A = m.add_species("A", D_surf, surface=True)
    P = m.add_species("P", D_surf * 0.5, surface=True)
    V = m.add_species("V", 1e-6)
    pb_vs = 2.0 * 1.0e11 * m.config.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * m.config.time_step / 1e-6)
    pb_ss = m.config.time_step * m.config.surface_grid_density / 6.0
    m.add_reaction_rule(["R'"], ["R'", "A'"], 3e4)
    m.add_reaction_rule(["L'", "R'"], ["P'", "A'"], 0.5 / pb_vs)
    m.add_reaction_rule(["P'"], ["A'", "A'"], 1e5)
    m.add_reaction_rule(["A'", "A'"], ["R'"], 0.2 / pb_ss)
    m.add_reaction_rule(["A'"], ["V,"], 5e4)
    sv, sf = create_icosphere(radius_um, subdivisions)
    m.add_geometry_object(sv, sf)
    bv, bf = create_box(box_um)
    m.add_geometry_object(bv, bf)
    n = n_r + n_a
    t = m.build(max_molecules=max_molecules or 4 * (n + n_lig) + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n_lig, box_um, t.length_unit, margin=1e-3)
    vol = MolArrays.from_positions(pos, L, schedule_unimol=True)
    surf = release_on_walls(rng, t, np.arange(len(sf), dtype=np.uint32), n, R, orientation=1, first_id=n_lig)
    sp = np.full(n, R, dtype=surf.species.dtype)
    sp[n_r:] = A
    surf.species[:] = rng.permutation(sp)
    return t, MolArrays.concat([vol, surf])


def unsupported_surface_surface_tables():
    """Surface-surface pathways outside the supported set: (what, tables).  Both the oracle and libmcx must
    refuse them instead of approximating."""
    out = []
    for what, products in (("needs vacant neighbour tiles while it keeps one surface reactant and consumes the other", ["C'", "D'", "A'"]),
                           ("frees two tiles, fills one, next to a volume product", ["C'", "V,"]),
                           ("two surface products on the two freed tiles and a volume product", ["C'", "D'", "V,"])):
        m = Model(Config(seed=1))
        for n in ("A", "B", "C", "D"):
            m.add_species(n, 1e-7, surface=True)
        m.add_species("V", 1e-6)
        m.add_reaction_rule(["A'", "B'"], products, 1e4)
        sv, sf = create_icosphere(0.1, 1)
        m.add_geometry_object(sv, sf)
        out.append((what, m.build(max_molecules=16)))
    return out


def counted_spheres(n=12000, seed=1, box_um=0.8, rng_mode=abi.MCX_RNG_PHILOX, p_target=0.4, max_molecules=None):
    """Counted volumes: two nested transparent icospheres, both counted, inside a counted
    reflective box; A + B -> C everywhere.  Volumes: {box}, {box, outer}, {box, outer, inner} (+ the empty set)."""
    m = Model(Config(seed=seed))
    m.add_species("A", 1e-6)
    m.add_species("B", 1e-6)
    m.add_species("C", 0.5e-6)
    pb = _pb_factor(m, 0, 1)
    m.add_reaction_rule(["A", "B"], ["C"], p_target / pb)
    ov, of = create_icosphere(0.3, 3)
    iv, if_ = create_icosphere(0.15, 2)
    m.add_geometry_object(ov, of, surf_class=0, counted=True)
    m.add_geometry_object(iv + 0.02, if_, surf_class=0, counted=True)
    bv, bf = create_box(box_um)
    m.add_geometry_object(bv, bf, counted=True)
    m.add_surface_property(0, abi.MCX_SURF_TRANSPARENT, species=None)
    t = m.build(max_molecules=max_molecules or 2 * n + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n, box_um, t.length_unit, margin=1e-3)
    mols = MolArrays.from_positions(pos, (np.arange(n) % 2).astype(np.uint32))
    mols.counted_volume[:] = counted_volume_of(t, pos)
    return t, mols


def intersecting_counted_spheres(n=12000, n_rec=0, seed=1, box_um=0.8, rng_mode=abi.MCX_RNG_PHILOX, p_target=0.4, k_off=3e5):
    """Counted objects that intersect (the waypoint case of update_counted_volume_id_when_crossing_wall,
    collision_utils.inl:1568-1694): two transparent, counted icospheres of radius 0.2 um whose centres are 0.2 um
    apart, inside a counted reflective box; A + B -> C everywhere.  Volumes: {box}, {box, S0}, {box, S1},
    {box, S0, S1}.  n_rec > 0: receptors on sphere 0 bind A from both sides and release it again (L R -> A + R: a volume
    product of a unimolecular surface reaction, on a wall that lies partly inside sphere 1)."""
    import math
    from mcell_b200.model import N_AV, MY_PI
    m = Model(Config(seed=seed))
    m.add_species("A", 1e-6)
    m.add_species("B", 1e-6)
    m.add_species("C", 0.5e-6)
    pb = _pb_factor(m, 0, 1)
    m.add_reaction_rule(["A", "B"], ["C"], p_target / pb)
    if n_rec:
        R = m.add_species("R", 0.0, surface=True)
        m.add_species("AR", 0.0, surface=True)
        pbs = 1.0e11 * m.config.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * m.config.time_step / 1e-6)
        m.add_reaction_rule(["A", "R"], ["AR'"], 0.5 / pbs)
        m.add_reaction_rule(["AR'"], ["A,", "R'"], k_off)
    v0, f0 = create_icosphere(0.2, 3)
    v1, f1 = create_icosphere(0.2, 3)
    m.add_geometry_object(np.asarray(v0) + [-0.1, 0.0, 0.0], f0, surf_class=0, counted=True)
    m.add_geometry_object(np.asarray(v1) + [0.1, 0.013, 0.007], f1, surf_class=0, counted=True)
    bv, bf = create_box(box_um)
    m.add_geometry_object(bv, bf, counted=True)
    m.add_surface_property(0, abi.MCX_SURF_TRANSPARENT, species=None)
    t = m.build(max_molecules=2 * (n + n_rec) + 64, rng_mode=rng_mode)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n, box_um, t.length_unit, margin=1e-3)
    mols = MolArrays.from_positions(pos, (np.arange(n) % 2).astype(np.uint32), schedule_unimol=True)
    mols.counted_volume[:] = counted_volume_of(t, pos)
    if n_rec:
        surf = release_on_walls(rng, t, np.arange(len(f0), dtype=np.uint32), n_rec, R, orientation=1, first_id=n)
        mols = MolArrays.concat([mols, surf])
    return t, mols


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
