"""BASELINE.json configs 2, 3 and 4 at their full sizes on the B200 through the C ABI.  The oracle cannot run these
sizes in seconds, so the checks are the size-independent properties the domain offers: conservation of every
conserved moiety, containment, one surface molecule per tile, agreement of the incremental device counts with a
recount of the downloaded population, the mass-action rate, and membrane tightness."""
import math

import numpy as np
import pytest

import common as cm
from mcell_b200 import abi
from mcell_b200.model import (Model, Config, MolArrays, create_box, create_icosphere, release_uniform_box,
                              release_on_walls, N_AV, MY_PI)

pytestmark = pytest.mark.gpu


def _engine(t):
    from mcell_b200 import Engine
    return Engine(t)


def _species_counts(m, n_species):
    return np.bincount(m.species[:m.n], minlength=n_species)[:n_species]


def test_config2_bimolecular_box_1e6():
    """A + B -> C, 1e6 molecules in a 2 um box, default 0.5 um subpartitions (15.6k molecules each), counts every
    10 iterations."""
    n = 1_000_000
    t, mols = cm.reactive_box(n=n, edge_um=2.0, seed=21, p_target=0.1, cap_factor=1.25)
    e = _engine(t)
    e.upload(mols)
    a0 = b0 = n // 2
    rows = []
    for _ in range(3):                       # "counts every 10 iterations": one plugin call per barrier window
        st = e.step(10)
        c = e.counts()[0]
        rows.append(c.copy())
        assert c[0] + c[2] == a0 and c[1] + c[2] == b0
        assert st.unresolved_conflicts == 0
    out = e.download()
    assert (_species_counts(out, 3) == rows[-1][:3]).all()
    assert out.n == int(rows[-1][:3].sum())
    for k in ("x", "y", "z"):
        v = getattr(out, k)[:out.n]
        assert v.min() >= -100.0 and v.max() <= 100.0
    assert len(np.unique(out.id[:out.n])) == out.n
    # mass action: dN_A/dt = -k N_A N_B / (N_AV V); p_target fixes k
    m = Model(Config(seed=1))
    m.add_species("A", 1e-6); m.add_species("B", 1e-6)
    k = 0.1 / cm._pb_factor(m, 0, 1)
    vol_l = 8.0e-15
    kk = k / (N_AV * vol_l) * 1e-6           # per iteration, per (molecule of A) per (molecule of B)
    a = float(a0)
    for _ in range(30):
        a -= kk * a * a
    got = rows[-1][0]
    assert abs(got - a) < 0.03 * (a0 - a) + 5 * math.sqrt(a0 - a), (got, a)


def _config3(n_lig, n_rec, seed):
    """Ligand-receptor on an icosphere of 20 480 triangles (create_icosphere(0.5 um, 6)) whose top cap absorbs the
    ligand and whose bottom cap is transparent, inside a reflective box."""
    m = Model(Config(seed=seed))
    L = m.add_species("L", 1e-6)
    R = m.add_species("R", 0.0, surface=True)
    LR = m.add_species("LR", 0.0, surface=True)
    pb = 2.0 * 1.0e11 * m.config.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * m.config.time_step / 1e-6)
    m.add_reaction_rule(["L'", "R'"], ["LR'"], 0.5 / pb)
    m.add_reaction_rule(["LR'"], ["L'", "R'"], 1e5)
    sv, sf = create_icosphere(0.5, 6)
    assert len(sf) == 20480
    cls = np.full(len(sf), abi.MCX_NONE, np.uint32)
    cz = sv[sf].mean(axis=1)[:, 2]
    cls[cz > 0.35] = 0
    cls[cz < -0.35] = 1
    m.add_geometry_object(sv, sf, cls)
    bv, bf = create_box(1.6)
    m.add_geometry_object(bv, bf)
    m.add_surface_property(0, abi.MCX_SURF_ABSORPTIVE, species="L")
    m.add_surface_property(1, abi.MCX_SURF_TRANSPARENT, species=None)
    t = m.build(max_molecules=2 * (n_lig + n_rec) + 64)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n_lig, 1.6, t.length_unit, margin=1e-3)
    vol = MolArrays.from_positions(pos, L, schedule_unimol=True)
    belt = np.flatnonzero(cls == abi.MCX_NONE).astype(np.uint32)   # receptors on the reflective belt
    surf = release_on_walls(rng, t, belt, n_rec, R, orientation=1, first_id=n_lig)
    return t, MolArrays.concat([vol, surf]), (L, R, LR)


def test_config3_ligand_receptor_icosphere_20k_triangles():
    n_lig, n_rec = 400_000, 8_000
    t, mols, (L, R, LR) = _config3(n_lig, n_rec, seed=22)
    e = _engine(t)
    e.upload(mols)
    absorbed = bound = unbound = transparent = 0
    for _ in range(5):
        st = e.step(10)
        absorbed += st.mol_wall_absorptions
        bound += st.bimol_rxns
        unbound += st.unimol_rxns
        transparent += st.mol_wall_transparent
        c = e.counts()[0]
        assert c[R] + c[LR] == n_rec
        assert c[L] + c[LR] + absorbed == n_lig
        assert st.unresolved_conflicts == 0
    assert absorbed > 1000 and bound > 500 and unbound > 100 and transparent > 1000
    out = e.download()
    assert (_species_counts(out, 3) == e.counts()[0][:3]).all()
    surf = out.wall[:out.n] != abi.MCX_NONE
    assert surf.sum() == n_rec
    tiles = out.wall[:out.n][surf].astype(np.uint64) << np.uint64(32) | out.tile[:out.n][surf].astype(np.uint64)
    assert len(np.unique(tiles)) == n_rec                      # Grid::molecules_per_tile: one molecule per tile
    vol = ~surf
    for k in ("x", "y", "z"):
        v = getattr(out, k)[:out.n][vol]
        assert v.min() >= -80.0 and v.max() <= 80.0
    # surface molecules sit on the sphere: within the sagitta of a subdivisions=6 face of the 50 lu radius
    r = np.sqrt(out.x[:out.n][surf] ** 2 + out.y[:out.n][surf] ** 2 + out.z[:out.n][surf] ** 2)
    assert r.max() <= 50.0 + 1e-9 and r.min() > 50.0 * (1 - 2e-3)


def _config4(n_total, seed):
    """Synapse-like: two nested icospheres of 81 920 triangles each (163 840 + 12 walls), Ca + CB <-> CaCB in the
    volume, pumps on the outer membrane taking Ca from the inside and releasing it outside; the inner membrane is
    transparent for Ca and reflective for the buffer."""
    m = Model(Config(seed=seed))
    Ca = m.add_species("Ca", 2e-6)
    CB = m.add_species("CB", 0.3e-6)
    CaCB = m.add_species("CaCB", 0.3e-6)
    P = m.add_species("P", 0.0, surface=True)
    CaP = m.add_species("CaP", 0.0, surface=True)
    lu, ts = m.length_unit, m.config.time_step
    eff = (m.space_step(2e-6) + m.space_step(0.3e-6)) * lu / ts
    Rr = m.rxn_radius_um
    pb = 1.0 / (2.0 * math.sqrt(MY_PI) * Rr * Rr * eff) * 1.0e15 / N_AV
    m.add_reaction_rule(["Ca", "CB"], ["CaCB"], 0.2 / pb)
    m.add_reaction_rule(["CaCB"], ["Ca", "CB"], 2e4)
    pbs = 2.0 * 1.0e11 * m.config.surface_grid_density / (2.0 * N_AV) * math.sqrt(MY_PI * ts / 2e-6)
    m.add_reaction_rule(["Ca,", "P'"], ["CaP'"], 0.5 / pbs)
    m.add_reaction_rule(["CaP'"], ["P'", "Ca'"], 2e5)
    ov, of = create_icosphere(1.9, 7)
    iv, if_ = create_icosphere(0.9, 7)
    assert len(of) == 81920 and len(if_) == 81920
    m.add_geometry_object(ov, of)
    m.add_geometry_object(iv, if_, surf_class=0)
    bv, bf = create_box(4.0)
    m.add_geometry_object(bv, bf)
    m.add_surface_property(0, abi.MCX_SURF_TRANSPARENT, species="Ca")
    n_pump = 60_000
    n_vol = n_total - n_pump
    t = m.build(max_molecules=int(1.25 * n_total) + 64)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n_vol, 4.0, t.length_unit, margin=1e-3)
    sp = rng.integers(0, 10, n_vol)
    species = np.select([sp < 4, sp < 8], [Ca, CB], CaCB).astype(np.uint32)
    vol = MolArrays.from_positions(pos, species, schedule_unimol=True)
    surf = release_on_walls(rng, t, np.arange(len(of), dtype=np.uint32), n_pump, P, orientation=1, first_id=n_vol)
    return t, MolArrays.concat([vol, surf]), (Ca, CB, CaCB, P, CaP), n_pump


def test_config4_synapse_like_mesh_160k_triangles_1e7_molecules():
    n_total = 10_000_000
    t, mols, (Ca, CB, CaCB, P, CaP), n_pump = _config4(n_total, seed=23)
    c0 = _species_counts(mols, 5)
    r0 = np.sqrt(mols.x ** 2 + mols.y ** 2 + mols.z ** 2)
    buf0 = (mols.species == CB) | (mols.species == CaCB)
    e = _engine(t)
    e.upload(mols)
    tot = {"bimol_rxns": 0, "unimol_rxns": 0, "mol_wall_reflections": 0, "mol_wall_transparent": 0}
    for _ in range(2):
        st = e.step(10)
        for k in tot:
            tot[k] += getattr(st, k)
        c = e.counts()[0]
        assert c[Ca] + c[CaCB] + c[CaP] == c0[Ca] + c0[CaCB]          # calcium
        assert c[CB] + c[CaCB] == c0[CB] + c0[CaCB]                    # buffer
        assert c[P] + c[CaP] == n_pump                                 # pumps
        assert st.unresolved_conflicts == 0
    assert tot["bimol_rxns"] > 10000 and tot["unimol_rxns"] > 10000
    assert tot["mol_wall_reflections"] > 10000 and tot["mol_wall_transparent"] > 1000
    out = e.download()
    n = out.n
    assert (_species_counts(out, 5) == e.counts()[0][:5]).all()
    assert len(np.unique(out.id[:n])) == n
    surf = out.wall[:n] != abi.MCX_NONE
    assert surf.sum() == n_pump
    tiles = out.wall[:n][surf].astype(np.uint64) << np.uint64(32) | out.tile[:n][surf].astype(np.uint64)
    assert len(np.unique(tiles)) == n_pump
    for k in ("x", "y", "z"):
        v = getattr(out, k)[:n][~surf]
        assert v.min() >= -200.0 and v.max() <= 200.0
    # membrane tightness for the buffer (both membranes reflect CB / CaCB; the sagitta of a subdivisions=7 face is
    # < 2e-4 of the radius): nobody gets from certainly-on-one-side to certainly-on-the-other side
    r1 = np.sqrt(out.x[:n] ** 2 + out.y[:n] ** 2 + out.z[:n] ** 2)
    buf1 = ((out.species[:n] == CB) | (out.species[:n] == CaCB)) & ~surf
    for radius in (190.0, 90.0):
        inner, outer = radius * (1 - 4e-4), radius
        assert (buf1 & (r1 < inner)).sum() <= (buf0 & (r0 < outer)).sum()
        assert (buf1 & (r1 > outer)).sum() <= (buf0 & (r0 > inner)).sum()
        # and the populations on either side did not drain: within Poisson noise of where they started (CaCB and CB
        # interconvert but stay on their side)
        n_in0 = float((buf0 & (r0 < inner)).sum())
        n_in1 = float((buf1 & (r1 < inner)).sum())
        assert abs(n_in1 - n_in0) < 12 * math.sqrt(n_in0) + 4e-4 * 3 * n_in0, (radius, n_in0, n_in1)


# ---- oracle comparison AT FULL SIZE ------------------------------------------------------------------------------------
# The oracle cannot run a whole iteration of these configurations in seconds, but the evaluation of ONE molecule depends
# on the start-of-iteration snapshot only.  So the device runs the full population — a few iterations to
# get products, fractional first steps and scheduled lifetimes into the state, then one traced iteration — and the
# oracle evaluates a sample of the molecules (every stride-th id) against the same state (orc_trace_sample).  Every
# sampled molecule that the device evaluated once (no conflict retry, not consumed as somebody's partner) must have the
# oracle's trace: event hash, partners, walls, reaction, words drawn bit for bit, position to 1e-12.
def _compare_full_size_sample(t, mols, stride, offset, warm_iterations):
    from oracle import oracle_py as O
    e = _engine(t)
    e.upload(mols)
    if warm_iterations:
        e.step(warm_iterations)
    state = e.download()
    n_ids = int(state.id[:state.n].max()) + 1
    e2 = _engine(t)
    e2.upload(state)                       # common state: ids, positions, times as the oracle gets them
    o = O.Oracle(t)
    o.upload(state)
    tr_g, st_g = e2.trace_step(n_ids)
    tr_o = o.trace_sample(stride, offset, n_ids)
    sampled = np.flatnonzero(tr_o["rounds"] > 0)
    once = sampled[(tr_g["rounds"][sampled] == 1) & (tr_g["outcome"][sampled] != abi.MCX_OUT_CONSUMED)]
    assert len(once) > 0.9 * len(sampled) and len(sampled) > 10000
    bad = cm.compare_traces(tr_o, tr_g, once)
    assert not bad, bad[:5]
    return tr_o[once], st_g


def test_config2_full_size_sample_matches_oracle():
    t, mols = cm.reactive_box(n=1_000_000, edge_um=2.0, seed=21, p_target=0.1, cap_factor=1.25)
    tr, st = _compare_full_size_sample(t, mols, stride=32, offset=5, warm_iterations=3)
    assert (tr["n_collisions"] > 0).sum() > 2000 and (tr["rxn_class"] != abi.MCX_NONE).sum() > 100
    assert st.bimol_rxns > 5000


def test_config3_full_size_sample_matches_oracle():
    t, mols, _ = _config3(400_000, 8_000, seed=3)
    tr, st = _compare_full_size_sample(t, mols, stride=4, offset=1, warm_iterations=3)
    assert (tr["n_wall_hits"] > 0).sum() > 1000


def test_config4_full_size_sample_matches_oracle():
    t, mols, _, _ = _config4(10_000_000, seed=4)
    tr, st = _compare_full_size_sample(t, mols, stride=512, offset=7, warm_iterations=2)
    assert (tr["n_wall_hits"] > 0).sum() > 100 and (tr["n_collisions"] > 0).sum() > 500


def test_config6_surface_surface_1e6_whole_iterations_match_oracle():
    """Surface-surface reactions at the size bench.py --config 6 times (1e6 surface molecules on 20 480 triangles, 3.5e6
    tiles): the oracle finishes an iteration of this in seconds, so WHOLE iterations are compared — traces of all
    molecules (tiles taken, partners in list order, class, pathway, tile / orientation bits, conflict rounds), statistics,
    species and rule counts — and the populations at the end."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from oracle import oracle_py as O
    t, mols, _, _, _ = bench.build_small_config(6)
    n = mols.n
    e, o = _engine(t), O.Oracle(t)
    e.upload(mols)
    o.upload(mols)
    rx = 0
    for it in range(3):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad[:3])
        for k in ("bimol_rxns", "unimol_rxns", "resolve_retries", "unresolved_conflicts", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        co, cg = o.counts(), e.counts()
        assert (np.asarray(cg[0]) == np.asarray(co[0])).all() and (np.asarray(cg[1]) == np.asarray(co[1])).all(), it
        rx += st_g.bimol_rxns
    assert rx > 50000 and st_g.unresolved_conflicts == 0
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    assert a.n == b.n and (a.id == b.id).all() and (a.species == b.species).all()
    assert (a.wall == b.wall).all() and (a.tile == b.tile).all() and (a.orientation == b.orientation).all()
    assert cm.rel_close(a.u, b.u, 1e-12).all() and cm.rel_close(a.v, b.v, 1e-12).all()
    s = b.wall != abi.MCX_NONE
    assert len(np.unique(np.stack([b.wall[s], b.tile[s]], 1), axis=0)) == int(s.sum())
