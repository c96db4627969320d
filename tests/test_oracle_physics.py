"""CPU tier: the oracle (restatement of the reference algorithm) against analytic expectations and structural
identities (the reference tree holds no golden vectors for the geometry/collision/reaction path),
and the oracle's two execution modes against each other."""
import ctypes as C
import math

import numpy as np
import pytest

import common as cm
from mcell_b200 import abi
from mcell_b200.model import N_AV
from oracle import oracle_py as O


def test_wall_constants_of_a_box():
    t, _ = cm.free_diffusion_box(n=10)
    o = O.Oracle(t)
    for wi in range(12):
        c = o.wall_constants(wi)
        n, d, u, v = c[0:3], c[3], c[4:7], c[7:10]
        assert np.linalg.norm(n) == pytest.approx(1.0, abs=1e-14)
        assert sorted(np.abs(n)) == pytest.approx([0, 0, 1], abs=1e-14)     # axis-aligned faces
        assert abs(d) == pytest.approx(50.0, abs=1e-12)
        assert d > 0                                                         # outward normals (create_box order)
        assert abs(np.dot(u, v)) < 1e-14 and abs(np.dot(u, n)) < 1e-14
        assert np.allclose(np.cross(n, u), v, atol=1e-14)
        assert c[13] == pytest.approx(0.5 * 100 * 100)                       # area


def test_walls_per_subpart_of_a_box():
    """Config 1: the 1 um cube fills 2x2x2 default subpartitions; its faces lie exactly on subpartition
    boundaries, so with the leeway of wall_subparts_collision_test (geometry_utils.inl:110-207) the walls are also
    registered in the touching outer layer: 4x4x4 subpartitions hold walls."""
    t, _ = cm.free_diffusion_box(n=10)
    o = O.Oracle(t)
    n = t.cfg.num_subparts_per_edge
    occupied = {}
    for s in range(n ** 3):
        w = o.subpart_walls(s)
        if len(w):
            occupied[s] = w
    assert len(occupied) == 64
    for s, w in occupied.items():
        x, y, z = s % n, (s // n) % n, s // (n * n)
        assert 8 <= x <= 11 and 8 <= y <= 11 and 8 <= z <= 11
        if x in (9, 10) and y in (9, 10) and z in (9, 10):
            assert 3 <= len(w) <= 6                                          # the three faces of a corner octant
        assert (np.diff(w.astype(int)) > 0).all()                            # ascending (uint_set order)
    allw = np.unique(np.concatenate(list(occupied.values())))
    assert (allw == np.arange(12)).all()


def test_free_diffusion_msd_and_containment():
    n = 40000
    t, mols = cm.free_diffusion_box(n=n, seed=5)
    o = O.Oracle(t)
    o.upload(mols)
    before = mols.sorted_by_id()
    st = o.step(1, 0)
    assert st.molecule_steps == n
    after = o.download().sorted_by_id()
    d2 = (after.x - before.x) ** 2 + (after.y - before.y) ** 2 + (after.z - before.z) ** 2
    inner = (np.abs(before.x) < 30) & (np.abs(before.y) < 30) & (np.abs(before.z) < 30)
    expect = 1.5 * t.species[0].space_step ** 2                              # 3*space_step^2/2 = 6 D dt
    assert abs(d2[inner].mean() - expect) < 5 * expect * math.sqrt(2.0 / 3.0 / inner.sum())
    st = o.step(60, 0)
    assert st.mol_wall_reflections > 1000
    a = o.download()
    assert a.n == n
    for k in ("x", "y", "z"):
        v = getattr(a, k)
        assert v.min() >= -50 and v.max() <= 50
        h, _ = np.histogram(v, bins=10, range=(-50, 50))
        assert np.all(np.abs(h - n / 10) < 6 * math.sqrt(n / 10))


def test_snapshot_replay_of_the_sequential_tape_is_identical_without_reactions():
    """Without reactions a molecule's outcome depends only on its own random words: the sequential mode
    (ONE global ISAAC64 stream, reference semantics) and the snapshot mode replaying each molecule's slice
    of that stream must agree bit for bit."""
    n = 6000
    t, mols = cm.free_diffusion_box(n=n, rng_mode=abi.MCX_RNG_TAPE, seed=2)
    a, b = O.Oracle(t), O.Oracle(t)
    a.upload(mols); b.upload(mols)
    for _ in range(3):
        tr_a, st_a = a.trace_step(0, n)
        words, off, ln = a.tape(n)
        tr_b, st_b = b.trace_step(2, n, words, off)
        assert not cm.compare_traces(tr_a, tr_b, np.arange(n))
        assert (tr_b["n_words"] == ln).all()
        assert st_a.mol_wall_reflections == st_b.mol_wall_reflections
    x, y = a.download().sorted_by_id(), b.download().sorted_by_id()
    assert (x.x == y.x).all() and (x.y == y.y).all() and (x.z == y.z).all()


def _mass_action_expected(t, n_a, n_b, edge_um, k):
    v_litres = (edge_um * 1e-5) ** 3             # 1 um = 1e-5 dm; dm^3 = litre
    return k * n_a * n_b / (N_AV * v_litres) * t.time_unit


@pytest.mark.parametrize("mode", [0, 1])
def test_bimolecular_rate_matches_mass_action(mode):
    """A + B -> C in a well-mixed reflective box: reactions per iteration = k [A][B] V dt (SURVEY §8c golden (2))."""
    n, edge = 40000, 0.5
    t, mols = cm.reactive_box(n=n, edge_um=edge, p_target=0.05, seed=9)
    k = 0.05 / cm._pb_factor(_two_species_model(), 0, 1)
    o = O.Oracle(t)
    o.upload(mols)
    got = expect = 0.0
    for _ in range(6):
        c, _r = o.counts()
        expect += _mass_action_expected(t, float(c[0]), float(c[1]), edge, k)
        st = o.step(1, mode)
        got += st.bimol_rxns
    assert got > 1500
    assert abs(got - expect) < 4 * math.sqrt(expect) + 0.03 * expect, (got, expect)


def _two_species_model():
    from mcell_b200.model import Model, Config
    m = Model(Config())
    m.add_species("A", 1e-6); m.add_species("B", 1e-6)
    return m


@pytest.mark.parametrize("mode", [0, 1])
def test_unimolecular_decay_is_exponential(mode):
    from mcell_b200.model import Model, Config, MolArrays, create_box, release_uniform_box
    m = Model(Config(seed=4))
    m.add_species("C", 1e-6); m.add_species("A", 1e-6)
    k = 5e4
    m.add_reaction_rule(["C"], ["A"], k)
    v, f = create_box(0.5)
    m.add_geometry_object(v, f)
    n = 30000
    t = m.build(max_molecules=2 * n)
    pos = release_uniform_box(np.random.default_rng(1), n, 0.5, t.length_unit, margin=1e-4)
    mols = MolArrays.from_positions(pos, 0, schedule_unimol=True)
    o = O.Oracle(t)
    o.upload(mols)
    iters = 10
    o.step(iters, mode)
    c, r = o.counts()
    surv = n * math.exp(-k * t.time_unit * iters)
    assert abs(c[0] - surv) < 5 * math.sqrt(n * (surv / n) * (1 - surv / n))
    assert c[0] + c[1] == n and r[0] == c[1]


def test_snapshot_and_sequential_semantics_agree_statistically():
    """The parallel (snapshot + conflict rounds) semantics must not bias reaction counts against the
    reference's sequential semantics: 3 sigma over seeds (north_star ensemble criterion, reduced size)."""
    seq, snap = [], []
    for seed in range(1, 9):
        t, mols = cm.reactive_box(n=8000, edge_um=0.3, p_target=0.3, seed=seed)
        for mode, acc in ((0, seq), (1, snap)):
            o = O.Oracle(t)
            o.upload(mols)
            o.step(15, mode)
            acc.append(float(o.counts()[0][2]))
    seq, snap = np.array(seq), np.array(snap)
    se = math.sqrt(seq.var(ddof=1) / len(seq) + snap.var(ddof=1) / len(snap))
    assert abs(seq.mean() - snap.mean()) < 3 * se + 0.005 * seq.mean(), (seq.mean(), snap.mean(), se)


def test_surface_classes_conserve_and_act():
    n = 12000
    t, mols = cm.sphere_classes(n=n, seed=3)
    for mode in (0, 1):
        o = O.Oracle(t)
        o.upload(mols)
        tot_abs = tot_tr = 0
        for _ in range(12):
            st = o.step(1, mode)
            tot_abs += st.mol_wall_absorptions
            tot_tr += st.mol_wall_transparent
        c, _ = o.counts()
        assert tot_abs > 30 and tot_tr > 60
        assert c[0] == n // 2 - tot_abs       # only L is absorbed
        assert c[1] == n // 2


def test_conflict_rounds_resolve_everything_at_default_depth():
    t, mols = cm.reactive_box(n=8000, edge_um=0.3, p_target=0.5, seed=3)
    o = O.Oracle(t)
    o.upload(mols)
    st = o.step(3, 1)
    assert st.bimol_rxns > 500 and st.resolve_retries > 0 and st.unresolved_conflicts == 0
    c, r = o.counts()
    assert c[0] + c[2] == 4000 and c[1] + c[2] == 4000 and r[0] == c[2]


# ---- surface molecules (collide_and_react_with_surf_mol, surface unimolecular) ---------------
def _surf_state(o, t):
    m = o.download()
    return m, m.wall != 0xFFFFFFFF


def _inside_sphere(t, m, n_faces=320):
    """Exact inside test for the (convex) icosphere made of the first n_faces walls."""
    v = t.vertices[t.tri[:n_faces]]
    nrm = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    nrm *= np.sign((nrm * v.mean(1)).sum(1))[:, None]     # outward
    p = np.c_[m.x, m.y, m.z]
    d = (p[:, None, :] - v[None, :, 0, :]) * nrm[None]
    return (d.sum(2) < 0).all(1)


@pytest.mark.parametrize("mode", [0, 1])
def test_surface_binding_conserves_tiles_and_respects_orientation(mode):
    """L' + R' -> LR' only from the front (outside), Ca, + P' -> CaP' only from the back (inside)."""
    t, mols = cm.ligand_receptor_sphere(n_lig=16000, n_rec=3000, n_pump=1500, seed=2, release_products=False,
                                        k_off=0.0, k_pump=0.0)
    inside = _inside_sphere(t, mols)
    n_L_in = int(((mols.species == 0) & inside).sum())
    n_Ca_out = int(((mols.species == 1) & ~inside).sum())
    o = O.Oracle(t)
    o.upload(mols)
    tot = 0
    for _ in range(25):
        tot += o.step(1, mode).bimol_rxns
    c, r = o.counts()
    assert tot > 150 and r[0] + r[2] == tot
    assert c[2] + c[3] == 3000 and c[4] + c[5] == 1500          # receptors / pumps are conserved on their tiles
    assert c[0] + c[3] == 8000 and c[1] + c[5] == 8000          # ligand / calcium only move into complexes
    m, is_surf = _surf_state(o, t)
    assert is_surf.sum() == 4500
    tiles = np.stack([m.wall[is_surf], m.tile[is_surf]], 1)
    assert len(np.unique(tiles, axis=0)) == 4500                   # one molecule per tile
    # ligands bind from the outside only, calcium from the inside only: the populations on the other side are intact
    assert c[3] > 50 and c[5] > 20
    ins = _inside_sphere(t, m)
    assert int(((m.species == 0) & ins).sum()) == n_L_in
    assert int(((m.species == 1) & ~ins).sum()) == n_Ca_out


@pytest.mark.parametrize("mode", [0, 1])
def test_surface_unbinding_places_products_on_the_named_side(mode):
    """LR' -> L' + R': L appears just outside the wall (2*16*EPS along the normal) and R keeps the tile;
    CaP' -> P' + Ca': calcium that was taken from the inside is released outside."""
    t, mols = cm.ligand_receptor_sphere(n_lig=16000, n_rec=3000, n_pump=1500, seed=3, k_off=3e5, k_pump=3e5)
    ins0 = _inside_sphere(t, mols)
    ca_in0 = int(((mols.species == 1) & ins0).sum())
    l_in0 = int(((mols.species == 0) & ins0).sum())
    o = O.Oracle(t)
    o.upload(mols)
    uni = 0
    for _ in range(30):
        uni += o.step(1, mode).unimol_rxns
    c, r = o.counts()
    assert uni > 60 and r[1] + r[3] == uni
    assert c[2] + c[3] == 3000 and c[4] + c[5] == 1500
    assert c[0] + c[3] == 8000 and c[1] + c[5] == 8000
    m, is_surf = _surf_state(o, t)
    ins = _inside_sphere(t, m)
    ca_in = int(((m.species == 1) & ins).sum())
    # every calcium taken by a pump came from the inside and none was released there
    assert ca_in == ca_in0 - int(r[2]), (ca_in, ca_in0, r[2], r[3])
    # unbound ligands reappear outside
    assert int(((m.species == 0) & ins).sum()) == l_in0


def test_surface_snapshot_and_sequential_agree_statistically():
    seq, snap = [], []
    for seed in range(1, 9):
        t, mols = cm.ligand_receptor_sphere(n_lig=8000, n_rec=2500, n_pump=1000, seed=seed)
        for mode, acc in ((0, seq), (1, snap)):
            o = O.Oracle(t)
            o.upload(mols)
            o.step(25, mode)
            acc.append([float(x) for x in o.counts()[1]])
    seq, snap = np.array(seq), np.array(snap)
    for k in range(4):
        se = math.sqrt(seq[:, k].var(ddof=1) / 8 + snap[:, k].var(ddof=1) / 8)
        assert abs(seq[:, k].mean() - snap[:, k].mean()) < 3 * se + 0.02 * seq[:, k].mean() + 1, (k, seq[:, k].mean(), snap[:, k].mean())


# ---- counted volumes ---------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1])
def test_counted_volume_index_follows_the_geometry(mode):
    """update_counted_volume_id_when_crossing_wall: after many transparent crossings of two nested counted
    spheres the index every molecule carries equals the one a fresh ray cast gives for its position; products
    inherit it; per-volume molecule and reaction counts add up to the world counts."""
    from mcell_b200.model import counted_volume_of
    t, mols = cm.counted_spheres(n=12000, seed=2)
    assert t.n_counted_volumes == 4 and sorted(len(s) for s in t.counted_volume_sets) == [0, 1, 2, 3]
    o = O.Oracle(t)
    o.upload(mols)
    tr = 0
    for _ in range(15):
        tr += o.step(1, mode).mol_wall_transparent
    assert tr > 1500
    m = o.download()
    want = counted_volume_of(t, np.c_[m.x, m.y, m.z])
    assert (m.counted_volume == want).all(), int((m.counted_volume != want).sum())
    mc, rc = o.counts_by_volume()
    c, r = o.counts()
    assert (mc.sum(1) == c).all() and (rc.sum(1) == r).all() and r[0] > 300
    assert (mc[:, 1:] > 0).all() and (rc[0, 1:] > 0).all() and mc[:, 0].sum() == 0
    for cv in range(t.n_counted_volumes):
        assert (np.bincount(m.species[m.counted_volume == cv], minlength=3) == mc[:, cv]).all()


# ---- surface diffusion --------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1])
def test_surface_diffusion_msd_tiles_and_edges(mode):
    """diffuse_surf_molecule / ray_trace_surf / traverse_surface: molecules stay on the mesh, keep one molecule per
    tile, cross triangle edges, and at low occupancy their mean square displacement per step is space_step^2
    (pick_surf_displacement: E|d|^2 = scale^2), measured over a few steps where the sphere is still locally flat."""
    t, mols = cm.diffusing_receptors(n_rec=400, n_lig=10, D_surf=2e-8, with_ligand=False, seed=5)
    o = O.Oracle(t)
    o.upload(mols)
    a = o.download().sorted_by_id()
    steps = 6
    o.step(steps, mode)
    b = o.download().sorted_by_id()
    s = a.wall != 0xFFFFFFFF
    assert (b.wall != 0xFFFFFFFF).sum() == 400 and (a.id == b.id).all()
    d2 = ((b.x - a.x) ** 2 + (b.y - a.y) ** 2 + (b.z - a.z) ** 2)[s]
    want = steps * t.species[1].space_step ** 2
    assert abs(d2.mean() - want) < 0.12 * want, (d2.mean(), want)
    assert ((b.wall != a.wall) & s).sum() > 40                                 # edges were crossed
    tiles = np.stack([b.wall[s], b.tile[s]], 1)
    assert len(np.unique(tiles, axis=0)) == 400
    # every molecule lies on its triangle: uv2xyz(u, v) is its position and the point is inside the wall
    tri = t.vertices[t.tri[b.wall[s]]]
    e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    p = np.c_[b.x, b.y, b.z][s] - tri[:, 0]
    n = np.cross(e1, e2)
    assert np.abs((p * n).sum(1) / np.linalg.norm(n, axis=1)).max() < 1e-9
    den = (e1 * e1).sum(1) * (e2 * e2).sum(1) - (e1 * e2).sum(1) ** 2
    bu = ((p * e1).sum(1) * (e2 * e2).sum(1) - (p * e2).sum(1) * (e1 * e2).sum(1)) / den
    bv = ((p * e2).sum(1) * (e1 * e1).sum(1) - (p * e1).sum(1) * (e1 * e2).sum(1)) / den
    assert (bu > -1e-9).all() and (bv > -1e-9).all() and (bu + bv < 1 + 1e-9).all()


def test_surface_diffusion_crowded_with_binding_agrees_between_semantics():
    """Crowded surface (38 % of the tiles taken) with ligand binding: sequential (reference) and snapshot semantics
    agree statistically on bound receptors; tiles stay exclusive in both."""
    seq, snap = [], []
    for seed in range(1, 7):
        t, mols = cm.diffusing_receptors(n_rec=3000, n_lig=8000, seed=seed)
        for mode, acc in ((0, seq), (1, snap)):
            o = O.Oracle(t)
            o.upload(mols)
            st = o.step(20, mode)
            m = o.download()
            s = m.wall != 0xFFFFFFFF
            assert len(np.unique(np.stack([m.wall[s], m.tile[s]], 1), axis=0)) == 3000
            acc.append(float(o.counts()[1][0]))
            if mode == 1:
                assert st.unresolved_conflicts == 0
    seq, snap = np.array(seq), np.array(snap)
    se = math.sqrt(seq.var(ddof=1) / len(seq) + snap.var(ddof=1) / len(snap))
    assert abs(seq.mean() - snap.mean()) < 3 * se + 0.03 * seq.mean(), (seq.mean(), snap.mean(), se)


def _inside(t, m, species):
    """(molecules of a volume species inside the sphere of transporter_sphere by a fresh ray cast, all of them)"""
    sel = (m.species[:m.n] == species) & (m.wall[:m.n] == abi.MCX_NONE)
    pos = np.stack([m.x[:m.n], m.y[:m.n], m.z[:m.n]], 1)[sel]
    inner = [i for i, s_ in enumerate(t.counted_volume_sets) if 0 in s_][0]
    return int((cm.counted_volume_of(t, pos) == inner).sum()), int(sel.sum())


@pytest.mark.parametrize("mode", [0, 1])
def test_kept_volume_reactants_of_surface_reactions(mode):
    """SURVEY 8 a17 (diffuse_react_event.cpp:945-975, 2689-2716): A' + T' -> A, + T' takes A through the wall (RX_FLIP:
    the molecule goes on behind the wall and its counted volume switches), S' + E' -> S' + E' + Pr' keeps both
    reactants (S reflects) and releases Pr in front.  The sphere is reflective otherwise, so A only ever enters, S and
    Pr never do; mode 0 = reference semantics, 1 = snapshot semantics."""
    t, mols = cm.transporter_sphere(n_vol=6000, n_trans=2500, n_enz=1500, seed=11)
    A, S, Pr, T, E = 0, 1, 2, 3, 4
    o = O.Oracle(t)
    o.upload(mols)
    a_in0, a_tot = _inside(t, mols, A)
    s_in0, s_tot = _inside(t, mols, S)
    prev = a_in0
    for it in range(8):
        o.step(2, mode)
        m = o.download()
        a_in, a_n = _inside(t, m, A)
        s_in, s_n = _inside(t, m, S)
        pr_in, pr_n = _inside(t, m, Pr)
        assert a_n == a_tot and s_n == s_tot                      # kept reactants are never consumed
        assert a_in >= prev and s_in == s_in0 and pr_in == 0       # one-way transport; S and Pr stay outside
        prev = a_in
        sp, rx = o.counts()
        assert sp[T] == 2500 and sp[E] == 1500 and sp[Pr] == rx[1] == pr_n
        assert rx[0] == a_in - a_in0                               # every transport event moved one A inside
        # counted volume index of every volume molecule == a fresh ray cast
        vol = m.wall[:m.n] == abi.MCX_NONE
        pos = np.stack([m.x[:m.n], m.y[:m.n], m.z[:m.n]], 1)[vol]
        assert (m.counted_volume[:m.n][vol] == cm.counted_volume_of(t, pos)).all()
    assert prev - a_in0 > 25 and rx[1] > 25, (prev - a_in0, rx)


def test_kept_volume_reactants_snapshot_matches_reference_semantics():
    """transport and turnover counts of the two execution modes agree statistically"""
    res = []
    for mode in (0, 1):
        tot = np.zeros(2)
        for seed in range(3):
            t, mols = cm.transporter_sphere(n_vol=8000, n_trans=2500, n_enz=1500, seed=20 + seed)
            o = O.Oracle(t)
            o.upload(mols)
            o.step(16, mode)
            tot += o.counts()[1][:2]
        res.append(tot)
    for k in range(2):
        a, b = res[0][k], res[1][k]
        assert abs(a - b) < 5 * math.sqrt(a + b), (k, res)


@pytest.mark.parametrize("mode", [0, 1])
def test_surface_region_counts(mode):
    """SURVEY 8 f2 (mol_or_rxn_count_event.cpp:528-534, 588-600; diffuse_react_event.cpp:2513-2521): surface molecules per
    set of counted regions == a count over the downloaded walls; reactions initiated by surface molecules (the
    unimolecular LR -> R, CaP -> P) are counted on their wall, reactions initiated by volume molecules in their volume."""
    from mcell_b200.engine import region_count
    t, mols = cm.ligand_receptor_sphere(n_lig=8000, n_rec=2500, n_pump=1200, seed=9, k_off=3e5, k_pump=4e5,
                                        release_products=False, regions=True)
    assert t.n_region_sets == 4 and set(t.region_sets) == {frozenset(), frozenset({0}), frozenset({1}), frozenset({0, 1})}
    o = O.Oracle(t)
    o.upload(mols)
    for it in range(4):
        o.step(5, mode)
        m = o.download()
        mc, rc = o.counts_by_surface_region()
        surf = m.wall[:m.n] != abi.MCX_NONE
        ref = np.zeros_like(mc)
        np.add.at(ref, (m.species[:m.n][surf], t.wall_region_set[m.wall[:m.n][surf]]), 1)
        assert (mc == ref).all()
        sp, rx = o.counts()
        assert (mc.sum(axis=1)[2:] == sp[2:]).all() and (mc[:2] == 0).all()       # every surface molecule is on some set
        assert (rc.sum(axis=1)[[1, 3]] == rx[[1, 3]]).all() and (rc[[0, 2]] == 0).all()
        north = region_count(t, mc, 0)
        band = region_count(t, mc, 1)
        both = mc[:, t.region_sets.index(frozenset({0, 1}))]
        assert (north + band - both <= sp).all()
    assert rx[1] > 20 and rx[3] > 10


@pytest.mark.parametrize("mode", [0, 1])
def test_finite_rate_surface_class_reactions(mode):
    """SURVEY 8 a18 (diffuse_react_event.cpp:991-1067, 1916-1988; rxn_utils.inl:593-626): permeation through a reactive
    surface in both directions, a consuming and a catalytic wall reaction.  Conservation, the side every product appears
    on, counted volumes against fresh ray casts, reaction counts against population changes."""
    t, mols = cm.permeable_sphere(n=9000, seed=13)
    A, B, Cc, D, E, F = range(6)
    inner = [i for i, s_ in enumerate(t.counted_volume_sets) if 0 in s_][0]
    o = O.Oracle(t)
    o.upload(mols)
    n0 = np.bincount(mols.species[:mols.n], minlength=6)
    for it in range(6):
        o.step(3, mode)
        m = o.download()
        sp, rx = o.counts()
        pos = np.stack([m.x[:m.n], m.y[:m.n], m.z[:m.n]], 1)
        cv = cm.counted_volume_of(t, pos)
        assert (m.counted_volume[:m.n] == cv).all()
        assert sp[A] == n0[A] and sp[E] == n0[E]                                  # kept reactants
        assert sp[B] + sp[Cc] == n0[B] and sp[Cc] == sp[D] == rx[2] and sp[F] == rx[3]
        s_ = m.species[:m.n]
        assert (cv[s_ == Cc] != inner).all() and (cv[s_ == D] == inner).all()      # C outside, D inside
        assert (cv[s_ == F] == inner).all() and (cv[s_ == E] != inner).sum() + (cv[s_ == E] == inner).sum() == n0[E]
        a_in = int((cv[s_ == A] == inner).sum())
    a_in0 = int((cm.counted_volume_of(t, np.stack([mols.x, mols.y, mols.z], 1)[mols.species == A]) == inner).sum())
    assert rx[0] - rx[1] == a_in - a_in0                                            # net inward crossings
    assert rx[0] > 60 and rx[1] > 10 and rx[2] > 60 and rx[3] > 40, rx
    e_in0 = int((cm.counted_volume_of(t, np.stack([mols.x, mols.y, mols.z], 1)[mols.species == E]) == inner).sum())
    assert int((cv[s_ == E] == inner).sum()) == e_in0                               # E never crosses


def test_surface_class_permeation_ratio_and_mode_agreement():
    """Hits from outside cross with p_in, hits from inside with p_out: the crossing / hit ratios of both directions
    follow the probabilities, and the two execution modes agree statistically."""
    res = []
    for mode in (0, 1):
        tot = np.zeros(2)
        for seed in range(3):
            t, mols = cm.permeable_sphere(n=9000, seed=30 + seed, products=False)
            o = O.Oracle(t)
            o.upload(mols)
            o.step(14, mode)
            tot += o.counts()[1][:2]
        res.append(tot)
    for k in range(2):
        a, b = res[0][k], res[1][k]
        assert abs(a - b) < 5 * math.sqrt(a + b), (k, res)
    assert res[0][0] > 200


@pytest.mark.parametrize("mode", [0, 1])
def test_region_borders_for_surface_molecules(mode):
    """SURVEY 8 a22, region borders (ray_trace_surf :1627-1665; reflect_absorb_inside_out / outside_in, diffusion_utils.inl:
    598-700): a REFLECTIVE class keeps surface molecules on their side of the outline of a reactive region, in both
    directions; an ABSORPTIVE one (absorptive region border) takes whoever reaches it; without the class they mix."""
    def run(border):
        t, mols = cm.diffusing_receptors(n_rec=3000, n_lig=0, seed=21, D_surf=4e-7, with_ligand=False, border=border)
        cap = None
        if border is not None:
            cap = np.flatnonzero(t.wall_region_set == t.region_sets.index(frozenset({0})))
            assert 0 < len(cap) < len(t.wall_region_set) and (t.wall_edge_border[cap] != 0).any()
            assert (t.wall_edge_border[np.setdiff1d(np.arange(len(t.wall_region_set)), cap)] == 0).all()
        o = O.Oracle(t)
        o.upload(mols)
        a = mols.sorted_by_id()
        o.step(40, mode)
        b = o.download().sorted_by_id()
        return t, cap, a, b
    # control: some receptors cross the outline
    t, _, a, b = run(None)
    t2, cap, _, _ = run(abi.MCX_SURF_REFLECTIVE)
    in0, in1 = np.isin(a.wall, cap), np.isin(b.wall, cap)
    assert b.n == a.n and (in0 != in1).sum() > 20
    # reflective border: nobody crosses, in either direction
    _, cap, a, b = run(abi.MCX_SURF_REFLECTIVE)
    assert b.n == a.n and (b.id == a.id).all()
    in0, in1 = np.isin(a.wall, cap), np.isin(b.wall, cap)
    assert (in0 == in1).all() and in0.sum() > 100 and (~in0).sum() > 100
    assert ((a.wall != b.wall) | (a.tile != b.tile)).sum() > 1000      # they do move
    # absorptive border: those that reach it are gone, the others stay on their side
    _, cap, a, b = run(abi.MCX_SURF_ABSORPTIVE)
    assert 20 < a.n - b.n < a.n // 2
    keep = np.isin(a.id, b.id)
    assert (np.isin(a.wall[keep], cap) == np.isin(b.wall, cap)).all()


@pytest.mark.parametrize("mode", [0, 1])
def test_intersecting_counted_objects(mode):
    """SURVEY 8 a20, counted objects that intersect (the waypoint case of update_counted_volume_id_when_crossing_wall,
    collision_utils.inl:1568-1694): crossing a wall of such an object toggles the object in the molecule's set of
    enclosing objects; volume products of unimolecular surface reactions on such walls get theirs from a ray cast at
    their first evaluation (MCX_MOL_CVI_PENDING until then).  Invariant: the counted volume of every volume molecule
    equals an independent ray cast in numpy, in all four volumes."""
    t, mols = cm.intersecting_counted_spheres(n=8000, n_rec=1500, seed=17)
    assert t.cv_intersecting == 0b011 and t.n_counted_volumes == 5
    o = O.Oracle(t)
    o.upload(mols)
    seen, crossings, pending_seen = set(), 0, 0
    for it in range(10):
        st = o.step(2, mode)
        crossings += st.mol_wall_transparent
        m = o.download()
        vol = m.wall[:m.n] == abi.MCX_NONE
        pend = (m.flags[:m.n] & abi.MCX_MOL_CVI_PENDING) != 0
        pending_seen += int(pend.sum())
        pos = np.stack([m.x[:m.n], m.y[:m.n], m.z[:m.n]], 1)
        ok = vol & ~pend
        cv = cm.counted_volume_of(t, pos[ok])
        assert (m.counted_volume[:m.n][ok] == cv).all(), it
        seen |= set(np.unique(cv).tolist())
        assert not (pend & ~vol).any()
    assert seen == {1, 2, 3, 4} and crossings > 2000
    sp, rx = o.counts()
    assert rx[2] > 30                                 # unbinding on the straddling walls happened ...
    assert pending_seen > 0 or mode == 0              # ... and was resolved later (the sequential mode evaluates products at once)
    mo, _ = o.counts_by_volume()
    assert mo[:3].sum() == sp[:3].sum()


@pytest.mark.parametrize("mode", [0, 1])
def test_counted_volume_computed_for_uploaded_molecules(mode):
    """Partition::add_volume_molecule computes the counted volume of a molecule that comes without one (partition.h:572-576);
    here such molecules are uploaded with MCX_MOL_CVI_PENDING and get it from a ray cast when they are first evaluated."""
    t, mols = cm.counted_spheres(n=6000, seed=23)
    truth = mols.counted_volume.copy()
    mols.counted_volume[:] = 0
    mols.flags[:] |= abi.MCX_MOL_CVI_PENDING
    assert (truth != 0).sum() > 4000 and len(np.unique(truth)) == 3
    o = O.Oracle(t)
    o.upload(mols)
    o.step(1, mode)
    m = o.download()
    assert not (m.flags[:m.n] & abi.MCX_MOL_CVI_PENDING).any()
    pos = np.stack([m.x[:m.n], m.y[:m.n], m.z[:m.n]], 1)
    assert (m.counted_volume[:m.n] == cm.counted_volume_of(t, pos)).all()


def _resume_case():
    return cm.reversible_box(n=6000, edge_um=0.4, seed=31)


def _resume_case_surface_surface():
    # few molecules per wall, so that walls gain their grids while the run goes on: Wall::has_initialized_grid is state
    return cm.surface_reactions(n_a=220, n_b=220, n_e=60, radius_um=0.25, subdivisions=3, seed=33, D_surf=6e-7)


def _run_resume(make_engine, total=8, stop=3, case=None, restore_wall_grids=True):
    """run `total` iterations at once, and `stop` iterations + download + a NEW engine at initial_iteration = stop +
    upload + the rest: both ends of the checkpoint must give the same population"""
    _resume_case = case or globals()["_resume_case"]
    t, mols = _resume_case()
    a = make_engine(t)
    a.upload(mols)
    for _ in range(total):
        a.step(1) if not hasattr(a, "SNAPSHOT") else a.step(1, 1)
    ref = a.download().sorted_by_id()
    ref_counts = a.counts()
    b = make_engine(t)
    b.upload(mols)
    for _ in range(stop):
        b.step(1) if not hasattr(b, "SNAPSHOT") else b.step(1, 1)
    saved, saved_counts, saved_next = b.download(), b.counts(), b.next_molecule_id()
    saved_grids = b.wall_grids()
    t2, _ = _resume_case()
    t2.cfg.initial_iteration = stop
    c = make_engine(t2)
    c.upload(saved)
    assert c.next_molecule_id(set_to=saved_next) == saved_next     # ids of molecules that are gone are not handed out again
    if restore_wall_grids:
        assert (c.wall_grids(set_to=saved_grids) == saved_grids).all()   # walls that held a molecule once keep their grid
    for _ in range(total - stop):
        c.step(1) if not hasattr(c, "SNAPSHOT") else c.step(1, 1)
    got = c.download().sorted_by_id()
    return ref, ref_counts, got, c.counts(), saved_counts


def test_checkpoint_resume_is_exact():
    """SURVEY 5.4 (checkpoint / resume): the per-molecule Philox streams are keyed by (seed, molecule id, iteration), so a
    run resumed from a downloaded population at Config.initial_iteration continues bit for bit; reaction counts of the
    two legs add up (the host adds initial_reactions_count like MolOrRxnCountTerm does)."""
    ref, ref_counts, got, counts, saved_counts = _run_resume(lambda t: O.Oracle(t))
    assert ref.n == got.n and (ref.id == got.id).all() and (ref.species == got.species).all()
    for k in ("x", "y", "z", "diffusion_time", "unimol_rxn_time"):
        assert (getattr(ref, k) == getattr(got, k)).all(), k
    assert (ref.flags == got.flags).all()
    assert (ref_counts[0] == counts[0]).all()
    assert (ref_counts[1] == counts[1] + saved_counts[1]).all() and ref_counts[1].sum() > 50


def test_checkpoint_resume_with_surface_surface_reactions_needs_the_wall_grids():
    """Wall::has_initialized_grid is part of the state of a model with surface-surface reactions (the neighbour search
    skips walls without a grid, which also sets the probability factor): restored with the population the resumed run is
    exact; the saved flags are a superset of the walls the saved population sits on."""
    ref, ref_counts, got, counts, saved_counts = _run_resume(lambda t: O.Oracle(t), total=10, stop=5, case=_resume_case_surface_surface)
    assert ref.n == got.n and (ref.id == got.id).all() and (ref.species == got.species).all()
    for k in ("x", "y", "z", "diffusion_time", "unimol_rxn_time", "wall", "tile", "u", "v"):
        assert (getattr(ref, k) == getattr(got, k)).all(), k
    assert (ref_counts[1] == counts[1] + saved_counts[1]).all() and ref_counts[1].sum() > 20
    # without the flags the resumed run is a different one
    _, _, lost, _, _ = _run_resume(lambda t: O.Oracle(t), total=10, stop=5, case=_resume_case_surface_surface, restore_wall_grids=False)
    assert lost.n != ref.n or not ((lost.wall == ref.wall).all() and (lost.tile == ref.tile).all() and (lost.species == ref.species).all())
    t, mols = _resume_case_surface_surface()
    o = O.Oracle(t)
    o.upload(mols)
    g0 = o.wall_grids()
    o.step(5, 1)
    g1 = o.wall_grids()
    d = o.download()
    on = np.zeros(len(t.tri), bool); on[d.wall[d.wall != 0xFFFFFFFF]] = True
    assert (g1 >= g0).all() and g1.sum() > g0.sum() and (g1[on] == 1).all() and (g1.astype(bool) & ~on).sum() > 0


# ---- surface-surface reactions -------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("static_b", [False, True])
def test_surface_surface_reactions_bookkeeping(mode, static_b):
    """react_2D_all_neighbors in both semantics: every rule moves the species counts the way it says, one molecule per
    tile throughout, every surface molecule lies on the tile its uv names, products of A' + B' -> C' and D' + D' -> A' + B'
    sit on tiles the reactants freed (no tile is gained or lost: surface molecules = tiles in use), the class whose
    orientations cannot match never fires, and a species that cannot diffuse still takes part (static_b)."""
    t, mols = cm.surface_reactions(seed=6, static_b=static_b)
    o = O.Oracle(t)
    o.upload(mols)
    for _ in range(4):
        o.step(10, mode)
        sp, rule = (np.asarray(a, dtype=np.int64) for a in o.counts())
        assert sp[1] == 1500 - rule[0] + rule[3] and sp[2] == rule[0] - rule[1] and sp[3] == rule[2] - 2 * rule[3] and sp[4] == 300
        assert sp[0] == 1500 - rule[0] + rule[1] - rule[2] + rule[3] and sp[5] == 0 and rule[4] == 0
        d = o.download()
        s = d.wall != 0xFFFFFFFF
        assert s.sum() == sp[:5].sum()
        assert len(np.unique(np.stack([d.wall[s], d.tile[s]], 1), axis=0)) == s.sum()
        L = load_library_for_grid()
        for i in np.flatnonzero(s)[::37]:
            v9 = np.ascontiguousarray(t.vertices[t.tri[d.wall[i]]].reshape(9))
            xyz = np.array([d.x[i], d.y[i], d.z[i]])
            assert L.mcx_xyz2grid(v9.ctypes.data_as(C.c_void_p), xyz.ctypes.data_as(C.c_void_p)) == d.tile[i]
    assert rule[0] > 150 and rule[2] > 40 and rule[3] > 5, rule


def load_library_for_grid():
    from mcell_b200.engine import load_library
    L = load_library()
    L.mcx_xyz2grid.restype = C.c_uint32
    L.mcx_xyz2grid.argtypes = [C.c_void_p, C.c_void_p]
    return L


def test_surface_surface_rate_is_mass_action_in_both_semantics():
    """A' + B' -> C' alone at low coverage: each A tests the B's on its ~12 neighbour tiles with the local probability
    factor 3 / (number of neighbour tiles) and vice versa, which adds up to the 2-D mass-action rate k [A][B] per area:
    reactions per step = k dt n_A n_B / area.  Sequential (reference) and snapshot semantics both give it."""
    from mcell_b200.model import Model, Config, create_icosphere, create_box, release_on_walls
    k_rate, steps = None, 20
    got = {0: 0, 1: 0}
    want = 0.0
    for seed in range(1, 7):
        m = Model(Config(seed=seed))
        A = m.add_species("A", 2e-7, surface=True)
        m.add_species("B", 1e-7, surface=True)
        m.add_species("C", 1e-7, surface=True)
        pb = m.config.time_step * m.config.surface_grid_density / 6.0
        k_rate = 0.03 / pb
        m.add_reaction_rule(["A'", "B'"], ["C'"], k_rate)
        sv, sf = create_icosphere(0.25, 3)
        m.add_geometry_object(sv, sf)
        bv, bf = create_box(0.8)
        m.add_geometry_object(bv, bf)
        n_a = n_b = 800
        t = m.build(max_molecules=4 * n_a)
        rng = np.random.default_rng(seed)
        mols = release_on_walls(rng, t, np.arange(len(sf), dtype=np.uint32), n_a + n_b, A, orientation=1, first_id=0)
        mols.species[:] = rng.permutation(np.r_[np.zeros(n_a, np.uint32), np.ones(n_b, np.uint32)]).astype(mols.species.dtype)
        tri = np.asarray(sv)[np.asarray(sf)]
        area_um2 = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1).sum()
        for mode in (0, 1):
            o = O.Oracle(t)
            o.upload(mols)
            na, nb = float(n_a), float(n_b)
            exp = 0.0
            for _ in range(steps):
                o.step(1, mode)
                sp, rule = o.counts()
                exp += k_rate * m.config.time_step * na * nb / area_um2
                na, nb = float(sp[0]), float(sp[1])
            got[mode] += int(rule[0])
            if mode == 0:
                want += exp
    for mode in (0, 1):
        assert abs(got[mode] - want) < 4 * np.sqrt(want) + 0.06 * want, (mode, got, want)
    assert abs(got[0] - got[1]) < 4 * np.sqrt(got[0] + got[1])


def test_oracle_refuses_unsupported_surface_surface_pathways():
    for what, t in cm.unsupported_surface_surface_tables():
        with pytest.raises(RuntimeError):
            O.Oracle(t)


# ---- surface products on vacant neighbour tiles ---------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1])
def test_products_on_vacant_neighbour_tiles_bookkeeping(mode):
    """find_surf_product_positions' general branch in both semantics: R' -> R' + A' (a kept reactant emits a surface
    molecule), L' + R' -> P' + A' (two surface products for one freed tile), P' -> A' + A' (a split), next to A' + A' -> R'
    and A' -> V,: every rule moves the counts the way it says, tiles stay exclusive, every surface molecule lies on the
    tile its position names, and a new molecule sits on a tile next to the one its reaction happened on."""
    t, mols = cm.vacant_tile_products(seed=5)
    o = O.Oracle(t)
    o.upload(mols)
    L = load_library_for_grid()
    for _ in range(4):
        o.step(5, mode)
        sp, r = (np.asarray(a, dtype=np.int64) for a in o.counts())
        assert sp[0] == 6000 - r[1] and sp[1] == 900 - r[1] + r[3] and sp[3] == r[1] - r[2] and sp[4] == r[4]
        assert sp[2] == 600 + r[0] + r[1] + 2 * r[2] - 2 * r[3] - r[4]
        d = o.download()
        s = d.wall != 0xFFFFFFFF
        assert s.sum() == sp[1:4].sum()
        assert len(np.unique(np.stack([d.wall[s], d.tile[s]], 1), axis=0)) == s.sum()
        for i in np.flatnonzero(s)[::23]:
            v9 = np.ascontiguousarray(t.vertices[t.tri[d.wall[i]]].reshape(9))
            xyz = np.array([d.x[i], d.y[i], d.z[i]])
            assert L.mcx_xyz2grid(v9.ctypes.data_as(C.c_void_p), xyz.ctypes.data_as(C.c_void_p)) == d.tile[i]
    assert r[0] > 200 and r[1] > 30 and r[2] > 15 and r[3] > 100 and r[4] > 100, r


def test_products_on_vacant_neighbour_tiles_agree_between_semantics():
    """Sequential (reference) and snapshot semantics on the same model: per-rule reaction counts after 12 iterations over 8
    seeds within 3 sigma (+ 10 % for the unimolecular rules, whose newborn reactants draw their lifetime one iteration late
    in the snapshot semantics)."""
    seq, snap = [], []
    for seed in range(1, 9):
        t, mols = cm.vacant_tile_products(seed=seed)
        for mode, acc in ((0, seq), (1, snap)):
            o = O.Oracle(t)
            o.upload(mols)
            o.step(12, mode)
            acc.append([float(x) for x in o.counts()[1][:5]])
    seq, snap = np.array(seq), np.array(snap)
    for r, floor in ((0, 0.05), (1, 0.0), (2, 0.10), (3, 0.05), (4, 0.10)):
        se = math.sqrt(seq[:, r].var(ddof=1) / 8 + snap[:, r].var(ddof=1) / 8)
        assert abs(seq[:, r].mean() - snap[:, r].mean()) <= 3 * se + floor * seq[:, r].mean(), (r, seq[:, r].mean(), snap[:, r].mean(), se)
