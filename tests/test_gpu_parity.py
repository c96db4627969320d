"""GPU parity tests (run on the B200 box): libmcx through its C ABI against the CPU oracle on the same
seeded inputs.  Bit-exact on event sequences / partner lists / reaction choices (integers), 1e-12 relative
on positions and event times (fp64), as BASELINE.json's north_star states."""
import numpy as np
import pytest

import common as cm
from mcell_b200 import abi

pytestmark = pytest.mark.gpu

POS_TOL = 1e-12  # relative, north_star


def _engine(t):
    from mcell_b200 import Engine
    return Engine(t)


def _oracle(t):
    from oracle import oracle_py as O
    return O.Oracle(t)


def _assert_same_population(a, b):
    """a, b: MolArrays sorted by id."""
    assert a.n == b.n
    assert (a.id == b.id).all()
    assert (a.species == b.species).all()
    for k in ("x", "y", "z"):
        assert cm.rel_close(getattr(a, k), getattr(b, k), POS_TOL).all(), k
    assert (a.flags == b.flags).all()
    assert cm.rel_close(a.diffusion_time, b.diffusion_time, POS_TOL).all()
    assert cm.rel_close(a.unimol_rxn_time, b.unimol_rxn_time, POS_TOL).all()
    # surface part (Molecule::s): wall, tile, orientation exact; uv to 1e-12
    assert (a.counted_volume == b.counted_volume).all()
    assert (a.wall == b.wall).all() and (a.tile == b.tile).all() and (a.orientation == b.orientation).all()
    assert cm.rel_close(a.u, b.u, POS_TOL).all() and cm.rel_close(a.v, b.v, POS_TOL).all()


def test_replay_reference_stream_free_diffusion():
    """The sequential oracle draws from ONE global ISAAC64 stream (reference semantics); the GPU replays
    each molecule's slice of that stream and must reproduce wall-hit sequences and positions."""
    n = 20000
    t, mols = cm.free_diffusion_box(n=n, rng_mode=abi.MCX_RNG_TAPE)
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    for it in range(3):
        tr_o, st_o = o.trace_step(0, n)
        words, off, ln = o.tape(n)
        tr_g, st_g = e.replay_step(words, off)
        ids = np.arange(n)
        bad = cm.compare_traces(tr_o, tr_g, ids)
        assert not bad, bad
        assert (tr_g["n_words"] == ln).all()
        assert st_g.mol_wall_reflections == st_o.mol_wall_reflections
        assert st_g.ray_polygon_tests == st_o.ray_polygon_tests
        assert st_g.molecule_steps == n
    assert st_o.mol_wall_reflections > 100
    _assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())


@pytest.mark.parametrize("subpart_um,cell_edge", [(0.5, 0.0), (0.05, 0.0), (0.5, 1.3), (0.05, 7.0)])
def test_replay_reactive_box_snapshot(subpart_um, cell_edge):
    """A + B -> C: per-molecule ISAAC64 tape slices; partner lists, reaction choices, conflict rounds."""
    n = 16000
    t, mols = cm.reactive_box(n=n, edge_um=0.4, p_target=0.5, rng_mode=abi.MCX_RNG_TAPE,
                              subpartition_dimension=subpart_um, cell_edge=cell_edge)
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    n_ids = n
    total_rxn = 0
    for it in range(4):
        words, off = cm.isaac_slices(100 + it, n_ids, 48)
        tr_o, st_o = o.trace_step(2, n_ids, words, off)
        tr_g, st_g = e.replay_step(words, off)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all()
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, bad
        for k in ("bimol_rxns", "vol_mol_vol_mol_collisions", "resolve_retries", "unresolved_conflicts",
                  "products_created", "mol_wall_reflections", "molecule_steps", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), k
        total_rxn += st_g.bimol_rxns
        assert (e.counts()[0] == o.counts()[0]).all()
        assert (e.counts()[1] == o.counts()[1]).all()
    assert total_rxn > 300
    _assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())


def test_philox_multi_iteration_reactive():
    """Production RNG path: per-molecule Philox streams, 12 iterations with products carried over."""
    n = 16000
    t, mols = cm.reactive_box(n=n, edge_um=0.4, p_target=0.4, seed=7)
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    for it in range(12):
        st_o = o.step(1, 1)
        st_g = e.step(1)
        assert st_g.bimol_rxns == st_o.bimol_rxns, it
        assert st_g.vol_mol_vol_mol_collisions == st_o.vol_mol_vol_mol_collisions, it
        assert st_g.molecule_steps == st_o.molecule_steps, it
    assert (e.counts()[0] == o.counts()[0]).all()
    _assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())


def test_philox_surface_classes_icosphere():
    """Absorptive / transparent / reflective faces of an icosphere (config 3 without receptors)."""
    n = 20000
    t, mols = cm.sphere_classes(n=n, seed=3)
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    tot = {"mol_wall_absorptions": 0, "mol_wall_transparent": 0, "mol_wall_reflections": 0}
    for it in range(10):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all()
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in tot:
            assert getattr(st_g, k) == getattr(st_o, k), k
            tot[k] += getattr(st_g, k)
    assert tot["mol_wall_absorptions"] > 50 and tot["mol_wall_transparent"] > 100 and tot["mol_wall_reflections"] > 1000
    assert (e.counts()[0] == o.counts()[0]).all()
    _assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())


def test_philox_ligand_receptor_surface_molecules():
    """BASELINE configs 3/4 surface chemistry: volume ligands binding receptors on the tiles of an icosphere
    (collide_and_react_with_surf_mol, orientation classes, tile recycling), pumps taking calcium from the inside,
    unimolecular unbinding of surface complexes.  Single-product pathways keep ids deterministic: traces, counts
    and the whole population (incl. wall / tile / orientation / uv) are compared over many iterations."""
    t, mols = cm.ligand_receptor_sphere(n_lig=24000, n_rec=3000, n_pump=1500, seed=4, release_products=False)
    n = mols.n
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    tot_bi = tot_uni = 0
    for it in range(12):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "unimol_rxns", "products_created", "mol_wall_reflections", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        tot_bi += st_g.bimol_rxns
        tot_uni += st_g.unimol_rxns
        assert (e.counts()[0] == o.counts()[0]).all(), it
    assert tot_bi > 100 and tot_uni > 20, (tot_bi, tot_uni)
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    _assert_same_population(a, b)
    assert (a.wall != abi.MCX_NONE).sum() == 4500      # every receptor / pump still owns its tile
    c = e.counts()[0]
    assert c[2] + c[3] == 3000 and c[4] + c[5] == 1500


def test_philox_surface_unbinding_two_products():
    """LR' -> L' + R' and CaP' -> P' + Ca': the volume product is placed 2*16*EPS off the wall on the side its
    orientation names, remembers where it was created (no immediate rebinding) and the surface product recycles the
    tile.  Second products take fresh ids from device atomics, so one iteration is compared as a multiset and the
    following ones statistically."""
    t, mols = cm.ligand_receptor_sphere(n_lig=24000, n_rec=3000, n_pump=1500, seed=6, k_off=4e5, k_pump=6e5)
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    # run until some complexes exist, comparing statistically; then one exact iteration from a common state
    for it in range(6):
        o.step(1, 1)
        e.step(1)
    co, cg = o.counts()[0].astype(float), e.counts()[0].astype(float)
    assert np.all(np.abs(co - cg) < 6 * np.sqrt(co + 1)), (co, cg)
    state = o.download()
    o2, e2 = _oracle(t), _engine(t)
    o2.upload(state)
    e2.upload(state)
    for _ in range(3):
        st_o = o2.step(1, 1)
        st_g = e2.step(1)
        for k in ("bimol_rxns", "unimol_rxns", "products_created", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), k
        if st_o.unimol_rxns:
            break
    assert st_o.unimol_rxns > 0
    a, b = o2.download(), e2.download()
    assert a.n == b.n

    def key(m):
        arr = np.c_[m.species.astype(float), m.x, m.y, m.z, m.wall.astype(float), m.tile.astype(float), m.orientation]
        return arr[np.lexsort(arr.T[::-1])]

    ka, kb = key(a), key(b)
    assert (ka[:, 0] == kb[:, 0]).all() and (ka[:, 4:] == kb[:, 4:]).all()
    assert cm.rel_close(ka[:, 1:4], kb[:, 1:4], POS_TOL).all()


def test_philox_transporter_flips_volume_reactant_through_the_wall():
    """SURVEY 8 a17, RX_FLIP (diffuse_react_event.cpp:945-970, 2694-2716): A' + T' -> A, + T' keeps both reactants and
    takes A through the wall of a counted sphere.  No new molecules, so ids stay deterministic: traces, statistics,
    per-volume counts and the whole population (incl. the counted volume of every molecule and the rebinding guard of
    the kept initiator, which shows in the next iteration's trace) are compared over many iterations."""
    t, mols = cm.transporter_sphere(n_vol=24000, n_trans=3500, n_enz=500, seed=5, enzyme=False)
    n = mols.n
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    flips = 0
    for it in range(14):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "products_created", "mol_wall_reflections", "resolve_retries", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        flips += st_g.bimol_rxns
        assert st_g.products_created == 0
        mo, ro = o.counts_by_volume()
        mg, rg = e.counts_by_volume()
        assert (mo == mg).all() and (ro == rg).all(), it
    assert flips > 150, flips
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    _assert_same_population(a, b)
    assert (a.counted_volume == b.counted_volume).all()
    vol = b.wall == abi.MCX_NONE
    pos = np.stack([b.x, b.y, b.z], 1)[vol]
    assert (b.counted_volume[vol] == cm.counted_volume_of(t, pos)).all()   # the index follows the molecule through the wall
    assert e.counts()[1][0] == flips


def test_philox_enzyme_keeps_volume_and_surface_reactant():
    """S' + E' -> S' + E' + Pr' next to the transporter rule: both reactants are kept (the volume
    reactant waits in front of the wall for the rest of its step), the product takes a fresh id — so every
    iteration starts from a common state and is compared bit for bit (traces, statistics, counts), the populations as
    multisets."""
    t, mols = cm.transporter_sphere(n_vol=24000, n_trans=2500, n_enz=2500, seed=6)
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)

    def key(m):
        arr = np.c_[m.species[:m.n].astype(float), m.x[:m.n], m.y[:m.n], m.z[:m.n], m.diffusion_time[:m.n],
                    m.flags[:m.n].astype(float), m.counted_volume[:m.n].astype(float), m.wall[:m.n].astype(float),
                    m.tile[:m.n].astype(float)]
        return arr[np.lexsort(arr.T[::-1])]

    made = 0
    state = mols
    for it in range(8):
        n_ids = int(state.id[:state.n].max()) + 1
        if it:
            e.upload(state)
            o.upload(state)
        tr_o, st_o = o.trace_step(1, n_ids)
        tr_g, st_g = e.trace_step(n_ids)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "products_created", "mol_wall_reflections", "resolve_retries", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        made += st_g.products_created
        assert (o.counts()[0] == e.counts()[0]).all(), it
        a, b = o.download(), e.download()
        assert a.n == b.n
        ka, kb = key(a), key(b)
        assert (ka[:, 0] == kb[:, 0]).all() and (ka[:, 4:] == kb[:, 4:]).all(), it
        assert cm.rel_close(ka[:, 1:4], kb[:, 1:4], POS_TOL).all(), it
        state = a
    assert made > 100, made
    c = e.counts()[0]
    assert c[0] + c[1] == 24000 and c[3] == 2500 and c[4] == 2500


def test_philox_surface_region_counts():
    """surface molecules per set of counted surface regions and the reactions initiated by surface
    molecules there (mcx_counts_by_surface_region) against the oracle, every iteration."""
    t, mols = cm.ligand_receptor_sphere(n_lig=24000, n_rec=3000, n_pump=1500, seed=12, k_off=3e5, k_pump=4e5,
                                        release_products=False, regions=True)
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)
    for it in range(10):
        o.step(1, 1)
        e.step(1)
        mo, ro = o.counts_by_surface_region()
        mg, rg = e.counts_by_surface_region()
        assert (mo == mg).all() and (ro == rg).all(), it
        assert (o.counts()[1] == e.counts()[1]).all(), it
    assert rg.sum() > 30 and (rg.sum(axis=1)[[1, 3]] == e.counts()[1][[1, 3]]).all()
    assert mg.sum() == 4500


def test_philox_surface_class_permeation():
    """SURVEY 8 a18, finite-rate reactions with a surface class (collide_and_react_with_walls -> test_intersect ->
    outcome_intersect): A crosses a reactive sphere in both directions with different probabilities (kept reactant,
    RX_FLIP).  Ids stay deterministic: traces, statistics, per-volume counts and the whole population over many
    iterations."""
    t, mols = cm.permeable_sphere(n=30000, seed=7, products=False)
    n = mols.n
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)
    crossed = 0
    for it in range(14):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "products_created", "mol_wall_reflections", "resolve_retries", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        crossed += st_g.bimol_rxns
        mo, ro = o.counts_by_volume()
        mg, rg = e.counts_by_volume()
        assert (mo == mg).all() and (ro == rg).all(), it
        assert (o.counts()[1] == e.counts()[1]).all(), it
    assert crossed > 300 and int((tr_g["outcome"][live] == abi.MCX_OUT_WALLRXN).sum()) > 5, crossed
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    _assert_same_population(a, b)
    pos = np.stack([b.x, b.y, b.z], 1)
    assert (b.counted_volume == cm.counted_volume_of(t, pos)).all()


def test_philox_surface_class_reactions_with_products():
    """... and the consuming (B' @ sc -> C' + D,) and catalytic (E' @ sc -> E' + F,) wall reactions, whose products take
    fresh ids: every iteration starts from a common state; traces, statistics and counts bit for bit, populations as
    multisets."""
    t, mols = cm.permeable_sphere(n=30000, seed=8)
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)

    def key(m):
        arr = np.c_[m.species[:m.n].astype(float), m.x[:m.n], m.y[:m.n], m.z[:m.n], m.diffusion_time[:m.n],
                    m.flags[:m.n].astype(float), m.counted_volume[:m.n].astype(float)]
        return arr[np.lexsort(arr.T[::-1])]

    made = 0
    state = mols
    for it in range(8):
        n_ids = int(state.id[:state.n].max()) + 1
        if it:
            e.upload(state)
            o.upload(state)
        tr_o, st_o = o.trace_step(1, n_ids)
        tr_g, st_g = e.trace_step(n_ids)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "products_created", "mol_wall_reflections", "resolve_retries", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        made += st_g.products_created
        assert (o.counts()[0] == e.counts()[0]).all(), it
        a, b = o.download(), e.download()
        assert a.n == b.n
        ka, kb = key(a), key(b)
        assert (ka[:, 0] == kb[:, 0]).all() and (ka[:, 4:] == kb[:, 4:]).all(), it
        assert cm.rel_close(ka[:, 1:4], kb[:, 1:4], POS_TOL).all(), it
        state = a
    assert made > 200, made


def test_philox_surface_diffusion_with_binding():
    """Surface diffusion (diffuse_surf_molecule, ray_trace_surf across triangle edges, tile claims between movers)
    together with ligand binding on the moving receptors: traces (incl. the tile every mover takes), conflict
    rounds, counts and the whole population over many iterations."""
    t, mols = cm.diffusing_receptors(n_rec=3000, n_lig=16000, seed=8)
    n = mols.n
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    moved = retries = rx = 0
    for it in range(12):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "unimol_rxns", "resolve_retries", "unresolved_conflicts", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        moved += int((tr_g["outcome"][live] == abi.MCX_OUT_SURFMOVE).sum())
        retries += st_g.resolve_retries
        rx += st_g.bimol_rxns
        assert (e.counts()[0] == o.counts()[0]).all(), it
    assert moved > 5000 and retries > 20 and rx > 50, (moved, retries, rx)
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    _assert_same_population(a, b)
    s = b.wall != abi.MCX_NONE
    assert len(np.unique(np.stack([b.wall[s], b.tile[s]], 1), axis=0)) == 3000


@pytest.mark.parametrize("border", [abi.MCX_SURF_REFLECTIVE, abi.MCX_SURF_ABSORPTIVE])
def test_philox_region_borders_for_surface_molecules(border):
    """SURVEY 8 a22, region borders: receptors diffusing on a sphere with a reactive region whose outline reflects or
    absorbs them, ligands binding on both sides: traces (every reflection changes the tile a mover takes, every
    absorption ends its trace), conflict rounds, counts and the whole population over many iterations."""
    t, mols = cm.diffusing_receptors(n_rec=3000, n_lig=12000, seed=9, D_surf=4e-7, border=border)
    n = mols.n
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)
    absorbed = moved = 0
    for it in range(12):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "unimol_rxns", "mol_wall_absorptions", "resolve_retries", "unresolved_conflicts", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        absorbed += st_g.mol_wall_absorptions
        moved += int((tr_g["outcome"][live] == abi.MCX_OUT_SURFMOVE).sum())
        assert (e.counts()[0] == o.counts()[0]).all(), it
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    _assert_same_population(a, b)
    assert moved > 5000
    if border == abi.MCX_SURF_ABSORPTIVE:
        assert absorbed > 20 and e.counts()[0][1] + e.counts()[0][2] == 3000 - absorbed
    else:
        assert absorbed == 0 and e.counts()[0][1] + e.counts()[0][2] == 3000


def test_philox_intersecting_counted_objects():
    """SURVEY 8 a20, counted objects that intersect: membership toggling on crossings, the lazy ray cast of surface-born
    volume products (MCX_MOL_CVI_PENDING) — every iteration from a common state (the second product of the unbinding
    takes a fresh id): traces, statistics, per-volume counts bit for bit, populations as multisets incl. the counted
    volume and the pending flag; and the device's counted volumes against an independent ray cast in numpy."""
    t, mols = cm.intersecting_counted_spheres(n=24000, n_rec=3000, seed=19)
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)

    def key(m):
        arr = np.c_[m.species[:m.n].astype(float), m.x[:m.n], m.y[:m.n], m.z[:m.n], m.diffusion_time[:m.n],
                    m.flags[:m.n].astype(float), m.counted_volume[:m.n].astype(float), m.wall[:m.n].astype(float),
                    m.tile[:m.n].astype(float)]
        return arr[np.lexsort(arr.T[::-1])]

    crossings = pending = 0
    state = mols
    for it in range(10):
        n_ids = int(state.id[:state.n].max()) + 1
        if it:
            e.upload(state)
            o.upload(state)
        tr_o, st_o = o.trace_step(1, n_ids)
        tr_g, st_g = e.trace_step(n_ids)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "unimol_rxns", "products_created", "mol_wall_transparent", "resolve_retries", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        crossings += st_g.mol_wall_transparent
        mo, ro = o.counts_by_volume()
        mg, rg = e.counts_by_volume()
        assert (mo == mg).all(), it
        a, b = o.download(), e.download()
        assert a.n == b.n
        ka, kb = key(a), key(b)
        assert (ka[:, 0] == kb[:, 0]).all() and (ka[:, 4:] == kb[:, 4:]).all(), it
        assert cm.rel_close(ka[:, 1:4], kb[:, 1:4], POS_TOL).all(), it
        vol = b.wall[:b.n] == abi.MCX_NONE
        pend = (b.flags[:b.n] & abi.MCX_MOL_CVI_PENDING) != 0
        pending += int(pend.sum())
        ok = vol & ~pend
        pos = np.stack([b.x[:b.n], b.y[:b.n], b.z[:b.n]], 1)[ok]
        assert (b.counted_volume[:b.n][ok] == cm.counted_volume_of(t, pos)).all(), it
        state = a
    assert crossings > 3000 and pending > 10, (crossings, pending)


def test_counted_volume_computed_for_uploaded_molecules():
    """Molecules uploaded with MCX_MOL_CVI_PENDING (the host does not know their counted volume): the device ray-casts it at
    their first evaluation; traces and populations against the oracle, counted volumes against numpy."""
    t, mols = cm.counted_spheres(n=16000, seed=23)
    mols.counted_volume[:] = 0
    mols.flags[:] |= abi.MCX_MOL_CVI_PENDING
    n = mols.n
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)
    for it in range(3):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        mo, ro = o.counts_by_volume()
        mg, rg = e.counts_by_volume()
        assert (mo == mg).all() and (ro == rg).all(), it
    b = e.download()
    assert not (b.flags[:b.n] & abi.MCX_MOL_CVI_PENDING).any()
    pos = np.stack([b.x[:b.n], b.y[:b.n], b.z[:b.n]], 1)
    assert (b.counted_volume[:b.n] == cm.counted_volume_of(t, pos)).all()


def test_checkpoint_resume_is_exact_on_the_device():
    """download at iteration k, a new handle with Config.initial_iteration = k, upload, restore
    next_molecule_id, go on — bit for bit the run that never stopped (streams are keyed by seed, molecule id, iteration)."""
    import test_oracle_physics as top
    ref, ref_counts, got, counts, saved_counts = top._run_resume(lambda t: _engine(t))
    _assert_same_population(ref, got)
    assert (ref_counts[0] == counts[0]).all()
    assert (ref_counts[1] == counts[1] + saved_counts[1]).all() and ref_counts[1].sum() > 50


def test_checkpoint_resume_with_surface_surface_reactions_on_the_device():
    """The same with surface-surface reactions: Wall::has_initialized_grid travels with the checkpoint
    (mcx_get_wall_grids / mcx_set_wall_grids); the device's flags equal the oracle's."""
    import test_oracle_physics as top
    ref, ref_counts, got, counts, saved_counts = top._run_resume(lambda t: _engine(t), total=10, stop=5,
                                                                 case=top._resume_case_surface_surface)
    _assert_same_population(ref, got)
    assert (ref_counts[1] == counts[1] + saved_counts[1]).all() and ref_counts[1].sum() > 20
    t, mols = top._resume_case_surface_surface()
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)
    assert (o.wall_grids() == e.wall_grids()).all()
    o.step(5, 1)
    e.step(5)
    assert (o.wall_grids() == e.wall_grids()).all() and e.wall_grids().sum() > 0


def test_more_than_256_species_and_rules():
    """The device counters hold 1024 species and 1024 reaction rules (the reference has no such limit; round 1 stopped at
    256): a chain of 600 species with one unimolecular rule each, populations and per-rule counts against the oracle."""
    from mcell_b200.model import Model, Config, MolArrays, create_box, release_uniform_box
    ns, n = 600, 30000
    m = Model(Config(seed=5))
    for k in range(ns):
        m.add_species("S%d" % k, 1e-6)
    for k in range(ns):
        m.add_reaction_rule(["S%d" % k], ["S%d" % ((k + 1) % ns)], 2e5 + 500.0 * k)
    bv, bf = create_box(0.5)
    m.add_geometry_object(bv, bf)
    t = m.build(max_molecules=2 * n + 64)
    rng = np.random.default_rng(5)
    pos = release_uniform_box(rng, n, 0.5, t.length_unit, margin=1e-3)
    mols = MolArrays.from_positions(pos, (np.arange(n) % ns).astype(np.uint32), schedule_unimol=True)
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)
    for it in range(6):
        st_o, st_g = o.step(1, 1), e.step(1)
        assert st_g.unimol_rxns == st_o.unimol_rxns, it
    co, cg = o.counts(), e.counts()
    assert len(cg[0]) == ns and len(cg[1]) == ns
    assert (co[0] == cg[0]).all() and (co[1] == cg[1]).all()
    assert cg[1][300:].sum() > 500 and cg[0].sum() == n
    _assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())


def test_philox_counted_volumes_nested_spheres():
    """Counted volumes: the index switches on transparent crossings of two nested counted spheres, products inherit
    it, and per-volume molecule / reaction counts (MolOrRxnCountEvent terms restricted to a volume) match."""
    t, mols = cm.counted_spheres(n=16000, seed=3)
    n = mols.n
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    crossings = 0
    for it in range(10):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        assert st_g.mol_wall_transparent == st_o.mol_wall_transparent and st_g.bimol_rxns == st_o.bimol_rxns
        crossings += st_g.mol_wall_transparent
        mo, ro = o.counts_by_volume()
        mg, rg = e.counts_by_volume()
        assert (mo == mg).all() and (ro == rg).all(), it
    assert crossings > 1000 and ro[0, 1:].min() > 0
    _assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())


def test_philox_reversible_binding_unimolecular():
    """Ca + CB <-> CaCB: unimolecular lifetimes, split steps, two-product unimolecular firing.
    Fresh ids of second products are allocated by atomics on the device, so populations are compared
    as (species, position) multisets rather than by id."""
    n = 9000
    t, mols = cm.reversible_box(n=n, seed=5)
    o = _oracle(t)
    o.upload(mols)
    e = _engine(t)
    e.upload(mols)
    for it in range(1):
        st_o = o.step(1, 1)
        st_g = e.step(1)
        for k in ("bimol_rxns", "unimol_rxns", "products_created", "molecule_steps", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
    a, b = o.download(), e.download()
    assert a.n == b.n

    def key(m):
        arr = np.c_[m.species.astype(float), m.x, m.y, m.z, m.diffusion_time, m.unimol_rxn_time]
        return arr[np.lexsort(arr.T[::-1])]

    ka, kb = key(a), key(b)
    assert (ka[:, 0] == kb[:, 0]).all()
    assert cm.rel_close(ka[:, 1:], kb[:, 1:], POS_TOL).all()
    # statistical agreement over more iterations (ids diverge after the first fresh allocation)
    for it in range(30):
        o.step(1, 1)
        e.step(1)
    co, cg = o.counts()[0].astype(float), e.counts()[0].astype(float)
    assert np.all(np.abs(co - cg) < 6 * np.sqrt(co + 1)), (co, cg)


def test_full_size_properties_free_diffusion():
    """BASELINE config 1 at full size (1e5 molecules, 1 um cube): size-independent properties —
    conservation, containment, uniform density, mean-square displacement 3*space_step^2/2 per step
    for molecules that did not touch a wall."""
    n = 100000
    t, mols = cm.free_diffusion_box(n=n, seed=11)
    e = _engine(t)
    e.upload(mols)
    before = mols.sorted_by_id()
    st = e.step(1)
    assert st.molecule_steps == n and st.n_live == n
    after = e.download().sorted_by_id()
    assert (after.id == before.id).all()
    d2 = (after.x - before.x) ** 2 + (after.y - before.y) ** 2 + (after.z - before.z) ** 2
    inner = (np.abs(before.x) < 30) & (np.abs(before.y) < 30) & (np.abs(before.z) < 30)
    msd = d2[inner].mean()
    expect = 1.5 * t.species[0].space_step ** 2
    assert abs(msd - expect) < 5 * expect * np.sqrt(2.0 / 3.0 / inner.sum()) * 1.5, (msd, expect)
    st = e.step(199)
    assert st.molecule_steps == 199 * n
    after = e.download().sorted_by_id()
    assert after.n == n
    for k in ("x", "y", "z"):
        v = getattr(after, k)
        assert v.min() >= -50.0 and v.max() <= 50.0
        hist, _ = np.histogram(v, bins=10, range=(-50, 50))
        assert np.all(np.abs(hist - n / 10) < 6 * np.sqrt(n / 10))


def test_errors_are_reported_not_fatal():
    from mcell_b200 import McxError
    t, mols = cm.free_diffusion_box(n=100)
    e = _engine(t)
    with pytest.raises(McxError):
        e.step(1)  # nothing uploaded
    mols.x[0] = 1e6  # outside the partition
    with pytest.raises(McxError) as ei:
        e.upload(mols)
    assert ei.value.code == abi.MCX_ERR_ESCAPED


def test_benchmark_chemistry_matches_oracle():
    """The chemistry bench.py times (config 5: 4 species, 6 reactions incl. the same-species class C + C -> B + D, the
    two-product bimolecular B + D -> C + C and two two-product unimoleculars), at bench.py's density, built by
    bench.build_model and populated by the same device release: every iteration starts from a common state and is
    compared bit for bit — traces of every molecule alive at the start, statistics, per-rule reaction counts — and the
    resulting populations as (species, position, times) multisets, because second products take fresh ids from device
    atomics.  Molecules whose id does not depend on that order are also compared by id."""
    import bench
    n = 20000
    t, edge_um = bench.build_model(n, seed=5, cap_factor=2.0)
    e, o = _engine(t), _oracle(t)
    assert bench.populate_by_release(e, n, edge_um, t.length_unit) == bench.populate_by_release(o, n, edge_um, t.length_unit)
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    _assert_same_population(a, b)                     # the released population itself is identical
    assert (np.bincount(a.species, minlength=4) == [8000, 8000, 2000, 2000]).all()
    nA0 = 8000 + 2000 + 2 * 2000                      # A + C + 2 D  (C = AB, D = A2B)
    nB0 = 8000 + 2000 + 2000                          # B + C + D
    rules_seen = np.zeros(6, np.int64)

    def key(m):
        arr = np.c_[m.species[:m.n].astype(float), m.x[:m.n], m.y[:m.n], m.z[:m.n], m.diffusion_time[:m.n],
                    m.unimol_rxn_time[:m.n], m.flags[:m.n].astype(float)]
        return arr[np.lexsort(arr.T[::-1])]

    state = a
    for it in range(12):
        # a common start: the oracle's population (the engines keep their own iteration counters and cumulative counts)
        n_ids = int(state.id.max()) + 1
        r_before_o, r_before_g = o.counts()[1].astype(np.int64), e.counts()[1].astype(np.int64)
        if it:
            e.upload(state)
            o.upload(state)
        tr_o, st_o = o.trace_step(1, n_ids)
        tr_g, st_g = e.trace_step(n_ids)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        assert set(live.tolist()) == set(state.id.tolist()), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "unimol_rxns", "vol_mol_vol_mol_collisions", "resolve_retries", "unresolved_conflicts",
                  "products_created", "mol_wall_reflections", "molecule_steps", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        d_o, d_g = o.counts()[1].astype(np.int64) - r_before_o, e.counts()[1].astype(np.int64) - r_before_g
        assert (d_o == d_g).all(), (it, d_o, d_g)
        rules_seen += d_g
        assert (o.counts()[0] == e.counts()[0]).all(), it
        c = e.counts()[0].astype(np.int64)
        assert c[0] + c[2] + 2 * c[3] == nA0 and c[1] + c[2] + c[3] == nB0, (it, c)
        a, b = o.download(), e.download()
        assert a.n == b.n
        ka, kb = key(a), key(b)
        assert (ka[:, 0] == kb[:, 0]).all() and (ka[:, 6] == kb[:, 6]).all(), it
        assert cm.rel_close(ka[:, 1:6], kb[:, 1:6], POS_TOL).all(), it
        # ids handed out before this iteration belong to the same molecules on both sides
        sa, sb = a.sorted_by_id(), b.sorted_by_id()
        old_a, old_b = sa.id < n_ids, sb.id < n_ids
        assert (sa.id[old_a] == sb.id[old_b]).all() and (sa.species[old_a] == sb.species[old_b]).all(), it
        assert cm.rel_close(np.c_[sa.x, sa.y, sa.z][old_a], np.c_[sb.x, sb.y, sb.z][old_b], POS_TOL).all(), it
        state = sa
    assert (rules_seen > 0).all(), rules_seen          # every one of the six rules fired, incl. C + C and B + D -> C + C
    assert st_g.unresolved_conflicts == 0


@pytest.mark.parametrize("static_b", [False, True])
def test_philox_surface_surface_reactions(static_b):
    """SURVEY 8 a23, react_2D_all_neighbors: surface molecules that react with the molecules on the tiles around their own —
    after their move, over the neighbour tiles of the static table (walls without a grid left out), one test with the
    local probability factor, several candidates through test_many_bimolecular; both consumed (product on the initiator's
    tile), catalytic (partner kept), two surface products over the two freed tiles (the random assignment), a class whose
    orientations never match.  static_b: a species that cannot diffuse but initiates.  Traces (partners in list order,
    class, pathway, orientation / tile bits in the event hash), conflict rounds, counts, the whole population."""
    t, mols = cm.surface_reactions(seed=4, static_b=static_b)
    n = mols.n
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)
    rx = retries = moved = 0
    for it in range(14):
        tr_o, st_o = o.trace_step(1, n)
        tr_g, st_g = e.trace_step(n)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "unimol_rxns", "resolve_retries", "unresolved_conflicts", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        rx += st_g.bimol_rxns
        retries += st_g.resolve_retries
        moved += int((tr_g["outcome"][live] == abi.MCX_OUT_SURFMOVE).sum())
        co, cg = o.counts(), e.counts()
        assert (cg[0] == co[0]).all() and (cg[1] == co[1]).all(), it
    assert rx > 150 and retries > 5 and moved > 3000, (rx, retries, moved)
    sp, rule = (np.asarray(a, dtype=np.int64) for a in e.counts())
    assert rule[0] > 100 and rule[2] > 20 and rule[3] > 3 and rule[4] == 0, rule
    # every rule moved the species counts the way it says (A, B, C, D, E, V)
    assert sp[1] == 1500 - rule[0] + rule[3] and sp[2] == rule[0] - rule[1] and sp[3] == rule[2] - 2 * rule[3] and sp[4] == 300
    assert sp[0] == 1500 - rule[0] + rule[1] - rule[2] + rule[3]
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    _assert_same_population(a, b)
    s = b.wall != abi.MCX_NONE
    assert len(np.unique(np.stack([b.wall[s], b.tile[s]], 1), axis=0)) == int(s.sum())


def test_unsupported_surface_surface_pathways_are_refused():
    """mcx_set_reactions refuses surface-surface pathways it cannot place (products on vacant neighbour tiles; the
    reference's non-terminating tile assignment) and the combination with region borders, with a message."""
    from mcell_b200 import McxError
    for what, t in cm.unsupported_surface_surface_tables():
        with pytest.raises(McxError) as ei:
            _engine(t)
        assert "surface-surface" in str(ei.value), (what, str(ei.value))


def test_replay_surface_surface_reactions():
    """The surface-surface scenario with per-molecule ISAAC64 tape slices (replay mode): the draws of the neighbour test,
    the tile assignment and the orientations come off the replayed stream in the reference's order on both sides."""
    t, mols = cm.surface_reactions(seed=11, rng_mode=abi.MCX_RNG_TAPE)
    n = mols.n
    o, e = _oracle(t), _engine(t)
    o.upload(mols)
    e.upload(mols)
    rx = 0
    for it in range(5):
        words, off = cm.isaac_slices(300 + it, n, 64)
        tr_o, st_o = o.trace_step(2, n, words, off)
        tr_g, st_g = e.replay_step(words, off)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "unimol_rxns", "resolve_retries", "unresolved_conflicts", "n_live"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        rx += st_g.bimol_rxns
        assert (np.asarray(e.counts()[1]) == np.asarray(o.counts()[1])).all(), it
    assert rx > 300, rx
    _assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())
