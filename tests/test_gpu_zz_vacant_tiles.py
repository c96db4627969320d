"""GPU parity of surface products on vacant neighbour tiles (the general branch of
find_surf_product_positions, src4/diffuse_react_event.cpp:2060-2100, 2155-2285): libmcx through its C ABI against the
CPU oracle on the same seeded inputs, bit for bit on traces, counts and populations."""
import numpy as np
import pytest

import common as cm
from mcell_b200 import abi

pytestmark = pytest.mark.gpu


def _run(t, mols, iterations, replay=False, seed0=500):
    """Every iteration starts from a common state (the oracle's population): products beyond the recycled ids take fresh
    ids, which the device hands out in its own global order, so ids created during an iteration are not compared — the
    populations are, as multisets, and every molecule that existed before the iteration by id."""
    from mcell_b200 import Engine
    from oracle import oracle_py as O
    o, e = O.Oracle(t), Engine(t)
    state = mols.sorted_by_id()
    totals = {"rx": 0, "uni": 0, "retries": 0, "rules": np.zeros(8, np.int64)}

    def key(m):
        arr = np.c_[m.species[:m.n].astype(float), m.wall[:m.n].astype(float), m.tile[:m.n].astype(float), m.flags[:m.n].astype(float),
                    m.orientation[:m.n].astype(float), m.x[:m.n], m.y[:m.n], m.z[:m.n], m.u[:m.n], m.v[:m.n], m.diffusion_time[:m.n]]
        return arr[np.lexsort(arr.T[::-1])]

    for it in range(iterations):
        e.upload(state)
        o.upload(state)
        n_ids = int(state.id.max()) + 1
        r_o, r_g = np.asarray(o.counts()[1], np.int64), np.asarray(e.counts()[1], np.int64)
        if replay:
            words, off = cm.isaac_slices(seed0 + it, n_ids, 64)
            tr_o, st_o = o.trace_step(2, n_ids, words, off)
            tr_g, st_g = e.replay_step(words, off)
        else:
            tr_o, st_o = o.trace_step(1, n_ids)
            tr_g, st_g = e.trace_step(n_ids)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert (np.flatnonzero(tr_g["rounds"] > 0) == live).all(), it
        bad = cm.compare_traces(tr_o, tr_g, live, check_rounds=True)
        assert not bad, (it, bad)
        for k in ("bimol_rxns", "unimol_rxns", "resolve_retries", "unresolved_conflicts", "n_live", "products_created"):
            assert getattr(st_g, k) == getattr(st_o, k), (it, k)
        d_o, d_g = np.asarray(o.counts()[1], np.int64) - r_o, np.asarray(e.counts()[1], np.int64) - r_g
        assert (d_o == d_g).all(), (it, d_o, d_g)
        totals["rules"][:len(d_g)] += d_g[:8]
        assert (np.asarray(o.counts()[0]) == np.asarray(e.counts()[0])).all(), it
        totals["rx"] += st_g.bimol_rxns; totals["uni"] += st_g.unimol_rxns; totals["retries"] += st_g.resolve_retries
        a, b = o.download(), e.download()
        assert a.n == b.n
        ka, kb = key(a), key(b)
        assert (ka[:, :5] == kb[:, :5]).all(), it                       # species, wall, tile, flags, orientation
        close = cm.rel_close(ka[:, 5:], kb[:, 5:], 1e-12)
        if not close.all():
            rows = np.flatnonzero(~close.all(axis=1))[:6]
            raise AssertionError((it, [(ka[r].tolist(), kb[r].tolist()) for r in rows]))   # position, uv, time
        sa, sb = a.sorted_by_id(), b.sorted_by_id()
        old_a, old_b = sa.id < n_ids, sb.id < n_ids
        assert (sa.id[old_a] == sb.id[old_b]).all() and (sa.species[old_a] == sb.species[old_b]).all(), it
        assert (sa.wall[old_a] == sb.wall[old_b]).all() and (sa.tile[old_a] == sb.tile[old_b]).all(), it
        s = sb.wall != abi.MCX_NONE
        assert len(np.unique(np.stack([sb.wall[s], sb.tile[s]], 1), axis=0)) == int(s.sum()), it
        assert (o.wall_grids() == e.wall_grids()).all(), it
        state = sa
    return totals


def test_philox_products_on_vacant_neighbour_tiles():
    """R' -> R' + A', L' + R' -> P' + A', P' -> A' + A', A' + A' -> R', A' -> V,: the tiles the extra products draw (the row of
    the neighbour-tile table from its back, rng % vacant), the claims on them in the conflict rounds, random points of the
    tiles, fresh molecule ids."""
    t, mols = cm.vacant_tile_products(seed=7)
    totals = _run(t, mols, 12)
    r = totals["rules"]
    assert r[0] > 100 and r[1] > 10 and r[2] > 2 and r[3] > 50 and r[4] > 50, r
    assert totals["retries"] > 20


def test_replay_products_on_vacant_neighbour_tiles():
    t, mols = cm.vacant_tile_products(seed=8, rng_mode=abi.MCX_RNG_TAPE)
    totals = _run(t, mols, 5, replay=True)
    assert totals["uni"] > 50


def test_crowded_surface_blocks_reactions_without_room():
    """RX_BLOCKED: on a nearly full surface most emissions find no vacant neighbour tile; the molecule lives on and draws
    a new lifetime, nothing is placed, tiles stay exclusive."""
    t, mols = cm.vacant_tile_products(n_r=6000, n_a=1500, n_lig=200, seed=9)
    totals = _run(t, mols, 6)
    assert totals["rules"][0] > 10
