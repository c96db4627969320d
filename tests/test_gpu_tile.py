"""The shared-memory tile path of the fast pass (k_diffuse_tile, csrc/mcx_tile.cuh: bulk copies + mbarrier, fp32
pre-filter, exact fp64 confirmation) against the gather walk and the oracle: same traces bit for bit."""
import os

import numpy as np
import pytest

import common as cm
from mcell_b200 import abi

pytestmark = pytest.mark.gpu


def _run(t, mols, n_ids, tile, iters, seeds):
    from mcell_b200 import Engine
    old = os.environ.get("MCX_TILE")
    os.environ["MCX_TILE"] = "1" if tile else "0"
    try:
        e = Engine(t)
        e.upload(mols)
        traces, stats = [], []
        for it in range(iters):
            words, off = cm.isaac_slices(seeds + it, n_ids, 48)
            tr, st = e.replay_step(words, off)
            assert e.fast_pass_kind() == (1 if tile else 0)
            traces.append(tr.copy()); stats.append(st)
        return traces, stats, e.download().sorted_by_id()
    finally:
        if old is None:
            os.environ.pop("MCX_TILE", None)
        else:
            os.environ["MCX_TILE"] = old


@pytest.mark.parametrize("n,edge_um,subpart_um", [(16000, 0.4, 0.5), (60000, 0.8, 0.05)])
def test_tile_path_equals_gather_walk_and_oracle(n, edge_um, subpart_um):
    from oracle import oracle_py as O
    t, mols = cm.reactive_box(n=n, edge_um=edge_um, p_target=0.5, rng_mode=abi.MCX_RNG_TAPE, subpartition_dimension=subpart_um)
    iters = 3
    tr_t, st_t, pop_t = _run(t, mols, n, True, iters, 300)
    tr_f, st_f, pop_f = _run(t, mols, n, False, iters, 300)
    o = O.Oracle(t)
    o.upload(mols)
    ids = np.arange(n)
    for it in range(iters):
        words, off = cm.isaac_slices(300 + it, n, 48)
        tr_o, st_o = o.trace_step(2, n, words, off)
        live = np.flatnonzero(tr_o["rounds"] > 0)
        assert not cm.compare_traces(tr_o, tr_t[it], live, check_rounds=True)
        assert not cm.compare_traces(tr_f[it], tr_t[it], ids, check_rounds=True)
        for k in ("bimol_rxns", "vol_mol_vol_mol_collisions", "mol_wall_reflections", "ray_polygon_tests", "molecule_steps"):
            assert getattr(st_t[it], k) == getattr(st_f[it], k) == getattr(st_o, k), k
    assert sum(s.bimol_rxns for s in st_t) > 50
    assert pop_t.n == pop_f.n and (pop_t.id == pop_f.id).all() and (pop_t.species == pop_f.species).all()
    for k in ("x", "y", "z"):
        assert (getattr(pop_t, k) == getattr(pop_f, k)).all()
