"""GPU tier (needs 2 devices; skipped otherwise): two ranks with NCCL halo exchange reproduce the single-GPU run
molecule for molecule.  The ranks run as two host threads of this process (one libmcx handle per device; ctypes
releases the GIL during the calls), which is the same C ABI sequence bench.py issues under torchrun."""
import threading

import numpy as np
import pytest

import common as cm
from mcell_b200 import abi, comm

pytestmark = pytest.mark.gpu


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run_ranks(tables_fn, mols, n_iter, world=2, halo_width=0.0, script=None):
    from mcell_b200 import Engine
    uid = comm.unique_id()
    out, errs = [None] * world, []
    barrier = threading.Barrier(world)

    def work(rank):
        try:
            t = tables_fn()
            t.cfg.device, t.cfg.rank, t.cfg.world_size, t.cfg.halo_width = rank, rank, world, halo_width
            e = Engine(t)
            barrier.wait()
            e.comm_init(uid)
            e.upload(mols)                      # every rank uploads everything; foreign slabs are dropped
            stats = script(e) if script else [e.step(1) for _ in range(n_iter)]
            by_volume = e.counts_by_volume() if len(t.counted_volume_sets) > 1 else None
            out[rank] = (e.download(), stats, e.counts(), e.slab_info(), by_volume)
            e.close()
        except Exception as ex:                 # noqa: BLE001
            errs.append((rank, ex))
            try:
                barrier.abort()
            except Exception:
                pass

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for x in th:
        x.start()
    for x in th:
        x.join(600)
    assert not errs, errs
    return out


@pytest.mark.skipif(_n_devices() < 2, reason="needs 2 CUDA devices")
@pytest.mark.parametrize("scenario", ["reactive", "free"])
def test_two_ranks_match_single_gpu(scenario):
    from mcell_b200 import Engine
    n, n_iter = 40000, 6
    if scenario == "reactive":
        make = lambda: cm.reactive_box(n=n, edge_um=1.0, p_target=0.5, seed=4, cap_factor=3)  # noqa: E731
    else:
        make = lambda: cm.free_diffusion_box(n=n, edge_um=1.0, seed=4, cap_factor=3)          # noqa: E731
    t, mols = make()
    single = Engine(t)
    single.upload(mols)
    st1 = [single.step(1) for _ in range(n_iter)]
    ref = single.download().sorted_by_id()
    ref_counts = single.counts()
    res = _run_ranks(lambda: make()[0], mols, n_iter, halo_width=45.0)
    parts = [r[0] for r in res]
    ids = np.concatenate([p.id[:p.n] for p in parts])
    assert len(ids) == ref.n and len(np.unique(ids)) == ref.n          # disjoint ownership, nothing lost
    o = np.argsort(ids, kind="stable")
    for k in ("id", "species", "x", "y", "z", "flags", "diffusion_time", "unimol_rxn_time"):
        got = np.concatenate([getattr(p, k)[:p.n] for p in parts])[o]
        assert (got == getattr(ref, k)[:ref.n]).all(), k               # bit for bit
    # ownership agrees with the host-side mirror of the slab arithmetic
    for p, r in zip(parts, res):
        assert (comm.rank_of(p.z[:p.n], r[3]) == r[3].rank).all()
    # counts: summed over ranks by the library; per-iteration statistics add up to the single-GPU ones
    assert (res[0][2][0] == ref_counts[0]).all() and (res[1][2][0] == ref_counts[0]).all()
    assert (res[0][2][1] == ref_counts[1]).all()
    for it in range(n_iter):
        for k in ("molecule_steps", "bimol_rxns", "mol_wall_reflections", "vol_mol_vol_mol_collisions"):
            assert sum(getattr(r[1][it], k) for r in res) == getattr(st1[it], k), (it, k)


@pytest.mark.skipif(_n_devices() < 2, reason="needs 2 CUDA devices")
@pytest.mark.parametrize("scenario", ["receptors", "surface_diffusion", "counted_volumes", "transporter", "permeable", "region_border",
                                      "surface_surface"])
def test_two_ranks_match_single_gpu_surface_and_counted(scenario):
    """Surface molecules (tiles, binding, unbinding, 2-D diffusion across the slab face) and counted volumes with two
    ranks: the halo records carry Molecule::s and the creation wall / tile of surface-born volume products, the counted
    volume rides in the record flags; the population and the per-volume counts equal the single-GPU run bit for bit."""
    from mcell_b200 import Engine
    n_iter = 8
    if scenario == "receptors":
        make = lambda: cm.ligand_receptor_sphere(n_lig=30000, n_rec=3000, n_pump=1500, radius_um=0.5, subdivisions=4,  # noqa: E731
                                                 box_um=1.6, seed=6, release_products=False)
        halo = 62.0   # 3 x (R + 6.993 * space_step of Ca, D = 2e-6)
    elif scenario == "surface_diffusion":
        make = lambda: cm.diffusing_receptors(n_rec=4000, n_lig=20000, radius_um=0.5, subdivisions=4, box_um=1.6, seed=7)  # noqa: E731
        halo = 45.0
    elif scenario == "transporter":   # kept volume reactant passing through the wall (RX_FLIP): the kept molecule's guard travels in the halo
        make = lambda: cm.transporter_sphere(n_vol=30000, n_trans=4000, n_enz=500, radius_um=0.5, subdivisions=4, box_um=1.6,  # noqa: E731
                                             seed=6, enzyme=False)
        halo = 45.0
    elif scenario == "permeable":     # finite-rate reactions with a surface class, both directions
        make = lambda: cm.permeable_sphere(n=30000, radius_um=0.5, subdivisions=4, box_um=1.6, seed=6, products=False)  # noqa: E731
        halo = 45.0
    elif scenario == "surface_surface":   # react_2D_all_neighbors across the slab face; Wall::has_initialized_grid shared by the ranks
        make = lambda: cm.surface_reactions(n_a=3000, n_b=3000, n_e=600, radius_um=0.5, subdivisions=4, box_um=1.6, seed=6)  # noqa: E731
        halo = 45.0
    elif scenario == "region_border":
        make = lambda: cm.diffusing_receptors(n_rec=4000, n_lig=20000, radius_um=0.5, subdivisions=4, box_um=1.6, seed=7,  # noqa: E731
                                              D_surf=4e-7, border=abi.MCX_SURF_REFLECTIVE)
        halo = 45.0
    else:
        make = lambda: cm.counted_spheres(n=30000, seed=8, box_um=1.2)  # noqa: E731
        halo = 45.0
    t, mols = make()
    single = Engine(t)
    single.upload(mols)
    st1 = [single.step(1) for _ in range(n_iter)]
    ref = single.download().sorted_by_id()
    ref_counts = single.counts()
    ref_by_volume = single.counts_by_volume() if len(t.counted_volume_sets) > 1 else None
    res = _run_ranks(lambda: make()[0], mols, n_iter, halo_width=halo)
    parts = [r[0] for r in res]
    ids = np.concatenate([p.id[:p.n] for p in parts])
    assert len(ids) == ref.n and len(np.unique(ids)) == ref.n
    o = np.argsort(ids, kind="stable")
    for k in ("id", "species", "x", "y", "z", "flags", "diffusion_time", "unimol_rxn_time", "wall", "tile", "orientation",
              "u", "v", "counted_volume"):
        got = np.concatenate([getattr(p, k)[:p.n] for p in parts])[o]
        assert (got == getattr(ref, k)[:ref.n]).all(), k
    assert (res[0][2][0] == ref_counts[0]).all() and (res[1][2][0] == ref_counts[0]).all()
    assert (res[0][2][1] == ref_counts[1]).all()
    for it in range(n_iter):
        for k in ("molecule_steps", "bimol_rxns", "unimol_rxns", "mol_wall_reflections", "mol_wall_transparent"):
            assert sum(getattr(r[1][it], k) for r in res) == getattr(st1[it], k), (it, k)
    assert sum(getattr(s_, "bimol_rxns") for s_ in st1) > 20
    if ref_by_volume is not None:
        for r in res:
            assert (r[4][0] == ref_by_volume[0]).all() and (r[4][1] == ref_by_volume[1]).all()


@pytest.mark.skipif(_n_devices() < 2, reason="needs 2 CUDA devices")
def test_two_ranks_release_on_device_matches_single_gpu():
    """mcx_release_volume_molecules with two ranks: every rank makes the same call and keeps the molecules of its slab;
    ids, positions and everything that follows equal the single-GPU run bit for bit (a sphere across the slab face and
    a shell released inside an iteration)."""
    from mcell_b200 import Engine
    n = 30000

    def make():
        return cm.reactive_box(n=n, edge_um=1.0, p_target=0.5, seed=12, cap_factor=4)

    def script(e):
        st = [e.step(1) for _ in range(3)]
        a = e.release(0, 6000, (0.0, 0.0, 0.0), (60.0, 60.0, 60.0), shape=abi.MCX_RELEASE_SPHERICAL)
        b = e.release(1, 4000, (5.0, -5.0, 2.0), (50.0, 40.0, 70.0), shape=abi.MCX_RELEASE_SPHERICAL_SHELL, release_time=3.5)
        assert (a, b) == (n, n + 6000)
        return st + [e.step(1) for _ in range(4)]

    t, mols = make()
    single = Engine(t)
    single.upload(mols)
    st1 = script(single)
    ref = single.download().sorted_by_id()
    ref_counts = single.counts()
    res = _run_ranks(lambda: make()[0], mols, 7, halo_width=45.0, script=script)
    parts = [r[0] for r in res]
    ids = np.concatenate([p.id[:p.n] for p in parts])
    assert len(ids) == ref.n and len(np.unique(ids)) == ref.n
    o = np.argsort(ids, kind="stable")
    for k in ("id", "species", "x", "y", "z", "flags", "diffusion_time", "unimol_rxn_time"):
        got = np.concatenate([getattr(p, k)[:p.n] for p in parts])[o]
        assert (got == getattr(ref, k)[:ref.n]).all(), k
    assert (res[0][2][0] == ref_counts[0]).all() and (res[1][2][0] == ref_counts[0]).all()
    for it in range(7):
        for k in ("molecule_steps", "bimol_rxns"):
            assert sum(getattr(r[1][it], k) for r in res) == getattr(st1[it], k), (it, k)


@pytest.mark.parametrize("world", [3, 4])
def test_interior_ranks_match_single_gpu(world):
    """Three and four ranks: the inner ranks have a neighbour on BOTH sides (two halos, both receive buffers, both
    parities of the exchange), which two ranks never exercise.  Slow species (D = 1e-7: reach 5 lu, halo 15 lu) keep
    every slab of the 1 um box thicker than its halo.  Population, counts and per-iteration statistics equal the
    single-GPU run bit for bit over 10 iterations."""
    if _n_devices() < world:
        pytest.skip("needs %d CUDA devices" % world)
    from mcell_b200 import Engine
    n, n_iter = 120000, 10
    make = lambda: cm.reactive_box(n=n, edge_um=1.0, p_target=0.7, seed=9, cap_factor=3, D=1e-7)  # noqa: E731
    t, mols = make()
    single = Engine(t)
    single.upload(mols)
    st1 = [single.step(1) for _ in range(n_iter)]
    ref = single.download().sorted_by_id()
    ref_counts = single.counts()
    assert sum(s_.bimol_rxns for s_ in st1) > 300
    res = _run_ranks(lambda: make()[0], mols, n_iter, world=world)
    infos = [r[3] for r in res]
    assert all(i.halo_layers > 0 for i in infos)
    assert all(infos[k].layer_hi == infos[k + 1].layer_lo for k in range(world - 1))
    parts = [r[0] for r in res]
    assert all(p.n > 0 for p in parts)
    ids = np.concatenate([p.id[:p.n] for p in parts])
    assert len(ids) == ref.n and len(np.unique(ids)) == ref.n
    o = np.argsort(ids, kind="stable")
    for k in ("id", "species", "x", "y", "z", "flags", "diffusion_time", "unimol_rxn_time"):
        got = np.concatenate([getattr(p, k)[:p.n] for p in parts])[o]
        assert (got == getattr(ref, k)[:ref.n]).all(), k
    for r in res:
        assert (r[2][0] == ref_counts[0]).all() and (r[2][1] == ref_counts[1]).all()
    for it in range(n_iter):
        for k in ("molecule_steps", "bimol_rxns", "mol_wall_reflections", "vol_mol_vol_mol_collisions"):
            assert sum(getattr(r[1][it], k) for r in res) == getattr(st1[it], k), (it, k)
