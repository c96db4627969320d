"""CPU tier: host-side logic of the multi-GPU slab decomposition with torch.distributed (gloo, world_size 2):
unique-id distribution, the numpy mirror of the device's slab-ownership arithmetic, count reduction."""
import os
import socket

import numpy as np
import pytest

from mcell_b200 import abi, comm
from mcell_b200.model import MolArrays


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = bytes((7 * i + 3) % 256 for i in range(comm.NCCL_UNIQUE_ID_BYTES))
        got = comm.broadcast_unique_id(dist, rank, make_id=lambda: blob)
        assert got == blob
        # every rank sees the same global population and keeps what its device would own
        rng = np.random.default_rng(5)
        n = 20000
        m = MolArrays(n)
        m.x[:], m.y[:], m.z[:] = rng.uniform(-50, 50, (3, n))
        m.z[:7] = [-50.0, 50.0, -1e9, 1e9, 0.0, -0.5 * 3.39, 3.39 * 14.5]     # faces, out of range, layer boundaries
        m.id[:] = np.arange(n)
        info = abi.mcx_slab_info()
        info.grid_origin_z, info.layer_rcp, info.n_layers = -50.0 - 0.5 * 3.39, 1.0 / 3.39, 31
        info.rank, info.world_size = rank, world
        info.layer_lo, info.layer_hi = comm.layer_range(info.n_layers, rank, world)
        mine = comm.select_owned(m, info)
        lay = comm.layer_of(mine.z, info.grid_origin_z, info.layer_rcp, info.n_layers)
        assert ((lay >= info.layer_lo) & (lay < info.layer_hi)).all()
        z_lo, z_hi = comm.owned_z_interval(info)
        inner = (mine.z > z_lo + 1e-9) & (mine.z < z_hi - 1e-9)
        assert inner.sum() >= mine.n - 4
        tot = comm.allreduce_sum(dist, [mine.n, float(mine.id.sum())])
        assert tot[0] == n and tot[1] == n * (n - 1) / 2          # disjoint cover of the population
        out_q.put((rank, mine.n))
    finally:
        dist.destroy_process_group()


def test_slab_ownership_and_plumbing_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(2))
    assert got[0] > 0 and got[1] > 0 and got[0] + got[1] == 20000


def test_layer_ranges_tile_the_grid():
    for n_layers in (7, 31, 2738):
        for world in (1, 2, 4, 8):
            edges = [comm.layer_range(n_layers, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n_layers
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
    # balanced slabs: with a halo of H layers per neighbour every rank evaluates the same thickness (+-1 layer)
    for n_layers, halo in ((334, 17), (2738, 40)):
        for world in (2, 4, 8):
            edges = [comm.layer_range(n_layers, r, world, halo) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n_layers
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            work = [(hi - lo) + halo * ((r > 0) + (r < world - 1)) for r, (lo, hi) in enumerate(edges)]
            assert max(work) - min(work) <= 1, (n_layers, halo, world, work)
            assert min(hi - lo for lo, hi in edges) >= halo
    z = np.array([-1e30, -3.0, 0.0, 2.999, 3.0, 1e30])
    assert comm.layer_of(z, 0.0, 1.0 / 3.0, 10).tolist() == [0, 0, 0, 0, 1, 9]
    assert comm.rank_of(z, (0.0, 1.0 / 3.0, 10), world=2).tolist() == [0, 0, 0, 0, 0, 1]


def test_bench_capacity_covers_the_halo_copies_of_every_rank_count():
    """A slab rank holds its own layers, a halo on each side (3x the one-step reach, configure_slab) and — between the
    halo refresh and the sort — the refreshed copies as well; bench.py must size max_molecules for that at every N the
    driver runs (N = 4 and 8 once ran out of slots with a fixed factor)."""
    import bench
    n = 100_000_000
    for world in (2, 4, 8):
        t, edge_um = bench.build_model(n, world=world, rank=1)
        edge_lu = edge_um / t.length_unit
        reach = t.cfg.rxn_radius_3d + 6.993 * max(sp.space_step for sp in t.species)
        halo = 3.0 * reach + 3.5                      # + one cell layer of rounding
        need = n / world * (1.0 + 4.0 * halo / (edge_lu / world)) * 1.05   # + products of one iteration
        assert t.cfg.max_molecules >= need, (world, t.cfg.max_molecules, need)


def test_select_owned_carries_surface_and_counted_volume_fields():
    """comm.select_owned hands a rank its molecules with EVERY field of the record: Molecule::s (wall, tile,
    orientation, uv) of surface molecules and the counted-volume index of volume molecules (a view() of arrays shorter
    than n silently drops them, which reset counted volumes to 0 and made surface uploads fail)."""
    from mcell_b200 import abi
    from mcell_b200.model import MolArrays
    n = 1000
    rng = np.random.default_rng(3)
    m = MolArrays(n)
    m.z[:] = rng.uniform(0, 30, n)
    m.id[:] = np.arange(n)
    m.species[:] = rng.integers(0, 4, n)
    m.counted_volume[:] = rng.integers(0, 5, n)
    surf = rng.random(n) < 0.3
    m.wall[surf] = rng.integers(0, 50, surf.sum())
    m.tile[surf] = rng.integers(0, 9, surf.sum())
    m.orientation[surf] = rng.choice([-1, 1], surf.sum())
    m.u[:] = rng.random(n)
    m.v[:] = rng.random(n)
    info = abi.mcx_slab_info()
    info.grid_origin_z, info.layer_rcp, info.n_layers, info.world_size, info.halo_layers = 0.0, 1.0 / 3.0, 10, 2, 0
    got = []
    for rank in range(2):
        info.rank = rank
        part = comm.select_owned(m, info)
        keep = np.flatnonzero(comm.rank_of(m.z, info) == rank)
        assert part.n == len(keep) > 0
        for k in MolArrays.FIELDS:
            assert len(getattr(part, k)) == part.n, k
            assert (getattr(part, k) == getattr(m, k)[keep]).all(), k
        v = part.view()
        assert v.wall and v.counted_volume          # the ABI view carries them (non-null pointers)
        got.append(part.n)
    assert sum(got) == n
    # a source without the optional arrays (bench.make_molecules) still works
    lean = MolArrays(0)
    for k in ("x", "y", "z", "id", "species", "flags", "diffusion_time", "unimol_rxn_time"):
        setattr(lean, k, getattr(m, k))
    lean.n = n
    info.rank = 0
    assert comm.select_owned(lean, info).n == got[0]
