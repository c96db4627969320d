"""Host-side plumbing of the multi-GPU slab decomposition (one process per GPU).

The data path is NCCL inside libmcx (csrc/mcx_comm.cu); this module only (a) creates / distributes the
ncclUniqueId, (b) mirrors the slab-ownership arithmetic of include/mcx.h (mcx_slab_info) in numpy so that a host
can hand each rank exactly the molecules its device owns, and (c) sums per-rank statistics.  torch.distributed is
used for rendezvous only (backend nccl on GPUs, gloo in the CPU tests)."""
import ctypes as C

import numpy as np

from . import abi

NCCL_UNIQUE_ID_BYTES = 128


def unique_id():
    """ncclGetUniqueId through libmcx (rank 0 only)."""
    from .engine import load_library, McxError
    buf = C.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
    n = load_library().mcx_comm_unique_id(C.cast(buf, C.c_void_p), NCCL_UNIQUE_ID_BYTES)
    if n < 0:
        raise McxError(n, "ncclGetUniqueId failed")
    return bytes(buf.raw[:NCCL_UNIQUE_ID_BYTES])


def broadcast_unique_id(dist, rank, device=None, make_id=unique_id):
    """Rank 0 creates the id, every rank receives its 128 bytes (works with nccl and gloo backends)."""
    import torch
    dev = device if device is not None else "cpu"
    t = torch.zeros(NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.tensor(list(make_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def layer_of(z, grid_origin_z, layer_rcp, n_layers):
    """Global z-layer of positions z — the arithmetic of cell_z() in csrc/mcx_device.cuh (IEEE double)."""
    c = np.floor((np.asarray(z, np.float64) - grid_origin_z) * layer_rcp)
    return np.clip(c, 0, n_layers - 1).astype(np.int64)   # clamp first: the device conversion saturates


def layer_range(n_layers, rank, world, halo_layers=0):
    """Layers [lo, hi) owned by `rank` (configure_slab in csrc/mcx_api.cu): the two outermost ranks, which have one
    halo only, own `halo_layers` more than the inner ones so that every rank evaluates the same number of layers."""
    def bound(k):
        if k <= 0:
            return 0
        if k >= world:
            return n_layers
        return halo_layers + ((n_layers - 2 * halo_layers) * k) // world
    return bound(rank), bound(rank + 1)


def rank_of(z, info_or_tuple, world=None):
    """Owning rank of every position."""
    halo = 0
    if isinstance(info_or_tuple, abi.mcx_slab_info):
        g0, rcp, n, world = info_or_tuple.grid_origin_z, info_or_tuple.layer_rcp, info_or_tuple.n_layers, info_or_tuple.world_size
        halo = info_or_tuple.halo_layers
    else:
        g0, rcp, n = info_or_tuple[:3]
        if len(info_or_tuple) > 3:
            halo = info_or_tuple[3]
    lay = layer_of(z, g0, rcp, n)
    bounds = np.array([layer_range(n, r, world, halo)[1] for r in range(world)], dtype=np.int64)
    return np.searchsorted(bounds, lay, side="right")


def owned_z_interval(info):
    """[z_lo, z_hi) of the owned layers in length units (outermost slabs extend to infinity by clamping)."""
    z_lo = info.grid_origin_z + info.layer_lo / info.layer_rcp
    z_hi = info.grid_origin_z + info.layer_hi / info.layer_rcp
    if info.layer_lo == 0:
        z_lo = -np.inf
    if info.layer_hi == info.n_layers:
        z_hi = np.inf
    return z_lo, z_hi


def select_owned(mols, info):
    """Subset of a MolArrays that this rank's device owns."""
    from .model import MolArrays
    r = rank_of(mols.z[:mols.n], info)
    keep = np.flatnonzero(r == info.rank)
    out = MolArrays(0)
    for k in MolArrays.FIELDS:   # incl. Molecule::s (wall, tile, orientation, u, v) and counted_volume
        a = getattr(mols, k)
        if a is None or len(a) < mols.n:   # an array the source does not carry (MolArrays.view() omits it too)
            continue
        setattr(out, k, np.ascontiguousarray(a[:mols.n][keep]))
    out.n = len(keep)
    return out


def allreduce_sum(dist, values, device=None):
    import torch
    t = torch.tensor(np.asarray(values, dtype=np.float64), dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
