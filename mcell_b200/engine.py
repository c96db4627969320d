"""Host-side binding of libmcx (ctypes over the C ABI in include/mcx.h).

The library is built in-tree (mcell_b200/build.py).  Importing works without a GPU (so the CPU
test tier can check the ABI); every compute entry point fails loudly with McxError when no CUDA
device is usable — there is no CPU fallback and nothing here touches oracle/."""
import ctypes as C
import os

import numpy as np

from . import abi
from .model import MolArrays

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCX_LIB: another build of the same library (tuning variants, tools/build_variants.sh)
LIB_PATH = os.environ.get("MCX_LIB") or os.path.join(_HERE, "libmcx.so")
_lib = None


class McxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmcx error %d: %s" % (code, msg))
        self.code = code


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McxError(abi.MCX_ERR_STATE, "libmcx.so is not built: run `python -m mcell_b200.build` "
                       "(__graft_entry__.build()); the CUDA extension is mandatory")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.mcx_create.argtypes = [C.POINTER(abi.mcx_config), C.POINTER(H)]
    L.mcx_destroy.argtypes = [H]
    L.mcx_destroy.restype = None
    L.mcx_last_error.argtypes = [H]
    L.mcx_last_error.restype = C.c_char_p
    L.mcx_set_geometry.argtypes = [H, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.mcx_set_species.argtypes = [H, C.c_void_p, C.c_uint32]
    L.mcx_set_reactions.argtypes = [H, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.mcx_set_surface_classes.argtypes = [H, C.c_void_p, C.c_uint32]
    L.mcx_upload_molecules.argtypes = [H, C.POINTER(abi.mcx_mol_soa)]
    L.mcx_download_molecules.argtypes = [H, C.POINTER(abi.mcx_mol_soa), C.c_uint64]
    L.mcx_num_molecules.argtypes = [H]
    L.mcx_num_molecules.restype = C.c_uint64
    L.mcx_step.argtypes = [H, C.c_uint32, C.POINTER(abi.mcx_step_stats)]
    L.mcx_replay_step.argtypes = [H, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(abi.mcx_step_stats)]
    L.mcx_trace_step.argtypes = [H, C.c_uint64, C.c_void_p, C.POINTER(abi.mcx_step_stats)]
    L.mcx_counts.argtypes = [H, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.mcx_comm_init.argtypes = [H, C.c_void_p, C.c_uint32]
    L.mcx_comm_unique_id.argtypes = [C.c_void_p, C.c_uint32]
    L.mcx_slab_info_get.argtypes = [H, C.POINTER(abi.mcx_slab_info)]
    L.mcx_comm_halo_path.argtypes = [H]
    L.mcx_fast_pass_kind.argtypes = [H]
    L.mcx_release_volume_molecules.argtypes = [H, C.POINTER(abi.mcx_release), C.POINTER(C.c_uint32)]
    L.mcx_release_list.argtypes = [H, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_uint32)]
    L.mcx_release_surface_molecules.argtypes = [H, C.POINTER(abi.mcx_surface_release), C.POINTER(C.c_uint32)]
    L.mcx_get_next_molecule_id.argtypes = [H, C.POINTER(C.c_uint32)]
    L.mcx_set_next_molecule_id.argtypes = [H, C.c_uint32]
    L.mcx_get_wall_grids.argtypes = [H, C.c_void_p, C.c_uint64]
    L.mcx_set_wall_grids.argtypes = [H, C.c_void_p, C.c_uint64]
    L.mcx_set_profiling.argtypes = [H, C.c_int]
    L.mcx_philox_block.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p]
    L.mcx_philox_block.restype = None
    L.mcx_set_counted_volumes.argtypes = [H, C.c_uint32, C.c_void_p, C.c_void_p]
    L.mcx_counts_by_volume.argtypes = [H, C.c_void_p, C.c_void_p]
    L.mcx_set_surface_regions.argtypes = [H, C.c_uint32, C.c_void_p]
    L.mcx_set_region_borders.argtypes = [H, C.c_void_p]
    L.mcx_set_counted_volume_objects.argtypes = [H, C.c_void_p, C.c_uint32]
    L.mcx_counts_by_surface_region.argtypes = [H, C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else None


class Engine:
    """One libmcx handle = the device-resident replacement of one DiffuseReactEvent + Partition."""

    def __init__(self, tables):
        self.L = load_library()
        self.t = tables
        self.h = C.c_void_p()
        rc = self.L.mcx_create(C.byref(tables.cfg), C.byref(self.h))
        if rc:
            msg = self.L.mcx_last_error(None).decode()
            self.h = None
            raise McxError(rc, msg)
        t = tables
        self._ck(self.L.mcx_set_species(self.h, C.cast(t.species, C.c_void_p), t.n_species))
        self._ck(self.L.mcx_set_reactions(self.h, C.cast(t.classes, C.c_void_p), t.n_classes,
                                          C.cast(t.pathways, C.c_void_p), t.n_pathways))
        self._ck(self.L.mcx_set_surface_classes(self.h, C.cast(t.surf_rules, C.c_void_p), t.n_surf_rules))
        self._ck(self.L.mcx_set_geometry(self.h, _vp(t.vertices), len(t.vertices), _vp(t.tri), len(t.tri),
                                         _vp(t.wall_surf_class), _vp(getattr(t, "wall_object", None))))
        if getattr(t, "n_counted_volumes", 0) > 1:
            self._ck(self.L.mcx_set_counted_volumes(self.h, t.n_counted_volumes, _vp(t.wall_cv_front), _vp(t.wall_cv_back)))
        if getattr(t, "n_counted_volumes", 0) > 1 and getattr(t, "cv_object_mask", None) is not None:
            self._ck(self.L.mcx_set_counted_volume_objects(self.h, _vp(t.cv_object_mask), t.cv_intersecting))
        if getattr(t, "n_region_sets", 0) > 1:
            self._ck(self.L.mcx_set_surface_regions(self.h, t.n_region_sets, _vp(t.wall_region_set)))
        if getattr(t, "wall_edge_border", None) is not None:
            self._ck(self.L.mcx_set_region_borders(self.h, _vp(t.wall_edge_border)))

    def _ck(self, rc):
        if rc:
            raise McxError(rc, self.L.mcx_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.mcx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, unique_id_bytes):
        buf = C.create_string_buffer(bytes(unique_id_bytes), len(unique_id_bytes))
        self._ck(self.L.mcx_comm_init(self.h, C.cast(buf, C.c_void_p), len(unique_id_bytes)))

    def halo_path(self):
        """0 = single device, 1 = NCCL send/recv, 2 = peer-memory stores over NVLink (mcx_comm_halo_path)."""
        return int(self.L.mcx_comm_halo_path(self.h))

    def fast_pass_kind(self):
        """0 = gather walk, 1 = shared-memory tiles (mcx_fast_pass_kind), for the last stepping call."""
        return int(self.L.mcx_fast_pass_kind(self.h))

    def slab_info(self):
        info = abi.mcx_slab_info()
        self._ck(self.L.mcx_slab_info_get(self.h, C.byref(info)))
        return info

    def set_profiling(self, enabled=True):
        self._ck(self.L.mcx_set_profiling(self.h, 1 if enabled else 0))

    def upload(self, mols):
        v = mols.view()
        self._ck(self.L.mcx_upload_molecules(self.h, C.byref(v)))

    def release(self, species, number, location, diameter, shape=abi.MCX_RELEASE_CUBIC, release_time=0.0, counted_volume_index=0, region_in=0, region_out=0, region_expr=()):
        """ReleaseEvent::release_ellipsoid_or_rectcuboid on the device (mcx_release_volume_molecules); returns the first
        id of the new molecules.  location / diameter in length units."""
        r = abi.mcx_release()
        r.species, r.shape, r.number = int(species), int(shape), int(number)
        r.location[:] = [float(v) for v in location]
        r.diameter[:] = [float(v) for v in diameter]
        r.release_time, r.counted_volume_index = float(release_time), int(counted_volume_index)
        r.region_in, r.region_out = int(region_in), int(region_out)
        r.region_expr_len = len(region_expr)   # postfix: object index, abi.MCX_REGION_UNION / _INTERSECT / _DIFFERENCE
        for q, op in enumerate(region_expr):
            r.region_expr[q] = int(op)
        first = C.c_uint32(0)
        self._ck(self.L.mcx_release_volume_molecules(self.h, C.byref(r), C.byref(first)))
        return int(first.value)

    def release_surface(self, species, number, walls, orientation=1, release_time=0.0, randomize_pos=True):
        """ReleaseEvent::release_onto_regions on the device (mcx_release_surface_molecules): `number` molecules of a
        surface species on vacant tiles of the listed walls; returns the first id."""
        wl = np.ascontiguousarray(walls, np.uint32)
        r = abi.mcx_surface_release()
        r.species, r.orientation, r.number, r.release_time = int(species), int(orientation), int(number), float(release_time)
        r.walls, r.n_walls, r.randomize_pos = wl.ctypes.data, len(wl), 1 if randomize_pos else 0
        first = C.c_uint32(0)
        self._ck(self.L.mcx_release_surface_molecules(self.h, C.byref(r), C.byref(first)))
        return int(first.value)

    def release_list(self, species, positions, counted_volume=None, release_time=0.0):
        """ReleaseEvent::release_list for volume molecules on the device (mcx_release_list): one molecule of species[k]
        at positions[k] (length units); returns the first id."""
        sp = np.ascontiguousarray(species, np.uint32)
        pos = np.asarray(positions, np.float64)
        x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
        cv = None if counted_volume is None else np.ascontiguousarray(counted_volume, np.uint32)
        first = C.c_uint32(0)
        self._ck(self.L.mcx_release_list(self.h, C.c_uint64(len(sp)), _vp(sp), _vp(x), _vp(y), _vp(z), _vp(cv) if cv is not None else None,
                                         C.c_double(release_time), C.byref(first)))
        return int(first.value)

    def next_molecule_id(self, set_to=None):
        """Partition::next_molecule_id: query, or (checkpoint resume) raise it to the saved value."""
        if set_to is not None:
            self._ck(self.L.mcx_set_next_molecule_id(self.h, C.c_uint32(int(set_to))))
        out = C.c_uint32(0)
        self._ck(self.L.mcx_get_next_molecule_id(self.h, C.byref(out)))
        return int(out.value)

    def wall_grids(self, set_to=None):
        """Wall::has_initialized_grid of every wall (one byte each): query, or (checkpoint resume) OR the saved flags in."""
        n = len(self.t.tri)
        if set_to is not None:
            a = np.ascontiguousarray(set_to, dtype=np.uint8)
            self._ck(self.L.mcx_set_wall_grids(self.h, C.c_void_p(a.ctypes.data), n))
        out = np.zeros(max(n, 1), np.uint8)
        self._ck(self.L.mcx_get_wall_grids(self.h, C.c_void_p(out.ctypes.data), n))
        return out[:n]

    def num_molecules(self):
        return int(self.L.mcx_num_molecules(self.h))

    def download(self, capacity=None):
        cap = int(capacity if capacity is not None else self.num_molecules() + 16)
        m = MolArrays(cap)
        v = m.view()
        self._ck(self.L.mcx_download_molecules(self.h, C.byref(v), cap))
        return m.truncated(int(v.n))

    def download_into(self, mols):
        v = mols.view()
        self._ck(self.L.mcx_download_molecules(self.h, C.byref(v), len(mols.x)))
        mols.n = int(v.n)
        return mols.n

    def step(self, n_iterations=1):
        st = abi.mcx_step_stats()
        self._ck(self.L.mcx_step(self.h, n_iterations, C.byref(st)))
        return st

    def trace_step(self, n_ids):
        tr = np.zeros(n_ids, dtype=abi.TRACE_DTYPE)
        st = abi.mcx_step_stats()
        self._ck(self.L.mcx_trace_step(self.h, n_ids, _vp(tr), C.byref(st)))
        return tr, st

    def replay_step(self, words, offsets):
        words = np.ascontiguousarray(words, np.uint32)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        tr = np.zeros(len(offsets), dtype=abi.TRACE_DTYPE)
        st = abi.mcx_step_stats()
        self._ck(self.L.mcx_replay_step(self.h, _vp(words), len(words), _vp(offsets), len(offsets), _vp(tr), C.byref(st)))
        return tr, st

    def counts(self):
        s = np.zeros(max(1, self.t.n_species), np.uint64)
        r = np.zeros(max(1, self.t.n_rules), np.uint64)
        self._ck(self.L.mcx_counts(self.h, _vp(s), self.t.n_species, _vp(r), self.t.n_rules))
        return s[:self.t.n_species], r[:self.t.n_rules]


def _counts_by_volume(self):
    """(molecules[species, counted volume], reactions[rule, counted volume]) — count terms restricted to a volume."""
    ncv = max(1, getattr(self.t, "n_counted_volumes", 1))
    m = np.zeros((max(1, self.t.n_species), ncv), np.uint64)
    r = np.zeros((max(1, self.t.n_rules), ncv), np.uint64)
    self._ck(self.L.mcx_counts_by_volume(self.h, _vp(m), _vp(r)))
    return m, r


Engine.counts_by_volume = _counts_by_volume


def _counts_by_surface_region(self):
    """(surface molecules[species, region set], reactions initiated by surface molecules[rule, region set]); a region
    expression is evaluated over Tables.region_sets (region_count below)."""
    nrs = max(1, getattr(self.t, "n_region_sets", 1))
    m = np.zeros((max(1, self.t.n_species), nrs), np.uint64)
    r = np.zeros((max(1, self.t.n_rules), nrs), np.uint64)
    self._ck(self.L.mcx_counts_by_surface_region(self.h, _vp(m), _vp(r)))
    return m, r


Engine.counts_by_surface_region = _counts_by_surface_region


def region_count(tables, per_set, region_index):
    """Sum of a [row, region set] count table over the sets that contain one region: the count 'on region R'."""
    cols = [k for k, s_ in enumerate(tables.region_sets) if region_index in s_]
    return per_set[:, cols].sum(axis=1)


def philox_block(seed, mol_id, iteration, block):
    out = np.zeros(4, np.uint32)
    load_library().mcx_philox_block(seed, mol_id, iteration, block, _vp(out))
    return out
