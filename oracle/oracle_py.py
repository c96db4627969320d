"""ctypes driver of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  It takes the same Tables object (mcell_b200.model.Model.build()) as the product
so both sides see identical inputs."""
import ctypes as C
import os
import subprocess

import numpy as np

from mcell_b200 import abi
from mcell_b200.model import MolArrays

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(native=False, force=False):
    """force: rebuild even when up to date (-march=native code must be compiled on the machine that runs it)."""
    target = "liboracle_native.so" if native else "liboracle.so"
    subprocess.run(["make", "-s", "-C", _HERE] + (["-B"] if force else []) + [target], check=True)
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)
    return os.path.join(_HERE, target)


def lib(native=False):
    global _LIB
    key = "native" if native else "core2"
    if _LIB is None:
        _LIB = {}
    if key not in _LIB:
        path = os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")
        if not os.path.exists(path):
            build(native)
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(abi.mcx_config)]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_num_molecules.restype = C.c_uint64
        L.orc_num_molecules.argtypes = [C.c_void_p]
        L.orc_tape_size.restype = C.c_uint64
        L.orc_tape_size.argtypes = [C.c_void_p]
        L.orc_subpart_wall_count.restype = C.c_uint64
        L.orc_subpart_wall_count.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_rng_new.restype = C.c_void_p
        L.orc_rng_dbl.restype = C.c_double
        L.orc_rng_gauss.restype = C.c_double
        L.orc_rng_uint.restype = C.c_uint32
        L.orc_rng_uses.restype = C.c_longlong
        for f in ("orc_rng_free", "orc_rng_uint", "orc_rng_dbl", "orc_rng_gauss", "orc_rng_uses"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_rng_fill_uint.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        L.orc_rng_fill_gauss.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        _LIB[key] = L
    return _LIB[key]


def ref_rng_lib():
    """The reference's own RNG (src/rng.c) compiled into oracle/_ref/ (build container only,
    travels prebuilt to the GPU box)."""
    path = os.path.join(_HERE, "_ref", "librefrng.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_rng_new.restype = C.c_void_p
    L.ref_rng_dbl.restype = C.c_double
    L.ref_rng_gauss.restype = C.c_double
    L.ref_rng_uint.restype = C.c_uint32
    L.ref_rng_uses.restype = C.c_longlong
    for f in ("ref_rng_free", "ref_rng_uint", "ref_rng_dbl", "ref_rng_gauss", "ref_rng_uses"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.ref_rng_fill_uint.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    L.ref_rng_fill_gauss.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    return L


class Oracle:
    SEQUENTIAL, SNAPSHOT = 0, 1

    def __init__(self, tables, native=False):
        self.L = lib(native)
        self.t = tables
        self.h = C.c_void_p(self.L.orc_create(C.byref(tables.cfg)))
        t = tables
        self._v = lambda x: C.c_void_p(x.ctypes.data) if x is not None and x.size else None
        self.L.orc_set_species(self.h, t.species, C.c_uint32(t.n_species))
        if self.L.orc_set_reactions(self.h, t.classes, C.c_uint32(t.n_classes), t.pathways, C.c_uint32(t.n_pathways)):
            raise RuntimeError(self.error())   # a table the oracle (like the product) refuses
        self.L.orc_set_surface_classes(self.h, t.surf_rules, C.c_uint32(t.n_surf_rules))
        self.L.orc_set_geometry(self.h, self._v(t.vertices), C.c_uint64(len(t.vertices)), self._v(t.tri),
                                C.c_uint64(len(t.tri)), self._v(t.wall_surf_class), self._v(getattr(t, "wall_object", None)))
        if getattr(t, "n_counted_volumes", 0) > 1:
            self.L.orc_set_counted_volumes(self.h, C.c_uint32(t.n_counted_volumes), self._v(t.wall_cv_front), self._v(t.wall_cv_back))
        if getattr(t, "n_counted_volumes", 0) > 1 and getattr(t, "cv_object_mask", None) is not None:
            self.L.orc_set_counted_volume_objects(self.h, self._v(t.cv_object_mask), C.c_uint32(t.cv_intersecting))
        if getattr(t, "n_region_sets", 0) > 1:
            self.L.orc_set_surface_regions(self.h, C.c_uint32(t.n_region_sets), self._v(t.wall_region_set))
        if getattr(t, "wall_edge_border", None) is not None:
            if self.L.orc_set_region_borders(self.h, self._v(t.wall_edge_border)):
                raise RuntimeError(self.error())

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def error(self):
        return self.L.orc_last_error(self.h).decode()

    def upload(self, mols):
        v = mols.view()
        rc = self.L.orc_upload_molecules(self.h, C.byref(v))
        if rc:
            raise RuntimeError(self.error())

    def release(self, species, number, location, diameter, shape=0, release_time=0.0, counted_volume_index=0, region_in=0, region_out=0, region_expr=()):
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
        rc = self.L.orc_release_volume_molecules(self.h, C.byref(r), C.byref(first))
        if rc:
            raise RuntimeError(self.error())
        return int(first.value)

    def release_surface(self, species, number, walls, orientation=1, release_time=0.0, randomize_pos=True):
        """ReleaseEvent::release_onto_regions on the device (mcx_release_surface_molecules): `number` molecules of a
        surface species on vacant tiles of the listed walls; returns the first id."""
        wl = np.ascontiguousarray(walls, np.uint32)
        r = abi.mcx_surface_release()
        r.species, r.orientation, r.number, r.release_time = int(species), int(orientation), int(number), float(release_time)
        r.walls, r.n_walls, r.randomize_pos = wl.ctypes.data, len(wl), 1 if randomize_pos else 0
        first = C.c_uint32(0)
        rc = self.L.orc_release_surface_molecules(self.h, C.byref(r), C.byref(first))
        if rc:
            raise RuntimeError(self.error())
        return int(first.value)

    def release_list(self, species, positions, counted_volume=None, release_time=0.0):
        sp = np.ascontiguousarray(species, np.uint32)
        pos = np.asarray(positions, np.float64)
        x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
        cv = None if counted_volume is None else np.ascontiguousarray(counted_volume, np.uint32)
        first = C.c_uint32(0)
        rc = self.L.orc_release_list(self.h, C.c_uint64(len(sp)), self._v(sp), self._v(x), self._v(y), self._v(z),
                                     self._v(cv) if cv is not None else None, C.c_double(release_time), C.byref(first))
        if rc:
            raise RuntimeError(self.error())
        return int(first.value)

    def next_molecule_id(self, set_to=None):
        if set_to is not None:
            self.L.orc_set_next_molecule_id(self.h, C.c_uint32(int(set_to)))
        out = C.c_uint32(0)
        self.L.orc_get_next_molecule_id(self.h, C.byref(out))
        return int(out.value)

    def wall_grids(self, set_to=None):
        n = len(self.t.tri)
        if set_to is not None:
            a = np.ascontiguousarray(set_to, dtype=np.uint8)
            self.L.orc_set_wall_grids(self.h, C.c_void_p(a.ctypes.data), C.c_uint64(n))
        out = np.zeros(max(n, 1), np.uint8)
        self.L.orc_get_wall_grids(self.h, C.c_void_p(out.ctypes.data), C.c_uint64(n))
        return out[:n]

    def num_molecules(self):
        return int(self.L.orc_num_molecules(self.h))

    def download(self):
        n = self.num_molecules()
        m = MolArrays(n)
        v = m.view()
        rc = self.L.orc_download_molecules(self.h, C.byref(v), C.c_uint64(n))
        assert rc == 0
        m.n = int(v.n)
        return m

    def step(self, n_iterations=1, mode=0):
        st = abi.mcx_step_stats()
        rc = self.L.orc_step(self.h, C.c_uint32(n_iterations), C.c_int(mode), C.byref(st))
        if rc:
            raise RuntimeError(self.error())
        return st

    def trace_step(self, mode, n_trace, words=None, offsets=None):
        """mode 0 sequential(+tape recording), 1 snapshot/Philox, 2 snapshot replaying a tape."""
        tr = np.zeros(n_trace, dtype=abi.TRACE_DTYPE)
        st = abi.mcx_step_stats()
        nw = 0 if words is None else len(words)
        ni = 0 if offsets is None else len(offsets)
        rc = self.L.orc_trace_step(self.h, C.c_int(mode), self._v(words), C.c_uint64(nw), self._v(offsets),
                                   C.c_uint64(ni), C.c_void_p(tr.ctypes.data), C.c_uint64(n_trace), C.byref(st))
        if rc:
            raise RuntimeError(self.error())
        return tr, st

    def trace_sample(self, stride, offset, n_trace):
        """First evaluation of the molecules with id % stride == offset against the current state (Philox streams); the
        state is not changed (orc_trace_sample)."""
        tr = np.zeros(n_trace, dtype=abi.TRACE_DTYPE)
        rc = self.L.orc_trace_sample(self.h, C.c_uint32(stride), C.c_uint32(offset), C.c_void_p(tr.ctypes.data), C.c_uint64(n_trace))
        if rc:
            raise RuntimeError(self.error())
        return tr

    def tape(self, n_ids):
        n = int(self.L.orc_tape_size(self.h))
        words = np.zeros(max(n, 1), np.uint32)
        off = np.zeros(n_ids, np.uint64)
        ln = np.zeros(n_ids, np.uint32)
        self.L.orc_tape_get(self.h, C.c_void_p(words.ctypes.data), C.c_void_p(off.ctypes.data),
                            C.c_void_p(ln.ctypes.data), C.c_uint64(n_ids))
        return words[:n], off, ln

    def counts(self):
        s = np.zeros(max(1, self.t.n_species), np.uint64)
        r = np.zeros(max(1, self.t.n_rules), np.uint64)
        self.L.orc_counts(self.h, C.c_void_p(s.ctypes.data), C.c_uint32(self.t.n_species),
                          C.c_void_p(r.ctypes.data), C.c_uint32(self.t.n_rules))
        return s[:self.t.n_species], r[:self.t.n_rules]

    def counts_by_volume(self):
        """(molecules[species, counted volume], reactions[rule, counted volume])"""
        ncv = max(1, getattr(self.t, "n_counted_volumes", 1))
        m = np.zeros((max(1, self.t.n_species), ncv), np.uint64)
        r = np.zeros((max(1, self.t.n_rules), ncv), np.uint64)
        self.L.orc_counts_by_volume(self.h, C.c_void_p(m.ctypes.data), C.c_void_p(r.ctypes.data))
        return m, r

    def counts_by_surface_region(self):
        """(surface molecules[species, region set], reactions initiated by surface molecules[rule, region set])"""
        nrs = max(1, getattr(self.t, "n_region_sets", 1))
        m = np.zeros((max(1, self.t.n_species), nrs), np.uint64)
        r = np.zeros((max(1, self.t.n_rules), nrs), np.uint64)
        rc = self.L.orc_counts_by_surface_region(self.h, C.c_void_p(m.ctypes.data), C.c_void_p(r.ctypes.data))
        if rc:
            raise RuntimeError("orc_counts_by_surface_region: %d" % rc)
        return m, r

    def subpart_walls(self, subpart):
        n = int(self.L.orc_subpart_wall_count(self.h, C.c_uint32(subpart)))
        out = np.zeros(max(n, 1), np.uint32)
        self.L.orc_subpart_walls(self.h, C.c_uint32(subpart), C.c_void_p(out.ctypes.data))
        return out[:n]

    def wall_constants(self, wi):
        out = np.zeros(16)
        self.L.orc_wall_constants(self.h, C.c_uint32(wi), C.c_void_p(out.ctypes.data))
        return out


def ref_mcell3_lib():
    """The reference's own wall/collision/reaction arithmetic (MCell3 originals of the MCell4 hot-path
    functions) compiled into oracle/_ref/libmcell3ref.so by `make -C oracle ref` (build container only;
    the .so travels prebuilt to the GPU box)."""
    path = os.path.join(_HERE, "_ref", "libmcell3ref.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref3_compute_pb_factor_volvol.restype = C.c_double
    L.ref3_compute_pb_factor_volvol.argtypes = [C.c_double] * 8 + [C.c_int, C.c_int]
    L.ref3_distinguishable.argtypes = [C.c_double, C.c_double, C.c_double]
    L.ref3_collide_mol.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    L.ref3_collide_wall.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref3_test_bimolecular.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint, C.c_uint, C.c_void_p]
    L.ref3_test_intersect.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint, C.c_uint, C.c_void_p]
    L.ref3_binary_search_double.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double]
    if hasattr(L, "ref3_exact_disk"):
        L.ref3_exact_disk.restype = C.c_double
        L.ref3_exact_disk.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
        L.ref3_grid_constants.argtypes = [C.c_void_p, C.c_void_p]
        L.ref3_xyz2grid.argtypes = [C.c_void_p, C.c_void_p]
        L.ref3_grid2uv.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref3_uv2xyz.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def mesh_edges(fn, verts, tris):
    """(neighbour wall per side or -1, is-forward flag, [cos, sin, tu, tv] per side) from ref3_mesh_edges or
    orc_unit_mesh_edges."""
    verts = np.ascontiguousarray(verts, np.float64); tris = np.ascontiguousarray(tris, np.uint32)
    nw = len(tris)
    nb = np.zeros(3 * nw, np.int32); fw = np.zeros(3 * nw, np.int32); tr = np.zeros((3 * nw, 4))
    rc = fn(C.c_void_p(verts.ctypes.data), C.c_uint(len(verts)), C.c_void_p(tris.ctypes.data), C.c_uint(nw),
            C.c_void_p(nb.ctypes.data), C.c_void_p(fw.ctypes.data), C.c_void_p(tr.ctypes.data))
    assert rc == 0
    return nb, fw, tr


def traverse_surface(fn, verts, tris, q_wall, q_side, q_uv):
    verts = np.ascontiguousarray(verts, np.float64); tris = np.ascontiguousarray(tris, np.uint32)
    q_wall = np.ascontiguousarray(q_wall, np.uint32); q_side = np.ascontiguousarray(q_side, np.int32)
    q_uv = np.ascontiguousarray(q_uv, np.float64)
    w = np.zeros(len(q_wall), np.int32); uv = np.zeros((len(q_wall), 2))
    rc = fn(C.c_void_p(verts.ctypes.data), C.c_uint(len(verts)), C.c_void_p(tris.ctypes.data), C.c_uint(len(tris)),
            C.c_void_p(q_wall.ctypes.data), C.c_void_p(q_side.ctypes.data), C.c_void_p(q_uv.ctypes.data), C.c_uint(len(q_wall)),
            C.c_void_p(w.ctypes.data), C.c_void_p(uv.ctypes.data))
    assert rc == 0
    return w, uv


def ray_trace_surf(fn, verts, tris, q_wall, q_uv, q_disp):
    """(end wall or -1, end uv) per query from ref3_ray_trace_2d or orc_unit_ray_trace_surf."""
    verts = np.ascontiguousarray(verts, np.float64); tris = np.ascontiguousarray(tris, np.uint32)
    q_wall = np.ascontiguousarray(q_wall, np.uint32)
    q_uv = np.ascontiguousarray(q_uv, np.float64); q_disp = np.ascontiguousarray(q_disp, np.float64)
    w = np.zeros(len(q_wall), np.int32); uv = np.zeros((len(q_wall), 2))
    rc = fn(C.c_void_p(verts.ctypes.data), C.c_uint(len(verts)), C.c_void_p(tris.ctypes.data), C.c_uint(len(tris)),
            C.c_void_p(q_wall.ctypes.data), C.c_void_p(q_uv.ctypes.data), C.c_void_p(q_disp.ctypes.data), C.c_uint(len(q_wall)),
            C.c_void_p(w.ctypes.data), C.c_void_p(uv.ctypes.data))
    assert rc == 0
    return w, uv


def find_edge_point(fn, v9, loc, disp):
    v9 = np.ascontiguousarray(v9, np.float64); loc = np.ascontiguousarray(loc, np.float64); disp = np.ascontiguousarray(disp, np.float64)
    pt = np.zeros(2)
    code = fn(C.c_void_p(v9.ctypes.data), C.c_void_p(loc.ctypes.data), C.c_void_p(disp.ctypes.data), C.c_void_p(pt.ctypes.data))
    return int(code), pt


_DDA_ARGS = [C.c_void_p, C.c_double, C.c_uint, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]


def ref_mcell4_lib():
    """MCell4's own subpartition walk (src4/collision_utils_subparts.inl) compiled unmodified into
    oracle/_ref/libmcell4ref.so by `make -C oracle ref` (oracle/ref_mcell4_shim.cpp); None where it is absent."""
    path = os.path.join(_HERE, "_ref", "libmcell4ref.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref4_collect_crossed_subparts.restype = C.c_uint
    L.ref4_collect_crossed_subparts.argtypes = _DDA_ARGS
    return L


def ref_mcell4_raytrace_lib():
    """MCell4's own ray_trace_vol (src4/diffuse_react_event.cpp:627-780) with sort_collisions_by_time and everything under
    it compiled unmodified into oracle/_ref/libmcell4raytrace.so (oracle/ref_mcell4_raytrace_shim.cpp); None where absent."""
    path = os.path.join(_HERE, "_ref", "libmcell4raytrace.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref4_ray_trace_vol.restype = C.c_int
    return L


class RayTraceScene:
    """One population for ref4_ray_trace_vol / orc_unit_ray_trace_vol: the oracle's world (tables + molecules) and the
    same data flattened for the reference shim (wall lists per subpartition taken from the oracle: that distribution is
    pinned separately against MCell4's own wall_subparts_collision_test)."""

    def __init__(self, tables, mols, cap=64):
        self.t, self.mols, self.cap = tables, mols, cap
        self.orc = Oracle(tables)
        self.orc.upload(mols)
        L = self.orc.L
        L.orc_unit_ray_trace_vol.restype = C.c_int
        cfg = tables.cfg
        self.n_sp = int(cfg.num_subparts_per_edge) ** 3
        lists = [self.orc.subpart_walls(s) for s in range(self.n_sp)]
        self.wall_off = np.zeros(self.n_sp + 1, np.uint32)
        self.wall_off[1:] = np.cumsum([len(x) for x in lists])
        self.walls = np.concatenate(lists + [np.zeros(1, np.uint32)]).astype(np.uint32)
        ns = int(tables.n_species)
        self.reacts = np.zeros((ns, ns), np.uint8)
        for k in range(int(tables.n_classes)):
            c = tables.classes[k]
            if c.kind == abi.MCX_RXN_BIMOL_VOLVOL:
                self.reacts[c.reactants[0], c.reactants[1]] = self.reacts[c.reactants[1], c.reactants[0]] = 1
        self.pos = np.ascontiguousarray(np.stack([mols.x, mols.y, mols.z], axis=1))
        self.species = np.ascontiguousarray(mols.species.astype(np.uint32))
        self.origin = np.array([cfg.origin[0], cfg.origin[1], cfg.origin[2]], np.float64)
        self.verts = np.ascontiguousarray(tables.vertices, np.float64)
        self.tri = np.ascontiguousarray(tables.tri, np.uint32)

    def wall_near(self, mol_id):
        """a wall of the molecule's own subpartition (MCX_NONE when it has none): a plausible last_hit_wall"""
        cfg = self.t.cfg
        rcp = cfg.num_subparts_per_edge / cfg.partition_edge_length
        i = ((self.pos[mol_id] - self.origin) * rcp).astype(np.int64)
        n = int(cfg.num_subparts_per_edge)
        s = int(i[0] + i[1] * n + i[2] * n * n)
        a, b = int(self.wall_off[s]), int(self.wall_off[s + 1])
        return int(self.walls[a]) if b > a else 0xFFFFFFFF

    def _out(self):
        cap = self.cap
        return (C.c_int(0), np.zeros(cap, np.int32), np.zeros(cap, np.float64), np.zeros(3 * cap, np.float64),
                np.zeros(cap, np.uint32), C.c_longlong(0))

    @staticmethod
    def _row(rc, disp, n, typ, tim, pos, what, used):
        k = n.value
        return dict(hit=rc, disp=disp.copy(), n=k, type=typ[:k].copy(), time=tim[:k].copy(), pos=pos[:3 * k].copy(),
                    what=what[:k].copy(), words=used.value)

    def oracle(self, mol_id, disp, last_hit_wall, words):
        vp = lambda a: C.c_void_p(a.ctypes.data)
        d = np.array(disp, np.float64)
        n, typ, tim, pos, what, used = self._out()
        words = np.ascontiguousarray(words, np.uint32)
        rc = self.orc.L.orc_unit_ray_trace_vol(self.orc.h, C.c_uint32(mol_id), vp(d), C.c_uint32(last_hit_wall), vp(words),
                                               C.c_uint64(len(words)), C.c_int(self.cap), C.byref(n), vp(typ), vp(tim), vp(pos),
                                               vp(what), C.byref(used))
        return self._row(rc, d, n, typ, tim, pos, what, used)

    def reference(self, R, mol_id, disp, last_hit_wall, seed, skip):
        vp = lambda a: C.c_void_p(a.ctypes.data)
        d = np.array(disp, np.float64)
        n, typ, tim, pos, what, used = self._out()
        after = np.zeros(3)
        sp_after = C.c_uint(0)
        cfg = self.t.cfg
        rc = R.ref4_ray_trace_vol(vp(self.origin), C.c_double(cfg.partition_edge_length), C.c_uint(cfg.num_subparts_per_edge),
                                  C.c_double(cfg.rxn_radius_3d), vp(self.verts), C.c_uint(len(self.verts)), vp(self.tri),
                                  C.c_uint(len(self.tri)), vp(self.wall_off), vp(self.walls), vp(self.species), vp(self.pos),
                                  C.c_uint(len(self.species)), vp(self.reacts), C.c_uint(self.reacts.shape[0]), C.c_uint(mol_id),
                                  vp(d), C.c_uint(last_hit_wall), C.c_uint(seed), C.c_uint(skip), C.c_int(self.cap), C.byref(n),
                                  vp(typ), vp(tim), vp(pos), vp(what), C.byref(used), vp(after), C.byref(sp_after))
        row = self._row(rc, d, n, typ, tim, pos, what, used)
        row["pos_after"] = after
        row["subpart_after"] = sp_after.value
        return row


_CWR_ARGS = None


def ref_mcell4_leaf_lib():
    """MCell4's own collide_mol / jump_away_line / collide_wall / get_closest_wall_collision / reflect_from_wall and
    Wall::initialize_wall_constants compiled unmodified into oracle/_ref/libmcell4leaf.so (oracle/ref_mcell4_leaf_shim.cpp);
    None where it is absent."""
    path = os.path.join(_HERE, "_ref", "libmcell4leaf.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref4_collide_wall.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref4_collide_mol.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    L.ref4_wall_constants.argtypes = [C.c_void_p, C.c_void_p]
    return L


def _cwr(fn, mesh, pos, move, last, tail_args):
    """common driver of ref4_closest_wall_and_reflect / orc_unit_closest_wall_and_reflect -> 19 numbers:
    found, wall, side, t, hit(3), pos_after(3), disp_after(3), t_steps, move_out(3), words, ray_polygon_tests"""
    v, f = mesh
    vp = lambda a: C.c_void_p(a.ctypes.data)
    move = np.ascontiguousarray(move, np.float64).copy()
    pos = np.ascontiguousarray(pos, np.float64)
    wall, side, t = C.c_uint(0), C.c_int(0), C.c_double(0)
    hit, pa, da = np.zeros(3), np.zeros(3), np.zeros(3)
    ts, used, tests = C.c_double(1.0), C.c_longlong(0), C.c_ulonglong(0)
    found = fn(vp(v), C.c_uint(len(v)), vp(f), C.c_uint(len(f)), vp(pos), vp(move), C.c_uint(last), *tail_args,
               C.byref(wall), C.byref(side), C.byref(t), vp(hit), vp(pa), vp(da), C.byref(ts), C.byref(used), C.byref(tests))
    if not found:
        return [0.0] + [0.0] * 12 + [1.0] + list(move) + [float(used.value), float(tests.value)]
    return [1.0, float(wall.value), float(side.value), t.value] + list(hit) + list(pa) + list(da) + [ts.value] + list(move) + \
           [float(used.value), float(tests.value)]


def ref4_closest_wall_and_reflect(L, mesh, pos, move, last, seed, skip):
    L.ref4_closest_wall_and_reflect.restype = C.c_int
    return _cwr(L.ref4_closest_wall_and_reflect, mesh, pos, move, last, (C.c_uint(seed), C.c_uint(skip)))


def orc_closest_wall_and_reflect(mesh, pos, move, last, words):
    L = lib()
    L.orc_unit_closest_wall_and_reflect.restype = C.c_int
    return _cwr(L.orc_unit_closest_wall_and_reflect, mesh, pos, move, last, (C.c_void_p(words.ctypes.data), C.c_uint64(len(words))))


def _collect(fn, grid, pos, disp, for_mols, for_walls, cap):
    origin, edge, n, expanded, radius = grid
    o = np.ascontiguousarray(origin, np.float64)
    pos = np.ascontiguousarray(pos, np.float64); disp = np.ascontiguousarray(disp, np.float64)
    w = np.zeros(cap, np.uint32); m = np.zeros(cap, np.uint32)
    nw, nm = C.c_uint(0), C.c_uint(0)
    d = fn(C.c_void_p(o.ctypes.data), float(edge), int(n), int(expanded), float(radius), C.c_void_p(pos.ctypes.data),
           C.c_void_p(disp.ctypes.data), int(for_mols), int(for_walls), C.c_void_p(w.ctypes.data), C.byref(nw),
           C.c_void_p(m.ctypes.data), C.byref(nm), cap)
    return int(d), w[:nw.value].copy(), m[:nm.value].copy()


def ref4_collect(L4, grid, pos, disp, for_mols, for_walls, cap=512):
    """(destination subpartition, ordered wall list, molecule set ascending) from the reference's compiled walk."""
    return _collect(L4.ref4_collect_crossed_subparts, grid, pos, disp, for_mols, for_walls, cap)


def orc_collect(grid, pos, disp, for_mols, for_walls, cap=512):
    """The same call on the oracle's restatement (molecule set in insertion order)."""
    L = lib()
    L.orc_unit_collect_crossed_subparts.restype = C.c_uint
    L.orc_unit_collect_crossed_subparts.argtypes = _DDA_ARGS
    return _collect(L.orc_unit_collect_crossed_subparts, grid, pos, disp, for_mols, for_walls, cap)


def unit_lib():
    """Typed access to the oracle's single-function entry points (orc_unit_*)."""
    L = lib()
    L.orc_unit_collide_mol.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    L.orc_unit_collide_wall.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_unit_distinguishable.argtypes = [C.c_double, C.c_double, C.c_double]
    L.orc_unit_test_bimolecular.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_uint64, C.c_void_p]
    L.orc_unit_pathway_for_probability.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.orc_unit_wall_in_box.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_unit_wall_constants.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_unit_exact_disk.restype = C.c_double
    L.orc_unit_exact_disk.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_unit_grid_constants.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_unit_xyz2grid.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_unit_grid2uv.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_unit_uv2xyz.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return L
