#!/usr/bin/env python3
"""Generate tests/golden/mcell3_ref_vectors.npz: outputs of the REFERENCE's own compiled arithmetic
(oracle/_ref/libmcell3ref.so, built by `make -C oracle ref` from /root/reference/src/*.c) on the deterministic
cases of mcell3_cases.py.  Run in the build container only; the .npz is committed and checked on every box."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import mcell3_cases as mc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

O.build()
R3 = O.ref_mcell3_lib()
assert R3 is not None


def vp(a):
    return C.c_void_p(a.ctypes.data)


tris = mc.triangles()
consts = np.zeros((len(tris), 16))
for i, t in enumerate(tris):
    R3.ref3_init_tri_wall(vp(tris[i]), vp(consts[i]))

rays = mc.wall_rays(tris)
ray_out = np.zeros((len(rays), 9))   # code, t, hit(3), move_out(3), words_used
for i, (ti, p, m) in enumerate(rays):
    p = np.ascontiguousarray(p, dtype=np.float64); m = np.ascontiguousarray(m, dtype=np.float64).copy()
    t = C.c_double(0); hit = np.zeros(3); used = C.c_longlong(0)
    code = R3.ref3_collide_wall(vp(p), vp(m), vp(tris[ti]), 77, i % 13, C.byref(t), vp(hit), C.byref(used))
    valid = code in (1, 2)   # COLLIDE_FRONT / COLLIDE_BACK
    ray_out[i] = [code, t.value if valid else 0.0] + (list(hit) if valid else [0, 0, 0]) + list(m) + [used.value]

p, mv, tg, Rr = mc.mol_pairs()
mol_out = np.zeros((len(p), 5))
for i in range(len(p)):
    t = C.c_double(0); hit = np.zeros(3)
    code = R3.ref3_collide_mol(vp(p[i]), vp(mv[i]), vp(tg[i]), Rr, C.byref(t), vp(hit))   # rows of C-contiguous arrays
    hitf = code == 3  # COLLIDE_VOL_M
    mol_out[i] = [1.0 if hitf else 0.0, t.value if hitf else 0.0] + (list(hit) if hitf else [0, 0, 0])

bx = mc.boxes(tris)
box_out = np.zeros(len(bx), np.int32)
for i, (ti, lo, hi) in enumerate(bx):
    lo = np.ascontiguousarray(lo, dtype=np.float64); hi = np.ascontiguousarray(hi, dtype=np.float64)
    box_out[i] = R3.ref3_wall_in_box(vp(tris[ti]), vp(lo), vp(hi)) != 0

rc = mc.rxn_cases()
no_rx = R3.ref3_rx_no_rx()
rxn_out = np.zeros((len(rc), 4), np.int64)   # bimol pathway(-1 none), words, intersect result (-1 none), words
for i, (cum, scaling, seed, skip) in enumerate(rc):
    used = C.c_longlong(0)
    c2 = np.ascontiguousarray(cum, dtype=np.float64).copy()
    r = R3.ref3_test_bimolecular(vp(c2), len(c2), scaling, seed, skip, C.byref(used))
    rxn_out[i, 0] = -1 if r == no_rx else r; rxn_out[i, 1] = used.value
    r = R3.ref3_test_intersect(vp(c2), len(c2), scaling, seed, skip, C.byref(used))
    rxn_out[i, 2] = -1 if r == no_rx else r; rxn_out[i, 3] = used.value

# compute_pb_factor for two volume reactants (react_util.c:163-181)
pb = []
for (ssa, ssb, ta, tb) in [(2.0, 2.0, 0, 0), (2.0, 1.4142135623730951, 0, 0), (2.0, 2.0, 1, 0), (2.828, 0.7, 0, 1), (0.0, 2.0, 0, 0)]:
    pb.append(R3.ref3_compute_pb_factor_volvol(1e-6, 0.01, 10000.0, 0.005641895835477563, ssa, 1.0, ssb, 1.0, ta, tb))

dist_cases = np.array([[1.0, 1.0 + 1e-13, 1e-12], [1.0, 1.0 + 3e-12, 1e-12], [1e6, 1e6 + 1e-7, 1e-12], [1e6, 1e6 + 1e-5, 1e-12],
                       [0.0, 1e-13, 1e-12], [0.0, 2e-12, 1e-12], [-5.0, 5.0, 1e-12], [1e-20, -1e-20, 1e-12], [0.3, 0.3, 1e-12]])
dist_out = np.array([R3.ref3_distinguishable(a, b, e) for a, b, e in dist_cases], np.int32)

# exact_disk (src/diffuse.c:1365) and the surface-grid arithmetic (src/grid_util.c)
R3.ref3_exact_disk.restype = C.c_double
R3.ref3_exact_disk.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
dc, Rd = mc.disk_cases()
disk_out = np.array([R3.ref3_exact_disk(vp(loc), vp(mv), Rd, vp(tg), len(walls), vp(walls)) for loc, mv, tg, walls in dc])
grid_tris = [i for i in range(len(tris)) if np.linalg.norm(np.cross(tris[i][3:6] - tris[i][0:3], tris[i][6:9] - tris[i][0:3])) > 0]
grid_consts = np.zeros((len(tris), 8))
for i in grid_tris:
    R3.ref3_grid_constants(vp(tris[i]), vp(grid_consts[i]))
gp = mc.grid_points(tris)
grid_idx = np.array([R3.ref3_xyz2grid(vp(tris[ti]), vp(pt)) for ti, pt in gp], np.int64)
grid_uv = np.zeros((len(gp), 2)); grid_xyz = np.zeros((len(gp), 3))
for k, (ti, pt) in enumerate(gp):
    R3.ref3_grid2uv(vp(tris[ti]), int(grid_idx[k]), vp(grid_uv[k]))
    R3.ref3_uv2xyz(vp(tris[ti]), vp(grid_uv[k]), vp(grid_xyz[k]))
print("disk: blocked %d full %d partial %d" % ((disk_out < 0).sum(), (disk_out == 1).sum(), ((disk_out >= 0) & (disk_out != 1)).sum()),
      " grid points:", len(gp), "max tiles", int(grid_consts[:, 7].max()))

np.savez_compressed(os.path.join(HERE, "mcell3_ref_vectors.npz"), disk_out=disk_out, grid_consts=grid_consts, grid_idx=grid_idx,
                    grid_uv=grid_uv, grid_xyz=grid_xyz, wall_constants=consts, ray_out=ray_out, mol_out=mol_out,
                    box_out=box_out, rxn_out=rxn_out, pb_factor=np.array(pb), dist_cases=dist_cases, dist_out=dist_out)
codes, cnt = np.unique(ray_out[:, 0], return_counts=True)
print("rays:", dict(zip(codes.tolist(), cnt.tolist())), " mol hits:", int(mol_out[:, 0].sum()), "/", len(mol_out),
      " boxes in:", int(box_out.sum()), "/", len(box_out), " rxn fired:", int((rxn_out[:, 0] >= 0).sum()), "/", len(rxn_out))
print("pb_factor:", pb)
