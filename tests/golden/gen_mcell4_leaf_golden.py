#!/usr/bin/env python3
"""Generate tests/golden/mcell4_leaf_vectors.npz: outputs of MCell4's OWN compiled leaf arithmetic
(oracle/_ref/libmcell4leaf.so = src4/collision_utils.inl:464-914,1711-1747 and wall.cpp:281-342 compiled unmodified,
oracle/ref_mcell4_leaf_shim.cpp) on the cases of mcell3_cases.py / mcell4_leaf_cases.py.  Run in the build container
only; the .npz is committed and checked on every box."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import mcell3_cases as mc  # noqa: E402
import mcell4_leaf_cases as lc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

O.build()
R4 = O.ref_mcell4_leaf_lib()
assert R4 is not None


def vp(a):
    return C.c_void_p(a.ctypes.data)


tris = mc.triangles()
consts = np.zeros((len(tris), 16))
for i in range(len(tris)):
    R4.ref4_wall_constants(vp(tris[i]), vp(consts[i]))

rays = mc.wall_rays(tris)
ray_out = np.zeros((len(rays), 9))   # code, t, hit(3), move_out(3), words_used
for i, (ti, p, m) in enumerate(rays):
    p = np.ascontiguousarray(p, dtype=np.float64); m = np.ascontiguousarray(m, dtype=np.float64).copy()
    t = C.c_double(0); hit = np.zeros(3); used = C.c_longlong(0)
    code = R4.ref4_collide_wall(vp(p), vp(m), vp(tris[ti]), 77, i % 13, C.byref(t), vp(hit), C.byref(used))
    valid = code in (1, 2)
    ray_out[i] = [code, t.value if valid else 0.0] + (list(hit) if valid else [0, 0, 0]) + list(m) + [used.value]

p, mv, tg, Rr = mc.mol_pairs()
mol_out = np.zeros((len(p), 5))
for i in range(len(p)):
    t = C.c_double(0); hit = np.zeros(3)
    code = R4.ref4_collide_mol(vp(p[i]), vp(mv[i]), vp(tg[i]), Rr, C.byref(t), vp(hit))
    hitf = code == 3
    mol_out[i] = [1.0 if hitf else 0.0, t.value if hitf else 0.0] + (list(hit) if hitf else [0, 0, 0])

mesh_out = np.array([O.ref4_closest_wall_and_reflect(R4, lc.meshes()[mi], pos, move, last, seed, skip)
                     for mi, pos, move, last, seed, skip in lc.mesh_rays()])
# pick_surf_displacement (diffusion_utils.inl:60-93): displacement and words drawn
surf_cases = lc.surf_displacement_cases()
surf_out = np.zeros((len(surf_cases), 3))
R4.ref4_pick_surf_displacement.restype = C.c_longlong
R4.ref4_pick_surf_displacement.argtypes = [C.c_double, C.c_uint, C.c_uint, C.c_void_p]
for i, (scale, seed, skip) in enumerate(surf_cases):
    o2 = np.zeros(2)
    used = R4.ref4_pick_surf_displacement(scale, seed, skip, vp(o2))
    surf_out[i] = [o2[0], o2[1], used]

# test_bimolecular (rxn_utils.inl:336-414) on the cases of the MCell3 golden
rc = mc.rxn_cases()
R4.ref4_test_bimolecular.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint, C.c_uint, C.c_void_p]
bimol_out = np.zeros((len(rc), 2), np.int64)
for i, (cum, scaling, seed, skip) in enumerate(rc):
    used = C.c_longlong(0)
    c2 = np.ascontiguousarray(cum, dtype=np.float64).copy()
    bimol_out[i] = [R4.ref4_test_bimolecular(vp(c2), len(c2), scaling, seed, skip, C.byref(used)), used.value]

# exact_disk (exact_disk_utils.inl:840-1145 and its helpers) on the cases of the MCell3 golden
R4.ref4_exact_disk.restype = C.c_double
R4.ref4_exact_disk.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
dcases, Rd = mc.disk_cases()
disk_out = np.array([R4.ref4_exact_disk(vp(loc), vp(mv), Rd, vp(tg), len(walls), vp(walls)) for loc, mv, tg, walls in dcases])

# surface grids (Grid::initialize, xyz2grid_tile_index, grid2uv, uv2grid_tile_index) and find_edge_point on the cases of
# the MCell3 goldens
import mcell3_surface_cases as sc  # noqa: E402
grid_tris = [i for i in range(len(tris)) if np.linalg.norm(np.cross(tris[i][3:6] - tris[i][0:3], tris[i][6:9] - tris[i][0:3])) > 0]
grid_consts = np.zeros((len(tris), 8))
for i in grid_tris:
    R4.ref4_grid_constants(vp(tris[i]), vp(grid_consts[i]))
gp = mc.grid_points(tris)
grid_idx = np.array([R4.ref4_xyz2grid(vp(tris[ti]), vp(pt)) for ti, pt in gp], np.int64)
grid_uv = np.zeros((len(gp), 2))
for k, (ti, pt) in enumerate(gp):
    R4.ref4_grid2uv(vp(tris[ti]), int(grid_idx[k]), vp(grid_uv[k]))
stris = sc.triangles()
moves = sc.edge_moves(stris)
fep_code = np.zeros(len(moves), np.int32); fep_pt = np.zeros((len(moves), 2))
for i, (ti, loc, disp) in enumerate(moves):
    fep_code[i], fep_pt[i] = O.find_edge_point(R4.ref4_find_edge_point, stris[ti], loc, disp)
big = sc.triangles(seed=14, n=40) * 5.0
uv2grid = np.array([R4.ref4_uv2grid(vp(big[ti]), vp(uv)) for ti, uv in sc.uv_points(big)], np.int32)

np.savez_compressed(os.path.join(HERE, "mcell4_leaf_vectors.npz"), wall_constants=consts, ray_out=ray_out, mol_out=mol_out,
                    mesh_out=mesh_out, surf_out=surf_out, bimol_out=bimol_out, grid_consts=grid_consts,
                    grid_idx=grid_idx, grid_uv=grid_uv, fep_code=fep_code, fep_pt=fep_pt, uv2grid=uv2grid, disk_out=disk_out)
print("wrote mcell4_leaf_vectors.npz:", consts.shape, ray_out.shape, mol_out.shape, mesh_out.shape,
      "hits", int(mesh_out[:, 0].sum()), "with redo words", int((mesh_out[:, 17] > 0).sum()))
