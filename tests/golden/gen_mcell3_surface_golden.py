#!/usr/bin/env python3
"""Generate tests/golden/mcell3_surface_vectors.npz from the REFERENCE's own compiled code (oracle/_ref/libmcell3ref.so:
surface_net, init_edge_transform, find_edge_point, traverse_surface of src/wall_util.c and ray_trace_2D of
src/diffuse.c, built unmodified by
`make -C oracle ref`) on the cases of mcell3_surface_cases.py.  Build container only; the .npz is committed."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import mcell3_surface_cases as sc  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

O.build()
R3 = O.ref_mcell3_lib()
assert R3 is not None
out = {}
for name, (v, f) in sc.meshes().items():
    nb, fw, tr = O.mesh_edges(R3.ref3_mesh_edges, v, f)
    qw, qs, quv = sc.traverse_queries(len(f))
    tw, tuv = O.traverse_surface(R3.ref3_traverse_surface, v, f, qw, qs, quv)
    out[name + "_nb"], out[name + "_fw"], out[name + "_tr"], out[name + "_tw"], out[name + "_tuv"] = nb, fw, tr, tw, tuv
    rw, ruv, rdisp = sc.ray_queries(v, f)
    out[name + "_rayw"], out[name + "_rayuv"] = O.ray_trace_surf(R3.ref3_ray_trace_2d, v, f, rw, ruv, rdisp)
    print(name, "walls", len(f), "paired sides", int((nb >= 0).sum()), "free", int((nb < 0).sum()))
tris = sc.triangles()
moves = sc.edge_moves(tris)
code = np.zeros(len(moves), np.int32)
pt = np.zeros((len(moves), 2))
for i, (ti, loc, disp) in enumerate(moves):
    code[i], pt[i] = O.find_edge_point(R3.ref3_find_edge_point, tris[ti], loc, disp)
out["fep_code"], out["fep_pt"] = code, pt
print("find_edge_point codes", {int(c): int((code == c).sum()) for c in np.unique(code)})
big = sc.triangles(seed=14, n=40) * 5.0      # larger walls: tens of tiles per side
pts = sc.uv_points(big)
out["uv2grid"] = np.array([R3.ref3_uv2grid(C.c_void_p(big[ti].ctypes.data), C.c_void_p(uv.ctypes.data)) for ti, uv in pts], np.int32)
print("uv2grid tiles up to", int(out["uv2grid"].max()))
np.savez_compressed(os.path.join(HERE, "mcell3_surface_vectors.npz"), **out)
