"""Deterministic cases for pinning the oracle's volume-path leaf arithmetic against MCell4's own compiled code
(oracle/_ref/libmcell4leaf.so): single-triangle rays, molecule pairs and wall constants are the cases of
mcell3_cases.py (so the MCell3 and MCell4 outputs can be compared with each other as well); mesh_rays() adds whole
meshes for get_closest_wall_collision + reflect_from_wall: random rays, rays aimed at vertices and at edge midpoints
(the REDO / jump_away_line paths, which restart the walk with a perturbed displacement), rays that skip a
last-hit wall."""
import numpy as np

from mcell_b200.model import create_box, create_icosphere


def meshes():
    out = []
    v, f = create_box(0.6)
    out.append((np.ascontiguousarray(v / 0.01, np.float64), np.ascontiguousarray(f, np.uint32)))          # 12 walls, 60 lu
    v, f = create_icosphere(0.3, 2)
    out.append((np.ascontiguousarray(v / 0.01, np.float64), np.ascontiguousarray(f, np.uint32)))          # 80 walls
    v, f = create_icosphere(0.25, 3)
    out.append((np.ascontiguousarray(v / 0.01 + np.array([3.0, -2.0, 1.5]), np.float64), np.ascontiguousarray(f, np.uint32)))  # 320
    return out


def mesh_rays(seed=21, per_mesh=700):
    """-> list of (mesh index, pos(3), move(3), last_hit_wall, rng seed, rng skip)"""
    rng = np.random.default_rng(seed)
    cases = []
    for mi, (v, f) in enumerate(meshes()):
        size = np.abs(v).max()
        for k in range(per_mesh):
            kind = k % 7
            pos = rng.uniform(-0.55, 0.55, 3) * size + v.mean(axis=0)
            if kind <= 2:                                   # random ray, often leaving the mesh
                move = rng.normal(size=3) * size * rng.uniform(0.05, 1.2)
            elif kind == 3:                                 # through a vertex
                target = v[rng.integers(0, len(v))]
                move = (target - pos) * rng.uniform(1.0, 2.0)
            elif kind == 4:                                 # through an edge midpoint
                t = f[rng.integers(0, len(f))]
                e = rng.integers(0, 3)
                target = 0.5 * (v[t[e]] + v[t[(e + 1) % 3]])
                move = (target - pos) * rng.uniform(1.0, 2.0)
            elif kind == 5:                                 # through a point on an edge
                t = f[rng.integers(0, len(f))]
                e = rng.integers(0, 3)
                a = rng.uniform(0, 1)
                target = v[t[e]] + a * (v[t[(e + 1) % 3]] - v[t[e]])
                move = (target - pos) * rng.uniform(1.0, 1.5)
            else:                                           # short move that stays inside
                move = rng.normal(size=3) * size * 0.02
            last = 0xFFFFFFFF if k % 5 else int(rng.integers(0, len(f)))
            cases.append((mi, np.ascontiguousarray(pos), np.ascontiguousarray(move), last, 31 + (k % 9), k % 11))
    return cases


def surf_displacement_cases(n=3000, seed=22):
    """-> list of (scale, rng seed, rng skip) for pick_surf_displacement"""
    rng = np.random.default_rng(seed)
    return [(float(rng.uniform(0.01, 5.0)), int(rng.integers(1, 50)), int(k % 97)) for k in range(n)]
