"""Deterministic cases for the surface-diffusion geometry: edge pairing and flattening transforms of a
mesh (surface_net + init_edge_transform), find_edge_point, traverse_surface.  Shared by the generator and the tests."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcell_b200.model import create_box, create_icosphere  # noqa: E402


def meshes():
    """name -> (vertices float64 [n,3] in length units, triangles uint32 [m,3])."""
    out = {}
    v, f = create_box(1.0)
    out["box"] = (np.ascontiguousarray(v * 100.0, np.float64), np.ascontiguousarray(f, np.uint32))
    out["open_box"] = (out["box"][0], np.ascontiguousarray(f[:-3], np.uint32))            # free edges: faces removed
    for sub in (2, 4):
        v, f = create_icosphere(0.5, sub)
        out["icosphere%d" % sub] = (np.ascontiguousarray(v * 100.0, np.float64), np.ascontiguousarray(f, np.uint32))
    rng = np.random.default_rng(5)
    v, f = create_icosphere(0.3, 3)
    v = v * 100.0 * (1.0 + 0.2 * rng.uniform(-1, 1, (len(v), 1))) + np.array([3.0, -7.0, 11.0])   # irregular triangles
    out["bumpy"] = (np.ascontiguousarray(v, np.float64), np.ascontiguousarray(f, np.uint32))
    return out


def triangles(seed=9, n=60):
    rng = np.random.default_rng(seed)
    t = rng.uniform(-20, 20, (n, 3, 3))
    t[:8] = rng.uniform(-0.05, 0.05, (8, 3, 3)) + 5.0          # tiny
    t[8:14, 2] = t[8:14, 0] + (t[8:14, 1] - t[8:14, 0]) * 0.5 + rng.normal(0, 0.01, (6, 3))  # slivers
    return np.ascontiguousarray(t.reshape(n, 9))


def edge_moves(tris, seed=10, per_tri=40):
    """(triangle index, loc[2], disp[2]) in the triangle's uv frame: starts inside, on an edge, at a vertex; moves that stay
    inside, leave through each edge, run along an edge or through a corner."""
    rng = np.random.default_rng(seed)
    out = []
    for ti, t in enumerate(tris):
        p0, p1, p2 = t[:3], t[3:6], t[6:9]
        u = (p1 - p0) / np.linalg.norm(p1 - p0)
        nrm = np.cross(u, p2 - p0)
        nrm /= np.linalg.norm(nrm)
        v = np.cross(nrm, u)
        uv = lambda p: np.array([np.dot(p - p0, u), np.dot(p - p0, v)])  # noqa: E731
        a, b, c = uv(p0), uv(p1), uv(p2)
        size = max(np.linalg.norm(b - a), np.linalg.norm(c - a))
        for k in range(per_tri):
            w = rng.dirichlet([1, 1, 1])
            kind = k % 8
            if kind == 5:
                w = np.array([w[0], 1 - w[0], 0.0])                # on edge 0
            elif kind == 6:
                w = np.array([1.0, 0.0, 0.0]) if k % 16 == 6 else np.array([0.0, 0.0, 1.0])  # at a vertex
            loc = w[0] * a + w[1] * b + w[2] * c
            if kind == 0:
                disp = rng.normal(0, 0.05 * size, 2)               # mostly stays inside
            elif kind == 7:
                disp = (b - a) * rng.uniform(-1.5, 1.5)            # parallel to edge 0
            elif kind == 4:
                disp = (c - loc) * rng.uniform(0.5, 2.0)           # through a corner
            else:
                disp = rng.normal(0, 1.0, 2) * size * rng.uniform(0.3, 3.0)
            out.append((ti, np.ascontiguousarray(loc), np.ascontiguousarray(disp)))
    return out


def traverse_queries(n_walls, seed=11, n=400):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, n_walls, n).astype(np.uint32), rng.integers(0, 3, n).astype(np.int32),
            np.ascontiguousarray(rng.uniform(-5, 15, (n, 2))))


def ray_queries(verts, tris, seed=12, n=1500):
    """(wall, start uv inside that wall, 2-D displacement of 0.05 to 4 edge lengths): most moves cross several edges; on
    the open mesh they reflect at free edges."""
    rng = np.random.default_rng(seed)
    qw = rng.integers(0, len(tris), n).astype(np.uint32)
    p0, p1, p2 = verts[tris[qw, 0]], verts[tris[qw, 1]], verts[tris[qw, 2]]
    u = (p1 - p0) / np.linalg.norm(p1 - p0, axis=1, keepdims=True)
    nrm = np.cross(u, p2 - p0)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    v = np.cross(nrm, u)
    b = np.stack([np.einsum("ij,ij->i", p1 - p0, u), np.zeros(n)], 1)
    c = np.stack([np.einsum("ij,ij->i", p2 - p0, u), np.einsum("ij,ij->i", p2 - p0, v)], 1)
    w = rng.dirichlet([1, 1, 1], n)
    uv = w[:, 1:2] * b + w[:, 2:3] * c
    size = np.linalg.norm(b, axis=1, keepdims=True)
    disp = rng.normal(0, 1, (n, 2)) * size * rng.uniform(0.05, 4.0, (n, 1))
    return qw, np.ascontiguousarray(uv), np.ascontiguousarray(disp)


def uv_points(tris, seed=13, per_tri=60):
    """(triangle index, uv) strictly inside the triangle or exactly at one of its vertices (uv2grid's special cases)."""
    rng = np.random.default_rng(seed)
    out = []
    for ti, t in enumerate(tris):
        p0, p1, p2 = t[:3], t[3:6], t[6:9]
        u = (p1 - p0) / np.linalg.norm(p1 - p0)
        nrm = np.cross(u, p2 - p0)
        nrm /= np.linalg.norm(nrm)
        v = np.cross(nrm, u)
        b = np.array([np.dot(p1 - p0, u), 0.0])
        c = np.array([np.dot(p2 - p0, u), np.dot(p2 - p0, v)])
        for k in range(per_tri):
            if k < 3:
                w = np.eye(3)[k]
                uv = w[1] * b + w[2] * c
            else:
                w = rng.dirichlet([1, 1, 1] if k % 4 else [0.2, 0.2, 0.2])     # also close to edges and corners
                uv = (w[1] * b + w[2] * c) * (1 - 1e-9) + 1e-9 * (b + c) / 3
            out.append((ti, np.ascontiguousarray(uv)))
    return out

