"""Deterministic input cases for pinning the oracle's restatement of collide_wall / collide_mol / wall_in_box /
init_tri_wall / test_bimolecular against the reference's own compiled code (oracle/_ref/libmcell3ref.so).
Shared by the golden generator (gen_mcell3_golden.py) and tests/test_oracle_vs_reference.py."""
import numpy as np


def triangles(n=300, seed=11):
    rng = np.random.default_rng(seed)
    t = rng.uniform(-60, 60, size=(n, 9))
    t[: n // 4] *= 0.05                                   # small triangles (icosphere-like, ~1 lu)
    # axis-aligned box faces with integer coordinates (exact edge/vertex hits are constructible)
    box = [[-50, -50, 50, -50, 50, -50, -50, -50, -50], [-50, 50, 50, 50, 50, -50, -50, 50, -50],
           [50, 50, 50, 50, -50, -50, 50, 50, -50], [50, -50, 50, -50, -50, -50, 50, -50, -50],
           [0, 0, 0, 8, 0, 0, 0, 8, 0], [0, 0, 0, 0, 8, 0, 8, 0, 0], [1, 1, 1, 5, 1, 1, 1, 1, 5],
           [0, 0, 0, 1, 1, 1, 2, 2, 2],                   # degenerate (collinear)
           [3, 3, 3, 3, 3, 3, 4, 4, 4]]                   # degenerate (repeated vertex)
    return np.concatenate([t, np.array(box, dtype=np.float64)])


def wall_rays(tris, per_tri=40, seed=12):
    """(tri index, point, move): generic rays aimed near the triangle plus adversarial ones that end or start
    on the plane and that hit edges / vertices exactly (REDO paths through jump_away_line)."""
    rng = np.random.default_rng(seed)
    out = []
    for ti, t in enumerate(tris):
        v0, v1, v2 = t[0:3], t[3:6], t[6:9]
        n = np.cross(v1 - v0, v2 - v0)
        ln = np.linalg.norm(n)
        if ln == 0:
            out.append((ti, v0 + 1.0, np.array([1.0, 0.5, 0.25])))
            continue
        n = n / ln
        for k in range(per_tri):
            a, b = rng.uniform(-0.3, 1.3, 2)
            target = v0 + a * (v1 - v0) + b * (v2 - v0)          # inside, outside and near edges
            d = rng.normal(size=3)
            length = abs(rng.normal()) * 3 + 0.01
            start = target - d * rng.uniform(0.0, 1.5) * length / np.linalg.norm(d)
            out.append((ti, start, d * length / np.linalg.norm(d)))
        # exact constructions (meaningful when coordinates are integers)
        for (p, q) in ((v0, v1), (v1, v2), (v2, v0)):
            mid = 0.5 * (p + q)
            out.append((ti, mid + 2 * n, -4 * n))                # through an edge midpoint
            out.append((ti, p + 2 * n, -4 * n))                  # through a vertex
        c = (v0 + v1 + v2) / 3
        out.append((ti, c + 2 * n, -2 * n))                      # ends exactly on the plane
        out.append((ti, c, (v1 - v0) * 0.1))                     # starts on the plane, moves inside it
        out.append((ti, c, n))                                   # starts on the plane, leaves it
        out.append((ti, c + 1e-13 * n, -n))                      # starts within EPS of the plane
        out.append((ti, c - 1e-13 * n, n))
    return out


def mol_pairs(n=4000, seed=13, R=0.5641895835477563):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-50, 50, size=(n, 3))
    mv = rng.normal(size=(n, 3)) * 1.41421356
    s = rng.uniform(-0.3, 1.3, size=(n, 1))
    off = rng.normal(size=(n, 3)) * R * 0.9
    target = p + s * mv + off
    target[::50] = p[::50]                                       # coincident
    target[1::50] = p[1::50] + mv[1::50]                         # exactly at the end of the move
    mv[2::97] = 0.0                                              # zero displacement
    return p, mv, target, R


def boxes(tris, per_tri=12, seed=14):
    rng = np.random.default_rng(seed)
    out = []
    for ti, t in enumerate(tris):
        lo, hi = t.reshape(3, 3).min(0), t.reshape(3, 3).max(0)
        for k in range(per_tri):
            c = rng.uniform(lo - 5, hi + 5)
            h = rng.uniform(0.1, 30, 3)
            out.append((ti, c - h, c + h))
        out.append((ti, lo, hi))                                 # the triangle's own bounding box
        out.append((ti, hi, hi + 1.0))                           # touching at a corner
        g = np.floor(lo / 50) * 50
        out.append((ti, g, g + 50.0))                            # a subpartition-shaped box
    return out


def rxn_cases(seed=15):
    rng = np.random.default_rng(seed)
    cases = []
    for k in range(400):
        n = int(rng.integers(1, 6))
        cum = np.cumsum(rng.uniform(0.001, 0.4, n))
        scaling = float(rng.choice([1.0, 1.0, 0.3, 2.5, float(rng.uniform(0.05, 3))]))
        cases.append((cum, scaling, 1000 + k, int(rng.integers(0, 20))))
    cases.append((np.array([1e140]), 1.0, 5, 0))                 # absorptive surface class: rate GIGANTIC
    cases.append((np.array([0.1]), 0.1, 6, 3))                   # max_fixed_p == scaling
    return cases


def box_triangles(h):
    """The 12 triangles of a centred cube with half edge h (create_box face order), 9 coordinates each."""
    v = np.array([[-h, -h, -h], [-h, -h, h], [-h, h, -h], [-h, h, h], [h, -h, -h], [h, -h, h], [h, h, -h], [h, h, h]], float)
    f = np.array([[1, 2, 0], [3, 6, 2], [7, 4, 6], [5, 0, 4], [6, 0, 2], [3, 5, 7], [1, 3, 2], [3, 7, 6], [7, 5, 4], [5, 1, 0],
                  [6, 4, 0], [3, 1, 5]])
    return np.ascontiguousarray(v[f].reshape(12, 9))


def disk_cases(n=3000, seed=16, R=0.5641895835477563):
    """(loc, mv, target, walls) for exact_disk: collisions next to the faces / edges / corners of a box and inside
    random triangle soups (multi-edge sweeps, crossings, parallel chords), targets inside the interaction disk."""
    rng = np.random.default_rng(seed)
    out = []
    box = box_triangles(10.0)
    for k in range(n):
        kind = k % 4
        if kind < 2:
            walls = box
            near = rng.uniform(0.01, 0.8, 3)
            far = np.array([1.0, float(rng.integers(0, 2)) + (0.0 if kind else 8 * rng.random()),
                            float(rng.integers(0, 2)) + (0.0 if kind else 8 * rng.random())])
            loc = np.array([10.0, 10.0, 10.0]) - near * far
        else:
            walls = np.ascontiguousarray(rng.uniform(-1.5, 1.5, (int(rng.integers(1, 7)), 9)))
            loc = rng.uniform(-0.5, 0.5, 3)
        mv = rng.normal(0, 1.4, 3)
        d = rng.normal(0, 1, 3)
        d -= d.dot(mv) / mv.dot(mv) * mv
        d *= rng.uniform(0, R) / np.linalg.norm(d)
        target = loc + d + mv * rng.uniform(-1e-3, 1e-3)
        if k % 97 == 0:
            target = loc.copy()                                  # hit the target exactly
        out.append((np.ascontiguousarray(loc), np.ascontiguousarray(mv), np.ascontiguousarray(target), walls))
    return out, R


def grid_points(tris, per_tri=12, seed=17):
    """(tri index, point on the triangle) incl. the three vertices, for xyz2grid."""
    rng = np.random.default_rng(seed)
    out = []
    for ti, t in enumerate(tris):
        v0, v1, v2 = t[0:3], t[3:6], t[6:9]
        if np.linalg.norm(np.cross(v1 - v0, v2 - v0)) == 0:
            continue
        for k in range(per_tri):
            a, b = rng.uniform(0.001, 0.998, 2)
            if a + b > 0.999:
                a, b = (1 - a) * 0.999, (1 - b) * 0.999
            out.append((ti, np.ascontiguousarray(v0 + a * (v1 - v0) + b * (v2 - v0))))
        out += [(ti, np.ascontiguousarray(v0.copy())), (ti, np.ascontiguousarray(v1.copy())), (ti, np.ascontiguousarray(v2.copy()))]
    return out
