"""Cases of the neighbour-tile pins (tests/test_oracle_vs_reference_mcell4_tiles.py, gen_mcell4_tiles_golden.py): closed
icospheres with perturbed vertices (walls with 1 to ~50 tiles, so neighbouring grids differ in size), an open fan of
triangles around one vertex (free edges, a vertex shared by many walls) and a flat strip; with every wall holding a grid
and with a random subset of the grids missing.  Reaction cases for test_bimolecular with a local probability factor and
test_many_bimolecular."""
import numpy as np

from mcell_b200.model import create_icosphere


def meshes():
    rng = np.random.default_rng(5)
    out = []
    for sub, scale in ((1, 40), (1, 120), (1, 60), (2, 200), (2, 90), (1, 25), (3, 300), (1, 8)):
        v, t = create_icosphere(0.05, sub)
        V = np.ascontiguousarray(np.asarray(v, np.float64) * scale)
        V = V * (1 + 0.25 * rng.random(V.shape))
        out.append((V, np.ascontiguousarray(np.asarray(t), np.uint32)))
    # a fan of 7 triangles around vertex 0 (open: the rim edges are free)
    n = 7
    ang = np.linspace(0, 1.7 * np.pi, n + 1)
    V = np.zeros((n + 2, 3)); V[1:, 0] = 5.5 * np.cos(ang) * (1 + 0.2 * rng.random(n + 1)); V[1:, 1] = 5.5 * np.sin(ang); V[1:, 2] = 0.3 * rng.random(n + 1)
    T = np.array([[0, i + 1, i + 2] for i in range(n)], np.uint32)
    out.append((np.ascontiguousarray(V), T))
    # a flat strip of 6 triangles with vertex orders rotated (every side of a wall gets to be the shared one)
    V = np.array([[0, 0, 0], [4, 0, 0], [0.5, 3.7, 0], [4.4, 3.9, 0], [8.1, 0.2, 0], [8.6, 4.2, 0], [12.2, 0.1, 0], [12.0, 4.0, 0]], np.float64)
    T = np.array([[0, 1, 2], [2, 1, 3], [3, 1, 4], [4, 5, 3], [6, 5, 4], [5, 6, 7]], np.uint32)
    out.append((V, T))
    return out


def grid_masks(n_walls, k):
    rng = np.random.default_rng(100 + k)
    return [None, (rng.random(n_walls) < 0.7).astype(np.uint8), (rng.random(n_walls) < 0.3).astype(np.uint8)]


def lpf_cases():
    """(cumulative pathway probabilities, scaling, local probability factor, seed, skip)"""
    rng = np.random.default_rng(11)
    out = []
    for i in range(400):
        npw = int(rng.integers(1, 5))
        cum = np.cumsum(rng.random(npw) * 10 ** rng.uniform(-3, 0.3))
        scaling = float(10 ** rng.uniform(-2, 1))
        lpf = 3.0 / float(rng.integers(1, 16))
        out.append((cum, scaling, lpf, int(rng.integers(1, 1000)), int(rng.integers(0, 50))))
    return out


def many_cases():
    """(list of cumulative pathway probabilities per class, scaling per class, local probability factor, seed, skip)"""
    rng = np.random.default_rng(12)
    out = []
    for i in range(400):
        n = int(rng.integers(2, 7))
        cums = [np.cumsum(rng.random(int(rng.integers(1, 4))) * 10 ** rng.uniform(-3, 0.2)) for _ in range(n)]
        scaling = 10 ** rng.uniform(-1.5, 1, n)
        lpf = 3.0 / float(rng.integers(1, 16))
        out.append((cums, scaling, lpf, int(rng.integers(1, 1000)), int(rng.integers(0, 50))))
    return out


# ---- find_surf_product_positions (products on vacant neighbour tiles) ------------------------------------------------------
# rule shapes: (kind, reactants are surface, entries of the rule's product list): 'S' new surface product, 'V' new volume
# product, 'K0' / 'K1' kept reactant 0 / 1.  kind: 1 unimolecular, 3 volume-surface (reactant 0 volume, 1 surface),
# 5 surface-surface.  All need more tiles than they free (the general branch).
PLACE_SHAPES = [
    (1, (1,), ("K0", "S")), (1, (1,), ("S", "K0")), (1, (1,), ("S", "S")), (1, (1,), ("S", "S", "S")), (1, (1,), ("S", "V", "S")),
    (1, (1,), ("K0", "S", "S")), (1, (1,), ("V", "K0", "S")),
    (3, (0, 1), ("S", "S")), (3, (0, 1), ("K1", "S")), (3, (0, 1), ("S", "K1")), (3, (0, 1), ("K0", "S", "S")), (3, (0, 1), ("S", "V", "S")),
    (3, (0, 1), ("K0", "K1", "S")), (3, (0, 1), ("S", "S", "S")),
    (5, (1, 1), ("S", "S", "S")), (5, (1, 1), ("K0", "K1", "S")), (5, (1, 1), ("S", "S", "S", "V")), (5, (1, 1), ("K1", "S", "K0")),
    (5, (1, 1), ("S", "V", "S", "S")),
]


def place_cases(n_per_mesh=60):
    """(mesh index, occupied (wall, tile) pairs, shape index, reactant sites [(wall, tile)] in rule order (None: volume),
    surface reactant index, seed, skip)"""
    rng = np.random.default_rng(21)
    out = []
    ms = meshes()
    for k in (0, 2, 3, 5, 8, 9):
        V, T = ms[k]
        area = 0.5 * np.linalg.norm(np.cross(V[T[:, 1]] - V[T[:, 0]], V[T[:, 2]] - V[T[:, 0]]), axis=1)
        n_axis = np.maximum(1, np.ceil(np.sqrt(area))).astype(int)
        per = n_axis * n_axis
        tiles = np.array([(w, t) for w in range(len(T)) for t in range(per[w])], np.uint32)
        for _ in range(n_per_mesh):
            frac = rng.choice([0.15, 0.5, 0.8, 0.93, 0.985])
            occ_mask = rng.random(len(tiles)) < frac
            si = int(rng.integers(0, len(PLACE_SHAPES)))
            kind, surf_flags, entries = PLACE_SHAPES[si]
            sites = []
            for f in surf_flags:
                if f:
                    j = int(rng.integers(0, len(tiles)))
                    occ_mask[j] = True
                    sites.append((int(tiles[j][0]), int(tiles[j][1])))
                else:
                    sites.append(None)
            if kind == 5 and sites[0] == sites[1]:
                continue
            surf_reac = 0 if kind == 1 else (1 if kind == 3 else int(rng.integers(0, 2)))
            out.append((k, np.ascontiguousarray(tiles[occ_mask]), si, sites, surf_reac, int(rng.integers(1, 1000)), int(rng.integers(0, 40))))
    return out


# surface-surface rule shapes that fit on the tiles they free (the recycled branches): same notation
# (two surface products next to a volume product — ("S", "S", "V") — make the reference divide by zero: after the two freed
# tiles are handed out the volume entry draws from an empty list of vacant tiles, :2232-2251; such tables are refused)
RECYCLE_SHAPES = [("S",), ("S", "S"), ("K0", "S"), ("S", "K1"), ("K1", "S", "V"), ("V", "K0", "S"), ("K0", "K1"), ("V",), ("K0", "V"), ("K1", "V", "S")]


def recycle_cases(n=240):
    """(shape index, initiator is reactant 1, seed, skip) on mesh 2 with the reactants on two fixed tiles"""
    rng = np.random.default_rng(22)
    return [(int(rng.integers(0, len(RECYCLE_SHAPES))), int(rng.integers(0, 2)), int(rng.integers(1, 1000)), int(rng.integers(0, 40)))
            for _ in range(n)]


# ---- react_2D_all_neighbors as a whole -------------------------------------------------------------------------------------
def react2d_models():
    """Static surface molecules of four species on a sphere (nothing diffuses: an evaluation is the neighbour test alone),
    surface-surface classes with and without orientation classes, with one and with several pathways; sparse and dense
    populations (walls without a molecule have no grid).  -> list of (tables, molecules, seeds)"""
    from mcell_b200.model import Model, Config, create_box, release_on_walls
    out = []
    for n, sub, radius, seed in ((700, 2, 0.12, 41), (2600, 3, 0.25, 42), (5200, 3, 0.25, 43)):
        m = Model(Config(seed=seed))
        for name in ("A", "B", "C", "D"):
            m.add_species(name, 0.0, surface=True)
        pb = m.config.time_step * m.config.surface_grid_density / 6.0
        m.add_reaction_rule(["A'", "B'"], ["C'"], 0.9 / pb)
        m.add_reaction_rule(["A'", "B'"], ["D'"], 0.6 / pb)           # second pathway of the same class
        m.add_reaction_rule(["A", "C"], ["D"], 1.4 / pb)               # no orientation class
        m.add_reaction_rule(["B'", "D,"], ["A'"], 2.5 / pb)            # opposite orientations
        m.add_reaction_rule(["C'", "C'"], ["A'", "B'"], 0.5 / pb)      # same species
        m.add_reaction_rule(["D'", "D'"], ["C'"], 4.0 / pb)            # probability above 1 with several partners
        sv, sf = create_icosphere(radius, sub)
        m.add_geometry_object(sv, sf)
        bv, bf = create_box(4 * radius)
        m.add_geometry_object(bv, bf)
        t = m.build(max_molecules=2 * n + 16, rng_mode=1)              # abi.MCX_RNG_TAPE
        rng = np.random.default_rng(seed)
        mols = release_on_walls(rng, t, np.arange(len(sf), dtype=np.uint32), n, 0, orientation=1, first_id=0, schedule_unimol=False)
        mols.species[:] = rng.integers(0, 4, n).astype(mols.species.dtype)
        mols.orientation[:] = rng.choice([-1, 1], n).astype(mols.orientation.dtype)
        out.append((t, mols, rng.integers(1, 100000, n).astype(np.uint32)))
    return out
