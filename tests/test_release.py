"""Release of volume molecules (ReleaseEvent::release_ellipsoid_or_rectcuboid, src4/release_event.cpp:953-1003).
CPU tier: the oracle's restatement against an independent numpy restatement of the same lines on the published Philox
stream, and against the shapes' geometry.  GPU tier: mcx_release_volume_molecules against the oracle, bit for bit, then
both stepped on."""
import numpy as np
import pytest

import common as cm
from mcell_b200 import abi, engine


def _oracle(t):
    from oracle import oracle_py as O
    return O.Oracle(t)


def _numpy_release(seed, iteration, first_id, number, shape, location, diameter):
    """The reference's loop (:962-991) per molecule, words from Philox4x32-10 (seed, id, iteration | 2^63)."""
    out = np.zeros((number, 3))
    for k in range(number):
        mid = first_id + k
        words, blk = [], 0

        def dbl():
            nonlocal blk
            if not words:
                words.extend(int(w) for w in engine.philox_block(seed, mid, iteration | (1 << 63), blk))
                blk += 1
            return 2.3283064365386962890625e-10 * float(words.pop(0))
        while True:
            p = np.array([dbl() - 0.5, dbl() - 0.5, dbl() - 0.5])
            if shape == abi.MCX_RELEASE_CUBIC or p[0] * p[0] + p[1] * p[1] + p[2] * p[2] < 0.25:
                break
        if shape == abi.MCX_RELEASE_SPHERICAL_SHELL:
            r = np.sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) * 2
            p = np.array([0.0, 0.0, 0.5]) if r == 0 else p / r
        out[k] = p * np.asarray(diameter, float) + np.asarray(location, float)
    return out


@pytest.mark.parametrize("shape", [abi.MCX_RELEASE_CUBIC, abi.MCX_RELEASE_SPHERICAL, abi.MCX_RELEASE_SPHERICAL_SHELL])
def test_oracle_release_matches_numpy_restatement(shape):
    t, mols = cm.free_diffusion_box(n=50, seed=11)
    o = _oracle(t)
    o.upload(mols)
    loc, dia = (3.0, -7.5, 11.25), (40.0, 30.0, 20.0)
    first = o.release(0, 300, loc, dia, shape=shape)
    assert first == 50
    got = o.download().sorted_by_id()
    assert got.n == 350 and (got.id[50:] == np.arange(50, 350)).all()
    want = _numpy_release(11, 0, first, 300, shape, loc, dia)
    assert (got.x[50:] == want[:, 0]).all() and (got.y[50:] == want[:, 1]).all() and (got.z[50:] == want[:, 2]).all()
    assert (got.flags[50:] & abi.MCX_MOL_SCHEDULE_UNIMOL).all() and not (got.flags[50:] & abi.MCX_MOL_PARTIAL).any()
    assert (o.counts()[0][0] == 350)


def test_oracle_release_shapes_and_moments():
    t, mols = cm.free_diffusion_box(n=10, seed=5)
    o = _oracle(t)
    o.upload(mols)
    n = 40000
    loc, d = np.array([5.0, -4.0, 2.0]), np.array([60.0, 50.0, 40.0])
    o.release(0, n, loc, d, shape=abi.MCX_RELEASE_CUBIC)
    o.release(0, n, loc, d, shape=abi.MCX_RELEASE_SPHERICAL)
    o.release(0, n, loc, d, shape=abi.MCX_RELEASE_SPHERICAL_SHELL, release_time=0.25)
    m = o.download().sorted_by_id()
    q = (np.stack([m.x, m.y, m.z], 1)[10:] - loc) / d
    cube, ball, shell = q[:n], q[n:2 * n], q[2 * n:]
    assert (np.abs(cube) <= 0.5).all()
    assert np.abs(cube.mean(0)).max() < 5 * np.sqrt(1 / 12 / n)
    assert np.abs(cube.var(0) - 1 / 12).max() < 0.003
    r = np.linalg.norm(ball, axis=1)
    assert (r < 0.5).all()
    assert abs((r < 0.25).mean() - 0.125) < 5 * np.sqrt(0.125 * 0.875 / n)      # uniform in the volume
    assert np.abs(np.linalg.norm(shell, axis=1) - 0.5).max() < 1e-12
    assert np.abs(shell.mean(0)).max() < 5 * np.sqrt(1 / 12 / n)                 # uniform on the sphere: var = r^2 / 3
    # a release inside the iteration starts its first step there (diffusion_time = release time)
    assert (m.flags[10 + 2 * n:] & abi.MCX_MOL_PARTIAL).all()
    assert (m.diffusion_time[10 + 2 * n:] == 0.25).all()
    # streams of the release domain differ from the diffusion streams of the same molecule and iteration
    assert (engine.philox_block(5, 77, 0, 0) != engine.philox_block(5, 77, 1 << 63, 0)).any()


def test_oracle_release_rejects_bad_requests():
    t, mols = cm.free_diffusion_box(n=10, seed=5)
    o = _oracle(t)
    o.upload(mols)
    with pytest.raises(RuntimeError):
        o.release(99, 10, (0, 0, 0), (1, 1, 1))
    with pytest.raises(RuntimeError):
        o.release(0, 10, (0, 0, 0), (1, 1, 1), shape=7)
    with pytest.raises(RuntimeError):
        o.release(0, 10, (0, 0, 0), (1, 1, 1), release_time=3.5)
    with pytest.raises(RuntimeError):
        o.release(0, 10, (1e9, 0, 0), (1, 1, 1))            # outside the partition


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [abi.MCX_RELEASE_CUBIC, abi.MCX_RELEASE_SPHERICAL, abi.MCX_RELEASE_SPHERICAL_SHELL])
def test_gpu_release_matches_oracle_and_steps_on(shape):
    """Device release = oracle release bit for bit (ids, positions, flags, counts); then both run on: the released
    molecules diffuse, react (A + B -> C) and are counted like uploaded ones."""
    import test_gpu_parity as gp
    from mcell_b200 import Engine
    t, mols = cm.reactive_box(n=8000, edge_um=0.4, p_target=0.4, seed=9)
    o = _oracle(t)
    o.upload(mols)
    e = Engine(t)
    e.upload(mols)
    for it in range(2):
        st_o, st_g = o.step(1, 1), e.step(1)
        assert st_g.bimol_rxns == st_o.bimol_rxns
    half = 0.4 / t.length_unit / 2
    for sp, frac, tr in ((0, 0.9, 0.0), (1, 0.5, 2.5)):
        dia = (2 * half * frac,) * 3
        a = o.release(sp, 3000, (0.0, 0.0, 0.0), dia, shape=shape, release_time=tr)
        b = e.release(sp, 3000, (0.0, 0.0, 0.0), dia, shape=shape, release_time=tr)
        assert a == b
    assert (e.counts()[0] == o.counts()[0]).all()
    gp._assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())
    for it in range(4):
        st_o, st_g = o.step(1, 1), e.step(1)
        assert st_g.bimol_rxns == st_o.bimol_rxns, it
        assert st_g.molecule_steps == st_o.molecule_steps, it
    assert (e.counts()[0] == o.counts()[0]).all()
    gp._assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())


@pytest.mark.gpu
def test_gpu_release_into_empty_population_and_errors():
    from mcell_b200 import Engine
    from mcell_b200.model import MolArrays
    t, mols = cm.free_diffusion_box(n=10, seed=3, cap_factor=20000)
    e = Engine(t)
    with pytest.raises(engine.McxError):
        e.release(0, 10, (0, 0, 0), (10, 10, 10))          # nothing uploaded yet
    e.upload(MolArrays(0))
    first = e.release(0, 100000, (0, 0, 0), (80, 80, 80), shape=abi.MCX_RELEASE_SPHERICAL)
    assert first == 0 and e.num_molecules() == 100000
    m = e.download().sorted_by_id()
    assert (m.id == np.arange(100000)).all()
    assert (np.sqrt(m.x ** 2 + m.y ** 2 + m.z ** 2) < 40).all()
    st = e.step(3)
    assert st.molecule_steps == 300000 and e.num_molecules() == 100000
    with pytest.raises(engine.McxError):
        e.release(0, 10, (0, 0, 0), (1, 1, 1), release_time=99.0)
    with pytest.raises(engine.McxError):
        e.release(0, 10, (1e9, 0, 0), (1, 1, 1))            # escapes the partition


def _region_case(seed=4):
    """counted_spheres: object 0 = outer icosphere (0.3 um), 1 = inner icosphere (0.15 um, shifted), 2 = the box."""
    t, mols = cm.counted_spheres(n=2000, seed=seed, max_molecules=80000)
    lu = t.length_unit
    outer_box = ((0.0, 0.0, 0.0), (2 * 0.3 / lu + 0.02,) * 3)     # centre, edges of the outer sphere's bounding box
    return t, mols, outer_box


def _inside(t, pos, obj):
    from mcell_b200.model import points_inside_mesh
    return points_inside_mesh(pos, t.vertices[t.tri][t.wall_object == obj])


def test_oracle_region_release_and_list_release():
    """SURVEY 8 f3 (release_event.cpp:904-951, 1008-1040): molecules released into 'outer - inner' lie inside the outer and
    outside the inner sphere, uniformly (compared with a rejection sample drawn in numpy), and carry the counted volume
    a fresh ray cast gives; a list release puts every molecule where the list says."""
    t, mols, (loc, dia) = _region_case()
    o = _oracle(t)
    o.upload(mols)
    n0 = mols.n
    first = o.release(2, 6000, loc, dia, shape=abi.MCX_RELEASE_REGION, region_in=1, region_out=2)
    assert first == n0
    m = o.download().sorted_by_id()
    new = m.id >= first
    pos = np.stack([m.x, m.y, m.z], 1)[new]
    assert new.sum() == 6000 and (m.species[new] == 2).all()
    assert _inside(t, pos, 0).all() and not _inside(t, pos, 1).any()
    assert (m.counted_volume[new] == cm.counted_volume_of(t, pos)).all()
    # uniform over the region: octant occupancies against a numpy rejection sample of the same region
    rng = np.random.default_rng(1)
    ref = rng.uniform(-0.5, 0.5, (60000, 3)) * np.asarray(dia) + np.asarray(loc)
    ref = ref[_inside(t, ref, 0) & ~_inside(t, ref, 1)]
    oct_new = np.bincount((pos > 0) @ np.array([1, 2, 4]), minlength=8) / len(pos)
    oct_ref = np.bincount((ref > 0) @ np.array([1, 2, 4]), minlength=8) / len(ref)
    assert np.abs(oct_new - oct_ref).max() < 5 * np.sqrt(0.125 / 6000 + 0.125 / len(ref))
    r_new, r_ref = np.linalg.norm(pos, axis=1), np.linalg.norm(ref, axis=1)
    assert abs(r_new.mean() - r_ref.mean()) < 5 * r_ref.std() * np.sqrt(1 / 6000 + 1 / len(ref))
    # the inner sphere alone: all three objects enclose it
    first2 = o.release(0, 1500, loc, dia, shape=abi.MCX_RELEASE_REGION, region_in=2)
    m = o.download().sorted_by_id()
    new2 = m.id >= first2
    pos2 = np.stack([m.x, m.y, m.z], 1)[new2]
    assert new2.sum() == 1500 and _inside(t, pos2, 1).all()
    assert (m.counted_volume[new2] == t.counted_volume_sets.index(frozenset({0, 1, 2}))).all()
    # list release
    lpos = rng.uniform(-20, 20, (500, 3))
    lsp = (np.arange(500) % 3).astype(np.uint32)
    lcv = cm.counted_volume_of(t, lpos)
    first3 = o.release_list(lsp, lpos, lcv)
    m = o.download().sorted_by_id()
    new3 = m.id >= first3
    assert (m.id[new3] == first3 + np.arange(500)).all() and (m.species[new3] == lsp).all()
    assert (np.stack([m.x, m.y, m.z], 1)[new3] == lpos).all() and (m.counted_volume[new3] == lcv).all()
    with pytest.raises(RuntimeError):
        o.release(0, 10, loc, dia, shape=abi.MCX_RELEASE_REGION, region_in=0)
    with pytest.raises(RuntimeError):
        o.release(0, 10, (200.0, 0, 0), (5, 5, 5), shape=abi.MCX_RELEASE_REGION, region_in=1)   # the box misses the object


@pytest.mark.gpu
def test_gpu_region_and_list_release_match_oracle_and_step_on():
    """mcx_release_volume_molecules(MCX_RELEASE_REGION) and mcx_release_list against the oracle, bit for bit (ids,
    positions, counted volumes), then both stepped on with per-volume counts compared."""
    import test_gpu_parity as gp
    from mcell_b200 import Engine
    t, mols, (loc, dia) = _region_case(seed=6)
    o = _oracle(t)
    o.upload(mols)
    e = Engine(t)
    e.upload(mols)
    o.step(1, 1)
    e.step(1)
    for args in ((2, 20000, 1, 2, 0.0), (0, 5000, 2, 0, 1.25), (1, 5000, 1, 0, 0.0)):
        sp, num, rin, rout, tr = args
        a = o.release(sp, num, loc, dia, shape=abi.MCX_RELEASE_REGION, release_time=tr, region_in=rin, region_out=rout)
        b = e.release(sp, num, loc, dia, shape=abi.MCX_RELEASE_REGION, release_time=tr, region_in=rin, region_out=rout)
        assert a == b
    rng = np.random.default_rng(2)
    lpos = rng.uniform(-35, 35, (3000, 3))
    lsp = (np.arange(3000) % 2).astype(np.uint32)
    lcv = cm.counted_volume_of(t, lpos)
    assert o.release_list(lsp, lpos, lcv) == e.release_list(lsp, lpos, lcv)
    assert (e.counts()[0] == o.counts()[0]).all()
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    gp._assert_same_population(a, b)
    pos = np.stack([b.x, b.y, b.z], 1)
    assert (b.counted_volume == cm.counted_volume_of(t, pos)).all()
    mo, _ = o.counts_by_volume()
    mg, _ = e.counts_by_volume()
    assert (mo == mg).all()
    for it in range(4):
        st_o, st_g = o.step(1, 1), e.step(1)
        assert st_g.bimol_rxns == st_o.bimol_rxns and st_g.mol_wall_transparent == st_o.mol_wall_transparent, it
    mo, ro = o.counts_by_volume()
    mg, rg = e.counts_by_volume()
    assert (mo == mg).all() and (ro == rg).all()
    with pytest.raises(engine.McxError):
        e.release(0, 10, (200.0, 0, 0), (5, 5, 5), shape=abi.MCX_RELEASE_REGION, region_in=1)


def _surface_case(seed=3, n_rec=1500):
    """ligand_receptor_sphere: 1280 sphere walls (object 0) with receptors / pumps already on some tiles"""
    t, mols = cm.ligand_receptor_sphere(n_lig=4000, n_rec=n_rec, n_pump=500, seed=seed, release_products=False, max_molecules=40000)
    tri = t.vertices[t.tri]
    cz = tri.mean(axis=1)[:, 2]
    north = np.flatnonzero((t.wall_object == 0) & (cz > 0)).astype(np.uint32)
    return t, mols, north


def test_oracle_surface_release_onto_region():
    """SURVEY 8 f3 (release_event.cpp:640-760): molecules land on distinct, previously vacant tiles of the listed walls
    only, spread by area (tile counts per wall against the multinomial expectation), at a random position inside their
    tile (xyz2grid of the position gives the tile back) or at its centre; a nearly full region is filled to the last
    tile by the fall-back; more molecules than vacant tiles are refused."""
    t, mols, north = _surface_case()
    o = _oracle(t)
    o.upload(mols)
    before = o.download()
    occupied = set(zip(before.wall[before.wall != abi.MCX_NONE].tolist(), before.tile[before.wall != abi.MCX_NONE].tolist()))
    first = o.release_surface(3, 3000, north, orientation=0)      # species 3 = LR
    m = o.download().sorted_by_id()
    new = m.id >= first
    assert new.sum() == 3000 and (m.species[new] == 3).all()
    assert np.isin(m.wall[new], north).all()
    placed = list(zip(m.wall[new].tolist(), m.tile[new].tolist()))
    assert len(set(placed)) == 3000 and not (set(placed) & occupied)
    assert set(np.unique(m.orientation[new]).tolist()) == {-1, 1} and abs(int(m.orientation[new].sum())) < 5 * np.sqrt(3000)
    # random position inside the tile: the position maps back to the tile
    L = engine.load_library()
    import ctypes as C
    L.mcx_xyz2grid.argtypes = [C.c_void_p, C.c_void_p]
    L.mcx_xyz2grid.restype = C.c_uint32
    tri = t.vertices[t.tri]
    pos = np.stack([m.x, m.y, m.z], 1)[new]
    for k in range(0, 3000, 7):
        v9 = np.ascontiguousarray(tri[m.wall[new][k]].reshape(9))
        q = np.ascontiguousarray(pos[k])
        assert L.mcx_xyz2grid(v9.ctypes.data, q.ctypes.data) == m.tile[new][k]
    # by area: walls of the region have nearly equal areas here, so counts per wall are ~ multinomial(3000, area / total)
    a = 0.5 * np.linalg.norm(np.cross(tri[north, 1] - tri[north, 0], tri[north, 2] - tri[north, 0]), axis=1)
    cnt = np.array([(m.wall[new] == wi).sum() for wi in north])
    exp = 3000 * a / a.sum()
    assert ((cnt - exp) ** 2 / exp).sum() < len(north) + 6 * np.sqrt(2 * len(north))     # chi-square
    # tile centres
    first2 = o.release_surface(2, 200, north, orientation=1, randomize_pos=False)
    m2 = o.download().sorted_by_id()
    new2 = m2.id >= first2
    L.mcx_grid2uv.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    uv = np.zeros(2)
    for k in range(0, 200, 5):
        v9 = np.ascontiguousarray(tri[m2.wall[new2][k]].reshape(9))
        L.mcx_grid2uv(v9.ctypes.data, int(m2.tile[new2][k]), uv.ctypes.data)
        assert m2.u[new2][k] == uv[0] and m2.v[new2][k] == uv[1]
    assert (m2.orientation[new2] == 1).all()
    # fill the region to the last vacant tile (fall-back), then one more is refused
    L.mcx_grid_num_tiles.argtypes = [C.c_void_p]
    L.mcx_grid_num_tiles.restype = C.c_uint32
    n_tiles = sum(int(L.mcx_grid_num_tiles(np.ascontiguousarray(tri[wi].reshape(9)).ctypes.data)) for wi in north)
    on_north = int(np.isin(m2.wall, north).sum())
    first3 = o.release_surface(3, n_tiles - on_north, north)
    m3 = o.download()
    s3 = np.isin(m3.wall, north)
    assert s3.sum() == n_tiles and len(set(zip(m3.wall[s3].tolist(), m3.tile[s3].tolist()))) == n_tiles
    with pytest.raises(RuntimeError):
        o.release_surface(3, 1, north)
    with pytest.raises(RuntimeError):
        o.release_surface(0, 10, north)                   # a volume species


@pytest.mark.gpu
def test_gpu_surface_release_matches_oracle_and_steps_on():
    """mcx_release_surface_molecules against the oracle, bit for bit (ids, walls, tiles, uv, orientations), incl. the
    fall-back fill of a nearly full region; then both stepped on (ligands bind the released receptors)."""
    import test_gpu_parity as gp
    from mcell_b200 import Engine
    t, mols, north = _surface_case(seed=5)
    o = _oracle(t)
    o.upload(mols)
    e = Engine(t)
    e.upload(mols)
    o.step(1, 1)
    e.step(1)
    south = np.setdiff1d(np.flatnonzero(t.wall_object == 0), north).astype(np.uint32)
    for sp, num, walls, orient, tr, rnd in ((2, 1500, north, 1, 0.0, True), (4, 800, north, 0, 1.5, True), (2, 500, south, -1, 0.0, False)):
        a = o.release_surface(sp, num, walls, orientation=orient, release_time=tr, randomize_pos=rnd)
        b = e.release_surface(sp, num, walls, orientation=orient, release_time=tr, randomize_pos=rnd)
        assert a == b
        assert (e.counts()[0] == o.counts()[0]).all()
    gp._assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())
    # nearly full: everything that is left on the southern walls but 40 tiles
    m = e.download()
    import ctypes as C
    L = engine.load_library()
    L.mcx_grid_num_tiles.argtypes = [C.c_void_p]
    L.mcx_grid_num_tiles.restype = C.c_uint32
    tri = t.vertices[t.tri]
    n_tiles = sum(int(L.mcx_grid_num_tiles(np.ascontiguousarray(tri[wi].reshape(9)).ctypes.data)) for wi in south)
    left = n_tiles - int(np.isin(m.wall, south).sum())
    assert o.release_surface(5, left - 40, south) == e.release_surface(5, left - 40, south)
    with pytest.raises(engine.McxError):
        e.release_surface(5, 41, south)
    a, b = o.download().sorted_by_id(), e.download().sorted_by_id()
    gp._assert_same_population(a, b)
    s = b.wall != abi.MCX_NONE
    assert len(np.unique(np.stack([b.wall[s], b.tile[s]], 1), axis=0)) == s.sum()      # one molecule per tile
    for it in range(4):
        st_o, st_g = o.step(1, 1), e.step(1)
        assert st_g.bimol_rxns == st_o.bimol_rxns and st_g.unimol_rxns == st_o.unimol_rxns, it
    assert (e.counts()[0] == o.counts()[0]).all()
    gp._assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())


def _expr_cases():
    U, I, D = abi.MCX_REGION_UNION, abi.MCX_REGION_INTERSECT, abi.MCX_REGION_DIFFERENCE
    # intersecting_counted_spheres: objects 0 and 1 = the two overlapping spheres, 2 = the box
    return [("union", (0, 1, U), lambda a, b: a | b), ("lens", (0, 1, I), lambda a, b: a & b),
            ("difference", (0, 1, D), lambda a, b: a & ~b), ("box minus both", (2, 0, D, 1, D), lambda a, b: ~a & ~b),
            ("symmetric difference", (0, 1, D, 1, 0, D, U), lambda a, b: a ^ b)]


def test_oracle_region_release_with_region_expressions():
    """Region expressions of a release (RegionExprNode trees: UNION / INTERSECT / DIFFERENCE of closed objects,
    release_event.cpp:787-813) as postfix programs over the memberships one ray cast gives: every released molecule lies
    where the expression says (independent point-in-mesh test in numpy), counted volumes included — the two spheres
    intersect, so the counted volume comes from the set of enclosing objects."""
    t, mols = cm.intersecting_counted_spheres(n=500, seed=3)
    lu = t.length_unit
    box = ((0.0, 0.0, 0.0), (0.62 / lu,) * 3)
    o = _oracle(t)
    o.upload(mols)
    for name, expr, where in _expr_cases():
        first = o.release(2, 1500, box[0], box[1], shape=abi.MCX_RELEASE_REGION, region_expr=expr)
        m = o.download().sorted_by_id()
        new = m.id >= first
        pos = np.stack([m.x, m.y, m.z], 1)[new]
        a, b = _inside(t, pos, 0), _inside(t, pos, 1)
        assert new.sum() == 1500 and where(a, b).all(), name
        assert (m.counted_volume[new] == cm.counted_volume_of(t, pos)).all(), name
    for bad in ((0, abi.MCX_REGION_UNION), (0, 1), (0, 1, 0x90), (40, 1, abi.MCX_REGION_UNION)):
        with pytest.raises(RuntimeError):
            o.release(2, 10, box[0], box[1], shape=abi.MCX_RELEASE_REGION, region_expr=bad)


@pytest.mark.gpu
def test_gpu_region_release_with_region_expressions_matches_oracle():
    import test_gpu_parity as gp
    from mcell_b200 import Engine
    t, mols = cm.intersecting_counted_spheres(n=500, seed=3)
    t.cfg.max_molecules = 40000
    lu = t.length_unit
    box = ((0.0, 0.0, 0.0), (0.62 / lu,) * 3)
    o = _oracle(t)
    o.upload(mols)
    e = Engine(t)
    e.upload(mols)
    for name, expr, _ in _expr_cases():
        assert o.release(2, 4000, box[0], box[1], shape=abi.MCX_RELEASE_REGION, region_expr=expr) == \
            e.release(2, 4000, box[0], box[1], shape=abi.MCX_RELEASE_REGION, region_expr=expr), name
    gp._assert_same_population(o.download().sorted_by_id(), e.download().sorted_by_id())
    with pytest.raises(engine.McxError):
        e.release(2, 10, box[0], box[1], shape=abi.MCX_RELEASE_REGION, region_expr=(0, abi.MCX_REGION_UNION))
