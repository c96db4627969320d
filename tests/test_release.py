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
