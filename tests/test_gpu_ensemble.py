"""Ensemble-level parity (BASELINE.json north_star, second level): libmcx on the B200 with its per-molecule
Philox streams against the CPU oracle in SEQUENTIAL mode, i.e. the reference's own semantics (one global
ISAAC64 stream, molecules processed in order, reactions applied immediately, products diffused in the same
iteration: src4/diffuse_react_event.cpp:67-161).

 * molecule and reaction counts over time agree within 3 sigma across 32 seeds;
 * the two-sample Kolmogorov-Smirnov test on per-axis displacements passes at p > 0.01.
"""
import math

import numpy as np
import pytest
from scipy import stats

import common as cm
from mcell_b200 import abi

pytestmark = pytest.mark.gpu

N_SEEDS = 32
CHECKPOINTS = (5, 10, 20)


def _engine(t):
    from mcell_b200 import Engine
    return Engine(t)


def _oracle(t):
    from oracle import oracle_py as O
    return O.Oracle(t)


def _three_sigma(a, b, rel_floor=0.0):
    a, b = np.asarray(a, float), np.asarray(b, float)
    se = math.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    return abs(a.mean() - b.mean()) <= 3.0 * se + rel_floor * abs(b.mean()), (a.mean(), b.mean(), se)


def test_counts_over_time_agree_within_3_sigma_over_32_seeds():
    """A + B -> C (BASELINE config 2 chemistry, reduced box): species counts and reaction counts at
    iterations 5/10/20, GPU (snapshot semantics + conflict rounds, Philox) vs oracle sequential (ISAAC64)."""
    gpu = {k: [] for k in CHECKPOINTS}
    ref = {k: [] for k in CHECKPOINTS}
    gpu_rxn, ref_rxn = [], []
    for seed in range(1, N_SEEDS + 1):
        t, mols = cm.reactive_box(n=4000, edge_um=0.25, p_target=0.3, seed=seed, subpartition_dimension=0.125)
        e, o = _engine(t), _oracle(t)
        e.upload(mols)
        o.upload(mols)
        done = 0
        for k in CHECKPOINTS:
            e.step(k - done)
            o.step(k - done, 0)
            done = k
            gpu[k].append(float(e.counts()[0][2]))
            ref[k].append(float(o.counts()[0][2]))
        gpu_rxn.append(float(e.counts()[1][0]))
        ref_rxn.append(float(o.counts()[1][0]))
        # conservation on both sides: A + C and B + C are invariant
        for c in (e.counts()[0], o.counts()[0]):
            assert c[0] + c[2] == 2000 and c[1] + c[2] == 2000
        e.close()
    for k in CHECKPOINTS:
        ok, info = _three_sigma(gpu[k], ref[k])
        assert ok, ("C count at iteration %d" % k, info)
        assert np.mean(ref[k]) > 20
    ok, info = _three_sigma(gpu_rxn, ref_rxn)
    assert ok, ("reaction count", info)


def test_reversible_binding_counts_agree_within_3_sigma_over_32_seeds():
    """Ca + CB <-> CaCB (config 4 chemistry): bimolecular + unimolecular, products take partial steps."""
    gpu, ref = [], []
    for seed in range(1, N_SEEDS + 1):
        t, mols = cm.reversible_box(n=3000, edge_um=0.25, seed=seed)
        e, o = _engine(t), _oracle(t)
        e.upload(mols)
        o.upload(mols)
        e.step(15)
        o.step(15, 0)
        gpu.append([float(x) for x in e.counts()[0]])
        ref.append([float(x) for x in o.counts()[0]])
        e.close()
    gpu, ref = np.array(gpu), np.array(ref)
    for s in range(3):
        ok, info = _three_sigma(gpu[:, s], ref[:, s])
        assert ok, ("species %d" % s, info)


def test_displacement_distribution_ks():
    """Per-axis displacement of one free step far from walls: GPU (Philox-driven Ziggurat) vs oracle sequential
    (ISAAC64-driven Ziggurat of src/rng.c:173-218), two-sample KS at p > 0.01; and against the analytic normal
    with sigma = space_step / sqrt(2) (diffusion_utils.inl:116-118)."""
    n = 200000
    t, mols = cm.free_diffusion_box(n=n, edge_um=4.0, seed=7)
    # keep only molecules far from the walls so that no reflection enters the displacement
    far = (np.abs(mols.x) < 150) & (np.abs(mols.y) < 150) & (np.abs(mols.z) < 150)
    e, o = _engine(t), _oracle(t)
    e.upload(mols)
    o.upload(mols)
    before = mols.sorted_by_id()
    e.step(1)
    o.step(1, 0)
    g, r = e.download().sorted_by_id(), o.download().sorted_by_id()
    assert g.n == n and r.n == n
    sel = far[np.argsort(mols.id, kind="stable")]
    sigma = t.species[0].space_step / math.sqrt(2.0)
    for ax in ("x", "y", "z"):
        dg = (getattr(g, ax) - getattr(before, ax))[sel]
        dr = (getattr(r, ax) - getattr(before, ax))[sel]
        p2 = stats.ks_2samp(dg, dr).pvalue
        p1 = stats.kstest(dg / sigma, "norm").pvalue
        assert p2 > 0.01, (ax, "two-sample", p2)
        assert p1 > 0.01, (ax, "normal", p1)
    # radial: |d|^2 / sigma^2 ~ chi^2(3)
    d2 = sum(((getattr(g, ax) - getattr(before, ax))[sel] / sigma) ** 2 for ax in ("x", "y", "z"))
    assert stats.kstest(d2, "chi2", args=(3,)).pvalue > 0.01
