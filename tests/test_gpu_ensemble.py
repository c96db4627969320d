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


def test_kept_reactant_and_surface_class_pathways_agree_with_reference_semantics():
    """The round-2 surface pathways at ensemble level: transporter (RX_FLIP) and enzyme (kept volume and surface reactant) on
    a sphere, and the finite-rate reactions with a surface class (permeation both ways, consuming and catalytic wall
    reactions) — per-rule reaction counts after 12 iterations, GPU against the oracle in the reference's sequential
    semantics, 16 seeds each.  The pathway that consumes its reactant must agree within 3 sigma.  The pathways that KEEP
    the volume reactant carry the documented deviation of the snapshot semantics (the kept molecule
    takes the rest of its step one iteration later with a freshly drawn displacement, where the reference carries on with
    what is left of the old one and therefore stays closer to the wall): their rates come out 3-10 % lower, measured on
    the CPU with the oracle's two modes (profiles/r02_zz_kept_reactant_bias.txt); the bound here is 3 sigma + 12 %."""
    n_seeds = 16
    kept_rules = {"transporter": (0, 1), "surface class": (0, 1, 3)}
    for name, make, n_rules in (("transporter", lambda s: cm.transporter_sphere(n_vol=5000, n_trans=1200, n_enz=900, seed=s), 2),
                                ("surface class", lambda s: cm.permeable_sphere(n=6000, seed=s), 4)):
        gpu, ref = [], []
        for seed in range(1, n_seeds + 1):
            t, mols = make(seed)
            e, o = _engine(t), _oracle(t)
            e.upload(mols)
            o.upload(mols)
            e.step(12)
            o.step(12, 0)
            gpu.append([float(x) for x in e.counts()[1][:n_rules]])
            ref.append([float(x) for x in o.counts()[1][:n_rules]])
            e.close()
        gpu, ref = np.array(gpu), np.array(ref)
        for r in range(n_rules):
            assert ref[:, r].mean() > 8, (name, r, ref[:, r].mean())
            ok, info = _three_sigma(gpu[:, r], ref[:, r], rel_floor=0.12 if r in kept_rules[name] else 0.0)
            assert ok, (name, "rule %d" % r, info)


def test_surface_surface_reactions_agree_with_reference_semantics():
    """SURVEY 8 a23 at ensemble level: A' + B' -> C', C' -> A', A' + E' -> D' + E' on a sphere (react_2D_all_neighbors),
    per-rule reaction counts after 12 iterations over 16 seeds, GPU against the oracle in the reference's sequential
    semantics.  The surface-surface rules must agree within 3 sigma; the unimolecular decay of the product C carries the
    documented lag of the snapshot semantics (a product draws its lifetime when it is first
    evaluated, one iteration after its birth): bound 3 sigma + 10 %."""
    n_seeds = 16
    gpu, ref = [], []
    for seed in range(1, n_seeds + 1):
        t, mols = cm.surface_reactions(seed=seed, p=0.08)
        e, o = _engine(t), _oracle(t)
        e.upload(mols)
        o.upload(mols)
        e.step(12)
        o.step(12, 0)
        gpu.append([float(x) for x in e.counts()[1][:3]])
        ref.append([float(x) for x in o.counts()[1][:3]])
        e.close()
    gpu, ref = np.array(gpu), np.array(ref)
    for r, floor in ((0, 0.0), (1, 0.10), (2, 0.0)):
        assert ref[:, r].mean() > 50, (r, ref[:, r].mean())
        ok, info = _three_sigma(gpu[:, r], ref[:, r], rel_floor=floor)
        assert ok, ("rule %d" % r, info)
