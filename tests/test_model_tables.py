"""CPU tier: host-side table builder (units, steps, reaction probabilities, partition, geometry generators)
against the formulas and structural identities of SURVEY.md §8c/§8d/§A.1."""
import math

import numpy as np
import pytest

import common as cm
from mcell_b200 import abi
from mcell_b200.model import Model, Config, create_box, create_icosphere, N_AV, MY_PI


def test_default_units_and_partition():
    m = Model(Config())
    m.add_species("A", 1e-6)
    v, f = create_box(1.0)
    m.add_geometry_object(v, f)
    t = m.build(max_molecules=10)
    assert t.length_unit == pytest.approx(0.01)                      # 1/sqrt(10000) um
    assert t.cfg.rxn_radius_3d == pytest.approx(0.56418958, rel=1e-7)  # 1/sqrt(pi*density) um in lu
    assert t.cfg.partition_edge_length == pytest.approx(1000.0)
    assert t.cfg.num_subparts_per_edge == 20
    assert [t.cfg.origin[k] for k in range(3)] == [-500.0, -500.0, -500.0]
    assert t.species[0].space_step == pytest.approx(2.0)             # sqrt(4e8*1e-6*1e-6)/0.01
    assert t.cfg.use_expanded_list == 0                              # no bimolecular vol rxns (converter :84-87)
    assert t.vertices.max() == pytest.approx(50.0) and t.vertices.min() == pytest.approx(-50.0)


def test_partition_grows_to_cover_geometry():
    m = Model(Config(partition_dimension=1.0))
    m.add_species("A", 1e-6)
    v, f = create_box(3.0)
    m.add_geometry_object(v, f)
    t = m.build(max_molecules=10)
    o = np.array([t.cfg.origin[k] for k in range(3)])
    e = t.cfg.partition_edge_length
    assert (o <= t.vertices.min(0)).all() and (o + e >= t.vertices.max(0)).all()
    sp = e / t.cfg.num_subparts_per_edge
    assert sp == pytest.approx(50.0)
    assert np.allclose(o / sp, np.round(o / sp))                     # origin aligned to the subpartition length


def test_bimolecular_probability_factor_config2():
    """pb_factor ~ 3.68e-10, k ~ 2.7e8 for max_fixed_p = 0.1."""
    t, _ = cm.reactive_box(n=10, p_target=0.1)
    m = Model(Config())
    m.add_species("A", 1e-6); m.add_species("B", 1e-6)
    pb = cm._pb_factor(m, 0, 1)
    R = 1.0 / math.sqrt(MY_PI * 10000.0)
    eff_vel = (2.0 + 2.0) * 0.01 / 1e-6
    assert pb == pytest.approx(1e15 / N_AV / (2 * math.sqrt(MY_PI) * R * R * eff_vel), rel=1e-14)
    assert pb == pytest.approx(3.68e-10, rel=5e-3)
    assert 0.1 / pb == pytest.approx(2.7e8, rel=2e-2)
    assert t.classes[0].max_fixed_p == pytest.approx(0.1, rel=1e-12)
    assert t.classes[0].kind == abi.MCX_RXN_BIMOL_VOLVOL
    assert t.pathways[0].n_products == 1 and t.pathways[0].keep_reactant_mask == 0
    assert t.cfg.use_expanded_list == 1


def test_unimolecular_probability_and_cumulative_pathways():
    m = Model(Config(time_step=2e-6))
    for n in "ABC":
        m.add_species(n, 1e-6)
    m.add_reaction_rule(["C"], ["A", "B"], 1e4)
    m.add_reaction_rule(["C"], ["A"], 3e4)
    t = m.build(max_molecules=10)
    assert t.n_classes == 1 and t.n_pathways == 2
    assert t.classes[0].kind == abi.MCX_RXN_UNIMOL
    assert t.pathways[0].cum_prob == pytest.approx(1e4 * 2e-6)
    assert t.pathways[1].cum_prob == pytest.approx(4e4 * 2e-6)
    assert t.classes[0].max_fixed_p == pytest.approx(t.pathways[1].cum_prob)


def test_kept_reactants_are_not_products():
    m = Model(Config())
    for n in "ABC":
        m.add_species(n, 1e-6)
    m.add_reaction_rule(["A", "B"], ["A", "C"], 1e8)   # A is a catalyst
    t = m.build(max_molecules=10)
    pw = t.pathways[0]
    assert pw.keep_reactant_mask == 1 and pw.n_products == 1 and pw.products[0] == 2


def test_reaction_radius_guard():
    m = Model(Config(partition_dimension=1.0, subpartition_dimension=0.012))
    m.add_species("A", 1e-6); m.add_species("B", 1e-6)
    m.add_reaction_rule(["A", "B"], [], 1e8)
    with pytest.raises(ValueError):
        m.build(max_molecules=10)


@pytest.mark.parametrize("sub", [1, 2, 3, 4, 6])
def test_icosphere_counts_and_watertight(sub):
    v, f = create_icosphere(0.5, sub)
    assert len(f) == 20 * 4 ** (sub - 1)
    assert len(v) == 10 * 4 ** (sub - 1) + 2
    assert np.allclose(np.linalg.norm(v, axis=1), 0.5, rtol=1e-12)
    if sub <= 4:
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        fwd = {(a, b) for a, b in e.tolist()}
        assert all((b, a) in fwd for a, b in fwd)               # every edge shared by exactly two faces
        assert len(fwd) == 3 * len(f)
        # outward orientation
        c = v[f].mean(1)
        n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
        assert (np.einsum("ij,ij->i", c, n) > 0).all()


def test_box_is_closed_and_outward():
    v, f = create_box(2.0)
    assert v.shape == (8, 3) and f.shape == (12, 3)
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert (np.einsum("ij,ij->i", v[f].mean(1), n) > 0).all()
    assert 0.5 * np.linalg.norm(n, axis=1).sum() == pytest.approx(6 * 4.0)


def test_rules_of_one_pair_with_different_orientations_are_refused():
    """The reference keeps one reaction class per reactant geometry (A' + R' and A, + R' are different classes); the
    device tables hold one class per species pair, so the table builder refuses such a model instead of merging the
    rules under the first rule's orientation."""
    import pytest
    from mcell_b200.model import Model, Config, create_icosphere
    m = Model(Config(seed=1))
    m.add_species("A", 1e-6)
    m.add_species("R", 0.0, surface=True)
    m.add_species("AR", 0.0, surface=True)
    m.add_reaction_rule(["A'", "R'"], ["AR'"], 1e7)
    m.add_reaction_rule(["A,", "R'"], ["AR'"], 2e7)
    v, f = create_icosphere(0.2, 1)
    m.add_geometry_object(v, f)
    with pytest.raises(ValueError, match="orientation"):
        m.build(max_molecules=100)
