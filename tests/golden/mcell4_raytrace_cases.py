"""Seeded inputs of the ray_trace_vol pin (tests/test_oracle_vs_reference_mcell4_raytrace.py and the generator of its
goldens, tests/golden/gen_mcell4_raytrace_golden.py): a population of volume molecules of three species around and
inside a reflective icosphere in a box, in a partition of 10^3 subpartitions, and one move per probed molecule."""
import numpy as np

from mcell_b200 import abi
from mcell_b200.model import Config, Model, MolArrays, create_box, create_icosphere, release_uniform_box

BOX_UM = 0.4
CAP = 64


def scene(seed=31, n=12000, subdivisions=2, interaction_radius=0.01, subpartition_dimension=0.042):
    """A + B -> A and A + A -> B react, N reacts with nothing; interaction radius 0.01 um against 0.04 um subpartitions,
    so that the neighbouring subpartitions of a move matter.  The partition is barely larger than the box: long moves
    of molecules next to the box end outside it (get_displacement_up_to_partition_boundary)."""
    m = Model(Config(seed=seed, partition_dimension=0.42, subpartition_dimension=subpartition_dimension,
                     interaction_radius=interaction_radius))
    m.add_species("A", 1e-6)
    m.add_species("B", 1e-6)
    m.add_species("N", 1e-6)
    m.add_reaction_rule(["A", "B"], ["A"], 1e8)
    m.add_reaction_rule(["A", "A"], ["B"], 1e8)
    sv, sf = create_icosphere(0.12, subdivisions)
    m.add_geometry_object(sv, sf)
    bv, bf = create_box(BOX_UM)
    m.add_geometry_object(bv, bf)
    t = m.build(max_molecules=2 * n + 64, rng_mode=abi.MCX_RNG_TAPE)
    rng = np.random.default_rng(seed)
    pos = release_uniform_box(rng, n, BOX_UM, t.length_unit, margin=1e-3)
    species = rng.choice(3, size=n, p=[0.45, 0.45, 0.1]).astype(np.uint32)
    return t, MolArrays.from_positions(pos, species)


def moves(t, mols, n_cases=2500, seed=97):
    """(molecule id, displacement, last_hit_wall, ISAAC seed, words to skip) per case: Gaussian moves of one to three
    subpartitions, every 7th axis-aligned (zero components), every 13th three times as long, every 17th aimed through
    the mesh vertex nearest to the molecule (edge / vertex hits: the REDO restarts that draw words and change the
    displacement), every 11th with a last_hit_wall (the wall the trace must skip: a wall of the molecule's own
    subpartition when there is one)."""
    rng = np.random.default_rng(seed)
    sp_len = t.cfg.partition_edge_length / t.cfg.num_subparts_per_edge
    out = []
    for k in range(n_cases):
        mid = int(rng.integers(0, mols.n))
        d = rng.normal(0.0, 0.9 * sp_len, 3)
        if k % 7 == 0:
            d[int(rng.integers(0, 3))] = 0.0
        if k % 13 == 0:
            d *= 3.0
        if k % 17 == 0:
            p = np.array([mols.x[mid], mols.y[mid], mols.z[mid]])
            v = np.asarray(t.vertices, np.float64).reshape(-1, 3)
            d = (v[np.argmin(((v - p) ** 2).sum(axis=1))] - p) * 1.3
        out.append((mid, d, k % 11 == 0, int(rng.integers(1, 1 << 30)), int(rng.integers(0, 50))))
    return out
