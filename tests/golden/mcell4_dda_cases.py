"""Deterministic cases for the subpartition walk (collect_crossed_subparts / collect_neighboring_subparts,
src4/collision_utils_subparts.inl:38-300): shared by the golden generator and the tests."""
import numpy as np

# (origin, partition_edge_length, subparts per edge, use_expanded_list, rxn_radius) — lengths in internal units
GRIDS = [
    ((-500.0, -500.0, -500.0), 1000.0, 20, 1, 0.5641895835477563),      # default 10 um partition, 0.5 um subpartitions
    ((-500.0, -500.0, -500.0), 1000.0, 20, 0, 0.5641895835477563),      # no volume-volume reactions: plain list
    ((-100.0, -100.0, -100.0), 200.0, 40, 1, 0.5641895835477563),       # 0.05 um subpartitions: several per step
    ((-64.0, -32.0, 0.0), 96.0, 7, 1, 2.5),                             # odd count, large radius, off-centre origin
]


def moves(seed=2024, n_per_grid=600):
    """List of (grid index, pos[3], disp[3], collect_for_molecules, collect_for_walls)."""
    rng = np.random.default_rng(seed)
    out = []
    for gi, (org, L, n, _exp, R) in enumerate(GRIDS):
        org = np.asarray(org)
        sp = L / n
        for k in range(n_per_grid):
            kind = k % 6
            pos = org + rng.uniform(0.02, 0.98, 3) * L
            if kind == 0:      # ordinary diffusion step
                disp = rng.normal(0, 1.4142, 3)
            elif kind == 1:    # long move across several subpartitions
                disp = rng.normal(0, 1.0, 3) * sp * rng.uniform(0.5, 3.0)
            elif kind == 2:    # start within the radius of faces / edges / corners of its subpartition
                cell = np.floor((pos - org) / sp)
                frac = np.where(rng.random(3) < 0.7, rng.choice([1e-9, 0.3 * R, 0.9 * R, sp - 0.9 * R, sp - 1e-9], 3), rng.uniform(0, sp, 3))
                pos = org + cell * sp + frac
                disp = rng.normal(0, 1.4142, 3)
            elif kind == 3:    # axis-aligned and zero components (guard_zero_div)
                disp = np.zeros(3)
                ax = rng.integers(0, 3)
                disp[ax] = rng.choice([-1, 1]) * sp * rng.uniform(0.1, 2.2)
                if k % 12 == 3:
                    disp[(ax + 1) % 3] = rng.normal(0, 1.0)
            elif kind == 4:    # ends exactly on a subpartition boundary / runs along one
                cell = np.floor((pos - org) / sp)
                target = org + (cell + rng.integers(-1, 3, 3)) * sp
                disp = np.where(rng.random(3) < 0.5, target - pos, rng.normal(0, 1.4142, 3))
            else:              # towards and past the partition boundary (the walk breaks out of range)
                pos = org + np.where(rng.random(3) < 0.5, rng.uniform(0.001, 0.02, 3), rng.uniform(0.98, 0.999, 3)) * L
                disp = rng.normal(0, 1.0, 3) * sp * 0.8
            dest = pos + disp
            # the reference asserts that the destination lies in the partition (ray_trace_vol clips the move first)
            if ((dest < org) | (dest >= org + L)).any() or ((pos < org) | (pos >= org + L)).any():
                continue
            out.append((gi, np.ascontiguousarray(pos), np.ascontiguousarray(disp), int(k % 5 != 4), int(k % 7 != 6)))
    return out
