#!/usr/bin/env python3
"""Device-timed iteration cost of BASELINE configs 1-4 (the parity-test configurations; bench.py carries the headline
config 5).  One JSON line per config: python tools/bench_configs.py [--iters 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(name, t, mols, iters, warmup=3):
    from mcell_b200 import Engine
    e = Engine(t)
    e.upload(mols)
    e.set_profiling(True)
    e.step(warmup)
    st = e.step(iters)
    n = max(1, int(st.profiled_iterations))
    line = {"config": name, "molecules": int(mols.n), "walls": int(len(t.tri)), "iterations": iters,
            "ms_per_iteration": st.device_ms / iters, "molecule_steps_per_sec": st.molecule_steps / (st.device_ms * 1e-3),
            "ms_fast_pass0": st.ms_diffuse / n, "ms_pass1_and_generic": st.ms_diffuse_slow / n, "ms_resolve": st.ms_resolve / n,
            "ms_sort": st.ms_sort / n, "deferred_fraction": st.deferred_molecules / max(1, st.molecule_steps),
            "bimol_rxns": int(st.bimol_rxns), "unimol_rxns": int(st.unimol_rxns), "wall_reflections": int(st.mol_wall_reflections),
            "unresolved_conflicts": int(st.unresolved_conflicts)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", type=int, default=0, help="run one config (1-4)")
    a = ap.parse_args()
    import common as cm
    import test_gpu_fullsize as fs
    if a.only in (0, 1):
        t, mols = cm.free_diffusion_box(n=100000, seed=1, cap_factor=1.25)
        run("1: free diffusion, 1e5 molecules, 1 um reflective cube", t, mols, a.iters)
    if a.only in (0, 2):
        t, mols = cm.reactive_box(n=1_000_000, edge_um=2.0, seed=2, p_target=0.1, cap_factor=1.25)
        run("2: A+B->C, 1e6 molecules, 2 um box", t, mols, a.iters)
    if a.only in (0, 3):
        t, mols, _ = fs._config3(400_000, 8_000, seed=3)
        run("3: ligand-receptor icosphere, 20 480 triangles", t, mols, a.iters)
    if a.only in (0, 4):
        t, mols, _, _ = fs._config4(10_000_000, seed=4)
        run("4: synapse-like, 163 840 triangles, 1e7 molecules", t, mols, a.iters)


if __name__ == "__main__":
    main()
