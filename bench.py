#!/usr/bin/env python3
"""bench.py — molecule-steps/sec of the diffuse-and-react hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # libmcx on N B200s (torchrun for N>1)
  python bench.py --impl reference --gpus N ...            # CPU arm: the mcell4-equivalent oracle

Workload (config.workload): BASELINE.json configs[4] — the reactive box with 4 species and
6 reactions at config 2's number density (125 molecules/um^3), 1e8 molecules, which fits one B200.
A "step" is one iteration (one DiffuseReactEvent::step) over every molecule.  At N>1 the same box is
slab-decomposed along z (strong scaling).  Inputs (3.2 GB of records + cell tables) are far larger
than the 126 MB L2, so no explicit L2 flush is needed between timed iterations.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DENSITY_PER_LU3 = 1.0e6 / 200.0 ** 3     # config 2: 1e6 molecules in a (2 um)^3 = (200 lu)^3 box
ITERS_PER_CALL = 10                      # "counts every 10 iterations": barrier window of one plugin call
B_ALG_DIFFUSE = 92.0                     # 32 read + 32 write + 28 neighbour staging
B_ALG_STEP = 160.0                       # + 68 B per-step sort
WORKLOAD = "reactive box, 4 species / 6 reactions, 1.25e5 molecules/um^3 = config 2's density (BASELINE configs[4])"


def build_model(n_total, seed=1, rank=0, world=1, cap_factor=1.25, cell_edge=0.0):
    """4 species, 6 reactions (4 bimolecular incl. a same-species one, 2 unimolecular)."""
    from mcell_b200.model import Model, Config, create_box, N_AV, MY_PI
    edge_lu = (n_total / DENSITY_PER_LU3) ** (1.0 / 3.0)
    edge_um = edge_lu * 0.01
    part_dim = max(10.0, math.ceil(edge_um + 1.0))
    m = Model(Config(seed=seed, partition_dimension=part_dim))
    for name in "ABCD":
        m.add_species(name, 1e-6)
    lu, ts = m.length_unit, m.config.time_step
    eff = 2 * m.space_step(1e-6) * lu / ts
    R = m.rxn_radius_um
    pb = 1.0 / (2.0 * math.sqrt(MY_PI) * R * R * eff) * 1.0e15 / N_AV
    k_bi = 0.1 / pb                      # max_fixed_p = 0.1
    m.add_reaction_rule(["A", "B"], ["C"], k_bi)
    m.add_reaction_rule(["C"], ["A", "B"], 1.0e4)
    m.add_reaction_rule(["A", "C"], ["D"], k_bi)
    m.add_reaction_rule(["D"], ["A", "C"], 1.0e4)
    m.add_reaction_rule(["B", "D"], ["C", "C"], k_bi)
    m.add_reaction_rule(["C", "C"], ["B", "D"], k_bi)
    v, f = create_box(edge_um)
    m.add_geometry_object(v, f)
    # slab ranks hold their own z-layers plus a halo on each side (3x the one-step reach, mcx_api.cu configure_slab),
    # and the halo refresh of an iteration is appended behind the results that still include the old halo copies:
    # own * (1 + 4 * halo / slab) slots, times the product headroom
    halo_lu = 3.0 * (m.rxn_radius_um / m.length_unit + 6.993 * m.space_step(1e-6)) + 2.0 * 3.5
    slab_lu = edge_lu / world
    per_rank = int(n_total / world * cap_factor * (1.0 + 4.0 * halo_lu / slab_lu if world > 1 else 1.0)) + 1024
    t = m.build(max_molecules=per_rank, rank=rank, world_size=world, cell_edge=cell_edge)
    return t, edge_um


def make_molecules(n_total, edge_um, length_unit, seed, slab=None, pinned=True):
    """Uniform positions; species in fixed proportions A:B:C:D = 4:4:1:1.  With a slab layout (mcx_slab_info of
    this rank) each rank generates only the molecules of its own z-slab, in one global id space; slab faces are
    the device's own layer boundaries (mcell_b200.comm mirrors the arithmetic)."""
    from mcell_b200.model import MolArrays
    from mcell_b200 import comm
    h = (edge_um / 2) / length_unit * (1 - 1e-9)
    rank, world = (slab.rank, slab.world_size) if slab is not None else (0, 1)
    # z-interval and molecule share of every rank (identical on all ranks, no communication)
    edges = [-h]
    for r in range(world - 1):
        hi_layer = comm.layer_range(slab.n_layers, r, world)[1]
        edges.append(min(h, max(-h, slab.grid_origin_z + hi_layer / slab.layer_rcp)))
    edges.append(h)
    cum = [int(round(n_total * (e + h) / (2 * h))) for e in edges]
    n, first_id = cum[rank + 1] - cum[rank], cum[rank]
    z_lo, z_hi = edges[rank], edges[rank + 1]
    rng = np.random.default_rng(seed * 1000 + rank)
    m = MolArrays(0)
    alloc = _pinned_alloc if pinned else (lambda shape, dt: np.zeros(shape, dt))
    m.x, m.y, m.z = alloc(n, np.float64), alloc(n, np.float64), alloc(n, np.float64)
    m.id, m.species, m.flags = alloc(n, np.uint32), alloc(n, np.uint32), alloc(n, np.uint32)
    m.diffusion_time, m.unimol_rxn_time = alloc(n, np.float64), alloc(n, np.float64)
    chunk = 1 << 22
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        m.x[s:e] = rng.uniform(-h, h, e - s)
        m.y[s:e] = rng.uniform(-h, h, e - s)
        z = rng.uniform(z_lo, z_hi, e - s)
        if slab is not None:  # a draw within rounding of a face could fall into the neighbour's layer
            bad = comm.rank_of(z, slab) != rank
            z[bad] = 0.5 * (z_lo + z_hi)
        m.z[s:e] = z
        r = rng.integers(0, 10, e - s)
        m.species[s:e] = np.select([r < 4, r < 8, r < 9], [0, 1, 2], 3)
    m.id[:] = np.arange(first_id, first_id + n, dtype=np.uint32)
    m.flags[:] = 2                         # MCX_MOL_SCHEDULE_UNIMOL: lifetimes drawn on first diffusion
    m.diffusion_time[:] = 0
    m.unimol_rxn_time[:] = -256.0
    m.n = n
    return m


def populate_by_release(eng, n_total, edge_um, length_unit):
    """The benchmark population, released ON the device (mcx_release_volume_molecules = ReleaseEvent::
    release_ellipsoid_or_rectcuboid): species A:B:C:D = 4:4:1:1 uniformly in the box, ids 0..n-1.  Every molecule's
    position comes from its own Philox stream of the release domain, so the population does not depend on the number
    of ranks (each rank keeps the molecules of its slab).  Returns the per-species numbers."""
    from mcell_b200 import abi
    from mcell_b200.model import MolArrays
    eng.upload(MolArrays(0))
    d = edge_um / length_unit * (1 - 1e-9)
    shares = [0.4, 0.4, 0.1, 0.1]
    numbers = [int(round(n_total * f)) for f in shares]
    numbers[0] += n_total - sum(numbers)
    for sp, k in enumerate(numbers):
        if k:
            eng.release(sp, k, (0.0, 0.0, 0.0), (d, d, d), shape=abi.MCX_RELEASE_CUBIC)
    return numbers


_PINNED_KEEP = []


def _pinned_alloc(n, dtype):
    """numpy view of pinned host memory (torch is plumbing for allocation only)."""
    import torch
    tdt = {np.float64: torch.float64, np.uint32: torch.int32}[dtype]
    try:
        tt = torch.empty(max(n, 1), dtype=tdt, pin_memory=True)
    except Exception:
        tt = torch.empty(max(n, 1), dtype=tdt)
    _PINNED_KEEP.append(tt)
    a = tt.numpy()[:n]
    return a.view(np.uint32) if dtype is np.uint32 else a


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason sampling during the timed region (NVML)."""

    def __init__(self, index=0, period=0.2):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def _traffic(kernel, molecules):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of this same command
    (profiles/traffic.json, written by tools/ncu_summary.py traffic); None when there is no capture at this size."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        e = t.get(kernel)
        if e and int(e.get("molecules", 0)) == int(molecules):
            return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- CPU arm
# The CPU arm is the mcell4-equivalent oracle in sequential (reference) semantics.  Its cost per molecule-step is set
# by the number of molecules per default 0.5 um subpartition (15.6k at this density: every molecule is tested against
# all of them, collision_utils.inl:520-552), so the bounded sample is a box of a x b x c WHOLE subpartitions of the
# same chemistry and density: what one core spends per molecule-step there is what it would spend in the 1e8 box.
SUBPART_UM = 0.5
MOLS_PER_SUBPART = DENSITY_PER_LU3 * (SUBPART_UM * 100.0) ** 3      # 15 625


def build_cpu_sample(dims, seed):
    """Reactive-box chemistry in a box of dims = (a, b, c) whole default subpartitions (faces 1 nm inside the
    subpartition boundaries)."""
    from mcell_b200.model import Model, Config, create_box, N_AV, MY_PI
    m = Model(Config(seed=seed))
    for name in "ABCD":
        m.add_species(name, 1e-6)
    lu, ts = m.length_unit, m.config.time_step
    eff = 2 * m.space_step(1e-6) * lu / ts
    R = m.rxn_radius_um
    pb = 1.0 / (2.0 * math.sqrt(MY_PI) * R * R * eff) * 1.0e15 / N_AV
    k_bi = 0.1 / pb
    m.add_reaction_rule(["A", "B"], ["C"], k_bi)
    m.add_reaction_rule(["C"], ["A", "B"], 1.0e4)
    m.add_reaction_rule(["A", "C"], ["D"], k_bi)
    m.add_reaction_rule(["D"], ["A", "C"], 1.0e4)
    m.add_reaction_rule(["B", "D"], ["C", "C"], k_bi)
    m.add_reaction_rule(["C", "C"], ["B", "D"], k_bi)
    v, f = create_box(1.0)                                   # unit cube centred at 0
    inset = 0.001
    size = np.array([d * SUBPART_UM - 2 * inset for d in dims])
    lo = np.array([-(d // 2) * SUBPART_UM + inset for d in dims])   # subpartition boundaries are multiples of 0.5 um
    v = (v + 0.5) * size + lo
    m.add_geometry_object(v, f)
    n = int(round(MOLS_PER_SUBPART * dims[0] * dims[1] * dims[2]))
    t = m.build(max_molecules=2 * n + 1024)
    return t, n, lo, size


def make_cpu_molecules(t, n, lo, size, seed):
    from mcell_b200.model import MolArrays
    rng = np.random.default_rng(seed)
    pos = (lo + size * (1e-6 + (1 - 2e-6) * rng.uniform(size=(n, 3)))) / t.length_unit
    r = rng.integers(0, 10, n)
    species = np.select([r < 4, r < 8, r < 9], [0, 1, 2], 3).astype(np.uint32)
    return MolArrays.from_positions(pos, species, schedule_unimol=True)


def _cpu_worker(args):
    """One independent seed of the bounded CPU sample (the reference's only scaling mode:
    utils/mcell4_runner/mcell4_runner.py:203-212): `warmup` untimed then `steps` timed iterations."""
    dims, seed, steps, warmup = args[:4]
    native = len(args) > 4 and args[4]
    from oracle import oracle_py as O
    t, n, lo, size = build_cpu_sample(dims, seed)
    mols = make_cpu_molecules(t, n, lo, size, seed)
    o = O.Oracle(t, native=native)
    o.upload(mols)
    if warmup:
        o.step(warmup, 0)
    t0 = time.perf_counter()
    st = o.step(steps, 0)
    dt = time.perf_counter() - t0
    return st.molecule_steps, dt


def cpu_sample(dims, steps, warmup, cores, native=False):
    """-> (aggregate molecule-steps/s, molecule-steps, slowest worker's seconds).  native: the -O3 -march=native build
    of the oracle instead of the reference's release flags (-O3 -march=core2, CMakeLists.txt:58)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(dims, 100 + i, steps, warmup, native) for i in range(cores)])
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return total / busy, total, busy


def cpu_single(dims, steps):
    """The same sample on ONE core with the others idle (BASELINE.json: the CPU column is quoted both single-threaded
    and as independent seeds on all cores) -> (molecule-steps/s, seconds)."""
    n_steps, dt = _cpu_worker((dims, 100, steps, 0))
    return n_steps / dt, dt


def _sample_text(dims, cores, steps):
    n = int(round(MOLS_PER_SUBPART * dims[0] * dims[1] * dims[2]))
    return ("%d independent seeds x %d molecules (%dx%dx%d whole default 0.5 um subpartitions, 15 625 molecules each) x %d "
            "iteration(s), same chemistry and density as the 1e8 box, sequential reference semantics" %
            (cores, n, dims[0], dims[1], dims[2], steps))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as O
    O.build()
    cores = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # one iteration costs ~2.5 s of one core per subpartition of the sample: size the sample so that the whole
    # --steps/--warmup run stays within ~2.5 minutes
    per_subpart_s = 2.6
    budget = 150.0
    n_sub = budget / ((steps + warmup) * per_subpart_s)
    dims = (2, 2, 2) if n_sub >= 8 else (1, 2, 2) if n_sub >= 4 else (1, 1, 2) if n_sub >= 2 else (1, 1, 1)
    value, total, busy = cpu_sample(dims, steps, warmup, cores)
    single, single_s = cpu_single(dims, 1)
    sample = _sample_text(dims, cores, steps)
    line = {
        "impl": "reference", "metric": "molecule_steps_per_sec", "value": value, "unit": "molecule-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * busy / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "molecules": args.molecules,
                   "cpu_sample_molecules_per_core": int(round(MOLS_PER_SUBPART * dims[0] * dims[1] * dims[2]))},
        "cpu_baseline": {"value": value, "unit": "molecule-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "single_thread": {"value": single, "cores": 1,
                                           "sample": "one of those seeds alone on the box, 1 iteration, %.1f s" % single_s}},
        "e2e": {"value": value, "unit": "molecule-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "mcell4-equivalent CPU oracle (oracle/), not the upstream binary: the reference cannot be built here (DESIGN.md 8)",
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libmcx has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from mcell_b200 import Engine, build as mb
    if rank == 0:
        mb.build()
    if dist:
        dist.barrier()

    n_total = args.molecules
    host_mols = None
    workload = WORKLOAD
    b_alg_step = B_ALG_STEP
    if args.config == 5:
        t, edge_um = build_model(n_total, seed=1, rank=rank, world=world, cell_edge=args.cell_edge)
    else:
        if world > 1:
            raise SystemExit("bench.py: configs 1-4 are single-GPU parity-test configurations (config 5 is the scaling workload)")
        t, host_mols, workload, b_alg_step, edge_um = build_small_config(args.config)
        n_total = int(host_mols.n)
    t.cfg.device = local_rank
    eng = Engine(t)
    slab = None
    if world > 1:
        from mcell_b200 import comm as mcomm
        eng.comm_init(mcomm.broadcast_unique_id(dist, rank, device="cuda"))
        slab = eng.slab_info()
    halo_text = {0: "one device, no halo", 1: "NCCL send/recv halo refresh per iteration",
                 2: "peer-memory halo refresh per iteration (NVLink stores + release/acquire flags)"}[eng.halo_path()]
    if host_mols is None:
        populate_by_release(eng, n_total, edge_um, t.length_unit)
    else:
        eng.upload(host_mols)

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing: inputs already in HBM
    eng.set_profiling(True)
    sync_all()
    if args.warmup:
        eng.step(args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all()
    st = eng.step(args.steps)
    sync_all()
    ms = torch.tensor([st.device_ms], dtype=torch.float64, device="cuda")
    steps_done = torch.tensor([float(st.molecule_steps)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(steps_done, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    mol_steps = float(steps_done.item())
    value = mol_steps / (ms_total * 1e-3)
    launches = int(st.kernel_launches)
    # global counts after warmup + steps iterations (summed over ranks by the library): the same population evolves
    # at every N, so these are directly comparable between the lines of a scaling run
    counts_species, counts_rules = eng.counts()
    iterations_done = args.warmup + args.steps
    prof_it = max(1, int(st.profiled_iterations))
    fast_ms = st.ms_diffuse / prof_it
    slow_ms = st.ms_diffuse_slow / prof_it
    mol_per_launch = float(st.molecule_steps) / max(1, args.steps)
    deferred_per_launch = float(st.deferred_molecules) / max(1, args.steps)
    peak, peak_src = _peaks()
    # algorithmic bytes of the dominant kernel: the fast pass reads every record (32 B) and its
    # neighbour staging (28 B) and writes the records it finishes (32 B); the slow pass does all three for the
    # molecules deferred to it
    fast_bytes = mol_per_launch * 60.0 + (mol_per_launch - deferred_per_launch) * 32.0
    slow_bytes = deferred_per_launch * B_ALG_DIFFUSE
    if fast_ms >= slow_ms:
        top_kernel, diffuse_ms, top_bytes = "k_diffuse_fast<0>", fast_ms, fast_bytes
    else:
        top_kernel, diffuse_ms, top_bytes = "k_diffuse_fast<1> + k_diffuse_slow", slow_ms, slow_bytes
    achieved = top_bytes / (diffuse_ms * 1e-3) / 1e9 if diffuse_ms > 0 else 0.0

    # ---- end to end through the C ABI with HOST buffers: upload -> ITERS_PER_CALL iterations -> download
    eng.set_profiling(False)
    e2e_steps = 0.0
    cap = int(t.cfg.max_molecules)
    from mcell_b200.model import MolArrays
    out = MolArrays(0)
    out.x, out.y, out.z = _pinned_alloc(cap, np.float64), _pinned_alloc(cap, np.float64), _pinned_alloc(cap, np.float64)
    out.id, out.species, out.flags = _pinned_alloc(cap, np.uint32), _pinned_alloc(cap, np.uint32), _pinned_alloc(cap, np.uint32)
    out.diffusion_time, out.unimol_rxn_time = _pinned_alloc(cap, np.float64), _pinned_alloc(cap, np.float64)
    bytes_per_mol = 52
    has_surf = host_mols is not None and bool((host_mols.wall[:host_mols.n] != 0xFFFFFFFF).any())
    if has_surf:  # Molecule::s travels too: wall, tile, orientation, u, v (+ counted volume)
        out.wall, out.tile, out.counted_volume = _pinned_alloc(cap, np.uint32), _pinned_alloc(cap, np.uint32), _pinned_alloc(cap, np.uint32)
        out.orientation = _pinned_alloc(cap, np.uint32).view(np.int32)
        out.u, out.v = _pinned_alloc(cap, np.float64), _pinned_alloc(cap, np.float64)
        bytes_per_mol += 32
    out.n = cap
    h2d = d2h = 0
    e2e_calls = max(1, args.e2e_calls)
    n_live = eng.download_into(out)
    eng.upload(_view(out, n_live))
    eng.step(ITERS_PER_CALL)  # warm the path once
    n_live = eng.download_into(out)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_calls):
        src = _view(out, n_live)
        eng.upload(src)
        h2d += n_live * bytes_per_mol
        s2 = eng.step(ITERS_PER_CALL)
        e2e_steps += s2.molecule_steps
        n_live = eng.download_into(out)
        counts = eng.counts()
        d2h += n_live * bytes_per_mol + 8 * (len(counts[0]) + len(counts[1]))
    sync_all()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    e2e_total = torch.tensor([e2e_steps], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_total, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_total.item()) / float(dt.item())
    clocks = sampler.result()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu and args.config == 5:
            from oracle import oracle_py as O
            O.build()
            cores = os.cpu_count() or 1
            v, steps, busy = cpu_sample((2, 2, 2), 1, 0, cores)
            single, single_s = cpu_single((2, 2, 2), 1)
            cpu = {"value": v, "unit": "molecule-steps/s", "cores": cores, "kind": "port",
                   "build": "-O3 -march=core2 -ffp-contract=off (the reference's release flags, CMakeLists.txt:58)",
                   "sample": _sample_text((2, 2, 2), cores, 1) + ", %.1f s of CPU work per core" % busy,
                   "single_thread": {"value": single, "cores": 1,
                                     "sample": "one of those seeds alone on the box, 1 iteration, %.1f s" % single_s}}
            try:   # BASELINE.md 3 build (b): the same sample with -O3 -march=native, compiled on this box
                O.build(native=True, force=True)
                vn, _, busy_n = cpu_sample((2, 2, 2), 1, 0, cores, native=True)
                cpu["march_native"] = {"value": vn, "cores": cores, "build": "-O3 -march=native -ffp-contract=off",
                                       "sample": "same sample, %.1f s of CPU work per core" % busy_n}
            except Exception as ex:   # noqa: BLE001
                cpu["march_native"] = {"unavailable": str(ex)[:200]}
        line = {
            "metric": "molecule_steps_per_sec", "value": value, "unit": "molecule-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / max(1, args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "baseline_config": args.config,
                       "molecules": n_total, "box_edge_um": edge_um, "iterations_per_plugin_call": ITERS_PER_CALL,
                       "l2": "inputs (>=3 GB at 1e8 molecules) larger than L2; no flush",
                       "parallelism": "z-slabs x%d, %s" % (world, halo_text), "rng": "philox4x32-10 per molecule"},
            "roofline": {"bound": "hbm", "kernel": top_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": _traffic(top_kernel, n_total) if world == 1 else None,
                         "peak_source": peak_src,
                         "alg_bytes_per_launch": top_bytes, "kernel_ms": diffuse_ms,
                         "kernel_share_of_step": diffuse_ms / (ms_total / max(1, args.steps)),
                         "whole_step_frac": value / world * b_alg_step / 1e9 / peak, "whole_step_alg_bytes_per_molecule": b_alg_step,
                         "ms_diffuse_fast": fast_ms, "ms_diffuse_slow": slow_ms,
                         "deferred_fraction": deferred_per_launch / max(1.0, mol_per_launch),
                         "deferred_by_reason": [int(x) for x in st.deferred_by_reason],
                         "ms_resolve": st.ms_resolve / prof_it, "ms_sort": st.ms_sort / prof_it},
            # one plugin call = upload + ITERS_PER_CALL iterations (= bench steps) + download + counts
            "e2e": {"value": e2e_value, "unit": "molecule-steps/s",
                    "h2d_bytes_per_step": h2d / e2e_calls / ITERS_PER_CALL, "d2h_bytes_per_step": d2h / e2e_calls / ITERS_PER_CALL,
                    "h2d_bytes_per_call": h2d / e2e_calls, "d2h_bytes_per_call": d2h / e2e_calls,
                    "iterations_per_call": ITERS_PER_CALL},
            "counts": {"after_iterations": iterations_done, "species": [int(x) for x in counts_species],
                       "rules": [int(x) for x in counts_rules]},
            "gpu_launches": launches, "clocks": clocks,
            "stats": {k: int(getattr(st, k)) for k in ("bimol_rxns", "unimol_rxns", "vol_mol_vol_mol_collisions",
                                                        "mol_wall_reflections", "resolve_retries", "unresolved_conflicts", "n_live")},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def build_small_config(which):
    """BASELINE.json configs 1-4 (the parity-test configurations) as bench workloads: -> (tables, molecules, workload
    text, algorithmic bytes per molecule-step of the whole step, box edge in um).  The scenario builders
    are the ones the GPU tests use (tests/common.py, tests/test_gpu_fullsize.py)."""
    tests_dir = os.path.join(ROOT, "tests")
    if tests_dir not in sys.path:
        sys.path.insert(0, tests_dir)
    import common as cm
    if which == 1:
        t, mols = cm.free_diffusion_box(n=100000, seed=1, cap_factor=1.25)
        return t, mols, "free diffusion of 1e5 volume molecules (D=1e-6 cm^2/s) in a reflective 1 um cube (BASELINE configs[0])", 64.0, 1.0
    if which == 2:
        t, mols = cm.reactive_box(n=1_000_000, edge_um=2.0, seed=2, p_target=0.1, cap_factor=1.25)
        return t, mols, "A+B->C in a 2 um box, 1e6 molecules, default subpartition grid (BASELINE configs[1])", 160.0, 2.0
    import test_gpu_fullsize as fs
    if which == 3:
        t, mols, _ = fs._config3(400_000, 8_000, seed=3)
        return t, mols, "ligand-receptor on an icosphere of 20 480 triangles, absorptive/transparent classes (BASELINE configs[2])", 64.0, 1.6
    if which == 4:
        t, mols, _, _ = fs._config4(10_000_000, seed=4)
        return t, mols, "synapse-like nested meshes, 163 840 triangles, Ca/calbindin/pumps, 1e7 molecules (BASELINE configs[3])", 160.0, 4.0
    if which == 6:
        # not a BASELINE config: the surface-surface path (react_2D_all_neighbors) at a size where it is measured
        t, mols = cm.surface_reactions(n_a=450_000, n_b=450_000, n_e=100_000, radius_um=5.0, subdivisions=6, box_um=10.4, seed=6,
                                       p=0.1, max_molecules=1_300_000)
        return t, mols, ("surface-surface reactions: 1e6 surface molecules of 5 species diffusing on an icosphere of 20 480 triangles "
                         "(3.5e6 tiles), A+B->C, C->A, A+E->D+E, D+D->A+B (extension, not in BASELINE.json)"), 160.0, 10.4
    raise SystemExit("bench.py: --config must be 1..6")


def _view(m, n):
    from mcell_b200.model import MolArrays
    v = MolArrays(0)
    for k in MolArrays.FIELDS:
        a = getattr(m, k)
        if a is not None and len(a) >= n and len(a) > 0:
            setattr(v, k, a[:n])
    v.n = n
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--molecules", type=int, default=100_000_000)
    ap.add_argument("--config", type=int, default=5, help="BASELINE.json config 1-5 (5 = the headline reactive box; 1-4 single GPU); 6 = surface-surface extension")
    ap.add_argument("--e2e-calls", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cell-edge", type=float, default=0.0, help="device neighbour-cell edge in length units (0 = auto)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
