#!/usr/bin/env python3
"""Summarise ncu output brought back in gpurun_out/ into text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/x_launches.csv            > profiles/rNN_launches.txt
  python tools/ncu_summary.py full gpurun_out/x.ncu-rep [kernel-symbol]      > profiles/rNN_kernel.txt

`launches`: per-kernel count / total / share of the `--metrics gpu__time_duration.sum` pass.
`full`: key counters of a `--set full` capture (DRAM bytes, throughput, occupancy, divergence) and, when the
in-tree libmcx.so still matches the capture, the hottest source lines (SASS joined with nvdisasm line info).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "nsecond": 1e-3}.get(u, 1.0)
        agg.setdefault(row["Kernel Name"].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-28s %6s %12s %10s %10s %7s" % ("kernel", "n", "sum_us", "mean_us", "max_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-28s %6d %12.1f %10.1f %10.1f %6.1f%%" % (k, len(v), sum(v), sum(v) / len(v), max(v), 100 * sum(v) / tot))
    print("%-28s %6d %12.1f" % ("TOTAL", sum(len(v) for v in agg.values()), tot))


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__local_memory_size" if False else "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def full(rep, symbol=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("== %s  (id %s)" % (name, r[0]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-68s %14s %s" % (k, r[i], units[i]))
        rd = float(r[hdr.index("dram__bytes_read.sum")]); ru = units[hdr.index("dram__bytes_read.sum")]
        wr = float(r[hdr.index("dram__bytes_write.sum")]); wu = units[hdr.index("dram__bytes_write.sum")]
        sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        print("  %-68s %14.1f MB per launch" % ("traffic = dram read + write", (rd * sc[ru] + wr * sc[wu]) / 1e6))
    if symbol:
        hot_lines(rep, symbol)


def hot_lines(rep, symbol):
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(sass.splitlines()))
    start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not start:
        return
    hdr = rows[start[0]]
    data = []
    for r in rows[start[0] + 1:]:
        if not r or not r[0].startswith("0x"):
            break
        data.append(r)
    ia, ie, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "mcell_b200", "libmcx.so")], cwd=tmp, capture_output=True)
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, "mcx_kernels.sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
    st = [i for i, l in enumerate(dis) if l.startswith(".text.") and symbol in l]
    if not st:
        print("(symbol %s not found in the in-tree build)" % symbol)
        return
    cur, insts = None, []
    for l in dis[st[0] + 1:]:
        if l.startswith("\t.section"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            insts.append((int(m.group(1), 16), cur))
    if len(insts) != len(data):
        print("(in-tree build differs from the captured binary: %d vs %d instructions; no line table)" % (len(insts), len(data)))
        return
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for (off, cur), r in zip(insts, data):
        a = agg[cur]
        a[0] += int(r[ie]); a[1] += int(r[it]); a[2] += int(r[isamp])
    tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
    print("\n# hottest source lines of %s (warp instructions %d, samples %d)" % (symbol, tot, tots))
    print("%-26s %8s %12s %9s" % ("file:line", "inst%", "avg_threads", "samples%"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][2])[:30]:
        print("%-26s %7.1f%% %12.1f %8.1f%%" % ("%s:%s" % k if k else "?", 100 * a[0] / tot, a[1] / max(1, a[0]), 100 * a[2] / max(1, tots)))


def traffic(rep, molecules):
    """profiles/traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel of a
    --set full capture, keyed by kernel name; bench.py reports it as roofline.traffic for the same workload size."""
    import json
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ik = hdr.index("Kernel Name")
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    out = {}
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[ik]).split("(")[0]
        b = float(r[ir]) * sc[units[ir]] + float(r[iw]) * sc[units[iw]]
        e = out.setdefault(name, {"molecules": int(molecules), "dram_bytes_per_launch": 0.0, "launches": 0})
        if b > e["dram_bytes_per_launch"]:
            e["dram_bytes_per_launch"] = b  # the largest launch of that name (the others are the near-empty rounds)
        e["launches"] += 1
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
