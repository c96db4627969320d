#!/usr/bin/env python3
"""Per-source-line and per-stall-reason breakdown of one kernel in an ncu report (needs the libmcx.so of the capture):
python tools/ncu_lines.py rep.ncu-rep <mangled-symbol> <kernel-name-substring> <dir holding mcell_b200/libmcx.so> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, symbol, want, root = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif r and r[0].startswith("0x") and cur is not None:
        cur["data"].append(r)
blk = [b for b in blocks if want in b["name"]][0]
hdr, data = blk["hdr"], blk["data"]
ie, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "mcell_b200", "libmcx.so")], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, "mcx_kernels.sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
st = [i for i, l in enumerate(dis) if l.startswith(".text.") and symbol in l][0]
cur, insts = None, []
for l in dis[st + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append(cur)
assert len(insts) == len(data), (len(insts), len(data))
agg = collections.OrderedDict()
stalls = collections.Counter()
tot = tots = 0
for ln, r in zip(insts, data):
    n, s = int(r[ie]), int(r[isamp])
    a = agg.setdefault(ln, [0, 0, 0, collections.Counter()])
    a[0] += n; a[1] += s; a[2] += int(r[it])
    for i, h in stall_cols:
        v = int(r[i] or 0)
        a[3][h] += v; stalls[h] += v
    tot += n; tots += s
print("# %s: %d warp instructions, %d samples" % (blk["name"], tot, tots))
print("# stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(1, sum(stalls.values()))) for k, v in stalls.most_common(8)))
print("%-26s %7s %6s %8s  %s" % ("file:line", "inst%", "thr", "samples%", "top stalls"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    ts = ", ".join("%s %d%%" % (h[6:], 100 * v / max(1, sum(a[3].values()))) for h, v in a[3].most_common(2))
    print("%-26s %6.1f%% %6.1f %7.1f%%  %s" % (k, 100 * a[0] / tot, a[2] / max(1, a[0]), 100 * a[1] / tots, ts))
