#!/usr/bin/env python3
"""Walk the SASS of one kernel in an ncu report with source lines and per-instruction active threads:
python tools/ncu_sass_walk.py rep.ncu-rep <mangled-symbol> [kernel-name-substring]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.environ.get("WALK_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rep, symbol = sys.argv[1], sys.argv[2]
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
# several kernels: pick the block whose "Kernel Name" row matches
blocks = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif r and r[0].startswith("0x") and cur is not None:
        cur["data"].append(r)
want = sys.argv[3] if len(sys.argv) > 3 else symbol
blk = [b for b in blocks if want in b["name"] or any(tok in b["name"] for tok in re.findall(r"k_[a-z_]+", symbol))]
blk = blk[0]
hdr, data = blk["hdr"], blk["data"]
ie, it, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "mcell_b200", "libmcx.so")], cwd=tmp, capture_output=True)
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
for k, (ln, r) in enumerate(zip(insts, data)):
    n = int(r[ie])
    if n == 0:
        continue
    print("%5d %-24s %10d %5.1f %6s  %s" % (k, ln, n, int(r[it]) / n, r[isamp], r[isrc][:70]))
