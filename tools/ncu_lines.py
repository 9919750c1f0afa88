#!/usr/bin/env python3
"""Attribute ncu per-SASS-instruction counts to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <libofdmx.so> [top N]
Joins `ncu --page source --csv` (per SASS address: instructions executed, stall samples) with
`nvdisasm --print-line-info` of the cubin embedded in the library."""
import csv
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, kre, lib = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kname = rows[0][1]
hdr = rows[1]
ia, iso, ie, ist = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
reason_cols = {c[6:]: i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c}
data, seen = [], set()
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        continue
    if r[ia] in seen:
        break
    seen.add(r[ia])
    data.append((int(r[ia], 16), r[iso], int(r[ie] or 0), int(r[ist] or 0),
                 {c: int(r[i] or 0) for c, i in reason_cols.items()}))
base = data[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
kbase = re.match(r"(?:void )?(\w+)", kname).group(1)
targs = re.search(r"<([^>]*)>", kname)
tm = ""
if targs:
    for a in targs.group(1).split(","):
        m = re.match(r"\s*\((\w+)\)(\d+)", a)
        tm += {"int": "Li%sE", "bool": "Lb%sE"}[m.group(1)] % m.group(2)
pat = r"(_Z\d*" + re.escape(kbase) + ("I" + tm + r"E\w*" if tm else r"\w*") + ")"
mangled = re.search(pat, dis).group(1)
sec = dis[dis.index(".section\t.text." + mangled):]
sec = sec[: sec.index(".section", 20)] if ".section" in sec[20:] else sec
cur, off2line = ("?", 0), {}
for line in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", line)
    if m:
        off2line[int(m.group(1), 16)] = cur
by = collections.Counter()
st = collections.Counter()
tot = 0
why = collections.defaultdict(collections.Counter)
allwhy = collections.Counter()
for a, s, n, w, rs in data:
    k = off2line.get(a - base, ("?", 0))
    by[k] += n
    st[k] += w
    tot += n
    why[k].update(rs)
    allwhy.update(rs)
print("kernel", kbase, "total warp-instr", tot, "stall samples", sum(st.values()))
srcs = {}
for (f, l), n in by.most_common(top):
    if f not in srcs:
        cands = glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", f))
        srcs[f] = open(cands[0]).read().splitlines() if cands else []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    top3 = " ".join("%s:%d" % kv for kv in why[(f, l)].most_common(3) if kv[1])
    print("%-22s L%-4d %9d %5.1f%%  stall %5.1f%%  [%s]  %s" % (f, l, n, 100.0 * n / tot, 100.0 * st[(f, l)] / max(1, sum(st.values())), top3, text[:70]))
print("stall reasons overall:", " ".join("%s:%.1f%%" % (k, 100.0 * v / max(1, sum(allwhy.values()))) for k, v in allwhy.most_common(8)))
# per-file totals
pf, ps = collections.Counter(), collections.Counter()
for (f, l), n in by.items():
    pf[f] += n
    ps[f] += st[(f, l)]
for f, n in pf.most_common():
    print("file %-24s instr %5.1f%%  stall %5.1f%%" % (f, 100.0 * n / tot, 100.0 * ps[f] / max(1, sum(st.values()))))
