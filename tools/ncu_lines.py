"""Stall samples of an .ncu-rep aggregated per source line (read here, no GPU).
    python tools/ncu_lines.py <report> <object.o> <mangled kernel name> [top=40]
The SASS of the report is matched by instruction index against `nvdisasm --print-line-info` of the object
(the innermost inlined location of every instruction)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

rep, obj, fun = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = dis.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(f".text.{fun}:"))
loc = None
locs = []
for l in lines[start + 1:]:
    if l.startswith("//---") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        locs.append(loc)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = rows[2:]
assert len(ins) == len(locs), (len(ins), len(locs))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
samp = Counter()
exe = Counter()
st = defaultdict(Counter)
for r, lc in zip(ins, locs):
    samp[lc] += int(r[ix["# Samples"]])
    exe[lc] += int(r[ix["Instructions Executed"]])
    for c in stall_cols:
        st[lc][c[6:]] += int(r[ix[c]])
ts, te = sum(samp.values()), sum(exe.values())
byfile = Counter()
for (f, _), v in samp.items():
    byfile[f] += v
print("samples per file:", ", ".join(f"{f}={100.0 * v / ts:.1f}%" for f, v in byfile.most_common()))
for lc, v in samp.most_common(top):
    print(f"{lc[0]:18s}:{lc[1]:4d}  {100.0 * v / ts:5.1f}% samples {100.0 * exe[lc] / te:5.1f}% exec   " + " ".join(f"{k}={100.0 * n / max(v, 1):.0f}%" for k, n in st[lc].most_common(4)))
