"""Per-instruction view of an .ncu-rep source page (read here, no GPU): executed instructions by opcode and
the SASS ranges with the most stall samples.   python tools/ncu_hot.py <report> [bucket=64]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 64
if rep.endswith(".gz"):
    import gzip

    raw = gzip.open(rep, "rt").read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = rows[2:]
ex = Counter()
tot_ex = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_stall = Counter()
for r in ins:
    op = r[ix["Source"]].split()[0]
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    n = int(r[ix["Instructions Executed"]])
    ex[op.split(".")[0]] += n
    tot_ex += n
    for c in stall_cols:
        tot_stall[c] += int(r[ix[c]])
print("executed warp instructions:", tot_ex, " static:", len(ins), " executed-at-least-once:", sum(1 for r in ins if int(r[ix["Instructions Executed"]]) > 0))
for op, n in ex.most_common(22):
    print(f"  {op:12s} {n:12d} {100.0 * n / tot_ex:5.1f} %")
ts = sum(tot_stall.values())
print("stall samples:", ", ".join(f"{c[6:]}={100.0 * v / ts:.1f}%" for c, v in tot_stall.most_common(10)))
print(f"\nSASS ranges of {bucket} instructions: samples, executed, top stalls, first instruction")
S = ix["# Samples"]
tot_s = sum(int(r[S]) for r in ins)
for b in range(0, len(ins), bucket):
    blk = ins[b:b + bucket]
    s = sum(int(r[S]) for r in blk)
    e = sum(int(r[ix["Instructions Executed"]]) for r in blk)
    if s * 200 < tot_s:
        continue
    st = Counter()
    for r in blk:
        for c in stall_cols:
            st[c[6:]] += int(r[ix[c]])
    ops = Counter(r[ix["Source"]].split()[0].split(".")[0] for r in blk)
    print(f"  [{b:5d}] {100.0 * s / tot_s:5.1f}% samples {100.0 * e / tot_ex:5.1f}% exec  " + " ".join(f"{k}={100.0 * v / max(s, 1):.0f}%" for k, v in st.most_common(4)) + "  | " + " ".join(f"{k}:{v}" for k, v in ops.most_common(5)))
