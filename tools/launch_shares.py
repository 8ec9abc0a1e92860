"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_shares.py <csv> [--exclude-microbench].
--exclude-microbench drops the kernels of tools/microbench/pipes (`k<MODE>`), which bench.py runs in a child process BEFORE the
timed region to measure the pipe rates of its roofline and which ncu lists with everything else, and the spin kernel (torch.cuda._sleep) the timed steps are queued behind."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
skip_mb = "--exclude-microbench" in sys.argv
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = re.sub(r"\(.*", "", r[ki]).strip()
    if skip_mb and (re.match(r"void k<\d+>", name) or "spin_kernel" in name):
        continue
    tot[name][0] += 1
    tot[name][1] += v * scale
total = sum(v[1] for v in tot.values())
print("kernel, launches, total_us, share_of_profiled_time")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k}, {n}, {us:.1f}, {100.0 * us / total:.1f}%")
