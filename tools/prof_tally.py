"""Single-GPU sweep of the tally kernel's shape knobs (blocks per SM, ballots per work item) at the per-rank size of
the 8-GPU run (131 072 ballots = 2.1 GB) and at the 1-GPU size (1M ballots = 16.4 GB).  CUDA events, 30 iterations.

    python tools/prof_tally.py            # prints one line per configuration
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import fheb200

QT = 1099511678977
n = 1024
fheb200.initialize(0)
dev = torch.device("cuda:0")
big = torch.empty((1 << 20, 2, n), dtype=torch.int64, device=dev)
fheb200.synth_ballots(big, 0, 1 << 20, n, QT, 0xB200)
out = torch.empty((2, n), dtype=torch.int64, device=dev)
ref = None


def timed(count, iters=30):
    global ref
    # rotate over the 16 GB so that every iteration streams from HBM (count * 16 KB >> L2 anyway)
    views = [big[i * count:(i + 1) * count] for i in range((1 << 20) // count)][:8]
    for v in views[:3]:
        fheb200.tally_votes(v, n, QT, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fheb200.tally_votes(views[i % len(views)], n, QT, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


if len(sys.argv) > 2 and sys.argv[1] == "--one":  # the shipped shape only (ncu capture): python tools/prof_tally.py --one 131072
    count = int(sys.argv[2])
    us = timed(count, 6)
    print(f"count={count}: {us:.1f} us  {count * 16384 / us / 1e6:.3f} TB/s")
    sys.exit(0)

for count in (131072, 1 << 20):
    for bpsm in (2, 3, 4, 5, 6, 8):
        for item in (8, 16, 32, 64, 128, 256):
            os.environ["FHEB_EXP_TALLY_BPSM"] = str(bpsm)
            os.environ["FHEB_EXP_TALLY_ITEM"] = str(item)
            us = timed(count, 30 if count == 131072 else 6)
            print(f"count={count:8d} bpsm={bpsm} item={item:4d}: {us:9.1f} us  {count * 16384 / us / 1e6:7.3f} TB/s", flush=True)
# result words must not depend on the shape
os.environ["FHEB_EXP_TALLY_BPSM"], os.environ["FHEB_EXP_TALLY_ITEM"] = "4", "32"
a = fheb200.tally_votes(big[:131072], n, QT).clone()
os.environ["FHEB_EXP_TALLY_BPSM"], os.environ["FHEB_EXP_TALLY_ITEM"] = "7", "24"
b = fheb200.tally_votes(big[:131072], n, QT).clone()
print("shape-independent words:", bool(torch.equal(a, b)))
