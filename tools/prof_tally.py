"""Driver for profiling the tally kernel: python tools/prof_tally.py [ballots]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
n, q = 1024, 1099511678977
cts = torch.empty((count, 2, n), dtype=torch.int64, device="cuda")
fheb200.synth_ballots(cts, 0, count, n, q, 1)
for _ in range(3):
    out = fheb200.tally_votes(cts, n, q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = fheb200.tally_votes(cts, n, q)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"tally {count} ballots: {ms * 1e3:.1f} us -> {count / ms / 1e3:.1f} M ballots/s, {count * 16384 / ms / 1e6:.0f} GB/s")
