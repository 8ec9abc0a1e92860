"""Driver for profiling the FHEV ballot ingest (validate + unpack kernels): python tools/prof_ingest.py [count] [N]."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200  # noqa: E402

cnt = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
q = 1099511678977
one = np.random.default_rng(3).integers(0, q, size=(64, 1, 2, n), dtype=np.uint64)
recs = [fheb200.serialize_ballot(one[i], q, i) for i in range(64)]
blob = b"".join(recs) * (cnt // 64)
wire = torch.from_numpy(np.frombuffer(blob + b"\0" * ((-len(blob)) % 8), dtype=np.uint8).copy()).cuda()
offs = np.arange(cnt + 1, dtype=np.uint64) * np.uint64(len(recs[0]))
out = torch.empty((cnt, 1, 2, n), dtype=torch.int64, device="cuda")
for _ in range(2):
    st = fheb200.ingest_ballots(wire, cnt, 1, n, q, offsets=offs, out=out)[1]
assert not st.any()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fheb200.ingest_ballots(wire, cnt, 1, n, q, offsets=offs, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
gb = cnt * (2 * len(recs[0]) + 16 * n) / 1e9
print(f"ingest {cnt} ballots N={n}: {ms:.3f} ms -> {cnt / ms * 1e3 / 1e6:.1f} M ballots/s, {gb / ms * 1e3:.0f} GB/s algorithmic")
