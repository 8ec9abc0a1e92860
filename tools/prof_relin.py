"""Timing of relinearisation: python tools/prof_relin.py [logn] [levels] [base_log] [batch] [q]  (FHEB_RELIN_UNFUSED=1: the five-launch path)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 12
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 4
base_log = int(sys.argv[3]) if len(sys.argv) > 3 else 16
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
q = int(sys.argv[5]) if len(sys.argv) > 5 else 4611686018326724609
n = 1 << logn
ring = fheb200.PolynomialRing(n, q)
keys = torch.randint(0, q, (levels, 2, n), dtype=torch.int64, device="cuda")
rk = fheb200.RelinearizationKey(ring, keys, base_log, levels)
ct3 = [torch.randint(0, q, (batch, 3, n), dtype=torch.int64, device="cuda") for _ in range(2)]
out = torch.empty((batch, 2, n), dtype=torch.int64, device="cuda")
for i in range(3):
    rk.relinearize(ct3[i % 2], out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
e0.record()
for i in range(iters):
    rk.relinearize(ct3[i % 2], out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"relinearize N={n} levels={levels} base_log={base_log} batch={batch} q={q} "
      f"{'unfused' if os.environ.get('FHEB_RELIN_UNFUSED') else 'fused'}: {ms:.4f} ms, {batch / ms / 1e3:.3f} M ciphertexts/s, "
      f"{(levels + 2) * n * batch / ms / 1e6:.1f} G coefficient-transforms/s")
