"""Driver for profiling the fused polynomial product: python tools/prof_polymul.py [logn] [q] [batch]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 14
q = int(sys.argv[2]) if len(sys.argv) > 2 else 4611686018326724609
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
n = 1 << logn
ring = fheb200.PolynomialRing(n, q)
a = [torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda") for _ in range(3)]
b = [torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda") for _ in range(3)]
c = torch.empty_like(a[0])
for i in range(3):
    ring.multiply(a[i], b[i], out=c)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(12):
    ring.multiply(a[i % 3], b[i % 3], out=c)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 12
print(f"polymul N={n} batch={batch}: {ms:.4f} ms -> {batch / ms * 1e3:.0f} products/s")
