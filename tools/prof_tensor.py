"""Timing of the tensor product (EncryptionEngine::multiply): python tools/prof_tensor.py [logn] [batch] [q]  (FHEB_TENSOR_UNFUSED=1: four launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 12
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
q = int(sys.argv[3]) if len(sys.argv) > 3 else 4611686018326724609
n = 1 << logn
ring = fheb200.PolynomialRing(n, q)
a = [torch.randint(0, q, (batch, 2, n), dtype=torch.int64, device="cuda") for _ in range(2)]
b = [torch.randint(0, q, (batch, 2, n), dtype=torch.int64, device="cuda") for _ in range(2)]
out = torch.empty((batch, 3, n), dtype=torch.int64, device="cuda")
for i in range(3):
    ring.tensor_multiply(a[i % 2], b[i % 2], out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
e0.record()
for i in range(iters):
    ring.tensor_multiply(a[i % 2], b[i % 2], out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"tensor product N={n} batch={batch} q={q} {'unfused' if os.environ.get('FHEB_TENSOR_UNFUSED') else 'fused'}: {ms:.4f} ms, "
      f"{batch / ms / 1e3:.3f} M products/s, {7 * n * batch / ms / 1e6:.1f} G coefficient-transforms/s, {56 * n * batch / ms / 1e6:.0f} GB/s algorithmic")
