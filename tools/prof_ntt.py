"""Small driver for timing/profiling the batched transform kernels: python tools/prof_ntt.py [logn] [q] [batch]."""
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
xs = [torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda") for _ in range(4)]
y = torch.empty_like(xs[0])
z = torch.empty_like(xs[0])


def timeit(fn, iters=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


f = timeit(lambda i: ring.to_ntt(xs[i % 4], out=y))
v = timeit(lambda i: ring.from_ntt(xs[i % 4], out=z))
m = timeit(lambda i: ring.multiply(xs[i % 4], xs[(i + 1) % 4], out=z))
c = batch * n
print(f"N={n} q={q} batch={batch}: forward {f:.4f} ms ({c / f / 1e6:.1f} Gcoeff/s, {16 * c / f / 1e6:.0f} GB/s)  "
      f"inverse {v:.4f} ms ({c / v / 1e6:.1f} Gcoeff/s)  polymul {m:.4f} ms ({batch / m * 1e3:.0f}/s, {24 * c / m / 1e6:.0f} GB/s)")
