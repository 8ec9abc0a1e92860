"""End-to-end timing of forward+inverse through the C ABI with pinned host buffers."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200
N, B, q = 16384, 1024, 4611686018326724609
ntt = fheb200.NTTProcessor(N, q)
hx = torch.randint(0, q, (B, N), dtype=torch.int64).pin_memory()
hy = torch.empty_like(hx).pin_memory(); hz = torch.empty_like(hx).pin_memory()
x, y, z = (t.numpy().view(np.uint64) for t in (hx, hy, hz))
for _ in range(2):
    ntt.forward_ntt(x, out=y); ntt.inverse_ntt(y, out=z)
t0 = time.perf_counter()
it = 6
for _ in range(it):
    ntt.forward_ntt(x, out=y); ntt.inverse_ntt(y, out=z)
dt = (time.perf_counter() - t0) / it
assert np.array_equal(z, x)
print(f"chunk {os.environ.get('FHEB_PIPE_CHUNK_MB','8')} MB: {dt*1e3:.2f} ms/step  {2*B*N/dt/1e9:.2f} Gcoeff/s  {4*B*N*8/dt/1e9:.1f} GB/s PCIe both ways")
