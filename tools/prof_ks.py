"""Timing of sample extraction + key switching at the tfhe-128-fast shape (synthetic key)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200
N, q, n, batch = 1024, 1099511678977, 742, 4096
rng = np.random.default_rng(1)
bsk = rng.integers(0, q, size=(n, 2, 2, N), dtype=np.uint64)
eng = fheb200.BootstrapEngine(N, q, n, 1, 23, 1, bsk)
for level, base_log in [(1, 23), (3, 4), (8, 5)]:
    ksk = torch.randint(0, q, (N * level, n + 1), dtype=torch.int64, device="cuda")
    eng.set_key_switch_key(ksk, n, base_log, level)
    ext = torch.randint(0, q, (batch, N + 1), dtype=torch.int64, device="cuda")
    out = torch.empty((batch, n + 1), dtype=torch.int64, device="cuda")
    for _ in range(2):
        eng.key_switch(ext, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.key_switch(ext, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    macs = batch * N * level * (n + 1)
    print(f"key switch level={level} base_log={base_log}: {ms:.3f} ms for {batch} ciphertexts ({macs / ms / 1e6:.1f} G mul-reduce/s)")
