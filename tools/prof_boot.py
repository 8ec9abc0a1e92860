"""Small driver for profiling the blind-rotation kernel: tfhe-128-fast shape, short batch."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 592
n = int(sys.argv[2]) if len(sys.argv) > 2 else 742
# optional: N q base_log level (default: tfhe-128-fast shape with the substitute prime)
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
q = int(sys.argv[4]) if len(sys.argv) > 4 else 1099511678977
base_log = int(sys.argv[5]) if len(sys.argv) > 5 else 23
level = int(sys.argv[6]) if len(sys.argv) > 6 else 1
rng = np.random.default_rng(1)
bsk = rng.integers(0, q, size=(n, 2 * level, 2, N), dtype=np.uint64)
eng = fheb200.BootstrapEngine(N, q, n, 1, base_log, level, bsk)
lwe = torch.randint(0, q, (batch, n + 1), dtype=torch.int64, device="cuda")
tp = torch.from_numpy(eng.get_default_test_poly().view(np.int64)).cuda()
for _ in range(2):
    out = eng.blind_rotate(lwe, tp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = eng.blind_rotate(lwe, tp)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"N {N} q {q} base_log {base_log} L {level} batch {batch} n {n}: {ms:.3f} ms -> {batch / ms * 1e3:.0f} blind rotations/s")
