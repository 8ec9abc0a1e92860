"""Small run of every kernel family for compute-sanitizer (memcheck): python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200  # noqa: E402

dev = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.int64)).cuda()
rng = np.random.default_rng(0)
for n, q in [(8, 97), (64, 1099511678977), (1024, 132120577), (4096, 4611686018326724609), (16384, 4611686018326724609), (16384, 132120577)]:
    ring = fheb200.PolynomialRing(n, q)
    a = dev(rng.integers(0, q, size=(5, n), dtype=np.uint64))
    b = dev(rng.integers(0, q, size=(5, n), dtype=np.uint64))
    ring.from_ntt(ring.to_ntt(a))
    ring.multiply(a, b)
    ring.add(a, b); ring.subtract(a, b); ring.negate(a); ring.multiply_scalar(a, 7); ring.pointwise_multiply(a, b)
    if n >= 64:
        ring.tensor_multiply(a.view(-1)[: 4 * n].view(2, 2, n), b.view(-1)[: 4 * n].view(2, 2, n))
ml = fheb200.MultiLimbModularArithmetic([0xFFFFFFFFFFFFFF43, 1])
u = dev(rng.integers(0, 2**63, size=(1001, 2), dtype=np.uint64))
ml.montgomery_mul(u, u); ml.mod_add(u, u); ml.mod_sub(u, u)
for N, q, n, k, bl, lv in [(128, 1099511678977, 5, 1, 4, 3), (256, 4611686018326724609, 3, 2, 10, 2), (1024, 1099511678977, 7, 1, 23, 1)]:
    bsk = rng.integers(0, q, size=(n, (k + 1) * lv, k + 1, N), dtype=np.uint64)
    eng = fheb200.BootstrapEngine(N, q, n, k, bl, lv, bsk)
    ksk = rng.integers(0, q, size=(k * N * 2, n + 1), dtype=np.uint64)
    eng.set_key_switch_key(ksk, n, 5, 2)
    lwe = dev(rng.integers(0, q, size=(9, n + 1), dtype=np.uint64))
    tp = dev(eng.get_default_test_poly())
    eng.bootstrap(lwe, tp)
    g = dev(rng.integers(0, q, size=(3, k + 1, N), dtype=np.uint64))
    eng.external_product(g, 1); eng.cmux(0, g, g.flip(0).contiguous())
cts = dev(rng.integers(0, 1099511678977, size=(777, 2, 1024), dtype=np.uint64))
fheb200.tally_votes(cts, 1024, 1099511678977)
acc = fheb200.CiphertextStreamAccumulator(1024, 1099511678977)
acc.add(cts[:1]); acc.add(cts[1:300]); acc.total()
fheb200.tally_votes(rng.integers(0, 97, size=(1200, 2, 1024), dtype=np.uint64), 1024, 97)  # host pipeline (19 MB)
torch.cuda.synchronize()
print("sanitize smoke done,", fheb200.launch_count(), "launches")
