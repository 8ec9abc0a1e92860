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
# relinearisation, wire formats, unreduced test polynomial (general blind-rotation kernel)
q62 = 4611686018326724609
ring = fheb200.PolynomialRing(4096, q62)
rk = fheb200.RelinearizationKey(ring, rng.integers(0, q62, size=(3, 2, 4096), dtype=np.uint64), 20, 3)
rk.relinearize(dev(rng.integers(0, q62, size=(5, 3, 4096), dtype=np.uint64)))
rk.relinearize(rng.integers(0, q62, size=(2, 3, 4096), dtype=np.uint64))
qt = 1099511678977
ballots = rng.integers(0, qt, size=(67, 2, 2, 1024), dtype=np.uint64)
recs = [fheb200.serialize_ballot(ballots[i], qt, i) for i in range(67)]
recs[3] = recs[3][:-9]
recs[5] = recs[5][:40]
offs = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.uint64)
blob = b"".join(recs)
fheb200.ingest_ballots(blob, 67, 2, 1024, qt, offsets=offs)
wire = torch.from_numpy(np.frombuffer(blob + b"\0" * ((-len(blob)) % 8), dtype=np.uint8).copy()).cuda()
fheb200.ingest_ballots(wire[: len(blob)], 67, 2, 1024, qt, offsets=offs)
small = [fheb200.serialize_ballot(ballots[i, :1, :, :16].copy(), qt, i) for i in range(9)]
fheb200.ingest_ballots(b"".join(small), 9, 1, 16, qt)
bsk = rng.integers(0, qt, size=(4, 2, 2, 1024), dtype=np.uint64)
eng = fheb200.BootstrapEngine(1024, qt, 4, 1, 23, 1, bsk)
tp = eng.get_default_test_poly().copy()
tp[::5] += np.uint64(qt)
eng.blind_rotate(dev(rng.integers(0, qt, size=(6, 5), dtype=np.uint64)), dev(tp))
torch.cuda.synchronize()
print("sanitize smoke done,", fheb200.launch_count(), "launches")
