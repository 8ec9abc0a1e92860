"""Bulk-async (TMA) landing buffer against plain global loads in the first pass of the transform kernels: forward and
inverse, degrees 2^6..2^13, the three arithmetic modes.  Bit-exactness of the TMA path is checked against the plain path on
every case (and the plain path is what the parity suite pins to the oracle).  One process per setting (the switch is read
once): python tools/prof_tma.py  ->  table."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    sys.path.insert(0, ROOT)
    import torch

    import fheb200

    torch.manual_seed(1234)  # the same inputs in both runs: the checksums must agree
    res = {}
    for logn in (6, 8, 10, 11, 12, 13):
        n = 1 << logn
        for q, tag in ((4611686018326724609, "int64"), (1099511678977, "fp64"), (132120577, "u32")):
            if (q - 1) % (2 * n):
                continue
            ring = fheb200.PolynomialRing(n, q)
            batch = (1 << 24) >> logn
            xs = [torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda") for _ in range(4)]
            y = torch.empty_like(xs[0])
            out = []
            for fn in (ring.to_ntt, ring.from_ntt):
                for i in range(3):
                    fn(xs[i], out=y)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(20):
                    fn(xs[i % 4], out=y)
                e1.record()
                torch.cuda.synchronize()
                out.append(e0.elapsed_time(e1) / 20 * 1e3)
            # checksum of a ragged batch (last group partly filled) for the bit-exactness comparison between the two runs
            r = ring.to_ntt(xs[0][: batch - 3].contiguous())
            r2 = ring.from_ntt(xs[1][:5].contiguous())
            res[f"{logn} {tag}"] = (out[0], out[1], int(r.sum().item()), int(r2.sum().item()))
            del xs, y
    for k, v in res.items():
        print(k, *v)
    sys.exit(0)

runs = {}
for tma in ("0", "1"):
    env = dict(os.environ, FHEB_TMA=tma)
    out = subprocess.run([sys.executable, __file__, "--worker"], env=env, capture_output=True, text=True)
    if out.returncode != 0:
        print(out.stderr[-3000:])
        sys.exit(1)
    for line in out.stdout.strip().splitlines():
        logn, tag, f, i, c1, c2 = line.split()
        runs.setdefault((int(logn), tag), {})[tma] = (float(f), float(i), c1, c2)
print("log2N mode   forward us (plain -> TMA)   inverse us (plain -> TMA)   same words")
for (logn, tag), r in sorted(runs.items()):
    a, b = r["0"], r["1"]
    print(f"{logn:5d} {tag:6s} {a[0]:8.1f} -> {b[0]:8.1f} ({100 * (a[0] / b[0] - 1):+5.1f} %)   {a[1]:8.1f} -> {b[1]:8.1f} ({100 * (a[1] / b[1] - 1):+5.1f} %)   {a[2:] == b[2:]}")
