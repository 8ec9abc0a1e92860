"""Secondary workloads of bench.py: the other BASELINE.json configs, each with its own roofline.

Every entry: {"value", "unit", "ms", "roofline": {...}} measured with CUDA events on torch's
current stream (the stream the library launches on), inputs resident in HBM and larger than L2
(or rotated) between timed iterations.
"""
from __future__ import annotations

import statistics

import numpy as np

Q62 = 4611686018326724609
QT = 1099511678977
Q27 = 132120577


def _time(torch, fn, iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in evs)


def _hbm(peak, algo_bytes, ms):
    ach = algo_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}


def _hash64(t):
    """FNV-1a over ALL result words (as 64-bit units), so that equal hashes across N mean equal tallies."""
    h = 0xCBF29CE484222325
    for w in t.detach().cpu().numpy().view(np.uint64).reshape(-1).tolist():
        h = ((h ^ w) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _tally_parity(fhe, torch, dist, world, rank, dev, n, q, per_rank=65536):
    """Parity of the SHARDED tally where the driver runs it: every rank tallies a 65 536-ballot slice of synthetic
    ballots through the same ShardedTally path as the timed loop (fused peer exchange when world > 1) and the result
    is compared, all 2N words, with the CPU oracle's tally of per-chunk oracle tallies (each rank folds its own slice
    with the oracle, the per-rank oracle tallies are gathered and folded by the oracle again).  The oracle is the
    checker here, outside every timed region."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    try:
        from oracle_bindings import Oracle

        orc = Oracle()
    except Exception as exc:  # the oracle is test infrastructure: say so rather than fail the bench
        return {"parity_vs_oracle": None, "parity_note": f"oracle unavailable: {exc}"}
    cts = torch.empty((per_rank, 2, n), dtype=torch.int64, device=dev)
    fhe.synth_ballots(cts, (1 << 30) + rank * per_rank, per_rank, n, q, 0x0B200)
    st = fhe.ShardedTally(n, q)
    got = None
    for _ in range(2):  # both inbox parities
        got = st.tally(cts)
    torch.cuda.synchronize()
    host = cts.cpu().numpy().view(np.uint64)
    mine = np.stack([orc.tally(host[i:i + 8192], q) for i in range(0, per_rank, 8192)])
    mine = orc.tally(mine, q)
    if world > 1:
        parts = [torch.empty((2, n), dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(mine.view(np.int64)).to(dev))
        exp = orc.tally(np.stack([p.cpu().numpy().view(np.uint64) for p in parts]), q)
    else:
        exp = mine
    ok = bool(np.array_equal(got.cpu().numpy().view(np.uint64), exp))
    flag = torch.tensor([int(ok)], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"parity_vs_oracle": bool(flag.item() == 1),
            "parity_note": f"{per_rank} ballots per rank through the timed path, all {2 * n} words vs the CPU oracle on every rank"}


def _native_latency():
    """tools/latency --json (built by __graft_entry__.build()): small-problem latency through the C ABI from C++."""
    import json
    import os
    import subprocess

    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "latency")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe, "--json"], capture_output=True, text=True, timeout=120)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
    except Exception:
        return None
    return None


def _pipe(mix_key, units, ms, pipes, unit_name, hbm=None):
    """Pipe roofline of a compute-bound line (bench_roofline.pipe_roofline) with the HBM fraction beside it."""
    import bench_roofline

    try:
        r = bench_roofline.pipe_roofline(mix_key, units, ms * 1e-3, pipes, unit_name)
    except Exception as exc:
        r = {"bound": "pipe", "achieved": None, "peak": None, "unit": None, "frac": None, "error": repr(exc)}
    if hbm is not None:
        r["hbm"] = {k: hbm[k] for k in ("achieved", "peak", "unit", "frac")}
    return r


def run(fhe, torch, dist, world, rank, dev, barrier, max_over_ranks, peak, pipes=None):
    out = {}
    pipes = pipes or {}
    gen = torch.Generator(device=dev).manual_seed(99 + rank)

    # ---- the published M4 Max rows use q = 132120577 at every size: same transform; 27-bit primes run the 32-bit kernels
    # (MODE_U32), the same prime with those switched off and the 41-bit tfhe prime run on the FP64 pipe (q < 2^42)
    import os

    for n, q, tag, mix, no_u32 in ((16384, Q27, "q132120577", "ntt_forward_n16384_q27_u32", False),
                                   (16384, Q27, "q132120577_fp64mode", "ntt_forward_n16384_q27_fp64", True),
                                   (1024, QT, "q1099511678977", "ntt_forward_n1024_qt_fp64", False)):
        if no_u32:
            os.environ["FHEB_NO_U32"] = "1"  # read by the library on every call
        ntt = fhe.NTTProcessor(n, q)
        batch = 1024 if n == 16384 else 16384
        sets = 4
        xs = [torch.randint(0, q, (batch, n), dtype=torch.int64, device=dev, generator=gen) for _ in range(sets)]
        y = torch.empty_like(xs[0])
        z = torch.empty_like(xs[0])

        def fwd_inv(i, ntt=ntt, xs=xs, y=y, z=z):
            ntt.forward_ntt(xs[i % sets], out=y)
            ntt.inverse_ntt(y, out=z)

        ms = _time(torch, fwd_inv, 10)
        os.environ.pop("FHEB_NO_U32", None)
        logn = n.bit_length() - 1
        out[f"ntt_n{n}_{tag}_b{batch}"] = {"value": 2.0 * batch * n / (ms * 1e-3), "unit": "coeff/s", "ms": ms,
                                            # forward + inverse: 2 * (N/2) log2 N butterflies per polynomial; the forward kernel's mix
                                            "roofline": _pipe(mix, 2.0 * batch * (n // 2) * logn, ms, pipes, "butterfly", _hbm(peak, 32.0 * n * batch, ms))}
        del xs, y, z

    # ---- C1: the reference's own CPU-runnable case (test_ntt_processor): N = 1024, q = 132120577, batch 1 - latency only
    ntt1 = fhe.NTTProcessor(1024, Q27)
    x1 = torch.randint(0, Q27, (1, 1024), dtype=torch.int64, device=dev, generator=gen)
    y1, z1 = torch.empty_like(x1), torch.empty_like(x1)

    def one(i):
        ntt1.forward_ntt(x1, out=y1)
        ntt1.inverse_ntt(y1, out=z1)

    ms = _time(torch, one, 50)
    assert torch.equal(z1, x1)
    # The reference's published row for this case is per-call latency of native code (M4 Max, Montgomery: 8.86 us per
    # transform, NTT_(degree=1024).csv:5), so the figure to put beside it is the C ABI called from C++ (tools/latency.cu:
    # CUDA events over 2000 back-to-back calls on one stream, and the same calls replayed from a CUDA graph); the time per
    # pair measured from THIS Python process is ctypes + interpreter time per call and is kept as a note.
    lat = _native_latency() if rank == 0 else None
    if lat and "c1_pair" in lat:
        out["ntt_n1024_q132120577_b1_latency"] = {
            "value": lat["c1_pair"]["stream_us"], "unit": "us per forward+inverse (two launches, device buffers, C ABI called from C++)",
            "ms": lat["c1_pair"]["stream_us"] * 1e-3, "cuda_graph_us": lat["c1_pair"]["graph_us"], "issue_to_done_us": lat["c1_pair"]["issue_to_done_us"],
            "forward_us": lat["c1_forward"]["stream_us"], "inverse_us": lat["c1_inverse"]["stream_us"], "python_ctypes_pair_us": ms * 1e3,
            "source": "tools/latency --json (this run)",
            "roofline": {"bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None}}
        out["mlimb2_montmul_n65536_latency"] = {
            "value": lat["c3_montmul_n65536"]["stream_us"], "unit": "us per call of 65536 two-limb products (device buffers, C ABI called from C++)",
            "ms": lat["c3_montmul_n65536"]["stream_us"] * 1e-3, "cuda_graph_us": lat["c3_montmul_n65536"]["graph_us"],
            "source": "tools/latency --json (this run)",
            "roofline": _hbm(peak, 48.0 * 65536, lat["c3_montmul_n65536"]["graph_us"] * 1e-3)}
    else:
        out["ntt_n1024_q132120577_b1_latency"] = {"value": ms * 1e3, "unit": "us per forward+inverse (two launches, device buffers, called from Python: ctypes time included)", "ms": ms,
                                                  "roofline": {"bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None}}

    # ---- C2: fused polynomial multiplication, batch 1024
    for n in (4096, 16384):
        ring = fhe.PolynomialRing(n, Q62)
        sets = 3 if n == 16384 else 8
        a = [torch.randint(0, Q62, (1024, n), dtype=torch.int64, device=dev, generator=gen) for _ in range(sets)]
        b = [torch.randint(0, Q62, (1024, n), dtype=torch.int64, device=dev, generator=gen) for _ in range(sets)]
        c = torch.empty_like(a[0])
        ms = _time(torch, lambda i: ring.multiply(a[i % sets], b[i % sets], out=c), 10)
        logn = n.bit_length() - 1
        out[f"polymul_n{n}_b1024"] = {"value": 1024 / (ms * 1e-3), "unit": "polymul/s", "ms": ms,
                                       # three transforms per product; the pointwise multiplies are inside the per-butterfly mix
                                       # of the fused kernel's capture (N = 16384; used for N = 4096 as well: same code, 12 stages)
                                       "roofline": _pipe("polymul_n16384_q62", 3.0 * 1024 * (n // 2) * logn, ms, pipes, "butterfly",
                                                         _hbm(peak, 24.0 * n * 1024, ms))}
        del a, b, c

    # ---- C3: two-limb Montgomery products (n = 65536 is launch-bound; 2^24 shows the bandwidth)
    ml = fhe.MultiLimbModularArithmetic([0xFFFFFFFFFFFFFF43, 1])
    for cnt in (65536, 1 << 24):
        a = torch.randint(0, 2**62, (cnt, 2), dtype=torch.int64, device=dev, generator=gen)
        a[:, 1] &= 1
        b = a.flip(0).contiguous()
        r = torch.empty_like(a)
        ms = _time(torch, lambda i: ml.montgomery_mul(a, b, out=r), 10)
        out[f"mlimb2_montmul_n{cnt}"] = {"value": cnt / (ms * 1e-3), "unit": "montmul/s", "ms": ms,
                                          "roofline": _hbm(peak, 48.0 * cnt, ms)}
        del a, b, r

    # ---- C5: ballot tally, sharded across ranks with one all-gather + the combine kernel
    n = 1024
    total_ballots = 1 << 20  # BASELINE config 5: 1M ballots (16.4 GB) sharded over the ranks (strong scaling; >> L2 on every rank)
    per_rank = total_ballots // world
    cts = torch.empty((per_rank, 2, n), dtype=torch.int64, device=dev)
    fhe.synth_ballots(cts, rank * per_rank, per_rank, n, QT, 0xB200)
    st = fhe.ShardedTally(n, QT)
    res = [None]

    def tally(i):
        res[0] = st.tally(cts)

    for i in range(3):
        tally(i)
    barrier()
    # the host-side barrier releases the ranks tens of microseconds apart; two untimed exchanges line the GPUs up on
    # the device before the timed region (every rank waits for every other inside the exchange)
    for i in range(2):
        tally(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    e0.record()
    for i in range(iters):
        tally(i)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / iters
    out["tally_n1024"] = {"value": world * per_rank / (ms * 1e-3), "unit": "ballots/s", "ms": ms,
                          "ballots": world * per_rank, "n_gpus": world, "scaling": "strong (1M ballots in total)",
                          "exchange": ("none (one rank)" if world == 1 else
                                       "fused into the tally kernel over peer memory" if getattr(st, "_peers", None) is not None else
                                       "NCCL all-gather + combine kernel"),
                          "roofline": dict(_hbm(peak, 16384.0 * per_rank, ms),
                                           note="peak = the measured COPY bandwidth (read + write); a read-only stream can exceed it"),
                          "hash64": _hash64(res[0])}
    out["tally_n1024"].update(_tally_parity(fhe, torch, dist, world, rank, dev, n, QT))
    if hasattr(st, "check"):
        st.check()  # raises when an exchange of the timed loop timed out
    del cts

    # ---- N4: multiply -> relinearize chain (tensor product + relinearisation with a pre-transformed key)
    n, levels, base_log = 4096, 4, 16
    ring = fhe.PolynomialRing(n, Q62)
    keys = torch.randint(0, Q62, (levels, 2, n), dtype=torch.int64, device=dev, generator=gen)
    rk = fhe.RelinearizationKey(ring, keys, base_log, levels)
    batch = 2048  # 201 MB of degree-2 ciphertexts per set
    ct3 = [torch.randint(0, Q62, (batch, 3, n), dtype=torch.int64, device=dev, generator=gen) for _ in range(2)]
    o2 = torch.empty((batch, 2, n), dtype=torch.int64, device=dev)
    ms = _time(torch, lambda i: rk.relinearize(ct3[i % 2], out=o2), 10)
    # algorithmic bytes: read 3 polynomials, write 2; integer work: `levels` forward + 2 inverse transforms
    out[f"relinearize_n{n}_L{levels}_b{batch}"] = {"value": batch / (ms * 1e-3), "unit": "ciphertexts/s", "ms": ms,
                                                     "transforms_per_ct": levels + 2,
                                                     "coeff_transforms_per_s": (levels + 2) * n * batch / (ms * 1e-3),
                                                     "launches_per_call": 1,
                                                     # one fused launch (relin_fused.cu); unit = butterflies of the levels forward + 2 inverse
                                                     # transforms, the multiply-accumulates are inside the per-butterfly mix of its own capture
                                                     "roofline": _pipe("relin_fused_n4096_l4", (levels + 2.0) * batch * (n // 2) * 12, ms, pipes,
                                                                       "butterfly", _hbm(peak, 40.0 * n * batch, ms))}
    del ct3, o2, keys
    a2 = [torch.randint(0, Q62, (batch, 2, n), dtype=torch.int64, device=dev, generator=gen) for _ in range(2)]
    b2 = [torch.randint(0, Q62, (batch, 2, n), dtype=torch.int64, device=dev, generator=gen) for _ in range(2)]
    o3 = torch.empty((batch, 3, n), dtype=torch.int64, device=dev)
    ms = _time(torch, lambda i: ring.tensor_multiply(a2[i % 2], b2[i % 2], out=o3), 10)
    out[f"tensor_multiply_n{n}_b{batch}"] = {"value": batch / (ms * 1e-3), "unit": "ciphertext products/s", "ms": ms, "transforms_per_product": 7,
                                              "coeff_transforms_per_s": 7.0 * n * batch / (ms * 1e-3), "launches_per_call": 1,
                                              "roofline": _pipe("tensor_fused_n4096", 7.0 * batch * (n // 2) * 12, ms, pipes, "butterfly",
                                                                _hbm(peak, 56.0 * n * batch, ms))}
    del a2, b2, o3

    # ---- N3: FHEV ballot records -> device ingest (checksum + realign), wire bytes already in HBM
    n, cnt = 1024, 32768
    one = np.random.default_rng(3).integers(0, QT, size=(64, 1, 2, n), dtype=np.uint64)
    recs = [fhe.serialize_ballot(one[i], QT, i) for i in range(64)]
    blob = b"".join(recs) * (cnt // 64)
    pad = (-len(blob)) % 8
    wire = torch.from_numpy(np.frombuffer(blob + b"\0" * pad, dtype=np.uint8).copy()).to(dev)
    offs = np.arange(cnt + 1, dtype=np.uint64) * np.uint64(len(recs[0]))
    ing = torch.empty((cnt, 1, 2, n), dtype=torch.int64, device=dev)
    status = [None]

    def ingest(i):
        status[0] = fhe.ingest_ballots(wire, cnt, 1, n, QT, offsets=offs, out=ing)[1]

    ms = _time(torch, ingest, 5, warm=2)
    assert not status[0].any()
    out["ballot_ingest_n1024"] = {"value": cnt / (ms * 1e-3), "unit": "ballots/s", "ms": ms, "ballots": cnt,
                                  "note": "includes the per-call status copy and synchronisation",
                                  # validate reads the record once, unpack reads it again and writes the aligned words
                                  "roofline": _hbm(peak, (2.0 * len(recs[0]) + 16384.0) * cnt, ms)}
    res = torch.empty((1, 2, n), dtype=torch.int64, device=dev)
    ms = _time(torch, lambda i: fhe.tally_wire(wire, cnt, 1, n, QT, offsets=offs, out=res), 5, warm=2)
    out["tally_wire_n1024"] = {"value": cnt / (ms * 1e-3), "unit": "ballots/s", "ms": ms, "ballots": cnt,
                               "note": "checksum validation + tally straight from the wire bytes (no unpacked ciphertexts)",
                               "roofline": _hbm(peak, 2.0 * len(recs[0]) * cnt, ms)}  # the wire is read twice
    del wire, ing, res
    out["bootstrap_tfhe128fast_shape"] = run_bootstrap(fhe, torch, dist, world, rank, dev, barrier, max_over_ranks, pipes=pipes)
    return out


def run_bootstrap(fhe, torch, dist, world, rank, dev, barrier, max_over_ranks, batch=4096, n=742, iters=2, pipes=None):
    """C4: tfhe-128-fast SHAPE (N=1024, k=1, n=742, base_log=23, L=1) with the substitute prime
    1099511678977 (the preset's 2^40+1 is composite) and a synthetic uniformly random key; each rank
    bootstraps its own `batch` LWE ciphertexts (blind rotation + sample extraction), no collective."""
    N, q, k, base_log, level = 1024, QT, 1, 23, 1
    rng = np.random.default_rng(742)
    bsk = rng.integers(0, q, size=(n, (k + 1) * level, k + 1, N), dtype=np.uint64)
    eng = fhe.BootstrapEngine(N, q, n, k, base_log, level, bsk)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    lwe = torch.randint(0, q, (batch, n + 1), dtype=torch.int64, device=dev, generator=gen)
    tp = torch.from_numpy(eng.get_default_test_poly().view(np.int64)).to(dev)
    out = torch.empty((batch, k * N + 1), dtype=torch.int64, device=dev)
    eng.bootstrap(lwe[:256], tp, out=out[:256])  # warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        eng.bootstrap(lwe, tp, out=out)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / iters
    # the full chain of BootstrapEngine::bootstrap with a key switching key (synthetic, the preset's gadget: one level)
    ksk = torch.randint(0, q, (k * N * level, n + 1), dtype=torch.int64, device=dev, generator=gen)
    eng.set_key_switch_key(ksk, n, base_log, level)
    out_ks = torch.empty((batch, n + 1), dtype=torch.int64, device=dev)
    eng.bootstrap(lwe[:256], tp, out=out_ks[:256])
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(iters):
        eng.bootstrap(lwe, tp, out=out_ks)
    k1.record()
    barrier()
    ms_ks = max_over_ranks(k0.elapsed_time(k1)) / iters
    # integer work actually executed per bootstrap: n steps x ((k+1)L forward + (k+1) inverse transforms of
    # (N/2) log2 N butterflies + (k+1)^2 L N multiply-accumulates)
    bfly = n * ((k + 1) * level + (k + 1)) * (N // 2) * 10
    macs = n * (k + 1) * (k + 1) * level * N
    return {"value": world * batch / (ms * 1e-3), "unit": "bootstraps/s", "ms": ms, "batch_per_gpu": batch, "n_gpus": world,
            "shape": f"N={N} k={k} n={n} base_log={base_log} L={level} q={q}",
            "with_key_switch": {"value": world * batch / (ms_ks * 1e-3), "unit": "bootstraps/s", "ms": ms_ks,
                                "note": "blind rotation + sample extraction + key switching (one level)"},
            "modmul_per_bootstrap": bfly + macs,
            "gmodmul_per_s_per_gpu": (bfly + macs) * batch / (ms * 1e-3) / 1e9,
            # FP64-pipe roofline per GPU: executed FP64 instructions per ciphertext-step from the kernel's ncu capture
            # (profiles/r02_opmix.json) against the DFMA rate measured by tools/microbench/pipes in this run
            "roofline": _pipe("boot_lean_tfhe128_step", float(n) * batch, ms, pipes or {}, "ciphertext-step")}


def cpu_reference(cores: int):
    """The reference's own CPU code (oracle/_ref, compiled from its sources) timed on bounded samples of the secondary
    workloads - SURVEY 8(d): PolynomialRing::multiply, MultiLimbModularArithmetic::montgomery_mul,
    BootstrapEngine::blind_rotate, batch_add - one engine call per unit spread over `cores` host threads.  Test
    infrastructure used as the reported baseline only; rank 0 at N=1."""
    import os
    import sys
    import time

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    from oracle_bindings import RefOracle, ref_available

    if not ref_available():
        return {"unavailable": "oracle/_ref/libref_oracle.so not built"}
    r = RefOracle()
    rng = np.random.default_rng(5)
    out = {}

    def entry(units, dt, unit, sample):
        return {"value": units / dt, "unit": unit, "cores": cores, "kind": "reference", "sample": f"{sample} ({dt:.2f} s)"}

    n = 16384
    ring = r.ring_create(n, Q62)
    a = rng.integers(0, Q62, size=(4 * cores, n), dtype=np.uint64)
    b = rng.integers(0, Q62, size=(4 * cores, n), dtype=np.uint64)
    t0 = time.perf_counter()
    r.ring_op(ring, "multiply", a, b, threads=cores)
    out["polymul_n16384_b1024"] = entry(a.shape[0], time.perf_counter() - t0, "polymul/s", f"{a.shape[0]} products, N={n}")
    r.ring_destroy(ring)

    q2 = np.array([0xFFFFFFFFFFFFFF43, 1], np.uint64)
    ml = r.mlimb_create(q2)
    x = rng.integers(0, 2**62, size=(1 << 20, 2), dtype=np.uint64)
    x[:, 1] &= np.uint64(1)
    t0 = time.perf_counter()
    r.mlimb_op(ml, "montmul", x, x[::-1].copy(), threads=cores)
    out["mlimb2_montmul_n16777216"] = entry(x.shape[0], time.perf_counter() - t0, "montmul/s", f"{x.shape[0]} two-limb products")
    r.mlimb_destroy(ml)

    N, nl = 1024, 742
    h = r.boot_create(N, QT, nl, 1, 23, 1, 4)
    r.boot_import_bsk(h, rng.integers(0, QT, size=(nl, 2, 2, N), dtype=np.uint64))
    lwe = rng.integers(0, QT, size=(cores, nl + 1), dtype=np.uint64)
    tp = r.boot_default_test_poly(h)
    t0 = time.perf_counter()
    r.boot_blind_rotate(h, lwe, tp, threads=cores)
    out["bootstrap_tfhe128fast_shape"] = entry(cores, time.perf_counter() - t0, "bootstraps/s",
                                               f"{cores} blind rotations, one per thread, tfhe-128-fast shape")
    r.boot_destroy(h)

    ring = r.ring_create(1024, QT)
    cts = rng.integers(0, QT, size=(8192, 2, 1024), dtype=np.uint64)
    t0 = time.perf_counter()
    r.tally(ring, cts)
    out["tally_n1024"] = dict(entry(cts.shape[0], time.perf_counter() - t0, "ballots/s", f"{cts.shape[0]} ballots, batch_add"), cores=1)
    r.ring_destroy(ring)
    return out
