"""Generate tests/golden/*.npz by running the REFERENCE'S OWN code.

Run in the build container (needs oracle/_ref/libref_oracle.so, i.e. /root/reference):

    python tests/golden/make_golden.py                 # everything (keys are random: every file changes)
    python tests/golden/make_golden.py boot_tfhe256    # only the files whose name starts with the given prefix

The reference ships no known-answer vectors for this path, so these files
are outputs of the reference itself (oracle/ref_harness.cpp -> NTTProcessor, PolynomialRing,
MultiLimbModularArithmetic, BootstrapEngine) on seeded inputs; keys come from the
reference's own KeyManager / encrypt_ggsw / generate_key_switch_key (random_device-seeded,
hence stored rather than regenerated).  They pin oracle/fhe_oracle.c on machines where
the reference cannot be built (the GPU box) and are the fixed fixtures of the GPU parity tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_bindings import RefOracle, mt19937_64_coeffs  # noqa: E402

Q27 = 132120577                 # cpp/tests/test_ntt_processor.cpp PRIME_1024
Q62 = 4611686018326724609       # cpp/tests/test_harness.h:143-147
QT = 1099511678977              # substitute for the composite tfhe-128-fast modulus


def ntt_case(r, n, q, seed, polys):
    h = r.ntt_create(n, q)
    fwd, inv, psi, psi_inv, inv_n = r.ntt_tables(h, n)
    x = mt19937_64_coeffs(seed, n * polys, q).reshape(polys, n)
    y = r.ntt_forward(h, x)
    z = r.ntt_inverse(h, x)  # inverse applied to the raw input as well (not only round trips)
    ring = r.ring_create(n, q)
    prod = r.ring_op(ring, "multiply", x[: polys // 2], x[polys // 2 : 2 * (polys // 2)])
    r.ring_destroy(ring)
    r.ntt_destroy(h)
    return dict(n=n, q=np.uint64(q), seed=seed, psi=np.uint64(psi), psi_inv=np.uint64(psi_inv),
                inv_n=np.uint64(inv_n), fwd_head=fwd[:8].copy(), inv_head=inv[:8].copy(),
                x=x, forward=y, inverse=z, product=prod)


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""


def want(name):
    return name.startswith(ONLY)


def save(name, **arrays):
    if want(name):
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)


def main():
    r = RefOracle()
    rng = np.random.default_rng(20261018)

    # --- transforms + polymul (C1, C2 shapes and the reference's own small configs)
    for n, q, seed, polys in [(8, 17, 42, 4), (8, 97, 42, 4), (16, 97, 42, 4), (1024, Q27, 42, 4),
                              (1024, QT, 7, 2), (4096, Q62, 42, 2), (16384, Q62, 123, 2)]:
        if want(f"ntt_n{n}_q{q}"):
            save(f"ntt_n{n}_q{q}", **ntt_case(r, n, q, seed, polys))

    # --- multi-limb (C3 modulus, cpp/tests/test_multi_limb.cpp:143)
    ql = np.array([0xFFFFFFFFFFFFFF43, 1], np.uint64)
    h = r.mlimb_create(ql)
    q_inv, r1, r2 = r.mlimb_constants(h, 2)
    qint = int(ql[0]) + (int(ql[1]) << 64)
    vals = [int.from_bytes(rng.bytes(17), "little") % qint for _ in range(512)]
    vals[:6] = [0, 1, qint - 1, qint - 2, (1 << 64) - 1, 1 << 64]
    ab = np.array([[v & (2**64 - 1), v >> 64] for v in vals], np.uint64)
    a, b = ab[:256], ab[256:]
    save("mlimb_q65", q=ql, q_inv=np.uint64(q_inv), r_mod_q=r1, r2_mod_q=r2,
                        a=a, b=b, montmul=r.mlimb_op(h, "montmul", a, b), add=r.mlimb_op(h, "add", a, b),
                        sub=r.mlimb_op(h, "sub", a, b), to_mont=r.mlimb_op(h, "to_mont", a),
                        from_mont=r.mlimb_op(h, "from_mont", a))
    r.mlimb_destroy(h)

    # --- bootstrap: two small shapes with keys drawn by the reference's own generators
    Q60 = 1152921504606584833  # primes::Q_60_1, cpp/src/parameter_set.cpp:24
    for tag, (N, QB, n, k, base_log, level, with_ksk) in {
        "n128_l3": (128, QT, 8, 1, 4, 3, True),          # engine defaults base_log 4 / level 3; n power of two -> KSK works
        "tfhe_shape_small_n": (1024, QT, 6, 1, 23, 1, False),  # tfhe-128-fast gadget (23,1) at N=1024, short LWE key
        # tfhe-256-secure (parameter_set.cpp:166-190: N=4096, Q_60_1, base_log 10, 3 levels) - the one TFHE preset
        # the reference can instantiate end to end (prime modulus, power-of-two LWE dimension) - with a 4-word LWE key
        "tfhe256_shape_small_n": (4096, Q60, 4, 1, 10, 3, True),
    }.items():
        if not want(f"boot_{tag}"):
            continue
        QT_ = QB
        h = r.boot_create(N, QB, n, k, base_log, level, 4)
        sk = r.boot_keygen(h, with_ksk)
        bsk = r.boot_export_bsk(h)
        lwe = r.boot_encrypt_lwe(h, np.array([0, 1, 2, 3], np.uint64), sk)
        lwe[3, 0] = 0            # exercises the rotation == 0 skip (bootstrap_engine.cpp:566)
        lwe[2, 1] = QT_ - 1      # rotation == 2N (not skipped, identity rotation)
        tp = r.boot_default_test_poly(h)
        acc = r.boot_blind_rotate(h, lwe, tp)
        glwe = rng.integers(0, QT_, size=(k + 1, N), dtype=np.uint64)
        out = dict(N=N, n=n, k=k, base_log=base_log, level=level, q=np.uint64(QT_), t=4, lwe_sk=sk, bsk=bsk, lwe=lwe,
                   test_poly=tp, blind_rotate=acc, glwe=glwe, external_product=r.boot_external_product(h, glwe, 1),
                   cmux=r.boot_cmux(h, 2, glwe, acc[0]), sample_extract=r.boot_sample_extract(h, acc),
                   decompose=r.boot_decompose(h, glwe[0], base_log, level), rotate_5=r.boot_rotate(h, glwe[0], 5),
                   rotate_m3=r.boot_rotate(h, glwe[0], -3), rotate_big=r.boot_rotate(h, glwe[0], N + 7),
                   lut_identity=r.boot_lut(h, 0, 4), lut_negation=r.boot_lut(h, 1, 4), lut_threshold=r.boot_lut(h, 2, 2, 4))
        if with_ksk:
            ksk = r.boot_export_ksk(h)
            out.update(ksk=ksk, key_switch=r.boot_key_switch(h, r.boot_sample_extract(h, acc), n),
                       bootstrap=r.boot_bootstrap(h, lwe, tp, n))
        save(f"boot_{tag}", **out)
        r.boot_destroy(h)

    # --- tally + tensor product (C5 shape scaled down)
    for n, q, m in [(64, QT, 5), (1024, QT, 9)]:
        ring = r.ring_create(n, q)
        cts = rng.integers(0, q, size=(m, 2, n), dtype=np.uint64)
        cts[1, 0, :4] = [q, q + 1, 2**64 - 1, 0]  # unreduced words: mod_add reduces inputs first
        save(f"tally_n{n}_m{m}", n=n, q=np.uint64(q), cts=cts,
                            linear=r.tally(ring, cts), tree=r.tally(ring, cts, tree=True), single=r.tally(ring, cts[1:2]),
                            tensor=r.tensor_multiply(ring, cts[0], cts[2]))
        r.ring_destroy(ring)
    # --- relinearisation: EncryptionEngine::relinearize composed from the reference's PolynomialRing
    # calls (ref_harness.cpp: ref_relinearize).  Keys are uniform words: the arithmetic does not care how they were made.
    rrng = np.random.default_rng(904)
    for n, q, key_count, base_log, level in [(64, Q27, 3, 8, 3), (1024, QT, 16, 0, 0), (4096, Q62, 4, 16, 4), (8, 17, 5, 6, 0)]:
        if not want(f"relin_n{n}"):
            continue
        ring = r.ring_create(n, q)
        ct = rrng.integers(0, q, size=(2, 3, n), dtype=np.uint64)
        ct[1, 0, :3] = [q, 2**64 - 1, q + 5]   # unreduced c0 words: add_inplace reduces both sides
        ct[1, 2, :2] = [2**64 - 1, 2**63 + 12345]  # c2 is only ever read through shifts and masks
        keys = rrng.integers(0, q, size=(key_count, 2, n), dtype=np.uint64)
        out = np.stack([r.relinearize(ring, ct[i], keys, base_log, level) for i in range(2)])
        nokey = r.relinearize(ring, ct[1], keys[:0], base_log, level)
        save(f"relin_n{n}", n=n, q=np.uint64(q), ct=ct, keys=keys, base_log=base_log, level=level, out=out, nokey=nokey)
        r.ring_destroy(ring)
    # --- wire formats: records written by the reference's own BallotSerializer / KeySerializer
    if want("wire_"):
        wrng = np.random.default_rng(709)
        n, q, choices = 64, QT, 2
        ballots = wrng.integers(0, q, size=(6, choices, 2, n), dtype=np.uint64)
        records = [r.serialize_ballot(ballots[i], q, 1_700_000_000 + i) for i in range(6)]
        sizes = np.array([len(x) for x in records], np.uint64)
        keys = wrng.integers(0, q, size=(3, 2, n), dtype=np.uint64)
        eval_key = r.serialize_eval_key(keys, q, 12, 3, 0xABCDEF)
        probe = bytes(wrng.integers(0, 256, size=1000, dtype=np.uint8))
        bsk = wrng.integers(0, q, size=(6, 4, 2, n), dtype=np.uint64)       # 6 GGSWs, (k+1)*L = 4 rows, k = 1
        kskp = wrng.integers(0, q, size=(2, 2, n), dtype=np.uint64)         # KeyManager-style pairs (skipped on ingest)
        boot_key = r.serialize_bootstrap_key(bsk, kskp, 4, 2, q, 0x5EED)
        save("wire_n64", n=n, q=np.uint64(q), choices=choices, ballots=ballots, sizes=sizes,
             records=np.frombuffer(b"".join(records), np.uint8), keys=keys, eval_key=np.frombuffer(eval_key, np.uint8),
             bsk=bsk, boot_key=np.frombuffer(boot_key, np.uint8), probe=np.frombuffer(probe, np.uint8), probe_crc=np.uint32(r.crc32(probe)), empty_crc=np.uint32(r.crc32(b"")))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
