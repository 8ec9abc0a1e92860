"""GPU parity tests of the TFHE bootstrap chain (run on the B200 with `-m gpu`), through the C ABI:
external product, CMux, blind rotation, sample extraction, key switching and the whole
bootstrap, bit-exact against

* tests/golden/boot_*.npz - outputs of the reference's own BootstrapEngine on keys made by the
  reference's own KeyManager / encrypt_ggsw / generate_key_switch_key, and
* the CPU oracle (oracle/fhe_oracle.c, pinned to the reference) on fresh seeded keys and inputs.

Parity is on raw ciphertext words (the reference's bootstrap output does not decrypt to its
input - SURVEY H10 - so decrypted values are never compared).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Q27 = 132120577
Q62 = 4611686018326724609
QT = 1099511678977


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b), f"{int((a != b).sum())} words differ; first at {np.argwhere(a != b)[:3].tolist()}"


@pytest.fixture(scope="module")
def fhe():
    import fheb200

    fheb200.initialize()
    return fheb200


@pytest.fixture(scope="module")
def torch():
    import torch

    assert torch.cuda.is_available()
    return torch


def dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("tag", ["n128_l3", "tfhe_shape_small_n", "tfhe256_shape_small_n"])
def test_golden_bootstrap(fhe, torch, tag):
    g = np.load(os.path.join(GOLDEN, f"boot_{tag}.npz"))
    N, n, k, base_log, level, q, t = (int(g[x]) for x in ("N", "n", "k", "base_log", "level", "q", "t"))
    eng = fhe.BootstrapEngine(N, q, n, k, base_log, level, g["bsk"], plaintext_modulus=t)
    eq(eng.get_default_test_poly(), g["test_poly"])
    eq(eng.create_identity_lut(4), g["lut_identity"])
    eq(eng.create_negation_lut(4), g["lut_negation"])
    eq(eng.create_threshold_lut(2, 4), g["lut_threshold"])
    glwe = g["glwe"]
    eq(eng.external_product(glwe[None], 1)[0], g["external_product"])                  # host buffers
    eq(host(eng.external_product(dev(torch, glwe[None]), 1))[0], g["external_product"])  # device buffers
    eq(eng.cmux(2, glwe[None], g["blind_rotate"][0:1])[0], g["cmux"])
    acc = eng.blind_rotate(g["lwe"], g["test_poly"])
    eq(acc, g["blind_rotate"])
    eq(host(eng.blind_rotate(dev(torch, g["lwe"]), dev(torch, g["test_poly"]))), g["blind_rotate"])
    eq(eng.sample_extract(acc), g["sample_extract"])
    eq(eng.bootstrap(g["lwe"], g["test_poly"]), g["sample_extract"])  # no KSK: chain stops after extraction
    if "ksk" in g.files:
        eng.set_key_switch_key(g["ksk"], n, base_log, level)
        eq(eng.key_switch(g["sample_extract"]), g["key_switch"])
        eq(eng.bootstrap(g["lwe"], g["test_poly"]), g["bootstrap"])
        eq(host(eng.bootstrap(dev(torch, g["lwe"]), dev(torch, g["test_poly"]))), g["bootstrap"])


SHAPES = [  # N, q, n, k, base_log, level, batch
    (32, Q62, 5, 1, 8, 2, 3),
    (64, Q62, 4, 2, 7, 2, 5),
    (128, QT, 8, 1, 4, 3, 9),
    (256, Q62, 4, 1, 15, 4, 4),
    (512, QT, 5, 3, 6, 2, 3),
    (512, Q62, 4, 2, 20, 1, 7),
    (1024, QT, 12, 1, 23, 1, 300),   # tfhe-128-fast shape (short key), more ciphertexts than resident blocks
    (2048, Q27, 3, 1, 9, 2, 3),
    (4096, Q62, 2, 1, 30, 1, 2),
    (2048, 1125899906826241, 3, 1, 15, 2, 3),      # tfhe-128-balanced shape (Q_50_1), short key
    (4096, 1152921504606584833, 2, 1, 10, 3, 3),   # tfhe-256-secure shape (Q_60_1): accumulator kept in global memory
    (4096, Q27, 2, 1, 9, 3, 2),                    # the same working set in FP64 mode
]


@pytest.mark.parametrize("N,q,n,k,base_log,level,batch", SHAPES)
def test_bootstrap_chain_matches_oracle(fhe, torch, oracle, N, q, n, k, base_log, level, batch):
    rng = np.random.default_rng(N + n + k)
    fwd, inv, _, _, inv_n = oracle.twiddles(N, q)
    p = oracle.boot_params(N, q, n, k, base_log, level, 4, fwd, inv, inv_n)
    rows = (k + 1) * level
    bsk = rng.integers(0, q, size=(n, rows, k + 1, N), dtype=np.uint64)
    eng = fhe.BootstrapEngine(N, q, n, k, base_log, level, bsk)
    test_poly = oracle.default_test_poly(p)
    eq(eng.get_default_test_poly(), test_poly)

    nb = min(batch, 4)  # the oracle is slow: check a few ciphertexts of the batch in full
    glwe = rng.integers(0, q, size=(nb, k + 1, N), dtype=np.uint64)
    other = rng.integers(0, 2**64, size=(nb, k + 1, N), dtype=np.uint64)  # unreduced words are reduced by subtract
    for i in range(nb):
        eq(host(eng.external_product(dev(torch, glwe[i:i + 1]), i % n))[0], oracle.external_product(p, glwe[i], bsk[i % n]))
        eq(host(eng.cmux(i % n, dev(torch, glwe[i:i + 1]), dev(torch, other[i:i + 1])))[0],
           oracle.cmux(p, bsk[i % n], glwe[i], other[i]))

    lwe = rng.integers(0, q, size=(batch, n + 1), dtype=np.uint64)
    lwe[0, :] = 0           # every rotation is zero: the accumulator passes through untouched
    lwe[1 % batch, 0] = 0   # a skipped step in the middle of the chain
    acc = host(eng.blind_rotate(dev(torch, lwe), dev(torch, test_poly)))
    sel = sorted(set([0, 1 % batch, batch // 2, batch - 1]))
    exp = oracle.blind_rotate(p, lwe[sel], bsk, test_poly)
    eq(acc[sel], exp)
    eq(host(eng.sample_extract(dev(torch, acc[sel]))), oracle.sample_extract(exp, k, N, q))
    if batch > 8:  # every ciphertext of a large batch: identical inputs must give identical outputs
        lwe2 = np.repeat(lwe[sel[-1]][None], batch, axis=0)
        acc2 = host(eng.blind_rotate(dev(torch, lwe2), dev(torch, test_poly)))
        eq(acc2, np.repeat(exp[-1][None], batch, axis=0))

    # key switching + full chain with a random key switching key
    ks_level, ks_base = 2, 5
    n_out = n
    ksk = rng.integers(0, q, size=(k * N * ks_level, n_out + 1), dtype=np.uint64)
    eng.set_key_switch_key(ksk, n_out, ks_base, ks_level)
    ext = oracle.sample_extract(exp, k, N, q)
    eq(host(eng.key_switch(dev(torch, ext))), oracle.key_switch(ext, q, ksk, n_out, ks_base, ks_level))
    eq(host(eng.bootstrap(dev(torch, lwe[sel]), dev(torch, test_poly))),
       oracle.bootstrap(p, lwe[sel], bsk, test_poly, ksk, n_out, ks_base, ks_level))


def test_key_switch_ragged_shapes(fhe, torch, oracle):
    """Output widths and batches that do not divide the kernel's tiles; zero digits are skipped."""
    N, q, n, k = 64, QT, 3, 1
    rng = np.random.default_rng(5)
    bsk = rng.integers(0, q, size=(n, 2, 2, N), dtype=np.uint64)
    eng = fhe.BootstrapEngine(N, q, n, k, 10, 1, bsk)
    # base_log above 32 takes the 64-bit digit variant of the kernel; 62-bit residues make the wrapped products matter
    for n_out, level, base_log, batch in [(1, 1, 3, 1), (127, 3, 4, 9), (128, 2, 23, 8), (300, 1, 7, 17), (65, 1, 40, 5), (33, 1, 63, 3)]:
        ksk = rng.integers(0, q, size=(k * N * level, n_out + 1), dtype=np.uint64)
        eng.set_key_switch_key(ksk, n_out, base_log, level)
        lwe = rng.integers(0, q, size=(batch, k * N + 1), dtype=np.uint64)
        lwe[0, :k * N] = 0  # all digits zero: result is (0, .., 0, b)
        eq(host(eng.key_switch(dev(torch, lwe))), oracle.key_switch(lwe, q, ksk, n_out, base_log, level))
        # unreduced b (bootstrap_engine.cpp:640,665): kept raw when every digit is zero, and `(b + q - t) % q` of the
        # first non-zero digit wraps modulo 2^64 - the kernel mirrors both
        raw = lwe.copy()
        raw[:, k * N] = rng.integers(2**64 - q, 2**64, size=batch, dtype=np.uint64)
        raw[batch // 2, k * N] = np.uint64(q)
        raw[-1, k * N] = np.uint64(2**64 - 1)
        raw[-1, :k * N - 1] = 0  # only the last coefficient has digits
        eq(host(eng.key_switch(dev(torch, raw))), oracle.key_switch(raw, q, ksk, n_out, base_log, level))


def test_bootstrap_errors(fhe):
    bsk = np.zeros((2, 2, 2, 64), np.uint64)
    eng = fhe.BootstrapEngine(64, QT, 2, 1, 10, 1, bsk)
    with pytest.raises(fhe.FheError):
        eng.external_product(np.zeros((1, 2, 64), np.uint64), 2)  # index out of range
    with pytest.raises(fhe.FheError):
        eng.key_switch(np.zeros((1, 65), np.uint64))  # no key switching key set
    with pytest.raises(fhe.FheError):
        fhe.BootstrapEngine(64, QT, 2, 1, 40, 2, np.zeros((2, 4, 2, 64), np.uint64))  # level * base_log > 64
    with pytest.raises(fhe.FheError):
        fhe.BootstrapEngine(16, 97, 2, 1, 3, 1, np.zeros((2, 2, 2, 16), np.uint64))  # N < 32
    assert eng.blind_rotate(np.zeros((0, 3), np.uint64), np.zeros(64, np.uint64)).shape == (0, 2, 64)


def test_full_batch_tfhe_shape_properties(fhe, torch, oracle):
    """BASELINE C4 shape (N=1024, k=1, n=742, base_log=23, L=1, batch 4096) with a synthetic key:
    eight ciphertexts against the oracle (about a second of CPU each) and batch-order independence on all 4096."""
    N, q, n, k, base_log, level, batch = 1024, QT, 742, 1, 23, 1, 4096
    rng = np.random.default_rng(742)
    bsk = rng.integers(0, q, size=(n, 2, 2, N), dtype=np.uint64)
    eng = fhe.BootstrapEngine(N, q, n, k, base_log, level, bsk)
    fwd, inv, _, _, inv_n = oracle.twiddles(N, q)
    p = oracle.boot_params(N, q, n, k, base_log, level, 4, fwd, inv, inv_n)
    test_poly = eng.get_default_test_poly()
    lwe = rng.integers(0, q, size=(batch, n + 1), dtype=np.uint64)
    lwe_d = dev(torch, lwe)
    acc = eng.blind_rotate(lwe_d, dev(torch, test_poly))
    sel = [0, 1, 591, 592, 2047, 3333, 4094, 4095]  # first / last of the batch, around a wave boundary (592 resident blocks), interior
    eq(host(acc[sel]), oracle.blind_rotate(p, lwe[sel], bsk, test_poly))
    perm = torch.randperm(batch, device="cuda")
    acc_p = eng.blind_rotate(lwe_d[perm].contiguous(), dev(torch, test_poly))
    assert torch.equal(acc_p, acc[perm])


@pytest.mark.parametrize("N,q,n,base_log,level", [(1024, QT, 6, 23, 1), (128, Q62, 5, 9, 2)])
def test_blind_rotation_with_unreduced_test_polynomial(fhe, torch, oracle, N, q, n, base_log, level):
    """A caller's test polynomial may hold words >= q; the reference's rotate / subtract / add treat them with their
    unsigned wrap-around formulas until the first executed step.  A device flag routes such calls to the general
    kernel (the lean one has that handling compiled out); both must reproduce the oracle."""
    rng = np.random.default_rng(N)
    fwd, inv, _, _, inv_n = oracle.twiddles(N, q)
    p = oracle.boot_params(N, q, n, 1, base_log, level, 4, fwd, inv, inv_n)
    bsk = rng.integers(0, q, size=(n, 2 * level, 2, N), dtype=np.uint64)
    eng = fhe.BootstrapEngine(N, q, n, 1, base_log, level, bsk)
    lwe = rng.integers(0, q, size=(8, n + 1), dtype=np.uint64)
    lwe[0, :] = 0     # no step executes: the raw words pass through the initial rotation only
    lwe[1, :2] = 0    # the first executed step comes late
    # a_i >= q - q/4N rounds to the RAW rotation 2N: the reference tests the int32 before rotate_polynomial normalises
    # it (bootstrap_engine.cpp:564-566), so it still runs a CMux with diff = 0, which canonicalises an unreduced
    # accumulator.  (q-1, q/3, 0, ..; b = 0) and (0, q/3, 0, ..) differ under the reference once the test polynomial
    # holds unreduced words; so do "only step is a 2N one" and "2N step after a real one" (the identity by then).
    lwe[5, :] = 0
    lwe[5, 0], lwe[5, 1] = q - 1, q // 3
    lwe[6, :] = 0
    lwe[6, 1] = q // 3
    lwe[7, :] = 0
    lwe[7, n - 1] = q - 1
    lwe[4, 0], lwe[4, 1] = q // 5, q - 2
    clean = oracle.default_test_poly(p)
    raw = clean.copy()
    raw[::7] += np.uint64(q)                          # same residues, unreduced
    raw[3], raw[N - 1] = np.uint64(2**64 - 1), np.uint64(q)
    for tp in (raw, clean, raw):                      # alternate: the flag is recomputed on every call
        eq(host(eng.blind_rotate(dev(torch, lwe), dev(torch, tp))), oracle.blind_rotate(p, lwe, bsk, tp))
    eq(eng.blind_rotate(lwe, raw), oracle.blind_rotate(p, lwe, bsk, raw))  # host buffers


def test_c4_full_shape_with_the_reference_own_keys(fhe, torch):
    """BASELINE C4 at its full shape (N=1024, k=1, n=742, base_log=23, L=1; substitute prime - SURVEY H6) with keys and
    inputs made by the reference ITSELF, live: KeyManager-style key generation + BootstrapEngine::encrypt_ggsw x 742
    (cpp/src/bootstrap_engine.cpp:268-364,391-421) and encrypt_lwe, through oracle/_ref (the reference's own sources
    compiled where they lie; the built library travels to the GPU box).  The GPU's blind-rotation and sample-extraction
    words must equal the reference's BootstrapEngine::blind_rotate / sample_extract (:547-577, :594-624) on 16
    ciphertexts.  A 24 MB key is too large for a committed fixture, hence live generation (about 15 s of host time)."""
    from oracle_bindings import RefOracle, ref_available

    if not ref_available():
        pytest.skip("oracle/_ref/libref_oracle.so not built (needs the reference sources at build time)")
    N, q, n, k, base_log, level, t = 1024, QT, 742, 1, 23, 1, 4
    r = RefOracle()
    h = r.boot_create(N, q, n, k, base_log, level, t)
    try:
        sk = r.boot_keygen(h)                      # the reference's own secret key and 742 GGSW ciphertexts
        assert set(np.unique(sk).tolist()) <= {0, 1}
        bsk = r.boot_export_bsk(h)
        assert bsk.shape == (n, (k + 1) * level, k + 1, N)
        msgs = np.arange(16, dtype=np.uint64) % np.uint64(t)
        lwe = r.boot_encrypt_lwe(h, msgs * np.uint64(q // t), sk)
        tp = r.boot_default_test_poly(h)
        cores = min(os.cpu_count() or 1, 16)
        exp_acc = r.boot_blind_rotate(h, lwe, tp, threads=cores)
        exp_ext = r.boot_sample_extract(h, exp_acc)
    finally:
        r.boot_destroy(h)
    eng = fhe.BootstrapEngine(N, q, n, k, base_log, level, bsk, plaintext_modulus=t)
    eq(eng.get_default_test_poly(), tp)
    acc = host(eng.blind_rotate(dev(torch, lwe), dev(torch, tp)))
    eq(acc, exp_acc)
    eq(host(eng.sample_extract(dev(torch, acc))), exp_ext)
    eq(eng.bootstrap(lwe, tp), exp_ext)            # host buffers, whole chain (no key switching key: stops after extraction)
