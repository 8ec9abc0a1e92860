"""CPU-only tests that pin the oracle (oracle/fhe_oracle.c).

1. against the golden vectors in tests/golden/ (outputs of the reference's own code,
   made by tests/golden/make_golden.py) - runs everywhere;
2. against the reference's own compiled classes (oracle/_ref/libref_oracle.so) on fresh
   seeded inputs - runs where that library exists;
3. the reference's own stand-alone test programs (cpp/tests/test_*.cpp built by
   oracle/build_ref.sh) must pass - runs where they exist.

Bit-exact (integer) comparisons throughout.
"""
import glob
import os
import subprocess

import numpy as np
import pytest

from oracle_bindings import ORACLE_DIR, mt19937_64_coeffs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Q27 = 132120577
Q62 = 4611686018326724609
QT = 1099511678977


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b), f"first mismatch at {np.argwhere(a != b)[:3].tolist()}"


# ------------------------------------------------------------------ golden --
def test_survey_anchor_values(oracle):
    """values obtained from the reference's NTTProcessor(8, 97)."""
    fwd, inv, psi, psi_inv, inv_n = oracle.twiddles(8, 97)
    assert (psi, psi_inv, inv_n) == (8, 85, 85)
    x = np.arange(1, 9, dtype=np.uint64)
    eq(oracle.forward(x[None], 97, fwd)[0], [36, 85, 31, 72, 93, 74, 58, 44])
    b = np.array([3, 1, 4, 1, 5, 9, 2, 6], np.uint64)
    eq(oracle.multiply(x[None], b[None], 97, fwd, inv, inv_n)[0], [63, 91, 86, 72, 83, 78, 41, 20])
    assert oracle.twiddles(1024, Q27)[2] == 113022246
    assert oracle.twiddles(4096, Q62)[2] == 2147730007686759558


def test_mt19937_64_matches_libstdcxx():
    # first outputs of std::mt19937_64(5489) are 14514284786278117030, 4620546740167642908
    assert mt19937_64_coeffs(5489, 2, 2**64 - 1).tolist() == [14514284786278117030 % (2**64 - 1),
                                                               4620546740167642908]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ntt_*.npz"))), ids=os.path.basename)
def test_golden_transforms(oracle, path):
    g = np.load(path)
    n, q = int(g["n"]), int(g["q"])
    fwd, inv, psi, psi_inv, inv_n = oracle.twiddles(n, q)
    assert (psi, psi_inv, inv_n) == (int(g["psi"]), int(g["psi_inv"]), int(g["inv_n"]))
    eq(fwd[:8], g["fwd_head"])
    eq(inv[:8], g["inv_head"])
    x = g["x"]
    eq(mt19937_64_coeffs(int(g["seed"]), x.size, q).reshape(x.shape), x)
    eq(oracle.forward(x, q, fwd), g["forward"])
    eq(oracle.inverse(x, q, inv, inv_n), g["inverse"])
    eq(oracle.inverse(g["forward"], q, inv, inv_n), x)  # round trip (test_ntt_processor.cpp:193-268)
    h = x.shape[0] // 2
    eq(oracle.multiply(x[:h], x[h : 2 * h], q, fwd, inv, inv_n), g["product"])


def test_golden_multi_limb(oracle):
    g = np.load(os.path.join(GOLDEN, "mlimb_q65.npz"))
    q = g["q"]
    q_inv, r1, r2 = oracle.mlimb_constants(q)
    assert q_inv == int(g["q_inv"])
    eq(r1, g["r_mod_q"])
    eq(r2, g["r2_mod_q"])
    a, b = g["a"], g["b"]
    eq(oracle.mlimb_montmul(a, b, q, q_inv), g["montmul"])
    eq(oracle.mlimb_add(a, b, q), g["add"])
    eq(oracle.mlimb_sub(a, b, q), g["sub"])
    r2b = np.broadcast_to(r2, a.shape).copy()
    eq(oracle.mlimb_montmul(a, r2b, q, q_inv), g["to_mont"])          # modular_arithmetic.cpp:673-678
    one = np.zeros_like(a)
    one[:, 0] = 1
    eq(oracle.mlimb_montmul(a, one, q, q_inv), g["from_mont"])        # :680-693 == reduce(a * 1)
    # independent check with Python integers
    qi = int(q[0]) + (int(q[1]) << 64)
    R = 1 << 128
    ai = [int(x[0]) + (int(x[1]) << 64) for x in a]
    bi = [int(x[0]) + (int(x[1]) << 64) for x in b]
    exp = [(x * y * pow(R, -1, qi)) % qi for x, y in zip(ai, bi)]
    got = [int(x[0]) + (int(x[1]) << 64) for x in g["montmul"]]
    assert exp == got


@pytest.mark.parametrize("tag", ["n128_l3", "tfhe_shape_small_n", "tfhe256_shape_small_n"])
def test_golden_bootstrap(oracle, tag):
    g = np.load(os.path.join(GOLDEN, f"boot_{tag}.npz"))
    N, n, k, base_log, level, q, t = (int(g[x]) for x in ("N", "n", "k", "base_log", "level", "q", "t"))
    fwd, inv, _, _, inv_n = oracle.twiddles(N, q)
    p = oracle.boot_params(N, q, n, k, base_log, level, t, fwd, inv, inv_n)
    eq(oracle.default_test_poly(p), g["test_poly"])
    eq(oracle.lookup_table(p, 0, 4), g["lut_identity"])
    eq(oracle.lookup_table(p, 1, 4), g["lut_negation"])
    eq(oracle.lookup_table(p, 2, 2, 4), g["lut_threshold"])
    glwe, bsk = g["glwe"], g["bsk"]
    eq(oracle.rotate(glwe[0], q, 5), g["rotate_5"])
    eq(oracle.rotate(glwe[0], q, -3), g["rotate_m3"])
    eq(oracle.rotate(glwe[0], q, N + 7), g["rotate_big"])
    eq(oracle.decompose(glwe[0], q, base_log, level), g["decompose"])
    eq(oracle.external_product(p, glwe, bsk[1]), g["external_product"])
    eq(oracle.cmux(p, bsk[2], glwe, g["blind_rotate"][0]), g["cmux"])
    acc = oracle.blind_rotate(p, g["lwe"], bsk, g["test_poly"])
    eq(acc, g["blind_rotate"])
    ext = oracle.sample_extract(acc, k, N, q)
    eq(ext, g["sample_extract"])
    if "ksk" in g.files:
        eq(oracle.key_switch(ext, q, g["ksk"], n, base_log, level), g["key_switch"])
        eq(oracle.bootstrap(p, g["lwe"], bsk, g["test_poly"], g["ksk"], n, base_log, level), g["bootstrap"])


@pytest.mark.parametrize("name", ["tally_n64_m5.npz", "tally_n1024_m9.npz"])
def test_golden_tally(oracle, name):
    g = np.load(os.path.join(GOLDEN, name))
    n, q, cts = int(g["n"]), int(g["q"]), g["cts"]
    eq(oracle.tally(cts, q), g["linear"])
    eq(oracle.tally(cts, q, tree=True), g["tree"])
    eq(g["linear"], g["tree"])                 # any reduction shape gives the same words
    eq(oracle.tally(cts[1:2], q), g["single"])  # one ballot: returned untouched, even unreduced words
    eq(g["single"], cts[1])
    with pytest.raises(ValueError):
        oracle.tally(cts[:0], q)
    fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
    eq(oracle.tensor_multiply(cts[0], cts[2], q, fwd, inv, inv_n), g["tensor"])


RELIN_GOLDEN = ["relin_n8.npz", "relin_n64.npz", "relin_n1024.npz", "relin_n4096.npz"]


@pytest.mark.parametrize("name", RELIN_GOLDEN)
def test_golden_relinearize(oracle, name):
    g = np.load(os.path.join(GOLDEN, name))
    n, q, bl, lv = int(g["n"]), int(g["q"]), int(g["base_log"]), int(g["level"])
    fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
    for i in range(2):
        eq(oracle.relinearize(g["ct"][i], g["keys"], bl, lv, q, fwd, inv, inv_n), g["out"][i])
    eq(oracle.relinearize(g["ct"][1], g["keys"][:0], bl, lv, q, fwd, inv, inv_n), g["nokey"])
    eq(g["nokey"], g["ct"][1, :2])  # no key pairs: c0 and c1 come back untouched, unreduced words included


# ----------------------------------------------------- reference, run live --
@pytest.mark.parametrize("n,q", [(4, 17), (8, 17), (16, 97), (64, QT), (512, Q27), (1024, Q27), (2048, 1125899906826241),
                                 (4096, Q62), (8192, Q62)])
def test_transforms_match_reference(oracle, ref, n, q):
    h = ref.ntt_create(n, q)
    try:
        rt = ref.ntt_tables(h, n)
        ot = oracle.twiddles(n, q)
        for a, b in zip(rt, ot):
            eq(a, b)
        fwd, inv, _, _, inv_n = ot
        rng = np.random.default_rng(n)
        cases = [mt19937_64_coeffs(123, 3 * n, q).reshape(3, n),
                 rng.integers(0, 2**64, size=(2, n), dtype=np.uint64),         # unreduced inputs
                 np.zeros((1, n), np.uint64), np.full((1, n), q - 1, np.uint64)]
        for x in cases:
            eq(oracle.forward(x, q, fwd), ref.ntt_forward(h, x))
            eq(oracle.inverse(x, q, inv, inv_n), ref.ntt_inverse(h, x))
        # adaptive_dispatcher's caller-supplied-table entry points agree on canonical inputs
        x = cases[0]
        if q % 3 and q % 5 and q % 17 and q % 257:  # its Montgomery setup is fine for any odd q; keep to parity configs
            eq(oracle.forward(x, q, fwd), ref.fast_ntt_forward(x, q, fwd))
            eq(oracle.fast_inverse(x, q, inv), ref.fast_ntt_inverse(x, q, inv))
    finally:
        ref.ntt_destroy(h)


def test_not_ntt_friendly_rejected(oracle, ref):
    with pytest.raises(ValueError):
        oracle.twiddles(1024, 1099511627777 - 2)  # q-1 not divisible by 2N
    with pytest.raises(Exception):
        ref.ntt_create(1024, 1099511627777 - 2)


@pytest.mark.parametrize("n,q", [(8, 97), (1024, Q27), (4096, Q62)])
def test_ring_ops_match_reference(oracle, ref, n, q):
    ring = ref.ring_create(n, q)
    try:
        fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
        rng = np.random.default_rng(7)
        a = rng.integers(0, q, size=(3, n), dtype=np.uint64)
        b = rng.integers(0, q, size=(3, n), dtype=np.uint64)
        au = rng.integers(0, 2**64, size=(3, n), dtype=np.uint64)
        bu = rng.integers(0, 2**64, size=(3, n), dtype=np.uint64)
        for x, y in [(a, b), (au, bu)]:
            eq(oracle.add(x, y, q), ref.ring_op(ring, "add", x, y))
            eq(oracle.sub(x, y, q), ref.ring_op(ring, "sub", x, y))
            eq(oracle.pointwise(x, y, q), ref.ring_op(ring, "pointwise", x, y))
            eq(oracle.negate(x, q), ref.ring_op(ring, "negate", x))
            eq(oracle.scalar(x, int(y[0, 0]), q), ref.ring_op(ring, "scalar", x, scalar=int(y[0, 0])))
            eq(oracle.multiply(x, y, q, fwd, inv, inv_n), ref.ring_op(ring, "multiply", x, y))
        eq(oracle.pointwise(a, b, q), ref.fast_modmul(a, b, q))
        cts = rng.integers(0, q, size=(11, 2, n), dtype=np.uint64)
        eq(oracle.tally(cts, q), ref.tally(ring, cts))
        eq(oracle.tally(cts, q, tree=True), ref.tally(ring, cts, tree=True))
        eq(oracle.tensor_multiply(cts[0], cts[1], q, fwd, inv, inv_n), ref.tensor_multiply(ring, cts[0], cts[1]))
        # relinearisation of a tensor product, default and explicit gadgets, fewer keys than levels
        ct3 = ref.tensor_multiply(ring, cts[0], cts[1])
        for key_count, bl, lv in [(16, 0, 0), (3, 0, 0), (2, 20, 2), (4, 7, 6)]:
            keys = rng.integers(0, q, size=(key_count, 2, n), dtype=np.uint64)
            eq(oracle.relinearize(ct3, keys, bl, lv, q, fwd, inv, inv_n), ref.relinearize(ring, ct3, keys, bl, lv))
    finally:
        ref.ring_destroy(ring)


@pytest.mark.parametrize("q_limbs", [[0xFFFFFFFFFFFFFF43, 1], [0xFFFFFFFFFFFFFFC5], [0x1D, 0, 1],
                                     [0xFFFFFFFFFFFFFF61, 0xFFFFFFFFFFFFFFFF]])
def test_multi_limb_matches_reference(oracle, ref, q_limbs):
    q = np.array(q_limbs, np.uint64)
    l = q.size
    h = ref.mlimb_create(q)
    try:
        rq_inv, rr1, rr2 = ref.mlimb_constants(h, l)
        q_inv, r1, r2 = oracle.mlimb_constants(q)
        assert q_inv == rq_inv
        eq(r1, rr1)
        eq(r2, rr2)
        qi = sum(int(v) << (64 * i) for i, v in enumerate(q))
        rng = np.random.default_rng(l)
        vals = [int.from_bytes(rng.bytes(8 * l + 1), "little") % qi for _ in range(400)]
        vals[:4] = [0, 1, qi - 1, qi // 2]
        arr = np.array([[(v >> (64 * i)) & (2**64 - 1) for i in range(l)] for v in vals], np.uint64)
        a, b = arr[:200], arr[200:]
        eq(oracle.mlimb_montmul(a, b, q, q_inv), ref.mlimb_op(h, "montmul", a, b))
        eq(oracle.mlimb_add(a, b, q), ref.mlimb_op(h, "add", a, b))
        eq(oracle.mlimb_sub(a, b, q), ref.mlimb_op(h, "sub", a, b))
    finally:
        ref.mlimb_destroy(h)


@pytest.mark.parametrize("N,n,k,base_log,level", [(64, 5, 1, 4, 3), (256, 4, 2, 7, 2), (1024, 3, 1, 23, 1)])
def test_bootstrap_matches_reference(oracle, ref, N, n, k, base_log, level):
    """Synthetic (uniform random) keys: parity is on raw ciphertext words."""
    q, t = QT, 4
    h = ref.boot_create(N, q, n, k, base_log, level, t)
    try:
        fwd, inv, _, _, inv_n = oracle.twiddles(N, q)
        p = oracle.boot_params(N, q, n, k, base_log, level, t, fwd, inv, inv_n)
        rng = np.random.default_rng(N + n)
        bsk = rng.integers(0, q, size=ref.bsk_shape(h), dtype=np.uint64)
        ref.boot_import_bsk(h, bsk)
        eq(ref.boot_export_bsk(h), bsk)
        glwe = rng.integers(0, q, size=(k + 1, N), dtype=np.uint64)
        glwe2 = rng.integers(0, q, size=(k + 1, N), dtype=np.uint64)
        for rot in (0, 1, -1, N - 1, N, 2 * N, 3 * N + 5, -(2 * N + 3)):
            eq(oracle.rotate(glwe[0], q, rot), ref.boot_rotate(h, glwe[0], rot))
        eq(oracle.decompose(glwe[0], q, base_log, level), ref.boot_decompose(h, glwe[0], base_log, level))
        eq(oracle.external_product(p, glwe, bsk[0]), ref.boot_external_product(h, glwe, 0))
        eq(oracle.cmux(p, bsk[1], glwe, glwe2), ref.boot_cmux(h, 1, glwe, glwe2))
        lwe = rng.integers(0, q, size=(3, n + 1), dtype=np.uint64)
        lwe[0, 0] = 0
        lwe[1, n] = 0
        tp = ref.boot_default_test_poly(h)
        eq(oracle.default_test_poly(p), tp)
        acc = ref.boot_blind_rotate(h, lwe, tp, threads=3)
        eq(oracle.blind_rotate(p, lwe, bsk, tp), acc)
        ext = ref.boot_sample_extract(h, acc)
        eq(oracle.sample_extract(acc, k, N, q), ext)
        n_out = 8
        ksk = rng.integers(0, q, size=(k * N * level, n_out + 1), dtype=np.uint64)
        ref.boot_import_ksk(h, ksk, base_log, level)
        eq(oracle.key_switch(ext, q, ksk, n_out, base_log, level), ref.boot_key_switch(h, ext, n_out))
        eq(oracle.bootstrap(p, lwe, bsk, tp, ksk, n_out, base_log, level), ref.boot_bootstrap(h, lwe, tp, n_out))
    finally:
        ref.boot_destroy(h)


@pytest.mark.parametrize("prog", ["test_ntt_processor", "test_multi_limb", "test_polynomial_ring",
                                  "test_neon_correctness"])
def test_reference_own_tests_pass(prog):
    """The reference's stand-alone test mains, built from its sources by oracle/build_ref.sh."""
    exe = os.path.join(ORACLE_DIR, "_ref", prog)
    if not os.path.exists(exe):
        pytest.skip("reference test program not built (needs /root/reference)")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
