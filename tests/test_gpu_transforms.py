"""GPU parity tests (run on the B200 with `-m gpu`): the CUDA path, called through the C ABI
(libfheb200.so via the host mirror in node-fhe-accelerate_b200/api.py), against

* the golden vectors under tests/golden/ (outputs of the reference's own code), and
* the CPU oracle (oracle/fhe_oracle.c, pinned to the reference) on fresh seeded inputs.

Everything is integer work: comparisons are bit-exact.  Modelled on the reference's
backend-equivalence tests (cpp/tests/test_metal_compute.cpp:109-181) and on
cpp/tests/test_ntt_processor.cpp / test_polynomial_ring.cpp / test_multi_limb.cpp.
"""
import glob
import os

import numpy as np
import pytest

from oracle_bindings import mt19937_64_coeffs

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Q27 = 132120577
Q62 = 4611686018326724609
QT = 1099511678977
Q50 = 1125899906826241


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
    """numpy uint64 -> CUDA int64 tensor holding the same words."""
    return torch.from_numpy(np.ascontiguousarray(x).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


# ------------------------------------------------------------------------------ library --
def test_device_is_b200_and_library_loaded(fhe):
    info = fhe.detect_hardware()
    assert info["has_cuda"] and info["compute_capability"][0] == 10
    assert not (info["has_metal"] or info["has_sme"] or info["has_neon"] or info["has_amx"])
    assert info["sm_count"] > 0
    with open("/proc/self/maps") as f:
        assert "libfheb200.so" in f.read()


# --------------------------------------------------------------------------- transforms --
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ntt_*.npz"))), ids=os.path.basename)
def test_golden_transforms(fhe, torch, path):
    g = np.load(path)
    n, q = int(g["n"]), int(g["q"])
    ntt = fhe.NTTProcessor(n, q)
    fwd, inv, psi, psi_inv, inv_n = ntt.get_twiddles()
    assert (psi, psi_inv, inv_n) == (int(g["psi"]), int(g["psi_inv"]), int(g["inv_n"]))
    eq(fwd[:8], g["fwd_head"])
    eq(inv[:8], g["inv_head"])
    x = g["x"]
    # host buffers (staged inside the call)
    eq(ntt.forward_ntt(x), g["forward"])
    eq(ntt.inverse_ntt(x), g["inverse"])
    eq(ntt.inverse_ntt(g["forward"]), x)
    # device buffers, out of place and in place
    xd = dev(torch, x)
    eq(host(ntt.forward_ntt(xd)), g["forward"])
    yd = xd.clone()
    ntt.forward_ntt(yd, out=yd)
    eq(host(yd), g["forward"])
    ntt.inverse_ntt(yd, out=yd)
    eq(host(yd), x)
    ring = fhe.PolynomialRing(n, q)
    h = x.shape[0] // 2
    eq(ring.multiply(x[:h], x[h:2 * h]), g["product"])
    eq(host(ring.multiply(dev(torch, x[:h]), dev(torch, x[h:2 * h]))), g["product"])


def _prime_for(logn, mode):
    """q62: 64-bit integer kernels; q<2^42: FP64-pipe kernels (41-bit prime up to N = 1024, then the 27-bit prime of the
    published rows with the 32-bit kernels switched off); q<2^27: the 32-bit kernels (132120577 and, at the top of the
    32 q window, the largest prime below 2^27 with 2^15 | q - 1)."""
    n = 1 << logn
    if mode == "q62":
        return Q62
    if mode == "q<2^42":
        return QT if n <= 1024 else Q27
    return Q27 if logn % 2 == 0 else 133857281


@pytest.mark.parametrize("lazy", ["q62", "q<2^42", "q<2^27", "q<2^27-paired"])
@pytest.mark.parametrize("logn", range(2, 15))
def test_transforms_match_oracle_all_degrees(fhe, torch, oracle, logn, lazy, monkeypatch):
    n, q = 1 << logn, _prime_for(logn, lazy)
    monkeypatch.delenv("FHEB_NO_U32", raising=False)
    monkeypatch.setenv("FHEB_U32_PAIR", "0")   # 32-bit kernels, one polynomial per 4-byte shared-memory slot
    if lazy == "q<2^42":
        monkeypatch.setenv("FHEB_NO_U32", "1")  # read per call by the library: keeps 27-bit primes on the FP64-pipe kernels
    if lazy == "q<2^27-paired":
        monkeypatch.setenv("FHEB_U32_PAIR", "1")   # 32-bit kernels, two polynomials per 8-byte slot
    ntt = fhe.NTTProcessor(n, q)
    fwd, inv, psi, psi_inv, inv_n = oracle.twiddles(n, q)
    gf, gi, gpsi, gpsi_inv, ginv_n = ntt.get_twiddles()
    eq(gf, fwd)
    eq(gi, inv)
    assert (gpsi, gpsi_inv, ginv_n) == (psi, psi_inv, inv_n)
    rng = np.random.default_rng(1000 + logn)
    batch = 37 if logn <= 10 else 5  # ragged against every polys-per-block setting
    cases = [rng.integers(0, q, size=(batch, n), dtype=np.uint64),
             rng.integers(0, 2**64, size=(3, n), dtype=np.uint64),  # unreduced inputs are reduced first
             np.zeros((1, n), np.uint64), np.full((2, n), q - 1, np.uint64)]
    for x in cases:
        xd = dev(torch, x)
        eq(host(ntt.forward_ntt(xd)), oracle.forward(x, q, fwd))
        eq(host(ntt.inverse_ntt(xd)), oracle.inverse(x, q, inv, inv_n))
    x = cases[0]
    eq(host(ntt.inverse_ntt_forward_network(dev(torch, x))), oracle.fast_inverse(x, q, inv))
    ring = fhe.PolynomialRing(n, q)
    a, b = cases[0], np.roll(cases[0], 1, axis=0)
    eq(host(ring.multiply(dev(torch, a), dev(torch, b))), oracle.multiply(a, b, q, fwd, inv, inv_n))
    au, bu = cases[1], cases[1][::-1].copy()
    eq(host(ring.multiply(dev(torch, au), dev(torch, bu))), oracle.multiply(au, bu, q, fwd, inv, inv_n))
    eq(ring.multiply(a[:2], a[:2]), oracle.multiply(a[:2], a[:2], q, fwd, inv, inv_n))  # aliased operands


@pytest.mark.parametrize("four_pass", [False, True], ids=["three-passes", "four-passes"])
def test_pass_plans_of_n16384_match_oracle(fhe, torch, oracle, four_pass, monkeypatch):
    """N = 16384: the plain forward / inverse kernels run three passes of 5+4+5 / 4+5+5 stages from twiddle tables of their
    own (plan keys 80 / 79, ntt_core.cuh); FHEB_NO_ALT_PLAN=1 puts them back on the 4+4+3+3 split every other kernel uses.
    Both against the oracle, in the integer and the 32-bit mode, ragged batch (more polynomials than resident blocks)."""
    if four_pass:
        monkeypatch.setenv("FHEB_NO_ALT_PLAN", "1")
    else:
        monkeypatch.delenv("FHEB_NO_ALT_PLAN", raising=False)
    n = 16384
    for q in (Q62, Q27):
        ntt = fhe.NTTProcessor(n, q)
        fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
        rng = np.random.default_rng(n + q % 97)
        batch = 148 * 3 + 5
        x = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
        x[2, 5000:5009] = rng.integers(q, 2**64, size=9, dtype=np.uint64)  # unreduced words are reduced on load
        check = [0, 2, 147, 148, batch - 1]
        xd = dev(torch, x)
        f = ntt.forward_ntt(xd)
        eq(host(f[check]), oracle.forward(x[check], q, fwd))
        eq(host(ntt.inverse_ntt(xd)[check]), oracle.inverse(x[check], q, inv, inv_n))
        xr = x.copy()
        xr[2] %= np.uint64(q)
        eq(host(ntt.inverse_ntt(f)), xr)  # round trip on the whole batch


@pytest.mark.parametrize("tma", ["1", "0"], ids=["tma-landing-buffer", "plain-loads"])
@pytest.mark.parametrize("logn", range(5, 15))
def test_first_pass_input_paths_match_oracle(fhe, torch, oracle, logn, tma, monkeypatch):
    """The first pass reads its words either with streaming global loads or from a shared-memory landing buffer filled by
    bulk-async (TMA) copies of the block's next group (cp.async.bulk + mbarrier, ntt_device.cuh).  The library picks per
    degree / mode / direction from measurements; FHEB_TMA forces either path.  Both, in all three arithmetic modes, against
    the oracle - ragged batches (last group partly filled), more groups than resident blocks, in-place operation."""
    monkeypatch.setenv("FHEB_TMA", tma)
    n = 1 << logn
    for q in (Q62, QT if n <= 1024 else 1099511592961 if (1099511592961 - 1) % (2 * n) == 0 else None, Q27):
        if q is None or (q - 1) % (2 * n) != 0:
            continue
        ntt = fhe.NTTProcessor(n, q)
        fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
        rng = np.random.default_rng(logn * 7 + 1)
        per_block = max(1, 1024 >> logn) if logn <= 9 else (2 if logn == 10 else 1)
        batch = 148 * 8 * per_block // (1 << max(0, logn - 7)) + 3  # several groups per block, ragged tail
        batch = min(batch, 1500)
        if logn >= 13:
            batch = 148 * 2 + 3  # (N = 16384 with 8-byte slots lands 12 of the 16 first-pass rows, the rest comes from caller memory)
        x = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
        x[1, :7] = rng.integers(q, 2**64, size=7, dtype=np.uint64)  # unreduced words are reduced on load
        x[batch - 2, -5:] = rng.integers(q, 2**64, size=5, dtype=np.uint64)  # (N = 16384: in the rows that are NOT landed)
        check = sorted(set([0, 1, batch // 2, batch - 2, batch - 1]))
        xd = dev(torch, x)
        f = host(ntt.forward_ntt(xd))
        eq(f[check], oracle.forward(x[check], q, fwd))
        i = host(ntt.inverse_ntt(xd))
        eq(i[check], oracle.inverse(x[check], q, inv, inv_n))
        y = dev(torch, x)
        ntt.forward_ntt(y, out=y)   # in place
        eq(host(y), f)


def test_reference_test_ntt_processor_configs(fhe, oracle):
    """cpp/tests/test_ntt_processor.cpp:198-235,276-300: seeds 42/123, (8,17), (16,97), (1024,132120577)."""
    for n, q, iters in [(8, 17, 100), (16, 97, 100), (1024, Q27, 20)]:
        ntt = fhe.NTTProcessor(n, q)
        fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
        for seed in (42, 123):
            x = mt19937_64_coeffs(seed, n * iters, q).reshape(iters, n)
            y = ntt.forward_ntt(x)
            eq(y, oracle.forward(x, q, fwd))
            eq(ntt.inverse_ntt(y), x)  # Property 1: round trip


def test_caller_supplied_cyclic_tables(fhe, torch, oracle):
    """fast_ntt_forward / benchmark backends feed an omega (N-th root) table through the same network."""
    n, q = 1024, Q27
    _, _, psi, psi_inv, inv_n = oracle.twiddles(n, q)
    omega, omega_inv = pow(psi, 2, q), pow(psi_inv, 2, q)
    fwd = np.array([pow(omega, i, q) for i in range(n)], np.uint64)
    inv = np.array([pow(omega_inv, i, q) for i in range(n)], np.uint64)
    ntt = fhe.NTTProcessor(n, q, fwd, inv, inv_n)
    x = np.random.default_rng(5).integers(0, q, size=(6, n), dtype=np.uint64)
    y = ntt.forward_ntt(x)
    eq(y, oracle.forward(x, q, fwd))
    eq(ntt.inverse_ntt_forward_network(y), oracle.fast_inverse(y, q, inv))
    eq(ntt.inverse_ntt_forward_network(y), x)  # with a cyclic table that IS the inverse transform


def test_plan_errors(fhe):
    for args, msg in [((1000, Q27), "power of 2"), ((2, 17), "between 4 and 65536"), ((1024, Q27 + 1), "odd"),
                      ((1024, 1099511627777 - 2), "NTT-friendly"), ((1024, 1099511627777), "primitive root")]:
        with pytest.raises(fhe.FheError) as e:
            fhe.NTTProcessor(*args)
        assert msg in str(e.value), str(e.value)
    ntt = fhe.NTTProcessor(8, 97)
    with pytest.raises(fhe.FheError):
        ntt.forward_ntt(np.zeros(7, np.uint64))
    assert ntt.forward_ntt(np.zeros((0, 8), np.uint64)).shape == (0, 8)  # empty batch


def test_full_size_properties(fhe, torch, oracle):
    """BASELINE C2 sizes (N = 4096 / 16384, batch 1024): round trip, linearity, ring axioms on the whole batch, and 32 rows
    spread over the batch (first, last, one per block of 33) against the CPU oracle for the forward transform and the product."""
    for n in (4096, 16384):
        q = Q62
        ring = fhe.PolynomialRing(n, q)
        g = torch.Generator(device="cuda").manual_seed(n)
        a = torch.randint(0, q, (1024, n), dtype=torch.int64, device="cuda", generator=g)
        b = torch.randint(0, q, (1024, n), dtype=torch.int64, device="cuda", generator=g)
        fa = ring.to_ntt(a)
        assert torch.equal(ring.from_ntt(fa), a)
        rows = sorted(set([0, 1023] + list(range(7, 1024, 33))))[:32]
        fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
        ah, bh = host(a[rows]), host(b[rows])
        eq(host(fa[rows]), oracle.forward(ah, q, fwd))
        eq(host(ring.multiply(a, b)[rows]), oracle.multiply(ah, bh, q, fwd, inv, inv_n))
        # T(a + b) == T(a) + T(b)
        assert torch.equal(ring.to_ntt(ring.add(a, b)), ring.add(fa, ring.to_ntt(b)))
        ab = ring.multiply(a, b)
        assert torch.equal(ab, ring.multiply(b, a))
        # fused product == unfused composition of the same library calls
        assert torch.equal(ab, ring.from_ntt(ring.pointwise_multiply(fa, ring.to_ntt(b))))
        one = torch.zeros_like(a)
        one[:, 0] = 1
        # multiply by T^-1(1,1,..,1) is the identity of this ring's product
        ident = ring.from_ntt(torch.ones_like(a))
        assert torch.equal(ring.multiply(a, ident), a)


# -------------------------------------------------------------------------- elementwise --
@pytest.mark.parametrize("q", [17, Q27, QT, Q62, (1 << 63) + 29, (1 << 64) - 59])
def test_elementwise_matches_oracle(fhe, torch, oracle, q):
    rng = np.random.default_rng(q % 1000)
    for count in (1, 2, 3, 1023, 4096, 100001):
        a = rng.integers(0, 2**64, size=count, dtype=np.uint64)
        b = rng.integers(0, 2**64, size=count, dtype=np.uint64)
        if count > 2:
            a[:3] = [0, q - 1, q]
            b[:3] = [q - 1, q - 1, 0]
        ad, bd = dev(torch, a), dev(torch, b)
        eq(host(fhe.modadd_batch(ad, bd, q)), oracle.add(a, b, q))
        eq(host(fhe.modsub_batch(ad, bd, q)), oracle.sub(a, b, q))
        eq(host(fhe.modmul_batch(ad, bd, q)), oracle.pointwise(a, b, q))
        eq(host(fhe.modneg_batch(dev(torch, a % np.uint64(q)), q)), oracle.negate(a % np.uint64(q), q))
        s = int(b[0])
        eq(host(fhe.modmul_scalar_batch(ad, s, q)), oracle.scalar(a, s, q))
    # host buffers, odd alignment (16-byte vector path must fall back) and aliasing r == a
    a = rng.integers(0, 2**64, size=1001, dtype=np.uint64)
    b = rng.integers(0, 2**64, size=1001, dtype=np.uint64)
    eq(fhe.modmul_batch(a[1:], b[1:], q), oracle.pointwise(a[1:], b[1:], q))
    ad = dev(torch, a)
    exp = oracle.add(a, b, q)
    fhe.modadd_batch(ad, dev(torch, b), q, out=ad)
    eq(host(ad), exp)


def test_golden_multi_limb(fhe, torch):
    g = np.load(os.path.join(GOLDEN, "mlimb_q65.npz"))
    ml = fhe.MultiLimbModularArithmetic(g["q"])
    assert ml.q_inv == int(g["q_inv"])
    eq(ml.r_mod_q, g["r_mod_q"])
    eq(ml.r2_mod_q, g["r2_mod_q"])
    a, b = g["a"], g["b"]
    for buf in (lambda x: x, lambda x: dev(torch, x)):
        out = lambda t: t if isinstance(t, np.ndarray) else host(t)
        eq(out(ml.montgomery_mul(buf(a), buf(b))), g["montmul"])
        eq(out(ml.mod_add(buf(a), buf(b))), g["add"])
        eq(out(ml.mod_sub(buf(a), buf(b))), g["sub"])
        eq(out(ml.to_montgomery(buf(a))), g["to_mont"])
        eq(out(ml.from_montgomery(buf(a))), g["from_mont"])


@pytest.mark.parametrize("q_limbs", [[0xFFFFFFFFFFFFFF43, 1], [0xFFFFFFFFFFFFFFC5], [0x1D, 0, 1],
                                     [0xFFFFFFFFFFFFFF61, 0xFFFFFFFFFFFFFFFF], [0x2F, 5, 0, 9], [3, 0, 0, 0, 0, 0, 0, 1]])
def test_multi_limb_matches_oracle(fhe, torch, oracle, q_limbs):
    q = np.array(q_limbs, np.uint64)
    l = q.size
    qi = sum(int(v) << (64 * i) for i, v in enumerate(q))
    q_inv, r1, r2 = oracle.mlimb_constants(q)
    ml = fhe.MultiLimbModularArithmetic(q)
    assert ml.q_inv == q_inv
    eq(ml.r_mod_q, r1)
    eq(ml.r2_mod_q, r2)
    rng = np.random.default_rng(l)
    count = 65536 if l == 2 else 3001  # BASELINE C3: n = 65536 two-limb elements
    vals = [int.from_bytes(rng.bytes(8 * l + 1), "little") % qi for _ in range(2 * count)]
    vals[:4] = [0, 1, qi - 1, qi - 2]
    arr = np.array([[(v >> (64 * i)) & (2**64 - 1) for i in range(l)] for v in vals], np.uint64)
    a, b = arr[:count], arr[count:]
    ad, bd = dev(torch, a), dev(torch, b)
    eq(host(ml.montgomery_mul(ad, bd)), oracle.mlimb_montmul(a, b, q, q_inv))
    eq(host(ml.mod_add(ad, bd)), oracle.mlimb_add(a, b, q))
    eq(host(ml.mod_sub(ad, bd)), oracle.mlimb_sub(a, b, q))
    if l == 2 and qi < 2**127:  # independent check with Python integers (the reference drops the top carry near 2^128)
        got = host(ml.montgomery_mul(ad[:64], bd[:64]))
        rinv = pow(1 << 128, -1, qi)
        for i in range(64):
            assert int(got[i, 0]) + (int(got[i, 1]) << 64) == vals[i] * vals[count + i] * rinv % qi


# -------------------------------------------------------------------------------- tally --
@pytest.mark.parametrize("name", ["tally_n64_m5.npz", "tally_n1024_m9.npz"])
def test_golden_tally(fhe, torch, name):
    g = np.load(os.path.join(GOLDEN, name))
    n, q, cts = int(g["n"]), int(g["q"]), g["cts"]
    eq(fhe.tally_votes(cts, n, q), g["linear"])
    eq(host(fhe.tally_votes(dev(torch, cts), n, q)), g["tree"])
    eq(fhe.tally_votes(cts[1:2], n, q), g["single"])  # one ballot: words returned untouched
    with pytest.raises(fhe.FheError) as e:
        fhe.tally_votes(cts[:0], n, q)
    assert "Cannot add empty vector of ciphertexts" in str(e.value)
    ring = fhe.PolynomialRing(n, q)
    eq(ring.tensor_multiply(cts[0:1], cts[2:3])[0], g["tensor"])


def test_rns_ring_every_limb_matches_the_oracle(fhe, torch, oracle):
    """PolynomialRing(degree, moduli): limb 0 is the reference's result (it only ever uses moduli[0]); here every limb
    of the chain is live and equals the reference method over its own modulus."""
    n, batch = 4096, 6
    moduli = [1152921504606584833, Q62, Q50, Q27]  # Q_60_1, the harness prime, Q_50_1 and an FP64-mode prime
    ring = fhe.RnsPolynomialRing(n, moduli)
    rng = np.random.default_rng(77)
    a = np.stack([rng.integers(0, q, size=(batch, n), dtype=np.uint64) for q in moduli])
    b = np.stack([rng.integers(0, q, size=(batch, n), dtype=np.uint64) for q in moduli])
    ad, bd = dev(torch, a), dev(torch, b)
    prod, fwd_, back, summ = host(ring.multiply(ad, bd)), host(ring.to_ntt(ad)), None, host(ring.add(ad, bd))
    back = host(ring.from_ntt(dev(torch, fwd_)))
    eq(back, a)
    for l, q in enumerate(moduli):
        fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
        eq(prod[l], oracle.multiply(a[l], b[l], q, fwd, inv, inv_n))
        eq(fwd_[l], oracle.forward(a[l], q, fwd))
        eq(summ[l], oracle.add(a[l], b[l], q))
    with pytest.raises(fhe.FheError) as e:
        fhe.RnsPolynomialRing(n, [])
    assert "At least one modulus required" in str(e.value)
    with pytest.raises(fhe.FheError):
        fhe.RnsPolynomialRing(n, [Q62, 1099511627777])  # 2^40 + 1 (the tfhe-128-fast preset's modulus) is not NTT friendly


# ----------------------------------------------------------------------- relinearisation --
@pytest.mark.parametrize("name", ["relin_n64.npz", "relin_n1024.npz", "relin_n4096.npz"])
def test_golden_relinearize(fhe, torch, name):
    """EncryptionEngine::relinearize against outputs of the reference's own code."""
    g = np.load(os.path.join(GOLDEN, name))
    n, q, bl, lv = int(g["n"]), int(g["q"]), int(g["base_log"]), int(g["level"])
    ring = fhe.PolynomialRing(n, q)
    key = fhe.RelinearizationKey(ring, g["keys"], bl, lv, key_id=7)
    eq(key.relinearize(g["ct"]), g["out"])                        # host buffers (pipelined staging)
    eq(host(key.relinearize(dev(torch, g["ct"]))), g["out"])      # device buffers
    eq(host(key.relinearize(dev(torch, g["ct"][1:2])))[0], g["out"][1])
    nokey = fhe.RelinearizationKey(ring, g["keys"][:0], bl, lv)
    assert nokey.levels == 0
    eq(host(nokey.relinearize(dev(torch, g["ct"][1:2])))[0], g["nokey"])
    with pytest.raises(fhe.FheError) as e:
        key.relinearize(g["ct"], ct_key_id=8)
    assert e.value.code == 2 and "Evaluation key does not match ciphertext key" in str(e.value)
    with pytest.raises(fhe.FheError):
        fhe.RelinearizationKey(ring, g["keys"], 40, 3)  # third shift would reach 80 bits
    ctd = dev(torch, g["ct"])
    with pytest.raises(fhe.FheError) as e2:   # 3 N words in, 2 N words out: the buffers must not overlap
        key.relinearize(ctd, out=ctd.view(-1)[: ctd.shape[0] * 2 * ctd.shape[-1]].view(ctd.shape[0], 2, ctd.shape[-1]))
    assert "overlap" in str(e2.value)


@pytest.mark.parametrize("n,q,key_count,bl,lv,batch", [(256, Q27, 16, 0, 0, 5), (2048, 1125899906826241, 3, 17, 3, 9),
                                                       (8192, Q62, 2, 31, 2, 3), (1024, QT, 6, 7, 6, 33)])
def test_relinearize_matches_oracle(fhe, torch, oracle, n, q, key_count, bl, lv, batch):
    rng = np.random.default_rng(n + batch)
    fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
    ring = fhe.PolynomialRing(n, q)
    a = rng.integers(0, q, size=(batch, 2, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(batch, 2, n), dtype=np.uint64)
    ct3 = ring.tensor_multiply(dev(torch, a), dev(torch, b))      # multiply -> relinearize, the reference's chain
    keys = rng.integers(0, q, size=(key_count, 2, n), dtype=np.uint64)
    key = fhe.RelinearizationKey(ring, dev(torch, keys), bl, lv)
    got = host(key.relinearize(ct3))
    ct3h = host(ct3)
    for i in range(batch):
        eq(got[i], oracle.relinearize(ct3h[i], keys, bl, lv, q, fwd, inv, inv_n))


def test_relinearize_fused_and_unfused_paths_agree(fhe, torch, oracle, monkeypatch):
    """The one-launch kernel (relin_fused.cu) and the five-launch path (relin.cu) give the oracle's words; shapes whose digit
    rows do not fit an SM (N = 8192 with four levels: 256 KB) fall back to the five-launch path by themselves."""
    for n, q, levels, bl, batch in [(4096, Q62, 4, 15, 3), (512, QT, 3, 13, 7), (8192, Q62, 4, 15, 2)]:
        rng = np.random.default_rng(n)
        fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
        ring = fhe.PolynomialRing(n, q)
        ct3 = rng.integers(0, q, size=(batch, 3, n), dtype=np.uint64)
        ct3[0, :2] = rng.integers(0, 2**64, size=(2, n), dtype=np.uint64)  # unreduced c0 / c1: add_inplace reduces them
        keys = rng.integers(0, q, size=(levels, 2, n), dtype=np.uint64)
        key = fhe.RelinearizationKey(ring, dev(torch, keys), bl, levels)
        exp = np.stack([oracle.relinearize(ct3[i], keys, bl, levels, q, fwd, inv, inv_n) for i in range(batch)])
        monkeypatch.delenv("FHEB_RELIN_UNFUSED", raising=False)
        eq(host(key.relinearize(dev(torch, ct3))), exp)
        monkeypatch.setenv("FHEB_RELIN_UNFUSED", "1")
        eq(host(key.relinearize(dev(torch, ct3))), exp)
        monkeypatch.delenv("FHEB_RELIN_UNFUSED", raising=False)


@pytest.mark.parametrize("count", [2, 3, 63, 64, 65, 1000, 20000])
def test_tally_matches_oracle(fhe, torch, oracle, count):
    n, q = 1024, QT
    rng = np.random.default_rng(count)
    cts = rng.integers(0, q, size=(count, 2, n), dtype=np.uint64)
    if count == 3:
        cts[1] = rng.integers(0, 2**64, size=(2, n), dtype=np.uint64)  # unreduced words are reduced by add
    exp = oracle.tally(cts, q)
    eq(host(fhe.tally_votes(dev(torch, cts), n, q)), exp)
    if count >= 64:  # sharded shape: per-shard partials, then the combine kernel
        parts = np.stack([host(fhe.tally_votes(dev(torch, cts[lo:hi]), n, q))
                          for lo, hi in [fhe.shard_range(count, r, 4) for r in range(4)]])
        eq(host(fhe.tally_combine(dev(torch, parts), n, q)), exp)


def test_synthetic_ballots_reproducible_and_tally_checksum(fhe, torch, oracle):
    """The on-device ballot generator is restated on the CPU; a 64k-ballot tally is checked through
    a checksum of per-chunk oracle tallies (size-independent: the tally of tallies)."""
    n, q, seed = 1024, QT, 0xB200
    count = 1 << 16
    cts = torch.empty((count, 2, n), dtype=torch.int64, device="cuda")
    fhe.synth_ballots(cts, 0, count, n, q, seed)

    def splitmix(idx):
        x = (idx + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))

    with np.errstate(over="ignore"):
        words = splitmix(np.uint64(seed) + np.arange(300 * 2 * n, dtype=np.uint64)) % np.uint64(q)
    eq(host(cts[:300]).reshape(-1), words)
    total = host(fhe.tally_votes(cts, n, q))
    chunks = np.stack([oracle.tally(host(cts[i:i + 8192]), q) for i in range(0, count, 8192)])
    eq(total, oracle.tally(chunks, q))


# ------------------------------------------------------------------- host-buffer pipeline --
def test_host_buffers_are_pipelined_in_chunks_and_match_device_path(fhe, torch, oracle):
    """Host buffers larger than one staging chunk (8 MB) go through the chunked copy/compute/copy
    pipeline; results must be identical to the device-buffer path (and to the oracle on a sample)."""
    n, q = 16384, Q62
    ring = fhe.PolynomialRing(n, q)
    rng = np.random.default_rng(77)
    a = rng.integers(0, q, size=(203, n), dtype=np.uint64)  # 26.6 MB: 4 chunks, the last one ragged
    b = rng.integers(0, q, size=(203, n), dtype=np.uint64)
    ad, bd = dev(torch, a), dev(torch, b)
    fa = ring.to_ntt(a)
    eq(fa, host(ring.to_ntt(ad)))
    eq(ring.from_ntt(fa), a)
    inplace = a.copy()
    ring.to_ntt(inplace, out=inplace)  # in == out on the host
    eq(inplace, fa)
    prod = ring.multiply(a, b)
    eq(prod, host(ring.multiply(ad, bd)))
    fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
    eq(prod[[0, 64, 202]], oracle.multiply(a[[0, 64, 202]], b[[0, 64, 202]], q, fwd, inv, inv_n))
    # element-wise and multi-limb vectors above the pipelining threshold, ragged tail
    x = rng.integers(0, 2**64, size=(1 << 21) + 4099, dtype=np.uint64)
    y = rng.integers(0, 2**64, size=(1 << 21) + 4099, dtype=np.uint64)
    eq(fhe.modmul_batch(x, y, q), host(fhe.modmul_batch(dev(torch, x), dev(torch, y), q)))
    eq(fhe.modneg_batch(x % np.uint64(q), q), oracle.negate(x % np.uint64(q), q))
    ml = fhe.MultiLimbModularArithmetic([0xFFFFFFFFFFFFFF43, 1])
    u = rng.integers(0, 2**64, size=(1 << 20, 2), dtype=np.uint64)
    u[:, 1] &= np.uint64(1)
    v = u[::-1].copy()
    eq(ml.montgomery_mul(u, v), host(ml.montgomery_mul(dev(torch, u), dev(torch, v))))
    # tally of host ballots: chunk partials folded at the end
    cts = rng.integers(0, QT, size=(3001, 2, 1024), dtype=np.uint64)
    eq(fhe.tally_votes(cts, 1024, QT), oracle.tally(cts, QT))


def test_tensor_multiply_batch_matches_oracle(fhe, torch, oracle):
    """EncryptionEngine::multiply tensor product over a batch (one fused pointwise launch)."""
    for n, q in [(1024, QT), (4096, Q62)]:
        fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
        rng = np.random.default_rng(n)
        ct1 = rng.integers(0, q, size=(9, 2, n), dtype=np.uint64)
        ct2 = rng.integers(0, q, size=(9, 2, n), dtype=np.uint64)
        ring = fhe.PolynomialRing(n, q)
        got = host(ring.tensor_multiply(dev(torch, ct1), dev(torch, ct2)))
        exp = np.stack([oracle.tensor_multiply(ct1[i], ct2[i], q, fwd, inv, inv_n) for i in range(9)])
        eq(got, exp)
        eq(ring.tensor_multiply(ct1[:2], ct2[:2]), exp[:2])  # host buffers
        # output overlapping an operand (the first operand sits inside the output buffer): same words (the library takes
        # the path that has read both operands before it writes)
        buf = torch.zeros((9 * 3 * n,), dtype=torch.int64, device="cuda")
        a_in = buf[n:n + 9 * 2 * n].view(9, 2, n)
        a_in.copy_(dev(torch, ct1))
        eq(host(ring.tensor_multiply(a_in, dev(torch, ct2), out=buf.view(9, 3, n))), exp)


def test_streaming_tally_accumulator(fhe, torch, oracle):
    """Running device-resident tally (stream_add semantics): chunks of ballots, host and device, any sizes."""
    n, q = 1024, QT
    rng = np.random.default_rng(9)
    cts = rng.integers(0, q, size=(1500, 2, n), dtype=np.uint64)
    cts[0] = rng.integers(0, 2**64, size=(2, n), dtype=np.uint64)  # the first ballot may hold unreduced words
    acc = fhe.CiphertextStreamAccumulator(n, q)
    with pytest.raises(fhe.FheError):
        acc.total()  # nothing added yet
    acc.add(cts[0:1])
    eq(acc.total(), cts[0])  # first ciphertext becomes the accumulator untouched
    done = 1
    for size, on_device in [(1, False), (2, True), (61, False), (700, True), (735, False)]:
        chunk = cts[done:done + size]
        acc.add(dev(torch, chunk) if on_device else chunk)
        done += size
        assert acc.count == done
        eq(acc.total(), oracle.tally(cts[:done], q))
    out = torch.empty((2, n), dtype=torch.int64, device="cuda")
    acc.total(out=out)
    eq(host(out), oracle.tally(cts, q))


# ---------------------------------------------------------------------------- edge cases --
def test_edge_cases_empty_aliased_unaligned(fhe, torch, oracle):
    n, q = 256, QT
    ring = fhe.PolynomialRing(n, q)
    fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
    rng = np.random.default_rng(3)
    # empty batches are no-ops everywhere
    e = np.zeros((0, n), np.uint64)
    assert ring.to_ntt(e).shape == ring.from_ntt(e).shape == ring.multiply(e, e).shape == ring.add(e, e).shape == (0, n)
    assert host(ring.multiply(dev(torch, e), dev(torch, e))).shape == (0, n)
    # product written over either operand, device and host
    a = rng.integers(0, q, size=(7, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(7, n), dtype=np.uint64)
    exp = oracle.multiply(a, b, q, fwd, inv, inv_n)
    ad, bd = dev(torch, a), dev(torch, b)
    ring.multiply(ad, bd, out=ad)
    eq(host(ad), exp)
    ad = dev(torch, a)
    ring.multiply(ad, bd, out=bd)
    eq(host(bd), exp)
    ah = a.copy()
    ring.multiply(ah, b, out=ah)
    eq(ah, exp)
    sq = oracle.multiply(a, a, q, fwd, inv, inv_n)
    ad = dev(torch, a)
    eq(host(ring.multiply(ad, ad)), sq)
    # 8-byte aligned but not 16-byte aligned device pointers (vector path must not be taken)
    buf = dev(torch, np.concatenate([[np.uint64(0)], a.reshape(-1)]))
    bufb = dev(torch, np.concatenate([[np.uint64(0)], b.reshape(-1)]))
    va, vb = buf[1:], bufb[1:]
    eq(host(fhe.modmul_batch(va, vb, q)), oracle.pointwise(a.reshape(-1), b.reshape(-1), q))
    eq(host(ring.to_ntt(va.view(7, n))), oracle.forward(a, q, fwd))
    cts = rng.integers(0, q, size=(33, 2, n), dtype=np.uint64)
    odd = dev(torch, np.concatenate([[np.uint64(0)], cts.reshape(-1)]))[1:].view(33, 2, n)
    eq(host(fhe.tally_votes(odd, n, q)), oracle.tally(cts, q))
    # all-zero and all-(q-1) polynomials, single polynomial batches
    for v in (0, q - 1):
        x = np.full((1, n), v, np.uint64)
        eq(ring.to_ntt(x), oracle.forward(x, q, fwd))
        eq(ring.multiply(x, x), oracle.multiply(x, x, q, fwd, inv, inv_n))


def test_randomised_shapes_match_oracle(fhe, torch, oracle):
    """Property-style sweep (fixed seed): random degree, modulus class, batch, value class."""
    rng = np.random.default_rng(2026)
    primes = {False: [Q62, 1152921504606584833, Q50], True: [QT, Q27, 97 * 0 + 1099511592961]}
    for trial in range(24):
        dp = bool(trial & 1)
        logn = int(rng.integers(2, 13))
        n = 1 << logn
        cands = [p for p in primes[dp] if (p - 1) % (2 * n) == 0]
        if not cands:
            continue
        q = cands[int(rng.integers(0, len(cands)))]
        batch = int(rng.integers(1, 40))
        fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
        ring = fhe.PolynomialRing(n, q)
        hi = 2**64 if trial % 3 == 0 else q
        a = rng.integers(0, hi, size=(batch, n), dtype=np.uint64)
        b = rng.integers(0, hi, size=(batch, n), dtype=np.uint64)
        ad, bd = dev(torch, a), dev(torch, b)
        eq(host(ring.to_ntt(ad)), oracle.forward(a, q, fwd))
        eq(host(ring.from_ntt(ad)), oracle.inverse(a, q, inv, inv_n))
        eq(host(ring.multiply(ad, bd)), oracle.multiply(a, b, q, fwd, inv, inv_n))
        eq(host(ring.subtract(ad, bd)), oracle.sub(a, b, q))


@pytest.mark.parametrize("q", [Q62, Q27], ids=["q62", "q27-fp64"])
@pytest.mark.parametrize("logn", [15, 16])
def test_large_degrees_two_launch_path(fhe, torch, oracle, logn, q):
    """Degrees above 2^14 (the reference accepts up to 65536): top stages + 2^14 sub-transforms."""
    n = 1 << logn
    ring = fhe.PolynomialRing(n, q)
    fwd, inv, psi, psi_inv, inv_n = oracle.twiddles(n, q)
    assert ring.ntt.get_twiddles()[2:] == (psi, psi_inv, inv_n)
    rng = np.random.default_rng(logn)
    x = rng.integers(0, q, size=(3, n), dtype=np.uint64)
    x[2] = rng.integers(0, 2**64, size=n, dtype=np.uint64)  # unreduced words
    xd = dev(torch, x)
    y = ring.to_ntt(xd)
    eq(host(y), oracle.forward(x, q, fwd))
    eq(host(ring.from_ntt(xd)), oracle.inverse(x, q, inv, inv_n))
    back = ring.from_ntt(y)
    eq(host(back), x % np.uint64(q))
    ring.to_ntt(xd, out=xd)  # in place
    eq(host(xd), oracle.forward(x, q, fwd))
    a, b = x[:2], x[1:3]
    eq(host(ring.multiply(dev(torch, a), dev(torch, b))), oracle.multiply(a, b, q, fwd, inv, inv_n))
    eq(ring.multiply(a, b), oracle.multiply(a, b, q, fwd, inv, inv_n))  # host buffers
