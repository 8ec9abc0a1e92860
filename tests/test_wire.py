"""Wire formats: the reference's FHEV ballot / FHEE eval-key / FHEB bootstrap-key containers
(cpp/src/key_serializer.cpp).  CPU tests pin the C restatement and the host-side C-ABI functions to the reference's
own serializer (compiled under oracle/_ref) and to records it wrote (tests/golden/wire_n64.npz); the GPU tests run
the device ingest path against both."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
QT = 1099511678977
Q62 = 4611686018326724609


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b)


@pytest.fixture(scope="module")
def wire():
    g = np.load(os.path.join(GOLDEN, "wire_n64.npz"))
    blob = g["records"].tobytes()
    offs = np.concatenate([[0], np.cumsum(g["sizes"])]).astype(np.uint64)
    recs = [blob[int(offs[i]):int(offs[i + 1])] for i in range(len(g["sizes"]))]
    return g, recs


def corruptions(rec):
    """(label, bytes, expected status or None) - the three rejections of deserialize_ballot.  None: whatever the
    reference's checksum makes of it.  With 196 of its 256 table entries zero the CRC state decays to the last few
    payload bytes, so a damaged payload is usually NOT detected (the reference accepts it too); parity, not
    detection, is what is tested for those."""
    flip = bytearray(rec)
    flip[200] ^= 0x40
    tail = bytearray(rec)
    tail[-1] ^= 0x01
    magic = bytearray(rec)
    magic[0] ^= 1
    crc = bytearray(rec)
    crc[47] ^= 0x80
    return [("ok", rec, 0), ("payload bit flipped", bytes(flip), None), ("last byte flipped", bytes(tail), None),
            ("bad magic", bytes(magic), 2), ("checksum field", bytes(crc), 3), ("truncated", rec[:-5], None),
            ("too small", rec[:63], 1), ("empty", b"", 1)]


# ------------------------------------------------------------------------------- CPU --
def test_crc_is_the_references_not_ieee(oracle, wire):
    import zlib

    g, _ = wire
    probe = g["probe"].tobytes()
    assert oracle.crc32(probe) == int(g["probe_crc"])
    assert oracle.crc32(b"") == int(g["empty_crc"]) == 0
    assert oracle.crc32(probe) != zlib.crc32(probe)  # only 60 table entries are initialised (key_serializer.cpp:21-32)


def test_oracle_parses_reference_records(oracle, wire):
    g, recs = wire
    n, q, ch = int(g["n"]), int(g["q"]), int(g["choices"])
    for i, rec in enumerate(recs):
        st, cts, ts = oracle.ballot_parse(rec, ch, n, q)
        assert st == 0 and ts == 1_700_000_000 + i
        eq(cts, g["ballots"][i])
    for label, bad, want in corruptions(recs[1]):
        st, cts, _ = oracle.ballot_parse(bad, ch, n, q)
        assert want is None or st == want, label
        if st:
            assert not cts.any()
    assert oracle.ballot_parse(recs[0], ch + 1, n, q)[0] == 4
    assert oracle.ballot_parse(recs[0], ch, n, q + 2)[0] == 4


def test_oracle_matches_reference_live(oracle, ref):
    rng = np.random.default_rng(5)
    for n, q, ch in [(8, 17, 1), (64, QT, 3), (1024, Q62, 1)]:
        cts = rng.integers(0, q, size=(ch, 2, n), dtype=np.uint64)
        rec = ref.serialize_ballot(cts, q, 42)
        got, ts = ref.deserialize_ballot(rec, n, q)
        eq(got, cts)
        st, mine, ts2 = oracle.ballot_parse(rec, ch, n, q)
        assert st == 0 and ts == ts2 == 42
        eq(mine, cts)
        messages = {1: "Input too small", 2: "Invalid magic bytes for ballot", 3: "Checksum verification failed"}
        from oracle_bindings import RefError

        for label, bad, want in corruptions(rec):
            st, mine, _ = oracle.ballot_parse(bad, ch, n, q)
            assert want is None or st == want, label
            if st in messages:
                with pytest.raises(RefError) as e:
                    ref.deserialize_ballot(bad, n, q)
                assert messages[st] in str(e.value), label
            elif st == 0:  # accepted by both, damaged words included
                eq(ref.deserialize_ballot(bad, n, q)[0], mine)
        for size in (0, 1, 59, 60, 61, 255, 4096):
            blob = bytes(rng.integers(0, 256, size=size, dtype=np.uint8))
            assert oracle.crc32(blob) == ref.crc32(blob)


def test_cabi_host_side_wire_functions(oracle, wire):
    """fheb_wire_crc32 / fheb_wire_header_read / fheb_ballot_serialize are host code: no GPU needed."""
    import fheb200

    g, recs = wire
    n, q, ch = int(g["n"]), int(g["q"]), int(g["choices"])
    assert fheb200.wire_crc32(g["probe"].tobytes()) == int(g["probe_crc"])
    for i, rec in enumerate(recs):
        assert fheb200.serialize_ballot(g["ballots"][i], q, 1_700_000_000 + i) == rec  # byte-identical to the reference's record
        h = fheb200.wire_header(rec)
        assert (h.magic, h.version, h.key_type, h.key_id) == (0x46484556, 1, 4, 1_700_000_000 + i)
        assert (h.poly_degree, h.modulus, h.data_size, h.checksum_type, h.compression) == (n, q, len(rec) - 49, 1, 0)
        assert h.checksum == oracle.crc32(rec[49:])
    h = fheb200.wire_header(g["eval_key"].tobytes())
    assert (h.magic, h.key_type, h.key_id, h.poly_degree, h.modulus) == (0x46484545, 2, 0xABCDEF, n, q)
    with pytest.raises(fheb200.FheError) as e:
        fheb200.wire_header(recs[0][:20])
    assert "Failed to read header" in str(e.value)


def test_empty_ballot_is_written_but_never_read_back(oracle, ref):
    """A ballot without choices serialises to 61 bytes - below sizeof(SerializationHeader) = 64, so the reference's own
    deserialize_ballot rejects what its serialize_ballot wrote ("Input too small").  Same bytes, same verdict here."""
    import fheb200
    from oracle_bindings import RefError

    none = np.zeros((0, 2, 8), np.uint64)
    rec = ref.serialize_ballot(none, 97, 5)
    assert len(rec) == 61 and fheb200.serialize_ballot(none, 97, 5) == rec
    assert oracle.ballot_parse(rec, 0, 8, 97)[0] == 1
    with pytest.raises(RefError) as e:
        ref.deserialize_ballot(rec, 8, 97)
    assert "Input too small" in str(e.value)


def test_speculative_checksum_schedule_is_exact():
    """The warp-parallel schedule of ballot_validate_kernel (csrc/wire.cu) restated lane by lane: 32 slices, start
    states guessed from `warmup` bytes run from state 0, neighbour checks, repair rounds until every check holds.
    Whatever the guesses are worth, the end state must be the serial checksum's (the table is not linear, so this is
    a property of the schedule, not of the CRC)."""
    table = [0] * 256
    for i in range(60):
        c = i
        for _ in range(8):
            c = (0xEDB88320 ^ (c >> 1)) if c & 1 else c >> 1
        table[i] = c

    def run(state, data):
        for b in data:
            state = table[(state ^ b) & 0xFF] ^ (state >> 8)
        return state

    rng = np.random.default_rng(2024)
    total_rounds = 0
    for trial in range(120):
        n = int(rng.integers(4096, 9000))
        kind = trial % 3
        if kind == 0:
            data = rng.integers(0, 256, size=n, dtype=np.uint8)
        elif kind == 1:   # 62-bit residues: the slowest to forget the start state
            data = rng.integers(0, 4611686018326724609, size=n // 8 + 1, dtype=np.uint64).view(np.uint8)[:n]
        else:             # low bytes only: long stretches of table entries below 60
            data = rng.integers(0, 60, size=n, dtype=np.uint8)
        data = [int(x) for x in data]
        warmup = (0, 8, 128)[trial % 3 if trial >= 60 else 2]
        chunk = (n + 31) // 32
        lo = [min(k * chunk, n) for k in range(32)]
        hi = [min(l + chunk, n) for l in lo]
        warm = [0 if l < warmup else l - warmup for l in lo]
        a = [run(0xFFFFFFFF if warm[k] == 0 else 0, data[warm[k]:lo[k]]) for k in range(32)]
        e = [run(a[k], data[lo[k]:hi[k]]) for k in range(32)]
        rounds = 0
        while True:
            ok = [k == 0 or (warm[k] == 0 and warmup != 0) or e[k - 1] == a[k] for k in range(32)]
            if all(ok):
                break
            prev = list(e)
            for k in range(32):
                if not ok[k]:
                    a[k] = prev[k - 1]
                    e[k] = run(a[k], data[lo[k]:hi[k]])
            rounds += 1
            assert rounds <= 32
        total_rounds += rounds
        assert e[31] == run(0xFFFFFFFF, data), (trial, warmup)
    assert total_rounds > 0  # the forced-miss trials (warmup 0 and 8) did exercise the repair rounds


# ------------------------------------------------------------------------------- GPU --
@pytest.fixture(scope="module")
def fhe():
    import fheb200

    fheb200.initialize()
    return fheb200


def host(t):
    return t.cpu().numpy().view(np.uint64) if hasattr(t, "cpu") else np.asarray(t)


@pytest.mark.gpu
def test_ingest_golden_records(fhe, wire, oracle):
    import torch

    g, recs = wire
    n, q, ch = int(g["n"]), int(g["q"]), int(g["choices"])
    # clean stream, self-describing (no offsets), host output
    cts, status, stamps = fhe.ingest_ballots(b"".join(recs), len(recs), ch, n, q)
    assert not status.any()
    eq(cts, g["ballots"])
    eq(stamps, 1_700_000_000 + np.arange(len(recs), dtype=np.uint64))
    # every rejection, explicit extents, device output; rejected records are zero ciphertexts
    mixed = [b for _, b, _ in corruptions(recs[1])] + [recs[3]]
    want = [oracle.ballot_parse(b, ch, n, q)[0] for b in mixed]
    assert {1, 2, 3} <= set(want) and want[0] == 0 and want[-1] == 0
    offs = np.concatenate([[0], np.cumsum([len(b) for b in mixed])]).astype(np.uint64)
    cts, status, stamps = fhe.ingest_ballots(b"".join(mixed), len(mixed), ch, n, q, offsets=offs, device="cuda")
    eq(status, np.array(want, np.uint8))
    got = host(cts)
    for i, rec in enumerate(mixed):
        st, exp, ts = oracle.ballot_parse(rec, ch, n, q)
        assert st == want[i]
        eq(got[i], exp)
        assert int(stamps[i]) == ts
    # wire bytes already on the device (8-byte aligned), unaligned records inside
    blob = np.frombuffer(b"".join(mixed), np.uint8)
    dcts, status2, _ = fhe.ingest_ballots(torch.from_numpy(blob.copy()).cuda(), len(mixed), ch, n, q, offsets=offs)
    eq(status2, status)
    eq(host(dcts), got)
    # wrong expected shape
    _, status3, _ = fhe.ingest_ballots(b"".join(recs), len(recs), ch, n, q + 2)
    assert (status3 == 4).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n,q,ch,count", [(1024, QT, 1, 300), (256, 132120577, 3, 77), (4096, Q62, 1, 9)])
def test_ingest_then_tally_matches_oracle(fhe, oracle, n, q, ch, count):
    """The production chain either side of C5: wire records -> device ingest -> tally, with damaged ballots dropped."""
    rng = np.random.default_rng(count)
    ballots = rng.integers(0, q, size=(count, ch, 2, n), dtype=np.uint64)
    recs = [fhe.serialize_ballot(ballots[i], q, i) for i in range(count)]
    damaged = [int(x) for x in rng.choice(count, size=count // 4, replace=False)]
    for k, i in enumerate(damaged):
        b = bytearray(recs[i])
        if k % 3 == 0:
            b[45 + k % 4] ^= 0x10  # the checksum field: always rejected
        elif k % 3 == 1:
            b[len(b) - 1 - k % 6] ^= 1 << (k % 8)  # the payload's tail: sometimes caught by the reference's CRC
        else:
            b[49 + int(rng.integers(0, len(b) - 49))] ^= 1 << int(rng.integers(0, 8))  # anywhere: usually not caught
        recs[i] = bytes(b)
    cts, status, _ = fhe.ingest_ballots(b"".join(recs), count, ch, n, q, device="cuda")
    got = host(cts)
    parsed = np.zeros_like(ballots)
    for i in range(count):  # record by record against the restated deserialize_ballot
        st, parsed[i], _ = oracle.ballot_parse(recs[i], ch, n, q)
        assert status[i] == st
        eq(got[i], parsed[i])
    good = [i for i in range(count) if status[i] == 0]
    assert 0 < len(good) < count and set(status) <= {0, 3}
    for c in range(ch):  # tally over the whole ingested buffer == oracle tally of the accepted ballots
        eq(host(fhe.tally_votes(cts[:, c].contiguous(), n, q)), oracle.tally(parsed[good][:, c], q))


@pytest.mark.gpu
def test_keys_from_wire(fhe, wire, oracle):
    import torch

    g, _ = wire
    n, q = int(g["n"]), int(g["q"])
    ring = fhe.PolynomialRing(n, q)
    key = fhe.RelinearizationKey.from_wire(ring, g["eval_key"].tobytes())
    assert key.key_id == 0xABCDEF and key.levels == 3
    fwd, inv, _, _, inv_n = oracle.twiddles(n, q)
    ct3 = np.random.default_rng(3).integers(0, q, size=(4, 3, n), dtype=np.uint64)
    got = key.relinearize(ct3)
    for i in range(4):
        eq(got[i], oracle.relinearize(ct3[i], g["keys"], 12, 3, q, fwd, inv, inv_n))
    bad = bytearray(g["eval_key"].tobytes())
    bad[46] ^= 2  # the checksum field (a damaged payload usually passes the reference's CRC, see corruptions())
    with pytest.raises(fhe.FheError) as e:
        fhe.RelinearizationKey.from_wire(ring, bytes(bad))
    assert "Checksum verification failed" in str(e.value)
    with pytest.raises(fhe.FheError) as e:
        fhe.RelinearizationKey.from_wire(ring, g["records"].tobytes())
    assert "Invalid magic bytes" in str(e.value)
    # FHEB container -> device bootstrap key: same external products as a key made from the raw words
    blob = g["boot_key"].tobytes()
    h = fhe.wire_header(blob)
    assert (h.magic, h.key_type, h.key_id, h.poly_degree, h.modulus) == (0x46484542, 3, 0x5EED, n, q)
    eng_w = fhe.BootstrapEngine.from_wire(blob, lwe_dimension=6, decomp_base_log=9, decomp_level=2)
    eng_r = fhe.BootstrapEngine(n, q, 6, 1, 9, 2, g["bsk"])
    glwe = np.random.default_rng(4).integers(0, q, size=(3, 2, n), dtype=np.uint64)
    p = oracle.boot_params(n, q, 6, 1, 9, 2, 4, fwd, inv, inv_n)
    for idx in (0, 5):
        a = eng_w.external_product(glwe, idx)
        eq(a, eng_r.external_product(glwe, idx))
        for i in range(3):
            eq(a[i], oracle.external_product(p, glwe[i], g["bsk"][idx]))
    with pytest.raises(fhe.FheError):
        fhe.BootstrapEngine.from_wire(blob, lwe_dimension=7, decomp_base_log=9, decomp_level=2)
    with pytest.raises(fhe.FheError):
        fhe.BootstrapEngine.from_wire(blob, lwe_dimension=6, decomp_base_log=9, decomp_level=3)


@pytest.mark.gpu
def test_checksum_repair_rounds_are_exact():
    """FHEB_WIRE_WARMUP=0 makes every lane's start-state guess wrong, so the warp-parallel checksum has to repair
    all 31 slices; the statuses must still be the serial checksum's (run in a subprocess: the knob is read once)."""
    import subprocess
    import sys

    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests")
import fheb200
from oracle_bindings import Oracle
fheb200.initialize(); orc = Oracle()
n, q, count = 1024, 4611686018326724609, 40
rng = np.random.default_rng(11)
ballots = rng.integers(0, q, size=(count, 1, 2, n), dtype=np.uint64)
recs = [fheb200.serialize_ballot(ballots[i], q, i) for i in range(count)]
for i in range(0, count, 3):
    b = bytearray(recs[i]); b[len(b) - 1 - i % 5] ^= 1 << (i % 8); recs[i] = bytes(b)
cts, status, _ = fheb200.ingest_ballots(b"".join(recs), count, 1, n, q)
want = [orc.ballot_parse(r, 1, n, q)[0] for r in recs]
assert list(status) == want, (list(status), want)
assert 0 in want and 3 in want
print("ok")
'''
    env = dict(os.environ, FHEB_WIRE_WARMUP="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("n,q,ch,count", [(1024, QT, 1, 500), (64, 132120577, 2, 333), (4096, Q62, 1, 70)])
def test_tally_straight_from_the_wire(fhe, oracle, n, q, ch, count):
    """fheb_tally_wire == deserialize_ballot per record + tally_votes per choice over the accepted ones."""
    import torch

    rng = np.random.default_rng(n + count)
    ballots = rng.integers(0, q, size=(count, ch, 2, n), dtype=np.uint64)
    ballots[:3, 0, 0, :3] = [q, q + 7, 2**64 - 1]  # unreduced words: add reduces them, a lone ballot keeps them
    recs = [fhe.serialize_ballot(ballots[i], q, i) for i in range(count)]
    for k, i in enumerate(int(x) for x in rng.choice(count, size=count // 5, replace=False)):
        b = bytearray(recs[i])
        if k % 2:
            b[45 + k % 4] ^= 0x21              # checksum field: rejected
        else:
            b[int(rng.integers(49, len(b)))] ^= 4  # payload: usually accepted by the reference's CRC, damaged words and all
        recs[i] = bytes(b)
    offs = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.uint64)
    parsed = np.zeros_like(ballots)
    want = np.zeros(count, np.uint8)
    for i in range(count):
        want[i], parsed[i], _ = oracle.ballot_parse(recs[i], ch, n, q)
    good = np.flatnonzero(want == 0)
    assert 0 < len(good) < count
    exp = np.stack([oracle.tally(parsed[good][:, c], q) for c in range(ch)])
    blob = b"".join(recs)
    got, status = fhe.tally_wire(blob, count, ch, n, q)                       # host wire, self-describing records
    eq(status, want)
    eq(got, exp)
    got2, status2 = fhe.tally_wire(blob, count, ch, n, q, offsets=offs, device="cuda")  # explicit extents, device result
    eq(status2, want)
    eq(host(got2), exp)
    dwire = torch.from_numpy(np.frombuffer(blob + b"\0" * ((-len(blob)) % 8), np.uint8).copy()).cuda()[: len(blob)]
    got3, status3 = fhe.tally_wire(dwire, count, ch, n, q, offsets=offs)      # wire bytes already in HBM
    eq(status3, want)
    eq(host(got3), exp)
    # the same numbers through ingest + tally_votes
    cts, _, _ = fhe.ingest_ballots(blob, count, ch, n, q, device="cuda")
    for c in range(ch):
        eq(host(fhe.tally_votes(cts[:, c].contiguous(), n, q)), exp[c])
    # one accepted record: returned untouched (unreduced words included); none: the reference's error
    li = int(good[0])
    bad3 = [bytes(bytearray(recs[i][:45]) + bytes([recs[i][45] ^ 1]) + bytearray(recs[i][46:])) for i in good[1:4]]
    lone = [recs[li]] + bad3
    got4, status4 = fhe.tally_wire(b"".join(lone), 4, ch, n, q)
    eq(status4, np.array([0, 3, 3, 3], np.uint8))
    eq(got4, parsed[li])
    with pytest.raises(fhe.FheError) as e:
        fhe.tally_wire(b"".join(lone[1:]), 3, ch, n, q)
    assert "Cannot add empty vector of ciphertexts" in str(e.value)
