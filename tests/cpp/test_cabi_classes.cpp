// C++ parity tests through include/fheb200.hpp (the header-only mirror of the reference's classes
// over the C ABI), written the way the reference's own stand-alone tests are
// (cpp/tests/test_ntt_processor.cpp, test_polynomial_ring.cpp, test_multi_limb.cpp: seeded
// mt19937_64 inputs, round trips, ring axioms) plus word-for-word comparison with the CPU oracle
// (oracle/fhe_oracle.c, test infrastructure).  Needs a B200; run by tests/test_gpu_cpp.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/fheb200.hpp"
#include "../../oracle/fhe_oracle.h"

static int failures = 0;
// the test is compiled with g++ (no CUDA headers): the device count comes through the C ABI
static void cudaGetDeviceCount_shim(int* n) { *n = fheb_device_count(); }
#define EXPECT(cond, ...)                        \
    do {                                         \
        if (!(cond)) {                           \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);            \
            std::printf("\n");                   \
            ++failures;                          \
        }                                        \
    } while (0)

struct TestRandom {  // cpp/tests/test_harness.h:29-47
    std::mt19937_64 rng;
    explicit TestRandom(uint64_t seed) : rng(seed) {}
    uint64_t next_coefficient(uint64_t q) { return rng() % q; }
};

static void test_ntt_processor() {
    const struct { uint32_t n; uint64_t q; int iters; } cfgs[] = {{8, 17, 100}, {16, 97, 100}, {1024, 132120577ULL, 20}};  // :203-207
    for (const auto& c : cfgs) {
        fheb200::NTTProcessor ntt(c.n, c.q);
        std::vector<uint64_t> fwd(c.n), inv(c.n);
        uint64_t sc[3];
        EXPECT(orc_precompute_twiddles(c.n, c.q, fwd.data(), inv.data(), sc) == 0, "oracle twiddles");
        const fheb200::TwiddleFactors t = ntt.get_twiddles();
        EXPECT(t.forward == fwd && t.inverse == inv && t.primitive_root == sc[0] && t.inv_n == sc[2], "twiddle tables n=%u", c.n);
        for (uint64_t seed : {42ULL, 123ULL}) {
            TestRandom r(seed);
            for (int it = 0; it < c.iters; ++it) {
                std::vector<uint64_t> x(c.n), y, ref;
                for (auto& v : x) v = r.next_coefficient(c.q);
                y = x;
                ntt.forward_ntt(y.data(), y.size());
                ref = x;
                orc_forward_ntt(ref.data(), c.n, c.q, fwd.data());
                EXPECT(y == ref, "forward_ntt n=%u seed=%llu iter=%d", c.n, (unsigned long long)seed, it);
                ntt.inverse_ntt(y.data(), y.size());
                EXPECT(y == x, "round trip n=%u", c.n);  // Property 1
            }
        }
        bool threw = false;
        try {
            std::vector<uint64_t> bad(c.n + 1);
            ntt.forward_ntt(bad.data(), bad.size());
        } catch (const std::invalid_argument&) { threw = true; }
        EXPECT(threw, "size mismatch must throw std::invalid_argument");
    }
    for (auto bad : {std::pair<uint32_t, uint64_t>{1000, 17}, {1024, 132120578ULL}, {1024, 1099511627775ULL}}) {
        bool threw = false;
        try { fheb200::NTTProcessor p(bad.first, bad.second); } catch (const std::invalid_argument&) { threw = true; }
        EXPECT(threw, "constructor must reject degree=%u q=%llu", bad.first, (unsigned long long)bad.second);
    }
}

static void test_polynomial_ring() {
    const uint32_t n = 4096;
    const uint64_t q = 4611686018326724609ULL;
    const size_t batch = 6;
    fheb200::PolynomialRing ring(n, q);
    std::vector<uint64_t> fwd(n), inv(n);
    uint64_t sc[3];
    orc_precompute_twiddles(n, q, fwd.data(), inv.data(), sc);
    TestRandom r(7);
    std::vector<uint64_t> a(batch * n), b(batch * n), c(batch * n), got(batch * n), ref(batch * n), t1(batch * n), t2(batch * n);
    for (auto& v : a) v = r.next_coefficient(q);
    for (auto& v : b) v = r.next_coefficient(q);
    for (auto& v : c) v = r.next_coefficient(q);
    ring.multiply(a.data(), b.data(), got.data(), batch);
    for (size_t i = 0; i < batch; ++i) orc_poly_multiply(&a[i * n], &b[i * n], &ref[i * n], n, q, fwd.data(), inv.data(), sc[2]);
    EXPECT(got == ref, "PolynomialRing::multiply vs oracle");
    ring.multiply(b.data(), a.data(), t1.data(), batch);
    EXPECT(t1 == got, "commutativity");
    // distributivity: a * (b + c) == a*b + a*c
    ring.add(b.data(), c.data(), t1.data(), batch);
    ring.multiply(a.data(), t1.data(), t1.data() /* aliased output */, batch);
    ring.multiply(a.data(), c.data(), t2.data(), batch);
    ring.add(got.data(), t2.data(), t2.data(), batch);
    EXPECT(t1 == t2, "distributivity");
    ring.subtract(a.data(), a.data(), t1.data(), batch);
    EXPECT(t1 == std::vector<uint64_t>(batch * n, 0), "a - a == 0");
    ring.negate(a.data(), t1.data(), batch);
    ring.add(a.data(), t1.data(), t1.data(), batch);
    EXPECT(t1 == std::vector<uint64_t>(batch * n, 0), "a + (-a) == 0");
    // tally + tensor product
    std::vector<uint64_t> tally(2 * n), tally_ref(2 * n);
    fheb200::tally_votes(a.data(), batch / 2, n, q, tally.data());
    EXPECT(orc_tally_linear(a.data(), batch / 2, n, q, tally_ref.data()) == 0 && tally == tally_ref, "tally_votes vs oracle");
    std::vector<uint64_t> ten(3 * n), ten_ref(3 * n);
    ring.tensor_multiply(a.data(), b.data(), ten.data(), 1);
    orc_tensor_multiply(a.data(), b.data(), ten_ref.data(), n, q, fwd.data(), inv.data(), sc[2]);
    EXPECT(ten == ten_ref, "tensor product vs oracle");
    // multiply -> relinearize (EncryptionEngine::relinearize, encryption.cpp:904-993) with a pre-transformed device key
    const uint32_t key_count = 3, base_log = 20, level = 3;
    std::vector<uint64_t> keys((size_t)key_count * 2 * n), rel(2 * n), rel_ref(2 * n);
    for (auto& v : keys) v = r.next_coefficient(q);
    fheb200::RelinearizationKey rk(ring, keys.data(), key_count, base_log, level, 99);
    EXPECT(rk.levels() == 3, "relinearisation levels");
    rk.relinearize(ten.data(), rel.data(), 1);
    orc_relinearize(ten.data(), keys.data(), key_count, base_log, level, rel_ref.data(), n, q, fwd.data(), inv.data(), sc[2]);
    EXPECT(rel == rel_ref, "relinearize vs oracle");
    // wire format: a record written here parses with the restated deserialize_ballot and survives the device ingest
    std::vector<uint8_t> rec(fheb_ballot_wire_size(1, n));
    size_t written = 0;
    EXPECT(fheb_ballot_serialize(a.data(), 1, n, q, 1234, rec.data(), rec.size(), &written) == 0 && written == rec.size(), "ballot serialize");
    std::vector<uint64_t> parsed(2 * n), ingested(2 * n);
    uint64_t stamp = 0, stamps[1] = {0};
    EXPECT(orc_ballot_parse(rec.data(), rec.size(), 1, n, q, parsed.data(), &stamp) == 0 && stamp == 1234, "oracle parses the record");
    EXPECT(std::memcmp(parsed.data(), a.data(), 2 * n * 8) == 0, "record payload");
    uint8_t status[1] = {9};
    size_t accepted = 0;
    EXPECT(fheb_ballots_ingest(rec.data(), rec.size(), nullptr, 1, 1, n, q, ingested.data(), status, stamps, &accepted, nullptr) == 0, "ingest");
    EXPECT(status[0] == FHEB_WIRE_OK && accepted == 1 && stamps[0] == 1234 && ingested == parsed, "device ingest vs oracle");
}

static void test_multi_limb() {
    const std::vector<uint64_t> q = {0xFFFFFFFFFFFFFF43ULL, 1};  // cpp/tests/test_multi_limb.cpp:143
    fheb200::MultiLimbModularArithmetic ml(q);
    std::vector<uint64_t> consts(5);
    orc_mlimb_constants(q.data(), 2, consts.data());
    EXPECT(ml.q_inv() == consts[0] && ml.r_mod_q()[0] == consts[1] && ml.r2_mod_q()[0] == consts[3], "multi-limb constants");
    const size_t count = 65536;  // BASELINE C3
    TestRandom r(11);
    std::vector<uint64_t> a(count * 2), b(count * 2), got(count * 2), ref(count * 2);
    for (size_t i = 0; i < count; ++i) { a[2 * i] = r.rng(); a[2 * i + 1] = r.rng() & 1; b[2 * i] = r.rng(); b[2 * i + 1] = 0; }
    ml.montgomery_mul_neon(a.data(), b.data(), got.data(), count);
    orc_mlimb_montmul(a.data(), b.data(), ref.data(), count, 2, q.data(), consts[0]);
    EXPECT(got == ref, "montgomery_mul vs oracle");
    ml.mod_add_neon(a.data(), b.data(), got.data(), count);
    orc_mlimb_add(a.data(), b.data(), ref.data(), count, 2, q.data());
    EXPECT(got == ref, "mod_add vs oracle");
    ml.mod_sub_neon(a.data(), b.data(), got.data(), count);
    orc_mlimb_sub(a.data(), b.data(), ref.data(), count, 2, q.data());
    EXPECT(got == ref, "mod_sub vs oracle");
}

static void test_bootstrap_engine() {
    const uint32_t N = 256, n = 5, k = 1, base_log = 6, level = 3;
    const uint64_t q = 1099511678977ULL;
    std::vector<uint64_t> fwd(N), inv(N);
    uint64_t sc[3];
    orc_precompute_twiddles(N, q, fwd.data(), inv.data(), sc);
    const orc_boot_params p{N, k, n, base_log, level, q, 4, fwd.data(), inv.data(), sc[2]};
    TestRandom r(3);
    std::vector<uint64_t> bsk((size_t)n * (k + 1) * level * (k + 1) * N);
    for (auto& v : bsk) v = r.next_coefficient(q);
    fheb200::BootstrapEngine eng(N, q, n, k, base_log, level, bsk.data());
    std::vector<uint64_t> tp = eng.get_default_test_poly(4), tp_ref(N);
    orc_default_test_poly(&p, tp_ref.data());
    EXPECT(tp == tp_ref, "default test polynomial");
    const size_t batch = 3;
    std::vector<uint64_t> lwe(batch * (n + 1)), got(batch * (k * N + 1)), ref(batch * (k * N + 1));
    for (auto& v : lwe) v = r.next_coefficient(q);
    eng.bootstrap_with_test_poly(lwe.data(), tp.data(), got.data(), batch);
    for (size_t i = 0; i < batch; ++i)
        orc_bootstrap(&p, &lwe[i * (n + 1)], bsk.data(), tp.data(), nullptr, 0, 0, 0, &ref[i * (k * N + 1)]);
    EXPECT(got == ref, "bootstrap (blind rotation + sample extraction) vs oracle");
}

// PolynomialRing::multiply's is_ntt short-circuit (polynomial_ring.cpp:421-447) and the tally's noise-budget metadata
static void test_is_ntt_and_noise_budget() {
    const uint32_t n = 1024;
    const uint64_t q = 132120577ULL;
    fheb200::PolynomialRing ring(n, q);
    std::vector<uint64_t> fwd(n), inv(n);
    uint64_t sc[3];
    orc_precompute_twiddles(n, q, fwd.data(), inv.data(), sc);
    TestRandom r(21);
    std::vector<uint64_t> a(n), b(n), ta(n), tb(n), got(n), ref(n);
    for (auto& v : a) v = r.next_coefficient(q);
    for (auto& v : b) v = r.next_coefficient(q);
    ring.to_ntt(a.data(), ta.data());
    ring.to_ntt(b.data(), tb.data());
    orc_poly_multiply(a.data(), b.data(), ref.data(), n, q, fwd.data(), inv.data(), sc[2]);
    EXPECT(ring.multiply(a.data(), false, b.data(), false, got.data()) == false && got == ref, "coefficient x coefficient");
    EXPECT(ring.multiply(ta.data(), true, b.data(), false, got.data()) == false && got == ref, "transform x coefficient");
    EXPECT(ring.multiply(a.data(), false, tb.data(), true, got.data()) == false && got == ref, "coefficient x transform");
    std::vector<uint64_t> pw(n);
    for (uint32_t i = 0; i < n; ++i) pw[i] = (uint64_t)(((unsigned __int128)ta[i] * tb[i]) % q);
    EXPECT(ring.multiply(ta.data(), true, tb.data(), true, got.data()) == true && got == pw, "both is_ntt: pointwise only, stays in transform form");

    // noise budgets: the reference's three bookkeeping rules, restated
    std::vector<double> bud = {30.0, 28.5, 31.0, 29.25, 27.0, 33.0, 30.5};
    double lo = bud[0];
    for (double v : bud) lo = std::min(lo, v);
    EXPECT(fheb200::tally_noise_budget(bud, fheb200::TallyVariant::BatchAdd) == lo - std::log2((double)bud.size()), "batch_add budget");
    std::vector<double> lvl = bud, nxt;
    while (lvl.size() > 1) {
        nxt.clear();
        for (size_t i = 0; i + 1 < lvl.size(); i += 2) nxt.push_back(std::min(lvl[i], lvl[i + 1]) - 1.0);
        if (lvl.size() % 2) nxt.push_back(lvl.back());
        lvl = nxt;
    }
    EXPECT(fheb200::tally_noise_budget(bud, fheb200::TallyVariant::BatchAddTree) == lvl[0], "batch_add_tree budget");
    EXPECT(fheb200::tally_noise_budget({12.5}) == 12.5, "single ballot keeps its budget");
    std::vector<uint64_t> cts(3 * 2 * n);
    for (auto& v : cts) v = r.next_coefficient(q);
    bool threw = false;
    try { fheb200::tally_votes(cts.data(), 3, n, q, {1, 2, 3}, {7, 7, 8}); } catch (const std::invalid_argument&) { threw = true; }
    EXPECT(threw, "ballots under different keys must be rejected");
    const fheb200::TallyResult tr = fheb200::tally_votes(cts.data(), 3, n, q, {20, 21, 22}, {7, 7, 7});
    std::vector<uint64_t> tref(2 * n);
    orc_tally_linear(cts.data(), 3, n, q, tref.data());
    EXPECT(tr.words == tref && tr.key_id == 7 && tr.noise_budget == 18.0, "tally_votes with metadata");  // level 1: min(20,21)-1 = 19, 22 carried; level 2: min(19,22)-1 = 18
}

// ONE process driving every visible GPU through the C ABI (no torch, no IPC): the sharded tally over peer memory, and
// host batches of transforms / products / bootstraps spread over the devices.  With one GPU the same code paths run
// with a group of one.
static void test_single_process_multi_gpu() {
    int ndev = 0;
    cudaGetDeviceCount_shim(&ndev);
    if (ndev > 8) ndev = 8;
    std::printf("single-process multi-GPU test on %d device(s)\n", ndev);
    const uint32_t n = 1024;
    const uint64_t q = 1099511678977ULL;
    TestRandom r(5);
    // ---- sharded tally, shards resident on their devices
    {
        fheb200::ShardedTallyGroup grp(n, q, (uint32_t)ndev);
        const size_t total = 97 * (size_t)ndev + 5;
        std::vector<uint64_t> cts(total * 2 * n), ref(2 * n), got(2 * n);
        for (auto& v : cts) v = r.next_coefficient(q);
        cts[3] = q + 9;  // unreduced word
        orc_tally_linear(cts.data(), total, n, q, ref.data());
        std::vector<const uint64_t*> shards(ndev);
        std::vector<size_t> counts(ndev);
        std::vector<void*> bufs(ndev, nullptr);
        size_t first = 0;
        for (int d = 0; d < ndev; ++d) {
            counts[d] = total / ndev + ((size_t)d < total % ndev ? 1 : 0);
            fheb200::initialize(d);  // makes device d current
            fheb200::check(fheb_device_alloc(&bufs[d], counts[d] * 2 * n * 8));
            fheb200::check(fheb_copy(bufs[d], cts.data() + first * 2 * n, counts[d] * 2 * n * 8, nullptr));
            shards[d] = static_cast<const uint64_t*>(bufs[d]);
            first += counts[d];
        }
        fheb200::initialize(0);
        for (int rep = 0; rep < 3; ++rep) {  // both inbox parities
            std::fill(got.begin(), got.end(), 0);
            grp.tally_votes(shards, counts, got.data());
            EXPECT(got == ref, "fheb_tally_sharded over %d device(s), call %d", ndev, rep);
        }
        for (int d = 0; d < ndev; ++d) fheb_device_free(bufs[d]);
    }
    // ---- host batches spread over the devices
    EXPECT(fheb200::set_devices() == ndev, "set_devices(all)");
    {
        const uint32_t N = 4096;
        const uint64_t Q = 4611686018326724609ULL;
        const size_t batch = 64 * (size_t)ndev + 3;  // 2 MB+ per device, ragged
        fheb200::PolynomialRing ring(N, Q);
        std::vector<uint64_t> fwd(N), inv(N);
        uint64_t sc[3];
        orc_precompute_twiddles(N, Q, fwd.data(), inv.data(), sc);
        std::vector<uint64_t> a(batch * N), b(batch * N), t(batch * N), c(batch * N);
        for (auto& v : a) v = r.next_coefficient(Q);
        for (auto& v : b) v = r.next_coefficient(Q);
        ring.to_ntt(a.data(), t.data(), batch);
        ring.multiply(a.data(), b.data(), c.data(), batch);
        for (size_t i : {(size_t)0, batch / 2, batch - 1}) {  // one polynomial of the first, a middle and the last share
            std::vector<uint64_t> x(a.begin() + i * N, a.begin() + (i + 1) * N), y(N);
            orc_forward_ntt(x.data(), N, Q, fwd.data());
            EXPECT(std::memcmp(x.data(), &t[i * N], N * 8) == 0, "spread forward transform, polynomial %zu", i);
            orc_poly_multiply(&a[i * N], &b[i * N], y.data(), N, Q, fwd.data(), inv.data(), sc[2]);
            EXPECT(std::memcmp(y.data(), &c[i * N], N * 8) == 0, "spread product, polynomial %zu", i);
        }
        ring.from_ntt(t.data(), t.data(), batch);
        EXPECT(t == a, "spread round trip");
        // host ballots spread over the devices
        std::vector<uint64_t> tal(2 * N), tal_ref(2 * N);
        fheb200::tally_votes(a.data(), batch / 2, N, Q, tal.data());
        orc_tally_linear(a.data(), batch / 2, N, Q, tal_ref.data());
        EXPECT(tal == tal_ref, "spread host tally");
    }
    {
        const uint32_t N = 256, nl = 4, k = 1, base_log = 6, level = 3;
        std::vector<uint64_t> fwd(N), inv(N);
        uint64_t sc[3];
        orc_precompute_twiddles(N, q, fwd.data(), inv.data(), sc);
        const orc_boot_params p{N, k, nl, base_log, level, q, 4, fwd.data(), inv.data(), sc[2]};
        std::vector<uint64_t> bsk((size_t)nl * (k + 1) * level * (k + 1) * N);
        for (auto& v : bsk) v = r.next_coefficient(q);
        fheb200::BootstrapEngine eng(N, q, nl, k, base_log, level, bsk.data());
        const std::vector<uint64_t> tp = eng.get_default_test_poly(4);
        const size_t batch = 256 * (size_t)ndev + 1;
        std::vector<uint64_t> lwe(batch * (nl + 1)), got(batch * (k * N + 1)), ref(k * N + 1);
        for (auto& v : lwe) v = r.next_coefficient(q);
        eng.bootstrap_with_test_poly(lwe.data(), tp.data(), got.data(), batch);
        for (size_t i : {(size_t)0, batch / 2, batch - 1}) {
            orc_bootstrap(&p, &lwe[i * (nl + 1)], bsk.data(), tp.data(), nullptr, 0, 0, 0, ref.data());
            EXPECT(std::memcmp(ref.data(), &got[i * (k * N + 1)], (k * N + 1) * 8) == 0, "spread bootstrap, ciphertext %zu", i);
        }
    }
    fheb200::check(fheb_set_devices(nullptr, 0));
}

int main() {
    try {
        fheb200::initialize();
        test_ntt_processor();
        test_polynomial_ring();
        test_multi_limb();
        test_bootstrap_engine();
        test_is_ntt_and_noise_budget();
        test_single_process_multi_gpu();
    } catch (const std::exception& e) {
        std::printf("FAIL: exception: %s\n", e.what());
        return 2;
    }
    if (failures == 0) std::printf("CPP CABI TESTS OK\n");
    return failures ? 1 : 0;
}
