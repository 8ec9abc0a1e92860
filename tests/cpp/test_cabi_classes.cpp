// C++ parity tests through include/fheb200.hpp (the header-only mirror of the reference's classes
// over the C ABI), written the way the reference's own stand-alone tests are
// (cpp/tests/test_ntt_processor.cpp, test_polynomial_ring.cpp, test_multi_limb.cpp: seeded
// mt19937_64 inputs, round trips, ring axioms) plus word-for-word comparison with the CPU oracle
// (oracle/fhe_oracle.c, test infrastructure).  Needs a B200; run by tests/test_gpu_cpp.py.
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/fheb200.hpp"
#include "../../oracle/fhe_oracle.h"

static int failures = 0;
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

int main() {
    try {
        fheb200::initialize();
        test_ntt_processor();
        test_polynomial_ring();
        test_multi_limb();
        test_bootstrap_engine();
    } catch (const std::exception& e) {
        std::printf("FAIL: exception: %s\n", e.what());
        return 2;
    }
    if (failures == 0) std::printf("CPP CABI TESTS OK\n");
    return failures ? 1 : 0;
}
