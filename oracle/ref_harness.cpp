/*
 * TEST INFRASTRUCTURE ONLY.  Not part of the product; never linked into it.
 *
 * extern "C" wrappers around the REFERENCE'S OWN C++ classes, compiled
 * together with the reference sources where they lie under /root/reference
 * (recipe: oracle/build_ref.sh) into oracle/_ref/libref_oracle.so.
 *
 * Purpose: (1) pin the plain-C restatement in oracle/fhe_oracle.c against the
 * reference's real code on identical inputs, (2) generate the golden vectors
 * under tests/golden/, (3) serve as the timed CPU baseline
 * (`cpu_baseline.kind == "reference"`).
 *
 * Nothing here re-implements arithmetic: every function forwards to a
 * reference method; the file:line of each is given next to the wrapper.
 * Flat little-endian u64 buffers in, flat buffers out.  All functions return
 * 0 on success, -1 when the reference threw (message via ref_last_error()).
 */
#define private public  /* test-only: reach decompose_polynomial / rotate_polynomial */
#include "bootstrap_engine.h"
#undef private
#include "adaptive_dispatcher.h"
#include "key_manager.h"
#include "key_serializer.h"
#include "modular_arithmetic.h"
#include "ntt_processor.h"
#include "parameter_set.h"
#include "polynomial_ring.h"

#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace fhe_accelerate;

static thread_local std::string g_err;

#define REF_TRY try {
#define REF_CATCH                                   \
    }                                               \
    catch (const std::exception& e) {               \
        g_err = e.what();                           \
        return -1;                                  \
    }                                               \
    catch (...) {                                   \
        g_err = "unknown exception";                \
        return -1;                                  \
    }                                               \
    return 0;

template <class F>
static void parallel_for(size_t count, int threads, F&& fn) {
    if (threads <= 1 || count <= 1) {
        for (size_t i = 0; i < count; ++i) fn(i, 0);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= count) break;
                fn(i, t);
            }
        });
    }
    for (auto& th : pool) th.join();
}

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

/* ---------------------------------------------------------------- NTT ---- */
/* NTTProcessor ctor: reference cpp/src/ntt_processor.cpp:134-160            */
int ref_ntt_create(uint32_t degree, uint64_t modulus, void** out) {
    REF_TRY
    *out = new NTTProcessor(degree, modulus);
    REF_CATCH
}
void ref_ntt_destroy(void* h) { delete static_cast<NTTProcessor*>(h); }

/* get_twiddles(): reference cpp/include/ntt_processor.h:84; tables built at
 * cpp/src/ntt_processor.cpp:168-208                                         */
int ref_ntt_tables(void* h, uint64_t* fwd, uint64_t* inv, uint64_t* scalars /*[3]: psi, psi_inv, inv_n*/) {
    REF_TRY
    const TwiddleFactors& t = static_cast<NTTProcessor*>(h)->get_twiddles();
    std::memcpy(fwd, t.forward.data(), t.forward.size() * 8);
    std::memcpy(inv, t.inverse.data(), t.inverse.size() * 8);
    scalars[0] = t.primitive_root;
    scalars[1] = t.inv_primitive_root;
    scalars[2] = t.inv_n;
    REF_CATCH
}

/* forward_ntt / inverse_ntt in-place: cpp/src/ntt_processor.cpp:262-311,
 * 325-380; the serial batch loop mirrors :394-408.  `threads` > 1 runs the
 * independent polynomials on several host threads (CPU baseline only).      */
int ref_ntt_forward(void* h, uint64_t* coeffs, size_t batch, int threads) {
    REF_TRY
    auto* p = static_cast<NTTProcessor*>(h);
    size_t n = p->get_degree();
    parallel_for(batch, threads, [&](size_t i, int) { p->forward_ntt(coeffs + i * n, n); });
    REF_CATCH
}
int ref_ntt_inverse(void* h, uint64_t* coeffs, size_t batch, int threads) {
    REF_TRY
    auto* p = static_cast<NTTProcessor*>(h);
    size_t n = p->get_degree();
    parallel_for(batch, threads, [&](size_t i, int) { p->inverse_ntt(coeffs + i * n, n); });
    REF_CATCH
}

/* fast_ntt_forward / fast_ntt_inverse / fast_modmul_batch:
 * cpp/src/adaptive_dispatcher.cpp:57-80,135-205 (caller-supplied table)     */
int ref_fast_ntt_forward(uint64_t* coeffs, size_t n, uint64_t q, const uint64_t* tw, size_t batch) {
    REF_TRY
    for (size_t i = 0; i < batch; ++i) fast_ntt_forward(coeffs + i * n, n, q, tw);
    REF_CATCH
}
int ref_fast_ntt_inverse(uint64_t* coeffs, size_t n, uint64_t q, const uint64_t* inv_tw, size_t batch) {
    REF_TRY
    for (size_t i = 0; i < batch; ++i) fast_ntt_inverse(coeffs + i * n, n, q, inv_tw);
    REF_CATCH
}
int ref_fast_modmul_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n, uint64_t q) {
    REF_TRY
    fast_modmul_batch(a, b, r, n, q);
    REF_CATCH
}

/* ------------------------------------------------------- PolynomialRing -- */
int ref_ring_create(uint32_t degree, uint64_t modulus, void** out) {
    REF_TRY
    *out = new PolynomialRing(degree, modulus);
    REF_CATCH
}
void ref_ring_destroy(void* h) { delete static_cast<PolynomialRing*>(h); }

static Polynomial make_poly(const uint64_t* src, uint32_t n, uint64_t q, bool ntt) {
    return Polynomial(std::vector<uint64_t>(src, src + n), q, ntt);
}

/* op: 0 add (polynomial_ring.cpp:263), 1 subtract (:311), 2 pointwise_multiply
 * (:493, operands flagged NTT), 3 multiply (:421, coefficient form),
 * 4 negate (:345; b ignored), 5 multiply_scalar (:454; scalar = b[0] of each
 * row... passed separately via `scalar`)                                    */
int ref_ring_binary(void* h, int op, const uint64_t* a, const uint64_t* b, uint64_t scalar,
                    uint64_t* r, size_t batch, int threads) {
    REF_TRY
    auto* ring = static_cast<PolynomialRing*>(h);
    uint32_t n = ring->degree();
    uint64_t q = ring->modulus();
    std::atomic<int> failed{0};
    std::string msg;
    parallel_for(batch, threads, [&](size_t i, int) {
        try {
            bool ntt = (op == 2);
            Polynomial pa = make_poly(a + i * n, n, q, ntt);
            Polynomial out(n, q);
            switch (op) {
                case 0: out = ring->add(pa, make_poly(b + i * n, n, q, ntt)); break;
                case 1: out = ring->subtract(pa, make_poly(b + i * n, n, q, ntt)); break;
                case 2: out = ring->pointwise_multiply(pa, make_poly(b + i * n, n, q, ntt)); break;
                case 3: out = ring->multiply(pa, make_poly(b + i * n, n, q, ntt)); break;
                case 4: out = ring->negate(pa); break;
                case 5: out = ring->multiply_scalar(pa, scalar); break;
                default: throw std::invalid_argument("bad op");
            }
            std::memcpy(r + i * n, out.data(), n * 8);
        } catch (const std::exception& e) {
            if (!failed.exchange(1)) msg = e.what();
        }
    });
    if (failed) throw std::runtime_error(msg);
    REF_CATCH
}

/* Tally, linear order: EncryptionEngine::batch_add, cpp/src/encryption.cpp:
 * 1327-1364, composed from the unmodified PolynomialRing::add_inplace
 * (polynomial_ring.cpp:274-309) because encryption.cpp does not compile as
 * shipped.  cts = [M][2][N]; out = [2][N].                      */
int ref_tally_linear(void* h, const uint64_t* cts, size_t m, uint64_t* out) {
    REF_TRY
    auto* ring = static_cast<PolynomialRing*>(h);
    uint32_t n = ring->degree();
    uint64_t q = ring->modulus();
    if (m == 0) throw std::invalid_argument("Cannot add empty vector of ciphertexts");
    Polynomial s0 = make_poly(cts, n, q, false);
    Polynomial s1 = make_poly(cts + n, n, q, false);
    for (size_t i = 1; i < m; ++i) {
        ring->add_inplace(s0, make_poly(cts + (2 * i) * n, n, q, false));
        ring->add_inplace(s1, make_poly(cts + (2 * i + 1) * n, n, q, false));
    }
    std::memcpy(out, s0.data(), n * 8);
    std::memcpy(out + n, s1.data(), n * 8);
    REF_CATCH
}

/* Tally, pairwise tree order: EncryptionEngine::batch_add_tree /
 * tally_votes, cpp/src/encryption.cpp:1061-1067,1366-1458 (sequential
 * branch), composed from PolynomialRing::add (polynomial_ring.cpp:263-272). */
int ref_tally_tree(void* h, const uint64_t* cts, size_t m, uint64_t* out) {
    REF_TRY
    auto* ring = static_cast<PolynomialRing*>(h);
    uint32_t n = ring->degree();
    uint64_t q = ring->modulus();
    if (m == 0) throw std::invalid_argument("Cannot add empty vector of ciphertexts");
    std::vector<std::pair<Polynomial, Polynomial>> level;
    for (size_t i = 0; i < m; ++i)
        level.emplace_back(make_poly(cts + 2 * i * n, n, q, false), make_poly(cts + (2 * i + 1) * n, n, q, false));
    while (level.size() > 1) {
        std::vector<std::pair<Polynomial, Polynomial>> next;
        for (size_t i = 0; i + 1 < level.size(); i += 2)
            next.emplace_back(ring->add(level[i].first, level[i + 1].first),
                              ring->add(level[i].second, level[i + 1].second));
        if (level.size() % 2 == 1) next.push_back(std::move(level.back()));
        level = std::move(next);
    }
    std::memcpy(out, level[0].first.data(), n * 8);
    std::memcpy(out + n, level[0].second.data(), n * 8);
    REF_CATCH
}

/* Tensor product: EncryptionEngine::multiply, cpp/src/encryption.cpp:737-798,
 * composed from PolynomialRing::{to_ntt,pointwise_multiply,add,from_ntt}.
 * ct = [2][N] each, out = [3][N], coefficient form in and out.              */
int ref_tensor_multiply(void* h, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out) {
    REF_TRY
    auto* ring = static_cast<PolynomialRing*>(h);
    uint32_t n = ring->degree();
    uint64_t q = ring->modulus();
    Polynomial a0 = make_poly(ct1, n, q, false), a1 = make_poly(ct1 + n, n, q, false);
    Polynomial b0 = make_poly(ct2, n, q, false), b1 = make_poly(ct2 + n, n, q, false);
    ring->to_ntt(a0); ring->to_ntt(a1); ring->to_ntt(b0); ring->to_ntt(b1);
    Polynomial r0 = ring->pointwise_multiply(a0, b0);
    Polynomial x = ring->pointwise_multiply(a0, b1);
    Polynomial y = ring->pointwise_multiply(a1, b0);
    Polynomial r1 = ring->add(x, y);
    Polynomial r2 = ring->pointwise_multiply(a1, b1);
    ring->from_ntt(r0); ring->from_ntt(r1); ring->from_ntt(r2);
    std::memcpy(out, r0.data(), n * 8);
    std::memcpy(out + n, r1.data(), n * 8);
    std::memcpy(out + 2 * n, r2.data(), n * 8);
    REF_CATCH
}

/* Relinearisation: EncryptionEngine::relinearize, cpp/src/encryption.cpp:904-993, composed from the same
 * unmodified PolynomialRing calls in the same order (clone, to_ntt on the digit polynomial AND on both key
 * polynomials every level, pointwise_multiply, from_ntt, add_inplace).  ct = [3][N] (c0, c1, c2);
 * keys = [key_count][2][N] ((a, b) per pair); base_log / level = the KeySwitchKey fields; out = [2][N]. */
int ref_relinearize(void* h, const uint64_t* ct, const uint64_t* keys, uint32_t key_count, uint32_t key_base_log,
                    uint32_t key_level, uint64_t* out) {
    REF_TRY
    auto* ring = static_cast<PolynomialRing*>(h);
    uint32_t degree = ring->degree();
    uint64_t modulus = ring->modulus();
    uint32_t decomp_base_log = key_base_log > 0 ? key_base_log : 4;
    uint64_t decomp_base = 1ULL << decomp_base_log;
    uint32_t num_levels = key_level > 0 ? key_level : static_cast<uint32_t>((64 + decomp_base_log - 1) / decomp_base_log);
    Polynomial result_c0 = make_poly(ct, degree, modulus, false);
    Polynomial result_c1 = make_poly(ct + degree, degree, modulus, false);
    const uint64_t* c2 = ct + 2 * (size_t)degree;
    for (uint32_t level = 0; level < num_levels && level < key_count; ++level) {
        Polynomial c2_digit(degree, modulus);
        uint64_t shift = level * decomp_base_log;
        uint64_t mask = decomp_base - 1;
        for (uint32_t i = 0; i < degree; ++i) c2_digit[i] = (c2[i] >> shift) & mask;
        Polynomial rlk_a_ntt = make_poly(keys + ((size_t)level * 2) * degree, degree, modulus, false);
        Polynomial rlk_b_ntt = make_poly(keys + ((size_t)level * 2 + 1) * degree, degree, modulus, false);
        Polynomial c2_digit_ntt = c2_digit.clone();
        ring->to_ntt(c2_digit_ntt);
        ring->to_ntt(rlk_a_ntt);
        ring->to_ntt(rlk_b_ntt);
        Polynomial prod_b = ring->pointwise_multiply(c2_digit_ntt, rlk_b_ntt);
        ring->from_ntt(prod_b);
        ring->add_inplace(result_c0, prod_b);
        Polynomial prod_a = ring->pointwise_multiply(c2_digit_ntt, rlk_a_ntt);
        ring->from_ntt(prod_a);
        ring->add_inplace(result_c1, prod_a);
    }
    std::memcpy(out, result_c0.data(), (size_t)degree * 8);
    std::memcpy(out + degree, result_c1.data(), (size_t)degree * 8);
    REF_CATCH
}

/* ------------------------------------------------------------ wire formats - */
/* KeySerializer / BallotSerializer, cpp/src/key_serializer.cpp (compiled unmodified). */
uint32_t ref_crc32(const uint8_t* data, size_t len) { return KeySerializer::compute_crc32(data, len); }

static int copy_out(const std::vector<uint8_t>& bytes, uint8_t* out, size_t cap, size_t* len) {
    *len = bytes.size();
    if (bytes.size() > cap) {
        g_err = "output buffer too small";
        return -1;
    }
    std::memcpy(out, bytes.data(), bytes.size());
    return 0;
}

/* BallotSerializer::serialize_ballot (:709-774): choices = [num_choices][2][n] */
int ref_serialize_ballot(const uint64_t* choices, uint32_t num_choices, uint32_t n, uint64_t q, uint64_t timestamp,
                         uint8_t* out, size_t cap, size_t* len) {
    REF_TRY
    std::vector<std::pair<Polynomial, Polynomial>> enc;
    for (uint32_t c = 0; c < num_choices; ++c)
        enc.emplace_back(make_poly(choices + ((size_t)c * 2) * n, n, q, false), make_poly(choices + ((size_t)c * 2 + 1) * n, n, q, false));
    BallotSerializer bs;
    std::vector<uint8_t> bytes;
    auto r = bs.serialize_ballot(enc, timestamp, bytes);
    if (!r.success) throw std::runtime_error(r.error_message);
    if (copy_out(bytes, out, cap, len) != 0) return -1;
    REF_CATCH
}

/* BallotSerializer::deserialize_ballot (:776-846): out = [max_choices][2][n]; returns -1 with the reference's
 * error message when it rejects the record */
int ref_deserialize_ballot(const uint8_t* in, size_t len, uint32_t n, uint64_t q, uint64_t* out, uint32_t max_choices,
                           uint32_t* num_choices, uint64_t* timestamp) {
    REF_TRY
    BallotSerializer bs;
    auto r = bs.deserialize_ballot(std::vector<uint8_t>(in, in + len), n, q);
    if (!r.success) throw std::runtime_error(r.error_message);
    *num_choices = (uint32_t)r.value->encrypted_choices.size();
    *timestamp = r.value->timestamp;
    for (uint32_t c = 0; c < *num_choices && c < max_choices; ++c) {
        const auto& ch = r.value->encrypted_choices[c];
        if (ch.first.degree() != n) throw std::runtime_error("choice degree differs from n");
        std::memcpy(out + ((size_t)c * 2) * n, ch.first.data(), (size_t)n * 8);
        std::memcpy(out + ((size_t)c * 2 + 1) * n, ch.second.data(), (size_t)n * 8);
    }
    REF_CATCH
}

/* KeySerializer::serialize_eval_key (:353-412): keys = [count][2][n] ((a, b) per pair) */
int ref_serialize_eval_key(const uint64_t* keys, uint32_t count, uint32_t n, uint64_t q, uint32_t base_log, uint32_t level,
                           uint64_t key_id, uint8_t* out, size_t cap, size_t* len) {
    REF_TRY
    EvaluationKey ek;
    ek.key_id = key_id;
    ek.relin_key.decomp_base_log = base_log;
    ek.relin_key.decomp_level = level;
    ek.relin_key.key_id = key_id;
    for (uint32_t i = 0; i < count; ++i)
        ek.relin_key.keys.emplace_back(make_poly(keys + ((size_t)i * 2) * n, n, q, false), make_poly(keys + ((size_t)i * 2 + 1) * n, n, q, false));
    KeySerializer ks;
    std::vector<uint8_t> bytes;
    auto r = ks.serialize_eval_key(ek, bytes);
    if (!r.success) throw std::runtime_error(r.error_message);
    if (copy_out(bytes, out, cap, len) != 0) return -1;
    REF_CATCH
}

/* KeySerializer::deserialize_eval_key (:414-466): keys_out = [max_count][2][n]; meta = {count, base_log, level} */
int ref_deserialize_eval_key(const uint8_t* in, size_t len, uint32_t n, uint64_t* keys_out, uint32_t max_count, uint32_t meta[3],
                             uint64_t* key_id) {
    REF_TRY
    KeySerializer ks;
    auto r = ks.deserialize_eval_key(std::vector<uint8_t>(in, in + len));
    if (!r.success) throw std::runtime_error(r.error_message);
    const auto& rk = r.value->relin_key;
    meta[0] = (uint32_t)rk.keys.size();
    meta[1] = rk.decomp_base_log;
    meta[2] = rk.decomp_level;
    *key_id = r.value->key_id;
    for (uint32_t i = 0; i < meta[0] && i < max_count; ++i) {
        std::memcpy(keys_out + ((size_t)i * 2) * n, rk.keys[i].first.data(), (size_t)n * 8);
        std::memcpy(keys_out + ((size_t)i * 2 + 1) * n, rk.keys[i].second.data(), (size_t)n * 8);
    }
    REF_CATCH
}

/* KeySerializer::serialize_bootstrap_key (:472-543): bsk = [n_lwe][rows][2][n]; ksk = [ksk_count][2][n] */
int ref_serialize_bootstrap_key(const uint64_t* bsk, uint32_t n_lwe, uint32_t rows, const uint64_t* ksk, uint32_t ksk_count,
                                uint32_t ksk_base_log, uint32_t ksk_level, uint32_t n, uint64_t q, uint64_t key_id,
                                uint8_t* out, size_t cap, size_t* len) {
    REF_TRY
    BootstrapKey bk;
    bk.key_id = key_id;
    bk.lwe_dimension = n_lwe;
    for (uint32_t i = 0; i < n_lwe; ++i) {
        std::vector<std::pair<Polynomial, Polynomial>> row;
        for (uint32_t j = 0; j < rows; ++j) {
            const uint64_t* p = bsk + (((size_t)i * rows + j) * 2) * n;
            row.emplace_back(make_poly(p, n, q, false), make_poly(p + n, n, q, false));
        }
        bk.bsk.push_back(std::move(row));
    }
    bk.ksk.decomp_base_log = ksk_base_log;
    bk.ksk.decomp_level = ksk_level;
    for (uint32_t i = 0; i < ksk_count; ++i)
        bk.ksk.keys.emplace_back(make_poly(ksk + ((size_t)i * 2) * n, n, q, false), make_poly(ksk + ((size_t)i * 2 + 1) * n, n, q, false));
    KeySerializer ks;
    std::vector<uint8_t> bytes;
    auto r = ks.serialize_bootstrap_key(bk, bytes);
    if (!r.success) throw std::runtime_error(r.error_message);
    if (copy_out(bytes, out, cap, len) != 0) return -1;
    REF_CATCH
}

/* ------------------------------------------------- scalar ModularArithmetic - */
/* the class behind the reference's N-API addon (src/native/lib.rs:44-120)      */
int ref_scalar_create(uint64_t modulus, void** out) {
    REF_TRY
    *out = new ModularArithmetic(modulus);
    REF_CATCH
}
void ref_scalar_destroy(void* h) { delete static_cast<ModularArithmetic*>(h); }
/* op: 0 montgomery_mul, 1 mod_add, 2 mod_sub, 3 to_montgomery, 4 from_montgomery, 5 get_modulus */
uint64_t ref_scalar_op(void* h, int op, uint64_t a, uint64_t b) {
    auto* m = static_cast<ModularArithmetic*>(h);
    switch (op) {
        case 0: return m->montgomery_mul(a, b);
        case 1: return m->mod_add(a, b);
        case 2: return m->mod_sub(a, b);
        case 3: return m->to_montgomery(a);
        case 4: return m->from_montgomery(a);
        default: return m->get_modulus();
    }
}

/* ------------------------------------------------------------ multi-limb - */
/* MultiLimbModularArithmetic: cpp/src/modular_arithmetic.cpp:471-693        */
int ref_mlimb_create(const uint64_t* q_limbs, size_t limbs, void** out) {
    REF_TRY
    *out = new MultiLimbModularArithmetic(MultiLimbInteger(std::vector<uint64_t>(q_limbs, q_limbs + limbs)));
    REF_CATCH
}
void ref_mlimb_destroy(void* h) { delete static_cast<MultiLimbModularArithmetic*>(h); }

/* constants: q_inv, then r_mod_q[limbs], then r2_mod_q[limbs]               */
int ref_mlimb_constants(void* h, uint64_t* out) {
    REF_TRY
    const auto& c = static_cast<MultiLimbModularArithmetic*>(h)->get_constants();
    size_t l = c.num_limbs;
    out[0] = c.q_inv;
    for (size_t i = 0; i < l; ++i) out[1 + i] = c.r_mod_q.get_limb(i);
    for (size_t i = 0; i < l; ++i) out[1 + l + i] = c.r2_mod_q.get_limb(i);
    REF_CATCH
}

/* op: 0 montgomery_mul (:614), 1 mod_add (:627), 2 mod_sub (:650),
 * 3 to_montgomery (:673), 4 from_montgomery (:680); element-major
 * [count][limbs] little-endian limbs                                        */
int ref_mlimb_batch(void* h, int op, const uint64_t* a, const uint64_t* b, uint64_t* r,
                    size_t count, size_t limbs, int threads) {
    REF_TRY
    auto* m = static_cast<MultiLimbModularArithmetic*>(h);
    parallel_for(count, threads, [&](size_t i, int) {
        MultiLimbInteger x(std::vector<uint64_t>(a + i * limbs, a + (i + 1) * limbs));
        MultiLimbInteger res;
        if (op == 3) res = m->to_montgomery(x);
        else if (op == 4) res = m->from_montgomery(x);
        else {
            MultiLimbInteger y(std::vector<uint64_t>(b + i * limbs, b + (i + 1) * limbs));
            res = (op == 0) ? m->montgomery_mul(x, y) : (op == 1) ? m->mod_add(x, y) : m->mod_sub(x, y);
        }
        for (size_t j = 0; j < limbs; ++j) r[i * limbs + j] = res.get_limb(j);
    });
    REF_CATCH
}

/* ------------------------------------------------------------ bootstrap -- */
struct RefBoot {
    ParameterSet params;
    std::unique_ptr<BootstrapEngine> engine;
    std::vector<std::unique_ptr<BootstrapEngine>> workers; /* one engine per host thread */
    ExtendedBootstrapKey key;
};

static ParameterSet make_params(uint32_t N, uint64_t q, uint32_t n, uint32_t k, uint32_t base_log,
                                uint32_t level, uint64_t t) {
    ParameterSet p;
    p.scheme = FHEScheme::TFHE;
    p.security = SecurityLevel::Bits128;
    p.poly_degree = N;
    p.moduli = {q};
    p.lwe_dimension = n;
    p.lwe_noise_std = 3.2e-11; /* TFHE_128_FAST value, cpp/src/parameter_set.cpp:121 */
    p.glwe_dimension = k;
    p.decomp_base_log = base_log;
    p.decomp_level = level;
    p.plaintext_modulus = t;
    return p;
}

/* BootstrapEngine ctor: cpp/src/bootstrap_engine.cpp:26-50                  */
int ref_boot_create(uint32_t N, uint64_t q, uint32_t n, uint32_t k, uint32_t base_log, uint32_t level,
                    uint64_t t, void** out) {
    REF_TRY
    auto* b = new RefBoot();
    b->params = make_params(N, q, n, k, base_log, level, t);
    b->engine = std::make_unique<BootstrapEngine>(b->params);
    b->key.lwe_dimension = n;
    b->key.glwe_dimension = k;
    b->key.poly_degree = N;
    b->key.decomp_base_log = base_log;
    b->key.decomp_level = level;
    *out = b;
    REF_CATCH
}
void ref_boot_destroy(void* h) { delete static_cast<RefBoot*>(h); }

/* get_default_test_poly(): built at cpp/src/bootstrap_engine.cpp:57-77      */
int ref_boot_default_test_poly(void* h, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    const Polynomial& p = b->engine->get_default_test_poly();
    std::memcpy(out, p.data(), p.degree() * 8);
    REF_CATCH
}

/* create_lookup_table family: cpp/src/bootstrap_engine.cpp:725-779.
 * kind: 0 identity(modulus=arg0), 1 negation(arg0), 2 threshold(arg0=thr, arg1=modulus) */
int ref_boot_lut(void* h, int kind, uint64_t arg0, uint64_t arg1, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    LookupTable lut = (kind == 0)   ? b->engine->create_identity_lut(arg0)
                      : (kind == 1) ? b->engine->create_negation_lut(arg0)
                                    : b->engine->create_threshold_lut(arg0, arg1);
    std::memcpy(out, lut.table.data(), lut.table.degree() * 8);
    REF_CATCH
}

/* Key material produced by the reference's own generators:
 * KeyManager::generate_secret_key (cpp/src/key_manager.cpp), then
 * BootstrapEngine::encrypt_ggsw per LWE key bit (bootstrap_engine.cpp:
 * 268-306) - the body of generate_bootstrap_key (:308-364) without its
 * key-switch-key step, whose Polynomial(n words) constructor throws when n is
 * not a power of two.  `with_ksk` additionally calls
 * generate_key_switch_key (:366-424).  lwe_sk_out[n] receives the binary LWE
 * key that was drawn.                                                       */
int ref_boot_keygen(void* h, int with_ksk, int64_t* lwe_sk_out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    KeyManager km(b->params);
    auto glwe_sk = km.generate_secret_key(SecretKeyDistribution::BINARY);
    SecureRandom rng;
    std::vector<int64_t> lwe_sk(b->params.lwe_dimension);
    for (auto& s : lwe_sk) s = static_cast<int64_t>(rng.sample_binary());
    b->key.bsk.clear();
    for (size_t i = 0; i < lwe_sk.size(); ++i) b->key.bsk.push_back(b->engine->encrypt_ggsw(lwe_sk[i], *glwe_sk));
    b->key.key_id = glwe_sk->key_id;
    if (with_ksk) b->key.ksk = b->engine->generate_key_switch_key(*glwe_sk, lwe_sk);
    std::memcpy(lwe_sk_out, lwe_sk.data(), lwe_sk.size() * 8);
    REF_CATCH
}

/* Flat BSK layout (row order of bootstrap_engine.cpp:281-303):
 * [n][(k+1)*L rows][k mask polys then body][N]                              */
int ref_boot_export_bsk(void* h, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint32_t N = b->params.poly_degree;
    for (const auto& g : b->key.bsk)
        for (const auto& row : g.matrix) {
            for (const auto& m : row.mask) { std::memcpy(out, m.data(), N * 8); out += N; }
            std::memcpy(out, row.body.data(), N * 8); out += N;
        }
    REF_CATCH
}
int ref_boot_import_bsk(void* h, const uint64_t* in) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint32_t N = b->params.poly_degree, k = b->params.glwe_dimension, L = b->params.decomp_level;
    uint64_t q = b->params.moduli[0];
    b->key.bsk.clear();
    for (uint32_t i = 0; i < b->params.lwe_dimension; ++i) {
        std::vector<GLWECiphertext> rows;
        for (uint32_t r = 0; r < (k + 1) * L; ++r) {
            std::vector<Polynomial> mask;
            for (uint32_t j = 0; j < k; ++j) { mask.push_back(make_poly(in, N, q, false)); in += N; }
            Polynomial body = make_poly(in, N, q, false); in += N;
            rows.emplace_back(std::move(mask), std::move(body), 0, false);
        }
        b->key.bsk.emplace_back(std::move(rows), b->params.decomp_base_log, L, 0);
    }
    REF_CATCH
}

/* Flat KSK layout: entries in generation order (coefficient-major, level-minor,
 * bootstrap_engine.cpp:391-421); each entry = n_out words of a then 1 word b */
int ref_boot_ksk_shape(void* h, uint64_t* entries, uint64_t* n_out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    *entries = b->key.ksk.keys.size();
    *n_out = b->key.ksk.keys.empty() ? 0 : b->key.ksk.keys[0].first.degree();
    REF_CATCH
}
int ref_boot_export_ksk(void* h, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    for (const auto& e : b->key.ksk.keys) {
        std::memcpy(out, e.first.data(), e.first.degree() * 8);
        out += e.first.degree();
        *out++ = e.second[0];
    }
    REF_CATCH
}
int ref_boot_import_ksk(void* h, const uint64_t* in, uint64_t entries, uint64_t n_out, uint32_t base_log,
                        uint32_t level) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint64_t q = b->params.moduli[0];
    b->key.ksk.keys.clear();
    b->key.ksk.decomp_base_log = base_log;
    b->key.ksk.decomp_level = level;
    for (uint64_t e = 0; e < entries; ++e) {
        Polynomial a = make_poly(in, static_cast<uint32_t>(n_out), q, false);
        in += n_out;
        Polynomial bb(std::vector<uint64_t>{*in++}, q, false);
        b->key.ksk.keys.emplace_back(std::move(a), std::move(bb));
    }
    REF_CATCH
}

/* encrypt_lwe: cpp/src/bootstrap_engine.cpp:785-817; out = [count][n+1] (a.., b) */
int ref_boot_encrypt_lwe(void* h, const uint64_t* values, size_t count, const int64_t* sk, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint32_t n = b->params.lwe_dimension;
    std::vector<int64_t> s(sk, sk + n);
    for (size_t i = 0; i < count; ++i) {
        LWECiphertext ct = b->engine->encrypt_lwe(values[i], s);
        std::memcpy(out + i * (n + 1), ct.a.data(), n * 8);
        out[i * (n + 1) + n] = ct.b;
    }
    REF_CATCH
}
int ref_boot_decrypt_lwe(void* h, const uint64_t* cts, size_t count, size_t dim, const int64_t* sk, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint64_t q = b->params.moduli[0];
    std::vector<int64_t> s(sk, sk + dim);
    for (size_t i = 0; i < count; ++i) {
        const uint64_t* c = cts + i * (dim + 1);
        LWECiphertext ct(std::vector<uint64_t>(c, c + dim), c[dim], q, 0);
        out[i] = b->engine->decrypt_lwe(ct, s);
    }
    REF_CATCH
}

static GLWECiphertext make_glwe(const uint64_t* src, uint32_t k, uint32_t N, uint64_t q) {
    std::vector<Polynomial> mask;
    for (uint32_t j = 0; j < k; ++j) mask.push_back(make_poly(src + j * N, N, q, false));
    return GLWECiphertext(std::move(mask), make_poly(src + k * N, N, q, false), 0, false);
}
static void store_glwe(const GLWECiphertext& g, uint64_t* dst) {
    uint32_t N = g.poly_degree();
    for (const auto& m : g.mask) { std::memcpy(dst, m.data(), N * 8); dst += N; }
    std::memcpy(dst, g.body.data(), N * 8);
}

/* decompose_polynomial: bootstrap_engine.cpp:152-185; out = [level][N]      */
int ref_boot_decompose(void* h, const uint64_t* poly, uint32_t base_log, uint32_t level, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint32_t N = b->params.poly_degree;
    auto d = b->engine->decompose_polynomial(make_poly(poly, N, b->params.moduli[0], false), base_log, level);
    for (uint32_t l = 0; l < level; ++l) std::memcpy(out + l * N, d[l].data(), N * 8);
    REF_CATCH
}
/* rotate_polynomial: bootstrap_engine.cpp:122-145                           */
int ref_boot_rotate(void* h, const uint64_t* poly, int32_t rotation, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint32_t N = b->params.poly_degree;
    Polynomial r = b->engine->rotate_polynomial(make_poly(poly, N, b->params.moduli[0], false), rotation);
    std::memcpy(out, r.data(), N * 8);
    REF_CATCH
}
/* external_product(glwe, bsk[index]): bootstrap_engine.cpp:431-518; glwe = [k+1][N] */
int ref_boot_external_product(void* h, const uint64_t* glwe, uint32_t index, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    GLWECiphertext g = make_glwe(glwe, b->params.glwe_dimension, b->params.poly_degree, b->params.moduli[0]);
    store_glwe(b->engine->external_product(g, b->key.bsk.at(index)), out);
    REF_CATCH
}
/* cmux(bsk[index], ct0, ct1): bootstrap_engine.cpp:520-540                  */
int ref_boot_cmux(void* h, uint32_t index, const uint64_t* ct0, const uint64_t* ct1, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint32_t k = b->params.glwe_dimension, N = b->params.poly_degree;
    uint64_t q = b->params.moduli[0];
    store_glwe(b->engine->cmux(b->key.bsk.at(index), make_glwe(ct0, k, N, q), make_glwe(ct1, k, N, q)), out);
    REF_CATCH
}

static void ensure_workers(RefBoot* b, int threads) {
    while (static_cast<int>(b->workers.size()) < threads)
        b->workers.push_back(std::make_unique<BootstrapEngine>(b->params));
}

/* blind_rotate: bootstrap_engine.cpp:547-577, on acc = (0,..,0,test_poly) as
 * bootstrap_with_test_poly builds it (:692-699).  lwe = [count][n+1];
 * out = [count][k+1][N].  One engine per host thread.                       */
int ref_boot_blind_rotate(void* h, const uint64_t* lwe, size_t count, const uint64_t* test_poly, uint64_t* out,
                          int threads) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint32_t k = b->params.glwe_dimension, N = b->params.poly_degree, n = b->params.lwe_dimension;
    uint64_t q = b->params.moduli[0];
    if (threads < 1) threads = 1;
    ensure_workers(b, threads);
    parallel_for(count, threads, [&](size_t i, int t) {
        const uint64_t* c = lwe + i * (n + 1);
        LWECiphertext ct(std::vector<uint64_t>(c, c + n), c[n], q, 0);
        GLWECiphertext acc(k, N, q, 0);
        for (auto& m : acc.mask) m.set_zero();
        acc.body = make_poly(test_poly, N, q, false);
        b->workers[t]->blind_rotate(acc, ct, b->key);
        store_glwe(acc, out + i * (k + 1) * N);
    });
    REF_CATCH
}
/* sample_extract: bootstrap_engine.cpp:594-624; out = [count][k*N+1]        */
int ref_boot_sample_extract(void* h, const uint64_t* glwe, size_t count, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint32_t k = b->params.glwe_dimension, N = b->params.poly_degree;
    uint64_t q = b->params.moduli[0];
    for (size_t i = 0; i < count; ++i) {
        LWECiphertext e = b->engine->sample_extract(make_glwe(glwe + i * (k + 1) * N, k, N, q));
        std::memcpy(out + i * (k * N + 1), e.a.data(), k * N * 8);
        out[i * (k * N + 1) + k * N] = e.b;
    }
    REF_CATCH
}
/* key_switch: bootstrap_engine.cpp:626-669; in = [count][dim_in+1], out = [count][n_out+1] */
int ref_boot_key_switch(void* h, const uint64_t* lwe, size_t count, size_t dim_in, uint64_t* out) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint64_t q = b->params.moduli[0];
    size_t n_out = b->key.ksk.keys.empty() ? 0 : b->key.ksk.keys[0].first.degree();
    for (size_t i = 0; i < count; ++i) {
        const uint64_t* c = lwe + i * (dim_in + 1);
        LWECiphertext ct(std::vector<uint64_t>(c, c + dim_in), c[dim_in], q, 0);
        LWECiphertext r = b->engine->key_switch(ct, b->key.ksk);
        std::memcpy(out + i * (n_out + 1), r.a.data(), n_out * 8);
        out[i * (n_out + 1) + n_out] = r.b;
    }
    REF_CATCH
}
/* bootstrap_with_test_poly: bootstrap_engine.cpp:684-711 (blind rotate +
 * sample extract + key switch); out = [count][n_out+1]                      */
int ref_boot_bootstrap(void* h, const uint64_t* lwe, size_t count, const uint64_t* test_poly, uint64_t* out,
                       int threads) {
    REF_TRY
    auto* b = static_cast<RefBoot*>(h);
    uint32_t N = b->params.poly_degree, n = b->params.lwe_dimension;
    uint64_t q = b->params.moduli[0];
    size_t n_out = b->key.ksk.keys.empty() ? 0 : b->key.ksk.keys[0].first.degree();
    if (threads < 1) threads = 1;
    ensure_workers(b, threads);
    Polynomial tp = make_poly(test_poly, N, q, false);
    parallel_for(count, threads, [&](size_t i, int t) {
        const uint64_t* c = lwe + i * (n + 1);
        LWECiphertext ct(std::vector<uint64_t>(c, c + n), c[n], q, 0);
        LWECiphertext r = b->workers[t]->bootstrap_with_test_poly(ct, b->key, tp);
        std::memcpy(out + i * (n_out + 1), r.a.data(), n_out * 8);
        out[i * (n_out + 1) + n_out] = r.b;
    });
    REF_CATCH
}

} /* extern "C" */
