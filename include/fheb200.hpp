// fheb200.hpp - header-only C++ mirror of the reference's classes for the hot path, over the C ABI
// of fheb200.h.  Same class and method names as the reference (namespace fhe_accelerate there,
// fheb200 here) so that its call sites and tests port by switching the include:
//
//   NTTProcessor                cpp/include/ntt_processor.h:49-303
//   PolynomialRing              cpp/include/polynomial_ring.h:312-513   (flat word buffers instead of Polynomial)
//   MultiLimbModularArithmetic  cpp/include/modular_arithmetic.h:124-194
//   BootstrapEngine             cpp/include/bootstrap_engine.h:176-508  (deterministic part; keys are uploaded)
//   tally_votes / batch_add     cpp/include/encryption.h tally slice
//
// Errors follow the reference: std::invalid_argument for bad parameters (with the reference's
// message text), std::runtime_error for everything else.  Buffers may be host or device memory.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <memory>
#include <vector>

#include "fheb200.h"

namespace fheb200 {

inline void check(int rc) {
    if (rc == FHEB_OK) return;
    const std::string msg = fheb_last_error();
    if (rc == FHEB_ERR_INVALID_PARAMETERS) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

inline void initialize(int device = -1) { check(fheb_init(device)); }
// One process, several GPUs: host-buffer batches of every class below are split over these devices (empty = all visible).
inline int set_devices(const std::vector<int>& devices = {}) {
    check(devices.empty() ? fheb_set_devices(nullptr, -1) : fheb_set_devices(devices.data(), (int)devices.size()));
    return fheb_get_devices(nullptr, 0);
}

struct TwiddleFactors {  // cpp/include/ntt_processor.h:29-41
    std::vector<uint64_t> forward, inverse;
    uint64_t primitive_root = 0, inv_primitive_root = 0, inv_n = 0;
};

class NTTProcessor {
   public:
    NTTProcessor(uint32_t degree, uint64_t modulus) : degree_(degree), modulus_(modulus) {
        check(fheb_ntt_plan_create(degree, modulus, &plan_));
    }
    // caller-supplied tables (fast_ntt_forward / MetalComputeContext::batch_ntt_forward shape)
    NTTProcessor(uint32_t degree, uint64_t modulus, const uint64_t* fwd, const uint64_t* inv, uint64_t inv_n)
        : degree_(degree), modulus_(modulus) {
        check(fheb_ntt_plan_create_with_tables(degree, modulus, fwd, inv, inv_n, &plan_));
    }
    ~NTTProcessor() { fheb_ntt_plan_destroy(plan_); }
    NTTProcessor(const NTTProcessor&) = delete;
    NTTProcessor& operator=(const NTTProcessor&) = delete;

    uint32_t get_degree() const { return degree_; }
    uint64_t get_modulus() const { return modulus_; }
    TwiddleFactors get_twiddles() const {
        TwiddleFactors t;
        t.forward.resize(degree_);
        t.inverse.resize(degree_);
        uint64_t s[3];
        check(fheb_ntt_plan_get_tables(plan_, t.forward.data(), t.inverse.data(), s));
        t.primitive_root = s[0];
        t.inv_primitive_root = s[1];
        t.inv_n = s[2];
        return t;
    }
    // in place, like NTTProcessor::forward_ntt(uint64_t*, size_t) - throws "Size must match polynomial degree"
    void forward_ntt(uint64_t* coeffs, size_t n, void* stream = nullptr) const {
        require_size(n);
        check(fheb_ntt_forward_batch(plan_, coeffs, coeffs, 1, stream));
    }
    void inverse_ntt(uint64_t* coeffs, size_t n, void* stream = nullptr) const {
        require_size(n);
        check(fheb_ntt_inverse_batch(plan_, coeffs, coeffs, 1, stream));
    }
    void forward_ntt(const uint64_t* in, uint64_t* out, size_t n, void* stream = nullptr) const {
        require_size(n);
        check(fheb_ntt_forward_batch(plan_, in, out, 1, stream));
    }
    void inverse_ntt(const uint64_t* in, uint64_t* out, size_t n, void* stream = nullptr) const {
        require_size(n);
        check(fheb_ntt_inverse_batch(plan_, in, out, 1, stream));
    }
    // contiguous batches [batch][N] (the reference's pointer-array overload is a serial loop over these)
    void forward_ntt_batch(const uint64_t* in, uint64_t* out, size_t batch, void* stream = nullptr) const {
        check(fheb_ntt_forward_batch(plan_, in, out, batch, stream));
    }
    void inverse_ntt_batch(const uint64_t* in, uint64_t* out, size_t batch, void* stream = nullptr) const {
        check(fheb_ntt_inverse_batch(plan_, in, out, batch, stream));
    }
    void forward_ntt_batch(uint64_t** coeffs_batch, size_t batch_size, size_t n) const {  // ntt_processor.h:143
        for (size_t i = 0; i < batch_size; ++i) forward_ntt(coeffs_batch[i], n);
    }
    void inverse_ntt_batch(uint64_t** coeffs_batch, size_t batch_size, size_t n) const {
        for (size_t i = 0; i < batch_size; ++i) inverse_ntt(coeffs_batch[i], n);
    }
    const fheb_ntt_plan* handle() const { return plan_; }

   private:
    void require_size(size_t n) const {
        if (n != degree_) throw std::invalid_argument("Size must match polynomial degree");  // ntt_processor.cpp:263-265
    }
    fheb_ntt_plan* plan_ = nullptr;
    uint32_t degree_;
    uint64_t modulus_;
};

class PolynomialRing {
   public:
    PolynomialRing(uint32_t degree, uint64_t modulus) : ntt_(degree, modulus) {}
    uint32_t degree() const { return ntt_.get_degree(); }
    uint64_t modulus() const { return ntt_.get_modulus(); }
    const NTTProcessor& ntt() const { return ntt_; }
    // element-wise over `batch` polynomials (words = batch * degree); r may alias a or b
    void add(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t batch = 1, void* s = nullptr) const {
        check(fheb_modadd_batch(a, b, r, batch * degree(), modulus(), s));
    }
    void subtract(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t batch = 1, void* s = nullptr) const {
        check(fheb_modsub_batch(a, b, r, batch * degree(), modulus(), s));
    }
    void negate(const uint64_t* a, uint64_t* r, size_t batch = 1, void* s = nullptr) const {
        check(fheb_modneg_batch(a, r, batch * degree(), modulus(), s));
    }
    void multiply_scalar(const uint64_t* a, uint64_t scalar, uint64_t* r, size_t batch = 1, void* s = nullptr) const {
        check(fheb_modmul_scalar_batch(a, scalar, r, batch * degree(), modulus(), s));
    }
    void pointwise_multiply(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t batch = 1, void* s = nullptr) const {
        check(fheb_modmul_batch(a, b, r, batch * degree(), modulus(), s));
    }
    void to_ntt(const uint64_t* a, uint64_t* r, size_t batch = 1, void* s = nullptr) const { ntt_.forward_ntt_batch(a, r, batch, s); }
    void from_ntt(const uint64_t* a, uint64_t* r, size_t batch = 1, void* s = nullptr) const { ntt_.inverse_ntt_batch(a, r, batch, s); }
    // coefficient-form product, PolynomialRing::multiply (polynomial_ring.cpp:421-447)
    void multiply(const uint64_t* a, const uint64_t* b, uint64_t* c, size_t batch = 1, void* s = nullptr) const {
        check(fheb_polymul_batch(ntt_.handle(), a, b, c, batch, s));
    }
    // The same with the reference's Polynomial::is_ntt flags on flat buffers (polynomial_ring.cpp:421-447): when BOTH
    // operands are already in transform form the product is the pointwise one and stays in transform form (:425-427);
    // otherwise operands not yet transformed are transformed, and the result comes back in coefficient form.
    // Returns the result's is_ntt flag.
    bool multiply(const uint64_t* a, bool a_is_ntt, const uint64_t* b, bool b_is_ntt, uint64_t* c, size_t batch = 1, void* s = nullptr) const {
        if (a_is_ntt && b_is_ntt) {
            pointwise_multiply(a, b, c, batch, s);
            return true;
        }
        if (!a_is_ntt && !b_is_ntt) {
            multiply(a, b, c, batch, s);
            return false;
        }
        // mixed: transform the coefficient-form operand into c, multiply pointwise, transform back (:433-444)
        const uint64_t* coeff = a_is_ntt ? b : a;
        const uint64_t* ntt = a_is_ntt ? a : b;
        if (c == ntt) throw std::invalid_argument("out must not alias the transform-form operand of a mixed product");
        to_ntt(coeff, c, batch, s);
        pointwise_multiply(c, ntt, c, batch, s);
        from_ntt(c, c, batch, s);
        return false;
    }
    // EncryptionEngine::multiply tensor product: [batch][2][N] x [batch][2][N] -> [batch][3][N]
    void tensor_multiply(const uint64_t* ct1, const uint64_t* ct2, uint64_t* out, size_t batch = 1, void* s = nullptr) const {
        check(fheb_tensor_multiply_batch(ntt_.handle(), ct1, ct2, out, batch, s));
    }

   private:
    NTTProcessor ntt_;
};

// Sharded tally over the GPUs of one box with the exchange fused into the tally kernel (one process per GPU):
// create on every rank, exchange handle() bytes by any transport, connect(), then run() per batch gives the global
// tally on every rank in one launch.  See fheb_tally_peers_* in fheb200.h.
class ShardedTallyPeers {
   public:
    ShardedTallyPeers(uint32_t degree, uint64_t modulus, uint32_t world, uint32_t rank) {
        check(fheb_tally_peers_create(degree, modulus, world, rank, &peers_, handle_));
    }
    ~ShardedTallyPeers() { fheb_tally_peers_destroy(peers_); }
    ShardedTallyPeers(const ShardedTallyPeers&) = delete;
    ShardedTallyPeers& operator=(const ShardedTallyPeers&) = delete;
    const uint8_t* handle() const { return handle_; }  // FHEB_PEER_HANDLE_BYTES bytes to send to every peer
    void connect(const uint8_t* all_handles /* [world][FHEB_PEER_HANDLE_BYTES], rank order */) { check(fheb_tally_peers_connect(peers_, all_handles)); }
    // local_cts = this rank's ballots [count][2][N] (device); out = [2][N] (device): the global tally
    void run(const uint64_t* local_cts, size_t count, uint64_t* out, void* s = nullptr) { check(fheb_tally_peers_run(peers_, local_cts, count, out, s)); }
    bool timed_out() const {  // after synchronising the stream: has any exchange issued so far timed out?
        int v = 0;
        check(fheb_tally_peers_status(peers_, &v));
        return v != 0;
    }
    uint32_t epoch() const { return fheb_tally_peers_epoch(peers_); }
    void reset(uint32_t epoch) { check(fheb_tally_peers_reset(peers_, epoch)); }
    void set_timeout(double seconds) { check(fheb_tally_peers_set_timeout(peers_, seconds)); }

   private:
    fheb_tally_peers* peers_ = nullptr;
    uint8_t handle_[FHEB_PEER_HANDLE_BYTES] = {};
};

// The same sharded tally driven by ONE process that owns several GPUs (the shape of the reference's single-process
// addon): shards[i] = counts[i] ballots in the memory of device i of the group.  See fheb_tally_group_*.
class ShardedTallyGroup {
   public:
    ShardedTallyGroup(uint32_t degree, uint64_t modulus, const std::vector<int>& devices) {
        check(fheb_tally_group_create(degree, modulus, devices.data(), (uint32_t)devices.size(), &group_));
    }
    ShardedTallyGroup(uint32_t degree, uint64_t modulus, uint32_t ndev) { check(fheb_tally_group_create(degree, modulus, nullptr, ndev, &group_)); }
    ~ShardedTallyGroup() { fheb_tally_group_destroy(group_); }
    ShardedTallyGroup(const ShardedTallyGroup&) = delete;
    ShardedTallyGroup& operator=(const ShardedTallyGroup&) = delete;
    uint32_t size() const { return fheb_tally_group_size(group_); }
    // EncryptionEngine::tally_votes over the shards; out = [2][N] in host memory or in any device's memory
    void tally_votes(const std::vector<const uint64_t*>& shards, const std::vector<size_t>& counts, uint64_t* out) {
        if (shards.size() != size() || counts.size() != size()) throw std::invalid_argument("one shard and one count per device of the group");
        check(fheb_tally_sharded(group_, shards.data(), counts.data(), out));
    }

   private:
    fheb_tally_group* group_ = nullptr;
};

// PolynomialRing(degree, moduli) (polynomial_ring.cpp:224-237) with every limb live: data limb-major
// [limbs][batch][N], limb l computed over moduli[l] by its own plan.  (The reference's multi-modulus ring only ever
// computes with moduli[0]; limb 0 here is exactly that.)
class RnsPolynomialRing {
   public:
    RnsPolynomialRing(uint32_t degree, const std::vector<uint64_t>& moduli) : degree_(degree) {
        if (moduli.empty()) throw std::invalid_argument("At least one modulus required");  // polynomial_ring.cpp:228-230
        for (uint64_t q : moduli) rings_.emplace_back(new PolynomialRing(degree, q));
    }
    size_t limbs() const { return rings_.size(); }
    const PolynomialRing& limb(size_t l) const { return *rings_[l]; }
    void to_ntt(const uint64_t* a, uint64_t* r, size_t batch, void* s = nullptr) const {
        for (size_t l = 0; l < limbs(); ++l) rings_[l]->to_ntt(a + l * batch * degree_, r + l * batch * degree_, batch, s);
    }
    void from_ntt(const uint64_t* a, uint64_t* r, size_t batch, void* s = nullptr) const {
        for (size_t l = 0; l < limbs(); ++l) rings_[l]->from_ntt(a + l * batch * degree_, r + l * batch * degree_, batch, s);
    }
    void add(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t batch, void* s = nullptr) const {
        for (size_t l = 0; l < limbs(); ++l) rings_[l]->add(a + l * batch * degree_, b + l * batch * degree_, r + l * batch * degree_, batch, s);
    }
    void multiply(const uint64_t* a, const uint64_t* b, uint64_t* c, size_t batch, void* s = nullptr) const {
        for (size_t l = 0; l < limbs(); ++l) rings_[l]->multiply(a + l * batch * degree_, b + l * batch * degree_, c + l * batch * degree_, batch, s);
    }

   private:
    uint32_t degree_;
    std::vector<std::unique_ptr<PolynomialRing>> rings_;
};

// EvaluationKey.relin_key on the device (cpp/include/key_manager.h:92-111), pre-transformed once.
class RelinearizationKey {
   public:
    // keys = [key_count][2][N]: (a, b) of every pair in generation order (KeyManager::generate_eval_key)
    RelinearizationKey(const PolynomialRing& ring, const uint64_t* keys, uint32_t key_count, uint32_t decomp_base_log,
                       uint32_t decomp_level, uint64_t key_id = 0)
        : key_id_(key_id) {
        check(fheb_relin_key_create(ring.ntt().handle(), keys, key_count, decomp_base_log, decomp_level, key_id, &key_));
    }
    ~RelinearizationKey() { fheb_relin_key_destroy(key_); }
    RelinearizationKey(const RelinearizationKey&) = delete;
    RelinearizationKey& operator=(const RelinearizationKey&) = delete;
    uint32_t levels() const { return fheb_relin_key_levels(key_); }
    // EncryptionEngine::relinearize (encryption.cpp:904-993): [batch][3][N] -> [batch][2][N]
    void relinearize(const uint64_t* cts, uint64_t* out, size_t batch = 1, void* s = nullptr) const {
        check(fheb_relinearize_batch(key_, cts, key_id_, out, batch, s));
    }

   private:
    fheb_relin_key* key_ = nullptr;
    uint64_t key_id_;
};

class MultiLimbModularArithmetic {
   public:
    explicit MultiLimbModularArithmetic(const std::vector<uint64_t>& q) : q_(q), consts_(1 + 2 * q.size()) {
        check(fheb_mlimb_constants(q_.data(), (uint32_t)q_.size(), consts_.data()));
    }
    size_t limbs() const { return q_.size(); }
    uint64_t q_inv() const { return consts_[0]; }
    const uint64_t* r_mod_q() const { return consts_.data() + 1; }
    const uint64_t* r2_mod_q() const { return consts_.data() + 1 + q_.size(); }
    // [count][limbs] little-endian words
    void montgomery_mul_neon(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, void* s = nullptr) const {
        check(fheb_mlimb_montmul_batch(a, b, r, count, (uint32_t)q_.size(), q_.data(), consts_[0], s));
    }
    void mod_add_neon(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, void* s = nullptr) const {
        check(fheb_mlimb_add_batch(a, b, r, count, (uint32_t)q_.size(), q_.data(), s));
    }
    void mod_sub_neon(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, void* s = nullptr) const {
        check(fheb_mlimb_sub_batch(a, b, r, count, (uint32_t)q_.size(), q_.data(), s));
    }

   private:
    std::vector<uint64_t> q_, consts_;
};

class BootstrapEngine {
   public:
    // bsk = [n][(k+1)*level][k+1][N] coefficient-form words as generate_bootstrap_key produces them
    BootstrapEngine(uint32_t poly_degree, uint64_t modulus, uint32_t lwe_dimension, uint32_t glwe_dimension,
                    uint32_t decomp_base_log, uint32_t decomp_level, const uint64_t* bsk)
        : ntt_(poly_degree, modulus), n_(lwe_dimension), k_(glwe_dimension) {
        const fheb_boot_params p{lwe_dimension, glwe_dimension, decomp_base_log, decomp_level};
        check(fheb_boot_key_create(ntt_.handle(), &p, bsk, &key_));
    }
    ~BootstrapEngine() { fheb_boot_key_destroy(key_); }
    BootstrapEngine(const BootstrapEngine&) = delete;
    BootstrapEngine& operator=(const BootstrapEngine&) = delete;

    void set_key_switch_key(const uint64_t* ksk, size_t entries, uint32_t n_out, uint32_t base_log, uint32_t level) {
        check(fheb_boot_key_set_ksk(key_, ksk, entries, n_out, base_log, level));
    }
    void external_product(const uint64_t* glwe, uint32_t index, uint64_t* out, size_t batch = 1, void* s = nullptr) const {
        check(fheb_external_product_batch(key_, index, glwe, out, batch, s));
    }
    void cmux(uint32_t index, const uint64_t* ct0, const uint64_t* ct1, uint64_t* out, size_t batch = 1, void* s = nullptr) const {
        check(fheb_cmux_batch(key_, index, ct0, ct1, out, batch, s));
    }
    void blind_rotate(const uint64_t* lwe, const uint64_t* test_poly, uint64_t* out, size_t batch = 1, void* s = nullptr) const {
        check(fheb_blind_rotate_batch(key_, lwe, test_poly, out, batch, s));
    }
    void sample_extract(const uint64_t* glwe, uint64_t* out, size_t batch = 1, void* s = nullptr) const {
        check(fheb_sample_extract_batch(key_, glwe, out, batch, s));
    }
    void key_switch(const uint64_t* lwe, uint64_t* out, size_t batch = 1, void* s = nullptr) const {
        check(fheb_key_switch_batch(key_, lwe, out, batch, s));
    }
    void bootstrap_with_test_poly(const uint64_t* lwe, const uint64_t* test_poly, uint64_t* out, size_t batch = 1, void* s = nullptr) const {
        check(fheb_bootstrap_batch(key_, lwe, test_poly, out, batch, s));
    }
    std::vector<uint64_t> get_default_test_poly(uint64_t plaintext_modulus = 4) const { return lut(3, plaintext_modulus, 0); }
    std::vector<uint64_t> create_identity_lut(uint64_t modulus) const { return lut(0, modulus, 0); }
    std::vector<uint64_t> create_negation_lut(uint64_t modulus) const { return lut(1, modulus, 0); }
    std::vector<uint64_t> create_threshold_lut(uint64_t threshold, uint64_t modulus) const { return lut(2, threshold, modulus); }

   private:
    std::vector<uint64_t> lut(int kind, uint64_t a0, uint64_t a1) const {
        std::vector<uint64_t> out(ntt_.get_degree());
        check(fheb_make_test_poly(ntt_.handle(), kind, a0, a1, out.data()));
        return out;
    }
    NTTProcessor ntt_;
    fheb_boot_key* key_ = nullptr;
    uint32_t n_, k_;
};

// EncryptionEngine::batch_add / tally_votes words: cts = [count][2][N] -> out = [2][N]
inline void tally_votes(const uint64_t* cts, size_t count, uint32_t degree, uint64_t modulus, uint64_t* out, void* s = nullptr) {
    check(fheb_tally(cts, count, degree, modulus, out, s));
}

// Ciphertext::noise_budget of the result (encryption.h:40-89): the words do not depend on the tally variant, this does.
enum class TallyVariant { BatchAdd = 0, BatchAddTree = 1, AddFold = 2 };  // encryption.cpp:1359-1360 / :1413,1437 / :613
inline double tally_noise_budget(const std::vector<double>& budgets, TallyVariant variant = TallyVariant::BatchAddTree) {
    double out = 0.0;
    check(fheb_tally_noise_budget(budgets.data(), budgets.size(), (int)variant, &out));
    return out;
}
// the reference's Ciphertext, as far as the tally touches it: words [2][N] + metadata
struct TallyResult {
    std::vector<uint64_t> words;  // c0 then c1
    double noise_budget = 0.0;
    uint64_t key_id = 0;
};
// EncryptionEngine::tally_votes (= batch_add_tree, encryption.cpp:1061-1067) with the metadata checks of :1345-1347
inline TallyResult tally_votes(const uint64_t* cts, size_t count, uint32_t degree, uint64_t modulus, const std::vector<double>& budgets,
                               const std::vector<uint64_t>& key_ids, TallyVariant variant = TallyVariant::BatchAddTree) {
    if (budgets.size() != count || key_ids.size() != count) throw std::invalid_argument("one noise budget and one key id per ballot");
    for (size_t i = 1; i < count; ++i)
        if (key_ids[i] != key_ids[0]) throw std::invalid_argument("All ciphertexts must be encrypted with the same key");
    TallyResult r;
    r.words.resize((size_t)2 * degree);
    tally_votes(cts, count, degree, modulus, r.words.data());
    r.noise_budget = tally_noise_budget(budgets, variant);
    r.key_id = count ? key_ids[0] : 0;
    return r;
}

}  // namespace fheb200
