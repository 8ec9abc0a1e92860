// Relinearisation of degree-2 ciphertexts: EncryptionEngine::relinearize,
// cpp/src/encryption.cpp:904-985, on the transform kernels of ntt.cu.
//
// Reference, per ciphertext (c0, c1, c2) and key pairs (rlk_a[l], rlk_b[l]):
//     digit_l[i] = (c2[i] >> (l * base_log)) & (base - 1)              raw word, level 0 = LOW bits   :948-955
//     c0 += T^-1( T(digit_l) . T(rlk_b[l]) )    (add_inplace: both sides reduced, result canonical)   :961-973
//     c1 += T^-1( T(digit_l) . T(rlk_a[l]) )                                                          :975-978
// for l < min(num_levels, keys.size()); the reference re-transforms both key polynomials on EVERY call.
// T is Z_q-linear and every partial result is a canonical residue, so
//     c0' = (c0 mod q + T^-1( sum_l T(digit_l) . T(rlk_b[l]) )) mod q
// gives the same words.  The key is therefore transformed ONCE at upload and kept in HBM with its Shoup
// companions; a ciphertext costs `levels` forward and two inverse transforms instead of 3*levels + 2*levels.
//
// C-ABI entry points here (include/fheb200.h): fheb_relin_key_create / _destroy / _levels, fheb_relinearize_batch.
#include <cstdlib>

#include "elementwise.hpp"
#include "modarith.cuh"
#include "plan.hpp"
#include "relin_fused.hpp"
#include "runtime.hpp"

namespace fheb {

struct RelinKey {
    const NttPlan* plan = nullptr;
    uint32_t base_log = 0;  // effective (the reference substitutes 4 for 0, :936)
    uint32_t levels = 0;    // levels actually applied: min(num_levels, key pairs supplied) (:947)
    uint64_t key_id = 0;
    uint64_t* d_key = nullptr;   // [levels][2 (b -> c0, a -> c1)][N] transformed, canonical
    uint64_t* d_keyp = nullptr;  // Shoup companions, same layout
    Tw* d_pack = nullptr;        // the same key in the packed position-order layout of the fused kernel (relin_fused.cu), times N^-1
};

// digits of c2: dig[ct][l][j] = (c2[ct][j] >> (l * base_log)) & mask, c2 = cts[ct][2][:]
__global__ void __launch_bounds__(256) relin_digits_kernel(const uint64_t* __restrict__ cts, uint64_t* __restrict__ dig,
                                                           size_t batch, uint32_t N, uint32_t levels, uint32_t base_log,
                                                           uint64_t mask) {
    const size_t total = batch * N;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t ct = i / N, j = i - ct * N;
        const uint64_t c2 = __ldcs(cts + (ct * 3 + 2) * N + j);
        uint64_t* d = dig + ct * levels * N + j;
        for (uint32_t l = 0; l < levels; ++l) d[(size_t)l * N] = (c2 >> (l * base_log)) & mask;
    }
}

// acc[ct][c][j] = sum_l T(digit)[ct][l][j] * key[l][c][j] mod q   (c = 0: rlk_b, c = 1: rlk_a)
__global__ void __launch_bounds__(256) relin_mac_kernel(const uint64_t* __restrict__ tdig, const uint64_t* __restrict__ key,
                                                        const uint64_t* __restrict__ keyp, uint64_t* __restrict__ acc,
                                                        size_t batch, uint32_t N, uint32_t levels, const ModQ m) {
    const size_t total = batch * N;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t ct = i / N, j = i - ct * N;
        const uint64_t* d = tdig + ct * levels * N + j;
        uint64_t s0 = 0, s1 = 0;  // running sums in [0, 2q)
        for (uint32_t l = 0; l < levels; ++l) {
            const uint64_t x = d[(size_t)l * N];
            const size_t k = ((size_t)l * 2) * N + j;
            s0 = csub(s0 + shoup_lazy(x, key[k], keyp[k], m), m.q2);
            s1 = csub(s1 + shoup_lazy(x, key[k + N], keyp[k + N], m), m.q2);
        }
        acc[(ct * 2) * N + j] = csub(s0, m.q);
        acc[(ct * 2 + 1) * N + j] = csub(s1, m.q);
    }
}

// out[ct][c][j] = (cts[ct][c][j] mod q + prod[ct][c][j]) mod q   (PolynomialRing::add_inplace, :973,978)
__global__ void __launch_bounds__(256) relin_finish_kernel(const uint64_t* __restrict__ cts, const uint64_t* prod, uint64_t* out,
                                                           size_t batch, uint32_t N, const ModQ m) {
    const size_t total = batch * 2 * N;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t ct = i / (2 * (size_t)N), r = i - ct * 2 * N;
        const uint64_t c = canon_any(__ldcs(cts + ct * 3 * N + r), m);
        out[i] = addmod_canon(c, prod[i], m.q);
    }
}

// no key pairs: the reference returns clones of c0 and c1 untouched (:982-989)
__global__ void __launch_bounds__(256) relin_copy_kernel(const uint64_t* __restrict__ cts, uint64_t* __restrict__ out, size_t batch,
                                                         uint32_t N) {
    const size_t total = batch * 2 * N;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t ct = i / (2 * (size_t)N), r = i - ct * 2 * N;
        out[i] = cts[ct * 3 * N + r];
    }
}

__global__ void __launch_bounds__(256) shoup_companions_kernel(const uint64_t* __restrict__ w, uint64_t* __restrict__ wp, size_t count,
                                                               const ModQ m) {
    // floor(w * 2^64 / q) for canonical w: the quotient of (w : 0) by q, from the 128-by-64 remainder
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        // long division, 64 steps of shift-subtract (set-up code, run once per key)
        uint64_t rem = w[i], quo = 0;
        for (int b = 0; b < 64; ++b) {
            const bool top = rem >> 63;
            rem <<= 1;
            quo <<= 1;
            if (top || rem >= m.q) {
                rem -= m.q;
                quo |= 1;
            }
        }
        wp[i] = quo;
    }
}

int relinearize_device(const RelinKey* k, const uint64_t* cts, uint64_t* out, size_t batch, cudaStream_t s) {
    const NttPlan* p = k->plan;
    const uint32_t N = p->degree;
    if (k->levels == 0) {
        relin_copy_kernel<<<stream_grid(batch * 2 * N, 256, 8), 256, 0, s>>>(cts, out, batch, N);
        FHEB_CHECK_LAUNCH();
        count_launch();
        return FHEB_OK;
    }
    if (k->d_pack && !getenv("FHEB_RELIN_UNFUSED")) {  // one launch (relin_fused.cu); FHEB_RELIN_UNFUSED=1: experiments and the parity suite
        RelinFusedArgs a{};
        a.cts = cts;
        a.out = out;
        a.batch = batch;
        a.key = k->d_pack;
        a.levels = k->levels;
        a.base_log = k->base_log;
        a.mask = (k->base_log >= 64) ? ~0ull : ((1ull << k->base_log) - 1);
        a.twf = p->d_fwd;
        a.twi = p->d_inv;
        a.ninv = p->ninv;
        a.m = p->mod;
        const int rc = relin_fused_launch(p->logn, p->mod.dp != 0, a, s);
        if (rc != RELIN_FUSED_UNSUPPORTED) return rc;
    }
    uint64_t* work = nullptr;  // digits [batch][levels][N] | products [batch][2][N]
    const size_t dig_words = batch * k->levels * N;
    FHEB_CUDA(cudaMallocAsync(&work, (dig_words + batch * 2 * N) * 8, s));
    uint64_t* dig = work;
    uint64_t* prod = work + dig_words;
    const uint64_t mask = (k->base_log >= 64) ? ~0ull : ((1ull << k->base_log) - 1);
    int rc = FHEB_OK;
    relin_digits_kernel<<<stream_grid(batch * N, 256, 8), 256, 0, s>>>(cts, dig, batch, N, k->levels, k->base_log, mask);
    if (cudaGetLastError() != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "relin_digits_kernel launch failed");
    count_launch();
    if (rc == FHEB_OK) rc = ntt_forward_device(p, dig, dig, batch * k->levels, s);
    if (rc == FHEB_OK) {
        relin_mac_kernel<<<stream_grid(batch * N, 256, 8), 256, 0, s>>>(dig, k->d_key, k->d_keyp, prod, batch, N, k->levels, p->mod);
        if (cudaGetLastError() != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "relin_mac_kernel launch failed");
        count_launch();
    }
    if (rc == FHEB_OK) rc = ntt_inverse_device(p, prod, prod, batch * 2, s);
    if (rc == FHEB_OK) {
        relin_finish_kernel<<<stream_grid(batch * 2 * N, 256, 8), 256, 0, s>>>(cts, prod, out, batch, N, p->mod);
        if (cudaGetLastError() != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "relin_finish_kernel launch failed");
        count_launch();
    }
    cudaFreeAsync(work, s);
    return rc;
}

}  // namespace fheb

using namespace fheb;

extern "C" {

int fheb_relin_key_create(const fheb_ntt_plan* plan, const uint64_t* keys, uint32_t key_count, uint32_t decomp_base_log,
                          uint32_t decomp_level, uint64_t key_id, fheb_relin_key** out) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(plan != nullptr && out != nullptr, "plan and out must not be null");
    FHEB_REQUIRE(key_count == 0 || keys != nullptr, "keys must not be null");
    const NttPlan* p = reinterpret_cast<const NttPlan*>(plan);
    // cpp/src/encryption.cpp:935-939: zero fields fall back to base_log 4 and ceil(64 / base_log) levels
    const uint32_t base_log = decomp_base_log > 0 ? decomp_base_log : 4;
    const uint32_t num_levels = decomp_level > 0 ? decomp_level : (64 + base_log - 1) / base_log;
    const uint32_t levels = num_levels < key_count ? num_levels : key_count;
    // `1ULL << base_log` and `c2 >> (level * base_log)` are undefined from 64 bits up in the reference
    FHEB_REQUIRE(base_log < 64, "decomp_base_log must be below 64 (got %u)", base_log);
    FHEB_REQUIRE(levels == 0 || (uint64_t)(levels - 1) * base_log < 64,
                 "decomposition shifts reach 64 bits (base_log %u, %u levels): undefined in the reference", base_log, levels);
    RelinKey* k = new RelinKey();
    k->plan = p;
    k->base_log = base_log;
    k->levels = levels;
    k->key_id = key_id;
    *out = reinterpret_cast<fheb_relin_key*>(k);
    if (levels == 0) return FHEB_OK;
    const size_t N = p->degree;
    const size_t words = (size_t)levels * 2 * N;
    cudaStream_t s = ctx().work;
    auto fail = [&](int rc) {
        if (k->d_key) cudaFree(k->d_key);
        if (k->d_keyp) cudaFree(k->d_keyp);
        if (k->d_pack) cudaFree(k->d_pack);
        delete k;
        *out = nullptr;
        return rc;
    };
    if (cudaMalloc(&k->d_key, words * 8) != cudaSuccess || cudaMalloc(&k->d_keyp, words * 8) != cudaSuccess)
        return fail(set_error(FHEB_ERR_OUT_OF_MEMORY, "cudaMalloc of the relinearisation key (%zu bytes) failed", words * 16));
    // caller layout [key_count][2 (a, b)][N]; device layout [levels][2 (b, a)][N]: component 0 feeds c0
    for (uint32_t l = 0; l < levels; ++l) {
        const uint64_t* a = keys + ((size_t)l * 2) * N;
        const uint64_t* b = a + N;
        if (cudaMemcpyAsync(k->d_key + ((size_t)l * 2) * N, b, N * 8, cudaMemcpyDefault, s) != cudaSuccess ||
            cudaMemcpyAsync(k->d_key + ((size_t)l * 2 + 1) * N, a, N * 8, cudaMemcpyDefault, s) != cudaSuccess)
            return fail(set_error(FHEB_ERR_NATIVE, "copying the relinearisation key to the device failed"));
    }
    int rc = ntt_forward_device(p, k->d_key, k->d_key, (size_t)levels * 2, s);  // ring_->to_ntt(rlk_*), :957-963
    if (rc != FHEB_OK) return fail(rc);
    shoup_companions_kernel<<<stream_grid(words, 256, 8), 256, 0, s>>>(k->d_key, k->d_keyp, words, p->mod);
    count_launch();
    if (p->logn >= 5 && p->logn <= 13 && p->top == 0) {  // degrees the fused kernel covers
        if (cudaMalloc(&k->d_pack, words * (p->mod.dp ? 8 : sizeof(Tw))) != cudaSuccess)
            return fail(set_error(FHEB_ERR_OUT_OF_MEMORY, "cudaMalloc of the packed relinearisation key failed"));
        rc = pack_key_rows_device(p, k->d_key, k->d_pack, 1, 2, levels, s);
        if (rc != FHEB_OK) return fail(rc);
    }
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
        return fail(set_error(FHEB_ERR_NATIVE, "relinearisation key set-up failed: %s", cudaGetErrorString(cudaGetLastError())));
    return FHEB_OK;
}

int fheb_relin_key_destroy(fheb_relin_key* key) {
    if (!key) return FHEB_OK;
    RelinKey* k = reinterpret_cast<RelinKey*>(key);
    if (k->d_key) cudaFree(k->d_key);
    if (k->d_keyp) cudaFree(k->d_keyp);
    if (k->d_pack) cudaFree(k->d_pack);
    delete k;
    return FHEB_OK;
}

uint32_t fheb_relin_key_levels(const fheb_relin_key* key) { return key ? reinterpret_cast<const RelinKey*>(key)->levels : 0; }

int fheb_relinearize_batch(const fheb_relin_key* key, const uint64_t* cts, uint64_t ct_key_id, uint64_t* out, size_t batch,
                           void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(key != nullptr, "key must not be null");
    const RelinKey* k = reinterpret_cast<const RelinKey*>(key);
    if (ct_key_id != k->key_id)  // cpp/src/encryption.cpp:912-914
        return set_error(FHEB_ERR_KEY_MISMATCH, "Evaluation key does not match ciphertext key");
    if (batch == 0) return FHEB_OK;
    FHEB_REQUIRE(cts != nullptr && out != nullptr, "ciphertext pointers must not be null");
    const size_t N = k->plan->degree;
    if (all_host({cts, out})) {
        FHEB_REQUIRE(!(cts < out + batch * 2 * N && out < cts + batch * 3 * N), "out must not overlap cts");
        return run_host_pipeline(batch, {{cts, 3 * N * 8, 0, true, false}, {out, 2 * N * 8, 0, false, true}},
                                 [&](void* const* dev, size_t, size_t n, cudaStream_t s) {
                                     return relinearize_device(k, (const uint64_t*)dev[0], (uint64_t*)dev[1], n, s);
                                 });
    }
    // ciphertexts are 3 N words in and 2 N words out: no overlap of the two buffers leaves every input word readable
    // until it is needed (the reference's relinearize_inplace works on ONE ciphertext object, :995-1003)
    FHEB_REQUIRE(!(cts < out + batch * 2 * N && out < cts + batch * 3 * N), "out must not overlap cts");
    cudaStream_t s = (cudaStream_t)stream;
    Staged si, so;
    FHEB_TRY(si.bind(cts, batch * 3 * N * 8, true, false, s));
    FHEB_TRY(so.bind(out, batch * 2 * N * 8, false, true, s));
    FHEB_TRY(relinearize_device(k, si.ptr<const uint64_t>(), so.ptr<uint64_t>(), batch, s));
    FHEB_TRY(so.finish());
    return sync_if_staged(s, {&si, &so});
}

}  // extern "C"
