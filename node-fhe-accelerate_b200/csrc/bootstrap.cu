// TFHE bootstrap chain: device-resident keys, blind rotation (persistent fused kernel in
// boot_kernel.cuh), sample extraction, key switching and the host-side LUT builders.
//
// C-ABI entry points here (include/fheb200.h): fheb_boot_key_create, fheb_boot_key_set_ksk,
// fheb_boot_key_destroy, fheb_external_product_batch, fheb_cmux_batch, fheb_blind_rotate_batch,
// fheb_sample_extract_batch, fheb_key_switch_batch, fheb_bootstrap_batch, fheb_make_test_poly.
#include <map>
#include <mutex>
#include <type_traits>

#include "boot_kernel.cuh"
#include "elementwise.hpp"
#include "ntt_plan.hpp"
#include "plan.hpp"
#include "relin_fused.hpp"

namespace fheb {

struct BootKey {
    const NttPlan* plan = nullptr;  // borrowed: the plan must outlive the key
    uint32_t n = 0, k = 0, base_log = 0, levels = 0;
    Tw* d_bsk = nullptr;            // [n][rows][k+1][N], transformed, position order: Shoup pairs, or doubles (mod.dp)
    uint64_t* d_ksk = nullptr;      // [entries][n_out + 1] raw words
    size_t ksk_entries = 0;
    uint32_t ksk_n_out = 0, ksk_base_log = 0, ksk_levels = 0;
    int* d_raw_flag = nullptr;      // "the test polynomial of the running blind rotation has words >= q" (lean / general kernel choice)
    // device that holds the buffers above, and copies of the key on other devices (made on first use when a host batch
    // is spread over several GPUs; owned by this key, dropped when the key switching key changes)
    int device = 0;
    mutable std::map<int, BootKey*> replicas;
    mutable std::mutex replica_mutex;
};

static size_t glwe_words(const BootKey* key);
static size_t ggsw_words(const BootKey* key);

static void boot_key_free(BootKey* key) {
    if (!key) return;
    for (auto& kv : key->replicas) boot_key_free(kv.second);
    if (key->d_bsk) cudaFree(key->d_bsk);
    if (key->d_ksk) cudaFree(key->d_ksk);
    if (key->d_raw_flag) cudaFree(key->d_raw_flag);
    delete key;
}

// the key on `device` (the calling thread's current device): the transformed key is copied device to device, the plan
// is replicated the same way
static const BootKey* boot_key_on_device(const BootKey* key, int device) {
    if (key->device == device) return key;
    std::lock_guard<std::mutex> lock(key->replica_mutex);
    auto it = key->replicas.find(device);
    if (it != key->replicas.end()) return it->second;
    const NttPlan* pd = plan_on_device(key->plan, device);
    if (!pd) return nullptr;
    BootKey* r = new BootKey();
    r->plan = pd;
    r->n = key->n;
    r->k = key->k;
    r->base_log = key->base_log;
    r->levels = key->levels;
    r->device = device;
    r->ksk_entries = key->ksk_entries;
    r->ksk_n_out = key->ksk_n_out;
    r->ksk_base_log = key->ksk_base_log;
    r->ksk_levels = key->ksk_levels;
    const size_t bsk_bytes = (size_t)key->n * ggsw_words(key) * (pd->mod.dp ? 8 : sizeof(Tw));
    const size_t ksk_bytes = key->d_ksk ? key->ksk_entries * ((size_t)key->ksk_n_out + 1) * 8 : 0;
    cudaError_t e = cudaMalloc(&r->d_bsk, bsk_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_bsk, key->d_bsk, bsk_bytes, cudaMemcpyDefault);
    if (e == cudaSuccess && ksk_bytes) e = cudaMalloc(&r->d_ksk, ksk_bytes);
    if (e == cudaSuccess && ksk_bytes) e = cudaMemcpy(r->d_ksk, key->d_ksk, ksk_bytes, cudaMemcpyDefault);
    if (e == cudaSuccess && cudaMalloc(&r->d_raw_flag, sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        r->d_raw_flag = nullptr;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error(e == cudaErrorMemoryAllocation ? FHEB_ERR_OUT_OF_MEMORY : FHEB_ERR_NATIVE, "replicating the bootstrapping key on device %d failed: %s",
                  device, cudaGetErrorString(e));
        boot_key_free(r);
        return nullptr;
    }
    key->replicas[device] = r;
    return r;
}

// ---- key preparation --------------------------------------------------------------------
// in: transforms y = [ggsw][row][j][N] in the reference's output order (index p of the permuted array).
// out, per GGSW: element (row, e, u, j) = y[row][j][bitrev(u * 2^rlast + e)] * N^-1 mod q, stored at
// ((row * E + e) * ITEMS + u) * KP1 + j with its Shoup companion floor(w * 2^64 / q), or as a double in DP mode.
// N^-1 is folded in here because the transform is linear: T^-1(sum D.G) = unscaled network of sum D.(G N^-1).
__global__ void __launch_bounds__(256) bsk_pack_kernel(const uint64_t* __restrict__ y, Tw* __restrict__ g, size_t words,
                                                       uint32_t logn, uint32_t rlast, uint32_t kp1, uint32_t rows,
                                                       const ModQ m, uint64_t ninv, int dp) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint32_t N = 1u << logn, E = 1u << rlast, items = N >> rlast;
    const size_t ggsw_words = (size_t)rows * kp1 * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride) {
        const size_t gidx = i / ggsw_words;
        uint32_t rem = (uint32_t)(i - gidx * ggsw_words);
        const uint32_t j = rem % kp1;
        rem /= kp1;
        const uint32_t u = rem % items;
        rem /= items;
        const uint32_t e = rem % E, row = rem / E;
        const uint32_t pos = (u << rlast) | e;
        const size_t src = gidx * ggsw_words + ((size_t)row * kp1 + j) * N + bitrev_rt(pos, (int)logn);
        const uint64_t w = mulmod(y[src], ninv, m);
        if (dp) {  // FP64 mode: the value as a double, 8 bytes per entry
            reinterpret_cast<uint64_t*>(g)[i] = double_to_bits((double)w);
        } else {
            Tw t;
            t.w = w;
            t.wp = (uint64_t)((((u128)w) << 64) / m.q);
            g[i] = t;
        }
    }
}

int pack_key_rows_device(const NttPlan* p, const uint64_t* y, Tw* g, size_t ggsws, uint32_t kp1, uint32_t rows, cudaStream_t s) {
    const size_t words = ggsws * rows * kp1 * p->degree;
    bsk_pack_kernel<<<stream_grid(words, 256, 8), 256, 0, s>>>(y, g, words, p->logn, last_pass_width(p->logn), kp1, rows, p->mod,
                                                                  p->inv_n % p->modulus, (int)p->mod.dp);
    if (cudaGetLastError() != cudaSuccess) return set_error(FHEB_ERR_NATIVE, "bsk_pack_kernel launch failed");
    count_launch();
    return FHEB_OK;
}

// ---- sample extraction ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_extract_kernel(const uint64_t* __restrict__ glwe, uint64_t* __restrict__ out,
                                                             size_t batch, uint32_t k, uint32_t N, const ModQ m) {
    const size_t width = (size_t)k * N + 1;
    const size_t total = batch * width;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t ct = i / width;
        const uint32_t idx = (uint32_t)(i - ct * width);
        out[i] = sample_extract_word(glwe + ct * ((size_t)k + 1) * N, idx, k, N, m);
    }
}

// ---- key switching (cpp/src/bootstrap_engine.cpp:626-669) -------------------------------------
// out[j] = (0 - sum_idx t(idx, j)) mod q with t = ((digit_idx * ksk[idx][j]) mod 2^64) % q (summed unreduced, see below) and
// digit_idx the low-bit digit of input coefficient idx / levels; the reference's running
// `(res + q - t) % q` is this sum in any order.  The b word starts from the input's b instead of 0.
// Block: KS_COLS output columns x KS_CTS ciphertexts; digits of the block's ciphertexts are staged
// in shared memory in chunks; every KSK word is read once per block and reused KS_CTS times.
constexpr int KS_COLS = 128;
constexpr int KS_CTS = 8;
constexpr int KS_CHUNK = 512;

// SMALL: digits below 2^32 (base_log <= 32) are kept as 32-bit words: the product is one wide and one narrow multiply
template <bool SMALL>
__global__ void __launch_bounds__(KS_COLS) key_switch_kernel(const uint64_t* __restrict__ lwe, const uint64_t* __restrict__ ksk,
                                                             uint64_t* __restrict__ out, size_t batch, uint32_t dim_in,
                                                             uint32_t n_out, uint32_t base_log, uint32_t levels, const ModQ m) {
    using Dig = typename std::conditional<SMALL, uint32_t, uint64_t>::type;
    __shared__ Dig digits[KS_CTS][KS_CHUNK];
    const uint32_t col = blockIdx.x * KS_COLS + threadIdx.x;
    const size_t ct0 = (size_t)blockIdx.y * KS_CTS;
    const uint32_t ncts = (uint32_t)((batch - ct0) < (size_t)KS_CTS ? (batch - ct0) : (size_t)KS_CTS);
    const size_t in_w = (size_t)dim_in + 1, out_w = (size_t)n_out + 1;
    const uint64_t mask = (base_log >= 64) ? ~0ull : ((1ull << base_log) - 1);
    const size_t entries = (size_t)dim_in * levels;
    // sum_e (x_e % q) == (sum_e x_e) % q for the wrapped 64-bit products x_e = digit * ksk mod 2^64: the raw products are
    // summed in a 128-bit counter per ciphertext and reduced once (one multiply per term instead of a Barrett
    // reduction per term; exact for any number of terms)
    uint64_t lo[KS_CTS], hi[KS_CTS];
#pragma unroll
    for (int c = 0; c < KS_CTS; ++c) lo[c] = hi[c] = 0;
    for (size_t e0 = 0; e0 < entries; e0 += KS_CHUNK) {
        const uint32_t chunk = (uint32_t)((entries - e0) < (size_t)KS_CHUNK ? (entries - e0) : (size_t)KS_CHUNK);
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < KS_CTS * chunk; t += KS_COLS) {
            const uint32_t c = t / chunk, e = t % chunk;
            uint64_t d = 0;
            if (c < ncts) {
                const size_t idx = e0 + e;
                const uint32_t i = (uint32_t)(idx / levels), l = (uint32_t)(idx % levels);
                d = (lwe[(ct0 + c) * in_w + i] >> ((levels - 1 - l) * base_log)) & mask;
            }
            digits[c][e] = (Dig)d;
        }
        __syncthreads();
        if (col < out_w) {
            const uint64_t* kp = ksk + e0 * out_w + col;
#pragma unroll 4
            for (uint32_t e = 0; e < chunk; ++e) {
                const uint64_t kv = __ldg(kp + (size_t)e * out_w);
#pragma unroll
                for (int c = 0; c < KS_CTS; ++c) {
                    acc128(lo[c], hi[c], (uint64_t)digits[c][e] * kv);  // zero digits contribute nothing (:654-657)
                }
            }
        }
    }
    if (col < out_w) {
        for (uint32_t c = 0; c < ncts; ++c) {
            uint64_t start = 0;
            const uint64_t total = fold128(hi[c], lo[c], m);
            if (col == n_out) {
                const uint64_t* in = lwe + (ct0 + c) * in_w;
                const uint64_t b = in[dim_in];
                start = b;
                if (b >= m.q) {
                    // Unreduced b (:640,665): the reference keeps lwe.b RAW until the first non-zero digit, whose update
                    // `(b + q - t1) % q` is computed in wrapping u64 arithmetic; with no non-zero digit b is returned as
                    // it came.  Rare path: one thread re-scans the digits for t1.
                    bool found = false;
                    uint64_t t1 = 0;
                    for (size_t idx = 0; idx < entries && !found; ++idx) {
                        const uint32_t i = (uint32_t)(idx / levels), l = (uint32_t)(idx % levels);
                        const uint64_t d = (in[i] >> ((levels - 1 - l) * base_log)) & mask;
                        if (d != 0) {
                            found = true;
                            t1 = reduce64(d * ksk[idx * out_w + n_out], m);
                        }
                    }
                    if (!found) {
                        out[(ct0 + c) * out_w + col] = b;
                        continue;
                    }
                    const uint64_t r1 = reduce64(b + (m.q - t1), m);  // wraps modulo 2^64 as the reference's sum does
                    start = addmod_canon(r1, t1, m.q);               // `total` below includes t1 again
                }
            }
            out[(ct0 + c) * out_w + col] = submod_canon(start, total, m.q);
        }
    }
}

static int boot_dispatch(const BootKey* key, const BootLaunch& a, cudaStream_t s) {
    const bool dp = key->plan->mod.dp != 0;
    switch (key->k) {
        case 1: return boot_launch_k1(key->plan->logn, dp, a, s);
        case 2: return boot_launch_k2(key->plan->logn, dp, a, s);
        case 3: return boot_launch_k3(key->plan->logn, dp, a, s);
    }
    return set_error(FHEB_ERR_INVALID_PARAMETERS, "glwe_dimension must be 1, 2 or 3 on this backend (got %u)", key->k);
}

static BootLaunch base_launch(const BootKey* key) {
    BootLaunch a{};
    a.n = key->n;
    a.levels = key->levels;
    a.base_log = key->base_log;
    a.twf = key->plan->d_fwd;
    a.twi = key->plan->d_inv;
    a.ninv = key->plan->ninv;
    a.m = key->plan->mod;
    return a;
}

static size_t glwe_words(const BootKey* key) { return ((size_t)key->k + 1) * key->plan->degree; }
static size_t ggsw_words(const BootKey* key) {
    return ((size_t)key->k + 1) * key->levels * ((size_t)key->k + 1) * key->plan->degree;
}

// flag = 1 when any word of the test polynomial is unreduced (>= q)
__global__ void __launch_bounds__(256) test_poly_raw_kernel(const uint64_t* __restrict__ tp, uint32_t N, uint64_t q, int* flag) {
    int raw = 0;
    for (uint32_t i = threadIdx.x; i < N; i += 256) raw |= (tp[i] >= q) ? 1 : 0;
    raw = __syncthreads_or(raw);
    if (threadIdx.x == 0) *flag = raw;
}

// device-pointer stages of the chain
static int blind_rotate_device(const BootKey* key, const uint64_t* lwe, const uint64_t* test_poly, uint64_t* out,
                               size_t batch, cudaStream_t s) {
    BootLaunch a = base_launch(key);
    a.mode = BOOT_BLIND;
    a.bsk = key->d_bsk;
    a.in0 = lwe;
    a.in1 = test_poly;
    a.out = out;
    a.batch = batch;
    if (key->d_raw_flag && key->k == 1) {
        test_poly_raw_kernel<<<1, 256, 0, s>>>(test_poly, key->plan->degree, key->plan->mod.q, key->d_raw_flag);
        FHEB_CHECK_LAUNCH();
        count_launch();
        a.raw_flag = key->d_raw_flag;
    }
    return boot_dispatch(key, a, s);
}

static int sample_extract_device(const BootKey* key, const uint64_t* glwe, uint64_t* out, size_t batch, cudaStream_t s) {
    const size_t total = batch * ((size_t)key->k * key->plan->degree + 1);
    sample_extract_kernel<<<stream_grid(total, 256, 8), 256, 0, s>>>(glwe, out, batch, key->k, key->plan->degree,
                                                                     key->plan->mod);
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

static int key_switch_device(const BootKey* key, const uint64_t* lwe, uint64_t* out, size_t batch, cudaStream_t s) {
    const uint32_t dim_in = key->k * key->plan->degree;
    const dim3 grid((key->ksk_n_out + 1 + KS_COLS - 1) / KS_COLS, (unsigned)((batch + KS_CTS - 1) / KS_CTS));
    if (key->ksk_base_log <= 32)
        key_switch_kernel<true><<<grid, KS_COLS, 0, s>>>(lwe, key->d_ksk, out, batch, dim_in, key->ksk_n_out, key->ksk_base_log,
                                                         key->ksk_levels, key->plan->mod);
    else
        key_switch_kernel<false><<<grid, KS_COLS, 0, s>>>(lwe, key->d_ksk, out, batch, dim_in, key->ksk_n_out, key->ksk_base_log,
                                                          key->ksk_levels, key->plan->mod);
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

// blind rotation -> sample extraction (-> key switching when a KSK is set), device pointers
static int bootstrap_chain_device(const BootKey* key, const uint64_t* lwe, const uint64_t* test_poly, uint64_t* out, size_t batch,
                                  cudaStream_t s) {
    const size_t ext_w = (size_t)key->k * key->plan->degree + 1;
    uint64_t* acc = nullptr;
    uint64_t* ext = nullptr;
    FHEB_CUDA(cudaMallocAsync(&acc, batch * glwe_words(key) * 8, s));
    int rc = blind_rotate_device(key, lwe, test_poly, acc, batch, s);
    if (rc == FHEB_OK && key->d_ksk) {
        if (cudaMallocAsync(&ext, batch * ext_w * 8, s) != cudaSuccess) rc = set_error(FHEB_ERR_OUT_OF_MEMORY, "cudaMallocAsync failed");
        if (rc == FHEB_OK) rc = sample_extract_device(key, acc, ext, batch, s);
        if (rc == FHEB_OK) rc = key_switch_device(key, ext, out, batch, s);
    } else if (rc == FHEB_OK) {
        rc = sample_extract_device(key, acc, out, batch, s);
    }
    cudaFreeAsync(acc, s);
    if (ext) cudaFreeAsync(ext, s);
    return rc;
}

// The chain is compute bound (tens of microseconds per ciphertext, a few KB each way): host batches are
// cut only into very large chunks so that every launch still fills the GPU for many waves.
constexpr size_t BOOT_HOST_CHUNK = 16384;
constexpr size_t BOOT_SPREAD_MIN = 128;  // ciphertexts per device below which spreading a host batch over several GPUs is not worth a thread

static int check_key(const fheb_boot_key* key) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(key != nullptr, "bootstrap key must not be null");
    return FHEB_OK;
}

}  // namespace fheb

using namespace fheb;

extern "C" {

int fheb_boot_key_create(const fheb_ntt_plan* plan, const fheb_boot_params* params, const uint64_t* bsk,
                         fheb_boot_key** out) {
    FHEB_REQUIRE(out != nullptr, "out must not be null");
    *out = nullptr;
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(plan != nullptr && params != nullptr && bsk != nullptr, "plan, params and bsk must not be null");
    const NttPlan* p = reinterpret_cast<const NttPlan*>(plan);
    FHEB_REQUIRE(params->glwe_dimension >= 1 && params->glwe_dimension <= 3, "glwe_dimension must be 1, 2 or 3 on this backend");
    FHEB_REQUIRE(params->decomp_level >= 1 && params->decomp_base_log >= 1 &&
                     (uint64_t)params->decomp_level * params->decomp_base_log <= 64 && params->decomp_base_log <= 63,
                 "decomposition must satisfy 1 <= base_log <= 63 and level * base_log <= 64");
    FHEB_REQUIRE(params->lwe_dimension >= 1, "lwe_dimension must be positive");
    FHEB_REQUIRE(p->logn >= 5 && p->logn <= 12, "bootstrap kernels support polynomial degrees 32..4096");
    FHEB_REQUIRE((params->glwe_dimension + 1) * params->decomp_level <= 64, "at most 64 gadget rows ((k+1) * level)");
    BootKey* key = new BootKey();
    key->plan = p;
    key->n = params->lwe_dimension;
    key->k = params->glwe_dimension;
    key->base_log = params->decomp_base_log;
    key->levels = params->decomp_level;
    key->device = ctx().device;
    if (cudaMalloc(&key->d_raw_flag, sizeof(int)) != cudaSuccess) {  // without it only the general kernel runs
        cudaGetLastError();
        key->d_raw_flag = nullptr;
    }
    const size_t words = (size_t)key->n * ggsw_words(key);
    const size_t polys = words / p->degree;
    cudaStream_t s = nullptr;
    int rc = FHEB_OK;
    uint64_t* tmp = nullptr;
    {
        Staged in;
        rc = in.bind(bsk, words * 8, true, false, s);
        if (rc == FHEB_OK && cudaMalloc(&tmp, words * 8) != cudaSuccess) rc = set_error(FHEB_ERR_OUT_OF_MEMORY, "cudaMalloc of the key staging buffer failed");
        if (rc == FHEB_OK && cudaMalloc(&key->d_bsk, words * (p->mod.dp ? 8 : sizeof(Tw))) != cudaSuccess) rc = set_error(FHEB_ERR_OUT_OF_MEMORY, "cudaMalloc of the device bootstrapping key failed");
        // T(row polynomial) once, here, instead of on every external product (bootstrap_engine.cpp:478-487)
        if (rc == FHEB_OK) rc = ntt_forward_device(p, in.ptr<const uint64_t>(), tmp, polys, s);
        if (rc == FHEB_OK) {
            rc = pack_key_rows_device(p, tmp, key->d_bsk, key->n, key->k + 1, (key->k + 1) * key->levels, s);
        }
        if (rc == FHEB_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "bootstrapping key upload failed");
    }
    if (tmp) cudaFree(tmp);
    if (rc != FHEB_OK) {
        fheb_boot_key_destroy(reinterpret_cast<fheb_boot_key*>(key));
        return rc;
    }
    *out = reinterpret_cast<fheb_boot_key*>(key);
    return FHEB_OK;
}

int fheb_boot_key_set_ksk(fheb_boot_key* key_, const uint64_t* ksk, size_t entries, uint32_t n_out, uint32_t base_log,
                          uint32_t level) {
    FHEB_TRY(check_key(key_));
    BootKey* key = reinterpret_cast<BootKey*>(key_);
    FHEB_REQUIRE(ksk != nullptr, "ksk must not be null");
    FHEB_REQUIRE(level >= 1 && base_log >= 1 && base_log <= 63 && (uint64_t)level * base_log <= 64,
                 "decomposition must satisfy 1 <= base_log <= 63 and level * base_log <= 64");
    FHEB_REQUIRE(entries == (size_t)key->k * key->plan->degree * level,
                 "key switching key must hold k * N * level entries (got %zu)", entries);
    FHEB_REQUIRE(n_out >= 1, "output dimension must be positive");
    {   // replicas carry the old key switching key: drop them, they are rebuilt on demand
        std::lock_guard<std::mutex> lock(key->replica_mutex);
        for (auto& kv : key->replicas) boot_key_free(kv.second);
        key->replicas.clear();
    }
    if (key->d_ksk) cudaFree(key->d_ksk);
    key->d_ksk = nullptr;
    const size_t bytes = entries * ((size_t)n_out + 1) * 8;
    FHEB_CUDA(cudaMalloc(&key->d_ksk, bytes));
    FHEB_CUDA(cudaMemcpy(key->d_ksk, ksk, bytes, cudaMemcpyDefault));
    key->ksk_entries = entries;
    key->ksk_n_out = n_out;
    key->ksk_base_log = base_log;
    key->ksk_levels = level;
    return FHEB_OK;
}

int fheb_boot_key_destroy(fheb_boot_key* key_) {
    boot_key_free(reinterpret_cast<BootKey*>(key_));
    return FHEB_OK;
}

static int one_step_entry(const fheb_boot_key* key_, uint32_t index, int mode, const uint64_t* in0, const uint64_t* in1,
                          uint64_t* out, size_t batch, void* stream) {
    FHEB_TRY(check_key(key_));
    const BootKey* key = reinterpret_cast<const BootKey*>(key_);
    FHEB_REQUIRE(index < key->n, "bootstrapping key index %u out of range (n = %u)", index, key->n);
    if (batch == 0) return FHEB_OK;
    FHEB_REQUIRE(in0 != nullptr && out != nullptr && (mode != BOOT_CMUX || in1 != nullptr), "ciphertext pointers must not be null");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = batch * glwe_words(key) * 8;
    Staged s0, s1, so;
    FHEB_TRY(s0.bind(in0, bytes, true, false, s));
    if (mode == BOOT_CMUX) FHEB_TRY(s1.bind(in1, bytes, true, false, s));
    FHEB_TRY(so.bind(out, bytes, false, true, s));
    BootLaunch a = base_launch(key);
    a.mode = mode;
    a.bsk = reinterpret_cast<const Tw*>(reinterpret_cast<const char*>(key->d_bsk) +
                                        (size_t)index * ggsw_words(key) * (key->plan->mod.dp ? 8 : sizeof(Tw)));
    a.in0 = s0.ptr<const uint64_t>();
    a.in1 = (mode == BOOT_CMUX) ? s1.ptr<const uint64_t>() : nullptr;
    a.out = so.ptr<uint64_t>();
    a.batch = batch;
    FHEB_REQUIRE(a.out != a.in0 && a.out != a.in1, "out must not alias an input");
    FHEB_TRY(boot_dispatch(key, a, s));
    FHEB_TRY(so.finish());
    return sync_if_staged(s, {&s0, &s1, &so});
}

int fheb_external_product_batch(const fheb_boot_key* key, uint32_t index, const uint64_t* glwe, uint64_t* out,
                                size_t batch, void* stream) {
    return one_step_entry(key, index, BOOT_EXT, glwe, nullptr, out, batch, stream);
}

int fheb_cmux_batch(const fheb_boot_key* key, uint32_t index, const uint64_t* ct0, const uint64_t* ct1, uint64_t* out,
                    size_t batch, void* stream) {
    return one_step_entry(key, index, BOOT_CMUX, ct0, ct1, out, batch, stream);
}

int fheb_blind_rotate_batch(const fheb_boot_key* key_, const uint64_t* lwe, const uint64_t* test_poly, uint64_t* out,
                            size_t batch, void* stream) {
    FHEB_TRY(check_key(key_));
    const BootKey* key = reinterpret_cast<const BootKey*>(key_);
    if (batch == 0) return FHEB_OK;
    FHEB_REQUIRE(lwe != nullptr && test_poly != nullptr && out != nullptr, "ciphertext pointers must not be null");
    cudaStream_t s = (cudaStream_t)stream;
    if (all_host({lwe, test_poly, out})) {
        if (device_list().size() > 1 && batch >= 2 * BOOT_SPREAD_MIN)
            return run_on_devices(batch, [&](int device, size_t first, size_t n) {
                const BootKey* kd = boot_key_on_device(key, device);
                if (!kd) return (int)FHEB_ERR_NATIVE;
                return run_host_pipeline(n, {{lwe + first * ((size_t)key->n + 1), ((size_t)key->n + 1) * 8, 0, true, false},
                                             {test_poly, 0, (size_t)key->plan->degree * 8, true, false},
                                             {out + first * glwe_words(key), glwe_words(key) * 8, 0, false, true}},
                                         [&](void* const* d, size_t, size_t m, cudaStream_t ps) {
                                             return blind_rotate_device(kd, static_cast<const uint64_t*>(d[0]), static_cast<const uint64_t*>(d[1]),
                                                                        static_cast<uint64_t*>(d[2]), m, ps);
                                         },
                                         BOOT_HOST_CHUNK);
            });
        return run_host_pipeline(batch, {{lwe, ((size_t)key->n + 1) * 8, 0, true, false}, {test_poly, 0, (size_t)key->plan->degree * 8, true, false},
                                         {out, glwe_words(key) * 8, 0, false, true}},
                                 [&](void* const* d, size_t, size_t n, cudaStream_t ps) {
                                     return blind_rotate_device(key, static_cast<const uint64_t*>(d[0]), static_cast<const uint64_t*>(d[1]),
                                                                static_cast<uint64_t*>(d[2]), n, ps);
                                 },
                                 BOOT_HOST_CHUNK);
    }
    Staged sl, st, so;
    FHEB_TRY(sl.bind(lwe, batch * ((size_t)key->n + 1) * 8, true, false, s));
    FHEB_TRY(st.bind(test_poly, (size_t)key->plan->degree * 8, true, false, s));
    FHEB_TRY(so.bind(out, batch * glwe_words(key) * 8, false, true, s));
    FHEB_TRY(blind_rotate_device(key, sl.ptr<const uint64_t>(), st.ptr<const uint64_t>(), so.ptr<uint64_t>(), batch, s));
    FHEB_TRY(so.finish());
    return sync_if_staged(s, {&sl, &st, &so});
}

int fheb_sample_extract_batch(const fheb_boot_key* key_, const uint64_t* glwe, uint64_t* out, size_t batch, void* stream) {
    FHEB_TRY(check_key(key_));
    const BootKey* key = reinterpret_cast<const BootKey*>(key_);
    if (batch == 0) return FHEB_OK;
    FHEB_REQUIRE(glwe != nullptr && out != nullptr, "ciphertext pointers must not be null");
    cudaStream_t s = (cudaStream_t)stream;
    Staged sg, so;
    FHEB_TRY(sg.bind(glwe, batch * glwe_words(key) * 8, true, false, s));
    FHEB_TRY(so.bind(out, batch * ((size_t)key->k * key->plan->degree + 1) * 8, false, true, s));
    FHEB_TRY(sample_extract_device(key, sg.ptr<const uint64_t>(), so.ptr<uint64_t>(), batch, s));
    FHEB_TRY(so.finish());
    return sync_if_staged(s, {&sg, &so});
}

int fheb_key_switch_batch(const fheb_boot_key* key_, const uint64_t* lwe, uint64_t* out, size_t batch, void* stream) {
    FHEB_TRY(check_key(key_));
    const BootKey* key = reinterpret_cast<const BootKey*>(key_);
    FHEB_REQUIRE(key->d_ksk != nullptr, "no key switching key has been set (fheb_boot_key_set_ksk)");
    if (batch == 0) return FHEB_OK;
    FHEB_REQUIRE(lwe != nullptr && out != nullptr, "ciphertext pointers must not be null");
    cudaStream_t s = (cudaStream_t)stream;
    Staged sl, so;
    FHEB_TRY(sl.bind(lwe, batch * ((size_t)key->k * key->plan->degree + 1) * 8, true, false, s));
    FHEB_TRY(so.bind(out, batch * ((size_t)key->ksk_n_out + 1) * 8, false, true, s));
    FHEB_TRY(key_switch_device(key, sl.ptr<const uint64_t>(), so.ptr<uint64_t>(), batch, s));
    FHEB_TRY(so.finish());
    return sync_if_staged(s, {&sl, &so});
}

int fheb_bootstrap_batch(const fheb_boot_key* key_, const uint64_t* lwe, const uint64_t* test_poly, uint64_t* out,
                         size_t batch, void* stream) {
    FHEB_TRY(check_key(key_));
    const BootKey* key = reinterpret_cast<const BootKey*>(key_);
    if (batch == 0) return FHEB_OK;
    FHEB_REQUIRE(lwe != nullptr && test_poly != nullptr && out != nullptr, "ciphertext pointers must not be null");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t ext_w = (size_t)key->k * key->plan->degree + 1;
    const size_t out_w = key->d_ksk ? (size_t)key->ksk_n_out + 1 : ext_w;
    if (all_host({lwe, test_poly, out})) {
        if (device_list().size() > 1 && batch >= 2 * BOOT_SPREAD_MIN)  // compute bound: worth spreading from a few hundred ciphertexts
            return run_on_devices(batch, [&](int device, size_t first, size_t n) {
                const BootKey* kd = boot_key_on_device(key, device);
                if (!kd) return (int)FHEB_ERR_NATIVE;
                return run_host_pipeline(n, {{lwe + first * ((size_t)key->n + 1), ((size_t)key->n + 1) * 8, 0, true, false},
                                             {test_poly, 0, (size_t)key->plan->degree * 8, true, false},
                                             {out + first * out_w, out_w * 8, 0, false, true}},
                                         [&](void* const* d, size_t, size_t m, cudaStream_t ps) {
                                             return bootstrap_chain_device(kd, static_cast<const uint64_t*>(d[0]), static_cast<const uint64_t*>(d[1]),
                                                                           static_cast<uint64_t*>(d[2]), m, ps);
                                         },
                                         BOOT_HOST_CHUNK);
            });
        return run_host_pipeline(batch, {{lwe, ((size_t)key->n + 1) * 8, 0, true, false}, {test_poly, 0, (size_t)key->plan->degree * 8, true, false},
                                         {out, out_w * 8, 0, false, true}},
                                 [&](void* const* d, size_t, size_t n, cudaStream_t ps) {
                                     return bootstrap_chain_device(key, static_cast<const uint64_t*>(d[0]), static_cast<const uint64_t*>(d[1]),
                                                                   static_cast<uint64_t*>(d[2]), n, ps);
                                 },
                                 BOOT_HOST_CHUNK);
    }
    Staged sl, st, so;
    FHEB_TRY(sl.bind(lwe, batch * ((size_t)key->n + 1) * 8, true, false, s));
    FHEB_TRY(st.bind(test_poly, (size_t)key->plan->degree * 8, true, false, s));
    FHEB_TRY(so.bind(out, batch * out_w * 8, false, true, s));
    FHEB_TRY(bootstrap_chain_device(key, sl.ptr<const uint64_t>(), st.ptr<const uint64_t>(), so.ptr<uint64_t>(), batch, s));
    FHEB_TRY(so.finish());
    return sync_if_staged(s, {&sl, &st, &so});
}

int fheb_make_test_poly(const fheb_ntt_plan* plan, int kind, uint64_t arg0, uint64_t arg1, uint64_t* out_host) {
    // Host-side set-up, integer arithmetic exactly as written in cpp/src/bootstrap_engine.cpp:57-77
    // (default test polynomial) and :725-779 (lookup tables); u64 products wrap as they do there.
    FHEB_REQUIRE(plan != nullptr && out_host != nullptr, "plan and out must not be null");
    const NttPlan* p = reinterpret_cast<const NttPlan*>(plan);
    const uint64_t N = p->degree, q = p->modulus;
    if (kind == 3) {
        const uint64_t t = arg0 > 0 ? arg0 : 4;
        const uint64_t delta = q / t;
        for (uint64_t i = 0; i < N; ++i) out_host[i] = (((i * t) / (2 * N)) * delta) % q;
        return FHEB_OK;
    }
    FHEB_REQUIRE(kind >= 0 && kind <= 2, "unknown lookup table kind %d", kind);
    const uint64_t in_mod = (kind == 2) ? arg1 : arg0;
    const uint64_t out_mod = (kind == 2) ? 2 : arg0;
    FHEB_REQUIRE(in_mod != 0 && out_mod != 0, "lookup table moduli must be non-zero");
    const uint64_t delta_out = q / out_mod;
    for (uint64_t i = 0; i < N; ++i) {
        uint64_t x = ((i * in_mod + N) / (2 * N)) % in_mod;
        uint64_t y;
        if (kind == 0) y = x;
        else if (kind == 1) y = (arg0 - x) % arg0;
        else y = (x >= arg0) ? 1 : 0;
        out_host[i] = ((y % out_mod) * delta_out) % q;
    }
    return FHEB_OK;
}

}  // extern "C"
