// Batched forward / inverse transform and fused polynomial multiplication for sm_100a.
//
// One thread block owns PPC whole polynomials in shared memory and walks the pass plan of
// ntt_core.cuh: the first pass reads coalesced from global memory straight into registers,
// the middle passes exchange through bank-swizzled shared memory, the last pass stores
// coalesced from registers (in the reference's bit-reversed order for the forward
// transform).  Blocks are persistent (grid = SMs x resident blocks) and walk the batch with
// a grid stride, prefetching their next polynomials into L2 with a bulk-async prefetch.
//
// C-ABI entry points here (include/fheb200.h): fheb_ntt_plan_*, fheb_ntt_forward_batch,
// fheb_ntt_inverse_batch, fheb_ntt_inverse_fwdnet_batch, fheb_polymul_batch.
#include "ntt_device.cuh"
#include "ntt_plan.hpp"
#include "plan.hpp"
#include "runtime.hpp"

namespace fheb {

int elementwise_device(int op, const uint64_t* a, const uint64_t* b, uint64_t scalar, uint64_t* r, size_t count, uint64_t q, cudaStream_t s);

// ---- host-side plan maths (mirrors the reference's constructor) --------------------------
static uint64_t h_mod_pow(uint64_t base, uint64_t exp, uint64_t mod) {  // ntt_processor.cpp:47-62
    uint64_t result = 1;
    base %= mod;
    while (exp > 0) {
        if (exp & 1) result = (uint64_t)(((u128)result * base) % mod);
        base = (uint64_t)(((u128)base * base) % mod);
        exp >>= 1;
    }
    return result;
}

static uint64_t h_mod_inverse(uint64_t a, uint64_t m) {  // ntt_processor.cpp:64-90 (signed Euclid)
    if (m == 1) return 0;
    int64_t m0 = (int64_t)m, x0 = 0, x1 = 1;
    int64_t as = (int64_t)(a % m), ms = (int64_t)m;
    while (as > 1) {
        int64_t quo = as / ms, t = ms;
        ms = as % ms;
        as = t;
        t = x0;
        x0 = x1 - quo * x0;
        x1 = t;
    }
    if (x1 < 0) x1 += m0;
    return (uint64_t)x1;
}

static bool h_find_primitive_root(uint32_t degree, uint64_t q, uint64_t* root) {  // ntt_processor.cpp:92-128
    const uint64_t two_n = (uint64_t)degree * 2;
    if ((q - 1) % two_n != 0) return false;
    const uint64_t exponent = (q - 1) / two_n;
    // The reference scans every g < q; a composite modulus (e.g. the tfhe-128-fast preset's
    // 2^40+1) makes that a ~10^12-step loop ending in the same exception, so the
    // scan is capped far beyond where any prime modulus finds its root.
    for (uint64_t g = 2; g < q && g < (1u << 20); g++) {
        const uint64_t omega = h_mod_pow(g, exponent, q);
        if (h_mod_pow(omega, two_n, q) == 1 && h_mod_pow(omega, degree, q) == q - 1) {
            *root = omega;
            return true;
        }
    }
    return false;
}

static int validate_degree_modulus(uint32_t degree, uint64_t modulus) {
    // messages follow NTTProcessor::NTTProcessor, cpp/src/ntt_processor.cpp:141-153
    FHEB_REQUIRE(degree > 0 && (degree & (degree - 1)) == 0, "Polynomial degree must be a power of 2");
    FHEB_REQUIRE(degree >= 4 && degree <= 65536, "Polynomial degree must be between 4 and 65536");
    FHEB_REQUIRE((modulus & 1) != 0, "Modulus must be odd");
    FHEB_REQUIRE(modulus > 2 && modulus < (1ULL << 62), "modulus must be below 2^62 on this backend");
    return FHEB_OK;
}

template <class T>
static int upload_heap(const std::vector<T>& heap, Tw** out) {
    FHEB_CUDA(cudaMalloc(out, heap.size() * sizeof(T)));
    FHEB_CUDA(cudaMemcpy(*out, heap.data(), heap.size() * sizeof(T), cudaMemcpyHostToDevice));
    return FHEB_OK;
}

static int plan_finish(NttPlan* p) {
    const uint64_t q = p->modulus;
    p->device = ctx().device;
    p->logn = log2_exact(p->degree);
    p->mod = make_modq(q);
    p->ninv = p->mod.dp ? Tw{double_to_bits((double)(p->inv_n % q)), 0} : Tw{p->inv_n % q, shoup_companion(p->inv_n % q, q)};
    p->unit_first = (p->fwd_table[0] % q == 1) && (p->inv_table[0] % q == 1);
    FHEB_REQUIRE(p->unit_first, "twiddle tables must start with 1 (root^0)");
    p->one = p->mod.dp ? Tw{double_to_bits(1.0), 0} : Tw{1, shoup_companion(1, q)};
    if (p->logn > 14) {  // sub-block tables + top-stage tables (ntt_device.cuh, degrees above 2^14)
        const uint32_t L = p->logn, D = L - 14;
        const bool dp = p->mod.dp != 0;
        const size_t epw = dp ? 1 : 2, sub = (size_t)1 << 14;
        p->top = D;
        for (int dir = 0; dir < 2; ++dir) {
            const uint64_t* table = dir ? p->inv_table.data() : p->fwd_table.data();
            // per sub-block: `sub` entries pass by pass + `sub` entries holding the last pass again with the block
            // index bit-reversed (append_bitrev_last_pass: coalesced twiddle loads in the bit-reversed last pass)
            std::vector<uint64_t> words(((size_t)(2 * sub) << D) * epw, 0), topw(8 * epw, 0);
            for (uint32_t h = 0; h < (1u << D); ++h) {
                for_each_sub_twiddle(L, D, h, [&](uint32_t at, uint32_t e) { put_twiddle(words, (size_t)h * 2 * sub + at, table[e] % q, q, dp); });
                append_bitrev_last_pass_words(words.data() + (size_t)h * 2 * sub * epw, 14, epw);
            }
            for_each_top_twiddle(L, D, [&](uint32_t at, uint32_t e) { put_twiddle(topw, at, table[e] % q, q, dp); });
            FHEB_TRY(upload_heap(words, dir ? &p->d_inv : &p->d_fwd));
            FHEB_TRY(upload_heap(topw, dir ? &p->d_top_inv : &p->d_top_fwd));
        }
        return FHEB_OK;
    }
    if (q < (1ULL << U32_QBITS)) {  // 32-bit mode of the plain transforms and the fused product (ntt_core.cuh MODE_U32)
        FHEB_TRY(upload_heap(build_heap_table_u32(p->fwd_table.data(), p->logn, q), &p->d_fwd32));
        FHEB_TRY(upload_heap(build_heap_table_u32(p->inv_table.data(), p->logn, q), &p->d_inv32));
        const uint64_t ni = p->inv_n % q;
        p->ninv32 = Tw{ni, (ni << 32) / q};
        if (p->logn == 14) FHEB_TRY(upload_heap(build_heap_table_u32(p->fwd_table.data(), p->logn, q, PLAN_KEY_ALT14_U32), &p->d_fwd32_alt));
        if (p->logn == 14) FHEB_TRY(upload_heap(build_heap_table_u32(p->inv_table.data(), p->logn, q, PLAN_KEY_ALT14_INV), &p->d_inv32_alt));
    }
    if (p->mod.dp) {  // FP64 mode: one double per twiddle
        FHEB_TRY(upload_heap(build_heap_table_dp(p->fwd_table.data(), p->logn, q), &p->d_fwd));
        FHEB_TRY(upload_heap(build_heap_table_dp(p->inv_table.data(), p->logn, q), &p->d_inv));
    } else {
        FHEB_TRY(upload_heap(build_heap_table(p->fwd_table.data(), p->logn, q), &p->d_fwd));
        FHEB_TRY(upload_heap(build_heap_table(p->inv_table.data(), p->logn, q), &p->d_inv));
        if (p->logn == 14) FHEB_TRY(upload_heap(build_heap_table(p->fwd_table.data(), p->logn, q, PLAN_KEY_ALT14), &p->d_fwd_alt));
        if (p->logn == 14) FHEB_TRY(upload_heap(build_heap_table(p->inv_table.data(), p->logn, q, PLAN_KEY_ALT14_INV), &p->d_inv_alt));
    }
    return FHEB_OK;
}

static void plan_free(NttPlan* p) {
    if (!p) return;
    for (auto& kv : p->replicas) plan_free(kv.second);
    if (p->d_fwd) cudaFree(p->d_fwd);
    if (p->d_inv) cudaFree(p->d_inv);
    if (p->d_top_fwd) cudaFree(p->d_top_fwd);
    if (p->d_top_inv) cudaFree(p->d_top_inv);
    if (p->d_fwd32) cudaFree(p->d_fwd32);
    if (p->d_inv32) cudaFree(p->d_inv32);
    if (p->d_fwd_alt) cudaFree(p->d_fwd_alt);
    if (p->d_fwd32_alt) cudaFree(p->d_fwd32_alt);
    if (p->d_inv_alt) cudaFree(p->d_inv_alt);
    if (p->d_inv32_alt) cudaFree(p->d_inv32_alt);
    delete p;
}

const NttPlan* plan_on_device(const NttPlan* p, int device) {
    if (p->device == device) return p;
    std::lock_guard<std::mutex> lock(p->replica_mutex);
    auto it = p->replicas.find(device);
    if (it != p->replicas.end()) return it->second;
    NttPlan* r = new NttPlan();  // same host-side tables, device tables rebuilt on the current device
    r->degree = p->degree;
    r->modulus = p->modulus;
    r->psi = p->psi;
    r->psi_inv = p->psi_inv;
    r->inv_n = p->inv_n;
    r->fwd_table = p->fwd_table;
    r->inv_table = p->inv_table;
    if (plan_finish(r) != FHEB_OK || r->device != device) {
        plan_free(r);
        return nullptr;  // the error text is set
    }
    p->replicas[device] = r;
    return r;
}

// ---- kernel dispatch --------------------------------------------------------------------
#if !defined(FHEB_U32_BIG_THREADS)
#define FHEB_U32_BIG_THREADS 256
#endif
template <int L, int DP>
struct Geometry {  // threads per block, polynomials per block
    static constexpr int UNITS = (L <= 9) ? (1024 >> L) : (L == 10 ? 2 : 1);  // work-buffer units (N slots each) per block
    static constexpr int PPC = (DP == MODE_U32P) ? 2 * UNITS : UNITS;          // pair mode: two polynomials per unit
#if defined(FHEB_EXP_R3)
    static constexpr int THREADS = (L <= 11) ? 128 : (L == 12 ? 256 : 1024);
#else
    // MODE_U32 at N >= 8192: 4-byte slots halve the footprint (64 KB at N = 16384), and 256-thread blocks let TWO blocks
    // share an SM's registers and shared memory, so one block's global loads and stores overlap the other's butterflies
    // (with one 512-thread block per SM the memory phases and the compute phases of a polynomial simply add up)
    static constexpr int THREADS = (L <= 11) ? 128 : (L == 12 ? 256 : (DP == MODE_U32 ? FHEB_U32_BIG_THREADS : 512));
#endif
    static constexpr size_t SLOT = smem_slot_bytes<DP>();
    static constexpr size_t SMEM = (Plan<L>::P > 1) ? (size_t)UNITS * (1u << L) * SLOT : 0;
};

// Attribute set-up and occupancy of a kernel are looked up once per (kernel, shared memory, threads, device) and cached:
// at batch 1 the two driver calls cost more than the transform itself.
template <class K>
static int configure(K kernel, size_t smem, int threads, int* blocks_per_sm) {
    struct Entry {
        const void* fn;
        size_t smem;
        int threads, device, bps;
    };
    static std::mutex mu;
    static std::vector<Entry> cache;
    const int device = ctx().device;
    const void* fn = reinterpret_cast<const void*>(kernel);
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : cache)
        if (e.fn == fn && e.smem == smem && e.threads == threads && e.device == device) {
            *blocks_per_sm = e.bps;
            return FHEB_OK;
        }
    if (smem > 48 * 1024) FHEB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FHEB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, threads, smem));
    if (*blocks_per_sm < 1) return set_error(FHEB_ERR_NATIVE, "kernel does not fit on an SM (smem %zu)", smem);
    cache.push_back({fn, smem, threads, device, *blocks_per_sm});
    return FHEB_OK;
}

static unsigned stream_grid_for(size_t work_items, int threads, int blocks_per_sm) {
    const size_t want = (work_items + threads - 1) / threads;
    const size_t cap = (size_t)ctx().sm_count * blocks_per_sm;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

static unsigned persistent_grid(size_t work_groups, int blocks_per_sm) {
    const size_t resident = (size_t)ctx().sm_count * (size_t)blocks_per_sm;
    return (unsigned)(work_groups < resident ? work_groups : resident);
}

enum { DIR_FWD = 0, DIR_INV = 1, DIR_INV_FWDNET = 2 };

// Where the bulk-async (TMA) landing buffer is the default: exactly the (degree, mode, direction) cells that measured
// faster in BOTH runs of tools/prof_tma.py (profiles/r02_tma_vs_plain.txt: +4..7 % forward at N <= 1024, +13 % inverse at
// N = 4096 in 32-bit mode; it LOSES 7..12 % for the inverse at N = 256 and 3..4 % at N = 2048, which stay on plain loads).
template <int L, int DP>
constexpr bool TMA_DEFAULT(bool inverse) {
    if (L == 6) return true;
    if (L == 8) return !inverse || DP == MODE_U32;  // 32-bit inverse: +1.2 / +2.7 / +5.0 % in three runs
    if (L == 10) return !inverse || DP != MODE_INT;
    if (L == 12) return inverse ? true : DP == MODE_INT;
    // N = 8192: 32-bit forward +15 % when measured first, 64-bit integer +3.5 / +4.4 %; re-measured with the leaner Shoup product
    // (profiles/r02_tma_vs_plain.txt, third table): the 64-bit forward now LOSES 6.4 % with the landing buffer (0.1579 vs 0.1480 ms
    // per 2^24 coefficients) and runs plain loads, the 64-bit inverse still gains 4.3 %.  32-bit N = 16384: no difference
    if (L == 13) return (DP == MODE_INT && inverse) || (!inverse && DP == MODE_U32);
    return false;
}

template <int L, int DP>
static int launch_transform(const NttPlan* p, int dir, const uint64_t* in, uint64_t* out, size_t batch, cudaStream_t s) {
    using G = Geometry<L, DP>;
    const Tw* d_fwd = is_u32(DP) ? p->d_fwd32 : p->d_fwd;
    const Tw* d_inv = is_u32(DP) ? p->d_inv32 : p->d_inv;
    const Tw ninv = is_u32(DP) ? p->ninv32 : p->ninv;
    const size_t groups = (batch + G::PPC - 1) / G::PPC;
    int bps = 0;
    // Bulk-async (TMA) landing buffer for the next group's words: multi-pass plans up to N = 4096 (two buffers fit several
    // blocks per SM), 16-byte aligned input, and only where it measured faster (FHEB_TMA=0/1 overrides: experiments).
    if constexpr (Plan<L>::P > 1 && DP != MODE_U32P) {
        // (N = 8192 / 16384 in the 32-bit mode: the 4-byte work buffer leaves room for a landing buffer of raw 8-byte
        // words, 96 / 192 KB in all; one 512-thread block per SM then, against two 256-thread blocks without it)
        const char* tma_s = getenv("FHEB_TMA");  // read per call: the parity suite forces both settings
        const int tma_env = tma_s ? atoi(tma_s) : -1;
        const bool tma = (tma_env < 0 ? TMA_DEFAULT<L, DP>(dir == DIR_INV) : tma_env != 0) && (reinterpret_cast<uintptr_t>(in) & 15u) == 0 && dir != DIR_INV_FWDNET;
        if (tma) {
            constexpr int T = (L >= 13) ? 512 : G::THREADS;
            // work | landing (raw 8-byte words; N = 16384 with 8-byte slots: the first LAND_PART_WORDS only) | mbarrier
            constexpr size_t SMEM_TMA = G::SMEM + ((L == 14 && G::SLOT == 8) ? (size_t)LAND_PART_WORDS : (size_t)G::PPC * (1u << L)) * 8 + 16;
            if (dir == DIR_INV) {
                auto k = ntt_inverse_kernel<L, DP, T, G::PPC, true>;
                FHEB_TRY(configure(k, SMEM_TMA, T, &bps));
                k<<<persistent_grid(groups, bps), T, SMEM_TMA, s>>>(in, out, batch, d_inv, ninv, p->mod);
            } else {
                auto k = ntt_forward_kernel<L, DP, T, G::PPC, false, true>;
                FHEB_TRY(configure(k, SMEM_TMA, T, &bps));
                k<<<persistent_grid(groups, bps), T, SMEM_TMA, s>>>(in, out, batch, d_fwd, ninv, p->mod);
            }
            FHEB_CHECK_LAUNCH();
            count_launch();
            return FHEB_OK;
        }
    }
    if (dir == DIR_INV) {
        if constexpr (L == 14 && (DP == MODE_INT || DP == MODE_U32)) {  // three passes of 4 + 5 + 5 stages (plan key 79)
            const Tw* alt = (DP == MODE_U32) ? p->d_inv32_alt : p->d_inv_alt;
            const bool want = (DP == MODE_INT) ? !getenv("FHEB_NO_ALT_PLAN") : (getenv("FHEB_U32_INV_ALT") != nullptr);  // 32-bit inverse: experiment switch
            if (alt != nullptr && want) {
                auto k = ntt_inverse_kernel<L, DP, G::THREADS, G::PPC, false, PLAN_KEY_ALT14_INV>;
                FHEB_TRY(configure(k, G::SMEM, G::THREADS, &bps));
                k<<<persistent_grid(groups, bps), G::THREADS, G::SMEM, s>>>(in, out, batch, alt, ninv, p->mod);
                FHEB_CHECK_LAUNCH();
                count_launch();
                return FHEB_OK;
            }
        }
        auto k = ntt_inverse_kernel<L, DP, G::THREADS, G::PPC>;
        FHEB_TRY(configure(k, G::SMEM, G::THREADS, &bps));
        k<<<persistent_grid(groups, bps), G::THREADS, G::SMEM, s>>>(in, out, batch, d_inv, ninv, p->mod);
    } else if (dir == DIR_FWD) {
        if constexpr (L == 14 && (DP == MODE_INT || DP == MODE_U32)) {  // three passes of 5 + 5 + 4 stages (plan key 78) from the table built for them
            const Tw* alt = (DP == MODE_U32) ? p->d_fwd32_alt : p->d_fwd_alt;
            if (alt != nullptr && !getenv("FHEB_NO_ALT_PLAN")) {
                auto k = ntt_forward_kernel<L, DP, G::THREADS, G::PPC, false, false, (DP == MODE_U32 ? PLAN_KEY_ALT14_U32 : PLAN_KEY_ALT14)>;
                FHEB_TRY(configure(k, G::SMEM, G::THREADS, &bps));
                k<<<persistent_grid(groups, bps), G::THREADS, G::SMEM, s>>>(in, out, batch, alt, ninv, p->mod);
                FHEB_CHECK_LAUNCH();
                count_launch();
                return FHEB_OK;
            }
        }
        auto k = ntt_forward_kernel<L, DP, G::THREADS, G::PPC, false>;
        FHEB_TRY(configure(k, G::SMEM, G::THREADS, &bps));
        k<<<persistent_grid(groups, bps), G::THREADS, G::SMEM, s>>>(in, out, batch, d_fwd, ninv, p->mod);
    } else {
        auto k = ntt_forward_kernel<L, DP, G::THREADS, G::PPC, true>;
        FHEB_TRY(configure(k, G::SMEM, G::THREADS, &bps));
        k<<<persistent_grid(groups, bps), G::THREADS, G::SMEM, s>>>(in, out, batch, d_inv, ninv, p->mod);
    }
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

template <int L, int DP>
static int launch_polymul(const NttPlan* p, const uint64_t* a, const uint64_t* b, uint64_t* c, size_t batch,
                          cudaStream_t s) {
    using G = Geometry<L, DP>;
    constexpr bool STASH_GLOBAL = (L >= 14);  // two 128 KB operands do not fit in shared memory
    constexpr size_t SMEM = STASH_GLOBAL ? G::SMEM : 2 * (size_t)G::UNITS * (1u << L) * G::SLOT;  // work buffer | stash (also for single-pass plans)
    const size_t groups = (batch + G::PPC - 1) / G::PPC;
    int bps = 0;
    auto k = polymul_kernel<L, DP, G::THREADS, G::PPC, STASH_GLOBAL>;
    FHEB_TRY(configure(k, SMEM, G::THREADS, &bps));
    const unsigned grid = persistent_grid(groups, bps);
    uint64_t* stash = nullptr;
    if (STASH_GLOBAL) {
        // per-block scratch for T(a); small enough to stay L2 resident (grid x 128 KB)
        FHEB_CUDA(cudaMallocAsync(&stash, (size_t)grid * G::UNITS * (1u << L) * 8, s));
    }
    if (is_u32(DP)) k<<<grid, G::THREADS, SMEM, s>>>(a, b, c, batch, p->d_fwd32, p->d_inv32, p->ninv32, p->mod, stash);
    else k<<<grid, G::THREADS, SMEM, s>>>(a, b, c, batch, p->d_fwd, p->d_inv, p->ninv, p->mod, stash);
    FHEB_CHECK_LAUNCH();
    count_launch();
    if (stash) FHEB_CUDA(cudaFreeAsync(stash, s));
    return FHEB_OK;
}

// Degrees 2^15 and 2^16: top stages + 2^D sub-transforms of 2^14 (two launches, one scratch pass).
template <int DP>
static int launch_transform_big(const NttPlan* p, int dir, const uint64_t* in, uint64_t* out, size_t batch, cudaStream_t s) {
    constexpr int LS = 14;
    using G = Geometry<LS, DP>;
    const uint32_t D = p->top;
    FHEB_REQUIRE(dir != DIR_INV_FWDNET, "the forward-network inverse is limited to degrees up to 16384 on this backend");
    const size_t N = (size_t)1 << p->logn, subs = batch << D;
    uint64_t* tmp = nullptr;
    FHEB_CUDA(cudaMallocAsync(&tmp, batch * N * 8, s));
    const unsigned top_grid = stream_grid_for(batch * (N >> D), 256, 8);
    int bps = 0;
    int rc = FHEB_OK;
    if (dir == DIR_FWD) {
        if (D == 1) ntt_top_kernel<1, DP, false><<<top_grid, 256, 0, s>>>(in, tmp, batch, p->logn, p->d_top_fwd, p->ninv, p->mod);
        else ntt_top_kernel<2, DP, false><<<top_grid, 256, 0, s>>>(in, tmp, batch, p->logn, p->d_top_fwd, p->ninv, p->mod);
        auto k = ntt_forward_sub_kernel<LS, DP, G::THREADS, false>;
        rc = configure(k, G::SMEM, G::THREADS, &bps);
        if (rc == FHEB_OK) k<<<persistent_grid(subs, bps), G::THREADS, G::SMEM, s>>>(tmp, out, subs, D, p->d_fwd, p->ninv, p->mod);
    } else {
        auto k = ntt_inverse_sub_kernel<LS, DP, G::THREADS>;
        rc = configure(k, G::SMEM, G::THREADS, &bps);
        if (rc == FHEB_OK) {
            k<<<persistent_grid(subs, bps), G::THREADS, G::SMEM, s>>>(in, tmp, subs, D, p->d_inv, p->one, p->mod);
            if (D == 1) ntt_top_kernel<1, DP, true><<<top_grid, 256, 0, s>>>(tmp, out, batch, p->logn, p->d_top_inv, p->ninv, p->mod);
            else ntt_top_kernel<2, DP, true><<<top_grid, 256, 0, s>>>(tmp, out, batch, p->logn, p->d_top_inv, p->ninv, p->mod);
        }
    }
    if (rc == FHEB_OK && cudaGetLastError() != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "large-degree transform launch failed");
    count_launch(2);
    cudaFreeAsync(tmp, s);
    return rc;
}

// Product for degrees above 2^14: T(a), T(b), pointwise, T^-1 as separate launches.
template <int DP>
static int launch_polymul_big(const NttPlan* p, const uint64_t* a, const uint64_t* b, uint64_t* c, size_t batch, cudaStream_t s) {
    const size_t words = batch << p->logn;
    uint64_t* ta = nullptr;
    FHEB_CUDA(cudaMallocAsync(&ta, 2 * words * 8, s));
    uint64_t* tb = ta + words;
    int rc = launch_transform_big<DP>(p, DIR_FWD, a, ta, batch, s);
    if (rc == FHEB_OK) rc = launch_transform_big<DP>(p, DIR_FWD, b, tb, batch, s);
    if (rc == FHEB_OK) rc = elementwise_device(2 /* mul */, ta, tb, 0, ta, words, p->modulus, s);
    if (rc == FHEB_OK) rc = launch_transform_big<DP>(p, DIR_INV, ta, c, batch, s);
    cudaFreeAsync(ta, s);
    return rc;
}

// 32-bit kernels: two polynomials per work-buffer slot (MODE_U32P) for multi-pass plans whenever the batch has a pair
#define FHEB_DISPATCH_L(FN, L_, ...)                                                                                   \
    case L_:                                                                                                           \
        if (u32 && pair && Plan<L_>::P > 1) return FN<L_, (Plan<L_>::P > 1 ? MODE_U32P : MODE_U32)>(__VA_ARGS__);     \
        return u32 ? FN<L_, MODE_U32>(__VA_ARGS__) : dp ? FN<L_, MODE_DP>(__VA_ARGS__) : FN<L_, MODE_INT>(__VA_ARGS__);

// FHEB_NO_U32=1 (read per call: the parity tests run both ways) keeps moduli below 2^27 on the FP64-pipe kernels
static bool use_u32(const NttPlan* p) { return p->d_fwd32 != nullptr && p->logn <= 14 && getenv("FHEB_NO_U32") == nullptr; }

// FHEB_U32_PAIR=0/1 (read per call: parity tests, experiments) overrides where the 32-bit kernels pack two polynomials
// per 8-byte slot (MODE_U32P) instead of one per 4-byte slot (MODE_U32); defaults from profiles/r02_u32_modes.txt
// (tools/prof_u32_modes.sh): the plain transforms are faster one polynomial per 4-byte slot at every degree but the
// forward at N = 16384 (92.5 vs 88.9 us, while the inverse is 91.3 vs 97.3 us); the fused product gains 2..6 % from the pair
// layout at N >= 8192 (its two operands then fit shared memory as ONE buffer each) and loses below.
static bool u32_pair_wanted(bool product, uint32_t logn) {
    const char* e = getenv("FHEB_U32_PAIR");
    if (e) return atoi(e) != 0;
    return product && logn >= 13;
}

static int dispatch_transform(const NttPlan* p, int dir, const uint64_t* in, uint64_t* out, size_t batch, cudaStream_t s) {
    const bool dp = p->mod.dp != 0;
    const bool u32 = use_u32(p);
    const bool pair = batch >= 2 && u32_pair_wanted(false, p->logn);
    switch (p->logn) {
        FHEB_DISPATCH_L(launch_transform, 2, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 3, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 4, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 5, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 6, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 7, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 8, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 9, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 10, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 11, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 12, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 13, p, dir, in, out, batch, s)
        FHEB_DISPATCH_L(launch_transform, 14, p, dir, in, out, batch, s)
        case 15:
        case 16: return dp ? launch_transform_big<MODE_DP>(p, dir, in, out, batch, s) : launch_transform_big<MODE_INT>(p, dir, in, out, batch, s);
    }
    return set_error(FHEB_ERR_INVALID_PARAMETERS, "unsupported degree 2^%u", p->logn);
}

static int dispatch_polymul(const NttPlan* p, const uint64_t* a, const uint64_t* b, uint64_t* c, size_t batch, cudaStream_t s) {
    const bool dp = p->mod.dp != 0;
    const bool u32 = use_u32(p);
    const bool pair = batch >= 2 && u32_pair_wanted(true, p->logn);
    switch (p->logn) {
        FHEB_DISPATCH_L(launch_polymul, 2, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 3, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 4, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 5, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 6, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 7, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 8, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 9, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 10, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 11, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 12, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 13, p, a, b, c, batch, s)
        FHEB_DISPATCH_L(launch_polymul, 14, p, a, b, c, batch, s)
        case 15:
        case 16: return dp ? launch_polymul_big<MODE_DP>(p, a, b, c, batch, s) : launch_polymul_big<MODE_INT>(p, a, b, c, batch, s);
    }
    return set_error(FHEB_ERR_INVALID_PARAMETERS, "unsupported degree 2^%u", p->logn);
}

// Device-pointer entry points used by other translation units (bootstrap, tally).
int ntt_forward_device(const NttPlan* p, const uint64_t* in, uint64_t* out, size_t batch, cudaStream_t s) {
    return batch ? dispatch_transform(p, DIR_FWD, in, out, batch, s) : FHEB_OK;
}
int ntt_inverse_device(const NttPlan* p, const uint64_t* in, uint64_t* out, size_t batch, cudaStream_t s) {
    return batch ? dispatch_transform(p, DIR_INV, in, out, batch, s) : FHEB_OK;
}

static int transform_entry(const fheb_ntt_plan* plan, int dir, const uint64_t* in, uint64_t* out, size_t batch, void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(plan != nullptr, "plan must not be null");
    if (batch == 0) return FHEB_OK;
    FHEB_REQUIRE(in != nullptr && out != nullptr, "coefficient pointers must not be null");
    const NttPlan* p = reinterpret_cast<const NttPlan*>(plan);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = batch * (size_t)p->degree * 8;
    if (all_host({in, out})) {  // host buffers: chunked copy-in / transform / copy-out pipeline
        const size_t row = (size_t)p->degree * 8;
        if (spread_over_devices(batch, bytes))  // one share, one host thread and one pipeline per configured GPU
            return run_on_devices(batch, [&](int device, size_t first, size_t n) {
                const NttPlan* pd = plan_on_device(p, device);
                if (!pd) return (int)FHEB_ERR_NATIVE;
                const uint64_t* in_d = in + first * p->degree;
                uint64_t* out_d = out + first * p->degree;
                return run_host_pipeline(n, {{in_d, row, 0, true, false}, {out_d, row, 0, false, true}},
                                         [&](void* const* d, size_t, size_t m, cudaStream_t ps) {
                                             return dispatch_transform(pd, dir, static_cast<const uint64_t*>(d[0]), static_cast<uint64_t*>(d[1]), m, ps);
                                         });
            });
        return run_host_pipeline(batch, {{in, row, 0, true, false}, {out, row, 0, false, true}},
                                 [&](void* const* d, size_t, size_t n, cudaStream_t ps) {
                                     return dispatch_transform(p, dir, static_cast<const uint64_t*>(d[0]), static_cast<uint64_t*>(d[1]), n, ps);
                                 });
    }
    Staged sin, sout;
    FHEB_TRY(sin.bind(in, bytes, true, false, s));
    if (out == in) FHEB_TRY(sout.bind_alias(sin, true));
    else FHEB_TRY(sout.bind(out, bytes, false, true, s));
    FHEB_TRY(dispatch_transform(p, dir, sin.ptr<const uint64_t>(), sout.ptr<uint64_t>(), batch, s));
    FHEB_TRY(sin.finish());
    FHEB_TRY(sout.finish());
    return sync_if_staged(s, {&sin, &sout});
}

}  // namespace fheb

using namespace fheb;

extern "C" {

int fheb_ntt_plan_create(uint32_t degree, uint64_t modulus, fheb_ntt_plan** out) {
    FHEB_REQUIRE(out != nullptr, "out must not be null");
    *out = nullptr;
    FHEB_TRY(ensure_ready());
    FHEB_TRY(validate_degree_modulus(degree, modulus));
    uint64_t psi = 0;
    // messages follow find_primitive_root, cpp/src/ntt_processor.cpp:103-105,127
    FHEB_REQUIRE((modulus - 1) % ((uint64_t)degree * 2) == 0, "Modulus is not NTT-friendly: q != 1 (mod 2N)");
    FHEB_REQUIRE(h_find_primitive_root(degree, modulus, &psi), "Could not find primitive root for given parameters");
    NttPlan* p = new NttPlan();
    p->degree = degree;
    p->modulus = modulus;
    p->psi = psi;
    p->psi_inv = h_mod_inverse(psi, modulus);
    p->inv_n = h_mod_inverse(degree, modulus);
    p->fwd_table.resize(degree);
    p->inv_table.resize(degree);
    p->fwd_table[0] = 1;
    p->inv_table[0] = 1;
    for (uint32_t i = 1; i < degree; i++) {  // cpp/src/ntt_processor.cpp:188-202
        p->fwd_table[i] = (uint64_t)(((u128)p->fwd_table[i - 1] * psi) % modulus);
        p->inv_table[i] = (uint64_t)(((u128)p->inv_table[i - 1] * p->psi_inv) % modulus);
    }
    int rc = plan_finish(p);
    if (rc != FHEB_OK) {
        fheb_ntt_plan_destroy(reinterpret_cast<fheb_ntt_plan*>(p));
        return rc;
    }
    *out = reinterpret_cast<fheb_ntt_plan*>(p);
    return FHEB_OK;
}

int fheb_ntt_plan_create_with_tables(uint32_t degree, uint64_t modulus, const uint64_t* fwd_table,
                                     const uint64_t* inv_table, uint64_t inv_n, fheb_ntt_plan** out) {
    FHEB_REQUIRE(out != nullptr, "out must not be null");
    *out = nullptr;
    FHEB_TRY(ensure_ready());
    FHEB_TRY(validate_degree_modulus(degree, modulus));
    FHEB_REQUIRE(fwd_table != nullptr && inv_table != nullptr, "twiddle tables must not be null");
    NttPlan* p = new NttPlan();
    p->degree = degree;
    p->modulus = modulus;
    p->inv_n = inv_n;
    p->fwd_table.assign(fwd_table, fwd_table + degree);
    p->inv_table.assign(inv_table, inv_table + degree);
    int rc = plan_finish(p);
    if (rc != FHEB_OK) {
        fheb_ntt_plan_destroy(reinterpret_cast<fheb_ntt_plan*>(p));
        return rc;
    }
    *out = reinterpret_cast<fheb_ntt_plan*>(p);
    return FHEB_OK;
}

int fheb_ntt_plan_destroy(fheb_ntt_plan* plan) {
    plan_free(reinterpret_cast<NttPlan*>(plan));
    return FHEB_OK;
}

int fheb_ntt_plan_get_tables(const fheb_ntt_plan* plan, uint64_t* fwd_table, uint64_t* inv_table, uint64_t scalars[3]) {
    FHEB_REQUIRE(plan != nullptr, "plan must not be null");
    const NttPlan* p = reinterpret_cast<const NttPlan*>(plan);
    if (fwd_table) std::memcpy(fwd_table, p->fwd_table.data(), p->degree * 8);
    if (inv_table) std::memcpy(inv_table, p->inv_table.data(), p->degree * 8);
    if (scalars) {
        scalars[0] = p->psi;
        scalars[1] = p->psi_inv;
        scalars[2] = p->inv_n;
    }
    return FHEB_OK;
}

uint32_t fheb_ntt_plan_degree(const fheb_ntt_plan* plan) { return plan ? reinterpret_cast<const NttPlan*>(plan)->degree : 0; }
uint64_t fheb_ntt_plan_modulus(const fheb_ntt_plan* plan) { return plan ? reinterpret_cast<const NttPlan*>(plan)->modulus : 0; }

int fheb_ntt_forward_batch(const fheb_ntt_plan* plan, const uint64_t* in, uint64_t* out, size_t batch, void* stream) {
    return transform_entry(plan, DIR_FWD, in, out, batch, stream);
}
int fheb_ntt_inverse_batch(const fheb_ntt_plan* plan, const uint64_t* in, uint64_t* out, size_t batch, void* stream) {
    return transform_entry(plan, DIR_INV, in, out, batch, stream);
}
int fheb_ntt_inverse_fwdnet_batch(const fheb_ntt_plan* plan, const uint64_t* in, uint64_t* out, size_t batch, void* stream) {
    return transform_entry(plan, DIR_INV_FWDNET, in, out, batch, stream);
}

int fheb_polymul_batch(const fheb_ntt_plan* plan, const uint64_t* a, const uint64_t* b, uint64_t* c, size_t batch, void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(plan != nullptr, "plan must not be null");
    if (batch == 0) return FHEB_OK;
    FHEB_REQUIRE(a != nullptr && b != nullptr && c != nullptr, "coefficient pointers must not be null");
    const NttPlan* p = reinterpret_cast<const NttPlan*>(plan);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = batch * (size_t)p->degree * 8;
    if (all_host({a, b, c})) {
        const size_t row = (size_t)p->degree * 8;
        if (spread_over_devices(batch, bytes))
            return run_on_devices(batch, [&](int device, size_t first, size_t n) {
                const NttPlan* pd = plan_on_device(p, device);
                if (!pd) return (int)FHEB_ERR_NATIVE;
                const size_t off = first * p->degree;
                return run_host_pipeline(n, {{a + off, row, 0, true, false}, {b + off, row, 0, true, false}, {c + off, row, 0, false, true}},
                                         [&](void* const* d, size_t, size_t m, cudaStream_t ps) {
                                             return dispatch_polymul(pd, static_cast<const uint64_t*>(d[0]), static_cast<const uint64_t*>(d[1]),
                                                                     static_cast<uint64_t*>(d[2]), m, ps);
                                         });
            });
        return run_host_pipeline(batch, {{a, row, 0, true, false}, {b, row, 0, true, false}, {c, row, 0, false, true}},
                                 [&](void* const* d, size_t, size_t n, cudaStream_t ps) {
                                     return dispatch_polymul(p, static_cast<const uint64_t*>(d[0]), static_cast<const uint64_t*>(d[1]),
                                                             static_cast<uint64_t*>(d[2]), n, ps);
                                 });
    }
    Staged sa, sb, sc;
    FHEB_TRY(sa.bind(a, bytes, true, false, s));
    if (b == a) FHEB_TRY(sb.bind_alias(sa, false));
    else FHEB_TRY(sb.bind(b, bytes, true, false, s));
    if (c == a) FHEB_TRY(sc.bind_alias(sa, true));
    else if (c == b) FHEB_TRY(sc.bind_alias(sb, true));
    else FHEB_TRY(sc.bind(c, bytes, false, true, s));
    FHEB_TRY(dispatch_polymul(p, sa.ptr<const uint64_t>(), sb.ptr<const uint64_t>(), sc.ptr<uint64_t>(), batch, s));
    FHEB_TRY(sa.finish());
    FHEB_TRY(sb.finish());
    FHEB_TRY(sc.finish());
    return sync_if_staged(s, {&sa, &sb, &sc});
}

}  // extern "C"
