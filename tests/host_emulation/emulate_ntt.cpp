// CPU emulation of the CUDA transform passes (no GPU needed).
//
// The device passes in node-fhe-accelerate_b200/csrc/ntt_core.cuh are written as
// host/device functions of (thread id, thread count).  Passes are separated by block
// barriers and every work item reads and writes only its own positions, so running the
// items of one pass sequentially on the host is equivalent to the barrier-synchronised
// device execution.  This program runs them that way for every supported degree and both
// range-tracking modes and checks the results word-for-word against the C oracle
// (oracle/fhe_oracle.c, which is pinned to the reference).  It validates index math,
// twiddle ordering, swizzling, bit-reversed I/O and the range bookkeeping of both arithmetic modes (integer / FP64); it says
// nothing about device-only code (launch geometry, PTX, memory spaces).
//
// Build + run: see tests/test_host_emulation.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../node-fhe-accelerate_b200/csrc/ntt_plan.hpp"
#include "../../oracle/fhe_oracle.h"

using namespace fheb;

// PK: plan key (the degree's own pass split, or an alternative one with its own table - ntt_core.cuh)
template <int L, int DP, int PASS, int PK = L>
static void run_fwd(uint32_t threads, uint32_t polys, const uint64_t* gin, uint64_t* gout, uint64_t* smem,
                    const Tw* tw, const ModQ& m) {
    constexpr int P = Plan<PK>::P;
    if constexpr (PASS < P) {
        for (uint32_t tid = 0; tid < threads; ++tid) {
            if constexpr (P == 1) fwd_pass<L, DP, PASS, IO_GLOBAL, IO_GLOBAL, true, false, 0, false, false, PK>(tid, threads, polys, gin, gout, smem, tw, m);
            else if constexpr (PASS == 0) fwd_pass<L, DP, PASS, IO_GLOBAL, IO_SMEM, true, false, 0, false, false, PK>(tid, threads, polys, gin, gout, smem, tw, m);
            else if constexpr (PASS == P - 1) fwd_pass<L, DP, PASS, IO_SMEM, IO_GLOBAL, true, false, 0, false, false, PK>(tid, threads, polys, gin, gout, smem, tw, m);
            else fwd_pass<L, DP, PASS, IO_SMEM, IO_SMEM, true, false, 0, false, false, PK>(tid, threads, polys, gin, gout, smem, tw, m);
        }
        run_fwd<L, DP, PASS + 1, PK>(threads, polys, gin, gout, smem, tw, m);
    }
}

template <int L, int DP, int PASS, int PK = L>
static void run_inv(uint32_t threads, uint32_t polys, const uint64_t* gin, uint64_t* gout, uint64_t* smem,
                    const Tw* tw, const Tw& ninv, const ModQ& m) {
    constexpr int P = Plan<PK>::P;
    if constexpr (PASS >= 0) {
        for (uint32_t tid = 0; tid < threads; ++tid) {
            if constexpr (P == 1) inv_pass<L, DP, PASS, IO_GLOBAL, IO_GLOBAL, true, 0, 1, false, false, PK>(tid, threads, polys, gin, gout, smem, tw, ninv, m);
            else if constexpr (PASS == P - 1) inv_pass<L, DP, PASS, IO_GLOBAL, IO_SMEM, true, 0, 1, false, false, PK>(tid, threads, polys, gin, gout, smem, tw, ninv, m);
            else if constexpr (PASS == 0) inv_pass<L, DP, PASS, IO_SMEM, IO_GLOBAL, true, 0, 1, false, false, PK>(tid, threads, polys, gin, gout, smem, tw, ninv, m);
            else inv_pass<L, DP, PASS, IO_SMEM, IO_SMEM, true, 0, 1, false, false, PK>(tid, threads, polys, gin, gout, smem, tw, ninv, m);
        }
        run_inv<L, DP, PASS - 1, PK>(threads, polys, gin, gout, smem, tw, ninv, m);
    }
}

// the device twiddle heap as raw words: (w, w') pairs, or one double per entry in DP mode
static std::vector<uint64_t> heap_words(const uint64_t* table, uint32_t L, uint64_t q, int mode, int key = 0) {
    if (is_u32(mode)) return build_heap_table_u32(table, L, q, key);
    if (mode == MODE_DP) return build_heap_table_dp(table, L, q);
    const std::vector<Tw> h = build_heap_table(table, L, q, key);
    std::vector<uint64_t> w(h.size() * 2);
    std::memcpy(w.data(), h.data(), w.size() * 8);
    return w;
}

static uint64_t pick_prime(int L, int mode) {
    // q == 1 (mod 2N), coprime to 2^64-1 where possible
    if (mode == MODE_U32) return (L <= 5) ? 97ULL + 0 * L : 132120577ULL;  // 2N | q - 1: 97 up to N = 16 ... see check()
    const bool lazy = mode == MODE_DP;
    if (lazy) return (L <= 10) ? 1099511678977ULL /* 41 bit */ : 132120577ULL /* 27 bit, 2N | q-1 up to 2^20 */;
    return 4611686018326724609ULL;  // 62 bit
}

template <int L, int DP, int PKF = L, int PKI = L>
static int check(uint32_t threads, uint32_t polys, uint64_t q_override = 0) {
    const uint32_t N = 1u << L;
    uint64_t q = q_override ? q_override : pick_prime(L, DP);
    if (L <= 3 && DP == MODE_INT) q = 4611686018326724609ULL;
    std::vector<uint64_t> fwd(N), inv(N);
    uint64_t sc[3];
    if (orc_precompute_twiddles(N, q, fwd.data(), inv.data(), sc) != 0) {
        std::printf("L=%d: prime %llu not NTT-friendly\n", L, (unsigned long long)q);
        return 1;
    }
    ModQ m = make_modq(q);
    if (!is_u32(DP) && (int)(m.dp != 0) != DP) {
        std::printf("L=%d: mode mismatch\n", L);
        return 1;
    }
    if (is_u32(DP) && (q >= (1ULL << U32_QBITS) || m.mu32 == 0)) {
        std::printf("L=%d: prime %llu too large for the 32-bit mode\n", L, (unsigned long long)q);
        return 1;
    }
    const std::vector<uint64_t> hfw = heap_words(fwd.data(), L, q, DP, PKF == L ? 0 : PKF), hiw = heap_words(inv.data(), L, q, DP, PKI == L ? 0 : PKI);
    const Tw* hf = reinterpret_cast<const Tw*>(hfw.data());
    const Tw* hi = reinterpret_cast<const Tw*>(hiw.data());
    const Tw ninv = is_u32(DP) ? Tw{sc[2], (sc[2] << 32) / q}
                    : DP == MODE_DP ? Tw{double_to_bits((double)sc[2]), 0} : Tw{sc[2], shoup_companion(sc[2], q)};
    std::mt19937_64 rng(1234 + L);
    std::vector<uint64_t> x((size_t)polys * N), ref, got((size_t)polys * N), smem((size_t)polys * N);
    int bad = 0;
    for (int variant = 0; variant < 3; ++variant) {
        for (auto& v : x) v = (variant == 0) ? rng() % q : (variant == 1 ? rng() : q - 1);  // canonical / unreduced / extreme
        ref = x;
        orc_forward_ntt_batch(ref.data(), polys, N, q, fwd.data());
        std::fill(smem.begin(), smem.end(), 0xDEADBEEFDEADBEEFULL);
        run_fwd<L, DP, 0, PKF>(threads, polys, x.data(), got.data(), smem.data(), hf, m);
        if (std::memcmp(ref.data(), got.data(), got.size() * 8) != 0) {
            std::printf("L=%d dp=%d variant=%d threads=%u polys=%u: FORWARD mismatch\n", L, DP, variant, threads, polys);
            ++bad;
        }
        ref = x;
        orc_inverse_ntt_batch(ref.data(), polys, N, q, inv.data(), sc[2]);
        run_inv<L, DP, Plan<PKI>::P - 1, PKI>(threads, polys, x.data(), got.data(), smem.data(), hi, ninv, m);
        if (std::memcmp(ref.data(), got.data(), got.size() * 8) != 0) {
            std::printf("L=%d dp=%d variant=%d threads=%u polys=%u: INVERSE mismatch\n", L, DP, variant, threads, polys);
            ++bad;
        }
    }
    return bad;
}

template <int L>
static int check_all() {
    int bad = 0;
    bad += check<L, MODE_INT>(64, 1);
    bad += check<L, MODE_DP>(64, 1);
    bad += check<L, MODE_INT>(96, 3);  // thread count not dividing the work, several polynomials per block
    bad += check<L, MODE_DP>(32, 2);
    // 32-bit mode (q < 2^27): the published rows' prime, the largest 27-bit prime with 2^15 | q - 1 (range bookkeeping at
    // the top of the 32 q window), and tiny primes where they are NTT-friendly
    bad += check<L, MODE_U32>(64, 1, 132120577ULL);
    bad += check<L, MODE_U32>(96, 3, 133857281ULL);  // the largest prime below 2^27 with 2^15 | q - 1
    if constexpr (L <= 4) bad += check<L, MODE_U32>(32, 2, 97ULL);
    if constexpr (Plan<L>::P > 1) {  // pair mode (two polynomials per work-buffer slot): even, odd and single batches
        bad += check<L, MODE_U32P>(64, 4, 132120577ULL);
        bad += check<L, MODE_U32P>(96, 3, 133857281ULL);
        bad += check<L, MODE_U32P>(32, 1, 132120577ULL);
    }
    if constexpr (L == 14) {  // the three-pass splits of the plain kernels: forward 5+4+5 (key 80), inverse 4+5+5 (key 79), and 5+5+4 (key 78)
        bad += check<L, MODE_INT, PLAN_KEY_ALT14, PLAN_KEY_ALT14_INV>(96, 2);
        bad += check<L, MODE_U32, PLAN_KEY_ALT14_U32, PLAN_KEY_ALT14_INV>(64, 1, 133857281ULL);
        bad += check<L, MODE_INT, 78, 78>(64, 1);
    }
    if constexpr (L < 14) bad += check_all<L + 1>();
    return bad;
}

static int check_arith() {
    // reduce128 / mulmod / reduce64 against __int128 arithmetic on edge and random operands
    std::mt19937_64 rng(99);
    const uint64_t primes[] = {17ULL, 97ULL, 132120577ULL, 1099511678977ULL, 1125899906826241ULL,
                               1152921504606584833ULL, 4611686018326724609ULL, 0xFFFFFFFFFFFFFFC5ULL, 0x8000000000000011ULL};
    int bad = 0;
    for (uint64_t q : primes) {
        ModQ m = make_modq(q);
        for (int i = 0; i < 200000; ++i) {
            uint64_t a = rng(), b = rng();
            if (i < 8) { a = (i & 1) ? q - 1 : 0; b = (i & 2) ? q - 1 : 1; if (i & 4) { a = ~0ULL; b = ~0ULL; } }
            uint64_t e = (uint64_t)(((u128)a * b) % q);
            if (mulmod_any(a, b, m) != e) { ++bad; break; }
            if (reduce64(a, m) != a % q) { ++bad; break; }
            if (q < (1ULL << 63)) {
                uint64_t w = b % q, wp = shoup_companion(w, q);
                uint64_t r = shoup_lazy(a, w, wp, m);
                if (r >= 2 * q || r % q != (uint64_t)(((u128)a * w) % q)) { ++bad; break; }
            }
            uint64_t ac = a % q, bc = b % q;
            if (addmod_canon(ac, bc, q) != (uint64_t)(((u128)ac + bc) % q)) { ++bad; break; }
            if (submod_canon(ac, bc, q) != (uint64_t)(((u128)ac + q - bc) % q)) { ++bad; break; }
        }
    }
    for (uint64_t q : {17ULL, 97ULL, 132120577ULL, 133857281ULL, (1ULL << 27) - 39}) {  // 32-bit primitives of MODE_U32
        ModQ m = make_modq(q);
        for (int i = 0; i < 200000; ++i) {
            uint32_t x = (uint32_t)rng();
            if (i < 4) x = (i & 1) ? 0xFFFFFFFFu : (uint32_t)(32 * q - 1);
            const uint32_t w = (uint32_t)(rng() % q), wp = (uint32_t)(((uint64_t)w << 32) / q);
            const uint32_t r = shoup32(x, w, wp, (uint32_t)q);
            if (r >= 2 * q || r % q != (uint64_t)x * w % q) { ++bad; break; }
            const uint32_t l = lazy32(x, m);
            if (l >= 2 * q || l % q != x % q) { ++bad; break; }
        }
    }
    if (bad) std::printf("arithmetic primitive mismatch (%d moduli)\n", bad);
    return bad;
}

int main() {
    int bad = check_arith();
    bad += check_all<2>();
    if (bad == 0) std::printf("HOST EMULATION OK\n");
    return bad ? 1 : 0;
}
