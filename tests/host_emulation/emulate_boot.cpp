// CPU emulation of the CUDA blind-rotation phases (no GPU needed).
//
// node-fhe-accelerate_b200/csrc/boot_core.cuh holds the phases of one CMux / external product
// as host/device functions of (thread id, thread count); the kernel separates them with block
// barriers.  Running the threads of one phase sequentially on the host is therefore
// equivalent.  This program drives them the way boot_kernel.cuh does (rotation table, initial
// monomial rotation, n steps) and checks the results word-for-word against the C oracle
// (oracle/fhe_oracle.c, pinned to the reference) for several shapes and both arithmetic modes
// (integer pipe for q < 2^62, FP64 pipe for q < 2^42).  It validates the algebra (pre-transformed key, transform-domain accumulation, fused
// passes, digit extraction, rotations); it says nothing about launch geometry or memory spaces.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../node-fhe-accelerate_b200/csrc/boot_core.cuh"
#include "../../node-fhe-accelerate_b200/csrc/ntt_plan.hpp"
#include "../../oracle/fhe_oracle.h"

using namespace fheb;

static std::vector<uint64_t> heap_words(const uint64_t* table, uint32_t L, uint64_t q, bool dp) {
    if (dp) return build_heap_table_dp(table, L, q);
    const std::vector<Tw> h = build_heap_table(table, L, q);
    std::vector<uint64_t> w(h.size() * 2);
    std::memcpy(w.data(), h.data(), w.size() * 8);
    return w;
}

template <int L, bool DP, int KP1, int PH = 0>
static void run_step(uint32_t threads, const BootStep& s, const Tw* twf, const Tw* twi, const Tw& ninv, const ModQ& m) {
    if constexpr (PH < boot_phases<L>()) {
        for (uint32_t tid = 0; tid < threads; ++tid) boot_phase<L, DP, KP1, PH>(tid, threads, s, twf, twi, ninv, m);
        run_step<L, DP, KP1, PH + 1>(threads, s, twf, twi, ninv, m);
    }
}

template <int L, bool DP, int KP1>
static int check(uint32_t threads, uint32_t levels, uint32_t base_log, uint32_t n, uint64_t q, bool raw_inputs) {
    const uint32_t N = 1u << L, k = KP1 - 1, rows = KP1 * levels;
    std::vector<uint64_t> fwd(N), inv(N);
    uint64_t sc[3];
    if (orc_precompute_twiddles(N, q, fwd.data(), inv.data(), sc) != 0) {
        std::printf("L=%d: prime %llu not NTT-friendly\n", L, (unsigned long long)q);
        return 1;
    }
    const ModQ m = make_modq(q);
    if ((m.dp != 0) != DP) {
        std::printf("mode mismatch for q=%llu\n", (unsigned long long)q);
        return 1;
    }
    const std::vector<uint64_t> hfw = heap_words(fwd.data(), L, q, DP), hiw = heap_words(inv.data(), L, q, DP);
    const Tw* hf = reinterpret_cast<const Tw*>(hfw.data());
    const Tw* hi = reinterpret_cast<const Tw*>(hiw.data());
    const Tw ninv = DP ? Tw{double_to_bits((double)sc[2]), 0} : Tw{sc[2], shoup_companion(sc[2], q)};
    orc_boot_params p{N, k, n, base_log, levels, q, 4, fwd.data(), inv.data(), sc[2]};

    std::mt19937_64 rng(77 * L + KP1 + levels);
    const size_t gw = (size_t)KP1 * N, ggsw_w = (size_t)rows * KP1 * N;
    std::vector<uint64_t> bsk((size_t)n * ggsw_w);
    for (auto& v : bsk) v = rng() % q;
    // key upload: T(poly) in position order + Shoup companions, or doubles in DP mode (what bsk_pack_kernel does)
    constexpr size_t GE = DP ? 1 : 2;  // words per key element
    std::vector<uint64_t> g(bsk.size() * GE);
    {
        std::vector<uint64_t> t(N);
        const uint32_t rl = last_pass_width(L), E = 1u << rl, items = N >> rl;
        for (size_t poly = 0; poly < bsk.size() / N; ++poly) {  // poly = (ggsw, row, j)
            std::memcpy(t.data(), bsk.data() + poly * N, N * 8);
            orc_forward_ntt(t.data(), N, q, fwd.data());
            const size_t gi = poly / ((size_t)rows * KP1), row = (poly / KP1) % rows, j = poly % KP1;
            for (uint32_t pos = 0; pos < N; ++pos) {
                const uint64_t w = mulmod(t[bitrev_c(pos, L)], sc[2], m);  // times N^-1
                const uint32_t e = pos & (E - 1), u = pos >> rl;
                const size_t at = gi * ggsw_w + ((row * E + e) * items + u) * KP1 + j;
                if (DP) g[at] = double_to_bits((double)w);
                else { g[2 * at] = w; g[2 * at + 1] = shoup_companion(w, q); }
            }
        }
    }
    int bad = 0;
    std::vector<uint64_t> acc(gw), work((size_t)rows * N), diff(gw), out(gw), ref(gw);
    BootStep s{};
    s.diff_sub = nullptr;
    s.acc = acc.data();
    s.work = work.data();
    s.levels = levels;
    s.rows = rows;
    s.base_log = base_log;

    // --- external product and cmux with bsk[1 % n]
    std::vector<uint64_t> ct0(gw), ct1(gw);
    for (auto& v : ct0) v = raw_inputs ? rng() : rng() % q;
    for (auto& v : ct1) v = raw_inputs ? rng() : rng() % q;
    const size_t gi = 1 % n;
    s.ggsw = reinterpret_cast<const Tw*>(g.data() + gi * ggsw_w * GE);
    s.gout = out.data();
    s.rot = 0;
    diff = ct0;
    s.diff = diff.data();
    s.add_acc = 0;
    s.maybe_raw = 0;
    run_step<L, DP, KP1>(threads, s, hf, hi, ninv, m);
    orc_external_product(&p, ct0.data(), bsk.data() + gi * ggsw_w, ref.data());
    if (out != ref) { std::printf("L=%d dp=%d kp1=%d levels=%u: EXTERNAL PRODUCT mismatch\n", L, DP, KP1, levels); ++bad; }
    for (size_t i = 0; i < gw; ++i) acc[i] = canon_any(ct0[i], m);  // the kernel loads ct0 reduced
    for (size_t i = 0; i < gw; ++i) diff[i] = submod_canon(canon_any(ct1[i], m), canon_any(ct0[i], m), q);
    s.add_acc = 1;
    run_step<L, DP, KP1>(threads, s, hf, hi, ninv, m);
    orc_cmux(&p, bsk.data() + gi * ggsw_w, ct0.data(), ct1.data(), ref.data());
    if (out != ref) { std::printf("L=%d dp=%d kp1=%d levels=%u: CMUX mismatch\n", L, DP, KP1, levels); ++bad; }

    // --- blind rotation
    for (int trial = 0; trial < 3; ++trial) {
        std::vector<uint64_t> lwe(n + 1), test(N);
        for (auto& v : lwe) v = rng() % q;
        if (trial == 1) { lwe[0] = 0; lwe[n] = 0; if (n > 2) lwe[2] = q - 1; }  // zero rotations are skipped
        // raw rotation 2N (a_i = q - 1): the reference runs a CMux with diff = 0 that canonicalises an unreduced accumulator
        if (trial == 2) { lwe[0] = q - 1; for (uint32_t i = 1; i < n; ++i) lwe[i] = (i == 1 && n > 3) ? q / 3 : 0; }
        for (auto& v : test) v = (raw_inputs && trial != 1) ? rng() : rng() % q;
        std::fill(ref.begin(), ref.end(), 0);
        std::memcpy(ref.data() + (size_t)k * N, test.data(), N * 8);
        orc_blind_rotate(&p, ref.data(), lwe.data(), bsk.data());
        const uint32_t rb = lwe_rotation(lwe[n], true, N, q);
        for (size_t i = 0; i < gw; ++i) acc[i] = (i / N == k) ? rotated_at(test.data(), (uint32_t)(i % N), rb, N, m) : 0;
        s.diff = nullptr;
        s.add_acc = 1;
        s.gout = nullptr;
        s.maybe_raw = 1;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t rot = lwe_step_rotation(lwe[i], N, q);  // same decision as boot_kernel
            if (!((rot & (2u * N - 1u)) != 0 || (rot != 0 && s.maybe_raw != 0))) continue;
            s.rot = rot & (2u * N - 1u);
            s.ggsw = reinterpret_cast<const Tw*>(g.data() + (size_t)i * ggsw_w * GE);
            run_step<L, DP, KP1>(threads, s, hf, hi, ninv, m);
            s.maybe_raw = 0;
        }
        if (acc != ref) { std::printf("L=%d dp=%d kp1=%d levels=%u trial=%d: BLIND ROTATE mismatch\n", L, DP, KP1, levels, trial); ++bad; }
        // sample extraction
        std::vector<uint64_t> ext((size_t)k * N + 1), ext_ref((size_t)k * N + 1);
        orc_sample_extract(ref.data(), k, N, q, ext_ref.data());
        for (uint32_t i = 0; i < ext.size(); ++i) ext[i] = sample_extract_word(acc.data(), i, k, N, m);
        if (ext != ext_ref) { std::printf("L=%d: SAMPLE EXTRACT mismatch\n", L); ++bad; }
    }
    return bad;
}

int main() {
    const uint64_t QT = 1099511678977ULL, Q27 = 132120577ULL, Q62 = 4611686018326724609ULL;
    int bad = 0;
    bad += check<7, true, 2>(64, 3, 4, 8, QT, false);     // tests/golden/boot_n128_l3 shape
    bad += check<10, true, 2>(128, 1, 23, 6, QT, false);  // tfhe-128-fast shape (small n)
    bad += check<10, true, 2>(96, 2, 10, 3, QT, true);    // raw (unreduced) inputs, ragged thread count
    bad += check<5, false, 2>(32, 2, 8, 4, Q62, false);
    bad += check<6, false, 3>(32, 2, 7, 3, Q62, true);
    bad += check<8, false, 2>(64, 4, 15, 3, Q62, false);
    bad += check<9, true, 4>(64, 2, 6, 3, QT, false);
    bad += check<9, false, 3>(100, 1, 20, 3, Q62, false);
    bad += check<11, true, 2>(256, 2, 9, 2, Q27, false);
    bad += check<12, false, 2>(512, 1, 30, 2, Q62, false);
    if (bad == 0) std::printf("BOOT EMULATION OK\n");
    return bad ? 1 : 0;
}
