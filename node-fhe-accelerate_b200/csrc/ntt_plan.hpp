// Host-side construction of the device twiddle tables (plain C++, no CUDA).
//
// Input: the reference's tables exactly as NTTProcessor::precompute_twiddles builds them
// (reference cpp/src/ntt_processor.cpp:168-208): table[i] = root^i mod q, i < N, natural
// exponent order, plus inv_n.  The kernels never index them as the reference does
// (`table[j * (n / group_size)]`, ntt_processor.cpp:286); instead the entries the network
// actually touches (only exponents below N/2) are re-ordered pass by pass as described in
// ntt_core.cuh (tw_index) and paired with their Shoup companions.
#pragma once
#include <cstdint>
#include <vector>

#include "modarith.cuh"
#include "ntt_core.cuh"

namespace fheb {

inline uint32_t log2_exact(uint32_t n) {
    uint32_t l = 0;
    while ((1u << l) < n) ++l;
    return l;
}

// Entry order (see tw_index in ntt_core.cuh): pass by pass, [a][g][blk]; the twiddle of block
// b = (blk << a) + g at stage s = S0 + a is table[bitrev_s(b) << (L-1-s)].  N - 1 entries, padded to N.
// key: plan key (ntt_core.cuh; the degree's own split unless an alternative one is asked for)
template <class Put>
inline void for_each_twiddle(uint32_t L, Put put, int key = 0) {
    int P, R[5];
    plan_runtime(key ? key : (int)L, P, R);
    uint32_t idx = 0;
    int s0 = 0;
    for (int p = 0; p < P; ++p) {
        for (int a = 0; a < R[p]; ++a) {
            const int s = s0 + a;
            for (uint32_t g = 0; g < (1u << a); ++g)
                for (uint32_t blk = 0; blk < (1u << s0); ++blk) {
                    const uint32_t b = (blk << a) + g;
                    put(idx + ((((1u << a) - 1u) + g) << s0) + blk, bitrev_c(b, s) << (L - 1 - s));
                }
        }
        idx += ((1u << R[p]) - 1u) << s0;
        s0 += R[p];
    }
}

// Second copy of the LAST pass's entries, from offset N on, with the block index bit-reversed:
// entry N + (slot << S0) + bitrev_S0(blk) = entry off_last + (slot << S0) + blk.  The transform kernels walk the last
// pass's items in bit-reversed order (so that the reference-order stores / loads are coalesced); with this copy
// consecutive lanes read consecutive twiddles there too, instead of one 32-byte sector per lane (ntt_core.cuh:
// fwd_pass / inv_pass, BRTW).  The first N entries are unchanged (fused product and bootstrap kernels use them).
template <class T>
inline void append_bitrev_last_pass(std::vector<T>& out, uint32_t L, int key = 0) {
    int P, R[5];
    plan_runtime(key ? key : (int)L, P, R);
    const size_t N = (size_t)1 << L;
    out.resize(2 * N, T{});
    uint32_t off = 0;
    int s0 = 0;
    for (int p = 0; p + 1 < P; ++p) {
        off += ((1u << R[p]) - 1u) << s0;
        s0 += R[p];
    }
    if (s0 == 0) return;  // single-pass plans: one block per stage, nothing to reorder
    const uint32_t slots = (1u << R[P - 1]) - 1u;
    for (uint32_t slot = 0; slot < slots; ++slot)
        for (uint32_t blk = 0; blk < (1u << s0); ++blk)
            out[N + ((size_t)slot << s0) + bitrev_c(blk, s0)] = out[off + ((size_t)slot << s0) + blk];
}

// the same on raw table words (epw words per entry) of one sub-block table with room for 2 * 2^L entries
inline void append_bitrev_last_pass_words(uint64_t* words, uint32_t L, size_t epw) {
    int P, R[5];
    plan_runtime((int)L, P, R);
    const size_t N = (size_t)1 << L;
    uint32_t off = 0;
    int s0 = 0;
    for (int p = 0; p + 1 < P; ++p) {
        off += ((1u << R[p]) - 1u) << s0;
        s0 += R[p];
    }
    if (s0 == 0) return;
    const uint32_t slots = (1u << R[P - 1]) - 1u;
    for (uint32_t slot = 0; slot < slots; ++slot)
        for (uint32_t blk = 0; blk < (1u << s0); ++blk)
            for (size_t w = 0; w < epw; ++w)
                words[(N + ((size_t)slot << s0) + bitrev_c(blk, s0)) * epw + w] = words[(off + ((size_t)slot << s0) + blk) * epw + w];
}

inline std::vector<Tw> build_heap_table(const uint64_t* table, uint32_t L, uint64_t q, int key = 0) {
    std::vector<Tw> out((size_t)1 << L, Tw{0, 0});
    for_each_twiddle(L, [&](uint32_t at, uint32_t e) {
        const uint64_t w = table[e] % q;
        out[at] = Tw{w, shoup_companion(w, q)};
    }, key);
    append_bitrev_last_pass(out, L, key);
    return out;
}

// DP mode (q < 2^42): each twiddle as a double, 8 bytes per entry
inline std::vector<uint64_t> build_heap_table_dp(const uint64_t* table, uint32_t L, uint64_t q) {
    std::vector<uint64_t> out((size_t)1 << L, 0);
    for_each_twiddle(L, [&](uint32_t at, uint32_t e) { out[at] = double_to_bits((double)(table[e] % q)); });
    append_bitrev_last_pass(out, L);
    return out;
}

// U32 mode (q < 2^27): value and floor(w * 2^32 / q) packed into one 8-byte entry (low half: w)
constexpr int U32_QBITS = 27;
inline std::vector<uint64_t> build_heap_table_u32(const uint64_t* table, uint32_t L, uint64_t q, int key = 0) {
    std::vector<uint64_t> out((size_t)1 << L, 0);
    for_each_twiddle(L, [&](uint32_t at, uint32_t e) {
        const uint64_t w = table[e] % q;
        out[at] = w | (((w << 32) / q) << 32);
    }, key);
    append_bitrev_last_pass(out, L, key);
    return out;
}

// ---- degrees above 2^14: tables of the 2^D sub-blocks and of the top D stages (ntt_device.cuh) -------
// Sub-block h runs stages D..L-1 of the big network; its stage s' = s - D has blocks b' whose big-network
// block is b = (h << s') + b', i.e. twiddle table[bitrev_s(b) << (L-1-s)].  Same entry order as above.
template <class Put>
inline void for_each_sub_twiddle(uint32_t L, uint32_t D, uint32_t h, Put put) {
    const uint32_t LS = L - D;
    int P, R[5];
    plan_runtime((int)LS, P, R);
    uint32_t idx = 0;
    int s0 = 0;
    for (int p = 0; p < P; ++p) {
        for (int a = 0; a < R[p]; ++a) {
            const int sp = s0 + a, s = sp + (int)D;
            for (uint32_t g = 0; g < (1u << a); ++g)
                for (uint32_t blk = 0; blk < (1u << s0); ++blk) {
                    const uint32_t b = (h << sp) + (blk << a) + g;
                    put(idx + ((((1u << a) - 1u) + g) << s0) + blk, bitrev_c(b, s) << (L - 1 - s));
                }
        }
        idx += ((1u << R[p]) - 1u) << s0;
        s0 += R[p];
    }
}
// top D stages as one register pass (S0 = 0): entry ((1 << a) - 1) + g = table[bitrev_a(g) << (L-1-a)]
template <class Put>
inline void for_each_top_twiddle(uint32_t L, uint32_t D, Put put) {
    for (uint32_t a = 0; a < D; ++a)
        for (uint32_t g = 0; g < (1u << a); ++g) put(((1u << a) - 1u) + g, bitrev_c(g, (int)a) << (L - 1 - a));
}

// raw words of a device table: (w, w') pairs or doubles
inline void put_twiddle(std::vector<uint64_t>& words, size_t at, uint64_t w, uint64_t q, bool dp) {
    if (dp) words[at] = double_to_bits((double)w);
    else {
        words[2 * at] = w;
        words[2 * at + 1] = shoup_companion(w, q);
    }
}

// width (log2) of the last register pass: the bootstrapping key is stored [..][e][u] for position u*2^R + e
inline uint32_t last_pass_width(uint32_t L) {
    int P, R[5];
    plan_runtime((int)L, P, R);
    return (uint32_t)R[P - 1];
}

}  // namespace fheb
