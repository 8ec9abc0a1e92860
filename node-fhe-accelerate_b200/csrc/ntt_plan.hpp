// Host-side construction of the device twiddle tables (plain C++, no CUDA).
//
// Input: the reference's tables exactly as NTTProcessor::precompute_twiddles builds them
// (reference cpp/src/ntt_processor.cpp:168-208): table[i] = root^i mod q, i < N, natural
// exponent order, plus inv_n.  The kernels never index them as the reference does
// (`table[j * (n / group_size)]`, ntt_processor.cpp:286); instead the entries the network
// actually touches (only exponents below N/2) are re-ordered into the block-ordered heap
// described in ntt_core.cuh and paired with their Shoup companions.
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

// heap[2^s + b] = table[bitrev_s(b) << (L-1-s)]  (entry 0 unused)
inline std::vector<Tw> build_heap_table(const uint64_t* table, uint32_t L, uint64_t q) {
    const uint32_t N = 1u << L;
    std::vector<Tw> heap(N);
    heap[0] = Tw{0, 0};
    for (uint32_t s = 0; s < L; ++s) {
        for (uint32_t b = 0; b < (1u << s); ++b) {
            const uint32_t e = bitrev_c(b, (int)s) << (L - 1 - s);
            const uint64_t w = table[e] % q;
            heap[(1u << s) + b] = Tw{w, shoup_companion(w, q)};
        }
    }
    return heap;
}

// DP mode (q < 2^42): the same heap holding each twiddle as a double, 8 bytes per entry
inline std::vector<uint64_t> build_heap_table_dp(const uint64_t* table, uint32_t L, uint64_t q) {
    const uint32_t N = 1u << L;
    std::vector<uint64_t> heap(N);
    heap[0] = 0;
    for (uint32_t s = 0; s < L; ++s) {
        for (uint32_t b = 0; b < (1u << s); ++b) {
            const uint32_t e = bitrev_c(b, (int)s) << (L - 1 - s);
            heap[(1u << s) + b] = double_to_bits((double)(table[e] % q));
        }
    }
    return heap;
}

}  // namespace fheb
