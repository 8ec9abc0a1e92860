// 64-bit modular arithmetic primitives for the B200 kernels.
//
// Everything here is exact integer arithmetic producing the same residues as the
// reference's `static_cast<__uint128_t>(a) * b % q`, `mod_add` and `mod_sub`
// (reference cpp/src/modular_arithmetic.cpp:122-153, cpp/src/ntt_processor.cpp:299-300),
// only computed without division: Shoup multiplication for fixed multiplicands
// (twiddles, N^-1), a Moeller-Granlund 128-by-64 reduction with a precomputed reciprocal
// for data x data products, and lazy (Harvey) butterflies whose value ranges are tracked
// at compile time.
//
// The header compiles under nvcc (device code) and under g++ (host emulation used by
// tests/test_host_emulation.py to validate index math and range tracking without a GPU).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FHEB_HD __host__ __device__ __forceinline__
#else
#define FHEB_HD inline
#endif

namespace fheb {

typedef unsigned __int128 u128;

struct Tw {  // one twiddle: value and its Shoup companion floor(w * 2^64 / q)
    uint64_t w;
    uint64_t wp;
};

// Per-modulus constants (built on the host by make_modq()).
struct ModQ {
    uint64_t q;
    uint64_t q2;     // 2q (valid when q < 2^63)
    uint64_t mu;     // floor(2^64 / q): Barrett constant for one-word reduction
    uint64_t dn;     // q << sh, normalised divisor (top bit set)
    uint64_t v;      // Moeller-Granlund reciprocal of dn: floor((2^128 - 1) / dn) - 2^64
    uint32_t sh;     // clz(q)
    uint32_t lazy;   // 1 when q < 2^46: butterflies skip intermediate reductions
};

FHEB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((u128)a * b) >> 64);
#endif
}

FHEB_HD void mul128(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    u128 p = (u128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}

// x * w mod q, lazily: result in [0, 2q) for ANY x < 2^64, given w < q and wp = floor(w*2^64/q).
FHEB_HD uint64_t shoup_lazy(uint64_t x, uint64_t w, uint64_t wp, uint64_t q) {
    uint64_t h = mulhi64(x, wp);
    return x * w - h * q;
}

FHEB_HD uint64_t csub(uint64_t x, uint64_t m) {  // x in [0, 2m) -> [0, m)
    return x >= m ? x - m : x;
}

// One-word Barrett: any x < 2^64 -> canonical [0, q).
FHEB_HD uint64_t reduce64(uint64_t x, const ModQ& m) {
    uint64_t r = x - mulhi64(x, m.mu) * m.q;  // in [0, 2q)
    return csub(r, m.q);
}

// 128-by-64 remainder, exact, for (hi:lo) < q * 2^64 (Moeller & Granlund, "Improved
// division by invariant integers", algorithm 4, on the normalised divisor dn = q << sh).
FHEB_HD uint64_t reduce128(uint64_t hi, uint64_t lo, const ModQ& m) {
    // normalise the dividend by the same shift
    uint64_t u1 = m.sh ? ((hi << m.sh) | (lo >> (64 - m.sh))) : hi;
    uint64_t u0 = lo << m.sh;
    uint64_t qh, ql;
    mul128(m.v, u1, qh, ql);
    // (qh:ql) += (u1:u0)
    ql += u0;
    qh += u1 + (ql < u0 ? 1 : 0);
    qh += 1;
    uint64_t r = u0 - qh * m.dn;
    if (r > ql) r += m.dn;
    if (r >= m.dn) r -= m.dn;
    return r >> m.sh;
}

// a * b mod q for canonical a, b (< q).
FHEB_HD uint64_t mulmod(uint64_t a, uint64_t b, const ModQ& m) {
    uint64_t hi, lo;
    mul128(a, b, hi, lo);
    return reduce128(hi, lo, m);
}

// a * b mod q for ARBITRARY 64-bit a, b - what `(u128)a * b % q` returns.
FHEB_HD uint64_t mulmod_any(uint64_t a, uint64_t b, const ModQ& m) {
    if (a >= m.q) a = reduce64(a, m);
    if (b >= m.q) b = reduce64(b, m);
    return mulmod(a, b, m);
}

// reference mod_add / mod_sub: inputs reduced first (modular_arithmetic.cpp:124-125,140-141)
FHEB_HD uint64_t addmod_canon(uint64_t a, uint64_t b, uint64_t q) {  // a, b < q, any q < 2^64
    uint64_t s = a + b;
    return (s < a || s >= q) ? s - q : s;
}
FHEB_HD uint64_t submod_canon(uint64_t a, uint64_t b, uint64_t q) {
    return a >= b ? a - b : q - (b - a);
}
FHEB_HD uint64_t canon_any(uint64_t x, const ModQ& m) {  // x % q for any x
    return x >= m.q ? reduce64(x, m) : x;
}

// ---- host-side constant builders (plan creation) ----------------------------------------
inline uint64_t shoup_companion(uint64_t w, uint64_t q) { return (uint64_t)((((u128)w) << 64) / q); }

inline ModQ make_modq(uint64_t q) {
    ModQ m;
    m.q = q;
    m.q2 = q << 1;
    m.mu = (uint64_t)((((u128)1) << 64) / q);
    m.sh = (uint32_t)__builtin_clzll(q);
    m.dn = q << m.sh;
    m.v = (uint64_t)((~(u128)0) / m.dn - (((u128)1) << 64));
    m.lazy = (q < (1ULL << 46)) ? 1u : 0u;
    return m;
}

}  // namespace fheb
