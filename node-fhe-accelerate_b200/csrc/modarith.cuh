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
#include <cmath>
#include <cstdint>
#include <cstring>

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
    uint32_t dp;     // 1 when q < 2^42: transforms run on the FP64 pipe (exact integer arithmetic in doubles)
    uint32_t mu32;   // floor(2^32 / q) when q < 2^31 (32-bit mode of the plain transforms, q < 2^27), else 0
    uint32_t pad_;
    double qd;       // (double)q
    double qinv;     // 1.0 / q, rounded to nearest
    uint64_t nq;     // 2^64 - q: x*w - h*q == lo64(x*w + h*nq), one multiply-add chain (shoup_lazy)
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
FHEB_HD uint64_t shoup_lazy(uint64_t x, uint64_t w, uint64_t wp, const ModQ& m) {
    const uint64_t h = mulhi64(x, wp);
#if defined(__CUDA_ARCH__) && !defined(FHEB_EXP_PLAIN_SHOUP)
    // x*w - h*q == lo64(x*w + h*(2^64 - q)): with the negated modulus as multiplicand both low products accumulate into
    // ONE chain of 32-bit limbs - 2 IMAD.WIDE + 4 IMAD, no negation, no subtraction.  From `x * w - h * q` ptxas builds
    // two separate low products, negates one and patches the high words together with IMAD.IADD / IMAD.X: 1.7 more
    // multiply-pipe and 1.6 more ALU instructions per product (forward N=16384: 2784 -> 2488 instructions, +5.4 %).
    const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), h0 = (uint32_t)h, h1 = (uint32_t)(h >> 32);
    const uint32_t w0 = (uint32_t)w, w1 = (uint32_t)(w >> 32), n0 = (uint32_t)m.nq, n1 = (uint32_t)(m.nq >> 32);
    uint64_t r;
    asm("{\n\t.reg .u64 t;\n\t.reg .u32 tl, th;\n\t"
        "mul.wide.u32 t, %1, %2;\n\t"
        "mad.wide.u32 t, %3, %4, t;\n\t"
        "mov.b64 {tl, th}, t;\n\t"
        "mad.lo.u32 th, %5, %2, th;\n\t"
        "mad.lo.u32 th, %1, %6, th;\n\t"
        "mad.lo.u32 th, %7, %4, th;\n\t"
        "mad.lo.u32 th, %3, %8, th;\n\t"
        "mov.b64 %0, {tl, th};\n\t}"
        : "=l"(r)
        : "r"(h0), "r"(n0), "r"(x0), "r"(w0), "r"(h1), "r"(n1), "r"(x1), "r"(w1));
    return r;
#else
    return x * w - h * m.q;
#endif
}

FHEB_HD uint64_t csub(uint64_t x, uint64_t m) {  // x in [0, 2m) -> [0, m)
#if defined(__CUDA_ARCH__) && !defined(FHEB_EXP_PLAIN_CSUB)
    // subtract with the borrow chained through the carry flag and select on the final borrow: ptxas turns this into
    // IADD3 + IADD3.X + 2 SEL (4 instructions) where `x >= m ? x - m : x` costs two ISETP more
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32), ml = (uint32_t)m, mh = (uint32_t)(m >> 32);
    uint32_t dl, dh, b;
    asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, 0;" : "=r"(dl), "=r"(dh), "=r"(b) : "r"(xl), "r"(xh), "r"(ml), "r"(mh));
    return ((uint64_t)(b ? xh : dh) << 32) | (b ? xl : dl);
#else
    return x >= m ? x - m : x;
#endif
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
#if defined(__CUDA_ARCH__) && !defined(FHEB_EXP_PLAIN_CSUB)
    // a - b with the borrow kept in the carry flag; add q back when it borrowed (same words as the expression below)
    const uint32_t al = (uint32_t)a, ah = (uint32_t)(a >> 32), bl = (uint32_t)b, bh = (uint32_t)(b >> 32);
    uint32_t dl, dh, br;
    asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, 0;" : "=r"(dl), "=r"(dh), "=r"(br) : "r"(al), "r"(ah), "r"(bl), "r"(bh));
    const uint64_t d = ((uint64_t)dh << 32) | dl;
    return br ? d + q : d;
#else
    return a >= b ? a - b : q - (b - a);
#endif
}
FHEB_HD uint64_t canon_any(uint64_t x, const ModQ& m) {  // x % q for any x
    return x >= m.q ? reduce64(x, m) : x;
}

// 128-bit running sum of raw words (exact for any count below 2^64) and its reduction
FHEB_HD void acc128(uint64_t& lo, uint64_t& hi, uint64_t v) {
    lo += v;
    hi += (lo < v) ? 1 : 0;
}
FHEB_HD uint64_t fold128(uint64_t hi, uint64_t lo, const ModQ& m) {
    return reduce128(canon_any(hi, m), lo, m);  // hi reduced first so that (hi:lo) < q * 2^64
}

// ---- exact modular arithmetic on the FP64 pipe (moduli below 2^42) ------------------------
// On B200 a DFMA issues at 64 lanes/clk/SM, the same rate as a 32-bit IMAD, while the 64-bit
// Shoup product needs six IMAD.WIDE (half rate) plus four IMAD (tools/microbench/pipes.cu).  For
// q < 2^42 a residue fits a double exactly, and a*b mod q is six FP64 operations:
//     h = rn(a*b); l = fma(a, b, -h)   (h + l == a*b exactly)
//     k = rint(h / q)                  (fma with the 1.5*2^52 rounding constant)
//     r = fma(-k, q, h) + l            (exact: |h - k q| < q, l tiny)  =>  r == a*b (mod q), |r| < q
// Values are integers held in doubles, |v| < CAP_DP * q <= 2^50, congruent to the true residue;
// every word handed back to the caller is converted to the canonical residue, so results are
// bit-identical to the integer path and to the reference.
constexpr int DP_QBITS = 42;
constexpr double DP_MAGIC = 6755399441055744.0;      // 1.5 * 2^52: x + MAGIC - MAGIC == rint(x) for |x| < 2^51
constexpr double DP_TWO52 = 4503599627370496.0;      // 2^52

FHEB_HD double bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    std::memcpy(&d, &b, 8);
    return d;
#endif
}
FHEB_HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    std::memcpy(&b, &d, 8);
    return b;
#endif
}
FHEB_HD double dp_mul(double a, double b) {  // rounded product, never contracted into an fma
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double p = a * b;
    return p;
#endif
}
FHEB_HD double dp_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
FHEB_HD double dp_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double s = a + b;
    return s;
#endif
}
// integer v < 2^51 -> double, and back for 0 <= r < 2^51 (no I2F / F2I conversion instructions)
FHEB_HD double dp_from_uint(uint64_t v) { return dp_add(bits_to_double(v | 0x4330000000000000ull), -DP_TWO52); }
FHEB_HD double dp_from_sint(int64_t v) {  // |v| < 2^50
    return dp_add(bits_to_double((uint64_t)(v + 0x4338000000000000ll)), -DP_MAGIC);
}
FHEB_HD uint64_t dp_to_uint(double r) { return double_to_bits(dp_add(r, DP_TWO52)) & 0x000FFFFFFFFFFFFFull; }

// a*b mod q, |result| < q, for |a*b| < 2^51 * q
FHEB_HD double dp_mulmod(double a, double b, const ModQ& m) {
    const double h = dp_mul(a, b);
    const double l = dp_fma(a, b, -h);
    const double k = dp_add(dp_fma(h, m.qinv, DP_MAGIC), -DP_MAGIC);
    return dp_add(dp_fma(-k, m.qd, h), l);
}
// |s| < 2^51 -> |result| <= q/2 + 1
FHEB_HD double dp_reduce(double s, const ModQ& m) {
    const double k = dp_add(dp_fma(s, m.qinv, DP_MAGIC), -DP_MAGIC);
    return dp_fma(-k, m.qd, s);
}
// |r| < q -> canonical residue as an integer word
FHEB_HD uint64_t dp_canon_word(double r, const ModQ& m) { return dp_to_uint(r < 0.0 ? dp_add(r, m.qd) : r); }

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
    m.dp = (q < (1ULL << DP_QBITS)) ? 1u : 0u;
    m.mu32 = (q < (1ULL << 31)) ? (uint32_t)((1ULL << 32) / q) : 0u;
    m.pad_ = 0;
    m.qd = (double)q;
    m.qinv = 1.0 / (double)q;
    m.nq = 0 - q;
    return m;
}

}  // namespace fheb
