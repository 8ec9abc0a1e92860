// Radix-2^R register passes of the reference's transform, in "position space".
//
// The reference's forward transform (cpp/src/ntt_processor.cpp:262-311) is: bit-reverse the
// input, then for stage s = 0..L-1 (m = 2^s) do (A,B) <- (A + w*B, A - w*B) on entries
// (k+j, k+j+m) of the PERMUTED array with w = table[j * N/(2m)].  Writing i = bitrev(p) for
// the position of permuted entry p in the ORIGINAL array, stage s pairs original positions
// that differ in bit (L-1-s), and its twiddle depends only on the TOP s bits of i:
//     table index  j*N/(2m)  with  j = bitrev_s(i >> (L-s)).
// So the network can run on the data where it lies (no permutation): distance N/2 first,
// distance 1 last, one twiddle per contiguous block of N/2^s positions, and the result for
// permuted entry p ends up at position bitrev(p).  The device twiddle table is therefore
// stored block-ordered: block b of stage s uses table[bitrev_s(b) << (L-1-s)] (layout: tw_index).
// The inverse (ntt_processor.cpp:325-380) is the same pairs walked backwards with
// (A,B) <- (A + B, (A - B)*w), then the bit-reversal (which brings entry p back to position
// bitrev(bitrev(i)) = i, i.e. nothing to move) and the scaling by N^-1.
//
// Two arithmetic modes, chosen per modulus at plan creation (ModQ::dp):
//   integer (q < 2^62): values kept lazily in [0, K*q), Shoup products on the integer pipe;
//   DP      (q < 2^42): values are integers held in doubles, |v| < K*q, products on the FP64 pipe
//                       (modarith.cuh); register arrays and shared memory carry the double's bits.
// K is tracked at compile time (Harvey-style lazy butterflies); every function states the bound it
// needs and the bound it leaves.  Outputs handed back to the caller are always canonical, so
// results are bit-identical to the reference's in both modes.
#pragma once
#include "modarith.cuh"

namespace fheb {

// Bank swizzle for 8-byte words in shared memory.  A warp-wide 8-byte access is served in two
// wavefronts of 16 lanes, and 16 lanes hit 16 distinct word-banks iff the four index bits that vary
// across them map to independent vectors of GF(2)^4.  The passes vary, depending on their geometry,
// index bits {0,1,2,3}, {0,1,2,6}, {0,1,2,7}, {0,1,5,6}, any four consecutive bits (last pass,
// natural item order) or the four top bits (last pass, bit-reversed item order).  Mapping index
// bit j to alpha^j, alpha a primitive element of GF(16) (x^4 + x + 1), makes every one of those
// sets independent (any 4 consecutive powers of alpha are a basis; the mixed sets were checked
// exhaustively, tools/check_swizzle.py).  alpha^0..alpha^3 are the unit vectors, so the map only
// XORs a function of the high index bits into the low nibble: a bijection on every aligned 16-word
// group, linear over GF(2) (swz(a | b) == swz(a) ^ swz(b) for disjoint a, b).
FHEB_HD uint32_t parity32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(v) & 1u;
#else
    return (uint32_t)__builtin_popcount(v) & 1u;
#endif
}
FHEB_HD uint32_t swz(uint32_t i) {
    const uint32_t h = i >> 4;  // bank bit k = parity of the index bits j >= 4 whose alpha^j has bit k set
    return i ^ (parity32(h & 0xF59u) | (parity32(h & 0x1EBu) << 1) | (parity32(h & 0x3D6u) << 2) | (parity32(h & 0x7ACu) << 3));
}

// The same construction for 4-byte words (MODE_U32 keeps one 32-bit word per shared-memory slot): a warp-wide 4-byte
// access is ONE wavefront of 32 lanes over 32 banks, so the five index bits that vary across the lanes must map to
// independent vectors of GF(2)^5: index bit j -> beta^j, beta a primitive element of GF(32) (x^5 + x^3 + 1; the other
// primitive polynomials leave conflicts in some pass geometry - tools/check_swizzle.py checks all plans exhaustively).
FHEB_HD uint32_t swz32(uint32_t i) {
    const uint32_t h = i >> 5;
    return i ^ (parity32(h & 0x375u) | (parity32(h & 0x6EAu) << 1) | (parity32(h & 0x5D4u) << 2) | (parity32(h & 0x0DDu) << 3) |
                (parity32(h & 0x1BAu) << 4));
}

FHEB_HD constexpr uint32_t bitrev_c(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

FHEB_HD uint32_t bitrev_rt(uint32_t x, int bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(x) >> (32 - bits)) : 0u;
#else
    return bitrev_c(x, bits);
#endif
}

// Arithmetic mode of a kernel instantiation (the template parameter called DP for historical reasons):
//   MODE_INT (0)  q < 2^62: 64-bit words, Shoup products from IMAD.WIDE chains, values in [0, 4q)
//   MODE_DP  (1)  q < 2^42: integers held in doubles, products on the FP64 pipe, |v| < 128 q
//   MODE_U32 (2)  q < 2^27: 32-bit words (kept zero-extended in the 64-bit register/shared-memory slots), Shoup products
//                 from ONE IMAD.HI + two IMAD, values in [0, 32q) - 14 lazy stages need no conditional subtraction.
//                 132120577, the modulus of every published reference row, is such a prime.
//   MODE_U32P (3) the same arithmetic on TWO polynomials at once: slot i of the work buffer holds word i of polynomial A
//                 in its low half and word i of polynomial B in its high half.  Index math, swizzle, bank behaviour and
//                 the twiddles are shared, so shared-memory instructions, twiddle loads, address arithmetic and block
//                 barriers per polynomial are halved; a 64-bit register pair simply carries both words.
constexpr int MODE_INT = 0, MODE_DP = 1, MODE_U32 = 2, MODE_U32P = 3;
constexpr bool is_u32(int mode) { return mode == MODE_U32 || mode == MODE_U32P; }
// work-buffer units (slots arrays) for `polys` polynomials: pairs in MODE_U32P
template <int DP>
FHEB_HD constexpr uint32_t units_of(uint32_t polys) { return DP == MODE_U32P ? (polys + 1) / 2 : polys; }
FHEB_HD uint32_t lo32(uint64_t x) { return (uint32_t)x; }
FHEB_HD uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
FHEB_HD uint64_t pack32(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// Shared-memory work buffer: N slots per unit.  A slot is 8 bytes (a 64-bit word, a double, or a pair of 32-bit words)
// except in MODE_U32, where it is one 4-byte word: half the footprint, so two blocks of N = 16384 share an SM.
template <int DP>
constexpr uint32_t smem_slot_bytes() { return DP == MODE_U32 ? 4u : 8u; }
template <int DP>
FHEB_HD uint32_t swzm(uint32_t i) { return DP == MODE_U32 ? swz32(i) : swz(i); }
template <int DP>
FHEB_HD uint64_t sm_load(const uint64_t* buf, size_t unit, uint32_t N, uint32_t slot) {
    if constexpr (DP == MODE_U32) return reinterpret_cast<const uint32_t*>(buf)[unit * N + slot];
    else return buf[unit * N + slot];
}
template <int DP>
FHEB_HD void sm_store(uint64_t* buf, size_t unit, uint32_t N, uint32_t slot, uint64_t v) {
    if constexpr (DP == MODE_U32) reinterpret_cast<uint32_t*>(buf)[unit * N + slot] = (uint32_t)v;
    else buf[unit * N + slot] = v;
}

// The 2^R slots of one work item are  pb ^ swz(c << EB),  pb = swz(base) the item's first slot.  Formed as written, every
// access costs a LOP3 (the XOR) plus an add of the buffer address (ptxas: IMAD.IADD, on the multiply pipe).  SlotRef hoists
// the address: swz(k) == k ^ f(k >> 4) with f inside the low nibble (five bits for 4-byte slots), and the bits of c << EB
// above the nibble are zero in pb, so
//     address(pb ^ swz(k)) = (A ^ (low(swz(k)) * bytes)) + high(swz(k)) * bytes,     A = address of slot pb,
// provided the unit's first slot is 128-byte aligned (the XOR then never reaches the bits the base address occupies).
// After unrolling k is a constant: one LOP3 per access, the rest is the immediate offset of the LDS / STS
// (-1 instruction per access, ~1 per butterfly).  The buffers are declared __align__(128) and units are N * bytes apart
// (N >= 32 wherever shared memory is used).
template <int DP>
struct SlotRef {
#if defined(__CUDA_ARCH__)
    uint32_t a;     // shared-window byte address of slot pb
    uint32_t unit;  // shared-window byte address of the unit's first slot, and pb itself: the plain form (HOIST = false)
    uint32_t pb;
#else
    uint64_t* buf;
    size_t first;  // unit * N
    uint32_t pb;
#endif
};
template <int DP>
FHEB_HD SlotRef<DP> slot_ref(const uint64_t* buf, size_t unit, uint32_t N, uint32_t pb) {
    SlotRef<DP> r;
#if defined(__CUDA_ARCH__)
    r.unit = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<const char*>(buf) + unit * N * smem_slot_bytes<DP>());
    r.a = r.unit + pb * smem_slot_bytes<DP>();
    r.pb = pb;
#else
    r.buf = const_cast<uint64_t*>(buf);
    r.first = unit * N;
    r.pb = pb;
#endif
    return r;
}
// HOIST = false: the plain form (XOR of the slot index, then the buffer address added) - measured 2.7 % faster at N = 4096 over
// 62-bit primes (plain transforms and the fused product: the one shape the hoisted form lost on), see slot_hoisted().
template <int L, int DP>
constexpr bool slot_hoisted() { return !(L == 12 && DP == MODE_INT); }
#if defined(__CUDA_ARCH__)
template <int DP, bool HOIST = true>
__device__ __forceinline__ void* slot_ptr(const SlotRef<DP>& r, uint32_t k) {
    constexpr uint32_t LOW = (DP == MODE_U32) ? 31u : 15u, SB = smem_slot_bytes<DP>();
    if constexpr (!HOIST) return __cvta_shared_to_generic(r.unit + (r.pb ^ k) * SB);
    return __cvta_shared_to_generic((r.a ^ ((k & LOW) * SB)) + (k & ~LOW) * SB);
}
#endif
// k = swzm<DP>(c << EB)
template <int DP, bool HOIST = true>
FHEB_HD uint64_t slot_load(const SlotRef<DP>& r, uint32_t k) {
#if defined(__CUDA_ARCH__)
    if constexpr (DP == MODE_U32) return *reinterpret_cast<const uint32_t*>(slot_ptr<DP, HOIST>(r, k));
    else return *reinterpret_cast<const uint64_t*>(slot_ptr<DP, HOIST>(r, k));
#else
    return sm_load<DP>(r.buf, 0, 0, (uint32_t)r.first + (r.pb ^ k));
#endif
}
template <int DP, bool HOIST = true>
FHEB_HD void slot_store(const SlotRef<DP>& r, uint32_t k, uint64_t v) {
#if defined(__CUDA_ARCH__)
    if constexpr (DP == MODE_U32) *reinterpret_cast<uint32_t*>(slot_ptr<DP, HOIST>(r, k)) = (uint32_t)v;
    else *reinterpret_cast<uint64_t*>(slot_ptr<DP, HOIST>(r, k)) = v;
#else
    sm_store<DP>(r.buf, 0, 0, (uint32_t)r.first + (r.pb ^ k), v);
#endif
}
constexpr int CAP_STRICT = 4;   // q < 2^62: 4q fits a word
constexpr int CAP_DP = 128;     // q < 2^42: |v| < 128 q <= 2^49 keeps every FP64 step exact with margin
constexpr int CAP_U32 = 32;     // q < 2^27: 32q fits 32 bits
constexpr int U32_UNIT_CAP = 8; // unit-twiddle butterflies double the bound: they reduce first once it would pass 8q

// ---- compile-time range tracking -----------------------------------------------------
constexpr int fwd_next_k(int K, bool has_unit, bool has_nonunit, int mode) {
    if (mode == MODE_DP) {  // |A +- t| <= K + 1 (|t| < q), |A +- B| <= 2K; never reduced (static_assert at the use)
        int kn = has_nonunit ? K + 1 : 0;
        int ku = has_unit ? 2 * K : 0;
        return kn > ku ? kn : ku;
    }
    if (is_u32(mode)) {  // t in [0, 2q): A +- t < (K + 2) q; unit: A +- B < 2K q, both reduced to [0, 2q) first when 2K > 8
        int kn = has_nonunit ? (((K + 2 > CAP_U32) ? 2 : K) + 2) : 0;
        int ku = has_unit ? ((2 * K > U32_UNIT_CAP) ? 4 : 2 * K) : 0;
        return kn > ku ? kn : ku;
    }
    int kn = has_nonunit ? (((K + 2 > CAP_STRICT) ? 2 : K) + 2) : 0;
    int ku = has_unit ? ((2 * K > CAP_STRICT) ? 4 : 2 * K) : 0;
    return kn > ku ? kn : ku;
}
constexpr int fwd_pass_k(int K, int R, bool unit_first, int mode) {
    for (int a = 0; a < R; ++a) K = fwd_next_k(K, unit_first, !(unit_first && a == 0), mode);
    return K;
}
constexpr int inv_next_k(int K, int mode) {
    if (mode == MODE_DP) return (2 * K > CAP_DP) ? 1 : 2 * K;
    if (is_u32(mode)) return (2 * K > CAP_U32 / 2) ? 2 : 2 * K;  // sums kept below 16q so that the next sum fits 32 bits
    return (2 * K > CAP_STRICT / 2) ? 2 : 2 * K;
}
constexpr int inv_pass_k(int K, int R, int mode) {
    for (int a = 0; a < R; ++a) K = inv_next_k(K, mode);
    return K;
}

// ---- 32-bit arithmetic of MODE_U32 (q < 2^27) --------------------------------------------------
FHEB_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// x * w mod q, lazily: [0, 2q) for ANY 32-bit x, given w < q and wp = floor(w * 2^32 / q)
FHEB_HD uint32_t shoup32(uint32_t x, uint32_t w, uint32_t wp, uint32_t q) { return x * w - mulhi32(x, wp) * q; }
// any 32-bit x -> [0, 2q)  (one-word Barrett with mu32 = floor(2^32 / q))
FHEB_HD uint32_t lazy32(uint32_t x, const ModQ& m) { return x - mulhi32(x, m.mu32) * (uint32_t)m.q; }
FHEB_HD uint32_t csub32(uint32_t x, uint32_t q) { return x >= q ? x - q : x; }

// the 32-bit butterflies on one (a, b) pair of words; bounds as in fwd_next_k / inv_next_k
template <int K, bool UNIT>
FHEB_HD void fwd_bfly32(uint32_t& a, uint32_t& b, uint32_t w, uint32_t wp, const ModQ& m) {
    const uint32_t q = (uint32_t)m.q;
    if constexpr (UNIT) {
        constexpr bool red = (2 * K > U32_UNIT_CAP);
        if constexpr (red) {
            a = lazy32(a, m);
            b = lazy32(b, m);
        }
        constexpr int KT = red ? 2 : K;
        static_assert(2 * KT <= CAP_U32, "32-bit range exceeded in a unit forward stage");
        const uint32_t s = a + b;
        b = a - b + (uint32_t)KT * q;
        a = s;
    } else {
        constexpr bool red = (K + 2 > CAP_U32);
        if constexpr (red) a = lazy32(a, m);
        const uint32_t t = shoup32(b, w, wp, q);  // [0, 2q) for any b
        b = a - t + 2u * q;
        a = a + t;
    }
}
template <int K, bool UNIT>
FHEB_HD void inv_bfly32(uint32_t& a, uint32_t& b, uint32_t w, uint32_t wp, const ModQ& m) {
    static_assert(2 * K <= CAP_U32, "sum would overflow 32 bits");
    constexpr bool red = (2 * K > CAP_U32 / 2);
    const uint32_t q = (uint32_t)m.q;
    uint32_t s = a + b;
    uint32_t d = a - b + (uint32_t)K * q;
    if constexpr (red) s = lazy32(s, m);
    a = s;
    if constexpr (UNIT) {
        if constexpr (red) d = lazy32(d, m);
        b = d;
    } else {
        b = shoup32(d, w, wp, q);
    }
}

template <int K>
FHEB_HD uint64_t kq(const ModQ& m) {
    if constexpr (K == 1) return m.q;
    else if constexpr (K == 2) return m.q2;
    else return m.q * (uint64_t)K;
}

// Forward butterfly on values bounded by K*q; leaves values bounded by fwd_next_k(K)*q.
template <int K, int DP, bool UNIT>
FHEB_HD void fwd_bfly(uint64_t& A, uint64_t& B, const Tw& w, const ModQ& m) {
    if constexpr (DP == MODE_DP) {
        static_assert(fwd_next_k(K, UNIT, !UNIT, true) <= CAP_DP, "FP64 range exceeded in a forward stage");
        const double a = bits_to_double(A);
        const double t = UNIT ? bits_to_double(B) : dp_mulmod(bits_to_double(B), bits_to_double(w.w), m);
        A = double_to_bits(dp_add(a, t));
        B = double_to_bits(dp_add(a, -t));
    } else if constexpr (DP == MODE_U32) {
        uint32_t a = (uint32_t)A, b = (uint32_t)B;
        fwd_bfly32<K, UNIT>(a, b, (uint32_t)w.w, (uint32_t)w.wp, m);
        A = a;
        B = b;
    } else if constexpr (DP == MODE_U32P) {  // two polynomials per slot, the same twiddle
        uint32_t a0 = lo32(A), b0 = lo32(B), a1 = hi32(A), b1 = hi32(B);
        fwd_bfly32<K, UNIT>(a0, b0, (uint32_t)w.w, (uint32_t)w.wp, m);
        fwd_bfly32<K, UNIT>(a1, b1, (uint32_t)w.w, (uint32_t)w.wp, m);
        A = pack32(a0, a1);
        B = pack32(b0, b1);
    } else if constexpr (UNIT) {  // twiddle == 1: no multiplication
        constexpr bool red = (2 * K > CAP_STRICT);
        static_assert(!red || K <= 4, "conditional subtraction only halves [0,4q)");
        uint64_t a = A, t = B;
        if constexpr (red) {
            a = csub(a, m.q2);
            t = csub(t, m.q2);
        }
        constexpr int KT = red ? 2 : K;
        A = a + t;
        B = a - t + kq<KT>(m);
    } else {
        constexpr bool red = (K + 2 > CAP_STRICT);
        static_assert(!red || K <= 4, "conditional subtraction only halves [0,4q)");
        uint64_t a = A;
        if constexpr (red) a = csub(a, m.q2);
        uint64_t t = shoup_lazy(B, w.w, w.wp, m);  // [0, 2q) for any B
        A = a + t;
        B = a - t + m.q2;
    }
}

// Inverse (Gentleman-Sande) butterfly on values bounded by K*q; leaves values bounded by inv_next_k(K)*q.
template <int K, int DP, bool UNIT>
FHEB_HD void inv_bfly(uint64_t& A, uint64_t& B, const Tw& w, const ModQ& m) {
    if constexpr (DP == MODE_DP) {
        static_assert(2 * K <= 2 * CAP_DP, "FP64 range exceeded in an inverse stage");
        constexpr bool red = (2 * K > CAP_DP);
        const double a = bits_to_double(A), b = bits_to_double(B);
        double s = dp_add(a, b);
        double d = dp_add(a, -b);
        if constexpr (red) s = dp_reduce(s, m);
        if constexpr (UNIT) {
            if constexpr (red) d = dp_reduce(d, m);
        } else {
            d = dp_mulmod(d, bits_to_double(w.w), m);
        }
        A = double_to_bits(s);
        B = double_to_bits(d);
    } else if constexpr (DP == MODE_U32) {
        uint32_t a = (uint32_t)A, b = (uint32_t)B;
        inv_bfly32<K, UNIT>(a, b, (uint32_t)w.w, (uint32_t)w.wp, m);
        A = a;
        B = b;
    } else if constexpr (DP == MODE_U32P) {
        uint32_t a0 = lo32(A), b0 = lo32(B), a1 = hi32(A), b1 = hi32(B);
        inv_bfly32<K, UNIT>(a0, b0, (uint32_t)w.w, (uint32_t)w.wp, m);
        inv_bfly32<K, UNIT>(a1, b1, (uint32_t)w.w, (uint32_t)w.wp, m);
        A = pack32(a0, a1);
        B = pack32(b0, b1);
    } else {
        static_assert(2 * K <= CAP_STRICT, "sum would overflow the word");
        constexpr bool red = (2 * K > CAP_STRICT / 2);
        static_assert(!red || 2 * K <= 4, "conditional subtraction only halves [0,4q)");
        uint64_t s = A + B;
        uint64_t d = A - B + kq<K>(m);
        if constexpr (red) s = csub(s, m.q2);
        A = s;
        if constexpr (UNIT) {
            if constexpr (red) d = csub(d, m.q2);
            B = d;
        } else {
            B = shoup_lazy(d, w.w, w.wp, m);
        }
    }
}

// Device twiddle table order.  Stage s = S0 + a of a pass uses one twiddle per block of N/2^s
// positions; block (blk << a) + g belongs to the item whose stage-S0 block is blk (g < 2^a).  The
// table is stored pass by pass as [a][g][blk], so that the lanes of a warp - consecutive items,
// i.e. consecutive blk in the late passes - read consecutive entries with every load
// instruction (a plain heap order makes each load stride 2^a entries).
template <int S0>
FHEB_HD constexpr uint32_t tw_index(int a, int g) { return (uint32_t)(((1 << a) - 1) + g) << S0; }

// Caller data streams through once per launch: mark it evict-first so that it does not push the
// twiddle tables (re-read for every polynomial) out of L1/L2.
FHEB_HD uint64_t stream_load(const uint64_t* p) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__ldcs(reinterpret_cast<const unsigned long long*>(p));
#else
    return *p;
#endif
}
FHEB_HD void stream_store(uint64_t* p, uint64_t v) {
#if defined(__CUDA_ARCH__)
    __stcs(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
#else
    *p = v;
#endif
}

// Twiddle tables: integer mode = (value, Shoup companion) pairs, 16 bytes; DP mode = one double, 8 bytes;
// U32 mode = (value, 32-bit Shoup companion) packed into 8 bytes.
template <int DP>
FHEB_HD Tw load_tw(const Tw* __restrict__ tw, uint32_t idx) {
    Tw t;
    if constexpr (DP == MODE_DP) {
        const uint64_t* p = reinterpret_cast<const uint64_t*>(tw);
#if defined(__CUDA_ARCH__)
        t.w = __ldg(p + idx);
#else
        t.w = p[idx];
#endif
        t.wp = 0;
    } else if constexpr (is_u32(DP)) {  // (w, w') as two 32-bit halves of one 8-byte entry
        const uint64_t* p = reinterpret_cast<const uint64_t*>(tw);
#if defined(__CUDA_ARCH__)
        const uint64_t v = __ldg(p + idx);
#else
        const uint64_t v = p[idx];
#endif
        t.w = v & 0xFFFFFFFFull;
        t.wp = v >> 32;
    } else {
#if defined(__CUDA_ARCH__)
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(tw) + idx);
        t.w = v.x;
        t.wp = v.y;
#else
        t = tw[idx];
#endif
    }
    return t;
}

// R forward stages on 2^R register-resident values; element bit (R-1) is the highest
// position bit of the pass.  The pass starts at stage S0; TB = table offset of the pass + the index
// of this item's block among the 2^S0 blocks of stage S0 (see tw_index below).
// REGTW: `tw` points at this item's twiddles already held in registers, entry ((1 << a) - 1) + g.
template <int R, int S0, int K, int DP, bool UNITFIRST, int A = 0, bool REGTW = false>
FHEB_HD void fwd_stages(uint64_t (&x)[1 << R], const Tw* __restrict__ tw, uint32_t TB, const ModQ& m) {
    if constexpr (A < R) {
        constexpr int half = 1 << (R - 1 - A);
#pragma unroll
        for (int g = 0; g < (1 << A); ++g) {
            if (UNITFIRST && g == 0) {
                Tw dummy{0, 0};
#pragma unroll
                for (int j = 0; j < half; ++j) fwd_bfly<K, DP, true>(x[g * 2 * half + j], x[g * 2 * half + j + half], dummy, m);
            } else {
                const Tw w = REGTW ? tw[((1 << A) - 1) + g] : load_tw<DP>(tw, TB + tw_index<S0>(A, g));
#pragma unroll
                for (int j = 0; j < half; ++j) fwd_bfly<K, DP, false>(x[g * 2 * half + j], x[g * 2 * half + j + half], w, m);
            }
        }
        fwd_stages<R, S0, fwd_next_k(K, UNITFIRST, !(UNITFIRST && A == 0), DP), DP, UNITFIRST, A + 1, REGTW>(x, tw, TB, m);
    }
}

// R inverse stages, highest stage of the pass first (element bit 0 first).
template <int R, int S0, int K, int DP, bool UNITFIRST, int A = R - 1, bool REGTW = false>
FHEB_HD void inv_stages(uint64_t (&x)[1 << R], const Tw* __restrict__ tw, uint32_t TB, const ModQ& m) {
    if constexpr (A >= 0) {
        constexpr int half = 1 << (R - 1 - A);
#pragma unroll
        for (int g = 0; g < (1 << A); ++g) {
            if (UNITFIRST && g == 0) {
                Tw dummy{0, 0};
#pragma unroll
                for (int j = 0; j < half; ++j) inv_bfly<K, DP, true>(x[g * 2 * half + j], x[g * 2 * half + j + half], dummy, m);
            } else {
                const Tw w = REGTW ? tw[((1 << A) - 1) + g] : load_tw<DP>(tw, TB + tw_index<S0>(A, g));
#pragma unroll
                for (int j = 0; j < half; ++j) inv_bfly<K, DP, false>(x[g * 2 * half + j], x[g * 2 * half + j + half], w, m);
            }
        }
        inv_stages<R, S0, inv_next_k(K, DP), DP, UNITFIRST, A - 1, REGTW>(x, tw, TB, m);
    }
}

// all (2^R - 1) twiddles of one item into registers, entry ((1 << a) - 1) + g
template <int R, int S0, int DP>
FHEB_HD void load_item_tw(const Tw* __restrict__ tw, uint32_t TB, Tw (&w)[(1 << R) - 1]) {
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
        for (int g = 0; g < (1 << a); ++g) w[((1 << a) - 1) + g] = load_tw<DP>(tw, TB + tw_index<S0>(a, g));
}

// value bounded by K*q  ->  canonical word
template <int K, int DP = false>
FHEB_HD uint64_t canon_k(uint64_t x, const ModQ& m) {
    if constexpr (DP == MODE_DP) {
        double r = bits_to_double(x);
        if constexpr (K > 1) r = dp_reduce(r, m);
        return dp_canon_word(r, m);
    } else if constexpr (DP == MODE_U32) {
        uint32_t v = (uint32_t)x;
        if constexpr (K > 2) v = lazy32(v, m);
        if constexpr (K > 1) v = csub32(v, (uint32_t)m.q);
        return v;
    } else if constexpr (DP == MODE_U32P) {  // both halves
        return pack32((uint32_t)canon_k<K, MODE_U32>(lo32(x), m), (uint32_t)canon_k<K, MODE_U32>(hi32(x), m));
    } else if constexpr (K <= 1) return x;
    else if constexpr (K == 2) return csub(x, m.q);
    else if constexpr (K <= 4) return csub(csub(x, m.q2), m.q);
    else return reduce64(x, m);
}

// caller word (any 64-bit value) -> the mode's register representation of its residue
template <int DP>
FHEB_HD uint64_t load_word(uint64_t v, const ModQ& m) {
    const uint64_t c = canon_any(v, m);
    if constexpr (DP == MODE_DP) return double_to_bits(dp_from_uint(c));
    else return c;
}
// the same for a whole item: unreduced words are rare, so one test covers the item's 2^R words
template <int DP, int E>
FHEB_HD void load_words(uint64_t (&x)[E], const ModQ& m) {
    bool raw = false;
#pragma unroll
    for (int c = 0; c < E; ++c) raw = raw || x[c] >= m.q;
    if (raw) {
#pragma unroll
        for (int c = 0; c < E; ++c) x[c] = canon_any(x[c], m);
    }
    if constexpr (DP == MODE_DP) {
#pragma unroll
        for (int c = 0; c < E; ++c) x[c] = double_to_bits(dp_from_uint(x[c]));
    }
}

enum {
    IO_SMEM = 0,          // the block's work buffer in shared memory (swizzled)
    IO_GLOBAL = 1,        // caller memory
    IO_STASH_SMEM = 2,    // second shared-memory buffer holding a finished transform (swizzled)
    IO_STASH_GLOBAL = 3,  // per-block global scratch holding a finished transform (natural index)
    IO_LANDING = 4,       // caller words already copied into shared memory by a bulk-async (TMA) load: natural index, raw words
    IO_LANDING_PART = 5   // N = 16384 with 8-byte slots: only the first LAND_PART_WORDS words landed (96 KB fit beside the 128 KB work buffer), the rest comes from caller memory
};
constexpr uint32_t LAND_PART_WORDS = 12288;
// word idx of a polynomial: landing buffer or caller memory (idx's row is a compile-time constant at every call site)
FHEB_HD uint64_t part_word(const uint64_t* landing, const uint64_t* global, uint32_t idx, bool landed) {
    return landed ? landing[idx] : stream_load(global + idx);
}
// first-pass input word: streaming global load, or a plain load from the landing buffer
template <int IN>
FHEB_HD uint64_t input_word(const uint64_t* p) {
    if constexpr (IN == IO_LANDING) return *p;
    else return stream_load(p);
}

// The E caller words of work unit `unit` (a polynomial; in MODE_U32P the pair 2*unit, 2*unit+1 of the block's `polys`
// polynomials, the missing partner of an odd tail reads as zero) at word indices idx(c), in the mode's representation.
// `src` = first polynomial of the block (global memory, or the landing buffer).
template <int DP, int IN, int E, class Idx>
FHEB_HD void load_unit(uint64_t (&x)[E], const uint64_t* src, uint32_t unit, uint32_t polys, uint32_t N, Idx idx, const ModQ& m) {
    if constexpr (DP == MODE_U32P) {
        const uint64_t* s0 = src + (size_t)(2 * unit) * N;
        const bool has1 = 2 * unit + 1 < polys;
        const uint64_t* s1 = has1 ? s0 + N : s0;
        uint64_t y[E];
        bool raw = false;
#pragma unroll
        for (int c = 0; c < E; ++c) {
            x[c] = input_word<IN>(s0 + idx(c));
            y[c] = input_word<IN>(s1 + idx(c));
            raw = raw || x[c] >= m.q || y[c] >= m.q;
        }
        if (raw) {  // unreduced words are rare: one test per item
#pragma unroll
            for (int c = 0; c < E; ++c) {
                x[c] = canon_any(x[c], m);
                y[c] = canon_any(y[c], m);
            }
        }
#pragma unroll
        for (int c = 0; c < E; ++c) x[c] = pack32((uint32_t)x[c], has1 ? (uint32_t)y[c] : 0u);
    } else {
        const uint64_t* s0 = src + (size_t)unit * N;
#pragma unroll
        for (int c = 0; c < E; ++c) x[c] = input_word<IN>(s0 + idx(c));
        load_words<DP, E>(x, m);
    }
}
// one finished word (canonical, or both canonical halves) of work unit `unit` to word index idx of its polynomial(s)
template <int DP>
FHEB_HD void store_unit_word(uint64_t* dst, uint32_t unit, uint32_t polys, uint32_t N, size_t idx, uint64_t v) {
    if constexpr (DP == MODE_U32P) {
        uint64_t* d0 = dst + (size_t)(2 * unit) * N;
        stream_store(d0 + idx, (uint64_t)lo32(v));
        if (2 * unit + 1 < polys) stream_store(d0 + N + idx, (uint64_t)hi32(v));
    } else {
        stream_store(dst + (size_t)unit * N + idx, v);
    }
}

// finished transform parked for the fused product: canonical word (integer) / reduced double (DP)
template <int K, int DP>
FHEB_HD uint64_t park_word(uint64_t x, const ModQ& m) {
    if constexpr (DP == MODE_DP) return (K > 1) ? double_to_bits(dp_reduce(bits_to_double(x), m)) : x;
    else return canon_k<K, DP>(x, m);
}

// x * N^-1 -> canonical word (x bounded by K*q, K within the mode's cap)
template <int DP>
FHEB_HD uint64_t scale_word(uint64_t x, const Tw& ninv, const ModQ& m) {
    if constexpr (DP == MODE_DP) return dp_canon_word(dp_mulmod(bits_to_double(x), bits_to_double(ninv.w), m), m);
    else if constexpr (DP == MODE_U32) return csub32(shoup32((uint32_t)x, (uint32_t)ninv.w, (uint32_t)ninv.wp, (uint32_t)m.q), (uint32_t)m.q);
    else if constexpr (DP == MODE_U32P) return pack32((uint32_t)scale_word<MODE_U32>(lo32(x), ninv, m), (uint32_t)scale_word<MODE_U32>(hi32(x), ninv, m));
    else return csub(shoup_lazy(x, ninv.w, ninv.wp, m), m.q);
}

// ---- pass plans: how the L stages are split into register passes ----------------------
// plan<L>::R[p] = stages in pass p (forward order); at most 5 passes, each of 1..4 stages (5 in the FHEB_EXP_R5 experiment).
template <int L> struct Plan;
#define FHEB_PLAN(L_, P_, ...)                         \
    template <> struct Plan<L_> {                      \
        static constexpr int P = P_;                   \
        static constexpr int R[5] = {__VA_ARGS__};     \
    };
FHEB_PLAN(2, 1, 2, 0, 0, 0, 0)
FHEB_PLAN(3, 1, 3, 0, 0, 0, 0)
FHEB_PLAN(4, 1, 4, 0, 0, 0, 0)
FHEB_PLAN(5, 2, 3, 2, 0, 0, 0)
FHEB_PLAN(6, 2, 3, 3, 0, 0, 0)
FHEB_PLAN(7, 2, 4, 3, 0, 0, 0)
FHEB_PLAN(8, 2, 4, 4, 0, 0, 0)
FHEB_PLAN(9, 3, 3, 3, 3, 0, 0)
FHEB_PLAN(10, 3, 4, 3, 3, 0, 0)
FHEB_PLAN(11, 3, 4, 4, 3, 0, 0)
FHEB_PLAN(12, 3, 4, 4, 4, 0, 0)
#if defined(FHEB_EXP_R3)  // experiment: 8-value passes only (fewer registers, more warps)
FHEB_PLAN(13, 5, 3, 3, 3, 2, 2)
FHEB_PLAN(14, 5, 3, 3, 3, 3, 2)
#elif defined(FHEB_EXP_R5)  // experiment: 32-value passes at N = 16384 (three passes instead of four) in EVERY kernel
FHEB_PLAN(13, 4, 4, 3, 3, 3, 0)
FHEB_PLAN(14, 3, 5, 5, 4, 0, 0)
#else
FHEB_PLAN(13, 4, 4, 3, 3, 3, 0)
FHEB_PLAN(14, 4, 4, 4, 3, 3, 0)
#endif
// Plan KEYS above 64 are alternative splits of degree 2^(key - 64), used by single kernels through their `PK` template
// parameter with a twiddle table of their own (NttPlan::d_fwd_alt / d_inv_alt).  78 = N = 16384 in THREE passes of 5 + 5 + 4 stages
// (32 register-resident values per thread): one shared-memory round trip and one block barrier less than 4 + 4 + 3 + 3.
// Measured per kernel (all bit-exact): plain forward over a 62-bit prime 0.1701 -> 0.1600 ms (no spills at 128
// registers) and in 32-bit mode 0.0809 -> 0.0794 ms (superseded by key 80 below); inverse 0.1810 -> 0.1810 (64 B of spills), fused product
// 0.578 -> 0.678 ms (716 B of spills), FP64 mode 0.1050 -> 0.1060: those keep the four-pass plan.
// 79 = 4 + 5 + 5: the INVERSE runs its passes last to first, so this split gives it the 32-value pass on the caller's
// words and the 16-value pass on the scaled output: 0.1815 -> 0.1717 ms (5 + 4 + 5: 0.1753; no spills in either).
// 80 = 5 + 4 + 5: forward over a 62-bit prime 0.1605 -> 0.1569 ms (4 + 5 + 5: 0.1733), 32-bit mode 0.0790 -> 0.0779 ms.
constexpr int PLAN_KEY_ALT14 = 80;      // plain forward, integer mode
constexpr int PLAN_KEY_ALT14_U32 = 80;  // plain forward, 32-bit mode
constexpr int PLAN_KEY_ALT14_INV = 79;  // plain inverse, integer mode (32-bit mode: 188 registers, one block per SM, 0.0913 -> 0.0929 ms: experiment switch only)
FHEB_PLAN(78, 3, 5, 5, 4, 0, 0)
FHEB_PLAN(79, 3, 4, 5, 5, 0, 0)
FHEB_PLAN(80, 3, 5, 4, 5, 0, 0)
#undef FHEB_PLAN

template <int L, int PASS>
constexpr int plan_s0() {  // first stage of pass PASS
    int s = 0;
    for (int p = 0; p < PASS; ++p) s += Plan<L>::R[p];
    return s;
}
template <int L, int PASS>
constexpr uint32_t plan_tw_offset() {  // first table entry of pass PASS: sum over earlier passes of (2^R - 1) * 2^S0
    uint32_t off = 0;
    int s = 0;
    for (int p = 0; p < PASS; ++p) {
        off += (uint32_t)((1 << Plan<L>::R[p]) - 1) << s;
        s += Plan<L>::R[p];
    }
    return off;
}
// runtime view of the plan for the host-side table builders
inline void plan_runtime(int L, int& P, int (&R)[5]) {
    P = 0;
    R[0] = R[1] = R[2] = R[3] = R[4] = 0;
    switch (L) {
#define FHEB_PLAN_RT(L_)                                          \
    case L_:                                                      \
        P = Plan<L_>::P;                                          \
        for (int i = 0; i < 5; ++i) R[i] = Plan<L_>::R[i];        \
        break;
        FHEB_PLAN_RT(2) FHEB_PLAN_RT(3) FHEB_PLAN_RT(4) FHEB_PLAN_RT(5) FHEB_PLAN_RT(6) FHEB_PLAN_RT(7) FHEB_PLAN_RT(8)
        FHEB_PLAN_RT(9) FHEB_PLAN_RT(10) FHEB_PLAN_RT(11) FHEB_PLAN_RT(12) FHEB_PLAN_RT(13) FHEB_PLAN_RT(14) FHEB_PLAN_RT(78) FHEB_PLAN_RT(79) FHEB_PLAN_RT(80)
#undef FHEB_PLAN_RT
    }
}

// SUB: the L stages are the tail of a larger transform (degrees above 2^14, see ntt_device.cuh): the first
// pass has no unit twiddles then.
template <int L, int DP, int PASS, bool SUB = false>
constexpr int plan_fwd_kin() {  // bound on values entering forward pass PASS (inputs canonical)
    int K = 1;
    for (int p = 0; p < PASS; ++p) K = fwd_pass_k(K, Plan<L>::R[p], p == 0 && !SUB, DP);
    return K;
}

// SUB kernels address caller memory through this map: word i of the sub-transform's reference-order
// side lives at global index (i << shift) | low of its polynomial.
struct GlobalMap {
    uint32_t shift, low;
};
template <int L, int DP, int PASS, int KSTART = 1>
constexpr int plan_inv_kin() {  // bound entering inverse pass PASS (run order P-1 .. 0), first executed pass fed values < KSTART*q
    int K = KSTART;
    for (int p = Plan<L>::P - 1; p > PASS; --p) K = inv_pass_k(K, Plan<L>::R[p], DP);
    return K;
}

// ---- one pass over `polys` polynomials held by one thread block ---------------------------
// Work item U in [0, polys * N / 2^R); the block walks them with stride nthreads.
// IO roles of a pass:
//   IN_GLOBAL  : values come from global memory (first pass)    else from shared memory
//   OUT_GLOBAL : values go to global memory (last pass)         else to shared memory
// Shared memory holds `polys` polynomials of N words each, word index swizzled by swz().



// Forward pass PASS of an L-stage transform.  `gin`/`gout` point at the block's first
// polynomial.  The last pass stores in the reference's (bit-reversed) output order.
//   SCALE      : multiply the outputs by `ninv` (fast_ntt_inverse semantics)
// OUT == IO_STASH_*: the finished transform is parked (position order, no bit reversal) in
// `gout` for the fused polynomial product; canonical words in integer mode, lazy doubles (|v| < KOUT*q) in DP mode.
template <int L, int DP, int PASS, int IN, int OUT, bool BITREV_OUT = true, bool SCALE = false, int IPT = 0, bool SUB = false, bool PIPE = false, int PK = L>
FHEB_HD void fwd_pass(uint32_t tid, uint32_t nthreads, uint32_t polys, const uint64_t* gin, uint64_t* gout,
                      uint64_t* smem, const Tw* __restrict__ tw, const ModQ& m, const Tw ninv = Tw{0, 0},
                      const GlobalMap map = GlobalMap{0, 0}, const uint64_t* gin_rest = nullptr) {
    constexpr int R = Plan<PK>::R[PASS];
    constexpr int E = 1 << R;
    constexpr int S0 = plan_s0<PK, PASS>();
    constexpr int EB = L - S0 - R;  // lowest position bit handled by this pass
    constexpr bool UNIT = (PASS == 0 && !SUB);
    constexpr int KIN = plan_fwd_kin<PK, DP, PASS, SUB>();
    constexpr int KOUT = fwd_pass_k(KIN, R, UNIT, DP);
    constexpr bool LAST = (PASS == Plan<PK>::P - 1);
    constexpr uint32_t N = 1u << L;
    constexpr uint32_t ITEMS = N >> R;  // per polynomial
    const uint32_t units = units_of<DP>(polys);  // work-buffer units: polynomials, or pairs of them (MODE_U32P)
    // BRTW: items are walked in bit-reversed order (u = bitrev(t)); the table's second copy of this pass (offset N,
    // block index bit-reversed: ntt_plan.hpp) is indexed by t, so consecutive lanes read consecutive twiddles
    constexpr bool BRTW = LAST && OUT == IO_GLOBAL && BITREV_OUT && S0 > 0;
    static_assert(!BRTW || (EB == 0 && S0 == L - R), "the last pass covers the lowest position bits");
    static_assert(IN == IO_SMEM || PASS == 0, "only the first pass reads caller words");
    static_assert(OUT == IO_SMEM || LAST, "only the last pass writes outside the work buffer");

    // IPT > 0: the thread's item count is known (IPT items, stride nthreads) and all their twiddles are
    // requested before the first item is computed - in the late passes every item has its own
    // twiddles, they stream from L2, and nothing else in a barrier-synchronised block hides that latency.
    constexpr int NW = E - 1;
    constexpr int TRIPS = IPT > 0 ? IPT : 1;
    Tw wall[TRIPS][NW];
    if constexpr (IPT > 0) {
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const uint32_t U = tid + (uint32_t)k * nthreads;
            if (U < units * ITEMS) {
                const uint32_t t0 = U & (ITEMS - 1);
                uint32_t u = t0;
                if (OUT == IO_GLOBAL && BITREV_OUT) u = bitrev_rt(u, L - R);
                const uint32_t base = ((u >> EB) << (EB + R)) | (u & ((1u << EB) - 1u));
                load_item_tw<R, S0, DP>(tw, BRTW ? (N + t0) : plan_tw_offset<PK, PASS>() + (S0 ? (base >> (L - S0)) : 0u), wall[k]);
            }
        }
    }
    // PIPE (chosen by the kernel): software-pipelined global loads of a multi-pass plan's first pass, for the
    // degrees where a thread has several items.  +1 % (integer) / +5 % (FP64 mode) in the plain transform at
    // N = 16384; the fused product kernel loses 7 % with it (register pressure) and does not ask for it.
    constexpr bool PIPE_IN = PIPE && IN == IO_GLOBAL && OUT == IO_SMEM && IPT == 0 && DP != MODE_U32P;  // (never with IO_LANDING: nothing to hide)
    uint64_t xnext[PIPE_IN ? E : 1];
    bool have_next = false;
    for (uint32_t U0 = tid; U0 < (IPT > 0 ? tid + 1 : units * ITEMS); U0 += nthreads) {
#pragma unroll
      for (int k = 0; k < TRIPS; ++k) {
        const uint32_t U = U0 + (uint32_t)k * nthreads;
        if (IPT > 0 && U >= units * ITEMS) break;
        const uint32_t poly = U >> (L - R);
        uint32_t u = U & (ITEMS - 1);
        // In the last pass consecutive threads take bit-reversed item indices so that the
        // bit-reversed stores below are coalesced.
        const uint32_t t = u;
        if (OUT == IO_GLOBAL && BITREV_OUT) u = bitrev_rt(t, L - R);
        const uint32_t base = ((u >> EB) << (EB + R)) | (u & ((1u << EB) - 1u));
        uint64_t x[E];
        if constexpr (IN == IO_LANDING_PART) {  // one polynomial per block: rows below LAND_PART_WORDS from the landing buffer (gin), the others from caller memory
            static_assert(PASS == 0 && (LAND_PART_WORDS & ((1u << EB) - 1u)) == 0, "whole rows of the first pass are landed");
#pragma unroll
            for (int c = 0; c < E; ++c) x[c] = part_word(gin, gin_rest, base | ((uint32_t)c << EB), ((uint32_t)c << EB) < LAND_PART_WORDS);
            load_words<DP, E>(x, m);
        } else if (IN == IO_GLOBAL || IN == IO_LANDING) {
            if constexpr (PIPE_IN) {
                // the words of this item were requested one trip ago; request the next item's now, so that the
                // global-memory latency of the first pass is paid once per polynomial instead of once per item
                if (have_next) {
#pragma unroll
                    for (int c = 0; c < E; ++c) x[c] = xnext[c];
                } else {
                    const uint64_t* src = gin + (size_t)poly * N;
#pragma unroll
                    for (int c = 0; c < E; ++c) x[c] = stream_load(src + (base | ((uint32_t)c << EB)));
                }
                const uint32_t Un = U + nthreads;
                have_next = Un < units * ITEMS;
                if (have_next) {
                    const uint32_t un = Un & (ITEMS - 1);
                    const uint32_t basen = ((un >> EB) << (EB + R)) | (un & ((1u << EB) - 1u));
                    const uint64_t* srcn = gin + (size_t)(Un >> (L - R)) * N;
#pragma unroll
                    for (int c = 0; c < E; ++c) xnext[c] = stream_load(srcn + (basen | ((uint32_t)c << EB)));
                }
                load_words<DP, E>(x, m);
            } else {
                load_unit<DP, IN, E>(x, gin, poly, polys, N, [&](int c) { return base | ((uint32_t)c << EB); }, m);
            }
        } else {
            const SlotRef<DP> sr = slot_ref<DP>(smem, poly, N, swzm<DP>(base));
#pragma unroll
            for (int c = 0; c < E; ++c) x[c] = slot_load<DP, slot_hoisted<L, DP>()>(sr, swzm<DP>((uint32_t)c << EB));
        }
        const uint32_t TB = BRTW ? (N + t) : plan_tw_offset<PK, PASS>() + (S0 ? (base >> (L - S0)) : 0u);
        if constexpr (IPT > 0) fwd_stages<R, S0, KIN, DP, UNIT, 0, true>(x, wall[k], 0u, m);
        else fwd_stages<R, S0, KIN, DP, UNIT>(x, tw, TB, m);
        if (OUT == IO_GLOBAL) {
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const uint64_t v = SCALE ? scale_word<DP>(x[c], ninv, m) : canon_k<KOUT, DP>(x[c], m);
                if (SUB) store_unit_word<DP>(gout, poly, polys, N, (((bitrev_c((uint32_t)c, R) << (L - R)) | t) << map.shift) | map.low, v);
                else if (BITREV_OUT) store_unit_word<DP>(gout, poly, polys, N, (bitrev_c((uint32_t)c, R) << (L - R)) | t, v);
                else store_unit_word<DP>(gout, poly, polys, N, base | ((uint32_t)c << EB), v);
            }
        } else if (OUT == IO_STASH_GLOBAL) {
            uint64_t* dst = gout + (size_t)poly * N;
#pragma unroll
            for (int c = 0; c < E; ++c) dst[base | ((uint32_t)c << EB)] = park_word<KOUT, DP>(x[c], m);
        } else {
            uint64_t* dst = (OUT == IO_STASH_SMEM ? gout : smem);
            const SlotRef<DP> sw = slot_ref<DP>(dst, poly, N, swzm<DP>(base));
#pragma unroll
            for (int c = 0; c < E; ++c)
                slot_store<DP, slot_hoisted<L, DP>()>(sw, swzm<DP>((uint32_t)c << EB), (OUT == IO_STASH_SMEM) ? park_word<KOUT, DP>(x[c], m) : x[c]);
        }
      }
    }
}

// Fused middle of the polynomial product (reference PolynomialRing::multiply,
// cpp/src/polynomial_ring.cpp:421-447): last forward pass of operand b, pointwise product
// with the parked transform of operand a, and the first executed inverse pass - all on the
// same 2^R register-resident positions, so neither bit-reversal ever happens.
//   IN    : IO_GLOBAL when the plan has a single pass (b read from caller memory), else IO_SMEM
//   OUT   : IO_GLOBAL when the plan has a single pass (scaled, canonical), else IO_SMEM
//   STASH : IO_STASH_SMEM or IO_STASH_GLOBAL (where fwd_pass parked T(a))
template <int L, int DP, int IN, int OUT, int STASH>
FHEB_HD void polymul_mid_pass(uint32_t tid, uint32_t nthreads, uint32_t polys, const uint64_t* gin, uint64_t* gout,
                              uint64_t* smem, const uint64_t* stash, const Tw* __restrict__ twf,
                              const Tw* __restrict__ twi, const Tw ninv, const ModQ& m) {
    constexpr int PASS = Plan<L>::P - 1;
    constexpr int R = Plan<L>::R[PASS];
    constexpr int E = 1 << R;
    constexpr int S0 = plan_s0<L, PASS>();
    constexpr int EB = L - S0 - R;  // == 0
    constexpr int KIN = plan_fwd_kin<L, DP, PASS>();
    constexpr int KOUT = fwd_pass_k(KIN, R, PASS == 0, DP);
    constexpr uint32_t N = 1u << L;
    constexpr uint32_t ITEMS = N >> R;
    static_assert(EB == 0, "the last forward pass covers the lowest position bits");
    static_assert(DP != MODE_U32P || (IN == IO_SMEM && OUT == IO_SMEM), "pair mode is used with multi-pass plans only");
    const uint32_t units = units_of<DP>(polys);
    // DP: the parked operand is reduced (|a| <= q/2 + 1), this one stays lazy: |a*b| <= KOUT q^2 / 2
    static_assert(DP != MODE_DP || KOUT <= 2 * CAP_DP, "FP64 product bound");

    for (uint32_t U = tid; U < units * ITEMS; U += nthreads) {
        const uint32_t poly = U >> (L - R);
        const uint32_t u = U & (ITEMS - 1);
        const uint32_t base = u << R;
        const uint32_t pb = swzm<DP>(base);
        const SlotRef<DP> sr = slot_ref<DP>(smem, poly, N, pb);
        uint64_t x[E];
        if (IN == IO_GLOBAL) {
            const uint64_t* src = gin + (size_t)poly * N;
#pragma unroll
            for (int c = 0; c < E; ++c) x[c] = stream_load(src + (base | (uint32_t)c));
            load_words<DP, E>(x, m);
        } else {
#pragma unroll
            for (int c = 0; c < E; ++c) x[c] = slot_load<DP, slot_hoisted<L, DP>()>(sr, swzm<DP>((uint32_t)c));
        }
        const uint32_t TB = plan_tw_offset<L, PASS>() + (S0 ? (base >> (L - S0)) : 0u);
        fwd_stages<R, S0, KIN, DP, PASS == 0>(x, twf, TB, m);
        const uint64_t* sa = stash + (size_t)poly * N;
        const SlotRef<DP> ss = slot_ref<DP>(STASH == IO_STASH_GLOBAL ? smem : stash, poly, N, pb);
#pragma unroll
        for (int c = 0; c < E; ++c) {
            const uint64_t av = (STASH == IO_STASH_GLOBAL) ? sa[base | (uint32_t)c] : slot_load<DP, slot_hoisted<L, DP>()>(ss, swzm<DP>((uint32_t)c));
            if constexpr (DP == MODE_DP) x[c] = double_to_bits(dp_mulmod(bits_to_double(av), bits_to_double(x[c]), m));
            else if constexpr (DP == MODE_U32P) {
                const uint64_t xc = canon_k<KOUT, DP>(x[c], m);
                x[c] = pack32((uint32_t)mulmod(lo32(av), lo32(xc), m), (uint32_t)mulmod(hi32(av), hi32(xc), m));
            }
            else x[c] = mulmod(av, canon_k<KOUT, DP>(x[c], m), m);  // canonical x canonical (U32 mode: generic 64-bit reduction, once per coefficient)
        }
        inv_stages<R, S0, 1, DP, PASS == 0>(x, twi, TB, m);
        if (OUT == IO_GLOBAL) {
            uint64_t* dst = gout + (size_t)poly * N;
#pragma unroll
            for (int c = 0; c < E; ++c) stream_store(dst + (base | (uint32_t)c), scale_word<DP>(x[c], ninv, m));
        } else {
#pragma unroll
            for (int c = 0; c < E; ++c) slot_store<DP, slot_hoisted<L, DP>()>(sr, swzm<DP>((uint32_t)c), x[c]);
        }
    }
}

// Inverse pass PASS (run in order P-1, P-2, .., 0).  The first one executed (PASS == P-1)
// may read the reference's input order (bit-reversed positions) from global memory; the
// last one executed (PASS == 0) multiplies by N^-1 (`ninv`, Shoup pair) and stores canonical
// values in natural order.
template <int L, int DP, int PASS, int IN, int OUT, bool BITREV_IN = true, int IPT = 0, int KSTART = 1, bool SUB = false, bool PIPE = false, int PK = L>
FHEB_HD void inv_pass(uint32_t tid, uint32_t nthreads, uint32_t polys, const uint64_t* gin, uint64_t* gout,
                      uint64_t* smem, const Tw* __restrict__ tw, const Tw ninv, const ModQ& m,
                      const GlobalMap map = GlobalMap{0, 0}, const uint64_t* gin_rest = nullptr) {
    constexpr int R = Plan<PK>::R[PASS];
    constexpr int E = 1 << R;
    constexpr int S0 = plan_s0<PK, PASS>();
    constexpr int EB = L - S0 - R;
    constexpr int KIN = plan_inv_kin<PK, DP, PASS, KSTART>();
    constexpr bool FIRST = (PASS == Plan<PK>::P - 1);
    constexpr uint32_t N = 1u << L;
    constexpr uint32_t ITEMS = N >> R;
    const uint32_t units = units_of<DP>(polys);
    constexpr bool EXT_IN = (IN == IO_GLOBAL || IN == IO_LANDING || IN == IO_LANDING_PART);  // caller words: from global memory or from the landing buffer
    constexpr bool BRTW = FIRST && EXT_IN && BITREV_IN && S0 > 0;  // see fwd_pass
    static_assert(!BRTW || (EB == 0 && S0 == L - R), "the last pass covers the lowest position bits");
    static_assert(IN == IO_SMEM || FIRST, "only the first executed pass reads caller words");
    static_assert(OUT == IO_SMEM || PASS == 0, "only the last executed pass writes global memory");

    constexpr int NW = E - 1;  // IPT > 0: all twiddles of the thread's IPT items requested up front (see fwd_pass)
    constexpr int TRIPS = IPT > 0 ? IPT : 1;
    Tw wall[TRIPS][NW];
    if constexpr (IPT > 0) {
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const uint32_t U = tid + (uint32_t)k * nthreads;
            if (U < units * ITEMS) {
                const uint32_t t0 = U & (ITEMS - 1);
                uint32_t u = t0;
                if (FIRST && EXT_IN && BITREV_IN) u = bitrev_rt(u, L - R);
                const uint32_t base = ((u >> EB) << (EB + R)) | (u & ((1u << EB) - 1u));
                load_item_tw<R, S0, DP>(tw, BRTW ? (N + t0) : plan_tw_offset<PK, PASS>() + (S0 ? (base >> (L - S0)) : 0u), wall[k]);
            }
        }
    }
    constexpr bool PIPE_IN = PIPE && FIRST && IN == IO_GLOBAL && BITREV_IN && !SUB && IPT == 0 && DP != MODE_U32P;  // see fwd_pass
    uint64_t xnext[PIPE_IN ? E : 1];
    bool have_next = false;
    for (uint32_t U0 = tid; U0 < (IPT > 0 ? tid + 1 : units * ITEMS); U0 += nthreads) {
#pragma unroll
      for (int k = 0; k < TRIPS; ++k) {
        const uint32_t U = U0 + (uint32_t)k * nthreads;
        if (IPT > 0 && U >= units * ITEMS) break;
        const uint32_t poly = U >> (L - R);
        uint32_t u = U & (ITEMS - 1);
        const uint32_t t = u;
        if (FIRST && EXT_IN && BITREV_IN) u = bitrev_rt(t, L - R);
        const uint32_t base = ((u >> EB) << (EB + R)) | (u & ((1u << EB) - 1u));
        uint64_t x[E];
        if constexpr (PIPE_IN) {  // plain transform, reference-order input: this item's words were requested one trip ago
            if (have_next) {
#pragma unroll
                for (int c = 0; c < E; ++c) x[c] = xnext[c];
            } else {
                const uint64_t* src = gin + (size_t)poly * N;
#pragma unroll
                for (int c = 0; c < E; ++c) x[c] = stream_load(src + ((bitrev_c((uint32_t)c, R) << (L - R)) | t));
            }
            const uint32_t Un = U + nthreads;
            have_next = Un < units * ITEMS;
            if (have_next) {
                const uint64_t* srcn = gin + (size_t)(Un >> (L - R)) * N;
                const uint32_t tn = Un & (ITEMS - 1);
#pragma unroll
                for (int c = 0; c < E; ++c) xnext[c] = stream_load(srcn + ((bitrev_c((uint32_t)c, R) << (L - R)) | tn));
            }
            load_words<DP, E>(x, m);
        } else if constexpr (IN == IO_LANDING_PART) {  // see fwd_pass; reference-order input: word (bitrev(c) << (L - R)) | t
            static_assert(FIRST && BITREV_IN && !SUB && (LAND_PART_WORDS & ((1u << (L - R)) - 1u)) == 0, "whole rows of the first executed pass are landed");
#pragma unroll
            for (int c = 0; c < E; ++c)
                x[c] = part_word(gin, gin_rest, (bitrev_c((uint32_t)c, R) << (L - R)) | t, (bitrev_c((uint32_t)c, R) << (L - R)) < LAND_PART_WORDS);
            load_words<DP, E>(x, m);
        } else if (EXT_IN) {
            if (SUB && BITREV_IN) load_unit<DP, IN, E>(x, gin, poly, polys, N, [&](int c) { return (((bitrev_c((uint32_t)c, R) << (L - R)) | t) << map.shift) | map.low; }, m);
            else if (BITREV_IN) load_unit<DP, IN, E>(x, gin, poly, polys, N, [&](int c) { return (bitrev_c((uint32_t)c, R) << (L - R)) | t; }, m);
            else load_unit<DP, IN, E>(x, gin, poly, polys, N, [&](int c) { return base | ((uint32_t)c << EB); }, m);
        } else {
            const SlotRef<DP> sr = slot_ref<DP>(smem, poly, N, swzm<DP>(base));
#pragma unroll
            for (int c = 0; c < E; ++c) x[c] = slot_load<DP, slot_hoisted<L, DP>()>(sr, swzm<DP>((uint32_t)c << EB));
        }
        const uint32_t TB = BRTW ? (N + t) : plan_tw_offset<PK, PASS>() + (S0 ? (base >> (L - S0)) : 0u);
        if constexpr (IPT > 0) inv_stages<R, S0, KIN, DP, (PASS == 0 && !SUB), R - 1, true>(x, wall[k], 0u, m);
        else inv_stages<R, S0, KIN, DP, (PASS == 0 && !SUB)>(x, tw, TB, m);
        if (OUT == IO_GLOBAL) {
#pragma unroll
            for (int c = 0; c < E; ++c) store_unit_word<DP>(gout, poly, polys, N, base | ((uint32_t)c << EB), scale_word<DP>(x[c], ninv, m));
        } else {
            const SlotRef<DP> sw = slot_ref<DP>(smem, poly, N, swzm<DP>(base));
#pragma unroll
            for (int c = 0; c < E; ++c) slot_store<DP, slot_hoisted<L, DP>()>(sw, swzm<DP>((uint32_t)c << EB), x[c]);
        }
      }
    }
}

}  // namespace fheb
