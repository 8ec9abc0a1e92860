// One CMux / external product of the TFHE blind rotation as a sequence of barrier-separated
// phases over one thread block's shared memory (host/device code: the same functions are run
// sequentially by tests/host_emulation to validate them without a GPU).
//
// Reference semantics (cpp/src/bootstrap_engine.cpp):
//   cmux(ggsw, ct0, ct1) = ct0 + external_product(ct1 - ct0, ggsw)                    :520-540
//   external_product     = sum over rows (c, l) of T^-1( T(digit_{c,l}) . T(ggsw[row][j]) ),
//                          accumulated in the coefficient domain, j = output component    :431-518
//   digit_{c,l}          = low-bit gadget digit of component c, centred without carry     :152-185
//   blind-rotate step    : ct1 = X^rot * acc (rotate_polynomial :122-145), ct0 = acc      :562-576
// T is Z_q-linear, so sum_rows T^-1(D_row . G_row,j) == T^-1(sum_rows D_row . G_row,j) exactly:
// the GGSW is transformed ONCE at key upload (kept in HBM in position order with its Shoup
// companions) and each step costs (k+1)*L forward and (k+1) inverse transforms instead of the
// reference's 10*L.  The last forward pass, the multiply-accumulate against the GGSW and the
// first inverse pass run on the same register-resident positions, so transformed digits
// never touch memory and no bit-reversal happens anywhere.
#pragma once
#include "ntt_core.cuh"

namespace fheb {

struct BootStep {          // block-uniform description of one step
    uint64_t* acc;         // smem [KP1][N], natural index: ct0 / running accumulator
    uint64_t* work;        // smem [max(rows, KP1)][N], swizzled index
    const uint64_t* diff;  // [KP1][N] natural index: precomputed ct1 - ct0 / the GLWE itself (shared or global); null = rotate acc
    const uint64_t* diff_sub;  // when non-null: the digits come from diff - diff_sub (both reduced first), read from global memory
    const Tw* ggsw;        // global [rows][E][N/E][KP1] (position u*E + e, component j at [e][u][j]; E = width of the last pass), times N^-1: (value, Shoup companion) pairs, or doubles in DP mode
    uint64_t* gout;        // when non-null the final pass stores here ([KP1][N], global) instead of acc
    uint32_t rot;          // normalised rotation in [0, 2N) (used when diff == null)
    uint32_t levels;
    uint32_t rows;         // digit rows summed by the multiply-accumulate: KP1 * levels (external product), levels (relinearisation)
    uint32_t base_log;
    uint32_t add_acc;      // 1: result = acc + product (cmux); 0: product only (external product)
    uint32_t maybe_raw;    // 1: acc may still hold unreduced caller words (>= q); cleared after the first executed step
};

// normalisation of rotate_polynomial (cpp/src/bootstrap_engine.cpp:127-128), int32 arithmetic as written
FHEB_HD uint32_t normalise_rotation(int32_t rotation, uint32_t N) {
    const int32_t two_n = 2 * (int32_t)N;
    return (uint32_t)(((rotation % two_n) + two_n) % two_n);
}

// (X^rot * p)[j] for rot in [0, 2N): p[j - rot] or its negation `(q - v) % q` in u64 arithmetic (:131-143)
FHEB_HD uint64_t rotated_at(const uint64_t* p, uint32_t j, uint32_t rot, uint32_t N, const ModQ& m) {
    const uint32_t s = (j + 2u * N - rot) & (2u * N - 1u);
    if (s < N) return p[s];
    const uint64_t t = m.q - p[s - N];
    return t >= m.q ? reduce64(t, m) : t;
}

// gadget digit l of coefficient c (decompose_polynomial, :165-178), as a canonical residue
// `small_base` (base <= q, block-uniform): both branches are already below q
FHEB_HD uint64_t gadget_digit(uint64_t c, uint32_t shift, uint64_t mask, uint64_t base, bool small_base, const ModQ& m) {
    const uint64_t d = (c >> shift) & mask;
    const uint64_t v = (d > (base >> 1)) ? (m.q - (base - d)) : d;
    return small_base ? v : canon_any(v, m);
}

template <int L>
constexpr int boot_phases() { return 2 * Plan<L>::P - 1; }

// Bound (in units of q) on the multiply-accumulate sums that enter the inverse network: the integer path
// keeps its running sum in [0, 2q); the FP64 path adds |t| < q per gadget row and reduces only when more
// than BOOT_DP_ROWS_LAZY rows were summed (block-uniform), so the common 2-row case skips the reduction.
constexpr int BOOT_DP_ROWS_LAZY = 2;
// Degrees from 2^11 up may keep the accumulator / operands in global memory when the working rows fill the SM
// (boot_kernel.cuh); the smaller, hot kernels are compiled without that path.
constexpr int BOOT_GLOBAL_MIN_L = 11;
template <bool DP>
constexpr int boot_kacc() { return 2; }

// ---- phase 0: difference, decomposition and the first forward pass ------------------------
// RAWOK = false compiles the handling of unreduced accumulator words out (the lean blind-rotation kernel, used when
// the caller's test polynomial is canonical - the only source of such words).
template <int L, bool DP, int KP1, bool RAWOK = true>
FHEB_HD void boot_first_pass(uint32_t tid, uint32_t nthreads, const BootStep& s, const Tw* __restrict__ tw,
                             const ModQ& m) {
    constexpr int R = Plan<L>::R[0];
    constexpr int E = 1 << R;
    constexpr int EB = L - R;
    constexpr uint32_t N = 1u << L;
    constexpr uint32_t ITEMS = N >> R;
    static_assert(Plan<L>::P >= 2, "bootstrap kernels need at least two passes (N >= 32)");
    const uint64_t base = 1ull << s.base_log;
    const uint64_t mask = base - 1;
    const bool small_base = base <= m.q;
    for (uint32_t U = tid; U < (uint32_t)KP1 * ITEMS; U += nthreads) {
        const uint32_t c = U >> (L - R);
        const uint32_t u = U & (ITEMS - 1);
        uint64_t d[E];
        if (s.diff) {
            const uint64_t* src = s.diff + (size_t)c * N;
#pragma unroll
            for (int e = 0; e < E; ++e) d[e] = src[u | ((uint32_t)e << EB)];
            if (L >= BOOT_GLOBAL_MIN_L && s.diff_sub) {  // PolynomialRing::subtract(ct1, ct0) on the fly (large shapes: no staging buffer)
                const uint64_t* sub = s.diff_sub + (size_t)c * N;
#pragma unroll
                for (int e = 0; e < E; ++e) d[e] = submod_canon(canon_any(d[e], m), canon_any(sub[u | ((uint32_t)e << EB)], m), m.q);
            }
        } else {
            // X^rot * acc - acc.  After the first executed step every accumulator word is canonical; only
            // the words of a caller's test polynomial can be unreduced (s.maybe_raw).  Then one test per
            // item picks between the plain formulas and the reference's exact unsigned wrap-around ones.
            const uint64_t* a = s.acc + (size_t)c * N;
            uint64_t ct0[E], src[E];
            bool raw = false;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const uint32_t pos = u | ((uint32_t)e << EB);
                const uint32_t sidx = (pos + 2u * N - s.rot) & (2u * N - 1u);
                ct0[e] = a[pos];
                src[e] = a[sidx & (N - 1u)];
            }
            if ((RAWOK && s.maybe_raw)) {  // block-uniform: only before the first executed step of a blind rotation
#pragma unroll
                for (int e = 0; e < E; ++e) raw = raw || ct0[e] >= m.q || src[e] >= m.q;
            }
            if (!raw) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint32_t pos = u | ((uint32_t)e << EB);
                    const uint32_t sidx = (pos + 2u * N - s.rot) & (2u * N - 1u);
                    const uint64_t ct1 = (sidx < N) ? src[e] : (src[e] ? m.q - src[e] : 0);  // (q - v) % q for canonical v
                    d[e] = submod_canon(ct1, ct0[e], m.q);  // PolynomialRing::subtract(ct1, ct0), :527-530
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint32_t pos = u | ((uint32_t)e << EB);
                    const uint64_t ct1 = canon_any(rotated_at(a, pos, s.rot, N, m), m);
                    d[e] = submod_canon(ct1, canon_any(ct0[e], m), m.q);
                }
            }
        }
        const uint32_t pb = swz(u);
        for (uint32_t l = 0; l < s.levels; ++l) {
            const uint32_t shift = (s.levels - 1 - l) * s.base_log;
            uint64_t x[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const uint64_t g = gadget_digit(d[e], shift, mask, base, small_base, m);
                x[e] = DP ? double_to_bits(dp_from_uint(g)) : g;
            }
            fwd_stages<R, 0, 1, DP, true>(x, tw, 0u, m);
            const SlotRef<MODE_INT> dst = slot_ref<MODE_INT>(s.work, c * s.levels + l, N, pb);
#pragma unroll
            for (int e = 0; e < E; ++e) slot_store<MODE_INT>(dst, swz((uint32_t)e << EB), x[e]);
        }
    }
}

// ---- phase P-1: last forward pass of every digit row, multiply-accumulate with the GGSW, first
//      inverse pass of every output component -------------------------------------------------
template <int L, bool DP, int KP1>
FHEB_HD void boot_mid_pass(uint32_t tid, uint32_t nthreads, const BootStep& s, const Tw* __restrict__ twf,
                           const Tw* __restrict__ twi, const ModQ& m) {
    constexpr int PASS = Plan<L>::P - 1;
    constexpr int R = Plan<L>::R[PASS];
    constexpr int E = 1 << R;
    constexpr int S0 = plan_s0<L, PASS>();
    constexpr int KIN = plan_fwd_kin<L, DP, PASS>();
    constexpr uint32_t N = 1u << L;
    constexpr uint32_t ITEMS = N >> R;
    static_assert(L - S0 - R == 0, "the last forward pass covers the lowest position bits");
    const uint32_t rows = s.rows;
    for (uint32_t u = tid; u < ITEMS; u += nthreads) {
        const uint32_t base = u << R;
        const uint32_t pb = swz(base);
        const uint32_t TB = plan_tw_offset<L, PASS>() + (base >> (L - S0));
        uint64_t out[KP1][E];
#pragma unroll
        for (int j = 0; j < KP1; ++j)
#pragma unroll
            for (int e = 0; e < E; ++e) out[j][e] = 0;
        Tw wf[E - 1];  // this item's forward twiddles: the same for every digit row
        load_item_tw<R, S0, DP>(twf, TB, wf);
        for (uint32_t row = 0; row < rows; ++row) {
            const SlotRef<MODE_INT> src = slot_ref<MODE_INT>(s.work, row, N, pb);
            uint64_t x[E];
#pragma unroll
            for (int e = 0; e < E; ++e) x[e] = slot_load<MODE_INT>(src, swz((uint32_t)e));
            fwd_stages<R, S0, KIN, DP, false, 0, true>(x, wf, 0u, m);
            // key element (row, position u*E + e, component j) lives at ((row*E + e)*ITEMS + u)*KP1 + j: lanes
            // (consecutive u) read consecutive entries, and the k+1 components of one position are adjacent
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const uint32_t g0 = ((row * E + (uint32_t)e) * ITEMS + u) * (uint32_t)KP1;
                Tw w[KP1];
                if constexpr (DP && KP1 == 2) {  // one 16-byte load for both components
#if defined(__CUDA_ARCH__)
                    const ulonglong2 v2 = __ldg(reinterpret_cast<const ulonglong2*>(s.ggsw) + (g0 >> 1));
                    w[0].w = v2.x;
                    w[1].w = v2.y;
#else
                    w[0] = load_tw<DP>(s.ggsw, g0);
                    w[1] = load_tw<DP>(s.ggsw, g0 + 1);
#endif
                } else {
#pragma unroll
                    for (int j = 0; j < KP1; ++j) w[j] = load_tw<DP>(s.ggsw, g0 + (uint32_t)j);
                }
#pragma unroll
                for (int j = 0; j < KP1; ++j) {
                    if constexpr (DP) {  // |t| < q: plain sums (rows <= CAP_DP checked at key upload)
                        const double t = dp_mulmod(bits_to_double(x[e]), bits_to_double(w[j].w), m);
                        out[j][e] = double_to_bits(dp_add(bits_to_double(out[j][e]), t));
                    } else {  // t in [0, 2q) for any x; keep the running sum in [0, 2q)
                        out[j][e] = csub(out[j][e] + shoup_lazy(x[e], w[j].w, w[j].wp, m), m.q2);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < KP1; ++j) {
            if constexpr (DP) {  // more rows than the static bound covers: back to |v| <= q/2 + 1
                if (rows > (uint32_t)BOOT_DP_ROWS_LAZY) {
#pragma unroll
                    for (int e = 0; e < E; ++e) out[j][e] = double_to_bits(dp_reduce(bits_to_double(out[j][e]), m));
                }
            }
            inv_stages<R, S0, boot_kacc<DP>(), DP, false>(out[j], twi, TB, m);
            const SlotRef<MODE_INT> dst = slot_ref<MODE_INT>(s.work, j, N, pb);
#pragma unroll
            for (int e = 0; e < E; ++e) slot_store<MODE_INT>(dst, swz((uint32_t)e), out[j][e]);
        }
    }
}

// ---- last phase: final inverse pass, scaling by N^-1, `+ ct0` ------------------------------------
template <int L, bool DP, int KP1, bool RAWOK = true>
FHEB_HD void boot_final_pass(uint32_t tid, uint32_t nthreads, const BootStep& s, const Tw* __restrict__ twi,
                             const Tw ninv, const ModQ& m) {
    constexpr int R = Plan<L>::R[0];
    constexpr int E = 1 << R;
    constexpr int EB = L - R;
    constexpr int KIN = plan_inv_kin<L, DP, 0, boot_kacc<DP>()>();
    constexpr int KFIN = inv_pass_k(KIN, R, DP);
    constexpr uint32_t N = 1u << L;
    constexpr uint32_t ITEMS = N >> R;
    for (uint32_t U = tid; U < (uint32_t)KP1 * ITEMS; U += nthreads) {
        const uint32_t c = U >> (L - R);
        const uint32_t u = U & (ITEMS - 1);
        const uint32_t pb = swz(u);
        const SlotRef<MODE_INT> src = slot_ref<MODE_INT>(s.work, c, N, pb);
        uint64_t x[E];
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] = slot_load<MODE_INT>(src, swz((uint32_t)e << EB));
        inv_stages<R, 0, KIN, DP, true>(x, twi, 0u, m);
        const uint64_t* a = s.acc + (size_t)c * N;
        uint64_t* dst = (s.gout ? s.gout : s.acc) + (size_t)c * N;
        uint64_t v[E], a0[E];
        bool raw = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            v[e] = canon_k<KFIN, DP>(x[e], m);  // N^-1 is folded into the uploaded key (the transform is linear)
            a0[e] = s.add_acc ? a[u | ((uint32_t)e << EB)] : 0;
        }
        if ((RAWOK && s.maybe_raw)) {
#pragma unroll
            for (int e = 0; e < E; ++e) raw = raw || a0[e] >= m.q;
        }
        if (raw) {  // unreduced ct0 words (caller data): PolynomialRing::add reduces them first
#pragma unroll
            for (int e = 0; e < E; ++e) a0[e] = canon_any(a0[e], m);
        }
#pragma unroll
        for (int e = 0; e < E; ++e)
            dst[u | ((uint32_t)e << EB)] = s.add_acc ? csub(v[e] + a0[e], m.q) : v[e];  // both canonical, q < 2^62  // PolynomialRing::add(result, ct0), :533-537
    }
}

// Phase PH of a step; the caller puts a block barrier after every phase.
template <int L, bool DP, int KP1, int PH, bool RAWOK = true>
FHEB_HD void boot_phase(uint32_t tid, uint32_t nthreads, const BootStep& s, const Tw* __restrict__ twf,
                        const Tw* __restrict__ twi, const Tw ninv, const ModQ& m) {
    constexpr int P = Plan<L>::P;
    static_assert(PH >= 0 && PH < 2 * P - 1, "phase out of range");
    if constexpr (PH == 0) {
        boot_first_pass<L, DP, KP1, RAWOK>(tid, nthreads, s, twf, m);
    } else if constexpr (PH < P - 1) {
        fwd_pass<L, DP, PH, IO_SMEM, IO_SMEM>(tid, nthreads, s.rows, nullptr, nullptr, s.work, twf, m);
    } else if constexpr (PH == P - 1) {
        boot_mid_pass<L, DP, KP1>(tid, nthreads, s, twf, twi, m);
    } else if constexpr (PH < 2 * P - 2) {
        inv_pass<L, DP, 2 * P - 2 - PH, IO_SMEM, IO_SMEM, true, 0, boot_kacc<DP>()>(tid, nthreads, (uint32_t)KP1, nullptr, nullptr, s.work, twi, ninv, m);
    } else {
        boot_final_pass<L, DP, KP1, RAWOK>(tid, nthreads, s, twi, ninv, m);
    }
}

// Rotation amounts of blind_rotate (cpp/src/bootstrap_engine.cpp:558,564): wrapping u64 product,
// truncation to int32, sign applied before the normalisation of rotate_polynomial.
FHEB_HD int32_t lwe_rotation_raw(uint64_t word, bool negate, uint32_t N, uint64_t q) {
    const uint64_t scaled = (word * 2ull * (uint64_t)N + q / 2) / q;
    int32_t r = (int32_t)(uint32_t)scaled;
    if (negate) r = (int32_t)(0u - (uint32_t)r);
    return r;
}
FHEB_HD uint32_t lwe_rotation(uint64_t word, bool negate, uint32_t N, uint64_t q) {
    return normalise_rotation(lwe_rotation_raw(word, negate, N, q), N);
}
// The reference skips a step on the RAW int32 rotation being zero (:564-566), before rotate_polynomial normalises it.
// A raw rotation that is a non-zero multiple of 2N (a_i >= q - q/4N rounds up to 2N) therefore still executes a CMux
// with ct1 = X^0 * acc: diff = 0, product = 0, and `product + ct0` leaves the accumulator CANONICALISED (:533-537).
// Encoded for the kernel as the value 2N (step code masks it to rotation 0): such a step must run while the
// accumulator may still hold unreduced caller words and is the identity afterwards.
FHEB_HD uint32_t lwe_step_rotation(uint64_t word, uint32_t N, uint64_t q) {
    const int32_t raw = lwe_rotation_raw(word, false, N, q);
    const uint32_t rot = normalise_rotation(raw, N);
    return (rot == 0 && raw != 0) ? 2u * N : rot;
}

// sample_extract (cpp/src/bootstrap_engine.cpp:594-624): word idx of the LWE [k*N + 1] extracted from glwe
FHEB_HD uint64_t sample_extract_word(const uint64_t* glwe, uint32_t idx, uint32_t k, uint32_t N, const ModQ& m) {
    if (idx == k * N) return glwe[(size_t)k * N];
    const uint32_t i = idx / N, j = idx % N;
    const uint64_t* mask = glwe + (size_t)i * N;
    if (j == 0) return mask[0];
    const uint64_t t = m.q - mask[N - j];
    return t >= m.q ? reduce64(t, m) : t;  // (q - v) % q in u64 arithmetic
}

}  // namespace fheb
