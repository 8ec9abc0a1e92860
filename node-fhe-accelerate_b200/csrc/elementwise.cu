// Element-wise modular kernels and the batched multi-limb Montgomery kernels (HBM-bound:
// 128-bit coalesced loads/stores, grid sized to the SM count, grid-stride loops).
//
// C-ABI entry points here (include/fheb200.h): fheb_modadd_batch, fheb_modsub_batch,
// fheb_modmul_batch, fheb_modneg_batch, fheb_modmul_scalar_batch, fheb_mlimb_montmul_batch,
// fheb_mlimb_add_batch, fheb_mlimb_sub_batch, fheb_mlimb_constants.
#include "elementwise.hpp"
#include "modarith.cuh"
#include "runtime.hpp"

namespace fheb {

enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_NEG = 3, OP_SCALAR = 4 };  // keep in step with elementwise.hpp

template <int OP>
__device__ __forceinline__ uint64_t apply(uint64_t a, uint64_t b, const ModQ& m) {
    if (OP == OP_ADD) return addmod_canon(canon_any(a, m), canon_any(b, m), m.q);  // modular_arithmetic.cpp:122-136
    if (OP == OP_SUB) return submod_canon(canon_any(a, m), canon_any(b, m), m.q);  // :138-153
    if (OP == OP_MUL) return mulmod_any(a, b, m);                                  // polynomial_ring.cpp:523-526
    if (OP == OP_NEG) return a == 0 ? 0 : m.q - a;                                 // polynomial_ring.cpp:356 (no input reduction)
    return mulmod_any(a, b, m);                                                    // OP_SCALAR: b = scalar % q, :466-470
}

// r[i] = op(a[i], b[i]); VEC = 2 uses 16-byte accesses (pointers 16-byte aligned, count even part)
template <int OP>
__global__ void __launch_bounds__(256) elementwise_kernel(const uint64_t* a, const uint64_t* b, uint64_t scalar,
                                                          uint64_t* r, size_t count, const ModQ m, int vec_ok) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (vec_ok) {
        const size_t pairs = count / 2;
        const ulonglong2* a2 = reinterpret_cast<const ulonglong2*>(a);
        const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(b);
        ulonglong2* r2 = reinterpret_cast<ulonglong2*>(r);
        for (size_t i = tid; i < pairs; i += stride) {
            const ulonglong2 av = a2[i];
            ulonglong2 bv;
            if (OP == OP_NEG) bv = make_ulonglong2(0, 0);
            else if (OP == OP_SCALAR) bv = make_ulonglong2(scalar, scalar);
            else bv = b2[i];
            ulonglong2 out;
            out.x = apply<OP>(av.x, bv.x, m);
            out.y = apply<OP>(av.y, bv.y, m);
            r2[i] = out;
        }
        done = pairs * 2;
    }
    for (size_t i = done + tid; i < count; i += stride) {
        const uint64_t bv = (OP == OP_NEG) ? 0 : (OP == OP_SCALAR ? scalar : b[i]);
        r[i] = apply<OP>(a[i], bv, m);
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

unsigned stream_grid(size_t work_items, int threads, int blocks_per_sm) {
    const size_t want = (work_items + threads - 1) / threads;
    const size_t cap = (size_t)ctx().sm_count * blocks_per_sm;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

template <int OP>
static int launch_elementwise(const uint64_t* a, const uint64_t* b, uint64_t scalar, uint64_t* r, size_t count,
                              uint64_t q, cudaStream_t s) {
    const ModQ m = make_modq(q);
    const int vec_ok = aligned16(a) && aligned16(r) && (b == nullptr || aligned16(b));
    elementwise_kernel<OP><<<stream_grid(count / 2 + 1, 256, 8), 256, 0, s>>>(a, b, scalar, r, count, m, vec_ok);
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

int elementwise_device(int op, const uint64_t* a, const uint64_t* b, uint64_t scalar, uint64_t* r, size_t count,
                       uint64_t q, cudaStream_t s) {
    if (count == 0) return FHEB_OK;
    switch (op) {
        case OP_ADD: return launch_elementwise<OP_ADD>(a, b, 0, r, count, q, s);
        case OP_SUB: return launch_elementwise<OP_SUB>(a, b, 0, r, count, q, s);
        case OP_MUL: return launch_elementwise<OP_MUL>(a, b, 0, r, count, q, s);
        case OP_NEG: return launch_elementwise<OP_NEG>(a, nullptr, 0, r, count, q, s);
        default: return launch_elementwise<OP_SCALAR>(a, nullptr, scalar % q, r, count, q, s);
    }
}

static int elementwise_entry(int op, const uint64_t* a, const uint64_t* b, uint64_t scalar, uint64_t* r, size_t count,
                             uint64_t q, void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(q >= 2, "Modulus must be at least 2");
    if (count == 0) return FHEB_OK;
    const bool binary = (op == OP_ADD || op == OP_SUB || op == OP_MUL);
    FHEB_REQUIRE(a != nullptr && r != nullptr && (!binary || b != nullptr), "operand pointers must not be null");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = count * 8;
    if (count >= (1u << 21) && all_host({a, b, r})) {  // large host vectors: pipelined in 4096-word items
        const size_t item = 4096, items = count / item, tail = count - items * item;
        FHEB_TRY(run_host_pipeline(items, {{a, item * 8, 0, true, false}, {binary ? b : nullptr, item * 8, 0, true, false}, {r, item * 8, 0, false, true}},
                                   [&](void* const* d, size_t, size_t n, cudaStream_t ps) {
                                       return elementwise_device(op, static_cast<const uint64_t*>(d[0]), static_cast<const uint64_t*>(d[1]), scalar,
                                                                 static_cast<uint64_t*>(d[2]), n * item, q, ps);
                                   }));
        if (tail == 0) return FHEB_OK;
        return elementwise_entry(op, a + items * item, binary ? b + items * item : nullptr, scalar, r + items * item, tail, q, stream);
    }
    Staged sa, sb, sr;
    FHEB_TRY(sa.bind(a, bytes, true, false, s));
    if (binary) {
        if (b == a) FHEB_TRY(sb.bind_alias(sa, false));
        else FHEB_TRY(sb.bind(b, bytes, true, false, s));
    }
    if (r == a) FHEB_TRY(sr.bind_alias(sa, true));
    else if (binary && r == b) FHEB_TRY(sr.bind_alias(sb, true));
    else FHEB_TRY(sr.bind(r, bytes, false, true, s));
    FHEB_TRY(elementwise_device(op, sa.ptr<const uint64_t>(), binary ? sb.ptr<const uint64_t>() : nullptr, scalar,
                                sr.ptr<uint64_t>(), count, q, s));
    FHEB_TRY(sa.finish());
    FHEB_TRY(sb.finish());
    FHEB_TRY(sr.finish());
    return sync_if_staged(s, {&sa, &sb, &sr});
}

// ---- multi-limb ------------------------------------------------------------------------
// Word-for-word the reference's algorithms (cpp/src/modular_arithmetic.cpp): schoolbook
// product (:525-545), word-serial Montgomery reduction with the carry propagation bound of
// :588 (a carry out of the top limb is dropped, as there), comparison from the top limb
// (:547-556) and the `a < b + borrow` borrow test of sub_limbs (:517-518).

struct MlimbQ {
    uint64_t q[8];
    uint64_t q_inv;
};

template <int LIMBS>
__device__ __forceinline__ bool limbs_less(const uint64_t* a, const uint64_t* b) {
#pragma unroll
    for (int i = LIMBS - 1; i >= 0; --i) {
        if (a[i] < b[i]) return true;
        if (a[i] > b[i]) return false;
    }
    return false;
}

template <int LIMBS>
__device__ __forceinline__ uint64_t add_limbs(const uint64_t* a, const uint64_t* b, uint64_t* r) {
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < LIMBS; ++i) {
        const u128 s = (u128)a[i] + b[i] + carry;
        r[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
    }
    return carry;
}

template <int LIMBS>
__device__ __forceinline__ uint64_t sub_limbs(const uint64_t* a, const uint64_t* b, uint64_t* r) {
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < LIMBS; ++i) {
        const uint64_t al = a[i], bl = b[i];
        const uint64_t diff = al - bl - borrow;
        borrow = (al < bl + borrow) ? 1 : 0;  // u64 wrap mirrored on purpose
        r[i] = diff;
    }
    return borrow;
}

template <int LIMBS, int OP>  // OP: 0 montmul, 1 add, 2 sub
__device__ __forceinline__ void mlimb_apply(const uint64_t* a, const uint64_t* b, uint64_t* out, const MlimbQ& mq) {
    if (OP == 0) {
        uint64_t t[2 * LIMBS];
#pragma unroll
        for (int i = 0; i < 2 * LIMBS; ++i) t[i] = 0;
#pragma unroll
        for (int i = 0; i < LIMBS; ++i) {
            uint64_t carry = 0;
#pragma unroll
            for (int j = 0; j < LIMBS; ++j) {
                const u128 p = (u128)a[i] * b[j] + t[i + j] + carry;
                t[i + j] = (uint64_t)p;
                carry = (uint64_t)(p >> 64);
            }
            t[i + LIMBS] = carry;
        }
#pragma unroll
        for (int i = 0; i < LIMBS; ++i) {
            const uint64_t mfac = t[i] * mq.q_inv;
            uint64_t carry = 0;
#pragma unroll
            for (int j = 0; j < LIMBS; ++j) {
                const u128 p = (u128)mfac * mq.q[j] + t[i + j] + carry;
                t[i + j] = (uint64_t)p;
                carry = (uint64_t)(p >> 64);
            }
#pragma unroll
            for (int j = LIMBS; j < 2 * LIMBS - i; ++j) {  // `&& carry` of :588 is a no-op when carry == 0
                const u128 s = (u128)t[i + j] + carry;
                t[i + j] = (uint64_t)s;
                carry = (uint64_t)(s >> 64);
            }
        }
        if (!limbs_less<LIMBS>(t + LIMBS, mq.q)) sub_limbs<LIMBS>(t + LIMBS, mq.q, out);
        else {
#pragma unroll
            for (int i = 0; i < LIMBS; ++i) out[i] = t[LIMBS + i];
        }
    } else if (OP == 1) {
        uint64_t s[LIMBS];
        const uint64_t carry = add_limbs<LIMBS>(a, b, s);
        if (carry || !limbs_less<LIMBS>(s, mq.q)) sub_limbs<LIMBS>(s, mq.q, out);
        else {
#pragma unroll
            for (int i = 0; i < LIMBS; ++i) out[i] = s[i];
        }
    } else {
        uint64_t d[LIMBS];
        const uint64_t borrow = sub_limbs<LIMBS>(a, b, d);
        if (borrow) add_limbs<LIMBS>(d, mq.q, out);
        else {
#pragma unroll
            for (int i = 0; i < LIMBS; ++i) out[i] = d[i];
        }
    }
}

template <int LIMBS, int OP>
__global__ void __launch_bounds__(256) mlimb_kernel(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count,
                                                    const MlimbQ mq, int vec_ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) {
        uint64_t x[LIMBS], y[LIMBS], o[LIMBS];
        if (vec_ok && (LIMBS % 2 == 0)) {
            const ulonglong2* a2 = reinterpret_cast<const ulonglong2*>(a + e * LIMBS);
            const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(b + e * LIMBS);
#pragma unroll
            for (int i = 0; i < LIMBS / 2; ++i) {
                const ulonglong2 av = a2[i], bv = b2[i];
                x[2 * i] = av.x;
                x[2 * i + 1] = av.y;
                y[2 * i] = bv.x;
                y[2 * i + 1] = bv.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < LIMBS; ++i) {
                x[i] = a[e * LIMBS + i];
                y[i] = b[e * LIMBS + i];
            }
        }
        mlimb_apply<LIMBS, OP>(x, y, o, mq);
        if (vec_ok && (LIMBS % 2 == 0)) {
            ulonglong2* r2 = reinterpret_cast<ulonglong2*>(r + e * LIMBS);
#pragma unroll
            for (int i = 0; i < LIMBS / 2; ++i) r2[i] = make_ulonglong2(o[2 * i], o[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < LIMBS; ++i) r[e * LIMBS + i] = o[i];
        }
    }
}

template <int LIMBS>
static int launch_mlimb(int op, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, const MlimbQ& mq,
                        cudaStream_t s) {
    const int vec_ok = aligned16(a) && aligned16(b) && aligned16(r);
    const unsigned grid = stream_grid(count, 256, 8);
    if (op == 0) mlimb_kernel<LIMBS, 0><<<grid, 256, 0, s>>>(a, b, r, count, mq, vec_ok);
    else if (op == 1) mlimb_kernel<LIMBS, 1><<<grid, 256, 0, s>>>(a, b, r, count, mq, vec_ok);
    else mlimb_kernel<LIMBS, 2><<<grid, 256, 0, s>>>(a, b, r, count, mq, vec_ok);
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

static int mlimb_entry(int op, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint32_t limbs,
                       const uint64_t* q_limbs, uint64_t q_inv, void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(limbs >= 1 && limbs <= 8, "limb count must be between 1 and 8");
    FHEB_REQUIRE(q_limbs != nullptr, "modulus limbs must not be null");
    bool zero = true;
    for (uint32_t i = 0; i < limbs; ++i) zero = zero && q_limbs[i] == 0;
    // message follows MultiLimbMontgomeryConstants, cpp/src/modular_arithmetic.cpp:474-476
    FHEB_REQUIRE(!zero && (q_limbs[0] & 1) != 0, "Modulus must be odd and non-zero for Montgomery arithmetic");
    if (count == 0) return FHEB_OK;
    FHEB_REQUIRE(a != nullptr && b != nullptr && r != nullptr, "operand pointers must not be null");
    MlimbQ mq{};
    for (uint32_t i = 0; i < limbs; ++i) mq.q[i] = q_limbs[i];
    mq.q_inv = q_inv;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = count * limbs * 8;
    auto launch = [&](const uint64_t* da, const uint64_t* db, uint64_t* dr, size_t n, cudaStream_t ps) {
        switch (limbs) {
            case 1: return launch_mlimb<1>(op, da, db, dr, n, mq, ps);
            case 2: return launch_mlimb<2>(op, da, db, dr, n, mq, ps);
            case 3: return launch_mlimb<3>(op, da, db, dr, n, mq, ps);
            case 4: return launch_mlimb<4>(op, da, db, dr, n, mq, ps);
            case 5: return launch_mlimb<5>(op, da, db, dr, n, mq, ps);
            case 6: return launch_mlimb<6>(op, da, db, dr, n, mq, ps);
            case 7: return launch_mlimb<7>(op, da, db, dr, n, mq, ps);
            default: return launch_mlimb<8>(op, da, db, dr, n, mq, ps);
        }
    };
    if (bytes >= (16u << 20) && all_host({a, b, r})) {
        const size_t row = (size_t)limbs * 8;
        return run_host_pipeline(count, {{a, row, 0, true, false}, {b, row, 0, true, false}, {r, row, 0, false, true}},
                                 [&](void* const* d, size_t, size_t n, cudaStream_t ps) {
                                     return launch(static_cast<const uint64_t*>(d[0]), static_cast<const uint64_t*>(d[1]), static_cast<uint64_t*>(d[2]), n, ps);
                                 });
    }
    Staged sa, sb, sr;
    FHEB_TRY(sa.bind(a, bytes, true, false, s));
    if (b == a) FHEB_TRY(sb.bind_alias(sa, false));
    else FHEB_TRY(sb.bind(b, bytes, true, false, s));
    if (r == a) FHEB_TRY(sr.bind_alias(sa, true));
    else if (r == b) FHEB_TRY(sr.bind_alias(sb, true));
    else FHEB_TRY(sr.bind(r, bytes, false, true, s));
    const uint64_t *da = sa.ptr<const uint64_t>(), *db = sb.ptr<const uint64_t>();
    uint64_t* dr = sr.ptr<uint64_t>();
    int rc = launch(da, db, dr, count, s);
    FHEB_TRY(rc);
    FHEB_TRY(sa.finish());
    FHEB_TRY(sb.finish());
    FHEB_TRY(sr.finish());
    return sync_if_staged(s, {&sa, &sb, &sr});
}

// ---- host-side constants (MultiLimbMontgomeryConstants ctor, modular_arithmetic.cpp:347-486)
static bool h_less(const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) {  // operator<, :315-329
    const size_t n = a.size() > b.size() ? a.size() : b.size();
    for (size_t i = n; i > 0; --i) {
        const uint64_t x = (i - 1 < a.size()) ? a[i - 1] : 0, y = (i - 1 < b.size()) ? b[i - 1] : 0;
        if (x < y) return true;
        if (x > y) return false;
    }
    return false;
}

// multi_limb_mod_proper, :361-429: bit-serial restoring division exactly as written there
// (shifted modulus truncated to the remainder width, strict greater-than test).
static std::vector<uint64_t> h_mod_proper(const std::vector<uint64_t>& a, const std::vector<uint64_t>& q) {
    const size_t l = q.size();
    if (a.size() <= l && h_less(a, q)) {
        std::vector<uint64_t> r = a;
        r.resize(l, 0);
        return r;
    }
    std::vector<uint64_t> rem = a;
    if (rem.size() < l) rem.resize(l, 0);
    const size_t rs = rem.size();
    std::vector<uint64_t> sh(rs);
    for (int bit = (int)(rs * 64) - 1; bit >= 0; --bit) {
        const size_t ls = (size_t)bit / 64, bs = (size_t)bit % 64;
        std::fill(sh.begin(), sh.end(), 0);
        for (size_t i = 0; i < l && i + ls < rs; ++i) {
            if (bs == 0) sh[i + ls] = q[i];
            else {
                sh[i + ls] |= q[i] << bs;
                if (i + ls + 1 < rs) sh[i + ls + 1] = q[i] >> (64 - bs);
            }
        }
        bool can = false;
        for (int i = (int)rs - 1; i >= 0; --i) {
            if (rem[i] > sh[i]) { can = true; break; }
            if (rem[i] < sh[i]) break;
        }
        if (can) {
            uint64_t borrow = 0;
            for (size_t i = 0; i < rs; ++i) {
                const uint64_t r = rem[i], s = sh[i];
                rem[i] = r - s - borrow;
                borrow = (r < s + borrow) ? 1 : 0;
            }
        }
    }
    rem.resize(l);
    return rem;
}

}  // namespace fheb

using namespace fheb;

extern "C" {

int fheb_modadd_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint64_t modulus, void* stream) {
    return elementwise_entry(OP_ADD, a, b, 0, r, count, modulus, stream);
}
int fheb_modsub_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint64_t modulus, void* stream) {
    return elementwise_entry(OP_SUB, a, b, 0, r, count, modulus, stream);
}
int fheb_modmul_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint64_t modulus, void* stream) {
    return elementwise_entry(OP_MUL, a, b, 0, r, count, modulus, stream);
}
int fheb_modneg_batch(const uint64_t* a, uint64_t* r, size_t count, uint64_t modulus, void* stream) {
    return elementwise_entry(OP_NEG, a, nullptr, 0, r, count, modulus, stream);
}
int fheb_modmul_scalar_batch(const uint64_t* a, uint64_t scalar, uint64_t* r, size_t count, uint64_t modulus, void* stream) {
    return elementwise_entry(OP_SCALAR, a, nullptr, scalar, r, count, modulus, stream);
}

int fheb_mlimb_montmul_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint32_t limbs,
                             const uint64_t* q_limbs, uint64_t q_inv, void* stream) {
    return mlimb_entry(0, a, b, r, count, limbs, q_limbs, q_inv, stream);
}
int fheb_mlimb_add_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint32_t limbs,
                         const uint64_t* q_limbs, void* stream) {
    return mlimb_entry(1, a, b, r, count, limbs, q_limbs, 0, stream);
}
int fheb_mlimb_sub_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint32_t limbs,
                         const uint64_t* q_limbs, void* stream) {
    return mlimb_entry(2, a, b, r, count, limbs, q_limbs, 0, stream);
}

int fheb_mlimb_constants(const uint64_t* q_limbs, uint32_t limbs, uint64_t* consts) {
    FHEB_REQUIRE(limbs >= 1 && limbs <= 8, "limb count must be between 1 and 8");
    FHEB_REQUIRE(q_limbs != nullptr && consts != nullptr, "pointers must not be null");
    std::vector<uint64_t> q(q_limbs, q_limbs + limbs);
    bool zero = true;
    for (uint64_t v : q) zero = zero && v == 0;
    FHEB_REQUIRE(!zero && (q[0] & 1) != 0, "Modulus must be odd and non-zero for Montgomery arithmetic");
    uint64_t x = q[0];  // compute_q_inv_limb, :347-358
    for (int i = 0; i < 5; ++i) x = x * (2 - q[0] * x);
    consts[0] = (~x) + 1;
    std::vector<uint64_t> r(limbs + 1, 0);  // compute_r_mod_q, :432-440
    r[limbs] = 1;
    const std::vector<uint64_t> r1 = h_mod_proper(r, q);
    std::vector<uint64_t> prod(2 * limbs, 0);  // compute_r2_mod_q, :443-468
    for (uint32_t i = 0; i < limbs; ++i) {
        uint64_t carry = 0;
        for (uint32_t j = 0; j < limbs; ++j) {
            u128 p = (u128)r1[i] * r1[j];
            p += prod[i + j];
            p += carry;
            prod[i + j] = (uint64_t)p;
            carry = (uint64_t)(p >> 64);
        }
        prod[i + limbs] = carry;
    }
    const std::vector<uint64_t> r2 = h_mod_proper(prod, q);
    for (uint32_t i = 0; i < limbs; ++i) {
        consts[1 + i] = r1[i];
        consts[1 + limbs + i] = r2[i];
    }
    return FHEB_OK;
}

}  // extern "C"
