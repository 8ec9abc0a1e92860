/*
 * TEST INFRASTRUCTURE ONLY - see fhe_oracle.h.  Plain-C restatement of the
 * reference's scalar CPU algorithms; each function names the reference
 * file:line it follows (paths relative to the reference's cpp/ directory).
 * Parity status: PINNED against oracle/_ref (the reference's own code built
 * here) and tests/golden/ - see tests/test_oracle.py.
 */
#include "fhe_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------ */
/* src/modular_arithmetic.cpp:122-136 - inputs reduced first, then one
 * conditional subtraction (wrap-around test for q >= 2^63).                 */
uint64_t orc_mod_add(uint64_t a, uint64_t b, uint64_t q) {
    a %= q;
    b %= q;
    uint64_t sum = a + b;
    if (sum < a || sum >= q) sum -= q;
    return sum;
}

/* src/modular_arithmetic.cpp:138-153 */
uint64_t orc_mod_sub(uint64_t a, uint64_t b, uint64_t q) {
    a %= q;
    b %= q;
    return (a >= b) ? (a - b) : (q - (b - a));
}

/* the `static_cast<__uint128_t>(x) * y % modulus` idiom used on every
 * butterfly, src/ntt_processor.cpp:299-300,362-363,377-378                  */
uint64_t orc_mod_mul(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }

/* src/ntt_processor.cpp:47-62 */
uint64_t orc_mod_pow(uint64_t base, uint64_t exp, uint64_t mod) {
    uint64_t result = 1;
    base %= mod;
    while (exp > 0) {
        if (exp & 1) result = (uint64_t)(((u128)result * base) % mod);
        base = (uint64_t)(((u128)base * base) % mod);
        exp >>= 1;
    }
    return result;
}

/* src/ntt_processor.cpp:64-90 - extended Euclid in SIGNED 64-bit           */
uint64_t orc_mod_inverse(uint64_t a, uint64_t m) {
    if (m == 1) return 0;
    int64_t m0 = (int64_t)m, x0 = 0, x1 = 1;
    int64_t a_s = (int64_t)(a % m), m_s = (int64_t)m;
    while (a_s > 1) {
        int64_t quo = a_s / m_s;
        int64_t t = m_s;
        m_s = a_s % m_s;
        a_s = t;
        t = x0;
        x0 = x1 - quo * x0;
        x1 = t;
    }
    if (x1 < 0) x1 += m0;
    return (uint64_t)x1;
}

/* src/ntt_processor.cpp:92-128 - smallest g >= 2 whose (q-1)/2N-th power has
 * order exactly 2N                                                          */
int orc_find_primitive_root(uint32_t degree, uint64_t q, uint64_t* root) {
    uint64_t two_n = (uint64_t)degree * 2;
    if ((q - 1) % two_n != 0) return -1;
    uint64_t exponent = (q - 1) / two_n;
    for (uint64_t g = 2; g < q; g++) {
        uint64_t omega = orc_mod_pow(g, exponent, q);
        if (orc_mod_pow(omega, two_n, q) == 1 && orc_mod_pow(omega, degree, q) == q - 1) {
            *root = omega;
            return 0;
        }
    }
    return -1;
}

/* src/ntt_processor.cpp:168-208 - tables of psi^i and psi^-i, i < N, standard
 * (non-Montgomery) form                                                     */
int orc_precompute_twiddles(uint32_t degree, uint64_t q, uint64_t* fwd, uint64_t* inv, uint64_t* scalars) {
    uint64_t psi;
    if (orc_find_primitive_root(degree, q, &psi) != 0) return -1;
    uint64_t psi_inv = orc_mod_inverse(psi, q);
    fwd[0] = 1;
    inv[0] = 1;
    for (uint32_t i = 1; i < degree; i++) {
        fwd[i] = (uint64_t)(((u128)fwd[i - 1] * psi) % q);
        inv[i] = (uint64_t)(((u128)inv[i - 1] * psi_inv) % q);
    }
    scalars[0] = psi;
    scalars[1] = psi_inv;
    scalars[2] = orc_mod_inverse(degree, q);
    return 0;
}

/* src/ntt_processor.cpp:38-45,214-223 */
static uint32_t bit_reverse(uint32_t index, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) {
        r = (r << 1) | (index & 1);
        index >>= 1;
    }
    return r;
}
static uint32_t log2_pow2(size_t n) {
    uint32_t l = 0;
    while (n > 1) {
        n >>= 1;
        l++;
    }
    return l;
}
void orc_bit_reverse_permutation(uint64_t* coeffs, size_t n) {
    uint32_t bits = log2_pow2(n);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t j = bit_reverse(i, bits);
        if (i < j) {
            uint64_t t = coeffs[i];
            coeffs[i] = coeffs[j];
            coeffs[j] = t;
        }
    }
}

/* src/ntt_processor.cpp:262-311 - bit-reverse, then DIT stages whose twiddle is
 * table[j * (n / group_size)] (table holds powers of a 2N-th root: SURVEY H4) */
void orc_forward_ntt(uint64_t* coeffs, size_t n, uint64_t q, const uint64_t* table) {
    uint32_t log_n = log2_pow2(n);
    orc_bit_reverse_permutation(coeffs, n);
    for (uint32_t stage = 0; stage < log_n; stage++) {
        uint32_t m = 1u << stage;
        uint32_t group = 2 * m;
        for (uint32_t k = 0; k < n; k += group) {
            for (uint32_t j = 0; j < m; j++) {
                uint64_t omega = table[j * (n / group)];
                uint64_t a = coeffs[k + j];
                uint64_t b = coeffs[k + j + m];
                uint64_t omega_b = (uint64_t)(((u128)omega * b) % q);
                coeffs[k + j] = orc_mod_add(a, omega_b, q);
                coeffs[k + j + m] = orc_mod_sub(a, omega_b, q);
            }
        }
    }
}

/* src/ntt_processor.cpp:325-380 - GS stages in reverse order, bit-reverse,
 * scale by N^-1                                                             */
void orc_inverse_ntt(uint64_t* coeffs, size_t n, uint64_t q, const uint64_t* table, uint64_t inv_n) {
    uint32_t log_n = log2_pow2(n);
    for (int32_t stage = (int32_t)log_n - 1; stage >= 0; stage--) {
        uint32_t m = 1u << stage;
        uint32_t group = 2 * m;
        for (uint32_t k = 0; k < n; k += group) {
            for (uint32_t j = 0; j < m; j++) {
                uint64_t omega_inv = table[j * (n / group)];
                uint64_t a = coeffs[k + j];
                uint64_t b = coeffs[k + j + m];
                uint64_t a_plus = orc_mod_add(a, b, q);
                uint64_t a_minus = orc_mod_sub(a, b, q);
                coeffs[k + j] = a_plus;
                coeffs[k + j + m] = (uint64_t)(((u128)a_minus * omega_inv) % q);
            }
        }
    }
    orc_bit_reverse_permutation(coeffs, n);
    for (size_t i = 0; i < n; i++) coeffs[i] = (uint64_t)(((u128)coeffs[i] * inv_n) % q);
}

/* src/ntt_processor.cpp:394-408 - serial loop over the batch                */
void orc_forward_ntt_batch(uint64_t* coeffs, size_t batch, size_t n, uint64_t q, const uint64_t* table) {
    for (size_t i = 0; i < batch; i++) orc_forward_ntt(coeffs + i * n, n, q, table);
}
void orc_inverse_ntt_batch(uint64_t* coeffs, size_t batch, size_t n, uint64_t q, const uint64_t* table,
                           uint64_t inv_n) {
    for (size_t i = 0; i < batch; i++) orc_inverse_ntt(coeffs + i * n, n, q, table, inv_n);
}

/* src/adaptive_dispatcher.cpp:171-205 - the SAME forward network fed the
 * inverse table, then scaling by n^(q-2) (Fermat).  The Montgomery detour of
 * :135-169 is value-preserving for canonical inputs, so it is not restated. */
void orc_fast_ntt_inverse(uint64_t* coeffs, size_t n, uint64_t q, const uint64_t* inv_table) {
    orc_forward_ntt(coeffs, n, q, inv_table);
    uint64_t n_inv = orc_mod_pow((uint64_t)n, q - 2, q);
    for (size_t i = 0; i < n; i++) coeffs[i] = (uint64_t)(((u128)coeffs[i] * n_inv) % q);
}

/* ------------------------------------------------------------------------ */
/* src/polynomial_ring.cpp:263-272,372-393 */
void orc_poly_add(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n, uint64_t q) {
    for (size_t i = 0; i < n; i++) r[i] = orc_mod_add(a[i], b[i], q);
}
/* src/polynomial_ring.cpp:311-320,395-415 */
void orc_poly_sub(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n, uint64_t q) {
    for (size_t i = 0; i < n; i++) r[i] = orc_mod_sub(a[i], b[i], q);
}
/* src/polynomial_ring.cpp:345-358 - NOTE: no input reduction (wraps if a >= q) */
void orc_poly_negate(const uint64_t* a, uint64_t* r, size_t n, uint64_t q) {
    for (size_t i = 0; i < n; i++) r[i] = (a[i] == 0) ? 0 : (q - a[i]);
}
/* src/polynomial_ring.cpp:454-473 */
void orc_poly_scalar(const uint64_t* a, uint64_t scalar, uint64_t* r, size_t n, uint64_t q) {
    scalar %= q;
    for (size_t i = 0; i < n; i++) r[i] = (uint64_t)(((u128)a[i] * scalar) % q);
}
/* src/polynomial_ring.cpp:493-530 */
void orc_poly_pointwise(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n, uint64_t q) {
    for (size_t i = 0; i < n; i++) r[i] = (uint64_t)(((u128)a[i] * b[i]) % q);
}
/* src/polynomial_ring.cpp:421-447 - coefficient-form operands: T^-1(T(a) . T(b)) */
void orc_poly_multiply(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n, uint64_t q,
                       const uint64_t* fwd_table, const uint64_t* inv_table, uint64_t inv_n) {
    uint64_t* ta = (uint64_t*)malloc(n * 8);
    uint64_t* tb = (uint64_t*)malloc(n * 8);
    memcpy(ta, a, n * 8);
    memcpy(tb, b, n * 8);
    orc_forward_ntt(ta, n, q, fwd_table);
    orc_forward_ntt(tb, n, q, fwd_table);
    orc_poly_pointwise(ta, tb, r, n, q);
    orc_inverse_ntt(r, n, q, inv_table, inv_n);
    free(ta);
    free(tb);
}

/* ------------------------------------------------------------------------ */
/* multi-limb helpers.  Limbs little-endian.                                 */

/* src/modular_arithmetic.cpp:315-329 (operator< with equal sizes)          */
static int limbs_less(const uint64_t* a, const uint64_t* b, size_t l) {
    for (size_t i = l; i > 0; --i) {
        if (a[i - 1] < b[i - 1]) return 1;
        if (a[i - 1] > b[i - 1]) return 0;
    }
    return 0;
}
/* src/modular_arithmetic.cpp:496-507 */
static uint64_t add_limbs(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t l) {
    uint64_t carry = 0;
    for (size_t i = 0; i < l; ++i) {
        u128 s = (u128)a[i] + b[i] + carry;
        r[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
    }
    return carry;
}
/* src/modular_arithmetic.cpp:509-523 - borrow test `a < b + borrow` in u64
 * (wraps when b == 2^64-1 with borrow in: SURVEY B13, mirrored)            */
static uint64_t sub_limbs(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t l) {
    uint64_t borrow = 0;
    for (size_t i = 0; i < l; ++i) {
        uint64_t al = a[i], bl = b[i];
        uint64_t diff = al - bl - borrow;
        borrow = (al < bl + borrow) ? 1 : 0;
        r[i] = diff;
    }
    return borrow;
}
/* src/modular_arithmetic.cpp:525-545 */
static void mul_limbs(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t l) {
    memset(r, 0, 2 * l * 8);
    for (size_t i = 0; i < l; ++i) {
        uint64_t carry = 0;
        for (size_t j = 0; j < l; ++j) {
            u128 p = (u128)a[i] * b[j];
            p += r[i + j];
            p += carry;
            r[i + j] = (uint64_t)p;
            carry = (uint64_t)(p >> 64);
        }
        r[i + l] = carry;
    }
}

/* src/modular_arithmetic.cpp:361-429 - bit-serial long division as written:
 * the shifted modulus is truncated to the remainder's width and the test is
 * a STRICT greater-than.  a has `asz` limbs, result has `l` limbs.          */
static void mlimb_mod_proper(const uint64_t* a, size_t asz, const uint64_t* q, size_t l, uint64_t* out) {
    if (asz <= l) {
        /* operator< pads the shorter operand with zeros (:315-329) */
        uint64_t* pad = (uint64_t*)calloc(l, 8);
        memcpy(pad, a, asz * 8);
        int less = limbs_less(pad, q, l);
        if (less) {
            /* reference returns `a` itself (asz limbs); callers then read l limbs
             * through get_limb, which yields 0 beyond asz (modular_arithmetic.h) */
            memcpy(out, pad, l * 8);
            free(pad);
            return;
        }
        free(pad);
    }
    size_t rsz = asz < l ? l : asz;
    uint64_t* rem = (uint64_t*)calloc(rsz, 8);
    uint64_t* sh = (uint64_t*)calloc(rsz, 8);
    memcpy(rem, a, asz * 8);
    for (int bit_pos = (int)(rsz * 64) - 1; bit_pos >= 0; --bit_pos) {
        size_t limb_shift = (size_t)bit_pos / 64, bit_shift = (size_t)bit_pos % 64;
        memset(sh, 0, rsz * 8);
        for (size_t i = 0; i < l && (i + limb_shift) < rsz; ++i) {
            uint64_t ml = q[i];
            if (bit_shift == 0) {
                sh[i + limb_shift] = ml;
            } else {
                sh[i + limb_shift] |= (ml << bit_shift);
                if (i + limb_shift + 1 < rsz) sh[i + limb_shift + 1] = (ml >> (64 - bit_shift));
            }
        }
        int can = 0;
        for (int i = (int)rsz - 1; i >= 0; --i) {
            if (rem[i] > sh[i]) {
                can = 1;
                break;
            } else if (rem[i] < sh[i]) {
                break;
            }
        }
        if (can) {
            uint64_t borrow = 0;
            for (size_t i = 0; i < rsz; ++i) {
                uint64_t r = rem[i], s = sh[i];
                uint64_t d = r - s - borrow;
                borrow = (r < s + borrow) ? 1 : 0;
                rem[i] = d;
            }
        }
    }
    memcpy(out, rem, l * 8);
    free(rem);
    free(sh);
}

/* src/modular_arithmetic.cpp:347-358 (Newton q_inv), :431-468 (R, R^2 mod q),
 * :471-486 (constructor)                                                    */
int orc_mlimb_constants(const uint64_t* q, size_t l, uint64_t* consts) {
    int zero = 1;
    for (size_t i = 0; i < l; ++i)
        if (q[i]) zero = 0;
    if (zero || (q[0] & 1) == 0) return -1;
    uint64_t x = q[0];
    for (int i = 0; i < 5; ++i) x = x * (2 - q[0] * x);
    consts[0] = (~x) + 1;
    uint64_t* r = (uint64_t*)calloc(l + 1, 8);
    r[l] = 1;
    mlimb_mod_proper(r, l + 1, q, l, consts + 1);
    free(r);
    uint64_t* prod = (uint64_t*)calloc(2 * l, 8);
    mul_limbs(consts + 1, consts + 1, prod, l);
    mlimb_mod_proper(prod, 2 * l, q, l, consts + 1 + l);
    free(prod);
    return 0;
}

/* src/modular_arithmetic.cpp:558-612 (montgomery_reduce), :614-625 (montgomery_mul) */
static void mlimb_reduce(uint64_t* t /*2l limbs, clobbered*/, uint64_t* out, size_t l, const uint64_t* q,
                         uint64_t q_inv) {
    for (size_t i = 0; i < l; ++i) {
        uint64_t m = t[i] * q_inv;
        uint64_t carry = 0;
        for (size_t j = 0; j < l; ++j) {
            u128 p = (u128)m * q[j];
            p += t[i + j];
            p += carry;
            t[i + j] = (uint64_t)p;
            carry = (uint64_t)(p >> 64);
        }
        for (size_t j = l; j < 2 * l - i && carry; ++j) {
            u128 s = (u128)t[i + j] + carry;
            t[i + j] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
    }
    if (!limbs_less(t + l, q, l)) {
        sub_limbs(t + l, q, out, l);
    } else {
        memcpy(out, t + l, l * 8);
    }
}
void orc_mlimb_montmul(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, size_t l, const uint64_t* q,
                       uint64_t q_inv) {
    uint64_t* t = (uint64_t*)malloc(2 * l * 8);
    uint64_t* o = (uint64_t*)malloc(l * 8);
    for (size_t e = 0; e < count; ++e) {
        mul_limbs(a + e * l, b + e * l, t, l);
        mlimb_reduce(t, o, l, q, q_inv);
        memcpy(r + e * l, o, l * 8);
    }
    free(t);
    free(o);
}
/* src/modular_arithmetic.cpp:627-648 */
void orc_mlimb_add(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, size_t l, const uint64_t* q) {
    uint64_t* s = (uint64_t*)malloc(l * 8);
    uint64_t* o = (uint64_t*)malloc(l * 8);
    for (size_t e = 0; e < count; ++e) {
        uint64_t carry = add_limbs(a + e * l, b + e * l, s, l);
        if (carry || !limbs_less(s, q, l)) {
            sub_limbs(s, q, o, l);
            memcpy(r + e * l, o, l * 8);
        } else {
            memcpy(r + e * l, s, l * 8);
        }
    }
    free(s);
    free(o);
}
/* src/modular_arithmetic.cpp:650-671 */
void orc_mlimb_sub(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, size_t l, const uint64_t* q) {
    uint64_t* d = (uint64_t*)malloc(l * 8);
    uint64_t* o = (uint64_t*)malloc(l * 8);
    for (size_t e = 0; e < count; ++e) {
        uint64_t borrow = sub_limbs(a + e * l, b + e * l, d, l);
        if (borrow) {
            add_limbs(d, q, o, l);
            memcpy(r + e * l, o, l * 8);
        } else {
            memcpy(r + e * l, d, l * 8);
        }
    }
    free(d);
    free(o);
}

/* ------------------------------------------------------------------------ */
/* src/bootstrap_engine.cpp:57-77 - coeff_i = ((i*t)/(2N)) * (q/t) % q with
 * `i * t` evaluated in u64                                                  */
void orc_default_test_poly(const orc_boot_params* p, uint64_t* out) {
    uint64_t t = p->t > 0 ? p->t : 4;
    uint64_t delta = p->q / t;
    for (uint32_t i = 0; i < p->N; ++i) {
        uint64_t value = ((uint64_t)i * t) / (2 * p->N);
        out[i] = (value * delta) % p->q;
    }
}

static uint64_t lut_eval(int kind, uint64_t arg0, uint64_t x) {
    switch (kind) {
        case 0: return x;                           /* identity  :774-779 */
        case 1: return (arg0 - x) % arg0;           /* negation  :760-765 */
        default: return x >= arg0 ? 1ULL : 0ULL;    /* threshold :767-772 */
    }
}
/* src/bootstrap_engine.cpp:725-758 */
void orc_lookup_table(const orc_boot_params* p, int kind, uint64_t arg0, uint64_t arg1, uint64_t* out) {
    uint64_t in_mod = (kind == 2) ? arg1 : arg0;
    uint64_t out_mod = (kind == 2) ? 2 : arg0;
    uint64_t delta_out = p->q / out_mod;
    for (uint32_t i = 0; i < p->N; ++i) {
        uint64_t input_val = ((uint64_t)i * in_mod + p->N) / (2 * p->N);
        input_val %= in_mod;
        uint64_t output_val = lut_eval(kind, arg0, input_val) % out_mod;
        out[i] = (output_val * delta_out) % p->q;
    }
}

/* src/bootstrap_engine.cpp:122-145 - int32 arithmetic as written            */
void orc_rotate_polynomial(const uint64_t* poly, uint32_t N, uint64_t q, int32_t rotation, uint64_t* out) {
    int32_t two_n = 2 * (int32_t)N;
    int32_t rot = ((rotation % two_n) + two_n) % two_n;
    memset(out, 0, (size_t)N * 8);
    for (uint32_t i = 0; i < N; ++i) {
        int32_t new_idx = ((int32_t)i + rot) % two_n;
        if (new_idx < (int32_t)N) {
            out[new_idx] = poly[i];
        } else {
            out[new_idx - (int32_t)N] = (q - poly[i]) % q;
        }
    }
}

/* src/bootstrap_engine.cpp:152-185 - LOW level*base_log bits, most significant
 * digit first, centred WITHOUT carry                            */
void orc_decompose_polynomial(const uint64_t* poly, uint32_t N, uint64_t q, uint32_t base_log, uint32_t level,
                              uint64_t* out) {
    uint64_t base = 1ULL << base_log;
    uint64_t mask = base - 1;
    for (uint32_t l = 0; l < level; ++l) {
        uint32_t shift = (level - 1 - l) * base_log;
        for (uint32_t i = 0; i < N; ++i) {
            uint64_t digit = (poly[i] >> shift) & mask;
            out[(size_t)l * N + i] = (digit > base / 2) ? ((q - (base - digit)) % q) : digit;
        }
    }
}

/* src/bootstrap_engine.cpp:431-518 - per GGSW row: T(digit), and for each of
 * the row's k+1 polynomials T(row poly) (re-transformed on every call),
 * pointwise product, T^-1, coefficient-domain accumulation; mask rows first,
 * then body rows                                                */
void orc_external_product(const orc_boot_params* p, const uint64_t* glwe, const uint64_t* ggsw, uint64_t* out) {
    uint32_t N = p->N, k = p->k, L = p->level;
    uint64_t q = p->q;
    size_t comps = (size_t)k + 1;
    uint64_t* dec = (uint64_t*)malloc(comps * L * N * 8);
    uint64_t* dn = (uint64_t*)malloc((size_t)N * 8);
    uint64_t* gn = (uint64_t*)malloc((size_t)N * 8);
    uint64_t* prod = (uint64_t*)malloc((size_t)N * 8);
    memset(out, 0, comps * N * 8);
    for (size_t c = 0; c < comps; ++c)
        orc_decompose_polynomial(glwe + c * N, N, q, p->base_log, L, dec + c * L * N);
    size_t row = 0;
    for (size_t c = 0; c < comps; ++c) { /* c < k: mask polynomials; c == k: body */
        for (uint32_t l = 0; l < L; ++l, ++row) {
            const uint64_t* grow = ggsw + row * comps * N;
            memcpy(dn, dec + (c * L + l) * N, (size_t)N * 8);
            orc_forward_ntt(dn, N, q, p->fwd_table);
            for (size_t j = 0; j < comps; ++j) { /* j < k: row mask j -> result mask j; j == k: body */
                memcpy(gn, grow + j * N, (size_t)N * 8);
                orc_forward_ntt(gn, N, q, p->fwd_table);
                orc_poly_pointwise(dn, gn, prod, N, q);
                orc_inverse_ntt(prod, N, q, p->inv_table, p->inv_n);
                orc_poly_add(out + j * N, prod, out + j * N, N, q);
            }
        }
    }
    free(dec);
    free(dn);
    free(gn);
    free(prod);
}

/* src/bootstrap_engine.cpp:520-540 - ct0 + ggsw [*] (ct1 - ct0)             */
void orc_cmux(const orc_boot_params* p, const uint64_t* ggsw, const uint64_t* ct0, const uint64_t* ct1,
              uint64_t* out) {
    size_t words = ((size_t)p->k + 1) * p->N;
    uint64_t* diff = (uint64_t*)malloc(words * 8);
    orc_poly_sub(ct1, ct0, diff, words, p->q);
    orc_external_product(p, diff, ggsw, out);
    orc_poly_add(out, ct0, out, words, p->q);
    free(diff);
}

/* src/bootstrap_engine.cpp:547-577 - rotation amounts in wrapping u64, then
 * int32 cast; iterations whose rotation is 0 are skipped        */
void orc_blind_rotate(const orc_boot_params* p, uint64_t* acc, const uint64_t* lwe, const uint64_t* bsk) {
    uint32_t N = p->N, k = p->k;
    uint64_t q = p->q;
    size_t comps = (size_t)k + 1, words = comps * N;
    size_t ggsw_words = comps * p->level * comps * N;
    uint64_t* rot = (uint64_t*)malloc(words * 8);
    uint64_t* next = (uint64_t*)malloc(words * 8);
    uint64_t b = lwe[p->n];
    int32_t b_rotation = -(int32_t)((b * 2 * N + q / 2) / q);
    for (size_t c = 0; c < comps; ++c) orc_rotate_polynomial(acc + c * N, N, q, b_rotation, rot + c * N);
    memcpy(acc, rot, words * 8);
    for (size_t i = 0; i < p->n; ++i) {
        int32_t a_rotation = (int32_t)((lwe[i] * 2 * N + q / 2) / q);
        if (a_rotation == 0) continue;
        for (size_t c = 0; c < comps; ++c) orc_rotate_polynomial(acc + c * N, N, q, a_rotation, rot + c * N);
        orc_cmux(p, bsk + i * ggsw_words, acc, rot, next);
        memcpy(acc, next, words * 8);
    }
    free(rot);
    free(next);
}

/* src/bootstrap_engine.cpp:594-624 */
void orc_sample_extract(const uint64_t* glwe, uint32_t k, uint32_t N, uint64_t q, uint64_t* out) {
    for (uint32_t i = 0; i < k; ++i) {
        const uint64_t* mask = glwe + (size_t)i * N;
        out[(size_t)i * N] = mask[0];
        for (uint32_t j = 1; j < N; ++j) out[(size_t)i * N + j] = (q - mask[N - j]) % q;
    }
    out[(size_t)k * N] = glwe[(size_t)k * N];
}

/* src/bootstrap_engine.cpp:626-669 - low-bit digits, zero digits skipped,
 * `digit * ksk` wraps mod 2^64 BEFORE `% q`                     */
void orc_key_switch(const uint64_t* lwe, size_t dim_in, uint64_t q, const uint64_t* ksk, size_t n_out,
                    uint32_t base_log, uint32_t level, uint64_t* out) {
    uint64_t base = 1ULL << base_log, mask = base - 1;
    memset(out, 0, n_out * 8);
    uint64_t res_b = lwe[dim_in];
    size_t idx = 0;
    for (size_t i = 0; i < dim_in; ++i) {
        uint64_t coeff = lwe[i];
        for (uint32_t l = 0; l < level; ++l) {
            uint32_t shift = (level - 1 - l) * base_log;
            uint64_t digit = (coeff >> shift) & mask;
            if (digit == 0) {
                idx++;
                continue;
            }
            const uint64_t* e = ksk + (idx++) * (n_out + 1);
            for (size_t j = 0; j < n_out; ++j) out[j] = (out[j] + q - (digit * e[j]) % q) % q;
            res_b = (res_b + q - (digit * e[n_out]) % q) % q;
        }
    }
    out[n_out] = res_b;
}

/* src/bootstrap_engine.cpp:684-711 */
void orc_bootstrap(const orc_boot_params* p, const uint64_t* lwe, const uint64_t* bsk, const uint64_t* test_poly,
                   const uint64_t* ksk, size_t n_out, uint32_t ksk_base_log, uint32_t ksk_level, uint64_t* out) {
    size_t comps = (size_t)p->k + 1, words = comps * p->N;
    uint64_t* acc = (uint64_t*)calloc(words, 8);
    memcpy(acc + (size_t)p->k * p->N, test_poly, (size_t)p->N * 8);
    orc_blind_rotate(p, acc, lwe, bsk);
    size_t dim = (size_t)p->k * p->N;
    if (ksk) {
        uint64_t* ext = (uint64_t*)malloc((dim + 1) * 8);
        orc_sample_extract(acc, p->k, p->N, p->q, ext);
        orc_key_switch(ext, dim, p->q, ksk, n_out, ksk_base_log, ksk_level, out);
        free(ext);
    } else {
        orc_sample_extract(acc, p->k, p->N, p->q, out);
    }
    free(acc);
}

/* ------------------------------------------------------------------------ */
/* src/encryption.cpp:1327-1364 - m == 1 returns the ballot untouched (raw
 * words); otherwise left fold with PolynomialRing::add_inplace              */
int orc_tally_linear(const uint64_t* cts, size_t m, size_t n, uint64_t q, uint64_t* out) {
    if (m == 0) return -1;
    memcpy(out, cts, 2 * n * 8);
    for (size_t i = 1; i < m; ++i) orc_poly_add(out, cts + i * 2 * n, out, 2 * n, q);
    return 0;
}
/* src/encryption.cpp:1366-1458 - pairwise levels, odd element carried       */
int orc_tally_tree(const uint64_t* cts, size_t m, size_t n, uint64_t q, uint64_t* out) {
    if (m == 0) return -1;
    size_t w = 2 * n;
    uint64_t* cur = (uint64_t*)malloc(m * w * 8);
    memcpy(cur, cts, m * w * 8);
    size_t count = m;
    while (count > 1) {
        size_t pairs = count / 2;
        for (size_t i = 0; i < pairs; ++i) orc_poly_add(cur + 2 * i * w, cur + (2 * i + 1) * w, cur + i * w, w, q);
        if (count % 2 == 1) memmove(cur + pairs * w, cur + (count - 1) * w, w * 8);
        count = pairs + (count % 2);
    }
    memcpy(out, cur, w * 8);
    free(cur);
    return 0;
}
/* src/encryption.cpp:737-798 */
void orc_tensor_multiply(const uint64_t* ct1, const uint64_t* ct2, uint64_t* out, size_t n, uint64_t q,
                         const uint64_t* fwd_table, const uint64_t* inv_table, uint64_t inv_n) {
    uint64_t* w = (uint64_t*)malloc(6 * n * 8);
    uint64_t *a0 = w, *a1 = w + n, *b0 = w + 2 * n, *b1 = w + 3 * n, *x = w + 4 * n, *y = w + 5 * n;
    memcpy(a0, ct1, 2 * n * 8);
    memcpy(b0, ct2, 2 * n * 8);
    orc_forward_ntt(a0, n, q, fwd_table);
    orc_forward_ntt(a1, n, q, fwd_table);
    orc_forward_ntt(b0, n, q, fwd_table);
    orc_forward_ntt(b1, n, q, fwd_table);
    orc_poly_pointwise(a0, b0, out, n, q);
    orc_poly_pointwise(a0, b1, x, n, q);
    orc_poly_pointwise(a1, b0, y, n, q);
    orc_poly_add(x, y, out + n, n, q);
    orc_poly_pointwise(a1, b1, out + 2 * n, n, q);
    orc_inverse_ntt(out, n, q, inv_table, inv_n);
    orc_inverse_ntt(out + n, n, q, inv_table, inv_n);
    orc_inverse_ntt(out + 2 * n, n, q, inv_table, inv_n);
    free(w);
}
/* src/encryption.cpp:904-993 (relinearize).  ct = [3][n]; keys = [key_count][2][n], (a, b) per pair; out = [2][n].
 * Level l takes the raw bits (c2 >> l*base_log) & (base-1) (:948-955), transforms the digit polynomial and both key
 * polynomials (:957-963), and adds T^-1 of the pointwise products to c0 (with b) and c1 (with a) (:965-978). */
void orc_relinearize(const uint64_t* ct, const uint64_t* keys, uint32_t key_count, uint32_t key_base_log, uint32_t key_level,
                     uint64_t* out, size_t n, uint64_t q, const uint64_t* fwd_table, const uint64_t* inv_table, uint64_t inv_n) {
    const uint32_t base_log = key_base_log > 0 ? key_base_log : 4;                       /* :936 */
    const uint64_t base = 1ULL << base_log;
    const uint32_t num_levels = key_level > 0 ? key_level : (64 + base_log - 1) / base_log; /* :938-939 */
    uint64_t* w = (uint64_t*)malloc(4 * n * 8);
    uint64_t *dig = w, *ka = w + n, *kb = w + 2 * n, *prod = w + 3 * n;
    const uint64_t* c2 = ct + 2 * n;
    memcpy(out, ct, 2 * n * 8); /* clones of c0 and c1 (:942-943) */
    for (uint32_t level = 0; level < num_levels && level < key_count; ++level) {
        const uint64_t shift = (uint64_t)level * base_log, mask = base - 1;
        for (size_t i = 0; i < n; ++i) dig[i] = (c2[i] >> shift) & mask;
        memcpy(ka, keys + ((size_t)level * 2) * n, n * 8);
        memcpy(kb, keys + ((size_t)level * 2 + 1) * n, n * 8);
        orc_forward_ntt(dig, n, q, fwd_table);
        orc_forward_ntt(ka, n, q, fwd_table);
        orc_forward_ntt(kb, n, q, fwd_table);
        orc_poly_pointwise(dig, kb, prod, n, q);
        orc_inverse_ntt(prod, n, q, inv_table, inv_n);
        orc_poly_add(out, prod, out, n, q);
        orc_poly_pointwise(dig, ka, prod, n, q);
        orc_inverse_ntt(prod, n, q, inv_table, inv_n);
        orc_poly_add(out + n, prod, out + n, n, q);
    }
    free(w);
}

/* ---- wire formats: src/key_serializer.cpp ------------------------------------------------------------ */
/* compute_crc32 (:34-40) over the table AS SHIPPED (:21-32): the initialiser list stops after 60 entries, the other
 * 196 are zero.  Entry i < 60 is the reflected IEEE 802.3 value (polynomial 0xEDB88320). */
static uint32_t orc_crc_entry(uint32_t i) {
    if (i >= 60) return 0;
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    return c;
}
uint32_t orc_crc32(const uint8_t* data, size_t len) {
    uint32_t crc = 0xFFFFFFFFu;
    for (size_t i = 0; i < len; ++i) crc = orc_crc_entry((crc ^ data[i]) & 0xFF) ^ (crc >> 8);
    return crc ^ 0xFFFFFFFFu;
}
static uint64_t orc_le(const uint8_t* p, int bytes) {
    uint64_t v = 0;
    for (int i = 0; i < bytes; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}
/* BallotSerializer::deserialize_ballot (:776-846) on one record of `len` bytes, then the shape test of the bulk
 * ingest path.  Returns 0 ok, 1 "Input too small" (:781: len < sizeof(SerializationHeader) = 64), 2 "Invalid magic
 * bytes for ballot" (:800), 3 "Checksum verification failed" (:808; payload read into a zero-filled buffer of
 * data_size bytes), 4 not `choices` ciphertexts of degree n over q.  out = [choices][2][n], zeroed unless 0. */
int orc_ballot_parse(const uint8_t* rec, size_t len, uint32_t choices, size_t n, uint64_t q, uint64_t* out, uint64_t* timestamp) {
    memset(out, 0, (size_t)choices * 2 * n * 8);
    *timestamp = 0;
    if (len < 64) return 1;
    if ((uint32_t)orc_le(rec, 4) != 0x46484556u) return 2;
    const uint32_t data_size = (uint32_t)orc_le(rec + 32, 4);
    const uint32_t checksum = (uint32_t)orc_le(rec + 45, 4);
    uint8_t* data = (uint8_t*)calloc(data_size ? data_size : 1, 1);
    const size_t avail = len - 49;
    memcpy(data, rec + 49, data_size < avail ? data_size : avail);
    const uint32_t crc = orc_crc32(data, data_size);
    int rc = 0;
    if (crc != checksum) {
        rc = 3;
    } else if (data_size > avail || data_size != 12 + (uint64_t)choices * (12 + 16 * n) || (uint32_t)orc_le(data + 8, 4) != choices) {
        rc = 4;
    } else {
        for (uint32_t c = 0; c < choices && rc == 0; ++c) {
            const uint8_t* ch = data + 12 + (size_t)c * (12 + 16 * n);
            if (orc_le(ch, 4) != n || orc_le(ch + 4, 8) != q) rc = 4;
        }
        if (rc == 0) {
            *timestamp = orc_le(data, 8);
            for (uint32_t c = 0; c < choices; ++c) {
                const uint8_t* ch = data + 12 + (size_t)c * (12 + 16 * n) + 12;
                for (size_t j = 0; j < 2 * n; ++j) out[(size_t)c * 2 * n + j] = orc_le(ch + 8 * j, 8);
            }
        }
    }
    free(data);
    return rc;
}
