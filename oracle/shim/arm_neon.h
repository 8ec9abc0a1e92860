/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). Not part of the product.
 *
 * x86 stand-in for <arm_neon.h>, found via -I oracle/shim when the reference's
 * own portable C++ is compiled into oracle/_ref/libref_oracle.so.
 *
 * Why it exists: the reference's modular_arithmetic.h includes <arm_neon.h>
 * unconditionally (reference cpp/include/modular_arithmetic.h:5) and the
 * scalar code paths still declare three NEON values that are never used
 * arithmetically (reference cpp/src/modular_arithmetic.cpp:189,194-195,
 * 707-708).  The reference also leans on libc++ transitive includes
 * (std::unique_ptr at modular_arithmetic.h:80, std::max at
 * modular_arithmetic.cpp:316), so the shim pulls those headers in as well.
 */
#pragma once
#include <algorithm>
#include <cstdint>
#include <memory>
#include <stdexcept>

struct uint64x2_t {
    uint64_t lane[2];
};

static inline uint64x2_t vdupq_n_u64(uint64_t v) {
    uint64x2_t r;
    r.lane[0] = v;
    r.lane[1] = v;
    return r;
}

static inline uint64x2_t vld1q_u64(const uint64_t* p) {
    uint64x2_t r;
    r.lane[0] = p[0];
    r.lane[1] = p[1];
    return r;
}
