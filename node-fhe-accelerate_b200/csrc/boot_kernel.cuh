// The persistent blind-rotation / CMux / external-product kernel (device-only part; the phase
// bodies live in boot_core.cuh).  One thread block owns one ciphertext: its accumulator stays in
// shared memory for all n CMux steps, the pre-transformed bootstrapping key streams in from
// HBM/L2 (it is shared by every block of the batch), and nothing but the final GLWE goes back.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "boot_core.cuh"
#include "runtime.hpp"

namespace fheb {

enum { BOOT_BLIND = 0, BOOT_CMUX = 1, BOOT_EXT = 2 };

struct BootLaunch {
    const Tw* bsk;           // BLIND: all n GGSWs; CMUX/EXT: the selected GGSW
    const uint64_t* in0;     // BLIND: lwe [batch][n+1]; CMUX: ct0 [batch][KP1][N]; EXT: glwe
    const uint64_t* in1;     // BLIND: test polynomial [N]; CMUX: ct1
    uint64_t* out;           // [batch][KP1][N]
    size_t batch;
    uint32_t n;              // LWE dimension (BLIND)
    uint32_t levels, base_log;
    int mode;
    int acc_global;          // 1: the accumulator lives in the output buffer (global memory, L2) instead of shared memory
    const int* raw_flag;     // BLIND, optional: device flag "the test polynomial has words >= q"; selects between the lean and the general kernel
    const Tw* twf;
    const Tw* twi;
    Tw ninv;
    ModQ m;
};

template <int L>
struct BootGeometry {
    // threads that work on ONE ciphertext; a block runs several ciphertexts in lockstep (groups of
    // TPC threads) so that its 16 warps share one instruction stream - the unrolled passes of a step
    // are ~80 KB of code, far beyond the instruction caches when four blocks run out of phase
    static constexpr int TPC = (L <= 6) ? 32 : (L <= 8) ? 64 : (L <= 10) ? 128 : (L == 11) ? 256 : 512;
    static constexpr int MAX_THREADS = 512;
    static constexpr int MAX_GROUPS = MAX_THREADS / TPC;
    // N = 1024 (the tfhe-128-fast shape) and N = 512: one ciphertext per 128-thread block and five blocks per SM
    // (96 registers) measured 7 % (N = 1024) and 21 % (N = 512) faster than four ciphertexts in lockstep in one
    // 512-thread block - the blocks drift out of phase, so one block's FP64 bursts overlap another's loads, digit
    // extraction and barriers.  N = 2048 measured the same either way and keeps the lockstep pair.
#if !defined(FHEB_EXP_BOOT_LB_BLOCKS)
#define FHEB_EXP_BOOT_LB_BLOCKS 5
#endif
    static constexpr int LB_THREADS = (L == 9 || L == 10) ? TPC : MAX_THREADS;
    static constexpr int LB_BLOCKS = (L == 9 || L == 10) ? FHEB_EXP_BOOT_LB_BLOCKS : 1;
};

// shared-memory bytes per ciphertext: acc [KP1][N] | work [rows][N] | (CMUX/EXT) diff [KP1][N] | (BLIND) rotations [n] (u32)
inline size_t boot_smem_bytes(int mode, uint32_t N, uint32_t kp1, uint32_t levels, uint32_t n, bool acc_global = false) {
    size_t words = (acc_global ? 0 : (size_t)kp1 * N) + (size_t)kp1 * levels * N;
    if (mode != BOOT_BLIND && !acc_global) words += (size_t)kp1 * N;
    size_t bytes = words * 8;
    if (mode == BOOT_BLIND) bytes += (size_t)n * 4;
    return (bytes + 127) & ~(size_t)127;  // every ciphertext's rows start 128-byte aligned (SlotRef, ntt_core.cuh)
}

// `active` is uniform per ciphertext group; idle groups still take part in the block barriers
template <int L, bool DP, int KP1, bool RAWOK, int PH = 0>
__device__ __forceinline__ void boot_run_step(bool active, uint32_t tid, uint32_t nthreads, const BootStep& s, const BootLaunch& a) {
    if constexpr (PH < boot_phases<L>()) {
        if (active) boot_phase<L, DP, KP1, PH, RAWOK>(tid, nthreads, s, a.twf, a.twi, a.ninv, a.m);
        __syncthreads();
        boot_run_step<L, DP, KP1, RAWOK, PH + 1>(active, tid, nthreads, s, a);
    }
}

// RAWOK = false: the lean blind-rotation kernel for canonical test polynomials (no unreduced-word code: 36 % fewer
// instructions, fewer instruction-cache misses).  Both variants are launched on a blind rotation that carries a
// raw_flag; the one the flag does not select returns at once.
template <int L, bool DP, int KP1, bool RAWOK = true>
__global__ void __launch_bounds__(BootGeometry<L>::LB_THREADS, BootGeometry<L>::LB_BLOCKS) boot_kernel(const BootLaunch a, const uint32_t ct_bytes) {
    extern __shared__ __align__(128) uint64_t smem[];
    if (a.raw_flag != nullptr && (*a.raw_flag != 0) != RAWOK) return;  // block-uniform
    constexpr uint32_t N = 1u << L;
    constexpr uint32_t TPC = BootGeometry<L>::TPC;
    constexpr uint32_t GW = (uint32_t)KP1 * N;  // words per GLWE
    const uint32_t group = threadIdx.x / TPC;    // which of the block's ciphertexts this thread works on
    const uint32_t groups = blockDim.x / TPC;
    const uint32_t tid = threadIdx.x % TPC;
    const uint32_t rows = (uint32_t)KP1 * a.levels;
    // large shapes (e.g. tfhe-256-secure: N=4096, three levels): the working rows alone nearly fill the SM, so
    // the accumulator is kept in the output GLWE in global memory (L2 resident, touched twice per step)
    const bool acc_global = (L >= BOOT_GLOBAL_MIN_L) && a.acc_global;
    uint64_t* sm_acc = smem + (size_t)group * (ct_bytes / 8);
    uint64_t* work = sm_acc + (acc_global ? 0 : GW);
    uint64_t* diff = work + (size_t)rows * N;                   // CMUX / EXT only
    uint32_t* rots = reinterpret_cast<uint32_t*>(work + (size_t)rows * N);  // BLIND only
    const size_t ggsw_words = (size_t)rows * KP1 * N;

    for (size_t ct0 = (size_t)blockIdx.x * groups; ct0 < a.batch; ct0 += (size_t)gridDim.x * groups) {
        const size_t ct = ct0 + group;
        const bool valid = ct < a.batch;
        uint64_t* gout = a.out + (valid ? ct : 0) * GW;
        // accumulator: shared memory, or (acc_global) the output GLWE itself / the caller's ct0 for a single CMux
        uint64_t* acc = !acc_global ? sm_acc
                        : (a.mode == BOOT_CMUX ? const_cast<uint64_t*>(a.in0 + (valid ? ct : 0) * GW) : gout);
        BootStep s;
        s.acc = acc;
        s.work = work;
        s.levels = a.levels;
        s.rows = rows;
        s.base_log = a.base_log;
        s.rot = 0;
        s.ggsw = a.bsk;
        s.diff = nullptr;
        s.diff_sub = nullptr;
        s.add_acc = 1;
        s.gout = nullptr;
        s.maybe_raw = (RAWOK && (a.mode == BOOT_BLIND || acc_global)) ? 1u : 0u;  // CMUX loads ct0 reduced into shared memory
        uint32_t nsteps = 1;
        if (a.mode == BOOT_BLIND) {
            if (valid) {
                const uint64_t* lwe = a.in0 + ct * ((size_t)a.n + 1);
                for (uint32_t i = tid; i < a.n; i += TPC) rots[i] = lwe_step_rotation(lwe[i], N, a.m.q);
                // acc = X^(-round(b * 2N / q)) * (0, .., 0, test_poly): multiply_glwe_by_monomial, :558-559
                const uint32_t rb = lwe_rotation(lwe[a.n], true, N, a.m.q);
                for (uint32_t i = tid; i < GW; i += TPC) {
                    const uint32_t c = i >> L, j = i & (N - 1);
                    acc[i] = (c == (uint32_t)KP1 - 1) ? rotated_at(a.in1, j, rb, N, a.m) : 0;
                }
            }
            nsteps = a.n;
        } else if (valid) {
            const uint64_t* g0 = a.in0 + ct * GW;
            if (acc_global) {  // no staging: the first pass reads the operands from global memory
                s.diff = (a.mode == BOOT_CMUX) ? a.in1 + ct * GW : g0;
                s.diff_sub = (a.mode == BOOT_CMUX) ? g0 : nullptr;
            } else {
                if (a.mode == BOOT_CMUX) {
                    const uint64_t* g1 = a.in1 + ct * GW;
                    for (uint32_t i = tid; i < GW; i += TPC) {
                        const uint64_t c0 = g0[i];
                        const uint64_t c0r = canon_any(c0, a.m);
                        acc[i] = c0r;
                        diff[i] = submod_canon(canon_any(g1[i], a.m), c0r, a.m.q);
                    }
                } else {
                    for (uint32_t i = tid; i < GW; i += TPC) diff[i] = g0[i];
                }
                s.diff = diff;
            }
            s.add_acc = (a.mode == BOOT_CMUX) ? 1 : 0;
            s.gout = gout;
        }
        __syncthreads();
        // one call site for the step body: the unrolled passes are large, keep a single copy of them
        for (uint32_t i = 0; i < nsteps; ++i) {
            bool active = valid;
            if (a.mode == BOOT_BLIND) {
                const uint32_t rot = valid ? rots[i] : 0u;
                // :564-566 - a zero RAW rotation skips the CMux (uniform per ciphertext).  rot == 2N marks a raw rotation
                // that is a non-zero multiple of 2N: the reference runs a CMux with diff = 0, whose only effect is to
                // canonicalise the accumulator - the identity once every word is canonical (maybe_raw == 0)
                active = (rot & (2u * N - 1u)) != 0 || (rot != 0 && s.maybe_raw != 0);
                s.rot = rot & (2u * N - 1u);
                s.ggsw = reinterpret_cast<const Tw*>(reinterpret_cast<const char*>(a.bsk) + (size_t)i * ggsw_words * (DP ? 8 : 16));
            }
            boot_run_step<L, DP, KP1, RAWOK>(active, tid, TPC, s, a);
            if (active) s.maybe_raw = 0;  // every accumulator word is now the output of a modular addition
        }
        if (a.mode == BOOT_BLIND && valid && !acc_global) {
            for (uint32_t i = tid; i < GW; i += TPC) gout[i] = acc[i];
        }
        __syncthreads();  // the next ciphertexts overwrite acc / rots / diff
    }
}

template <int L, bool DP, int KP1>
int boot_launch_one(const BootLaunch& a_in, cudaStream_t stream) {
    using G = BootGeometry<L>;
    BootLaunch a = a_in;
    size_t ct_bytes = boot_smem_bytes(a.mode, 1u << L, KP1, a.levels, a.n);
    const size_t cap = (size_t)ctx().prop.sharedMemPerBlockOptin;
    auto k = boot_kernel<L, DP, KP1>;
    a.acc_global = 0;
    if (ct_bytes > cap && L >= BOOT_GLOBAL_MIN_L) {  // keep the accumulator in the output buffer instead
        a.acc_global = 1;
        ct_bytes = boot_smem_bytes(a.mode, 1u << L, KP1, a.levels, a.n, true);
    }
    if (ct_bytes > cap)
        return set_error(FHEB_ERR_INVALID_PARAMETERS,
                         "bootstrap working set (%zu bytes) exceeds the shared memory of one SM; reduce N, k or the level count",
                         ct_bytes);
    size_t groups = cap / ct_bytes;  // ciphertexts per block: as many as fit, at most 512 threads' worth
    if (groups > (size_t)G::MAX_GROUPS) groups = G::MAX_GROUPS;
    if (groups > (size_t)(G::LB_THREADS / G::TPC)) groups = G::LB_THREADS / G::TPC;
    if (groups > a.batch) groups = a.batch;
    if (const char* e = getenv("FHEB_BOOT_GROUPS")) {  // experiment knob: ciphertexts per block
        const size_t g = (size_t)atoi(e);
        if (g >= 1 && g < groups) groups = g;
    }
    const size_t smem = groups * ct_bytes;
    const int threads = (int)groups * G::TPC;
    if (smem > 48 * 1024) FHEB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    FHEB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, threads, smem));
    if (bps < 1) return set_error(FHEB_ERR_NATIVE, "bootstrap kernel does not fit on an SM");
    const size_t resident = (size_t)ctx().sm_count * (size_t)bps;
    const size_t blocks = (a.batch + groups - 1) / groups;
    const unsigned grid = (unsigned)(blocks < resident ? blocks : resident);
    // the lean variant exists for k = 1 blind rotations kept in shared memory (the hot configuration)
    const bool split = (KP1 == 2) && a.mode == BOOT_BLIND && !a.acc_global && a.raw_flag != nullptr;
    if (!split) a.raw_flag = nullptr;
    k<<<grid, threads, smem, stream>>>(a, (uint32_t)ct_bytes);
    FHEB_CHECK_LAUNCH();
    count_launch();
    if constexpr (KP1 == 2) {
        if (split) {
            auto lean = boot_kernel<L, DP, KP1, false>;
            if (smem > 48 * 1024) FHEB_CUDA(cudaFuncSetAttribute(lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            lean<<<grid, threads, smem, stream>>>(a, (uint32_t)ct_bytes);
            FHEB_CHECK_LAUNCH();
            count_launch();
        }
    }
    return FHEB_OK;
}

template <int KP1>
int boot_launch_kp1(uint32_t logn, bool dp, const BootLaunch& a, cudaStream_t stream) {
#define FHEB_BOOT_CASE(L_) \
    case L_:               \
        return dp ? boot_launch_one<L_, true, KP1>(a, stream) : boot_launch_one<L_, false, KP1>(a, stream);
    switch (logn) {
        FHEB_BOOT_CASE(5)
        FHEB_BOOT_CASE(6)
        FHEB_BOOT_CASE(7)
        FHEB_BOOT_CASE(8)
        FHEB_BOOT_CASE(9)
        FHEB_BOOT_CASE(10)
        FHEB_BOOT_CASE(11)
        FHEB_BOOT_CASE(12)
    }
#undef FHEB_BOOT_CASE
    return set_error(FHEB_ERR_INVALID_PARAMETERS, "bootstrap kernels support polynomial degrees 32..4096 (got 2^%u)", logn);
}

// one translation unit per GLWE dimension (parallel compilation)
int boot_launch_k1(uint32_t logn, bool dp, const BootLaunch& a, cudaStream_t stream);
int boot_launch_k2(uint32_t logn, bool dp, const BootLaunch& a, cudaStream_t stream);
int boot_launch_k3(uint32_t logn, bool dp, const BootLaunch& a, cudaStream_t stream);

}  // namespace fheb
