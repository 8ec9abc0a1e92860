// __global__ kernels of the transform path (device-only part; the pass bodies live in
// ntt_core.cuh so that the host emulation can run them too).
#pragma once
#include <cuda_runtime.h>

#include "ntt_core.cuh"

// Software-pipelined first-pass loads (PIPE, ntt_core.cuh) pay for the 8-byte kernels, whose one block per SM has nothing
// else to hide the load latency behind.  The 32-bit kernel with 4-byte slots is the opposite case: without the pipeline's
// 32 staging registers it needs 64 registers instead of 100, so THREE 256-thread blocks share an SM (64 KB of work buffer
// each) and cover each other's loads - forward N = 16384, q = 132120577: 0.0916 -> 0.0815 ms (3.30 TB/s, 0.50 of the HBM peak).
// The inverse fits 64 registers either way; without the pipeline it measured 0.0890-0.0894 ms against 0.0901-0.0910 with it.
#if defined(FHEB_EXP_U32_PIPE)
#define FHEB_PIPE_MODE(DP) true
#else
#define FHEB_PIPE_MODE(DP) ((DP) != MODE_U32)
#endif
#define FHEB_PIPE_MODE_INV(DP) ((DP) != MODE_U32)

namespace fheb {

// Hint the next polynomials of this block into L2 while the current ones are processed
// (cp.async.bulk.prefetch.L2: one bulk request from one thread, no register or SM traffic).
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (bytes & 15u) == 0 && bytes != 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- bulk-async (TMA) loads of the next polynomials into a landing buffer ------------------------------------
// One thread asks the copy engine for the whole PPC x N x 8 bytes of the block's NEXT group (cp.async.bulk, 1-D: the rows
// of a batch are contiguous, no tensor map needed) while the block transforms the current one; completion is counted on an
// mbarrier in shared memory.  The first pass then reads its words from shared memory instead of waiting on global loads.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the initialised barrier visible to the copy engine
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

template <int L, int DP, int PASS, bool SUB = false, int PK = L>
__device__ __forceinline__ void fwd_middle(uint32_t tid, uint32_t nthreads, uint32_t polys, uint64_t* smem,
                                           const Tw* __restrict__ tw, const ModQ& m) {
    if constexpr (PASS < Plan<PK>::P - 1) {
        fwd_pass<L, DP, PASS, IO_SMEM, IO_SMEM, true, false, 0, SUB, false, PK>(tid, nthreads, polys, nullptr, nullptr, smem, tw, m);
        __syncthreads();
        fwd_middle<L, DP, PASS + 1, SUB, PK>(tid, nthreads, polys, smem, tw, m);
    }
}

template <int L, int DP, int PASS, bool SUB = false, int PK = L>
__device__ __forceinline__ void inv_middle(uint32_t tid, uint32_t nthreads, uint32_t polys, uint64_t* smem,
                                           const Tw* __restrict__ tw, const Tw ninv, const ModQ& m) {
    if constexpr (PASS > 0) {
        inv_pass<L, DP, PASS, IO_SMEM, IO_SMEM, true, 0, 1, SUB, false, PK>(tid, nthreads, polys, nullptr, nullptr, smem, tw, ninv, m);
        __syncthreads();
        inv_middle<L, DP, PASS - 1, SUB, PK>(tid, nthreads, polys, smem, tw, ninv, m);
    }
}

// Forward transform of `batch` polynomials, [batch][N] -> [batch][N] (in may equal out).
// SCALE = true gives fast_ntt_inverse semantics when `tw` is the inverse table.
// TMA = true (multi-pass plans, degrees up to 4096): shared memory = work buffer | landing buffer | mbarrier; the
// next group's words are bulk-copied into the landing buffer while this group's later passes run.
// PK: plan key (ntt_core.cuh): the pass split this kernel runs, with the matching table in `tw`.
template <int L, int DP, int THREADS, int PPC, bool SCALE, bool TMA = false, int PK = L>
__global__ void __launch_bounds__(THREADS) ntt_forward_kernel(const uint64_t* in, uint64_t* out, size_t batch,
                                                              const Tw* __restrict__ tw, const Tw ninv, const ModQ m) {
    extern __shared__ __align__(128) uint64_t smem[];
    constexpr int P = Plan<PK>::P;
    static_assert(PK == L || !TMA, "alternative plans run the plain-load form");
    constexpr size_t N = (size_t)1 << L;
    const uint32_t tid = threadIdx.x;
    const size_t groups = (batch + PPC - 1) / PPC;
    if constexpr (TMA && P > 1) {
        // work buffer (PPC x N slots of the mode's width) | landing buffer (PPC x N raw 8-byte words) | mbarrier
        // PART (N = 16384, 8-byte slots, one polynomial per block): the 128 KB work buffer leaves 99 KB, so only the first
        // LAND_PART_WORDS words (12 of the 16 rows the first pass reads) are landed; the last 4 rows are prefetched into L2
        // and read from caller memory
        constexpr bool PART = (L == 14) && smem_slot_bytes<DP>() == 8;
        static_assert(!PART || PPC == 1, "partial landing: one polynomial per block");
        constexpr size_t LWORDS = PART ? (size_t)LAND_PART_WORDS : (size_t)PPC * N;
        constexpr int IN0 = PART ? IO_LANDING_PART : IO_LANDING;
        uint64_t* landing = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(smem) + (size_t)PPC * N * smem_slot_bytes<DP>());
        uint64_t* bar = landing + LWORDS;
        auto group_bytes = [&](size_t g) {
            const size_t q0 = g * PPC;
            return PART ? (uint32_t)(LWORDS * 8) : (uint32_t)(((batch - q0) < (size_t)PPC ? (batch - q0) : (size_t)PPC) * N * 8);
        };
        auto request = [&](size_t g) {  // one thread: bulk copy of group g's (first) words, L2 prefetch of the rest
            tma_load_1d(landing, in + g * PPC * N, group_bytes(g), bar);
            if constexpr (PART) prefetch_l2_bulk(in + g * N + LWORDS, (uint32_t)((N - LWORDS) * 8));
        };
        if (tid == 0) {
            mbar_init(bar, 1);
            if ((size_t)blockIdx.x < groups) request(blockIdx.x);
        }
        __syncthreads();
        uint32_t parity = 0;
        for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
            const size_t p0 = g * PPC;
            const uint32_t polys = (uint32_t)((batch - p0) < (size_t)PPC ? (batch - p0) : (size_t)PPC);
            uint64_t* gout = out + p0 * N;
            mbar_wait(bar, parity);  // this group's words have landed
            parity ^= 1u;
            fwd_pass<L, DP, 0, IN0, IO_SMEM>(tid, THREADS, polys, landing, gout, smem, tw, m, Tw{0, 0}, GlobalMap{0, 0}, in + p0 * N);
            __syncthreads();  // every thread is done with the landing buffer: the next group may land
            if (tid == 0 && g + gridDim.x < groups) request(g + gridDim.x);
            fwd_middle<L, DP, 1>(tid, THREADS, polys, smem, tw, m);
            fwd_pass<L, DP, P - 1, IO_SMEM, IO_GLOBAL, true, SCALE>(tid, THREADS, polys, nullptr, gout, smem, tw, m, ninv);
            __syncthreads();
        }
        return;
    }
    for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const size_t p0 = g * PPC;
        const uint32_t polys = (uint32_t)((batch - p0) < (size_t)PPC ? (batch - p0) : (size_t)PPC);
        const uint64_t* gin = in + p0 * N;
        uint64_t* gout = out + p0 * N;
        if (tid == 0 && g + gridDim.x < groups) {
            const size_t q0 = (g + gridDim.x) * PPC;
            const size_t np = (batch - q0) < (size_t)PPC ? (batch - q0) : (size_t)PPC;
            prefetch_l2_bulk(in + q0 * N, (uint32_t)(np * N * 8));
        }
        if constexpr (P == 1) {
            fwd_pass<L, DP, 0, IO_GLOBAL, IO_GLOBAL, true, SCALE>(tid, THREADS, polys, gin, gout, smem, tw, m, ninv);
        } else {
            fwd_pass<L, DP, 0, IO_GLOBAL, IO_SMEM, true, false, 0, false, (L >= 13 && FHEB_PIPE_MODE(DP)), PK>(tid, THREADS, polys, gin, gout, smem, tw, m);
            __syncthreads();
            fwd_middle<L, DP, 1, false, PK>(tid, THREADS, polys, smem, tw, m);
            fwd_pass<L, DP, P - 1, IO_SMEM, IO_GLOBAL, true, SCALE, 0, false, false, PK>(tid, THREADS, polys, gin, gout, smem, tw, m, ninv);
            __syncthreads();  // the next group's first pass overwrites the work buffer
        }
    }
}

// Inverse transform (Gentleman-Sande network, bit-reversal folded into the loads, N^-1 folded
// into the last pass).
template <int L, int DP, int THREADS, int PPC, bool TMA = false, int PK = L>
__global__ void __launch_bounds__(THREADS) ntt_inverse_kernel(const uint64_t* in, uint64_t* out, size_t batch,
                                                              const Tw* __restrict__ tw, const Tw ninv, const ModQ m) {
    extern __shared__ __align__(128) uint64_t smem[];
    constexpr int P = Plan<PK>::P;
    static_assert(PK == L || !TMA, "alternative plans run the plain-load form");
    constexpr size_t N = (size_t)1 << L;
    const uint32_t tid = threadIdx.x;
    const size_t groups = (batch + PPC - 1) / PPC;
    if constexpr (TMA && P > 1) {  // see ntt_forward_kernel
        constexpr bool PART = (L == 14) && smem_slot_bytes<DP>() == 8;
        static_assert(!PART || PPC == 1, "partial landing: one polynomial per block");
        constexpr size_t LWORDS = PART ? (size_t)LAND_PART_WORDS : (size_t)PPC * N;
        constexpr int IN0 = PART ? IO_LANDING_PART : IO_LANDING;
        uint64_t* landing = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(smem) + (size_t)PPC * N * smem_slot_bytes<DP>());
        uint64_t* bar = landing + LWORDS;
        auto group_bytes = [&](size_t g) {
            const size_t q0 = g * PPC;
            return PART ? (uint32_t)(LWORDS * 8) : (uint32_t)(((batch - q0) < (size_t)PPC ? (batch - q0) : (size_t)PPC) * N * 8);
        };
        auto request = [&](size_t g) {
            tma_load_1d(landing, in + g * PPC * N, group_bytes(g), bar);
            if constexpr (PART) prefetch_l2_bulk(in + g * N + LWORDS, (uint32_t)((N - LWORDS) * 8));
        };
        if (tid == 0) {
            mbar_init(bar, 1);
            if ((size_t)blockIdx.x < groups) request(blockIdx.x);
        }
        __syncthreads();
        uint32_t parity = 0;
        for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
            const size_t p0 = g * PPC;
            const uint32_t polys = (uint32_t)((batch - p0) < (size_t)PPC ? (batch - p0) : (size_t)PPC);
            uint64_t* gout = out + p0 * N;
            mbar_wait(bar, parity);
            parity ^= 1u;
            inv_pass<L, DP, P - 1, IN0, IO_SMEM>(tid, THREADS, polys, landing, gout, smem, tw, ninv, m, GlobalMap{0, 0}, in + p0 * N);
            __syncthreads();
            if (tid == 0 && g + gridDim.x < groups) request(g + gridDim.x);
            inv_middle<L, DP, P - 2>(tid, THREADS, polys, smem, tw, ninv, m);
            inv_pass<L, DP, 0, IO_SMEM, IO_GLOBAL>(tid, THREADS, polys, nullptr, gout, smem, tw, ninv, m);
            __syncthreads();
        }
        return;
    }
    for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const size_t p0 = g * PPC;
        const uint32_t polys = (uint32_t)((batch - p0) < (size_t)PPC ? (batch - p0) : (size_t)PPC);
        const uint64_t* gin = in + p0 * N;
        uint64_t* gout = out + p0 * N;
        if (tid == 0 && g + gridDim.x < groups) {
            const size_t q0 = (g + gridDim.x) * PPC;
            const size_t np = (batch - q0) < (size_t)PPC ? (batch - q0) : (size_t)PPC;
            prefetch_l2_bulk(in + q0 * N, (uint32_t)(np * N * 8));
        }
        if constexpr (P == 1) {
            inv_pass<L, DP, 0, IO_GLOBAL, IO_GLOBAL>(tid, THREADS, polys, gin, gout, smem, tw, ninv, m);
        } else {
            inv_pass<L, DP, P - 1, IO_GLOBAL, IO_SMEM, true, 0, 1, false, (L >= 14 && FHEB_PIPE_MODE_INV(DP)), PK>(tid, THREADS, polys, gin, gout, smem, tw, ninv, m);
            __syncthreads();
            inv_middle<L, DP, P - 2, false, PK>(tid, THREADS, polys, smem, tw, ninv, m);
            inv_pass<L, DP, 0, IO_SMEM, IO_GLOBAL, true, 0, 1, false, false, PK>(tid, THREADS, polys, gin, gout, smem, tw, ninv, m);
            __syncthreads();
        }
    }
}

// ---- degrees above 2^14 (one polynomial no longer fits the shared memory of an SM) ----------------
// The first D = L - 14 stages pair positions N/2, N/4 apart; after them the 2^D contiguous sub-blocks
// of 2^14 positions are independent.  So: a streaming kernel runs the top D stages in registers
// (global -> scratch), then the 14-stage shared-memory kernel transforms every sub-block with its own
// twiddle table (the tail of the big network: no unit twiddles; 2 * 2^14 entries each, see ntt_plan.hpp) and scatters the results into the
// reference's output order, index (r << D) | bitrev_D(h) for word r of sub-block h.  The inverse runs
// the same two kernels backwards.
template <int D, int DP, bool INVERSE>
__global__ void __launch_bounds__(256) ntt_top_kernel(const uint64_t* in, uint64_t* out, size_t batch, uint32_t L,
                                                      const Tw* __restrict__ tw, const Tw ninv, const ModQ m) {
    constexpr int E = 1 << D;
    const size_t N = (size_t)1 << L, items = N >> D;
    const size_t total = batch * items;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t poly = i / items, j = i - poly * items;
        const uint64_t* src = in + poly * N + j;
        uint64_t* dst = out + poly * N + j;
        uint64_t x[E];
#pragma unroll
        for (int c = 0; c < E; ++c) x[c] = stream_load(src + (size_t)c * items);
        load_words<DP, E>(x, m);
        if constexpr (INVERSE) {
            inv_stages<D, 0, 1, DP, true>(x, tw, 0u, m);
#pragma unroll
            for (int c = 0; c < E; ++c) stream_store(dst + (size_t)c * items, scale_word<DP>(x[c], ninv, m));
        } else {
            fwd_stages<D, 0, 1, DP, true>(x, tw, 0u, m);
            constexpr int KOUT = fwd_pass_k(1, D, true, DP);
#pragma unroll
            for (int c = 0; c < E; ++c) dst[(size_t)c * items] = canon_k<KOUT, DP>(x[c], m);  // re-read soon: keep in L2
        }
    }
}

// 14-stage tail of a larger forward transform: sub-block g of `tmp` (natural positions, canonical)
// -> caller memory in the reference's order.  tables = 2^D consecutive twiddle tables.
template <int L, int DP, int THREADS, bool SCALE>
__global__ void __launch_bounds__(THREADS) ntt_forward_sub_kernel(const uint64_t* tmp, uint64_t* out, size_t subs, uint32_t D,
                                                                  const Tw* __restrict__ tables, const Tw ninv, const ModQ m) {
    extern __shared__ __align__(128) uint64_t smem[];
    constexpr int P = Plan<L>::P;
    constexpr size_t N = (size_t)1 << L;
    const uint32_t tid = threadIdx.x;
    for (size_t g = blockIdx.x; g < subs; g += gridDim.x) {
        const size_t poly = g >> D;
        const uint32_t h = (uint32_t)(g & ((1u << D) - 1u));
        const Tw* tw = reinterpret_cast<const Tw*>(reinterpret_cast<const char*>(tables) + (size_t)h * 2 * N * (DP ? 8 : 16));
        const GlobalMap map{D, bitrev_rt(h, (int)D)};
        fwd_pass<L, DP, 0, IO_GLOBAL, IO_SMEM, true, false, 0, true>(tid, THREADS, 1, tmp + g * N, nullptr, smem, tw, m);
        __syncthreads();
        fwd_middle<L, DP, 1, true>(tid, THREADS, 1, smem, tw, m);
        fwd_pass<L, DP, P - 1, IO_SMEM, IO_GLOBAL, true, SCALE, 0, true>(tid, THREADS, 1, nullptr, out + (poly << (L + D)), smem, tw, m, ninv, map);
        __syncthreads();
    }
}

// 14-stage head of a larger inverse transform: caller memory (reference order) -> `tmp` (natural
// positions, canonical, unscaled); the top D stages and the scaling follow in ntt_top_kernel.
template <int L, int DP, int THREADS>
__global__ void __launch_bounds__(THREADS) ntt_inverse_sub_kernel(const uint64_t* in, uint64_t* tmp, size_t subs, uint32_t D,
                                                                  const Tw* __restrict__ tables, const Tw one, const ModQ m) {
    extern __shared__ __align__(128) uint64_t smem[];
    constexpr int P = Plan<L>::P;
    constexpr size_t N = (size_t)1 << L;
    const uint32_t tid = threadIdx.x;
    for (size_t g = blockIdx.x; g < subs; g += gridDim.x) {
        const size_t poly = g >> D;
        const uint32_t h = (uint32_t)(g & ((1u << D) - 1u));
        const Tw* tw = reinterpret_cast<const Tw*>(reinterpret_cast<const char*>(tables) + (size_t)h * 2 * N * (DP ? 8 : 16));
        const GlobalMap map{D, bitrev_rt(h, (int)D)};
        inv_pass<L, DP, P - 1, IO_GLOBAL, IO_SMEM, true, 0, 1, true>(tid, THREADS, 1, in + (poly << (L + D)), nullptr, smem, tw, one, m, map);
        __syncthreads();
        inv_middle<L, DP, P - 2, true>(tid, THREADS, 1, smem, tw, one, m);
        inv_pass<L, DP, 0, IO_SMEM, IO_GLOBAL, true, 0, 1, true>(tid, THREADS, 1, nullptr, tmp + g * N, smem, tw, one, m);
        __syncthreads();
    }
}

// c = T^-1(T(a) . T(b)) in one launch (PolynomialRing::multiply, polynomial_ring.cpp:421-447).
// T(a) is parked either in a second shared-memory buffer or, when two operands do not fit
// (N = 16384), in a per-block global scratch that stays L2 resident.
template <int L, int DP, int THREADS, int PPC, bool STASH_GLOBAL>
__global__ void __launch_bounds__(THREADS) polymul_kernel(const uint64_t* a, const uint64_t* b, uint64_t* c, size_t batch,
                                                          const Tw* __restrict__ twf, const Tw* __restrict__ twi,
                                                          const Tw ninv, const ModQ m, uint64_t* scratch) {
    extern __shared__ __align__(128) uint64_t smem[];
    constexpr int P = Plan<L>::P;
    constexpr size_t N = (size_t)1 << L;
    constexpr int STASH = STASH_GLOBAL ? IO_STASH_GLOBAL : IO_STASH_SMEM;
    constexpr int UPB = (DP == MODE_U32P) ? PPC / 2 : PPC;  // work-buffer units per block (pairs of polynomials in pair mode)
    uint64_t* stash = STASH_GLOBAL ? scratch + (size_t)blockIdx.x * UPB * N
                                   : reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(smem) + (size_t)UPB * N * smem_slot_bytes<DP>());
    const uint32_t tid = threadIdx.x;
    const size_t groups = (batch + PPC - 1) / PPC;
    for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const size_t p0 = g * PPC;
        const uint32_t polys = (uint32_t)((batch - p0) < (size_t)PPC ? (batch - p0) : (size_t)PPC);
        const uint64_t* ga = a + p0 * N;
        const uint64_t* gb = b + p0 * N;
        uint64_t* gc = c + p0 * N;
        if (tid == 0) prefetch_l2_bulk(gb, (uint32_t)(polys * N * 8));
        if constexpr (P == 1) {
            fwd_pass<L, DP, 0, IO_GLOBAL, STASH, false>(tid, THREADS, polys, ga, stash, smem, twf, m);
            if (!STASH_GLOBAL) __syncthreads();
            polymul_mid_pass<L, DP, IO_GLOBAL, IO_GLOBAL, STASH>(tid, THREADS, polys, gb, gc, smem, stash, twf, twi, ninv, m);
            __syncthreads();
        } else {
            // operand a: all passes, result parked in the stash
            fwd_pass<L, DP, 0, IO_GLOBAL, IO_SMEM>(tid, THREADS, polys, ga, nullptr, smem, twf, m);
            __syncthreads();
            fwd_middle<L, DP, 1>(tid, THREADS, polys, smem, twf, m);
            fwd_pass<L, DP, P - 1, IO_SMEM, STASH, false>(tid, THREADS, polys, nullptr, stash, smem, twf, m);
            __syncthreads();
            // operand b: all but the last pass, then the fused product + first inverse pass
            fwd_pass<L, DP, 0, IO_GLOBAL, IO_SMEM>(tid, THREADS, polys, gb, nullptr, smem, twf, m);
            __syncthreads();
            fwd_middle<L, DP, 1>(tid, THREADS, polys, smem, twf, m);
            polymul_mid_pass<L, DP, IO_SMEM, IO_SMEM, STASH>(tid, THREADS, polys, nullptr, nullptr, smem, stash, twf, twi, ninv, m);
            __syncthreads();
            inv_middle<L, DP, P - 2>(tid, THREADS, polys, smem, twi, ninv, m);
            inv_pass<L, DP, 0, IO_SMEM, IO_GLOBAL>(tid, THREADS, polys, nullptr, gc, smem, twi, ninv, m);
            __syncthreads();
        }
    }
}

}  // namespace fheb
