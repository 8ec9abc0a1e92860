// Relinearisation as ONE launch (EncryptionEngine::relinearize, cpp/src/encryption.cpp:904-993; see relin.cu for the
// algebra).  A thread block owns a ciphertext: the raw-bit digits of c2 are cut in registers while the first forward pass
// reads c2 ONCE, the `levels` digit rows stay in shared memory through the forward passes, the last forward pass of
// every row, the multiply-accumulate against the transformed key and the first inverse pass run on the same
// register-resident positions (boot_mid_pass: the relinearisation key is packed exactly like a GGSW with rows = levels
// and two components, N^-1 folded in), and the last inverse pass adds c0 / c1 (reduced first, PolynomialRing::add_inplace)
// and stores the result.  Digits, transformed digits and the products never touch global memory; the unfused path of
// relin.cu (5 launches, three round trips through HBM) remains for shapes whose rows do not fit an SM.
#include <cstdlib>

#include "boot_core.cuh"
#include "plan.hpp"
#include "relin_fused.hpp"
#include "runtime.hpp"

namespace fheb {

// first forward pass: digit l of c2 = (c2 >> l * base_log) & mask (raw word, level 0 = low bits, :948-955)
template <int L, bool DP>
__device__ __forceinline__ void relin_first_pass(uint32_t tid, uint32_t nthreads, const uint64_t* __restrict__ c2, uint64_t* work,
                                                 uint32_t levels, uint32_t base_log, uint64_t mask, const Tw* __restrict__ tw,
                                                 const ModQ& m) {
    constexpr int R = Plan<L>::R[0];
    constexpr int E = 1 << R;
    constexpr int EB = L - R;
    constexpr uint32_t N = 1u << L;
    constexpr uint32_t ITEMS = N >> R;
    for (uint32_t u = tid; u < ITEMS; u += nthreads) {
        uint64_t d[E];
#pragma unroll
        for (int e = 0; e < E; ++e) d[e] = stream_load(c2 + (u | ((uint32_t)e << EB)));
        const uint32_t pb = swz(u);
        for (uint32_t l = 0; l < levels; ++l) {
            const uint32_t shift = l * base_log;
            uint64_t x[E];
#pragma unroll
            for (int e = 0; e < E; ++e) x[e] = load_word<DP>((d[e] >> shift) & mask, m);  // the transform reduces its input words
            fwd_stages<R, 0, 1, DP, true>(x, tw, 0u, m);
            const SlotRef<MODE_INT> dst = slot_ref<MODE_INT>(work, l, N, pb);
#pragma unroll
            for (int e = 0; e < E; ++e) slot_store<MODE_INT>(dst, swz((uint32_t)e << EB), x[e]);
        }
    }
}

template <int L, bool DP, int PH>
__device__ __forceinline__ void relin_phases(uint32_t tid, uint32_t nthreads, const BootStep& s, const RelinFusedArgs& a) {
    if constexpr (PH < boot_phases<L>()) {
        boot_phase<L, DP, 2, PH, true>(tid, nthreads, s, a.twf, a.twi, a.ninv, a.m);
        __syncthreads();
        relin_phases<L, DP, PH + 1>(tid, nthreads, s, a);
    }
}

template <int L, bool DP, int THREADS>
__global__ void __launch_bounds__(THREADS) relin_fused_kernel(const RelinFusedArgs a) {
    extern __shared__ __align__(128) uint64_t smem[];
    constexpr size_t N = (size_t)1 << L;
    const uint32_t tid = threadIdx.x;
    for (size_t ct = blockIdx.x; ct < a.batch; ct += gridDim.x) {
        const uint64_t* c = a.cts + ct * 3 * N;
        BootStep s;
        s.acc = const_cast<uint64_t*>(c);  // (c0, c1): read by the final pass only
        s.work = smem;
        s.diff = nullptr;
        s.diff_sub = nullptr;
        s.ggsw = a.key;
        s.gout = a.out + ct * 2 * N;
        s.rot = 0;
        s.levels = a.levels;
        s.rows = a.levels;
        s.base_log = a.base_log;
        s.add_acc = 1;
        s.maybe_raw = 1;  // caller words: add_inplace reduces both sides
        relin_first_pass<L, DP>(tid, THREADS, c + 2 * N, smem, a.levels, a.base_log, a.mask, a.twf, a.m);
        __syncthreads();
        relin_phases<L, DP, 1>(tid, THREADS, s, a);  // every phase ends with a barrier: the next ciphertext may overwrite the rows
    }
}

template <int L, bool DP, int THREADS_ = 0>
static int relin_launch_one(const RelinFusedArgs& a, cudaStream_t stream) {
    // threads per ciphertext = items of the first and of the fused middle phase (N / 16 from N = 2048 up): both walk one
    // item per thread over ALL rows, so more threads than items idle there.  Measured at N = 4096, four levels, 62-bit
    // prime: 128 / 256 / 512 threads 0.940 / 0.506 / 0.746 ms per 2048 ciphertexts (the five-launch path: 0.693 ms); N = 2048,
    // three levels: 128 / 256 threads 0.382 / 0.367 ms per 4096.
    constexpr int THREADS = THREADS_ ? THREADS_ : (L <= 8) ? 64 : (L <= 10) ? 128 : (L <= 12) ? 256 : 512;
    const size_t rows = a.levels < 2 ? 2 : a.levels;  // the two product components reuse rows 0 and 1
    const size_t smem = rows * ((size_t)8 << L);
    if (smem > (size_t)ctx().prop.sharedMemPerBlockOptin) return RELIN_FUSED_UNSUPPORTED;
    auto k = relin_fused_kernel<L, DP, THREADS>;
    if (smem > 48 * 1024) FHEB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    FHEB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, THREADS, smem));
    if (bps < 1) return RELIN_FUSED_UNSUPPORTED;
    const size_t resident = (size_t)ctx().sm_count * (size_t)bps;
    const unsigned grid = (unsigned)(a.batch < resident ? a.batch : resident);
    k<<<grid, THREADS, smem, stream>>>(a);
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

int relin_fused_launch(uint32_t logn, bool dp, const RelinFusedArgs& a, cudaStream_t stream) {
#define FHEB_RELIN_CASE(L_) \
    case L_:                \
        return dp ? relin_launch_one<L_, true>(a, stream) : relin_launch_one<L_, false>(a, stream);
    if (const char* e = getenv("FHEB_RELIN_THREADS")) {  // experiment
        const int t = atoi(e);
        if (logn == 12 && !dp && t == 512) return relin_launch_one<12, false, 512>(a, stream);
        if (logn == 11 && !dp && t == 128) return relin_launch_one<11, false, 128>(a, stream);
    }
    switch (logn) {
        FHEB_RELIN_CASE(5)
        FHEB_RELIN_CASE(6)
        FHEB_RELIN_CASE(7)
        FHEB_RELIN_CASE(8)
        FHEB_RELIN_CASE(9)
        FHEB_RELIN_CASE(10)
        FHEB_RELIN_CASE(11)
        FHEB_RELIN_CASE(12)
        FHEB_RELIN_CASE(13)
    }
#undef FHEB_RELIN_CASE
    return RELIN_FUSED_UNSUPPORTED;
}

}  // namespace fheb
