// EncryptionEngine::multiply (cpp/src/encryption.cpp:737-798) as ONE launch: T on the four operand polynomials,
// c0 = a0.b0, c1 = a0.b1 + a1.b0, c2 = a1.b1, T^-1 on the three results.  A thread block owns a pair of ciphertexts: the
// four polynomials stay in shared memory (position order, so no bit reversal happens in either direction), the products
// are formed in place and only the degree-2 ciphertext goes back to HBM - 32 N bytes in, 24 N bytes out, against four
// launches and 136 N bytes of traffic through a cudaMallocAsync'ed scratch for the unfused path (tally.cu), which stays
// for the shapes whose four polynomials do not fit an SM (N >= 8192) and for the single-pass degrees (N <= 16).
#include "ntt_device.cuh"
#include "plan.hpp"
#include "runtime.hpp"
#include "tensor_fused.hpp"

namespace fheb {

template <int L, int DP, int THREADS>
__global__ void __launch_bounds__(THREADS) tensor_fused_kernel(const uint64_t* ct1, const uint64_t* ct2, uint64_t* out, size_t batch,
                                                               const Tw* __restrict__ twf, const Tw* __restrict__ twi, const Tw ninv,
                                                               const ModQ m) {
    extern __shared__ __align__(128) uint64_t smem[];
    constexpr int P = Plan<L>::P;
    constexpr uint32_t N = 1u << L;
    constexpr int KOUT = fwd_pass_k(plan_fwd_kin<L, DP, P - 1>(), Plan<L>::R[P - 1], false, DP);  // bound after the last forward pass
    static_assert(P >= 2, "single-pass degrees use the unfused path");
    static_assert(DP == MODE_INT || DP == MODE_DP, "integer or FP64 arithmetic");
    const uint32_t tid = threadIdx.x;
    for (size_t ct = blockIdx.x; ct < batch; ct += gridDim.x) {
        // rows 0, 1 = a0, a1; rows 2, 3 = b0, b1
        fwd_pass<L, DP, 0, IO_GLOBAL, IO_SMEM>(tid, THREADS, 2, ct1 + ct * 2 * N, nullptr, smem, twf, m);
        fwd_pass<L, DP, 0, IO_GLOBAL, IO_SMEM>(tid, THREADS, 2, ct2 + ct * 2 * N, nullptr, smem + 2 * (size_t)N, twf, m);
        __syncthreads();
        fwd_middle<L, DP, 1>(tid, THREADS, 4, smem, twf, m);
        fwd_pass<L, DP, P - 1, IO_SMEM, IO_SMEM>(tid, THREADS, 4, nullptr, nullptr, smem, twf, m);
        __syncthreads();
        // the same slot of every row holds the same position: the products need no index math at all
        for (uint32_t i = tid; i < N; i += THREADS) {
            if constexpr (DP == MODE_DP) {
                const double a0 = dp_reduce(bits_to_double(smem[i]), m), a1 = dp_reduce(bits_to_double(smem[N + i]), m);
                const double b0 = bits_to_double(smem[2 * N + i]), b1 = bits_to_double(smem[3 * N + i]);  // lazy: |b| < KOUT q
                static_assert(KOUT <= 2 * CAP_DP, "FP64 product bound");
                smem[i] = double_to_bits(dp_mulmod(a0, b0, m));
                smem[N + i] = double_to_bits(dp_reduce(dp_add(dp_mulmod(a0, b1, m), dp_mulmod(a1, b0, m)), m));
                smem[2 * N + i] = double_to_bits(dp_mulmod(a1, b1, m));
            } else {
                const uint64_t a0 = canon_k<KOUT, DP>(smem[i], m), a1 = canon_k<KOUT, DP>(smem[N + i], m);
                const uint64_t b0 = canon_k<KOUT, DP>(smem[2 * N + i], m), b1 = canon_k<KOUT, DP>(smem[3 * N + i], m);
                smem[i] = mulmod(a0, b0, m);
                smem[N + i] = addmod_canon(mulmod(a0, b1, m), mulmod(a1, b0, m), m.q);
                smem[2 * N + i] = mulmod(a1, b1, m);
            }
        }
        __syncthreads();
        inv_middle<L, DP, P - 1>(tid, THREADS, 3, smem, twi, ninv, m);
        inv_pass<L, DP, 0, IO_SMEM, IO_GLOBAL>(tid, THREADS, 3, nullptr, out + ct * 3 * N, smem, twi, ninv, m);
        __syncthreads();  // the next pair overwrites the rows
    }
}

template <int L, int DP>
static int tensor_launch_one(const NttPlan* p, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out, size_t batch, cudaStream_t stream) {
    constexpr int THREADS = (L <= 8) ? 64 : (L <= 10) ? 128 : (L == 11) ? 256 : 512;
    const size_t smem = (size_t)32 << L;  // four rows of N eight-byte slots
    if (smem > (size_t)ctx().prop.sharedMemPerBlockOptin) return TENSOR_FUSED_UNSUPPORTED;
    auto k = tensor_fused_kernel<L, DP, THREADS>;
    if (smem > 48 * 1024) FHEB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    FHEB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, THREADS, smem));
    if (bps < 1) return TENSOR_FUSED_UNSUPPORTED;
    const size_t resident = (size_t)ctx().sm_count * (size_t)bps;
    k<<<(unsigned)(batch < resident ? batch : resident), THREADS, smem, stream>>>(ct1, ct2, out, batch, p->d_fwd, p->d_inv, p->ninv, p->mod);
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

int tensor_fused_launch(const NttPlan* p, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out, size_t batch, cudaStream_t stream) {
    if (p->top != 0) return TENSOR_FUSED_UNSUPPORTED;
    const bool dp = p->mod.dp != 0;
#define FHEB_TENSOR_CASE(L_) \
    case L_:                 \
        return dp ? tensor_launch_one<L_, MODE_DP>(p, ct1, ct2, out, batch, stream) : tensor_launch_one<L_, MODE_INT>(p, ct1, ct2, out, batch, stream);
    switch (p->logn) {
        FHEB_TENSOR_CASE(5)
        FHEB_TENSOR_CASE(6)
        FHEB_TENSOR_CASE(7)
        FHEB_TENSOR_CASE(8)
        FHEB_TENSOR_CASE(9)
        FHEB_TENSOR_CASE(10)
        FHEB_TENSOR_CASE(11)
        FHEB_TENSOR_CASE(12)
    }
#undef FHEB_TENSOR_CASE
    return TENSOR_FUSED_UNSUPPORTED;
}

}  // namespace fheb
