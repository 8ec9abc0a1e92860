// The object behind fheb_ntt_plan: the reference's TwiddleFactors (host copies, natural
// exponent order - cpp/include/ntt_processor.h:29-41) plus the device-side heap tables.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <vector>

#include "modarith.cuh"

namespace fheb {

struct NttPlan {
    uint32_t degree = 0;
    uint32_t logn = 0;
    uint64_t modulus = 0;
    uint64_t psi = 0, psi_inv = 0, inv_n = 0;
    std::vector<uint64_t> fwd_table, inv_table;  // as the reference holds them
    bool unit_first = false;
    ModQ mod{};
    Tw ninv{};
    Tw* d_fwd = nullptr;  // heap-ordered twiddles, N entries: (w, w') pairs, or doubles when mod.dp
    Tw* d_inv = nullptr;
    // degrees above 2^14 (top = log2 N - 14 > 0): d_fwd / d_inv hold 2^top consecutive sub-block tables
    // of 2^14 entries, d_top_* the few twiddles of the first `top` stages, `one` the multiplicand 1
    uint32_t top = 0;
    Tw* d_top_fwd = nullptr;
    Tw* d_top_inv = nullptr;
    Tw one{};
    // moduli below 2^27 (e.g. 132120577): a second table set for the 32-bit kernels of the plain transforms and the fused
    // product (8 bytes per entry: value and 32-bit Shoup companion); everything else keeps using the tables above
    Tw* d_fwd32 = nullptr;
    Tw* d_inv32 = nullptr;
    Tw ninv32{};
    // N = 16384: the plain forward / inverse kernels run three-pass splits (plan keys 78 / 79, ntt_core.cuh) from tables of their own
    Tw* d_fwd_alt = nullptr;    // integer mode
    Tw* d_fwd32_alt = nullptr;  // 32-bit mode
    Tw* d_inv_alt = nullptr;    // inverse: 4 + 5 + 5 (plan key 79), integer mode
    Tw* d_inv32_alt = nullptr;  // 32-bit mode
    // the device whose memory holds the tables above, and copies of the plan on other devices (made on first use when
    // a host batch is spread over several GPUs; owned by this plan)
    int device = 0;
    mutable std::map<int, NttPlan*> replicas;
    mutable std::mutex replica_mutex;
};

// the plan's tables on `device` (the calling thread's current device): the plan itself or a replica built on demand
const NttPlan* plan_on_device(const NttPlan* p, int device);

// device-pointer entry points shared between translation units
int ntt_forward_device(const NttPlan* p, const uint64_t* in, uint64_t* out, size_t batch, cudaStream_t s);
int ntt_inverse_device(const NttPlan* p, const uint64_t* in, uint64_t* out, size_t batch, cudaStream_t s);

}  // namespace fheb
