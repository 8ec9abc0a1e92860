// Interface between tally.cu (C ABI of the tensor product) and tensor_fused.cu (the one-launch kernel).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace fheb {

struct NttPlan;
constexpr int TENSOR_FUSED_UNSUPPORTED = -1001;  // four polynomials do not fit an SM (N >= 8192), or a single-pass degree (N <= 16)
int tensor_fused_launch(const NttPlan* p, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out, size_t batch, cudaStream_t stream);

}  // namespace fheb
