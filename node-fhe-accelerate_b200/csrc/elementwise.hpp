// Device-pointer entry points of elementwise.cu shared with other translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace fheb {

// op: 0 add, 1 sub, 2 mul, 3 neg, 4 scalar-mul (reference semantics, see elementwise.cu)
int elementwise_device(int op, const uint64_t* a, const uint64_t* b, uint64_t scalar, uint64_t* r, size_t count,
                       uint64_t q, cudaStream_t s);
// sums `count` device rows of `width` words into out[width] (canonical words; tally.cu)
int tally_rows_device(const uint64_t* rows, size_t count, uint32_t width, uint64_t q, uint64_t* out, cudaStream_t s);
// grid for a streaming kernel: enough blocks for the work, capped at SMs x blocks_per_sm
unsigned stream_grid(size_t work_items, int threads, int blocks_per_sm);

}  // namespace fheb
