// Interface between relin.cu (key handling, C ABI) and relin_fused.cu (the one-launch kernel).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "modarith.cuh"

namespace fheb {

struct RelinFusedArgs {
    const uint64_t* cts;  // [batch][3][N]
    uint64_t* out;        // [batch][2][N]
    size_t batch;
    const Tw* key;        // packed like a GGSW: [levels rows][E][N/E][2], times N^-1 (Shoup pairs, or doubles in FP64 mode)
    uint32_t levels, base_log;
    uint64_t mask;
    const Tw* twf;
    const Tw* twi;
    Tw ninv;
    ModQ m;
};

constexpr int RELIN_FUSED_UNSUPPORTED = -1000;  // shape outside the fused kernel (rows do not fit an SM, degree below 32 or above 8192)
int relin_fused_launch(uint32_t logn, bool dp, const RelinFusedArgs& a, cudaStream_t stream);

// bootstrap.cu: transforms in the reference's output order -> the packed key layout of boot_mid_pass
// (y = [ggsws][rows][kp1][N]; g = [ggsws][rows][E][N/E][kp1] entries, times N^-1)
struct NttPlan;
int pack_key_rows_device(const NttPlan* p, const uint64_t* y, Tw* g, size_t ggsws, uint32_t kp1, uint32_t rows, cudaStream_t s);

}  // namespace fheb
