// Back-to-back fheb_tally calls on device buffers, timed with CUDA events: separates kernel time from launch gaps.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -Iinclude tools/microbench/tally_loop.cu -Lnode-fhe-accelerate_b200 -lfheb200 -o build/tally_loop
#include <cstdio>
#include <cuda_runtime.h>
#include "fheb200.h"
int main(int argc, char** argv) {
    const size_t count = argc > 1 ? atol(argv[1]) : 131072;
    const uint32_t n = 1024;
    const uint64_t q = 1099511678977ULL;
    fheb_init(0);
    uint64_t *cts, *out;
    cudaMalloc(&cts, count * 2 * n * 8);
    cudaMalloc(&out, 2 * n * 8);
    fheb_synth_ballots(cts, 0, count, n, q, 7, nullptr);
    cudaStream_t s;
    cudaStreamCreate(&s);
    for (int i = 0; i < 5; ++i) fheb_tally(cts, count, n, q, out, s);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 50;
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters; ++i) fheb_tally(cts, count, n, q, out, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("fheb_tally %zu ballots: %.1f us per call (%.0f GB/s)\n", count, ms * 1e3 / iters, count * 16384.0 / (ms / iters * 1e-3) / 1e9);
    return 0;
}
