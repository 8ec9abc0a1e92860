// Store-path microbenchmark: one persistent block per SM (128 KB of dynamic shared memory, like the
// N=16384 transform kernel) writes 128 KB "polynomials" with different patterns; no compute.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 16384, THREADS = 512;

template <int MODE>
__global__ void __launch_bounds__(THREADS) k(uint64_t* out, const uint64_t* in, int polys, int smem_words) {
    extern __shared__ uint64_t sm[];
    const int tid = threadIdx.x;
    uint64_t acc = 0;
    for (int p = blockIdx.x; p < polys; p += gridDim.x) {
        uint64_t* dst = out + (size_t)p * N;
        const uint64_t* src = in + (size_t)p * N;
        if (MODE == 0) {  // the transform kernel's pattern: 4 iterations x 8 runs, 8-byte stores, 256 B per warp
            for (int k2 = 0; k2 < N / 8 / THREADS; ++k2)
#pragma unroll
                for (int r = 0; r < 8; ++r) dst[r * (N / 8) + k2 * THREADS + tid] = (uint64_t)p + r + tid;
        } else if (MODE == 1) {  // contiguous 16-byte stores
            for (int i = tid; i < N / 2; i += THREADS) reinterpret_cast<ulonglong2*>(dst)[i] = make_ulonglong2(p, i);
        } else if (MODE == 2) {  // contiguous 8-byte stores
            for (int i = tid; i < N; i += THREADS) dst[i] = (uint64_t)p + i;
        } else if (MODE == 3) {  // pass-0 load pattern only: 16 x 8-byte loads strided by 8 KB
            for (int it = 0; it < N / 16 / THREADS; ++it)
#pragma unroll
                for (int c = 0; c < 16; ++c) acc += src[c * (N / 16) + it * THREADS + tid];
        } else if (MODE == 4) {  // load pattern + store pattern
            for (int it = 0; it < N / 16 / THREADS; ++it)
#pragma unroll
                for (int c = 0; c < 16; ++c) acc += src[c * (N / 16) + it * THREADS + tid];
            for (int k2 = 0; k2 < N / 8 / THREADS; ++k2)
#pragma unroll
                for (int r = 0; r < 8; ++r) dst[r * (N / 8) + k2 * THREADS + tid] = acc + r;
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}

static int ROT = 1;
template <int MODE>
void run(const char* name, uint64_t* out, const uint64_t* in, int polys, int blocks_per_sm, double bytes_per_poly) {
    size_t smem = blocks_per_sm == 1 ? 128 * 1024 : 32 * 1024;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = 148 * blocks_per_sm;
    k<MODE><<<grid, THREADS, smem>>>(out, in, polys, 0);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) k<MODE><<<grid, THREADS, smem>>>(out + (size_t)(i % ROT) * polys * N, in + (size_t)(i % ROT) * polys * N, polys, 0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    ms /= 10;
    printf("%-44s blocks/SM=%d  %7.4f ms  %7.1f GB/s\n", name, blocks_per_sm, ms, bytes_per_poly * polys / ms / 1e6);
}

int main() {
    const int polys = 1024;
    uint64_t *out, *in;
    cudaMalloc(&out, (size_t)polys * N * 8 * 8);
    cudaMalloc(&in, (size_t)polys * N * 8 * 8);
    cudaMemset(in, 1, (size_t)polys * N * 8 * 8);
    for (int rot : {1, 8})
    for (int b : {1, 4}) {
        ROT = rot;
        printf("-- %d rotating buffers of 134 MB --\n", rot);
        run<0>("store: 8 runs x 8 B (transform pattern)", out, in, polys, b, N * 8.0);
        run<1>("store: contiguous 16 B", out, in, polys, b, N * 8.0);
        run<2>("store: contiguous 8 B", out, in, polys, b, N * 8.0);
        run<3>("load : 16 strided 8 B (pass-0 pattern)", out, in, polys, b, N * 8.0);
        run<4>("load + store (transform patterns)", out, in, polys, b, N * 16.0);
    }
    return 0;
}
