// Register-resident butterfly rate: the library's own lazy butterflies (csrc/ntt_core.cuh, the code the transform kernels
// run) on 16 values per thread with the twiddles held in registers - no shared memory, no global memory, no barriers.
// What this loop reaches is the ceiling of ANY kernel built from these butterflies on this GPU, with the real instruction
// mix (wide and narrow multiplies, carries, conditional subtractions) issued together; the per-pipe microbenchmarks
// (pipes.cu) time every instruction class alone.
//   bfly [--json] [device]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <cuda_runtime.h>

#include "../../node-fhe-accelerate_b200/csrc/ntt_core.cuh"

using namespace fheb;

template <int MODE, bool INVERSE, int THREADS>
__global__ void __launch_bounds__(THREADS) bfly_loop(uint64_t* out, const Tw* tw, int iters, const ModQ m) {
    extern __shared__ uint64_t pad[];  // only to pin the number of resident blocks
    uint64_t x[16];
    Tw w[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) w[i] = load_tw<MODE>(tw, (threadIdx.x * 15 + i) & 1023);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        uint64_t v = ((uint64_t)(threadIdx.x * 16 + i) * 0x9E3779B97F4A7C15ull) % m.q;
        x[i] = load_word<MODE>(v, m);
    }
    constexpr int K = (MODE == MODE_INT) ? 4 : (MODE == MODE_DP ? 1 : 2);
    for (int it = 0; it < iters; ++it) {
        if constexpr (INVERSE) inv_stages<4, 4, (MODE == MODE_INT ? 2 : K), MODE, false, 3, true>(x, w, 0u, m);
        else fwd_stages<4, 4, K, MODE, false, 0, true>(x, w, 0u, m);
        if constexpr (MODE == MODE_DP) {  // keep the FP64 range bounded the way a pass boundary does
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = double_to_bits(dp_reduce(bits_to_double(x[i]), m));
        } else if constexpr (MODE == MODE_U32) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = lazy32((uint32_t)x[i], m);
        }
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= x[i];
    out[(size_t)blockIdx.x * THREADS + threadIdx.x] = acc;
    if (pad[0] == 0x1234567) out[0] = pad[1];
}

// One middle pass of the N=16384 transform in steady state: the library's own fwd_pass (shared memory -> registers ->
// shared memory, swizzled, twiddles from the table in global memory) + the block barrier, repeated on one polynomial.
// NOBAR: the same without the barrier (wrong results, timing only): what the barrier itself costs.
template <int MODE, int PASS, int THREADS, bool NOBAR>
__global__ void __launch_bounds__(THREADS) pass_loop(uint64_t* out, const Tw* tw, int iters, const ModQ m) {
    extern __shared__ __align__(16) uint64_t smem[];
    constexpr uint32_t N = 1u << 14;
    for (uint32_t i = threadIdx.x; i < N; i += THREADS) {
        uint64_t v = ((uint64_t)(blockIdx.x * N + i) * 0x9E3779B97F4A7C15ull) % m.q;
        sm_store<MODE>(smem, 0, N, i, load_word<MODE>(v, m));
    }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        fwd_pass<14, MODE, PASS, IO_SMEM, IO_SMEM>(threadIdx.x, THREADS, 1, nullptr, nullptr, smem, tw, m);
        if constexpr (MODE == MODE_DP) {  // FP64 range: reduce in place now and then (not timed separately; 1 of 8 iterations)
            if ((it & 7) == 7) {
                __syncthreads();
                for (uint32_t i = threadIdx.x; i < N; i += THREADS) smem[i] = double_to_bits(dp_reduce(bits_to_double(smem[i]), m));
            }
        }
        if constexpr (!NOBAR) __syncthreads();
    }
    __syncthreads();
    uint64_t acc = 0;
    for (uint32_t i = threadIdx.x; i < N; i += THREADS) acc ^= sm_load<MODE>(smem, 0, N, i);
    out[(size_t)blockIdx.x * THREADS + threadIdx.x] = acc;
}

static bool g_json = false;

template <int MODE, int PASS, int THREADS, bool NOBAR>
double run_pass(const char* name, uint64_t q, const Tw* d_tw, int dev) {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const ModQ m = make_modq(q);
    auto kern = pass_loop<MODE, PASS, THREADS, NOBAR>;
    const int smem = (1 << 14) * (MODE == MODE_U32 ? 4 : 8);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, THREADS, smem);
    const int grid = sms * bps;
    uint64_t* out;
    cudaMalloc(&out, (size_t)grid * THREADS * 8);
    const int iters = 2000;
    kern<<<grid, THREADS, smem>>>(out, d_tw, 20, m);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<<<grid, THREADS, smem>>>(out, d_tw, iters, m);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "%s failed\n", name); exit(1); }
    const int R = Plan<14>::R[PASS];
    const double bf = (double)grid * iters * 8192.0 * R;
    const double rate = bf / (ms * 1e-3);
    if (!g_json)
        printf("%-44s %d x %d threads/SM  %8.3f ms  %8.1f G butterfly/s  %5.2f butterflies/clk/SM\n", name, bps, THREADS, ms, rate / 1e9, rate / sms / (clk * 1e3));
    cudaFree(out);
    return rate;
}

template <int MODE, bool INVERSE, int THREADS>
double run(const char* name, uint64_t q, int blocks_per_sm, const Tw* d_tw, int dev) {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const ModQ m = make_modq(q);
    auto kern = bfly_loop<MODE, INVERSE, THREADS>;
    const int smem = (int)(220 * 1024 / blocks_per_sm) & ~1023;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    uint64_t* out;
    const int grid = sms * blocks_per_sm;
    cudaMalloc(&out, (size_t)grid * THREADS * 8);
    const int iters = 4000;
    kern<<<grid, THREADS, smem>>>(out, d_tw, 50, m);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<<<grid, THREADS, smem>>>(out, d_tw, iters, m);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "%s failed\n", name); exit(1); }
    const double bf = (double)grid * THREADS * iters * 32.0;
    const double rate = bf / (ms * 1e-3);
    if (!g_json)
        printf("%-44s %d x %d threads/SM  %8.3f ms  %8.1f G butterfly/s  %5.2f butterflies/clk/SM (at %d MHz nominal)\n", name, blocks_per_sm, THREADS, ms,
               rate / 1e9, rate / sms / (clk * 1e3), clk / 1000);
    cudaFree(out);
    return rate;
}

int main(int argc, char** argv) {
    int dev = 0;
    for (int i = 1; i < argc; ++i) {
        if (std::string(argv[i]) == "--json") g_json = true;
        else dev = atoi(argv[i]);
    }
    if (cudaSetDevice(dev) != cudaSuccess) { fprintf(stderr, "no device %d\n", dev); return 1; }
    // twiddle stand-ins: residues with their Shoup companions (the values do not matter for the rate)
    const uint64_t q62 = 4611686018326724609ull, q40 = 1099511678977ull, q27 = 132120577ull;
    constexpr int TWN = 1 << 16;
    static Tw h[3][TWN];
    for (int t = 0; t < 3; ++t) {
        const uint64_t q = t == 0 ? q62 : (t == 1 ? q40 : q27);
        for (int i = 0; i < TWN; ++i) {
            const uint64_t w = ((uint64_t)(i + 3) * 0xD1B54A32D192ED03ull) % q;
            if (t == 0) h[t][i] = Tw{w, shoup_companion(w, q)};
            else if (t == 1) {
                double d = (double)w, dq = d / (double)q;
                uint64_t b, bq;
                memcpy(&b, &d, 8);
                memcpy(&bq, &dq, 8);
                (void)bq;  // (a table of (w, w/q) pairs was tried: register-resident +3 %, last pass -32 %: twice the twiddle traffic)
                reinterpret_cast<uint64_t*>(h[t])[i] = b;
            }
            else reinterpret_cast<uint64_t*>(h[t])[i] = w | ((((uint64_t)w << 32) / q) << 32);
        }
    }
    Tw* d_tw;
    cudaMalloc(&d_tw, sizeof(h));
    cudaMemcpy(d_tw, h, sizeof(h), cudaMemcpyHostToDevice);
    double r[8];
    r[0] = run<MODE_INT, false, 512>("integer forward (q < 2^62)", q62, 1, d_tw, dev);
    r[1] = run<MODE_INT, true, 512>("integer inverse (q < 2^62)", q62, 1, d_tw, dev);
    r[2] = run<MODE_INT, false, 256>("integer forward (q < 2^62)", q62, 2, d_tw, dev);
    r[3] = run<MODE_DP, false, 512>("FP64 forward (q < 2^42)", q40, 1, d_tw + TWN, dev);
    r[4] = run<MODE_DP, true, 512>("FP64 inverse (q < 2^42)", q40, 1, d_tw + TWN, dev);
    r[5] = run<MODE_U32, false, 256>("32-bit forward (q < 2^27)", q27, 2, d_tw + 2 * TWN, dev);
    r[6] = run<MODE_U32, true, 256>("32-bit inverse (q < 2^27)", q27, 2, d_tw + 2 * TWN, dev);
    if (!g_json) {
        run_pass<MODE_INT, 1, 512, false>("integer pass 1 (4 stages, smem <-> regs)", q62, d_tw, dev);
        run_pass<MODE_INT, 1, 512, true>("integer pass 1, no barrier (timing only)", q62, d_tw, dev);
        run_pass<MODE_INT, 2, 512, false>("integer pass 2 (3 stages)", q62, d_tw, dev);
        run_pass<MODE_INT, 3, 512, false>("integer pass 3 (3 stages, per-item twiddles)", q62, d_tw, dev);
        run_pass<MODE_INT, 3, 512, true>("integer pass 3, no barrier (timing only)", q62, d_tw, dev);
        run_pass<MODE_DP, 1, 512, false>("FP64 pass 1 (4 stages)", q40, d_tw + TWN, dev);
        run_pass<MODE_DP, 3, 512, false>("FP64 pass 3 (3 stages)", q40, d_tw + TWN, dev);
        run_pass<MODE_U32, 1, 256, false>("32-bit pass 1 (4 stages)", q27, d_tw + 2 * TWN, dev);
        run_pass<MODE_U32, 3, 256, false>("32-bit pass 3 (3 stages)", q27, d_tw + 2 * TWN, dev);
    }
    if (g_json)
        printf("{\"int_fwd_gbfly\": %.2f, \"int_inv_gbfly\": %.2f, \"int_fwd_2x256_gbfly\": %.2f, \"dp_fwd_gbfly\": %.2f, \"dp_inv_gbfly\": %.2f, "
               "\"u32_fwd_gbfly\": %.2f, \"u32_inv_gbfly\": %.2f}\n", r[0] / 1e9, r[1] / 1e9, r[2] / 1e9, r[3] / 1e9, r[4] / 1e9, r[5] / 1e9, r[6] / 1e9);
    return 0;
}
