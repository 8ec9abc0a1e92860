// Pipe-throughput microbenchmark: DFMA vs IMAD.WIDE vs IMAD (per SM per clock), used to choose
// between the integer and the FP64 formulation of the modular butterfly.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(uint64_t* out, int iters, double da, uint64_t ua) {
    double d[8];
    uint64_t u[8];
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = threadIdx.x + i; u[i] = threadIdx.x * 77 + i; v[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) d[i] = fma(d[i], da, d[(i + 1) & 7]);
            if (MODE == 1) u[i] = (uint64_t)(uint32_t)u[i] * (uint32_t)ua + u[(i + 1) & 7];   // IMAD.WIDE.U32
            if (MODE == 2) v[i] = v[i] * (uint32_t)ua + v[(i + 1) & 7];                        // IMAD
            if (MODE == 3) u[i] = __umul64hi(u[i], ua) + u[(i + 1) & 7];
            if (MODE == 5) v[i] = __funnelshift_l(v[i], v[(i + 1) & 7], 3) ^ (uint32_t)ua;       // SHF + LOP3 (ALU pipe)
            if (MODE == 6) v[i] = __umulhi(v[i], (uint32_t)ua) + v[(i + 1) & 7];                   // IMAD.HI.U32
            if (MODE == 7) d[i] = rint(d[i] * da) + d[(i + 1) & 7];                                // DMUL + FRND.F64 + DADD
            if (MODE == 8) d[i] = rint(d[i]);                                                       // FRND.F64 alone (dependent chain per element)
            if (MODE == 4) { d[i] = fma(d[i], da, d[(i + 1) & 7]); u[i] = (uint64_t)(uint32_t)u[i] * (uint32_t)ua + u[(i + 1) & 7]; }
        }
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += (uint64_t)d[i] + u[i] + v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static bool g_json = false;
static double g_tops[10];

template <int MODE>
void run(const char* name, int ops_per_iter) {
    uint64_t* out;
    cudaMalloc(&out, 148 * 4 * 512 * 8);
    const int iters = 20000;
    k<MODE><<<148 * 4, 512>>>(out, 100, 1.0000001, 12345);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<148 * 4, 512>>>(out, iters, 1.0000001, 12345);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)148 * 4 * 512 * iters * ops_per_iter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    g_tops[MODE] = ops / ms / 1e9;
    if (!g_json)
        printf("%-28s %8.3f ms  %7.2f Tops/s  %6.1f ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms / 1e9, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}

// usage: pipes [--json] [device]   (--json: one line of thread-operations per second, for bench.py's roofline)
int main(int argc, char** argv) {
    int dev = 0;
    for (int i = 1; i < argc; ++i) {
        if (std::string(argv[i]) == "--json") g_json = true;
        else dev = atoi(argv[i]);
    }
    if (cudaSetDevice(dev) != cudaSuccess) { fprintf(stderr, "no device %d\n", dev); return 1; }
    run<0>("DFMA", 8);
    run<1>("IMAD.WIDE.U32", 8);
    run<2>("IMAD (32-bit)", 8);
    run<3>("umul64hi", 8);
    run<4>("DFMA + IMAD.WIDE together", 16);
    run<5>("SHF + LOP3 (ALU pipe)", 16);
    run<6>("IMAD.HI.U32", 8);
    if (!g_json) {
        run<7>("DMUL + FRND.F64 + DADD (3 ops)", 24);
        run<8>("FRND.F64 alone", 8);
    }
    if (g_json) {
        int clk = 0, sms = 0;
        cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        printf("{\"dfma_tops\": %.4f, \"imad_wide_tops\": %.4f, \"imad_tops\": %.4f, \"umul64hi_tops\": %.4f, "
               "\"dfma_plus_imad_wide_tops\": %.4f, \"alu_tops\": %.4f, \"imad_hi_tops\": %.4f, \"sm_count\": %d, \"nominal_mhz\": %d}\n",
               g_tops[0], g_tops[1], g_tops[2], g_tops[3], g_tops[4], g_tops[5], g_tops[6], sms, clk / 1000);
    }
    return 0;
}
