// Small-problem latency through the C ABI, measured from C++ (no Python in the loop): BASELINE config C1 (forward + inverse
// transform, N = 1024, q = 132120577, batch 1 - the published M4 Max Montgomery row is 8.86 us per transform,
// NTT_(degree=1024).csv:5) and C3 at its BASELINE size (two-limb Montgomery products, n = 65536).
//   nvcc -O2 -o tools/latency tools/latency.cu -Inode-fhe-accelerate_b200/../include -Lnode-fhe-accelerate_b200 -lfheb200
// Three figures per workload: host time to ISSUE a call (asynchronous, device buffers), device time per call in a
// back-to-back stream (CUDA events over 2000 calls), and the time from issue to completion of ONE call (host clock around
// call + synchronize).  With --graph the back-to-back figure is also taken from a captured CUDA graph of 100 calls.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "fheb200.h"

#define CK(x)                                                                  \
    do {                                                                       \
        int rc_ = (x);                                                         \
        if (rc_ != 0) {                                                        \
            std::printf("FAILED %s: %s\n", #x, fheb_last_error());             \
            return 1;                                                          \
        }                                                                      \
    } while (0)

struct Row {
    const char* key;
    double issue_us, stream_us, done_us, graph_us;
};
static std::vector<Row> g_rows;
static bool g_json = false;

template <class F>
static void measure(const char* key, const char* name, cudaStream_t s, F call, bool graph) {
    using clk = std::chrono::steady_clock;
    for (int i = 0; i < 50; ++i) call();
    cudaStreamSynchronize(s);
    const int iters = 2000;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const auto t0 = clk::now();
    cudaEventRecord(a, s);
    for (int i = 0; i < iters; ++i) call();
    cudaEventRecord(b, s);
    const auto t1 = clk::now();
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double issue_us = std::chrono::duration<double, std::micro>(t1 - t0).count() / iters;
    double single = 0;
    for (int i = 0; i < 200; ++i) {
        const auto u0 = clk::now();
        call();
        cudaStreamSynchronize(s);
        single += std::chrono::duration<double, std::micro>(clk::now() - u0).count();
    }
    Row row{key, issue_us, ms * 1e3 / iters, single / 200, -1.0};
    if (!g_json) std::printf("%-44s issue %6.2f us/call   stream %6.2f us/call   issue->done %6.2f us", name, issue_us, ms * 1e3 / iters, single / 200);
    if (graph) {
        cudaGraph_t g;
        cudaGraphExec_t ge;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        for (int i = 0; i < 100; ++i) call();
        if (cudaStreamEndCapture(s, &g) == cudaSuccess && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) {
            cudaGraphLaunch(ge, s);
            cudaStreamSynchronize(s);
            cudaEventRecord(a, s);
            for (int i = 0; i < 20; ++i) cudaGraphLaunch(ge, s);
            cudaEventRecord(b, s);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            row.graph_us = ms * 1e3 / 2000;
            if (!g_json) std::printf("   graph of 100: %6.2f us/call", ms * 1e3 / 2000);
        } else if (!g_json) {
            std::printf("   graph capture failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    if (!g_json) std::printf("\n");
    g_rows.push_back(row);
}

int main(int argc, char** argv) {
    g_json = argc > 1 && std::strcmp(argv[1], "--json") == 0;  // one JSON line for bench.py (graph figures included)
    const bool graph = g_json || (argc > 1 && std::strcmp(argv[1], "--graph") == 0);
    CK(fheb_init(0));
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    {
        fheb_ntt_plan* plan = nullptr;
        CK(fheb_ntt_plan_create(1024, 132120577ULL, &plan));
        uint64_t *x, *y;
        cudaMalloc(&x, 1024 * 8);
        cudaMalloc(&y, 1024 * 8);
        cudaMemset(x, 1, 1024 * 8);
        measure("c1_forward", "C1 forward N=1024 q=132120577 batch 1", s, [&] { fheb_ntt_forward_batch(plan, x, y, 1, s); }, graph);
        measure("c1_inverse", "C1 inverse N=1024 q=132120577 batch 1", s, [&] { fheb_ntt_inverse_batch(plan, y, x, 1, s); }, graph);
        measure("c1_pair", "C1 forward+inverse pair", s, [&] { fheb_ntt_forward_batch(plan, x, y, 1, s); fheb_ntt_inverse_batch(plan, y, x, 1, s); }, graph);
        fheb_ntt_plan_destroy(plan);
    }
    {
        const uint64_t q[2] = {0xFFFFFFFFFFFFFF43ULL, 1};
        uint64_t consts[5];
        CK(fheb_mlimb_constants(q, 2, consts));
        const size_t n = 65536;
        uint64_t *a, *b, *r;
        cudaMalloc(&a, n * 16);
        cudaMalloc(&b, n * 16);
        cudaMalloc(&r, n * 16);
        cudaMemset(a, 0, n * 16);
        cudaMemset(b, 0, n * 16);
        measure("c3_montmul_n65536", "C3 two-limb montgomery_mul n=65536", s, [&] { fheb_mlimb_montmul_batch(a, b, r, n, 2, q, consts[0], s); }, graph);
    }
    if (g_json) {
        std::printf("{");
        for (size_t i = 0; i < g_rows.size(); ++i)
            std::printf("%s\"%s\": {\"issue_us\": %.3f, \"stream_us\": %.3f, \"issue_to_done_us\": %.3f, \"graph_us\": %.3f}", i ? ", " : "", g_rows[i].key,
                        g_rows[i].issue_us, g_rows[i].stream_us, g_rows[i].done_us, g_rows[i].graph_us);
        std::printf("}\n");
    }
    return 0;
}
