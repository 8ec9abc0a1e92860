// Prints the shared-window address of a kernel's dynamic shared memory (decides whether swizzled slot addresses can be
// formed with XOR alone, i.e. whether the base has any of the low 17 bits set).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned* out) {
    extern __shared__ __align__(16) unsigned long long smem[];
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned)__cvta_generic_to_shared(smem);
}
int main() {
    unsigned* d; unsigned h[2];
    cudaMalloc(&d, 8);
    for (int kb : {16, 64, 128, 200, 227}) {
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
        probe<<<2, 32, kb * 1024>>>(d);
        cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
        printf("dynamic smem %3d KB: base 0x%x / 0x%x (%s)\n", kb, h[0], h[1], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
