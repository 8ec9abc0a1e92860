// Blind-rotation kernel instantiations for GLWE dimension k = 3 (see boot_kernel.cuh).
#include "boot_kernel.cuh"

namespace fheb {
int boot_launch_k3(uint32_t logn, bool dp, const BootLaunch& a, cudaStream_t stream) {
    return boot_launch_kp1<4>(logn, dp, a, stream);
}
}  // namespace fheb
