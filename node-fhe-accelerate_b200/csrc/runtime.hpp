// Host runtime shared by the C-ABI translation units: status/error plumbing, the device
// context, pointer classification and host-buffer staging.  No CPU compute path lives
// here: without a usable sm_100 device every entry point fails.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <functional>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fheb200.h"

namespace fheb {

struct Context {
    bool ready = false;
    int device = -1;
    cudaDeviceProp prop{};
    int sm_count = 0;
    cudaStream_t copy_in = nullptr;   // staging streams for host buffers
    cudaStream_t copy_out = nullptr;
    cudaStream_t work = nullptr;
    static constexpr int PIPE_SLOTS = 3;
    cudaStream_t pipe[PIPE_SLOTS] = {nullptr, nullptr, nullptr};  // host-buffer pipeline: copy-in / kernels / copy-out per chunk
};

Context& ctx();            // context of the calling thread's current device (valid after ensure_ready())
Context& ctx_of(int device);
int set_error(int code, const char* fmt, ...);
int ensure_ready();  // lazily initialises the current device's context; FHEB_ERR_HARDWARE_UNAVAILABLE if no sm_100 GPU
extern std::atomic<uint64_t> g_launches;

inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define FHEB_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::fheb::set_error(_e == cudaErrorMemoryAllocation ? FHEB_ERR_OUT_OF_MEMORY : FHEB_ERR_NATIVE, \
                                     "%s failed: %s", #expr, cudaGetErrorString(_e));                \
    } while (0)

#define FHEB_CHECK_LAUNCH()                                                                          \
    do {                                                                                             \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e != cudaSuccess) return ::fheb::set_error(FHEB_ERR_NATIVE, "kernel launch failed: %s", cudaGetErrorString(_e)); \
    } while (0)

#define FHEB_REQUIRE(cond, ...)                                                                      \
    do {                                                                                             \
        if (!(cond)) return ::fheb::set_error(FHEB_ERR_INVALID_PARAMETERS, __VA_ARGS__);             \
    } while (0)

#define FHEB_TRY(expr)                 \
    do {                               \
        int _rc = (expr);              \
        if (_rc != FHEB_OK) return _rc; \
    } while (0)

bool is_device_pointer(const void* p);

// A view of caller memory on the device.  Device pointers are used in place; host pointers
// get a device staging buffer, filled on `stream` when `copy_in` is set and written back by
// finish().  finish() synchronises the stream only when something was staged from the host.
class Staged {
   public:
    Staged() = default;
    ~Staged();
    Staged(const Staged&) = delete;
    Staged& operator=(const Staged&) = delete;
    int bind(const void* user, size_t bytes, bool copy_in, bool copy_out, cudaStream_t stream);
    int bind_alias(Staged& other, bool copy_out);  // same user pointer as `other`: share its device buffer
    template <class T>
    T* ptr() const { return reinterpret_cast<T*>(dev_); }
    int finish();  // copy back (if needed); caller syncs once at the end via sync_if_staged
    bool staged() const { return owned_; }
    const void* user() const { return user_; }

   private:
    void* dev_ = nullptr;
    const void* user_ = nullptr;
    size_t bytes_ = 0;
    bool owned_ = false;
    bool copy_out_ = false;
    cudaStream_t stream_ = nullptr;
};

int sync_if_staged(cudaStream_t stream, std::initializer_list<const Staged*> bufs);

// ---- host-buffer pipeline -------------------------------------------------------------------
// When every buffer of a call lives in host memory the batch is cut into chunks and each chunk's
// host->device copy, kernels and device->host copy run on one of PIPE_SLOTS streams, so the two
// PCIe directions and the SMs work concurrently (instead of copy-all, compute, copy-all).
struct PipeArg {
    const void* host;   // caller pointer; arguments with the same pointer share one device buffer
    size_t stride;      // bytes per item; 0 = one block of `total` bytes shared by all items (staged once)
    size_t total;       // bytes when stride == 0
    bool in, out;       // copied to the device before / back to the host after the chunk's kernels
};
using PipeFn = std::function<int(void* const* dev, size_t first_item, size_t n_items, cudaStream_t s)>;
// Chunk schedule of the pipeline (pure host logic, unit-tested on the CPU): full-size chunks in the middle,
// geometrically smaller ones (chunk/8, chunk/4, chunk/2) mirrored at both ends when `ramp` is set and there is room,
// so that the time before the first kernel can start (one copy-in) and after the last one ends (one copy-out) is
// that of a small chunk.  The sizes are positive, none exceeds `chunk`, and they sum to `items`.
inline std::vector<size_t> pipeline_chunk_sizes(size_t items, size_t chunk, bool ramp) {
    std::vector<size_t> sizes;
    if (items == 0) return sizes;
    if (chunk < 1) chunk = 1;
    if (chunk > items) chunk = items;
    std::vector<size_t> head;
    size_t used = 0;
    if (ramp && chunk >= 16 && items > 2 * chunk)
        for (size_t c = chunk / 8; c < chunk && used + 2 * c + chunk <= items; c *= 2) {
            head.push_back(c);
            used += 2 * c;  // mirrored at the tail
        }
    sizes = head;
    size_t left = items - used;
    while (left > 0) {
        const size_t c = left < chunk ? left : chunk;
        sizes.push_back(c);
        left -= c;
    }
    for (size_t i = head.size(); i-- > 0;) sizes.push_back(head[i]);
    return sizes;
}

// ---- several GPUs driven by one process -------------------------------------------------------
// fheb_set_devices() names the GPUs that HOST-buffer batches are spread over (the reference's addon is one process,
// src/native/lib.rs:23-133; every unit of the path is independent).  Each device gets a contiguous share
// and its own host thread running the chunked copy/compute/copy pipeline on its own PCIe link; plans and keys are
// replicated on a device the first time it is used.  Device buffers always run where they live, on the current device.
std::vector<int> device_list();                        // empty or one entry: no spreading
bool spread_over_devices(size_t items, size_t bytes);  // more than one device configured and the batch is worth splitting
using DeviceFn = std::function<int(int device, size_t first, size_t n)>;
int run_on_devices(size_t items, const DeviceFn& fn);  // fn runs with `device` current and its context ready; first error wins

bool all_host(std::initializer_list<const void*> ptrs);  // true when no non-null pointer is device memory
int run_host_pipeline(size_t items, std::vector<PipeArg> args, const PipeFn& fn, size_t chunk_items = 0 /* 0: ~16 MB per buffer */);

}  // namespace fheb
