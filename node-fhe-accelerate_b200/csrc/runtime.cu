// Library context, error reporting, buffer helpers and host-buffer staging.
// C-ABI entry points here: fheb_init, fheb_shutdown, fheb_version, fheb_last_error,
// fheb_device_info_get, fheb_device_alloc/free, fheb_host_alloc/free, fheb_copy,
// fheb_synchronize, fheb_launch_count (declared in include/fheb200.h).
#include "runtime.hpp"

namespace fheb {

static thread_local std::string t_error;
static std::mutex g_ctx_mutex;
static Context g_ctx;
std::atomic<uint64_t> g_launches{0};

Context& ctx() { return g_ctx; }

int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
    return code;
}

static int init_locked(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(FHEB_ERR_HARDWARE_UNAVAILABLE,
                         "no CUDA device available (%s); this backend has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) return set_error(FHEB_ERR_INVALID_PARAMETERS, "device %d out of range (%d devices)", device, count);
    FHEB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FHEB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        return set_error(FHEB_ERR_HARDWARE_UNAVAILABLE,
                         "device %d (%s) is sm_%d%d; this library ships sm_100a code only", device, prop.name,
                         prop.major, prop.minor);
    }
    if (g_ctx.ready && g_ctx.device != device) {
        // re-target: drop the old staging streams
        cudaStreamDestroy(g_ctx.copy_in);
        cudaStreamDestroy(g_ctx.copy_out);
        cudaStreamDestroy(g_ctx.work);
        g_ctx.ready = false;
    }
    if (!g_ctx.ready) {
        g_ctx.device = device;
        g_ctx.prop = prop;
        g_ctx.sm_count = prop.multiProcessorCount;
        FHEB_CUDA(cudaStreamCreateWithFlags(&g_ctx.copy_in, cudaStreamNonBlocking));
        FHEB_CUDA(cudaStreamCreateWithFlags(&g_ctx.copy_out, cudaStreamNonBlocking));
        FHEB_CUDA(cudaStreamCreateWithFlags(&g_ctx.work, cudaStreamNonBlocking));
        g_ctx.ready = true;
    }
    return FHEB_OK;
}

int ensure_ready() {
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    if (g_ctx.ready) {
        // Kernels must run on the device the context was made for.
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != g_ctx.device) return init_locked(cur);
        return FHEB_OK;
    }
    return init_locked(-1);
}

bool is_device_pointer(const void* p) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

Staged::~Staged() {
    if (owned_ && dev_) cudaFreeAsync(dev_, stream_);
}

int Staged::bind(const void* user, size_t bytes, bool copy_in, bool copy_out, cudaStream_t stream) {
    user_ = user;
    bytes_ = bytes;
    stream_ = stream;
    copy_out_ = copy_out;
    if (bytes == 0 || user == nullptr) {
        dev_ = const_cast<void*>(user);
        return FHEB_OK;
    }
    if (is_device_pointer(user)) {
        dev_ = const_cast<void*>(user);
        owned_ = false;
        return FHEB_OK;
    }
    FHEB_CUDA(cudaMallocAsync(&dev_, bytes, stream));
    owned_ = true;
    if (copy_in) FHEB_CUDA(cudaMemcpyAsync(dev_, user, bytes, cudaMemcpyHostToDevice, stream));
    return FHEB_OK;
}

int Staged::bind_alias(Staged& other, bool copy_out) {
    user_ = other.user_;
    bytes_ = other.bytes_;
    stream_ = other.stream_;
    dev_ = other.dev_;
    owned_ = false;
    copy_out_ = false;
    if (copy_out) other.copy_out_ = true;
    return FHEB_OK;
}

int Staged::finish() {
    if (owned_ && copy_out_ && bytes_)
        FHEB_CUDA(cudaMemcpyAsync(const_cast<void*>(user_), dev_, bytes_, cudaMemcpyDeviceToHost, stream_));
    return FHEB_OK;
}

int sync_if_staged(cudaStream_t stream, std::initializer_list<const Staged*> bufs) {
    bool any = false;
    for (const Staged* b : bufs) any = any || b->staged();
    if (any) FHEB_CUDA(cudaStreamSynchronize(stream));
    return FHEB_OK;
}

}  // namespace fheb

using namespace fheb;

extern "C" {

int fheb_init(int device) {
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    return init_locked(device);
}

int fheb_shutdown(void) {
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    if (g_ctx.ready) {
        cudaStreamDestroy(g_ctx.copy_in);
        cudaStreamDestroy(g_ctx.copy_out);
        cudaStreamDestroy(g_ctx.work);
        g_ctx = Context{};
    }
    return FHEB_OK;
}

const char* fheb_version(void) { return "0.1.0-b200"; }

const char* fheb_last_error(void) { return t_error.c_str(); }

int fheb_device_info_get(fheb_device_info* out) {
    FHEB_REQUIRE(out != nullptr, "out must not be null");
    FHEB_TRY(ensure_ready());
    std::memset(out, 0, sizeof(*out));
    const cudaDeviceProp& p = ctx().prop;
    out->has_cuda = 1;
    out->cc_major = p.major;
    out->cc_minor = p.minor;
    out->sm_count = p.multiProcessorCount;
    out->device_memory_bytes = p.totalGlobalMem;
    out->l2_bytes = (uint64_t)p.l2CacheSize;
    out->smem_per_block_optin = p.sharedMemPerBlockOptin;
    std::strncpy(out->name, p.name, sizeof(out->name) - 1);
    return FHEB_OK;
}

int fheb_device_alloc(void** out, size_t bytes) {
    FHEB_REQUIRE(out != nullptr, "out must not be null");
    FHEB_TRY(ensure_ready());
    FHEB_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return FHEB_OK;
}

int fheb_device_free(void* p) {
    if (p) FHEB_CUDA(cudaFree(p));
    return FHEB_OK;
}

int fheb_host_alloc(void** out, size_t bytes) {
    FHEB_REQUIRE(out != nullptr, "out must not be null");
    FHEB_TRY(ensure_ready());
    FHEB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return FHEB_OK;
}

int fheb_host_free(void* p) {
    if (p) FHEB_CUDA(cudaFreeHost(p));
    return FHEB_OK;
}

int fheb_copy(void* dst, const void* src, size_t bytes, void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    if (!is_device_pointer(dst) || !is_device_pointer(src)) FHEB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return FHEB_OK;
}

int fheb_synchronize(void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return FHEB_OK;
}

uint64_t fheb_launch_count(int reset) {
    uint64_t v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

}  // extern "C"
