// Library context, error reporting, buffer helpers and host-buffer staging.
// C-ABI entry points here: fheb_init, fheb_shutdown, fheb_version, fheb_last_error,
// fheb_device_info_get, fheb_device_alloc/free, fheb_host_alloc/free, fheb_copy,
// fheb_synchronize, fheb_launch_count (declared in include/fheb200.h).
#include "runtime.hpp"

#include <sched.h>

#include <cctype>
#include <thread>

namespace fheb {

static thread_local std::string t_error;
// One context per device (a process may drive several GPUs, from one thread that switches devices or from one thread
// per device): looked up by the calling thread's CURRENT device, created on first use, never re-targeted.
constexpr int MAX_DEVICES = 64;
static std::mutex g_ctx_mutex;
static Context g_ctx[MAX_DEVICES];
static std::atomic<bool> g_ctx_ready[MAX_DEVICES];
std::atomic<uint64_t> g_launches{0};

static int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) {
        cudaGetLastError();
        d = 0;
    }
    return (d >= 0 && d < MAX_DEVICES) ? d : 0;
}

Context& ctx() { return g_ctx[current_device()]; }
Context& ctx_of(int device) { return g_ctx[(device >= 0 && device < MAX_DEVICES) ? device : 0]; }

int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
    return code;
}

// makes `device` (or, when negative, the current device) current and creates its context if it has none
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
    if (device >= count || device >= MAX_DEVICES)
        return set_error(FHEB_ERR_INVALID_PARAMETERS, "device %d out of range (%d devices)", device, count);
    FHEB_CUDA(cudaSetDevice(device));
    Context& c = g_ctx[device];
    if (c.ready) return FHEB_OK;
    cudaDeviceProp prop;
    FHEB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        return set_error(FHEB_ERR_HARDWARE_UNAVAILABLE,
                         "device %d (%s) is sm_%d%d; this library ships sm_100a code only", device, prop.name,
                         prop.major, prop.minor);
    }
    c.device = device;
    c.prop = prop;
    c.sm_count = prop.multiProcessorCount;
    FHEB_CUDA(cudaStreamCreateWithFlags(&c.copy_in, cudaStreamNonBlocking));
    FHEB_CUDA(cudaStreamCreateWithFlags(&c.copy_out, cudaStreamNonBlocking));
    FHEB_CUDA(cudaStreamCreateWithFlags(&c.work, cudaStreamNonBlocking));
    for (auto& ps : c.pipe) FHEB_CUDA(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    {   // staging buffers come from the stream-ordered allocator on every call: keep its memory cached
        // across synchronisations instead of returning it to the driver each time
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    c.ready = true;
    g_ctx_ready[device].store(true, std::memory_order_release);
    return FHEB_OK;
}

int ensure_ready() {
    const int d = current_device();
    if (g_ctx_ready[d].load(std::memory_order_acquire)) return FHEB_OK;  // kernels run on the calling thread's current device
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
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

// ---- host placement: the CPUs and memory next to a GPU ----------------------------------------
// A pinned buffer that lives on the other socket crosses the inter-socket link on every copy.  The CPUs local to a GPU
// come from sysfs (/sys/bus/pci/devices/<bdf>/local_cpulist); binding the calling thread to them while cudaHostAlloc
// faults the pages in places the buffer on the GPU's NUMA node (first touch).  No-op where sysfs has no answer.
static bool device_local_cpus(int device, cpu_set_t* set) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    for (char* c = bdf; *c; ++c) *c = (char)tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bdf + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char line[4096] = {0};
    const bool got = fgets(line, sizeof(line), f) != nullptr;
    fclose(f);
    if (!got) return false;
    CPU_ZERO(set);
    int any = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int lo = 0, hi = 0;
        const int k = sscanf(tok, "%d-%d", &lo, &hi);
        if (k < 1) continue;
        if (k == 1) hi = lo;
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c) {
            CPU_SET(c, set);
            ++any;
        }
    }
    return any > 0;
}

class LocalCpuScope {  // binds the calling thread to the CPUs next to `device` for its lifetime
   public:
    explicit LocalCpuScope(int device) {
        cpu_set_t want;
        if (getenv("FHEB_NO_NUMA_BIND") == nullptr && device_local_cpus(device, &want) &&
            sched_getaffinity(0, sizeof(old_), &old_) == 0) {
            cpu_set_t both;
            CPU_AND(&both, &want, &old_);  // never leave the set the process was given (containers, taskset)
            if (CPU_COUNT(&both) > 0 && sched_setaffinity(0, sizeof(both), &both) == 0) bound_ = true;
        }
    }
    ~LocalCpuScope() {
        if (bound_) sched_setaffinity(0, sizeof(old_), &old_);
    }

   private:
    cpu_set_t old_;
    bool bound_ = false;
};

static std::mutex g_devices_mutex;
static std::vector<int> g_devices;

std::vector<int> device_list() {
    std::lock_guard<std::mutex> lock(g_devices_mutex);
    return g_devices;
}

bool spread_over_devices(size_t items, size_t bytes) {
    size_t n;
    {
        std::lock_guard<std::mutex> lock(g_devices_mutex);
        n = g_devices.size();
    }
    return n > 1 && items >= 2 * n && bytes >= ((size_t)4 << 20) * n;  // a few MB per device at least: below that one GPU wins
}

int run_on_devices(size_t items, const DeviceFn& fn) {
    const std::vector<int> devs = device_list();
    const size_t nd = devs.empty() ? 1 : devs.size();
    std::vector<int> rcs(nd, FHEB_OK);
    std::vector<std::string> msgs(nd);
    std::vector<std::thread> threads;
    const size_t base = items / nd, rem = items % nd;
    size_t first = 0;
    for (size_t d = 0; d < nd; ++d) {
        const size_t n = base + (d < rem ? 1 : 0);
        const size_t f = first;
        first += n;
        if (n == 0) continue;
        threads.emplace_back([&, d, f, n] {
            int rc = FHEB_OK;
            LocalCpuScope near_gpu(devs.empty() ? current_device() : devs[d]);  // the thread that feeds a GPU runs on the CPUs next to it
            if (!devs.empty() && cudaSetDevice(devs[d]) != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "cudaSetDevice(%d) failed", devs[d]);
            if (rc == FHEB_OK) rc = ensure_ready();
            if (rc == FHEB_OK) rc = fn(devs.empty() ? ctx().device : devs[d], f, n);
            rcs[d] = rc;
            if (rc != FHEB_OK) msgs[d] = t_error;  // error text is thread-local: hand it to the caller's thread
        });
    }
    for (auto& t : threads) t.join();
    for (size_t d = 0; d < nd; ++d)
        if (rcs[d] != FHEB_OK) return set_error(rcs[d], "device %d: %s", devs.empty() ? -1 : devs[d], msgs[d].c_str());
    return FHEB_OK;
}

bool all_host(std::initializer_list<const void*> ptrs) {
    for (const void* p : ptrs)
        if (p != nullptr && is_device_pointer(p)) return false;
    return true;
}

int run_host_pipeline(size_t items, std::vector<PipeArg> args, const PipeFn& fn, size_t chunk_items) {
    constexpr int SLOTS = Context::PIPE_SLOTS;
    Context& c = ctx();
    const size_t na = args.size();
    std::vector<size_t> rep(na);  // first argument with the same host pointer owns the device buffer
    size_t max_stride = 1;
    for (size_t i = 0; i < na; ++i) {
        rep[i] = i;
        for (size_t j = 0; j < i; ++j)
            if (args[j].host == args[i].host && args[j].stride == args[i].stride) {
                rep[i] = j;
                args[j].in = args[j].in || args[i].in;
                args[j].out = args[j].out || args[i].out;
                break;
            }
        if (args[i].stride > max_stride) max_stride = args[i].stride;
    }
    // ~16 MB per buffer per chunk: long enough for PCIe efficiency, short enough to pipeline
    static const size_t chunk_bytes = [] {  // FHEB_PIPE_CHUNK_MB: tuning knob
        const char* e = getenv("FHEB_PIPE_CHUNK_MB");
        const long mb = e ? atol(e) : 0;
        return (size_t)(mb > 0 ? mb : 16) << 20;  // 16 MB measured best on PCIe Gen5 (tools/prof_e2e.py)
    }();
    size_t chunk = chunk_items ? chunk_items : chunk_bytes / max_stride;
    if (chunk < 1) chunk = 1;
    if (chunk > items) chunk = items;
    const std::vector<size_t> sizes = pipeline_chunk_sizes(items, chunk, chunk_items == 0);
    const size_t nchunks = sizes.size();
    const int slots = (int)(nchunks < (size_t)SLOTS ? nchunks : (size_t)SLOTS);

    std::vector<void*> shared(na, nullptr);
    std::vector<std::vector<void*>> dev(slots, std::vector<void*>(na, nullptr));
    cudaEvent_t shared_ready = nullptr;
    int rc = FHEB_OK;
    auto fail = [&](cudaError_t e, const char* what) {
        if (rc == FHEB_OK) rc = set_error(e == cudaErrorMemoryAllocation ? FHEB_ERR_OUT_OF_MEMORY : FHEB_ERR_NATIVE, "%s failed: %s", what, cudaGetErrorString(e));
    };
    bool any_shared = false;
    for (size_t i = 0; i < na && rc == FHEB_OK; ++i) {
        if (args[i].stride != 0 || rep[i] != i || args[i].host == nullptr) continue;
        cudaError_t e = cudaMallocAsync(&shared[i], args[i].total ? args[i].total : 1, c.pipe[0]);
        if (e != cudaSuccess) { fail(e, "cudaMallocAsync"); break; }
        if (args[i].in) {
            e = cudaMemcpyAsync(shared[i], args[i].host, args[i].total, cudaMemcpyHostToDevice, c.pipe[0]);
            if (e != cudaSuccess) fail(e, "cudaMemcpyAsync");
        }
        any_shared = true;
    }
    if (rc == FHEB_OK && any_shared) {
        cudaError_t e = cudaEventCreateWithFlags(&shared_ready, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(shared_ready, c.pipe[0]);
        for (int sl = 1; sl < slots && e == cudaSuccess; ++sl) e = cudaStreamWaitEvent(c.pipe[sl], shared_ready, 0);
        if (e != cudaSuccess) fail(e, "cudaEventRecord");
    }
    for (int sl = 0; sl < slots && rc == FHEB_OK; ++sl)
        for (size_t i = 0; i < na && rc == FHEB_OK; ++i) {
            if (args[i].stride == 0 || rep[i] != i || args[i].host == nullptr) continue;
            cudaError_t e = cudaMallocAsync(&dev[sl][i], chunk * args[i].stride, c.pipe[sl]);
            if (e != cudaSuccess) fail(e, "cudaMallocAsync");
        }
    std::vector<void*> ptrs(na);
    size_t first = 0;
    for (size_t k = 0; k < nchunks && rc == FHEB_OK; first += sizes[k], ++k) {
        const int sl = (int)(k % slots);
        cudaStream_t s = c.pipe[sl];
        const size_t n = sizes[k];
        for (size_t i = 0; i < na; ++i) ptrs[i] = args[i].host == nullptr ? nullptr : (args[i].stride ? dev[sl][rep[i]] : shared[rep[i]]);
        for (size_t i = 0; i < na && rc == FHEB_OK; ++i) {
            if (args[i].stride == 0 || rep[i] != i || !args[i].in || args[i].host == nullptr) continue;
            cudaError_t e = cudaMemcpyAsync(dev[sl][i], static_cast<const char*>(args[i].host) + first * args[i].stride,
                                            n * args[i].stride, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) fail(e, "cudaMemcpyAsync");
        }
        if (rc == FHEB_OK) rc = fn(ptrs.data(), first, n, s);
        for (size_t i = 0; i < na && rc == FHEB_OK; ++i) {
            if (args[i].stride == 0 || rep[i] != i || !args[i].out || args[i].host == nullptr) continue;
            cudaError_t e = cudaMemcpyAsync(const_cast<char*>(static_cast<const char*>(args[i].host)) + first * args[i].stride,
                                            dev[sl][i], n * args[i].stride, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) fail(e, "cudaMemcpyAsync");
        }
    }
    // shared outputs (none today) would be copied here; release and drain
    for (int sl = 0; sl < slots; ++sl) {
        for (size_t i = 0; i < na; ++i)
            if (dev[sl][i]) cudaFreeAsync(dev[sl][i], c.pipe[sl]);
    }
    for (int sl = 0; sl < slots; ++sl) {
        cudaError_t e = cudaStreamSynchronize(c.pipe[sl]);
        if (e != cudaSuccess) fail(e, "cudaStreamSynchronize");
    }
    for (size_t i = 0; i < na; ++i)
        if (shared[i]) cudaFree(shared[i]);
    if (shared_ready) cudaEventDestroy(shared_ready);
    return rc;
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
    int cur = -1;
    cudaGetDevice(&cur);
    for (int d = 0; d < MAX_DEVICES; ++d) {
        Context& c = g_ctx[d];
        if (!c.ready) continue;
        g_ctx_ready[d].store(false, std::memory_order_release);
        cudaSetDevice(d);
        cudaStreamDestroy(c.copy_in);
        cudaStreamDestroy(c.copy_out);
        cudaStreamDestroy(c.work);
        for (auto& ps : c.pipe) cudaStreamDestroy(ps);
        c = Context{};
    }
    if (cur >= 0) cudaSetDevice(cur);
    cudaGetLastError();
    return FHEB_OK;
}

int fheb_device_count(void) {
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int usable = 0;
    for (int d = 0; d < visible; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++usable;
    }
    cudaGetLastError();
    return usable;
}

int fheb_set_devices(const int* devices, int count) {
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) {
        cudaGetLastError();
        return set_error(FHEB_ERR_HARDWARE_UNAVAILABLE, "no CUDA device available; this backend has no CPU fallback");
    }
    std::vector<int> list;
    if (count < 0) count = visible;  // all visible devices
    for (int k = 0; k < count; ++k) {
        const int d = devices ? devices[k] : k;
        FHEB_REQUIRE(d >= 0 && d < visible && d < MAX_DEVICES, "device %d out of range (%d devices)", d, visible);
        for (int seen : list) FHEB_REQUIRE(seen != d, "device %d listed twice", d);
        list.push_back(d);
    }
    int cur = -1;
    cudaGetDevice(&cur);
    {
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        for (int d : list) {
            const int rc = init_locked(d);  // every device must be an sm_100 part; creates its context
            if (rc != FHEB_OK) {
                if (cur >= 0) cudaSetDevice(cur);
                return rc;
            }
        }
    }
    if (cur >= 0) cudaSetDevice(cur);
    std::lock_guard<std::mutex> lock(g_devices_mutex);
    g_devices = list;
    return FHEB_OK;
}

int fheb_get_devices(int* out, int capacity) {
    const std::vector<int> list = device_list();
    for (int k = 0; k < (int)list.size() && k < capacity; ++k)
        if (out) out[k] = list[k];
    return (int)list.size();
}

const char* fheb_version(void) { return "0.2.0-b200"; }

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
    LocalCpuScope near_gpu(current_device());  // pages are faulted in by this thread: place them on the current GPU's NUMA node
    FHEB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
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

// ---- scalar ModularArithmetic (the addon's existing surface; host-side, see include/fheb200.h) ------
struct ScalarModArith {
    uint64_t modulus, r2_mod_q, q_inv;
};
typedef unsigned __int128 u128_t;

static uint64_t scalar_mod_inverse(uint64_t a, uint64_t m) {  // cpp/src/modular_arithmetic.cpp:8-31, AArch64 division
    if (m == 0) return 0;
    int64_t m0 = (int64_t)m, x0 = 0, x1 = 1;
    if (m == 1) return 0;
    while (a > 1) {
        const int64_t quo = (m != 0) ? (int64_t)(a / m) : 0;
        int64_t t = (int64_t)m;
        m = (m != 0) ? (a % m) : a;
        a = (uint64_t)t;
        t = x0;
        x0 = x1 - quo * x0;
        x1 = t;
    }
    if (x1 < 0) x1 += m0;
    return (uint64_t)x1;
}

static uint64_t scalar_montgomery_reduce(const ScalarModArith* c, uint64_t hi, uint64_t lo) {  // :84-110
    const uint64_t m = lo * c->q_inv;
    const u128_t mq = (u128_t)m * c->modulus;
    const u128_t sum = ((u128_t)hi << 64) + lo + mq;  // wraps modulo 2^128 exactly as the reference's sum does
    uint64_t t = (uint64_t)(sum >> 64);
    if (t >= c->modulus) t -= c->modulus;
    return t;
}

int fheb_modarith_create(uint64_t modulus, fheb_modarith** out) {
    FHEB_REQUIRE(out != nullptr, "out must not be null");
    *out = nullptr;
    // message follows MontgomeryConstants::MontgomeryConstants, cpp/src/modular_arithmetic.cpp:53-55
    FHEB_REQUIRE(modulus != 0 && (modulus & 1) != 0, "Modulus must be odd and non-zero for Montgomery arithmetic");
    ScalarModArith* c = new ScalarModArith();
    c->modulus = modulus;
    const uint64_t r_mod_q = (uint64_t)((((u128_t)1) << 64) % modulus);
    c->r2_mod_q = (uint64_t)(((u128_t)r_mod_q * r_mod_q) % modulus);
    c->q_inv = (~scalar_mod_inverse(modulus, UINT64_MAX)) + 1;  // :69-70
    *out = reinterpret_cast<fheb_modarith*>(c);
    return FHEB_OK;
}
int fheb_modarith_destroy(fheb_modarith* m) {
    delete reinterpret_cast<ScalarModArith*>(m);
    return FHEB_OK;
}
uint64_t fheb_modarith_montgomery_mul(const fheb_modarith* m, uint64_t a, uint64_t b) {  // :112-120
    const u128_t p = (u128_t)a * b;
    return scalar_montgomery_reduce(reinterpret_cast<const ScalarModArith*>(m), (uint64_t)(p >> 64), (uint64_t)p);
}
uint64_t fheb_modarith_mod_add(const fheb_modarith* m, uint64_t a, uint64_t b) {  // :122-136
    const uint64_t q = reinterpret_cast<const ScalarModArith*>(m)->modulus;
    a %= q;
    b %= q;
    uint64_t s = a + b;
    if (s < a || s >= q) s -= q;
    return s;
}
uint64_t fheb_modarith_mod_sub(const fheb_modarith* m, uint64_t a, uint64_t b) {  // :138-153
    const uint64_t q = reinterpret_cast<const ScalarModArith*>(m)->modulus;
    a %= q;
    b %= q;
    return a >= b ? a - b : q - (b - a);
}
uint64_t fheb_modarith_to_montgomery(const fheb_modarith* m, uint64_t a) {  // :155-159
    return fheb_modarith_montgomery_mul(m, a, reinterpret_cast<const ScalarModArith*>(m)->r2_mod_q);
}
uint64_t fheb_modarith_from_montgomery(const fheb_modarith* m, uint64_t a) {  // :161-165
    return scalar_montgomery_reduce(reinterpret_cast<const ScalarModArith*>(m), 0, a);
}
uint64_t fheb_modarith_get_modulus(const fheb_modarith* m) { return reinterpret_cast<const ScalarModArith*>(m)->modulus; }

uint64_t fheb_launch_count(int reset) {
    uint64_t v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

}  // extern "C"
