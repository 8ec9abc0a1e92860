// Encrypted-ballot tally: column sums of [count][2][N] ciphertext words modulo q.
//
// The reference folds ballots with PolynomialRing::add_inplace (linear, cpp/src/encryption.cpp:
// 1327-1364) or pairwise (tree, :1366-1458); both reduce every operand first and keep the
// running sum canonical, so the result words are sum_i(x_i mod q) mod q = (sum_i x_i) mod q
// for any grouping.  The kernel therefore accumulates the RAW words in a 128-bit counter per
// column (exact for any count < 2^64) and reduces once.  HBM-bound: every ballot word is read
// exactly once with 16-byte streaming loads, 8 independent loads in flight per thread.
//
// C-ABI entry points here (include/fheb200.h): fheb_tally, fheb_tally_combine,
// fheb_tensor_multiply_batch, fheb_synth_ballots.
#include <cstdlib>
#include <map>

#include "elementwise.hpp"
#include "modarith.cuh"
#include "plan.hpp"
#include "runtime.hpp"

namespace fheb {


// grid = (column chunks, slabs).  Block: 256 threads x 2 columns.  Slab y sums ballots
// [y*per_slab, min(count, (y+1)*per_slab)) and writes canonical partial sums to
// partial[y][width].
constexpr int TALLY_THREADS = 256;
constexpr int TALLY_UNROLL = 8;
constexpr uint32_t TALLY_MAX_PEERS = 16;

// Cross-GPU stage of the sharded tally, fused into the tally kernel (peer memory over NVLink instead of an
// all-gather + a combine kernel): the block that folds a column chunk stores its 512 words into row `rank` of EVERY
// peer's inbox, publishes a per-(rank, chunk) flag with the call's epoch (release, system scope), waits for the same
// flag from every peer in its own inbox (acquire), and sums the `world` rows.  Inboxes are double-buffered by epoch
// parity: a peer can be at most one call ahead (it needs this rank's flag of call k+1 to finish call k+1, and this
// rank sends that only after its call k has completed on its stream).
struct TallyPeerArgs {
    uint64_t* inbox[TALLY_MAX_PEERS];   // peer p's inbox of this epoch's parity: [world][width] words
    unsigned* flags[TALLY_MAX_PEERS];   // peer p's flags: [world][chunks]
    unsigned* status;                   // local: set to 1 when a wait timed out
    uint32_t world, rank, epoch;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t ld_relaxed_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(TALLY_THREADS) tally_kernel(const uint64_t* __restrict__ cts, size_t count,
                                                              size_t per_slab, uint32_t width /* 2N words */,
                                                              uint64_t* partial, const ModQ m, int vec_ok,
                                                              unsigned* done = nullptr, uint64_t* final_out = nullptr,
                                                              const int use_peers = 0, const TallyPeerArgs pa = TallyPeerArgs{}) {
    const uint32_t col = (blockIdx.x * TALLY_THREADS + threadIdx.x) * 2;
    const bool live = col < width;  // (threads past the row's end still take part in the barriers below)
    const size_t first = (size_t)blockIdx.y * per_slab;
    size_t last = first + per_slab;
    if (last > count) last = count;
    uint64_t lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
    const bool pair = (col + 1 < width);
    if (!live) {
    } else if (vec_ok && pair) {
        const ulonglong2* base = reinterpret_cast<const ulonglong2*>(cts + col);
        const size_t row = width / 2;  // ulonglong2 per ballot
        size_t i = first;
        for (; i + TALLY_UNROLL <= last; i += TALLY_UNROLL) {
            ulonglong2 v[TALLY_UNROLL];
#pragma unroll
            for (int u = 0; u < TALLY_UNROLL; ++u) v[u] = __ldcs(base + (i + u) * row);
#pragma unroll
            for (int u = 0; u < TALLY_UNROLL; ++u) {
                acc128(lo0, hi0, v[u].x);
                acc128(lo1, hi1, v[u].y);
            }
        }
        for (; i < last; ++i) {
            const ulonglong2 v = __ldcs(base + i * row);
            acc128(lo0, hi0, v.x);
            acc128(lo1, hi1, v.y);
        }
    } else {
        for (size_t i = first; i < last; ++i) {
            acc128(lo0, hi0, cts[i * width + col]);
            if (pair) acc128(lo1, hi1, cts[i * width + col + 1]);
        }
    }
    if (live) {
        uint64_t* out = partial + (size_t)blockIdx.y * width + col;
        out[0] = fold128(hi0, lo0, m);
        if (pair) out[1] = fold128(hi1, lo1, m);
    }
    if (done == nullptr) return;
    // Second stage without a second launch (a 4-block fold kernel between two 0.3 ms launches costs ~25 us of drain,
    // launch and ramp-up): the LAST block of a column chunk to finish folds that chunk's slab partials.
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(done + blockIdx.x, 1u) == gridDim.y - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (live) {
        lo0 = hi0 = lo1 = hi1 = 0;
        const uint64_t* p = partial + col;  // written by other SMs: read through L2, many loads in flight
        constexpr unsigned FOLD_UNROLL = 16;
        unsigned y = 0;
        if (pair && (width % 2 == 0)) {  // partial rows are 16-byte aligned (stream-ordered allocation, even width, even col)
            const ulonglong2* p2 = reinterpret_cast<const ulonglong2*>(p);
            const size_t row = width / 2;
            for (; y + FOLD_UNROLL <= gridDim.y; y += FOLD_UNROLL) {
                ulonglong2 v[FOLD_UNROLL];
#pragma unroll
                for (unsigned u = 0; u < FOLD_UNROLL; ++u) v[u] = __ldcg(p2 + (size_t)(y + u) * row);
#pragma unroll
                for (unsigned u = 0; u < FOLD_UNROLL; ++u) {
                    acc128(lo0, hi0, v[u].x);
                    acc128(lo1, hi1, v[u].y);
                }
            }
        }
        for (; y < gridDim.y; ++y) {
            acc128(lo0, hi0, __ldcg(p + (size_t)y * width));
            if (pair) acc128(lo1, hi1, __ldcg(p + (size_t)y * width + 1));
        }
        if (!use_peers) {
            final_out[col] = fold128(hi0, lo0, m);
            if (pair) final_out[col + 1] = fold128(hi1, lo1, m);
        }
    }
    if (threadIdx.x == 0) done[blockIdx.x] = 0;  // ready for the next call that draws this counter slot
    if (!use_peers) return;

    // ---- fused exchange + combine over peer memory ------------------------------------------------
    const uint32_t chunks = gridDim.x;
    if (live) {
        const uint64_t v0 = fold128(hi0, lo0, m), v1 = pair ? fold128(hi1, lo1, m) : 0;
        for (uint32_t p = 0; p < pa.world; ++p) {  // NVLink stores into every inbox (the own one included)
            uint64_t* dst = pa.inbox[p] + (size_t)pa.rank * width + col;
            dst[0] = v0;
            if (pair) dst[1] = v1;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < pa.world) {
        st_release_sys(pa.flags[threadIdx.x] + (size_t)pa.rank * chunks + blockIdx.x, pa.epoch);
        const unsigned* mine = pa.flags[pa.rank] + (size_t)threadIdx.x * chunks + blockIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) != pa.epoch) {
            if (clock64() - t0 > (1ll << 35)) {  // ~17 s at 2 GHz: a peer never arrived
                *pa.status = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (live) {
        lo0 = hi0 = lo1 = hi1 = 0;
        const uint64_t* row = pa.inbox[pa.rank] + col;
        for (uint32_t r = 0; r < pa.world; ++r) {
            acc128(lo0, hi0, ld_relaxed_sys(row + (size_t)r * width));
            if (pair) acc128(lo1, hi1, ld_relaxed_sys(row + (size_t)r * width + 1));
        }
        final_out[col] = fold128(hi0, lo0, m);
        if (pair) final_out[col + 1] = fold128(hi1, lo1, m);
    }
}

// Transform-domain tensor product of EncryptionEngine::multiply (cpp/src/encryption.cpp:760-785) for a whole
// batch in one launch: t = [batch][4][N] holds T(a0), T(a1), T(b0), T(b1); out = [batch][3][N] gets
// a0.b0, a0.b1 + a1.b0, a1.b1 (canonical words).
__global__ void __launch_bounds__(256) tensor_pointwise_kernel(const uint64_t* __restrict__ ta, const uint64_t* __restrict__ tb,
                                                               uint64_t* __restrict__ out, size_t batch, uint32_t N, const ModQ m) {
    const size_t total = batch * N;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t ct = i / N, j = i - ct * N;
        const uint64_t a0 = ta[(ct * 2) * N + j], a1 = ta[(ct * 2 + 1) * N + j];
        const uint64_t b0 = tb[(ct * 2) * N + j], b1 = tb[(ct * 2 + 1) * N + j];
        uint64_t* o = out + ct * 3 * N + j;
        o[0] = mulmod(a0, b0, m);
        o[N] = addmod_canon(mulmod(a0, b1, m), mulmod(a1, b0, m), m.q);
        o[2 * (size_t)N] = mulmod(a1, b1, m);
    }
}

// splitmix64 (public-domain constants) - counter-based, reproducible on the CPU
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) synth_ballots_kernel(uint64_t* out, size_t first_word, size_t words, uint64_t seed,
                                                            const ModQ m) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride)
        out[i] = reduce64(splitmix64(seed + first_word + i), m);
}

// "Blocks done" counters of the in-kernel second stage: TALLY_SLOTS sets of TALLY_MAX_CHUNKS counters, all zero between
// calls (the folding block resets its own).  Calls draw sets round-robin, so up to TALLY_SLOTS tallies may be in flight
// on different streams at once.
constexpr unsigned TALLY_SLOTS = 64, TALLY_MAX_CHUNKS = 256;
static unsigned* tally_counters() {
    static std::map<int, unsigned*> pools;  // one pool per device the library has been pointed at
    static std::atomic<unsigned> next{0};
    static std::mutex mu;
    unsigned* pool = nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        const int dev = ctx().device;
        auto it = pools.find(dev);
        if (it == pools.end()) {
            const size_t bytes = (size_t)TALLY_SLOTS * TALLY_MAX_CHUNKS * sizeof(unsigned);
            if (cudaMalloc(&pool, bytes) != cudaSuccess || cudaMemset(pool, 0, bytes) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;  // falls back to the two-launch form
            }
            pools[dev] = pool;
        } else {
            pool = it->second;
        }
    }
    return pool + (size_t)(next.fetch_add(1) % TALLY_SLOTS) * TALLY_MAX_CHUNKS;
}

// Sums `count` rows of `width` words; out = [width].  single_raw: a lone row is copied verbatim
// (EncryptionEngine::batch_add returns the ciphertext untouched for size 1, encryption.cpp:1332-1334).
static int tally_device(const uint64_t* cts, size_t count, uint32_t width, uint64_t q, uint64_t* out, bool single_raw,
                        cudaStream_t s) {
    if (count == 1 && single_raw) {
        FHEB_CUDA(cudaMemcpyAsync(out, cts, (size_t)width * 8, cudaMemcpyDeviceToDevice, s));
        return FHEB_OK;
    }
    const ModQ m = make_modq(q);
    const unsigned chunks = (width + 2 * TALLY_THREADS - 1) / (2 * TALLY_THREADS);
    // enough slabs to put ~4 blocks on every SM, but at least 64 ballots per slab
    static const int bpsm = getenv("FHEB_EXP_TALLY_BPSM") ? atoi(getenv("FHEB_EXP_TALLY_BPSM")) : 4;  // tuning knob
    size_t slabs = ((size_t)ctx().sm_count * bpsm + chunks - 1) / chunks;
    const size_t max_slabs = (count + 63) / 64;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    const size_t per_slab = (count + slabs - 1) / slabs;
    slabs = (count + per_slab - 1) / per_slab;
    const int vec_ok = ((reinterpret_cast<uintptr_t>(cts) & 15u) == 0) && (width % 2 == 0);
    if (slabs == 1) {
        tally_kernel<<<dim3(chunks, 1), TALLY_THREADS, 0, s>>>(cts, count, per_slab, width, out, m, vec_ok);
        FHEB_CHECK_LAUNCH();
        count_launch();
        return FHEB_OK;
    }
    uint64_t* partial = nullptr;
    FHEB_CUDA(cudaMallocAsync(&partial, slabs * (size_t)width * 8, s));
    unsigned* done = chunks <= TALLY_MAX_CHUNKS ? tally_counters() : nullptr;
    if (done) {
        tally_kernel<<<dim3(chunks, (unsigned)slabs), TALLY_THREADS, 0, s>>>(cts, count, per_slab, width, partial, m, vec_ok, done, out);
        FHEB_CHECK_LAUNCH();
        count_launch();
    } else {
        tally_kernel<<<dim3(chunks, (unsigned)slabs), TALLY_THREADS, 0, s>>>(cts, count, per_slab, width, partial, m, vec_ok);
        FHEB_CHECK_LAUNCH();
        tally_kernel<<<dim3(chunks, 1), TALLY_THREADS, 0, s>>>(partial, slabs, slabs, width, out, m, 1);
        FHEB_CHECK_LAUNCH();
        count_launch(2);
    }
    FHEB_CUDA(cudaFreeAsync(partial, s));
    return FHEB_OK;
}

// device rows -> one row (any count, canonical result); used by the wire-format tally (wire.cu)
int tally_rows_device(const uint64_t* rows, size_t count, uint32_t width, uint64_t q, uint64_t* out, cudaStream_t s) {
    return tally_device(rows, count, width, q, out, false, s);
}

// ---- sharded tally with the exchange fused into the kernel (peer memory) --------------------------------
struct TallyPeers {
    uint32_t degree = 0, world = 0, rank = 0, chunks = 0, epoch = 0;
    uint64_t modulus = 0;
    void* local = nullptr;                 // this rank's exported allocation: inbox[2][world][width] | flags[world][chunks] | status
    void* peer[TALLY_MAX_PEERS] = {};      // opened allocations of the peers (peer[rank] == local)
    bool connected = false;
    size_t inbox_bytes() const { return (size_t)world * 2 * degree * 8; }
    size_t flags_offset() const { return 2 * inbox_bytes(); }
    size_t status_offset() const { return flags_offset() + (((size_t)world * chunks * 4 + 15) & ~(size_t)15); }
    size_t total_bytes() const { return status_offset() + 16; }
};

static int tally_peers_run_device(TallyPeers* tp, const uint64_t* cts, size_t count, uint64_t* out, cudaStream_t s) {
    const uint32_t width = 2 * tp->degree;
    const ModQ m = make_modq(tp->modulus);
    const unsigned chunks = tp->chunks;
    size_t slabs = ((size_t)ctx().sm_count * 4 + chunks - 1) / chunks;
    const size_t max_slabs = (count + 63) / 64;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    const size_t per_slab = count ? (count + slabs - 1) / slabs : 1;
    slabs = count ? (count + per_slab - 1) / per_slab : 1;
    unsigned* done = tally_counters();
    FHEB_REQUIRE(done != nullptr, "tally counters unavailable");
    TallyPeerArgs pa{};
    pa.world = tp->world;
    pa.rank = tp->rank;
    pa.epoch = ++tp->epoch;
    const size_t parity = (pa.epoch & 1u) * tp->inbox_bytes();
    for (uint32_t p = 0; p < tp->world; ++p) {
        pa.inbox[p] = reinterpret_cast<uint64_t*>(static_cast<char*>(tp->peer[p]) + parity);
        pa.flags[p] = reinterpret_cast<unsigned*>(static_cast<char*>(tp->peer[p]) + tp->flags_offset());
    }
    pa.status = reinterpret_cast<unsigned*>(static_cast<char*>(tp->local) + tp->status_offset());
    uint64_t* partial = nullptr;
    FHEB_CUDA(cudaMallocAsync(&partial, slabs * (size_t)width * 8, s));
    const int vec_ok = ((reinterpret_cast<uintptr_t>(cts) & 15u) == 0);
    tally_kernel<<<dim3(chunks, (unsigned)slabs), TALLY_THREADS, 0, s>>>(cts, count, per_slab, width, partial, m, vec_ok, done, out, 1, pa);
    FHEB_CHECK_LAUNCH();
    count_launch();
    FHEB_CUDA(cudaFreeAsync(partial, s));
    return FHEB_OK;
}

static int tally_entry(const uint64_t* cts, size_t count, uint32_t degree, uint64_t q, uint64_t* out, bool single_raw,
                       void* stream) {
    FHEB_TRY(ensure_ready());
    // message follows EncryptionEngine::batch_add, cpp/src/encryption.cpp:1328-1330
    FHEB_REQUIRE(count != 0, "Cannot add empty vector of ciphertexts");
    FHEB_REQUIRE(degree > 0 && (degree & (degree - 1)) == 0, "Polynomial degree must be a power of 2");
    FHEB_REQUIRE(q >= 2, "Modulus must be at least 2");
    FHEB_REQUIRE(cts != nullptr && out != nullptr, "ciphertext pointers must not be null");
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t width = 2 * degree;
    if (count > 1 && all_host({cts, out}) && count * (size_t)width * 8 >= (16u << 20)) {
        // host ballots: each chunk is copied in and folded to one partial tally while the next chunk is in
        // flight; the partials are folded at the end (any grouping gives the same words)
        const size_t row = (size_t)width * 8;
        size_t chunk = (8u << 20) / row;
        if (chunk < 1) chunk = 1;
        const size_t nchunks = (count + chunk - 1) / chunk;
        uint64_t* partial = nullptr;
        FHEB_CUDA(cudaMalloc(&partial, nchunks * row));
        int rc = run_host_pipeline(count, {{cts, row, 0, true, false}},
                                   [&](void* const* d, size_t first, size_t n, cudaStream_t ps) {
                                       return tally_device(static_cast<const uint64_t*>(d[0]), n, width, q, partial + (first / chunk) * width, false, ps);
                                   },
                                   chunk);
        uint64_t* dout = nullptr;
        if (rc == FHEB_OK && cudaMalloc(&dout, row) != cudaSuccess) rc = set_error(FHEB_ERR_OUT_OF_MEMORY, "cudaMalloc failed");
        if (rc == FHEB_OK) rc = tally_device(partial, nchunks, width, q, dout, false, s);
        if (rc == FHEB_OK && cudaMemcpyAsync(out, dout, row, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "copy of the tally failed");
        if (rc == FHEB_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "tally failed");
        cudaFree(partial);
        if (dout) cudaFree(dout);
        return rc;
    }
    Staged sin, sout;
    FHEB_TRY(sin.bind(cts, count * (size_t)width * 8, true, false, s));
    FHEB_TRY(sout.bind(out, (size_t)width * 8, false, true, s));
    FHEB_TRY(tally_device(sin.ptr<const uint64_t>(), count, width, q, sout.ptr<uint64_t>(), single_raw, s));
    FHEB_TRY(sin.finish());
    FHEB_TRY(sout.finish());
    return sync_if_staged(s, {&sin, &sout});
}

}  // namespace fheb

using namespace fheb;

extern "C" {

int fheb_tally(const uint64_t* cts, size_t count, uint32_t degree, uint64_t modulus, uint64_t* out, void* stream) {
    return tally_entry(cts, count, degree, modulus, out, true, stream);
}

int fheb_tally_combine(const uint64_t* partials, size_t parts, uint32_t degree, uint64_t modulus, uint64_t* out,
                       void* stream) {
    return tally_entry(partials, parts, degree, modulus, out, true, stream);
}

int fheb_tally_peers_create(uint32_t degree, uint64_t modulus, uint32_t world, uint32_t rank, fheb_tally_peers** out,
                            uint8_t* handle_out) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(out != nullptr && handle_out != nullptr, "out and handle_out must not be null");
    FHEB_REQUIRE(degree > 0 && (degree & (degree - 1)) == 0, "Polynomial degree must be a power of 2");
    FHEB_REQUIRE(modulus >= 2, "Modulus must be at least 2");
    FHEB_REQUIRE(world >= 2 && world <= TALLY_MAX_PEERS && rank < world, "world must be 2..%u and rank below it", TALLY_MAX_PEERS);
    static_assert(sizeof(cudaIpcMemHandle_t) == FHEB_PEER_HANDLE_BYTES, "IPC handle size");
    TallyPeers* tp = new TallyPeers();
    tp->degree = degree;
    tp->modulus = modulus;
    tp->world = world;
    tp->rank = rank;
    tp->chunks = (2 * degree + 2 * TALLY_THREADS - 1) / (2 * TALLY_THREADS);
    if (tp->chunks > TALLY_MAX_CHUNKS) {
        delete tp;
        return set_error(FHEB_ERR_INVALID_PARAMETERS, "degree too large for the fused sharded tally");
    }
    cudaIpcMemHandle_t h;
    if (cudaMalloc(&tp->local, tp->total_bytes()) != cudaSuccess || cudaMemset(tp->local, 0, tp->total_bytes()) != cudaSuccess ||
        cudaIpcGetMemHandle(&h, tp->local) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        if (tp->local) cudaFree(tp->local);
        delete tp;
        return set_error(FHEB_ERR_NATIVE, "peer buffer set-up failed: %s", cudaGetErrorString(e));
    }
    cudaDeviceSynchronize();
    std::memcpy(handle_out, &h, sizeof(h));
    tp->peer[rank] = tp->local;
    *out = reinterpret_cast<fheb_tally_peers*>(tp);
    return FHEB_OK;
}

int fheb_tally_peers_connect(fheb_tally_peers* peers, const uint8_t* handles) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(peers != nullptr && handles != nullptr, "peers and handles must not be null");
    TallyPeers* tp = reinterpret_cast<TallyPeers*>(peers);
    for (uint32_t p = 0; p < tp->world; ++p) {
        if (p == tp->rank || tp->peer[p]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)p * FHEB_PEER_HANDLE_BYTES, sizeof(h));
        const cudaError_t e = cudaIpcOpenMemHandle(&tp->peer[p], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            tp->peer[p] = nullptr;
            return set_error(FHEB_ERR_NATIVE, "opening the inbox of rank %u failed: %s", p, cudaGetErrorString(e));
        }
    }
    tp->connected = true;
    return FHEB_OK;
}

int fheb_tally_peers_run(fheb_tally_peers* peers, const uint64_t* cts, size_t count, uint64_t* out, void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(peers != nullptr, "peers must not be null");
    TallyPeers* tp = reinterpret_cast<TallyPeers*>(peers);
    FHEB_REQUIRE(tp->connected, "fheb_tally_peers_connect has not succeeded on this handle");
    FHEB_REQUIRE(out != nullptr && (count == 0 || cts != nullptr), "ciphertext pointers must not be null");
    FHEB_REQUIRE((count == 0 || is_device_pointer(cts)) && is_device_pointer(out), "the fused sharded tally takes device buffers");
    return tally_peers_run_device(tp, cts, count, out, (cudaStream_t)stream);
}

int fheb_tally_peers_status(const fheb_tally_peers* peers, int* timed_out) {
    FHEB_REQUIRE(peers != nullptr && timed_out != nullptr, "peers and timed_out must not be null");
    const TallyPeers* tp = reinterpret_cast<const TallyPeers*>(peers);
    unsigned v = 0;
    FHEB_CUDA(cudaMemcpy(&v, static_cast<const char*>(tp->local) + tp->status_offset(), 4, cudaMemcpyDeviceToHost));
    *timed_out = (int)v;
    return FHEB_OK;
}

int fheb_tally_peers_destroy(fheb_tally_peers* peers) {
    if (!peers) return FHEB_OK;
    TallyPeers* tp = reinterpret_cast<TallyPeers*>(peers);
    cudaDeviceSynchronize();
    for (uint32_t p = 0; p < tp->world; ++p)
        if (p != tp->rank && tp->peer[p]) cudaIpcCloseMemHandle(tp->peer[p]);
    if (tp->local) cudaFree(tp->local);
    delete tp;
    return FHEB_OK;
}

// ---- streaming accumulator ------------------------------------------------------------------
struct TallyStream {
    uint32_t degree = 0;
    uint64_t modulus = 0;
    uint64_t count = 0;        // ballots folded so far
    uint64_t* d_total = nullptr;   // [2][N]: raw words of the only ballot while count == 1, canonical sums afterwards
    uint64_t* d_part = nullptr;    // [2][N] scratch
};

int fheb_tally_stream_create(uint32_t degree, uint64_t modulus, fheb_tally_stream** out) {
    FHEB_REQUIRE(out != nullptr, "out must not be null");
    *out = nullptr;
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(degree > 0 && (degree & (degree - 1)) == 0, "Polynomial degree must be a power of 2");
    FHEB_REQUIRE(modulus >= 2, "Modulus must be at least 2");
    TallyStream* t = new TallyStream();
    t->degree = degree;
    t->modulus = modulus;
    const size_t bytes = (size_t)2 * degree * 8;
    if (cudaMalloc(&t->d_total, bytes) != cudaSuccess || cudaMalloc(&t->d_part, bytes) != cudaSuccess) {
        fheb_tally_stream_destroy(reinterpret_cast<fheb_tally_stream*>(t));
        return set_error(FHEB_ERR_OUT_OF_MEMORY, "cudaMalloc of the tally accumulator failed");
    }
    *out = reinterpret_cast<fheb_tally_stream*>(t);
    return FHEB_OK;
}

int fheb_tally_stream_add(fheb_tally_stream* ts, const uint64_t* cts, size_t count, void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(ts != nullptr, "tally stream must not be null");
    if (count == 0) return FHEB_OK;
    FHEB_REQUIRE(cts != nullptr, "ciphertext pointer must not be null");
    TallyStream* t = reinterpret_cast<TallyStream*>(ts);
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t width = 2 * t->degree;
    // first ballot ever: it becomes the accumulator as it is (stream_add, :482-484)
    uint64_t* dst = (t->count == 0) ? t->d_total : t->d_part;
    const bool raw_single = (t->count == 0 && count == 1);
    if (is_device_pointer(cts)) {
        FHEB_TRY(tally_device(cts, count, width, t->modulus, dst, raw_single, s));
    } else {  // host chunk: reuse the one-shot entry point's staging / pipelining into a device result
        FHEB_TRY(tally_entry(cts, count, t->degree, t->modulus, dst, raw_single, stream));
    }
    if (t->count != 0) FHEB_TRY(elementwise_device(0 /* add: reduces both inputs first */, t->d_total, t->d_part, 0, t->d_total, width, t->modulus, s));
    t->count += count;
    return FHEB_OK;
}

int fheb_tally_stream_total(const fheb_tally_stream* ts, uint64_t* out, void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(ts != nullptr && out != nullptr, "tally stream and out must not be null");
    const TallyStream* t = reinterpret_cast<const TallyStream*>(ts);
    // message follows EncryptionEngine::batch_add on an empty input, cpp/src/encryption.cpp:1328-1330
    FHEB_REQUIRE(t->count != 0, "Cannot add empty vector of ciphertexts");
    cudaStream_t s = (cudaStream_t)stream;
    FHEB_CUDA(cudaMemcpyAsync(out, t->d_total, (size_t)2 * t->degree * 8, cudaMemcpyDefault, s));
    if (!is_device_pointer(out)) FHEB_CUDA(cudaStreamSynchronize(s));
    return FHEB_OK;
}

uint64_t fheb_tally_stream_count(const fheb_tally_stream* ts) { return ts ? reinterpret_cast<const TallyStream*>(ts)->count : 0; }

int fheb_tally_stream_destroy(fheb_tally_stream* ts) {
    TallyStream* t = reinterpret_cast<TallyStream*>(ts);
    if (!t) return FHEB_OK;
    if (t->d_total) cudaFree(t->d_total);
    if (t->d_part) cudaFree(t->d_part);
    delete t;
    return FHEB_OK;
}

int fheb_tensor_multiply_batch(const fheb_ntt_plan* plan, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out,
                               size_t batch, void* stream) {
    // EncryptionEngine::multiply, cpp/src/encryption.cpp:737-798: T on the four operand
    // polynomials, c0 = a0.b0, c1 = a0.b1 + a1.b0, c2 = a1.b1, T^-1 on the three results.
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(plan != nullptr, "plan must not be null");
    if (batch == 0) return FHEB_OK;
    FHEB_REQUIRE(ct1 != nullptr && ct2 != nullptr && out != nullptr, "ciphertext pointers must not be null");
    const NttPlan* p = reinterpret_cast<const NttPlan*>(plan);
    const size_t N = p->degree;
    cudaStream_t s = (cudaStream_t)stream;
    Staged s1, s2, so;
    FHEB_TRY(s1.bind(ct1, batch * 2 * N * 8, true, false, s));
    FHEB_TRY(s2.bind(ct2, batch * 2 * N * 8, true, false, s));
    FHEB_TRY(so.bind(out, batch * 3 * N * 8, false, true, s));
    uint64_t* work = nullptr;  // T(ct1), T(ct2): [batch][2][N] each
    FHEB_CUDA(cudaMallocAsync(&work, batch * 4 * N * 8, s));
    uint64_t* ta = work;
    uint64_t* tb = work + batch * 2 * N;
    int rc = ntt_forward_device(p, s1.ptr<const uint64_t>(), ta, batch * 2, s);
    if (rc == FHEB_OK) rc = ntt_forward_device(p, s2.ptr<const uint64_t>(), tb, batch * 2, s);
    uint64_t* o = so.ptr<uint64_t>();
    if (rc == FHEB_OK) {
        tensor_pointwise_kernel<<<stream_grid(batch * N, 256, 8), 256, 0, s>>>(ta, tb, o, batch, (uint32_t)N, p->mod);
        if (cudaGetLastError() != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "tensor_pointwise_kernel launch failed");
        count_launch();
    }
    if (rc == FHEB_OK) rc = ntt_inverse_device(p, o, o, batch * 3, s);
    cudaFreeAsync(work, s);
    FHEB_TRY(rc);
    FHEB_TRY(so.finish());
    return sync_if_staged(s, {&s1, &s2, &so});
}

int fheb_synth_ballots(uint64_t* cts_device, size_t first_ballot, size_t count, uint32_t degree, uint64_t modulus,
                       uint64_t seed, void* stream) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(cts_device != nullptr && is_device_pointer(cts_device), "cts_device must be a device pointer");
    FHEB_REQUIRE(modulus >= 2, "Modulus must be at least 2");
    const size_t words = count * 2 * (size_t)degree;
    if (words == 0) return FHEB_OK;
    synth_ballots_kernel<<<stream_grid(words, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        cts_device, first_ballot * 2 * (size_t)degree, words, seed, make_modq(modulus));
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

}  // extern "C"
