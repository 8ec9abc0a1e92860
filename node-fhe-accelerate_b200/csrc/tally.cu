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
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>

#include "tensor_fused.hpp"
#include "elementwise.hpp"
#include "modarith.cuh"
#include "plan.hpp"
#include "runtime.hpp"

namespace fheb {


// Block: 256 threads x 2 columns = one 512-column chunk of the row.  Two shapes of the same kernel:
//  * slab form (slot.lo == nullptr): grid = (chunks, slabs); slab y sums ballots [y*per_slab, (y+1)*per_slab) and
//    writes canonical partial sums to partial[y][width] (small inputs, very wide rows, and the second stage of those);
//  * single-launch form: grid = chunks * workers persistent blocks.  The blocks of a column chunk draw work ITEMS
//    (runs of `item` ballots) from a per-chunk counter, so every SM stays busy until the input is exhausted (a static
//    split ends with a tail: 2.1 GB per rank ran at 6.7 TB/s against 7.4 TB/s for 16 GB).  Each block flushes its
//    128-bit column sums ONCE into 128-bit accumulators in global memory (atomic add of the low word, the carry and the
//    high word follow), and the last block of a chunk to finish reduces the chunk's accumulators - no slab partials,
//    no serial fold of 148 rows at the end of the launch.
constexpr int TALLY_THREADS = 256;
constexpr int TALLY_UNROLL = 8;
// measured on the B200 (tools/prof_tally.py, profiles/r02_tally_sweep.txt): 5 resident blocks per SM (48 registers) and
// 16-ballot work items: 131 072 ballots 298.6 us (7.19 TB/s), 1M ballots 2.326 ms (7.39 TB/s); 4 blocks: 309 / 2409 us
constexpr int TALLY_BLOCKS_PER_SM = 5;
constexpr uint32_t TALLY_MAX_PEERS = 16;

// Cross-GPU stage of the sharded tally, fused into the tally kernel (peer memory over NVLink instead of an
// all-gather + a combine kernel): the block that reduces a column chunk stores its 512 words into row `rank` of EVERY
// peer's inbox, publishes a per-(rank, chunk) flag with the call's epoch (release, system scope), waits until the same
// flag of every peer in its own inbox has REACHED that epoch (acquire; epochs only grow, so a peer that is already one
// call ahead and has overwritten its flag with epoch+1 still satisfies the wait), and sums the `world` rows.  Inboxes
// are double-buffered by epoch parity: a peer can be at most one call ahead (it needs this rank's flag of call k+1 to
// finish call k+1, and this rank sends that only after its call k has completed on its stream), so the rows of epoch
// k are intact while they are read.
struct TallyPeerArgs {
    uint64_t* inbox[TALLY_MAX_PEERS];   // peer p's inbox of this epoch's parity: [world][width] words
    unsigned* flags[TALLY_MAX_PEERS];   // peer p's flags: [world][chunks]
    unsigned* status;                   // host-mapped: first epoch whose wait timed out (0 = none)
    uint32_t world, rank, epoch;
    long long timeout_clocks;
};

struct TallySlot {   // scratch of one single-launch tally (all zero between calls: the reducing block resets what it used)
    unsigned* next;  // [chunks] next work item of the chunk
    unsigned* done;  // [chunks] blocks of the chunk that have flushed
    uint64_t* lo;    // [width] 128-bit column accumulators
    uint64_t* hi;
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

// sums ballots [first, last) of this thread's column pair into (lo0,hi0,lo1,hi1)
__device__ __forceinline__ void tally_rows(const uint64_t* __restrict__ cts, size_t first, size_t last, uint32_t width,
                                           uint32_t col, bool pair, int vec_ok, uint64_t& lo0, uint64_t& hi0,
                                           uint64_t& lo1, uint64_t& hi1) {
    if (vec_ok && pair) {
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
}

// 128-bit atomic accumulate: the carry out of the low word is recovered from the value the atomic returns
__device__ __forceinline__ void atomic_acc128(uint64_t* glo, uint64_t* ghi, uint64_t lo, uint64_t hi) {
    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(glo), (unsigned long long)lo);
    hi += (old + lo < old) ? 1 : 0;
    if (hi) atomicAdd(reinterpret_cast<unsigned long long*>(ghi), (unsigned long long)hi);
}

__global__ void __launch_bounds__(TALLY_THREADS, TALLY_BLOCKS_PER_SM) tally_kernel(const uint64_t* __restrict__ cts, size_t count,
                                                              size_t per_slab, uint32_t width /* 2N words */,
                                                              uint64_t* partial, const ModQ m, int vec_ok,
                                                              const TallySlot slot = TallySlot{}, uint32_t workers = 0, uint32_t item = 0,
                                                              uint64_t* final_out = nullptr, const int use_peers = 0,
                                                              const TallyPeerArgs pa = TallyPeerArgs{}) {
    uint64_t lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
    if (slot.lo == nullptr) {  // ---- slab form
        const uint32_t col = (blockIdx.x * TALLY_THREADS + threadIdx.x) * 2;
        if (col >= width) return;
        const bool pair = (col + 1 < width);
        const size_t first = (size_t)blockIdx.y * per_slab;
        size_t last = first + per_slab;
        if (last > count) last = count;
        tally_rows(cts, first, last, width, col, pair, vec_ok, lo0, hi0, lo1, hi1);
        uint64_t* out = partial + (size_t)blockIdx.y * width + col;
        out[0] = fold128(hi0, lo0, m);
        if (pair) out[1] = fold128(hi1, lo1, m);
        return;
    }
    // ---- single-launch form
    const uint32_t chunks = gridDim.x / workers;
    const uint32_t chunk = blockIdx.x % chunks;
    const uint32_t col = (chunk * TALLY_THREADS + threadIdx.x) * 2;
    const bool live = col < width;  // (threads past the row's end still take part in the barriers below)
    const bool pair = (col + 1 < width);
    __shared__ unsigned s_item[2];
    __shared__ bool is_last;
    const size_t nitems = (count + item - 1) / item;
    if (threadIdx.x == 0) s_item[0] = atomicAdd(slot.next + chunk, 1u);
    __syncthreads();
    unsigned it = s_item[0], par = 0;
    while (it < nitems) {
        if (threadIdx.x == 0) s_item[par ^ 1] = atomicAdd(slot.next + chunk, 1u);  // the next item, while this one streams in
        const size_t first = (size_t)it * item;
        size_t last = first + item;
        if (last > count) last = count;
        if (live) tally_rows(cts, first, last, width, col, pair, vec_ok, lo0, hi0, lo1, hi1);
        __syncthreads();
        par ^= 1;
        it = s_item[par];
    }
    if (live) {
        if (lo0 | hi0) atomic_acc128(slot.lo + col, slot.hi + col, lo0, hi0);
        if (pair && (lo1 | hi1)) atomic_acc128(slot.lo + col + 1, slot.hi + col + 1, lo1, hi1);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(slot.done + chunk, 1u) == workers - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    uint64_t v0 = 0, v1 = 0;
    if (live) {  // accumulators live in L2 (atomics): read them there, and leave them zero for the next call
        v0 = fold128(__ldcg(slot.hi + col), __ldcg(slot.lo + col), m);
        slot.lo[col] = 0;
        slot.hi[col] = 0;
        if (pair) {
            v1 = fold128(__ldcg(slot.hi + col + 1), __ldcg(slot.lo + col + 1), m);
            slot.lo[col + 1] = 0;
            slot.hi[col + 1] = 0;
        }
    }
    if (threadIdx.x == 0) {
        slot.done[chunk] = 0;
        slot.next[chunk] = 0;
    }
    if (!use_peers) {
        if (live) {
            final_out[col] = v0;
            if (pair) final_out[col + 1] = v1;
        }
        return;
    }

    // ---- fused exchange + combine over peer memory ------------------------------------------------
    if (live) {
        for (uint32_t p = 0; p < pa.world; ++p) {  // NVLink stores into every inbox (the own one included)
            uint64_t* dst = pa.inbox[p] + (size_t)pa.rank * width + col;
            dst[0] = v0;
            if (pair) dst[1] = v1;
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    if (threadIdx.x < pa.world) {
        st_release_sys(pa.flags[threadIdx.x] + (size_t)pa.rank * chunks + chunk, pa.epoch);
        const unsigned* mine = pa.flags[pa.rank] + (size_t)threadIdx.x * chunks + chunk;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(mine) - pa.epoch) < 0) {  // epochs are monotonic: "reached", not "equals"
            if (clock64() - t0 > pa.timeout_clocks) {         // a peer never arrived
                timed_out = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (timed_out) {
        // Visible per call: the result words of this chunk become all-ones (never a residue: q < 2^64), and the
        // host-mapped status word records the first failing epoch (checked by the host on every later call).
        if (threadIdx.x == 0) {
            volatile unsigned* st = pa.status;  // plain stores: PCIe atomics on host memory are not guaranteed
            if (*st == 0u) *st = pa.epoch;
            __threadfence_system();
        }
        if (live) {
            final_out[col] = ~0ull;
            if (pair) final_out[col + 1] = ~0ull;
        }
        return;
    }
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

// Scratch slots of the single-launch form: per device TALLY_SLOTS sets of (work counters, blocks-done counters,
// 128-bit column accumulators), all zero between calls (the reducing block resets what it used).  Calls draw slots
// round-robin, so up to TALLY_SLOTS tallies may be in flight on different streams of a device at once.
constexpr unsigned TALLY_SLOTS = 32, TALLY_MAX_CHUNKS = 64;
constexpr uint32_t TALLY_ACC_WORDS = TALLY_MAX_CHUNKS * 2 * TALLY_THREADS;  // rows up to 2 * 16384 words; wider rows use the slab form
static bool tally_slot(TallySlot* out) {
    struct Pool {
        unsigned* counters = nullptr;
        uint64_t* acc = nullptr;
        unsigned next = 0;
    };
    static std::map<int, Pool> pools;  // one pool per device the library has been pointed at
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const int dev = ctx().device;
    Pool& pool = pools[dev];
    if (pool.counters == nullptr) {
        const size_t cbytes = (size_t)TALLY_SLOTS * 2 * TALLY_MAX_CHUNKS * sizeof(unsigned);
        const size_t abytes = (size_t)TALLY_SLOTS * 2 * TALLY_ACC_WORDS * sizeof(uint64_t);
        unsigned* c = nullptr;
        uint64_t* a = nullptr;
        if (cudaMalloc(&c, cbytes) != cudaSuccess || cudaMalloc(&a, abytes) != cudaSuccess ||
            cudaMemset(c, 0, cbytes) != cudaSuccess || cudaMemset(a, 0, abytes) != cudaSuccess ||
            cudaDeviceSynchronize() != cudaSuccess) {
            cudaGetLastError();
            if (c) cudaFree(c);
            if (a) cudaFree(a);
            return false;  // the caller falls back to the slab form
        }
        pool.counters = c;
        pool.acc = a;
    }
    const unsigned k = pool.next++ % TALLY_SLOTS;
    out->next = pool.counters + (size_t)k * 2 * TALLY_MAX_CHUNKS;
    out->done = out->next + TALLY_MAX_CHUNKS;
    out->lo = pool.acc + (size_t)k * 2 * TALLY_ACC_WORDS;
    out->hi = out->lo + TALLY_ACC_WORDS;
    return true;
}

// grid shape of the single-launch form: persistent blocks, `bpsm` per SM, split evenly over the column chunks
struct TallyShape {
    unsigned chunks, workers, item;
};
static TallyShape tally_shape(size_t count, uint32_t width) {
    // tuning knobs, read on every call (tools/prof_tally.py sweeps them inside one process)
    const char* e_bpsm = getenv("FHEB_EXP_TALLY_BPSM");
    const char* e_item = getenv("FHEB_EXP_TALLY_ITEM");
    const int bpsm = e_bpsm ? atoi(e_bpsm) : TALLY_BLOCKS_PER_SM;
    const int item_env = e_item ? atoi(e_item) : 16;
    TallyShape sh;
    sh.chunks = (width + 2 * TALLY_THREADS - 1) / (2 * TALLY_THREADS);
    sh.item = (unsigned)(item_env > 0 ? item_env : 16);
    size_t workers = ((size_t)ctx().sm_count * (bpsm > 0 ? bpsm : TALLY_BLOCKS_PER_SM)) / sh.chunks;
    const size_t nitems = (count + sh.item - 1) / sh.item;
    if (workers > nitems) workers = nitems;
    if (workers < 1) workers = 1;
    sh.workers = (unsigned)workers;
    return sh;
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
    const int vec_ok = ((reinterpret_cast<uintptr_t>(cts) & 15u) == 0) && (width % 2 == 0);
    const unsigned chunks = (width + 2 * TALLY_THREADS - 1) / (2 * TALLY_THREADS);
    TallySlot slot{};
    if (count <= 64) {  // one block per column chunk
        tally_kernel<<<dim3(chunks, 1), TALLY_THREADS, 0, s>>>(cts, count, count, width, out, m, vec_ok);
        FHEB_CHECK_LAUNCH();
        count_launch();
        return FHEB_OK;
    }
    if (width <= TALLY_ACC_WORDS && tally_slot(&slot)) {
        const TallyShape sh = tally_shape(count, width);
        tally_kernel<<<sh.chunks * sh.workers, TALLY_THREADS, 0, s>>>(cts, count, 0, width, nullptr, m, vec_ok, slot, sh.workers, sh.item, out);
        FHEB_CHECK_LAUNCH();
        count_launch();
        return FHEB_OK;
    }
    // very wide rows: slab partials, then one more pass over them
    size_t slabs = ((size_t)ctx().sm_count * 4 + chunks - 1) / chunks;
    const size_t max_slabs = (count + 63) / 64;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    const size_t per_slab = (count + slabs - 1) / slabs;
    slabs = (count + per_slab - 1) / per_slab;
    uint64_t* partial = nullptr;
    FHEB_CUDA(cudaMallocAsync(&partial, slabs * (size_t)width * 8, s));
    tally_kernel<<<dim3(chunks, (unsigned)slabs), TALLY_THREADS, 0, s>>>(cts, count, per_slab, width, partial, m, vec_ok);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        tally_kernel<<<dim3(chunks, 1), TALLY_THREADS, 0, s>>>(partial, slabs, slabs, width, out, m, 1);
        e = cudaGetLastError();
    }
    count_launch(2);
    cudaFreeAsync(partial, s);
    if (e != cudaSuccess) return set_error(FHEB_ERR_NATIVE, "kernel launch failed: %s", cudaGetErrorString(e));
    return FHEB_OK;
}

// device rows -> one row (any count, canonical result); used by the wire-format tally (wire.cu)
int tally_rows_device(const uint64_t* rows, size_t count, uint32_t width, uint64_t q, uint64_t* out, cudaStream_t s) {
    return tally_device(rows, count, width, q, out, false, s);
}

// ---- sharded tally with the exchange fused into the kernel (peer memory) --------------------------------
// One TallyPeers per rank (GPU).  The ranks are processes (inboxes mapped with CUDA IPC: fheb_tally_peers_*) or
// the devices of ONE process (peer access enabled, plain device pointers: fheb_tally_group_*).
struct TallyPeers {
    uint32_t degree = 0, world = 0, rank = 0, chunks = 0, epoch = 0;
    int device = 0;
    uint64_t modulus = 0;
    void* local = nullptr;                 // this rank's exported allocation: inbox[2][world][width] | flags[world][chunks]
    void* peer[TALLY_MAX_PEERS] = {};      // the peers' allocations as this rank addresses them (peer[rank] == local)
    bool ipc[TALLY_MAX_PEERS] = {};        // opened with cudaIpcOpenMemHandle (to be closed)
    unsigned* h_status = nullptr;          // pinned, mapped: first epoch whose exchange timed out (0 = none)
    unsigned* d_status = nullptr;          // device address of the same word
    TallySlot slot{};                      // own scratch (not drawn from the shared pool: a stuck call must not poison it)
    long long timeout_clocks = 1ll << 35;  // ~17 s at 2 GHz
    bool connected = false;
    size_t inbox_bytes() const { return (size_t)world * 2 * degree * 8; }
    size_t flags_offset() const { return 2 * inbox_bytes(); }
    size_t scratch_offset() const { return flags_offset() + (((size_t)world * chunks * 4 + 15) & ~(size_t)15); }
    size_t total_bytes() const { return scratch_offset() + (size_t)2 * chunks * 4 + 16 + (size_t)2 * 2 * degree * 8; }
};

static int tally_peers_alloc(TallyPeers* tp) {
    FHEB_CUDA(cudaMalloc(&tp->local, tp->total_bytes()));
    FHEB_CUDA(cudaMemset(tp->local, 0, tp->total_bytes()));
    FHEB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&tp->h_status), 64, cudaHostAllocMapped | cudaHostAllocPortable));
    *tp->h_status = 0;
    FHEB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&tp->d_status), tp->h_status, 0));
    char* sc = static_cast<char*>(tp->local) + tp->scratch_offset();
    tp->slot.next = reinterpret_cast<unsigned*>(sc);
    tp->slot.done = tp->slot.next + tp->chunks;
    tp->slot.lo = reinterpret_cast<uint64_t*>(sc + (((size_t)2 * tp->chunks * 4 + 15) & ~(size_t)15));
    tp->slot.hi = tp->slot.lo + (size_t)2 * tp->degree;
    FHEB_CUDA(cudaDeviceSynchronize());
    tp->peer[tp->rank] = tp->local;
    return FHEB_OK;
}

static void tally_peers_free(TallyPeers* tp) {
    for (uint32_t p = 0; p < tp->world; ++p)
        if (p != tp->rank && tp->peer[p] && tp->ipc[p]) cudaIpcCloseMemHandle(tp->peer[p]);
    if (tp->local) cudaFree(tp->local);
    if (tp->h_status) cudaFreeHost(tp->h_status);
    cudaGetLastError();
    delete tp;
}

static int tally_peers_failed(const TallyPeers* tp) {
    const unsigned e = *reinterpret_cast<volatile unsigned*>(tp->h_status);
    if (e == 0) return FHEB_OK;
    return set_error(FHEB_ERR_NATIVE,
                     "fused sharded tally: a peer rank did not arrive in the exchange of call %u; the results of that call "
                     "(all-ones words) and of later calls are invalid - destroy the handle or call fheb_tally_peers_reset on every rank",
                     e);
}

static int tally_peers_run_device(TallyPeers* tp, const uint64_t* cts, size_t count, uint64_t* out, cudaStream_t s) {
    FHEB_TRY(tally_peers_failed(tp));  // an earlier call of this handle timed out: fail loudly instead of returning stale sums
    const uint32_t width = 2 * tp->degree;
    const ModQ m = make_modq(tp->modulus);
    TallyShape sh = tally_shape(count ? count : 1, width);
    TallyPeerArgs pa{};
    pa.world = tp->world;
    pa.rank = tp->rank;
    pa.epoch = tp->epoch + 1;  // committed only when the launch succeeded: a failed call must not desynchronise the ranks
    pa.timeout_clocks = tp->timeout_clocks;
    const size_t parity = (pa.epoch & 1u) * tp->inbox_bytes();
    for (uint32_t p = 0; p < tp->world; ++p) {
        pa.inbox[p] = reinterpret_cast<uint64_t*>(static_cast<char*>(tp->peer[p]) + parity);
        pa.flags[p] = reinterpret_cast<unsigned*>(static_cast<char*>(tp->peer[p]) + tp->flags_offset());
    }
    pa.status = tp->d_status;
    const int vec_ok = ((reinterpret_cast<uintptr_t>(cts) & 15u) == 0);
    tally_kernel<<<sh.chunks * sh.workers, TALLY_THREADS, 0, s>>>(cts, count, 0, width, nullptr, m, vec_ok, tp->slot, sh.workers, sh.item, out, 1, pa);
    FHEB_CHECK_LAUNCH();
    tp->epoch = pa.epoch;
    count_launch();
    return FHEB_OK;
}

static int tally_entry_single(const uint64_t* cts, size_t count, uint32_t degree, uint64_t q, uint64_t* out, void* stream, bool single_raw = false);

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
    if (count > 1 && all_host({cts, out}) && spread_over_devices(count, count * (size_t)width * 8)) {
        // host ballots, several GPUs: every device folds a contiguous share (own thread, own PCIe link) into one row of
        // `parts`; the rows are canonical partial tallies and are folded on the current device
        const std::vector<int> devs = device_list();
        std::vector<uint64_t> parts(devs.size() * (size_t)width, 0);
        std::vector<char> used(devs.size(), 0);
        FHEB_TRY(run_on_devices(count, [&](int device, size_t first, size_t n) {
            size_t slot = 0;
            while (slot < devs.size() && devs[slot] != device) ++slot;
            used[slot] = 1;
            // single_raw = false: a share of one ballot must come back reduced, it is an operand of the final fold;
            // the caller's stream belongs to the caller's device: every worker uses its own device's default stream
            return tally_entry_single(cts + first * width, n, degree, q, parts.data() + slot * (size_t)width, nullptr, false);
        }));
        std::vector<uint64_t> rows;
        for (size_t d = 0; d < devs.size(); ++d)
            if (used[d]) rows.insert(rows.end(), parts.begin() + d * width, parts.begin() + (d + 1) * width);
        const size_t nrows = rows.size() / width;
        Staged sin, sout;
        FHEB_TRY(sin.bind(rows.data(), rows.size() * 8, true, false, s));
        FHEB_TRY(sout.bind(out, (size_t)width * 8, false, true, s));
        FHEB_TRY(tally_device(sin.ptr<const uint64_t>(), nrows, width, q, sout.ptr<uint64_t>(), false, s));
        FHEB_TRY(sout.finish());
        return sync_if_staged(s, {&sin, &sout});
    }
    return tally_entry_single(cts, count, degree, q, out, stream, single_raw);
}

static int tally_entry_single(const uint64_t* cts, size_t count, uint32_t degree, uint64_t q, uint64_t* out, void* stream, bool single_raw) {
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t width = 2 * degree;
    if (count > 1 && all_host({cts}) && count * (size_t)width * 8 >= (16u << 20)) {
        // host ballots (the result may live on either side: the streaming accumulator passes a device row): each
        // chunk is copied in and folded to one partial tally while the next chunk is in flight; the partials are
        // folded at the end (any grouping gives the same words)
        const size_t row = (size_t)width * 8;
        size_t chunk = (8u << 20) / row;
        if (chunk < 1) chunk = 1;
        const size_t nchunks = (count + chunk - 1) / chunk;
        const bool out_dev = is_device_pointer(out);
        uint64_t* partial = nullptr;
        FHEB_CUDA(cudaMalloc(&partial, (nchunks + 1) * row));
        uint64_t* dout = out_dev ? out : partial + nchunks * (size_t)width;
        int rc = run_host_pipeline(count, {{cts, row, 0, true, false}},
                                   [&](void* const* d, size_t first, size_t n, cudaStream_t ps) {
                                       return tally_device(static_cast<const uint64_t*>(d[0]), n, width, q, partial + (first / chunk) * width, false, ps);
                                   },
                                   chunk);
        if (rc == FHEB_OK) rc = tally_device(partial, nchunks, width, q, dout, false, s);
        if (rc == FHEB_OK && !out_dev && cudaMemcpyAsync(out, dout, row, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "copy of the tally failed");
        if (rc == FHEB_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "tally failed");
        cudaFree(partial);
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

static int tally_peers_new(uint32_t degree, uint64_t modulus, uint32_t world, uint32_t rank, TallyPeers** out) {
    FHEB_REQUIRE(degree > 0 && (degree & (degree - 1)) == 0, "Polynomial degree must be a power of 2");
    FHEB_REQUIRE(modulus >= 2, "Modulus must be at least 2");
    FHEB_REQUIRE(world >= 1 && world <= TALLY_MAX_PEERS && rank < world, "world must be 1..%u and rank below it", TALLY_MAX_PEERS);
    TallyPeers* tp = new TallyPeers();
    tp->degree = degree;
    tp->modulus = modulus;
    tp->world = world;
    tp->rank = rank;
    tp->device = ctx().device;
    tp->chunks = (2 * degree + 2 * TALLY_THREADS - 1) / (2 * TALLY_THREADS);
    if (tp->chunks > TALLY_MAX_CHUNKS) {
        delete tp;
        return set_error(FHEB_ERR_INVALID_PARAMETERS, "degree too large for the fused sharded tally");
    }
    const int rc = tally_peers_alloc(tp);
    if (rc != FHEB_OK) {
        cudaGetLastError();
        tally_peers_free(tp);
        return rc;
    }
    *out = tp;
    return FHEB_OK;
}

int fheb_tally_peers_create(uint32_t degree, uint64_t modulus, uint32_t world, uint32_t rank, fheb_tally_peers** out,
                            uint8_t* handle_out) {
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(out != nullptr && handle_out != nullptr, "out and handle_out must not be null");
    FHEB_REQUIRE(world >= 2, "world must be 2..%u and rank below it", TALLY_MAX_PEERS);
    static_assert(sizeof(cudaIpcMemHandle_t) == FHEB_PEER_HANDLE_BYTES, "IPC handle size");
    TallyPeers* tp = nullptr;
    FHEB_TRY(tally_peers_new(degree, modulus, world, rank, &tp));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, tp->local);
    if (e != cudaSuccess) {
        cudaGetLastError();
        tally_peers_free(tp);
        return set_error(FHEB_ERR_NATIVE, "peer buffer set-up failed: %s", cudaGetErrorString(e));
    }
    std::memcpy(handle_out, &h, sizeof(h));
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
        tp->ipc[p] = true;
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
    *timed_out = (int)*reinterpret_cast<volatile unsigned*>(tp->h_status);  // host-mapped word: no copy, no synchronisation
    return FHEB_OK;
}

uint32_t fheb_tally_peers_epoch(const fheb_tally_peers* peers) {
    return peers ? reinterpret_cast<const TallyPeers*>(peers)->epoch : 0;
}

int fheb_tally_peers_reset(fheb_tally_peers* peers, uint32_t epoch) {
    // After a failed exchange: every rank synchronises its stream, agrees on an epoch not used before (e.g. the
    // maximum of fheb_tally_peers_epoch over the ranks, plus 2) and calls this; flags of earlier epochs compare as old.
    FHEB_REQUIRE(peers != nullptr, "peers must not be null");
    TallyPeers* tp = reinterpret_cast<TallyPeers*>(peers);
    int cur = -1;
    cudaGetDevice(&cur);
    FHEB_CUDA(cudaSetDevice(tp->device));
    FHEB_CUDA(cudaDeviceSynchronize());
    // own scratch back to zero (a call that timed out has reset it already; a failed launch never touched it)
    FHEB_CUDA(cudaMemset(static_cast<char*>(tp->local) + tp->scratch_offset(), 0, tp->total_bytes() - tp->scratch_offset()));
    FHEB_CUDA(cudaDeviceSynchronize());
    tp->epoch = epoch;
    *tp->h_status = 0;
    if (cur >= 0) cudaSetDevice(cur);
    return FHEB_OK;
}

int fheb_tally_peers_set_timeout(fheb_tally_peers* peers, double seconds) {
    FHEB_REQUIRE(peers != nullptr && seconds > 0, "peers must not be null and the timeout positive");
    TallyPeers* tp = reinterpret_cast<TallyPeers*>(peers);
    tp->timeout_clocks = (long long)(seconds * 1e3 * (double)ctx_of(tp->device).prop.clockRate);  // clockRate is in kHz
    if (tp->timeout_clocks < 1000) tp->timeout_clocks = 1000;
    return FHEB_OK;
}

int fheb_tally_peers_destroy(fheb_tally_peers* peers) {
    if (!peers) return FHEB_OK;
    TallyPeers* tp = reinterpret_cast<TallyPeers*>(peers);
    int cur = -1;
    cudaGetDevice(&cur);
    cudaSetDevice(tp->device);
    cudaDeviceSynchronize();
    tally_peers_free(tp);
    if (cur >= 0) cudaSetDevice(cur);
    return FHEB_OK;
}

// ---- the same exchange between the GPUs of ONE process (the reference's addon is one process,
// src/native/lib.rs:23-133).  Peer access instead of CUDA IPC; one host thread launches every device's kernel.
struct TallyGroup {
    uint32_t ndev = 0, degree = 0;
    int device[TALLY_MAX_PEERS] = {};
    TallyPeers* rank[TALLY_MAX_PEERS] = {};
    cudaStream_t stream[TALLY_MAX_PEERS] = {};
    uint64_t* result[TALLY_MAX_PEERS] = {};  // [2][N] on each device (every rank ends with the global tally)
};

int fheb_tally_group_destroy(fheb_tally_group* group) {
    TallyGroup* g = reinterpret_cast<TallyGroup*>(group);
    if (!g) return FHEB_OK;
    int cur = -1;
    cudaGetDevice(&cur);
    for (uint32_t i = 0; i < g->ndev; ++i) {
        cudaSetDevice(g->device[i]);
        cudaDeviceSynchronize();
        if (g->rank[i]) tally_peers_free(g->rank[i]);
        if (g->result[i]) cudaFree(g->result[i]);
        if (g->stream[i]) cudaStreamDestroy(g->stream[i]);
    }
    cudaGetLastError();
    if (cur >= 0) cudaSetDevice(cur);
    delete g;
    return FHEB_OK;
}

int fheb_tally_group_create(uint32_t degree, uint64_t modulus, const int* devices, uint32_t ndev, fheb_tally_group** out) {
    FHEB_REQUIRE(out != nullptr, "out must not be null");
    *out = nullptr;
    FHEB_REQUIRE(ndev >= 1 && ndev <= TALLY_MAX_PEERS, "the group takes 1..%u devices", TALLY_MAX_PEERS);
    int cur = -1;
    cudaGetDevice(&cur);
    TallyGroup* g = new TallyGroup();
    g->ndev = ndev;
    g->degree = degree;
    int rc = FHEB_OK;
    for (uint32_t i = 0; i < ndev && rc == FHEB_OK; ++i) {
        g->device[i] = devices ? devices[i] : (int)i;
        for (uint32_t j = 0; j < i; ++j)
            if (g->device[j] == g->device[i]) rc = set_error(FHEB_ERR_INVALID_PARAMETERS, "device %d listed twice", g->device[i]);
        if (rc == FHEB_OK) rc = fheb_init(g->device[i]);  // makes it current and creates its context
        if (rc == FHEB_OK) rc = tally_peers_new(degree, modulus, ndev, i, &g->rank[i]);
        if (rc == FHEB_OK && cudaStreamCreateWithFlags(&g->stream[i], cudaStreamNonBlocking) != cudaSuccess)
            rc = set_error(FHEB_ERR_NATIVE, "stream creation failed on device %d", g->device[i]);
        if (rc == FHEB_OK && cudaMalloc(&g->result[i], (size_t)2 * degree * 8) != cudaSuccess)
            rc = set_error(FHEB_ERR_OUT_OF_MEMORY, "cudaMalloc failed on device %d", g->device[i]);
    }
    for (uint32_t i = 0; i < ndev && rc == FHEB_OK; ++i) {
        cudaSetDevice(g->device[i]);
        for (uint32_t j = 0; j < ndev && rc == FHEB_OK; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->device[i], g->device[j]);
            if (!can) {
                rc = set_error(FHEB_ERR_HARDWARE_UNAVAILABLE, "device %d cannot map the memory of device %d", g->device[i], g->device[j]);
                break;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(g->device[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                rc = set_error(FHEB_ERR_NATIVE, "enabling peer access %d -> %d failed: %s", g->device[i], g->device[j], cudaGetErrorString(e));
            cudaGetLastError();
            g->rank[i]->peer[j] = g->rank[j]->local;
        }
        if (rc == FHEB_OK) g->rank[i]->connected = true;
    }
    if (cur >= 0) cudaSetDevice(cur);
    if (rc != FHEB_OK) {
        const std::string keep = fheb_last_error();
        fheb_tally_group_destroy(reinterpret_cast<fheb_tally_group*>(g));
        return set_error(rc, "%s", keep.c_str());
    }
    *out = reinterpret_cast<fheb_tally_group*>(g);
    return FHEB_OK;
}

int fheb_tally_sharded(fheb_tally_group* group, const uint64_t* const* cts, const size_t* counts, uint64_t* out) {
    // EncryptionEngine::tally_votes over ballots that are sharded across the group's GPUs (cpp/src/encryption.cpp:
    // 1061-1067,1327-1458): cts[i] = device pointer on device i to counts[i] ballots [count][2][N]; out = [2][N],
    // host memory or memory of any device of the group.  Any split gives the reference's words.
    FHEB_REQUIRE(group != nullptr && cts != nullptr && counts != nullptr && out != nullptr, "group, cts, counts and out must not be null");
    TallyGroup* g = reinterpret_cast<TallyGroup*>(group);
    size_t total = 0;
    for (uint32_t i = 0; i < g->ndev; ++i) {
        total += counts[i];
        FHEB_REQUIRE(counts[i] == 0 || cts[i] != nullptr, "cts[%u] must not be null", i);
    }
    // message follows EncryptionEngine::batch_add, cpp/src/encryption.cpp:1328-1330
    FHEB_REQUIRE(total != 0, "Cannot add empty vector of ciphertexts");
    int cur = -1;
    cudaGetDevice(&cur);
    int rc = FHEB_OK;
    uint32_t launched = 0;
    for (uint32_t i = 0; i < g->ndev && rc == FHEB_OK; ++i, ++launched) {
        if (cudaSetDevice(g->device[i]) != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "cudaSetDevice(%d) failed", g->device[i]);
        if (rc == FHEB_OK) rc = tally_peers_run_device(g->rank[i], cts[i], counts[i], g->result[i], g->stream[i]);
    }
    if (rc != FHEB_OK && launched > 1) {
        // some ranks are in the exchange and the rest never will be: they time out; say what has to happen next
        const std::string keep = fheb_last_error();
        rc = set_error(rc, "%s (ranks already launched will time out; destroy the group)", keep.c_str());
    }
    if (rc == FHEB_OK) {
        cudaSetDevice(g->device[0]);
        if (cudaMemcpyAsync(out, g->result[0], (size_t)2 * g->degree * 8, cudaMemcpyDefault, g->stream[0]) != cudaSuccess)
            rc = set_error(FHEB_ERR_NATIVE, "copy of the tally failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    for (uint32_t i = 0; i < launched; ++i) {
        cudaSetDevice(g->device[i]);
        const cudaError_t e = cudaStreamSynchronize(g->stream[i]);
        if (e != cudaSuccess && rc == FHEB_OK) rc = set_error(FHEB_ERR_NATIVE, "sharded tally failed on device %d: %s", g->device[i], cudaGetErrorString(e));
    }
    for (uint32_t i = 0; i < launched && rc == FHEB_OK; ++i) rc = tally_peers_failed(g->rank[i]);
    if (cur >= 0) cudaSetDevice(cur);
    return rc;
}

uint32_t fheb_tally_group_size(const fheb_tally_group* group) { return group ? reinterpret_cast<const TallyGroup*>(group)->ndev : 0; }

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
    } else {  // host chunk: the one-shot entry point pipelines it (chunked copies overlapped with the folds) into the device row
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
    auto overlaps = [&](const uint64_t* in) {  // [in, in + 2 N batch) against [out, out + 3 N batch)
        const uint64_t* o0 = so.ptr<const uint64_t>();
        return in < o0 + batch * 3 * N && o0 < in + batch * 2 * N;
    };
    if (!getenv("FHEB_TENSOR_UNFUSED") && !overlaps(s1.ptr<const uint64_t>()) && !overlaps(s2.ptr<const uint64_t>())) {
        // one launch (tensor_fused.cu); an output that overlaps an operand takes the unfused path below, which has
        // transformed both operands into its scratch before the first output word is written
        const int frc = tensor_fused_launch(p, s1.ptr<const uint64_t>(), s2.ptr<const uint64_t>(), so.ptr<uint64_t>(), batch, s);
        if (frc != TENSOR_FUSED_UNSUPPORTED) {
            FHEB_TRY(frc);
            FHEB_TRY(so.finish());
            return sync_if_staged(s, {&s1, &s2, &so});
        }
    }
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

int fheb_tally_noise_budget(const double* budgets, size_t count, int variant, double* out) {
    // The noise_budget metadata of the reference's tally variants (the polynomial words do not depend on the variant;
    // this number does - SURVEY B11).  Host arithmetic on doubles, same operations in the same order:
    //   0 batch_add        min over all - log2(count)                       cpp/src/encryption.cpp:1337-1360
    //   1 batch_add_tree   min(pair) - 1 per level, odd element carried      cpp/src/encryption.cpp:1390-1456 (:1413,:1437)
    //   2 add, left fold   min(running, next) - 1 per ciphertext             cpp/src/encryption.cpp:613 (stream_add's accumulate)
    // A single ciphertext keeps its budget in every variant (:1332-1334, :1374-1376).
    FHEB_REQUIRE(budgets != nullptr && out != nullptr, "budgets and out must not be null");
    FHEB_REQUIRE(count != 0, "Cannot add empty vector of ciphertexts");
    FHEB_REQUIRE(variant >= 0 && variant <= 2, "unknown tally variant %d", variant);
    if (count == 1) {
        *out = budgets[0];
        return FHEB_OK;
    }
    if (variant == 0) {
        double lo = budgets[0];
        for (size_t i = 1; i < count; ++i) lo = std::min(lo, budgets[i]);
        *out = lo - std::log2(static_cast<double>(count));
        return FHEB_OK;
    }
    if (variant == 2) {
        double acc = budgets[0];
        for (size_t i = 1; i < count; ++i) acc = std::min(acc, budgets[i]) - 1.0;
        *out = acc;
        return FHEB_OK;
    }
    std::vector<double> level(budgets, budgets + count), next;
    while (level.size() > 1) {
        next.clear();
        for (size_t i = 0; i + 1 < level.size(); i += 2) next.push_back(std::min(level[i], level[i + 1]) - 1.0);
        if (level.size() % 2 == 1) next.push_back(level.back());
        level.swap(next);
    }
    *out = level[0];
    return FHEB_OK;
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
