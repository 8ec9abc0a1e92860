// Wire formats either side of the tally and of key upload: the FHEV ballot container and
// the FHEE / FHEB key containers of cpp/include/key_serializer.h:28-84, cpp/src/key_serializer.cpp.
//
// Container = packed 49-byte header (write_header, key_serializer.cpp:96-108: magic u32 | version u32 |
// key_type u32 | key_id u64 | poly_degree u32 | modulus u64 | data_size u32 | checksum_type u8 | compression u8 |
// reserved[7] | checksum u32, little endian, no padding) followed by data_size payload bytes; checksum = CRC of
// the payload.  The reference's CRC is table driven (:34-40) with a table that holds only the first 60 of the 256
// IEEE 802.3 entries (:21-32; the remaining initialisers are missing, i.e. zero), so it is NOT the standard
// CRC-32 and not GF(2)-linear: a record's checksum can only be computed byte by byte.  Parity means the same
// table, so that ballots written by the reference verify here and vice versa.
//
// Ballots (high volume, C5 feeds on them) are validated and unpacked ON THE DEVICE: one thread walks one record's
// checksum (independent records in parallel), then a coalesced kernel realigns the payload words (the payload
// starts 73 bytes into a record, never 8-byte aligned) into the [count][choices][2][N] layout fheb_tally reads.
// Key containers are parsed once on the host and handed to the device key constructors.
//
// C-ABI entry points here (include/fheb200.h): fheb_wire_crc32, fheb_wire_header_read, fheb_ballot_serialize,
// fheb_ballots_ingest, fheb_relin_key_from_wire, fheb_boot_key_from_wire.
#include <cstdlib>
#include <memory>

#include "elementwise.hpp"
#include "plan.hpp"
#include "runtime.hpp"

namespace fheb {

constexpr uint32_t WIRE_HEADER_BYTES = 49;     // bytes write_header emits
constexpr uint32_t WIRE_HEADER_STRUCT = 64;    // sizeof(SerializationHeader) with natural alignment: the "Input too small" bound (:781)
constexpr uint32_t WIRE_CRC_ENTRIES = 60;      // initialisers present in the reference's table (:21-32)
constexpr uint32_t MAGIC_BALLOT = 0x46484556;  // "FHEV"
constexpr uint32_t MAGIC_EVAL = 0x46484545;    // "FHEE"
constexpr uint32_t MAGIC_BOOT = 0x46484542;    // "FHEB"

// entry i of the reflected IEEE 802.3 table (polynomial 0xEDB88320), zero from WIRE_CRC_ENTRIES up
__host__ __device__ inline uint32_t wire_crc_entry(uint32_t i) {
    if (i >= WIRE_CRC_ENTRIES) return 0;
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    return c;
}

struct CrcTable {
    uint32_t t[256];
    CrcTable() {
        for (uint32_t i = 0; i < 256; ++i) t[i] = wire_crc_entry(i);
    }
};
static const CrcTable& crc_table() {
    static const CrcTable tab;
    return tab;
}

static uint32_t crc32_host(const uint8_t* data, size_t len, size_t zero_pad = 0) {
    const uint32_t* t = crc_table().t;
    uint32_t crc = 0xFFFFFFFFu;
    for (size_t i = 0; i < len; ++i) crc = t[(crc ^ data[i]) & 0xFF] ^ (crc >> 8);
    for (size_t i = 0; i < zero_pad; ++i) crc = t[crc & 0xFF] ^ (crc >> 8);
    return crc ^ 0xFFFFFFFFu;
}

template <class T>
__host__ __device__ inline T load_le(const uint8_t* p) {  // unaligned little-endian field
    T v = 0;
    for (unsigned i = 0; i < sizeof(T); ++i) v |= (T)p[i] << (8 * i);
    return v;
}
template <class T>
static void store_le(uint8_t* p, T v) {
    for (unsigned i = 0; i < sizeof(T); ++i) p[i] = (uint8_t)(v >> (8 * i));
}

static void parse_header(const uint8_t* p, fheb_wire_header* h) {
    h->magic = load_le<uint32_t>(p);
    h->version = load_le<uint32_t>(p + 4);
    h->key_type = load_le<uint32_t>(p + 8);
    h->key_id = load_le<uint64_t>(p + 12);
    h->poly_degree = load_le<uint32_t>(p + 20);
    h->modulus = load_le<uint64_t>(p + 24);
    h->data_size = load_le<uint32_t>(p + 32);
    h->checksum_type = p[36];
    h->compression = p[37];
    h->checksum = load_le<uint32_t>(p + 45);
}

// ---- device: ballot validation ----------------------------------------------------------------
// status: FHEB_WIRE_* of include/fheb200.h.  The checks and their order are BallotSerializer::deserialize_ballot's
// (:776-812): size, magic, checksum over data_size bytes - the reference reads the payload into a zero-filled
// buffer, so a truncated record is checksummed with zero padding - and then the shape this bulk path needs
// (choices, degree and modulus of every choice, exact payload length).
//
// One WARP per record.  The checksum is a byte-serial recurrence, but with 196 zero table entries its state is
// forgotten quickly: the state after byte i is T[x_i] ^ T[x_{i-1}]>>8 ^ T[x_{i-2}]>>16 ^ T[x_{i-3}]>>24, and most
// T[x] are zero.  So every lane takes one 1/32 slice of the payload and GUESSES its start state by running from
// state 0 over the WIRE_WARMUP bytes in front of the slice; afterwards lane k checks its guess against lane k-1's
// end state.  Lane 0 starts from the true initial state, so when every check holds the chain is exact by
// induction; a lane whose guess was wrong re-runs its slice from the neighbour's end state until all checks hold
// (at most 31 rounds, in practice none: measured miss rate per boundary below 0.2 % on 62-bit residues).  The
// result is always the exact serial checksum, only the schedule is speculative.
constexpr int WIRE_THREADS = 256;
constexpr uint32_t WIRE_WARMUP = 128;
constexpr size_t WIRE_TAB_BYTES = 256 * 32 * 4;
constexpr uint32_t WIRE_PARALLEL_MIN = 4096;  // payloads shorter than this are walked by one lane

// tab: all 256 entries replicated per lane ([entry][lane], 32 KB) so that 32 random lookups never share a bank and
// no range test is needed; tab_lane = 32-bit shared address of this lane's column.
__device__ __forceinline__ uint32_t crc_lookup(uint32_t tab_lane, uint32_t idx) {
    uint32_t t;
    asm("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(tab_lane + (idx << 7)));
    return t;
}
__device__ __forceinline__ uint32_t crc_step(uint32_t crc, uint32_t byte, uint32_t tab_lane) {
    return crc_lookup(tab_lane, (crc ^ byte) & 0xFF) ^ (crc >> 8);
}
// four bytes: XOR the word in once - the bytes reach the low end of the shift register exactly when their step
// comes (plain algebra of the recurrence, independent of the table's contents)
__device__ __forceinline__ uint32_t crc_word(uint32_t crc, uint32_t w, uint32_t tab_lane) {
    crc ^= w;
#pragma unroll
    for (int k = 0; k < 4; ++k) crc = crc_lookup(tab_lane, crc & 0xFF) ^ (crc >> 8);
    return crc;
}
__device__ __forceinline__ uint32_t crc_words(uint32_t crc, const uint4& v, uint32_t tab_lane) {
    crc = crc_word(crc, v.x, tab_lane);
    crc = crc_word(crc, v.y, tab_lane);
    crc = crc_word(crc, v.z, tab_lane);
    return crc_word(crc, v.w, tab_lane);
}

// The serial walk of one slice.  Lanes of a warp stream through different slices, so their loads do not coalesce;
// the next 64 bytes are requested before the current 64 are folded in, which hides the L2 latency behind the
// dependent chain of table steps.
__device__ __forceinline__ uint32_t crc_run(uint32_t crc, const uint8_t* p, uint64_t n, uint32_t tab_lane) {
    while (n && ((uintptr_t)p & 15)) {  // up to the first 16-byte boundary
        crc = crc_step(crc, *p++, tab_lane);
        --n;
    }
    if (n >= 64) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 cur[4] = {__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3)};
        while (n >= 128) {
            const uint4 nxt[4] = {__ldg(q + 4), __ldg(q + 5), __ldg(q + 6), __ldg(q + 7)};
#pragma unroll
            for (int k = 0; k < 4; ++k) crc = crc_words(crc, cur[k], tab_lane);
#pragma unroll
            for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
            q += 4;
            n -= 64;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) crc = crc_words(crc, cur[k], tab_lane);
        q += 4;
        n -= 64;
        p = reinterpret_cast<const uint8_t*>(q);
    }
    for (; n >= 16; n -= 16, p += 16) crc = crc_words(crc, __ldg(reinterpret_cast<const uint4*>(p)), tab_lane);
    for (; n; --n) crc = crc_step(crc, *p++, tab_lane);
    return crc;
}

__global__ void __launch_bounds__(WIRE_THREADS) ballot_validate_kernel(const uint8_t* __restrict__ wire, const uint64_t* __restrict__ offsets,
                                                                     size_t count, uint32_t choices, uint32_t N, uint64_t q,
                                                                     uint8_t* __restrict__ status, uint64_t* __restrict__ timestamps,
                                                                     uint32_t warmup) {
    extern __shared__ uint32_t tab[];  // [256][32]
    for (uint32_t i = threadIdx.x; i < 256 * 32; i += WIRE_THREADS) tab[i] = wire_crc_entry(i >> 5);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t tab_lane = (uint32_t)__cvta_generic_to_shared(tab + lane);
    const size_t warps = (size_t)gridDim.x * (WIRE_THREADS / 32);
    for (size_t r = (size_t)blockIdx.x * (WIRE_THREADS / 32) + (threadIdx.x >> 5); r < count; r += warps) {
        const uint64_t off = offsets[r], end = offsets[r + 1];
        const uint64_t span = end - off;
        const uint8_t* rec = wire + off;
        uint8_t st = FHEB_WIRE_OK;
        uint64_t stamp = 0;
        if (span < WIRE_HEADER_STRUCT) {
            st = FHEB_WIRE_TOO_SMALL;
        } else if (load_le<uint32_t>(rec) != MAGIC_BALLOT) {
            st = FHEB_WIRE_BAD_MAGIC;
        } else {
            const uint32_t data_size = load_le<uint32_t>(rec + 32);
            const uint32_t expected = load_le<uint32_t>(rec + 45);
            const uint64_t avail = span - WIRE_HEADER_BYTES;
            const uint64_t n = data_size < avail ? data_size : avail;
            const uint64_t pad = data_size - n;
            const uint8_t* d = rec + WIRE_HEADER_BYTES;
            uint32_t crc;
            if (n < WIRE_PARALLEL_MIN) {
                crc = crc_run(0xFFFFFFFFu, d, n, tab_lane);  // every lane walks it (same addresses: broadcast loads)
            } else {
                const uint64_t chunk = (n + 31) / 32;
                const uint64_t lo = lane * chunk < n ? lane * chunk : n;
                const uint64_t hi = lo + chunk < n ? lo + chunk : n;
                const uint64_t warm = lo < warmup ? 0 : lo - warmup;
                uint32_t a = crc_run(warm == 0 ? 0xFFFFFFFFu : 0u, d + warm, lo - warm, tab_lane);  // guessed start state
                uint32_t e = crc_run(a, d + lo, hi - lo, tab_lane);
                for (;;) {
                    const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, e, 1);
                    const bool ok = lane == 0 || (warm == 0 && warmup != 0) || prev == a;
                    if (__all_sync(0xFFFFFFFFu, ok)) break;
                    if (!ok) {
                        a = prev;
                        e = crc_run(a, d + lo, hi - lo, tab_lane);
                    }
                }
                crc = __shfl_sync(0xFFFFFFFFu, e, 31);
            }
            // zero padding of a truncated record; under zero input the state 0 is a fixed point (T[0] == 0)
            for (uint64_t i = 0; i < pad && crc != 0; ++i) crc = crc_step(crc, 0, tab_lane);
            if ((crc ^ 0xFFFFFFFFu) != expected) {
                st = FHEB_WIRE_BAD_CHECKSUM;
            } else {
                // shape: timestamp u64 | num_choices u32 | per choice: degree u32 | modulus u64 | a[N] | b[N]   (:718-735)
                const uint64_t per_choice = 12 + 16 * (uint64_t)N;
                if (pad != 0 || data_size != 12 + (uint64_t)choices * per_choice || load_le<uint32_t>(d + 8) != choices) {
                    st = FHEB_WIRE_SHAPE_MISMATCH;
                } else {
                    for (uint32_t c = 0; c < choices && st == FHEB_WIRE_OK; ++c) {
                        const uint8_t* ch = d + 12 + c * per_choice;
                        if (load_le<uint32_t>(ch) != N || load_le<uint64_t>(ch + 4) != q) st = FHEB_WIRE_SHAPE_MISMATCH;
                    }
                    if (st == FHEB_WIRE_OK) stamp = load_le<uint64_t>(d);
                }
            }
        }
        if (lane == 0) {
            status[r] = st;
            if (timestamps) timestamps[r] = stamp;
        }
    }
}

// Payload words of accepted records -> cts[count][choices][2][N]; rejected records become zero ciphertexts (the
// additive identity: they drop out of a tally of the whole buffer).  One block walks one record at a time; every
// thread builds its output word from the two aligned words that straddle it.
__global__ void __launch_bounds__(256) ballot_unpack_kernel(const uint8_t* __restrict__ wire, const uint8_t* wire_end,
                                                            const uint64_t* __restrict__ offsets,
                                                            const uint8_t* __restrict__ status, size_t count, uint32_t choices,
                                                            uint32_t N, uint64_t* __restrict__ cts) {
    const size_t words_per_choice = 2 * (size_t)N;
    const size_t words = choices * words_per_choice;
    for (size_t r = blockIdx.x; r < count; r += gridDim.x) {
        uint64_t* out = cts + r * words;
        if (status[r] != FHEB_WIRE_OK) {
            for (size_t i = threadIdx.x; i < words; i += blockDim.x) __stcs(out + i, 0ull);
            continue;
        }
        const uint8_t* d = wire + offsets[r] + WIRE_HEADER_BYTES + 12;
        for (size_t i = threadIdx.x; i < words; i += blockDim.x) {
            const size_t c = i / words_per_choice, j = i - c * words_per_choice;
            const uint8_t* src = d + c * (12 + 8 * words_per_choice) + 12 + 8 * j;
            const uintptr_t a = (uintptr_t)src;
            const uint64_t* al = reinterpret_cast<const uint64_t*>(a & ~(uintptr_t)7);
            const uint32_t sh = (uint32_t)(a & 7) * 8;
            uint64_t v;
            if (reinterpret_cast<const uint8_t*>(al + 2) <= wire_end) {
                v = __ldg(al);
                if (sh) v = (v >> sh) | (__ldg(al + 1) << (64 - sh));
            } else {  // the last words of the buffer: no read past its end
                v = load_le<uint64_t>(src);
            }
            __stcs(out + i, v);
        }
    }
}

// Tally straight from the wire bytes: column pair (2 words) per thread, one slab of records per blockIdx.y, raw words
// of the ACCEPTED records summed in 128-bit counters and reduced once (the words of batch_add / tally_votes for any
// grouping, see tally.cu).  The payload is never 8-byte aligned: the thread reads the three aligned words that cover
// its 16 bytes and funnel-shifts by the record's misalignment (uniform across the block).  width = choices * 2N.
constexpr int WTALLY_THREADS = 256;
constexpr int WTALLY_UNROLL = 4;

__global__ void __launch_bounds__(WTALLY_THREADS) tally_wire_kernel(const uint8_t* __restrict__ wire, const uint8_t* wire_end,
                                                                    const uint64_t* __restrict__ offsets, const uint8_t* __restrict__ status,
                                                                    size_t count, size_t per_slab, uint32_t words_per_choice,
                                                                    uint32_t width, uint64_t* __restrict__ partial, const ModQ m) {
    const uint32_t col = (blockIdx.x * WTALLY_THREADS + threadIdx.x) * 2;  // words_per_choice is even: a pair never straddles choices
    if (col >= width) return;
    const uint32_t c = col / words_per_choice, j = col - c * words_per_choice;
    const size_t rel = WIRE_HEADER_BYTES + 12 + (size_t)c * (12 + 8 * (size_t)words_per_choice) + 12 + 8 * (size_t)j;
    const size_t first = (size_t)blockIdx.y * per_slab;
    size_t last = first + per_slab;
    if (last > count) last = count;
    uint64_t lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
    auto fetch = [&](size_t r, uint64_t (&w)[3], uint32_t& sh, bool& ok) {
        ok = status[r] == FHEB_WIRE_OK;
        sh = 0;
        w[0] = w[1] = w[2] = 0;
        if (!ok) return;
        const uint8_t* src = wire + offsets[r] + rel;
        const uintptr_t a = (uintptr_t)src;
        const uint64_t* al = reinterpret_cast<const uint64_t*>(a & ~(uintptr_t)7);
        sh = (uint32_t)(a & 7) * 8;
        if (reinterpret_cast<const uint8_t*>(al + 3) <= wire_end) {
            w[0] = __ldcs(al);
            w[1] = __ldcs(al + 1);
            w[2] = sh ? __ldcs(al + 2) : 0;
        } else {  // the very end of the buffer: no read past it
            w[0] = load_le<uint64_t>(src);
            w[1] = load_le<uint64_t>(src + 8);
            sh = 0;
        }
    };
    auto add = [&](const uint64_t (&w)[3], uint32_t sh, bool ok) {
        if (!ok) return;
        const uint64_t v0 = sh ? (w[0] >> sh) | (w[1] << (64 - sh)) : w[0];
        const uint64_t v1 = sh ? (w[1] >> sh) | (w[2] << (64 - sh)) : w[1];
        acc128(lo0, hi0, v0);
        acc128(lo1, hi1, v1);
    };
    size_t r = first;
    for (; r + WTALLY_UNROLL <= last; r += WTALLY_UNROLL) {
        uint64_t w[WTALLY_UNROLL][3];
        uint32_t sh[WTALLY_UNROLL];
        bool ok[WTALLY_UNROLL];
#pragma unroll
        for (int u = 0; u < WTALLY_UNROLL; ++u) fetch(r + u, w[u], sh[u], ok[u]);
#pragma unroll
        for (int u = 0; u < WTALLY_UNROLL; ++u) add(w[u], sh[u], ok[u]);
    }
    for (; r < last; ++r) {
        uint64_t w[3];
        uint32_t sh;
        bool ok;
        fetch(r, w, sh, ok);
        add(w, sh, ok);
    }
    uint64_t* out = partial + (size_t)blockIdx.y * width + col;
    out[0] = fold128(hi0, lo0, m);
    out[1] = fold128(hi1, lo1, m);
}

int ballots_validate_device(const uint8_t* wire, const uint64_t* offsets, size_t count, uint32_t choices, uint32_t N, uint64_t q,
                            uint8_t* status, uint64_t* timestamps, cudaStream_t s) {
    // FHEB_WIRE_WARMUP: test knob - 0 makes every start-state guess wrong, which exercises the repair rounds
    static const uint32_t warmup = [] {
        const char* e = getenv("FHEB_WIRE_WARMUP");
        return e ? (uint32_t)atoi(e) : WIRE_WARMUP;
    }();
    const size_t per_block = WIRE_THREADS / 32;
    const size_t want = (count + per_block - 1) / per_block;
    int bps = 0;
    FHEB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, ballot_validate_kernel, WIRE_THREADS, WIRE_TAB_BYTES));
    const size_t resident = (size_t)ctx().sm_count * (size_t)(bps > 0 ? bps : 1);
    ballot_validate_kernel<<<(unsigned)(want < resident ? want : resident), WIRE_THREADS, WIRE_TAB_BYTES, s>>>(wire, offsets, count, choices, N, q,
                                                                                                  status, timestamps, warmup);
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

int ballots_ingest_device(const uint8_t* wire, size_t wire_bytes, const uint64_t* offsets, size_t count, uint32_t choices, uint32_t N, uint64_t q,
                          uint64_t* cts, uint8_t* status, uint64_t* timestamps, cudaStream_t s) {
    FHEB_TRY(ballots_validate_device(wire, offsets, count, choices, N, q, status, timestamps, s));
    const size_t cap = (size_t)ctx().sm_count * 8;
    ballot_unpack_kernel<<<(unsigned)(count < cap ? count : cap), 256, 0, s>>>(wire, wire + wire_bytes, offsets, status, count, choices, N, cts);
    FHEB_CHECK_LAUNCH();
    count_launch();
    return FHEB_OK;
}

// ---- host: key containers ---------------------------------------------------------------------
struct Cursor {
    const uint8_t* p;
    size_t left;
    bool ok = true;
    template <class T>
    T get() {
        if (left < sizeof(T)) {
            ok = false;
            left = 0;
            return 0;
        }
        T v = load_le<T>(p);
        p += sizeof(T);
        left -= sizeof(T);
        return v;
    }
    // read_polynomial (:128-141): degree u32 (must equal the header's), then degree words
    int polynomial(uint32_t degree, uint64_t* dst) {
        const uint32_t d = get<uint32_t>();
        if (!ok) return set_error(FHEB_ERR_INVALID_PARAMETERS, "key container payload is truncated");
        if (d != degree) return set_error(FHEB_ERR_INVALID_PARAMETERS, "Polynomial degree mismatch");
        if (left < (size_t)degree * 8) return set_error(FHEB_ERR_INVALID_PARAMETERS, "key container payload is truncated");
        std::memcpy(dst, p, (size_t)degree * 8);  // little-endian host, as the reference assumes
        p += (size_t)degree * 8;
        left -= (size_t)degree * 8;
        return FHEB_OK;
    }
};

// common front of deserialize_eval_key / deserialize_bootstrap_key (:421-437, :552-568): header, magic, checksum
static int open_container(const uint8_t* bytes, size_t len, uint32_t magic, fheb_wire_header* h, Cursor* cur) {
    FHEB_REQUIRE(bytes != nullptr, "bytes must not be null");
    if (len < WIRE_HEADER_BYTES) return set_error(FHEB_ERR_INVALID_PARAMETERS, "Failed to read header");
    parse_header(bytes, h);
    if (h->magic != magic) return set_error(FHEB_ERR_INVALID_PARAMETERS, "Invalid magic bytes");
    const size_t avail = len - WIRE_HEADER_BYTES;
    const size_t n = h->data_size < avail ? h->data_size : avail;
    // verify_checksum (:76-94): NONE and unknown types pass; CRC32 and "SHA256" both use compute_crc32
    if (h->checksum_type == 1 || h->checksum_type == 2) {
        if (crc32_host(bytes + WIRE_HEADER_BYTES, n, h->data_size - n) != h->checksum)
            return set_error(FHEB_ERR_INVALID_PARAMETERS, "Checksum verification failed");
    }
    cur->p = bytes + WIRE_HEADER_BYTES;
    cur->left = n;
    return FHEB_OK;
}

}  // namespace fheb

using namespace fheb;

extern "C" {

uint32_t fheb_wire_crc32(const void* data, size_t len) { return crc32_host(static_cast<const uint8_t*>(data), data ? len : 0); }

int fheb_wire_header_read(const void* bytes, size_t len, fheb_wire_header* out) {
    FHEB_REQUIRE(bytes != nullptr && out != nullptr, "bytes and out must not be null");
    if (len < WIRE_HEADER_BYTES) return set_error(FHEB_ERR_INVALID_PARAMETERS, "Failed to read header");
    parse_header(static_cast<const uint8_t*>(bytes), out);
    return FHEB_OK;
}

size_t fheb_ballot_wire_size(uint32_t num_choices, uint32_t degree) {
    return WIRE_HEADER_BYTES + 12 + (size_t)num_choices * (12 + 16 * (size_t)degree);
}

int fheb_ballot_serialize(const uint64_t* choices, uint32_t num_choices, uint32_t degree, uint64_t modulus, uint64_t timestamp,
                          void* out, size_t capacity, size_t* written) {
    // BallotSerializer::serialize_ballot, :709-774.  Host-side packing of ONE record (e.g. a tally result).
    FHEB_REQUIRE(out != nullptr && written != nullptr, "out and written must not be null");
    FHEB_REQUIRE(num_choices == 0 || choices != nullptr, "choices must not be null");
    FHEB_REQUIRE(!is_device_pointer(choices) && !is_device_pointer(out), "fheb_ballot_serialize packs host buffers");
    const size_t total = fheb_ballot_wire_size(num_choices, degree);
    *written = total;
    FHEB_REQUIRE(total <= capacity, "output buffer too small: %zu bytes needed", total);
    FHEB_REQUIRE(total - WIRE_HEADER_BYTES <= 0xFFFFFFFFull, "ballot payload exceeds the 32-bit data_size field");
    uint8_t* o = static_cast<uint8_t*>(out);
    uint8_t* d = o + WIRE_HEADER_BYTES;
    store_le<uint64_t>(d, timestamp);
    store_le<uint32_t>(d + 8, num_choices);
    uint8_t* p = d + 12;
    for (uint32_t c = 0; c < num_choices; ++c) {
        store_le<uint32_t>(p, degree);
        store_le<uint64_t>(p + 4, modulus);
        std::memcpy(p + 12, choices + (size_t)c * 2 * degree, (size_t)degree * 16);
        p += 12 + (size_t)degree * 16;
    }
    const uint32_t data_size = (uint32_t)(total - WIRE_HEADER_BYTES);
    std::memset(o, 0, WIRE_HEADER_BYTES);
    store_le<uint32_t>(o, MAGIC_BALLOT);
    store_le<uint32_t>(o + 4, 1);  // SERIALIZATION_VERSION
    store_le<uint32_t>(o + 8, 4);  // key_type: ballot (:743)
    store_le<uint64_t>(o + 12, timestamp);  // the timestamp doubles as the id (:744)
    store_le<uint32_t>(o + 20, num_choices ? degree : 0);
    store_le<uint64_t>(o + 24, num_choices ? modulus : 0);
    store_le<uint32_t>(o + 32, data_size);
    o[36] = 1;  // ChecksumType::CRC32
    o[37] = 0;  // CompressionType::NONE
    store_le<uint32_t>(o + 45, crc32_host(d, data_size));
    return FHEB_OK;
}

int fheb_ballots_ingest(const void* wire, size_t wire_bytes, const uint64_t* offsets, size_t count, uint32_t num_choices,
                        uint32_t degree, uint64_t modulus, uint64_t* cts, uint8_t* status, uint64_t* timestamps,
                        size_t* accepted, void* stream) {
    FHEB_TRY(ensure_ready());
    if (accepted) *accepted = 0;
    if (count == 0) return FHEB_OK;
    FHEB_REQUIRE(wire != nullptr && cts != nullptr && status != nullptr, "wire, cts and status must not be null");
    FHEB_REQUIRE(degree >= 1 && num_choices >= 1, "degree and num_choices must be positive");
    FHEB_REQUIRE(!is_device_pointer(status) && (timestamps == nullptr || !is_device_pointer(timestamps)) &&
                     (offsets == nullptr || !is_device_pointer(offsets)),
                 "status, timestamps and offsets are host arrays");
    const uint8_t* w = static_cast<const uint8_t*>(wire);
    const bool wire_on_device = is_device_pointer(wire);
    // record extents: caller supplied (count + 1 entries), or walked from the self-describing headers
    std::vector<uint64_t> walked;
    if (!offsets) {
        FHEB_REQUIRE(!wire_on_device, "offsets are required when the wire buffer is device memory");
        walked.resize(count + 1);
        size_t pos = 0;
        for (size_t r = 0; r < count; ++r) {
            walked[r] = pos;
            FHEB_REQUIRE(wire_bytes - pos >= WIRE_HEADER_BYTES, "wire buffer ends inside record %zu", r);
            const size_t next = pos + WIRE_HEADER_BYTES + load_le<uint32_t>(w + pos + 32);
            pos = next < wire_bytes ? next : wire_bytes;
        }
        walked[count] = pos;
        offsets = walked.data();
    }
    for (size_t r = 0; r < count; ++r)
        FHEB_REQUIRE(offsets[r] <= offsets[r + 1] && offsets[r + 1] <= wire_bytes, "record %zu lies outside the wire buffer", r);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t ct_bytes = count * (size_t)num_choices * 2 * degree * 8;
    Staged sc;
    FHEB_TRY(sc.bind(cts, ct_bytes, false, true, s));
    uint8_t* d_wire = nullptr;
    if (wire_on_device) {
        FHEB_REQUIRE(((uintptr_t)w & 7) == 0, "a device wire buffer must be 8-byte aligned");
        d_wire = const_cast<uint8_t*>(w);
    } else {
        FHEB_CUDA(cudaMallocAsync(&d_wire, wire_bytes, s));
        const cudaError_t e = cudaMemcpyAsync(d_wire, w, wire_bytes, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) {  // do not leave the staging copy of the wire behind
            cudaFreeAsync(d_wire, s);
            return set_error(FHEB_ERR_NATIVE, "copy of the wire bytes failed: %s", cudaGetErrorString(e));
        }
    }
    uint64_t* d_off = nullptr;
    uint8_t* d_status = nullptr;
    uint64_t* d_ts = nullptr;
    int rc = FHEB_OK;
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == FHEB_OK) rc = set_error(FHEB_ERR_NATIVE, "%s failed: %s", what, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    cuda_ok(cudaMallocAsync(&d_off, (count + 1) * 8, s), "cudaMallocAsync");
    cuda_ok(cudaMallocAsync(&d_status, count, s), "cudaMallocAsync");
    if (timestamps) cuda_ok(cudaMallocAsync(&d_ts, count * 8, s), "cudaMallocAsync");
    if (rc == FHEB_OK) cuda_ok(cudaMemcpyAsync(d_off, offsets, (count + 1) * 8, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync");
    if (rc == FHEB_OK)
        rc = ballots_ingest_device(d_wire, wire_bytes, d_off, count, num_choices, degree, modulus, sc.ptr<uint64_t>(), d_status, d_ts, s);
    if (rc == FHEB_OK) cuda_ok(cudaMemcpyAsync(status, d_status, count, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync");
    if (rc == FHEB_OK && timestamps) cuda_ok(cudaMemcpyAsync(timestamps, d_ts, count * 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync");
    if (rc == FHEB_OK) rc = sc.finish();
    if (rc == FHEB_OK) cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize");  // status is a host array: the call completes here
    if (d_off) cudaFreeAsync(d_off, s);
    if (d_status) cudaFreeAsync(d_status, s);
    if (d_ts) cudaFreeAsync(d_ts, s);
    if (!wire_on_device && d_wire) cudaFreeAsync(d_wire, s);
    FHEB_TRY(rc);
    if (accepted) {
        size_t ok = 0;
        for (size_t r = 0; r < count; ++r) ok += status[r] == FHEB_WIRE_OK;
        *accepted = ok;
    }
    return FHEB_OK;
}

int fheb_tally_wire(const void* wire, size_t wire_bytes, const uint64_t* offsets, size_t count, uint32_t num_choices,
                    uint32_t degree, uint64_t modulus, uint64_t* out, uint8_t* status, size_t* accepted, void* stream) {
    // BallotSerializer::deserialize_ballot per record (:776-846) + EncryptionEngine::tally_votes per choice
    // (encryption.cpp:1061-1067,1327-1364) without materialising the ciphertexts: validate, then sum the accepted
    // records' words straight out of the wire bytes.
    FHEB_TRY(ensure_ready());
    if (accepted) *accepted = 0;
    FHEB_REQUIRE(count != 0, "Cannot add empty vector of ciphertexts");  // encryption.cpp:1328-1330
    FHEB_REQUIRE(wire != nullptr && out != nullptr && status != nullptr, "wire, out and status must not be null");
    FHEB_REQUIRE(degree >= 1 && num_choices >= 1, "degree and num_choices must be positive");
    FHEB_REQUIRE(modulus >= 2, "Modulus must be at least 2");
    FHEB_REQUIRE(!is_device_pointer(status) && (offsets == nullptr || !is_device_pointer(offsets)), "status and offsets are host arrays");
    const uint8_t* w = static_cast<const uint8_t*>(wire);
    const bool wire_on_device = is_device_pointer(wire);
    std::vector<uint64_t> walked;
    if (!offsets) {
        FHEB_REQUIRE(!wire_on_device, "offsets are required when the wire buffer is device memory");
        walked.resize(count + 1);
        size_t pos = 0;
        for (size_t r = 0; r < count; ++r) {
            walked[r] = pos;
            FHEB_REQUIRE(wire_bytes - pos >= WIRE_HEADER_BYTES, "wire buffer ends inside record %zu", r);
            const size_t next = pos + WIRE_HEADER_BYTES + load_le<uint32_t>(w + pos + 32);
            pos = next < wire_bytes ? next : wire_bytes;
        }
        walked[count] = pos;
        offsets = walked.data();
    }
    for (size_t r = 0; r < count; ++r)
        FHEB_REQUIRE(offsets[r] <= offsets[r + 1] && offsets[r + 1] <= wire_bytes, "record %zu lies outside the wire buffer", r);
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t wpc = 2 * degree, width = num_choices * wpc;
    Staged so;
    FHEB_TRY(so.bind(out, (size_t)width * 8, false, true, s));
    uint8_t* d_wire = nullptr;
    uint64_t* d_off = nullptr;
    uint8_t* d_status = nullptr;
    uint64_t* partial = nullptr;
    int rc = FHEB_OK;
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == FHEB_OK) rc = set_error(FHEB_ERR_NATIVE, "%s failed: %s", what, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    if (wire_on_device) {
        if (((uintptr_t)w & 7) != 0) rc = set_error(FHEB_ERR_INVALID_PARAMETERS, "a device wire buffer must be 8-byte aligned");
        d_wire = const_cast<uint8_t*>(w);
    } else if (cuda_ok(cudaMallocAsync(&d_wire, wire_bytes, s), "cudaMallocAsync")) {
        cuda_ok(cudaMemcpyAsync(d_wire, w, wire_bytes, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync");
    }
    if (rc == FHEB_OK) cuda_ok(cudaMallocAsync(&d_off, (count + 1) * 8, s), "cudaMallocAsync");
    if (rc == FHEB_OK) cuda_ok(cudaMallocAsync(&d_status, count, s), "cudaMallocAsync");
    if (rc == FHEB_OK) cuda_ok(cudaMemcpyAsync(d_off, offsets, (count + 1) * 8, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync");
    if (rc == FHEB_OK) rc = ballots_validate_device(d_wire, d_off, count, num_choices, degree, modulus, d_status, nullptr, s);
    if (rc == FHEB_OK) cuda_ok(cudaMemcpyAsync(status, d_status, count, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync");
    if (rc == FHEB_OK) cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize");  // the number of accepted records picks the path
    size_t ok = 0, lone = 0;
    if (rc == FHEB_OK) {
        for (size_t r = 0; r < count; ++r)
            if (status[r] == FHEB_WIRE_OK) {
                ++ok;
                lone = r;
            }
        if (accepted) *accepted = ok;
        if (ok == 0) rc = set_error(FHEB_ERR_INVALID_PARAMETERS, "Cannot add empty vector of ciphertexts");
    }
    if (rc == FHEB_OK && ok == 1) {
        // a single ciphertext is returned untouched, unreduced words included (encryption.cpp:1332-1334)
        const size_t cap = 1;
        ballot_unpack_kernel<<<(unsigned)cap, 256, 0, s>>>(d_wire, d_wire + wire_bytes, d_off + lone, d_status + lone, 1, num_choices, degree,
                                                           so.ptr<uint64_t>());
        if (cudaGetLastError() != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "ballot_unpack_kernel launch failed");
        count_launch();
    } else if (rc == FHEB_OK) {
        const unsigned chunks = (width + 2 * WTALLY_THREADS - 1) / (2 * WTALLY_THREADS);
        size_t slabs = ((size_t)ctx().sm_count * 4 + chunks - 1) / chunks;
        const size_t max_slabs = (count + 63) / 64;
        if (slabs > max_slabs) slabs = max_slabs;
        if (slabs < 1) slabs = 1;
        const size_t per_slab = (count + slabs - 1) / slabs;
        slabs = (count + per_slab - 1) / per_slab;
        if (cuda_ok(cudaMallocAsync(&partial, slabs * (size_t)width * 8, s), "cudaMallocAsync")) {
            tally_wire_kernel<<<dim3(chunks, (unsigned)slabs), WTALLY_THREADS, 0, s>>>(d_wire, d_wire + wire_bytes, d_off, d_status, count, per_slab,
                                                                                 wpc, width, partial, make_modq(modulus));
            if (cudaGetLastError() != cudaSuccess) rc = set_error(FHEB_ERR_NATIVE, "tally_wire_kernel launch failed");
            count_launch();
            // fold the slab partials; with two or more accepted ballots every word is the output of a modular addition
            if (rc == FHEB_OK) rc = tally_rows_device(partial, slabs, width, modulus, so.ptr<uint64_t>(), s);
        }
    }
    if (rc == FHEB_OK) rc = so.finish();
    if (rc == FHEB_OK && so.staged()) cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    if (partial) cudaFreeAsync(partial, s);
    if (d_off) cudaFreeAsync(d_off, s);
    if (d_status) cudaFreeAsync(d_status, s);
    if (!wire_on_device && d_wire) cudaFreeAsync(d_wire, s);
    return rc;
}

int fheb_relin_key_from_wire(const fheb_ntt_plan* plan, const void* bytes, size_t len, fheb_relin_key** out) {
    // KeySerializer::deserialize_eval_key, :414-466, then fheb_relin_key_create on the pairs
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(plan != nullptr && out != nullptr, "plan and out must not be null");
    const NttPlan* p = reinterpret_cast<const NttPlan*>(plan);
    fheb_wire_header h;
    Cursor cur{nullptr, 0};
    FHEB_TRY(open_container(static_cast<const uint8_t*>(bytes), len, MAGIC_EVAL, &h, &cur));
    const uint32_t base_log = cur.get<uint32_t>();
    const uint32_t level = cur.get<uint32_t>();
    const uint32_t num_keys = cur.get<uint32_t>();
    FHEB_REQUIRE(cur.ok, "key container payload is truncated");
    if (num_keys) {
        FHEB_REQUIRE(h.poly_degree == p->degree, "Polynomial degree mismatch");
        FHEB_REQUIRE(h.modulus == p->modulus, "Polynomial modulus mismatch");
    }
    FHEB_REQUIRE((uint64_t)num_keys * 2 * (4 + 8ull * p->degree) <= cur.left, "key container payload is truncated");
    std::vector<uint64_t> keys((size_t)num_keys * 2 * p->degree);
    for (uint32_t i = 0; i < num_keys * 2; ++i) FHEB_TRY(cur.polynomial(p->degree, keys.data() + (size_t)i * p->degree));
    return fheb_relin_key_create(plan, keys.data(), num_keys, base_log, level, h.key_id, out);
}

int fheb_boot_key_from_wire(const fheb_ntt_plan* plan, const fheb_boot_params* params, const void* bytes, size_t len,
                            fheb_boot_key** out) {
    // KeySerializer::deserialize_bootstrap_key, :545-615.  The container holds bsk[lwe_dimension][row_size] pairs of
    // polynomials: for glwe_dimension 1 that is exactly the GGSW layout [n][(k+1)*L rows][k+1][N] fheb_boot_key_create
    // takes.  The KeyManager-style key-switching pairs that follow are skipped: BootstrapEngine::key_switch consumes
    // a different structure (fheb_boot_key_set_ksk), and the reference has no conversion between the two.
    FHEB_TRY(ensure_ready());
    FHEB_REQUIRE(plan != nullptr && params != nullptr && out != nullptr, "plan, params and out must not be null");
    FHEB_REQUIRE(params->glwe_dimension == 1, "the FHEB container stores polynomial pairs: glwe_dimension must be 1");
    const NttPlan* p = reinterpret_cast<const NttPlan*>(plan);
    fheb_wire_header h;
    Cursor cur{nullptr, 0};
    FHEB_TRY(open_container(static_cast<const uint8_t*>(bytes), len, MAGIC_BOOT, &h, &cur));
    const uint32_t lwe_dimension = cur.get<uint32_t>();
    const uint32_t bsk_size = cur.get<uint32_t>();
    FHEB_REQUIRE(cur.ok, "key container payload is truncated");
    FHEB_REQUIRE(lwe_dimension == params->lwe_dimension && bsk_size == params->lwe_dimension,
                 "container holds %u GGSW ciphertexts for LWE dimension %u, expected %u", bsk_size, lwe_dimension, params->lwe_dimension);
    FHEB_REQUIRE(h.poly_degree == p->degree, "Polynomial degree mismatch");
    FHEB_REQUIRE(h.modulus == p->modulus, "Polynomial modulus mismatch");
    const uint32_t rows = 2 * params->decomp_level;
    const size_t N = p->degree;
    FHEB_REQUIRE((uint64_t)bsk_size * (4 + (uint64_t)rows * 2 * (4 + 8 * N)) <= cur.left, "key container payload is truncated");
    std::vector<uint64_t> bsk((size_t)bsk_size * rows * 2 * N);
    for (uint32_t i = 0; i < bsk_size; ++i) {
        const uint32_t row_size = cur.get<uint32_t>();
        FHEB_REQUIRE(cur.ok && row_size == rows, "GGSW %u has %u rows, expected (k+1)*L = %u", i, row_size, rows);
        for (uint32_t j = 0; j < rows * 2; ++j) FHEB_TRY(cur.polynomial(p->degree, bsk.data() + ((size_t)i * rows * 2 + j) * N));
    }
    return fheb_boot_key_create(plan, params, bsk.data(), out);
}

}  // extern "C"
