/*
 * fheb200.h - C ABI of the B200 (sm_100a) backend for node-fhe-accelerate's data-parallel
 * hot path: batched forward/inverse transform, NTT-domain polynomial multiplication,
 * element-wise and multi-limb modular arithmetic, the TFHE blind-rotation chain and the
 * encrypted-ballot tally.
 *
 * This is the drop-in boundary: plain pointers, sizes and opaque handles, no C++ or torch
 * types.  Each entry point names the reference interface it replaces (paths relative to
 * the reference repository).  The reference's N-API addon (src/native/lib.rs,
 * src/native/bridge.rs) exposes only scalar ModularArithmetic today; the bulk entry points
 * below are what its cxx bridge would bind for the classes in cpp/include (see
 * INTEGRATION.md for the binding stubs).
 *
 * Conventions
 *  - All polynomial data are unsigned 64-bit words, row-major, [batch][N].
 *  - Every data pointer may be a DEVICE pointer (cudaMalloc / torch CUDA tensor) or a HOST
 *    pointer (pinned or pageable).  Host buffers are staged through device memory inside
 *    the call (chunked, copy/compute overlapped); device buffers are used in place.
 *  - `stream` is a cudaStream_t passed as void* (NULL = the library's own stream).  Calls
 *    on device buffers are asynchronous with respect to the host; calls on host buffers
 *    return after the results are in the host buffer.
 *  - Results are bit-identical to the reference's portable scalar C++ path.
 *  - There is NO CPU fallback: without an sm_100 device every call fails with
 *    FHEB_ERR_HARDWARE_UNAVAILABLE.
 *  - Return value: 0 on success, otherwise an fheb_status; the message is available from
 *    fheb_last_error() (thread-local).  No C++ exception crosses this boundary; the codes
 *    line up with FHEErrorCode (src/api/types.ts:140-151) and the message text follows the
 *    reference's std::invalid_argument messages where one exists.
 *  - Handles are thread-compatible (do not use one handle from two threads at once),
 *    matching the reference's lock-free objects.
 */
#ifndef FHEB200_H
#define FHEB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FHEB_API __declspec(dllexport)
#else
#define FHEB_API __attribute__((visibility("default")))
#endif

typedef enum fheb_status {
    FHEB_OK = 0,
    FHEB_ERR_INVALID_PARAMETERS = 1,   /* FHEErrorCode.INVALID_PARAMETERS; std::invalid_argument */
    FHEB_ERR_KEY_MISMATCH = 2,         /* FHEErrorCode.KEY_MISMATCH                               */
    FHEB_ERR_HARDWARE_UNAVAILABLE = 3, /* FHEErrorCode.HARDWARE_UNAVAILABLE: no sm_100 device     */
    FHEB_ERR_NATIVE = 4,               /* FHEErrorCode.NATIVE_ERROR: CUDA runtime failure         */
    FHEB_ERR_OUT_OF_MEMORY = 5
} fheb_status;

typedef struct fheb_ntt_plan fheb_ntt_plan; /* replaces NTTProcessor / PolynomialRing state   */
typedef struct fheb_boot_key fheb_boot_key; /* replaces ExtendedBootstrapKey (device resident) */
typedef struct fheb_relin_key fheb_relin_key; /* replaces EvaluationKey.relin_key (device resident, pre-transformed) */

/* ---- library / device ---------------------------------------------------------------- */

/* replaces initialize(): src/native/lib.rs:23-30.  device < 0 selects the current device. */
FHEB_API int fheb_init(int device);
FHEB_API int fheb_shutdown(void);
/* One process, several GPUs (the reference's addon is a single Node process: src/native/lib.rs:23-133).  After
 * fheb_set_devices every entry point whose batch lives in HOST memory (transforms, products, blind rotation,
 * bootstrap, tally) splits the batch into contiguous shares, one per device, each copied and computed on its own
 * GPU over its own PCIe link by its own host thread; plans and keys are replicated on a device on first use.
 * Calls on DEVICE buffers run where the buffers live (make that device current first).  count < 0 selects every
 * visible device; devices == NULL selects 0..count-1; count == 0 or 1 turns spreading off.
 * fheb_get_devices returns the number configured (and fills out[0..capacity)). */
FHEB_API int fheb_device_count(void); /* visible sm_100 devices (0 when there is none) */
FHEB_API int fheb_set_devices(const int* devices, int count);
FHEB_API int fheb_get_devices(int* out, int capacity);
/* replaces version(): src/native/lib.rs:129-133 */
FHEB_API const char* fheb_version(void);
FHEB_API const char* fheb_last_error(void);

/* replaces detect_hardware() / HardwareCapabilities: src/native/lib.rs:32-42,122-127,
 * cpp/include/fhe_types.h:19-26.  The Apple fields are reported truthfully as absent. */
typedef struct fheb_device_info {
    int32_t has_sme, has_metal, has_neon, has_amx; /* always 0 on this backend            */
    int32_t has_cuda;                              /* 1                                    */
    int32_t cc_major, cc_minor;                    /* 10, 0 on B200                        */
    int32_t sm_count;                              /* 148 on B200 (metal_gpu_cores analogue) */
    uint64_t device_memory_bytes;                  /* unified_memory_size analogue         */
    uint64_t l2_bytes;
    uint64_t smem_per_block_optin;
    char name[128];
} fheb_device_info;
FHEB_API int fheb_device_info_get(fheb_device_info* out);

/* The scalar class the reference's addon exports today: ModularArithmetic::{new, montgomery_mul, mod_add,
 * mod_sub, to_montgomery, from_montgomery, get_modulus} (src/native/lib.rs:44-120, cpp/src/modular_arithmetic.cpp:
 * 52-165).  One scalar per call: host-side integer code with no GPU value, kept so that the
 * addon's existing surface has a home behind this boundary.  It repeats the reference word for word,
 * INCLUDING its Montgomery constant q_inv = -(q^-1 mod (2^64 - 1)) (mathematically not a Montgomery
 * constant) with AArch64 division semantics (x / 0 = 0, x % 0 = x).  Not used by any kernel. */
typedef struct fheb_modarith fheb_modarith;
FHEB_API int fheb_modarith_create(uint64_t modulus, fheb_modarith** out);
FHEB_API int fheb_modarith_destroy(fheb_modarith* m);
FHEB_API uint64_t fheb_modarith_montgomery_mul(const fheb_modarith* m, uint64_t a, uint64_t b);
FHEB_API uint64_t fheb_modarith_mod_add(const fheb_modarith* m, uint64_t a, uint64_t b);
FHEB_API uint64_t fheb_modarith_mod_sub(const fheb_modarith* m, uint64_t a, uint64_t b);
FHEB_API uint64_t fheb_modarith_to_montgomery(const fheb_modarith* m, uint64_t a);
FHEB_API uint64_t fheb_modarith_from_montgomery(const fheb_modarith* m, uint64_t a);
FHEB_API uint64_t fheb_modarith_get_modulus(const fheb_modarith* m);

/* Buffer helpers; replace MetalComputeContext::create_buffer / release_buffer /
 * copy_to_buffer / copy_from_buffer / synchronize: cpp/include/metal_compute.h:46-50,79. */
FHEB_API int fheb_device_alloc(void** out, size_t bytes);
FHEB_API int fheb_device_free(void* p);
FHEB_API int fheb_host_alloc(void** out, size_t bytes); /* pinned host memory */
FHEB_API int fheb_host_free(void* p);
FHEB_API int fheb_copy(void* dst, const void* src, size_t bytes, void* stream); /* any direction */
FHEB_API int fheb_synchronize(void* stream);

/* ---- transform plans -------------------------------------------------------------------- */

/* replaces NTTProcessor::NTTProcessor + precompute_twiddles: cpp/src/ntt_processor.cpp:
 * 134-160,168-208.  Finds the same primitive 2N-th root as find_primitive_root (:92-128:
 * smallest g >= 2) and builds the same tables.  Errors mirror the constructor's
 * (":141-153": degree not a power of two / outside [4, 65536] / even modulus; ":103-105":
 * modulus not NTT-friendly).  Supported on this backend: the reference's full range N = 4 .. 65536 (one
 * launch up to 16384, two launches above), q < 2^62; bootstrap kernels N = 32 .. 4096. */
FHEB_API int fheb_ntt_plan_create(uint32_t degree, uint64_t modulus, fheb_ntt_plan** out);

/* Caller-supplied tables, the shape of MetalComputeContext::batch_ntt_forward(..., twiddles)
 * (cpp/include/metal_compute.h:66-73) and fast_ntt_forward / fast_ntt_inverse
 * (cpp/include/adaptive_dispatcher.h:55-71): fwd_table[i], inv_table[i], i < N, in natural
 * exponent order exactly as TwiddleFactors holds them (cpp/include/ntt_processor.h:29-41);
 * the network indexes them as table[j * N / (2m)] (cpp/src/ntt_processor.cpp:286). */
FHEB_API int fheb_ntt_plan_create_with_tables(uint32_t degree, uint64_t modulus, const uint64_t* fwd_table,
                                              const uint64_t* inv_table, uint64_t inv_n, fheb_ntt_plan** out);
FHEB_API int fheb_ntt_plan_destroy(fheb_ntt_plan* plan);

/* replaces NTTProcessor::get_twiddles(): cpp/include/ntt_processor.h:84.  Any pointer may
 * be NULL.  scalars = {primitive_root, inv_primitive_root, inv_n} (0,0,inv_n for plans made
 * from caller tables). */
FHEB_API int fheb_ntt_plan_get_tables(const fheb_ntt_plan* plan, uint64_t* fwd_table, uint64_t* inv_table,
                                      uint64_t scalars[3]);
FHEB_API uint32_t fheb_ntt_plan_degree(const fheb_ntt_plan* plan);   /* NTTProcessor::get_degree  */
FHEB_API uint64_t fheb_ntt_plan_modulus(const fheb_ntt_plan* plan);  /* NTTProcessor::get_modulus */

/* ---- transforms ---------------------------------------------------------------------------- */

/* replace NTTProcessor::forward_ntt / inverse_ntt (in-place and out-of-place) and
 * forward_ntt_batch / inverse_ntt_batch: cpp/src/ntt_processor.cpp:262-319,325-388,394-408;
 * MetalComputeContext::batch_ntt_forward / batch_ntt_inverse: cpp/include/metal_compute.h:66-73.
 * in == out is allowed.  Unreduced input words are reduced first, as mod_add/mod_sub do
 * (cpp/src/modular_arithmetic.cpp:124-125,140-141); outputs are canonical. */
FHEB_API int fheb_ntt_forward_batch(const fheb_ntt_plan* plan, const uint64_t* in, uint64_t* out, size_t batch,
                                    void* stream);
FHEB_API int fheb_ntt_inverse_batch(const fheb_ntt_plan* plan, const uint64_t* in, uint64_t* out, size_t batch,
                                    void* stream);
/* replaces fast_ntt_inverse: cpp/src/adaptive_dispatcher.cpp:171-205 - the FORWARD network fed
 * the inverse table, then scaling by N^-1.  Inputs must be canonical (< q). */
FHEB_API int fheb_ntt_inverse_fwdnet_batch(const fheb_ntt_plan* plan, const uint64_t* in, uint64_t* out,
                                           size_t batch, void* stream);

/* ---- polynomial ring --------------------------------------------------------------------- */

/* replaces PolynomialRing::multiply on coefficient-form operands: cpp/src/polynomial_ring.cpp:
 * 421-447, and MetalComputeContext::batch_poly_mul: cpp/include/metal_compute.h:75-78.
 * c[i] = T^-1(T(a[i]) . T(b[i])), one fused launch. */
FHEB_API int fheb_polymul_batch(const fheb_ntt_plan* plan, const uint64_t* a, const uint64_t* b, uint64_t* c,
                                size_t batch, void* stream);

/* Element-wise, count words, modulus any non-zero 64-bit value; r may alias a or b
 * (cpp/include/adaptive_dispatcher.h:30).  Replace PolynomialRing::add/subtract/negate/
 * multiply_scalar/pointwise_multiply (cpp/src/polynomial_ring.cpp:263-366,454-562),
 * MetalComputeContext::batch_modmul / batch_modadd (cpp/include/metal_compute.h:57-64),
 * fast_modmul_batch (cpp/src/adaptive_dispatcher.cpp:57-80) and the modmul/modadd/modsub/
 * modneg shader kernels (cpp/shaders/modular).  Semantics are the reference's scalar ones:
 * add/sub reduce their inputs first; mul is (a*b) % q on the raw words; neg is
 * (a == 0 ? 0 : q - a) WITHOUT input reduction. */
FHEB_API int fheb_modadd_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint64_t modulus, void* stream);
FHEB_API int fheb_modsub_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint64_t modulus, void* stream);
FHEB_API int fheb_modmul_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint64_t modulus, void* stream);
FHEB_API int fheb_modneg_batch(const uint64_t* a, uint64_t* r, size_t count, uint64_t modulus, void* stream);
FHEB_API int fheb_modmul_scalar_batch(const uint64_t* a, uint64_t scalar, uint64_t* r, size_t count, uint64_t modulus, void* stream);

/* ---- multi-limb Montgomery ------------------------------------------------------------------ */

/* replace MultiLimbModularArithmetic::montgomery_mul_neon / mod_add_neon / mod_sub_neon:
 * cpp/src/modular_arithmetic.cpp:796-824 (scalar bodies :496-693).  Element-major
 * [count][limbs], little-endian limbs, 1 <= limbs <= 8.  q_inv = -q[0]^-1 mod 2^64 as
 * MultiLimbMontgomeryConstants computes it (:347-358,484-485). */
FHEB_API int fheb_mlimb_montmul_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint32_t limbs,
                                      const uint64_t* q_limbs, uint64_t q_inv, void* stream);
FHEB_API int fheb_mlimb_add_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint32_t limbs,
                                  const uint64_t* q_limbs, void* stream);
FHEB_API int fheb_mlimb_sub_batch(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t count, uint32_t limbs,
                                  const uint64_t* q_limbs, void* stream);
/* replaces the MultiLimbMontgomeryConstants constructor (:471-486): consts = {q_inv,
 * R mod q [limbs], R^2 mod q [limbs]}.  Host-side integer set-up, no device work. */
FHEB_API int fheb_mlimb_constants(const uint64_t* q_limbs, uint32_t limbs, uint64_t* consts);

/* ---- TFHE bootstrap --------------------------------------------------------------------------- */

typedef struct fheb_boot_params { /* the TFHE fields of ParameterSet: cpp/include/parameter_set.h:70-120 */
    uint32_t lwe_dimension;   /* n                                  */
    uint32_t glwe_dimension;  /* k                                  */
    uint32_t decomp_base_log; /* gadget base log                    */
    uint32_t decomp_level;    /* gadget levels L                    */
} fheb_boot_params;

/* replaces ExtendedBootstrapKey.bsk (cpp/include/bootstrap_engine.h:139-158) as produced by
 * BootstrapEngine::generate_bootstrap_key / encrypt_ggsw (cpp/src/bootstrap_engine.cpp:
 * 268-364).  bsk = [n][(k+1)*L rows][k+1 polys (k mask, then body)][N] coefficient-form words
 * (host or device).  The key is transformed ONCE here and kept resident in HBM.  Supported: N = 32 .. 4096,
 * k = 1 .. 3, any level count whose working rows fit one SM's shared memory (all three TFHE presets of
 * cpp/src/parameter_set.cpp:108-190 do).  The plan must outlive the key. */
FHEB_API int fheb_boot_key_create(const fheb_ntt_plan* plan, const fheb_boot_params* params, const uint64_t* bsk,
                                  fheb_boot_key** out);
/* replaces ExtendedBootstrapKey.ksk / KeySwitchKey (cpp/include/key_manager.h:92-99):
 * ksk = [entries][n_out + 1] (a words then b), entries = k*N*ksk_level in generation order
 * (cpp/src/bootstrap_engine.cpp:391-421).  Optional: only key_switch/bootstrap need it. */
FHEB_API int fheb_boot_key_set_ksk(fheb_boot_key* key, const uint64_t* ksk, size_t entries, uint32_t n_out,
                                   uint32_t base_log, uint32_t level);
FHEB_API int fheb_boot_key_destroy(fheb_boot_key* key);

/* replaces BootstrapEngine::external_product / cmux with bsk[index]:
 * cpp/src/bootstrap_engine.cpp:431-518,520-540.  glwe, ct0, ct1, out = [batch][k+1][N]. */
FHEB_API int fheb_external_product_batch(const fheb_boot_key* key, uint32_t index, const uint64_t* glwe, uint64_t* out,
                                         size_t batch, void* stream);
FHEB_API int fheb_cmux_batch(const fheb_boot_key* key, uint32_t index, const uint64_t* ct0, const uint64_t* ct1,
                             uint64_t* out, size_t batch, void* stream);
/* replaces BootstrapEngine::blind_rotate on acc = (0,..,0,test_poly) as
 * bootstrap_with_test_poly sets it up: cpp/src/bootstrap_engine.cpp:547-577,692-699.
 * lwe = [batch][n+1] (a words then b); test_poly = [N] (canonical words);
 * out = [batch][k+1][N]. */
FHEB_API int fheb_blind_rotate_batch(const fheb_boot_key* key, const uint64_t* lwe, const uint64_t* test_poly,
                                     uint64_t* out, size_t batch, void* stream);
/* replaces BootstrapEngine::sample_extract: cpp/src/bootstrap_engine.cpp:594-624.
 * glwe = [batch][k+1][N]; out = [batch][k*N+1]. */
FHEB_API int fheb_sample_extract_batch(const fheb_boot_key* key, const uint64_t* glwe, uint64_t* out, size_t batch,
                                       void* stream);
/* replaces BootstrapEngine::key_switch: cpp/src/bootstrap_engine.cpp:626-669.
 * lwe = [batch][k*N+1]; out = [batch][n_out+1]. */
FHEB_API int fheb_key_switch_batch(const fheb_boot_key* key, const uint64_t* lwe, uint64_t* out, size_t batch,
                                   void* stream);
/* replaces BootstrapEngine::bootstrap_with_test_poly / programmable_bootstrap:
 * cpp/src/bootstrap_engine.cpp:684-723.  With a KSK: out = [batch][n_out+1]; without one the
 * chain stops after sample extraction: out = [batch][k*N+1]. */
FHEB_API int fheb_bootstrap_batch(const fheb_boot_key* key, const uint64_t* lwe, const uint64_t* test_poly,
                                  uint64_t* out, size_t batch, void* stream);

/* Host-side LUT builders; replace init_default_test_poly and create_lookup_table /
 * create_identity_lut / create_negation_lut / create_threshold_lut:
 * cpp/src/bootstrap_engine.cpp:57-77,725-779.  kind: 0 identity(arg0 = modulus),
 * 1 negation(arg0 = modulus), 2 threshold(arg0 = threshold, arg1 = modulus),
 * 3 default test polynomial (arg0 = plaintext modulus t). */
FHEB_API int fheb_make_test_poly(const fheb_ntt_plan* plan, int kind, uint64_t arg0, uint64_t arg1, uint64_t* out_host);

/* ---- encrypted-ballot tally ---------------------------------------------------------------- */

/* replaces EncryptionEngine::batch_add / batch_add_tree / tally_votes:
 * cpp/src/encryption.cpp:1061-1067,1327-1458.  cts = [count][2][N] (c0 then c1 of each
 * ballot); out = [2][N].  count == 0 fails with the reference's message; count == 1 returns
 * the ballot's words untouched, as the reference does (:1332-1334). */
FHEB_API int fheb_tally(const uint64_t* cts, size_t count, uint32_t degree, uint64_t modulus, uint64_t* out,
                        void* stream);
/* noise_budget metadata of the tally variants (host arithmetic; the words above do not depend on the variant):
 * replaces the budget bookkeeping of EncryptionEngine::batch_add (variant 0: min - log2(count),
 * cpp/src/encryption.cpp:1337-1360), batch_add_tree / tally_votes (variant 1: min(pair) - 1 per level, odd element
 * carried, :1390-1456) and a left fold with add (variant 2: :613). */
FHEB_API int fheb_tally_noise_budget(const double* budgets, size_t count, int variant, double* out);
/* Second step of the sharded tally: folds `parts` partial tallies ([parts][2][N], canonical,
 * e.g. the result of an NCCL all-gather of per-GPU fheb_tally outputs) into out = [2][N]. */
FHEB_API int fheb_tally_combine(const uint64_t* partials, size_t parts, uint32_t degree, uint64_t modulus,
                                uint64_t* out, void* stream);
/* Sharded tally with the cross-GPU step FUSED into the tally kernel (one process per GPU of one box).  Same result
 * words as fheb_tally on every rank + an all-gather + fheb_tally_combine, in ONE launch per rank: the block that
 * finishes a column chunk stores its words into every peer's inbox over NVLink (CUDA IPC peer memory), publishes a
 * flag, waits for the peers' flags and sums the rows.  Set-up: every rank creates its handle (which exports a
 * 64-byte IPC handle), the ranks exchange those bytes by any means (the Python mirror uses one torch.distributed
 * all-gather), then connect.  run() must be called by all ranks the same number of times (each call is one
 * epoch; the epoch advances only when the launch succeeded).  A rank that never arrives makes its peers give up
 * after the timeout (~17 s unless set): the result words of that call become all-ones (never a residue), the
 * host-mapped status word records the failing epoch, and every later run() on the handle fails until reset. */
typedef struct fheb_tally_peers fheb_tally_peers;
#define FHEB_PEER_HANDLE_BYTES 64
FHEB_API int fheb_tally_peers_create(uint32_t degree, uint64_t modulus, uint32_t world, uint32_t rank, fheb_tally_peers** out,
                                     uint8_t* handle_out /* [FHEB_PEER_HANDLE_BYTES] */);
FHEB_API int fheb_tally_peers_connect(fheb_tally_peers* peers, const uint8_t* handles /* [world][FHEB_PEER_HANDLE_BYTES], rank order */);
/* cts = this rank's ballots [count][2][N] (device), out = [2][N] (device): the GLOBAL tally, on every rank */
FHEB_API int fheb_tally_peers_run(fheb_tally_peers* peers, const uint64_t* cts, size_t count, uint64_t* out, void* stream);
/* timed_out = epoch of the first exchange that timed out, 0 if none.  Reads a host-mapped word: no copy and no
 * synchronisation; call it after synchronising the stream to learn about the calls issued so far. */
FHEB_API int fheb_tally_peers_status(const fheb_tally_peers* peers, int* timed_out);
FHEB_API uint32_t fheb_tally_peers_epoch(const fheb_tally_peers* peers); /* calls launched so far */
/* After a failure: all ranks synchronise, agree on an unused epoch (e.g. max over ranks of _epoch() + 2), reset. */
FHEB_API int fheb_tally_peers_reset(fheb_tally_peers* peers, uint32_t epoch);
FHEB_API int fheb_tally_peers_set_timeout(fheb_tally_peers* peers, double seconds);
FHEB_API int fheb_tally_peers_destroy(fheb_tally_peers* peers);

/* The same sharded tally driven from ONE process that owns several GPUs (the reference's addon is a single Node
 * process: src/native/lib.rs:23-133).  The group enables peer access between its devices; fheb_tally_sharded
 * launches the fused kernel on every device from the calling thread and returns when the global tally is in `out`.
 * replaces EncryptionEngine::tally_votes over sharded ballots: cpp/src/encryption.cpp:1061-1067,1327-1458.
 * devices = NULL selects devices 0..ndev-1.  cts[i] = counts[i] ballots [count][2][N] in the memory of device i of
 * the group (counts may be zero for some, not all, devices); out = [2][N], host memory or any device's memory. */
typedef struct fheb_tally_group fheb_tally_group;
FHEB_API int fheb_tally_group_create(uint32_t degree, uint64_t modulus, const int* devices, uint32_t ndev, fheb_tally_group** out);
FHEB_API int fheb_tally_sharded(fheb_tally_group* group, const uint64_t* const* cts, const size_t* counts, uint64_t* out);
FHEB_API uint32_t fheb_tally_group_size(const fheb_tally_group* group);
FHEB_API int fheb_tally_group_destroy(fheb_tally_group* group);

/* Streaming tally: replaces the running accumulator of
 * CiphertextStreamProcessor::stream_add (cpp/src/streaming_processor.cpp:460-526) and the accumulate path
 * of ChunkedCiphertextProcessor: ballots arrive in chunks, the running total stays on the device.
 * Semantics follow stream_add: the first ciphertext becomes the accumulator untouched; every later one is
 * folded with EncryptionEngine::add, so after >= 2 ballots the total is canonical.  add() accepts host or
 * device chunks ([count][2][N]); total() writes the running total ([2][N]) and may be called at any time. */
typedef struct fheb_tally_stream fheb_tally_stream;
FHEB_API int fheb_tally_stream_create(uint32_t degree, uint64_t modulus, fheb_tally_stream** out);
FHEB_API int fheb_tally_stream_add(fheb_tally_stream* ts, const uint64_t* cts, size_t count, void* stream);
FHEB_API int fheb_tally_stream_total(const fheb_tally_stream* ts, uint64_t* out, void* stream);
FHEB_API uint64_t fheb_tally_stream_count(const fheb_tally_stream* ts);
FHEB_API int fheb_tally_stream_destroy(fheb_tally_stream* ts);

/* replaces EncryptionEngine::multiply (tensor product): cpp/src/encryption.cpp:737-798.
 * ct1, ct2 = [batch][2][N]; out = [batch][3][N]. */
FHEB_API int fheb_tensor_multiply_batch(const fheb_ntt_plan* plan, const uint64_t* ct1, const uint64_t* ct2,
                                        uint64_t* out, size_t batch, void* stream);

/* replaces EvaluationKey / KeySwitchKey as produced by KeyManager::generate_eval_key (cpp/include/key_manager.h:
 * 92-111, cpp/src/key_manager.cpp:266-331) for EncryptionEngine::relinearize.  keys = [key_count][2][N]
 * coefficient-form words, (a, b) of every pair in generation order (host or device).  decomp_base_log /
 * decomp_level are the KeySwitchKey fields; zero values fall back as the reference does (base_log 4,
 * ceil(64 / base_log) levels: cpp/src/encryption.cpp:935-939) and min(levels, key_count) levels are applied.
 * Both key polynomials of every level are transformed ONCE here (the reference re-transforms them on every
 * call).  Shift amounts of 64 bits or more are undefined in the reference and rejected here.  The plan must
 * outlive the key. */
FHEB_API int fheb_relin_key_create(const fheb_ntt_plan* plan, const uint64_t* keys, uint32_t key_count,
                                   uint32_t decomp_base_log, uint32_t decomp_level, uint64_t key_id, fheb_relin_key** out);
FHEB_API int fheb_relin_key_destroy(fheb_relin_key* key);
FHEB_API uint32_t fheb_relin_key_levels(const fheb_relin_key* key); /* levels actually applied */
/* replaces EncryptionEngine::relinearize / relinearize_inplace on degree-2 ciphertexts: cpp/src/encryption.cpp:
 * 904-1003.  cts = [batch][3][N] (c0, c1, c2), coefficient form, e.g. the output of fheb_tensor_multiply_batch;
 * out = [batch][2][N].  ct_key_id must equal the key's id (FHEB_ERR_KEY_MISMATCH, the reference's message).
 * With no key pairs c0 and c1 are returned untouched, as the reference does (:980-989).  out must not overlap cts
 * (FHEB_ERR_INVALID_PARAMETERS). */
FHEB_API int fheb_relinearize_batch(const fheb_relin_key* key, const uint64_t* cts, uint64_t ct_key_id, uint64_t* out,
                                    size_t batch, void* stream);

/* ---- wire formats ----------------------------------------------------------------- */

/* SerializationHeader as written by KeySerializer::write_header (cpp/include/key_serializer.h:59-84, cpp/src/
 * key_serializer.cpp:96-108): 49 packed little-endian bytes in front of data_size payload bytes. */
typedef struct fheb_wire_header {
    uint32_t magic;         /* FHES / FHEP / FHEE / FHEB / FHEV: key_serializer.h:33-37 */
    uint32_t version;
    uint32_t key_type;      /* 0 secret, 1 public, 2 eval, 3 bootstrap, 4 ballot */
    uint64_t key_id;        /* ballots: the timestamp */
    uint32_t poly_degree;
    uint64_t modulus;
    uint32_t data_size;
    uint8_t checksum_type;  /* ChecksumType: 0 none, 1 CRC32, 2 "SHA256" (the reference falls back to its CRC32) */
    uint8_t compression;    /* CompressionType; the reference never compresses */
    uint32_t checksum;
} fheb_wire_header;
FHEB_API int fheb_wire_header_read(const void* bytes, size_t len, fheb_wire_header* out); /* KeySerializer::read_header, :623-641 */
/* replaces KeySerializer::compute_crc32 (:34-40) INCLUDING its table as shipped: only the first 60 of the 256
 * IEEE 802.3 entries are initialised (:21-32), so this is not the standard CRC-32.  Host-side. */
FHEB_API uint32_t fheb_wire_crc32(const void* data, size_t len);

/* replaces BallotSerializer::serialize_ballot (:709-774) for one record: choices = [num_choices][2][degree] host
 * words.  *written receives the record size (also when the buffer is too small); fheb_ballot_wire_size is
 * estimate_ballot_size's exact counterpart (:848-854 counts sizeof(SerializationHeader) = 64 instead of the 49
 * bytes actually written). */
FHEB_API size_t fheb_ballot_wire_size(uint32_t num_choices, uint32_t degree);
FHEB_API int fheb_ballot_serialize(const uint64_t* choices, uint32_t num_choices, uint32_t degree, uint64_t modulus,
                                   uint64_t timestamp, void* out, size_t capacity, size_t* written);

/* Per-record result of fheb_ballots_ingest.  1..3 are the three rejections of BallotSerializer::deserialize_ballot
 * (:781-783 "Input too small", :800-802 "Invalid magic bytes for ballot", :808-811 "Checksum verification failed");
 * 4 is this bulk path's own: the record is well formed but not num_choices ciphertexts of the given degree/modulus. */
enum {
    FHEB_WIRE_OK = 0,
    FHEB_WIRE_TOO_SMALL = 1,
    FHEB_WIRE_BAD_MAGIC = 2,
    FHEB_WIRE_BAD_CHECKSUM = 3,
    FHEB_WIRE_SHAPE_MISMATCH = 4
};
/* replaces a loop of BallotSerializer::deserialize_ballot (:776-846) over `count` FHEV records, validated and
 * unpacked ON THE DEVICE straight into the layout fheb_tally reads.  wire = the records' bytes (host, or 8-byte
 * aligned device memory); offsets = count + 1 host entries, record r = bytes [offsets[r], offsets[r+1]) - or
 * NULL (host wire only): records are back to back and their extents are walked from the headers.
 * cts = [count][num_choices][2][degree] (host or device): accepted records' words exactly as stored; REJECTED
 * records become all-zero ciphertexts, the additive identity, so the buffer can be tallied as a whole.
 * status = [count] host bytes (FHEB_WIRE_*); timestamps = [count] host words or NULL; *accepted (optional) =
 * number of FHEB_WIRE_OK records.  Returns after the results are in place. */
FHEB_API int fheb_ballots_ingest(const void* wire, size_t wire_bytes, const uint64_t* offsets, size_t count,
                                 uint32_t num_choices, uint32_t degree, uint64_t modulus, uint64_t* cts, uint8_t* status,
                                 uint64_t* timestamps, size_t* accepted, void* stream);

/* The receive-and-count path in one call: fheb_ballots_ingest's validation followed by EncryptionEngine::tally_votes
 * (cpp/src/encryption.cpp:1061-1067,1327-1364) over every choice of the ACCEPTED records, summed straight out of the
 * wire bytes (the ciphertexts are never materialised: the wire is read once for the checksums and once for the sum).
 * out = [num_choices][2][degree] (host or device), choice c = the tally of ciphertext c of every accepted ballot.
 * wire / offsets / status / accepted as in fheb_ballots_ingest.  No accepted record fails with the reference's
 * "Cannot add empty vector of ciphertexts"; a single accepted record is returned untouched, as batch_add does. */
FHEB_API int fheb_tally_wire(const void* wire, size_t wire_bytes, const uint64_t* offsets, size_t count, uint32_t num_choices,
                             uint32_t degree, uint64_t modulus, uint64_t* out, uint8_t* status, size_t* accepted, void* stream);

/* replace KeySerializer::deserialize_eval_key (:414-466) / deserialize_bootstrap_key (:545-615) followed by the
 * device key constructors above.  Errors carry the reference's messages ("Failed to read header", "Invalid magic
 * bytes", "Checksum verification failed", "Polynomial degree mismatch").  FHEB containers: glwe_dimension 1 only
 * (the container stores polynomial pairs); the KeyManager-style key-switching pairs at its end are skipped -
 * BootstrapEngine::key_switch consumes a different structure (fheb_boot_key_set_ksk). */
FHEB_API int fheb_relin_key_from_wire(const fheb_ntt_plan* plan, const void* bytes, size_t len, fheb_relin_key** out);
FHEB_API int fheb_boot_key_from_wire(const fheb_ntt_plan* plan, const fheb_boot_params* params, const void* bytes,
                                     size_t len, fheb_boot_key** out);

/* Synthetic ballots generated ON DEVICE (bench / scaling runs: 16 GB never crosses PCIe).
 * word(ballot, comp, j) = splitmix64(seed + (ballot*2 + comp)*N + j) % modulus, reproducible on
 * the CPU (tests/ restate it).  Not part of the reference; test/bench support only. */
FHEB_API int fheb_synth_ballots(uint64_t* cts_device, size_t first_ballot, size_t count, uint32_t degree,
                                uint64_t modulus, uint64_t seed, void* stream);

/* Number of kernel launches issued by this library since the last reset (bench bookkeeping). */
FHEB_API uint64_t fheb_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* FHEB200_H */
