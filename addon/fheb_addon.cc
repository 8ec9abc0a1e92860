// Node-API addon over libfheb200.so: the surface of the reference's napi-rs addon
// (index.d.ts:14-44, src/native/lib.rs:23-133) - initialize, detectHardware, version, class ModularArithmetic -
// plus the bulk entry points src/api/fhe-engine.ts needs to stop returning `{handle: 0n}` placeholders
// (fhe-engine.ts:209-321): NttProcessor, BootstrapEngine, RelinearizationKey, tallyVotes, ingestBallots.
//
// Bulk data cross as BigUint64Array / Buffer (no copy on the JS side; the C ABI stages host buffers through HBM);
// 64-bit scalars that do not fit a JS number (moduli, key ids) cross as BigInt.  The existing scalar class keeps
// its `number` signatures exactly as index.d.ts declares them, precision loss above 2^53 included.
//
// Plain Node-API C calls (ABI-stable, Node-API version 6): builds with node-gyp or cmake-js against any Node >= 12.17
// (addon/binding.gyp).  The build image has no Node toolchain; the CPU test suite type-checks this file against
// addon/stub/node_api.h (tests/test_cabi_cpu.py) and INTEGRATION.md shows the napi-rs route as well.
#include <node_api.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "fheb200.h"

namespace {

#define NAPI_OK(call)                                                       \
    do {                                                                    \
        if ((call) != napi_ok) {                                            \
            napi_throw_error(env, "NATIVE_ERROR", "N-API call failed: " #call); \
            return nullptr;                                                 \
        }                                                                   \
    } while (0)

// FHEErrorCode names of src/api/types.ts:140-151 for the status codes of fheb200.h
const char* code_name(int rc) {
    switch (rc) {
        case FHEB_ERR_INVALID_PARAMETERS: return "INVALID_PARAMETERS";
        case FHEB_ERR_KEY_MISMATCH: return "KEY_MISMATCH";
        case FHEB_ERR_HARDWARE_UNAVAILABLE: return "HARDWARE_UNAVAILABLE";
        case FHEB_ERR_OUT_OF_MEMORY: return "OUT_OF_MEMORY";
        default: return "NATIVE_ERROR";
    }
}

// throws FHEError-shaped exceptions: `code` = the FHEErrorCode name, message = the library's (the reference's) text
#define FHEB_OK_OR_THROW(call)                                   \
    do {                                                         \
        const int rc_ = (call);                                  \
        if (rc_ != FHEB_OK) {                                    \
            napi_throw_error(env, code_name(rc_), fheb_last_error()); \
            return nullptr;                                      \
        }                                                        \
    } while (0)

struct Args {
    size_t argc = 8;
    napi_value argv[8];
    napi_value self = nullptr;
};

bool get_args(napi_env env, napi_callback_info info, Args* a, size_t need) {
    if (napi_get_cb_info(env, info, &a->argc, a->argv, &a->self, nullptr) != napi_ok || a->argc < need) {
        napi_throw_error(env, "INVALID_PARAMETERS", "wrong number of arguments");
        return false;
    }
    return true;
}

bool get_u64(napi_env env, napi_value v, uint64_t* out) {  // BigInt, or a non-negative integral number
    bool lossless = false;
    if (napi_get_value_bigint_uint64(env, v, out, &lossless) == napi_ok) return true;
    double d = 0;
    if (napi_get_value_double(env, v, &d) == napi_ok && d >= 0 && d <= 9007199254740991.0) {
        *out = (uint64_t)d;
        return true;
    }
    napi_throw_error(env, "INVALID_PARAMETERS", "expected a BigInt or a non-negative integer");
    return false;
}

bool get_u32(napi_env env, napi_value v, uint32_t* out) {
    if (napi_get_value_uint32(env, v, out) == napi_ok) return true;
    napi_throw_error(env, "INVALID_PARAMETERS", "expected an unsigned 32-bit integer");
    return false;
}

// BigUint64Array -> (words, count)
bool get_words(napi_env env, napi_value v, uint64_t** data, size_t* count) {
    napi_typedarray_type type;
    void* p = nullptr;
    if (napi_get_typedarray_info(env, v, &type, count, &p, nullptr, nullptr) != napi_ok || type != napi_biguint64_array) {
        napi_throw_error(env, "INVALID_PARAMETERS", "expected a BigUint64Array");
        return false;
    }
    *data = static_cast<uint64_t*>(p);
    return true;
}

napi_value new_words(napi_env env, size_t count, uint64_t** data) {
    void* p = nullptr;
    napi_value buf, arr;
    NAPI_OK(napi_create_arraybuffer(env, count * 8, &p, &buf));
    NAPI_OK(napi_create_typedarray(env, napi_biguint64_array, count, buf, 0, &arr));
    *data = static_cast<uint64_t*>(p);
    return arr;
}

napi_value undefined(napi_env env) {
    napi_value u;
    napi_get_undefined(env, &u);
    return u;
}

// ---- initialize / detectHardware / version (index.d.ts:14-27) -------------------------------------------------
napi_value Initialize(napi_env env, napi_callback_info) {
    FHEB_OK_OR_THROW(fheb_init(-1));
    return undefined(env);
}

napi_value DetectHardware(napi_env env, napi_callback_info) {
    fheb_device_info i;
    FHEB_OK_OR_THROW(fheb_device_info_get(&i));
    napi_value o, f, cores, mem;
    NAPI_OK(napi_create_object(env, &o));
    NAPI_OK(napi_get_boolean(env, false, &f));  // the Apple units are truthfully absent
    NAPI_OK(napi_set_named_property(env, o, "hasSme", f));
    NAPI_OK(napi_set_named_property(env, o, "hasMetal", f));
    NAPI_OK(napi_set_named_property(env, o, "hasNeon", f));
    NAPI_OK(napi_set_named_property(env, o, "hasAmx", f));
    NAPI_OK(napi_create_uint32(env, (uint32_t)i.sm_count, &cores));  // the "GPU cores" analogue: 148 SMs
    NAPI_OK(napi_set_named_property(env, o, "metalGpuCores", cores));
    NAPI_OK(napi_create_double(env, (double)i.device_memory_bytes, &mem));
    NAPI_OK(napi_set_named_property(env, o, "unifiedMemorySize", mem));
    return o;
}

napi_value Version(napi_env env, napi_callback_info) {
    napi_value s;
    NAPI_OK(napi_create_string_utf8(env, fheb_version(), NAPI_AUTO_LENGTH, &s));
    return s;
}

// ---- class ModularArithmetic (index.d.ts:29-44, lib.rs:44-120): numbers in, numbers out ----------------------------
void ModArithFinalize(napi_env, void* data, void*) { fheb_modarith_destroy(static_cast<fheb_modarith*>(data)); }

napi_value ModArithNew(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 1)) return nullptr;
    double m = 0;
    if (napi_get_value_double(env, a.argv[0], &m) != napi_ok || m <= 0) {  // lib.rs:53-55
        napi_throw_error(env, "INVALID_PARAMETERS", "Modulus must be positive");
        return nullptr;
    }
    fheb_modarith* h = nullptr;
    FHEB_OK_OR_THROW(fheb_modarith_create((uint64_t)m, &h));
    NAPI_OK(napi_wrap(env, a.self, h, ModArithFinalize, nullptr, nullptr));
    return a.self;
}

template <uint64_t (*OP)(const fheb_modarith*, uint64_t, uint64_t)>
napi_value ModArithBinary(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 2)) return nullptr;
    void* h = nullptr;
    NAPI_OK(napi_unwrap(env, a.self, &h));
    double x = 0, y = 0;
    if (napi_get_value_double(env, a.argv[0], &x) != napi_ok || napi_get_value_double(env, a.argv[1], &y) != napi_ok || x < 0 || y < 0) {
        napi_throw_error(env, "INVALID_PARAMETERS", "Inputs must be non-negative");  // lib.rs:64-66
        return nullptr;
    }
    napi_value r;
    NAPI_OK(napi_create_double(env, (double)OP(static_cast<fheb_modarith*>(h), (uint64_t)x, (uint64_t)y), &r));
    return r;
}

template <uint64_t (*OP)(const fheb_modarith*, uint64_t)>
napi_value ModArithUnary(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 1)) return nullptr;
    void* h = nullptr;
    NAPI_OK(napi_unwrap(env, a.self, &h));
    double x = 0;
    if (napi_get_value_double(env, a.argv[0], &x) != napi_ok || x < 0) {
        napi_throw_error(env, "INVALID_PARAMETERS", "Input must be non-negative");
        return nullptr;
    }
    napi_value r;
    NAPI_OK(napi_create_double(env, (double)OP(static_cast<fheb_modarith*>(h), (uint64_t)x), &r));
    return r;
}

napi_value ModArithGetModulus(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 0)) return nullptr;
    void* h = nullptr;
    NAPI_OK(napi_unwrap(env, a.self, &h));
    napi_value r;
    NAPI_OK(napi_create_double(env, (double)fheb_modarith_get_modulus(static_cast<fheb_modarith*>(h)), &r));
    return r;
}

// ---- class NttProcessor(degree, modulus: bigint): forwardBatch / inverseBatch / polymulBatch -----------------------
struct Ntt {
    fheb_ntt_plan* plan = nullptr;
    uint32_t degree = 0;
};
void NttFinalize(napi_env, void* data, void*) {
    Ntt* n = static_cast<Ntt*>(data);
    fheb_ntt_plan_destroy(n->plan);
    delete n;
}

napi_value NttNew(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 2)) return nullptr;
    uint32_t degree = 0;
    uint64_t modulus = 0;
    if (!get_u32(env, a.argv[0], &degree) || !get_u64(env, a.argv[1], &modulus)) return nullptr;
    Ntt* n = new Ntt();
    n->degree = degree;
    const int rc = fheb_ntt_plan_create(degree, modulus, &n->plan);  // NTTProcessor's constructor messages
    if (rc != FHEB_OK) {
        delete n;
        napi_throw_error(env, code_name(rc), fheb_last_error());
        return nullptr;
    }
    NAPI_OK(napi_wrap(env, a.self, n, NttFinalize, nullptr, nullptr));
    return a.self;
}

// coeffs: BigUint64Array of batch * degree words, transformed in place
template <int (*FN)(const fheb_ntt_plan*, const uint64_t*, uint64_t*, size_t, void*)>
napi_value NttInPlace(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 1)) return nullptr;
    void* h = nullptr;
    NAPI_OK(napi_unwrap(env, a.self, &h));
    Ntt* n = static_cast<Ntt*>(h);
    uint64_t* data = nullptr;
    size_t count = 0;
    if (!get_words(env, a.argv[0], &data, &count)) return nullptr;
    if (count % n->degree != 0) {
        napi_throw_error(env, "INVALID_PARAMETERS", "Size must match polynomial degree");  // ntt_processor.cpp:263-265
        return nullptr;
    }
    FHEB_OK_OR_THROW(FN(n->plan, data, data, count / n->degree, nullptr));
    return a.argv[0];
}

napi_value NttPolymul(napi_env env, napi_callback_info info) {  // (a, b) -> new BigUint64Array, PolynomialRing::multiply
    Args a;
    if (!get_args(env, info, &a, 2)) return nullptr;
    void* h = nullptr;
    NAPI_OK(napi_unwrap(env, a.self, &h));
    Ntt* n = static_cast<Ntt*>(h);
    uint64_t *x = nullptr, *y = nullptr, *z = nullptr;
    size_t cx = 0, cy = 0;
    if (!get_words(env, a.argv[0], &x, &cx) || !get_words(env, a.argv[1], &y, &cy)) return nullptr;
    if (cx != cy || cx % n->degree != 0) {
        napi_throw_error(env, "INVALID_PARAMETERS", "Polynomial degree mismatch");
        return nullptr;
    }
    napi_value out = new_words(env, cx, &z);
    if (!out) return nullptr;
    FHEB_OK_OR_THROW(fheb_polymul_batch(n->plan, x, y, z, cx / n->degree, nullptr));
    return out;
}

// ---- tallyVotes(ballots: BigUint64Array [count][2][N], degree, modulus: bigint) -> BigUint64Array [2][N] ------------
napi_value TallyVotes(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 3)) return nullptr;
    uint64_t* cts = nullptr;
    size_t words = 0;
    uint32_t degree = 0;
    uint64_t modulus = 0;
    if (!get_words(env, a.argv[0], &cts, &words) || !get_u32(env, a.argv[1], &degree) || !get_u64(env, a.argv[2], &modulus)) return nullptr;
    if (degree == 0 || words % (2 * (size_t)degree) != 0) {
        napi_throw_error(env, "INVALID_PARAMETERS", "ballots must hold count * 2 * degree words");
        return nullptr;
    }
    uint64_t* out = nullptr;
    napi_value res = new_words(env, 2 * (size_t)degree, &out);
    if (!res) return nullptr;
    FHEB_OK_OR_THROW(fheb_tally(cts, words / (2 * (size_t)degree), degree, modulus, out, nullptr));
    return res;
}

// ---- ingestBallots(wire: Buffer, count, numChoices, degree, modulus) -> {ciphertexts, status, accepted} ---------------
napi_value IngestBallots(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 5)) return nullptr;
    void* wire = nullptr;
    size_t wire_bytes = 0;
    uint32_t count = 0, choices = 0, degree = 0;
    uint64_t modulus = 0;
    if (napi_get_buffer_info(env, a.argv[0], &wire, &wire_bytes) != napi_ok) {
        napi_throw_error(env, "INVALID_PARAMETERS", "wire must be a Buffer");
        return nullptr;
    }
    if (!get_u32(env, a.argv[1], &count) || !get_u32(env, a.argv[2], &choices) || !get_u32(env, a.argv[3], &degree) ||
        !get_u64(env, a.argv[4], &modulus))
        return nullptr;
    uint64_t* cts = nullptr;
    napi_value cts_arr = new_words(env, (size_t)count * choices * 2 * degree, &cts);
    if (!cts_arr) return nullptr;
    void* status = nullptr;
    napi_value status_buf, status_arr, accepted_v, o;
    NAPI_OK(napi_create_arraybuffer(env, count, &status, &status_buf));
    NAPI_OK(napi_create_typedarray(env, napi_uint8_array, count, status_buf, 0, &status_arr));
    size_t accepted = 0;
    FHEB_OK_OR_THROW(fheb_ballots_ingest(wire, wire_bytes, nullptr, count, choices, degree, modulus, cts,
                                         static_cast<uint8_t*>(status), nullptr, &accepted, nullptr));
    NAPI_OK(napi_create_object(env, &o));
    NAPI_OK(napi_create_uint32(env, (uint32_t)accepted, &accepted_v));
    NAPI_OK(napi_set_named_property(env, o, "ciphertexts", cts_arr));
    NAPI_OK(napi_set_named_property(env, o, "status", status_arr));  // FHEB_WIRE_* per record
    NAPI_OK(napi_set_named_property(env, o, "accepted", accepted_v));
    return o;
}

// ---- class BootstrapEngine(degree, modulus, lweDimension, glweDimension, baseLog, level, bsk) --------------------------
struct Boot {
    fheb_ntt_plan* plan = nullptr;
    fheb_boot_key* key = nullptr;
    uint32_t n = 0, k = 0, degree = 0;
};
void BootFinalize(napi_env, void* data, void*) {
    Boot* b = static_cast<Boot*>(data);
    fheb_boot_key_destroy(b->key);
    fheb_ntt_plan_destroy(b->plan);
    delete b;
}

napi_value BootNew(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 7)) return nullptr;
    uint32_t degree = 0, n = 0, k = 0, base_log = 0, level = 0;
    uint64_t modulus = 0;
    uint64_t* bsk = nullptr;
    size_t words = 0;
    if (!get_u32(env, a.argv[0], &degree) || !get_u64(env, a.argv[1], &modulus) || !get_u32(env, a.argv[2], &n) ||
        !get_u32(env, a.argv[3], &k) || !get_u32(env, a.argv[4], &base_log) || !get_u32(env, a.argv[5], &level) ||
        !get_words(env, a.argv[6], &bsk, &words))
        return nullptr;
    if (words != (size_t)n * (k + 1) * level * (k + 1) * degree) {
        napi_throw_error(env, "INVALID_PARAMETERS", "bootstrap key has the wrong number of words");
        return nullptr;
    }
    Boot* b = new Boot();
    b->n = n;
    b->k = k;
    b->degree = degree;
    const fheb_boot_params params = {n, k, base_log, level};
    int rc = fheb_ntt_plan_create(degree, modulus, &b->plan);
    if (rc == FHEB_OK) rc = fheb_boot_key_create(b->plan, &params, bsk, &b->key);
    if (rc != FHEB_OK) {
        BootFinalize(env, b, nullptr);
        napi_throw_error(env, code_name(rc), fheb_last_error());
        return nullptr;
    }
    NAPI_OK(napi_wrap(env, a.self, b, BootFinalize, nullptr, nullptr));
    return a.self;
}

// bootstrapBatch(lwe: BigUint64Array [batch][n+1], testPoly: BigUint64Array [N]) -> BigUint64Array [batch][k*N+1]
napi_value BootBootstrap(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 2)) return nullptr;
    void* h = nullptr;
    NAPI_OK(napi_unwrap(env, a.self, &h));
    Boot* b = static_cast<Boot*>(h);
    uint64_t *lwe = nullptr, *tp = nullptr, *out = nullptr;
    size_t lw = 0, tw = 0;
    if (!get_words(env, a.argv[0], &lwe, &lw) || !get_words(env, a.argv[1], &tp, &tw)) return nullptr;
    if (lw % (b->n + 1) != 0 || tw != b->degree) {
        napi_throw_error(env, "INVALID_PARAMETERS", "lwe must hold batch * (n + 1) words and testPoly N words");
        return nullptr;
    }
    const size_t batch = lw / (b->n + 1), ow = (size_t)b->k * b->degree + 1;
    napi_value res = new_words(env, batch * ow, &out);
    if (!res) return nullptr;
    FHEB_OK_OR_THROW(fheb_bootstrap_batch(b->key, lwe, tp, out, batch, nullptr));
    return res;
}

// ---- multiplyRelinearize(degree, modulus, evalKeyWire: Buffer, ct1, ct2) -> BigUint64Array [batch][2][N] --------------
// EncryptionEngine::multiply followed by relinearize with an FHEE container as KeySerializer writes it
napi_value MultiplyRelinearize(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 5)) return nullptr;
    uint32_t degree = 0;
    uint64_t modulus = 0;
    void* wire = nullptr;
    size_t wire_bytes = 0;
    uint64_t *c1 = nullptr, *c2 = nullptr, *out = nullptr;
    size_t w1 = 0, w2 = 0;
    if (!get_u32(env, a.argv[0], &degree) || !get_u64(env, a.argv[1], &modulus)) return nullptr;
    if (napi_get_buffer_info(env, a.argv[2], &wire, &wire_bytes) != napi_ok) {
        napi_throw_error(env, "INVALID_PARAMETERS", "evalKeyWire must be a Buffer");
        return nullptr;
    }
    if (!get_words(env, a.argv[3], &c1, &w1) || !get_words(env, a.argv[4], &c2, &w2)) return nullptr;
    if (degree == 0 || w1 != w2 || w1 % (2 * (size_t)degree) != 0) {
        napi_throw_error(env, "INVALID_PARAMETERS", "ciphertexts must hold batch * 2 * degree words each");
        return nullptr;
    }
    const size_t batch = w1 / (2 * (size_t)degree);
    fheb_ntt_plan* plan = nullptr;
    fheb_relin_key* key = nullptr;
    void* ct3 = nullptr;
    fheb_wire_header hdr;
    napi_value res = new_words(env, batch * 2 * degree, &out);
    if (!res) return nullptr;
    int rc = fheb_ntt_plan_create(degree, modulus, &plan);
    if (rc == FHEB_OK) rc = fheb_wire_header_read(wire, wire_bytes, &hdr);
    if (rc == FHEB_OK) rc = fheb_relin_key_from_wire(plan, wire, wire_bytes, &key);
    if (rc == FHEB_OK) rc = fheb_device_alloc(&ct3, batch * 3 * degree * 8);  // the degree-2 ciphertexts never leave HBM
    if (rc == FHEB_OK) rc = fheb_tensor_multiply_batch(plan, c1, c2, static_cast<uint64_t*>(ct3), batch, nullptr);
    if (rc == FHEB_OK) rc = fheb_relinearize_batch(key, static_cast<uint64_t*>(ct3), hdr.key_id, out, batch, nullptr);
    std::string msg = rc != FHEB_OK ? fheb_last_error() : "";
    if (ct3) fheb_device_free(ct3);
    fheb_relin_key_destroy(key);
    fheb_ntt_plan_destroy(plan);
    if (rc != FHEB_OK) {
        napi_throw_error(env, code_name(rc), msg.c_str());
        return nullptr;
    }
    return res;
}

// ---- modAddBatch / modSubBatch(a, b, modulus: bigint) -> new BigUint64Array: PolynomialRing::add / subtract over a batch -------
template <int (*FN)(const uint64_t*, const uint64_t*, uint64_t*, size_t, uint64_t, void*)>
napi_value ElementwiseBinary(napi_env env, napi_callback_info info) {
    Args a;
    if (!get_args(env, info, &a, 3)) return nullptr;
    uint64_t *x = nullptr, *y = nullptr, *z = nullptr;
    size_t cx = 0, cy = 0;
    uint64_t modulus = 0;
    if (!get_words(env, a.argv[0], &x, &cx) || !get_words(env, a.argv[1], &y, &cy) || !get_u64(env, a.argv[2], &modulus)) return nullptr;
    if (cx != cy) {
        napi_throw_error(env, "INVALID_PARAMETERS", "Polynomial degree mismatch");  // polynomial_ring.cpp:241-257
        return nullptr;
    }
    napi_value out = new_words(env, cx, &z);
    if (!out) return nullptr;
    FHEB_OK_OR_THROW(FN(x, y, z, cx, modulus, nullptr));
    return out;
}

// ---- setDevices(): one Node process, every visible GPU (fheb_set_devices); returns the number of devices in use ----------
napi_value SetDevices(napi_env env, napi_callback_info) {
    FHEB_OK_OR_THROW(fheb_set_devices(nullptr, -1));
    napi_value v;
    NAPI_OK(napi_create_uint32(env, (uint32_t)fheb_get_devices(nullptr, 0), &v));
    return v;
}

#define METHOD(name, fn) {name, nullptr, fn, nullptr, nullptr, nullptr, napi_default, nullptr}

}  // namespace

NAPI_MODULE_INIT() {
    const napi_property_descriptor functions[] = {
        METHOD("initialize", Initialize),
        METHOD("detectHardware", DetectHardware),
        METHOD("version", Version),
        METHOD("tallyVotes", TallyVotes),
        METHOD("ingestBallots", IngestBallots),
        METHOD("multiplyRelinearize", MultiplyRelinearize),
        METHOD("modAddBatch", ElementwiseBinary<fheb_modadd_batch>),
        METHOD("modSubBatch", ElementwiseBinary<fheb_modsub_batch>),
        METHOD("setDevices", SetDevices),
    };
    NAPI_OK(napi_define_properties(env, exports, sizeof(functions) / sizeof(functions[0]), functions));

    const napi_property_descriptor modarith[] = {
        METHOD("montgomeryMul", ModArithBinary<fheb_modarith_montgomery_mul>),
        METHOD("modAdd", ModArithBinary<fheb_modarith_mod_add>),
        METHOD("modSub", ModArithBinary<fheb_modarith_mod_sub>),
        METHOD("toMontgomery", ModArithUnary<fheb_modarith_to_montgomery>),
        METHOD("fromMontgomery", ModArithUnary<fheb_modarith_from_montgomery>),
        METHOD("getModulus", ModArithGetModulus),
    };
    napi_value cls;
    NAPI_OK(napi_define_class(env, "ModularArithmetic", NAPI_AUTO_LENGTH, ModArithNew, nullptr,
                              sizeof(modarith) / sizeof(modarith[0]), modarith, &cls));
    NAPI_OK(napi_set_named_property(env, exports, "ModularArithmetic", cls));

    const napi_property_descriptor ntt[] = {
        METHOD("forwardBatch", NttInPlace<fheb_ntt_forward_batch>),
        METHOD("inverseBatch", NttInPlace<fheb_ntt_inverse_batch>),
        METHOD("polymulBatch", NttPolymul),
    };
    NAPI_OK(napi_define_class(env, "NttProcessor", NAPI_AUTO_LENGTH, NttNew, nullptr, sizeof(ntt) / sizeof(ntt[0]), ntt, &cls));
    NAPI_OK(napi_set_named_property(env, exports, "NttProcessor", cls));

    const napi_property_descriptor boot[] = {METHOD("bootstrapBatch", BootBootstrap)};
    NAPI_OK(napi_define_class(env, "BootstrapEngine", NAPI_AUTO_LENGTH, BootNew, nullptr, sizeof(boot) / sizeof(boot[0]), boot, &cls));
    NAPI_OK(napi_set_named_property(env, exports, "BootstrapEngine", cls));
    return exports;
}
