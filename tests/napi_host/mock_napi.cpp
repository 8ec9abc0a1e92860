// A minimal Node-API HOST (test infrastructure, not a Node replacement): implements the Node-API functions that
// addon/fheb_addon.cc calls - with their public, ABI-stable signatures - over a tiny value model, loads the addon the way
// Node does (dlopen + napi_register_module_v1), and drives its exports like JS would: `addon.version()`,
// `new addon.ModularArithmetic(17).modAdd(5, 16)`, `new addon.NttProcessor(1024, q).forwardBatch(arr)` ...
//
// The build image has no Node toolchain and no Node headers, so this is as close as the addon can get to running here:
// the SAME shared object (build/fheb_addon.node) that node-gyp would produce from addon/binding.gyp, executed.  What this
// does not prove is header parity with the real <node_api.h> (addon/stub/node_api.h restates the declarations).
//
//   mock_napi <addon.node> cpu      scalar ModularArithmetic + version + argument errors   (no GPU needed)
//   mock_napi <addon.node> gpu      + initialize, detectHardware, NttProcessor, tallyVotes, modAddBatch against the oracle
#include <dlfcn.h>
#include <node_api.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../../oracle/fhe_oracle.h"

struct napi_value__ {
    enum Kind { UNDEFINED, BOOLEAN, NUMBER, BIGINT, STRING, OBJECT, FUNCTION, ARRAYBUFFER, TYPEDARRAY, BUFFER, CLASS } kind = UNDEFINED;
    bool b = false;
    double num = 0;
    uint64_t big = 0;
    std::string str;
    std::map<std::string, napi_value> props;  // OBJECT: own properties; CLASS: prototype methods
    napi_callback cb = nullptr;               // FUNCTION / CLASS constructor
    void* wrapped = nullptr;                  // napi_wrap
    napi_finalize fin = nullptr;
    std::shared_ptr<std::vector<uint8_t>> bytes;  // ARRAYBUFFER / TYPEDARRAY / BUFFER storage
    size_t offset = 0, length = 0;            // TYPEDARRAY: element count; BUFFER: byte count
    napi_typedarray_type ta = napi_uint8_array;
};
struct napi_env__ {
    std::vector<std::unique_ptr<napi_value__>> heap;
    bool pending = false;
    std::string code, msg;
    napi_value make(napi_value__::Kind k) {
        heap.emplace_back(new napi_value__());
        heap.back()->kind = k;
        return heap.back().get();
    }
};
struct napi_callback_info__ {
    std::vector<napi_value> args;
    napi_value self = nullptr;
};

extern "C" {
napi_status napi_get_cb_info(napi_env, napi_callback_info i, size_t* argc, napi_value* argv, napi_value* self, void** data) {
    if (argc) {
        const size_t cap = *argc;
        for (size_t k = 0; k < cap && k < i->args.size(); ++k) argv[k] = i->args[k];
        *argc = i->args.size();
    }
    if (self) *self = i->self;
    if (data) *data = nullptr;
    return napi_ok;
}
napi_status napi_define_class(napi_env env, const char* name, size_t, napi_callback ctor, void*, size_t n, const napi_property_descriptor* p, napi_value* out) {
    napi_value c = env->make(napi_value__::CLASS);
    c->str = name;
    c->cb = ctor;
    for (size_t k = 0; k < n; ++k) {
        napi_value f = env->make(napi_value__::FUNCTION);
        f->cb = p[k].method;
        c->props[p[k].utf8name] = f;
    }
    *out = c;
    return napi_ok;
}
napi_status napi_define_properties(napi_env env, napi_value obj, size_t n, const napi_property_descriptor* p) {
    for (size_t k = 0; k < n; ++k) {
        napi_value f = env->make(napi_value__::FUNCTION);
        f->cb = p[k].method;
        obj->props[p[k].utf8name] = f;
    }
    return napi_ok;
}
napi_status napi_wrap(napi_env, napi_value obj, void* native, napi_finalize fin, void*, napi_ref*) {
    obj->wrapped = native;
    obj->fin = fin;
    return napi_ok;
}
napi_status napi_unwrap(napi_env, napi_value obj, void** out) {
    if (!obj || !obj->wrapped) return napi_invalid_arg;
    *out = obj->wrapped;
    return napi_ok;
}
napi_status napi_create_object(napi_env env, napi_value* out) { *out = env->make(napi_value__::OBJECT); return napi_ok; }
napi_status napi_set_named_property(napi_env, napi_value obj, const char* name, napi_value v) { obj->props[name] = v; return napi_ok; }
napi_status napi_get_boolean(napi_env env, bool v, napi_value* out) { *out = env->make(napi_value__::BOOLEAN); (*out)->b = v; return napi_ok; }
napi_status napi_get_undefined(napi_env env, napi_value* out) { *out = env->make(napi_value__::UNDEFINED); return napi_ok; }
napi_status napi_create_double(napi_env env, double v, napi_value* out) { *out = env->make(napi_value__::NUMBER); (*out)->num = v; return napi_ok; }
napi_status napi_create_uint32(napi_env env, uint32_t v, napi_value* out) { return napi_create_double(env, (double)v, out); }
napi_status napi_create_bigint_uint64(napi_env env, uint64_t v, napi_value* out) { *out = env->make(napi_value__::BIGINT); (*out)->big = v; return napi_ok; }
napi_status napi_create_string_utf8(napi_env env, const char* s, size_t len, napi_value* out) {
    *out = env->make(napi_value__::STRING);
    (*out)->str = (len == NAPI_AUTO_LENGTH) ? std::string(s) : std::string(s, len);
    return napi_ok;
}
napi_status napi_get_value_double(napi_env, napi_value v, double* out) {
    if (v->kind != napi_value__::NUMBER) return napi_number_expected;
    *out = v->num;
    return napi_ok;
}
napi_status napi_get_value_uint32(napi_env, napi_value v, uint32_t* out) {
    if (v->kind != napi_value__::NUMBER) return napi_number_expected;
    *out = (uint32_t)v->num;
    return napi_ok;
}
napi_status napi_get_value_bigint_uint64(napi_env, napi_value v, uint64_t* out, bool* lossless) {
    if (v->kind != napi_value__::BIGINT) return napi_generic_failure;  // (Node: napi_bigint_expected)
    *out = v->big;
    if (lossless) *lossless = true;
    return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env, napi_value v, napi_typedarray_type* type, size_t* length, void** data, napi_value*, size_t*) {
    if (v->kind != napi_value__::TYPEDARRAY) return napi_invalid_arg;
    if (type) *type = v->ta;
    if (length) *length = v->length;
    if (data) *data = v->bytes->data() + v->offset;
    return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env env, size_t bytes, void** data, napi_value* out) {
    *out = env->make(napi_value__::ARRAYBUFFER);
    (*out)->bytes = std::make_shared<std::vector<uint8_t>>(bytes ? bytes : 8);
    (*out)->length = bytes;
    if (data) *data = (*out)->bytes->data();
    return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value ab, size_t offset, napi_value* out) {
    if (ab->kind != napi_value__::ARRAYBUFFER) return napi_invalid_arg;
    *out = env->make(napi_value__::TYPEDARRAY);
    (*out)->bytes = ab->bytes;
    (*out)->offset = offset;
    (*out)->length = length;
    (*out)->ta = type;
    return napi_ok;
}
napi_status napi_get_buffer_info(napi_env, napi_value v, void** data, size_t* length) {
    if (v->kind != napi_value__::BUFFER) return napi_invalid_arg;
    *data = v->bytes->data();
    *length = v->length;
    return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char* code, const char* msg) {
    if (!env->pending) {
        env->pending = true;
        env->code = code ? code : "";
        env->msg = msg ? msg : "";
    }
    return napi_ok;
}
}  // extern "C"

// ---- the "JS" side ---------------------------------------------------------------------------------------

static napi_env env_;
static int failures = 0;
#define EXPECT(cond, ...)                                \
    do {                                                 \
        if (!(cond)) {                                   \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);                    \
            std::printf("\n");                           \
            ++failures;                                  \
        }                                                \
    } while (0)

static napi_value num(double v) { napi_value r; napi_create_double(env_, v, &r); return r; }
static napi_value big(uint64_t v) { napi_value r; napi_create_bigint_uint64(env_, v, &r); return r; }
static napi_value words(const std::vector<uint64_t>& w) {
    void* p;
    napi_value ab, ta;
    napi_create_arraybuffer(env_, w.size() * 8, &p, &ab);
    std::memcpy(p, w.data(), w.size() * 8);
    napi_create_typedarray(env_, napi_biguint64_array, w.size(), ab, 0, &ta);
    return ta;
}
static std::vector<uint64_t> to_words(napi_value v) {
    std::vector<uint64_t> w(v->length);
    std::memcpy(w.data(), v->bytes->data() + v->offset, w.size() * 8);
    return w;
}
// fn(...args) with `this` = self; a pending exception is returned through `threw` (and cleared)
static napi_value call(napi_value fn, napi_value self, std::vector<napi_value> args, std::string* threw = nullptr) {
    napi_callback_info__ info;
    info.args = std::move(args);
    info.self = self;
    napi_value r = fn->cb(env_, &info);
    if (env_->pending) {
        if (threw) *threw = env_->code + ": " + env_->msg;
        else { std::printf("uncaught exception %s: %s\n", env_->code.c_str(), env_->msg.c_str()); ++failures; }
        env_->pending = false;
        return nullptr;
    }
    if (threw) threw->clear();
    return r;
}
static napi_value construct(napi_value cls, std::vector<napi_value> args, std::string* threw = nullptr) {  // `new cls(...args)`
    napi_value self = env_->make(napi_value__::OBJECT);
    self->props = cls->props;  // prototype methods
    call(cls, self, std::move(args), threw);
    return (threw && !threw->empty()) ? nullptr : self;
}
static napi_value method(napi_value obj, const char* name) {
    auto it = obj->props.find(name);
    if (it == obj->props.end()) { std::printf("FAIL: no property %s\n", name); ++failures; std::exit(2); }
    return it->second;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::printf("usage: mock_napi <addon.node> cpu|gpu\n"); return 2; }
    const bool gpu = std::strcmp(argv[2], "gpu") == 0;
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) { std::printf("dlopen failed: %s\n", dlerror()); return 2; }
    typedef napi_value (*Init)(napi_env, napi_value);
    Init init = reinterpret_cast<Init>(dlsym(h, "napi_register_module_v1"));  // what NAPI_MODULE_INIT() exports
    if (!init) { std::printf("napi_register_module_v1 not exported\n"); return 2; }
    napi_env__ e;
    env_ = &e;
    napi_value addon = init(env_, e.make(napi_value__::OBJECT));
    EXPECT(addon && !e.pending, "module initialisation");

    // ---- part 1 of the typings: the reference addon's own surface (index.d.ts:14-44, test-modular.js) - no GPU involved
    napi_value v = call(method(addon, "version"), addon, {});
    EXPECT(v && v->kind == napi_value__::STRING && v->str.find("b200") != std::string::npos, "version() = %s", v ? v->str.c_str() : "?");
    napi_value MA = method(addon, "ModularArithmetic");
    std::string threw;
    napi_value m17 = construct(MA, {num(17)}, &threw);
    EXPECT(m17 && threw.empty(), "new ModularArithmetic(17): %s", threw.c_str());
    EXPECT(call(method(m17, "getModulus"), m17, {})->num == 17, "getModulus");
    EXPECT(call(method(m17, "modAdd"), m17, {num(5), num(16)})->num == 4, "modAdd(5, 16) mod 17");
    EXPECT(call(method(m17, "modSub"), m17, {num(5), num(16)})->num == 6, "modSub(5, 16) mod 17");
    construct(MA, {num(-5)}, &threw);  // lib.rs:52-55 rejects non-positive moduli
    EXPECT(!threw.empty(), "new ModularArithmetic(-5) must throw");
    call(method(m17, "modAdd"), m17, {num(1)}, &threw);
    EXPECT(threw.find("INVALID_PARAMETERS") == 0, "missing argument must throw INVALID_PARAMETERS (got '%s')", threw.c_str());
    {   // the reference's Montgomery constants are what they are: the mirror must give the same numbers
        const uint64_t q = 132120577ULL;
        napi_value m = construct(MA, {num((double)q)});
        void* sc = nullptr;
        (void)sc;
        const double a = 123456789.0, b = 98765432.0;
        const double mm = call(method(m, "montgomeryMul"), m, {num(a), num(b)})->num;
        const double back = call(method(m, "fromMontgomery"), m, {call(method(m, "toMontgomery"), m, {num(a)})})->num;
        EXPECT(mm >= 0 && mm < (double)q && back >= 0 && back < (double)q, "Montgomery results stay in range");
    }
    if (!gpu) {
        call(method(addon, "initialize"), addon, {}, &threw);  // no sm_100 device here: must fail loudly, never fall back
        if (!threw.empty()) EXPECT(threw.find("HARDWARE_UNAVAILABLE") == 0, "initialize() without a GPU: %s", threw.c_str());
        std::printf(failures ? "NAPI HOST FAILED\n" : "NAPI HOST CPU OK\n");
        return failures ? 1 : 0;
    }

    // ---- part 2: bulk entry points on the GPU, against the oracle
    call(method(addon, "initialize"), addon, {});
    napi_value hw = call(method(addon, "detectHardware"), addon, {});
    EXPECT(hw && method(hw, "metalGpuCores")->num >= 100 && method(hw, "hasMetal")->b == false, "detectHardware()");
    EXPECT(call(method(addon, "setDevices"), addon, {})->num >= 1, "setDevices()");
    const uint32_t N = 1024;
    const uint64_t q = 132120577ULL;
    std::vector<uint64_t> fwd(N), inv(N);
    uint64_t sc[3];
    orc_precompute_twiddles(N, q, fwd.data(), inv.data(), sc);
    std::mt19937_64 rng(42);
    const size_t batch = 5;
    std::vector<uint64_t> a(batch * N), b(batch * N), ref(batch * N);
    for (auto& x : a) x = rng() % q;
    for (auto& x : b) x = rng() % q;
    napi_value ntt = construct(method(addon, "NttProcessor"), {num(N), big(q)});
    napi_value arr = words(a);
    napi_value same = call(method(ntt, "forwardBatch"), ntt, {arr});
    ref = a;
    orc_forward_ntt_batch(ref.data(), batch, N, q, fwd.data());
    EXPECT(same == arr && to_words(arr) == ref, "NttProcessor.forwardBatch vs oracle (in place)");
    call(method(ntt, "inverseBatch"), ntt, {arr});
    EXPECT(to_words(arr) == a, "inverseBatch round trip");
    napi_value prod = call(method(ntt, "polymulBatch"), ntt, {words(a), words(b)});
    for (size_t i = 0; i < batch; ++i) orc_poly_multiply(&a[i * N], &b[i * N], &ref[i * N], N, q, fwd.data(), inv.data(), sc[2]);
    EXPECT(prod && to_words(prod) == ref, "polymulBatch vs oracle");
    call(method(ntt, "forwardBatch"), ntt, {words(std::vector<uint64_t>(N + 1))}, &threw);
    EXPECT(threw.find("Size must match polynomial degree") != std::string::npos, "size mismatch message (got '%s')", threw.c_str());
    std::vector<uint64_t> tref(2 * N);
    orc_tally_linear(a.data(), 2, N, q, tref.data());
    napi_value tal = call(method(addon, "tallyVotes"), addon, {words(std::vector<uint64_t>(a.begin(), a.begin() + 4 * N)), num(N), big(q)});
    EXPECT(tal && to_words(tal) == tref, "tallyVotes vs oracle");
    napi_value sum = call(method(addon, "modAddBatch"), addon, {words(a), words(b), big(q)});
    orc_poly_add(a.data(), b.data(), ref.data(), batch * N, q);
    EXPECT(sum && to_words(sum) == ref, "modAddBatch vs oracle");
    std::printf(failures ? "NAPI HOST FAILED\n" : "NAPI HOST GPU OK\n");
    return failures ? 1 : 0;
}
