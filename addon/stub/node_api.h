/*
 * SYNTAX-CHECK STAND-IN for Node's <node_api.h> - NOT the real header and never shipped or linked.
 *
 * The build image has no Node toolchain, so addon/fheb_addon.cc cannot be compiled against the real N-API here.
 * This file declares just the handful of N-API (Node-API version 6) types and functions the addon uses, with
 * their public, ABI-stable signatures, so that `g++ -fsyntax-only -Iaddon/stub` type-checks the addon in the CPU
 * test suite.  A real build uses node-gyp / cmake-js, which put Node's own node_api.h first on the include path
 * (see addon/binding.gyp); this directory is not on it.
 */
#ifndef FHEB_STUB_NODE_API_H
#define FHEB_STUB_NODE_API_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_ref__* napi_ref;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected,
               napi_function_expected, napi_number_expected, napi_boolean_expected, napi_array_expected,
               napi_generic_failure, napi_pending_exception } napi_status;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
               napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array,
               napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_default = 0, napi_writable = 1 << 0, napi_enumerable = 1 << 1, napi_configurable = 1 << 2,
               napi_static = 1 << 10 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef struct {
    const char* utf8name;
    napi_value name;
    napi_callback method;
    napi_callback getter;
    napi_callback setter;
    napi_value value;
    napi_property_attributes attributes;
    void* data;
} napi_property_descriptor;
#define NAPI_AUTO_LENGTH SIZE_MAX
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_define_class(napi_env env, const char* utf8name, size_t length, napi_callback constructor, void* data,
                              size_t property_count, const napi_property_descriptor* properties, napi_value* result);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor* properties);
napi_status napi_wrap(napi_env env, napi_value js_object, void* native_object, napi_finalize finalize_cb, void* finalize_hint, napi_ref* result);
napi_status napi_unwrap(napi_env env, napi_value js_object, void** result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_get_boolean(napi_env env, bool value, napi_value* result);
napi_status napi_get_undefined(napi_env env, napi_value* result);
napi_status napi_create_double(napi_env env, double value, napi_value* result);
napi_status napi_create_uint32(napi_env env, uint32_t value, napi_value* result);
napi_status napi_create_bigint_uint64(napi_env env, uint64_t value, napi_value* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_value_bigint_uint64(napi_env env, napi_value value, uint64_t* result, bool* lossless);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value* result);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void** data, size_t* length);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
#ifdef __cplusplus
}
#define NAPI_MODULE_INIT() extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports)
#else
#define NAPI_MODULE_INIT() napi_value napi_register_module_v1(napi_env env, napi_value exports)
#endif
#endif
