/*
 * node_api.h — STUB for syntax-checking ts/addon/rt_napi.c in an image without node.
 *
 * Only the types and the prototypes rt_napi.c uses, with the signatures of Node-API version 8 (node >= 12.22) as
 * published in the Node.js documentation ("Node-API" chapter).  It is NOT the real header: nothing here is linked or
 * executed; `tests/test_host_and_abi.py::test_napi_shim_compiles_against_the_stub_header` runs
 * `gcc -fsyntax-only -Wall -Wextra -Werror` over the shim with this directory on the include path, which catches
 * typos, wrong argument counts and type mismatches against the C ABI (include/rt_b200.h).  A maintainer builds the
 * addon with node-gyp against the real header (ts/addon/binding.gyp).
 */
#ifndef RT_STUB_NODE_API_H
#define RT_STUB_NODE_API_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef struct napi_deferred__* napi_deferred;
typedef struct napi_async_work__* napi_async_work;
typedef struct napi_ref__* napi_ref;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_generic_failure = 9, napi_pending_exception = 10, napi_cancelled = 11 } napi_status;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
               napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_default = 0, napi_writable = 1 << 0, napi_enumerable = 1 << 1, napi_configurable = 1 << 2, napi_static = 1 << 10 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);
typedef struct {
  const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter; napi_value value;
  napi_property_attributes attributes; void* data;
} napi_property_descriptor;
typedef napi_value (*napi_addon_register_func)(napi_env env, napi_value exports);
typedef struct napi_module {
  int nm_version; unsigned int nm_flags; const char* nm_filename; napi_addon_register_func nm_register_func; const char* nm_modname;
  void* nm_priv; void* reserved[4];
} napi_module;
void napi_module_register(napi_module* mod);
#define NAPI_AUTO_LENGTH SIZE_MAX
#define NODE_GYP_MODULE_NAME rt_b200
#define NAPI_MODULE(modname, regfunc) \
  static napi_module _module = {1, 0, __FILE__, regfunc, #modname, NULL, {0}}; \
  static void _register_##modname(void) __attribute__((constructor)); \
  static void _register_##modname(void) { napi_module_register(&_module); }

napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_named_property(napi_env env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_has_named_property(napi_env env, napi_value object, const char* utf8name, bool* result);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_create_double(napi_env env, double value, napi_value* result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_get_undefined(napi_env env, napi_value* result);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor* properties);
napi_status napi_adjust_external_memory(napi_env env, int64_t change_in_bytes, int64_t* adjusted_value);
napi_status napi_create_promise(napi_env env, napi_deferred* deferred, napi_value* promise);
napi_status napi_resolve_deferred(napi_env env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_async_work(napi_env env, napi_value async_resource, napi_value async_resource_name, napi_async_execute_callback execute,
                                   napi_async_complete_callback complete, void* data, napi_async_work* result);
napi_status napi_queue_async_work(napi_env env, napi_async_work work);
napi_status napi_delete_async_work(napi_env env, napi_async_work work);
napi_status napi_create_reference(napi_env env, napi_value value, uint32_t initial_refcount, napi_ref* result);
napi_status napi_delete_reference(napi_env env, napi_ref ref);
#endif
