/*
 * Declaration-only stand-in for Node's <node_api.h>, written from the documented N-API C
 * interface (https://nodejs.org/api/n-api.html) so that tests/test_node_binding.py can type-check
 * bindings/node/src/addon.c in an image without Node. It declares exactly the subset addon.c
 * uses and is never linked or shipped; the real header comes with node-gyp.
 */
#ifndef SPXB_NODE_API_STUB_H
#define SPXB_NODE_API_STUB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;

typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_generic_failure = 9 } napi_status;

typedef enum {
  napi_default = 0,
  napi_writable = 1 << 0,
  napi_enumerable = 1 << 1,
  napi_configurable = 1 << 2
} napi_property_attributes;

typedef enum {
  napi_int8_array,
  napi_uint8_array,
  napi_uint8_clamped_array,
  napi_int16_array,
  napi_uint16_array,
  napi_int32_array,
  napi_uint32_array,
  napi_float32_array,
  napi_float64_array
} napi_typedarray_type;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);

typedef struct {
  const char *utf8name;
  napi_value name;
  napi_callback method;
  napi_callback getter;
  napi_callback setter;
  napi_value value;
  napi_property_attributes attributes;
  void *data;
} napi_property_descriptor;

#define NAPI_AUTO_LENGTH SIZE_MAX

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv,
                             napi_value *this_arg, void **data);
napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_get_value_double(napi_env env, napi_value value, double *result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result);
napi_status napi_create_string_utf8(napi_env env, const char *str, size_t length, napi_value *result);
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint,
                                 napi_value *result);
napi_status napi_get_value_external(napi_env env, napi_value value, void **result);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void **data, size_t *length);
napi_status napi_create_buffer(napi_env env, size_t length, void **data, napi_value *result);
napi_status napi_create_buffer_copy(napi_env env, size_t length, const void *data, void **result_data,
                                    napi_value *result);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t *result);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value *result);
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value);
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value *result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type,
                                     size_t *length, void **data, napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count,
                                   const napi_property_descriptor *properties);

/* module registration, reduced to the init function's signature */
#define NAPI_MODULE_INIT() napi_value napi_register_module_v1(napi_env env, napi_value exports)

#ifdef __cplusplus
}
#endif
#endif
