/*
 * fake_napi.c -- a minimal in-process stand-in for the part of Node's N-API that
 * bindings/node/src/addon.c uses, so that the addon can be EXECUTED (on the GPU box, against
 * libspeexb200.so) in an image that has no Node. Values are small tagged structs, never freed
 * (test process); errors thrown by the addon are kept as a pending exception like in Node.
 * Test infrastructure only; see addon_selftest.c and tests/test_node_binding.py.
 */
#include <node_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fake_napi.h"

struct napi_env__ {
  int pending;
  char message[256];
};

struct napi_callback_info__ {
  size_t argc;
  napi_value *argv;
};

static napi_value new_value(fake_kind k) {
  napi_value v = (napi_value)calloc(1, sizeof(struct napi_value__));
  v->kind = k;
  return v;
}

napi_env fake_env_new(void) { return (napi_env)calloc(1, sizeof(struct napi_env__)); }
int fake_exception_pending(napi_env env) { return env->pending; }
const char *fake_exception_message(napi_env env) { return env->message; }
void fake_exception_clear(napi_env env) {
  env->pending = 0;
  env->message[0] = 0;
}

napi_value fake_number(double d) {
  napi_value v = new_value(FAKE_NUMBER);
  v->number = d;
  return v;
}

napi_value fake_buffer(const void *data, size_t len) {
  napi_value v = new_value(FAKE_BUFFER);
  /* like N-API, a zero-length Buffer has no storage: its data pointer is NULL */
  v->data = len ? (uint8_t *)malloc(len) : NULL;
  if (len) memcpy(v->data, data, len);
  v->length = len;
  return v;
}

napi_value fake_array(size_t n) {
  napi_value v = new_value(FAKE_ARRAY);
  v->items = (napi_value *)calloc(n ? n : 1, sizeof(napi_value));
  v->length = n;
  return v;
}

napi_value fake_uint32_array(const uint32_t *data, size_t n) {
  napi_value v = new_value(FAKE_UINT32_ARRAY);
  v->data = (uint8_t *)malloc((n ? n : 1) * 4);
  if (n) memcpy(v->data, data, n * 4);
  v->length = n;
  return v;
}

napi_value fake_object(void) { return new_value(FAKE_OBJECT); }

/* exports.<name>(args...) */
napi_value fake_call(napi_env env, napi_value exports, const char *name, size_t argc, napi_value *argv) {
  size_t i;
  for (i = 0; i < exports->n_props; ++i)
    if (strcmp(exports->prop_names[i], name) == 0) {
      struct napi_callback_info__ info;
      info.argc = argc;
      info.argv = argv;
      return exports->prop_methods[i](env, &info);
    }
  fprintf(stderr, "fake_napi: no export %s\n", name);
  exit(2);
}

/* ---- the N-API subset ---- */
napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t *argc, napi_value *argv,
                             napi_value *this_arg, void **data) {
  size_t i, cap = *argc;
  (void)env;
  for (i = 0; i < cap; ++i) argv[i] = i < info->argc ? info->argv[i] : NULL;
  *argc = info->argc;
  if (this_arg) *this_arg = NULL;
  if (data) *data = NULL;
  return napi_ok;
}

napi_status napi_throw_error(napi_env env, const char *code, const char *msg) {
  (void)code;
  env->pending = 1;
  snprintf(env->message, sizeof env->message, "%s", msg ? msg : "");
  return napi_ok;
}

napi_status napi_get_value_double(napi_env env, napi_value value, double *result) {
  (void)env;
  if (!value || value->kind != FAKE_NUMBER) return napi_invalid_arg;
  *result = value->number;
  return napi_ok;
}

napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result) {
  (void)env;
  *result = fake_number(value);
  return napi_ok;
}

napi_status napi_create_string_utf8(napi_env env, const char *str, size_t length, napi_value *result) {
  napi_value v = new_value(FAKE_STRING);
  (void)env;
  if (length == NAPI_AUTO_LENGTH) length = strlen(str);
  v->data = (uint8_t *)calloc(length + 1, 1);
  memcpy(v->data, str, length);
  v->length = length;
  *result = v;
  return napi_ok;
}

napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint,
                                 napi_value *result) {
  napi_value v = new_value(FAKE_EXTERNAL);
  (void)env;
  v->external = data;
  v->finalize = finalize_cb;
  v->finalize_hint = finalize_hint;
  *result = v;
  return napi_ok;
}

napi_status napi_get_value_external(napi_env env, napi_value value, void **result) {
  (void)env;
  if (!value || value->kind != FAKE_EXTERNAL) return napi_invalid_arg;
  *result = value->external;
  return napi_ok;
}

napi_status napi_get_buffer_info(napi_env env, napi_value value, void **data, size_t *length) {
  (void)env;
  if (!value || value->kind != FAKE_BUFFER) return napi_invalid_arg;
  if (data) *data = value->data;
  if (length) *length = value->length;
  return napi_ok;
}

napi_status napi_create_buffer(napi_env env, size_t length, void **data, napi_value *result) {
  napi_value v = new_value(FAKE_BUFFER);
  (void)env;
  v->data = length ? (uint8_t *)calloc(length, 1) : NULL; /* NULL data for an empty Buffer, like N-API */
  v->length = length;
  if (data) *data = v->data;
  *result = v;
  return napi_ok;
}

napi_status napi_create_buffer_copy(napi_env env, size_t length, const void *data, void **result_data,
                                    napi_value *result) {
  (void)env;
  *result = fake_buffer(data, length);
  if (result_data) *result_data = (*result)->data;
  return napi_ok;
}

napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t *result) {
  (void)env;
  if (!value || value->kind != FAKE_ARRAY) return napi_invalid_arg;
  *result = (uint32_t)value->length;
  return napi_ok;
}

napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value *result) {
  (void)env;
  if (!object || object->kind != FAKE_ARRAY || index >= object->length) return napi_invalid_arg;
  *result = object->items[index];
  return napi_ok;
}

napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value) {
  (void)env;
  if (!object || object->kind != FAKE_ARRAY || index >= object->length) return napi_invalid_arg;
  object->items[index] = value;
  return napi_ok;
}

napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value *result) {
  (void)env;
  *result = fake_array(length);
  return napi_ok;
}

napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type,
                                     size_t *length, void **data, napi_value *arraybuffer, size_t *byte_offset) {
  (void)env;
  if (!typedarray || typedarray->kind != FAKE_UINT32_ARRAY) return napi_invalid_arg;
  if (type) *type = napi_uint32_array;
  if (length) *length = typedarray->length;
  if (data) *data = typedarray->data;
  if (arraybuffer) *arraybuffer = NULL;
  if (byte_offset) *byte_offset = 0;
  return napi_ok;
}

napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count,
                                   const napi_property_descriptor *properties) {
  size_t i;
  (void)env;
  object->prop_names = (const char **)calloc(property_count, sizeof(char *));
  object->prop_methods = (napi_callback *)calloc(property_count, sizeof(napi_callback));
  for (i = 0; i < property_count; ++i) {
    object->prop_names[i] = properties[i].utf8name;
    object->prop_methods[i] = properties[i].method;
  }
  object->n_props = property_count;
  return napi_ok;
}
