/* fake_napi.h -- value model of the in-process N-API stand-in (test infrastructure only) */
#ifndef SPXB_FAKE_NAPI_H
#define SPXB_FAKE_NAPI_H
#include <node_api.h>

typedef enum {
  FAKE_UNDEFINED = 0, FAKE_NUMBER, FAKE_STRING, FAKE_BUFFER, FAKE_ARRAY, FAKE_EXTERNAL,
  FAKE_UINT32_ARRAY, FAKE_OBJECT
} fake_kind;

struct napi_value__ {
  fake_kind kind;
  double number;
  uint8_t *data;   /* Buffer / string / Uint32Array bytes */
  size_t length;   /* bytes (Buffer, string) or elements (Array, Uint32Array) */
  napi_value *items;
  void *external;
  napi_finalize finalize;
  void *finalize_hint;
  const char **prop_names;
  napi_callback *prop_methods;
  size_t n_props;
};

napi_env fake_env_new(void);
int fake_exception_pending(napi_env env);
const char *fake_exception_message(napi_env env);
void fake_exception_clear(napi_env env);
napi_value fake_number(double d);
napi_value fake_buffer(const void *data, size_t len);
napi_value fake_array(size_t n);
napi_value fake_uint32_array(const uint32_t *data, size_t n);
napi_value fake_object(void);
napi_value fake_call(napi_env env, napi_value exports, const char *name, size_t argc, napi_value *argv);
napi_value napi_register_module_v1(napi_env env, napi_value exports);
#endif
