/*
 * addon.c -- N-API (plain C, ABI-stable) shim between node-speex-resampler's TypeScript surface
 * and libspeexb200.so. It replaces what src/speex_wasm.js + the WASM-heap staging of
 * src/index.ts:59-115 do in the reference: the five symbols that file binds
 * (src/index.ts:6-16, exported by scripts/build_emscripten.sh:20) are called here natively on
 * the Buffer's own bytes, and the batched entry behind SpeexResampler.processChunks is added.
 *
 * Exports (all synchronous, like the reference's processChunk):
 *   deviceCount(): number
 *   lastError(): string
 *   init(channels, inRate, outRate, quality): External      throws Error(strerror) on failure
 *   process(handle, chunk: Buffer, capFrames): Buffer       one processChunk (src/index.ts:89-115)
 *   destroy(handle): void
 *   batchCreate(nStreams, channels, inRate, outRate, quality, device): External
 *   batchProcess(batch, chunks: Buffer[], capFrames: Uint32Array): Buffer[]
 *   batchAdopt(batch, streamIndex, handle): void            migrate a single stream's state
 *   batchDestroy(batch): void
 *   setKernel(handle, kernel): void / batchSetKernel(batch, kernel): void
 *                                                   0 auto, 1 strict (bit-exact), 2 tiled, 3 tensor
 *
 * Built by binding.gyp (node-gyp) against include/speexb200.h; no Node toolchain exists in the
 * image this repository is developed in, so tests/test_node_binding.py compiles this file
 * against a declaration-only stand-in for <node_api.h> (bindings/node/test/stub) and checks the
 * TypeScript side's rules against the Python mirror.
 */
#include <node_api.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "speexb200.h"

#define SPX_MAX_ARGS 6

typedef struct {
  SpeexResamplerState *st;
  uint32_t channels;
} spx_handle;

typedef struct {
  spxb_batch *b;
  uint32_t n_streams, channels;
  /* grow-only pinned staging (spxb_host_alloc): packed rows in, packed rows out */
  int16_t *h_in, *h_out;
  size_t in_cap, out_cap; /* int16 elements */
  uint32_t *in_frames, *out_frames;
} spx_batch;

static int16_t empty_frame; /* stands in for the data pointer of a zero-length Buffer */

static napi_value throw_text(napi_env env, const char *msg) {
  napi_throw_error(env, NULL, msg);
  return NULL;
}

/* src/index.ts:63-65,104-106: a non-zero code surfaces as Error(strerror(code)) */
static napi_value throw_code(napi_env env, int err) { return throw_text(env, speex_resampler_strerror(err)); }

static int get_args(napi_env env, napi_callback_info info, size_t want, napi_value *argv) {
  size_t argc = SPX_MAX_ARGS;
  if (napi_get_cb_info(env, info, &argc, argv, NULL, NULL) != napi_ok) return 0;
  if (argc < want) {
    throw_text(env, "Invalid argument.");
    return 0;
  }
  return 1;
}

static int get_u32(napi_env env, napi_value v, uint32_t *out) {
  /* the reference passes JS numbers straight into WASM i32 parameters: ToUint32 semantics */
  double d = 0;
  if (napi_get_value_double(env, v, &d) != napi_ok) {
    throw_text(env, "Invalid argument.");
    return 0;
  }
  *out = (uint32_t)(int64_t)d;
  return 1;
}

static void finalize_handle(napi_env env, void *data, void *hint) {
  spx_handle *h = (spx_handle *)data;
  (void)env;
  (void)hint;
  if (h->st) speex_resampler_destroy(h->st); /* the reference never frees; here GC does */
  free(h);
}

static void release_batch(spx_batch *bt) {
  if (bt->b) spxb_batch_destroy(bt->b);
  if (bt->h_in) spxb_host_free(bt->h_in);
  if (bt->h_out) spxb_host_free(bt->h_out);
  free(bt->in_frames);
  free(bt->out_frames);
  bt->b = NULL;
  bt->h_in = bt->h_out = NULL;
  bt->in_frames = bt->out_frames = NULL;
  bt->in_cap = bt->out_cap = 0;
}

static void finalize_batch(napi_env env, void *data, void *hint) {
  (void)env;
  (void)hint;
  release_batch((spx_batch *)data);
  free(data);
}

static napi_value js_device_count(napi_env env, napi_callback_info info) {
  napi_value r;
  (void)info;
  napi_create_int32(env, spxb_device_count(), &r);
  return r;
}

static napi_value js_last_error(napi_env env, napi_callback_info info) {
  napi_value r;
  (void)info;
  napi_create_string_utf8(env, spxb_last_error(), NAPI_AUTO_LENGTH, &r);
  return r;
}

/* init(channels, inRate, outRate, quality) -- src/index.ts:59-68 */
static napi_value js_init(napi_env env, napi_callback_info info) {
  napi_value argv[SPX_MAX_ARGS], r;
  uint32_t ch, in_rate, out_rate, q;
  int err = 0;
  spx_handle *h;
  if (!get_args(env, info, 4, argv)) return NULL;
  if (!get_u32(env, argv[0], &ch) || !get_u32(env, argv[1], &in_rate) || !get_u32(env, argv[2], &out_rate) ||
      !get_u32(env, argv[3], &q))
    return NULL;
  h = (spx_handle *)calloc(1, sizeof(*h));
  if (!h) return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
  h->st = speex_resampler_init(ch, in_rate, out_rate, (int)q, &err);
  h->channels = ch;
  if (!h->st) {
    free(h);
    return throw_code(env, err ? err : RESAMPLER_ERR_ALLOC_FAILED);
  }
  if (napi_create_external(env, h, finalize_handle, NULL, &r) != napi_ok) {
    speex_resampler_destroy(h->st);
    free(h);
    return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
  }
  return r;
}

/* process(handle, chunk, capFrames) -> fresh Buffer with the frames written.
 * The input Buffer is read in place (no copy into a staging heap); the consumed count is
 * dropped exactly as src/index.ts:108 drops it. */
static napi_value js_process(napi_env env, napi_callback_info info) {
  napi_value argv[SPX_MAX_ARGS], r;
  spx_handle *h = NULL;
  void *in = NULL, *out = NULL;
  size_t in_bytes = 0;
  uint32_t cap = 0, in_len, out_len;
  int err;
  if (!get_args(env, info, 3, argv)) return NULL;
  if (napi_get_value_external(env, argv[0], (void **)&h) != napi_ok || !h || !h->st)
    return throw_code(env, RESAMPLER_ERR_BAD_STATE);
  if (napi_get_buffer_info(env, argv[1], &in, &in_bytes) != napi_ok) return throw_code(env, RESAMPLER_ERR_INVALID_ARG);
  if (!get_u32(env, argv[2], &cap)) return NULL;
  in_len = (uint32_t)(in_bytes / 2 / h->channels);
  out_len = cap;
  if (napi_create_buffer(env, (size_t)cap * h->channels * 2, &out, &r) != napi_ok)
    return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
  /* N-API may hand out NULL for the data of a zero-length Buffer; the reference returns an empty
   * Buffer for an empty chunk (its loop does not run), so never pass NULL down */
  if (!in) in = &empty_frame;
  if (!out) out = &empty_frame;
  err = speex_resampler_process_interleaved_int(h->st, (const int16_t *)in, &in_len, (int16_t *)out, &out_len);
  if (err) return throw_code(env, err);
  if (out_len != cap) {
    /* a fresh Buffer of exactly the written bytes (src/index.ts:111-115) */
    napi_value exact;
    void *dst = NULL;
    if (napi_create_buffer_copy(env, (size_t)out_len * h->channels * 2, out, &dst, &exact) != napi_ok)
      return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
    return exact;
  }
  return r;
}

static napi_value js_destroy(napi_env env, napi_callback_info info) {
  napi_value argv[SPX_MAX_ARGS];
  spx_handle *h = NULL;
  if (!get_args(env, info, 1, argv)) return NULL;
  if (napi_get_value_external(env, argv[0], (void **)&h) == napi_ok && h && h->st) {
    speex_resampler_destroy(h->st);
    h->st = NULL;
  }
  return NULL;
}

/* batchCreate(nStreams, channels, inRate, outRate, quality, device) */
static napi_value js_batch_create(napi_env env, napi_callback_info info) {
  napi_value argv[SPX_MAX_ARGS], r;
  uint32_t n, ch, in_rate, out_rate, q, dev;
  int err = 0;
  spx_batch *bt;
  if (!get_args(env, info, 6, argv)) return NULL;
  if (!get_u32(env, argv[0], &n) || !get_u32(env, argv[1], &ch) || !get_u32(env, argv[2], &in_rate) ||
      !get_u32(env, argv[3], &out_rate) || !get_u32(env, argv[4], &q) || !get_u32(env, argv[5], &dev))
    return NULL;
  bt = (spx_batch *)calloc(1, sizeof(*bt));
  if (!bt) return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
  bt->b = spxb_batch_create(n, ch, in_rate, out_rate, (int)q, (int)dev, &err);
  bt->n_streams = n;
  bt->channels = ch;
  bt->in_frames = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  bt->out_frames = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  if (!bt->b || !bt->in_frames || !bt->out_frames) {
    release_batch(bt);
    free(bt);
    return throw_code(env, err ? err : RESAMPLER_ERR_ALLOC_FAILED);
  }
  if (napi_create_external(env, bt, finalize_batch, NULL, &r) != napi_ok) {
    release_batch(bt);
    free(bt);
    return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
  }
  return r;
}

static int grow_pinned(int16_t **p, size_t *cap, size_t want) {
  int16_t *np;
  if (want <= *cap) return 1;
  np = (int16_t *)spxb_host_alloc(want * sizeof(int16_t));
  if (!np) return 0;
  if (*p) spxb_host_free(*p);
  *p = np;
  *cap = want;
  return 1;
}

/* batchProcess(batch, chunks, capFrames) -> Buffer[]: one processChunk per stream in one launch.
 * chunks[s] is stream s's interleaved int16 input (lengths may differ), capFrames[s] its output
 * capacity in frames (the caller applies the grow-only rule of src/index.ts:80-95 per stream). */
static napi_value js_batch_process(napi_env env, napi_callback_info info) {
  napi_value argv[SPX_MAX_ARGS], result, elem;
  spx_batch *bt = NULL;
  uint32_t n = 0, s, max_in = 0, max_cap = 0;
  napi_typedarray_type ty;
  size_t caps_len = 0, in_stride, out_stride;
  void *caps = NULL;
  int err;
  if (!get_args(env, info, 3, argv)) return NULL;
  if (napi_get_value_external(env, argv[0], (void **)&bt) != napi_ok || !bt || !bt->b)
    return throw_code(env, RESAMPLER_ERR_BAD_STATE);
  if (napi_get_array_length(env, argv[1], &n) != napi_ok || n != bt->n_streams)
    return throw_code(env, RESAMPLER_ERR_INVALID_ARG);
  if (napi_get_typedarray_info(env, argv[2], &ty, &caps_len, &caps, NULL, NULL) != napi_ok ||
      ty != napi_uint32_array || caps_len != n)
    return throw_code(env, RESAMPLER_ERR_INVALID_ARG);
  /* pass 1: lengths */
  for (s = 0; s < n; ++s) {
    void *data = NULL;
    size_t bytes = 0;
    if (napi_get_element(env, argv[1], s, &elem) != napi_ok ||
        napi_get_buffer_info(env, elem, &data, &bytes) != napi_ok)
      return throw_code(env, RESAMPLER_ERR_INVALID_ARG);
    if (bytes % ((size_t)bt->channels * 2) != 0)
      return throw_text(env, "Chunk length should be a multiple of channels * 2 bytes");
    bt->in_frames[s] = (uint32_t)(bytes / 2 / bt->channels);
    bt->out_frames[s] = ((const uint32_t *)caps)[s];
    if (bt->in_frames[s] > max_in) max_in = bt->in_frames[s];
    if (bt->out_frames[s] > max_cap) max_cap = bt->out_frames[s];
  }
  in_stride = max_in ? max_in : 1;
  out_stride = max_cap ? max_cap : 1;
  if (!grow_pinned(&bt->h_in, &bt->in_cap, in_stride * bt->channels * n) ||
      !grow_pinned(&bt->h_out, &bt->out_cap, out_stride * bt->channels * n))
    return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
  /* pass 2: pack the rows into pinned memory (one flat DMA on the library side) */
  for (s = 0; s < n; ++s) {
    void *data = NULL;
    size_t bytes = 0;
    napi_get_element(env, argv[1], s, &elem);
    napi_get_buffer_info(env, elem, &data, &bytes);
    if (bytes) memcpy(bt->h_in + (size_t)s * in_stride * bt->channels, data, bytes);
  }
  err = spxb_batch_process(bt->b, bt->h_in, in_stride, bt->in_frames, bt->h_out, out_stride, bt->out_frames);
  if (err) return throw_code(env, err);
  if (napi_create_array_with_length(env, n, &result) != napi_ok) return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
  for (s = 0; s < n; ++s) {
    void *dst = NULL;
    if (napi_create_buffer_copy(env, (size_t)bt->out_frames[s] * bt->channels * 2,
                                bt->h_out + (size_t)s * out_stride * bt->channels, &dst, &elem) != napi_ok)
      return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
    napi_set_element(env, result, s, elem);
  }
  return result;
}

/* batchAdopt(batch, streamIndex, handle): copy a stream that already ran through processChunk
 * (its last_sample / samp_frac_num / history) into slot streamIndex of the batch */
static napi_value js_batch_adopt(napi_env env, napi_callback_info info) {
  napi_value argv[SPX_MAX_ARGS];
  spx_batch *bt = NULL;
  spx_handle *h = NULL;
  uint32_t idx = 0, frac = 0, magic = 0, in_rate = 0, out_rate = 0;
  int32_t last = 0;
  int16_t *hist;
  int err, latency;
  if (!get_args(env, info, 3, argv)) return NULL;
  if (napi_get_value_external(env, argv[0], (void **)&bt) != napi_ok || !bt || !bt->b ||
      napi_get_value_external(env, argv[2], (void **)&h) != napi_ok || !h || !h->st)
    return throw_code(env, RESAMPLER_ERR_BAD_STATE);
  if (!get_u32(env, argv[1], &idx)) return NULL;
  speex_resampler_get_rate(h->st, &in_rate, &out_rate);
  latency = speex_resampler_get_input_latency(h->st); /* filt_len / 2 */
  hist = (int16_t *)calloc((size_t)(2 * latency + 2) * h->channels, sizeof(int16_t));
  if (!hist) return throw_code(env, RESAMPLER_ERR_ALLOC_FAILED);
  err = spxb_batch_get_state(spxb_resampler_batch(h->st), 0, &last, &frac, &magic, hist);
  if (!err) err = spxb_batch_set_state(bt->b, idx, last, frac, hist);
  free(hist);
  if (err) return throw_code(env, err);
  return NULL;
}

static napi_value js_batch_destroy(napi_env env, napi_callback_info info) {
  napi_value argv[SPX_MAX_ARGS];
  spx_batch *bt = NULL;
  if (!get_args(env, info, 1, argv)) return NULL;
  if (napi_get_value_external(env, argv[0], (void **)&bt) == napi_ok && bt) release_batch(bt);
  return NULL;
}

/* setKernel(handle, kernel) / batchSetKernel(batch, kernel): kernel family of include/speexb200.h
 * (SPXB_KERNEL_*). A single stream defaults to the bit-exact kernel, a batch to AUTO. */
static napi_value js_set_kernel(napi_env env, napi_callback_info info) {
  napi_value argv[SPX_MAX_ARGS];
  spx_handle *h = NULL;
  uint32_t k = 0;
  int err;
  if (!get_args(env, info, 2, argv)) return NULL;
  if (napi_get_value_external(env, argv[0], (void **)&h) != napi_ok || !h || !h->st)
    return throw_code(env, RESAMPLER_ERR_BAD_STATE);
  if (!get_u32(env, argv[1], &k)) return NULL;
  err = spxb_batch_set_kernel(spxb_resampler_batch(h->st), (int)k);
  if (err) return throw_code(env, err);
  return NULL;
}

static napi_value js_batch_set_kernel(napi_env env, napi_callback_info info) {
  napi_value argv[SPX_MAX_ARGS];
  spx_batch *bt = NULL;
  uint32_t k = 0;
  int err;
  if (!get_args(env, info, 2, argv)) return NULL;
  if (napi_get_value_external(env, argv[0], (void **)&bt) != napi_ok || !bt || !bt->b)
    return throw_code(env, RESAMPLER_ERR_BAD_STATE);
  if (!get_u32(env, argv[1], &k)) return NULL;
  err = spxb_batch_set_kernel(bt->b, (int)k);
  if (err) return throw_code(env, err);
  return NULL;
}

#define SPX_METHOD(name, fn) \
  { name, NULL, fn, NULL, NULL, NULL, napi_default, NULL }

NAPI_MODULE_INIT() {
  const napi_property_descriptor props[] = {
      SPX_METHOD("deviceCount", js_device_count),   SPX_METHOD("lastError", js_last_error),
      SPX_METHOD("init", js_init),                  SPX_METHOD("process", js_process),
      SPX_METHOD("destroy", js_destroy),            SPX_METHOD("batchCreate", js_batch_create),
      SPX_METHOD("batchProcess", js_batch_process), SPX_METHOD("batchAdopt", js_batch_adopt),
      SPX_METHOD("batchDestroy", js_batch_destroy), SPX_METHOD("setKernel", js_set_kernel),
      SPX_METHOD("batchSetKernel", js_batch_set_kernel),
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}
