/*
 * addon_selftest.c -- executes bindings/node/src/addon.c through the in-process N-API stand-in
 * (fake_napi.c) against the real libspeexb200.so: the calls index.ts makes, in the order it
 * makes them. Prints one line per result ("label frames fnv1a64"); tests/test_parity_gpu.py
 * recomputes every line with the CPU oracle (the single stream runs the bit-exact kernel by default,
 * the batches are switched to it with batchSetKernel) and compares.
 * Needs a B200; test infrastructure only.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fake_napi.h"

#define CH 2
#define IN_RATE 44100
#define OUT_RATE 48000
#define QUALITY 7

/* deterministic PCM shared with the Python side: x[k] of stream s */
static int16_t pcm_sample(uint32_t s, uint32_t k) {
  uint32_t v = (k + 1u) * 2654435761u + s * 40503u;
  v ^= v >> 15;
  return (int16_t)((int32_t)(v & 0x3fffu) - 8192);
}

static napi_value pcm_buffer(uint32_t s, uint32_t first_frame, uint32_t frames) {
  int16_t *x = (int16_t *)malloc((size_t)frames * CH * 2 + 2);
  uint32_t i;
  napi_value b;
  for (i = 0; i < frames * CH; ++i) x[i] = pcm_sample(s, first_frame * CH + i);
  b = fake_buffer(x, (size_t)frames * CH * 2);
  free(x);
  return b;
}

static uint64_t fnv1a64(const uint8_t *p, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  size_t i;
  for (i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
  return h;
}

static void report(const char *label, unsigned idx, napi_value buf) {
  printf("%s%u %zu %016llx\n", label, idx, buf->length / (CH * 2), (unsigned long long)fnv1a64(buf->data, buf->length));
}

static napi_value must(napi_env env, napi_value v, const char *what) {
  if (fake_exception_pending(env)) {
    printf("UNEXPECTED exception in %s: %s\n", what, fake_exception_message(env));
    exit(1);
  }
  return v;
}

/* the capacity rule of src/index.ts:80-95, grow-only per stream */
static uint32_t capacity_frames(double *out_buffer_size, size_t bytes) {
  /* ceil(bytes * out / in), on integers */
  const double target = (double)(((unsigned long long)bytes * OUT_RATE + IN_RATE - 1) / IN_RATE);
  if (*out_buffer_size < target) *out_buffer_size = target;
  return (uint32_t)(*out_buffer_size / CH / 2);
}

int main(void) {
  napi_env env = fake_env_new();
  napi_value exports = fake_object();
  napi_value argv[6], h, out, batch, chunks, caps_v, res;
  double cap_state = -1;
  uint32_t k, s;
  napi_register_module_v1(env, exports);

  out = must(env, fake_call(env, exports, "deviceCount", 0, argv), "deviceCount");
  if (out->number < 1) {
    printf("no CUDA device\n");
    return 3;
  }

  /* error path first: Error(strerror(err)) like src/index.ts:63-65 */
  argv[0] = fake_number(0); argv[1] = fake_number(IN_RATE); argv[2] = fake_number(OUT_RATE); argv[3] = fake_number(QUALITY);
  fake_call(env, exports, "init", 4, argv);
  printf("init_error %d %s\n", fake_exception_pending(env), fake_exception_message(env));
  fake_exception_clear(env);

  /* one stream: init + three 20 ms hops with a ragged one in between */
  argv[0] = fake_number(CH);
  h = must(env, fake_call(env, exports, "init", 4, argv), "init");
  {
    const uint32_t hops[4] = {882, 441, 7, 882};
    uint32_t pos = 0;
    for (k = 0; k < 4; ++k) {
      napi_value chunk = pcm_buffer(0, pos, hops[k]);
      argv[0] = h; argv[1] = chunk; argv[2] = fake_number(capacity_frames(&cap_state, chunk->length));
      out = must(env, fake_call(env, exports, "process", 3, argv), "process");
      report("single", k, out);
      pos += hops[k];
    }
    /* an empty chunk: the reference returns an empty Buffer (N-API gives NULL data for it) */
    {
      napi_value chunk = fake_buffer("", 0);
      argv[0] = h; argv[1] = chunk; argv[2] = fake_number(capacity_frames(&cap_state, 0));
      out = must(env, fake_call(env, exports, "process", 3, argv), "process(empty)");
      printf("empty %zu\n", out->length);
    }
  }

  /* a batch of 8 streams, ragged chunk lengths, two calls */
  {
    const uint32_t S = 8;
    double cap_states[8];
    uint32_t pos[8] = {0};
    uint32_t caps[8];
    argv[0] = fake_number(S); argv[1] = fake_number(CH); argv[2] = fake_number(IN_RATE); argv[3] = fake_number(OUT_RATE);
    argv[4] = fake_number(QUALITY); argv[5] = fake_number(0);
    batch = must(env, fake_call(env, exports, "batchCreate", 6, argv), "batchCreate");
    argv[0] = batch; argv[1] = fake_number(1); /* SPXB_KERNEL_STRICT */
    must(env, fake_call(env, exports, "batchSetKernel", 2, argv), "batchSetKernel");
    for (s = 0; s < S; ++s) cap_states[s] = -1;
    for (k = 0; k < 2; ++k) {
      chunks = fake_array(S);
      for (s = 0; s < S; ++s) {
        const uint32_t frames = k == 0 ? 882 : 300 + 41 * s;
        chunks->items[s] = pcm_buffer(10 + s, pos[s], frames);
        caps[s] = capacity_frames(&cap_states[s], chunks->items[s]->length);
        pos[s] += frames;
      }
      caps_v = fake_uint32_array(caps, S);
      argv[0] = batch; argv[1] = chunks; argv[2] = caps_v;
      res = must(env, fake_call(env, exports, "batchProcess", 3, argv), "batchProcess");
      for (s = 0; s < S; ++s) report(k == 0 ? "batchA" : "batchB", s, res->items[s]);
    }
    argv[0] = batch;
    must(env, fake_call(env, exports, "batchDestroy", 1, argv), "batchDestroy");
  }

  /* migrate the single stream (already 4 hops in) into slot 1 of a 2-stream batch, then continue it */
  {
    uint32_t caps[2];
    double fresh = -1;
    argv[0] = fake_number(2); argv[1] = fake_number(CH); argv[2] = fake_number(IN_RATE); argv[3] = fake_number(OUT_RATE);
    argv[4] = fake_number(QUALITY); argv[5] = fake_number(0);
    batch = must(env, fake_call(env, exports, "batchCreate", 6, argv), "batchCreate");
    argv[0] = batch; argv[1] = fake_number(1); /* SPXB_KERNEL_STRICT */
    must(env, fake_call(env, exports, "batchSetKernel", 2, argv), "batchSetKernel");
    argv[0] = batch; argv[1] = fake_number(1); argv[2] = h;
    must(env, fake_call(env, exports, "batchAdopt", 3, argv), "batchAdopt");
    chunks = fake_array(2);
    chunks->items[0] = pcm_buffer(30, 0, 882);
    chunks->items[1] = pcm_buffer(0, 882 + 441 + 7 + 882, 882);
    caps[0] = capacity_frames(&fresh, chunks->items[0]->length);
    caps[1] = capacity_frames(&cap_state, chunks->items[1]->length);
    caps_v = fake_uint32_array(caps, 2);
    argv[0] = batch; argv[1] = chunks; argv[2] = caps_v;
    res = must(env, fake_call(env, exports, "batchProcess", 3, argv), "batchProcess");
    report("adopt", 0, res->items[0]);
    report("adopt", 1, res->items[1]);
    /* misaligned chunk: the fixed message of src/index.ts:56 */
    chunks->items[0] = fake_buffer("abc", 3);
    fake_call(env, exports, "batchProcess", 3, argv);
    printf("align_error %d %s\n", fake_exception_pending(env), fake_exception_message(env));
    fake_exception_clear(env);
  }
  argv[0] = h;
  must(env, fake_call(env, exports, "destroy", 1, argv), "destroy");
  printf("done\n");
  return 0;
}
