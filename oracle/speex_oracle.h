/*
 * speex_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar, strict IEEE-754, no FMA contraction) of the
 * arithmetic the reference runs behind SpeexResampler.processChunk:
 *   /root/reference/deps/speex/resample.c   (filter bank, four inner kernels,
 *                                            streaming driver)
 *   /root/reference/deps/speex/arch.h:208   (WORD2INT)
 *   /root/reference/src/index.ts:50-116     (processChunk capacity rule)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this. The shipped library
 * (node_speex_resampler_b200/csrc) never includes, links or calls it.
 *
 * Parity pinning: the reference has NO golden vectors (src/test.ts:40,74 only
 * assert duration). This restatement is pinned bit-for-bit against a native
 * gcc build of the reference's own resample.c (oracle/_ref, see Makefile) on the
 * reference's three resources/ .pcm fixtures and on seeded synthetic inputs;
 * the resulting hashes are frozen in tests/golden/. It is also pinned against the
 * reference's SHIPPED WebAssembly module, executed through oracle/wasm2c_lite.py
 * (oracle/_ref/libspeex_wasm.so): tables and PCM bit-identical on the whole parity
 * matrix and on the fixtures (tests/test_oracle.py).
 */
#ifndef SPEEX_ORACLE_H
#define SPEEX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_resampler orc_resampler;

/* mirrors the error enum at speex_resampler.h:104-113 */
enum {
  ORC_OK = 0,
  ORC_ERR_ALLOC = 1,
  ORC_ERR_BAD_STATE = 2,
  ORC_ERR_INVALID_ARG = 3,
  ORC_ERR_PTR_OVERLAP = 4,
  ORC_ERR_OVERFLOW = 5
};

typedef struct {
  uint32_t num, den;          /* gcd-reduced in/out ratio (resample.c:1125-1128) */
  uint32_t filt_len;          /* N */
  uint32_t oversample;
  int32_t int_advance, frac_advance;
  float cutoff;
  int32_t use_direct;         /* 1: per-phase table den*N ; 0: oversampled table os*N+8 */
  int32_t use_double;         /* 1: quality > 8 -> f64 accumulators */
  uint32_t table_len;
  uint32_t channels;
  int32_t quality;
} orc_params;

orc_resampler *orc_create(uint32_t channels, uint32_t in_rate, uint32_t out_rate,
                          int quality, int *err);
void orc_destroy(orc_resampler *r);

/* same contract as speex_resampler_process_interleaved_int (resample.c:1061):
 * *in_frames / *out_frames are per-channel frame counts, in: available/capacity,
 * out: consumed/written (values of the last channel). */
int orc_process_interleaved_int16(orc_resampler *r, const int16_t *in,
                                  uint32_t *in_frames, int16_t *out,
                                  uint32_t *out_frames);

/* same contract as speex_resampler_process_interleaved_float of the float build
 * (resample.c:1038-1059 over :927-963): float samples in, the kernels' f32 results out
 * unrounded; int16 and float calls may be mixed on one state (the history is float). */
int orc_process_interleaved_float(orc_resampler *r, const float *in, uint32_t *in_frames,
                                  float *out, uint32_t *out_frames);

void orc_get_params(const orc_resampler *r, orc_params *p);
const float *orc_table(const orc_resampler *r);
/* per-channel streaming state: last_sample, samp_frac_num and the N-1 history
 * floats at the head of that channel's work buffer */
void orc_get_state(const orc_resampler *r, uint32_t channel, int32_t *last_sample,
                   uint32_t *samp_frac_num, const float **history);

/* SpeexResampler.processChunk (src/index.ts:50-116) over the restatement:
 * grow-only output capacity = running max of ceil(bytes*out/in), frames =
 * trunc(cap/channels/2), consumed-input count ignored. `cap_bytes_state` is the
 * instance's _outBufferSize (start it at -1). Returns bytes written to `out`
 * (caller provides at least ceil(bytes*out/in) bytes) or -(err) on error;
 * -100 if nbytes is not a multiple of channels*2. */
long orc_process_chunk(orc_resampler *r, double *cap_bytes_state, uint32_t in_rate,
                       uint32_t out_rate, const uint8_t *chunk, size_t nbytes,
                       uint8_t *out);

uint64_t orc_fnv1a64(const uint8_t *p, size_t n);

#ifdef __cplusplus
}
#endif
#endif
