/*
 * speexb200.h -- C ABI of libspeexb200.so, the B200-native (sm_100a) Speex-compatible
 * resampler. Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Part 1 is the drop-in surface: the exact symbols the reference's Emscripten build
 * exports (scripts/build_emscripten.sh:20) and src/index.ts:6-16 binds, with the
 * signatures, length conventions and error codes of deps/speex/speex_resampler.h.
 * Part 2 is the batched entry the north star adds (`processChunks`): thousands of
 * independent streams of one (channels, in_rate, out_rate, quality) per launch, state
 * (last_sample, samp_frac_num, magic_samples, history) resident in HBM.
 * Part 3 is host-only introspection (filter bank, call planning) used by the CPU tests.
 *
 * All paths are relative to /root/reference unless they start with csrc/.
 */
#ifndef SPEEXB200_H
#define SPEEXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SPXB_API
#else
#define SPXB_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------ */
/* Part 1: drop-in Speex symbols (replaces deps/speex/resample.c)      */
/* ------------------------------------------------------------------ */

/* error codes: deps/speex/speex_resampler.h:104-113 */
enum {
  RESAMPLER_ERR_SUCCESS = 0,
  RESAMPLER_ERR_ALLOC_FAILED = 1,
  RESAMPLER_ERR_BAD_STATE = 2,
  RESAMPLER_ERR_INVALID_ARG = 3,
  RESAMPLER_ERR_PTR_OVERLAP = 4,
  RESAMPLER_ERR_OVERFLOW = 5,
  RESAMPLER_ERR_MAX_ERROR
};

struct SpeexResamplerState_;
typedef struct SpeexResamplerState_ SpeexResamplerState;

/* replaces resample.c:794 (speex_resampler.h:127-131). One stream, state on the GPU.
 * Returns NULL and *err = RESAMPLER_ERR_INVALID_ARG for channels/rates == 0 or quality
 * outside 0..10 (resample.c:804-809); *err = RESAMPLER_ERR_ALLOC_FAILED when no CUDA
 * device / memory is available (there is NO CPU fallback). */
SPXB_API SpeexResamplerState *speex_resampler_init(uint32_t nb_channels, uint32_t in_rate,
                                                   uint32_t out_rate, int quality, int *err);

/* replaces resample.c:868 (speex_resampler.h:157) */
SPXB_API void speex_resampler_destroy(SpeexResamplerState *st);

/* replaces resample.c:1089 (speex_resampler.h:237-239) */
SPXB_API void speex_resampler_get_rate(SpeexResamplerState *st, uint32_t *in_rate,
                                       uint32_t *out_rate);

/* replaces resample.c:1061 (speex_resampler.h:217-221). `in`/`out` are HOST pointers to
 * interleaved int16 [frame][channel]; *in_len / *out_len are frames per channel: on
 * entry available / capacity, on return consumed / written. Synchronous. */
SPXB_API int speex_resampler_process_interleaved_int(SpeexResamplerState *st,
                                                     const int16_t *in, uint32_t *in_len,
                                                     int16_t *out, uint32_t *out_len);

/* replaces resample.c:1222 (speex_resampler.h:338); same five strings + default */
SPXB_API const char *speex_resampler_strerror(int err);

/* rest of the Speex C API that does not change the filter mid-stream
 * (speex_resampler.h:250-334) */
SPXB_API void speex_resampler_get_ratio(SpeexResamplerState *st, uint32_t *ratio_num,
                                        uint32_t *ratio_den);           /* resample.c:1147 */
SPXB_API void speex_resampler_get_quality(SpeexResamplerState *st, int *quality); /* :1165 */
SPXB_API int speex_resampler_get_input_latency(SpeexResamplerState *st);  /* resample.c:1190 */
SPXB_API int speex_resampler_get_output_latency(SpeexResamplerState *st); /* resample.c:1195 */
/* replaces resample.c:1038 of the float build (speex_resampler.h:189-207): float samples in,
 * the kernels' f32 results out -- no scaling, rounding or saturation -- bit-identical to the
 * reference's float entry. The first float call moves the state's history to f32 (exactly);
 * int16 and float calls may then be mixed freely, as with the reference, but the state runs
 * on the bit-exact strict kernel from then on. */
SPXB_API int speex_resampler_process_interleaved_float(SpeexResamplerState *st, const float *in,
                                                       uint32_t *in_len, float *out,
                                                       uint32_t *out_len);
/* replaces resample.c:799 / :1107 / :1084 / :1153 (speex_resampler.h:140-147, 226-235, 262-276).
 * The ratio may be given apart from the nominal rates; the filter depends on the reduced ratio.
 * set_rate / set_rate_frac / set_quality follow the reference call for call: before the first
 * resampled sample the filter is rebuilt and its memory zeroed; mid-stream a change that keeps
 * the filter length (any ratio change while up-sampling, e.g. clock-drift correction) rescales
 * samp_frac_num, and one that lengthens it re-anchors the history behind zeros and advances
 * last_sample by half the growth (resample.c:727-758, :1131-1140); one that SHORTENS it leaves
 * the surplus history behind as "magic samples" that the next calls resample before their own
 * input, on either entry, with the reference's lengths even when the capacity binds (resample.c:
 * 759-776, :904-922, :940-941, :993-1016; its memory never shrinks, so its input block grows).
 * Only a second change of the filter LENGTH while magic samples are still pending is refused
 * (RESAMPLER_ERR_BAD_STATE, state untouched). */
SPXB_API SpeexResamplerState *speex_resampler_init_frac(uint32_t nb_channels, uint32_t ratio_num,
                                                        uint32_t ratio_den, uint32_t in_rate,
                                                        uint32_t out_rate, int quality, int *err);
SPXB_API int speex_resampler_set_rate(SpeexResamplerState *st, uint32_t in_rate, uint32_t out_rate);
SPXB_API int speex_resampler_set_rate_frac(SpeexResamplerState *st, uint32_t ratio_num,
                                           uint32_t ratio_den, uint32_t in_rate, uint32_t out_rate);
SPXB_API int speex_resampler_set_quality(SpeexResamplerState *st, int quality);
/* Per-channel entries and strides (speex_resampler.h:169-193, :285-309; resample.c:925-1036,
 * :1170-1188): channel `channel_index` of the state alone, reading in[i * in_stride] and writing
 * out[m * out_stride]. The first such call turns the state into a planar one (one mono device
 * stream per channel, each with its own position), which also serves the interleaved entries
 * from then on -- bit-exactly, on the strict kernel. */
SPXB_API int speex_resampler_process_int(SpeexResamplerState *st, uint32_t channel_index,
                                         const int16_t *in, uint32_t *in_len, int16_t *out,
                                         uint32_t *out_len);
SPXB_API int speex_resampler_process_float(SpeexResamplerState *st, uint32_t channel_index,
                                           const float *in, uint32_t *in_len, float *out,
                                           uint32_t *out_len);
SPXB_API void speex_resampler_set_input_stride(SpeexResamplerState *st, uint32_t stride);   /* :1170 */
SPXB_API void speex_resampler_get_input_stride(SpeexResamplerState *st, uint32_t *stride);  /* :1175 */
SPXB_API void speex_resampler_set_output_stride(SpeexResamplerState *st, uint32_t stride);  /* :1180 */
SPXB_API void speex_resampler_get_output_stride(SpeexResamplerState *st, uint32_t *stride); /* :1185 */
SPXB_API int speex_resampler_skip_zeros(SpeexResamplerState *st);         /* resample.c:1200 */
SPXB_API int speex_resampler_reset_mem(SpeexResamplerState *st);          /* resample.c:1208 */

/* ------------------------------------------------------------------ */
/* Part 2: batched streams (backs SpeexResampler.processChunks)        */
/* ------------------------------------------------------------------ */

typedef struct spxb_batch spxb_batch;

/* which device kernel family runs the FIR */
enum {
  SPXB_KERNEL_AUTO = 0,   /* tensor when the call qualifies, else tiled, else strict */
  SPXB_KERNEL_STRICT = 1, /* one thread per output, the reference's own operation order:
                             bit-exact against the scalar reference */
  SPXB_KERNEL_TILED = 2,  /* register-tiled per-phase FIR (precomputed per-phase taps,
                             fp32 accumulate): <= 1 LSB from the reference */
  SPXB_KERNEL_TENSOR = 3  /* tcgen05 int8 tensor-core FIR: exact integer dot product with
                             24-bit fixed-point per-phase taps, one rounding at the end:
                             <= 1 LSB from the reference */
};

SPXB_API int spxb_device_count(void);
/* human-readable text of the last CUDA / argument failure on this thread ("" if none) */
SPXB_API const char *spxb_last_error(void);

/* n_streams independent streams sharing (channels, in_rate, out_rate, quality) on CUDA
 * device `device`. Fresh state: history = 0, last_sample = samp_frac_num = 0
 * (resample.c:721-725, :838-843). */
SPXB_API spxb_batch *spxb_batch_create(uint32_t n_streams, uint32_t channels,
                                       uint32_t in_rate, uint32_t out_rate, int quality,
                                       int device, int *err);
/* Same, with the history kept as f32 (what the reference's `mem` is): serves float in/out
 * (spxb_batch_process_f32) and int16 in/out on one state, bit-exactly, on the strict kernel. */
SPXB_API spxb_batch *spxb_batch_create_f32(uint32_t n_streams, uint32_t channels,
                                           uint32_t in_rate, uint32_t out_rate, int quality,
                                           int device, int *err);
SPXB_API int spxb_batch_is_f32(const spxb_batch *b);
SPXB_API void spxb_batch_destroy(spxb_batch *b);
SPXB_API int spxb_batch_set_kernel(spxb_batch *b, int kernel); /* SPXB_KERNEL_* */
SPXB_API int spxb_batch_get_kernel(const spxb_batch *b);       /* family used by last call */
/* launch geometry of the last tensor-kernel call: {outputs per tile, 32-frame MMA steps per
 * tile, tiles per series group, series groups (128 series each), shared-memory stages,
 * dynamic shared memory bytes}; RESAMPLER_ERR_BAD_STATE when no tensor call was planned yet */
SPXB_API int spxb_batch_tensor_geometry(const spxb_batch *b, uint32_t *geom6);
/* debug: with SPXB_UMMA_TRACE=1 in the environment the tensor kernel records a clock64
 * timeline (32 words per CTA) of its last launch; copies it out, returns CTAs copied */
SPXB_API long spxb_batch_tensor_trace(spxb_batch *b, uint64_t *dst, size_t cap_words);

/* One processChunk-equivalent for every stream, HOST buffers (pageable or pinned).
 * Stream s reads in + s*in_stride_frames*channels (in_frames[s] frames) and writes
 * out + s*out_stride_frames*channels (capacity out_frames[s] frames). in_frames /
 * out_frames are [n_streams] in-out arrays with the Speex convention (consumed /
 * written on return). Equals n_streams independent calls of
 * speex_resampler_process_interleaved_int. Synchronous. */
SPXB_API int spxb_batch_process(spxb_batch *b, const int16_t *in, size_t in_stride_frames,
                                uint32_t *in_frames, int16_t *out,
                                size_t out_stride_frames, uint32_t *out_frames);

/* speex_resampler_process_interleaved_float for every stream of a float batch: HOST float
 * buffers, strides and lengths in frames as above. Lengths follow the float entry's block walk
 * (resample.c:927-963: no 1024-frame output block). Synchronous. */
SPXB_API int spxb_batch_process_f32(spxb_batch *b, const float *in, size_t in_stride_frames,
                                    uint32_t *in_frames, float *out, size_t out_stride_frames,
                                    uint32_t *out_frames);

/* Scaled float PCM (+-1.0 full scale) in and out on an int16 batch: converted to int16 as the
 * kernel loads it (round to nearest even of x*32768, saturated) and back as it stores; history,
 * arithmetic and lengths are the int16 path's. Strict kernel, bit-exact against the oracle fed the
 * converted samples. */
SPXB_API int spxb_batch_process_pcm_f32(spxb_batch *b, const float *in, size_t in_stride_frames,
                                        uint32_t *in_frames, float *out, size_t out_stride_frames,
                                        uint32_t *out_frames);

/* Strided call, all streams of the batch (backs the per-channel Speex entries): stream s reads
 * sample f of its channel c at in[s*in_stream_stride + f*in_step + c] and writes
 * out[s*out_stream_stride + m*out_step + c]; strides and steps count samples of the call's format
 * (float_io != 0: f32, float batch only). Samples between the strided ones are neither read nor
 * written. Synchronous; bit-exact (strict kernel). */
SPXB_API int spxb_batch_process_strided(spxb_batch *b, const void *in, size_t in_stream_stride,
                                        uint32_t in_step, uint32_t *in_frames, void *out,
                                        size_t out_stream_stride, uint32_t out_step,
                                        uint32_t *out_frames, int float_io);

/* Same, split for pipelining host<->device copies against the kernel: submit() stages
 * H2D + kernel + D2H on the batch's streams and returns a ticket at once (in_frames /
 * out_frames are final on return: the lengths are a pure function of the stream state);
 * wait() blocks until that ticket's output bytes are in `out`. Up to
 * spxb_batch_pipeline_depth() tickets may be in flight; `in`/`out` must stay valid and
 * (for full overlap) be pinned, e.g. from spxb_host_alloc. */
SPXB_API int spxb_batch_submit(spxb_batch *b, const int16_t *in, size_t in_stride_frames,
                               uint32_t *in_frames, int16_t *out, size_t out_stride_frames,
                               uint32_t *out_frames, uint64_t *ticket);
SPXB_API int spxb_batch_wait(spxb_batch *b, uint64_t ticket);
SPXB_API int spxb_batch_pipeline_depth(const spxb_batch *b);

/* DEVICE buffers (already in HBM; e.g. the output of a GPU decoder). Asynchronous on the
 * batch's compute stream; lengths as above (host arrays, final on return). */
SPXB_API int spxb_batch_process_device(spxb_batch *b, const int16_t *d_in,
                                       size_t in_stride_frames, uint32_t *in_frames,
                                       int16_t *d_out, size_t out_stride_frames,
                                       uint32_t *out_frames);
/* Uniform variant: every stream gets n_in frames and capacity out_cap; returns the
 * (common or last-stream) consumed/written through *in_used / *out_written (may be NULL).
 * No per-stream host arrays are touched when all streams share one state. */
SPXB_API int spxb_batch_process_device_uniform(spxb_batch *b, const int16_t *d_in,
                                               size_t in_stride_frames, uint32_t n_in,
                                               int16_t *d_out, size_t out_stride_frames,
                                               uint32_t out_cap, uint32_t *in_used,
                                               uint32_t *out_written);

/* A queue of hops already resident in HBM: step k (first_step <= k < first_step + steps)
 * reads slot k % ring of d_in (slots in_slot_elems int16 apart) and writes slot k % ring of
 * d_out, every stream n_in frames in / out_cap capacity, state carried hop to hop. One
 * uniform call per hop, launched back to back on the compute stream with no host sync. */
SPXB_API int spxb_batch_process_device_ring(spxb_batch *b, const int16_t *d_in,
                                            size_t in_stride_frames, size_t in_slot_elems,
                                            int16_t *d_out, size_t out_stride_frames,
                                            size_t out_slot_elems, uint32_t ring, uint32_t n_in,
                                            uint32_t out_cap, uint32_t first_step, uint32_t steps);

/* the one-stream batch behind a SpeexResamplerState (state migration into a larger batch) */
SPXB_API spxb_batch *spxb_resampler_batch(SpeexResamplerState *st);

/* run the batch's kernels on a caller-owned cudaStream_t (e.g. torch's current stream;
 * NULL is the legacy default stream); spxb_batch_use_own_stream goes back to the batch's own */
SPXB_API int spxb_batch_set_stream(spxb_batch *b, void *cuda_stream);
SPXB_API int spxb_batch_use_own_stream(spxb_batch *b);
SPXB_API int spxb_batch_synchronize(spxb_batch *b);

/* stream state (checkpoint / resume). history is interleaved
 * [filt_len-1][channels] int16. Both synchronise the batch first. */
SPXB_API int spxb_batch_get_state(spxb_batch *b, uint32_t stream, int32_t *last_sample,
                                  uint32_t *samp_frac_num, uint32_t *magic_samples,
                                  int16_t *history);
SPXB_API int spxb_batch_set_state(spxb_batch *b, uint32_t stream, int32_t last_sample,
                                  uint32_t samp_frac_num, const int16_t *history);
/* float view of the history, for both kinds of batch (set on an int16 batch requires
 * int16-valued samples) */
SPXB_API int spxb_batch_get_state_f32(spxb_batch *b, uint32_t stream, int32_t *last_sample,
                                      uint32_t *samp_frac_num, uint32_t *magic_samples,
                                      float *history);
SPXB_API int spxb_batch_set_state_f32(spxb_batch *b, uint32_t stream, int32_t last_sample,
                                      uint32_t samp_frac_num, const float *history);
SPXB_API int spxb_batch_reset(spxb_batch *b);      /* resample.c:1208 for every stream */
SPXB_API int spxb_batch_skip_zeros(spxb_batch *b); /* resample.c:1200 for every stream */

/* counters since creation: kernels launched by this library for this batch, bytes copied */
typedef struct {
  uint64_t kernel_launches;
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  uint64_t calls;
} spxb_counters;
SPXB_API int spxb_batch_counters(const spxb_batch *b, spxb_counters *c);

/* pinned host memory for the zero-staging path */
SPXB_API void *spxb_host_alloc(size_t bytes);
SPXB_API void spxb_host_free(void *p);

/* Measurement aid (bench.py): achieved FP32 rate, in FLOP/s (2 per FMA), of a register-resident
 * FFMA loop on the current device -- the measured denominator of the FP32-FMA roofline SURVEY 8d
 * asks for (MEASURED_PEAKS.json has no fp32 entry). `iters` <= 0 picks a default. */
SPXB_API double spxb_measure_fp32_peak(int iters);

/* ------------------------------------------------------------------ */
/* Part 3: host-only introspection (no GPU needed)                     */
/* ------------------------------------------------------------------ */

/* what update_filter (resample.c:605-701) derives for a fresh resampler */
typedef struct {
  uint32_t num, den;      /* gcd-reduced in/out (resample.c:1125-1128) */
  uint32_t filt_len;      /* N */
  uint32_t oversample;
  int32_t int_advance, frac_advance;
  float cutoff;
  int32_t use_direct;     /* 1: den*N per-phase table, 0: oversample*N+8 prototype */
  int32_t use_double;     /* 1: quality > 8 (f64 accumulators in the reference) */
  uint32_t table_len;     /* floats in the reference-layout sinc table */
} spxb_filter_info;

SPXB_API int spxb_filter_describe(uint32_t in_rate, uint32_t out_rate, int quality,
                                  spxb_filter_info *info);
/* the sinc table in the reference's own layout (bit-identical to st->sinc_table);
 * dst has room for `cap` floats; returns floats written or -err */
SPXB_API long spxb_filter_table(uint32_t in_rate, uint32_t out_rate, int quality, float *dst,
                                size_t cap);
/* per-phase taps h[phase][j] (den*N floats) the tiled kernel contracts with */
SPXB_API long spxb_filter_phase_taps(uint32_t in_rate, uint32_t out_rate, int quality,
                                     float *dst, size_t cap);

/* tensor kernel (SPXB_KERNEL_TENSOR): the per-phase taps as signed 24-bit fixed point,
 * h[phase][j] = round(tap * 2^shift) with |h| <= 8355711 (three balanced base-256 digits) */
SPXB_API long spxb_filter_fixed_taps(uint32_t in_rate, uint32_t out_rate, int quality,
                                     int32_t *dst, size_t cap, int *shift);
/* output tiles of one uniform call at stream position (last_sample, samp_frac_num): per tile
 * {first output m0, first frame of the K axis kf0 (history is < 0), phase of m0, q0 - kf0};
 * dst has room for cap_tiles x 4 ints; *ksteps = 32-frame MMA steps per tile. Returns tiles. */
SPXB_API long spxb_tensor_plan(uint32_t in_rate, uint32_t out_rate, int quality,
                               int32_t last_sample, uint32_t samp_frac_num, uint32_t n_out,
                               uint32_t nt, int32_t *dst, size_t cap_tiles, uint32_t *ksteps);
/* the int8 tap tile of (phase0, delta) exactly as the kernel consumes it:
 * [chunk < 2*ksteps][row < 3*nt][16 B], rows = digit d2 | d1 | d0 of each of the nt outputs */
SPXB_API long spxb_tensor_tap_tile(uint32_t in_rate, uint32_t out_rate, int quality, uint32_t nt,
                                   uint32_t phase0, uint32_t delta, int8_t *dst, size_t cap);
/* The packed ("resident") tap-tile layout of the persistent tensor kernel: per K step only the
 * 16-column blocks of each tap digit that can be non-zero are stored (csrc/umma_plan.h). Per K
 * step 18 words: byte offset / 16, rows per chunk, MMA entries, 3 x (first row, columns, first
 * accumulator column of the hi plane), first block of d2/d1/d0, end block of d2/d1/d0.
 * Returns the number of K steps (or -error); *tile_bytes receives the packed tile size. */
SPXB_API long spxb_tensor_packed_plan(uint32_t in_rate, uint32_t out_rate, int quality, uint32_t nt,
                                      uint32_t *dst, size_t cap_words, uint32_t *tile_bytes);
SPXB_API long spxb_tensor_tap_tile_packed(uint32_t in_rate, uint32_t out_rate, int quality,
                                          uint32_t nt, uint32_t phase0, uint32_t delta, int8_t *dst,
                                          size_t cap);

/* lengths and next state of one speex_resampler_process_int call (resample.c:968-1036
 * with :878-902), without touching samples: a pure function of the stream position. */
typedef struct {
  uint32_t n_out;        /* frames written */
  uint32_t consumed;     /* input frames consumed */
  int32_t last_sample;   /* new last_sample */
  uint32_t samp_frac_num;/* new samp_frac_num */
} spxb_call_plan;
SPXB_API int spxb_plan_call(uint32_t in_rate, uint32_t out_rate, int32_t last_sample,
                            uint32_t samp_frac_num, uint32_t n_in, uint32_t out_cap,
                            spxb_call_plan *plan);

/* the same for the float entry, whose output block is unbounded (resample.c:944) */
SPXB_API int spxb_plan_call_f32(uint32_t in_rate, uint32_t out_rate, int32_t last_sample,
                                uint32_t samp_frac_num, uint32_t n_in, uint32_t out_cap,
                                spxb_call_plan *plan);

/* the general form: `magic` pending samples in front of the input (what a filter change that
 * shortened the filter leaves behind, resample.c:759-776, :904-922), either entry's block walk,
 * and the input block of a state whose memory outgrew its filter (in_block = mem_alloc_size -
 * (filt_len - 1), 160 for a fresh state). plan->consumed counts real input frames. */
SPXB_API int spxb_plan_call_ex(uint32_t in_rate, uint32_t out_rate, int32_t last_sample,
                               uint32_t samp_frac_num, uint32_t magic, uint32_t n_in,
                               uint32_t out_cap, int float_entry, uint32_t in_block,
                               spxb_call_plan *plan, uint32_t *magic_used);

SPXB_API const char *spxb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SPEEXB200_H */
