/*
 * speex_oracle.c -- TEST INFRASTRUCTURE ONLY (see speex_oracle.h).
 *
 * Scalar CPU restatement of the Speex resampler exactly as the reference's WASM
 * build configures it (scripts/build_emscripten.sh:18-19: FLOATING_POINT,
 * OUTSIDE_SPEEX): every sample value is float32, q9/q10 accumulate float32
 * products in float64. Compile with -ffp-contract=off and without -ffast-math so
 * that each C operation is one IEEE operation, like the WASM f32.mul / f32.add.
 *
 * Parity: pinned bit-for-bit to a native build of the reference's resample.c
 * (oracle/_ref/libspeex_ref.so) by tests/test_oracle.py; hashes frozen in
 * tests/golden/oracle_hashes.json.
 *
 * Each function names the reference lines it restates (paths relative to
 * /root/reference/deps/speex/).
 */
#include "speex_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORC_IN_BLOCK 160   /* st->buffer_size, resample.c:835 */
#define ORC_OUT_BLOCK 1024 /* FIXED_STACK_ALLOC without VAR_ARRAYS, resample.c:111 */

/* ---- Kaiser window lookup tables: numeric data from resample.c:148-192 ---- */
static const double win_k12[68] = {
    0.99859849, 1.00000000, 0.99859849, 0.99440475, 0.98745105, 0.97779076, 0.96549770,
    0.95066529, 0.93340547, 0.91384741, 0.89213598, 0.86843014, 0.84290116, 0.81573067,
    0.78710866, 0.75723148, 0.72629970, 0.69451601, 0.66208321, 0.62920216, 0.59606986,
    0.56287762, 0.52980938, 0.49704014, 0.46473455, 0.43304576, 0.40211431, 0.37206735,
    0.34301800, 0.31506490, 0.28829195, 0.26276832, 0.23854851, 0.21567274, 0.19416736,
    0.17404546, 0.15530766, 0.13794294, 0.12192957, 0.10723616, 0.09382272, 0.08164178,
    0.07063950, 0.06075685, 0.05193064, 0.04409466, 0.03718069, 0.03111947, 0.02584161,
    0.02127838, 0.01736250, 0.01402878, 0.01121463, 0.00886058, 0.00691064, 0.00531256,
    0.00401805, 0.00298291, 0.00216702, 0.00153438, 0.00105297, 0.00069463, 0.00043489,
    0.00025272, 0.00013031, 0.0000527734, 0.00001000, 0.00000000};
static const double win_k10[36] = {
    0.99537781, 1.00000000, 0.99537781, 0.98162644, 0.95908712, 0.92831446, 0.89005583,
    0.84522401, 0.79486424, 0.74011713, 0.68217934, 0.62226347, 0.56155915, 0.50119680,
    0.44221549, 0.38553619, 0.33194107, 0.28205962, 0.23636152, 0.19515633, 0.15859932,
    0.12670280, 0.09935205, 0.07632451, 0.05731132, 0.04193980, 0.02979584, 0.02044510,
    0.01345224, 0.00839739, 0.00488951, 0.00257636, 0.00115101, 0.00035515, 0.00000000,
    0.00000000};
static const double win_k8[36] = {
    0.99635258, 1.00000000, 0.99635258, 0.98548012, 0.96759014, 0.94302200, 0.91223751,
    0.87580811, 0.83439927, 0.78875245, 0.73966538, 0.68797126, 0.63451750, 0.58014482,
    0.52566725, 0.47185369, 0.41941150, 0.36897272, 0.32108304, 0.27619388, 0.23465776,
    0.19672670, 0.16255380, 0.13219758, 0.10562887, 0.08273982, 0.06335451, 0.04724088,
    0.03412321, 0.02369490, 0.01563093, 0.00959968, 0.00527363, 0.00233883, 0.00050000,
    0.00000000};
static const double win_k6[36] = {
    0.99733006, 1.00000000, 0.99733006, 0.98935595, 0.97618418, 0.95799003, 0.93501423,
    0.90755855, 0.87598009, 0.84068475, 0.80211977, 0.76076565, 0.71712752, 0.67172623,
    0.62508937, 0.57774224, 0.53019925, 0.48295561, 0.43647969, 0.39120616, 0.34752997,
    0.30580127, 0.26632152, 0.22934058, 0.19505503, 0.16360756, 0.13508755, 0.10953262,
    0.08693120, 0.06722600, 0.05031820, 0.03607231, 0.02432151, 0.01487334, 0.00752000,
    0.00000000};

typedef struct {
  const double *lut;
  int lut_oversample;
} orc_window;

static const orc_window WIN6 = {win_k6, 32}, WIN8 = {win_k8, 32}, WIN10 = {win_k10, 32},
                        WIN12 = {win_k12, 64}; /* resample.c:199-206 */

/* quality -> (base length, oversample, down bw, up bw, window); resample.c:226-238 */
typedef struct {
  int base_len;
  int oversample;
  float bw_down, bw_up;
  const orc_window *win;
} orc_quality;

static const orc_quality QTAB[11] = {
    {8, 4, 0.830f, 0.860f, &WIN6},    {16, 4, 0.850f, 0.880f, &WIN6},
    {32, 4, 0.882f, 0.910f, &WIN6},   {48, 8, 0.895f, 0.917f, &WIN8},
    {64, 8, 0.921f, 0.940f, &WIN8},   {80, 16, 0.922f, 0.940f, &WIN10},
    {96, 16, 0.940f, 0.945f, &WIN10}, {128, 16, 0.950f, 0.950f, &WIN10},
    {160, 16, 0.960f, 0.960f, &WIN10}, {192, 32, 0.968f, 0.968f, &WIN12},
    {256, 32, 0.975f, 0.975f, &WIN12}};

struct orc_resampler {
  uint32_t in_rate, out_rate, num, den;
  int quality;
  uint32_t channels;
  uint32_t N;         /* filt_len */
  uint32_t work_len;  /* per-channel work buffer: N-1 history + ORC_IN_BLOCK */
  int int_adv, frac_adv;
  float cutoff;
  uint32_t oversample;
  int use_direct, use_double;
  int32_t *pos;       /* last_sample per channel */
  uint32_t *frac;     /* samp_frac_num per channel */
  float *work;        /* channels * work_len */
  float *table;
  uint32_t table_len;
};

/* Kaiser window value by cubic interpolation into the LUT; resample.c:240-258.
 * frac^3 is formed in float32, the polynomial and the 4-tap blend in float64. */
static double window_at(float x, const orc_window *w) {
  float y = x * w->lut_oversample;
  int idx = (int)floor(y);
  float t = y - idx;
  float t2 = t * t;
  float t3 = t * t * t;
  double c3 = -0.1666666667 * t + 0.1666666667 * t3;
  double c2 = t + 0.5 * t2 - 0.5 * t3;
  double c0 = -0.3333333333 * t + 0.5 * t2 - 0.1666666667 * t3;
  double c1 = 1.f - c3 - c2 - c0;
  return c0 * w->lut[idx] + c1 * w->lut[idx + 1] + c2 * w->lut[idx + 2] +
         c3 * w->lut[idx + 3];
}

/* one windowed-sinc tap; resample.c:288-298 (FLOATING_POINT branch) */
static float sinc_tap(float cutoff, float x, int N, const orc_window *w) {
  float xs = x * cutoff;
  if (fabs(x) < 1e-6) return cutoff;
  if (fabs(x) > .5 * N) return 0;
  return cutoff * sin(M_PI * xs) / (M_PI * xs) * window_at(fabs(2. * x / N), w);
}

/* resample.c:593-603 */
static int scale_u32(uint32_t *res, uint32_t v, uint32_t mul, uint32_t div) {
  uint32_t hi = v / div, lo = v % div;
  if (lo > UINT32_MAX / mul || hi > UINT32_MAX / mul ||
      hi * mul > UINT32_MAX - lo * mul / div)
    return ORC_ERR_OVERFLOW;
  *res = lo * mul / div + hi * mul;
  return ORC_OK;
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) { /* resample.c:1095-1105 */
  while (b) {
    uint32_t t = a % b;
    a = b;
    b = t;
  }
  return a;
}

/* Filter-bank derivation for a fresh (not yet started) resampler:
 * resample.c:605-726 minus the mid-stream length-change branches (:727-782),
 * which the TypeScript wrapper can never reach (it never calls set_rate /
 * set_quality after init). */
static int build_filter(orc_resampler *r) {
  const orc_quality *q = &QTAB[r->quality];
  r->int_adv = r->num / r->den;
  r->frac_adv = r->num % r->den;
  r->oversample = q->oversample;
  r->N = q->base_len;

  if (r->num > r->den) { /* down-sampling: widen the filter, lower the cutoff */
    r->cutoff = q->bw_down * r->den / r->num;
    if (scale_u32(&r->N, r->N, r->num, r->den) != ORC_OK) return ORC_ERR_ALLOC;
    r->N = ((r->N - 1) & (~0x7u)) + 8;
    if (2 * r->den < r->num) r->oversample >>= 1;
    if (4 * r->den < r->num) r->oversample >>= 1;
    if (8 * r->den < r->num) r->oversample >>= 1;
    if (16 * r->den < r->num) r->oversample >>= 1;
    if (r->oversample < 1) r->oversample = 1;
  } else {
    r->cutoff = q->bw_up;
  }

  /* smaller table wins; products deliberately in uint32 like the reference (:647) */
  r->use_direct = r->N * r->den <= r->N * r->oversample + 8 &&
                  INT_MAX / sizeof(float) / r->den >= r->N;
  r->use_double = r->quality > 8;
  if (r->use_direct) {
    r->table_len = r->N * r->den;
  } else {
    if ((INT_MAX / sizeof(float) - 8) / r->oversample < r->N) return ORC_ERR_ALLOC;
    r->table_len = r->N * r->oversample + 8;
  }
  r->table = (float *)malloc(sizeof(float) * r->table_len);
  if (!r->table) return ORC_ERR_ALLOC;

  if (r->use_direct) { /* :668-678 one row of N taps per output phase */
    for (uint32_t ph = 0; ph < r->den; ph++)
      for (int32_t j = 0; j < (int32_t)r->N; j++)
        r->table[ph * r->N + j] =
            sinc_tap(r->cutoff, (j - (int32_t)r->N / 2 + 1) - ((float)ph) / r->den, r->N,
                     q->win);
  } else { /* :689-691 one oversampled prototype, 4 guard taps each side */
    int32_t end = (int32_t)(r->oversample * r->N + 4);
    for (int32_t i = -4; i < end; i++)
      r->table[i + 4] =
          sinc_tap(r->cutoff, (i / (float)r->oversample - r->N / 2), r->N, q->win);
  }

  r->work_len = r->N - 1 + ORC_IN_BLOCK; /* :709 */
  if (INT_MAX / sizeof(float) / r->channels < r->work_len) return ORC_ERR_ALLOC;
  r->work = (float *)calloc((size_t)r->channels * r->work_len, sizeof(float));
  return r->work ? ORC_OK : ORC_ERR_ALLOC;
}

orc_resampler *orc_create(uint32_t channels, uint32_t in_rate, uint32_t out_rate,
                          int quality, int *err) {
  int e = ORC_OK;
  orc_resampler *r = NULL;
  /* resample.c:804-809 */
  if (channels == 0 || in_rate == 0 || out_rate == 0 || quality > 10 || quality < 0) {
    e = ORC_ERR_INVALID_ARG;
    goto done;
  }
  r = (orc_resampler *)calloc(1, sizeof(*r));
  if (!r) {
    e = ORC_ERR_ALLOC;
    goto done;
  }
  r->channels = channels;
  r->quality = quality;
  r->in_rate = in_rate;
  r->out_rate = out_rate;
  {
    uint32_t g = gcd_u32(in_rate, out_rate); /* :1125-1128 */
    r->num = in_rate / g;
    r->den = out_rate / g;
  }
  r->pos = (int32_t *)calloc(channels, sizeof(int32_t));
  r->frac = (uint32_t *)calloc(channels, sizeof(uint32_t));
  if (!r->pos || !r->frac) {
    e = ORC_ERR_ALLOC;
  } else {
    e = build_filter(r);
  }
  if (e != ORC_OK) {
    orc_destroy(r);
    r = NULL;
  }
done:
  if (err) *err = e;
  return r;
}

void orc_destroy(orc_resampler *r) {
  if (!r) return;
  free(r->pos);
  free(r->frac);
  free(r->work);
  free(r->table);
  free(r);
}

/* cubic blend weights between 4 neighbouring oversampled taps; resample.c:318-328 */
static void blend_weights(float t, float w[4]) {
  w[0] = -0.16667f * t + 0.16667f * t * t * t;
  w[1] = t + 0.5f * t * t - 0.5f * t * t * t;
  w[3] = -0.33333f * t + 0.5f * t * t - 0.16667f * t * t * t;
  w[2] = 1. - w[0] - w[1] - w[3];
}

/* Produce outputs from the channel's work buffer x[0 .. N-1+n_in) until the read
 * position passes n_in or `room` outputs exist. One routine for the four
 * reference kernels (resample.c:331-384, 389-435, 438-496, 501-558); the
 * summation order of each is kept exactly. Returns outputs written. */
static uint32_t run_kernel(orc_resampler *r, uint32_t ch, const float *x, uint32_t n_in,
                           float *y, uint32_t room) {
  const int N = (int)r->N;
  const uint32_t den = r->den, os = r->oversample;
  const float *T = r->table;
  int32_t pos = r->pos[ch];
  uint32_t fr = r->frac[ch];
  uint32_t made = 0;

  while (pos < (int32_t)n_in && made < room) {
    const float *seg = x + pos;
    float v;
    if (r->use_direct) {
      const float *h = T + (size_t)fr * N;
      if (!r->use_double) { /* :352 single float accumulator, j ascending */
        float s = 0;
        for (int j = 0; j < N; j++) s += h[j] * seg[j];
        v = s;
      } else { /* :409-417 four f64 lanes strided by 4, products rounded to f32 */
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int j = 0; j < N; j += 4) {
          a0 += h[j] * seg[j];
          a1 += h[j + 1] * seg[j + 1];
          a2 += h[j + 2] * seg[j + 2];
          a3 += h[j + 3] * seg[j + 3];
        }
        v = (float)(a0 + a1 + a2 + a3);
      }
    } else {
      const int off = (int)(fr * os / den);                /* :454 */
      const float t = ((float)((fr * os) % den)) / den;    /* :458 */
      const float *tp = T + 4 + os - off - 2;              /* tap k of input j at tp[j*os+k] */
      float w[4];
      if (!r->use_double) { /* :464-476 */
        float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int j = 0; j < N; j++) {
          const float s = seg[j];
          const float *c = tp + (size_t)j * os;
          a0 += s * c[0];
          a1 += s * c[1];
          a2 += s * c[2];
          a3 += s * c[3];
        }
        blend_weights(t, w);
        v = w[0] * a0 + w[1] * a1 + w[2] * a2 + w[3] * a3;
      } else { /* :527-539: f32 product (MULT16_16 casts, arch.h:180) into f64 sums */
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int j = 0; j < N; j++) {
          const float s = seg[j];
          const float *c = tp + (size_t)j * os;
          a0 += s * c[0];
          a1 += s * c[1];
          a2 += s * c[2];
          a3 += s * c[3];
        }
        blend_weights(t, w);
        v = (float)(w[0] * a0 + w[1] * a1 + w[2] * a2 + w[3] * a3);
      }
    }
    y[made++] = v;
    /* :372-378 advance the rational read position */
    pos += r->int_adv;
    fr += r->frac_adv;
    if (fr >= den) {
      fr -= den;
      pos++;
    }
  }
  r->pos[ch] = pos;
  r->frac[ch] = fr;
  return made;
}

/* arch.h:208-209 (float build): asymmetric saturation, round half up in f64 */
static int16_t to_int16(float v) {
  if (v < -32767.5f) return -32768;
  if (v > 32766.5f) return 32767;
  return (int16_t)floor(.5 + v);
}

/* One channel of a strided int16 stream; resample.c:968-1036 (int entry of the
 * float build) + :878-902 (process_native). The magic-sample branches are inert
 * here because the filter never changes after init. */
static void run_channel(orc_resampler *r, uint32_t ch, const int16_t *in, uint32_t stride,
                        uint32_t *in_frames, int16_t *out, uint32_t *out_frames) {
  float ybuf[ORC_OUT_BLOCK];
  float *x = r->work + (size_t)ch * r->work_len;
  const uint32_t hist = r->N - 1;
  uint32_t left_in = *in_frames, left_out = *out_frames;

  while (left_in && left_out) {
    uint32_t take = left_in > ORC_IN_BLOCK ? ORC_IN_BLOCK : left_in;
    uint32_t room = left_out > ORC_OUT_BLOCK ? ORC_OUT_BLOCK : left_out;
    if (in)
      for (uint32_t j = 0; j < take; j++) x[hist + j] = in[(size_t)j * stride];
    else
      for (uint32_t j = 0; j < take; j++) x[hist + j] = 0;

    uint32_t made = run_kernel(r, ch, x, take, ybuf, room);
    /* :891-899 commit only what the read position has passed, slide history */
    uint32_t used = take;
    if (r->pos[ch] < (int32_t)take) used = (uint32_t)r->pos[ch];
    r->pos[ch] -= (int32_t)used;
    for (uint32_t j = 0; j < hist; j++) x[j] = x[j + used];

    for (uint32_t j = 0; j < made; j++) out[(size_t)j * stride] = to_int16(ybuf[j]);
    left_in -= used;
    left_out -= made;
    out += (size_t)made * stride;
    if (in) in += (size_t)used * stride;
  }
  *in_frames -= left_in;
  *out_frames -= left_out;
}

/* One channel of a strided float stream; resample.c:927-963, the NATIVE entry of the float
 * build (speex_resampler_process_float): the same 160-frame input blocks, but an output block
 * is all the capacity that is left (there is no 1024-sample stack buffer on this path) and the
 * kernel's f32 results are stored as they are -- no scaling, rounding or saturation. */
static int run_channel_float(orc_resampler *r, uint32_t ch, const float *in, uint32_t stride,
                             uint32_t *in_frames, float *out, uint32_t *out_frames) {
  float *x = r->work + (size_t)ch * r->work_len;
  const uint32_t hist = r->N - 1;
  uint32_t left_in = *in_frames, left_out = *out_frames;
  float *ybuf = (float *)malloc(sizeof(float) * (left_out ? left_out : 1));
  if (!ybuf) return ORC_ERR_ALLOC;

  while (left_in && left_out) {
    uint32_t take = left_in > ORC_IN_BLOCK ? ORC_IN_BLOCK : left_in;
    if (in)
      for (uint32_t j = 0; j < take; j++) x[hist + j] = in[(size_t)j * stride];
    else
      for (uint32_t j = 0; j < take; j++) x[hist + j] = 0;

    uint32_t made = run_kernel(r, ch, x, take, ybuf, left_out); /* :944 ochunk = olen */
    uint32_t used = take;
    if (r->pos[ch] < (int32_t)take) used = (uint32_t)r->pos[ch];
    r->pos[ch] -= (int32_t)used;
    for (uint32_t j = 0; j < hist; j++) x[j] = x[j + used];

    for (uint32_t j = 0; j < made; j++) out[(size_t)j * stride] = ybuf[j];
    left_in -= used;
    left_out -= made;
    out += (size_t)made * stride;
    if (in) in += (size_t)used * stride;
  }
  free(ybuf);
  *in_frames -= left_in;
  *out_frames -= left_out;
  return ORC_OK;
}

int orc_process_interleaved_float(orc_resampler *r, const float *in, uint32_t *in_frames, float *out,
                                  uint32_t *out_frames) {
  /* resample.c:1038-1059 in the float build: channel loop as for int16 */
  const uint32_t n_in = *in_frames, cap = *out_frames;
  for (uint32_t c = 0; c < r->channels; c++) {
    *in_frames = n_in;
    *out_frames = cap;
    int e = run_channel_float(r, c, in ? in + c : NULL, r->channels, in_frames, out + c, out_frames);
    if (e) return e;
  }
  return ORC_OK;
}

int orc_process_interleaved_int16(orc_resampler *r, const int16_t *in, uint32_t *in_frames,
                                  int16_t *out, uint32_t *out_frames) {
  /* resample.c:1061-1082: every channel restarts from the caller's lengths; the
   * values left in *in_frames / *out_frames are those of the last channel */
  const uint32_t n_in = *in_frames, cap = *out_frames;
  for (uint32_t c = 0; c < r->channels; c++) {
    *in_frames = n_in;
    *out_frames = cap;
    run_channel(r, c, in ? in + c : NULL, r->channels, in_frames, out + c, out_frames);
  }
  return ORC_OK;
}

void orc_get_params(const orc_resampler *r, orc_params *p) {
  p->num = r->num;
  p->den = r->den;
  p->filt_len = r->N;
  p->oversample = r->oversample;
  p->int_advance = r->int_adv;
  p->frac_advance = r->frac_adv;
  p->cutoff = r->cutoff;
  p->use_direct = r->use_direct;
  p->use_double = r->use_double;
  p->table_len = r->table_len;
  p->channels = r->channels;
  p->quality = r->quality;
}

const float *orc_table(const orc_resampler *r) { return r->table; }

void orc_get_state(const orc_resampler *r, uint32_t channel, int32_t *last_sample,
                   uint32_t *samp_frac_num, const float **history) {
  if (last_sample) *last_sample = r->pos[channel];
  if (samp_frac_num) *samp_frac_num = r->frac[channel];
  if (history) *history = r->work + (size_t)channel * r->work_len;
}

/* src/index.ts:50-116 */
long orc_process_chunk(orc_resampler *r, double *cap_bytes_state, uint32_t in_rate,
                       uint32_t out_rate, const uint8_t *chunk, size_t nbytes,
                       uint8_t *out) {
  const uint32_t ch = r->channels;
  if (nbytes % (ch * 2u) != 0) return -100; /* index.ts:55-57 */
  /* index.ts:80-87 grow-only output staging, JS double arithmetic */
  double want = ceil((double)nbytes * (double)out_rate / (double)in_rate);
  if (*cap_bytes_state < want) *cap_bytes_state = want;
  uint32_t in_frames = (uint32_t)(nbytes / ch / 2);                    /* index.ts:90 */
  uint32_t out_frames = (uint32_t)(int32_t)(*cap_bytes_state / ch / 2); /* :95, i32 store truncates */
  int e = orc_process_interleaved_int16(r, (const int16_t *)chunk, &in_frames,
                                        (int16_t *)out, &out_frames);
  if (e != ORC_OK) return -(long)e;
  return (long)out_frames * ch * 2; /* index.ts:108-115; consumed count is ignored */
}

/* FNV-1a-64 of a byte string: the digest tests/golden freezes oracle outputs with */
uint64_t orc_fnv1a64(const uint8_t *p, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 0x100000001b3ull;
  return h;
}
