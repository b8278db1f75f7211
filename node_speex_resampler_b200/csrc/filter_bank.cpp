// filter_bank.cpp -- see filter_bank.h. Host only; compiled with -ffp-contract=off so
// every arithmetic operation below is exactly one IEEE operation.
#include "filter_bank.h"

#include <climits>
#include <cmath>
#include <cstddef>

namespace spxb {
namespace {

constexpr double kPi = 3.14159265358979323846;

// Kaiser window samples (numeric data of resample.c:148-192), one array per window
// family, sampled on [0, 1] with `density` points per unit plus guard points.
struct WindowLut {
  const double *samples;
  int density;
};

const double kKaiser12[68] = {
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
const double kKaiser10[36] = {
    0.99537781, 1.00000000, 0.99537781, 0.98162644, 0.95908712, 0.92831446, 0.89005583,
    0.84522401, 0.79486424, 0.74011713, 0.68217934, 0.62226347, 0.56155915, 0.50119680,
    0.44221549, 0.38553619, 0.33194107, 0.28205962, 0.23636152, 0.19515633, 0.15859932,
    0.12670280, 0.09935205, 0.07632451, 0.05731132, 0.04193980, 0.02979584, 0.02044510,
    0.01345224, 0.00839739, 0.00488951, 0.00257636, 0.00115101, 0.00035515, 0.00000000,
    0.00000000};
const double kKaiser8[36] = {
    0.99635258, 1.00000000, 0.99635258, 0.98548012, 0.96759014, 0.94302200, 0.91223751,
    0.87580811, 0.83439927, 0.78875245, 0.73966538, 0.68797126, 0.63451750, 0.58014482,
    0.52566725, 0.47185369, 0.41941150, 0.36897272, 0.32108304, 0.27619388, 0.23465776,
    0.19672670, 0.16255380, 0.13219758, 0.10562887, 0.08273982, 0.06335451, 0.04724088,
    0.03412321, 0.02369490, 0.01563093, 0.00959968, 0.00527363, 0.00233883, 0.00050000,
    0.00000000};
const double kKaiser6[36] = {
    0.99733006, 1.00000000, 0.99733006, 0.98935595, 0.97618418, 0.95799003, 0.93501423,
    0.90755855, 0.87598009, 0.84068475, 0.80211977, 0.76076565, 0.71712752, 0.67172623,
    0.62508937, 0.57774224, 0.53019925, 0.48295561, 0.43647969, 0.39120616, 0.34752997,
    0.30580127, 0.26632152, 0.22934058, 0.19505503, 0.16360756, 0.13508755, 0.10953262,
    0.08693120, 0.06722600, 0.05031820, 0.03607231, 0.02432151, 0.01487334, 0.00752000,
    0.00000000};

const WindowLut kWin6{kKaiser6, 32}, kWin8{kKaiser8, 32}, kWin10{kKaiser10, 32},
    kWin12{kKaiser12, 64};  // resample.c:199-206

// One row per quality 0..10 (resample.c:226-238).
struct QualityRow {
  uint32_t base_taps;
  uint32_t oversample;
  float down_bw, up_bw;
  const WindowLut *window;
};
const QualityRow kQuality[11] = {
    {8, 4, 0.830f, 0.860f, &kWin6},     {16, 4, 0.850f, 0.880f, &kWin6},
    {32, 4, 0.882f, 0.910f, &kWin6},    {48, 8, 0.895f, 0.917f, &kWin8},
    {64, 8, 0.921f, 0.940f, &kWin8},    {80, 16, 0.922f, 0.940f, &kWin10},
    {96, 16, 0.940f, 0.945f, &kWin10},  {128, 16, 0.950f, 0.950f, &kWin10},
    {160, 16, 0.960f, 0.960f, &kWin10}, {192, 32, 0.968f, 0.968f, &kWin12},
    {256, 32, 0.975f, 0.975f, &kWin12}};

uint32_t gcd32(uint32_t a, uint32_t b) {
  while (b != 0) {
    const uint32_t r = a % b;
    a = b;
    b = r;
  }
  return a;
}

// value * mul / div without 64-bit intermediates, refusing on 32-bit overflow exactly
// where the reference does (resample.c:593-603)
bool mul_div_u32(uint32_t value, uint32_t mul, uint32_t div, uint32_t *out) {
  const uint32_t whole = value / div, rest = value % div;
  if (rest > UINT32_MAX / mul || whole > UINT32_MAX / mul ||
      whole * mul > UINT32_MAX - rest * mul / div)
    return false;
  *out = rest * mul / div + whole * mul;
  return true;
}

// Window amplitude at x in [0,1]: 4-point cubic interpolation of the LUT.
// resample.c:240-258 -- the position and its powers are f32, the polynomial is f64.
double window_amplitude(float x, const WindowLut &lut) {
  const float scaled = x * lut.density;
  const int cell = static_cast<int>(std::floor(scaled));
  const float u = scaled - cell;
  const float u2 = u * u;
  const float u3 = u * u * u;
  const double w3 = -0.1666666667 * u + 0.1666666667 * u3;
  const double w2 = u + 0.5 * u2 - 0.5 * u3;
  const double w0 = -0.3333333333 * u + 0.5 * u2 - 0.1666666667 * u3;
  const double w1 = 1.f - w3 - w2 - w0;
  const double *s = lut.samples + cell;
  return w0 * s[0] + w1 * s[1] + w2 * s[2] + w3 * s[3];
}

// One tap of the windowed sinc at distance x (in input samples) from the centre.
// resample.c:288-298, FLOATING_POINT branch.
float windowed_sinc(float cutoff, float x, int taps, const WindowLut &lut) {
  const float arg = x * cutoff;
  const double dist = std::fabs(static_cast<double>(x));
  if (dist < 1e-6) return cutoff;
  if (dist > .5 * taps) return 0.f;
  const double v = cutoff * std::sin(kPi * arg) / (kPi * arg) *
                   window_amplitude(static_cast<float>(std::fabs(2. * x / taps)), lut);
  return static_cast<float>(v);
}

}  // namespace

int derive_filter_spec(uint32_t in_rate, uint32_t out_rate, int quality, FilterSpec *spec) {
  if (in_rate == 0 || out_rate == 0 || quality < 0 || quality > 10) return 3;
  FilterSpec s;
  s.in_rate = in_rate;
  s.out_rate = out_rate;
  s.quality = quality;
  const uint32_t g = gcd32(in_rate, out_rate);
  s.num = in_rate / g;
  s.den = out_rate / g;
  s.int_advance = static_cast<int32_t>(s.num / s.den);
  s.frac_advance = static_cast<int32_t>(s.num % s.den);

  const QualityRow &q = kQuality[quality];
  s.taps = q.base_taps;
  s.oversample = q.oversample;
  if (s.num > s.den) {
    // decimating: stretch the prototype by num/den (rounded up to a multiple of 8) and
    // thin the oversampling for large ratios (resample.c:618-635)
    s.cutoff = q.down_bw * s.den / s.num;
    if (!mul_div_u32(s.taps, s.num, s.den, &s.taps)) return 1;
    s.taps = ((s.taps - 1) & ~0x7u) + 8;
    for (uint32_t k = 2; k <= 16; k <<= 1)
      if (k * s.den < s.num) s.oversample >>= 1;
    if (s.oversample < 1) s.oversample = 1;
  } else {
    s.cutoff = q.up_bw;
  }
  // the smaller of the two table shapes, compared in wrapping uint32 like resample.c:647
  s.direct = (s.taps * s.den <= s.taps * s.oversample + 8) &&
             (INT_MAX / sizeof(float) / s.den >= s.taps);
  if (s.direct) {
    s.table_len = s.taps * s.den;
  } else {
    if ((INT_MAX / sizeof(float) - 8) / s.oversample < s.taps) return 1;
    s.table_len = s.taps * s.oversample + 8;
  }
  s.wide_accum = quality > 8;
  *spec = s;
  return 0;
}

std::vector<float> build_reference_table(const FilterSpec &s) {
  std::vector<float> table(s.table_len);
  const WindowLut &lut = *kQuality[s.quality].window;
  const int taps = static_cast<int>(s.taps);
  if (s.direct) {
    // resample.c:668-678: row `phase` holds the taps for a read position phase/den past
    // an input sample; tap j sits (j - N/2 + 1) samples from the centre
    for (uint32_t phase = 0; phase < s.den; ++phase) {
      const float shift = static_cast<float>(phase) / s.den;
      float *row = table.data() + static_cast<size_t>(phase) * s.taps;
      for (int j = 0; j < taps; ++j)
        row[j] = windowed_sinc(s.cutoff, (j - taps / 2 + 1) - shift, taps, lut);
    }
  } else {
    // resample.c:689-691: prototype sampled `oversample` times per input sample, with
    // 4 guard points on each side for the cubic blend
    const int32_t last = static_cast<int32_t>(s.oversample * s.taps + 4);
    for (int32_t i = -4; i < last; ++i)
      table[static_cast<size_t>(i + 4)] =
          windowed_sinc(s.cutoff, i / static_cast<float>(s.oversample) - s.taps / 2, taps, lut);
  }
  return table;
}

void cubic_weights(float t, float w[4]) {
  // resample.c:318-328 (float branch); w[2] is the f64 remainder so the four sum to 1
  w[0] = -0.16667f * t + 0.16667f * t * t * t;
  w[1] = t + 0.5f * t * t - 0.5f * t * t * t;
  w[3] = -0.33333f * t + 0.5f * t * t - 0.16667f * t * t * t;
  w[2] = static_cast<float>(1. - w[0] - w[1] - w[3]);
}

std::vector<float> build_phase_taps(const FilterSpec &s, const std::vector<float> &ref) {
  const size_t N = s.taps;
  if (s.direct) return ref;  // already one row per phase
  std::vector<float> taps(static_cast<size_t>(s.den) * N);
  for (uint32_t phase = 0; phase < s.den; ++phase) {
    // resample.c:454-458: which prototype cell the phase falls in, and how far into it
    const uint32_t scaled = phase * s.oversample;
    const uint32_t cell = scaled / s.den;
    const float t = static_cast<float>(scaled % s.den) / s.den;
    float w[4];
    cubic_weights(t, w);
    // input j meets prototype taps ref[4 + (j+1)*os - cell - 2 + k], k = 0..3 (:467-473)
    const float *p = ref.data() + 4 + s.oversample - cell - 2;
    float *row = taps.data() + static_cast<size_t>(phase) * N;
    for (size_t j = 0; j < N; ++j) {
      const float *c = p + j * s.oversample;
      const double v = static_cast<double>(w[0]) * c[0] + static_cast<double>(w[1]) * c[1] +
                       static_cast<double>(w[2]) * c[2] + static_cast<double>(w[3]) * c[3];
      row[j] = static_cast<float>(v);
    }
  }
  return taps;
}

}  // namespace spxb

namespace spxb {

uint32_t max_window_advance(uint32_t n, uint32_t num, uint32_t den) {
  return static_cast<uint32_t>((static_cast<uint64_t>(n) * num + den - 1) / den);
}

bool build_band_table(const FilterSpec &s, const std::vector<float> &taps, size_t max_bytes, BandTable *out) {
  const uint32_t N = s.taps;
  BandTable t;
  t.kp = (3 + max_window_advance(7, s.num, s.den) + N + 3) / 4 * 4;
  t.pad = (max_window_advance(8, s.num, s.den) + 3 + 3) / 4 * 4;
  t.row = t.kp + 2 * t.pad;
  const uint64_t floats = static_cast<uint64_t>(s.den) * 4 * 8 * t.row;
  if (floats * sizeof(float) > max_bytes) return false;
  t.data.assign(floats, 0.f);
  for (uint32_t p0 = 0; p0 < s.den; ++p0) {
    for (uint32_t al = 0; al < 4; ++al) {
      for (uint32_t r = 0; r < 8; ++r) {
        const uint64_t acc = static_cast<uint64_t>(p0) + static_cast<uint64_t>(r) * s.num;
        const uint32_t phase = static_cast<uint32_t>(acc % s.den);
        const uint32_t shift = al + static_cast<uint32_t>(acc / s.den);  // first tap's column
        float *dst = t.data.data() + ((static_cast<uint64_t>(p0) * 4 + al) * 8 + r) * t.row + t.pad + shift;
        const float *src = taps.data() + static_cast<size_t>(phase) * N;
        for (uint32_t j = 0; j < N; ++j) dst[j] = src[j];
      }
    }
  }
  *out = std::move(t);
  return true;
}

}  // namespace spxb
