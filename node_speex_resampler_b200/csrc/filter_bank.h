// filter_bank.h -- host-side windowed-sinc filter bank of the Speex resampler.
//
// Derives, for one (in_rate, out_rate, quality), everything the reference's
// update_filter (deps/speex/resample.c:605-701) derives for a fresh resampler, and
// generates the Kaiser-windowed sinc table with the reference's exact mixed f32/f64
// evaluation order (resample.c:240-258 compute_func, :288-298 sinc), so that the table
// uploaded to the GPU is bit-identical to st->sinc_table. Runs once per ratio, on the
// host; the result is uploaded once and shared by every stream of a batch.
#pragma once

#include <cstdint>
#include <vector>

namespace spxb {

struct FilterSpec {
  uint32_t in_rate = 0, out_rate = 0;
  uint32_t num = 0, den = 0;  // reduced ratio: num input steps per den output steps
  int quality = 0;
  uint32_t taps = 0;          // filt_len N
  uint32_t oversample = 0;
  int32_t int_advance = 0, frac_advance = 0;
  float cutoff = 0.f;
  bool direct = false;        // per-phase table (den rows of N) vs oversampled prototype
  bool wide_accum = false;    // quality > 8: reference accumulates in f64
  uint32_t table_len = 0;
};

// Speex error codes (speex_resampler.h:104-113) are returned as ints: 0 ok, 1 alloc,
// 3 invalid argument.
int derive_filter_spec(uint32_t in_rate, uint32_t out_rate, int quality, FilterSpec *spec);

// Table in the reference layout: direct -> T[phase*N + j]; otherwise T[i+4] for the
// oversampled prototype i in [-4, oversample*N+4).
std::vector<float> build_reference_table(const FilterSpec &spec);

// cubic blend weights of resample.c:318-328 for a fractional offset t in [0,1)
void cubic_weights(float t, float w[4]);

// Per-phase taps h[phase*N + j], j ascending over the input window, for every phase
// in [0, den): the direct table itself, or the cubic blend of the four neighbouring
// prototype taps folded into one tap (evaluated in f64, rounded once to f32).
std::vector<float> build_phase_taps(const FilterSpec &spec, const std::vector<float> &ref_table);

}  // namespace spxb

namespace spxb {

// Banded tap tiles for the streaming FIR kernel. A tile is 8 consecutive outputs whose first
// output has phase p0 and whose first window frame q0 has (q0 & 3) == al. Column k of every
// row multiplies window frame (q0 - al) + k, so row r (phase (p0 + r*num) % den, window
// start q0 + floor((p0 + r*num)/den)) holds its N taps shifted right by al + that advance,
// with zeros elsewhere. Rows are `row` floats long: `pad` zero columns, then columns
// 0 .. kp-1, then `pad` more, so neighbouring tiles of one warp can run a common column range.
// Layout: data[((p0*4 + al)*8 + r)*row + pad + k].
struct BandTable {
  std::vector<float> data;
  uint32_t kp = 0;   // columns that can hold a tap: 3 + max advance over 7 outputs + N, /4 up
  uint32_t pad = 0;  // zero margin on each side (multiple of 4)
  uint32_t row = 0;  // kp + 2*pad
};

// largest advance of the window start over n outputs: ceil(n*num/den)
uint32_t max_window_advance(uint32_t n, uint32_t num, uint32_t den);

// false when the table would exceed `max_bytes` (huge denominators): such batches are
// served by the strict kernel
bool build_band_table(const FilterSpec &spec, const std::vector<float> &phase_taps, size_t max_bytes,
                      BandTable *out);

}  // namespace spxb
