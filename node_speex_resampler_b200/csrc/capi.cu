// capi.cu -- the five symbols the reference's WASM exports (scripts/build_emscripten.sh:20,
// bound at src/index.ts:6-16) plus the read-only rest of deps/speex/speex_resampler.h, each
// implemented over a one-stream device batch. Same signatures, length conventions and error
// codes as deps/speex/resample.c; the arithmetic runs on the GPU (no CPU path).
#include <algorithm>
#include <cstdint>
#include <new>
#include <vector>

#include "../../include/speexb200.h"
#include "call_plan.h"
#include "filter_bank.h"

#include <string>

namespace spxb {
const FilterSpec &batch_spec(const spxb_batch *b);
int batch_kernel_pref(const spxb_batch *b);
void batch_set_in_block(spxb_batch *b, uint32_t in_block);
void batch_force_plan(spxb_batch *b, const CallPlan *plan);
void batch_set_planar_state(spxb_batch *b, bool planar);
uint32_t batch_in_block(const spxb_batch *b);
void set_error(const std::string &msg);
}

struct SpeexResamplerState_ {
  spxb_batch *batch = nullptr;
  uint32_t in_rate = 0, out_rate = 0, channels = 0;
  uint32_t ratio_num = 0, ratio_den = 0;  // as given (the filter depends on their reduced ratio only)
  int quality = 0;
  bool started = false;  // a call has reached the resampling loop (resample.c:881 `st->started = 1`)
  // the reference's memory only grows (resample.c:709-719): samples per channel it has room for.
  // mem_alloc - (filt_len - 1) is the input block of its walk (160 unless a filter got shorter).
  uint32_t mem_alloc = 0;
  // "magic samples" (resample.c:759-776): frames a filter shortening left over, resampled before
  // the next input; interleaved, in the batch's sample format (one of the two vectors is used)
  uint32_t magic = 0;
  std::vector<int16_t> magic_i;
  std::vector<float> magic_f;
  std::vector<int16_t> joined_i;  // magic ++ input of the call being issued
  std::vector<float> joined_f;
  std::vector<int16_t> silence;  // stands in for in == NULL (resample.c:1007-1010)
  std::vector<float> fsilence;   // the same for the float entry (resample.c:950-952)
  // st->in_stride / st->out_stride of the reference (resample.c:836-837, :1170-1188): used by the
  // per-channel entries only (the interleaved ones override them with nb_channels, :1066-1068)
  uint32_t in_stride = 1, out_stride = 1;
  // A state that has used a per-channel entry is PLANAR: its batch holds one mono stream per channel,
  // each with its own position (the reference keeps last_sample / samp_frac_num / mem per channel).
  bool planar = false;
  std::vector<uint32_t> lens_in, lens_out;
};

extern "C" {

SpeexResamplerState *speex_resampler_init(uint32_t nb_channels, uint32_t in_rate, uint32_t out_rate,
                                          int quality, int *err) {
  return speex_resampler_init_frac(nb_channels, in_rate, out_rate, in_rate, out_rate, quality, err);
}

SpeexResamplerState *speex_resampler_init_frac(uint32_t nb_channels, uint32_t ratio_num, uint32_t ratio_den,
                                               uint32_t in_rate, uint32_t out_rate, int quality, int *err) {
  // resample.c:804-809: argument check precedes any allocation
  if (nb_channels == 0 || ratio_num == 0 || ratio_den == 0 || quality > 10 || quality < 0) {
    if (err) *err = RESAMPLER_ERR_INVALID_ARG;
    return nullptr;
  }
  SpeexResamplerState *st = new (std::nothrow) SpeexResamplerState_();
  if (!st) {
    if (err) *err = RESAMPLER_ERR_ALLOC_FAILED;
    return nullptr;
  }
  int e = 0;
  int device = 0;
  st->batch = spxb_batch_create(1, nb_channels, ratio_num, ratio_den, quality, device, &e);
  if (!st->batch) {
    delete st;
    if (err) *err = e ? e : RESAMPLER_ERR_ALLOC_FAILED;
    return nullptr;
  }
  st->in_rate = in_rate;
  st->out_rate = out_rate;
  st->ratio_num = ratio_num;
  st->ratio_den = ratio_den;
  st->channels = nb_channels;
  st->quality = quality;
  st->mem_alloc = spxb::batch_spec(st->batch).taps - 1 + spxb::kInBlock;  // resample.c:835, :709
  // The reference's own entry points return the reference's bytes: a single-stream state runs the
  // bit-exact kernel unless the caller opts into the tensor kernel
  // (spxb_batch_set_kernel(spxb_resampler_batch(st), SPXB_KERNEL_AUTO / _TENSOR): +-1 LSB).
  spxb_batch_set_kernel(st->batch, SPXB_KERNEL_STRICT);
  if (err) *err = RESAMPLER_ERR_SUCCESS;
  return st;
}

void speex_resampler_destroy(SpeexResamplerState *st) {
  if (!st) return;
  spxb_batch_destroy(st->batch);
  delete st;
}

void speex_resampler_get_rate(SpeexResamplerState *st, uint32_t *in_rate, uint32_t *out_rate) {
  *in_rate = st->in_rate;
  *out_rate = st->out_rate;
}

void speex_resampler_get_ratio(SpeexResamplerState *st, uint32_t *ratio_num, uint32_t *ratio_den) {
  const spxb::FilterSpec &s = spxb::batch_spec(st->batch);
  *ratio_num = s.num;
  *ratio_den = s.den;
}

void speex_resampler_get_quality(SpeexResamplerState *st, int *quality) { *quality = st->quality; }

int speex_resampler_get_input_latency(SpeexResamplerState *st) {
  return static_cast<int>(spxb::batch_spec(st->batch).taps / 2);
}

int speex_resampler_get_output_latency(SpeexResamplerState *st) {
  const spxb::FilterSpec &s = spxb::batch_spec(st->batch);
  return static_cast<int>(((s.taps / 2) * s.den + (s.num >> 1)) / s.num);
}

// Filter changes (resample.c:1107-1163 over update_filter's memory branches, :703-782).
//  * before the first sample has been resampled: the filter is rebuilt and its memory zeroed
//    (:721-725) -- a fresh batch for the new ratio / quality, keeping last_sample;
//  * mid-stream, same filter length (any ratio change while up-sampling: clock-drift correction):
//    only the table changes; history and last_sample stay, samp_frac_num is rescaled to the new
//    denominator (:1131-1140, multiply_frac :593-603);
//  * mid-stream, longer filter (:727-758 with no magic samples pending): the old history moves to
//    the end of the new one behind zeros and last_sample advances by half the growth;
//  * mid-stream, shorter filter: the surplus history stays behind as "magic samples" that the next
//    calls resample before their own input (:759-776, :904-922; process_with_magic below);
//  * any of these again while such samples are still pending (reshape_memory below).
// The reference's memory around a filter-length change, per channel (resample.c:727-776), as one
// sequence of frames V = history (N_old - 1 frames) ++ pending magic frames (M):
//  * shorter filter (:759-776): m = (N_old - N_new) / 2 frames move out of the history into the
//    magic region: history' = V[m, m + N_new - 1), magic' = the next m + M frames;
//  * longer filter (:727-758): the pending frames are first folded back "as if nothing had
//    happened" -- B = M zero frames ++ V, olen = N_old + 2M -- then, if N_new > olen, history' =
//    zeros ++ B and last_sample advances by (N_new - olen) / 2; otherwise m2 = (olen - N_new) / 2
//    frames of B turn into magic again: history' = B[m2, m2 + N_new - 1), magic' = the next m2.
// All channels of a state move together here, so the frames stay interleaved.
extern "C++" {
template <typename T>
static void reshape_memory(std::vector<T> &V, size_t ch, uint32_t n_old, uint32_t n_new, uint32_t M,
                           std::vector<T> *hist, std::vector<T> *magic, uint32_t *magic_frames, int32_t *last_sample) {
  auto frames = [&](const std::vector<T> &src, size_t f0, size_t n) {
    std::vector<T> out(n * ch, T(0));
    for (size_t f = 0; f < n; ++f)
      for (size_t c = 0; c < ch; ++c)
        if ((f0 + f) * ch + c < src.size()) out[f * ch + c] = src[(f0 + f) * ch + c];
    return out;
  };
  if (n_new < n_old) {
    const uint32_t m = (n_old - n_new) / 2;
    *hist = frames(V, m, n_new - 1);
    *magic = frames(V, static_cast<size_t>(m) + n_new - 1, static_cast<size_t>(m) + M);
    *magic_frames = m + M;
  } else {
    const uint32_t olen = n_old + 2 * M;
    std::vector<T> B(static_cast<size_t>(M) * ch, T(0));
    B.insert(B.end(), V.begin(), V.begin() + static_cast<size_t>(n_old - 1 + M) * ch);
    if (n_new > olen) {
      hist->assign(static_cast<size_t>(n_new - olen) * ch, T(0));
      hist->insert(hist->end(), B.begin(), B.end());
      magic->clear();
      *magic_frames = 0;
      *last_sample += static_cast<int32_t>((n_new - olen) / 2);
    } else {
      const uint32_t m2 = (olen - n_new) / 2;
      *hist = frames(B, m2, n_new - 1);
      *magic = frames(B, static_cast<size_t>(m2) + n_new - 1, m2);
      *magic_frames = m2;
    }
  }
}
}  // extern "C++"

static int refilter(SpeexResamplerState *st, uint32_t ratio_num, uint32_t ratio_den, int quality) {
  spxb::FilterSpec next;
  if (int e = spxb::derive_filter_spec(ratio_num, ratio_den, quality, &next)) return e;
  const spxb::FilterSpec old = spxb::batch_spec(st->batch);
  if (st->planar && st->started) {
    spxb::set_error("changing the filter mid-stream on a state that has used the per-channel entries is not supported");
    return RESAMPLER_ERR_BAD_STATE;
  }
  const bool f32 = spxb_batch_is_f32(st->batch) != 0;
  int32_t last = 0;
  uint32_t frac = 0, magic = 0;
  const size_t ch = st->channels;
  const size_t old_live = static_cast<size_t>(old.taps - 1) * ch;
  std::vector<float> hist_f(f32 ? old_live + 1 : 1), new_hist_f, new_magic_f;
  std::vector<int16_t> hist_i(f32 ? 1 : old_live + 1), new_hist_i, new_magic_i;
  int e = f32 ? spxb_batch_get_state_f32(st->batch, 0, &last, &frac, &magic, st->started ? hist_f.data() : nullptr)
              : spxb_batch_get_state(st->batch, 0, &last, &frac, &magic, st->started ? hist_i.data() : nullptr);
  if (e) return e;
  uint32_t new_magic = st->magic;
  const bool reshape = st->started && next.taps != old.taps;
  if (st->started) {
    // :1131-1140: samp_frac_num * den_new / den_old (it is < den_old, so the product fits 64 bits), clamped
    uint64_t scaled = static_cast<uint64_t>(frac) * next.den / old.den;
    if (frac != 0 && frac > 0xffffffffu / next.den) return RESAMPLER_ERR_OVERFLOW;  // multiply_frac's check
    if (scaled >= next.den) scaled = next.den - 1;
    frac = static_cast<uint32_t>(scaled);
    if (f32) {
      hist_f.resize(old_live);
      if (st->magic_f.empty() && !st->magic_i.empty()) st->magic_f.assign(st->magic_i.begin(), st->magic_i.end());
      if (reshape) {
        std::vector<float> V(hist_f);
        V.insert(V.end(), st->magic_f.begin(), st->magic_f.begin() + static_cast<size_t>(st->magic) * ch);
        reshape_memory<float>(V, ch, old.taps, next.taps, st->magic, &new_hist_f, &new_magic_f, &new_magic, &last);
      } else {
        new_hist_f = hist_f;  // same length: only the table changes, pending frames stay pending
        new_magic_f = st->magic_f;
      }
    } else {
      hist_i.resize(old_live);
      if (reshape) {
        std::vector<int16_t> V(hist_i);
        V.insert(V.end(), st->magic_i.begin(), st->magic_i.begin() + static_cast<size_t>(st->magic) * ch);
        reshape_memory<int16_t>(V, ch, old.taps, next.taps, st->magic, &new_hist_i, &new_magic_i, &new_magic, &last);
      } else {
        new_hist_i = hist_i;
        new_magic_i = st->magic_i;
      }
    }
  } else {
    frac = 0;
  }
  new_hist_f.resize(static_cast<size_t>(next.taps - 1) * ch + 1);
  new_hist_i.resize(static_cast<size_t>(next.taps - 1) * ch + 1);
  spxb_batch *nb = f32 ? spxb_batch_create_f32(1, st->channels, ratio_num, ratio_den, quality, 0, &e)
                       : spxb_batch_create(1, st->channels, ratio_num, ratio_den, quality, 0, &e);
  if (!nb) return e ? e : RESAMPLER_ERR_ALLOC_FAILED;
  e = f32 ? spxb_batch_set_state_f32(nb, 0, last, frac, st->started ? new_hist_f.data() : nullptr)
          : spxb_batch_set_state(nb, 0, last, frac, st->started ? new_hist_i.data() : nullptr);
  if (e) {
    spxb_batch_destroy(nb);
    return e;
  }
  spxb_batch_set_kernel(nb, spxb::batch_kernel_pref(st->batch));
  spxb_batch_destroy(st->batch);
  st->batch = nb;
  if (st->started) {
    st->magic = new_magic;
    if (f32) {
      st->magic_f = new_magic_f;
      st->magic_i.clear();
    } else {
      st->magic_i = new_magic_i;
    }
  }
  st->mem_alloc = std::max(st->mem_alloc, next.taps - 1 + spxb::kInBlock);  // :709-719: never shrinks
  spxb::batch_set_in_block(nb, st->mem_alloc - (next.taps - 1));
  return RESAMPLER_ERR_SUCCESS;
}

// A call while magic samples are pending (resample.c:993-1016 int16 entry, :940-962 float entry):
// the pending frames go in front of the caller's input, the lengths follow the reference's walk
// (call_plan.h: plan_call_magic) and the kernels see one call over the joined input.
extern "C++" {
template <typename T>
static int process_with_magic(SpeexResamplerState *st, const T *in, uint32_t *in_len, T *out, uint32_t *out_len,
                              std::vector<T> &magic, std::vector<T> &joined, bool float_entry) {
  const spxb::FilterSpec &s = spxb::batch_spec(st->batch);
  const size_t ch = st->channels;
  int32_t last = 0;
  uint32_t frac = 0, mg = 0;
  if (int e = spxb_batch_get_state(st->batch, 0, &last, &frac, &mg, nullptr)) return e;
  spxb::StreamPos pos;
  pos.last_sample = last;
  pos.samp_frac_num = frac;
  const uint32_t in_block = st->mem_alloc - (s.taps - 1);
  const spxb::MagicPlan mp = spxb::plan_call_magic(s.num, s.den, pos, st->magic, *in_len, *out_len, float_entry, in_block);
  joined.assign(magic.begin(), magic.begin() + static_cast<size_t>(st->magic) * ch);
  joined.insert(joined.end(), in, in + static_cast<size_t>(*in_len) * ch);
  spxb::CallPlan forced = mp.plan;
  forced.consumed = mp.magic_used + mp.plan.consumed;  // what the history slides by
  uint32_t n_total = st->magic + *in_len, n_cap = *out_len;
  spxb::batch_force_plan(st->batch, &forced);
  const int e = float_entry
                    ? spxb_batch_process_f32(st->batch, reinterpret_cast<const float *>(joined.data()), n_total, &n_total,
                                             reinterpret_cast<float *>(out), n_cap, &n_cap)
                    : spxb_batch_process(st->batch, reinterpret_cast<const int16_t *>(joined.data()), n_total, &n_total,
                                         reinterpret_cast<int16_t *>(out), n_cap, &n_cap);
  spxb::batch_force_plan(st->batch, nullptr);
  if (e) return e;
  magic.erase(magic.begin(), magic.begin() + static_cast<size_t>(mp.magic_used) * ch);
  st->magic -= mp.magic_used;
  *in_len = mp.plan.consumed;
  *out_len = mp.plan.n_out;
  return RESAMPLER_ERR_SUCCESS;
}
}  // extern "C++"

// ---- per-channel entries (resample.c:925-1036) ------------------------------------------------
// The interleaved batch keeps ONE position for all channels of the stream (they move together as
// long as only the interleaved entries are used). The first per-channel call splits it: one mono
// stream per channel, same history, same position, free to diverge from then on.
static int to_planar(SpeexResamplerState *st, bool want_f32) {
  const bool f32 = spxb_batch_is_f32(st->batch) != 0;
  if (st->planar && (f32 || !want_f32)) return RESAMPLER_ERR_SUCCESS;
  if (st->magic != 0) {
    spxb::set_error("per-channel calls while magic samples are pending are not supported");
    return RESAMPLER_ERR_BAD_STATE;
  }
  const spxb::FilterSpec s = spxb::batch_spec(st->batch);
  const uint32_t ch = st->channels, live = s.taps - 1;
  const uint32_t src_streams = st->planar ? ch : 1, src_ch = st->planar ? 1 : ch;
  const bool to_f32 = f32 || want_f32;
  int e = 0;
  spxb_batch *nb = to_f32 ? spxb_batch_create_f32(ch, 1, st->ratio_num, st->ratio_den, st->quality, 0, &e)
                          : spxb_batch_create(ch, 1, st->ratio_num, st->ratio_den, st->quality, 0, &e);
  if (!nb) return e ? e : RESAMPLER_ERR_ALLOC_FAILED;
  std::vector<float> hist(static_cast<size_t>(live) * src_ch + 1), mono(live + 1);
  for (uint32_t src = 0; src < src_streams && !e; ++src) {
    int32_t last = 0;
    uint32_t frac = 0, magic = 0;
    e = spxb_batch_get_state_f32(st->batch, src, &last, &frac, &magic, hist.data());
    for (uint32_t c = 0; c < src_ch && !e; ++c) {
      for (uint32_t j = 0; j < live; ++j) mono[j] = hist[static_cast<size_t>(j) * src_ch + c];
      e = spxb_batch_set_state_f32(nb, st->planar ? src : c, last, frac, mono.data());
    }
  }
  if (e) {
    spxb_batch_destroy(nb);
    return e;
  }
  spxb_batch_set_kernel(nb, SPXB_KERNEL_STRICT);
  spxb::batch_set_in_block(nb, st->mem_alloc - live);
  spxb::batch_set_planar_state(nb, true);
  spxb_batch_destroy(st->batch);
  st->batch = nb;
  st->planar = true;
  return RESAMPLER_ERR_SUCCESS;
}

extern "C++" {
// one call on a planar state: channel `only` alone (per-channel entries) or, with only == ~0u, every
// channel of an interleaved buffer (step nb_channels, channel c starting at element c)
template <typename T>
static int planar_call(SpeexResamplerState *st, uint32_t only, const T *in, uint32_t *in_len, T *out, uint32_t *out_len,
                       uint32_t in_step, uint32_t out_step, bool float_io) {
  const uint32_t ch = st->channels;
  st->lens_in.assign(ch, 0u);
  st->lens_out.assign(ch, 0u);
  for (uint32_t c = 0; c < ch; ++c)
    if (only == ~0u || c == only) {
      st->lens_in[c] = *in_len;
      st->lens_out[c] = *out_len;
    }
  // stream c starts at element c of an interleaved buffer; a per-channel call passes its own pointer
  const size_t stream_stride = only == ~0u ? 1 : 0;
  const int e = spxb_batch_process_strided(st->batch, in, stream_stride, in_step, st->lens_in.data(), out, stream_stride,
                                           out_step, st->lens_out.data(), float_io ? 1 : 0);
  if (e) return e;
  const uint32_t rep = only == ~0u ? ch - 1 : only;  // the interleaved entries report the last channel's lengths
  *in_len = st->lens_in[rep];
  *out_len = st->lens_out[rep];
  return RESAMPLER_ERR_SUCCESS;
}
}  // extern "C++"

int speex_resampler_set_rate_frac(SpeexResamplerState *st, uint32_t ratio_num, uint32_t ratio_den, uint32_t in_rate,
                                  uint32_t out_rate) {
  if (!st || ratio_num == 0 || ratio_den == 0) return RESAMPLER_ERR_INVALID_ARG;
  if (st->in_rate == in_rate && st->out_rate == out_rate && st->ratio_num == ratio_num && st->ratio_den == ratio_den)
    return RESAMPLER_ERR_SUCCESS;
  // the same reduced ratio keeps the filter (resample.c compares the given numbers, then reduces;
  // rebuilding an identical filter before the first sample is not observable)
  const spxb::FilterSpec &s = spxb::batch_spec(st->batch);
  const bool same_ratio = static_cast<uint64_t>(ratio_num) * s.den == static_cast<uint64_t>(ratio_den) * s.num;
  if (!same_ratio)
    if (int e = refilter(st, ratio_num, ratio_den, st->quality)) return e;
  st->in_rate = in_rate;
  st->out_rate = out_rate;
  st->ratio_num = ratio_num;
  st->ratio_den = ratio_den;
  return RESAMPLER_ERR_SUCCESS;
}

int speex_resampler_set_rate(SpeexResamplerState *st, uint32_t in_rate, uint32_t out_rate) {
  return speex_resampler_set_rate_frac(st, in_rate, out_rate, in_rate, out_rate);  // resample.c:1084-1087
}

int speex_resampler_set_quality(SpeexResamplerState *st, int quality) {
  if (!st || quality > 10 || quality < 0) return RESAMPLER_ERR_INVALID_ARG;
  if (st->quality == quality) return RESAMPLER_ERR_SUCCESS;
  if (int e = refilter(st, st->ratio_num, st->ratio_den, quality)) return e;
  st->quality = quality;
  return RESAMPLER_ERR_SUCCESS;
}

int speex_resampler_skip_zeros(SpeexResamplerState *st) { return spxb_batch_skip_zeros(st->batch); }

int speex_resampler_reset_mem(SpeexResamplerState *st) {
  st->magic = 0;  // resample.c:1213
  st->magic_i.clear();
  st->magic_f.clear();
  return spxb_batch_reset(st->batch);
}

int speex_resampler_process_interleaved_int(SpeexResamplerState *st, const int16_t *in, uint32_t *in_len,
                                            int16_t *out, uint32_t *out_len) {
  if (!st || !in_len || !out_len) return RESAMPLER_ERR_INVALID_ARG;
  if (*in_len == 0 || *out_len == 0) {
    // resample.c:988: the block loop does not run -- nothing is read, written or consumed, so the
    // buffers may be NULL (N-API hands out NULL for a zero-length Buffer)
    *in_len = 0;
    *out_len = 0;
    return RESAMPLER_ERR_SUCCESS;
  }
  if (!out) return RESAMPLER_ERR_INVALID_ARG;
  if (!in) {
    st->silence.assign(static_cast<size_t>(*in_len) * st->channels, 0);
    in = st->silence.data();
  }
  st->started = true;
  if (st->planar)  // every channel of the interleaved buffers, each from its own position (resample.c:1066-1078)
    return planar_call<int16_t>(st, ~0u, in, in_len, out, out_len, st->channels, st->channels, false);
  if (st->magic != 0) {
    if (spxb_batch_is_f32(st->batch)) {
      spxb::set_error("int16 call on a float-history state with magic samples pending is not supported");
      return RESAMPLER_ERR_BAD_STATE;
    }
    return process_with_magic<int16_t>(st, in, in_len, out, out_len, st->magic_i, st->joined_i, false);
  }
  // one stream: the strides are irrelevant, the lengths are the in-out cells
  return spxb_batch_process(st->batch, in, *in_len, in_len, out, *out_len, out_len);
}

int speex_resampler_process_interleaved_float(SpeexResamplerState *st, const float *in, uint32_t *in_len,
                                              float *out, uint32_t *out_len) {
  if (!st || !in_len || !out_len) return RESAMPLER_ERR_INVALID_ARG;
  if (*out_len == 0 || (*in_len == 0 && st->magic == 0)) {
    // resample.c:939-943: the loop does not run (pending magic samples are drained even without
    // input, :937-938, hence the second condition); buffers may be NULL
    *in_len = 0;
    *out_len = 0;
    return RESAMPLER_ERR_SUCCESS;
  }
  if (!out) return RESAMPLER_ERR_INVALID_ARG;
  if (st->planar) {
    if (int e = to_planar(st, true)) return e;
    if (!in) {
      st->fsilence.assign(static_cast<size_t>(*in_len) * st->channels, 0.f);
      in = st->fsilence.data();
    }
    st->started = true;
    return planar_call<float>(st, ~0u, in, in_len, out, out_len, st->channels, st->channels, true);
  }
  if (!spxb_batch_is_f32(st->batch)) {
    // first float call: the state moves to a float-history batch (int16 history converts exactly)
    int e = 0;
    spxb_batch *fb = spxb_batch_create_f32(1, st->channels, st->ratio_num, st->ratio_den, st->quality, 0, &e);
    if (!fb) return e ? e : RESAMPLER_ERR_ALLOC_FAILED;
    const spxb::FilterSpec &s = spxb::batch_spec(st->batch);
    std::vector<float> hist(static_cast<size_t>(s.taps ? s.taps - 1 : 0) * st->channels + 1);
    int32_t last = 0;
    uint32_t frac = 0, magic = 0;
    e = spxb_batch_get_state_f32(st->batch, 0, &last, &frac, &magic, hist.data());
    if (!e) e = spxb_batch_set_state_f32(fb, 0, last, frac, hist.data());
    if (e) {
      spxb_batch_destroy(fb);
      return e;
    }
    if (spxb::batch_kernel_pref(st->batch) == SPXB_KERNEL_STRICT) spxb_batch_set_kernel(fb, SPXB_KERNEL_STRICT);
    // the reference's memory never shrinks (resample.c:709-719): a state whose filter got shorter
    // keeps its enlarged input block on the float side too
    spxb::batch_set_in_block(fb, st->mem_alloc - (s.taps - 1));
    spxb_batch_destroy(st->batch);
    st->batch = fb;
  }
  if (!in) {
    st->fsilence.assign(static_cast<size_t>(*in_len) * st->channels, 0.f);
    in = st->fsilence.data();
  }
  if (*in_len != 0 && *out_len != 0) st->started = true;
  if (st->magic != 0) {
    if (st->magic_f.empty()) {  // the state turned float after the filter change
      st->magic_f.assign(st->magic_i.begin(), st->magic_i.end());
      st->magic_i.clear();
    }
    return process_with_magic<float>(st, in, in_len, out, out_len, st->magic_f, st->joined_f, true);
  }
  return spxb_batch_process_f32(st->batch, in, *in_len, in_len, out, *out_len, out_len);
}

int speex_resampler_process_int(SpeexResamplerState *st, uint32_t channel_index, const int16_t *in, uint32_t *in_len,
                                int16_t *out, uint32_t *out_len) {
  if (!st || !in_len || !out_len || channel_index >= st->channels) return RESAMPLER_ERR_INVALID_ARG;
  if (*in_len == 0 || *out_len == 0) {  // resample.c:988: the loop does not run
    *in_len = 0;
    *out_len = 0;
    return RESAMPLER_ERR_SUCCESS;
  }
  if (!out) return RESAMPLER_ERR_INVALID_ARG;
  if (int e = to_planar(st, false)) return e;
  uint32_t in_step = st->in_stride;
  if (!in) {  // resample.c:1007-1010: zeros instead of input
    st->silence.assign(*in_len, 0);
    in = st->silence.data();
    in_step = 1;
  }
  st->started = true;
  return planar_call<int16_t>(st, channel_index, in, in_len, out, out_len, in_step, st->out_stride, false);
}

int speex_resampler_process_float(SpeexResamplerState *st, uint32_t channel_index, const float *in, uint32_t *in_len,
                                  float *out, uint32_t *out_len) {
  if (!st || !in_len || !out_len || channel_index >= st->channels) return RESAMPLER_ERR_INVALID_ARG;
  if (*in_len == 0 || *out_len == 0) {  // resample.c:939
    *in_len = 0;
    *out_len = 0;
    return RESAMPLER_ERR_SUCCESS;
  }
  if (!out) return RESAMPLER_ERR_INVALID_ARG;
  if (int e = to_planar(st, true)) return e;
  uint32_t in_step = st->in_stride;
  if (!in) {
    st->fsilence.assign(*in_len, 0.f);
    in = st->fsilence.data();
    in_step = 1;
  }
  st->started = true;
  return planar_call<float>(st, channel_index, in, in_len, out, out_len, in_step, st->out_stride, true);
}

void speex_resampler_set_input_stride(SpeexResamplerState *st, uint32_t stride) { st->in_stride = stride; }
void speex_resampler_get_input_stride(SpeexResamplerState *st, uint32_t *stride) { *stride = st->in_stride; }
void speex_resampler_set_output_stride(SpeexResamplerState *st, uint32_t stride) { st->out_stride = stride; }
void speex_resampler_get_output_stride(SpeexResamplerState *st, uint32_t *stride) { *stride = st->out_stride; }

spxb_batch *spxb_resampler_batch(SpeexResamplerState *st) { return st ? st->batch : nullptr; }

const char *speex_resampler_strerror(int err) {
  // resample.c:1222-1239: five texts and a default (code 5 has no text of its own there)
  switch (err) {
    case RESAMPLER_ERR_SUCCESS:
      return "Success.";
    case RESAMPLER_ERR_ALLOC_FAILED:
      return "Memory allocation failed.";
    case RESAMPLER_ERR_BAD_STATE:
      return "Bad resampler state.";
    case RESAMPLER_ERR_INVALID_ARG:
      return "Invalid argument.";
    case RESAMPLER_ERR_PTR_OVERLAP:
      return "Input and output buffers overlap.";
    default:
      return "Unknown error. Bad error code or strange version mismatch.";
  }
}

}  // extern "C"
