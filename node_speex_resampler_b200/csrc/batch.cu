// batch.cu -- host runtime behind the C ABI (include/speexb200.h): owns the CUDA streams,
// pinned staging, the uploaded filter bank and the device-resident per-stream state
// (last_sample, samp_frac_num, magic_samples, history) of a batch of streams.
//
// This file replaces what src/speex_wasm.js + the WASM heap staging of src/index.ts:59-115
// do in the reference (malloc'd in/out staging, length cells, state lifetime). There is no
// CPU fallback: without a CUDA device creation fails with RESAMPLER_ERR_ALLOC_FAILED.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/speexb200.h"
#include "call_plan.h"
#include "device_types.h"
#include "filter_bank.h"
#include "launch.h"
#include "umma_plan.h"

namespace spxb {

thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

#define SPXB_CUDA(expr)                                                               \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
      return RESAMPLER_ERR_ALLOC_FAILED;                                              \
    }                                                                                 \
  } while (0)

constexpr int kMaxPipelineSlots = 8;
// slots of the host-buffer pipeline; slots - 1 calls in flight. SPXB_PIPELINE_SLOTS overrides (2..8).
static int pipeline_slots() {
  static const int v = [] {
    const char *e = getenv("SPXB_PIPELINE_SLOTS");
    const int n = e ? atoi(e) : 4;
    return n < 2 ? 2 : n > kMaxPipelineSlots ? kMaxPipelineSlots : n;
  }();
  return v;
}

inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// one in-flight call: device staging + (optional) pinned bounce buffers + events
struct Slot {
  int16_t *d_in = nullptr;
  size_t d_in_cap = 0;  // int16 elements
  int16_t *d_out = nullptr;
  size_t d_out_cap = 0;
  int16_t *h_in = nullptr;  // pinned bounce for pageable / ragged input
  size_t h_in_cap = 0;
  int16_t *h_out = nullptr;
  size_t h_out_cap = 0;
  StreamCall *h_calls = nullptr;  // pinned, n_streams
  StreamCall *d_calls = nullptr;
  uint32_t *h_ids = nullptr;      // pinned, n_streams: stream-id lists of a ragged call's cohorts
  uint32_t *d_ids = nullptr;
  cudaEvent_t ev_h2d = nullptr, ev_kernel = nullptr, ev_done = nullptr;
  bool busy = false;
  uint64_t ticket = 0;
  uint32_t io_words = 1;  // int16 units per sample of the call in flight (2: float in/out)
  // deferred copy-out (pageable or ragged output)
  bool bounce_out = false;
  int16_t *user_out = nullptr;
  size_t user_out_stride = 0;  // elements
  size_t dev_out_stride = 0;   // elements
  std::vector<uint32_t> out_counts;  // frames per stream (ragged); empty => uniform
  uint32_t uniform_out = 0;
};

}  // namespace spxb

using namespace spxb;

struct spxb_batch {
  int device = 0;
  int sm_count = 148;
  FilterSpec spec;
  uint32_t n_streams = 0, channels = 0;
  // Sample format. A float batch (spxb_batch_create_f32) keeps its history as f32 -- what the
  // reference's `mem` is -- and serves float in/out (io_words == 2 during such a call) as well as
  // int16 in/out, bit-exactly, on the strict kernel. Buffers and strides stay in int16 units; a
  // float sample occupies two of them.
  bool f32 = false;
  uint32_t hist_words = 1;  // int16 units per history sample
  uint32_t in_block = kInBlock;           // input frames per block of the reference walk (call_plan.h)
  const CallPlan *forced_plan = nullptr;  // set by the C API around a call that has magic samples pending
  uint32_t io_words = 1;    // int16 units per in/out sample of the call being issued
  // strided call in progress (spxb_batch_process_strided): samples between consecutive frames of a
  // series in the caller's buffers; 0 = rows are channels-interleaved (every other entry)
  uint32_t in_step = 0, out_step = 0;
  // the streams of this batch are the CHANNELS of one Speex state that has used the per-channel
  // entries (capi.cu): reset_mem reproduces the reference's channel-major layout quirk across them
  bool planar_state = false;
  bool pcm_float = false;  // the call being issued takes / returns scaled float PCM (CallArgs::fmt 3)
  // filter bank in HBM
  float *d_table = nullptr, *d_taps = nullptr, *d_blend = nullptr, *d_band = nullptr;
  uint32_t band_kp = 0, band_pad = 0, band_row = 0;
  // stream state in HBM
  int16_t *d_hist[2] = {nullptr, nullptr};
  int hist_cur = 0;
  uint32_t hist_frames = 0, hist_stride = 0;
  int32_t *d_last_sample = nullptr;
  uint32_t *d_samp_frac = nullptr;
  uint32_t *d_magic = nullptr;
  // host shadow of the positions (lengths of a call are a pure function of these)
  std::vector<StreamPos> pos;
  bool uniform_pos = true;
  // streams
  cudaStream_t s_own = nullptr, s_compute = nullptr, s_in = nullptr, s_out = nullptr;
  bool external_stream = false;
  cudaEvent_t ev_state = nullptr;
  Slot slots[kMaxPipelineSlots];
  uint64_t next_ticket = 1;
  UmmaContext *umma = nullptr;  // tensor kernel state (nullptr: filter not covered)
  int kernel_pref = SPXB_KERNEL_AUTO;
  int last_kernel = SPXB_KERNEL_AUTO;
  spxb_counters counters{};
  // CUDA graphs of hop sequences (spxb_batch_process_device_ring): key -> instantiated graph.
  // `seen` remembers, per key, how many stream operations / allocations the planner needed the
  // last time the sequence ran launch by launch (only sequences that needed none are captured).
  struct RingGraph {
    cudaGraphExec_t exec = nullptr;
    uint64_t last_use = 0;
  };
  std::unordered_map<std::string, RingGraph> ring_graphs;
  std::unordered_map<std::string, uint64_t> ring_seen;
  uint64_t ring_clock = 0;
  bool dry_run = false;  // plan and advance the host-side state, launch nothing (graph replay)
  bool defer_pos_mirror = false;  // uniform hops update pos[0] only; mirrored at the end of the sequence
  // memo of the last uniform plan
  bool memo_valid = false;
  StreamPos memo_pos;
  uint32_t memo_n_in = 0, memo_cap = 0, memo_out_block = 0;
  CallPlan memo_plan;
};

namespace spxb {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

static CallPlan plan_memo(spxb_batch *b, StreamPos p, uint32_t n_in, uint32_t cap) {
  // the float entry has no 1024-sample output block (resample.c:944)
  // (scaled float PCM is the int16 path behind a format conversion: it keeps the int16 walk)
  const uint32_t out_block = b->io_words == 2 && !b->pcm_float ? kOutBlockUnbounded : kOutBlock;
  if (b->memo_valid && b->memo_pos.last_sample == p.last_sample &&
      b->memo_pos.samp_frac_num == p.samp_frac_num && b->memo_n_in == n_in &&
      b->memo_cap == cap && b->memo_out_block == out_block)
    return b->memo_plan;
  CallPlan pl = plan_call(b->spec.num, b->spec.den, p, n_in, cap, out_block, b->in_block);
  b->memo_out_block = out_block;
  b->memo_valid = true;
  b->memo_pos = p;
  b->memo_n_in = n_in;
  b->memo_cap = cap;
  b->memo_plan = pl;
  return pl;
}

static StreamCall to_stream_call(StreamPos p, uint32_t n_in, const CallPlan &pl) {
  StreamCall sc;
  sc.ls0 = p.last_sample;
  sc.frac0 = p.samp_frac_num;
  sc.n_in = n_in;
  sc.n_out = pl.n_out;
  sc.consumed = pl.consumed;
  sc.ls1 = pl.next.last_sample;
  sc.frac1 = pl.next.samp_frac_num;
  sc.pad_ = 0;
  return sc;
}

static bool is_pinned_or_device(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice ||
         at.type == cudaMemoryTypeManaged;
}

static int grow_device(int16_t **p, size_t *cap, size_t need) {
  if (*cap >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const size_t want = round_up(need + need / 4, 4096);
  SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(p), want * sizeof(int16_t)));
  *cap = want;
  return 0;
}

static int grow_pinned(int16_t **p, size_t *cap, size_t need) {
  if (*cap >= need) return 0;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *cap = 0;
  const size_t want = round_up(need + need / 4, 4096);
  SPXB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(p), want * sizeof(int16_t), cudaHostAllocDefault));
  *cap = want;
  return 0;
}

// Finish a slot: wait for its D2H, run the deferred bounce copy, mark it free.
static int retire_slot(spxb_batch *b, Slot &sl) {
  if (!sl.busy) return 0;
  SPXB_CUDA(cudaEventSynchronize(sl.ev_done));
  if (sl.bounce_out && sl.user_out) {
    const size_t ch = static_cast<size_t>(b->channels) * sl.io_words;
    if (sl.out_counts.empty()) {
      const size_t row = static_cast<size_t>(sl.uniform_out) * ch;
      for (uint32_t s = 0; s < b->n_streams; ++s)
        std::memcpy(sl.user_out + s * sl.user_out_stride, sl.h_out + s * sl.dev_out_stride,
                    row * sizeof(int16_t));
    } else {
      for (uint32_t s = 0; s < b->n_streams; ++s)
        std::memcpy(sl.user_out + s * sl.user_out_stride, sl.h_out + s * sl.dev_out_stride,
                    static_cast<size_t>(sl.out_counts[s]) * ch * sizeof(int16_t));
    }
  }
  sl.busy = false;
  sl.bounce_out = false;
  sl.user_out = nullptr;
  return 0;
}

// Build the kernel arguments for a call whose per-stream plans are already decided, launch
// the FIR (+ fused history slide) on the compute stream and flip the history ping-pong.
struct RaggedHost {  // host copy of a ragged call's per-stream plans + this slot's id-list buffers
  const StreamCall *calls = nullptr;
  uint32_t *h_ids = nullptr, *d_ids = nullptr;
};
static int launch_grouped(spxb_batch *b, const CallArgs &a, const RaggedHost &rh, uint32_t *launches, bool *done);

static int launch_call(spxb_batch *b, const int16_t *d_in, size_t in_stride_elems, int16_t *d_out,
                       size_t out_stride_elems, const StreamCall *d_calls,
                       const StreamCall &uniform, uint32_t max_n_out, const RaggedHost &rh = RaggedHost()) {
  CallArgs a;
  a.filt.num = b->spec.num;
  a.filt.den = b->spec.den;
  a.filt.taps = b->spec.taps;
  a.filt.oversample = b->spec.oversample;
  a.filt.direct = b->spec.direct ? 1 : 0;
  a.filt.wide_accum = b->spec.wide_accum ? 1 : 0;
  a.filt.table = b->d_table;
  a.filt.phase_taps = b->d_taps;
  a.filt.blend = b->d_blend;
  a.filt.band = b->d_band;
  a.filt.band_kp = b->band_kp;
  a.filt.band_pad = b->band_pad;
  a.filt.band_row = b->band_row;
  a.n_streams = b->n_streams;
  a.channels = b->channels;
  a.in = d_in;
  a.in_stride = in_stride_elems;
  a.out = d_out;
  a.out_stride = out_stride_elems;
  a.hist_src = b->d_hist[b->hist_cur];
  a.hist_dst = b->d_hist[b->hist_cur ^ 1];
  a.hist_stride = b->hist_stride;
  a.hist_frames = b->hist_frames;
  a.last_sample = b->d_last_sample;
  a.samp_frac = b->d_samp_frac;
  a.per_stream = d_calls;
  a.uniform = uniform;
  a.max_n_out = max_n_out;
  a.fmt = b->pcm_float ? 3u : !b->f32 ? 0u : (b->io_words == 2 ? 2u : 1u);
  a.ids = nullptr;
  a.n_ids = 0;
  a.in_step = b->in_step ? b->in_step : b->channels;
  a.out_step = b->out_step ? b->out_step : b->channels;
  const bool strided = a.in_step != b->channels || a.out_step != b->channels;

  uint32_t launches = 0;
  cudaError_t ce = cudaSuccess;
  int used = SPXB_KERNEL_STRICT;
  TiledConfig cfg;
  const int pref = b->kernel_pref;
  // float batches run the strict kernel only (the fast families are built around int16 history)
  // float batches and strided calls run the strict kernel only (the fast families are built around
  // int16 history and channels-interleaved rows)
  const bool want_tensor = !b->f32 && !b->pcm_float && !strided && (pref == SPXB_KERNEL_AUTO || pref == SPXB_KERNEL_TENSOR);
  const bool want_tiled = !b->f32 && !b->pcm_float && !strided && (pref == SPXB_KERNEL_AUTO || pref == SPXB_KERNEL_TILED);
  bool grouped = false;
  if (want_tensor && d_calls && rh.calls && max_n_out != 0 && b->umma && !b->dry_run) {
    // ragged batch: groups of streams that share one position run on the tensor kernel
    if (int e = launch_grouped(b, a, rh, &launches, &grouped)) return e;
  }
  if (grouped) {
    used = SPXB_KERNEL_TENSOR;
  } else if (want_tensor && max_n_out != 0 && umma_prepare(b->umma, a, b->s_compute, &ce)) {
    if (b->dry_run) launches += 1;
    else ce = launch_umma(b->umma, a, b->s_compute, &launches);
    used = SPXB_KERNEL_TENSOR;
  } else if (ce != cudaSuccess) {
    // planning failed on a CUDA error (not merely "not covered")
  } else if (pref == SPXB_KERNEL_TENSOR && max_n_out != 0 && !strided) {
    set_error("SPXB_KERNEL_TENSOR requested but this call does not qualify for the tensor kernel");
    return RESAMPLER_ERR_BAD_STATE;
  } else if (want_tiled && b->d_band && tiled_qualifies(a, b->sm_count, &cfg)) {
    if (b->dry_run) launches += 1;
    else ce = launch_tiled(a, cfg, b->s_compute, &launches);
    used = SPXB_KERNEL_TILED;
  } else if (pref == SPXB_KERNEL_TILED && max_n_out != 0 && !strided) {
    set_error("SPXB_KERNEL_TILED requested but this call does not qualify for the tiled kernel");
    return RESAMPLER_ERR_BAD_STATE;
  } else {
    if (b->dry_run) launches += 1;
    else ce = launch_strict(a, b->s_compute, &launches);
  }
  if (ce != cudaSuccess) {
    set_error(std::string("kernel launch: ") + cudaGetErrorString(ce));
    return RESAMPLER_ERR_BAD_STATE;
  }
  b->last_kernel = used;
  b->counters.kernel_launches += launches;
  b->hist_cur ^= 1;
  return 0;
}

// A ragged batch is rarely random: streams that were fed the same chunk sizes since they started
// move in cohorts that share (last_sample, samp_frac_num) and the call's lengths. Cohorts of at
// least kMinCohort streams (at most kMaxCohorts of them) each get one tensor-kernel launch over
// their stream-id list; whatever is left goes to the strict kernel in one launch over its list.
// *done stays false (nothing launched) when no cohort qualifies; the caller then takes the
// whole batch to the strict kernel as before.
static int launch_grouped(spxb_batch *b, const CallArgs &a, const RaggedHost &rh, uint32_t *launches, bool *done) {
  constexpr uint32_t kMinCohort = 32, kMaxCohorts = 8;
  *done = false;
  const uint32_t S = b->n_streams;
  const StreamCall *h_calls = rh.calls;
  struct Cohort {
    StreamCall call;
    std::vector<uint32_t> ids;
  };
  auto same = [](const StreamCall &x, const StreamCall &y) { return std::memcmp(&x, &y, sizeof(StreamCall)) == 0; };
  std::vector<Cohort> cohorts;
  std::unordered_map<uint64_t, std::vector<uint32_t>> by_hash;  // plan hash -> cohort indices
  uint32_t last = 0xffffffffu;  // neighbours are usually in the same cohort
  for (uint32_t s = 0; s < S; ++s) {
    // (streams that sit this call out form a cohort too: their history still has to cross to the
    // other half of the ping-pong, which the strict launch over the remainder does)
    const StreamCall &c = h_calls[s];
    if (last != 0xffffffffu && same(cohorts[last].call, c)) {
      cohorts[last].ids.push_back(s);
      continue;
    }
    uint64_t h = 0xcbf29ce484222325ull;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(&c);
    for (size_t k = 0; k < sizeof(StreamCall) / 4; ++k) h = (h ^ w[k]) * 0x100000001b3ull;
    std::vector<uint32_t> &cands = by_hash[h];
    uint32_t found = 0xffffffffu;
    for (uint32_t k : cands)
      if (same(cohorts[k].call, c)) found = k;
    if (found == 0xffffffffu) {
      found = static_cast<uint32_t>(cohorts.size());
      cands.push_back(found);
      cohorts.push_back(Cohort{c, {}});
      if (cohorts.size() > 4 * kMaxCohorts + S / kMinCohort) return 0;  // too fragmented to be worth grouping
    }
    cohorts[found].ids.push_back(s);
    last = found;
  }
  std::vector<uint32_t> big;
  for (uint32_t k = 0; k < cohorts.size(); ++k)
    if (cohorts[k].ids.size() >= kMinCohort && cohorts[k].call.n_out != 0) big.push_back(k);
  if (big.empty() || big.size() > kMaxCohorts) return 0;
  uint32_t *h_ids = rh.h_ids, *d_ids = rh.d_ids;
  // layout of the id list: [cohort big[0]] [cohort big[1]] ... [everything else]
  std::vector<uint32_t> offset(big.size() + 1, 0);
  uint32_t n = 0;
  std::vector<bool> is_big(cohorts.size(), false);
  for (uint32_t k = 0; k < big.size(); ++k) {
    is_big[big[k]] = true;
    offset[k] = n;
    for (uint32_t s : cohorts[big[k]].ids) h_ids[n++] = s;
  }
  offset[big.size()] = n;
  for (uint32_t k = 0; k < cohorts.size(); ++k)
    if (!is_big[k])
      for (uint32_t s : cohorts[k].ids) h_ids[n++] = s;
  const uint32_t n_listed = n;
  SPXB_CUDA(cudaMemcpyAsync(d_ids, h_ids, n_listed * sizeof(uint32_t), cudaMemcpyHostToDevice, b->s_compute));
  b->counters.h2d_bytes += n_listed * sizeof(uint32_t);

  std::vector<uint32_t> fallback;  // cohorts the tensor planner turned down
  for (uint32_t k = 0; k < big.size(); ++k) {
    CallArgs g = a;
    g.per_stream = nullptr;
    g.uniform = cohorts[big[k]].call;
    g.max_n_out = g.uniform.n_out;
    g.ids = d_ids + offset[k];
    g.n_ids = offset[k + 1] - offset[k];
    cudaError_t ce = cudaSuccess;
    if (umma_prepare(b->umma, g, b->s_compute, &ce)) {
      ce = launch_umma(b->umma, g, b->s_compute, launches);
    } else if (ce == cudaSuccess) {
      // not covered (e.g. the tile pool is full): the strict kernel takes this cohort too
      g.per_stream = a.per_stream;
      ce = launch_strict(g, b->s_compute, launches);
    }
    if (ce != cudaSuccess) {
      set_error(std::string("kernel launch (cohort): ") + cudaGetErrorString(ce));
      return RESAMPLER_ERR_BAD_STATE;
    }
  }
  if (n_listed > offset[big.size()]) {
    CallArgs r = a;
    r.ids = d_ids + offset[big.size()];
    r.n_ids = n_listed - offset[big.size()];
    const cudaError_t ce = launch_strict(r, b->s_compute, launches);
    if (ce != cudaSuccess) {
      set_error(std::string("kernel launch (remainder): ") + cudaGetErrorString(ce));
      return RESAMPLER_ERR_BAD_STATE;
    }
  }
  *done = true;
  return 0;
}

// Decide every stream's plan for (in_frames, out_frames); fills per-stream StreamCalls into
// `calls` when the batch is ragged. Updates the host shadow and the in/out length arrays.
struct Decided {
  bool uniform = true;
  StreamCall uni{};
  uint32_t max_n_in = 0, max_n_out = 0;
  bool any_work = false;
  StreamPos next_uniform{};  // uniform call: where every stream stands afterwards
};

// decide() / decide_uniform() only PLAN: the host shadow of the positions moves in
// commit_positions(), after the call's kernels have been launched successfully -- a call that
// fails (a forced kernel family that does not cover it, a CUDA launch error) leaves the shadow,
// the ping-pong half and the device state where they were.

static Decided decide_uniform(spxb_batch *b, uint32_t n_in, uint32_t cap) {
  Decided d;
  const StreamPos p = b->pos[0];
  const CallPlan pl = b->forced_plan ? *b->forced_plan : plan_memo(b, p, n_in, cap);
  d.uni = to_stream_call(p, n_in, pl);
  d.max_n_in = n_in;
  d.max_n_out = pl.n_out;
  // (a forced plan -- magic samples pending -- may consume without producing: resample.c:904-922)
  d.any_work = b->forced_plan ? (pl.consumed != 0 || pl.n_out != 0) : (n_in != 0 && cap != 0);
  d.next_uniform = pl.next;
  return d;
}

static void commit_positions(spxb_batch *b, const Decided &d, const StreamCall *calls) {
  if (!d.any_work) return;
  if (d.uniform) {
    // all shadows advance together; inside a hop sequence only pos[0] is kept exact and the
    // rest are mirrored once at its end (ring_hops)
    if (b->defer_pos_mirror) b->pos[0] = d.next_uniform;
    else for (auto &q : b->pos) q = d.next_uniform;
    return;
  }
  const uint32_t S = b->n_streams;
  for (uint32_t s = 0; s < S; ++s) {  // (a stream that sits the call out has ls1 == ls0, frac1 == frac0)
    b->pos[s].last_sample = calls[s].ls1;
    b->pos[s].samp_frac_num = calls[s].frac1;
  }
  // positions may have diverged
  b->uniform_pos = true;
  for (uint32_t s = 1; s < S && b->uniform_pos; ++s)
    b->uniform_pos = b->pos[s].last_sample == b->pos[0].last_sample &&
                     b->pos[s].samp_frac_num == b->pos[0].samp_frac_num;
}

static Decided decide(spxb_batch *b, uint32_t *in_frames, uint32_t *out_frames, StreamCall *calls) {
  const uint32_t S = b->n_streams;
  bool same = b->uniform_pos;
  for (uint32_t s = 1; same && s < S; ++s)
    same = in_frames[s] == in_frames[0] && out_frames[s] == out_frames[0];
  if (same) {
    Decided d = decide_uniform(b, in_frames[0], out_frames[0]);
    for (uint32_t s = 0; s < S; ++s) {
      in_frames[s] = d.uni.consumed;
      out_frames[s] = d.uni.n_out;
    }
    return d;
  }
  Decided d;
  d.uniform = false;
  for (uint32_t s = 0; s < S; ++s) {
    const StreamPos p = b->pos[s];
    const uint32_t n_in = in_frames[s], cap = out_frames[s];
    const CallPlan pl = plan_memo(b, p, n_in, cap);
    calls[s] = to_stream_call(p, n_in, pl);
    d.max_n_in = std::max(d.max_n_in, n_in);
    d.max_n_out = std::max(d.max_n_out, pl.n_out);
    if (n_in != 0 && cap != 0) d.any_work = true;
    in_frames[s] = pl.consumed;
    out_frames[s] = pl.n_out;
  }
  return d;
}

static void free_batch(spxb_batch *b) {
  if (!b) return;
  DeviceGuard g(b->device);
  cudaDeviceSynchronize();
  for (auto &sl : b->slots) {
    if (sl.d_in) cudaFree(sl.d_in);
    if (sl.d_out) cudaFree(sl.d_out);
    if (sl.h_in) cudaFreeHost(sl.h_in);
    if (sl.h_out) cudaFreeHost(sl.h_out);
    if (sl.h_calls) cudaFreeHost(sl.h_calls);
    if (sl.d_calls) cudaFree(sl.d_calls);
    if (sl.h_ids) cudaFreeHost(sl.h_ids);
    if (sl.d_ids) cudaFree(sl.d_ids);
    if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
    if (sl.ev_kernel) cudaEventDestroy(sl.ev_kernel);
    if (sl.ev_done) cudaEventDestroy(sl.ev_done);
  }
  for (auto &kv : b->ring_graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);

  if (b->ev_state) cudaEventDestroy(b->ev_state);
  if (b->d_table) cudaFree(b->d_table);
  if (b->d_taps) cudaFree(b->d_taps);
  if (b->d_blend) cudaFree(b->d_blend);
  if (b->d_band) cudaFree(b->d_band);
  umma_destroy(b->umma);
  if (b->d_hist[0]) cudaFree(b->d_hist[0]);
  if (b->d_hist[1]) cudaFree(b->d_hist[1]);
  if (b->d_last_sample) cudaFree(b->d_last_sample);
  if (b->d_samp_frac) cudaFree(b->d_samp_frac);
  if (b->d_magic) cudaFree(b->d_magic);
  if (b->s_own) cudaStreamDestroy(b->s_own);
  if (b->s_in) cudaStreamDestroy(b->s_in);
  if (b->s_out) cudaStreamDestroy(b->s_out);
  cudaGetLastError();
  delete b;
}

static int create_batch(spxb_batch *b) {
  SPXB_CUDA(cudaSetDevice(b->device));
  cudaDeviceProp prop;
  SPXB_CUDA(cudaGetDeviceProperties(&prop, b->device));
  b->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    set_error("libspeexb200 is built for sm_100a (B200) only; device is sm_" +
              std::to_string(prop.major) + std::to_string(prop.minor));
    return RESAMPLER_ERR_ALLOC_FAILED;
  }
  SPXB_CUDA(tiled_prepare_device());
  SPXB_CUDA(cudaStreamCreateWithFlags(&b->s_own, cudaStreamNonBlocking));
  SPXB_CUDA(cudaStreamCreateWithFlags(&b->s_in, cudaStreamNonBlocking));
  SPXB_CUDA(cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking));
  b->s_compute = b->s_own;
  SPXB_CUDA(cudaEventCreateWithFlags(&b->ev_state, cudaEventDisableTiming));
  for (auto &sl : b->slots) {
    SPXB_CUDA(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
    SPXB_CUDA(cudaEventCreateWithFlags(&sl.ev_kernel, cudaEventDisableTiming));
    SPXB_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
  }

  // filter bank: generated on the host exactly like update_filter, uploaded once
  const FilterSpec &sp = b->spec;
  std::vector<float> table = build_reference_table(sp);
  SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&b->d_table), table.size() * sizeof(float)));
  SPXB_CUDA(cudaMemcpy(b->d_table, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice));
  // per-phase taps only while the den*N table stays modest (64 MiB); beyond that the strict
  // kernel (which needs only the oversampled prototype) serves the batch
  const uint64_t phase_floats = static_cast<uint64_t>(sp.den) * sp.taps;
  if (!b->f32 && phase_floats * sizeof(float) <= (64ull << 20)) {
    std::vector<float> taps = build_phase_taps(sp, table);
    SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&b->d_taps), taps.size() * sizeof(float)));
    SPXB_CUDA(cudaMemcpy(b->d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    // pre-shifted tap tiles of the streaming kernel, while they stay modest too
    BandTable band;
    if (build_band_table(sp, taps, 64ull << 20, &band)) {
      SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&b->d_band), band.data.size() * sizeof(float)));
      SPXB_CUDA(cudaMemcpy(b->d_band, band.data.data(), band.data.size() * sizeof(float), cudaMemcpyHostToDevice));
      b->band_kp = band.kp;
      b->band_pad = band.pad;
      b->band_row = band.row;
    }
  }
  // tensor kernel: fixed-point taps + tap-tile pool (nullptr when the filter is not covered)
  if (!b->f32) b->umma = umma_create(sp, table, b->channels, b->sm_count);
  if (!sp.direct) {
    std::vector<float> blend(static_cast<size_t>(sp.den) * 4);
    for (uint32_t ph = 0; ph < sp.den; ++ph) {
      const uint32_t scaled = ph * sp.oversample;  // uint32 like resample.c:458
      cubic_weights(static_cast<float>(scaled % sp.den) / sp.den, &blend[static_cast<size_t>(ph) * 4]);
    }
    SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&b->d_blend), blend.size() * sizeof(float)));
    SPXB_CUDA(cudaMemcpy(b->d_blend, blend.data(), blend.size() * sizeof(float), cudaMemcpyHostToDevice));
  }

  // stream state, zeroed: resample.c:721-725 and the calloc'd per-channel arrays :838-843
  b->hist_frames = static_cast<uint32_t>(round_up(sp.taps - 1, 16));  // 16-frame K chunks (tensor kernel)
  b->hist_stride = static_cast<uint32_t>(
      round_up(static_cast<size_t>(b->hist_frames) * b->channels * b->hist_words, 8));
  const size_t hist_bytes = static_cast<size_t>(b->n_streams) * b->hist_stride * sizeof(int16_t);
  for (int i = 0; i < 2; ++i) {
    SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&b->d_hist[i]), std::max<size_t>(hist_bytes, 16)));
    SPXB_CUDA(cudaMemset(b->d_hist[i], 0, std::max<size_t>(hist_bytes, 16)));
  }
  const size_t nS = b->n_streams;
  SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&b->d_last_sample), nS * sizeof(int32_t)));
  SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&b->d_samp_frac), nS * sizeof(uint32_t)));
  SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&b->d_magic), nS * sizeof(uint32_t)));
  SPXB_CUDA(cudaMemset(b->d_last_sample, 0, nS * sizeof(int32_t)));
  SPXB_CUDA(cudaMemset(b->d_samp_frac, 0, nS * sizeof(uint32_t)));
  SPXB_CUDA(cudaMemset(b->d_magic, 0, nS * sizeof(uint32_t)));
  b->pos.assign(nS, StreamPos{});
  b->uniform_pos = true;
  SPXB_CUDA(cudaDeviceSynchronize());
  return 0;
}

// Shared body of submit(): host buffers in, ticket out.
static int submit_host(spxb_batch *b, const int16_t *in, size_t in_stride_frames, uint32_t *in_frames,
                       int16_t *out, size_t out_stride_frames, uint32_t *out_frames,
                       uint64_t *ticket) {
  DeviceGuard g(b->device);
  const uint32_t S = b->n_streams;
  const size_t ch = static_cast<size_t>(b->channels) * b->io_words;  // int16 units per frame
  Slot &sl = b->slots[b->next_ticket % pipeline_slots()];
  if (int e = retire_slot(b, sl)) return e;
  sl.io_words = b->io_words;

  // remember the offered lengths: decide() overwrites the arrays with consumed / written
  const bool offered_uniform = [&] {
    for (uint32_t s = 1; s < S; ++s)
      if (in_frames[s] != in_frames[0] || out_frames[s] != out_frames[0]) return false;
    return true;
  }();
  std::vector<uint32_t> offered_in;
  if (!offered_uniform) offered_in.assign(in_frames, in_frames + S);
  const uint32_t offered_in0 = in_frames[0];

  if (!sl.h_calls) {
    SPXB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_calls), S * sizeof(StreamCall), cudaHostAllocDefault));
    SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&sl.d_calls), S * sizeof(StreamCall)));
    SPXB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_ids), S * sizeof(uint32_t), cudaHostAllocDefault));
    SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&sl.d_ids), S * sizeof(uint32_t)));
  }
  Decided d = decide(b, in_frames, out_frames, sl.h_calls);

  // Device rows mirror densely packed pinned host rows (one flat DMA, no per-row pitch walk);
  // anything else is re-pitched to 16-byte rows.
  const bool dense_in = d.uniform && in_stride_frames == d.max_n_in && is_pinned_or_device(in);
  const bool dense_out = d.uniform && out_stride_frames == d.max_n_out && is_pinned_or_device(out);
  const size_t dev_in_stride = dense_in ? static_cast<size_t>(d.max_n_in) * ch
                                        : round_up(static_cast<size_t>(d.max_n_in) * ch, 8);
  const size_t dev_out_stride = dense_out ? static_cast<size_t>(d.max_n_out) * ch
                                          : round_up(static_cast<size_t>(d.max_n_out) * ch, 8);
  if (int e = grow_device(&sl.d_in, &sl.d_in_cap, std::max<size_t>(dev_in_stride * S, 8))) return e;
  if (int e = grow_device(&sl.d_out, &sl.d_out_cap, std::max<size_t>(dev_out_stride * S, 8))) return e;
  // size the idle slots alike now, so that no cudaMalloc (a device-wide sync) lands in the middle
  // of a pipelined run the first time each slot comes up
  for (int k = 0; k < pipeline_slots(); ++k) {
    Slot &o = b->slots[k];
    if (&o == &sl || o.busy) continue;
    if (int e = grow_device(&o.d_in, &o.d_in_cap, std::max<size_t>(dev_in_stride * S, 8))) return e;
    if (int e = grow_device(&o.d_out, &o.d_out_cap, std::max<size_t>(dev_out_stride * S, 8))) return e;
  }

  sl.ticket = b->next_ticket++;
  if (ticket) *ticket = sl.ticket;
  sl.busy = true;
  sl.bounce_out = false;
  sl.user_out = nullptr;
  sl.out_counts.clear();

  if (!d.any_work) {
    SPXB_CUDA(cudaEventRecord(sl.ev_done, b->s_out));
    return 0;
  }

  // Small calls (a single stream's 20 ms chunk is 3.5 KB) are latency, not bandwidth: both copies and
  // the kernel go down ONE stream with no event hand-overs between three (each costs several
  // microseconds of a ~25 us call); large calls keep the three-stream pipeline that overlaps the
  // copy of step k+1 with the kernel of step k and the read-back of step k-1.
  const size_t in_row_bytes = static_cast<size_t>(d.max_n_in) * ch * sizeof(int16_t);
  const size_t out_bytes_est = static_cast<size_t>(d.max_n_out) * ch * sizeof(int16_t) * S;
  const bool small_call = in_row_bytes * S + out_bytes_est <= (256u << 10);
  const cudaStream_t s_in = small_call ? b->s_compute : b->s_in;
  const cudaStream_t s_out = small_call ? b->s_compute : b->s_out;

  // ---- H2D ----
  const bool direct_in = d.uniform && is_pinned_or_device(in);
  if (dense_in) {
    SPXB_CUDA(cudaMemcpyAsync(sl.d_in, in, in_row_bytes * S, cudaMemcpyDefault, s_in));
  } else if (direct_in) {
    SPXB_CUDA(cudaMemcpy2DAsync(sl.d_in, dev_in_stride * sizeof(int16_t), in,
                                in_stride_frames * ch * sizeof(int16_t), in_row_bytes, S,
                                cudaMemcpyDefault, s_in));
  } else {
    if (int e = grow_pinned(&sl.h_in, &sl.h_in_cap, dev_in_stride * S)) return e;
    for (uint32_t s = 0; s < S; ++s) {
      const uint32_t n = offered_in.empty() ? offered_in0 : offered_in[s];
      std::memcpy(sl.h_in + s * dev_in_stride, in + s * in_stride_frames * ch, n * ch * sizeof(int16_t));
    }
    SPXB_CUDA(cudaMemcpyAsync(sl.d_in, sl.h_in, dev_in_stride * S * sizeof(int16_t),
                              cudaMemcpyHostToDevice, s_in));
  }
  b->counters.h2d_bytes += in_row_bytes * S;
  if (!d.uniform) {
    SPXB_CUDA(cudaMemcpyAsync(sl.d_calls, sl.h_calls, S * sizeof(StreamCall), cudaMemcpyHostToDevice, s_in));
    b->counters.h2d_bytes += S * sizeof(StreamCall);
  }
  if (!small_call) SPXB_CUDA(cudaEventRecord(sl.ev_h2d, s_in));

  // ---- kernel ----
  if (!small_call) SPXB_CUDA(cudaStreamWaitEvent(b->s_compute, sl.ev_h2d, 0));
  if (int e = launch_call(b, sl.d_in, dev_in_stride, sl.d_out, dev_out_stride,
                          d.uniform ? nullptr : sl.d_calls, d.uni, d.max_n_out,
                          d.uniform ? RaggedHost() : RaggedHost{sl.h_calls, sl.h_ids, sl.d_ids})) {
    // nothing was launched: the slot goes back (its H2D is harmless), the positions were never moved
    cudaEventRecord(sl.ev_done, s_in);
    return e;
  }
  commit_positions(b, d, sl.h_calls);
  if (!small_call) SPXB_CUDA(cudaEventRecord(sl.ev_kernel, b->s_compute));

  // ---- D2H ----
  if (!small_call) SPXB_CUDA(cudaStreamWaitEvent(s_out, sl.ev_kernel, 0));
  const size_t out_row_bytes = static_cast<size_t>(d.max_n_out) * ch * sizeof(int16_t);
  if (d.max_n_out != 0) {
    const bool direct_out = d.uniform && is_pinned_or_device(out);
    if (dense_out) {
      SPXB_CUDA(cudaMemcpyAsync(out, sl.d_out, out_row_bytes * S, cudaMemcpyDefault, s_out));
    } else if (direct_out) {
      SPXB_CUDA(cudaMemcpy2DAsync(out, out_stride_frames * ch * sizeof(int16_t), sl.d_out,
                                  dev_out_stride * sizeof(int16_t), out_row_bytes, S,
                                  cudaMemcpyDefault, s_out));
    } else {
      if (int e = grow_pinned(&sl.h_out, &sl.h_out_cap, dev_out_stride * S)) return e;
      SPXB_CUDA(cudaMemcpyAsync(sl.h_out, sl.d_out, dev_out_stride * S * sizeof(int16_t),
                                cudaMemcpyDeviceToHost, s_out));
      sl.bounce_out = true;
      sl.user_out = out;
      sl.user_out_stride = out_stride_frames * ch;
      sl.dev_out_stride = dev_out_stride;
      if (d.uniform)
        sl.uniform_out = d.max_n_out;
      else
        sl.out_counts.assign(out_frames, out_frames + S);
    }
    b->counters.d2h_bytes += out_row_bytes * S;
  }
  SPXB_CUDA(cudaEventRecord(sl.ev_done, s_out));
  b->counters.calls += 1;
  return 0;
}

}  // namespace spxb

// speex_resampler_reset_mem (resample.c:1208-1220) for every stream. The reference zeroes the FIRST
// nb_channels * (filt_len - 1) floats of `mem`, whose channels are mem_alloc_size apart -- so only
// channel 0's history is certainly cleared; element j of channel c's history goes to zero iff
// c * mem_alloc_size + j < nb_channels * (filt_len - 1). Reproduced as it is (a stereo stream keeps
// its right-channel history across a reset in the reference, and so it does here).
__global__ void reset_hist_kernel(int16_t *hist, uint32_t n_streams, uint32_t hist_stride, uint32_t hist_frames,
                                  uint32_t channels, uint32_t words, uint32_t live, uint32_t mem_alloc,
                                  uint32_t planar_state) {
  const size_t per_stream = static_cast<size_t>(hist_frames) * channels;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= per_stream * n_streams) return;
  const uint32_t s = static_cast<uint32_t>(idx / per_stream), e = static_cast<uint32_t>(idx % per_stream);
  const uint32_t frame = e / channels, c = e % channels, lead = hist_frames - live;
  bool zero = frame < lead;  // padding in front of the live history is always zero
  if (!zero) {
    const uint64_t j = frame - lead;
    // (a planar state: its streams are the channels of one Speex state)
    zero = planar_state ? static_cast<uint64_t>(s) * mem_alloc + j < static_cast<uint64_t>(n_streams) * live
                        : static_cast<uint64_t>(c) * mem_alloc + j < static_cast<uint64_t>(channels) * live;
  }
  if (zero)
    for (uint32_t w = 0; w < words; ++w) hist[static_cast<size_t>(s) * hist_stride + static_cast<size_t>(e) * words + w] = 0;
}

// ---------------------------------------------------------------------------
// C ABI, part 2 (batched streams)
// ---------------------------------------------------------------------------
extern "C" {

int spxb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char *spxb_last_error(void) { return g_last_error.c_str(); }

const char *spxb_version(void) { return "speexb200 0.1 (sm_100a)"; }

static spxb_batch *create_any(uint32_t n_streams, uint32_t channels, uint32_t in_rate, uint32_t out_rate,
                              int quality, int device, bool f32, int *err);

spxb_batch *spxb_batch_create(uint32_t n_streams, uint32_t channels, uint32_t in_rate,
                              uint32_t out_rate, int quality, int device, int *err) {
  return create_any(n_streams, channels, in_rate, out_rate, quality, device, false, err);
}

spxb_batch *spxb_batch_create_f32(uint32_t n_streams, uint32_t channels, uint32_t in_rate,
                                  uint32_t out_rate, int quality, int device, int *err) {
  return create_any(n_streams, channels, in_rate, out_rate, quality, device, true, err);
}

int spxb_batch_is_f32(const spxb_batch *b) { return b && b->f32 ? 1 : 0; }

static spxb_batch *create_any(uint32_t n_streams, uint32_t channels, uint32_t in_rate, uint32_t out_rate,
                              int quality, int device, bool f32, int *err) {
  int e = RESAMPLER_ERR_SUCCESS;
  spxb_batch *b = nullptr;
  FilterSpec spec;
  if (n_streams == 0 || channels == 0) {
    e = RESAMPLER_ERR_INVALID_ARG;
  } else {
    e = derive_filter_spec(in_rate, out_rate, quality, &spec);
  }
  if (e == 0) {
    if (spxb_device_count() <= device || device < 0) {
      set_error("no CUDA device " + std::to_string(device) + " (libspeexb200 has no CPU path)");
      e = RESAMPLER_ERR_ALLOC_FAILED;
    }
  }
  if (e == 0) {
    b = new (std::nothrow) spxb_batch();
    if (!b) {
      e = RESAMPLER_ERR_ALLOC_FAILED;
    } else {
      b->device = device;
      b->spec = spec;
      b->n_streams = n_streams;
      b->channels = channels;
      b->f32 = f32;
      b->hist_words = f32 ? 2 : 1;
      int prev = -1;
      cudaGetDevice(&prev);
      e = create_batch(b);
      if (prev >= 0) cudaSetDevice(prev);
      if (e != 0) {
        free_batch(b);
        b = nullptr;
      }
    }
  }
  if (err) *err = e;
  return b;
}

void spxb_batch_destroy(spxb_batch *b) { free_batch(b); }

int spxb_batch_set_kernel(spxb_batch *b, int kernel) {
  if (!b || kernel < SPXB_KERNEL_AUTO || kernel > SPXB_KERNEL_TENSOR) return RESAMPLER_ERR_INVALID_ARG;
  b->kernel_pref = kernel;
  return 0;
}

int spxb_batch_get_kernel(const spxb_batch *b) { return b ? b->last_kernel : 0; }

int spxb_batch_tensor_geometry(const spxb_batch *b, uint32_t *geom6) {
  if (!b || !geom6) return RESAMPLER_ERR_INVALID_ARG;
  umma_geometry(b->umma, geom6);
  return geom6[0] ? 0 : RESAMPLER_ERR_BAD_STATE;
}

long spxb_batch_tensor_trace(spxb_batch *b, uint64_t *dst, size_t cap_words) {
  if (!b || !dst) return -RESAMPLER_ERR_INVALID_ARG;
  if (spxb_batch_synchronize(b)) return -RESAMPLER_ERR_BAD_STATE;
  DeviceGuard g(b->device);
  return umma_read_trace(b->umma, reinterpret_cast<unsigned long long *>(dst), cap_words);
}

int spxb_batch_pipeline_depth(const spxb_batch *) { return pipeline_slots() - 1; }

int spxb_batch_submit(spxb_batch *b, const int16_t *in, size_t in_stride_frames, uint32_t *in_frames,
                      int16_t *out, size_t out_stride_frames, uint32_t *out_frames,
                      uint64_t *ticket) {
  if (!b || !in_frames || !out_frames || !in || !out) return RESAMPLER_ERR_INVALID_ARG;
  return submit_host(b, in, in_stride_frames, in_frames, out, out_stride_frames, out_frames, ticket);
}

int spxb_batch_wait(spxb_batch *b, uint64_t ticket) {
  if (!b) return RESAMPLER_ERR_INVALID_ARG;
  DeviceGuard g(b->device);
  for (auto &sl : b->slots)
    if (sl.busy && sl.ticket == ticket) return retire_slot(b, sl);
  return 0;  // already retired
}

int spxb_batch_process(spxb_batch *b, const int16_t *in, size_t in_stride_frames, uint32_t *in_frames,
                       int16_t *out, size_t out_stride_frames, uint32_t *out_frames) {
  uint64_t t = 0;
  int e = spxb_batch_submit(b, in, in_stride_frames, in_frames, out, out_stride_frames, out_frames, &t);
  if (e) return e;
  return spxb_batch_wait(b, t);
}

// Float in/out (speex_resampler_process_interleaved_float for every stream): same staging and
// pipeline as the int16 entry with two int16 units per sample; lengths follow the float entry's
// block walk (no 1024-frame output block).
int spxb_batch_process_f32(spxb_batch *b, const float *in, size_t in_stride_frames, uint32_t *in_frames,
                           float *out, size_t out_stride_frames, uint32_t *out_frames) {
  if (!b) return RESAMPLER_ERR_INVALID_ARG;
  if (!b->f32) {
    set_error("spxb_batch_process_f32 needs a batch made by spxb_batch_create_f32 (float history)");
    return RESAMPLER_ERR_BAD_STATE;
  }
  b->io_words = 2;
  uint64_t t = 0;
  int e = submit_host(b, reinterpret_cast<const int16_t *>(in), in_stride_frames, in_frames,
                      reinterpret_cast<int16_t *>(out), out_stride_frames, out_frames, &t);
  b->io_words = 1;
  if (e) return e;
  return spxb_batch_wait(b, t);
}

// Scaled float PCM in and out (+-1.0 full scale) on an int16 batch: samples are converted to int16 as
// the kernel loads them (round to nearest even of x * 32768, saturated) and back as it stores them;
// history, arithmetic and lengths are the int16 path's. Fused into the strict kernel (bit-exact
// against the oracle fed the converted samples); the step in front of / behind the resampler in
// float pipelines.
int spxb_batch_process_pcm_f32(spxb_batch *b, const float *in, size_t in_stride_frames, uint32_t *in_frames,
                               float *out, size_t out_stride_frames, uint32_t *out_frames) {
  if (!b) return RESAMPLER_ERR_INVALID_ARG;
  if (b->f32) {
    set_error("spxb_batch_process_pcm_f32 takes an int16 batch (spxb_batch_create)");
    return RESAMPLER_ERR_BAD_STATE;
  }
  b->io_words = 2;
  b->pcm_float = true;
  uint64_t t = 0;
  int e = submit_host(b, reinterpret_cast<const int16_t *>(in), in_stride_frames, in_frames,
                      reinterpret_cast<int16_t *>(out), out_stride_frames, out_frames, &t);
  b->io_words = 1;
  b->pcm_float = false;
  if (e) return e;
  return spxb_batch_wait(b, t);
}

// Strided call (the per-channel entries of the Speex API, resample.c:925-1036): stream s reads sample
// f of its channel c at in[s * in_stream_stride + f * in_step + c] and writes out[s * out_stream_stride
// + m * out_step + c]; strides and steps count SAMPLES of the call's format (float_io: f32, float
// batch only). Elements between the strided ones are never read or written. Synchronous, strict kernel.
int spxb_batch_process_strided(spxb_batch *b, const void *in, size_t in_stream_stride, uint32_t in_step,
                               uint32_t *in_frames, void *out, size_t out_stream_stride, uint32_t out_step,
                               uint32_t *out_frames, int float_io) {
  if (!b || !in_frames || !out_frames || !in || !out || in_step == 0 || out_step == 0) return RESAMPLER_ERR_INVALID_ARG;
  if (float_io && !b->f32) {
    set_error("float samples need a batch made by spxb_batch_create_f32 (float history)");
    return RESAMPLER_ERR_BAD_STATE;
  }
  if (int e = spxb_batch_synchronize(b)) return e;
  DeviceGuard g(b->device);
  const uint32_t S = b->n_streams, ch = b->channels;
  const uint32_t words = float_io ? 2u : 1u;  // int16 units per sample
  Slot &sl = b->slots[0];
  if (!sl.h_calls) {
    SPXB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_calls), S * sizeof(StreamCall), cudaHostAllocDefault));
    SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&sl.d_calls), S * sizeof(StreamCall)));
    SPXB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_ids), S * sizeof(uint32_t), cudaHostAllocDefault));
    SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&sl.d_ids), S * sizeof(uint32_t)));
  }
  std::vector<uint32_t> offered_in(in_frames, in_frames + S);
  b->io_words = words;
  // every stream gets its own plan (a per-channel call moves one stream of a planar state only)
  const bool was_uniform = b->uniform_pos;
  b->uniform_pos = false;
  Decided d = decide(b, in_frames, out_frames, sl.h_calls);
  b->uniform_pos = was_uniform;
  b->io_words = 1;
  if (!d.any_work) return 0;
  // spans of the caller's buffers the call touches, in samples
  size_t span_in = 0, span_out = 0;
  for (uint32_t s = 0; s < S; ++s) {
    if (offered_in[s]) span_in = std::max(span_in, s * in_stream_stride + static_cast<size_t>(offered_in[s] - 1) * in_step + ch);
    if (out_frames[s]) span_out = std::max(span_out, s * out_stream_stride + static_cast<size_t>(out_frames[s] - 1) * out_step + ch);
  }
  if (int e = grow_device(&sl.d_in, &sl.d_in_cap, std::max<size_t>(span_in * words, 8))) return e;
  if (int e = grow_device(&sl.d_out, &sl.d_out_cap, std::max<size_t>(span_out * words, 8))) return e;
  if (int e = grow_pinned(&sl.h_in, &sl.h_in_cap, std::max<size_t>(span_in * words, 8))) return e;
  if (int e = grow_pinned(&sl.h_out, &sl.h_out_cap, std::max<size_t>(span_out * words, 8))) return e;
  std::memcpy(sl.h_in, in, span_in * words * sizeof(int16_t));
  SPXB_CUDA(cudaMemcpyAsync(sl.d_in, sl.h_in, span_in * words * sizeof(int16_t), cudaMemcpyHostToDevice, b->s_compute));
  SPXB_CUDA(cudaMemcpyAsync(sl.d_calls, sl.h_calls, S * sizeof(StreamCall), cudaMemcpyHostToDevice, b->s_compute));
  b->counters.h2d_bytes += span_in * words * sizeof(int16_t) + S * sizeof(StreamCall);
  b->in_step = in_step;
  b->out_step = out_step;
  b->io_words = words;
  const int e = launch_call(b, sl.d_in, in_stream_stride * words, sl.d_out, out_stream_stride * words, sl.d_calls, d.uni,
                            d.max_n_out);
  b->in_step = b->out_step = 0;
  b->io_words = 1;
  if (e) return e;
  d.uniform = false;
  commit_positions(b, d, sl.h_calls);
  if (span_out) {
    SPXB_CUDA(cudaMemcpyAsync(sl.h_out, sl.d_out, span_out * words * sizeof(int16_t), cudaMemcpyDeviceToHost, b->s_compute));
    b->counters.d2h_bytes += span_out * words * sizeof(int16_t);
  }
  SPXB_CUDA(cudaStreamSynchronize(b->s_compute));
  // only the samples the call produced go to the caller's buffer
  for (uint32_t s = 0; s < S; ++s)
    for (uint32_t m = 0; m < out_frames[s]; ++m) {
      const size_t at = (s * out_stream_stride + static_cast<size_t>(m) * out_step) * words;
      std::memcpy(static_cast<int16_t *>(out) + at, sl.h_out + at, static_cast<size_t>(ch) * words * sizeof(int16_t));
    }
  b->counters.calls += 1;
  return 0;
}

int spxb_batch_process_device(spxb_batch *b, const int16_t *d_in, size_t in_stride_frames,
                              uint32_t *in_frames, int16_t *d_out, size_t out_stride_frames,
                              uint32_t *out_frames) {
  if (!b || !in_frames || !out_frames) return RESAMPLER_ERR_INVALID_ARG;
  DeviceGuard g(b->device);
  Slot &sl = b->slots[b->next_ticket % pipeline_slots()];
  if (int e = retire_slot(b, sl)) return e;
  const uint32_t S = b->n_streams;
  if (!sl.h_calls) {
    SPXB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_calls), S * sizeof(StreamCall), cudaHostAllocDefault));
    SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&sl.d_calls), S * sizeof(StreamCall)));
    SPXB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_ids), S * sizeof(uint32_t), cudaHostAllocDefault));
    SPXB_CUDA(cudaMalloc(reinterpret_cast<void **>(&sl.d_ids), S * sizeof(uint32_t)));
  }
  Decided d = decide(b, in_frames, out_frames, sl.h_calls);
  if (!d.any_work) return 0;
  sl.ticket = b->next_ticket++;
  if (!d.uniform) {
    // the plan upload rides the compute stream so it is ordered before the kernel
    SPXB_CUDA(cudaMemcpyAsync(sl.d_calls, sl.h_calls, S * sizeof(StreamCall), cudaMemcpyHostToDevice, b->s_compute));
    b->counters.h2d_bytes += S * sizeof(StreamCall);
    sl.busy = true;  // h_calls must outlive the async copy
  }
  int e = launch_call(b, d_in, in_stride_frames * b->channels * b->io_words, d_out,
                      out_stride_frames * b->channels * b->io_words,
                      d.uniform ? nullptr : sl.d_calls, d.uni, d.max_n_out,
                      d.uniform ? RaggedHost() : RaggedHost{sl.h_calls, sl.h_ids, sl.d_ids});
  if (e) {
    if (!d.uniform) cudaEventRecord(sl.ev_done, b->s_compute);  // the slot's plan upload is in flight
    return e;
  }
  commit_positions(b, d, sl.h_calls);
  if (!d.uniform) SPXB_CUDA(cudaEventRecord(sl.ev_done, b->s_compute));
  b->counters.calls += 1;
  return 0;
}

int spxb_batch_process_device_uniform(spxb_batch *b, const int16_t *d_in, size_t in_stride_frames,
                                      uint32_t n_in, int16_t *d_out, size_t out_stride_frames,
                                      uint32_t out_cap, uint32_t *in_used, uint32_t *out_written) {
  if (!b) return RESAMPLER_ERR_INVALID_ARG;
  if (!b->uniform_pos) {
    set_error("process_device_uniform needs all streams at one position; use process_device");
    return RESAMPLER_ERR_BAD_STATE;
  }
  DeviceGuard g(b->device);
  Decided d = decide_uniform(b, n_in, out_cap);
  if (in_used) *in_used = d.uni.consumed;
  if (out_written) *out_written = d.uni.n_out;
  if (!d.any_work) return 0;
  int e = launch_call(b, d_in, in_stride_frames * b->channels * b->io_words, d_out,
                      out_stride_frames * b->channels * b->io_words,
                      nullptr, d.uni, d.max_n_out);
  if (e) return e;
  commit_positions(b, d, nullptr);
  b->counters.calls += 1;
  return 0;
}

// Hop sequences are launch-bound at small shapes (C3: 7.5 us per 20 ms hop, a third of it launch
// latency), so a sequence that repeats -- same buffers, same ring phase, same stream position --
// is captured into a CUDA graph the second time it is seen unchanged and replayed from then on.
// The host-side planning (lengths, positions, ping-pong) still runs per hop, in dry-run mode, so
// the batch's state after a replay is by construction what the launches would have left.
static bool ring_graphs_enabled() {
  static const bool v = [] {
    const char *e = getenv("SPXB_RING_GRAPH");
    return !e || atoi(e) != 0;
  }();
  return v;
}

static int ring_hops(spxb_batch *b, const int16_t *d_in, size_t in_stride_frames, size_t in_slot_elems,
                     int16_t *d_out, size_t out_stride_frames, size_t out_slot_elems, uint32_t ring,
                     uint32_t n_in, uint32_t out_cap, uint32_t first_step, uint32_t steps) {
  int e = 0;
  b->defer_pos_mirror = true;
  for (uint32_t k = first_step; k < first_step + steps && !e; ++k) {
    const size_t slot = k % ring;
    e = spxb_batch_process_device_uniform(b, d_in + slot * in_slot_elems, in_stride_frames, n_in,
                                          d_out + slot * out_slot_elems, out_stride_frames, out_cap, nullptr,
                                          nullptr);
  }
  b->defer_pos_mirror = false;
  for (auto &q : b->pos) q = b->pos[0];
  return e;
}

int spxb_batch_process_device_ring(spxb_batch *b, const int16_t *d_in, size_t in_stride_frames,
                                   size_t in_slot_elems, int16_t *d_out, size_t out_stride_frames,
                                   size_t out_slot_elems, uint32_t ring, uint32_t n_in, uint32_t out_cap,
                                   uint32_t first_step, uint32_t steps) {
  if (!b || ring == 0) return RESAMPLER_ERR_INVALID_ARG;
  auto plain = [&] {
    return ring_hops(b, d_in, in_stride_frames, in_slot_elems, d_out, out_stride_frames, out_slot_elems, ring,
                     n_in, out_cap, first_step, steps);
  };
  // (the legacy default stream cannot be captured)
  if (!ring_graphs_enabled() || steps < 4 || !b->uniform_pos || b->s_compute == nullptr) return plain();

  // everything the captured launches depend on
  struct Key {
    const void *in, *out;
    size_t in_stride, in_slot, out_stride, out_slot;
    uint32_t ring, n_in, out_cap, phase, steps;
    int32_t ls;
    uint32_t frac;
    int hist_cur, kernel_pref;
    uint64_t pool_generation;
    void *stream;
  } key;
  std::memset(&key, 0, sizeof(key));
  key.in = d_in;
  key.out = d_out;
  key.in_stride = in_stride_frames;
  key.in_slot = in_slot_elems;
  key.out_stride = out_stride_frames;
  key.out_slot = out_slot_elems;
  key.ring = ring;
  key.n_in = n_in;
  key.out_cap = out_cap;
  key.phase = first_step % ring;
  key.steps = steps;
  key.ls = b->pos[0].last_sample;
  key.frac = b->pos[0].samp_frac_num;
  key.hist_cur = b->hist_cur;
  key.kernel_pref = b->kernel_pref;
  key.pool_generation = umma_pool_generation(b->umma);
  key.stream = b->s_compute;
  const std::string k(reinterpret_cast<const char *>(&key), sizeof(key));
  DeviceGuard g(b->device);

  auto hit = b->ring_graphs.find(k);
  if (hit != b->ring_graphs.end()) {
    // the GPU starts at once; the host-side bookkeeping of the hops runs beside it
    hit->second.last_use = ++b->ring_clock;
    if (cudaGraphLaunch(hit->second.exec, b->s_compute) != cudaSuccess) {
      // nothing ran and nothing moved yet: drop the graph and run the hops launch by launch
      cudaGetLastError();
      cudaGraphExecDestroy(hit->second.exec);
      b->ring_graphs.erase(hit);
      return plain();
    }
    b->dry_run = true;
    const int e = plain();
    b->dry_run = false;
    return e;
  }

  auto seen = b->ring_seen.find(k);
  const bool capture = seen != b->ring_seen.end() && seen->second == 0;
  if (!capture) {
    // launch by launch, remembering whether the planner had to touch the stream on the way
    const uint64_t ops0 = umma_stream_ops(b->umma);
    const int e = plain();
    if (b->ring_seen.size() > 4096) b->ring_seen.clear();
    b->ring_seen[k] = umma_stream_ops(b->umma) - ops0;
    return e;
  }

  // second unchanged sighting: capture the launches (they still execute: the graph is launched below).
  // Capturing runs the host-side planning of every hop, which moves the position shadow, the
  // ping-pong half and the counters although no kernel executes yet: if the capture or the
  // instantiation fails, all of that is put back and the sequence runs launch by launch instead.
  const StreamPos pos_before = b->pos[0];
  const int hist_before = b->hist_cur;
  const spxb_counters counters_before = b->counters;
  auto fall_back = [&](const char *what, cudaError_t ce) {
    for (auto &q : b->pos) q = pos_before;
    b->hist_cur = hist_before;
    b->counters = counters_before;
    b->memo_valid = false;
    b->ring_seen.erase(k);  // do not try to capture this sequence again right away
    cudaGetLastError();
    set_error(std::string("ring graph ") + what + ": " + cudaGetErrorString(ce) + " (ran launch by launch)");
    return plain();
  };
  cudaGraph_t graph = nullptr;
  SPXB_CUDA(cudaStreamBeginCapture(b->s_compute, cudaStreamCaptureModeThreadLocal));
  umma_set_frozen(b->umma, true);
  const int e = plain();
  umma_set_frozen(b->umma, false);
  const cudaError_t ce = cudaStreamEndCapture(b->s_compute, &graph);
  if (e || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    return fall_back("capture", ce != cudaSuccess ? ce : cudaErrorUnknown);
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ci = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ci != cudaSuccess || !exec) return fall_back("instantiate", ci != cudaSuccess ? ci : cudaErrorUnknown);
  if (b->ring_graphs.size() >= 64) {  // evict the least recently used
    auto victim = b->ring_graphs.begin();
    for (auto it = b->ring_graphs.begin(); it != b->ring_graphs.end(); ++it)
      if (it->second.last_use < victim->second.last_use) victim = it;
    cudaGraphExecDestroy(victim->second.exec);
    b->ring_graphs.erase(victim);
  }
  spxb_batch::RingGraph rg;
  rg.exec = exec;
  rg.last_use = ++b->ring_clock;
  b->ring_graphs.emplace(k, rg);
  const cudaError_t cl = cudaGraphLaunch(exec, b->s_compute);
  if (cl != cudaSuccess) {
    cudaGraphExecDestroy(exec);
    b->ring_graphs.erase(k);
    return fall_back("launch", cl);
  }
  return 0;
}

int spxb_batch_set_stream(spxb_batch *b, void *cuda_stream) {
  if (!b) return RESAMPLER_ERR_INVALID_ARG;
  DeviceGuard g(b->device);
  SPXB_CUDA(cudaStreamSynchronize(b->s_compute));
  b->s_compute = static_cast<cudaStream_t>(cuda_stream);
  b->external_stream = true;
  return 0;
}

int spxb_batch_use_own_stream(spxb_batch *b) {
  if (!b) return RESAMPLER_ERR_INVALID_ARG;
  DeviceGuard g(b->device);
  SPXB_CUDA(cudaStreamSynchronize(b->s_compute));
  b->s_compute = b->s_own;
  b->external_stream = false;
  return 0;
}

int spxb_batch_synchronize(spxb_batch *b) {
  if (!b) return RESAMPLER_ERR_INVALID_ARG;
  DeviceGuard g(b->device);
  for (auto &sl : b->slots)
    if (int e = retire_slot(b, sl)) return e;
  SPXB_CUDA(cudaStreamSynchronize(b->s_in));
  SPXB_CUDA(cudaStreamSynchronize(b->s_compute));
  SPXB_CUDA(cudaStreamSynchronize(b->s_out));
  return 0;
}

int spxb_batch_get_state(spxb_batch *b, uint32_t stream, int32_t *last_sample, uint32_t *samp_frac_num,
                         uint32_t *magic_samples, int16_t *history) {
  if (!b || stream >= b->n_streams) return RESAMPLER_ERR_INVALID_ARG;
  if (b->f32 && history) {
    set_error("this batch keeps a float history: use spxb_batch_get_state_f32");
    return RESAMPLER_ERR_INVALID_ARG;
  }
  if (int e = spxb_batch_synchronize(b)) return e;
  DeviceGuard g(b->device);
  if (last_sample)
    SPXB_CUDA(cudaMemcpy(last_sample, b->d_last_sample + stream, sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (samp_frac_num)
    SPXB_CUDA(cudaMemcpy(samp_frac_num, b->d_samp_frac + stream, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (magic_samples)
    SPXB_CUDA(cudaMemcpy(magic_samples, b->d_magic + stream, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (history) {
    const size_t live = static_cast<size_t>(b->spec.taps - 1) * b->channels;
    const size_t lead = static_cast<size_t>(b->hist_frames - (b->spec.taps - 1)) * b->channels;
    if (live)
      SPXB_CUDA(cudaMemcpy(history, b->d_hist[b->hist_cur] + static_cast<size_t>(stream) * b->hist_stride + lead,
                           live * sizeof(int16_t), cudaMemcpyDeviceToHost));
  }
  return 0;
}

int spxb_batch_set_state(spxb_batch *b, uint32_t stream, int32_t last_sample, uint32_t samp_frac_num,
                         const int16_t *history) {
  if (!b || stream >= b->n_streams || last_sample < 0 || samp_frac_num >= b->spec.den)
    return RESAMPLER_ERR_INVALID_ARG;
  if (b->f32 && history) {
    set_error("this batch keeps a float history: use spxb_batch_set_state_f32");
    return RESAMPLER_ERR_INVALID_ARG;
  }
  if (int e = spxb_batch_synchronize(b)) return e;
  DeviceGuard g(b->device);
  SPXB_CUDA(cudaMemcpy(b->d_last_sample + stream, &last_sample, sizeof(int32_t), cudaMemcpyHostToDevice));
  SPXB_CUDA(cudaMemcpy(b->d_samp_frac + stream, &samp_frac_num, sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (history) {
    const size_t live = static_cast<size_t>(b->spec.taps - 1) * b->channels;
    const size_t lead = static_cast<size_t>(b->hist_frames - (b->spec.taps - 1)) * b->channels;
    if (live)
      SPXB_CUDA(cudaMemcpy(b->d_hist[b->hist_cur] + static_cast<size_t>(stream) * b->hist_stride + lead, history,
                           live * sizeof(int16_t), cudaMemcpyHostToDevice));
  }
  b->pos[stream].last_sample = last_sample;
  b->pos[stream].samp_frac_num = samp_frac_num;
  b->uniform_pos = true;
  for (uint32_t s = 1; s < b->n_streams && b->uniform_pos; ++s)
    b->uniform_pos = b->pos[s].last_sample == b->pos[0].last_sample &&
                     b->pos[s].samp_frac_num == b->pos[0].samp_frac_num;
  return 0;
}

// Float view of a stream's state, for both kinds of batch (an int16 history converts exactly).
int spxb_batch_get_state_f32(spxb_batch *b, uint32_t stream, int32_t *last_sample, uint32_t *samp_frac_num,
                             uint32_t *magic_samples, float *history) {
  if (!b || stream >= b->n_streams) return RESAMPLER_ERR_INVALID_ARG;
  const size_t live = static_cast<size_t>(b->spec.taps - 1) * b->channels;
  if (!b->f32) {
    std::vector<int16_t> h(live ? live : 1);
    if (int e = spxb_batch_get_state(b, stream, last_sample, samp_frac_num, magic_samples, history ? h.data() : nullptr))
      return e;
    if (history)
      for (size_t i = 0; i < live; ++i) history[i] = static_cast<float>(h[i]);
    return 0;
  }
  if (int e = spxb_batch_get_state(b, stream, last_sample, samp_frac_num, magic_samples, nullptr)) return e;
  if (history && live) {
    DeviceGuard g(b->device);
    const size_t lead = static_cast<size_t>(b->hist_frames - (b->spec.taps - 1)) * b->channels;
    const float *row = reinterpret_cast<const float *>(b->d_hist[b->hist_cur] + static_cast<size_t>(stream) * b->hist_stride);
    SPXB_CUDA(cudaMemcpy(history, row + lead, live * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return 0;
}

int spxb_batch_set_state_f32(spxb_batch *b, uint32_t stream, int32_t last_sample, uint32_t samp_frac_num,
                             const float *history) {
  if (!b || stream >= b->n_streams) return RESAMPLER_ERR_INVALID_ARG;
  const size_t live = static_cast<size_t>(b->spec.taps - 1) * b->channels;
  if (!b->f32) {
    // an int16 history can only hold what int16 input left there
    std::vector<int16_t> h(live ? live : 1);
    if (history)
      for (size_t i = 0; i < live; ++i) {
        const float v = history[i];
        if (!(v >= -32768.f && v <= 32767.f) || v != static_cast<float>(static_cast<int16_t>(v))) {
          set_error("set_state_f32: history is not int16-valued; use a float batch (spxb_batch_create_f32)");
          return RESAMPLER_ERR_INVALID_ARG;
        }
        h[i] = static_cast<int16_t>(v);
      }
    return spxb_batch_set_state(b, stream, last_sample, samp_frac_num, history ? h.data() : nullptr);
  }
  if (int e = spxb_batch_set_state(b, stream, last_sample, samp_frac_num, nullptr)) return e;
  if (history && live) {
    DeviceGuard g(b->device);
    const size_t lead = static_cast<size_t>(b->hist_frames - (b->spec.taps - 1)) * b->channels;
    float *row = reinterpret_cast<float *>(b->d_hist[b->hist_cur] + static_cast<size_t>(stream) * b->hist_stride);
    SPXB_CUDA(cudaMemcpy(row + lead, history, live * sizeof(float), cudaMemcpyHostToDevice));
  }
  return 0;
}

int spxb_batch_reset(spxb_batch *b) {
  if (!b) return RESAMPLER_ERR_INVALID_ARG;
  if (int e = spxb_batch_synchronize(b)) return e;
  DeviceGuard g(b->device);
  const uint32_t live = b->spec.taps - 1, mem_alloc = live + b->in_block;  // resample.c:709, :835
  const size_t elems = static_cast<size_t>(b->n_streams) * b->hist_frames * b->channels;
  if (elems) {
    reset_hist_kernel<<<static_cast<unsigned>((elems + 255) / 256), 256, 0, b->s_compute>>>(
        b->d_hist[b->hist_cur], b->n_streams, b->hist_stride, b->hist_frames, b->channels, b->hist_words, live,
        mem_alloc, b->planar_state ? 1u : 0u);
    SPXB_CUDA(cudaGetLastError());
  }
  SPXB_CUDA(cudaMemsetAsync(b->d_last_sample, 0, b->n_streams * sizeof(int32_t), b->s_compute));
  SPXB_CUDA(cudaMemsetAsync(b->d_samp_frac, 0, b->n_streams * sizeof(uint32_t), b->s_compute));
  SPXB_CUDA(cudaMemsetAsync(b->d_magic, 0, b->n_streams * sizeof(uint32_t), b->s_compute));
  SPXB_CUDA(cudaStreamSynchronize(b->s_compute));
  b->pos.assign(b->n_streams, StreamPos{});
  b->uniform_pos = true;
  b->memo_valid = false;
  return 0;
}

int spxb_batch_skip_zeros(spxb_batch *b) {
  if (!b) return RESAMPLER_ERR_INVALID_ARG;
  if (int e = spxb_batch_synchronize(b)) return e;
  DeviceGuard g(b->device);
  std::vector<int32_t> ls(b->n_streams, static_cast<int32_t>(b->spec.taps / 2));
  SPXB_CUDA(cudaMemcpy(b->d_last_sample, ls.data(), ls.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  for (auto &p : b->pos) p.last_sample = static_cast<int32_t>(b->spec.taps / 2);
  return 0;
}

int spxb_batch_counters(const spxb_batch *b, spxb_counters *c) {
  if (!b || !c) return RESAMPLER_ERR_INVALID_ARG;
  *c = b->counters;
  return 0;
}

void *spxb_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    set_error("cudaHostAlloc failed");
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void spxb_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------
// C ABI, part 3 (host-only introspection)
// ---------------------------------------------------------------------------
int spxb_filter_describe(uint32_t in_rate, uint32_t out_rate, int quality, spxb_filter_info *info) {
  if (!info) return RESAMPLER_ERR_INVALID_ARG;
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, quality, &s)) return e;
  info->num = s.num;
  info->den = s.den;
  info->filt_len = s.taps;
  info->oversample = s.oversample;
  info->int_advance = s.int_advance;
  info->frac_advance = s.frac_advance;
  info->cutoff = s.cutoff;
  info->use_direct = s.direct;
  info->use_double = s.wide_accum;
  info->table_len = s.table_len;
  return 0;
}

long spxb_filter_table(uint32_t in_rate, uint32_t out_rate, int quality, float *dst, size_t cap) {
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, quality, &s)) return -e;
  if (!dst || cap < s.table_len) return -RESAMPLER_ERR_INVALID_ARG;
  std::vector<float> t = build_reference_table(s);
  std::memcpy(dst, t.data(), t.size() * sizeof(float));
  return static_cast<long>(t.size());
}

long spxb_filter_phase_taps(uint32_t in_rate, uint32_t out_rate, int quality, float *dst, size_t cap) {
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, quality, &s)) return -e;
  const size_t n = static_cast<size_t>(s.den) * s.taps;
  if (!dst || cap < n) return -RESAMPLER_ERR_INVALID_ARG;
  std::vector<float> t = build_phase_taps(s, build_reference_table(s));
  std::memcpy(dst, t.data(), n * sizeof(float));
  return static_cast<long>(n);
}

long spxb_filter_fixed_taps(uint32_t in_rate, uint32_t out_rate, int quality, int32_t *dst, size_t cap,
                            int *shift) {
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, quality, &s)) return -e;
  const size_t n = static_cast<size_t>(s.den) * s.taps;
  if (!dst || !shift || cap < n) return -RESAMPLER_ERR_INVALID_ARG;
  FixedTaps ft;
  if (!build_fixed_taps(s, build_reference_table(s), &ft)) return -RESAMPLER_ERR_BAD_STATE;
  std::memcpy(dst, ft.h.data(), n * sizeof(int32_t));
  *shift = ft.shift;
  return static_cast<long>(n);
}

long spxb_tensor_plan(uint32_t in_rate, uint32_t out_rate, int quality, int32_t last_sample,
                      uint32_t samp_frac_num, uint32_t n_out, uint32_t nt, int32_t *dst, size_t cap_tiles,
                      uint32_t *ksteps) {
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, quality, &s)) return -e;
  if (!dst || !ksteps || nt < 16 || nt > 128 || nt % 16 != 0 || last_sample < 0 || samp_frac_num >= s.den)
    return -RESAMPLER_ERR_INVALID_ARG;
  std::vector<UmmaTile> tiles;
  std::vector<UmmaTileKey> keys;
  const uint32_t hist_frames = static_cast<uint32_t>(round_up(s.taps - 1, 16));
  plan_umma_tiles(s.num, s.den, s.taps, hist_frames, last_sample, samp_frac_num, n_out, nt, &tiles, &keys);
  if (tiles.size() > cap_tiles) return -RESAMPLER_ERR_INVALID_ARG;
  for (size_t i = 0; i < tiles.size(); ++i) {
    dst[4 * i + 0] = static_cast<int32_t>(tiles[i].m0);
    dst[4 * i + 1] = tiles[i].kf0;
    dst[4 * i + 2] = static_cast<int32_t>(keys[i].phase0);
    dst[4 * i + 3] = static_cast<int32_t>(keys[i].delta);
  }
  *ksteps = umma_ksteps(s.taps, s.num, s.den, nt);
  return static_cast<long>(tiles.size());
}

long spxb_tensor_tap_tile(uint32_t in_rate, uint32_t out_rate, int quality, uint32_t nt, uint32_t phase0,
                          uint32_t delta, int8_t *dst, size_t cap) {
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, quality, &s)) return -e;
  if (!dst || nt < 16 || nt > 128 || nt % 16 != 0 || phase0 >= s.den || delta >= 16)
    return -RESAMPLER_ERR_INVALID_ARG;
  const uint32_t ks = umma_ksteps(s.taps, s.num, s.den, nt);
  const size_t bytes = static_cast<size_t>(2) * ks * 3 * nt * 16;
  if (cap < bytes) return -RESAMPLER_ERR_INVALID_ARG;
  FixedTaps ft;
  if (!build_fixed_taps(s, build_reference_table(s), &ft)) return -RESAMPLER_ERR_BAD_STATE;
  UmmaTileKey key;
  key.phase0 = phase0;
  key.delta = delta;
  fill_tap_tile_host(ft, s.num, s.den, s.taps, nt, ks, key, dst);
  return static_cast<long>(bytes);
}

long spxb_tensor_packed_plan(uint32_t in_rate, uint32_t out_rate, int quality, uint32_t nt, uint32_t *dst,
                             size_t cap_words, uint32_t *tile_bytes) {
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, quality, &s)) return -e;
  if (!dst || !tile_bytes) return -RESAMPLER_ERR_INVALID_ARG;
  FixedTaps ft;
  if (!build_fixed_taps(s, build_reference_table(s), &ft)) return -RESAMPLER_ERR_BAD_STATE;
  UmmaPackedPlan plan;
  if (!build_packed_plan(s, ft, nt, &plan)) return -RESAMPLER_ERR_INVALID_ARG;
  if (cap_words < static_cast<size_t>(plan.ksteps) * 18) return -RESAMPLER_ERR_INVALID_ARG;
  for (uint32_t k = 0; k < plan.ksteps; ++k) {
    const UmmaKStep &ks = plan.k[k];
    uint32_t *w = dst + static_cast<size_t>(k) * 18;
    w[0] = ks.off16;
    w[1] = ks.rows;
    w[2] = ks.n_ent;
    for (uint32_t e = 0; e < 3; ++e) {
      w[3 + 3 * e] = e < ks.n_ent ? ks.ent[e].row : 0;
      w[4 + 3 * e] = e < ks.n_ent ? ks.ent[e].n : 0;
      w[5 + 3 * e] = e < ks.n_ent ? ks.ent[e].dcol : 0;
    }
    for (int i = 0; i < 3; ++i) {
      w[12 + i] = ks.b0[i];
      w[15 + i] = ks.b1[i];
    }
  }
  *tile_bytes = plan.tile_bytes;
  return static_cast<long>(plan.ksteps);
}

long spxb_tensor_tap_tile_packed(uint32_t in_rate, uint32_t out_rate, int quality, uint32_t nt, uint32_t phase0,
                                 uint32_t delta, int8_t *dst, size_t cap) {
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, quality, &s)) return -e;
  if (!dst || phase0 >= s.den || delta >= 16) return -RESAMPLER_ERR_INVALID_ARG;
  FixedTaps ft;
  if (!build_fixed_taps(s, build_reference_table(s), &ft)) return -RESAMPLER_ERR_BAD_STATE;
  UmmaPackedPlan plan;
  if (!build_packed_plan(s, ft, nt, &plan)) return -RESAMPLER_ERR_INVALID_ARG;
  if (cap < plan.tile_bytes) return -RESAMPLER_ERR_INVALID_ARG;
  UmmaTileKey key;
  key.phase0 = phase0;
  key.delta = delta;
  fill_tap_tile_packed_host(ft, s.num, s.den, s.taps, plan, key, dst);
  return static_cast<long>(plan.tile_bytes);
}

static int plan_call_any(uint32_t in_rate, uint32_t out_rate, int32_t last_sample, uint32_t samp_frac_num,
                         uint32_t n_in, uint32_t out_cap, uint32_t out_block, spxb_call_plan *plan);

int spxb_plan_call(uint32_t in_rate, uint32_t out_rate, int32_t last_sample, uint32_t samp_frac_num,
                   uint32_t n_in, uint32_t out_cap, spxb_call_plan *plan) {
  return plan_call_any(in_rate, out_rate, last_sample, samp_frac_num, n_in, out_cap, kOutBlock, plan);
}

int spxb_plan_call_ex(uint32_t in_rate, uint32_t out_rate, int32_t last_sample, uint32_t samp_frac_num,
                      uint32_t magic, uint32_t n_in, uint32_t out_cap, int float_entry, uint32_t in_block,
                      spxb_call_plan *plan, uint32_t *magic_used) {
  if (!plan || in_rate == 0 || out_rate == 0 || last_sample < 0 || in_block == 0) return RESAMPLER_ERR_INVALID_ARG;
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, 0, &s)) return e;
  if (samp_frac_num >= s.den) return RESAMPLER_ERR_INVALID_ARG;
  StreamPos p;
  p.last_sample = last_sample;
  p.samp_frac_num = samp_frac_num;
  const MagicPlan mp = plan_call_magic(s.num, s.den, p, magic, n_in, out_cap, float_entry != 0, in_block);
  plan->n_out = mp.plan.n_out;
  plan->consumed = mp.plan.consumed;
  plan->last_sample = mp.plan.next.last_sample;
  plan->samp_frac_num = mp.plan.next.samp_frac_num;
  if (magic_used) *magic_used = mp.magic_used;
  return 0;
}

int spxb_plan_call_f32(uint32_t in_rate, uint32_t out_rate, int32_t last_sample, uint32_t samp_frac_num,
                       uint32_t n_in, uint32_t out_cap, spxb_call_plan *plan) {
  return plan_call_any(in_rate, out_rate, last_sample, samp_frac_num, n_in, out_cap, kOutBlockUnbounded, plan);
}

static int plan_call_any(uint32_t in_rate, uint32_t out_rate, int32_t last_sample, uint32_t samp_frac_num,
                         uint32_t n_in, uint32_t out_cap, uint32_t out_block, spxb_call_plan *plan) {
  if (!plan || in_rate == 0 || out_rate == 0 || last_sample < 0) return RESAMPLER_ERR_INVALID_ARG;
  FilterSpec s;
  if (int e = derive_filter_spec(in_rate, out_rate, 0, &s)) return e;
  if (samp_frac_num >= s.den) return RESAMPLER_ERR_INVALID_ARG;
  StreamPos p;
  p.last_sample = last_sample;
  p.samp_frac_num = samp_frac_num;
  const CallPlan pl = plan_call(s.num, s.den, p, n_in, out_cap, out_block);
  plan->n_out = pl.n_out;
  plan->consumed = pl.consumed;
  plan->last_sample = pl.next.last_sample;
  plan->samp_frac_num = pl.next.samp_frac_num;
  return 0;
}

}  // extern "C"

// exposed to capi.cpp
namespace spxb {
const FilterSpec &batch_spec(const spxb_batch *b) { return b->spec; }
int batch_kernel_pref(const spxb_batch *b) { return b->kernel_pref; }
void batch_set_in_block(spxb_batch *b, uint32_t in_block) {
  b->in_block = in_block ? in_block : kInBlock;
  b->memo_valid = false;
}
void batch_force_plan(spxb_batch *b, const CallPlan *plan) { b->forced_plan = plan; }
void batch_set_planar_state(spxb_batch *b, bool planar) { b->planar_state = planar; }
uint32_t batch_in_block(const spxb_batch *b) { return b->in_block; }
}
