// kernels_umma.cu -- tensor-core FIR for sm_100a (tcgen05.mma.kind::i8, accumulators in TMEM);
// the "tensor" kernel family of include/speexb200.h.
//
// One launch does the whole hot path of speex_resampler_process_interleaved_int
// (deps/speex/resample.c:1061-1082 over :968-1036 and the four resampler_basic_* kernels
// :331-558) for a batch of streams at a common stream position, as an EXACT integer banded
// GEMM (umma_plan.h): int16 samples split into (hi s8, lo u8) byte planes, per-phase taps
// quantised to 24-bit fixed point and split into three s8 digits, int32 accumulation on the
// tensor cores, one rounding at the end (WORD2INT, arch.h:208-209).
//
// CTA = one output tile: 128 series (64 stereo / 128 mono streams) x nt consecutive outputs.
//   D[128 x 4nt] (TMEM, s32) column blocks P0..P3 with weights 2^24, 2^16, 2^8, 1:
//     A = hi plane (s8) x B rows [d2 | d1 | d0]  -> columns [0, 3nt)
//     A = lo plane (u8) x the same B             -> columns [nt, 4nt)
//   The K (window) axis is streamed through a ring of shared-memory stages of 64 frames:
//     X stage   [plane 2][K step 2][half 2][row 128][16 B]  K-major, no swizzle; row = stream
//               (mono) or 2*stream + channel (stereo); chunk strides padded so that a converter
//               store instruction hits 32 distinct banks (x_lbo / x_kstep below)
//     tap stage [chunk 4][row 3nt][16 B]                   one 1-D bulk copy from the tile pool
// Warp roles (352 threads):
//   0-7  converters: fetch int16 PCM (history for frames < 0, the call's input after) with lanes
//        walking ALONG a stream's segment (one LDG.128 = four whole lines), two stages of loads in
//        flight in registers, split into byte planes with PRMT, store in UMMA layout; afterwards
//        the epilogue: tcgen05.ld, 32-bit nested-floor recombination (exact), lane-pair exchange for
//        stereo, cvt.pack.sat (= WORD2INT's saturation) and 16-byte stores straight to the output;
//   8    lane 0 owns the mbarriers and issues the tap bulk copies (first ring-full before the
//        programmatic grid dependency resolves);
//   9    owns TMEM; the whole warp walks the stages, one elected lane issues the MMAs from
//        uniform registers (four per K step) and releases stages with tcgen05.commit;
//   10   slides the history (resample.c:898-899) and publishes the new stream position, beside
//        the FIR, four streams at a time from the front of the CTA's share; for long histories
//        the converter warps take single streams from its back once their last stage is stored
//        (they would otherwise idle until the accumulator is complete).
// Instantiations: CH (mono / stereo) x FAST (all rows 16-byte aligned, no cluster) x IDS (the launch
// covers a stream-id list: one cohort of a ragged batch) x PACED (long K loops; see the kernel).
// Stages are handed over with mbarriers (full: 256 converter arrivals + the bulk copy's byte
// count; empty: tcgen05.commit). Launched with programmatic stream serialization: the next call's
// grid runs its prologue while this one drains and blocks in griddepcontrol.wait before touching
// PCM or history. Measured limits and what was tried: DESIGN.md section 4.3.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "kernels_common.cuh"
#include "launch.h"
#include "umma_plan.h"
#include "umma_common.cuh"
#include "umma_context.h"
#include "umma_ptx.cuh"

namespace spxb {

namespace {

using namespace ptx;
using namespace ummac;

constexpr int kMaxStages = 6;

constexpr uint32_t kInlineTiles = 64;  // tile table carried in the kernel parameters up to this many tiles
struct InlineTile {
  int32_t kf0;    // first frame of the tile's K axis (UmmaTile::kf0)
  uint32_t slot;  // its tap tile in the pool
};

struct UmmaArgs {
  const UmmaTile *tiles;
  uint32_t n_tiles;
  uint32_t n_inline;          // != 0: inl[] holds the tile table (no dependent global load, no
                              // 64-bit division in the prologue)
  uint32_t n_groups;          // series groups in the grid (padded to a multiple of `cluster`)
  uint32_t cluster;           // CTAs per cluster (1, 2 or 4): same tile, consecutive series groups;
                              // each loads 1/cluster of every tap stage and multicasts it
  const int8_t *pool;
  uint32_t tile_bytes;
  uint32_t nt;
  uint32_t ksteps;
  uint32_t stages;
  uint32_t tmem_cols;
  int shift;
  unsigned long long *trace;  // optional per-CTA timeline (kTraceSlots words per CTA), else nullptr
  uint32_t debug;             // SPXB_UMMA_DEBUG bits (timing experiments only, results are garbage):
                              // 1 = no PCM loads, 2 = no MMAs, 4 = no PCM conversion/stores
  InlineTile inl[kInlineTiles];
};

// Timing-experiment knobs cost ~10 % on the long-filter shapes even when off (they perturb the
// converters' load scheduling), so they exist only in builds made with -DSPXB_DEBUG_KNOBS.
#ifdef SPXB_DEBUG_KNOBS
#define SPXB_DEBUG_BITS(u) ((u).debug)
#else
#define SPXB_DEBUG_BITS(u) 0u
#endif
constexpr int kTraceSlots = 32;
// The timeline marks exist only in the PACED instantiations of the kernel (see umma_fir_kernel):
// inside it SPXB_TRACE_PTR(u) is u.trace, elsewhere a compile-time null and the marks vanish.
#define SPXB_TRACE_PTR(u) (PACED ? (u).trace : static_cast<unsigned long long *>(nullptr))
template <bool PACED>
__device__ __forceinline__ void trace_mark_t(const UmmaArgs &u, int slot) {
  if (SPXB_TRACE_PTR(u))
    SPXB_TRACE_PTR(u)[static_cast<size_t>(blockIdx.x) * kTraceSlots + slot] = static_cast<unsigned long long>(clock64());
}
#define trace_mark(u, slot) trace_mark_t<PACED>(u, slot)

// History slide of the streams first, first + step, ... (n_total of them) by whichever warps are
// free, claiming from one packed counter (low half: streams taken from the front, high half:
// streams taken from the back). The history warp takes four streams at a time from the front
// (four streams x four vectors of type V -- the widest the shift and the row alignment allow --
// per lane loaded before any store: 16 loads in flight); converter warps that have run out of
// stages take single streams from the back. The history warp's loads queue behind the converters'
// in the SM's memory pipeline, and by how much depends on the build (identical code has been
// measured 2x apart), so without the second end it can become the straggler the whole CTA waits for.
template <typename V, bool IDS, bool FRONT>
__device__ __forceinline__ void slide_rows(const CallArgs &a, uint32_t step, uint32_t first, uint32_t n_total,
                                           uint32_t hist_elems, size_t shift, int lane, uint32_t *claims) {
  constexpr uint32_t VW = sizeof(V) / 2;  // int16 elements per vector
  for (;;) {
    uint32_t old = 0;
    if (lane == 0) old = atomicAdd(claims, FRONT ? 4u : 0x10000u);
    old = __shfl_sync(0xffffffffu, old, 0);
    const uint32_t from_front = old & 0xffffu, from_back = old >> 16;
    if (from_front + from_back >= n_total) break;
    // this unit: streams k0 .. n_mine-1
    const uint32_t k0 = FRONT ? from_front : n_total - 1 - from_back;
    const uint32_t n_mine = FRONT ? min(n_total - from_back, from_front + 4u) : k0 + 1;
    for (uint32_t e0 = lane * VW; e0 < hist_elems; e0 += 32 * VW * 4) {
      V val[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const size_t s = !IDS ? first + static_cast<size_t>(k0 + kk) * step
                              : k0 + kk < n_mine ? a.ids[first + (k0 + kk) * step] : 0;
        const int16_t *hsrc = a.hist_src + s * a.hist_stride;
        const int16_t *isrc = a.in + s * a.in_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t e = e0 + j * 32 * VW;
          const size_t src = shift + e;
          if (k0 + kk < n_mine && e < hist_elems)
            val[kk][j] = __ldg(reinterpret_cast<const V *>(src < hist_elems ? hsrc + src : isrc + (src - hist_elems)));
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const size_t s = !IDS ? first + static_cast<size_t>(k0 + kk) * step
                              : k0 + kk < n_mine ? a.ids[first + (k0 + kk) * step] : 0;
        int16_t *hdst = a.hist_dst + s * a.hist_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t e = e0 + j * 32 * VW;
          if (k0 + kk < n_mine && e < hist_elems) *reinterpret_cast<V *>(hdst + e) = val[kk][j];
        }
      }
    }
  }
}

// FAST = every input and output row starts on a 16-byte boundary and the CTA is not part of a
// cluster: the instantiation every BASELINE shape takes. It drops the narrower load / store
// variants and the multicast paths, which is worth having because warps of five roles run
// different parts of this kernel at the same time and share one instruction cache.
// IDS = the launch covers the stream subset a.ids[0 .. a.n_ids) (a cohort of a ragged batch); the
// plain instantiation carries no indirection at all (it costs the long-filter shapes 25 %).
// PACED = the instantiation for long K loops (more than kLeanStages stages). It is the kernel with
// its timeline probes compiled in (never taken unless SPXB_UMMA_TRACE is set) and the converters
// asking for the stage two ahead BEFORE they hand the current one over. Measured, same source
// otherwise: on the 13-16-stage loops of C4 / C5 this instantiation is 8 % faster than the lean one
// (12.6 / 76.8 us against 13.7 / 83.3 us), on C3's 4-stage loop the lean one (no probes, stage
// handed over first) is 10 % faster (6.3 against 7.0 us). Which probes matter was bisected
// (profiles/umma_paced_ab_r1.log): not the ones that used to sit inside the converters' stage
// step (removed: without them the paced kernel gained another 2 %), but the ones around the
// converter loop (slots 2, 3, 4, 30) -- a clock read between the two unrolled stage steps and at
// the loop's ends changes how ptxas schedules the loop; explicit clock-read fences at the end of
// every stage step do not reproduce it. The mechanism is not understood; the two instantiations
// are kept because both measurements are solid.
constexpr uint32_t kLeanStages = 8;
template <int CH, bool FAST, bool IDS, bool PACED>
__global__ void __launch_bounds__(kThreads, 1) umma_fir_kernel(const CallArgs a, const UmmaArgs u) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], acc_bar;
  __shared__ uint32_t tmem_slot, slide_next;

  constexpr int kStreams = kUmmaRows / CH;  // streams per series group
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // series group fastest: CTAs running at the same time share a tap tile (measured: tile-fastest
  // order is slower, the tap stream is bounded by bytes into the SM, not by hot L2 lines)
  const uint32_t g = blockIdx.x % u.n_groups, t = blockIdx.x / u.n_groups;
  const StreamCall sc = a.uniform;
  const uint32_t nt = u.nt;
  // tile geometry: from the kernel parameters (constant bank), else in closed form
  // (umma_plan.cpp: plan_umma_tiles) with the tap-tile slot read from the table in HBM
  const uint32_t m0 = t * nt;
  int kf0;
  if (u.n_inline) {
    kf0 = u.inl[t].kf0;
  } else {
    const unsigned long long tt = static_cast<unsigned long long>(sc.frac0) +
                                  static_cast<unsigned long long>(m0) * a.filt.num;
    const long long q0 = static_cast<long long>(sc.ls0) - (static_cast<long long>(a.filt.taps) - 1) +
                         static_cast<long long>(tt / a.filt.den);
    const long long from_hist = q0 + a.hist_frames;
    kf0 = static_cast<int>(from_hist - (from_hist % kUmmaChunkFrames) - a.hist_frames);
  }
  const uint32_t tap_chunk = 3 * nt * 16;
  const uint32_t tap_stage = kStageChunks * tap_chunk;
  constexpr uint32_t kXPlaneBytes = x_plane(CH), kXStageBytes = x_stage(CH);
  const uint32_t stage_bytes = kXStageBytes + tap_stage;
  const uint32_t n_chunks = 2 * u.ksteps;
  const uint32_t n_iters = (n_chunks + kStageChunks - 1) / kStageChunks;
  const uint32_t S = u.stages;
  const uint32_t n_rows = IDS ? a.n_ids : a.n_streams;  // streams this launch covers
  auto stream_of = [&](uint32_t i) -> size_t { return IDS ? a.ids[i] : i; };
  const uint32_t cluster = FAST ? 1u : u.cluster;
  const uint32_t cta_rank = cluster > 1 ? cluster_ctarank() : 0u;
  const uint16_t cluster_mask = static_cast<uint16_t>((1u << cluster) - 1u);
  // alignment every input row start shares (16-byte items start at multiples of 16 B in a row)
  const uint32_t row_bits = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(a.in)) |
                            (static_cast<uint32_t>(a.in_stride) * 2u);
  const int in_align = FAST ? 16 : (row_bits & 15u) == 0 ? 16 : (row_bits & 7u) == 0 ? 8 : (row_bits & 3u) == 0 ? 4 : 2;

  // Programmatic dependent launch: let the next call's grid start its prologue (barrier init,
  // TMEM allocation) while this grid drains; it blocks in griddepcontrol.wait below until this
  // grid has completed, before it touches anything this grid reads or writes.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) {
    trace_mark(u, 0);
    if (SPXB_TRACE_PTR(u)) {
      unsigned long long gt;
      uint32_t smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      SPXB_TRACE_PTR(u)[static_cast<size_t>(blockIdx.x) * kTraceSlots + 14] = gt;
      SPXB_TRACE_PTR(u)[static_cast<size_t>(blockIdx.x) * kTraceSlots + 16] = smid;
    }
  }
  // The tap-tile producer owns the barriers: it initialises them and has the first tap stages
  // in flight before the grid dependency resolves. (The tile table and the tap pool are only ever
  // rewritten by stream-ordered copies, and a call that re-planned launches without the
  // programmatic edge, so they are safe to read here.)
  uint32_t taps_issued = 0;
  if (warp == kTmaWarp && lane == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], kConvThreads + 1);
      mbar_init(&empty_bar[s], cluster);  // every CTA of the cluster is done with the slot
    }
    mbar_init(&acc_bar, 1);
    slide_next = 0;
    fence_mbar_init();
  }
  // peers multicast into this CTA's stages and arrive on its barriers: all initialised first
  if (cluster > 1) cluster_sync_all();
  auto load_taps = [&](uint32_t it, uint32_t slot, const int8_t *src) {
    const uint32_t chunks_here = min(static_cast<uint32_t>(kStageChunks), n_chunks - it * kStageChunks);
    const uint32_t bytes = chunks_here * tap_chunk;
    uint8_t *dst = smem + slot * stage_bytes + kXStageBytes;
    const int8_t *from = src + static_cast<size_t>(it) * tap_stage;
    mbar_arrive_expect_tx(&full_bar[slot], bytes);
    if (cluster == 1) {
      bulk_g2s(dst, from, bytes, &full_bar[slot]);
    } else {
      const uint32_t part = bytes >> (cluster >> 1), off = part * cta_rank;  // cluster is 2 or 4
      bulk_g2s_multicast(dst + off, from + off, part, &full_bar[slot], cluster_mask);
    }
  };
  if (warp == kTmaWarp && lane == 0) {
    const int8_t *src = u.pool + static_cast<size_t>(u.n_inline ? u.inl[t].slot : u.tiles[t].slot) * u.tile_bytes;
    for (; taps_issued < min(S, n_iters); ++taps_issued) load_taps(taps_issued, taps_issued, src);
    trace_mark(u, 12);
  }
  if (warp == kMmaWarp) {
    tmem_alloc(&tmem_slot, u.tmem_cols);
    tmem_relinquish();
  }

  // ---- converter state: PCM items in flight, two stages deep ----
  // A stage is 64 frames of 128 series = kStreams stream segments of 64*CH*2 bytes. Lanes of a
  // warp walk ALONG a segment in 16-byte items (PPS items per stream, SPI streams per warp
  // instruction), so one LDG.128 covers four full 128-byte lines; each thread owns kItems items
  // per stage, item i of warp w belonging to stream (4w + i) * SPI + lane / PPS.
  constexpr int FPI = 8 / CH;          // frames per 16-byte item
  constexpr int PPS = 64 / FPI;        // items per stream per stage (16 stereo, 8 mono)
  constexpr int SPI = 32 / PPS;        // streams per warp instruction (2 stereo, 4 mono)
  constexpr int kItems = 4;
  const int conv_p = lane % PPS;       // item position inside the stage segment
  // Rows past the end of the batch read row 0 (their results are never stored).
  uint32_t conv_off[kItems];           // byte offset of the item's hi/left word(s) inside an X stage
  const char *cur[kItems];             // where this thread's item of the NEXT fetched stage lives
  int f_next = kf0 + conv_p * FPI;     // its first frame: history for f < 0, this call's input after
  auto input_ptr = [&](int i, int f) {
    const uint32_t sg = g * kStreams + static_cast<uint32_t>((4 * (warp & 7) + i) * SPI + lane / PPS);
    const size_t r = stream_of(sg < n_rows ? sg : 0);
    return reinterpret_cast<const char *>(a.in + r * a.in_stride + static_cast<ptrdiff_t>(f) * CH);
  };
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const uint32_t sl = static_cast<uint32_t>((4 * (warp & 7) + i) * SPI + lane / PPS);
    const uint32_t sg = g * kStreams + sl;
    const size_t r = stream_of(sg < n_rows ? sg : 0);
    cur[i] = f_next < 0 ? reinterpret_cast<const char *>(a.hist_src + r * a.hist_stride +
                                                         (static_cast<ptrdiff_t>(a.hist_frames) + f_next) * CH)
                        : input_ptr(i, f_next);
    // chunk j = 16 frames; inside the chunk row sl holds 16 bytes = 16 frames of one plane
    const uint32_t j = static_cast<uint32_t>(conv_p * FPI) / kUmmaChunkFrames;
    const uint32_t byte_in_row = static_cast<uint32_t>(conv_p * FPI) % kUmmaChunkFrames;
    // stereo: left channel of stream sl -> row 2 sl, right channel -> row 2 sl + 1
    conv_off[i] = (j >> 1) * x_kstep(CH) + (j & 1) * x_lbo(CH) + sl * (16 * CH) + byte_in_row;
  }
  constexpr int kStageFrames = kStageChunks * kUmmaChunkFrames;  // 64
  uint4 raw0[kItems], raw1[kItems];  // two stages of loads in flight (three measured no better)
  // fetch the next stage (stages are fetched strictly in order): pointers just advance by one
  // stage of bytes, except once, where a thread's frames cross from the history into the input
  auto fetch = [&](uint4 (&raw)[kItems]) {
    const int f = f_next;
    f_next += kStageFrames;
    if (f >= 0 && f < kStageFrames) {
#pragma unroll
      for (int i = 0; i < kItems; ++i) cur[i] = input_ptr(i, f);
    }
    const int rem = static_cast<int>(sc.n_in) - f;  // input frames left from f (when f >= 0)
    if (SPXB_DEBUG_BITS(u) & 1u) {
#pragma unroll
      for (int i = 0; i < kItems; ++i) raw[i] = make_uint4(0u, 0u, 0u, 0u);
    } else if (f < 0 || (rem >= FPI && in_align == 16)) {
#pragma unroll
      for (int i = 0; i < kItems; ++i) raw[i] = __ldg(reinterpret_cast<const uint4 *>(cur[i]));
    } else if (rem >= FPI && in_align == 8) {
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        const uint2 lo = __ldg(reinterpret_cast<const uint2 *>(cur[i]));
        const uint2 hi = __ldg(reinterpret_cast<const uint2 *>(cur[i]) + 1);
        raw[i] = make_uint4(lo.x, lo.y, hi.x, hi.y);
      }
    } else if (rem >= FPI && in_align == 4) {
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(cur[i]);
        raw[i] = make_uint4(__ldg(w), __ldg(w + 1), __ldg(w + 2), __ldg(w + 3));
      }
    } else if (rem <= 0) {
#pragma unroll
      for (int i = 0; i < kItems; ++i) raw[i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      // the item holding the end of the input, or rows only 2-byte aligned: sample by sample
#pragma unroll
      for (int i = 0; i < kItems; ++i)
        raw[i] = fetch_item_slow(reinterpret_cast<const int16_t *>(cur[i]), min(rem, FPI) * CH);
    }
#pragma unroll
    for (int i = 0; i < kItems; ++i) cur[i] += kStageFrames * CH * 2;
  };
  // Everything above touched only kernel parameters and this CTA's own resources. The previous
  // call's grid (which reads the history buffer this call overwrites, and writes the one this call
  // reads) must have completed before any global access below.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // the first two stages' loads go out before the setup barrier
  if (warp < kConvWarps) {
    fetch(raw0);
    if (n_iters > 1) fetch(raw1);
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) trace_mark(u, 1);

  if (warp < kConvWarps) {
    // ================= converters: PCM -> byte planes in UMMA layout =================
    auto convert_store = [&](uint8_t *xs, const uint4 (&raw)[kItems]) {
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        uint8_t *base = xs + conv_off[i];
        const uint4 w = raw[i];
        if (CH == 1) {
          // 8 frames; word = (x[2k+1] << 16) | x[2k]: bytes lo0 hi0 lo1 hi1 -> 8 bytes per plane
          const uint2 hi = make_uint2(__byte_perm(w.x, w.y, 0x7531), __byte_perm(w.z, w.w, 0x7531));
          const uint2 lo = make_uint2(__byte_perm(w.x, w.y, 0x6420), __byte_perm(w.z, w.w, 0x6420));
          *reinterpret_cast<uint2 *>(base) = hi;
          *reinterpret_cast<uint2 *>(base + kXPlaneBytes) = lo;
        } else {
          // 4 frames; word f = (R_f << 16) | L_f -> one word per plane and channel (rows 2 sl, 2 sl + 1)
          const uint32_t ul = __byte_perm(w.x, w.y, 0x5140), vl = __byte_perm(w.z, w.w, 0x5140);
          const uint32_t ur = __byte_perm(w.x, w.y, 0x7362), vr = __byte_perm(w.z, w.w, 0x7362);
          *reinterpret_cast<uint32_t *>(base) = __byte_perm(ul, vl, 0x7632);
          *reinterpret_cast<uint32_t *>(base + 16) = __byte_perm(ur, vr, 0x7632);
          *reinterpret_cast<uint32_t *>(base + kXPlaneBytes) = __byte_perm(ul, vl, 0x5410);
          *reinterpret_cast<uint32_t *>(base + kXPlaneBytes + 16) = __byte_perm(ur, vr, 0x5410);
        }
      }
    };
    uint32_t cslot = 0, cpar = 1;  // ring slot of the stage being filled, parity its `empty` wait expects
    uint8_t *cstage = smem;
    auto stage_step = [&](uint32_t it, uint4 (&raw)[kItems]) {
      const uint32_t slot = cslot, par = cpar ^ 1u;
      mbar_wait(&empty_bar[slot], par ^ 1u);
      if (!(SPXB_DEBUG_BITS(u) & 4u)) convert_store(cstage, raw);
      if (PACED) {
        if (it + 2 < n_iters) fetch(raw);
        fence_proxy_async_smem();
        mbar_arrive(&full_bar[slot]);
      } else {
        // short loops: the tensor core gets the stage first, then the loads for two stages ahead go out
        fence_proxy_async_smem();
        mbar_arrive(&full_bar[slot]);
        if (it + 2 < n_iters) fetch(raw);
      }
      cstage += stage_bytes;
      if (++cslot == S) {
        cslot = 0;
        cpar ^= 1u;
        cstage = smem;
      }
    };

    if (tid == 0) trace_mark(u, 2);
    for (uint32_t it = 0; it < n_iters; it += 2) {
      stage_step(it, raw0);
      if (tid == 0 && it == 0) trace_mark(u, 3);
      if (it + 1 < n_iters) stage_step(it + 1, raw1);
    }
    if (tid == 0) trace_mark(u, 4);
    if (tid == kConvThreads - 32) trace_mark(u, 30);

    // out of stages to fill: help with the history slide (long histories, 16-byte vectors only)
    {
      const uint32_t hist_elems = a.hist_frames * CH;
      const size_t shift = static_cast<size_t>(sc.consumed) * CH;
      if (hist_elems >= 512 && shift % 8 == 0 && in_align == 16) {
        const uint32_t first = g * kStreams + t;
        const uint32_t in_group = t < static_cast<uint32_t>(kStreams) ? (kStreams - t + u.n_tiles - 1) / u.n_tiles : 0u;
        const uint32_t in_batch = first < n_rows ? (n_rows - first + u.n_tiles - 1) / u.n_tiles : 0u;
        slide_rows<uint4, IDS, false>(a, u.n_tiles, first, min(in_group, in_batch), hist_elems, shift, lane, &slide_next);
      }
    }

    // ================= epilogue =================
    mbar_wait(&acc_bar, 0);
    tc_fence_after_sync();
    if (tid == 0) trace_mark(u, 5);
    // Straight from TMEM to the interleaved int16 output: a lane owns one series (TMEM lane) and 16
    // consecutive outputs per column group; mono packs them into 32 contiguous bytes, stereo first
    // swaps halves with the neighbouring lane (the other channel of the same stream) so that each
    // lane of the pair holds 8 whole frames = 32 contiguous bytes.
    const uint32_t n_valid = min(nt, sc.n_out - m0);
    const uint32_t row = (warp & 3) * 32 + lane;  // TMEM lane
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t sl_out = CH == 2 ? row >> 1 : row, ch_out = CH == 2 ? (row & 1u) : 0u;
    const uint32_t s_out = g * kStreams + sl_out;
    const bool live_out = s_out < n_rows;
    int16_t *out_row = a.out + stream_of(live_out ? s_out : 0) * a.out_stride +
                       static_cast<size_t>(m0) * CH;
    const uint32_t out_bits = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(a.out)) |
                              (static_cast<uint32_t>(a.out_stride) * 2u);
    const int out_align = FAST ? 16 : (out_bits & 15u) == 0 ? 16 : (out_bits & 3u) == 0 ? 4 : 2;
    for (uint32_t cg = warp >> 2; cg * 16 < n_valid; cg += 2) {
      uint32_t p0[16], p1[16], p2[16], p3[16];
      tmem_ld16(lane_addr + cg * 16, p0);
      tmem_ld16(lane_addr + nt + cg * 16, p1);
      tmem_ld16(lane_addr + 2 * nt + cg * 16, p2);
      tmem_ld16(lane_addr + 3 * nt + cg * 16, p3);
      tmem_ld_wait();
      int r16[16];
      combine16(p0, p1, p2, p3, u.shift, r16);  // rounded, not yet saturated
      uint32_t w[8];       // this lane's 16 int16 values = 32 contiguous output bytes
      uint32_t first_elem;  // their position in the stream's row, in int16 elements from m0
      if (CH == 2) {
        // lane pair (left, right): left keeps frames [0,8), right keeps frames [8,16)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int send = ch_out == 0 ? r16[8 + j] : r16[j];
          const int recv = __shfl_xor_sync(0xffffffffu, send, 1);
          w[j] = ch_out == 0 ? pack_sat_s16x2(recv, r16[j]) : pack_sat_s16x2(r16[8 + j], recv);
        }
        first_elem = (cg * 16 + ch_out * 8) * 2;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = pack_sat_s16x2(r16[2 * j + 1], r16[2 * j]);
        first_elem = cg * 16;
      }
      if (!live_out) continue;
      const uint32_t total = n_valid * CH;
      const uint32_t n_here = first_elem >= total ? 0u : min(16u, total - first_elem);  // int16 elements
      int16_t *dst = out_row + first_elem;
      if (n_here == 16 && out_align == 16) {
        reinterpret_cast<uint4 *>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint4 *>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
      } else if (out_align >= 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (2u * j + 1 < n_here) reinterpret_cast<uint32_t *>(dst)[j] = w[j];
        if (n_here & 1u) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (2u * j + 1 == n_here) dst[2 * j] = static_cast<int16_t>(w[j] & 0xffffu);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (2u * j < n_here) dst[2 * j] = static_cast<int16_t>(w[j] & 0xffffu);
          if (2u * j + 1 < n_here) dst[2 * j + 1] = static_cast<int16_t>(w[j] >> 16);
        }
      }
    }
    if (tid == 0) trace_mark(u, 6);
    if (tid == 0) trace_mark(u, 7);
  } else if (warp == kTmaWarp) {
    // ================= tap tiles: one bulk copy per stage =================
    if (lane == 0) {
      const int8_t *src = u.pool + static_cast<size_t>(u.n_inline ? u.inl[t].slot : u.tiles[t].slot) * u.tile_bytes;
      uint32_t slot = taps_issued % S, par = (taps_issued / S) & 1u;
      for (uint32_t it = taps_issued; it < n_iters; ++it) {
        mbar_wait(&empty_bar[slot], par ^ 1u);
        load_taps(it, slot, src);
        if (++slot == S) {
          slot = 0;
          par ^= 1u;
        }
      }
      trace_mark(u, 13);
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================= MMA issue =================
    // The whole warp walks the stages (uniform control flow, descriptors in uniform registers);
    // one elected lane issues. N = 3nt splits into pieces of at most 256 columns.
    const uint32_t n3 = 3 * nt;
    const uint32_t np0 = min(n3, 256u), np1 = n3 - np0;            // [0, 3nt)
    const uint32_t nq0 = min(2 * nt, 256u), nq1 = 2 * nt - nq0;    // [0, 2nt) (first K step, lo plane)
    const uint32_t id_hi0 = umma_idesc_i8(128, np0, true, true), id_hi1 = umma_idesc_i8(128, np1, true, true);
    const uint32_t id_lo0 = umma_idesc_i8(128, np0, false, true), id_lo1 = umma_idesc_i8(128, np1, false, true);
    const uint32_t id_lq0 = umma_idesc_i8(128, nq0, false, true), id_lq1 = umma_idesc_i8(128, nq1, false, true);
    const uint32_t id_lf = umma_idesc_i8(128, nt, false, true);
    const uint64_t a_base = umma_smem_desc(smem_u32(smem), x_lbo(CH), 128);
    const uint64_t b_base = umma_smem_desc(smem_u32(smem) + kXStageBytes, tap_chunk, 128);
    const uint32_t stage_step16 = stage_bytes >> 4;
    const uint32_t a_ks16 = x_kstep(CH) >> 4, a_lo16 = kXPlaneBytes >> 4, b_ks16 = (2 * tap_chunk) >> 4;
    // One K step = hi plane x B into [0,3nt), lo plane x the same B into [nt,4nt).
    auto kstep = [&](uint64_t a_hi, uint64_t b) {
      const uint64_t a_lo = a_hi + a_lo16;
      umma_i8(tmem, a_hi, b, id_hi0, 1u);
      if (np1) umma_i8(tmem + 256, a_hi, b + 256, id_hi1, 1u);
      umma_i8(tmem + nt, a_lo, b, id_lo0, 1u);
      if (np1) umma_i8(tmem + nt + 256, a_lo, b + 256, id_lo1, 1u);
    };
    uint32_t slot = 0, par = 0;
    uint64_t a_st = a_base, b_st = b_base;
    const bool odd_tail = (n_chunks % kStageChunks) != 0;  // last stage holds one K step
    for (uint32_t it = 0; it < n_iters; ++it) {
      mbar_wait(&full_bar[slot], par);
      tc_fence_after_sync();
      if (SPXB_TRACE_PTR(u) && lane == 0 && it < 12) trace_mark(u, 20 + it);
      const bool last = it + 1 == n_iters;
      if (elect_one()) {
        if (!(SPXB_DEBUG_BITS(u) & 2u)) {
          if (it == 0) {
            umma_i8(tmem, a_st, b_st, id_hi0, 0u);
            if (np1) umma_i8(tmem + 256, a_st, b_st + 256, id_hi1, 0u);
            // columns [nt,3nt) already hold hi*B: accumulate; columns [3nt,4nt) are fresh
            umma_i8(tmem + nt, a_st + a_lo16, b_st, id_lq0, 1u);
            if (nq1) umma_i8(tmem + nt + 256, a_st + a_lo16, b_st + 256, id_lq1, 1u);
            umma_i8(tmem + 3 * nt, a_st + a_lo16, b_st + 2 * nt, id_lf, 0u);
          } else {
            kstep(a_st, b_st);
          }
          if (!(last && odd_tail)) kstep(a_st + a_ks16, b_st + b_ks16);
        }
        if (cluster == 1) umma_commit(&empty_bar[slot]);
        else umma_commit_multicast(&empty_bar[slot], cluster_mask);
        if (last) umma_commit(&acc_bar);
      }
      __syncwarp();
      a_st += stage_step16;
      b_st += stage_step16;
      if (++slot == S) {
        slot = 0;
        par ^= 1u;
        a_st = a_base;
        b_st = b_base;
      }
    }
    if (lane == 0) trace_mark(u, 11);
  } else {
    // ================= history slide (resample.c:898-899) and the new position =================
    // Runs beside the FIR: it reads the old history and this call's input, writes the other half
    // of the ping-pong. Stream sl of the group is handled by the tile CTA with t == sl % n_tiles;
    // new history element e = element consumed*CH + e of (old history || input).
    const uint32_t hist_elems = a.hist_frames * CH;
    const size_t shift = static_cast<size_t>(sc.consumed) * CH;
    const int vshift = (shift % 8 == 0) ? 8 : (shift % 4 == 0) ? 4 : (shift % 2 == 0) ? 2 : 1;
    const int vw = min(vshift, in_align / 2);
    // streams of this CTA: sl = t + k * n_tiles while the stream exists
    const uint32_t first = g * kStreams + t;
    const uint32_t in_group = t < static_cast<uint32_t>(kStreams) ? (kStreams - t + u.n_tiles - 1) / u.n_tiles : 0u;
    const uint32_t in_batch = first < n_rows ? (n_rows - first + u.n_tiles - 1) / u.n_tiles : 0u;
    const uint32_t n_mine = min(in_group, in_batch);
    if (vw == 8) slide_rows<uint4, IDS, true>(a, u.n_tiles, first, n_mine, hist_elems, shift, lane, &slide_next);
    else if (vw == 4) slide_rows<uint2, IDS, true>(a, u.n_tiles, first, n_mine, hist_elems, shift, lane, &slide_next);
    else if (vw == 2) slide_rows<uint32_t, IDS, true>(a, u.n_tiles, first, n_mine, hist_elems, shift, lane, &slide_next);
    else slide_rows<uint16_t, IDS, true>(a, u.n_tiles, first, n_mine, hist_elems, shift, lane, &slide_next);
    for (uint32_t k = lane; k < n_mine; k += 32) {
      const size_t s = stream_of(first + k * u.n_tiles);
      a.last_sample[s] = sc.ls1;
      a.samp_frac[s] = sc.frac1;
    }
    if (lane == 0) trace_mark(u, 8);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, u.tmem_cols);
  // no CTA may leave while a peer can still multicast into it or arrive on its barriers
  if (cluster > 1) cluster_sync_all();
  if (tid == 0 && SPXB_TRACE_PTR(u)) {
    trace_mark(u, 9);
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    SPXB_TRACE_PTR(u)[static_cast<size_t>(blockIdx.x) * kTraceSlots + 15] = gt;
  }
}

// One thread per (chunk, row): 16 digits -> one 16-byte store. jobs[i] = {slot, phase0, delta}.
__global__ void build_tap_tiles_kernel(const int32_t *__restrict__ h, uint32_t num, uint32_t den, uint32_t taps,
                                       uint32_t nt, uint32_t ksteps, const uint32_t *__restrict__ jobs,
                                       int8_t *pool, uint32_t tile_bytes) {
  const uint32_t slot = jobs[3 * blockIdx.y], phase0 = jobs[3 * blockIdx.y + 1], delta = jobs[3 * blockIdx.y + 2];
  const uint32_t rows = 3 * nt, cells = 2 * ksteps * rows;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells) return;
  const uint32_t c = idx / rows, r = idx % rows;
  const uint32_t n = r % nt, digit = 2 - r / nt;
  const unsigned long long tt = static_cast<unsigned long long>(phase0) + static_cast<unsigned long long>(n) * num;
  const uint32_t phase = static_cast<uint32_t>(tt % den);
  const long long first = static_cast<long long>(delta) + static_cast<long long>(tt / den);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const long long j = static_cast<long long>(c) * 16 + e - first;
    int d = 0;
    if (j >= 0 && j < static_cast<long long>(taps)) {
      const int32_t v = h[static_cast<size_t>(phase) * taps + static_cast<size_t>(j)];
      const int32_t d0 = ((v + 128) & 255) - 128;
      const int32_t r1 = (v - d0) >> 8;
      const int32_t d1 = ((r1 + 128) & 255) - 128;
      const int32_t d2 = (r1 - d1) >> 8;
      d = digit == 0 ? d0 : digit == 1 ? d1 : d2;
    }
    w[e >> 2] |= (static_cast<uint32_t>(d) & 0xffu) << (8 * (e & 3));
  }
  *reinterpret_cast<uint4 *>(pool + static_cast<size_t>(slot) * tile_bytes + static_cast<size_t>(idx) * 16) =
      make_uint4(w[0], w[1], w[2], w[3]);
}

uint32_t pow2_cols(uint32_t cols) {
  uint32_t c = 32;
  while (c < cols) c <<= 1;
  return c;
}

}  // namespace

namespace {

constexpr size_t kMaxPoolBytes = 256ull << 20;

bool inline_tiles_enabled() {
  static const bool v = [] {
    const char *e = getenv("SPXB_UMMA_INLINE_TILES");
    return !e || atoi(e) != 0;
  }();
  return v;
}

uint32_t forced_nt() {
  static const uint32_t v = [] {
    const char *e = getenv("SPXB_UMMA_NT");
    return e ? static_cast<uint32_t>(atoi(e)) : 0u;
  }();
  return v;
}

// Tile width: fewest estimated cycles for the whole grid, from the measured timeline of a tile
// (profiles/umma_trace_r1.log): ~5500 cycles of prologue + epilogue, then per 64-frame stage the
// slower of the MMA floor (two K steps x two planes x 3nt/2 cycles) and the stage's operand bytes
// (PCM planes + tap tile) at the ~30 B/clk an SM gets from L2 when every SM streams; the grid
// runs in waves of one CTA per SM.
uint32_t pick_nt(const UmmaContext &c, uint32_t n_groups, uint32_t n_out) {
  if (forced_nt() >= 16 && forced_nt() <= 128 && forced_nt() % 16 == 0) return forced_nt();
  uint32_t best = 16;
  double best_cost = 1e300;
  for (uint32_t nt = 16; nt <= 128; nt += 16) {
    const uint32_t tiles = (n_out + nt - 1) / nt;
    const double ctas = static_cast<double>(tiles) * n_groups;
    const double waves = std::ceil(ctas / c.sm_count);
    const uint32_t ks = umma_ksteps(c.spec.taps, c.spec.num, c.spec.den, nt);
    const double stage_cycles = std::max(6.0 * nt, (1.0 * x_stage(static_cast<int>(c.channels)) + 4.0 * 48.0 * nt) / 30.0);
    const double tile_cycles = 5500.0 + stage_cycles * ((ks + 1) / 2) + 10.0 * nt;
    const double cost = waves * tile_cycles;
    if (cost < best_cost) {
      best_cost = cost;
      best = nt;
    }
  }
  return best;
}

void drop_pool(UmmaContext *c) {
  c->pool_generation += 1;
  if (c->d_pool) cudaFree(c->d_pool);
  c->d_pool = nullptr;
  c->pool_cap = 0;
  c->slot_of.clear();
  c->memo = false;
}

}  // namespace

UmmaContext *umma_create(const FilterSpec &spec, const std::vector<float> &ref_table, uint32_t channels,
                         int sm_count) {
  if (channels != 1 && channels != 2) return nullptr;
  if (spec.taps > 32768u) return nullptr;  // int32 accumulators: N * 255 * 128 * 2 < 2^31
  if (static_cast<uint64_t>(spec.den) * spec.taps * sizeof(int32_t) > (64ull << 20)) return nullptr;
  UmmaContext *c = new UmmaContext;
  c->spec = spec;
  c->channels = channels;
  c->sm_count = sm_count;
  if (!build_fixed_taps(spec, ref_table, &c->ft)) {
    delete c;
    return nullptr;
  }
  if (cudaMalloc(reinterpret_cast<void **>(&c->d_h), c->ft.h.size() * sizeof(int32_t)) != cudaSuccess ||
      cudaMemcpy(c->d_h, c->ft.h.data(), c->ft.h.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    if (c->d_h) cudaFree(c->d_h);
    delete c;
    return nullptr;
  }
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    auto big_smem = [](auto kernel) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    };
    big_smem(umma_fir_kernel<1, false, false, false>);
    big_smem(umma_fir_kernel<1, false, false, true>);
    big_smem(umma_fir_kernel<2, false, false, false>);
    big_smem(umma_fir_kernel<2, false, false, true>);
    big_smem(umma_fir_kernel<1, true, false, false>);
    big_smem(umma_fir_kernel<1, true, false, true>);
    big_smem(umma_fir_kernel<2, true, false, false>);
    big_smem(umma_fir_kernel<2, true, false, true>);
    big_smem(umma_fir_kernel<1, false, true, false>);
    big_smem(umma_fir_kernel<1, false, true, true>);
    big_smem(umma_fir_kernel<2, false, true, false>);
    big_smem(umma_fir_kernel<2, false, true, true>);
    umma2_configure_device();
    configured_dev = dev;
  }
  return c;
}

void umma_destroy(UmmaContext *c) {
  if (!c) return;
  if (c->d_h) cudaFree(c->d_h);
  if (c->d_pool) cudaFree(c->d_pool);
  if (c->d_tiles) cudaFree(c->d_tiles);
  if (c->d_jobs) cudaFree(c->d_jobs);
  if (c->d_kplan) cudaFree(c->d_kplan);
  if (c->d_kdev) cudaFree(c->d_kdev);
  if (c->d_trace) cudaFree(c->d_trace);
  delete c;
}

int umma_shift(const UmmaContext *c) { return c ? c->ft.shift : 0; }

void umma_geometry(const UmmaContext *c, uint32_t out[6]) {
  for (int i = 0; i < 6; ++i) out[i] = 0;
  if (!c || !c->memo) return;
  out[0] = c->nt;
  out[1] = c->ksteps;
  out[2] = c->n_tiles;
  out[3] = c->m_groups;
  out[4] = c->stages;
  out[5] = c->smem_bytes;
}

bool umma_prepare(UmmaContext *c, const CallArgs &a, cudaStream_t stream, cudaError_t *err) {
  *err = cudaSuccess;
  if (!c || a.per_stream != nullptr) return false;  // streams at different positions
  if (a.channels != c->channels || a.uniform.n_out == 0) return false;
  if ((reinterpret_cast<uintptr_t>(a.hist_src) & 15) != 0 || a.hist_stride % 8 != 0 || a.hist_frames % 16 != 0)
    return false;
  if (a.uniform.n_in > 0x3fffffffu || a.uniform.ls0 > 0x3fffffff || a.uniform.ls0 < 0) return false;
  const uint32_t n_rows = a.ids ? a.n_ids : a.n_streams;
  const uint32_t n_groups = (n_rows * a.channels + kUmmaRows - 1) / kUmmaRows;
  const StreamCall &sc = a.uniform;
  // Which tensor kernel: the persistent, TMA-fed one (kernels_umma2.cu) for long filters -- 8 or more
  // 64-frame stages per tile, where a stage's fixed costs and the PCM path dominate (C4 -9 %, C5 -10 %)
  // -- when the call is one it covers (whole batch, 16-byte aligned rows); the one-tile-per-CTA kernel
  // otherwise (short filters such as C3 are a single wave of a few stages per CTA; there its shorter
  // prologue wins, 6.3 vs 6.9 us). SPXB_UMMA_RESIDENT=0 / 1 forces the choice.
  static const int force_resident = [] {
    const char *e = getenv("SPXB_UMMA_RESIDENT");
    return e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  const bool long_filter = umma_ksteps(c->spec.taps, c->spec.num, c->spec.den, 64) >= 15;
  const bool want_resident = (force_resident == 1 || (force_resident < 0 && long_filter)) && umma2_covers(a);
  if (c->memo && c->m_ls0 == sc.ls0 && c->m_frac0 == sc.frac0 && c->m_n_out == sc.n_out &&
      c->m_hist_frames == a.hist_frames && c->m_groups == n_groups && c->resident_wanted == want_resident) {
    return true;  // steady state: same tiles as the previous call
  }
  const uint64_t ops_before = c->stream_ops;

  uint32_t nt = pick_nt(*c, n_groups, sc.n_out);
  // the persistent kernel keeps its A ring behind the accumulator: 4 nt + 64 columns of the 512
  if (want_resident && umma2_planes_in_tmem() && nt > 112) nt = 112;
  if ((nt != c->nt || want_resident != c->resident_wanted) && c->frozen) {
    *err = cudaErrorNotSupported;
    return false;
  }
  if (nt != c->nt || want_resident != c->resident_wanted) {
    drop_pool(c);
    c->nt = nt;
    c->resident_wanted = want_resident;
    c->ksteps = umma_ksteps(c->spec.taps, c->spec.num, c->spec.den, nt);
    c->tmem_cols = pow2_cols(4 * nt);
    // the persistent kernel with a packed, shared-memory-resident tap tile when that tile fits
    c->resident = false;
    uint32_t x_stages = 0;
    if (want_resident && build_packed_plan(c->spec, c->ft, nt, &c->packed))
      x_stages = umma2_x_stages(a.channels, c->packed.tile_bytes, c->ksteps);
    if (x_stages >= 2) {
      c->stream_ops += 1;
      if (c->d_kplan) {
        cudaStreamSynchronize(stream);
        cudaFree(c->d_kplan);
        c->d_kplan = nullptr;
      }
      if ((*err = cudaMalloc(reinterpret_cast<void **>(&c->d_kplan), c->ksteps * sizeof(UmmaKStep))) != cudaSuccess ||
          (*err = cudaMemcpyAsync(c->d_kplan, c->packed.k.data(), c->ksteps * sizeof(UmmaKStep), cudaMemcpyHostToDevice,
                                  stream)) != cudaSuccess) {
        c->nt = 0;
        return false;
      }
      cudaStreamSynchronize(stream);  // pageable source; geometry changes are rare
      if ((*err = umma2_upload_plan(c, stream)) != cudaSuccess) {
        c->nt = 0;
        return false;
      }
      c->resident = true;
      // two accumulator sets only in the variant with the byte planes in shared memory (SPXB_UMMA_ATMEM=0);
      // the default kernel keeps one set and the A ring in the remaining TMEM columns
      c->n_acc = (!umma2_planes_in_tmem() && 8 * nt <= 512) ? 2u : 1u;
      c->tmem_cols = pow2_cols(c->n_acc * 4 * nt);
      c->tile_bytes = c->packed.tile_bytes;
      c->stages = x_stages;
      c->smem_bytes = x_stages * x_slot(static_cast<int>(a.channels)) + c->tile_bytes;
    } else {
      c->tile_bytes = 2 * c->ksteps * 3 * nt * 16;
      const uint32_t stage_bytes = x_stage(static_cast<int>(a.channels)) + kStageChunks * 3 * nt * 16;
      const uint32_t n_iters = (2 * c->ksteps + kStageChunks - 1) / kStageChunks;
      uint32_t stages = std::min<uint32_t>(kMaxStages, kMaxSmem / stage_bytes);
      static const uint32_t stage_cap = [] {  // experiment: shared memory given up = L1 gained
        const char *e = getenv("SPXB_UMMA_STAGES");
        return e ? static_cast<uint32_t>(atoi(e)) : 0u;
      }();
      if (stage_cap) stages = std::min(stages, stage_cap);
      stages = std::max(1u, std::min(stages, n_iters));
      c->stages = stages;
      c->smem_bytes = stages * stage_bytes;
      if (c->smem_bytes > kMaxSmem) return false;
    }
  }
  std::vector<UmmaTile> tiles;
  std::vector<UmmaTileKey> keys;
  plan_umma_tiles(c->spec.num, c->spec.den, c->spec.taps, a.hist_frames, sc.ls0, sc.frac0, sc.n_out, nt, &tiles,
                  &keys);
  // tap tiles this call needs that the pool lacks
  std::vector<uint32_t> jobs;
  size_t next_slot = c->slot_of.size();
  std::unordered_map<uint64_t, uint32_t> fresh;
  for (size_t i = 0; i < tiles.size(); ++i) {
    const uint64_t key = static_cast<uint64_t>(keys[i].phase0) * 16 + keys[i].delta;
    auto it = c->slot_of.find(key);
    if (it != c->slot_of.end()) {
      tiles[i].slot = it->second;
      continue;
    }
    auto f = fresh.find(key);
    if (f != fresh.end()) {
      tiles[i].slot = f->second;
      continue;
    }
    const uint32_t slot = static_cast<uint32_t>(next_slot++);
    fresh.emplace(key, slot);
    tiles[i].slot = slot;
    jobs.push_back(slot);
    jobs.push_back(keys[i].phase0);
    jobs.push_back(keys[i].delta);
  }
  if (next_slot * static_cast<size_t>(c->tile_bytes) > kMaxPoolBytes) return false;
  const bool table_in_params = c->resident ? tiles.size() <= 32 : inline_tiles_enabled() && tiles.size() <= kInlineTiles;
  if (c->frozen && (!jobs.empty() || next_slot > c->pool_cap || !table_in_params)) {
    *err = cudaErrorNotSupported;  // would need the stream / an allocation while it is being captured
    return false;
  }
  if (next_slot > c->pool_cap) {
    c->stream_ops += 1;
    c->pool_generation += 1;
    // grow: existing tiles are rebuilt rather than copied (rare; geometry changes only)
    const size_t want = std::max(next_slot, c->pool_cap * 2);
    const size_t cap = std::min(want, kMaxPoolBytes / c->tile_bytes);
    int8_t *np = nullptr;
    if ((*err = cudaMalloc(reinterpret_cast<void **>(&np), cap * c->tile_bytes)) != cudaSuccess) return false;
    if (c->d_pool) {
      // queued kernels may still read the old pool
      cudaStreamSynchronize(stream);
      cudaMemcpyAsync(np, c->d_pool, c->slot_of.size() * static_cast<size_t>(c->tile_bytes),
                      cudaMemcpyDeviceToDevice, stream);
      cudaStreamSynchronize(stream);
      cudaFree(c->d_pool);
    }
    c->d_pool = np;
    c->pool_cap = cap;
  }
  if (!jobs.empty()) {
    c->stream_ops += 1;
    const size_t n_jobs = jobs.size() / 3;
    if (n_jobs > c->jobs_cap) {
      if (c->d_jobs) {
        cudaStreamSynchronize(stream);
        cudaFree(c->d_jobs);
      }
      c->jobs_cap = std::max<size_t>(n_jobs, 64);
      if ((*err = cudaMalloc(reinterpret_cast<void **>(&c->d_jobs), c->jobs_cap * 3 * sizeof(uint32_t))) != cudaSuccess)
        return false;
    }
    // pageable source: the copy is staged before the call returns, so `jobs` may go away
    if ((*err = cudaMemcpyAsync(c->d_jobs, jobs.data(), jobs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                stream)) != cudaSuccess)
      return false;
    if (c->resident) {
      if ((*err = umma2_build_tiles(c, c->d_jobs, n_jobs, stream)) != cudaSuccess) return false;
    } else {
      const uint32_t cells = 2 * c->ksteps * 3 * nt;
      const dim3 grid((cells + 255) / 256, static_cast<unsigned>(n_jobs));
      build_tap_tiles_kernel<<<grid, 256, 0, stream>>>(c->d_h, c->spec.num, c->spec.den, c->spec.taps, nt, c->ksteps,
                                                       c->d_jobs, c->d_pool, c->tile_bytes);
      if ((*err = cudaGetLastError()) != cudaSuccess) return false;
    }
    for (auto &kv : fresh) c->slot_of.emplace(kv.first, kv.second);
  }
  // the tile table travels in the kernel parameters when it fits; only longer tables go to HBM
  if (!table_in_params) {
    c->stream_ops += 1;
    if (tiles.size() > c->tiles_cap) {
      if (c->d_tiles) {
        cudaStreamSynchronize(stream);
        cudaFree(c->d_tiles);
      }
      c->tiles_cap = std::max<size_t>(tiles.size(), 64);
      if ((*err = cudaMalloc(reinterpret_cast<void **>(&c->d_tiles), c->tiles_cap * sizeof(UmmaTile))) != cudaSuccess)
        return false;
    }
    if ((*err = cudaMemcpyAsync(c->d_tiles, tiles.data(), tiles.size() * sizeof(UmmaTile), cudaMemcpyHostToDevice,
                                stream)) != cudaSuccess)
      return false;
  }
  // a re-plan that uploaded or built something launches without the programmatic edge (the next
  // kernel's prologue reads the tile table / tap pool before its grid dependency resolves)
  if (c->stream_ops != ops_before) c->fresh_plan = true;
  c->n_tiles = static_cast<uint32_t>(tiles.size());
  c->h_tiles = tiles;
  c->memo = true;
  c->m_ls0 = sc.ls0;
  c->m_frac0 = sc.frac0;
  c->m_n_out = sc.n_out;
  c->m_hist_frames = a.hist_frames;
  c->m_groups = n_groups;
  // Cluster: CTAs of one tile share the tap stream; multicast cuts each SM's requests for it.
  // Only worth it with several groups; padding CTAs (no live streams) cost a whole tile each.
  static const int forced_cluster = [] {
    const char *e = getenv("SPXB_UMMA_CLUSTER");
    return e ? atoi(e) : 0;
  }();
  uint32_t cl = 1;
  if (forced_cluster == 1 || forced_cluster == 2 || forced_cluster == 4) {
    cl = static_cast<uint32_t>(forced_cluster);
  } else {
    // measured (profiles/umma_cluster_sweep_r1.log): multicast pairs / quads are 7-20 % SLOWER than
    // independent CTAs -- the tap stream is not the binding resource, shared-memory bandwidth is,
    // and the lock-step of a cluster costs more than the saved L2 requests. Kept as an option.
    cl = 1;
  }
  c->cluster = cl;
  c->grid_groups = (n_groups + cl - 1) / cl * cl;
  return true;
}

void umma_set_frozen(UmmaContext *c, bool frozen) {
  if (c) c->frozen = frozen;
}
uint64_t umma_stream_ops(const UmmaContext *c) { return c ? c->stream_ops : 0; }
uint64_t umma_pool_generation(const UmmaContext *c) { return c ? c->pool_generation : 0; }

cudaError_t launch_umma(UmmaContext *c, const CallArgs &a, cudaStream_t stream, uint32_t *launches) {
  if (c->resident) return launch_umma2(c, a, stream, launches);
  UmmaArgs u;
  u.tiles = c->d_tiles;
  u.n_tiles = c->n_tiles;
  u.n_inline = (inline_tiles_enabled() && c->n_tiles <= kInlineTiles) ? c->n_tiles : 0u;
  for (uint32_t i = 0; i < kInlineTiles; ++i) {
    u.inl[i].kf0 = i < u.n_inline ? c->h_tiles[i].kf0 : 0;
    u.inl[i].slot = i < u.n_inline ? c->h_tiles[i].slot : 0u;
  }
  u.n_groups = c->grid_groups;
  u.cluster = c->cluster;
  u.pool = c->d_pool;
  u.tile_bytes = c->tile_bytes;
  u.nt = c->nt;
  u.ksteps = c->ksteps;
  u.stages = c->stages;
  u.tmem_cols = c->tmem_cols;
  u.shift = c->ft.shift;
  u.trace = nullptr;
  static const uint32_t debug_bits = [] {
    const char *e = getenv("SPXB_UMMA_DEBUG");
    return e ? static_cast<uint32_t>(atoi(e)) : 0u;
  }();
  u.debug = debug_bits;
  const uint64_t grid = static_cast<uint64_t>(u.n_groups) * u.n_tiles;
  if (grid == 0 || grid > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  static const bool want_trace = getenv("SPXB_UMMA_TRACE") != nullptr;
  if (want_trace) {
    if (c->trace_ctas < grid) {
      if (c->d_trace) cudaFree(c->d_trace);
      c->d_trace = nullptr;
      if (cudaMalloc(reinterpret_cast<void **>(&c->d_trace), grid * kTraceSlots * sizeof(unsigned long long)) ==
          cudaSuccess)
        c->trace_ctas = grid;
    }
    u.trace = c->d_trace;
  }
  static const bool use_pdl = [] {
    const char *e = getenv("SPXB_UMMA_PDL");
    return !e || atoi(e) != 0;
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = c->smem_bytes;
  static const uint32_t pad_smem = [] {  // experiment: what a smaller L1 carve-out costs this kernel
    const char *e = getenv("SPXB_UMMA_PAD_SMEM");
    return e ? static_cast<uint32_t>(atoi(e)) : 0u;
  }();
  if (pad_smem) cfg.dynamicSmemBytes = std::min<uint32_t>(kMaxSmem, c->smem_bytes + pad_smem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n_attr = 0;
  if (c->cluster > 1) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = c->cluster;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  if (use_pdl && !c->fresh_plan) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  static const bool allow_fast = [] {
    const char *e = getenv("SPXB_UMMA_FAST");
    return !e || atoi(e) != 0;
  }();
  const uintptr_t row_bits = reinterpret_cast<uintptr_t>(a.in) | (a.in_stride * 2u) |
                             reinterpret_cast<uintptr_t>(a.out) | (a.out_stride * 2u);
  const bool fast = allow_fast && c->cluster == 1 && (row_bits & 15u) == 0;
  // long K loops take the paced instantiation (and so does a traced launch: the marks live there)
  const uint32_t n_stages_run = (2 * c->ksteps + kStageChunks - 1) / kStageChunks;
  static const int forced_paced = [] {
    const char *e = getenv("SPXB_UMMA_PACED");
    return e ? atoi(e) : -1;
  }();
  const bool paced = u.trace != nullptr || (forced_paced >= 0 ? forced_paced != 0 : n_stages_run > kLeanStages);
  auto launch = [&](auto kernel) { return cudaLaunchKernelEx(&cfg, kernel, a, u); };
  auto by_pace = [&](auto lean_kernel, auto paced_kernel) { return paced ? launch(paced_kernel) : launch(lean_kernel); };
  cudaError_t e;
  if (a.ids) {  // a cohort of a ragged batch: the generic instantiation with the id indirection
    e = a.channels == 2 ? by_pace(umma_fir_kernel<2, false, true, false>, umma_fir_kernel<2, false, true, true>)
                        : by_pace(umma_fir_kernel<1, false, true, false>, umma_fir_kernel<1, false, true, true>);
  } else if (a.channels == 2) {
    e = fast ? by_pace(umma_fir_kernel<2, true, false, false>, umma_fir_kernel<2, true, false, true>)
             : by_pace(umma_fir_kernel<2, false, false, false>, umma_fir_kernel<2, false, false, true>);
  } else {
    e = fast ? by_pace(umma_fir_kernel<1, true, false, false>, umma_fir_kernel<1, true, false, true>)
             : by_pace(umma_fir_kernel<1, false, false, false>, umma_fir_kernel<1, false, false, true>);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) c->fresh_plan = false;
  if (e == cudaSuccess && launches) *launches += 1;
  return e;
}

// debug timeline of the last traced launch: kTraceSlots words per CTA; returns CTAs copied
long umma_read_trace(const UmmaContext *c, unsigned long long *dst, size_t cap_words) {
  if (!c || !c->d_trace || !c->memo) return 0;
  if (c->resident) {  // persistent kernel: 128 words per CTA, one CTA per SM at most
    const size_t n = std::min<size_t>(std::min<size_t>(static_cast<size_t>(c->n_tiles) * c->m_groups, c->sm_count),
                                      cap_words / 128);
    if (cudaMemcpy(dst, c->d_trace, n * 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return static_cast<long>(n);
  }
  const size_t ctas = std::min<size_t>(static_cast<size_t>(c->n_tiles) * c->grid_groups, cap_words / kTraceSlots);
  if (cudaMemcpy(dst, c->d_trace, ctas * kTraceSlots * sizeof(unsigned long long), cudaMemcpyDeviceToHost) !=
      cudaSuccess)
    return -1;
  return static_cast<long>(ctas);
}

}  // namespace spxb
