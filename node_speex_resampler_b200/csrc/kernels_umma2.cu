// kernels_umma2.cu -- persistent tensor-core FIR for sm_100a: packed tap tiles RESIDENT in shared
// memory (the second tensor kernel; kernels_umma.cu, one tile per CTA with streamed tap tiles,
// stays as the path for filters whose packed tile does not fit).
//
// Same arithmetic as kernels_umma.cu -- the whole hot path of speex_resampler_process_interleaved_int
// (deps/speex/resample.c:1061-1082 over :968-1036 and the four resampler_basic_* kernels :331-558)
// as an EXACT integer banded GEMM on tcgen05.mma.kind::i8, one rounding at the end (WORD2INT,
// arch.h:208-209) -- but organised around what the round-1 profile showed to bind the long-filter
// shapes: bytes from L2 into the SM (tap tiles re-streamed by every CTA, 57 % of the traffic) and
// per-tile fixed cost (prologue + epilogue = 30 % of a CTA's life).
//   * One CTA per SM, persistent: CTA b walks the contiguous share [b*W/grid, (b+1)*W/grid) of the
//     tile list ordered tile-index-major (w = t * groups + g), so consecutive tiles of a CTA share
//     their output tile index t and with it the tap tile. Barriers, TMEM and the instruction cache
//     are set up once per CTA instead of once per tile.
//   * The tap tile of t stays in shared memory for the whole run of tiles that share it and is
//     loaded once per run (one bulk copy per K stage, each with its own mbarrier, so the first
//     tile's MMAs start as the stages arrive). PCM is then the only per-tile stream into the SM.
//   * The tile is PACKED (umma_plan.h): per K step only the 16-column blocks of each tap digit
//     that can be non-zero are stored -- the band's corners, the high digit outside the main lobe
//     and the middle digit near the filter's ends are skipped -- which cuts the tap bytes to
//     ~0.6 and the MMA columns with them. The MMA warp walks a per-K-step table of at most three
//     entries (B rows, D columns) carried in the kernel parameters.
// Warp roles (352 threads):
//   0-7  converters + epilogue: PCM (history for frames < 0, the call's input after) -> byte planes
//        in UMMA layout through a ring of 64-frame stages; two stages of loads in flight in
//        registers, also ACROSS tile boundaries (the next tile's first loads fly during the
//        epilogue); epilogue straight from TMEM to the interleaved int16 output;
//   8    lane 0 owns the mbarriers and loads the tap tile of each run;
//   9    owns TMEM; one elected lane issues the MMAs, releases PCM stages with tcgen05.commit,
//        hands the accumulator to the epilogue (acc_full) and takes it back (acc_empty);
//   10   slides the history (resample.c:898-899) of this CTA's share of streams beside the FIR and
//        publishes the new stream position.
// One instantiation per (CH, FAST, IDS); the schedule is written down, not induced by probes.
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels_common.cuh"
#include "launch.h"
#include "umma_common.cuh"
#include "umma_context.h"
#include "umma_plan.h"
#include "umma_ptx.cuh"

namespace spxb {

namespace {

using namespace ptx;
using namespace ummac;

constexpr int kMaxXStages = 6;
// 16 warps (128 registers per thread), every role on its own warps:
//   0-7   converters, two groups of four on alternate stages;
//   8-11  epilogue, one warp per TMEM lane quarter (warp % 4 selects the quarter);
//   12    MMA issue (owns TMEM and the barriers, loads the tap tile of each run);
//   13    warms the constant cache with the MMA records, then idles;
//   14-15 history slide.
// (Sixteen converter warps, two per stage share, were tried: at 96 registers they spill, and with
// the shared memory this kernel takes L1 is ~25 KB, so spills go to L2.)
constexpr int kConvWarps2 = 8, kGroupWarps = 4;
constexpr int kEpiWarp0 = 8, kEpiWarps = 4;
constexpr int kMmaWarp2 = 12, kSpareWarp = 13, kHistWarp0 = 14, kHistWarps = 2;
constexpr int kThreads2 = 16 * 32;
constexpr uint32_t kMaxTapStages = kUmmaMaxKsteps / 2;
constexpr uint32_t kInlineTiles2 = 32;
// dynamic shared memory this kernel may ask for: 227 KB minus its static part (barriers, 1 KB with
// the alignment of the dynamic array)
constexpr uint32_t kMaxSmem2 = 227u * 1024u - 5120u;

// What the kernel needs of the packed plan, ready to use (built on the host, carried in the kernel
// parameters): per 64-frame stage the tap bytes to load and its run of MMA records; per record
// everything one hi/lo pair of MMAs needs except the ring slot -- the issuing lane adds three bases
// and goes (a table of raw block ranges cost ~190 cycles of descriptor building per MMA in that
// one lane).
struct MmaRec {
  uint32_t b;         // added to the B descriptor: (K step offset + first row) | rows per chunk << 16
  uint32_t idesc_hi;  // instruction descriptor of the hi-plane MMA (the lo plane clears the A sign bit)
  uint32_t d_a;       // first accumulator column (hi plane) | A offset of the K step inside the stage << 16
  uint32_t pad_;
};
constexpr uint32_t kMaxRecs = kUmmaMaxKsteps * kUmmaMaxEntries;

struct InlineTile2 {
  int32_t kf0;
  uint32_t slot;
};

struct Umma2Args {
  const UmmaTile *tiles;  // tile table in HBM (used when n_inline == 0)
  uint32_t n_tiles;       // output tiles T
  uint32_t n_inline;
  uint32_t n_groups;      // series groups G
  uint32_t n_work;        // T * G
  const int8_t *pool;
  uint32_t tile_bytes;
  uint32_t nt;
  uint32_t ksteps;
  uint32_t x_stages;
  uint32_t tmem_cols;
  uint32_t dense;         // every K step is one MMA pair over all 3 nt columns (no packing, 3 nt <= 256): the issue loop needs no records
  uint32_t n_acc;         // accumulator sets in TMEM (2 when 8 nt <= 512: the epilogue of a tile runs under the next tile's MMAs)
  int shift;
  unsigned long long *trace;
  InlineTile2 inl[kInlineTiles2];
  uint32_t n_rec;
  uint32_t stage_off[kMaxTapStages + 1];  // byte offset of each 64-frame stage inside the packed tile
  uint16_t stage_rec[kMaxTapStages + 2];  // first MMA record of each stage (K step 0 has its own code)
  uint16_t stage_mid[kMaxTapStages + 2];  // first record of the stage's second K step
  // The MMA records live in the kernel parameters (constant bank): the issuing lane reads them with
  // uniform loads straight into uniform registers. (From shared memory every operand of a
  // tcgen05.mma went through a register-to-uniform move: ~150 cycles of issue per MMA, three times
  // what the tensor pipe needs for it.) A spare warp touches them during the prologue so that the
  // first tile does not pay a cold constant-cache miss per line.
  MmaRec rec[kMaxRecs + 1];
};

// PCM loads of the converters. With the tap tile resident the CTA's shared memory leaves the SM's
// unified L1 only ~25 KB, less than the two stages of loads (32 KB) the converters keep in flight:
// SPXB_LD_NOALLOC=1 asks for loads that do not allocate L1 lines.
#ifndef SPXB_LD_NOALLOC
#define SPXB_LD_NOALLOC 0
#endif
__device__ __forceinline__ uint4 ld_pcm16(const void *p) {
#if defined(SPXB_DBG_NOLOAD)
  return make_uint4(0u, 0u, 0u, reinterpret_cast<uintptr_t>(p) == 1 ? 1u : 0u);  // timing experiment: no PCM loads
#elif SPXB_LD_NOALLOC == 2
  return *reinterpret_cast<const uint4 *>(p);  // plain ld.global (LSU path instead of the read-only path)
#elif SPXB_LD_NOALLOC
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
#else
  return __ldg(reinterpret_cast<const uint4 *>(p));
#endif
}

#ifdef SPXB_UMMA2_TRACE
constexpr int kTraceSlots2 = 128;
#define TRACE2(u, slot)                                                                                       \
  do {                                                                                                        \
    if ((u).trace) (u).trace[static_cast<size_t>(blockIdx.x) * kTraceSlots2 + (slot)] = static_cast<unsigned long long>(clock64()); \
  } while (0)
#else
#define TRACE2(u, slot) \
  do {                  \
  } while (0)
#endif

// History slide (resample.c:898-899) of streams first, first + step, ... (n_mine of them), four at
// a time: four streams x four vectors of type V per lane are loaded before any store.
template <typename V, bool IDS>
__device__ __forceinline__ void slide_streams(const CallArgs &a, uint32_t step, uint32_t first, uint32_t n_mine,
                                              uint32_t hist_elems, size_t shift, int lane, uint32_t part,
                                              uint32_t parts) {
  constexpr uint32_t VW = sizeof(V) / 2;  // int16 elements per vector
  for (uint32_t k0 = 4 * part; k0 < n_mine; k0 += 4 * parts) {
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

// FAST = every input and output row starts on a 16-byte boundary (every BASELINE shape): drops the
// narrower load / store variants. IDS = the launch covers the stream subset a.ids[0 .. a.n_ids)
// (a cohort of a ragged batch).
//
// Converter schedule. One converter warp's share of a stage is a serial chain -- wait for the ring
// slot, split the bytes, fence, arrive, address and issue the next loads -- of 650-800 cycles
// however little data it covers (measured with loads and MMAs compiled out), and global loads come
// back after ~2000 cycles. So the 8 converter warps work as TWO groups of four on ALTERNATE stages
// (group 0 the even stages of the CTA's stage sequence, group 1 the odd ones), each thread taking
// 8 items of its group's stage: two stages are in the making at any time, the chain is paid once
// per two stages' worth of data, and with two of its own stages of loads in registers per warp
// 64 KB are in flight per SM. The converters never look at tile boundaries: the CTA's tiles are one
// sequence of stages to them.
//
// Accumulators. With 8 nt <= 512 TMEM holds two accumulator sets: the MMA warp moves on to the next
// tile the moment the last MMA of a tile is issued, and the four epilogue warps read the finished
// set out underneath. With one set (wider tiles) the MMA warp waits for the epilogue's last
// tcgen05.ld before the next tile's first MMA.
template <int CH, bool FAST, bool IDS>
__global__ void __launch_bounds__(kThreads2, 1)
    umma2_fir_kernel(const __grid_constant__ CallArgs a, const __grid_constant__ Umma2Args u) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t x_full[kMaxXStages], x_empty[kMaxXStages], tap_full[kMaxTapStages];
  __shared__ uint64_t taps_free, acc_full[2], acc_empty[2], tmem_ready;
  __shared__ uint32_t tmem_slot;

  constexpr int kStreams = kUmmaRows / CH;  // streams per series group
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const StreamCall sc = a.uniform;
  const uint32_t nt = u.nt, G = u.n_groups, T = u.n_tiles;
  // this CTA's share of the tile list (tile index major: w = t * G + g)
  const uint32_t w_begin = static_cast<uint32_t>(static_cast<unsigned long long>(blockIdx.x) * u.n_work / gridDim.x);
  const uint32_t w_end = static_cast<uint32_t>(static_cast<unsigned long long>(blockIdx.x + 1) * u.n_work / gridDim.x);
  const uint32_t n_tiles_mine = w_end - w_begin;
  constexpr uint32_t kXPlaneBytes = x_plane(CH), kXStageBytes = x_stage(CH);
  const uint32_t S = u.x_stages;
  const uint32_t n_iters = (u.ksteps + 1) / 2;  // 64-frame stages per tile
  uint8_t *const tap_smem = smem + S * kXStageBytes;
  const uint32_t n_rows = IDS ? a.n_ids : a.n_streams;  // streams this launch covers
  auto stream_of = [&](uint32_t i) -> size_t { return IDS ? a.ids[i] : i; };
  auto tile_kf0 = [&](uint32_t t) -> int { return u.n_inline ? u.inl[t].kf0 : u.tiles[t].kf0; };
  auto tile_slot = [&](uint32_t t) -> uint32_t { return u.n_inline ? u.inl[t].slot : u.tiles[t].slot; };
  // alignment every input row start shares (16-byte items start at multiples of 16 B in a row)
  const uint32_t row_bits = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(a.in)) |
                            (static_cast<uint32_t>(a.in_stride) * 2u);
  const int in_align = FAST ? 16 : (row_bits & 15u) == 0 ? 16 : (row_bits & 7u) == 0 ? 8 : (row_bits & 3u) == 0 ? 4 : 2;

  // Programmatic dependent launch: the next call's grid may be scheduled as SMs drain; it blocks in
  // griddepcontrol.wait below until this grid has completed, before it touches PCM, history or output.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) TRACE2(u, 0);
  if (tid == kThreads2 - 64) {
#ifdef SPXB_UMMA2_TRACE
    if (u.trace) {
      unsigned long long gt;
      uint32_t smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      u.trace[static_cast<size_t>(blockIdx.x) * kTraceSlots2 + 14] = gt;
      u.trace[static_cast<size_t>(blockIdx.x) * kTraceSlots2 + 13] = smid;
    }
#endif
  }

  // ---- prologue: the barriers, one per lane of the MMA warp; then everybody may proceed ----
  if (warp == kMmaWarp2) {
    const uint32_t n_bar = 2 * S + n_iters + 6;
    for (uint32_t i = lane; i < n_bar; i += 32) {
      if (i < S) mbar_init(&x_full[i], kGroupWarps);             // one elected arrival per warp of a group
      else if (i < 2 * S) mbar_init(&x_empty[i - S], 1);
      else if (i < 2 * S + n_iters) mbar_init(&tap_full[i - 2 * S], 1);
      else if (i == 2 * S + n_iters) mbar_init(&taps_free, 1);
      else if (i <= 2 * S + n_iters + 2) mbar_init(&acc_full[i - (2 * S + n_iters + 1)], 1);
      else if (i <= 2 * S + n_iters + 4) mbar_init(&acc_empty[i - (2 * S + n_iters + 3)], kEpiWarps);
      else mbar_init(&tmem_ready, 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // One bulk copy per K stage of the packed tile, each completing its own barrier: lane `it` of the
  // MMA warp issues stage `it` (offsets and sizes come with the kernel parameters).
  auto load_tap_tile = [&](uint32_t t) {
    const int8_t *src = u.pool + static_cast<size_t>(tile_slot(t)) * u.tile_bytes;
    if (static_cast<uint32_t>(lane) < n_iters) {
      const uint32_t off = u.stage_off[lane], bytes = u.stage_off[lane + 1] - off;
      mbar_arrive_expect_tx(&tap_full[lane], bytes);
      bulk_g2s(tap_smem + off, src + off, bytes, &tap_full[lane]);
    }
  };

  if (warp < kConvWarps2) {
    // ================= converters: PCM -> byte planes in UMMA layout =================
    // A stage is 64 frames of 128 series = kStreams stream segments of 64*CH*2 bytes. Lanes of a warp
    // walk ALONG a segment in 16-byte items (PPS items per stream, SPI streams per warp instruction),
    // so one LDG.128 covers four full 128-byte lines; each thread owns kItems items per stage, item i
    // of warp gw (of its group) belonging to stream (8 gw + i) * SPI + lane / PPS of the tile's group.
    constexpr int FPI = 8 / CH;          // frames per 16-byte item
    constexpr int PPS = 64 / FPI;        // items per stream per stage (16 stereo, 8 mono)
    constexpr int SPI = 32 / PPS;        // streams per warp instruction (2 stereo, 4 mono)
    constexpr int kItems = 8;
    constexpr int kStageFrames = kStageChunks * kUmmaChunkFrames;  // 64
    const uint32_t group = static_cast<uint32_t>(warp) >> 2, gw = warp & 3;
    const int conv_p = lane % PPS;       // item position inside the stage segment
    // byte offset of item i's hi/left word(s) inside an X stage: conv_off0 + i * kItemStride
    // (item i belongs to stream sl = (8 gw + i) * SPI + lane / PPS; stereo: left channel of stream sl
    // -> row 2 sl, right channel -> row 2 sl + 1)
    constexpr uint32_t kItemStride = SPI * 16 * CH;
    const uint32_t sl0 = static_cast<uint32_t>(8 * gw * SPI + lane / PPS);
    const uint32_t conv_off0 = [&] {
      const uint32_t j = static_cast<uint32_t>(conv_p * FPI) / kUmmaChunkFrames;
      const uint32_t byte_in_row = static_cast<uint32_t>(conv_p * FPI) % kUmmaChunkFrames;
      return (j >> 1) * x_kstep(CH) + (j & 1) * x_lbo(CH) + sl0 * (16 * CH) + byte_in_row;
    }();
    const uint32_t total_stages = n_tiles_mine * n_iters;

    // fetch cursor: (tile index, stage) whose loads go out next; advances two stages at a time
    uint32_t f_tile = 0, f_it = group;
    while (f_it >= n_iters && f_tile < n_tiles_mine) {
      f_it -= n_iters;
      ++f_tile;
    }
    int f_kf0 = 0;
    uint32_t f_sg0 = 0;  // first of this thread's streams in the fetch tile, as an index into the launch's rows
    auto setup_fetch = [&]() {
      const uint32_t w = w_begin + f_tile, t = w / G, g = w - t * G;
      f_kf0 = tile_kf0(t);
      f_sg0 = g * kStreams + sl0;
    };
    // row of the batch of item i (rows past the end of the batch read row 0; never stored)
    auto item_row = [&](int i) -> size_t {
      const uint32_t sg = f_sg0 + static_cast<uint32_t>(i * SPI);
      return stream_of(sg < n_rows ? sg : 0);
    };
    auto fetch = [&](uint4 (&raw)[kItems]) {
      if (f_tile >= n_tiles_mine) return;
      const int f = f_kf0 + static_cast<int>(f_it) * kStageFrames + conv_p * FPI;  // this thread's first frame
      const int rem = static_cast<int>(sc.n_in) - f;  // input frames left from f (when f >= 0)
      if (f < 0) {
#pragma unroll
        for (int i = 0; i < kItems; ++i)
          raw[i] = ld_pcm16(a.hist_src + item_row(i) * a.hist_stride + (static_cast<ptrdiff_t>(a.hist_frames) + f) * CH);
      } else if (rem >= FPI && in_align == 16) {
#pragma unroll
        for (int i = 0; i < kItems; ++i)
          raw[i] = ld_pcm16(a.in + item_row(i) * a.in_stride + static_cast<ptrdiff_t>(f) * CH);
      } else if (rem <= 0) {
#pragma unroll
        for (int i = 0; i < kItems; ++i) raw[i] = make_uint4(0u, 0u, 0u, 0u);
      } else {
        // the item holding the end of the input, or rows less than 16-byte aligned
#pragma unroll
        for (int i = 0; i < kItems; ++i)
          raw[i] = fetch_item_any(a.in + item_row(i) * a.in_stride + static_cast<ptrdiff_t>(f) * CH, min(rem, FPI) * CH,
                                  in_align);
      }
      f_it += 2;
      if (f_it >= n_iters) {
        do {
          f_it -= n_iters;
          ++f_tile;
        } while (f_it >= n_iters);
        if (f_tile < n_tiles_mine) setup_fetch();
      }
    };
    uint4 raw0[kItems], raw1[kItems];  // two of this warp's stages of loads in flight
    // Everything above touched only kernel parameters and this CTA's own resources. The previous
    // call's grid (which reads the history buffer this call overwrites, and writes the one this call
    // reads) must have completed before any global access below.
    if (tid == 0) TRACE2(u, 3);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) TRACE2(u, 4);
    if (f_tile < n_tiles_mine) setup_fetch();

    auto convert_store = [&](uint8_t *xs, const uint4 (&raw)[kItems]) {
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        uint8_t *base = xs + conv_off0 + i * kItemStride;
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

    // ---- this warp's stages: q = group, group + 2, ... of the CTA's stage sequence ----
    uint32_t slot = group % S, par = 1u ^ ((group / S) & 1u);  // ring slot of stage q, parity its `empty` wait expects
#ifdef SPXB_UMMA2_TRACE
    uint32_t step_no = 0;
#define STEP_MARK(k) \
  if (tid == 0 && step_no == n_iters + 3) TRACE2(u, 56 + (k))
#else
#define STEP_MARK(k)
#endif
    auto convert_step = [&](const uint4 (&raw)[kItems]) {
      STEP_MARK(0);
      mbar_wait(&x_empty[slot], par);
      STEP_MARK(1);
      convert_store(smem + slot * kXStageBytes, raw);
      STEP_MARK(2);
      // every thread makes its own stores visible to the async proxy, then one lane per warp arrives
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_full[slot]);
      STEP_MARK(3);
      slot += 2;
      if (slot >= S) {
        slot -= S;
        par ^= 1u;
      }
    };
    // The two register sets take turns strictly (never moved: a move of a register that a load in
    // flight will write waits for that load).
    fetch(raw0);
    fetch(raw1);
    if (tid == 0) TRACE2(u, 2);
    for (uint32_t q = group; q < total_stages; q += 4) {
      convert_step(raw0);
      fetch(raw0);
      STEP_MARK(4);
#ifdef SPXB_UMMA2_TRACE
      ++step_no;
#endif
      if (q + 2 >= total_stages) break;
      convert_step(raw1);
      fetch(raw1);
#ifdef SPXB_UMMA2_TRACE
      ++step_no;
#endif
    }
    if (tid == 0) TRACE2(u, 10);
  } else if (warp < kEpiWarp0 + kEpiWarps) {
    // ================= epilogue: straight from TMEM to the interleaved int16 output =================
    // A lane owns one series (TMEM lane) and 16 consecutive outputs per column group; mono packs them
    // into 32 contiguous bytes, stereo first swaps halves with the neighbouring lane (the other
    // channel of the same stream) so that each lane of the pair holds 8 whole frames = 32 bytes.
    // (Groups of 8 columns took 1.7x as long: an iteration costs ~600 cycles of TMEM-load and
    // dependent-arithmetic latency whatever its width.)
    const uint32_t out_bits = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(a.out)) |
                              (static_cast<uint32_t>(a.out_stride) * 2u);
    const int out_align = FAST ? 16 : (out_bits & 15u) == 0 ? 16 : (out_bits & 3u) == 0 ? 4 : 2;
    const uint32_t quarter = static_cast<uint32_t>(warp) & 3u;
    const uint32_t row = quarter * 32 + lane;  // TMEM lane = series of the tile
    const uint32_t sl_out = CH == 2 ? row >> 1 : row, ch_out = CH == 2 ? (row & 1u) : 0u;
    // the previous call's grid may still be writing the output rows this call overwrites
    asm volatile("griddepcontrol.wait;" ::: "memory");
    mbar_wait(&tmem_ready, 0);
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = tmem + ((quarter * 32u) << 16);
    uint32_t buf = 0, use_par = 0;  // accumulator set of the tile, parity of its `full` barrier
    for (uint32_t tile_no = 0; tile_no < n_tiles_mine; ++tile_no) {
      const uint32_t w = w_begin + tile_no, t = w / G, g = w - t * G;
      const uint32_t m0 = t * nt;
      const uint32_t n_valid = min(nt, sc.n_out - m0);
      const uint32_t s_out = g * kStreams + sl_out;
      const bool live_out = s_out < n_rows;
      int16_t *out_row = a.out + stream_of(live_out ? s_out : 0) * a.out_stride + static_cast<size_t>(m0) * CH;
      if (lane == 0 && quarter == 0) TRACE2(u, 16 + 8 * min(tile_no, 4u) + 1);
      mbar_wait(&acc_full[buf], use_par);
      tc_fence_after_sync();
      const uint32_t lane_addr = lane_base + buf * 4u * nt;
      const uint32_t cg_end = (n_valid + 15) / 16;
      for (uint32_t cg = 0; cg < cg_end; ++cg) {
        uint32_t p0[16], p1[16], p2[16], p3[16];
        tmem_ld16(lane_addr + cg * 16, p0);
        tmem_ld16(lane_addr + nt + cg * 16, p1);
        tmem_ld16(lane_addr + 2 * nt + cg * 16, p2);
        tmem_ld16(lane_addr + 3 * nt + cg * 16, p3);
        tmem_ld_wait();
        if (cg + 1 == cg_end) {
          // this warp's last read of the accumulator set: hand it back before the arithmetic and
          // the stores of this group
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        int r16[16];
        combine16(p0, p1, p2, p3, u.shift, r16);  // rounded, not yet saturated
        uint32_t wv[8];       // this lane's 16 int16 values = 32 contiguous output bytes
        uint32_t first_elem;  // their position in the stream's row, in int16 elements from m0
        if (CH == 2) {
          // lane pair (left, right): left keeps frames [0,8), right keeps frames [8,16)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int send = ch_out == 0 ? r16[8 + j] : r16[j];
            const int recv = __shfl_xor_sync(0xffffffffu, send, 1);
            wv[j] = ch_out == 0 ? pack_sat_s16x2(recv, r16[j]) : pack_sat_s16x2(r16[8 + j], recv);
          }
          first_elem = (cg * 16 + ch_out * 8) * 2;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) wv[j] = pack_sat_s16x2(r16[2 * j + 1], r16[2 * j]);
          first_elem = cg * 16;
        }
        if (!live_out) continue;
        const uint32_t total = n_valid * CH;
        const uint32_t n_here = first_elem >= total ? 0u : min(16u, total - first_elem);  // int16 elements
        int16_t *dst = out_row + first_elem;
        if (n_here == 16 && out_align == 16) {
          reinterpret_cast<uint4 *>(dst)[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          reinterpret_cast<uint4 *>(dst)[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
        } else if (out_align >= 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (2u * j + 1 < n_here) reinterpret_cast<uint32_t *>(dst)[j] = wv[j];
          if (n_here & 1u) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (2u * j + 1 == n_here) dst[2 * j] = static_cast<int16_t>(wv[j] & 0xffffu);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (2u * j < n_here) dst[2 * j] = static_cast<int16_t>(wv[j] & 0xffffu);
            if (2u * j + 1 < n_here) dst[2 * j + 1] = static_cast<int16_t>(wv[j] >> 16);
          }
        }
      }
      if (lane == 0 && quarter == 0) TRACE2(u, 16 + 8 * min(tile_no, 4u) + 3);
      if (++buf == u.n_acc) {
        buf = 0;
        use_par ^= 1u;
      }
    }
  } else if (warp == kMmaWarp2) {
    // ================= MMA issue =================
    // The whole warp walks the tiles and stages (uniform control flow, descriptors in uniform
    // registers); one elected lane issues. First: the tap tile of the first run (the tile table and
    // the tap pool are only ever rewritten by stream-ordered work, and a call that re-planned launches
    // without the programmatic edge: safe to read before the grid dependency resolves), then TMEM.
    load_tap_tile(w_begin / G);
    if (lane == 0) TRACE2(u, 12);
    tmem_alloc(&tmem_slot, u.tmem_cols);
    tmem_relinquish();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tmem_ready);
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t n3 = 3 * nt;
    const uint32_t np0 = min(n3, 256u), np1 = n3 - np0;            // [0, 3nt)
    const uint32_t nq0 = min(2 * nt, 256u), nq1 = 2 * nt - nq0;    // [0, 2nt) (first K step, lo plane)
    const uint32_t id_hi = umma_idesc_i8(128, 0, true, true), id_lo = umma_idesc_i8(128, 0, false, true);
    auto with_n = [](uint32_t idesc, uint32_t n) { return idesc | ((n >> 3) << 17); };
    const uint64_t a_base = umma_smem_desc(smem_u32(smem), x_lbo(CH), 128);
    // B: the K halves of one MMA are rows*16 bytes apart (per K step), 8-row groups 128 bytes apart
    const uint64_t b_fixed = umma_smem_desc(smem_u32(tap_smem), 0, 128);
    constexpr uint32_t a_stage16 = kXStageBytes >> 4, a_lo16 = kXPlaneBytes >> 4;
    uint32_t slot = 0, par = 0, tile_no = 0, run = 0;
    uint32_t buf = 0, buf_use = 0;  // accumulator set of the tile, how many times it has been used before
    uint64_t a_st = a_base;
    uint32_t cur_t = 0xffffffffu;
    for (uint32_t w = w_begin; w < w_end; ++w, ++tile_no) {
      const uint32_t t = w / G;
      const bool new_run = t != cur_t;
      uint32_t tap_par = 0;
      if (new_run) {
        if (run) {
          // every MMA that reads the resident tile has completed (this warp's own commit at the end
          // of the previous run): load the tile of the new run
          mbar_wait(&taps_free, (run - 1) & 1u);
          load_tap_tile(t);
        }
        cur_t = t;
        tap_par = run & 1u;
        ++run;
      }
      const bool run_ends = w + 1 == w_end || (w + 1) / G != t;
      const uint32_t tr = 16 + 8 * min(tile_no, 4u);  // trace slots of this tile
      if (buf_use) {
        // the epilogue has read this set's previous tile out of TMEM
        mbar_wait(&acc_empty[buf], (buf_use - 1) & 1u);
        tc_fence_after_sync();
      }
      const uint32_t acc = tmem + buf * 4u * nt;
      if (lane == 0) TRACE2(u, tr + 4);
      for (uint32_t it = 0; it < n_iters; ++it) {
        if (tile_no == 1 && it < 16 && lane == 0) TRACE2(u, 64 + 3 * it);
        if (new_run) mbar_wait(&tap_full[it], tap_par);
        mbar_wait(&x_full[slot], par);
        tc_fence_after_sync();
        if (tile_no == 1 && it < 16 && lane == 0) TRACE2(u, 65 + 3 * it);
        if (it == 0 && lane == 0) TRACE2(u, tr + 5);
        const bool last = it + 1 == n_iters;
        if (elect_one()) {
          if (it == 0) {
            // K step 0 is stored whole and initialises every accumulator column
            const uint64_t b_k = b_fixed + (static_cast<uint64_t>(n3) << 16), a_lo = a_st + a_lo16;
            umma_i8(acc, a_st, b_k, with_n(id_hi, np0), 0u);
            if (np1) umma_i8(acc + 256, a_st, b_k + 256, with_n(id_hi, np1), 0u);
            // columns [nt,3nt) already hold hi*B: accumulate; columns [3nt,4nt) are fresh
            umma_i8(acc + nt, a_lo, b_k, with_n(id_lo, nq0), 1u);
            if (nq1) umma_i8(acc + nt + 256, a_lo, b_k + 256, with_n(id_lo, nq1), 1u);
            umma_i8(acc + 3 * nt, a_lo, b_k + 2 * nt, with_n(id_lo, nt), 0u);
          }
          if (u.dense) {
            // one MMA pair per K step, every operand a constant step from the previous one: nothing
            // but uniform adds between the MMAs (the record walk below costs ~100 cycles of dependent
            // uniform loads and arithmetic per MMA -- more than the tensor pipe needs to run one)
            const uint32_t id_h3 = with_n(id_hi, n3), id_l3 = with_n(id_lo, n3);
            const uint64_t b_row = b_fixed + (static_cast<uint64_t>(n3) << 16);
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
              const uint32_t k = 2 * it + h;
              if (k == 0 || k >= u.ksteps) continue;
              const uint64_t b = b_row + k * (2 * n3);
              const uint64_t a_hi = a_st + h * (x_kstep(CH) >> 4);
#ifndef SPXB_DBG_NOMMA
              umma_i8(acc, a_hi, b, id_h3, 1u);
              umma_i8(acc + nt, a_hi + a_lo16, b, id_l3, 1u);
#endif
            }
          } else {
            const uint32_t m_end = u.stage_rec[it + 1];
#pragma unroll 1
            for (uint32_t m = u.stage_rec[it]; m < m_end; ++m) {
              const uint32_t rb = u.rec[m].b, ri = u.rec[m].idesc_hi, rd = u.rec[m].d_a;
              const uint64_t b = b_fixed + rb;
              const uint64_t a_hi = a_st + (rd >> 16);
              const uint32_t d_hi = acc + (rd & 0xffffu);
#ifndef SPXB_DBG_NOMMA
              umma_i8(d_hi, a_hi, b, ri, 1u);
              umma_i8(d_hi + nt, a_hi + a_lo16, b, ri & ~(1u << 7), 1u);
#else
              if (b == 1 && a_hi == 2 && d_hi == 3) umma_i8(d_hi, a_hi, b, ri, 1u);  // timing experiment: no MMAs
#endif
            }
          }
          umma_commit(&x_empty[slot]);
          if (last) {
            umma_commit(&acc_full[buf]);
            if (run_ends) umma_commit(&taps_free);
          }
        }
        __syncwarp();
        if (tile_no == 1 && it < 16 && lane == 0) TRACE2(u, 66 + 3 * it);
        if (last && lane == 0) TRACE2(u, tr + 6);
        a_st += a_stage16;
        if (++slot == S) {
          slot = 0;
          par ^= 1u;
          a_st = a_base;
        }
      }
      if (++buf == u.n_acc) {
        buf = 0;
        ++buf_use;
      }
    }
    if (lane == 0) TRACE2(u, 11);
  } else if (warp == kSpareWarp) {
    // ================= spare warp: pull the MMA records into the constant cache =================
    uint32_t acc = 0;
    for (uint32_t i = lane * 4; i <= u.n_rec; i += 32 * 4) acc ^= u.rec[i].b;  // one read per 64-byte line
    if (acc == 0xdeadbeefu && u.n_rec == 0xffffffffu) a.samp_frac[0] = acc;     // (keeps the reads alive)
  } else {
    // ================= history slide (resample.c:898-899) and the new position =================
    // Runs beside the FIR: it reads the old history and this call's input, writes the other half
    // of the ping-pong. Stream sl of group g is handled with the tile (t, g) for which
    // t == sl % T; new history element e = element consumed*CH + e of (old history || input).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t hist_elems = a.hist_frames * CH;
    const size_t shift = static_cast<size_t>(sc.consumed) * CH;
    const int vshift = (shift % 8 == 0) ? 8 : (shift % 4 == 0) ? 4 : (shift % 2 == 0) ? 2 : 1;
    const int vw = min(vshift, in_align / 2);
    const uint32_t hw = static_cast<uint32_t>(warp - kHistWarp0);  // the history warps take turns at units of 4 streams
    for (uint32_t w = w_begin; w < w_end; ++w) {
      const uint32_t t = w / G, g = w - t * G;
      if (t >= static_cast<uint32_t>(kStreams)) continue;
      const uint32_t first = g * kStreams + t;
      const uint32_t in_group = (kStreams - t + T - 1) / T;
      const uint32_t in_batch = first < n_rows ? (n_rows - first + T - 1) / T : 0u;
      const uint32_t n_mine = min(in_group, in_batch);
      if (vw == 8) slide_streams<uint4, IDS>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      else if (vw == 4) slide_streams<uint2, IDS>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      else if (vw == 2) slide_streams<uint32_t, IDS>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      else slide_streams<uint16_t, IDS>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      for (uint32_t k = lane + 32 * hw; k < n_mine; k += 32 * kHistWarps) {
        const size_t s = stream_of(first + k * T);
        a.last_sample[s] = sc.ls1;
        a.samp_frac[s] = sc.frac1;
      }
    }
    if (lane == 0 && hw == 0) TRACE2(u, 8);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp2) tmem_dealloc(tmem_slot, u.tmem_cols);
  if (tid == 0) {
    TRACE2(u, 9);
#ifdef SPXB_UMMA2_TRACE
    if (u.trace) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      u.trace[static_cast<size_t>(blockIdx.x) * kTraceSlots2 + 15] = gt;
    }
#endif
  }
}

// One thread per 16-byte cell of the packed tile. jobs[i] = {slot, phase0, delta}; `plan` is the
// host's UmmaKStep table (umma_plan.h) in HBM.
__global__ void build_packed_tiles_kernel(const int32_t *__restrict__ h, uint32_t num, uint32_t den, uint32_t taps,
                                          uint32_t ksteps, const UmmaKStep *__restrict__ plan,
                                          const uint32_t *__restrict__ jobs, int8_t *pool, uint32_t tile_bytes) {
  const uint32_t slot = jobs[3 * blockIdx.y], phase0 = jobs[3 * blockIdx.y + 1], delta = jobs[3 * blockIdx.y + 2];
  const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= tile_bytes / 16) return;
  uint32_t k = 0;
  while (k + 1 < ksteps && plan[k + 1].off16 <= cell) ++k;
  const UmmaKStep ks = plan[k];
  uint32_t r = cell - ks.off16;
  const uint32_t half = r / ks.rows;
  r -= half * ks.rows;
  uint32_t digit = 0, n = 0;
  for (int i = 0; i < 3; ++i) {
    const uint32_t cnt = 16u * (ks.b1[i] - ks.b0[i]);
    if (r < cnt) {
      digit = i;
      n = 16u * ks.b0[i] + r;
      break;
    }
    r -= cnt;
  }
  const unsigned long long tt = static_cast<unsigned long long>(phase0) + static_cast<unsigned long long>(n) * num;
  const uint32_t phase = static_cast<uint32_t>(tt % den);
  const long long first = static_cast<long long>(delta) + static_cast<long long>(tt / den);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const long long j = static_cast<long long>(32 * k + 16 * half + e) - first;
    int d = 0;
    if (j >= 0 && j < static_cast<long long>(taps)) {
      const int32_t v = h[static_cast<size_t>(phase) * taps + static_cast<size_t>(j)];
      const int32_t d0 = ((v + 128) & 255) - 128;
      const int32_t r1 = (v - d0) >> 8;
      const int32_t d1 = ((r1 + 128) & 255) - 128;
      const int32_t d2 = (r1 - d1) >> 8;
      d = digit == 2 ? d0 : digit == 1 ? d1 : d2;  // digit index 0 = d2, 1 = d1, 2 = d0
    }
    w[e >> 2] |= (static_cast<uint32_t>(d) & 0xffu) << (8 * (e & 3));
  }
  *reinterpret_cast<uint4 *>(pool + static_cast<size_t>(slot) * tile_bytes + static_cast<size_t>(cell) * 16) =
      make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace

void umma2_configure_device() {
  auto big_smem = [](auto kernel) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem2); };
  big_smem(umma2_fir_kernel<1, false, false>);
  big_smem(umma2_fir_kernel<2, false, false>);
  big_smem(umma2_fir_kernel<1, true, false>);
  big_smem(umma2_fir_kernel<2, true, false>);
  big_smem(umma2_fir_kernel<1, false, true>);
  big_smem(umma2_fir_kernel<2, false, true>);
}

// the MMA records and per-stage tables the kernel reads (they travel in the kernel parameters)
cudaError_t umma2_upload_plan(UmmaContext *c, cudaStream_t) {
  const uint32_t n_iters = (c->ksteps + 1) / 2;
  const uint32_t a_ks16 = x_kstep(static_cast<int>(c->channels)) >> 4;
  c->recs.clear();
  c->stage_off.assign(kMaxTapStages + 1, 0u);
  c->stage_rec.assign(kMaxTapStages + 2, 0u);
  c->stage_mid.assign(kMaxTapStages + 2, 0u);
  for (uint32_t it = 0; it < n_iters; ++it) {
    c->stage_off[it] = c->packed.k[2 * it].off16 * 16u;
    c->stage_rec[it] = static_cast<uint16_t>(c->recs.size() / 4);
    for (uint32_t h = 0; h < 2; ++h) {
      const uint32_t k = 2 * it + h;
      if (h == 1) c->stage_mid[it] = static_cast<uint16_t>(c->recs.size() / 4);
      if (k >= c->ksteps) break;
      const UmmaKStep &ks = c->packed.k[k];
      if (k == 0) continue;  // K step 0 has its own code in the kernel
      for (uint32_t e = 0; e < ks.n_ent; ++e) {
        c->recs.push_back((ks.off16 + ks.ent[e].row) | (static_cast<uint32_t>(ks.rows) << 16));
        c->recs.push_back(umma_idesc_i8(128, ks.ent[e].n, true, true));
        c->recs.push_back(ks.ent[e].dcol | ((h ? a_ks16 : 0u) << 16));
        c->recs.push_back(0u);
      }
    }
  }
  for (uint32_t it = n_iters; it <= kMaxTapStages; ++it) c->stage_off[it] = c->packed.tile_bytes;
  for (uint32_t it = n_iters; it < kMaxTapStages + 2; ++it) c->stage_rec[it] = static_cast<uint16_t>(c->recs.size() / 4);
  return c->recs.size() / 4 <= kMaxRecs ? cudaSuccess : cudaErrorInvalidValue;
}

// x stages the resident kernel would run for this geometry (0: the packed tile does not fit)
uint32_t umma2_x_stages(uint32_t channels, uint32_t tile_bytes, uint32_t ksteps) {
  const uint32_t xs = x_stage(static_cast<int>(channels));
  if (ksteps > kUmmaMaxKsteps || tile_bytes + 2 * xs > kMaxSmem2) return 0;
  const uint32_t n_iters = (ksteps + 1) / 2;
  uint32_t stages = std::min<uint32_t>(kMaxXStages, (kMaxSmem2 - tile_bytes) / xs);
  static const uint32_t cap = [] {
    const char *e = getenv("SPXB_UMMA2_XSTAGES");
    return e ? static_cast<uint32_t>(atoi(e)) : 0u;
  }();
  if (cap >= 2) stages = std::min(stages, cap);
  return std::max(1u, std::min(stages, std::max(2u, n_iters)));
}

cudaError_t umma2_build_tiles(UmmaContext *c, const uint32_t *d_jobs, size_t n_jobs, cudaStream_t stream) {
  const uint32_t cells = c->tile_bytes / 16;
  const dim3 grid((cells + 255) / 256, static_cast<unsigned>(n_jobs));
  build_packed_tiles_kernel<<<grid, 256, 0, stream>>>(c->d_h, c->spec.num, c->spec.den, c->spec.taps, c->ksteps,
                                                      c->d_kplan, d_jobs, c->d_pool, c->tile_bytes);
  return cudaGetLastError();
}

cudaError_t launch_umma2(UmmaContext *c, const CallArgs &a, cudaStream_t stream, uint32_t *launches) {
  Umma2Args u;
  std::memset(&u, 0, sizeof(u));
  u.tiles = c->d_tiles;
  u.n_tiles = c->n_tiles;
  u.n_inline = c->n_tiles <= kInlineTiles2 ? c->n_tiles : 0u;
  for (uint32_t i = 0; i < u.n_inline; ++i) {
    u.inl[i].kf0 = c->h_tiles[i].kf0;
    u.inl[i].slot = c->h_tiles[i].slot;
  }
  u.n_groups = c->m_groups;
  const uint64_t work = static_cast<uint64_t>(c->m_groups) * c->n_tiles;
  if (work == 0 || work > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  u.n_work = static_cast<uint32_t>(work);
  u.pool = c->d_pool;
  u.tile_bytes = c->tile_bytes;
  u.nt = c->nt;
  u.ksteps = c->ksteps;
  u.x_stages = c->stages;
  u.tmem_cols = c->tmem_cols;
  u.n_acc = c->n_acc;
  u.dense = 3 * c->nt <= 256 ? 1u : 0u;
  for (uint32_t k = 0; k < c->ksteps && u.dense; ++k) {
    const UmmaKStep &ks = c->packed.k[k];
    if (ks.n_ent != 1 || ks.ent[0].row != 0 || ks.ent[0].dcol != 0 || ks.ent[0].n != 3 * c->nt || ks.off16 != k * 6 * c->nt)
      u.dense = 0u;
  }
  u.shift = c->ft.shift;
  u.n_rec = static_cast<uint32_t>(c->recs.size() / 4);
  std::memcpy(u.rec, c->recs.data(), c->recs.size() * sizeof(uint32_t));
  for (uint32_t i = 0; i <= kMaxTapStages; ++i) u.stage_off[i] = c->stage_off[i];
  for (uint32_t i = 0; i < kMaxTapStages + 2; ++i) u.stage_rec[i] = c->stage_rec[i];
  for (uint32_t i = 0; i < kMaxTapStages + 2; ++i) u.stage_mid[i] = c->stage_mid[i];
  u.trace = nullptr;
#ifdef SPXB_UMMA2_TRACE
  static const bool want_trace = getenv("SPXB_UMMA_TRACE") != nullptr;
  const uint32_t grid_for_trace = static_cast<uint32_t>(std::min<uint64_t>(work, static_cast<uint64_t>(c->sm_count)));
  if (want_trace) {
    if (c->trace_ctas < grid_for_trace) {
      if (c->d_trace) cudaFree(c->d_trace);
      c->d_trace = nullptr;
      if (cudaMalloc(reinterpret_cast<void **>(&c->d_trace), grid_for_trace * kTraceSlots2 * sizeof(unsigned long long)) ==
          cudaSuccess)
        c->trace_ctas = grid_for_trace;
    }
    u.trace = c->d_trace;
  }
#endif
  static const bool use_pdl = [] {
    const char *e = getenv("SPXB_UMMA_PDL");
    return !e || atoi(e) != 0;
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(std::min<uint64_t>(work, static_cast<uint64_t>(c->sm_count))));
  cfg.blockDim = dim3(kThreads2);
  cfg.dynamicSmemBytes = c->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  unsigned n_attr = 0;
  if (use_pdl && !c->fresh_plan) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  const uintptr_t row_bits = reinterpret_cast<uintptr_t>(a.in) | (a.in_stride * 2u) |
                             reinterpret_cast<uintptr_t>(a.out) | (a.out_stride * 2u);
  const bool fast = (row_bits & 15u) == 0;
  auto launch = [&](auto kernel) { return cudaLaunchKernelEx(&cfg, kernel, a, u); };
  cudaError_t e;
  if (a.ids) {
    e = a.channels == 2 ? launch(umma2_fir_kernel<2, false, true>) : launch(umma2_fir_kernel<1, false, true>);
  } else if (a.channels == 2) {
    e = fast ? launch(umma2_fir_kernel<2, true, false>) : launch(umma2_fir_kernel<2, false, false>);
  } else {
    e = fast ? launch(umma2_fir_kernel<1, true, false>) : launch(umma2_fir_kernel<1, false, false>);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) c->fresh_plan = false;
  if (e == cudaSuccess && launches) *launches += 1;
  return e;
}

}  // namespace spxb
