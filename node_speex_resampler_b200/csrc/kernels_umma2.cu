// kernels_umma2.cu -- persistent, TMA-fed tensor-core FIR kernels for sm_100a: what long filters run
// (kernels_umma.cu, one tile per CTA with streamed tap tiles and LDG-fed converters, stays as the
// path for short filters, ragged cohorts, input rows off 16-byte boundaries and tap tiles that do
// not fit). Two kernels live here:
//   umma3_fir_kernel  (default)  the byte planes go to TENSOR MEMORY as the MMA's A operand;
//   umma2_fir_kernel  (SPXB_UMMA_ATMEM=0)  the byte planes are converted in place in the ring slot in
//                     shared memory; with two accumulator sets for tiles of <= 64 outputs.
//
// Same arithmetic as kernels_umma.cu -- the whole hot path of speex_resampler_process_interleaved_int
// (deps/speex/resample.c:1061-1082 over :968-1036 and the four resampler_basic_* kernels :331-558)
// as an EXACT integer banded GEMM on tcgen05.mma.kind::i8, one rounding at the end (WORD2INT,
// arch.h:208-209) -- organised around what the measurements of round 2 showed to bound a 64-frame
// stage of the long-filter shapes:
//   * PCM reaches the SM through the copy engine. LDG.128 into registers delivers 13-14 B/clk/SM
//     from 8 warps and 22 from 16, whatever is in flight and whatever L1 is left (csrc/ldg_rate.cu);
//     that alone is ~1200 cycles per 16 KB stage -- the stage time of every earlier version.
//     cp.async 16 B does 25.6, 1-D bulk copies cost ~61 cycles of issue EACH (4 B/clk for 256-byte
//     row pieces), one tensor-map box of 64 rows x 256 B does 42 B/clk/SM = 387 cycles per stage.
//     A loader lane issues the boxes of a stage into a ring slot. No thread computes a PCM address,
//     no registers hold loads in flight, the ragged end of the input and the rows past the end of
//     the batch are the tensor map's zero fill.
//   * The raw bytes become the A operand without a plane ring in shared memory (umma3): a converter
//     thread owns one series = one TMEM lane, splits its row's bytes in registers and writes them with
//     tcgen05.st behind the accumulator columns. The raw slot is free as soon as it has been read, the
//     tensor core stops re-reading the planes from shared memory, and the shared memory the plane ring
//     took is ring depth for the boxes. (umma2: the group splits the bytes in place in the slot,
//     behind a group barrier; the slot is held until the MMAs that read it have completed.)
//   * One CTA per SM, persistent: CTA b walks the contiguous share [b*W/grid, (b+1)*W/grid) of the
//     tile list ordered tile-index-major (w = t * groups + g), so consecutive tiles of a CTA share
//     their output tile index t and with it the tap tile. Barriers, TMEM and the instruction cache
//     are set up once per CTA instead of once per tile.
//   * The tap tile of t stays in shared memory for the whole run of tiles that share it (one bulk
//     copy per K stage, each with its own mbarrier, so the first tile's MMAs start as the stages
//     arrive): every CTA re-streaming its tap tile was 57 % of the L2 -> SM traffic.
//   * The tile is PACKED (umma_plan.h): per K step only the 16-column blocks of each tap digit
//     that can be non-zero are stored, ~0.6 of the dense bytes. The MMA lane walks a per-K-step
//     table of at most three records carried in the kernel parameters, one record ahead of the
//     MMAs it issues.
//   * Every mbarrier has exactly ONE waiting party (a parity wait tells a phase only from the one
//     before it), and a slot is handed back only after the loads from it have completed -- both
//     learnt from launches that hung or were wrong in one row out of thousands.
// Warp roles (384 threads; 512 in umma2's instantiation with two accumulator sets):
//   0-7   converters, two groups of four on alternate stages; they also run the epilogue (straight
//         from TMEM to the interleaved int16 output) when there is one accumulator set;
//   8     owns TMEM and the barriers, loads the tap tile of each run; one elected lane issues the
//         MMAs, releases slots with tcgen05.commit, hands the accumulator over (acc_full) and takes
//         it back (acc_empty);
//   9     lane 0: the TMA boxes;
//   10-11 slide the history (resample.c:898-899) of this CTA's share of streams beside the FIR and
//         publish the new stream position;
//   12-15 (umma2, 8 nt <= 512: two accumulator sets) the epilogue of a tile under the next tile's MMAs.
#include <cuda.h>

#include <algorithm>
#include <cstddef>
#include <chrono>
#include <cstdio>
#include <thread>
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
// 12 warps (168 registers per thread):
//   0-7   converters + epilogue, two groups of four on alternate stages (warp % 4 = TMEM lane quarter);
//   8     MMA issue (owns TMEM and the barriers, loads the tap tile of each run);
//   9     lane 0 feeds the ring: one tensor-map TMA box of raw PCM per stage;
//   10-11 history slide.
constexpr int kConvWarps2 = 8, kGroupWarps = 4;
constexpr int kMmaWarp2 = 8, kLoadWarp = 9, kHistWarp0 = 10, kHistWarps = 2;
constexpr int kThreads2 = 12 * 32;
// DB instantiation (two accumulator sets, 8 nt <= 512): four more warps, 12-15, run the epilogue of a
// tile under the next tile's MMAs (warp % 4 = TMEM lane quarter); 128 registers per thread
constexpr int kEpiWarp0 = 12, kEpiWarps = 4;
constexpr int kThreadsDB = 16 * 32;
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
};
constexpr uint32_t kMaxRecs = kUmmaMaxKsteps * kUmmaMaxEntries;

struct InlineTile2 {
  int32_t kf0;
  uint32_t slot;
};

// Tensor maps over the call's PCM (bytes; row = stream): the input rows and the history rows, each
// with a box of one whole stage (64 frames x the tile's streams) and a box of a quarter stage (16
// frames) for the one stage of a tile that straddles the end of the history. Elements outside the
// tensor (frames past the call's input, streams past the end of the batch) arrive as zeros.
struct alignas(64) Umma2Maps {
  CUtensorMap in64, in16, hist64, hist16;
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
  int shift;
  unsigned long long *trace;
  InlineTile2 inl[kInlineTiles2];
  uint32_t n_rec;
  uint32_t stage_off[kMaxTapStages + 1];  // byte offset of each 64-frame stage inside the packed tile
  uint16_t stage_rec[kMaxTapStages + 2];  // first MMA record of each stage (K step 0 has its own code)
  // The MMA records live in the kernel parameters (constant bank): the issuing lane reads them with
  // uniform loads straight into uniform registers. (From shared memory every operand of a
  // tcgen05.mma went through a register-to-uniform move: ~150 cycles of issue per MMA, three times
  // what the tensor pipe needs for it.)
  MmaRec rec[kMaxRecs + 1];
};

#ifndef SPXB_PF_DIST
#define SPXB_PF_DIST 0
#endif
#ifndef SPXB_PREFETCH_MAPS
#define SPXB_PREFETCH_MAPS 0
#endif
#ifndef SPXB_DEFER_EPILOGUE
#define SPXB_DEFER_EPILOGUE 1
#endif

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

// SPXB_UMMA2_WATCHDOG: a barrier wait that gives up after ~a second, records who waited for what in
// a host-mapped buffer (u.trace) and carries on (the results are garbage; the point is the record)
#ifdef SPXB_UMMA2_WATCHDOG
__device__ __noinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity, uint32_t tag, unsigned long long *dbg,
                                          const volatile uint32_t *prog = nullptr) {
  for (uint32_t i = 0; i < (1u << 22); ++i)
    if (mbar_try_wait(bar, parity)) return;
  if (dbg) {
    const unsigned long long n = atomicAdd(dbg, 1ull);
    if (n < 500) {
      dbg[1 + n] = (static_cast<unsigned long long>(blockIdx.x) << 48) | (static_cast<unsigned long long>(threadIdx.x) << 32) | tag;
      __threadfence_system();
    }
    if (prog && n < 40) {
      for (int w = 0; w < 12; ++w) {
        const unsigned long long k = atomicAdd(dbg, 1ull);
        if (k < 500) dbg[1 + k] = (static_cast<unsigned long long>(blockIdx.x) << 48) | (static_cast<unsigned long long>(w * 32) << 32) | 0xE0000000ull | (prog[w] & 0x0fffffffu) | ((prog[w] >> 24) << 24 & 0x0f000000u);
      }
      __threadfence_system();
    }
  }
}
#define WAIT(bar, par, tag) mbar_wait_wd(bar, par, tag, u.trace)
#define WAIT_L(bar, par, tag) mbar_wait_wd(bar, par, tag, u.trace, wd_prog)
#else
#define WAIT(bar, par, tag) mbar_wait(bar, par)
#define WAIT_L(bar, par, tag) mbar_wait(bar, par)
#endif

__device__ __forceinline__ void tma_box_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
#ifndef SPXB_DBG_NOLOAD
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
#endif
}

// the same box, only as far as L2 (no shared memory, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int x, int y) {
#ifndef SPXB_DBG_NOLOAD
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
#endif
}

// named barrier of one converter group (ids 1, 2; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(uint32_t group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupWarps * 32) : "memory");
}

// History slide (resample.c:898-899) of streams first, first + step, ... (n_mine of them), four at
// a time: four streams x four vectors of type V per lane are loaded before any store.
template <typename V>
__device__ __forceinline__ void slide_streams(const CallArgs &a, uint32_t step, uint32_t first, uint32_t n_mine,
                                              uint32_t hist_elems, size_t shift, int lane, uint32_t part,
                                              uint32_t parts) {
  constexpr uint32_t VW = sizeof(V) / 2;  // int16 elements per vector
  for (uint32_t k0 = 4 * part; k0 < n_mine; k0 += 4 * parts) {
    for (uint32_t e0 = lane * VW; e0 < hist_elems; e0 += 32 * VW * 4) {
      V val[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const size_t s = first + static_cast<size_t>(k0 + kk) * step;
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
        const size_t s = first + static_cast<size_t>(k0 + kk) * step;
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

// The launch covers whole batches whose input, output and history rows start on 16-byte boundaries
// (every BASELINE shape); anything else -- cohorts of a ragged batch, odd row pitches -- runs on the
// one-tile-per-CTA kernel (kernels_umma.cu).
//
// A ring slot holds one 64-frame stage of the tile's 128 series, first as raw PCM (a TMA box: one row
// of 64 * CH * 2 bytes per stream), then, converted IN PLACE by a converter group, as the two byte
// planes in UMMA layout:
//   loader lane : x_empty[slot] -> TMA box(es) -> raw_full[slot] (transaction bytes)
//   group q & 1 : raw_full[slot] -> 8 items per thread into registers, bytes split -> group barrier
//                 (every thread has read its raw bytes) -> planes stored -> fence -> x_full[slot]
//   MMA lane    : x_full[slot] -> MMAs -> tcgen05.commit -> x_empty[slot]
// PCM therefore reaches the SM through the copy engine (42 B/clk/SM measured in this access pattern,
// csrc/ldg_rate.cu) instead of LDG.128 into registers (13-14 B/clk/SM from 8 warps, whatever is in
// flight -- what bounded the earlier kernels at ~1100 cycles per stage), no registers hold loads in
// flight, and the converters never compute a global address.
template <int CH, bool DB>
__global__ void __launch_bounds__(DB ? kThreadsDB : kThreads2, 1)
    umma2_fir_kernel(const __grid_constant__ CallArgs a, const __grid_constant__ Umma2Args u,
                     const __grid_constant__ Umma2Maps maps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // all barriers in one array, at fixed offsets, so that one convergent instruction initialises 32 of them
  constexpr uint32_t kBarTap = 3 * kMaxXStages, kBarMisc = kBarTap + kMaxTapStages, kBars = kBarMisc + 6;
  __shared__ uint64_t bars[kBars];
  uint64_t *const raw_full = bars, *const x_full = bars + kMaxXStages, *const x_empty = bars + 2 * kMaxXStages;
  uint64_t *const tap_full = bars + kBarTap;
  uint64_t &taps_free = bars[kBarMisc], &tmem_ready = bars[kBarMisc + 5];
  uint64_t *const acc_full = bars + kBarMisc + 1, *const acc_empty = bars + kBarMisc + 3;
  __shared__ uint32_t tmem_slot, raw_seen[kConvWarps2];
#ifdef SPXB_UMMA2_WATCHDOG
  __shared__ volatile uint32_t wd_prog[16];
#define PROG(code, v) \
  do {                \
    if ((threadIdx.x & 31) == 0) wd_prog[threadIdx.x >> 5] = ((code) << 24) | (v); \
  } while (0)
#else
#define PROG(code, v) \
  do {                \
  } while (0)
#endif

  constexpr int kStreams = kUmmaRows / CH;  // streams per series group
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const StreamCall sc = a.uniform;
  const uint32_t nt = u.nt, G = u.n_groups, T = u.n_tiles;
  // this CTA's share of the tile list (tile index major: w = t * G + g)
  const uint32_t w_begin = static_cast<uint32_t>(static_cast<unsigned long long>(blockIdx.x) * u.n_work / gridDim.x);
  const uint32_t w_end = static_cast<uint32_t>(static_cast<unsigned long long>(blockIdx.x + 1) * u.n_work / gridDim.x);
  const uint32_t n_tiles_mine = w_end - w_begin;
  constexpr uint32_t kXPlaneBytes = x_plane(CH), kXStageBytes = x_slot(CH);  // slot pitch of the ring
  constexpr int kStageFrames = kStageChunks * kUmmaChunkFrames;  // 64
  constexpr uint32_t kRowBytes = kStageFrames * CH * 2;          // raw bytes of one stream in a stage box
  constexpr uint32_t kRawBytes = kStreams * kRowBytes;           // 16 KB: the raw stage sits at the start of its slot
  constexpr uint32_t kQuarterRow = kRowBytes / 4, kQuarterBytes = kRawBytes / 4;
  static_assert(kRawBytes <= x_stage(CH) && kXStageBytes % 128 == 0, "a raw stage fits its slot; slots take TMA boxes");
  const uint32_t S = u.x_stages;
  const uint32_t n_iters = (u.ksteps + 1) / 2;  // 64-frame stages per tile
  uint8_t *const tap_smem = smem + S * kXStageBytes;
  const uint32_t n_rows = a.n_streams;
  auto tile_kf0 = [&](uint32_t t) -> int { return u.n_inline ? u.inl[t].kf0 : u.tiles[t].kf0; };
  auto tile_slot = [&](uint32_t t) -> uint32_t { return u.n_inline ? u.inl[t].slot : u.tiles[t].slot; };

  // Programmatic dependent launch: the next call's grid may be scheduled as SMs drain; it blocks in
  // griddepcontrol.wait below until this grid has completed, before it touches PCM, history or output.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) TRACE2(u, 0);
  if (tid == 11 * 32) {
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
    if (lane == 0) TRACE2(u, 5);
    // Lane i initialises barriers i and i + 32: two convergent mbarrier.init instructions, ~130 cycles
    // each (a chain of per-lane branches over the barrier kinds cost ~1700 cycles of every CTA's
    // prologue, one lane looping over them 60-90 cycles per barrier; csrc/mbar_probe.cu).
#pragma unroll
    for (uint32_t i = lane; i < kBars; i += 32) {
      const bool group_arrivals = i >= kMaxXStages && i < 2 * kMaxXStages;          // x_full
      const bool epi_arrivals = i == kBarMisc + 3 || i == kBarMisc + 4;             // acc_empty
      mbar_init(&bars[i], group_arrivals ? kGroupWarps : epi_arrivals ? (DB ? kEpiWarps : kConvWarps2) : 1u);
    }
    fence_mbar_init();
    if (lane == 0) TRACE2(u, 6);
  }
  __syncthreads();
  if (tid == 0) TRACE2(u, 1);

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

  // ---- epilogue of one tile: straight from TMEM to the interleaved int16 output ----
  // a lane owns one series (TMEM lane) and 16 consecutive outputs per column group; mono packs them
  // into 32 contiguous bytes, stereo first swaps halves with the neighbouring lane (the other
  // channel of the same stream) so that each lane of the pair holds 8 whole frames = 32 bytes.
  // Without the DB warps the two converter warps of a TMEM lane quarter (one of each group) share the
  // column groups of a tile; with them, one epilogue warp per quarter takes them all.
  const uint32_t gw_e = static_cast<uint32_t>(warp) & 3u;
  const uint32_t row = gw_e * 32 + lane;  // TMEM lane = series of the tile
  const uint32_t sl_out = CH == 2 ? row >> 1 : row, ch_out = CH == 2 ? (row & 1u) : 0u;
  uint32_t tmem = 0;
  // (cg0, cg_step): this warp's column groups; (buf, par): accumulator set of the tile and the parity
  // of its `full` phase
  auto epilogue = [&](uint32_t tile_no, uint32_t cg0, uint32_t cg_step, uint32_t buf, uint32_t par) {
    const uint32_t w = w_begin + tile_no, t = w / G, g = w - t * G;
    const uint32_t m0 = t * nt;
    const uint32_t n_valid = min(nt, sc.n_out - m0);
    const uint32_t s_out = g * kStreams + sl_out;
    const bool live_out = s_out < n_rows;
    int16_t *out_row = a.out + static_cast<size_t>(live_out ? s_out : 0) * a.out_stride + static_cast<size_t>(m0) * CH;
    if (tile_no == 0) {
      WAIT(&tmem_ready, 0, 0x01000000u);
      tc_fence_after_sync();
      tmem = tmem_slot;
    }
    PROG(5u, tile_no);
    if (tid == 0) TRACE2(u, 16 + 8 * min(tile_no, 4u) + 1);
    WAIT(&acc_full[buf], par, 0x02000000u | tile_no);
    tc_fence_after_sync();
    const uint32_t lane_addr = tmem + ((gw_e * 32u) << 16) + buf * 4u * nt;
    const uint32_t cg_end = (n_valid + 15) / 16;
    for (uint32_t cg = cg0; cg < cg_end; cg += cg_step) {
      uint32_t p0[16], p1[16], p2[16], p3[16];
      tmem_ld16(lane_addr + cg * 16, p0);
      tmem_ld16(lane_addr + nt + cg * 16, p1);
      tmem_ld16(lane_addr + 2 * nt + cg * 16, p2);
      tmem_ld16(lane_addr + 3 * nt + cg * 16, p3);
      tmem_ld_wait();
      if (cg + cg_step >= cg_end) {
        // this warp's last read of the accumulator: hand it back before the arithmetic and the
        // stores of this group, so the next tile's MMAs start underneath them
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
      if (n_here == 16) {
        reinterpret_cast<uint4 *>(dst)[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        reinterpret_cast<uint4 *>(dst)[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
      } else {
        // the ragged end of the call's output: rows are 16-byte aligned, so whole 32-bit words, then
        // possibly one last int16
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (2u * j + 1 < n_here) reinterpret_cast<uint32_t *>(dst)[j] = wv[j];
        if (n_here & 1u) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (2u * j + 1 == n_here) dst[2 * j] = static_cast<int16_t>(wv[j] & 0xffffu);
        }
      }
    }
    if (cg0 >= cg_end) {
      // (a warp with no column group in this tile still owes its arrival)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    PROG(6u, tile_no);
    if (tid == 0) TRACE2(u, 16 + 8 * min(tile_no, 4u) + 3);
  };

  if (warp < kConvWarps2) {
    // ================= converters: raw PCM -> byte planes in UMMA layout, in place; epilogue =================
    // A raw stage is kStreams rows of kRowBytes. Lanes of a warp walk ALONG a row in 16-byte items
    // (PPS items per stream, SPI streams per warp instruction: conflict-free LDS.128); each thread owns
    // kItems items per stage, item i of warp gw (of its group) belonging to stream
    // (8 gw + i) * SPI + lane / PPS of the tile's group.
    constexpr int FPI = 8 / CH;          // frames per 16-byte item
    constexpr int PPS = 64 / FPI;        // items per stream per stage (16 stereo, 8 mono)
    constexpr int SPI = 32 / PPS;        // streams per warp instruction (2 stereo, 4 mono)
    constexpr int kItems = 8;
    const uint32_t group = static_cast<uint32_t>(warp) >> 2, gw = warp & 3;
    const int conv_p = lane % PPS;       // item position inside the stream's row
    const uint32_t sl0 = static_cast<uint32_t>(8 * gw * SPI + lane / PPS);
    // where item i sits in a raw stage: whole-stage box, or four quarter boxes (16 frames each)
    const uint32_t raw_off0 = sl0 * kRowBytes + conv_p * 16;
    constexpr uint32_t kRawItemStride = SPI * kRowBytes;
    constexpr int kItemsPerQuarter = PPS / 4;
    const uint32_t rawq_off0 = static_cast<uint32_t>(conv_p / kItemsPerQuarter) * kQuarterBytes + sl0 * kQuarterRow +
                               static_cast<uint32_t>(conv_p % kItemsPerQuarter) * 16;
    constexpr uint32_t kRawqItemStride = SPI * kQuarterRow;
    // byte offset of item i's hi/left word(s) inside a converted stage: conv_off0 + i * kItemStride
    // (stereo: left channel of stream sl -> row 2 sl, right channel -> row 2 sl + 1)
    constexpr uint32_t kItemStride = SPI * 16 * CH;
    const uint32_t conv_off0 = [&] {
      const uint32_t j = static_cast<uint32_t>(conv_p * FPI) / kUmmaChunkFrames;
      const uint32_t byte_in_row = static_cast<uint32_t>(conv_p * FPI) % kUmmaChunkFrames;
      return (j >> 1) * x_kstep(CH) + (j & 1) * x_lbo(CH) + sl0 * (16 * CH) + byte_in_row;
    }();
    const uint32_t total_stages = n_tiles_mine * n_iters;

    // The previous call's grid may still be writing the output rows this call overwrites.
    if (tid == 0) TRACE2(u, 3);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) TRACE2(u, 4);

    // ---- this group's stages: q = group, group + 2, ... of the CTA's stage sequence ----
    // Every thread watches EVERY phase of the raw_full barriers in order, also those of the other
    // group's stages: a parity wait tells "this phase" from "the one before" only. With an odd number
    // of slots a group meets a slot every other time it is used; skipping a phase would let the wait
    // for use k + 1 be satisfied by the completion of use k - 1 (seen as a hang, then a fault, on the
    // first shape that got three slots).
    uint32_t slot = 0, par = 0;         // ring slot of stage q, parity of its raw_full phase
    uint32_t c_tile = 0, c_it = 0;      // (tile, stage) of stage q
    uint32_t done_tile = 0;             // tiles whose epilogue this warp has run
    int c_kf0 = 0;
    bool have_kf0 = false;
#ifdef SPXB_UMMA2_TRACE
    uint32_t step_no = 0;
#define STEP_MARK(k) \
  if (tid == 0 && step_no == n_iters / 2 + 2) TRACE2(u, 56 + (k))
#else
#define STEP_MARK(k)
#endif
    for (uint32_t q = 0; q < total_stages; ++q, ++c_it) {
      if ((q & 1u) != group) {
        // the other group's stage: only follow the barrier
        WAIT(&raw_full[slot], par, 0x09000000u | (slot << 20) | (par << 16) | q);
        if (++slot == S) {
          slot = 0;
          par ^= 1u;
        }
        continue;
      }
      while (c_it >= n_iters) {
        c_it -= n_iters;
        ++c_tile;
        have_kf0 = false;
      }
      // every stage of the tiles before c_tile is stored: their epilogues are due (the MMAs of
      // c_tile cannot start before the accumulator of c_tile - 1 has been read out)
      if (!DB)
        for (; done_tile < c_tile; ++done_tile) epilogue(done_tile, group, 2, 0, done_tile & 1u);
      if (!have_kf0) {
        c_kf0 = tile_kf0((w_begin + c_tile) / G);
        have_kf0 = true;
      }
      const int f0 = c_kf0 + static_cast<int>(c_it) * kStageFrames;
      const bool quarters = f0 < 0 && f0 + kStageFrames > 0;  // the stage that straddles the end of the history
      uint8_t *xs = smem + slot * kXStageBytes;
      STEP_MARK(0);
      PROG(1u, q);
      WAIT(&raw_full[slot], par, 0x03000000u | (slot << 20) | (par << 16) | q);
      STEP_MARK(1);
      const uint8_t *rp = xs + (quarters ? rawq_off0 : raw_off0);
      const uint32_t rstride = quarters ? kRawqItemStride : kRawItemStride;
      uint4 w[kItems];
#pragma unroll
      for (int i = 0; i < kItems; ++i) w[i] = *reinterpret_cast<const uint4 *>(rp + i * rstride);
      // split into the byte planes (in registers: the values depend on the loads, so every raw byte
      // of this thread has been read when it reaches the barrier)
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        const uint4 v = w[i];
        if (CH == 1) {
          // 8 frames; word = (x[2k+1] << 16) | x[2k]: bytes lo0 hi0 lo1 hi1 -> 8 bytes per plane
          w[i] = make_uint4(__byte_perm(v.x, v.y, 0x7531), __byte_perm(v.z, v.w, 0x7531),
                            __byte_perm(v.x, v.y, 0x6420), __byte_perm(v.z, v.w, 0x6420));
        } else {
          // 4 frames; word f = (R_f << 16) | L_f -> one word per plane and channel (rows 2 sl, 2 sl + 1)
          const uint32_t ul = __byte_perm(v.x, v.y, 0x5140), vl = __byte_perm(v.z, v.w, 0x5140);
          const uint32_t ur = __byte_perm(v.x, v.y, 0x7362), vr = __byte_perm(v.z, v.w, 0x7362);
          w[i] = make_uint4(__byte_perm(ul, vl, 0x7632), __byte_perm(ur, vr, 0x7632),
                            __byte_perm(ul, vl, 0x5410), __byte_perm(ur, vr, 0x5410));
        }
      }
      PROG(2u, q);
      {
        // (a store of a value that depends on every word: the warp cannot reach the barrier with loads
        // from the slot still in flight, whatever the compiler does with the byte splitting above)
        uint32_t seen = 0;
#pragma unroll
        for (int i = 0; i < kItems; ++i) seen ^= w[i].x ^ w[i].y ^ w[i].z ^ w[i].w;
        seen = __reduce_xor_sync(0xffffffffu, seen);
        if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&raw_seen[warp]) = seen;
      }
      group_sync(group);  // every thread of the group holds its share of the slot in registers
      STEP_MARK(2);
      PROG(3u, q);
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        uint8_t *base = xs + conv_off0 + i * kItemStride;
        if (CH == 1) {
          *reinterpret_cast<uint2 *>(base) = make_uint2(w[i].x, w[i].y);
          *reinterpret_cast<uint2 *>(base + kXPlaneBytes) = make_uint2(w[i].z, w[i].w);
        } else {
          *reinterpret_cast<uint32_t *>(base) = w[i].x;
          *reinterpret_cast<uint32_t *>(base + 16) = w[i].y;
          *reinterpret_cast<uint32_t *>(base + kXPlaneBytes) = w[i].z;
          *reinterpret_cast<uint32_t *>(base + kXPlaneBytes + 16) = w[i].w;
        }
      }
      STEP_MARK(3);
      // every thread makes its own stores visible to the async proxy, then one lane per warp arrives
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_full[slot]);
      PROG(4u, q);
      STEP_MARK(4);
#ifdef SPXB_UMMA2_TRACE
      ++step_no;
#endif
      if (++slot == S) {
        slot = 0;
        par ^= 1u;
      }
    }
    if (!DB)
      for (; done_tile < n_tiles_mine; ++done_tile) epilogue(done_tile, group, 2, 0, done_tile & 1u);
    if (tid == 0) TRACE2(u, 10);
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
    const uint32_t acc0 = tmem_slot;
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
          WAIT(&taps_free, (run - 1) & 1u, 0x04000000u | run);
          load_tap_tile(t);
        }
        cur_t = t;
        tap_par = run & 1u;
        ++run;
      }
      const bool run_ends = w + 1 == w_end || (w + 1) / G != t;
      const uint32_t tr = 16 + 8 * min(tile_no, 4u);  // trace slots of this tile
      const uint32_t buf = DB ? (tile_no & 1u) : 0u, buf_use = DB ? (tile_no >> 1) : tile_no;
      if (buf_use) {
        // the epilogue has read this set's previous tile out of TMEM
        WAIT(&acc_empty[buf], (buf_use - 1) & 1u, 0x05000000u | tile_no);
        tc_fence_after_sync();
      }
      const uint32_t acc = acc0 + buf * 4u * nt;
      if (lane == 0) TRACE2(u, tr + 4);
      for (uint32_t it = 0; it < n_iters; ++it) {
        if (tile_no == 1 && it < 16 && lane == 0) TRACE2(u, 64 + 3 * it);
        PROG(1u, (tile_no << 8) | it);
        if (new_run) WAIT(&tap_full[it], tap_par, 0x06000000u | (it << 8) | run);
        WAIT(&x_full[slot], par, 0x07000000u | (slot << 20) | (par << 16) | (tile_no << 8) | it);
        tc_fence_after_sync();
        PROG(2u, (tile_no << 8) | it);
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
            // but uniform adds between the MMAs
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
            // The record walk, one record ahead: the uniform loads of record m + 1 are issued before
            // the MMAs of record m, so their latency is not part of the chain between two MMAs.
            const uint32_t m_end = u.stage_rec[it + 1];
            uint32_t m = u.stage_rec[it];
            uint32_t rb = u.rec[m].b, ri = u.rec[m].idesc_hi, rd = u.rec[m].d_a;
#pragma unroll 1
            for (; m < m_end; ++m) {
              const uint32_t nb = u.rec[m + 1].b, ni = u.rec[m + 1].idesc_hi, nd = u.rec[m + 1].d_a;
              const uint64_t b = b_fixed + rb;
              const uint64_t a_hi = a_st + (rd >> 16);
              const uint32_t d_hi = acc + (rd & 0xffffu);
#ifndef SPXB_DBG_NOMMA
              umma_i8(d_hi, a_hi, b, ri, 1u);
              umma_i8(d_hi + nt, a_hi + a_lo16, b, ri & ~(1u << 7), 1u);
#else
              if (b == 1 && a_hi == 2 && d_hi == 3) umma_i8(d_hi, a_hi, b, ri, 1u);  // timing experiment: no MMAs
#endif
              rb = nb;
              ri = ni;
              rd = nd;
            }
          }
          umma_commit(&x_empty[slot]);
          if (last) {
            umma_commit(&acc_full[buf]);
            if (run_ends) umma_commit(&taps_free);
          }
        }
        __syncwarp();
        PROG(3u, (tile_no << 8) | it);
        if (tile_no == 1 && it < 16 && lane == 0) TRACE2(u, 66 + 3 * it);
        if (last && lane == 0) TRACE2(u, tr + 6);
        a_st += a_stage16;
        if (++slot == S) {
          slot = 0;
          par ^= 1u;
          a_st = a_base;
        }
      }
    }
    if (lane == 0) TRACE2(u, 11);
  } else if (warp == kLoadWarp) {
    // ================= loader: one TMA box of raw PCM per stage =================
    if (lane == 0) {
      // the previous call's grid wrote the history this call reads
      asm volatile("griddepcontrol.wait;" ::: "memory");
      TRACE2(u, 7);
      uint32_t slot = 0, par = 1;  // a fresh barrier passes a wait on the phase "before the first"
      const int hist_frames = static_cast<int>(a.hist_frames);
      for (uint32_t tile_no = 0; tile_no < n_tiles_mine; ++tile_no) {
        const uint32_t w = w_begin + tile_no, t = w / G, g = w - t * G;
        const int kf0 = tile_kf0(t);
        const int row0 = static_cast<int>(g * kStreams);
        for (uint32_t it = 0; it < n_iters; ++it) {
          // Ask L2 for the input of a stage several slots ahead (SPXB_PF_DIST stages; the frames a call
          // brings in come from HBM, and a ring slot is too precious to sit through that latency)
          if (SPXB_PF_DIST) {
            const int fp = kf0 + static_cast<int>(it + SPXB_PF_DIST) * kStageFrames;
            if (it + SPXB_PF_DIST < n_iters && fp >= 0 && fp < static_cast<int>(sc.n_in))
              tma_prefetch_2d(&maps.in64, fp * CH * 2, row0);
          }
          WAIT_L(&x_empty[slot], par, 0x08000000u | (slot << 20) | (par << 16) | (tile_no << 8) | it);
          uint8_t *dst = smem + slot * kXStageBytes;
          const int f0 = kf0 + static_cast<int>(it) * kStageFrames;
#ifdef SPXB_DBG_NOLOAD
          mbar_arrive(&raw_full[slot]);  // timing experiment: no PCM traffic
#else
          mbar_arrive_expect_tx(&raw_full[slot], kRawBytes);
#endif
          if (f0 >= 0) {
            tma_box_2d(dst, &maps.in64, f0 * CH * 2, row0, &raw_full[slot]);
          } else if (f0 + kStageFrames <= 0) {
            tma_box_2d(dst, &maps.hist64, (hist_frames + f0) * CH * 2, row0, &raw_full[slot]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int f = f0 + 16 * j;
              if (f < 0) tma_box_2d(dst + j * kQuarterBytes, &maps.hist16, (hist_frames + f) * CH * 2, row0, &raw_full[slot]);
              else tma_box_2d(dst + j * kQuarterBytes, &maps.in16, f * CH * 2, row0, &raw_full[slot]);
            }
          }
          if (tile_no == 0 && it == 0) TRACE2(u, 2);
          if (++slot == S) {
            slot = 0;
            par ^= 1u;
          }
        }
      }
    }
  } else if (DB && warp >= kEpiWarp0) {
    // ================= epilogue warps (two accumulator sets) =================
    // the previous call's grid may still be writing the output rows this call overwrites
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (uint32_t tile_no = 0; tile_no < n_tiles_mine; ++tile_no) epilogue(tile_no, 0, 1, tile_no & 1u, (tile_no >> 1) & 1u);
  } else {
    // ================= history slide (resample.c:898-899) and the new position =================
    // Runs beside the FIR: it reads the old history and this call's input, writes the other half
    // of the ping-pong. Stream sl of group g is handled with the tile (t, g) for which
    // t == sl % T; new history element e = element consumed*CH + e of (old history || input).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t hist_elems = a.hist_frames * CH;
    const size_t shift = static_cast<size_t>(sc.consumed) * CH;
    const int vw = (shift % 8 == 0) ? 8 : (shift % 4 == 0) ? 4 : (shift % 2 == 0) ? 2 : 1;
    const uint32_t hw = static_cast<uint32_t>(warp - kHistWarp0);  // the history warps take turns at units of 4 streams
    for (uint32_t w = w_begin; w < w_end; ++w) {
      const uint32_t t = w / G, g = w - t * G;
      if (t >= static_cast<uint32_t>(kStreams)) continue;
      const uint32_t first = g * kStreams + t;
      const uint32_t in_group = (kStreams - t + T - 1) / T;
      const uint32_t in_batch = first < n_rows ? (n_rows - first + T - 1) / T : 0u;
      const uint32_t n_mine = min(in_group, in_batch);
      if (vw == 8) slide_streams<uint4>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      else if (vw == 4) slide_streams<uint2>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      else if (vw == 2) slide_streams<uint32_t>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      else slide_streams<uint16_t>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      for (uint32_t k = lane + 32 * hw; k < n_mine; k += 32 * kHistWarps) {
        const size_t s = first + static_cast<size_t>(k) * T;
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

// ---------------------------------------------------------------------------------------------
// Variant with the A operand in TENSOR MEMORY (opt-in: SPXB_UMMA_ATMEM=1; single accumulator set).
// The byte planes never touch shared memory: a converter thread owns one series (= TMEM lane), reads
// its stream's raw bytes of a stage from the ring slot, splits them in registers and writes them with
// tcgen05.st into the A ring behind the accumulator columns; the MMAs take A from there
// (csrc/atmem_probe.cu: row = lane, column j = K bytes 4j .. 4j+3, little-endian). What that buys:
//   * a raw slot is free again as soon as its bytes are in registers -- not after the MMAs that consume
//     the stage, as with the in-place conversion -- so the three slots a 168 KB tap tile leaves cover
//     the TMA latency; the plane ring that replaces it is TMEM columns, not shared memory;
//   * the tensor core stops reading the planes from shared memory (43 % of its operand bytes).
// Raw boxes are 128 bytes wide with the 128-byte swizzle (16-byte chunk c of row r at c ^ (r & 7)), so
// that 32 lanes reading 32 different rows at the same chunk do not collide: two boxes per stage for
// stereo (one per K step), one for mono. The stage that straddles the end of the history arrives as
// four unswizzled quarter boxes, as in the kernel above.
struct alignas(64) Umma3Maps {
  CUtensorMap in_sw, hist_sw, in16, hist16;
};

__device__ __forceinline__ void umma_i8_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

constexpr uint32_t kMaxRawSlots = 6, kMaxASlots = 4, kRawSlotBytes = 16384;
#ifdef SPXB_UMMA2_WATCHDOG
#define PROG3(code, v)                                                                                              \
  do {                                                                                                              \
    if ((threadIdx.x & 31) == 0 && blockIdx.x < 4 && u.trace)                                                      \
      reinterpret_cast<volatile unsigned long long *>(u.trace)[600 + blockIdx.x * 16 + (threadIdx.x >> 5)] =        \
          (static_cast<unsigned long long>(code) << 32) | (v);                                                      \
  } while (0)
#else
#define PROG3(code, v) \
  do {                 \
  } while (0)
#endif

template <int CH>
__global__ void __launch_bounds__(kThreads2, 1)
    umma3_fir_kernel(const __grid_constant__ CallArgs a, const __grid_constant__ Umma2Args u,
                     const __grid_constant__ Umma3Maps maps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // barriers at fixed offsets: raw_full [12], raw_empty [6], a_full, a_empty [4 each], tap_full [32], 4 more.
  // Stage q's box lands in raw slot q % R but completes barrier raw_full[q % 2R]: with two converter
  // groups on alternate stages and 2R even, each of those barriers is waited on by ONE group, which
  // then sees every one of its phases -- whatever R is. (A parity wait tells a phase only from the one
  // before it: a barrier whose uses alternate between the groups let a group that met it every other
  // time be satisfied by a stale phase. Seen as a hang, twice.) The A ring has an even number of slots
  // for the same reason; raw_empty, a_full, tap_full and the accumulator's have one waiting party.
  constexpr uint32_t kBarRE = 2 * kMaxRawSlots, kBarA = kBarRE + kMaxRawSlots, kBarTap = kBarA + 2 * kMaxASlots,
                     kBarMisc = kBarTap + kMaxTapStages, kBars = kBarMisc + 4;
  static_assert(kBars <= 64, "two init instructions per lane");
  __shared__ uint64_t bars[kBars];
  uint64_t *const raw_full = bars, *const raw_empty = bars + kBarRE;
  uint64_t *const a_full = bars + kBarA, *const a_empty = bars + kBarA + kMaxASlots;
  uint64_t *const tap_full = bars + kBarTap;
  uint64_t &taps_free = bars[kBarMisc], &acc_full = bars[kBarMisc + 1], &acc_empty = bars[kBarMisc + 2],
           &tmem_ready = bars[kBarMisc + 3];
  __shared__ uint32_t tmem_slot, raw_seen[kConvWarps2];

  constexpr int kStreams = kUmmaRows / CH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const StreamCall sc = a.uniform;
  const uint32_t nt = u.nt, G = u.n_groups, T = u.n_tiles;
  const uint32_t w_begin = static_cast<uint32_t>(static_cast<unsigned long long>(blockIdx.x) * u.n_work / gridDim.x);
  const uint32_t w_end = static_cast<uint32_t>(static_cast<unsigned long long>(blockIdx.x + 1) * u.n_work / gridDim.x);
  const uint32_t n_tiles_mine = w_end - w_begin;
  constexpr int kStageFrames = kStageChunks * kUmmaChunkFrames;  // 64
  constexpr uint32_t kRowBytes = kStageFrames * CH * 2;          // raw bytes of one stream in a stage
  constexpr uint32_t kQuarterRow = kRowBytes / 4, kQuarterBytes = kRawSlotBytes / 4;
  static_assert(kStreams * kRowBytes == kRawSlotBytes, "a raw stage is 16 KB");
  const uint32_t R = u.x_stages;                     // raw ring slots
  const uint32_t A = (512u - 4u * nt) / 32u >= 4u ? 4u : 2u;  // A ring slots (32 columns each) behind the accumulator; even
  const uint32_t a_col0 = 4u * nt;
  const uint32_t n_iters = (u.ksteps + 1) / 2;
  uint8_t *const tap_smem = smem + R * kRawSlotBytes;
  const uint32_t n_rows = a.n_streams;
  auto tile_kf0 = [&](uint32_t t) -> int { return u.n_inline ? u.inl[t].kf0 : u.tiles[t].kf0; };
  auto tile_slot = [&](uint32_t t) -> uint32_t { return u.n_inline ? u.inl[t].slot : u.tiles[t].slot; };

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) TRACE2(u, 0);
#if SPXB_PREFETCH_MAPS
  // the first box of a launch otherwise waits for its (cold) descriptor
  if (warp == kLoadWarp && lane < 4)
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<const char *>(&maps) + lane * sizeof(CUtensorMap)) : "memory");
#endif
#ifdef SPXB_UMMA2_TRACE
  if (tid == 11 * 32 && u.trace) {
    unsigned long long gt;
    uint32_t smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    u.trace[static_cast<size_t>(blockIdx.x) * kTraceSlots2 + 14] = gt;
    u.trace[static_cast<size_t>(blockIdx.x) * kTraceSlots2 + 13] = smid;
  }
#endif
  if (warp == kMmaWarp2) {
#pragma unroll
    for (uint32_t i = lane; i < kBars; i += 32) {
      const bool four = (i >= kBarRE && i < kBarA + kMaxASlots);  // raw_empty, a_full: one arrival per warp of a group
      mbar_init(&bars[i], four ? kGroupWarps : i == kBarMisc + 2 ? kConvWarps2 : 1u);
    }
    fence_mbar_init();
  }
  __syncthreads();

  auto load_tap_tile = [&](uint32_t t) {
    const int8_t *src = u.pool + static_cast<size_t>(tile_slot(t)) * u.tile_bytes;
    if (static_cast<uint32_t>(lane) < n_iters) {
      const uint32_t off = u.stage_off[lane], bytes = u.stage_off[lane + 1] - off;
      mbar_arrive_expect_tx(&tap_full[lane], bytes);
      bulk_g2s(tap_smem + off, src + off, bytes, &tap_full[lane]);
    }
  };

  if (warp < kConvWarps2) {
    // ================= converters (one series per thread) + epilogue =================
    const uint32_t group = static_cast<uint32_t>(warp) >> 2, quarter = warp & 3;
    const uint32_t row = quarter * 32 + lane;                       // series of the tile = TMEM lane
    const uint32_t sl = CH == 2 ? row >> 1 : row, ch = CH == 2 ? (row & 1u) : 0u;  // stream of the group, channel
    const uint32_t sw = (sl & 7u) * 16u;                            // the 128-byte swizzle of this stream's row
    // stereo: selectors that pick this thread's channel out of two (R << 16 | L) words
    const uint32_t sel = ch ? 0x7632u : 0x5410u;
    const uint32_t lane_taddr = (quarter * 32u) << 16;
    const uint32_t total_stages = n_tiles_mine * n_iters;
    uint32_t tmem = 0;

    // ---- epilogue of one tile (as in the kernel above, the two warps of a lane quarter share the column groups) ----
    const uint32_t sl_out = sl, ch_out = ch;
    const uint32_t out_bits = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(a.out)) | (static_cast<uint32_t>(a.out_stride) * 2u);
    const int out_align = (out_bits & 15u) == 0 ? 16 : (out_bits & 3u) == 0 ? 4 : 2;
    auto epilogue = [&](uint32_t tile_no) {
      const uint32_t w = w_begin + tile_no, t = w / G, g = w - t * G;
      const uint32_t m0 = t * nt;
      const uint32_t n_valid = min(nt, sc.n_out - m0);
      const uint32_t s_out = g * kStreams + sl_out;
      const bool live_out = s_out < n_rows;
      int16_t *out_row = a.out + static_cast<size_t>(live_out ? s_out : 0) * a.out_stride + static_cast<size_t>(m0) * CH;
      if (tid == 0) TRACE2(u, 16 + 8 * min(tile_no, 4u) + 1);
      PROG3(0x20u, tile_no);
      WAIT(&acc_full, tile_no & 1u, 0x12000000u | tile_no);
      tc_fence_after_sync();
      const uint32_t lane_addr = tmem + lane_taddr;
      const uint32_t cg_end = (n_valid + 15) / 16;
      for (uint32_t cg = group; cg < cg_end; cg += 2) {
        uint32_t p0[16], p1[16], p2[16], p3[16];
        tmem_ld16(lane_addr + cg * 16, p0);
        tmem_ld16(lane_addr + nt + cg * 16, p1);
        tmem_ld16(lane_addr + 2 * nt + cg * 16, p2);
        tmem_ld16(lane_addr + 3 * nt + cg * 16, p3);
        tmem_ld_wait();
        if (cg + 2 >= cg_end) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty);
        }
        int r16[16];
        combine16(p0, p1, p2, p3, u.shift, r16);
        uint32_t wv[8];
        uint32_t first_elem;
        if (CH == 2) {
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
        const uint32_t n_here = first_elem >= total ? 0u : min(16u, total - first_elem);
        int16_t *dst = out_row + first_elem;
        if (n_here == 16 && out_align == 16) {
          reinterpret_cast<uint4 *>(dst)[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          reinterpret_cast<uint4 *>(dst)[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
        } else if (out_align >= 4) {
          // the ragged end of the call's output, or output rows that are only 4-byte aligned (e.g. 882
          // stereo frames per row): whole 32-bit words, then possibly one last int16
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
      if (group >= cg_end) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty);
      }
      if (tid == 0) TRACE2(u, 16 + 8 * min(tile_no, 4u) + 3);
    };

    asm volatile("griddepcontrol.wait;" ::: "memory");
    WAIT(&tmem_ready, 0, 0x11000000u);
    tc_fence_after_sync();
    tmem = tmem_slot;

    // this group's stages: q = group, group + 2, ...; raw slot q % R, raw_full barrier q % 2R, A slot q % A
    uint32_t rs = group % R, rb = group, rpar = 0;  // (2R >= 4 > group)
    uint32_t as = group, apar = 1;                  // (A >= 2; a fresh barrier passes a wait on the phase before the first)
    uint32_t c_tile = 0, c_it = group, done_tile = 0;
    int c_kf0 = 0;
    bool have_kf0 = false;
    for (uint32_t q = group; q < total_stages; q += 2, c_it += 2) {
      // The epilogue of tile t is due once this group has stored its last stage of t. It is put off by
      // ONE stage when that is possible: this group's first stage of tile t + 1 goes into the A ring
      // first (its slot is released by MMAs of tile t, not by the epilogue), so that the next tile's
      // MMAs find their operands the moment the accumulator is handed back.
      bool epilogue_after = false;
      while (c_it >= n_iters) {
        c_it -= n_iters;
        ++c_tile;
        have_kf0 = false;
      }
      if (SPXB_DEFER_EPILOGUE && done_tile + 1 == c_tile && c_it < 2 && n_iters >= 2) epilogue_after = true;
      else
        for (; done_tile < c_tile; ++done_tile) epilogue(done_tile);
      if (!have_kf0) {
        c_kf0 = tile_kf0((w_begin + c_tile) / G);
        have_kf0 = true;
      }
      PROG3(0x21u, q);
      WAIT(&raw_full[rb], rpar, 0x13000000u | (rb << 20) | (rpar << 16) | q);
      PROG3(0x22u, q);
      {
        if (quarter == 0 && lane == 0 && c_tile == 1 && c_it < 12) TRACE2(u, 100 + c_it);
        const int f0 = c_kf0 + static_cast<int>(c_it) * kStageFrames;
        const bool quarters = f0 < 0 && f0 + kStageFrames > 0;
        const uint8_t *slot = smem + rs * kRawSlotBytes;
        // this thread's series, 64 frames: hi and lo bytes, 16 words each (8 per K step)
        uint32_t hi[16], lo[16];
        if (CH == 2) {
          // a 16-byte chunk = 4 frames of (L, R); 8 chunks per K step
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            uint4 v;
            if (!quarters) {
              // box h = c / 8 (K step), row sl of 128 bytes, chunk c % 8 swizzled
              v = *reinterpret_cast<const uint4 *>(slot + (c >> 3) * (kRawSlotBytes / 2) + sl * 128 + (((c & 7) * 16) ^ sw));
            } else {
              // quarter box c / 4: rows of 64 bytes, not swizzled
              v = *reinterpret_cast<const uint4 *>(slot + (c >> 2) * kQuarterBytes + sl * kQuarterRow + (c & 3) * 16);
            }
            const uint32_t p01 = __byte_perm(v.x, v.y, sel), p23 = __byte_perm(v.z, v.w, sel);
            hi[c] = __byte_perm(p01, p23, 0x7531);
            lo[c] = __byte_perm(p01, p23, 0x6420);
          }
        } else {
          // a 16-byte chunk = 8 frames; 4 chunks per K step, two words per plane and chunk
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 v;
            if (!quarters) v = *reinterpret_cast<const uint4 *>(slot + sl * 128 + ((c * 16) ^ sw));
            else v = *reinterpret_cast<const uint4 *>(slot + (c >> 1) * kQuarterBytes + sl * kQuarterRow + (c & 1) * 16);
            hi[2 * c] = __byte_perm(v.x, v.y, 0x7531);
            hi[2 * c + 1] = __byte_perm(v.z, v.w, 0x7531);
            lo[2 * c] = __byte_perm(v.x, v.y, 0x6420);
            lo[2 * c + 1] = __byte_perm(v.z, v.w, 0x6420);
          }
        }
        // The slot can be refilled once every lane's raw bytes are IN registers. Having issued the loads
        // is not enough (one row of one tile came out wrong in ~2 % of the launches: a chunk of the NEXT
        // box read in place of this one's). Lane 0 stores a value reduced over words that depend on every
        // load of every lane before it arrives: the store cannot issue until those loads have completed.
        {
          uint32_t seen = 0;
#pragma unroll
          for (int c = 0; c < 16; ++c) seen ^= hi[c] ^ lo[c];
          seen = __reduce_xor_sync(0xffffffffu, seen);
          if (lane == 0) {
            *reinterpret_cast<volatile uint32_t *>(&raw_seen[warp]) = seen;
            mbar_arrive(&raw_empty[rs]);
          }
        }
        // the MMAs that read this A slot last have completed
        PROG3(0x23u, q);
        WAIT(&a_empty[as], apar, 0x14000000u | (as << 20) | (apar << 16) | q);
        tc_fence_after_sync();
        PROG3(0x24u, q);
        const uint32_t ta = tmem + lane_taddr + a_col0 + as * 32u;
        uint32_t r8[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int j = 0; j < 8; ++j) r8[j] = hi[8 * h + j];
          tmem_st8(ta + 16 * h, r8);
#pragma unroll
          for (int j = 0; j < 8; ++j) r8[j] = lo[8 * h + j];
          tmem_st8(ta + 16 * h + 8, r8);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
        PROG3(0x25u, q);
        if (epilogue_after)
          for (; done_tile < c_tile; ++done_tile) epilogue(done_tile);
      }
      rs += 2;
      if (rs >= R) rs -= R;
      rb += 2;
      if (rb >= 2 * R) {
        rb -= 2 * R;
        rpar ^= 1u;
      }
      as += 2;
      if (as >= A) {
        as -= A;
        apar ^= 1u;
      }
    }
    PROG3(0x26u, done_tile);
    for (; done_tile < n_tiles_mine; ++done_tile) epilogue(done_tile);
    PROG3(0x27u, done_tile);
  } else if (warp == kMmaWarp2) {
    // ================= MMA issue =================
    PROG3(0x33u, 0);
    load_tap_tile(w_begin / G);
    tmem_alloc(&tmem_slot, 512);
    PROG3(0x34u, 0);
    tmem_relinquish();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tmem_ready);
    tc_fence_after_sync();
    const uint32_t acc = tmem_slot;
    const uint32_t n3 = 3 * nt;
    const uint32_t np0 = min(n3, 256u), np1 = n3 - np0;
    const uint32_t nq0 = min(2 * nt, 256u), nq1 = 2 * nt - nq0;
    const uint32_t id_hi = umma_idesc_i8(128, 0, true, true), id_lo = umma_idesc_i8(128, 0, false, true);
    auto with_n = [](uint32_t idesc, uint32_t n) { return idesc | ((n >> 3) << 17); };
    const uint64_t b_fixed = umma_smem_desc(smem_u32(tap_smem), 0, 128);
    // A ring: group g = stage parity owns slots [a_base[g], a_base[g] + a_n[g])
    uint32_t as = 0, apar = 0, tile_no = 0, run = 0;
    uint32_t cur_t = 0xffffffffu;
    for (uint32_t w = w_begin; w < w_end; ++w, ++tile_no) {
      const uint32_t t = w / G;
      const bool new_run = t != cur_t;
      uint32_t tap_par = 0;
      if (new_run) {
        if (run) {
          WAIT(&taps_free, (run - 1) & 1u, 0x16000000u | run);
          load_tap_tile(t);
        }
        cur_t = t;
        tap_par = run & 1u;
        ++run;
      }
      const bool run_ends = w + 1 == w_end || (w + 1) / G != t;
      if (tile_no) {
        WAIT(&acc_empty, (tile_no - 1) & 1u, 0x17000000u | tile_no);
        tc_fence_after_sync();
      }
      const uint32_t tr = 16 + 8 * min(tile_no, 4u);
      if (lane == 0) TRACE2(u, tr + 4);
      for (uint32_t it = 0; it < n_iters; ++it) {
        if (tile_no == 1 && it < 12 && lane == 0) TRACE2(u, 64 + 3 * it);
        if (new_run) WAIT(&tap_full[it], tap_par, 0x18000000u | (it << 8) | run);
        PROG3(0x30u, (tile_no << 8) | it);
        WAIT(&a_full[as], apar, 0x19000000u | (as << 20) | (apar << 16) | (tile_no << 8) | it);
        PROG3(0x31u, (tile_no << 8) | it);
        tc_fence_after_sync();
        if (tile_no == 1 && it < 12 && lane == 0) TRACE2(u, 65 + 3 * it);
        if (it == 0 && lane == 0) TRACE2(u, tr + 5);
        const bool last = it + 1 == n_iters;
        const uint32_t a_st = acc + a_col0 + as * 32u;  // [K step 0: hi 8 cols, lo 8 cols][K step 1: hi, lo]
        if (elect_one()) {
          if (it == 0) {
            const uint64_t b_k = b_fixed + (static_cast<uint64_t>(n3) << 16);
            umma_i8_ta(acc, a_st, b_k, with_n(id_hi, np0), 0u);
            if (np1) umma_i8_ta(acc + 256, a_st, b_k + 256, with_n(id_hi, np1), 0u);
            umma_i8_ta(acc + nt, a_st + 8, b_k, with_n(id_lo, nq0), 1u);
            if (nq1) umma_i8_ta(acc + nt + 256, a_st + 8, b_k + 256, with_n(id_lo, nq1), 1u);
            umma_i8_ta(acc + 3 * nt, a_st + 8, b_k + 2 * nt, with_n(id_lo, nt), 0u);
          }
          const uint32_t m_end = u.stage_rec[it + 1];
          uint32_t m = u.stage_rec[it];
          uint32_t rb = u.rec[m].b, ri = u.rec[m].idesc_hi, rd = u.rec[m].d_a;
#pragma unroll 1
          for (; m < m_end; ++m) {
            const uint32_t nb = u.rec[m + 1].b, ni = u.rec[m + 1].idesc_hi, nd = u.rec[m + 1].d_a;
            const uint64_t b = b_fixed + rb;
            const uint32_t a_hi = a_st + ((rd >> 16) ? 16u : 0u);
            const uint32_t d_hi = acc + (rd & 0xffffu);
            umma_i8_ta(d_hi, a_hi, b, ri, 1u);
            umma_i8_ta(d_hi + nt, a_hi + 8, b, ri & ~(1u << 7), 1u);
            rb = nb;
            ri = ni;
            rd = nd;
          }
          umma_commit(&a_empty[as]);
          if (last) {
            umma_commit(&acc_full);
            if (run_ends) umma_commit(&taps_free);
          }
        }
        __syncwarp();
        if (tile_no == 1 && it < 12 && lane == 0) TRACE2(u, 66 + 3 * it);
        if (last && lane == 0) TRACE2(u, tr + 6);
        if (++as == A) {
          as = 0;
          apar ^= 1u;
        }
      }
    }
    PROG3(0x32u, tile_no);
    if (lane == 0) TRACE2(u, 11);
  } else if (warp == kLoadWarp) {
    // ================= loader: the raw boxes of one stage per ring slot =================
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      uint32_t rs = 0, rb = 0, rpar = 1;  // raw slot, raw_full barrier; a fresh barrier passes a wait on the phase before the first
      const int hist_frames = static_cast<int>(a.hist_frames);
      constexpr int kBoxes = CH;                       // 128-byte boxes per stage
      constexpr int kBoxFrames = kStageFrames / kBoxes;
      for (uint32_t tile_no = 0; tile_no < n_tiles_mine; ++tile_no) {
        const uint32_t w = w_begin + tile_no, t = w / G, g = w - t * G;
        const int kf0 = tile_kf0(t);
        const int row0 = static_cast<int>(g * kStreams);
        for (uint32_t it = 0; it < n_iters; ++it) {
          PROG3(0x40u, (tile_no << 8) | it);
          WAIT(&raw_empty[rs], rpar, 0x1a000000u | (rs << 20) | (rpar << 16) | (tile_no << 8) | it);
          uint8_t *dst = smem + rs * kRawSlotBytes;
          const int f0 = kf0 + static_cast<int>(it) * kStageFrames;
          if (tile_no == 1 && it < 12) TRACE2(u, 112 + it);
          mbar_arrive_expect_tx(&raw_full[rb], kRawSlotBytes);
          if (f0 >= 0 || f0 + kStageFrames <= 0) {
#pragma unroll
            for (int k = 0; k < kBoxes; ++k) {
              const int f = f0 + k * kBoxFrames;
              if (f0 >= 0) tma_box_2d(dst + k * (kRawSlotBytes / kBoxes), &maps.in_sw, f * CH * 2, row0, &raw_full[rb]);
              else tma_box_2d(dst + k * (kRawSlotBytes / kBoxes), &maps.hist_sw, (hist_frames + f) * CH * 2, row0, &raw_full[rb]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int f = f0 + 16 * j;
              if (f < 0) tma_box_2d(dst + j * kQuarterBytes, &maps.hist16, (hist_frames + f) * CH * 2, row0, &raw_full[rb]);
              else tma_box_2d(dst + j * kQuarterBytes, &maps.in16, f * CH * 2, row0, &raw_full[rb]);
            }
          }
          if (++rs == R) {
            rs = 0;
            rpar ^= 1u;
          }
          if (++rb == 2 * R) rb = 0;
        }
      }
    }
  } else {
    // ================= history slide (resample.c:898-899) and the new position =================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t hist_elems = a.hist_frames * CH;
    const size_t shift = static_cast<size_t>(sc.consumed) * CH;
    const int vw = (shift % 8 == 0) ? 8 : (shift % 4 == 0) ? 4 : (shift % 2 == 0) ? 2 : 1;
    const uint32_t hw = static_cast<uint32_t>(warp - kHistWarp0);
    for (uint32_t w = w_begin; w < w_end; ++w) {
      const uint32_t t = w / G, g = w - t * G;
      if (t >= static_cast<uint32_t>(kStreams)) continue;
      const uint32_t first = g * kStreams + t;
      const uint32_t in_group = (kStreams - t + T - 1) / T;
      const uint32_t in_batch = first < n_rows ? (n_rows - first + T - 1) / T : 0u;
      const uint32_t n_mine = min(in_group, in_batch);
      if (vw == 8) slide_streams<uint4>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      else if (vw == 4) slide_streams<uint2>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      else if (vw == 2) slide_streams<uint32_t>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      else slide_streams<uint16_t>(a, T, first, n_mine, hist_elems, shift, lane, hw, kHistWarps);
      for (uint32_t k = lane + 32 * hw; k < n_mine; k += 32 * kHistWarps) {
        const size_t s = first + static_cast<size_t>(k) * T;
        a.last_sample[s] = sc.ls1;
        a.samp_frac[s] = sc.frac1;
      }
    }
  }

  PROG3(0x50u, 0);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp2) tmem_dealloc(tmem_slot, 512);
  PROG3(0x51u, 0);
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
  big_smem(umma2_fir_kernel<1, false>);
  big_smem(umma2_fir_kernel<2, false>);
  big_smem(umma2_fir_kernel<1, true>);
  big_smem(umma2_fir_kernel<2, true>);
  big_smem(umma3_fir_kernel<1>);
  big_smem(umma3_fir_kernel<2>);
}

// what the TMA-fed kernel needs of a call: the whole batch (no stream subset), an input to read, and
// every row -- input, output, history -- starting on a 16-byte boundary
bool umma2_planes_in_tmem() {
  static const bool v = [] {
    const char *e = getenv("SPXB_UMMA_ATMEM");
    return !e || atoi(e) != 0;
  }();
  return v;
}

bool umma2_covers(const CallArgs &a) {
  if (a.ids != nullptr || a.uniform.n_in == 0) return false;
  // TMA needs the input and history rows on 16-byte boundaries; the output rows only matter to the kernel
  // that stores whole vectors unconditionally (planes in shared memory)
  uintptr_t bits = reinterpret_cast<uintptr_t>(a.in) | (a.in_stride * 2u) | reinterpret_cast<uintptr_t>(a.hist_src) |
                   (a.hist_stride * 2u);
  if (!umma2_planes_in_tmem()) bits |= reinterpret_cast<uintptr_t>(a.out) | (a.out_stride * 2u);
  return (bits & 15u) == 0;
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#ifdef SPXB_UMMA2_WATCHDOG
unsigned long long *g_wd = nullptr;
void wd_dump() {
  static unsigned long long seen = 0;
  if (!g_wd) return;
  unsigned long long n = *reinterpret_cast<volatile unsigned long long *>(g_wd);
  if (n > 500) n = 500;
  for (; seen < n; ++seen) {
    const unsigned long long r = g_wd[1 + seen];
    fprintf(stderr, "watchdog: cta %llu thread %llu (warp %llu) tag %08llx\n", r >> 48, (r >> 32) & 0xffff, ((r >> 32) & 0xffff) / 32, r & 0xffffffffull);
  }
}
#endif

EncodeTiledFn encode_tiled() {
  static const EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// bytes x rows view of `rows` PCM rows of `row_bytes` valid bytes, `pitch` bytes apart; box = box_bytes x box_rows
bool pcm_map(CUtensorMap *m, const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch, uint32_t box_bytes,
             uint32_t box_rows, bool swizzle128 = false) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn || row_bytes == 0 || rows == 0) return false;
  const cuuint64_t gdim[2] = {row_bytes, rows};
  const cuuint64_t gstride[1] = {pitch};
  const cuuint32_t box[2] = {box_bytes, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// the MMA records and per-stage tables the kernel reads (they travel in the kernel parameters)
cudaError_t umma2_upload_plan(UmmaContext *c, cudaStream_t) {
  const uint32_t n_iters = (c->ksteps + 1) / 2;
  const uint32_t a_ks16 = x_kstep(static_cast<int>(c->channels)) >> 4;
  c->recs.clear();
  c->stage_off.assign(kMaxTapStages + 1, 0u);
  c->stage_rec.assign(kMaxTapStages + 2, 0u);
  for (uint32_t it = 0; it < n_iters; ++it) {
    c->stage_off[it] = c->packed.k[2 * it].off16 * 16u;
    c->stage_rec[it] = static_cast<uint16_t>(c->recs.size() / 3);
    for (uint32_t h = 0; h < 2; ++h) {
      const uint32_t k = 2 * it + h;
      if (k >= c->ksteps) break;
      const UmmaKStep &ks = c->packed.k[k];
      if (k == 0) continue;  // K step 0 has its own code in the kernel
      for (uint32_t e = 0; e < ks.n_ent; ++e) {
        c->recs.push_back((ks.off16 + ks.ent[e].row) | (static_cast<uint32_t>(ks.rows) << 16));
        c->recs.push_back(umma_idesc_i8(128, ks.ent[e].n, true, true));
        c->recs.push_back(ks.ent[e].dcol | ((h ? a_ks16 : 0u) << 16));
      }
    }
  }
  for (uint32_t it = n_iters; it <= kMaxTapStages; ++it) c->stage_off[it] = c->packed.tile_bytes;
  for (uint32_t it = n_iters; it < kMaxTapStages + 2; ++it) c->stage_rec[it] = static_cast<uint16_t>(c->recs.size() / 3);
  return c->recs.size() / 3 <= kMaxRecs ? cudaSuccess : cudaErrorInvalidValue;
}

// x stages the resident kernel would run for this geometry (0: the packed tile does not fit)
uint32_t umma2_x_stages(uint32_t channels, uint32_t tile_bytes, uint32_t ksteps) {
  const uint32_t xs = x_slot(static_cast<int>(channels));
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
  u.dense = 3 * c->nt <= 256 ? 1u : 0u;
  for (uint32_t k = 0; k < c->ksteps && u.dense; ++k) {
    const UmmaKStep &ks = c->packed.k[k];
    if (ks.n_ent != 1 || ks.ent[0].row != 0 || ks.ent[0].dcol != 0 || ks.ent[0].n != 3 * c->nt || ks.off16 != k * 6 * c->nt)
      u.dense = 0u;
  }
  u.shift = c->ft.shift;
  u.n_rec = static_cast<uint32_t>(c->recs.size() / 3);
  std::memcpy(u.rec, c->recs.data(), c->recs.size() * sizeof(uint32_t));
  for (uint32_t i = 0; i <= kMaxTapStages; ++i) u.stage_off[i] = c->stage_off[i];
  for (uint32_t i = 0; i < kMaxTapStages + 2; ++i) u.stage_rec[i] = c->stage_rec[i];
  u.trace = nullptr;
#ifdef SPXB_UMMA2_WATCHDOG
  {
    static unsigned long long *d_dbg = nullptr;
    if (!g_wd && cudaHostAlloc(reinterpret_cast<void **>(&g_wd), 1024 * 8, cudaHostAllocMapped) == cudaSuccess) {
      std::memset(g_wd, 0, 1024 * 8);
      std::thread([] {
        std::this_thread::sleep_for(std::chrono::seconds(25));
        wd_dump();
        for (int cta = 0; cta < 4; ++cta)
          for (int w = 0; w < 12; ++w)
            fprintf(stderr, "progress: cta %d warp %d code %02llx value %llu\n", cta, w, g_wd[600 + cta * 16 + w] >> 32, g_wd[600 + cta * 16 + w] & 0xffffffffull);
        fflush(stderr);
      }).detach();
      cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_dbg), g_wd, 0);
      atexit(wd_dump);
    }
    wd_dump();
    u.trace = d_dbg;
  }
#endif
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
  if (!umma2_covers(a)) return cudaErrorNotSupported;  // (umma_prepare only plans this kernel for calls it covers)
  // Tensor maps over this call's input and history rows. Host arithmetic only; they travel in the
  // kernel parameters, so a captured launch keeps its own.
  Umma2Maps maps;
  const uint32_t ch = a.channels, streams = kUmmaRows / ch, row_box = 64 * ch * 2;
  const uint64_t in_bytes = static_cast<uint64_t>(a.uniform.n_in) * ch * 2, hist_bytes = static_cast<uint64_t>(a.hist_frames) * ch * 2;
  if (!pcm_map(&maps.in64, a.in, in_bytes, a.n_streams, a.in_stride * 2, row_box, streams) ||
      !pcm_map(&maps.in16, a.in, in_bytes, a.n_streams, a.in_stride * 2, row_box / 4, streams) ||
      !pcm_map(&maps.hist64, a.hist_src, hist_bytes, a.n_streams, static_cast<uint64_t>(a.hist_stride) * 2, row_box, streams) ||
      !pcm_map(&maps.hist16, a.hist_src, hist_bytes, a.n_streams, static_cast<uint64_t>(a.hist_stride) * 2, row_box / 4, streams))
    return cudaErrorInvalidValue;
  cudaError_t e;
  const bool out_rows_aligned = ((reinterpret_cast<uintptr_t>(a.out) | (a.out_stride * 2u)) & 15u) == 0;
  // planes in tensor memory: wide tiles (one accumulator set), and any tile whose output rows are not on
  // 16-byte boundaries (the kernels with planes in shared memory store whole vectors)
  if (umma2_planes_in_tmem() && (c->n_acc == 1 || !out_rows_aligned) && 4 * c->nt + 64 <= 512) {
    // raw ring: as many 16 KB slots as fit beside the tap tile
    const uint32_t slots = std::min<uint32_t>(kMaxRawSlots, (kMaxSmem2 - c->tile_bytes) / kRawSlotBytes);
    if (slots < 2) return cudaErrorInvalidConfiguration;
    u.x_stages = slots;
    cfg.dynamicSmemBytes = slots * kRawSlotBytes + c->tile_bytes;
    Umma3Maps m3;
    if (!pcm_map(&m3.in_sw, a.in, in_bytes, a.n_streams, a.in_stride * 2, 128, streams, true) ||
        !pcm_map(&m3.hist_sw, a.hist_src, hist_bytes, a.n_streams, static_cast<uint64_t>(a.hist_stride) * 2, 128, streams, true))
      return cudaErrorInvalidValue;
    m3.in16 = maps.in16;
    m3.hist16 = maps.hist16;
    e = a.channels == 2 ? cudaLaunchKernelEx(&cfg, umma3_fir_kernel<2>, a, u, m3)
                        : cudaLaunchKernelEx(&cfg, umma3_fir_kernel<1>, a, u, m3);
  } else if (!out_rows_aligned) {
    return cudaErrorNotSupported;  // (the kernels below store whole vectors)
  } else if (c->n_acc == 2) {
    cfg.blockDim = dim3(kThreadsDB);
    e = a.channels == 2 ? cudaLaunchKernelEx(&cfg, umma2_fir_kernel<2, true>, a, u, maps)
                        : cudaLaunchKernelEx(&cfg, umma2_fir_kernel<1, true>, a, u, maps);
  } else {
    e = a.channels == 2 ? cudaLaunchKernelEx(&cfg, umma2_fir_kernel<2, false>, a, u, maps)
                        : cudaLaunchKernelEx(&cfg, umma2_fir_kernel<1, false>, a, u, maps);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) c->fresh_plan = false;
  if (e == cudaSuccess && launches) *launches += 1;
  return e;
}

}  // namespace spxb
