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
//     X stage   [plane 2][chunk 4][row 128][16 B]   K-major, no swizzle (LBO 2048, SBO 128)
//     tap stage [chunk 4][row 3nt][16 B]            one 1-D bulk copy from the tile pool
// Warp roles: 0-7 fetch int16 PCM (history for frames < 0, the call's input after), split it
// into byte planes with PRMT and store it in UMMA layout, later run the epilogue
// (tcgen05.ld, 64-bit recombination, round-half-up + saturate, interleaved int16 stores through
// shared memory); warp 8 lane 0 issues the tap bulk copies; warp 9 owns TMEM and lane 0
// issues the MMAs. Stages are handed over with mbarriers (full: 256 converter arrivals + the
// bulk copy's byte count; empty: tcgen05.commit).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "kernels_common.cuh"
#include "launch.h"
#include "umma_plan.h"
#include "umma_ptx.cuh"

namespace spxb {

namespace {

using namespace ptx;

constexpr int kConvWarps = 8;
constexpr int kConvThreads = kConvWarps * 32;
constexpr int kTmaWarp = 8, kMmaWarp = 9;
constexpr int kThreads = 320;
constexpr int kStageChunks = 4;                                  // 64 frames, 2 K steps
constexpr uint32_t kChunkBytesX = kUmmaRows * 16;                // 2048
constexpr uint32_t kXPlaneBytes = kStageChunks * kChunkBytesX;   // 8192
constexpr uint32_t kXStageBytes = 2 * kXPlaneBytes;              // hi + lo planes
constexpr int kMaxStages = 6;
constexpr uint32_t kMaxSmem = 227u * 1024u - 2048u;              // dynamic part; barriers are static

struct UmmaArgs {
  const UmmaTile *tiles;
  uint32_t n_tiles;
  uint32_t n_groups;
  const int8_t *pool;
  uint32_t tile_bytes;
  uint32_t nt;
  uint32_t ksteps;
  uint32_t stages;
  uint32_t tmem_cols;
  int shift;
};

// one MMA per <= 256 rows of B: D[:, col0 + (r - row0)] (+)= A * B[r]^T
__device__ __forceinline__ void issue_rows(uint32_t tmem, uint64_t adesc, bool a_signed, uint32_t b_addr,
                                           uint32_t b_lbo, uint32_t row0, uint32_t nrows, uint32_t col0,
                                           uint32_t acc) {
  for (uint32_t r = 0; r < nrows; r += 256) {
    const uint32_t n = min(256u, nrows - r);
    const uint64_t bdesc = umma_smem_desc(b_addr + (row0 + r) * 16, b_lbo, 128);
    umma_i8(tmem + col0 + r, adesc, bdesc, umma_idesc_i8(128, n, a_signed, true), acc);
  }
}

template <int CH>
__global__ void __launch_bounds__(kThreads, 1) umma_fir_kernel(const CallArgs a, const UmmaArgs u) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], acc_bar;
  __shared__ uint32_t tmem_slot;

  constexpr int kStreams = kUmmaRows / CH;  // streams per series group
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t g = blockIdx.x % u.n_groups, t = blockIdx.x / u.n_groups;
  const UmmaTile tile = u.tiles[t];
  const StreamCall sc = a.uniform;
  const uint32_t nt = u.nt;
  const uint32_t tap_chunk = 3 * nt * 16;
  const uint32_t tap_stage = kStageChunks * tap_chunk;
  const uint32_t stage_bytes = kXStageBytes + tap_stage;
  const uint32_t n_chunks = 2 * u.ksteps;
  const uint32_t n_iters = (n_chunks + kStageChunks - 1) / kStageChunks;
  const uint32_t S = u.stages;
  // alignment every input row start shares (16-byte items start at multiples of 16 B in a row)
  const uint32_t row_bits = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(a.in)) |
                            (static_cast<uint32_t>(a.in_stride) * 2u);
  const int in_align = (row_bits & 15u) == 0 ? 16 : (row_bits & 7u) == 0 ? 8 : (row_bits & 3u) == 0 ? 4 : 2;

  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], kConvThreads + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(&tmem_slot, u.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < kConvWarps) {
    // ================= converters: PCM -> byte planes in UMMA layout =================
    constexpr int kTasks = 2 / CH;    // (stream, chunk) tasks per thread per stage
    constexpr int kItems = 2 * CH;    // 16-byte items per task (16 frames)
    constexpr int FPI = 8 / CH;       // frames per item
    uint4 raw[kTasks][kItems];
    auto fetch = [&](uint32_t it) {
#pragma unroll
      for (int q = 0; q < kTasks; ++q) {
        const int id = tid + kConvThreads * q;
        const int sl = id % kStreams, j = id / kStreams;
        const int f0 = tile.kf0 + static_cast<int>((it * kStageChunks + j) * kUmmaChunkFrames);
#pragma unroll
        for (int i = 0; i < kItems; ++i)
          raw[q][i] = fetch_raw16<CH>(a, sc, g * kStreams + sl, f0 + i * FPI, in_align);
      }
    };
    auto convert_store = [&](uint8_t *xs) {
#pragma unroll
      for (int q = 0; q < kTasks; ++q) {
        const int id = tid + kConvThreads * q;
        const int sl = id % kStreams, j = id / kStreams;
        uint8_t *base = xs + j * kChunkBytesX;
        if (CH == 1) {
          // word = (x[2i+1] << 16) | x[2i]: bytes lo0 hi0 lo1 hi1
          uint4 lo, hi;
          lo.x = __byte_perm(raw[q][0].x, raw[q][0].y, 0x6420);
          lo.y = __byte_perm(raw[q][0].z, raw[q][0].w, 0x6420);
          lo.z = __byte_perm(raw[q][1].x, raw[q][1].y, 0x6420);
          lo.w = __byte_perm(raw[q][1].z, raw[q][1].w, 0x6420);
          hi.x = __byte_perm(raw[q][0].x, raw[q][0].y, 0x7531);
          hi.y = __byte_perm(raw[q][0].z, raw[q][0].w, 0x7531);
          hi.z = __byte_perm(raw[q][1].x, raw[q][1].y, 0x7531);
          hi.w = __byte_perm(raw[q][1].z, raw[q][1].w, 0x7531);
          *reinterpret_cast<uint4 *>(base + sl * 16) = hi;
          *reinterpret_cast<uint4 *>(base + kXPlaneBytes + sl * 16) = lo;
        } else {
          // word f = (R_f << 16) | L_f. Four frames (one item) -> one word of each plane/channel.
          uint32_t llo[4], lhi[4], rlo[4], rhi[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 w = raw[q][i];
            const uint32_t ul = __byte_perm(w.x, w.y, 0x5140), vl = __byte_perm(w.z, w.w, 0x5140);
            const uint32_t ur = __byte_perm(w.x, w.y, 0x7362), vr = __byte_perm(w.z, w.w, 0x7362);
            llo[i] = __byte_perm(ul, vl, 0x5410);
            lhi[i] = __byte_perm(ul, vl, 0x7632);
            rlo[i] = __byte_perm(ur, vr, 0x5410);
            rhi[i] = __byte_perm(ur, vr, 0x7632);
          }
          // rows: left channel of stream sl -> row sl, right channel -> row 64 + sl
          *reinterpret_cast<uint4 *>(base + sl * 16) = make_uint4(lhi[0], lhi[1], lhi[2], lhi[3]);
          *reinterpret_cast<uint4 *>(base + (64 + sl) * 16) = make_uint4(rhi[0], rhi[1], rhi[2], rhi[3]);
          *reinterpret_cast<uint4 *>(base + kXPlaneBytes + sl * 16) = make_uint4(llo[0], llo[1], llo[2], llo[3]);
          *reinterpret_cast<uint4 *>(base + kXPlaneBytes + (64 + sl) * 16) =
              make_uint4(rlo[0], rlo[1], rlo[2], rlo[3]);
        }
      }
    };

    fetch(0);
    for (uint32_t it = 0; it < n_iters; ++it) {
      const uint32_t slot = it % S, par = (it / S) & 1u;
      mbar_wait(&empty_bar[slot], par ^ 1u);
      convert_store(smem + slot * stage_bytes);
      if (it + 1 < n_iters) fetch(it + 1);
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[slot]);
    }

    // ================= epilogue =================
    mbar_wait(&acc_bar, 0);
    tc_fence_after_sync();
    const uint32_t n_valid = min(nt, sc.n_out - tile.m0);
    const uint32_t row = (warp & 3) * 32 + lane;  // TMEM lane
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t pitch_w = (CH == 2 ? nt : nt / 2) + 1;  // words per staged row (odd)
    uint16_t *stage16 = reinterpret_cast<uint16_t *>(smem);
    const long long half_ulp = 1ll << (u.shift - 1);
    for (uint32_t cg = warp >> 2; cg * 16 < n_valid; cg += 2) {
      uint32_t p0[16], p1[16], p2[16], p3[16];
      tmem_ld16(lane_addr + cg * 16, p0);
      tmem_ld16(lane_addr + nt + cg * 16, p1);
      tmem_ld16(lane_addr + 2 * nt + cg * 16, p2);
      tmem_ld16(lane_addr + 3 * nt + cg * 16, p3);
      tmem_ld_wait();
      int r16[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        long long v = static_cast<long long>(static_cast<int>(p3[i]));
        v += static_cast<long long>(static_cast<int>(p2[i])) << 8;
        v += static_cast<long long>(static_cast<int>(p1[i])) << 16;
        v += static_cast<long long>(static_cast<int>(p0[i])) << 24;
        // floor(y + 1/2) with y = v * 2^-shift, then saturate (arch.h:208-209)
        const long long r = (v + half_ulp) >> u.shift;
        r16[i] = static_cast<int>(max(-32768ll, min(32767ll, r)));
      }
      if (CH == 2) {
        const uint32_t sl = row & 63, c = row >> 6;
        uint16_t *dst = stage16 + (sl * pitch_w) * 2 + (cg * 16) * 2 + c;
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[2 * i] = static_cast<uint16_t>(r16[i]);
      } else {
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem) + row * pitch_w + cg * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = (static_cast<uint32_t>(r16[2 * i]) & 0xffffu) | (static_cast<uint32_t>(r16[2 * i + 1]) << 16);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");
    // staged rows -> global, coalesced along the stream's interleaved output
    const uint32_t *stage32 = reinterpret_cast<const uint32_t *>(smem);
    for (uint32_t r = warp; r < static_cast<uint32_t>(kStreams); r += kConvWarps) {
      const uint32_t s = g * kStreams + r;
      if (s >= a.n_streams) break;
      int16_t *dst = a.out + static_cast<size_t>(s) * a.out_stride + static_cast<size_t>(tile.m0) * CH;
      const uint32_t n_elems = n_valid * CH;
      if ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
        for (uint32_t i = lane; i < n_elems / 2; i += 32) reinterpret_cast<uint32_t *>(dst)[i] = stage32[r * pitch_w + i];
        if ((n_elems & 1u) && lane == 0) dst[n_elems - 1] = static_cast<int16_t>(stage16[r * pitch_w * 2 + n_elems - 1]);
      } else {
        for (uint32_t i = lane; i < n_elems; i += 32) dst[i] = static_cast<int16_t>(stage16[r * pitch_w * 2 + i]);
      }
    }
  } else if (warp == kTmaWarp) {
    // ================= tap tiles: one bulk copy per stage =================
    if (lane == 0) {
      const int8_t *src = u.pool + static_cast<size_t>(tile.slot) * u.tile_bytes;
      for (uint32_t it = 0; it < n_iters; ++it) {
        const uint32_t slot = it % S, par = (it / S) & 1u;
        mbar_wait(&empty_bar[slot], par ^ 1u);
        const uint32_t chunks_here = min(static_cast<uint32_t>(kStageChunks), n_chunks - it * kStageChunks);
        const uint32_t bytes = chunks_here * tap_chunk;
        mbar_arrive_expect_tx(&full_bar[slot], bytes);
        bulk_g2s(smem + slot * stage_bytes + kXStageBytes, src + static_cast<size_t>(it) * tap_stage, bytes,
                 &full_bar[slot]);
      }
    }
    __syncwarp();
  } else {
    // ================= MMA issue =================
    if (lane == 0) {
      for (uint32_t it = 0; it < n_iters; ++it) {
        const uint32_t slot = it % S, par = (it / S) & 1u;
        mbar_wait(&full_bar[slot], par);
        tc_fence_after_sync();
        const uint32_t xs = smem_u32(smem + slot * stage_bytes);
        const uint32_t ts = xs + kXStageBytes;
        const uint32_t ks_here = min(static_cast<uint32_t>(kStageChunks), n_chunks - it * kStageChunks) / 2;
        for (uint32_t ks = 0; ks < ks_here; ++ks) {
          const uint64_t a_hi = umma_smem_desc(xs + ks * 2 * kChunkBytesX, kChunkBytesX, 128);
          const uint64_t a_lo = umma_smem_desc(xs + kXPlaneBytes + ks * 2 * kChunkBytesX, kChunkBytesX, 128);
          const uint32_t b = ts + ks * 2 * tap_chunk;
          const bool first = (it == 0 && ks == 0);
          issue_rows(tmem, a_hi, true, b, tap_chunk, 0, 3 * nt, 0, first ? 0u : 1u);
          if (first) {
            // columns [nt,3nt) already hold hi*B: accumulate; columns [3nt,4nt) are fresh
            issue_rows(tmem, a_lo, false, b, tap_chunk, 0, 2 * nt, nt, 1u);
            issue_rows(tmem, a_lo, false, b, tap_chunk, 2 * nt, nt, 3 * nt, 0u);
          } else {
            issue_rows(tmem, a_lo, false, b, tap_chunk, 0, 3 * nt, nt, 1u);
          }
        }
        umma_commit(&empty_bar[slot]);
      }
      umma_commit(&acc_bar);
    }
    __syncwarp();
  }

  // ---- this CTA's slice of the history slide (resample.c:898-899) and the new position ----
  // stream sl of the group is handled by the tile CTA with t == sl % n_tiles; new history
  // element e = element consumed*CH + e of (old history || input)
  {
    const uint32_t hist_elems = a.hist_frames * CH;
    const size_t shift = static_cast<size_t>(sc.consumed) * CH;
    const int vshift = (shift % 8 == 0) ? 8 : (shift % 4 == 0) ? 4 : (shift % 2 == 0) ? 2 : 1;
    const int vw = min(vshift, in_align / 2);
    for (uint32_t sl = t; sl < static_cast<uint32_t>(kStreams); sl += u.n_tiles) {
      const uint32_t s = g * kStreams + sl;
      if (s >= a.n_streams) break;
      for (uint32_t e = tid * vw; e < hist_elems; e += kThreads * vw) {
        const size_t src = shift + e;
        const int16_t *p = (src < hist_elems) ? a.hist_src + static_cast<size_t>(s) * a.hist_stride + src
                                              : a.in + static_cast<size_t>(s) * a.in_stride + (src - hist_elems);
        int16_t *d = a.hist_dst + static_cast<size_t>(s) * a.hist_stride + e;
        if (vw == 8) *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(p);
        else if (vw == 4) *reinterpret_cast<uint2 *>(d) = *reinterpret_cast<const uint2 *>(p);
        else if (vw == 2) *reinterpret_cast<uint32_t *>(d) = *reinterpret_cast<const uint32_t *>(p);
        else *d = *p;
      }
      if (tid == 0) {
        a.last_sample[s] = sc.ls1;
        a.samp_frac[s] = sc.frac1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, u.tmem_cols);
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

// Per-batch state of the tensor kernel: fixed-point taps in HBM, the pool of tap tiles keyed
// by (first phase, K-origin offset), and the tile list of the last planned call geometry.
struct UmmaContext {
  FilterSpec spec;
  uint32_t channels = 0;
  int sm_count = 148;
  FixedTaps ft;
  int32_t *d_h = nullptr;
  // geometry (changes only when the tile width changes)
  uint32_t nt = 0, ksteps = 0, tile_bytes = 0, stages = 0, tmem_cols = 0, smem_bytes = 0;
  int8_t *d_pool = nullptr;
  size_t pool_cap = 0;  // tiles
  std::unordered_map<uint64_t, uint32_t> slot_of;
  UmmaTile *d_tiles = nullptr;
  size_t tiles_cap = 0;
  uint32_t n_tiles = 0;
  uint32_t *d_jobs = nullptr;
  size_t jobs_cap = 0;
  // memo of the planned geometry
  bool memo = false;
  int32_t m_ls0 = 0;
  uint32_t m_frac0 = 0, m_n_out = 0, m_hist_frames = 0, m_groups = 0;
};

namespace {

constexpr size_t kMaxPoolBytes = 256ull << 20;

uint32_t forced_nt() {
  static const uint32_t v = [] {
    const char *e = getenv("SPXB_UMMA_NT");
    return e ? static_cast<uint32_t>(atoi(e)) : 0u;
  }();
  return v;
}

// Tile width: fewest estimated cycles for the whole grid. A tile costs its MMA floor
// (3*nt cycles per 32-frame K step: two MMAs of N = 3nt at N/2 cycles each) plus a fixed
// prologue/epilogue; the grid runs in waves of one CTA per SM.
uint32_t pick_nt(const UmmaContext &c, uint32_t n_groups, uint32_t n_out) {
  if (forced_nt() >= 16 && forced_nt() <= 128 && forced_nt() % 16 == 0) return forced_nt();
  uint32_t best = 16;
  double best_cost = 1e300;
  for (uint32_t nt = 16; nt <= 128; nt += 16) {
    const uint32_t tiles = (n_out + nt - 1) / nt;
    const double ctas = static_cast<double>(tiles) * n_groups;
    const double waves = std::ceil(ctas / c.sm_count);
    const uint32_t ks = umma_ksteps(c.spec.taps, c.spec.num, c.spec.den, nt);
    const double tile_cycles = 3000.0 + 3.0 * nt * ks + 10.0 * nt;
    const double cost = waves * tile_cycles;
    if (cost < best_cost) {
      best_cost = cost;
      best = nt;
    }
  }
  return best;
}

void drop_pool(UmmaContext *c) {
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
    cudaFuncSetAttribute(umma_fir_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(umma_fir_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
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
  const uint32_t n_groups = (a.n_streams * a.channels + kUmmaRows - 1) / kUmmaRows;
  const StreamCall &sc = a.uniform;
  if (c->memo && c->m_ls0 == sc.ls0 && c->m_frac0 == sc.frac0 && c->m_n_out == sc.n_out &&
      c->m_hist_frames == a.hist_frames && c->m_groups == n_groups)
    return true;  // steady state: same tiles as the previous call

  const uint32_t nt = pick_nt(*c, n_groups, sc.n_out);
  if (nt != c->nt) {
    drop_pool(c);
    c->nt = nt;
    c->ksteps = umma_ksteps(c->spec.taps, c->spec.num, c->spec.den, nt);
    c->tile_bytes = 2 * c->ksteps * 3 * nt * 16;
    c->tmem_cols = pow2_cols(4 * nt);
    const uint32_t stage_bytes = kXStageBytes + kStageChunks * 3 * nt * 16;
    const uint32_t n_iters = (2 * c->ksteps + kStageChunks - 1) / kStageChunks;
    uint32_t stages = std::min<uint32_t>(kMaxStages, kMaxSmem / stage_bytes);
    stages = std::max(1u, std::min(stages, n_iters));
    // the epilogue stages the output tile over the (drained) ring
    const uint32_t out_bytes = (a.channels == 2 ? 64u * (nt + 1) : 128u * (nt / 2 + 1)) * 4u;
    while (stages * stage_bytes < out_bytes) ++stages;
    c->stages = stages;
    c->smem_bytes = stages * stage_bytes;
    if (c->smem_bytes > kMaxSmem) return false;
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
  if (next_slot > c->pool_cap) {
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
    const uint32_t cells = 2 * c->ksteps * 3 * nt;
    const dim3 grid((cells + 255) / 256, static_cast<unsigned>(n_jobs));
    build_tap_tiles_kernel<<<grid, 256, 0, stream>>>(c->d_h, c->spec.num, c->spec.den, c->spec.taps, nt, c->ksteps,
                                                     c->d_jobs, c->d_pool, c->tile_bytes);
    if ((*err = cudaGetLastError()) != cudaSuccess) return false;
    for (auto &kv : fresh) c->slot_of.emplace(kv.first, kv.second);
  }
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
  c->n_tiles = static_cast<uint32_t>(tiles.size());
  c->memo = true;
  c->m_ls0 = sc.ls0;
  c->m_frac0 = sc.frac0;
  c->m_n_out = sc.n_out;
  c->m_hist_frames = a.hist_frames;
  c->m_groups = n_groups;
  return true;
}

cudaError_t launch_umma(UmmaContext *c, const CallArgs &a, cudaStream_t stream, uint32_t *launches) {
  UmmaArgs u;
  u.tiles = c->d_tiles;
  u.n_tiles = c->n_tiles;
  u.n_groups = c->m_groups;
  u.pool = c->d_pool;
  u.tile_bytes = c->tile_bytes;
  u.nt = c->nt;
  u.ksteps = c->ksteps;
  u.stages = c->stages;
  u.tmem_cols = c->tmem_cols;
  u.shift = c->ft.shift;
  const uint64_t grid = static_cast<uint64_t>(u.n_groups) * u.n_tiles;
  if (grid == 0 || grid > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  if (a.channels == 2)
    umma_fir_kernel<2><<<static_cast<unsigned>(grid), kThreads, c->smem_bytes, stream>>>(a, u);
  else
    umma_fir_kernel<1><<<static_cast<unsigned>(grid), kThreads, c->smem_bytes, stream>>>(a, u);
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && launches) *launches += 1;
  return e;
}

}  // namespace spxb
