// umma_plan.h -- host-side planning for the tensor-core FIR kernel (kernels_umma.cu).
//
// The kernel evaluates the per-phase FIR of deps/speex/resample.c:331-558 as an exact
// integer banded GEMM on the int8 tensor cores:
//   * every per-phase tap (the direct table, or the cubic blend of resample.c:467-476 folded
//     into one tap in f64) is quantised once to a signed 24-bit fixed-point integer
//     h = round(tap * 2^shift) and split into three balanced base-256 digits d2,d1,d0 in
//     [-128,127]  (h = d2*65536 + d1*256 + d0);
//   * every int16 input sample splits exactly into x = hi*256 + lo, hi in [-128,127] (s8),
//     lo in [0,255] (u8);
//   * x*h = hi*d2*2^24 + (hi*d1 + lo*d2)*2^16 + (hi*d0 + lo*d1)*2^8 + lo*d0, four int32
//     accumulators per output, summed over the window by tcgen05.mma.kind::i8 -- no rounding
//     anywhere until the final  floor(y*2^-shift + 1/2)  (WORD2INT, arch.h:208-209).
//
// An output tile is `nt` consecutive outputs of 128 series. Its window starts at frame
// q0 = last_sample - (N-1) + floor((frac + m0*num)/den); the K axis of the GEMM starts at
// the 16-frame boundary kf0 <= q0 (boundaries counted from the start of the history buffer),
// so the tile's banded tap matrix depends only on (phase of its first output, q0 - kf0).
#pragma once

#include <cstdint>
#include <vector>

#include "filter_bank.h"

namespace spxb {

constexpr uint32_t kUmmaChunkFrames = 16;  // frames per 16-byte K chunk of int8 operands
constexpr uint32_t kUmmaStepFrames = 32;   // frames per tcgen05.mma (K = 32 bytes)
constexpr uint32_t kUmmaRows = 128;        // series per tile (UMMA M)

// Per-phase taps as 24-bit fixed point: h[phase*N + j] = round(tap * 2^shift), |h| <= 8355711
// so that the three balanced digits fit int8. Returns false when the filter does not fit
// (degenerate all-zero table).
struct FixedTaps {
  std::vector<int32_t> h;  // [den][N]
  int shift = 0;
};
bool build_fixed_taps(const FilterSpec &spec, const std::vector<float> &ref_table, FixedTaps *out);

// balanced base-256 digits of a fixed-point tap
inline void split_digits(int32_t h, int *d2, int *d1, int *d0) {
  const int32_t lo = ((h + 128) & 255) - 128;
  const int32_t r1 = (h - lo) >> 8;
  const int32_t mid = ((r1 + 128) & 255) - 128;
  *d0 = lo;
  *d1 = mid;
  *d2 = (r1 - mid) >> 8;
}

struct UmmaTile {
  uint32_t m0;    // first output of the tile
  int32_t kf0;    // first frame of the K axis, X~ coordinates (history is f < 0)
  uint32_t slot;  // tap tile in the pool
};

struct UmmaTileKey {
  uint32_t phase0;  // phase of output m0
  uint32_t delta;   // q0 - kf0, in [0, 16)
};

// K steps (of 32 frames) every tile of this geometry runs
uint32_t umma_ksteps(uint32_t taps, uint32_t num, uint32_t den, uint32_t nt);

// Tiles of one uniform call. keys[i] identifies the tap matrix tile i needs.
void plan_umma_tiles(uint32_t num, uint32_t den, uint32_t taps, uint32_t hist_frames, int32_t ls0,
                     uint32_t frac0, uint32_t n_out, uint32_t nt, std::vector<UmmaTile> *tiles,
                     std::vector<UmmaTileKey> *keys);

// Reference (host) fill of one tap tile in the layout the kernel consumes:
// [chunk c < 2*ksteps][row r < 3*nt][16 bytes], row r = digit (2 - r/nt) of output n = r % nt,
// byte e of chunk c = frame k = 16c + e of the tile's K axis; tap index j = k - delta - adv(n).
void fill_tap_tile_host(const FixedTaps &ft, uint32_t num, uint32_t den, uint32_t taps, uint32_t nt,
                        uint32_t ksteps, UmmaTileKey key, int8_t *dst);

// ---------------------------------------------------------------------------------------------
// Packed ("resident") tap tiles: the banded tap matrix of a tile is mostly zeros -- the corners
// outside the band, and, because a windowed sinc decays, the high digit d2 is zero outside the
// main lobe and the middle digit d1 near the filter's ends. A packed tile stores, per K step, only
// the 16-column blocks of each digit that can hold a non-zero for SOME tile key (phase0, delta), so
// that one layout -- one table of MMA segments -- serves every tile of the geometry:
//   K step k:  chunk 0 rows [d2 blocks | d1 blocks | d0 blocks], then chunk 1 rows alike
//              (rows_k rows of 16 bytes per chunk; the two K halves of one MMA are rows_k*16 apart)
// K step 0 is stored whole: its MMAs initialise every accumulator column.
// Each K step runs at most three MMAs per byte plane ("entries": a run of B rows and the D columns
// it accumulates into; neighbouring digits whose block ranges are whole fuse into one entry).
constexpr uint32_t kUmmaMaxKsteps = 64;
constexpr uint32_t kUmmaMaxEntries = 3;

struct UmmaKStep {
  uint32_t off16;    // byte offset of the K step inside the packed tile, / 16
  uint16_t rows;     // rows per chunk (= LBO / 16)
  uint16_t n_ent;
  struct Entry {
    uint16_t row;    // first B row inside the chunk
    uint16_t n;      // columns (multiple of 16, <= 256)
    uint16_t dcol;   // first accumulator column for the hi plane (the lo plane adds nt)
  } ent[kUmmaMaxEntries];
  // block ranges per digit, index 0 = d2, 1 = d1, 2 = d0 (16-column blocks [b0, b1))
  uint8_t b0[3], b1[3];
};

struct UmmaPackedPlan {
  uint32_t nt = 0, ksteps = 0;
  uint32_t tile_bytes = 0;
  std::vector<UmmaKStep> k;  // [ksteps]
};

// false when the geometry is not covered (too many K steps)
bool build_packed_plan(const FilterSpec &spec, const FixedTaps &ft, uint32_t nt, UmmaPackedPlan *out);

// host fill of one packed tile (tests; the device builder in kernels_umma2.cu must agree)
void fill_tap_tile_packed_host(const FixedTaps &ft, uint32_t num, uint32_t den, uint32_t taps,
                               const UmmaPackedPlan &plan, UmmaTileKey key, int8_t *dst);

}  // namespace spxb
