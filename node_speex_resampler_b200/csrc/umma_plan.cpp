// umma_plan.cpp -- see umma_plan.h.
#include "umma_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace spxb {

namespace {

// per-phase tap in f64: the direct table entry, or the cubic blend of the four neighbouring
// prototype taps (deps/speex/resample.c:454-476 by linearity), not yet rounded
double phase_tap_f64(const FilterSpec &s, const std::vector<float> &ref, uint32_t phase, uint32_t j,
                     const float w[4], uint32_t cell) {
  if (s.direct) return static_cast<double>(ref[static_cast<size_t>(phase) * s.taps + j]);
  const float *c = ref.data() + 4 + s.oversample - cell - 2 + static_cast<size_t>(j) * s.oversample;
  return static_cast<double>(w[0]) * c[0] + static_cast<double>(w[1]) * c[1] +
         static_cast<double>(w[2]) * c[2] + static_cast<double>(w[3]) * c[3];
}

constexpr int32_t kMaxFixed = 127 * 65536 + 127 * 256 + 127;  // 8355711

}  // namespace

bool build_fixed_taps(const FilterSpec &s, const std::vector<float> &ref, FixedTaps *out) {
  const size_t N = s.taps;
  std::vector<double> v(static_cast<size_t>(s.den) * N);
  double vmax = 0.0;
  for (uint32_t phase = 0; phase < s.den; ++phase) {
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t cell = 0;
    if (!s.direct) {
      const uint32_t scaled = phase * s.oversample;  // uint32 like resample.c:458
      cell = scaled / s.den;
      cubic_weights(static_cast<float>(scaled % s.den) / s.den, w);
    }
    for (uint32_t j = 0; j < N; ++j) {
      const double t = phase_tap_f64(s, ref, phase, j, w, cell);
      v[static_cast<size_t>(phase) * N + j] = t;
      vmax = std::max(vmax, std::fabs(t));
    }
  }
  if (!(vmax > 0.0) || !std::isfinite(vmax)) return false;
  int shift = 30;
  while (shift > 0 && std::ldexp(vmax, shift) > static_cast<double>(kMaxFixed) - 1.0) --shift;
  if (std::ldexp(vmax, shift) > static_cast<double>(kMaxFixed) - 1.0) return false;
  out->shift = shift;
  out->h.resize(v.size());
  for (size_t i = 0; i < v.size(); ++i) out->h[i] = static_cast<int32_t>(std::llround(std::ldexp(v[i], shift)));
  return true;
}

uint32_t umma_ksteps(uint32_t taps, uint32_t num, uint32_t den, uint32_t nt) {
  // last output of a tile starts at most floor((den-1 + (nt-1)*num)/den) frames after the
  // first; the first starts at most 15 frames after the K origin
  const uint64_t adv = (static_cast<uint64_t>(den) - 1 + static_cast<uint64_t>(nt - 1) * num) / den;
  const uint64_t frames = (kUmmaChunkFrames - 1) + adv + taps;
  return static_cast<uint32_t>((frames + kUmmaStepFrames - 1) / kUmmaStepFrames);
}

void plan_umma_tiles(uint32_t num, uint32_t den, uint32_t taps, uint32_t hist_frames, int32_t ls0,
                     uint32_t frac0, uint32_t n_out, uint32_t nt, std::vector<UmmaTile> *tiles,
                     std::vector<UmmaTileKey> *keys) {
  tiles->clear();
  keys->clear();
  for (uint32_t m0 = 0; m0 < n_out; m0 += nt) {
    const uint64_t t = static_cast<uint64_t>(frac0) + static_cast<uint64_t>(m0) * num;
    const int64_t q0 = static_cast<int64_t>(ls0) - (static_cast<int64_t>(taps) - 1) + static_cast<int64_t>(t / den);
    // chunk boundaries are counted from the first frame of the history buffer
    const int64_t from_hist = q0 + hist_frames;  // >= 0 because ls0 >= 0 and hist_frames >= N-1
    const int64_t kf = from_hist - (from_hist % kUmmaChunkFrames) - hist_frames;
    UmmaTile tl;
    tl.m0 = m0;
    tl.kf0 = static_cast<int32_t>(kf);
    tl.slot = 0;
    tiles->push_back(tl);
    UmmaTileKey k;
    k.phase0 = static_cast<uint32_t>(t % den);
    k.delta = static_cast<uint32_t>(q0 - kf);
    keys->push_back(k);
  }
}

void fill_tap_tile_host(const FixedTaps &ft, uint32_t num, uint32_t den, uint32_t taps, uint32_t nt,
                        uint32_t ksteps, UmmaTileKey key, int8_t *dst) {
  const uint32_t chunks = 2 * ksteps, rows = 3 * nt;
  for (uint32_t c = 0; c < chunks; ++c)
    for (uint32_t r = 0; r < rows; ++r) {
      const uint32_t n = r % nt, digit = 2 - r / nt;
      const uint64_t t = static_cast<uint64_t>(key.phase0) + static_cast<uint64_t>(n) * num;
      const uint32_t phase = static_cast<uint32_t>(t % den);
      const int64_t first = static_cast<int64_t>(key.delta) + static_cast<int64_t>(t / den);
      for (uint32_t e = 0; e < 16; ++e) {
        const int64_t j = static_cast<int64_t>(c) * 16 + e - first;
        int d[3] = {0, 0, 0};
        if (j >= 0 && j < static_cast<int64_t>(taps))
          split_digits(ft.h[static_cast<size_t>(phase) * taps + static_cast<size_t>(j)], &d[2], &d[1], &d[0]);
        dst[(static_cast<size_t>(c) * rows + r) * 16 + e] = static_cast<int8_t>(d[digit]);
      }
    }
}

bool build_packed_plan(const FilterSpec &s, const FixedTaps &ft, uint32_t nt, UmmaPackedPlan *out) {
  const uint32_t ksteps = umma_ksteps(s.taps, s.num, s.den, nt);
  if (ksteps == 0 || ksteps > kUmmaMaxKsteps || nt % 16 != 0 || nt == 0 || nt > 128) return false;
  const uint32_t N = s.taps, nb = nt / 16;
  // tap indices at which each digit is non-zero for some phase
  int64_t jlo[3] = {N, N, N}, jhi[3] = {-1, -1, -1};
  for (uint32_t phase = 0; phase < s.den; ++phase)
    for (uint32_t j = 0; j < N; ++j) {
      int d[3];
      split_digits(ft.h[static_cast<size_t>(phase) * N + j], &d[0], &d[1], &d[2]);  // d2, d1, d0
      for (int i = 0; i < 3; ++i)
        if (d[i] != 0) {
          jlo[i] = std::min<int64_t>(jlo[i], j);
          jhi[i] = std::max<int64_t>(jhi[i], j);
        }
    }
  out->nt = nt;
  out->ksteps = ksteps;
  out->k.assign(ksteps, UmmaKStep{});
  uint32_t off16 = 0;
  for (uint32_t k = 0; k < ksteps; ++k) {
    UmmaKStep &ks = out->k[k];
    for (int i = 0; i < 3; ++i) {
      uint32_t b0 = nb, b1 = 0;
      static const bool dense = [] {  // experiment: store every block (one fused MMA per plane when 3nt <= 256)
        const char *e = getenv("SPXB_UMMA_DENSE");
        return e && atoi(e) != 0;
      }();
      if (k == 0 || dense) {
        b0 = 0;
        b1 = nb;
      } else if (jhi[i] >= 0) {
        // column n of a tile with key (phase0, delta) starts its window at frame
        // first(n) = delta + floor((phase0 + n*num)/den) of the tile's K axis, between
        // floor(n*num/den) and 15 + floor((den-1 + n*num)/den); digit i of column n touches
        // frames first(n) + [jlo, jhi]; K step k covers frames [32k, 32k+32)
        for (uint32_t n = 0; n < nt; ++n) {
          const int64_t fmin = static_cast<int64_t>(static_cast<uint64_t>(n) * s.num / s.den);
          const int64_t fmax = 15 + static_cast<int64_t>((static_cast<uint64_t>(s.den) - 1 + static_cast<uint64_t>(n) * s.num) / s.den);
          const bool hit = fmin + jlo[i] <= static_cast<int64_t>(32 * k + 31) && fmax + jhi[i] >= static_cast<int64_t>(32 * k);
          if (hit) {
            b0 = std::min(b0, n / 16);
            b1 = std::max(b1, n / 16 + 1);
          }
        }
      }
      if (b1 <= b0) b0 = b1 = 0;
      ks.b0[i] = static_cast<uint8_t>(b0);
      ks.b1[i] = static_cast<uint8_t>(b1);
    }
    // rows of a chunk: d2 blocks, d1 blocks, d0 blocks; entries fuse neighbouring digits whose runs
    // are contiguous in both the B rows (always) and the D columns (digit i ends at nt and digit
    // i+1 starts at 0), up to 256 columns
    uint32_t row = 0;
    ks.n_ent = 0;
    bool open = false;
    for (int i = 0; i < 3; ++i) {
      const uint32_t n = 16u * (ks.b1[i] - ks.b0[i]);
      if (n == 0) {
        open = false;
        continue;
      }
      const uint32_t dcol = static_cast<uint32_t>(i) * nt + 16u * ks.b0[i];
      UmmaKStep::Entry *last = ks.n_ent ? &ks.ent[ks.n_ent - 1] : nullptr;
      if (open && last && ks.b0[i] == 0 && last->dcol + last->n == dcol && last->n + n <= 256) {
        last->n = static_cast<uint16_t>(last->n + n);
      } else {
        ks.ent[ks.n_ent].row = static_cast<uint16_t>(row);
        ks.ent[ks.n_ent].n = static_cast<uint16_t>(n);
        ks.ent[ks.n_ent].dcol = static_cast<uint16_t>(dcol);
        ks.n_ent += 1;
      }
      open = ks.b1[i] == nb;
      row += n;
    }
    ks.rows = static_cast<uint16_t>(row);
    ks.off16 = off16;
    off16 += 2 * row;
  }
  out->tile_bytes = off16 * 16;
  return true;
}

void fill_tap_tile_packed_host(const FixedTaps &ft, uint32_t num, uint32_t den, uint32_t taps,
                               const UmmaPackedPlan &plan, UmmaTileKey key, int8_t *dst) {
  for (uint32_t k = 0; k < plan.ksteps; ++k) {
    const UmmaKStep &ks = plan.k[k];
    for (uint32_t half = 0; half < 2; ++half) {
      uint32_t row = 0;
      for (int i = 0; i < 3; ++i)
        for (uint32_t n = 16u * ks.b0[i]; n < 16u * ks.b1[i]; ++n, ++row) {
          const uint64_t t = static_cast<uint64_t>(key.phase0) + static_cast<uint64_t>(n) * num;
          const uint32_t phase = static_cast<uint32_t>(t % den);
          const int64_t first = static_cast<int64_t>(key.delta) + static_cast<int64_t>(t / den);
          int8_t *cell = dst + (static_cast<size_t>(ks.off16) + static_cast<size_t>(half) * ks.rows + row) * 16;
          for (uint32_t e = 0; e < 16; ++e) {
            const int64_t j = static_cast<int64_t>(32 * k + 16 * half + e) - first;
            int d[3] = {0, 0, 0};
            if (j >= 0 && j < static_cast<int64_t>(taps))
              split_digits(ft.h[static_cast<size_t>(phase) * taps + static_cast<size_t>(j)], &d[0], &d[1], &d[2]);
            cell[e] = static_cast<int8_t>(d[i]);
          }
        }
    }
  }
}

}  // namespace spxb
