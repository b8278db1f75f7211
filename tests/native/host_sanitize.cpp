// host_sanitize.cpp -- the host-side planners (filter bank, call planner, tensor-kernel planner)
// exercised under AddressSanitizer + UndefinedBehaviorSanitizer. Built and run by
// tests/test_sanitizers_cpu.py; prints "ok <checksum>" when every case ran clean.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "call_plan.h"
#include "filter_bank.h"
#include "umma_plan.h"

using namespace spxb;

int main() {
  static const uint32_t cases[][3] = {{24000, 48000, 5},  {24000, 24000, 5},  {24000, 48000, 10}, {44100, 48000, 7},
                                      {44100, 48000, 10}, {44100, 48000, 1},  {44100, 24000, 5},  {24000, 44100, 3},
                                      {48000, 16000, 10}, {96000, 44100, 10}, {44100, 48000, 0},  {8000, 48000, 3},
                                      {48000, 44100, 4},  {16000, 8000, 8},   {44100, 48001, 3},  {8000, 96000, 2},
                                      {1, 1, 4},          {7, 3, 2},          {192000, 8000, 10}};
  uint64_t sum = 0;
  for (const auto &c : cases) {
    FilterSpec spec;
    if (derive_filter_spec(c[0], c[1], static_cast<int>(c[2]), &spec) != 0) {
      std::printf("derive_filter_spec failed for %u %u %u\n", c[0], c[1], c[2]);
      return 1;
    }
    const std::vector<float> table = build_reference_table(spec);
    sum += table.size();
    const uint64_t phase_floats = static_cast<uint64_t>(spec.den) * spec.taps;
    if (phase_floats * 4 <= (64ull << 20)) {
      const std::vector<float> taps = build_phase_taps(spec, table);
      sum += taps.size();
      BandTable band;
      if (build_band_table(spec, taps, 64ull << 20, &band)) sum += band.data.size();
      FixedTaps ft;
      if (spec.taps <= 32768 && build_fixed_taps(spec, table, &ft)) {
        sum += static_cast<uint64_t>(ft.shift);
        for (uint32_t nt : {16u, 80u, 112u, 128u}) {
          const uint32_t ks = umma_ksteps(spec.taps, spec.num, spec.den, nt);
          const uint32_t hist_frames = (spec.taps - 1 + 15) / 16 * 16;
          std::vector<UmmaTile> tiles;
          std::vector<UmmaTileKey> keys;
          plan_umma_tiles(spec.num, spec.den, spec.taps, hist_frames, 3, spec.den / 2, 1000, nt, &tiles, &keys);
          if (!keys.empty() && static_cast<uint64_t>(ks) * nt * 96 <= (8ull << 20)) {
            std::vector<int8_t> tile(static_cast<size_t>(2) * ks * 3 * nt * 16);
            fill_tap_tile_host(ft, spec.num, spec.den, spec.taps, nt, ks, keys.back(), tile.data());
            sum += static_cast<uint8_t>(tile[tile.size() / 2]);
          }
          sum += tiles.size();
        }
      }
    }
    StreamPos pos;
    for (uint32_t k = 0; k < 200; ++k) {
      const uint32_t n_in = (k * 7919u) % 2000u, cap = (k * 104729u) % 3000u;
      const CallPlan a = plan_call(spec.num, spec.den, pos, n_in, cap);
      const CallPlan b = plan_call(spec.num, spec.den, pos, n_in, cap, kOutBlockUnbounded);
      sum += a.n_out + a.consumed + b.n_out;
      pos = a.next;
    }
  }
  std::printf("ok %llu\n", static_cast<unsigned long long>(sum));
  return 0;
}
