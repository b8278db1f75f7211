// umma_context.h -- per-batch host state of the tensor-core FIR kernels.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <unordered_map>
#include <vector>

#include "filter_bank.h"
#include "umma_plan.h"

namespace spxb {

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
  uint32_t cluster = 1, grid_groups = 0;  // CTAs per cluster; series groups padded to a multiple of it
  // resident: the persistent kernel (kernels_umma2.cu) with packed tap tiles serves this geometry;
  // otherwise the one-tile-per-CTA kernel with dense, streamed tap tiles (kernels_umma.cu)
  bool resident = false;
  uint32_t n_acc = 1;  // accumulator sets of the persistent kernel in TMEM (2 when 8 nt <= 512)
  bool resident_wanted = false;  // what the call that fixed the geometry asked for (kernels_umma2.cu: umma2_covers)
  UmmaPackedPlan packed;
  UmmaKStep *d_kplan = nullptr;  // the packed plan in HBM (tile builder)
  void *d_kdev = nullptr;        // the plan the kernel reads, in HBM (kernels_umma2.cu: Plan2)
  std::vector<uint32_t> recs;       // MMA records of the persistent kernel, 4 words each (kernels_umma2.cu: MmaRec)
  std::vector<uint32_t> stage_off;  // per 64-frame stage: byte offset inside the packed tile (+ end)
  std::vector<uint16_t> stage_rec;  // per stage: first MMA record (+ end)
  int8_t *d_pool = nullptr;
  size_t pool_cap = 0;  // tiles
  std::unordered_map<uint64_t, uint32_t> slot_of;
  UmmaTile *d_tiles = nullptr;
  size_t tiles_cap = 0;
  uint32_t n_tiles = 0;
  std::vector<UmmaTile> h_tiles;  // host copy of the planned tile table (kernel parameters)
  // CUDA-graph support (batch.cu: ring graphs): while `frozen`, planning must not touch the
  // stream or allocate (the stream is being captured); stream_ops counts every such operation,
  // pool_generation changes whenever cached launches would point at stale tap tiles
  bool frozen = false;
  uint64_t stream_ops = 0, pool_generation = 0;
  uint32_t *d_jobs = nullptr;
  size_t jobs_cap = 0;
  unsigned long long *d_trace = nullptr;  // SPXB_UMMA_TRACE=1: timeline of the last launch
  size_t trace_ctas = 0;
  // memo of the planned geometry
  bool memo = false;
  // the tile table / tap pool were (re)written on the stream since the last tensor-kernel launch:
  // that launch goes without the programmatic edge. Sticky until a launch succeeds, so a failed
  // launch cannot make the next one read a pool that is still being built.
  bool fresh_plan = false;
  int32_t m_ls0 = 0;
  uint32_t m_frac0 = 0, m_n_out = 0, m_hist_frames = 0, m_groups = 0;
};

}  // namespace spxb
