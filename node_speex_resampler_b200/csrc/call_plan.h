// call_plan.h -- lengths and next stream position of one resampler call, computed
// without touching samples.
//
// The reference walks each channel through blocks of at most 160 input frames
// (st->buffer_size, resample.c:835/:977/:990) and 1024 output frames
// (FIXED_STACK_ALLOC, resample.c:111/:982-991); inside a block the read position
// advances by num/den per output (resample.c:372-378) and the block commits
// min(block length, position) input frames (resample.c:891-894). Output VALUES depend
// only on absolute positions, but how many frames a call consumes when the output
// capacity binds depends on where the block boundaries fall, so the walk is reproduced
// block by block, each block in closed form.
#pragma once

#include <cstdint>

namespace spxb {

constexpr uint32_t kInBlock = 160;    // resample.c:835
constexpr uint32_t kOutBlock = 1024;  // resample.c:111

struct StreamPos {
  int32_t last_sample = 0;     // index of the next read window relative to unconsumed input
  uint32_t samp_frac_num = 0;  // fractional position in units of 1/den
};

struct CallPlan {
  uint32_t n_out = 0;     // frames written
  uint32_t consumed = 0;  // input frames consumed
  StreamPos next;         // position after the call
};

// Outputs m = 0 .. n_out-1 of the call read the window starting at
//   p(m) = pos.last_sample + floor((pos.samp_frac_num + m*num) / den)
// of X = history(N-1 frames) || input, with phase (pos.samp_frac_num + m*num) % den.
// out_block: output frames one block may produce -- kOutBlock on the int16 entry (its 1024-sample
// stack buffer), unbounded on the float entry, which writes straight into the caller's buffer
// (resample.c:944 `ochunk = olen`).
constexpr uint32_t kOutBlockUnbounded = 0xffffffffu;
// in_block: input frames one block may take -- mem_alloc_size - (filt_len - 1) of the reference
// state: kInBlock for a fresh state, larger once a filter change has shortened the filter (the
// reference never shrinks its memory, resample.c:709-719).
CallPlan plan_call(uint32_t num, uint32_t den, StreamPos pos, uint32_t n_in, uint32_t out_cap,
                   uint32_t out_block = kOutBlock, uint32_t in_block = kInBlock);

// A call on a state that holds `magic` pending samples (left in its memory by a filter change
// that shortened the filter, resample.c:759-776): they are resampled first, as one block of their
// own (speex_resampler_magic, :904-922), sharing the first iteration's output block with the first
// input block on the int16 entry (:993-1016) and taking all the capacity left on the float entry
// (:940-941). plan.consumed counts REAL input frames; magic_used the pending samples consumed.
struct MagicPlan {
  CallPlan plan;
  uint32_t magic_used = 0;
};
MagicPlan plan_call_magic(uint32_t num, uint32_t den, StreamPos pos, uint32_t magic, uint32_t n_in,
                          uint32_t out_cap, bool float_entry, uint32_t in_block);

}  // namespace spxb
