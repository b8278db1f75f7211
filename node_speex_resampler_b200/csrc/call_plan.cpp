// call_plan.cpp -- see call_plan.h.
#include "call_plan.h"

#include <algorithm>

namespace spxb {

namespace {

// number of outputs whose window start stays below `limit`, starting from (ls, frac):
// smallest m with ls + floor((frac + m*num)/den) >= limit
inline uint64_t outputs_before(int64_t ls, uint64_t frac, uint64_t limit, uint64_t num,
                               uint64_t den) {
  if (ls >= static_cast<int64_t>(limit)) return 0;
  const uint64_t span = (limit - static_cast<uint64_t>(ls)) * den - frac;  // > 0: frac < den
  return (span + num - 1) / num;
}

}  // namespace

CallPlan plan_call(uint32_t num, uint32_t den, StreamPos pos, uint32_t n_in, uint32_t out_cap,
                   uint32_t out_block, uint32_t in_block) {
  CallPlan plan;
  int64_t ls = pos.last_sample;
  uint64_t frac = pos.samp_frac_num;
  uint64_t left_in = n_in, left_out = out_cap;

  // Fast path: capacity can never bind -> every block commits all of its input, so the
  // call consumes everything and the position simply advances by the output count.
  const uint64_t possible = outputs_before(ls, frac, n_in, num, den);
  if (possible < out_cap) {
    const uint64_t adv = frac + possible * num;
    plan.n_out = static_cast<uint32_t>(possible);
    plan.consumed = n_in;
    plan.next.last_sample = static_cast<int32_t>(ls + static_cast<int64_t>(adv / den) - n_in);
    plan.next.samp_frac_num = static_cast<uint32_t>(adv % den);
    return plan;
  }

  // resample.c:988 `while (ilen && olen)`
  while (left_in != 0 && left_out != 0) {
    const uint64_t take = std::min<uint64_t>(left_in, in_block);
    const uint64_t room = std::min<uint64_t>(left_out, out_block);
    const uint64_t made = std::min(room, outputs_before(ls, frac, take, num, den));
    const uint64_t adv = frac + made * num;
    ls += static_cast<int64_t>(adv / den);
    frac = adv % den;
    // resample.c:891-894
    const uint64_t used = (ls < static_cast<int64_t>(take)) ? static_cast<uint64_t>(ls) : take;
    ls -= static_cast<int64_t>(used);
    left_in -= used;
    left_out -= made;
  }
  plan.n_out = static_cast<uint32_t>(out_cap - left_out);
  plan.consumed = static_cast<uint32_t>(n_in - left_in);
  plan.next.last_sample = static_cast<int32_t>(ls);
  plan.next.samp_frac_num = static_cast<uint32_t>(frac);
  return plan;
}

MagicPlan plan_call_magic(uint32_t num, uint32_t den, StreamPos pos, uint32_t magic, uint32_t n_in,
                          uint32_t out_cap, bool float_entry, uint32_t in_block) {
  int64_t ls = pos.last_sample;
  uint64_t frac = pos.samp_frac_num;
  uint64_t left_in = n_in, left_out = out_cap, m = magic;
  // one block of `len` input frames with room for `room` outputs (process_native, resample.c:878-902)
  auto block = [&](uint64_t len, uint64_t room, uint64_t *made_out) {
    const uint64_t made = std::min(room, outputs_before(ls, frac, len, num, den));
    const uint64_t adv = frac + made * num;
    ls += static_cast<int64_t>(adv / den);
    frac = adv % den;
    const uint64_t used = (ls < static_cast<int64_t>(len)) ? static_cast<uint64_t>(ls) : len;
    ls -= static_cast<int64_t>(used);
    *made_out = made;
    return used;
  };
  if (float_entry) {  // resample.c:940-962
    if (m) {
      uint64_t made = 0;
      m -= block(m, left_out, &made);
      left_out -= made;
    }
    if (!m) {
      while (left_in != 0 && left_out != 0) {
        uint64_t made = 0;
        left_in -= block(std::min<uint64_t>(left_in, in_block), left_out, &made);
        left_out -= made;
      }
    }
  } else {  // resample.c:988-1029
    while (left_in != 0 && left_out != 0) {
      uint64_t room = std::min<uint64_t>(left_out, kOutBlock);
      if (m) {
        uint64_t made = 0;
        m -= block(m, room, &made);
        room -= made;
        left_out -= made;
      }
      if (!m) {
        uint64_t made = 0;
        left_in -= block(std::min<uint64_t>(left_in, in_block), room, &made);
        left_out -= made;
      }
    }
  }
  MagicPlan r;
  r.plan.n_out = static_cast<uint32_t>(out_cap - left_out);
  r.plan.consumed = static_cast<uint32_t>(n_in - left_in);
  r.plan.next.last_sample = static_cast<int32_t>(ls);
  r.plan.next.samp_frac_num = static_cast<uint32_t>(frac);
  r.magic_used = static_cast<uint32_t>(magic - m);
  return r;
}

}  // namespace spxb
