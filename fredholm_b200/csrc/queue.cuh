// Warp-level queue primitives: dynamic work fetch for persistent kernels and
// warp-aggregated append (one atomic per warp, ballot + popc for the offsets).
#pragma once
#include <cstdint>

#include "vecmath.cuh"

namespace frd
{

FR_D uint32_t lane_id() { return threadIdx.x & 31u; }

// Every lane of the warp must call this.  Grabs the next 32 consecutive items;
// returns false for the whole warp when the queue is exhausted.  `item` may be
// >= n for the tail lanes of the last batch.
FR_D bool fetch_batch(uint32_t* cursor, uint32_t n, uint32_t& item)
{
  uint32_t base = 0;
  if (lane_id() == 0) base = atomicAdd(cursor, 32u);
  base = __shfl_sync(0xffffffffu, base, 0);
  item = base + lane_id();
  return base < n;
}

// Every lane of the warp must call this.  Returns the slot reserved for this lane
// (only meaningful when `want`).
FR_D uint32_t queue_reserve(uint32_t* counter, bool want)
{
  const uint32_t mask = __ballot_sync(0xffffffffu, want);
  if (mask == 0u) return 0u;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if ((int)lane_id() == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(mask & ((1u << lane_id()) - 1u));
}

}  // namespace frd
