// Quantisation of child boxes into a CWBVH node frame, shared by the builder's collapse (bvh_build.cu) and the
// instance-tree refit (accel.cu): conservative -- lower bounds round down, upper bounds round up.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace frd
{

// biased exponent e such that extent <= 255 * 2^(e-127)
__device__ __forceinline__ uint32_t grid_exponent(float extent)
{
  const float step = __fdiv_ru(extent, 255.0f);
  uint32_t e = (__float_as_uint(step) + 0x007fffffu) >> 23;
  return min(max(e, 1u), 254u);
}

// 1 / 2^(e-127) = 2^(127-e): biased exponent 254 - e
__device__ __forceinline__ float grid_inverse_step(uint32_t e) { return __uint_as_float((254u - e) << 23); }

__device__ __forceinline__ uint8_t quantize_lo(float v, float origin, float inv_step)
{
  return (uint8_t)fminf(fmaxf(floorf(__fmul_rd(__fsub_rd(v, origin), inv_step)), 0.0f), 255.0f);
}
__device__ __forceinline__ uint8_t quantize_hi(float v, float origin, float inv_step)
{
  return (uint8_t)fminf(fmaxf(ceilf(__fmul_ru(__fsub_ru(v, origin), inv_step)), 0.0f), 255.0f);
}

}  // namespace frd
