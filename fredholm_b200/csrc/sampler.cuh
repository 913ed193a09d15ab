// Per-path sampler of the wavefront integrator: bit-exact with the reference's
// draw streams, but carried as 5 words instead of the reference's 72-byte
// SamplerState (shared.h:66-96) -- the PCG and blue-noise members there are
// initialised and never drawn (pt.cu:382-398), so they are not carried.
//
//   2-D draws: Correlated Multi-Jittered, 4x4 strata (Kensler 2013), scramble =
//              xxhash32(n_spp/16, pixel, draw#, xxhash32(seed))   cmj.cu:60-80
//   1-D draws: Sobol' with Laine-Karras Owen scrambling (Burley 2020), index =
//              owen(pixel + n_spp*W*H mod 2^32), dimension++ from 1 sobol.cu:10733
//
// All of this is integer arithmetic and must match the reference bit for bit.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "vecmath.cuh"

namespace frd
{

// g_sobol_matrices (1024 dims x 32 columns, one 128-byte line per dimension) comes
// from tables.cuh, which the including .cu file must include first.

// xxhash32 avalanche over 1 / 4 words, https://www.shadertoy.com/view/XlGcRh
// (reference shared.h:282-291, 306-319)
FR_HD uint32_t rotl17(uint32_t h) { return (h << 17) | (h >> 15); }
FR_HD uint32_t xxh_fin(uint32_t h)
{
  h = 2246822519u * (h ^ (h >> 15));
  h = 3266489917u * (h ^ (h >> 13));
  return h ^ (h >> 16);
}
FR_HD uint32_t xxhash32(uint32_t p)
{
  return xxh_fin(668265263u * rotl17(p + 374761393u));
}
FR_HD uint32_t xxhash32(uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
  uint32_t h = w + 374761393u + x * 3266489917u;
  h = 668265263u * rotl17(h);
  h += y * 3266489917u;
  h = 668265263u * rotl17(h);
  h += z * 3266489917u;
  h = 668265263u * rotl17(h);
  return xxh_fin(h);
}

// Kensler's cycle-walking permutation of [0,l) (cmj.cu:12-43)
FR_HD uint32_t kensler_permute(uint32_t i, uint32_t l, uint32_t p)
{
  uint32_t w = l - 1;
  w |= w >> 1;
  w |= w >> 2;
  w |= w >> 4;
  w |= w >> 8;
  w |= w >> 16;
  do {
    i ^= p;
    i *= 0xe170893du;
    i ^= p >> 16;
    i ^= (i & w) >> 4;
    i ^= p >> 8;
    i *= 0x0929eb3fu;
    i ^= p >> 23;
    i ^= (i & w) >> 1;
    i *= 1u | p >> 27;
    i *= 0x6935fa69u;
    i ^= (i & w) >> 11;
    i *= 0x74dcb303u;
    i ^= (i & w) >> 2;
    i *= 0x9e501cc3u;
    i ^= (i & w) >> 2;
    i *= 0xc860a3dfu;
    i &= w;
    i ^= i >> 5;
  } while (i >= l);
  return (i + p) % l;
}

// Kensler's randfloat (cmj.cu:45-58)
FR_HD float kensler_randfloat(uint32_t i, uint32_t p)
{
  i ^= p;
  i ^= i >> 17;
  i ^= i >> 10;
  i *= 0xb36534e5u;
  i ^= i >> 12;
  i ^= i >> 21;
  i *= 0x93fc4795u;
  i ^= 0xdf6e307fu;
  i ^= i >> 17;
  i *= 1u | p >> 18;
  // the product must be rounded on its own: callers add to it, and a fused
  // multiply-add would differ from the host reference in the last bit
#ifdef __CUDA_ARCH__
  return __fmul_rn((float)i, 1.0f / 4294967808.0f);
#else
  return i * (1.0f / 4294967808.0f);
#endif
}

// one CMJ point of the 4x4 pattern (cmj.cu:60-69)
FR_HD float2 cmj4x4(uint32_t index, uint32_t scramble)
{
  index = kensler_permute(index, 16u, scramble * 0x51633e2du);
  const uint32_t col = index & 3u, row = index >> 2;
  const uint32_t sx = kensler_permute(col, 4u, scramble * 0xa511e9b3u);
  const uint32_t sy = kensler_permute(row, 4u, scramble * 0x63d83595u);
  const float jx = kensler_randfloat(index, scramble * 0xa399d265u);
  const float jy = kensler_randfloat(index, scramble * 0x711ad6a5u);
  return make_float2((col + (sy + jx) / 4) / 4, (row + (sx + jy) / 4) / 4);
}

FR_HD uint32_t reverse_bits32(uint32_t x)
{
#ifdef __CUDA_ARCH__
  return __brev(x);
#else
  x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
  x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
  x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
  return (x >> 16) | (x << 16);
#endif
}

// Laine-Karras hash based Owen scramble (sobol.cu:10706-10731)
FR_HD uint32_t owen_scramble(uint32_t x, uint32_t seed)
{
  x = reverse_bits32(x);
  x += seed;
  x ^= x * 0x6c50b47cu;
  x ^= x * 0xb82f1e52u;
  x ^= x * 0xc7afe638u;
  x ^= x * 0x8d22f6e6u;
  return reverse_bits32(x);
}

FR_HD uint32_t seed_mix(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }

// 5-word sampler state; `draws` packs the two running counters.
struct PathSampler {
  uint32_t pixel;       // image_idx
  uint32_t n_spp;       // sample index of this path
  uint32_t sobol_index;  // owen(pixel + n_spp*W*H), fixed for the path
  uint32_t seed_hash;   // xxhash32(seed)
  uint32_t cmj_draws;   // number of 2-D draws so far
  uint32_t sobol_dim;   // next Sobol dimension (starts at 1)

  // pt.cu:378-399
  FR_HD void init(uint32_t pixel_, uint32_t n_spp_, uint32_t n_pixels, uint32_t seed)
  {
    pixel = pixel_;
    n_spp = n_spp_;
    seed_hash = xxhash32(seed);
    sobol_index = owen_scramble(pixel_ + n_spp_ * n_pixels, seed_hash);
    cmj_draws = 0;
    sobol_dim = 1;
  }

  FR_HD float2 next2d()
  {
    const uint32_t scramble = xxhash32(n_spp / 16u, pixel, cmj_draws, seed_hash);
    cmj_draws++;
    return cmj4x4(n_spp % 16u, scramble);
  }

#ifdef FR_HAVE_TABLES
  FR_D float next1d()
  {
    // the 32 columns of this dimension are one 128-byte line: eight independent 16-byte
    // loads (usually warp-uniform) and 32 masked XORs, no data-dependent load chain
    const uint4* m = reinterpret_cast<const uint4*>(g_sobol_matrices + 32u * sobol_dim);
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 w = __ldg(m + q);
      const uint32_t b = sobol_index >> (4 * q);
      v ^= w.x & (0u - (b & 1u));
      v ^= w.y & (0u - ((b >> 1) & 1u));
      v ^= w.z & (0u - ((b >> 2) & 1u));
      v ^= w.w & (0u - ((b >> 3) & 1u));
    }
    const uint32_t r = owen_scramble(v, seed_mix(seed_hash, sobol_dim));
    sobol_dim++;
    return r * (1.0f / 4294967296.0f);
  }
#endif
};

}  // namespace frd
