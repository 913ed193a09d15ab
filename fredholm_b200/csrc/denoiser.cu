// Denoise stage: edge-avoiding a-trous wavelet filter guided by the first-hit normal and
// albedo AOVs.  Stands in for the reference's OptiX AI denoiser
// (fredholm/include/fredholm/denoiser.h:14-145, invoked between render and post-process,
// app/rtcamp8.cpp:194-198); see include/fredholm/denoiser.h for why it is a different filter.
//
// k_prepare demodulates (c = beauty / max(albedo, floor)) and clamps fireflies against the 8
// direct neighbours; then one kernel per pass: every pixel gathers 25 taps at stride 2^i, the
// last pass remodulates.  HBM-bound: 48 B read (colour + normal + albedo) + 16 B written per
// pixel and launch; the 9x / 25x tap reuse is served by L1/L2 (tiles of 32x8 pixels per CTA).
//
// oracle/denoise_np.py restates the same arithmetic in numpy (parity tolerance 1e-4: expf).
#include "fredholm/denoiser.h"

#include <algorithm>
#include <cmath>

#include "cuda_util.h"
#include "vecmath.cuh"

namespace
{

using namespace frd;

struct AtrousArgs {
  const float4* color;   // pass input: demodulated colour (for k_prepare: the beauty layer)
  const float4* normal;
  const float4* albedo;
  float4* out;
  int width, height, step;
  float inv_sigma_c2, inv_sigma_a2, albedo_floor, firefly_k;
};

__device__ __forceinline__ float3 demod_albedo(const float4& a, float floor_)
{
  return f3(fmaxf(a.x, floor_), fmaxf(a.y, floor_), fmaxf(a.z, floor_));
}
// range domain of the colour weight: log(1 + c) per channel, i.e. relative differences, so
// that emitters and fireflies do not leak into their surroundings
__device__ __forceinline__ float3 compress(const float3& c)
{
  return f3(log1pf(fmaxf(c.x, 0.0f)), log1pf(fmaxf(c.y, 0.0f)), log1pf(fmaxf(c.z, 0.0f)));
}
__device__ __forceinline__ float pow64(float x)
{
  x *= x;
  x *= x;
  x *= x;
  x *= x;
  x *= x;
  return x * x;
}
// the normal AOV is a mean over the pixel's samples (shorter than 1 on geometric edges,
// zero on primary misses): the weights compare directions
__device__ __forceinline__ float3 unit_or_zero(const float4& n)
{
  const float d = n.x * n.x + n.y * n.y + n.z * n.z;
  if (!(d > 0.0f)) return f3(0.0f);
  const float inv = 1.0f / sqrtf(d);
  return f3(n.x * inv, n.y * inv, n.z * inv);
}
__device__ __forceinline__ float normal_weight(const float3& a, const float3& b)
{
  return pow64(fminf(fmaxf(a.x * b.x + a.y * b.y + a.z * b.z, 0.0f), 1.0f));
}
__device__ __forceinline__ float albedo_exponent(const float4& a, const float4& b, float inv_sigma_a2)
{
  const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return (dx * dx + dy * dy + dz * dz) * inv_sigma_a2;
}

// pass "-1": demodulate by the albedo and suppress fireflies.  A pixel whose brightest channel
// exceeds k x the brightest channel among its (up to 8) direct neighbours on the same surface
// (normal weight x albedo weight >= 0.5) is scaled down to that limit (+1e-3).
__global__ void __launch_bounds__(256) k_prepare(AtrousArgs a)
{
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= a.width || y >= a.height) return;
  const int p = x + a.width * y;
  const float4 bp = __ldg(a.color + p), np4 = __ldg(a.normal + p), ap4 = __ldg(a.albedo + p);
  const float3 mod_p = demod_albedo(ap4, a.albedo_floor);
  float3 c = f3(bp.x / mod_p.x, bp.y / mod_p.y, bp.z / mod_p.z);
  if (a.firefly_k > 0.0f) {
    const float3 np = unit_or_zero(np4);
    float m = -1.0f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        if (dx == 0 && dy == 0) continue;
        const int xx = x + dx, yy = y + dy;
        if (xx < 0 || xx >= a.width || yy < 0 || yy >= a.height) continue;
        const int q = xx + a.width * yy;
        const float4 bq = __ldg(a.color + q), nq4 = __ldg(a.normal + q), aq4 = __ldg(a.albedo + q);
        const float g = normal_weight(np, unit_or_zero(nq4)) * expf(-albedo_exponent(ap4, aq4, a.inv_sigma_a2));
        if (g >= 0.5f) {
          const float3 mod_q = demod_albedo(aq4, a.albedo_floor);
          m = fmaxf(m, fmaxf(bq.x / mod_q.x, fmaxf(bq.y / mod_q.y, bq.z / mod_q.z)));
        }
      }
    }
    const float top = fmaxf(c.x, fmaxf(c.y, c.z));
    const float limit = a.firefly_k * m + 1e-3f;
    if (m >= 0.0f && top > limit) {
      const float s = limit / top;
      c = f3(c.x * s, c.y * s, c.z * s);
    }
  }
  a.out[p] = make_float4(c.x, c.y, c.z, bp.w);
}

template <bool LAST>
__global__ void __launch_bounds__(256) k_atrous(AtrousArgs a)
{
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= a.width || y >= a.height) return;
  const int p = x + a.width * y;
  const float4 cp4 = __ldg(a.color + p), np4 = __ldg(a.normal + p), ap4 = __ldg(a.albedo + p);
  const float3 np = unit_or_zero(np4);
  const float3 rp = compress(f3(cp4));
  const float kw[5] = {1.0f / 16.0f, 1.0f / 4.0f, 3.0f / 8.0f, 1.0f / 4.0f, 1.0f / 16.0f};
  float3 sum = f3(0.0f);
  float wsum = 0.0f;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    const int yy = y + dy * a.step;
    if (yy < 0 || yy >= a.height) continue;
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int xx = x + dx * a.step;
      if (xx < 0 || xx >= a.width) continue;
      const int q = xx + a.width * yy;
      const float4 cq4 = __ldg(a.color + q), nq4 = __ldg(a.normal + q), aq4 = __ldg(a.albedo + q);
      const float3 rq = compress(f3(cq4));
      const float wn = (dx == 0 && dy == 0) ? 1.0f : normal_weight(np, unit_or_zero(nq4));
      const float drx = rp.x - rq.x, dry = rp.y - rq.y, drz = rp.z - rq.z;
      const float e = albedo_exponent(ap4, aq4, a.inv_sigma_a2) + (drx * drx + dry * dry + drz * drz) * a.inv_sigma_c2;
      const float w = kw[dx + 2] * kw[dy + 2] * wn * expf(-e);
      sum.x += w * cq4.x;
      sum.y += w * cq4.y;
      sum.z += w * cq4.z;
      wsum += w;
    }
  }
  const float inv = 1.0f / wsum;  // the centre tap alone contributes 9/64
  float3 o = f3(sum.x * inv, sum.y * inv, sum.z * inv);
  if (LAST) {
    const float3 mod_p = demod_albedo(ap4, a.albedo_floor);
    o = f3(o.x * mod_p.x, o.y * mod_p.y, o.z * mod_p.z);
  }
  // alpha of the beauty layer travels in .w untouched
  a.out[p] = make_float4(o.x, o.y, o.z, cp4.w);
}

// 2x bilinear upscale (pixel centres), the reference's UPSCALE2X model kind stands for this
__global__ void __launch_bounds__(256) k_upscale2x(const float4* __restrict__ in, int width, int height,
                                                   float4* __restrict__ out)
{
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= 2 * width || y >= 2 * height) return;
  const float sx = (x + 0.5f) * 0.5f - 0.5f, sy = (y + 0.5f) * 0.5f - 0.5f;
  const float fx = floorf(sx), fy = floorf(sy);
  const float tx = sx - fx, ty = sy - fy;
  const int x0 = min(max((int)fx, 0), width - 1), x1 = min(max((int)fx + 1, 0), width - 1);
  const int y0 = min(max((int)fy, 0), height - 1), y1 = min(max((int)fy + 1, 0), height - 1);
  const float4 c00 = __ldg(in + x0 + width * y0), c10 = __ldg(in + x1 + width * y0);
  const float4 c01 = __ldg(in + x0 + width * y1), c11 = __ldg(in + x1 + width * y1);
  const float w00 = (1.0f - tx) * (1.0f - ty), w10 = tx * (1.0f - ty), w01 = (1.0f - tx) * ty, w11 = tx * ty;
  out[x + 2 * width * y] = make_float4(w00 * c00.x + w10 * c10.x + w01 * c01.x + w11 * c11.x,
                                       w00 * c00.y + w10 * c10.y + w01 * c01.y + w11 * c11.y,
                                       w00 * c00.z + w10 * c10.z + w01 * c01.z + w11 * c11.z,
                                       w00 * c00.w + w10 * c10.w + w01 * c01.w + w11 * c11.w);
}

}  // namespace

namespace fredholm
{

struct Denoiser::Impl {
  uint32_t width = 0, height = 0;
  const float4* beauty = nullptr;
  const float4* normal = nullptr;
  const float4* albedo = nullptr;
  float4* denoised = nullptr;
  bool upscale = false;
  cudaStream_t stream = 0;
  DenoiserParams params;
  frd::DevBuf<float4> ping, pong;
};

Denoiser::Denoiser(uint32_t width, uint32_t height, const float4* d_beauty, const float4* d_normal,
                   const float4* d_albedo, float4* d_denoised, bool upscale, cudaStream_t stream)
    : m_impl(std::make_unique<Impl>())
{
  if (width == 0 || height == 0) throw std::runtime_error("Denoiser: empty image");
  if (!d_beauty || !d_normal || !d_albedo || !d_denoised) throw std::runtime_error("Denoiser: null layer pointer");
  m_impl->width = width;
  m_impl->height = height;
  m_impl->beauty = d_beauty;
  m_impl->normal = d_normal;
  m_impl->albedo = d_albedo;
  m_impl->denoised = d_denoised;
  m_impl->upscale = upscale;
  m_impl->stream = stream;
  m_impl->ping.alloc((size_t)width * height);
  m_impl->pong.alloc((size_t)width * height);
}

namespace
{
uint32_t on_device(int device, uint32_t width)
{
  FR_CUDA_CHECK(cudaSetDevice(device));
  return width;
}
}  // namespace

Denoiser::Denoiser(int context, uint32_t width, uint32_t height, const float4* d_beauty, const float4* d_normal,
                   const float4* d_albedo, const float4* d_denoised, bool upscale)
    : Denoiser(on_device(context, width), height, d_beauty, d_normal, d_albedo, const_cast<float4*>(d_denoised), upscale,
               /*stream=*/0)
{
}

Denoiser::~Denoiser() noexcept(false) {}

void Denoiser::set_params(const DenoiserParams& params)
{
  if (params.iterations < 1 || params.iterations > 12) throw std::runtime_error("Denoiser: iterations out of range");
  if (!(params.sigma_color > 0.0f) || !(params.sigma_albedo > 0.0f) || !(params.albedo_floor > 0.0f))
    throw std::runtime_error("Denoiser: sigmas and albedo floor must be positive");
  m_impl->params = params;
}
const DenoiserParams& Denoiser::get_params() const { return m_impl->params; }

void Denoiser::denoise()
{
  Impl& d = *m_impl;
  const dim3 block(32, 8);
  const dim3 grid((d.width + 31) / 32, (d.height + 7) / 8);
  const int n = d.params.iterations;
  AtrousArgs a;
  a.normal = d.normal;
  a.albedo = d.albedo;
  a.width = (int)d.width;
  a.height = (int)d.height;
  a.step = 1;
  a.inv_sigma_c2 = 0.0f;
  a.inv_sigma_a2 = 1.0f / (d.params.sigma_albedo * d.params.sigma_albedo);
  a.albedo_floor = d.params.albedo_floor;
  a.firefly_k = d.params.firefly_k;
  a.color = d.beauty;
  a.out = d.pong.get();
  k_prepare<<<grid, block, 0, d.stream>>>(a);
  FR_CUDA_LAUNCH_CHECK();
  const float4* src = a.out;
  for (int i = 0; i < n; ++i) {
    const bool last = i == n - 1;
    float4* dst = (last && !d.upscale) ? d.denoised : ((i & 1) ? d.pong.get() : d.ping.get());
    a.color = src;
    a.out = dst;
    a.step = 1 << i;
    const float sc = d.params.sigma_color / (float)(1 << i);
    a.inv_sigma_c2 = 1.0f / (sc * sc);
    if (last)
      k_atrous<true><<<grid, block, 0, d.stream>>>(a);
    else
      k_atrous<false><<<grid, block, 0, d.stream>>>(a);
    FR_CUDA_LAUNCH_CHECK();
    src = dst;
  }
  if (d.upscale) {
    const dim3 grid2((2 * d.width + 31) / 32, (2 * d.height + 7) / 8);
    k_upscale2x<<<grid2, block, 0, d.stream>>>(src, (int)d.width, (int)d.height, d.denoised);
    FR_CUDA_LAUNCH_CHECK();
  }
}

void Denoiser::wait_for_completion() const
{
  FR_CUDA_CHECK(cudaStreamSynchronize(m_impl->stream));
  FR_CUDA_CHECK(cudaGetLastError());
}

}  // namespace fredholm
