// fredholm::Denoiser -- albedo/normal-guided denoise stage between render and post-process.
//
// Same call surface as the reference's wrapper around the OptiX AI denoiser
// (fredholm/include/fredholm/denoiser.h:14-145; used by app/rtcamp8.cpp:113-118,194-198 and
// app/controller.cpp): construct once with the device pointers of the beauty / normal /
// albedo AOV layers and of the output image, then call denoise() after every render.
// The OptiX denoiser is a proprietary network that does not exist without OptiX, so the
// stage is replaced by a hand-written edge-avoiding a-trous wavelet filter (Dammertz et al.
// 2010) over the same three guide layers: the beauty layer is demodulated by the first-hit
// albedo, fireflies are clamped against the 8 direct neighbours, then 5 passes of a 5x5
// B3-spline kernel at strides 1,2,4,8,16 whose taps are weighted by normal, albedo and
// (log-domain) colour similarity filter the image before it is remodulated.
// Output is therefore NOT comparable with the reference's denoised pixels; its parity
// oracle is the numpy restatement of this filter (oracle/denoise_np.py).
//
// Intentional signature change (as for Renderer): no OptixDeviceContext; an optional CUDA
// stream instead (0 = default stream, the reference's behaviour).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>

namespace fredholm
{

struct DenoiserParams {
  int iterations = 5;           // a-trous passes, stride 2^i
  float sigma_color = 0.5f;     // range sigma of pass 0 on log(1 + c); halved every pass
  float sigma_albedo = 0.1f;    // sigma on the raw albedo difference
  float albedo_floor = 0.01f;   // demodulation: c = beauty / max(albedo, floor)
  float firefly_k = 2.0f;       // pre-pass: clamp a pixel to k x its brightest same-surface neighbour; 0 = off
};

class Denoiser
{
 public:
  Denoiser(uint32_t width, uint32_t height, const float4* d_beauty, const float4* d_normal, const float4* d_albedo,
           float4* d_denoised, bool upscale = false, cudaStream_t stream = 0);
  // the reference's argument order (denoiser.h:17-20: context first), so that applications written against it
  // compile unchanged with include/optwl/optwl.h: `context` is the CUDA device index (optwl::Context::m_context)
  Denoiser(int context, uint32_t width, uint32_t height, const float4* d_beauty, const float4* d_normal,
           const float4* d_albedo, const float4* d_denoised, bool upscale = false);
  ~Denoiser() noexcept(false);
  Denoiser(const Denoiser&) = delete;
  Denoiser& operator=(const Denoiser&) = delete;

  void set_params(const DenoiserParams& params);
  const DenoiserParams& get_params() const;

  // enqueues the filter on the stream; output is width x height (2x each when upscale)
  void denoise();
  void wait_for_completion() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> m_impl;
};

}  // namespace fredholm
