// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" entry points around the reference's OWN post-process kernels
// (fredholm/kernels/src/post-process.cu:5-153), which the Makefile compiles with nvcc for
// sm_100a from the sources where they lie under /root/reference.  The result
// (oracle/_ref/libpostprocess_ref.so) is the reference implementation of bloom /
// chromatic aberration / tone mapping running on the same B200 as the product kernels;
// the GPU parity test compares the two.  All pointers are DEVICE pointers.
#include <cuda_runtime.h>

#include <exception>
#include <string>

#include "kernels/post-process.h"

// post-process.cu:37-47 (the declaration in the reference header is stale, post-process.h:130-135)
void tone_mapping_kernel_launch(const float4* beauty_in, int width, int height, float ISO,
                                float chromatic_aberration, float4* beauty_out);

static std::string g_err;

extern "C" {

const char* ppr_last_error() { return g_err.c_str(); }

int ppr_post_process(const void* beauty_in, void* high, void* temp, int width, int height, int use_bloom,
                     float bloom_threshold, float bloom_sigma, float ISO, float chromatic_aberration, void* out)
{
  try {
    PostProcessParams p;
    p.use_bloom = use_bloom != 0;
    p.bloom_threshold = bloom_threshold;
    p.bloom_sigma = bloom_sigma;
    p.ISO = ISO;
    p.chromatic_aberration = chromatic_aberration;
    post_process_kernel_launch(static_cast<const float4*>(beauty_in), static_cast<float4*>(high),
                               static_cast<float4*>(temp), width, height, p, static_cast<float4*>(out));
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      g_err = cudaGetErrorString(e);
      return -1;
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

int ppr_tone_mapping(const void* beauty_in, int width, int height, float ISO, float chromatic_aberration, void* out)
{
  try {
    tone_mapping_kernel_launch(static_cast<const float4*>(beauty_in), width, height, ISO, chromatic_aberration,
                               static_cast<float4*>(out));
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      g_err = cudaGetErrorString(e);
      return -1;
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
}
