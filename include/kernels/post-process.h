// Post-process stage: bloom, chromatic aberration, exposure, tone mapping, sRGB.
//
// Same free-function interface as the reference's kernels library
// (fredholm/kernels/include/kernels/post-process.h:4-10,120-135 and
// fredholm/kernels/src/post-process.cu:5-47).  All pointers are device pointers to
// width*height float4 images; the calls run on the default stream and return after
// enqueueing the last kernel (like the reference, which only synchronises between
// the bloom passes).
#pragma once
#include <cuda_runtime.h>

struct PostProcessParams {
  bool use_bloom;
  float bloom_threshold;
  float bloom_sigma;
  float ISO;
  float chromatic_aberration;
};

void post_process_kernel_launch(const float4* beauty_in, float4* beauty_high_luminance, float4* beauty_temp,
                                int width, int height, const PostProcessParams& params, float4* beauty_out);

void tone_mapping_kernel_launch(const float4* beauty_in, int width, int height, float ISO,
                                float chromatic_aberration, float4* beauty_out);
