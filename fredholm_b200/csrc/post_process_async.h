// Stream-taking forms of the post-process stage (post_process.cu) for callers that pipeline
// frames (batch.cpp).  The public free functions of kernels/post-process.h run on the
// default stream like the reference's (post-process.cu:5-47).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "kernels/post-process.h"

namespace frd
{

void post_process_async(const float4* beauty_in, float4* beauty_high_luminance, float4* beauty_temp, int width,
                        int height, const PostProcessParams& params, float4* beauty_out, cudaStream_t stream);

// float4 -> RGBA8 as the reference's applications convert on the host (app/rtcamp8.cpp:268-280):
// (unsigned char)clamp(255 * v, 0, 255), alpha = 255
void quantize_rgba8_async(const float4* in, size_t n_pixels, uchar4* out, cudaStream_t stream);

}  // namespace frd
