// cwl::CUDATexture<T> -- RAII CUDA texture object with the reference's sampler state (reference
// cwl/include/cwl/texture.h:13-74: normalized coordinates, wrap addressing, linear filter, 8-bit -> [0,1],
// optional sRGB decode).  Provided for applications that create texture objects themselves; the renderer of
// this core keeps its scene textures as plain RGBA8 arrays and filters them in the shade stage with the same
// sampler rule in fp32 (csrc/surface.cuh; DESIGN.md "textures").
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "cwl/util.h"

namespace cwl
{

template <typename T>
class CUDATexture
{
 public:
  CUDATexture(uint32_t width, uint32_t height, const T* data, bool srgb_to_linear = false)
      : size(make_uint2(width, height))
  {
    const cudaChannelFormatDesc channel_desc = cudaCreateChannelDesc<T>();
    CUDA_CHECK(cudaMallocArray(&m_array, &channel_desc, width, height));
    const size_t pitch = size_t(width) * sizeof(T);
    CUDA_CHECK(cudaMemcpy2DToArray(m_array, 0, 0, data, pitch, pitch, height, cudaMemcpyHostToDevice));

    cudaResourceDesc res_desc = {};
    res_desc.resType = cudaResourceTypeArray;
    res_desc.res.array.array = m_array;

    cudaTextureDesc tex_desc = {};
    tex_desc.addressMode[0] = cudaAddressModeWrap;
    tex_desc.addressMode[1] = cudaAddressModeWrap;
    tex_desc.filterMode = cudaFilterModeLinear;
    tex_desc.readMode = std::is_same<T, uchar4>::value ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    tex_desc.normalizedCoords = 1;
    tex_desc.maxAnisotropy = 1;
    tex_desc.maxMipmapLevelClamp = 99;
    tex_desc.minMipmapLevelClamp = 0;
    tex_desc.mipmapFilterMode = cudaFilterModePoint;
    tex_desc.sRGB = srgb_to_linear ? 1 : 0;
    CUDA_CHECK(cudaCreateTextureObject(&m_texture_object, &res_desc, &tex_desc, nullptr));
  }

  CUDATexture(const CUDATexture& other) = delete;
  CUDATexture& operator=(const CUDATexture& other) = delete;
  CUDATexture(CUDATexture&& other) noexcept
      : size(other.size), m_array(other.m_array), m_texture_object(other.m_texture_object)
  {
    other.m_array = nullptr;
    other.m_texture_object = 0;
  }

  ~CUDATexture() noexcept(false)
  {
    if (m_texture_object) CUDA_CHECK(cudaDestroyTextureObject(m_texture_object));
    if (m_array) CUDA_CHECK(cudaFreeArray(m_array));
  }

  uint2 get_size() const { return size; }
  cudaTextureObject_t get_texture_object() const { return m_texture_object; }

 private:
  uint2 size;
  cudaArray_t m_array = nullptr;
  cudaTextureObject_t m_texture_object = 0;
};

}  // namespace cwl
