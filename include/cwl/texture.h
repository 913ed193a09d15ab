// cwl::CUDATexture<T> of the B200 core: an owning CUDA texture object whose sampler is the one the reference
// gives its scene textures (reference cwl/include/cwl/texture.h:13-74) -- normalized coordinates, wrap addressing
// in both directions, bilinear filter, 8-bit texels read as [0,1] floats, optional sRGB decode.
//
// Provided so that applications written against the reference's helper library compile and run; the renderer of
// this core does not use texture objects (it keeps RGBA8 arrays and filters in the shade stage in fp32 so that the
// fetch is reproducible on the host, csrc/surface.cuh).  Storage and sampler creation are not templates: the
// template only fixes the channel format and the read mode.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "cwl/util.h"

namespace cwl
{
namespace detail
{

// the reference's sampler state as a descriptor
inline cudaTextureDesc reference_sampler(bool texels_as_unit_floats, bool decode_srgb)
{
  cudaTextureDesc d = {};
  d.normalizedCoords = 1;
  for (cudaTextureAddressMode& mode : d.addressMode) mode = cudaAddressModeWrap;
  d.filterMode = cudaFilterModeLinear;
  d.mipmapFilterMode = cudaFilterModePoint;
  d.minMipmapLevelClamp = 0;
  d.maxMipmapLevelClamp = 99;
  d.maxAnisotropy = 1;
  d.readMode = texels_as_unit_floats ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
  d.sRGB = decode_srgb ? 1 : 0;
  return d;
}

// a 2-D CUDA array with one texture object on it
class TextureStorage
{
 public:
  TextureStorage(const cudaChannelFormatDesc& format, size_t texel_bytes, uint32_t width, uint32_t height, const void* texels,
                 const cudaTextureDesc& sampler)
  {
    CUDA_CHECK(cudaMallocArray(&m_array, &format, width, height));
    const size_t row_bytes = texel_bytes * width;
    CUDA_CHECK(cudaMemcpy2DToArray(m_array, 0, 0, texels, row_bytes, row_bytes, height, cudaMemcpyHostToDevice));
    cudaResourceDesc resource = {};
    resource.resType = cudaResourceTypeArray;
    resource.res.array.array = m_array;
    CUDA_CHECK(cudaCreateTextureObject(&m_object, &resource, &sampler, nullptr));
  }
  TextureStorage(const TextureStorage&) = delete;
  TextureStorage& operator=(const TextureStorage&) = delete;
  TextureStorage(TextureStorage&& other) noexcept : m_array(other.m_array), m_object(other.m_object)
  {
    other.m_array = nullptr;
    other.m_object = 0;
  }
  ~TextureStorage() noexcept(false)
  {
    if (m_object) CUDA_CHECK(cudaDestroyTextureObject(m_object));
    if (m_array) CUDA_CHECK(cudaFreeArray(m_array));
  }
  cudaTextureObject_t object() const { return m_object; }

 private:
  cudaArray_t m_array = nullptr;
  cudaTextureObject_t m_object = 0;
};

}  // namespace detail

template <typename T>
class CUDATexture
{
 public:
  CUDATexture(uint32_t width, uint32_t height, const T* data, bool srgb_to_linear = false)
      : m_size(make_uint2(width, height)),
        m_storage(cudaCreateChannelDesc<T>(), sizeof(T), width, height, data,
                  detail::reference_sampler(std::is_same<T, uchar4>::value, srgb_to_linear))
  {
  }
  CUDATexture(const CUDATexture&) = delete;
  CUDATexture& operator=(const CUDATexture&) = delete;
  CUDATexture(CUDATexture&&) noexcept = default;

  uint2 get_size() const { return m_size; }
  cudaTextureObject_t get_texture_object() const { return m_storage.object(); }

 private:
  uint2 m_size;
  detail::TextureStorage m_storage;
};

}  // namespace cwl
