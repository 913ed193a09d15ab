// cwl::CUDABuffer<T> -- the RAII device buffer the reference's applications own their AOV layers with
// (reference cwl/include/cwl/buffer.h:18-85; app/rtcamp8.cpp:86-111, app/controller.cpp:80-124) and read the
// framebuffer back through (copy_from_device_to_host, buffer.h:64-69).  Same constructors and method names,
// same 32-bit element count; the storage comes from the CUDA runtime API instead of the driver API.
// CUDAGLBuffer (buffer.h:87-150, OpenGL interop of the viewer) is not provided: the GL viewer is outside the
// hot path.
#pragma once
#include <cstdint>
#include <vector>

#include "cwl/util.h"

namespace cwl
{

template <typename T>
class CUDABuffer
{
 public:
  explicit CUDABuffer(uint32_t buffer_size) : m_buffer_size(buffer_size)
  {
    if (buffer_size == 0) return;
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&m_d_ptr), size_t(m_buffer_size) * sizeof(T)));
  }
  // every 32-bit word of the buffer is set to `value` (buffer.h:27-33)
  CUDABuffer(uint32_t buffer_size, uint32_t value) : CUDABuffer<T>(buffer_size)
  {
    if (buffer_size == 0) return;
    fill_words(value);
  }
  explicit CUDABuffer(const std::vector<T>& values) : CUDABuffer<T>(static_cast<uint32_t>(values.size()))
  {
    if (values.empty()) return;
    copy_from_host_to_device(values);
  }
  CUDABuffer(const CUDABuffer<T>& other) = delete;
  CUDABuffer& operator=(const CUDABuffer<T>& other) = delete;
  CUDABuffer(CUDABuffer<T>&& other) noexcept : m_d_ptr(other.m_d_ptr), m_buffer_size(other.m_buffer_size)
  {
    other.m_d_ptr = nullptr;
    other.m_buffer_size = 0;
  }
  ~CUDABuffer() noexcept(false)
  {
    if (m_d_ptr) CUDA_CHECK(cudaFree(m_d_ptr));
  }

  void clear() const { fill_words(0u); }

  void copy_from_host_to_device(const std::vector<T>& value) const
  {
    if (value.size() < m_buffer_size) throw std::runtime_error("CUDABuffer: host vector smaller than the buffer");
    CUDA_CHECK(cudaMemcpy(m_d_ptr, value.data(), size_t(m_buffer_size) * sizeof(T), cudaMemcpyHostToDevice));
  }

  void copy_from_device_to_host(std::vector<T>& value) const
  {
    value.resize(m_buffer_size);
    if (m_buffer_size == 0) return;
    CUDA_CHECK(cudaMemcpy(value.data(), m_d_ptr, size_t(m_buffer_size) * sizeof(T), cudaMemcpyDeviceToHost));
  }

  T* get_device_ptr() { return m_d_ptr; }
  const T* get_const_device_ptr() const { return m_d_ptr; }

  uint32_t get_size() const { return m_buffer_size; }
  uint32_t get_size_in_bytes() const { return m_buffer_size * static_cast<uint32_t>(sizeof(T)); }

 private:
  void fill_words(uint32_t value) const
  {
    if (m_buffer_size == 0) return;
    const size_t bytes = size_t(m_buffer_size) * sizeof(T);
    if (value == 0u) {
      CUDA_CHECK(cudaMemset(m_d_ptr, 0, bytes));
      return;
    }
    // cuMemsetD32 semantics with the runtime API: a 2-D memset of 4-byte "rows" would be slow, so stage one word
    // pattern through the host (constructor-time only)
    std::vector<uint32_t> words(bytes / sizeof(uint32_t), value);
    CUDA_CHECK(cudaMemcpy(m_d_ptr, words.data(), words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }

  T* m_d_ptr = nullptr;
  uint32_t m_buffer_size = 0;
};

}  // namespace cwl
