// cwl/util.h of the B200 core.
//
// Source compatibility with what the reference's applications and headers use from its CUDA helper library
// (reference cwl/include/cwl/util.h): the CUDA_CHECK / CUDA_SYNC_CHECK macros (cwl/util.h:11-36), cudaCheckError
// (cwl/util.h:41-58) and DeviceObject<T> (cwl/util.h:60-81).  Differences by design: runtime API only (no driver
// API, no OptiX header, so an application links libcudart and nothing else), one out-of-line formatting function
// instead of a stringstream per call site, and DeviceObject can be refreshed and read back.
// Error contract as in the reference: std::runtime_error whose text names the failing call and file:line.
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

namespace cwl
{
namespace detail
{

[[noreturn]] inline void fail(const std::string& head, cudaError_t code, const char* file, int line)
{
  throw std::runtime_error(head + "'" + cudaGetErrorString(code) + "' (" + file + ":" + std::to_string(line) + ")\n");
}

inline void check_call(cudaError_t code, const char* call_text, const char* file, int line)
{
  if (code != cudaSuccess) fail(std::string("CUDA call (") + call_text + " ) failed with error: ", code, file, line);
}

inline void check_sync(const char* file, int line)
{
  cudaDeviceSynchronize();
  const cudaError_t code = cudaGetLastError();
  if (code != cudaSuccess) fail("CUDA error on synchronize with error ", code, file, line);
}

}  // namespace detail

// same contract for a result value the caller already holds (the reference's overload takes a driver-API CUresult)
inline void cudaCheckError(cudaError_t result, const char* file = __builtin_FILE(), int line = __builtin_LINE(),
                           const char* function = __builtin_FUNCTION())
{
  if (result == cudaSuccess) return;
  throw std::runtime_error(std::string(file) + "(" + std::to_string(line) + ") " + function + ": " +
                           cudaGetErrorName(result) + ": " + cudaGetErrorString(result) + "\n");
}

// One host object mirrored in device memory (launch parameters and the like), freed with the wrapper.
template <typename T>
class DeviceObject
{
 public:
  explicit DeviceObject(const T& value)
  {
    detail::check_call(cudaMalloc(&m_storage, sizeof(T)), "cudaMalloc", __FILE__, __LINE__);
    upload(value);
  }
  DeviceObject(const DeviceObject&) = delete;
  DeviceObject& operator=(const DeviceObject&) = delete;
  ~DeviceObject() noexcept(false) { detail::check_call(cudaFree(m_storage), "cudaFree", __FILE__, __LINE__); }

  T* get_device_ptr() const { return static_cast<T*>(m_storage); }

  // extensions: refresh the device copy / read it back
  void upload(const T& value)
  {
    detail::check_call(cudaMemcpy(m_storage, &value, sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy", __FILE__, __LINE__);
  }
  T download() const
  {
    T value;
    detail::check_call(cudaMemcpy(&value, m_storage, sizeof(T), cudaMemcpyDeviceToHost), "cudaMemcpy", __FILE__, __LINE__);
    return value;
  }

 private:
  void* m_storage = nullptr;
};

}  // namespace cwl

#define CUDA_CHECK(call) ::cwl::detail::check_call((call), #call, __FILE__, __LINE__)
#define CUDA_SYNC_CHECK() ::cwl::detail::check_sync(__FILE__, __LINE__)
