// cwl/util.h of the B200 core: the error macros and the single-object device wrapper the reference's
// applications and headers are written against (reference cwl/include/cwl/util.h:11-81), without OptiX and
// without the CUDA driver API -- everything goes through the runtime API, so an application links libcudart
// only.
//   CUDA_CHECK(call)          cwl/util.h:11-22   std::runtime_error with the call text and file:line
//   CUDA_SYNC_CHECK()         cwl/util.h:24-36   cudaDeviceSynchronize + cudaGetLastError
//   cwl::cudaCheckError(r)    cwl/util.h:41-58   same contract for a result value; takes cudaError_t here (the
//                                                reference's takes a driver-API CUresult)
//   cwl::DeviceObject<T>      cwl/util.h:60-81   RAII copy of one host object in device memory
#pragma once
#include <cuda_runtime.h>

#include <sstream>
#include <stdexcept>

#define CUDA_CHECK(call)                                                                          \
  do {                                                                                            \
    const cudaError_t cwl_error_ = (call);                                                        \
    if (cwl_error_ != cudaSuccess) {                                                              \
      std::stringstream cwl_ss_;                                                                  \
      cwl_ss_ << "CUDA call (" << #call << " ) failed with error: '" << cudaGetErrorString(cwl_error_) \
              << "' (" << __FILE__ << ":" << __LINE__ << ")\n";                                   \
      throw std::runtime_error(cwl_ss_.str());                                                    \
    }                                                                                             \
  } while (0)

#define CUDA_SYNC_CHECK()                                                                             \
  do {                                                                                                \
    cudaDeviceSynchronize();                                                                          \
    const cudaError_t cwl_error_ = cudaGetLastError();                                                \
    if (cwl_error_ != cudaSuccess) {                                                                  \
      std::stringstream cwl_ss_;                                                                      \
      cwl_ss_ << "CUDA error on synchronize with error '" << cudaGetErrorString(cwl_error_) << "' ("  \
              << __FILE__ << ":" << __LINE__ << ")\n";                                                \
      throw std::runtime_error(cwl_ss_.str());                                                        \
    }                                                                                                 \
  } while (0)

namespace cwl
{

inline void cudaCheckError(cudaError_t result, const char* file = __builtin_FILE(), int line = __builtin_LINE(),
                           const char* function = __builtin_FUNCTION())
{
  if (result == cudaSuccess) return;
  std::stringstream ss;
  ss << file << "(" << line << ") " << function << ": " << cudaGetErrorName(result) << ": "
     << cudaGetErrorString(result) << std::endl;
  throw std::runtime_error(ss.str());
}

// RAII wrapper for one object in device memory
template <typename T>
class DeviceObject
{
 public:
  explicit DeviceObject(const T& object)
  {
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&m_device_ptr), sizeof(T)));
    CUDA_CHECK(cudaMemcpy(m_device_ptr, &object, sizeof(T), cudaMemcpyHostToDevice));
  }
  DeviceObject(const DeviceObject&) = delete;
  DeviceObject& operator=(const DeviceObject&) = delete;
  ~DeviceObject() noexcept(false) { CUDA_CHECK(cudaFree(reinterpret_cast<void*>(m_device_ptr))); }

  T* get_device_ptr() const { return m_device_ptr; }

 private:
  T* m_device_ptr = nullptr;
};

}  // namespace cwl
