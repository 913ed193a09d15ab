// Error handling and a minimal owning device buffer for the core's internals.
// Error behaviour follows the reference (cwl/util.h:11-56): every failed CUDA
// call becomes a std::runtime_error carrying the call text and file:line.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <sstream>
#include <stdexcept>
#include <utility>
#include <vector>

#define FR_CUDA_CHECK(call)                                                          \
  do {                                                                               \
    const cudaError_t fr_err_ = (call);                                              \
    if (fr_err_ != cudaSuccess) {                                                    \
      std::stringstream fr_ss_;                                                      \
      fr_ss_ << "CUDA call (" << #call << ") failed with error: '"                   \
             << cudaGetErrorString(fr_err_) << "' (" << __FILE__ << ":" << __LINE__ \
             << ")";                                                                 \
      throw std::runtime_error(fr_ss_.str());                                        \
    }                                                                                \
  } while (0)

#define FR_CUDA_LAUNCH_CHECK() FR_CUDA_CHECK(cudaGetLastError())

namespace frd
{

// Grid of a persistent kernel = SMs x resident CTAs per SM of the CURRENT device, cached per device: one process
// may drive several devices from several host threads (MultiGpuRenderer), so a plain static would be both racy and
// wrong on a mixed box.
class GridCache
{
 public:
  int get(const void* kernel, int block)
  {
    int dev = 0;
    FR_CUDA_CHECK(cudaGetDevice(&dev));
    const bool cached = dev >= 0 && dev < kMaxDevices;
    if (cached) {
      const int g = m_grid[dev].load(std::memory_order_relaxed);
      if (g) return g;
    }
    int sms = 0, per_sm = 0;
    FR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0));
    const int g = sms * (per_sm > 0 ? per_sm : 1);
    if (cached) m_grid[dev].store(g, std::memory_order_relaxed);
    return g;
  }

 private:
  static constexpr int kMaxDevices = 64;
  std::atomic<int> m_grid[kMaxDevices] = {};
};

template <typename T>
class DevBuf
{
 public:
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p_(o.p_), n_(o.n_)
  {
    o.p_ = nullptr;
    o.n_ = 0;
  }
  DevBuf& operator=(DevBuf&& o) noexcept
  {
    if (this != &o) {
      release();
      p_ = o.p_;
      n_ = o.n_;
      o.p_ = nullptr;
      o.n_ = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }

  void alloc(size_t n)
  {
    release();
    if (n) {
      // size and pointer change together: a failed cudaMalloc leaves an EMPTY buffer, never {nullptr, n}
      T* p = nullptr;
      const cudaError_t err = cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(T));
      if (err != cudaSuccess) {
        cudaGetLastError();  // clear the sticky-free error so the context stays usable
        std::stringstream ss;
        ss << "cudaMalloc of " << n * sizeof(T) << " bytes failed: '" << cudaGetErrorString(err) << "'";
        throw std::runtime_error(ss.str());
      }
      p_ = p;
      n_ = n;
    }
  }
  // grow-only (contents are NOT preserved)
  void reserve(size_t n)
  {
    if (n > n_) alloc(n);
  }
  void release()
  {
    if (p_) cudaFree(p_);
    p_ = nullptr;
    n_ = 0;
  }
  void upload(const T* host, size_t n, cudaStream_t s = 0)
  {
    reserve(n);
    if (n) FR_CUDA_CHECK(cudaMemcpyAsync(p_, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& v, cudaStream_t s = 0) { upload(v.data(), v.size(), s); }
  void zero(cudaStream_t s = 0)
  {
    if (n_) FR_CUDA_CHECK(cudaMemsetAsync(p_, 0, n_ * sizeof(T), s));
  }
  T* get() const { return p_; }
  size_t size() const { return n_; }
  size_t bytes() const { return n_ * sizeof(T); }

 private:
  T* p_ = nullptr;
  size_t n_ = 0;
};

}  // namespace frd
