// Sample-sharded rendering over several GPUs with ONE ncclReduce of the accumulation buffers
// (include/fredholm/multi_gpu.h).  Host code only: the kernels are the single-GPU ones.
#include "fredholm/multi_gpu.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <exception>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <thread>

#include "nvtx.h"
#include "renderer_impl.h"

namespace fredholm
{

namespace
{

// ---- NCCL, resolved at run time -------------------------------------------------------------------------
// dlopen by SONAME returns the copy the process already holds (a Python host that imported torch has loaded
// torch's bundled libnccl.so.2), otherwise the system library.
struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;

  template <typename F>
  void sym(F& f, const char* name)
  {
    f = reinterpret_cast<F>(dlsym(handle, name));
    if (!f) throw std::runtime_error(std::string("multi-GPU: libnccl has no symbol ") + name);
  }

  Nccl()
  {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle)
      throw std::runtime_error("multi-GPU rendering needs NCCL: libnccl.so.2 could not be loaded (no fallback path)");
    sym(GetUniqueId, "ncclGetUniqueId");
    sym(CommInitRank, "ncclCommInitRank");
    sym(CommInitAll, "ncclCommInitAll");
    sym(CommDestroy, "ncclCommDestroy");
    sym(Reduce, "ncclReduce");
    sym(GroupStart, "ncclGroupStart");
    sym(GroupEnd, "ncclGroupEnd");
    sym(GetErrorString, "ncclGetErrorString");
    sym(GetVersion, "ncclGetVersion");
  }
};

Nccl& nccl()
{
  static Nccl n;  // throws on every call until the library can be loaded (function-local static retry)
  return n;
}

#define FR_NCCL_CHECK(call)                                                                              \
  do {                                                                                                   \
    const ncclResult_t fr_r_ = (call);                                                                   \
    if (fr_r_ != ncclSuccess) {                                                                          \
      std::stringstream fr_ss_;                                                                          \
      fr_ss_ << "NCCL call (" << #call << ") failed: '" << nccl().GetErrorString(fr_r_) << "' (" << __FILE__ \
             << ":" << __LINE__ << ")";                                                                  \
      throw std::runtime_error(fr_ss_.str());                                                            \
    }                                                                                                    \
  } while (0)

constexpr uint32_t kCmjPattern = 16;

}  // namespace

CommId make_comm_id()
{
  static_assert(sizeof(ncclUniqueId) == sizeof(CommId), "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  FR_NCCL_CHECK(nccl().GetUniqueId(&id));
  CommId out;
  std::memcpy(out.bytes, &id, sizeof(out.bytes));
  return out;
}

void sample_slice(uint32_t total_spp, int rank, int world, uint32_t& first, uint32_t& count)
{
  if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("sample_slice: rank outside the world");
  const uint64_t blocks = ((uint64_t)total_spp + kCmjPattern - 1) / kCmjPattern;
  const uint64_t b0 = (uint64_t)rank * blocks / world, b1 = ((uint64_t)rank + 1) * blocks / world;
  first = (uint32_t)std::min<uint64_t>(b0 * kCmjPattern, total_spp);
  count = (uint32_t)std::min<uint64_t>(b1 * kCmjPattern, total_spp) - first;
}

// ---- one rank ---------------------------------------------------------------------------------------------
struct ShardedRenderer::Impl {
  Renderer* renderer = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

ShardedRenderer::ShardedRenderer(Renderer& renderer, const CommId& id, int rank, int world) : m_impl(new Impl())
{
  if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("ShardedRenderer: rank outside the world");
  m_impl->renderer = &renderer;
  m_impl->rank = rank;
  m_impl->world = world;
  FR_CUDA_CHECK(cudaSetDevice(renderer.get_device()));
  ncclUniqueId nid;
  std::memcpy(&nid, id.bytes, sizeof(nid));
  FR_NCCL_CHECK(nccl().CommInitRank(&m_impl->comm, world, nid, rank));
}

ShardedRenderer::~ShardedRenderer()
{
  if (m_impl && m_impl->comm) {
    cudaSetDevice(m_impl->renderer->get_device());
    cudaStreamSynchronize(m_impl->renderer->get_stream());
    nccl().CommDestroy(m_impl->comm);
  }
}

int ShardedRenderer::rank() const { return m_impl->rank; }
int ShardedRenderer::world() const { return m_impl->world; }
Renderer& ShardedRenderer::renderer() { return *m_impl->renderer; }

void ShardedRenderer::reduce(const RenderLayer& layer, uint32_t total_spp, int root)
{
  FR_NVTX_RANGE("reduce_film");
  Renderer& r = *m_impl->renderer;
  FR_CUDA_CHECK(cudaSetDevice(r.get_device()));
  if (root < 0 || root >= m_impl->world) throw std::invalid_argument("ShardedRenderer: root outside the world");
  const Renderer::Impl* im = r.impl();
  const size_t n_pixels = (size_t)im->width * im->height;
  cudaStream_t s = r.get_stream();
  // the single exchange step: every bound layer, in place on the root, one group = one launch per rank
  FR_NCCL_CHECK(nccl().GroupStart());
  auto red = [&](void* p, size_t n_floats) {
    if (p) FR_NCCL_CHECK(nccl().Reduce(p, p, n_floats, ncclFloat, ncclSum, root, m_impl->comm, s));
  };
  red(layer.beauty, 4 * n_pixels);
  red(layer.position, 4 * n_pixels);
  red(layer.normal, 4 * n_pixels);
  red(layer.texcoord, 4 * n_pixels);
  red(layer.albedo, 4 * n_pixels);
  red(layer.depth, n_pixels);
  FR_NCCL_CHECK(nccl().GroupEnd());
  if (m_impl->rank == root && total_spp > 0) r.scale_layers(layer, 1.0f / (float)total_spp);
}

void ShardedRenderer::render(const CameraParams& camera, const float3& bg_color, const RenderLayer& layer,
                             uint32_t total_spp, uint32_t max_depth, int root)
{
  Renderer& r = *m_impl->renderer;
  uint32_t first = 0, count = 0;
  sample_slice(total_spp, m_impl->rank, m_impl->world, first, count);
  const FilmMode mode0 = r.get_film_mode();
  const uint32_t count0 = r.get_sample_count();
  r.set_film_mode(FilmMode::SUM);
  r.set_sample_offset(count0 + first);
  if (count > 0) r.render(camera, bg_color, layer, count, max_depth);
  reduce(layer, total_spp, root);
  r.set_film_mode(mode0);
  r.set_sample_offset(count0 + total_spp);  // the frame now holds total_spp more samples, on every rank
}

void ShardedRenderer::render(const Camera& camera, const float3& bg_color, const RenderLayer& layer, uint32_t total_spp,
                             uint32_t max_depth, int root)
{
  render(m_impl->renderer->camera_params(camera), bg_color, layer, total_spp, max_depth, root);
}

// ---- one process, several devices ---------------------------------------------------------------------------
struct MultiGpuRenderer::Impl {
  std::vector<int> devices;
  std::vector<std::unique_ptr<Renderer>> renderers;
  std::vector<std::unique_ptr<ShardedRenderer>> ranks;
  // accumulation buffers of the ranks > 0 (rank 0 accumulates into the caller's layers)
  struct Layers {
    frd::DevBuf<float4> beauty, position, normal, texcoord, albedo;
    frd::DevBuf<float> depth;
  };
  std::vector<Layers> own;
};

MultiGpuRenderer::MultiGpuRenderer(const std::vector<int>& devices) : m_impl(new Impl())
{
  m_impl->devices = devices;
  if (m_impl->devices.empty()) {
    int n = 0;
    FR_CUDA_CHECK(cudaGetDeviceCount(&n));
    for (int d = 0; d < n; ++d) m_impl->devices.push_back(d);
  }
  const int n = (int)m_impl->devices.size();
  if (n == 0) throw std::runtime_error("MultiGpuRenderer: no CUDA device (there is no CPU fallback)");
  m_impl->renderers.resize(n);
  m_impl->ranks.resize(n);
  m_impl->own.resize(n);
  for (int i = 0; i < n; ++i) m_impl->renderers[i] = std::make_unique<Renderer>(m_impl->devices[i]);
  // ncclCommInitRank from one thread per rank (ncclCommInitAll's behaviour, with our own rank objects)
  const CommId id = make_comm_id();
  for_each([&](int rank, Renderer& r) { m_impl->ranks[rank] = std::make_unique<ShardedRenderer>(r, id, rank, n); });
}

MultiGpuRenderer::~MultiGpuRenderer() noexcept(false)
{
  if (!m_impl) return;
  for (auto& r : m_impl->ranks) r.reset();
  for (size_t i = 0; i < m_impl->own.size(); ++i) {
    cudaSetDevice(m_impl->devices[i]);
    m_impl->own[i] = Impl::Layers();
  }
  m_impl->renderers.clear();
}

int MultiGpuRenderer::size() const { return (int)m_impl->devices.size(); }
Renderer& MultiGpuRenderer::renderer(int rank) { return *m_impl->renderers.at(rank); }

void MultiGpuRenderer::for_each(const std::function<void(int, Renderer&)>& f)
{
  const int n = size();
  std::vector<std::thread> threads;
  std::mutex m;
  std::exception_ptr first_error;
  for (int rank = 0; rank < n; ++rank) {
    threads.emplace_back([&, rank] {
      try {
        FR_CUDA_CHECK(cudaSetDevice(m_impl->devices[rank]));
        f(rank, *m_impl->renderers[rank]);
      } catch (...) {
        std::lock_guard<std::mutex> lock(m);
        if (!first_error) first_error = std::current_exception();
      }
    });
  }
  for (auto& t : threads) t.join();
  if (first_error) std::rethrow_exception(first_error);
}

void MultiGpuRenderer::load_scene(const std::filesystem::path& filepath, bool clear)
{
  // parse once, upload everywhere
  renderer(0).load_scene(filepath, clear);
  const Scene& s = renderer(0).get_scene();
  for_each([&](int rank, Renderer& r) {
    if (rank > 0) r.set_scene(s);
  });
}
void MultiGpuRenderer::set_scene(const Scene& scene)
{
  for_each([&](int, Renderer& r) { r.set_scene(scene); });
}
void MultiGpuRenderer::build_gas()
{
  for_each([&](int, Renderer& r) { r.build_gas(); });
}
void MultiGpuRenderer::set_directional_light(const float3& le, const float3& dir, float angle)
{
  for_each([&](int, Renderer& r) { r.set_directional_light(le, dir, angle); });
}
void MultiGpuRenderer::load_arhosek_sky(float turbidity, float albedo)
{
  for_each([&](int, Renderer& r) { r.load_arhosek_sky(turbidity, albedo); });
}
void MultiGpuRenderer::set_resolution(uint32_t width, uint32_t height)
{
  for_each([&](int, Renderer& r) { r.set_resolution(width, height); });
}
void MultiGpuRenderer::set_max_wave_paths(size_t n_paths)
{
  for_each([&](int, Renderer& r) { r.set_max_wave_paths(n_paths); });
}

void MultiGpuRenderer::render(const Camera& camera, const float3& bg_color, const RenderLayer& layer, uint32_t total_spp,
                              uint32_t max_depth)
{
  render(renderer(0).camera_params(camera), bg_color, layer, total_spp, max_depth);
}

void MultiGpuRenderer::render(const CameraParams& camera, const float3& bg_color, const RenderLayer& layer,
                              uint32_t total_spp, uint32_t max_depth)
{
  for_each([&](int rank, Renderer& r) {
    RenderLayer mine = layer;
    if (rank > 0) {
      // same layers as the caller bound, on this device, zeroed on this device's stream
      Impl::Layers& o = m_impl->own[rank];
      const size_t n = (size_t)r.impl()->width * r.impl()->height;
      cudaStream_t s = r.get_stream();
      auto prep4 = [&](frd::DevBuf<float4>& b, float4* want) -> float4* {
        if (!want) return nullptr;
        b.reserve(n);
        FR_CUDA_CHECK(cudaMemsetAsync(b.get(), 0, sizeof(float4) * n, s));
        return b.get();
      };
      mine.beauty = prep4(o.beauty, layer.beauty);
      mine.position = prep4(o.position, layer.position);
      mine.normal = prep4(o.normal, layer.normal);
      mine.texcoord = prep4(o.texcoord, layer.texcoord);
      mine.albedo = prep4(o.albedo, layer.albedo);
      mine.depth = nullptr;
      if (layer.depth) {
        o.depth.reserve(n);
        FR_CUDA_CHECK(cudaMemsetAsync(o.depth.get(), 0, sizeof(float) * n, s));
        mine.depth = o.depth.get();
      }
    }
    m_impl->ranks[rank]->render(camera, bg_color, mine, total_spp, max_depth, 0);
  });
}

void MultiGpuRenderer::wait_for_completion()
{
  for_each([&](int, Renderer& r) { r.wait_for_completion(); });
}

}  // namespace fredholm
