// Multi-GPU rendering inside the C++ core.
//
// The path shards by SAMPLE INDEX: samples are independent given (pixel, sample index, seed)
// (reference init_sampler_state, fredholm/modules/pt.cu:378-399), so every GPU holds the whole scene and
// BVH, renders a contiguous slice of the sample indices of every pixel into SUM accumulators, and ONE
// ncclReduce of the accumulation buffers over NVLink followed by a division by the total sample count
// gives the image a single GPU would have rendered, up to fp32 summation order.  There is no per-bounce
// communication.  Slices are multiples of 16 samples so that every 4x4 CMJ pattern (cmj.cu:71-80) stays on
// one GPU.
//
// Two shapes, same arithmetic:
//   ShardedRenderer    one rank of a world: wraps the Renderer of ONE device.  The ranks may be threads of
//                      one process or one process per GPU (torchrun): the 128-byte communicator id made on
//                      rank 0 (make_comm_id) reaches the others by any channel the host has.
//   MultiGpuRenderer   one process, all (or the given) devices of a box: one Renderer and one host thread
//                      per device -- the shape of the reference's batch application (one render thread,
//                      app/rtcamp8.cpp:159-246), times N.
//
// NCCL is loaded at run time (libnccl.so.2, whichever copy the process already holds), so the library has
// no link-time dependency on it; every entry point below throws std::runtime_error when it is missing.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <vector>

#include "fredholm/renderer.h"

namespace fredholm
{

struct CommId {
  char bytes[128];  // ncclUniqueId
};
// rank 0 makes the id (ncclGetUniqueId); every rank of the world passes the same bytes to ShardedRenderer
CommId make_comm_id();

// rank's slice of the samples [0, total_spp): contiguous, disjoint, whole CMJ patterns (16) where possible
void sample_slice(uint32_t total_spp, int rank, int world, uint32_t& first, uint32_t& count);

class ShardedRenderer
{
 public:
  // collective over the world: ncclCommInitRank on the renderer's device
  ShardedRenderer(Renderer& renderer, const CommId& id, int rank, int world);
  ~ShardedRenderer();
  ShardedRenderer(const ShardedRenderer&) = delete;
  ShardedRenderer& operator=(const ShardedRenderer&) = delete;

  int rank() const;
  int world() const;
  Renderer& renderer();

  // Collective.  Renders this rank's slice of a total_spp-sample frame into `layer` (which must be zeroed:
  // sums are accumulated), reduces every bound layer onto `root` with ncclReduce on the renderer's stream
  // and turns the sums into means there.  Asynchronous on the renderer's stream; the other ranks' layers
  // hold their partial sums afterwards.  Restores the renderer's film mode and sample offset.
  void render(const Camera& camera, const float3& bg_color, const RenderLayer& layer, uint32_t total_spp,
              uint32_t max_depth, int root = 0);
  void render(const CameraParams& camera, const float3& bg_color, const RenderLayer& layer, uint32_t total_spp,
              uint32_t max_depth, int root = 0);
  // the exchange step alone (sums already in `layer`): reduce to root + scale by 1 / total_spp there
  void reduce(const RenderLayer& layer, uint32_t total_spp, int root = 0);

  struct Impl;

 private:
  std::unique_ptr<Impl> m_impl;
};

class MultiGpuRenderer
{
 public:
  // devices: CUDA device indices, one rank each (rank 0 = devices[0] owns the result); empty = all devices
  explicit MultiGpuRenderer(const std::vector<int>& devices = {});
  ~MultiGpuRenderer() noexcept(false);
  MultiGpuRenderer(const MultiGpuRenderer&) = delete;
  MultiGpuRenderer& operator=(const MultiGpuRenderer&) = delete;

  int size() const;
  Renderer& renderer(int rank);
  // runs f(rank, renderer) for every rank on that rank's host thread (device set) and waits for all;
  // the first exception is rethrown.  Scene loading, build_gas, lights, resolution go through this.
  void for_each(const std::function<void(int, Renderer&)>& f);

  // Renderer's call sequence, replicated on every device
  void load_scene(const std::filesystem::path& filepath, bool clear = true);
  void set_scene(const Scene& scene);
  void build_gas();
  void set_directional_light(const float3& le, const float3& dir, float angle);
  void load_arhosek_sky(float turbidity, float albedo);
  void set_resolution(uint32_t width, uint32_t height);
  void set_max_wave_paths(size_t n_paths);

  // total_spp samples per pixel, split over the devices; `layer` holds DEVICE pointers on devices[0]
  // (the caller owns and clears them, as with Renderer::render; note that Renderer methods leave the calling
  // thread on their own device, so allocate after cudaSetDevice(devices[0])); the other devices accumulate
  // into buffers the object owns.  Asynchronous; wait_for_completion() synchronises every device.
  void render(const Camera& camera, const float3& bg_color, const RenderLayer& layer, uint32_t total_spp,
              uint32_t max_depth);
  void render(const CameraParams& camera, const float3& bg_color, const RenderLayer& layer, uint32_t total_spp,
              uint32_t max_depth);
  void wait_for_completion();

 private:
  struct Impl;
  std::unique_ptr<Impl> m_impl;
};

}  // namespace fredholm
