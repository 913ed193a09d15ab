// fredholm::Renderer -- the rendering core behind the application-facing API.
//
// Same call surface as the reference's header-only OptiX renderer
// (fredholm/include/fredholm/renderer.h:29-846): applications call, in order,
//   create_module -> create_program_group -> create_pipeline -> set_resolution ->
//   load_scene -> build_gas -> build_ias -> create_sbt -> render(...) ->
//   wait_for_completion
// (app/controller.cpp:60-68,126-134; app/rtcamp8.cpp:79-124).  Here the OptiX
// pipeline steps are no-ops kept for source compatibility: the kernels are
// hand-written sm_100a code linked into libfredholm_b200.so, build_gas() builds
// the GPU LBVH -> CWBVH over all sub-meshes and build_ias() is folded into it
// (one world-space tree).  The only intentional signature change is the
// constructor (no OptixDeviceContext): Renderer(int cuda_device).
//
// render() is "canonical mode" of the reference: n_samples samples are taken as
// n_samples launches of one sample each would take them.
// Errors surface as std::runtime_error with file:line, like the reference's
// CUDA_CHECK / OPTIX_CHECK.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <filesystem>
#include <memory>

#include "fredholm/camera.h"
#include "fredholm/scene.h"
#include "fredholm/shared.h"

namespace fredholm
{

enum class FilmMode : int {
  MEAN = 0,  // streaming mean into the layers (reference behaviour)
  SUM = 1,   // running sums; divide by the sample count after a multi-GPU reduce
};

struct RenderStatistics {
  unsigned long long paths = 0;
  unsigned long long rays_radiance = 0;
  unsigned long long rays_shadow = 0;
  unsigned long long rays_light = 0;
  unsigned long long kernel_launches = 0;
  // visibility / MIS rays whose contribution is exactly zero: the reference traces them (pt.cu:766-925), this
  // core does not.  total_rays() + rays_skipped = the reference's optixTrace call count for the same samples.
  unsigned long long rays_skipped = 0;
  // only filled while set_traversal_counting(true): BVH nodes visited / triangles tested by the
  // radiance, shadow and MIS rays (index 0, 1, 2)
  unsigned long long nodes_visited[3] = {0, 0, 0};
  unsigned long long tris_tested[3] = {0, 0, 0};
  unsigned long long total_rays() const { return rays_radiance + rays_shadow + rays_light; }
};

struct AccelInfo {
  uint32_t n_faces = 0;
  uint32_t n_nodes = 0;
  uint32_t depth = 0;
  float build_ms = 0.0f;
  size_t bytes = 0;
  // two-level structure (instanced scenes): distinct meshes, triangles actually stored, time of the last
  // instance-tree update (set_time / transform change)
  bool two_level = false;
  uint32_t n_instances = 0, n_meshes = 0, n_stored_faces = 0;
  float tlas_update_ms = 0.0f;
  bool tlas_refitted = false;  // the last update refitted the instance tree in place (else: rebuilt it)
};

// Acceleration-structure layout (extension).  FLAT: every instance's triangles in world space in ONE tree (fastest
// traversal; a transform change rebuilds it).  TWO_LEVEL: one object-space tree per distinct mesh + an instance
// tree, like the reference's GAS + IAS (renderer.h:434-552); a transform change only rebuilds the instance tree.
// AUTO: FLAT (bit-exact against the oracle, fastest) unless building it would take more than a third of the
// device's memory and at least half of the scene's triangles are copies of another sub-mesh.
enum class AccelMode : int { AUTO = 0, FLAT = 1, TWO_LEVEL = 2 };

class Renderer
{
 public:
  explicit Renderer(int cuda_device = 0);
  ~Renderer() noexcept(false);
  Renderer(const Renderer&) = delete;
  Renderer& operator=(const Renderer&) = delete;

  // ---- OptiX pipeline steps of the reference: no-ops here ----
  void create_module(const std::filesystem::path& filepath);
  void create_program_group();
  void create_pipeline();
  void create_sbt();

  // ---- scene ----
  void load_scene(const std::filesystem::path& filepath, bool clear = true);
  // extension: adopt an already filled Scene (procedural content)
  void set_scene(const Scene& scene);
  const Scene& get_scene() const;
  void build_gas();
  void build_ias();
  void set_time(float time);
  void set_accel_mode(AccelMode mode);  // extension; takes effect at the next build_gas()
  // extension: replace the sub-mesh transforms (column-major 4x4 each) and bring the acceleration structure up to
  // date -- an instance-tree update in TWO_LEVEL mode, a rebuild in FLAT mode
  void set_transforms(const float* transforms16, uint32_t n_submeshes);

  // ---- lights / sky ----
  void set_directional_light(const float3& le, const float3& dir, float angle);
  void clear_directional_light();  // extension
  void set_sky_intensity(float sky_intensity);
  void load_ibl(const std::filesystem::path& filepath);
  void set_ibl(const float4* texels, uint32_t width, uint32_t height);  // extension
  void clear_ibl();
  void load_arhosek_sky(float turbidity, float albedo);
  void clear_arhosek_sky();

  // ---- film ----
  void set_resolution(uint32_t width, uint32_t height);
  void init_render_states();

  // ---- render ----
  void render(const Camera& camera, const float3& bg_color, const RenderLayer& render_layer, uint32_t n_samples,
              uint32_t max_depth);
  // same, with the camera given as the packed parameter block
  void render(const CameraParams& camera, const float3& bg_color, const RenderLayer& render_layer,
              uint32_t n_samples, uint32_t max_depth);
  void wait_for_completion();

  // ---- extensions for multi-GPU sample slicing and tuning ----
  void set_sample_offset(uint32_t first_sample);  // next render starts at this sample index
  uint32_t get_sample_count() const;
  void set_film_mode(FilmMode mode);
  FilmMode get_film_mode() const;
  // the packed camera block render() uses for `camera` (a camera node of the scene overrides its transform)
  CameraParams camera_params(const Camera& camera) const;
  void scale_layers(const RenderLayer& render_layer, float scale);
  void set_max_wave_paths(size_t n_paths);
  size_t get_wave_state_bytes() const;  // device memory the integrator holds for path state and ray queues
  // two waves in flight on two CUDA streams (max_wave_paths is split between them): the late, nearly empty
  // bounces of one wave run under the full launches of the next.  Sample values and film order are unchanged.
  void set_wave_overlap(bool on);
  // Wave compaction (on by default): in a render of several waves, the few paths a wave still has alive after
  // `depth` bounces move to a dense straggler set that finishes the remaining bounces of up to eight waves
  // together, so small waves no longer pay their own nearly empty late launches.  Bit-identical layers; the host
  // waits once per wave for a path count.  Not in single-launch mode.
  void set_wave_compaction(bool on, uint32_t depth = 3);
  // One render(n_samples) call behaves like ONE reference launch of n_samples: payload.firsthit and the
  // first-hit AOVs outlive the sample loop (pt.cu:432-433, 744-759) -- what app/rtcamp8.cpp produces.  Off by
  // default: render(n_samples) equals n_samples launches of one sample, what the reference GUI produces.
  void set_single_launch(bool on);
  // samples of one pixel block that share a warp (1, 2, 4, 8, 16, 32): a scheduling choice, no sample changes
  void set_samples_per_warp(uint32_t spw);
  // measurement: run the counting instantiations of the traversal kernels (process-wide switch)
  void set_traversal_counting(bool on);
  // per-stage device time, measured with CUDA events on the renderer's stream;
  // stages: generate, trace_closest, shade, trace_shadow, trace_light, advance, film
  static constexpr int kStageCount = 7;
  void set_stage_timing(bool on);
  void get_stage_times(double ms[kStageCount], unsigned long long launches[kStageCount]);
  RenderStatistics get_statistics();
  void reset_statistics();
  AccelInfo get_accel_info() const;
  cudaStream_t get_stream() const;
  int get_device() const;

  struct Impl;
  Impl* impl() { return m_impl.get(); }

 private:
  std::unique_ptr<Impl> m_impl;
};

// Hosek-Wilkie RGB sky coefficients for (turbidity, ground albedo, solar elevation):
// out = 3 x 9 configuration values followed by 3 mean radiances
// (reference arhosek.h:145-323).
void arhosek_rgb_cook(float turbidity, float albedo, float elevation, float out30[30]);

}  // namespace fredholm
