#include "integrator.h"

#include "nvtx.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace frd
{

namespace
{
uint32_t env_u32(const char* name, uint32_t fallback)
{
  const char* e = getenv(name);
  return e ? (uint32_t)strtoul(e, nullptr, 0) : fallback;
}
}  // namespace

Integrator::Integrator(cudaStream_t stream) : m_stream(stream)
{
  m_sort_mask = env_u32("FRD_SORT", 0u);
  m_sort_bits = std::min(std::max(env_u32("FRD_SORT_BITS", 4u), 1u), 7u);
  set_samples_per_warp(env_u32("FRD_SAMPLES_PER_WARP", kDefaultSamplesPerWarp));
}

void Integrator::set_samples_per_warp(uint32_t spw)
{
  uint32_t l = 0;
  while (l < 5u && (2u << l) <= spw) l++;
  m_spw_log2 = l;
}

void Integrator::set_coherence_sort(uint32_t queue_mask, uint32_t cell_bits)
{
  m_sort_mask = queue_mask;
  m_sort_bits = std::min(std::max(cell_bits, 1u), 7u);
}

// sorts queue `which` and returns the order to trace it in (nullptr: sort disabled for it)
const uint32_t* Integrator::sorted(const SceneView& scene, const WaveBuffers& wb, int which, bool use_octant)
{
  SortGrid g;
  g.lo = scene.bounds_lo;
  const float cells = (float)(1u << m_sort_bits);
  const float ex = std::max(scene.bounds_hi.x - scene.bounds_lo.x, 1e-20f);
  const float ey = std::max(scene.bounds_hi.y - scene.bounds_lo.y, 1e-20f);
  const float ez = std::max(scene.bounds_hi.z - scene.bounds_lo.z, 1e-20f);
  g.inv_cell = make_float3(cells / ex, cells / ey, cells / ez);
  g.cell_bits = m_sort_bits;
  g.use_octant = use_octant ? 1u : 0u;
  m_sort_bins.reserve(size_t(1) << (3 * m_sort_bits + 3));
  launch_coherence_sort(m_stream, wb, g, which, m_sort_keys.get(), m_sort_bins.get(), m_sort_out.get());
  m_launches += 2;  // three kernels, one of them counted by stage()
  return m_sort_out.get();
}

Integrator::~Integrator()
{
  for (auto& t : m_timed) {
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  for (auto e : m_event_pool) cudaEventDestroy(e);
}

cudaEvent_t Integrator::get_event()
{
  if (!m_event_pool.empty()) {
    cudaEvent_t e = m_event_pool.back();
    m_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  FR_CUDA_CHECK(cudaEventCreate(&e));
  return e;
}

namespace
{
const char* const kStageNames[STAGE_COUNT] = {"generate", "trace_closest", "shade", "trace_shadow",
                                              "trace_light", "advance", "film"};
}

template <typename F>
void Integrator::stage(int id, F&& launch)
{
  FR_NVTX_RANGE(kStageNames[id]);
  if (!m_time_stages) {
    launch();
  } else {
    TimedLaunch t{id, get_event(), get_event()};
    FR_CUDA_CHECK(cudaEventRecord(t.e0, m_stream));
    launch();
    FR_CUDA_CHECK(cudaEventRecord(t.e1, m_stream));
    m_timed.push_back(t);
  }
  m_launches++;
}

StageTimes Integrator::stage_times()
{
  StageTimes out;
  FR_CUDA_CHECK(cudaStreamSynchronize(m_stream));
  for (auto& t : m_timed) {
    float ms = 0.0f;
    FR_CUDA_CHECK(cudaEventElapsedTime(&ms, t.e0, t.e1));
    out.ms[t.stage] += ms;
    out.launches[t.stage]++;
    m_event_pool.push_back(t.e0);
    m_event_pool.push_back(t.e1);
  }
  m_timed.clear();
  return out;
}

void Integrator::ensure_capacity(size_t n_slots)
{
  if (m_ctl.size() == 0) {
    m_ctl.alloc(1);
    m_ctl.zero(m_stream);
  }
  if (n_slots <= m_capacity) return;
  // buffers are in use by work already queued on the stream
  FR_CUDA_CHECK(cudaStreamSynchronize(m_stream));
  // Not valid until every buffer below has its new size: if one allocation throws (a 64 Mi-path wave is ~25 GB),
  // the capacity stays 0 and all wave buffers are released, so a later render() with a smaller wave reallocates
  // everything instead of launching on a half-grown set.
  m_capacity = 0;
  try {
    grow_wave_buffers(n_slots);
  } catch (...) {
    release_wave_buffers();
    throw;
  }
  m_capacity = n_slots;
}

// Buffers only some renders need -- the first-hit AOV words (48 B / path), the sun and area-light NEE queues
// (48 B / path each), the coherence-sort scratch (8 B / path) -- are allocated when a render first needs them, at
// the capacity of the core set: a beauty-only render of a scene without emitters keeps 268 B / path instead of 372.
void Integrator::ensure_optional(const WaveNeeds& need)
{
  bool synced = false;
  auto grow = [&](auto& buf, bool wanted) {
    if (!wanted || buf.size() >= m_capacity) return;
    if (!synced) FR_CUDA_CHECK(cudaStreamSynchronize(m_stream));
    synced = true;
    buf.alloc(m_capacity);
  };
  grow(m_aov0, need.aov);
  grow(m_aov1, need.aov);
  grow(m_aov2, need.aov);
  grow(m_shadow[0], need.sun_queue);
  grow(m_shadow[2], need.area_queue);
  grow(m_sort_keys, need.sort);
  grow(m_sort_out, need.sort);
  m_state_bytes = m_capacity * kWaveBytesCore + m_aov0.bytes() + m_aov1.bytes() + m_aov2.bytes() + m_shadow[0].bytes() +
                  m_shadow[2].bytes() + m_sort_keys.bytes() + m_sort_out.bytes();
}

void Integrator::release_wave_buffers()
{
  m_ray_o.release();
  m_ray_d.release();
  m_hit.release();
  m_thr.release();
  m_L.release();
  m_aov0.release();
  m_aov1.release();
  m_aov2.release();
  m_queue[0].release();
  m_queue[1].release();
  for (auto& s : m_shadow) s.release();
  for (auto& q : m_class_queue) q.release();
  m_light.release();
  m_sort_keys.release();
  m_sort_out.release();
  m_state_bytes = 0;
}

void Integrator::grow_wave_buffers(size_t n_slots)
{
  m_ray_o.alloc(n_slots);
  m_ray_d.alloc(n_slots);
  m_hit.alloc(n_slots);
  m_thr.alloc(n_slots);
  m_L.alloc(n_slots);
  m_queue[0].alloc(n_slots);
  m_queue[1].alloc(n_slots);
  m_shadow[1].alloc(n_slots);
  for (auto& q : m_class_queue) q.alloc(n_slots);
  m_light.alloc(n_slots);
  // the optional sets follow on demand (ensure_optional); what exists is dropped so that it regrows to the new size
  m_aov0.release();
  m_aov1.release();
  m_aov2.release();
  m_shadow[0].release();
  m_shadow[2].release();
  m_sort_keys.release();
  m_sort_out.release();
}

void Integrator::render(const SceneView& scene, const fredholm::CameraParams& camera, uint32_t width,
                        uint32_t height, const fredholm::RenderLayer& layers, uint32_t sample_base,
                        uint32_t n_samples, uint32_t max_depth, uint32_t seed, int film_mode, uint32_t class_mask)
{
  if (width == 0 || height == 0 || n_samples == 0) return;
  // samples per warp: never more than the call's sample count needs (a 1-spp render keeps 8x4 tiles)
  uint32_t spw_log2 = m_spw_log2;
  while (spw_log2 > 0 && (1u << spw_log2) > n_samples) spw_log2--;
  const FilmGeom film = make_film_geom(width, height, spw_log2);
  WaveNeeds need;
  need.aov = (layers.position || layers.normal || layers.depth || layers.texcoord || layers.albedo) && !m_single_launch;
  need.sun_queue = scene.has_dir_light != 0;
  need.area_queue = scene.n_lights > 0;
  need.sort = m_sort_mask != 0;
  // a wave = whole sample groups (spw samples of every pixel), at least one
  size_t groups_per_wave = std::max<size_t>(
      1, std::min<size_t>(film_groups(film, n_samples), m_max_wave_paths / film.slots_per_group));
  if (groups_per_wave * film.slots_per_group > m_capacity) {
    // growing: never ask for more than the device can give (90 % of what is free plus what the wave holds now)
    size_t free_b = 0, total_b = 0;
    FR_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    const size_t fit = (size_t)(0.9 * (double)(free_b + m_state_bytes)) / (wave_bytes_per_slot(need) * film.slots_per_group);
    groups_per_wave = std::max<size_t>(1, std::min(groups_per_wave, fit));
  }
  const uint32_t per_wave = (uint32_t)(groups_per_wave << spw_log2);
  ensure_capacity(groups_per_wave * film.slots_per_group);
  ensure_optional(need);

  WaveBuffers wb;
  wb.ray_o = m_ray_o.get();
  wb.ray_d = m_ray_d.get();
  wb.hit = m_hit.get();
  wb.thr = m_thr.get();
  wb.L = m_L.get();
  wb.aov0 = m_aov0.get();
  wb.aov1 = m_aov1.get();
  wb.aov2 = m_aov2.get();
  wb.queue[0] = m_queue[0].get();
  wb.queue[1] = m_queue[1].get();
  for (int k = 0; k < 3; ++k) wb.shadow[k] = m_shadow[k].get();
  for (int c = 0; c < CLS_COUNT; ++c) wb.class_queue[c] = m_class_queue[c].get();
  wb.light = m_light.get();
  wb.ctl = m_ctl.get();
  wb.first_hit = nullptr;
  wb.pix_aov0 = wb.pix_aov1 = wb.pix_aov2 = nullptr;
  if (m_single_launch) {
    const size_t n_pixels = (size_t)width * height;
    m_first_hit.reserve(n_pixels);
    FR_CUDA_CHECK(cudaMemsetAsync(m_first_hit.get(), 0xff, n_pixels * sizeof(uint32_t), m_stream));
    for (int k = 0; k < 3; ++k) {
      m_pix_aov[k].reserve(n_pixels);
      FR_CUDA_CHECK(cudaMemsetAsync(m_pix_aov[k].get(), 0, n_pixels * sizeof(float4), m_stream));
    }
    wb.first_hit = m_first_hit.get();
    wb.pix_aov0 = m_pix_aov[0].get();
    wb.pix_aov1 = m_pix_aov[1].get();
    wb.pix_aov2 = m_pix_aov[2].get();
  }

  FR_NVTX_RANGE("render");
  for (uint32_t done = 0; done < n_samples; done += per_wave) {
    FR_NVTX_RANGE("wave");
    WaveParams wp;
    wp.film = film;
    wp.n_samples = std::min(per_wave, n_samples - done);
    wp.sample_base = sample_base + done;
    wp.max_depth = max_depth;
    wp.seed = seed;
    wp.want_aov = (layers.position || layers.normal || layers.depth || layers.texcoord || layers.albedo) ? 1u : 0u;
    wp.single_launch = m_single_launch ? 1u : 0u;
    wp.camera = camera;

    stage(STAGE_ADVANCE, [&] { launch_wave_begin(m_stream, wb, (unsigned long long)wp.n_samples * width * height); });
    stage(STAGE_GENERATE, [&] { launch_generate(m_stream, wp, wb); });
    for (uint32_t depth = 0; depth < max_depth; ++depth) {
#if FR_HAVE_NVTX
      char bounce_name[24];
      snprintf(bounce_name, sizeof(bounce_name), "bounce %u", depth);
      FR_NVTX_RANGE(bounce_name);
#endif
      // coherence sort (queue management, booked under "advance"): each queue is sorted right
      // before it is traced, so one scratch order buffer serves all of them
      const uint32_t* order = nullptr;
      auto sort_queue = [&](uint32_t bit, int which, bool use_octant) {
        order = nullptr;
        if (m_sort_mask & bit) stage(STAGE_ADVANCE, [&] { order = sorted(scene, wb, which, use_octant); });
      };
      if (depth > 0) sort_queue(1u, SORT_RADIANCE0 + (int)(depth & 1u), true);
      stage(STAGE_TRACE_CLOSEST, [&] { launch_trace_closest(m_stream, scene, wb, depth, order); });
      if (depth == 0 && m_single_launch) stage(STAGE_SHADE, [&] { launch_first_hit(m_stream, wp, wb); });
      if (depth == 0) stage(STAGE_SHADE, [&] { launch_miss(m_stream, wp, scene, wb); });
      for (int c = 0; c < CLS_MISS; ++c)
        if (class_mask & (1u << c)) stage(STAGE_SHADE, [&] { launch_shade(m_stream, wp, scene, wb, depth, c); });
      if (scene.has_dir_light) {
        sort_queue(2u, SORT_SHADOW0, false);  // all sun rays point the same way
        stage(STAGE_TRACE_SHADOW, [&] { launch_trace_shadow(m_stream, scene, wb, 0, order, depth == 0); });
      }
      sort_queue(4u, SORT_SHADOW1, true);
      stage(STAGE_TRACE_SHADOW, [&] { launch_trace_shadow(m_stream, scene, wb, 1, order); });
      if (scene.n_lights > 0) {
        sort_queue(8u, SORT_SHADOW2, true);
        stage(STAGE_TRACE_SHADOW, [&] { launch_trace_shadow(m_stream, scene, wb, 2, order); });
      }
      sort_queue(16u, SORT_LIGHT, true);
      stage(STAGE_TRACE_LIGHT, [&] { launch_trace_light(m_stream, scene, wb, order); });
      stage(STAGE_ADVANCE, [&] { launch_advance(m_stream, wb); });
    }
    stage(STAGE_FILM, [&] { launch_film(m_stream, wp, wb, layers, film_mode); });
  }
}

void Integrator::scale_layers(const fredholm::RenderLayer& layers, uint32_t n_pixels, float scale)
{
  launch_scale_layers(m_stream, layers, n_pixels, scale);
  m_launches++;
}

RenderStats Integrator::stats()
{
  RenderStats s;
  s.launches = m_launches;
  if (m_ctl.size() == 0) return s;
  WaveControl h;
  FR_CUDA_CHECK(cudaStreamSynchronize(m_stream));
  FR_CUDA_CHECK(cudaMemcpy(&h, m_ctl.get(), sizeof(h), cudaMemcpyDeviceToHost));
  s.paths = h.paths;
  s.rays_closest = h.rays_closest;
  s.rays_shadow = h.rays_shadow;
  s.rays_light = h.rays_light;
  s.rays_skipped = h.rays_skipped;
  for (int i = 0; i < 3; ++i) {
    s.nodes[i] = h.nodes[i];
    s.tris[i] = h.tris[i];
  }
  return s;
}

void Integrator::reset_stats()
{
  m_launches = 0;
  if (m_ctl.size()) {
    FR_CUDA_CHECK(cudaStreamSynchronize(m_stream));
    m_ctl.zero(m_stream);
  }
}

}  // namespace frd
