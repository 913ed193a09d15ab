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
  m_overlap = env_u32("FRD_WAVE_OVERLAP", 0u) != 0u;
  set_wave_compaction(env_u32("FRD_WAVE_COMPACTION", 1u) != 0u, env_u32("FRD_WAVE_COMPACTION_DEPTH", kDefaultCompactionDepth));
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
const uint32_t* Integrator::sorted(cudaStream_t s, WaveSet& set, const SceneView& scene, const WaveBuffers& wb, int which,
                                   bool use_octant)
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
  set.sort_bins.reserve(size_t(1) << (3 * m_sort_bits + 3));
  launch_coherence_sort(s, wb, g, which, set.sort_keys.get(), set.sort_bins.get(), set.sort_out.get());
  m_launches += 2;  // three kernels, one of them counted by stage()
  return set.sort_out.get();
}

Integrator::~Integrator()
{
  for (auto& t : m_timed) {
    if (t.owns_e0) cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  for (auto e : m_event_pool) cudaEventDestroy(e);
  if (m_ev_start) cudaEventDestroy(m_ev_start);
  if (m_ev_alive) cudaEventDestroy(m_ev_alive);
  if (m_alive_host) cudaFreeHost(m_alive_host);
  for (auto e : m_ev_film)
    if (e) cudaEventDestroy(e);
  if (m_aux_stream) {
    cudaStreamSynchronize(m_aux_stream);
    cudaStreamDestroy(m_aux_stream);
  }
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
void Integrator::stage(cudaStream_t s, int id, F&& launch)
{
  FR_NVTX_RANGE(kStageNames[id]);
  if (!m_time_stages) {
    launch();
  } else {
    const bool chained = m_chain_event != nullptr && m_chain_stream == s;
    TimedLaunch t{id, chained ? m_chain_event : get_event(), get_event(), !chained};
    if (!chained) FR_CUDA_CHECK(cudaEventRecord(t.e0, s));
    launch();
    FR_CUDA_CHECK(cudaEventRecord(t.e1, s));
    m_timed.push_back(t);
    m_chain_event = t.e1;
    m_chain_stream = s;
  }
  m_launches++;
}

void Integrator::sync_all_streams()
{
  FR_CUDA_CHECK(cudaStreamSynchronize(m_stream));
  if (m_aux_stream) FR_CUDA_CHECK(cudaStreamSynchronize(m_aux_stream));
}

StageTimes Integrator::stage_times()
{
  StageTimes out;
  sync_all_streams();
  for (auto& t : m_timed) {
    float ms = 0.0f;
    FR_CUDA_CHECK(cudaEventElapsedTime(&ms, t.e0, t.e1));
    out.ms[t.stage] += ms;
    out.launches[t.stage]++;
    if (t.owns_e0) m_event_pool.push_back(t.e0);
    m_event_pool.push_back(t.e1);
  }
  m_timed.clear();
  m_chain_event = nullptr;
  return out;
}

// ---- wave sets -------------------------------------------------------------------------------------------
void Integrator::WaveSet::grow_core(size_t n_slots, size_t l_slots)
{
  ray_o.alloc(n_slots);
  ray_d.alloc(n_slots);
  hit.alloc(n_slots);
  thr.alloc(n_slots);
  L.alloc(l_slots);
  queue[0].alloc(n_slots);
  queue[1].alloc(n_slots);
  shadow[1].alloc(n_slots);
  for (auto& q : class_queue) q.alloc(n_slots);
  light.alloc(n_slots);
  // the optional buffers follow on demand; what exists is dropped so that it regrows to the new size
  aov0.release();
  aov1.release();
  aov2.release();
  shadow[0].release();
  shadow[2].release();
  sort_keys.release();
  sort_out.release();
  origin.release();
}

void Integrator::WaveSet::release()
{
  ray_o.release();
  ray_d.release();
  hit.release();
  thr.release();
  L.release();
  aov0.release();
  aov1.release();
  aov2.release();
  queue[0].release();
  queue[1].release();
  for (auto& s : shadow) s.release();
  for (auto& q : class_queue) q.release();
  light.release();
  sort_keys.release();
  sort_out.release();
  origin.release();
  capacity = 0;
  L_capacity = 0;
}

size_t Integrator::WaveSet::bytes() const
{
  size_t b = ray_o.bytes() + ray_d.bytes() + hit.bytes() + thr.bytes() + L.bytes() + aov0.bytes() + aov1.bytes() + aov2.bytes() +
             queue[0].bytes() + queue[1].bytes() + light.bytes() + sort_keys.bytes() + sort_out.bytes() + origin.bytes();
  for (const auto& s : shadow) b += s.bytes();
  for (const auto& q : class_queue) b += q.bytes();
  return b;
}

WaveBuffers Integrator::WaveSet::view() const
{
  WaveBuffers wb;
  wb.ray_o = ray_o.get();
  wb.ray_d = ray_d.get();
  wb.hit = hit.get();
  wb.thr = thr.get();
  wb.L = L.get();
  wb.aov0 = aov0.get();
  wb.aov1 = aov1.get();
  wb.aov2 = aov2.get();
  wb.queue[0] = queue[0].get();
  wb.queue[1] = queue[1].get();
  for (int k = 0; k < 3; ++k) wb.shadow[k] = shadow[k].get();
  for (int c = 0; c < CLS_COUNT; ++c) wb.class_queue[c] = class_queue[c].get();
  wb.light = light.get();
  wb.ctl = ctl.get();
  wb.first_hit = nullptr;
  wb.pix_aov0 = wb.pix_aov1 = wb.pix_aov2 = nullptr;
  wb.origin = nullptr;  // set by the caller for a straggler set
  wb.origin_stride = wb.origin_samples = 0;
  return wb;
}

// Core buffers for n_slots paths plus the optional ones this render needs -- the first-hit AOV words (48 B / path),
// the sun and area-light NEE queues (48 B / path each), the coherence-sort scratch (8 B / path): a beauty-only
// render of a scene without emitters keeps 268 B / path instead of 372.
void Integrator::ensure_capacity(WaveSet& set, size_t n_slots, const WaveNeeds& need, size_t l_slots, bool with_origin)
{
  l_slots = std::max(l_slots, n_slots);
  if (set.ctl.size() == 0) {
    set.ctl.alloc(1);
    set.ctl.zero(m_stream);
  }
  bool synced = false;
  auto sync_once = [&] {
    // the buffers are in use by work already queued
    if (!synced) sync_all_streams();
    synced = true;
  };
  if (n_slots > set.capacity) {
    sync_once();
    // Not valid until every buffer has its new size: if one allocation throws (a 64 Mi-path wave is tens of GB),
    // the capacity stays 0 and the set is released, so a later render() with a smaller wave reallocates
    // everything instead of launching on a half-grown set.
    set.capacity = 0;
    try {
      set.grow_core(n_slots, std::max(l_slots, set.L_capacity));
    } catch (...) {
      set.release();
      throw;
    }
    set.capacity = n_slots;
    set.L_capacity = std::max(l_slots, set.L_capacity);
  } else if (l_slots > set.L_capacity) {
    sync_once();
    set.L.alloc(l_slots);
    set.L_capacity = l_slots;
  }
  auto grow = [&](auto& buf, bool wanted) {
    if (!wanted || buf.size() >= set.capacity) return;
    sync_once();
    buf.alloc(set.capacity);
  };
  grow(set.aov0, need.aov);
  grow(set.aov1, need.aov);
  grow(set.aov2, need.aov);
  grow(set.shadow[0], need.sun_queue);
  grow(set.shadow[2], need.area_queue);
  grow(set.sort_keys, need.sort);
  grow(set.sort_out, need.sort);
  grow(set.origin, with_origin);
  m_state_bytes = m_set[0].bytes() + m_set[1].bytes();
}

// one wave: camera rays, max_depth bounces of { closest hit, shade per material class, visibility rays, MIS rays }
// (everything up to, not including, the film)
void Integrator::render_wave(cudaStream_t s, WaveSet& set, const WaveBuffers& wb, const WaveParams& wp, const SceneView& scene,
                             uint32_t class_mask)
{
  FR_NVTX_RANGE("wave");
  start_wave(s, wb, wp);
  run_bounces(s, set, wb, wp, scene, class_mask, 0, wp.max_depth);
}

void Integrator::start_wave(cudaStream_t s, const WaveBuffers& wb, const WaveParams& wp)
{
  stage(s, STAGE_ADVANCE, [&] {
    launch_wave_begin(s, wb, (unsigned long long)wp.n_samples * wp.film.width * wp.film.height);
  });
  stage(s, STAGE_GENERATE, [&] { launch_generate(s, wp, wb); });
}

void Integrator::ensure_aux_stream()
{
  if (m_aux_stream) return;
  FR_CUDA_CHECK(cudaStreamCreateWithFlags(&m_aux_stream, cudaStreamNonBlocking));
  FR_CUDA_CHECK(cudaEventCreateWithFlags(&m_ev_start, cudaEventDisableTiming));
  for (auto& e : m_ev_film) FR_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  FR_CUDA_CHECK(cudaEventCreateWithFlags(&m_ev_alive, cudaEventDisableTiming));
  FR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&m_alive_host), sizeof(uint32_t), cudaHostAllocDefault));
}

void Integrator::run_bounces(cudaStream_t s, WaveSet& set, const WaveBuffers& wb, const WaveParams& wp, const SceneView& scene,
                             uint32_t class_mask, uint32_t depth_begin, uint32_t depth_end, bool probe_alive)
{
  for (uint32_t depth = depth_begin; depth < depth_end; ++depth) {
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
      if (m_sort_mask & bit) stage(s, STAGE_ADVANCE, [&] { order = sorted(s, set, scene, wb, which, use_octant); });
    };
    if (depth > 0) sort_queue(1u, SORT_RADIANCE0 + (int)(depth & 1u), true);
    stage(s, STAGE_TRACE_CLOSEST, [&] { launch_trace_closest(s, scene, wb, depth, order); });
    if (depth == 0 && m_single_launch) stage(s, STAGE_SHADE, [&] { launch_first_hit(s, wp, wb); });
    if (depth == 0) stage(s, STAGE_SHADE, [&] { launch_miss(s, wp, scene, wb); });
    for (int c = 0; c < CLS_MISS; ++c)
      if (class_mask & (1u << c)) stage(s, STAGE_SHADE, [&] { launch_shade(s, wp, scene, wb, depth, c); });
    if (probe_alive && depth + 1 == depth_end) {
      FR_CUDA_CHECK(cudaEventRecord(m_ev_start, s));
      FR_CUDA_CHECK(cudaStreamWaitEvent(m_aux_stream, m_ev_start, 0));
      FR_CUDA_CHECK(cudaMemcpyAsync(m_alive_host, &wb.ctl->n[Q_NEXT], sizeof(uint32_t), cudaMemcpyDeviceToHost, m_aux_stream));
      FR_CUDA_CHECK(cudaEventRecord(m_ev_alive, m_aux_stream));
    }
    if (scene.has_dir_light) {
      sort_queue(2u, SORT_SHADOW0, false);  // all sun rays point the same way
      stage(s, STAGE_TRACE_SHADOW, [&] { launch_trace_shadow(s, scene, wb, 0, order, depth == 0); });
    }
    sort_queue(4u, SORT_SHADOW1, true);
    stage(s, STAGE_TRACE_SHADOW, [&] { launch_trace_shadow(s, scene, wb, 1, order); });
    if (scene.n_lights > 0) {
      sort_queue(8u, SORT_SHADOW2, true);
      stage(s, STAGE_TRACE_SHADOW, [&] { launch_trace_shadow(s, scene, wb, 2, order); });
    }
    sort_queue(16u, SORT_LIGHT, true);
    stage(s, STAGE_TRACE_LIGHT, [&] { launch_trace_light(s, scene, wb, order); });
    stage(s, STAGE_ADVANCE, [&] { launch_advance(s, wb); });
  }
}

void Integrator::render(const SceneView& scene, const fredholm::CameraParams& camera, uint32_t width,
                        uint32_t height, const fredholm::RenderLayer& layers, uint32_t sample_base,
                        uint32_t n_samples, uint32_t max_depth, uint32_t seed, int film_mode, uint32_t class_mask)
{
  if (width == 0 || height == 0 || n_samples == 0) return;
  m_chain_event = nullptr;  // stage timing: the caller may have queued work on the stream since the last launch
  // samples per warp: never more than the call's sample count needs (a 1-spp render keeps 8x4 tiles)
  uint32_t spw_log2 = m_spw_log2;
  while (spw_log2 > 0 && (1u << spw_log2) > n_samples) spw_log2--;
  const FilmGeom film = make_film_geom(width, height, spw_log2);
  WaveNeeds need;
  need.aov = (layers.position || layers.normal || layers.depth || layers.texcoord || layers.albedo) && !m_single_launch;
  need.sun_queue = scene.has_dir_light != 0;
  need.area_queue = scene.n_lights > 0;
  need.sort = m_sort_mask != 0;
  // paths in flight = whole sample groups (spw samples of every pixel), at least one
  const size_t groups_total = film_groups(film, n_samples);
  size_t groups_in_flight = std::max<size_t>(1, std::min<size_t>(groups_total, m_max_wave_paths / film.slots_per_group));
  // wave compaction: a pass of up to kMaxWavesPerPass waves keeps all their radiance arrays and a straggler set
  const bool can_compact = m_compaction && !m_overlap && !m_single_launch && max_depth > m_compaction_depth;
  auto waves_per_pass = [&](size_t groups_per_wave) {
    const size_t n_waves = (groups_total + groups_per_wave - 1) / groups_per_wave;
    const size_t by_index = (size_t(1) << 32) / (groups_per_wave * film.slots_per_group);  // WaveBuffers::origin is 32 bits
    return !can_compact ? size_t(1) : std::max<size_t>(1, std::min({n_waves, (size_t)kMaxWavesPerPass, by_index}));
  };
  auto bytes_per_slot = [&](size_t per_pass) {
    const size_t b = wave_bytes_per_slot(need);
    WaveNeeds late_need = need;
    late_need.aov = false;
    const size_t late = wave_bytes_per_slot(late_need) + sizeof(uint32_t);
    return per_pass <= 1 ? b : b + sizeof(float4) * (per_pass - 1) + late / kStragglerDivisor + 1;
  };
  if (groups_in_flight * film.slots_per_group > m_set[0].capacity + m_set[1].capacity) {
    // growing: never ask for more than the device can give (90 % of what is free plus what the waves hold now)
    size_t free_b = 0, total_b = 0;
    FR_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    const size_t per_slot = bytes_per_slot(waves_per_pass(groups_in_flight));
    const size_t fit = (size_t)(0.9 * (double)(free_b + m_state_bytes)) / (per_slot * film.slots_per_group);
    groups_in_flight = std::max<size_t>(1, std::min(groups_in_flight, fit));
  }
  const size_t per_pass = waves_per_pass(groups_in_flight);
  if (per_pass >= 2) {
    const size_t n_slots = groups_in_flight * film.slots_per_group;
    ensure_capacity(m_set[0], n_slots, need, per_pass * n_slots);
    WaveNeeds late_need = need;
    late_need.aov = false;  // the first-hit words are complete after the first bounce and stay with the wave
    ensure_capacity(m_set[1], std::max<size_t>(n_slots / kStragglerDivisor, 1024), late_need, 0, true);
    WaveParams wp;
    wp.film = film;
    wp.n_samples = 0;
    wp.sample_base = sample_base;
    wp.max_depth = max_depth;
    wp.seed = seed;
    wp.want_aov = need.aov ? 1u : 0u;
    wp.single_launch = 0u;
    wp.camera = camera;
    render_compacted(scene, wp, layers, n_samples, (uint32_t)(groups_in_flight << spw_log2), (uint32_t)per_pass, film_mode,
                     class_mask);
    return;
  }
  // two waves in flight when asked for and when the render has at least two waves' worth of samples
  const bool overlap = m_overlap && !m_time_stages && !m_single_launch && groups_in_flight >= 2 && groups_total >= 2;
  const size_t groups_per_wave = overlap ? groups_in_flight / 2 : groups_in_flight;
  const uint32_t per_wave = (uint32_t)(groups_per_wave << spw_log2);
  const int n_sets = overlap && groups_total > groups_per_wave ? 2 : 1;
  for (int k = 0; k < n_sets; ++k) ensure_capacity(m_set[k], groups_per_wave * film.slots_per_group, need);
  if (n_sets == 1 && m_set[1].capacity && !m_overlap && !can_compact) {
    // overlap was switched off: give the second set back
    sync_all_streams();
    m_set[1].release();
    m_state_bytes = m_set[0].bytes();
  }

  WaveBuffers wb[2] = {m_set[0].view(), m_set[1].view()};
  if (m_single_launch) {
    const size_t n_pixels = (size_t)width * height;
    m_first_hit.reserve(n_pixels);
    FR_CUDA_CHECK(cudaMemsetAsync(m_first_hit.get(), 0xff, n_pixels * sizeof(uint32_t), m_stream));
    for (int k = 0; k < 3; ++k) {
      m_pix_aov[k].reserve(n_pixels);
      FR_CUDA_CHECK(cudaMemsetAsync(m_pix_aov[k].get(), 0, n_pixels * sizeof(float4), m_stream));
    }
    wb[0].first_hit = m_first_hit.get();
    wb[0].pix_aov0 = m_pix_aov[0].get();
    wb[0].pix_aov1 = m_pix_aov[1].get();
    wb[0].pix_aov2 = m_pix_aov[2].get();
  }

  cudaStream_t streams[2] = {m_stream, m_stream};
  if (n_sets == 2) {
    ensure_aux_stream();
    streams[1] = m_aux_stream;
    // the second stream starts after whatever the caller queued on the renderer's stream (layer clears)
    FR_CUDA_CHECK(cudaEventRecord(m_ev_start, m_stream));
    FR_CUDA_CHECK(cudaStreamWaitEvent(m_aux_stream, m_ev_start, 0));
  }

  FR_NVTX_RANGE("render");
  uint32_t wave = 0;
  for (uint32_t done = 0; done < n_samples; done += per_wave, ++wave) {
    const int k = n_sets == 2 ? (int)(wave & 1u) : 0;
    cudaStream_t s = streams[k];
    WaveParams wp;
    wp.film = film;
    wp.n_samples = std::min(per_wave, n_samples - done);
    wp.sample_base = sample_base + done;
    wp.max_depth = max_depth;
    wp.seed = seed;
    wp.want_aov = (layers.position || layers.normal || layers.depth || layers.texcoord || layers.albedo) ? 1u : 0u;
    wp.single_launch = m_single_launch ? 1u : 0u;
    wp.camera = camera;
    render_wave(s, m_set[k], wb[k], wp, scene, class_mask);
    // the film applies the waves in sample order (streaming mean): wave w's film runs after wave w-1's
    if (n_sets == 2 && wave > 0) FR_CUDA_CHECK(cudaStreamWaitEvent(s, m_ev_film[(wave - 1) & 1u], 0));
    stage(s, STAGE_FILM, [&] { launch_film(s, wp, wb[k], layers, film_mode); });
    if (n_sets == 2) FR_CUDA_CHECK(cudaEventRecord(m_ev_film[wave & 1u], s));
  }
  // whatever follows on the renderer's stream (read-back, post-process) sees the finished layers
  if (n_sets == 2 && wave > 0 && ((wave - 1) & 1u) == 1u) FR_CUDA_CHECK(cudaStreamWaitEvent(m_stream, m_ev_film[1], 0));
}

// Wave compaction (integrator.h, set_wave_compaction): passes of up to `per_pass` waves.  m_set[0] is the wave (its
// radiance array holds per_pass waves' worth), m_set[1] the straggler set.
void Integrator::render_compacted(const SceneView& scene, WaveParams wp, const fredholm::RenderLayer& layers, uint32_t n_samples,
                                  uint32_t per_wave, uint32_t per_pass, int film_mode, uint32_t class_mask)
{
  FR_NVTX_RANGE("render (wave compaction)");
  ensure_aux_stream();
  cudaStream_t s = m_stream;
  WaveSet& wave_set = m_set[0];
  WaveSet& late_set = m_set[1];
  const uint32_t stride = film_groups(wp.film, per_wave) * wp.film.slots_per_group;
  const WaveBuffers wb0 = wave_set.view();
  WaveBuffers late = late_set.view();
  late.origin = late_set.origin.get();
  late.origin_stride = stride;
  late.origin_samples = per_wave;
  const size_t late_capacity = late_set.capacity;
  const uint32_t depth_move = m_compaction_depth;
  const uint32_t first_sample = wp.sample_base;
  fredholm::RenderLayer first_hit_layers = layers, beauty_layer = {};
  first_hit_layers.beauty = nullptr;
  beauty_layer.beauty = layers.beauty;

  for (uint32_t done = 0; done < n_samples;) {
    // the stragglers' sample index is relative to the first wave of the pass
    WaveParams pass_wp = wp;
    pass_wp.sample_base = first_sample + done;
    WaveParams wave_wp[kMaxWavesPerPass];
    uint32_t n_waves = 0;
    size_t n_late = 0;
    stage(s, STAGE_ADVANCE, [&] { launch_wave_begin(s, late, 0ull); });
    auto finish_stragglers = [&] {
      if (n_late == 0) return;
      FR_NVTX_RANGE("stragglers");
      run_bounces(s, late_set, late, pass_wp, scene, class_mask, depth_move, wp.max_depth);
      stage(s, STAGE_ADVANCE, [&] { launch_migrate_back(s, late, (uint32_t)n_late, wave_set.L.get()); });
      stage(s, STAGE_ADVANCE, [&] { launch_wave_begin(s, late, 0ull); });
      n_late = 0;
    };
    auto wave_view = [&](uint32_t k) {
      WaveBuffers wb = wb0;
      wb.L = wave_set.L.get() + (size_t)k * stride;
      return wb;
    };
    while (n_waves < per_pass && done < n_samples) {
      FR_NVTX_RANGE("wave");
      WaveParams w = wp;
      w.n_samples = std::min(per_wave, n_samples - done);
      w.sample_base = first_sample + done;
      const WaveBuffers wb = wave_view(n_waves);
      start_wave(s, wb, w);
      // how many paths go on is known after the last shade launch: the host reads it on the second stream while
      // that bounce's visibility rays are still being traced, so the GPU does not wait for the decision below
      run_bounces(s, wave_set, wb, w, scene, class_mask, 0, depth_move, true);
      FR_CUDA_CHECK(cudaEventSynchronize(m_ev_alive));
      const uint32_t alive = *m_alive_host;
      if (alive > late_capacity) {
        // a scene that keeps its paths (an interior): this wave finishes where it is
        run_bounces(s, wave_set, wb, w, scene, class_mask, depth_move, wp.max_depth);
      } else if (alive > 0) {
        if (n_late + alive > late_capacity) finish_stragglers();
        stage(s, STAGE_ADVANCE, [&] { launch_migrate(s, wb, late, depth_move, n_waves * stride); });
        m_launches++;
        n_late += alive;
      }
      // the first-hit layers are complete after the first bounce and the wave's slots are about to be reused:
      // they go to the film now (in sample order, as every wave does this), the beauty layer at the end of the pass
      if (w.want_aov) stage(s, STAGE_FILM, [&] { launch_film(s, w, wb, first_hit_layers, film_mode); });
      wave_wp[n_waves++] = w;
      done += w.n_samples;
    }
    finish_stragglers();
    // the film applies the waves in sample order (streaming mean)
    for (uint32_t k = 0; k < n_waves; ++k) {
      WaveParams w = wave_wp[k];
      w.want_aov = 0u;
      stage(s, STAGE_FILM, [&] { launch_film(s, w, wave_view(k), beauty_layer, film_mode); });
    }
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
  sync_all_streams();
  for (const WaveSet& set : m_set) {
    if (set.ctl.size() == 0) continue;
    WaveControl h;
    FR_CUDA_CHECK(cudaMemcpy(&h, set.ctl.get(), sizeof(h), cudaMemcpyDeviceToHost));
    s.paths += h.paths;
    s.rays_closest += h.rays_closest;
    s.rays_shadow += h.rays_shadow;
    s.rays_light += h.rays_light;
    s.rays_skipped += h.rays_skipped;
    for (int i = 0; i < 3; ++i) {
      s.nodes[i] += h.nodes[i];
      s.tris[i] += h.tris[i];
    }
  }
  return s;
}

void Integrator::reset_stats()
{
  m_launches = 0;
  sync_all_streams();
  for (WaveSet& set : m_set)
    if (set.ctl.size()) set.ctl.zero(m_stream);
  FR_CUDA_CHECK(cudaStreamSynchronize(m_stream));
}

}  // namespace frd
