// Host driver of the wavefront integrator: owns the per-path state and the ray
// queues in HBM and sequences the stages of one render call
// (replaces the single optixLaunch of Renderer::render, renderer.h:730-733).
#pragma once
#include <cstdint>
#include <vector>

#include "cuda_util.h"
#include "wavefront.h"
#include "wavefront_kernels.h"

namespace frd
{

struct RenderStats {
  unsigned long long paths = 0;         // camera samples started
  unsigned long long rays_closest = 0;  // radiance rays traced
  unsigned long long rays_shadow = 0;   // visibility rays traced
  unsigned long long rays_light = 0;    // MIS rays traced
  unsigned long long rays_skipped = 0;  // rays the reference traces whose contribution is exactly zero (not traced)
  unsigned long long launches = 0;      // kernels launched by the integrator
  // counting builds (set_traversal_counting): nodes visited / triangles tested per ray type
  unsigned long long nodes[3] = {0, 0, 0}, tris[3] = {0, 0, 0};
};

enum Stage : int {
  STAGE_GENERATE = 0,
  STAGE_TRACE_CLOSEST,
  STAGE_SHADE,
  STAGE_TRACE_SHADOW,
  STAGE_TRACE_LIGHT,
  STAGE_ADVANCE,
  STAGE_FILM,
  STAGE_COUNT
};

struct StageTimes {
  double ms[STAGE_COUNT] = {};
  unsigned long long launches[STAGE_COUNT] = {};
};

class Integrator
{
 public:
  explicit Integrator(cudaStream_t stream);

  // paths kept in flight per wave; rounded down to whole samples (at least one)
  void set_max_wave_paths(size_t n) { m_max_wave_paths = n; }
  // single-launch mode: one render() call behaves like ONE reference launch of n_samples (payload.firsthit and
  // the first-hit AOVs outlive the sample loop, pt.cu:432-433, 744-759); default off = one launch per sample
  void set_single_launch(bool on) { m_single_launch = on; }
  bool single_launch() const { return m_single_launch; }
  size_t max_wave_paths() const { return m_max_wave_paths; }
  // samples of one pixel block that share a warp (1, 2, 4, ... 32; wavefront.h FilmGeom): camera rays and
  // first-bounce shadow rays of a warp then form a one-pixel beam.  Does not change any sample's value.
  void set_samples_per_warp(uint32_t spw);
  uint32_t samples_per_warp() const { return 1u << m_spw_log2; }
  static constexpr uint32_t kDefaultSamplesPerWarp = 1;

  // Renders samples [sample_base, sample_base + n_samples) of every pixel into
  // `layers` (device pointers).  Asynchronous on the stream.  class_mask: bit c set if
  // the scene has materials of ShadeClass c (only those shade kernels are launched).
  void render(const SceneView& scene, const fredholm::CameraParams& camera, uint32_t width, uint32_t height,
              const fredholm::RenderLayer& layers, uint32_t sample_base, uint32_t n_samples, uint32_t max_depth,
              uint32_t seed, int film_mode, uint32_t class_mask);

  void scale_layers(const fredholm::RenderLayer& layers, uint32_t n_pixels, float scale);

  // blocks until the stream is idle, then reads the device counters
  RenderStats stats();
  void reset_stats();

  size_t state_bytes() const { return m_state_bytes; }

  // Coherence sort (sort.cu).  queue_mask: bit 0 radiance rays (bounces >= 1), bit 1 sun NEE rays,
  // bit 2 sky NEE rays, bit 3 area-light NEE rays, bit 4 MIS rays.  cell_bits: 2^bits origin
  // cells per axis.  Defaults come from FRD_SORT / FRD_SORT_BITS.
  void set_coherence_sort(uint32_t queue_mask, uint32_t cell_bits);
  uint32_t coherence_sort_mask() const { return m_sort_mask; }

  // Two waves in flight on two streams (off by default): `max_wave_paths` is then split between them.  Sample
  // values and the order in which the film applies them do not change.  Ignored while stage timing is on (the
  // per-stage times would include the other stream's kernels) and in single-launch mode.
  void set_wave_overlap(bool on) { m_overlap = on; }
  bool wave_overlap() const { return m_overlap; }

  // Wave compaction (on by default).  A render of several waves pays the late bounces of every wave -- launches
  // that carry a few per cent of the paths and run at the latency of one ray, ~3.7 ms per wave of the bench frame.
  // With compaction a wave runs its first `depth` bounces, then the paths it still has alive move to a dense
  // straggler set (4 float4 per path) and the wave's slots are free for the next samples; the stragglers of up to
  // kMaxWavesPerPass waves finish their remaining bounces together, their radiance goes back to the slot it came
  // from and the film applies the waves in sample order as before.  Every sample's arithmetic is unchanged, the
  // image is bit-identical to the uncompacted render.  The host waits once per wave for the number of paths that
  // go on (read on a second stream under that bounce's visibility launches, the GPU does not idle); a wave that
  // keeps more than the straggler set holds finishes in place.
  // With first-hit AOV layers bound, those go to the film when their wave has finished its first bounces and the
  // beauty layer at the end of the pass.  Not used in single-launch mode, with wave overlap, or for a one-wave render.
  void set_wave_compaction(bool on, uint32_t depth = kDefaultCompactionDepth)
  {
    m_compaction = on;
    m_compaction_depth = depth < 1u ? 1u : depth;
  }
  bool wave_compaction() const { return m_compaction; }
  static constexpr uint32_t kDefaultCompactionDepth = 3;
  static constexpr uint32_t kMaxWavesPerPass = 8;
  // straggler slots per wave slot (1/4: the bench scene keeps 5 % of its paths after three bounces)
  static constexpr size_t kStragglerDivisor = 4;

  // per-stage device time (CUDA events around every launch, on the launching stream)
  void set_stage_timing(bool on) { m_time_stages = on; }
  StageTimes stage_times();  // synchronises, returns and clears the accumulated times
  ~Integrator();

 private:
  // wave state per path slot.  Core set: 5 float4 words (ray origin, direction, hit, throughput, radiance),
  // 2 + CLS_COUNT queue entries, the sky-NEE and MIS ray records = 220 B; optional sets: first-hit AOV words
  // (3 float4), sun / area-light NEE records (48 B each), coherence-sort scratch (8 B)
  static constexpr size_t kWaveBytesCore = 5 * sizeof(float4) + (2 + CLS_COUNT) * sizeof(uint32_t) + sizeof(ShadowRay) +
                                           sizeof(LightRay);
  struct WaveNeeds {
    bool aov = false, sun_queue = false, area_queue = false, sort = false;
  };
  static constexpr size_t wave_bytes_per_slot(const WaveNeeds& n)
  {
    return kWaveBytesCore + (n.aov ? 3 * sizeof(float4) : 0) + (n.sun_queue ? sizeof(ShadowRay) : 0) +
           (n.area_queue ? sizeof(ShadowRay) : 0) + (n.sort ? 2 * sizeof(uint32_t) : 0);
  }

  // Everything one wave in flight owns: path state, queues, control block, sort scratch.  The integrator keeps
  // two of them so that (set_wave_overlap) consecutive waves of a render can run on two streams, the late,
  // nearly empty bounces of one under the full launches of the next.
  struct WaveSet {
    DevBuf<float4> ray_o, ray_d, hit, thr, L, aov0, aov1, aov2;
    DevBuf<uint32_t> queue[2];
    DevBuf<uint32_t> class_queue[CLS_COUNT];
    DevBuf<ShadowRay> shadow[3];
    DevBuf<LightRay> light;
    DevBuf<WaveControl> ctl;
    DevBuf<uint32_t> sort_keys, sort_out, sort_bins;
    DevBuf<uint32_t> origin;  // straggler set only (WaveBuffers::origin)
    size_t capacity = 0;    // path slots of the core set; 0 while the set is not valid
    size_t L_capacity = 0;  // slots of the radiance array: with wave compaction it holds several waves' worth

    void grow_core(size_t n_slots, size_t l_slots);
    void release();
    size_t bytes() const;
    WaveBuffers view() const;
  };
  // l_slots (>= n_slots): size of the radiance array; with_origin: the set is a straggler set
  void ensure_capacity(WaveSet& set, size_t n_slots, const WaveNeeds& need, size_t l_slots = 0, bool with_origin = false);
  void sync_all_streams();

  cudaStream_t m_stream;
  cudaStream_t m_aux_stream = nullptr;            // second wave in flight (created on first use)
  cudaEvent_t m_ev_start = nullptr, m_ev_film[2] = {nullptr, nullptr}, m_ev_alive = nullptr;
  uint32_t* m_alive_host = nullptr;  // pinned
  bool m_overlap = false;
  bool m_compaction = true;
  uint32_t m_compaction_depth = kDefaultCompactionDepth;
  size_t m_max_wave_paths = size_t(1) << 26;  // 64 Mi paths in flight (17.6 GB of wave state for a beauty-only frame)
  uint32_t m_spw_log2 = 0;
  size_t m_state_bytes = 0;
  unsigned long long m_launches = 0;

  // Stage timing: consecutive launches on one stream share an event (the end of one is the start of the next), so
  // a timed frame records one event per launch, not two.
  struct TimedLaunch {
    int stage;
    cudaEvent_t e0, e1;
    bool owns_e0;
  };
  cudaEvent_t m_chain_event = nullptr;  // end of the last timed launch ...
  cudaStream_t m_chain_stream = nullptr;  // ... on this stream; reset whenever other work may have been queued
  bool m_time_stages = false;
  std::vector<TimedLaunch> m_timed;
  std::vector<cudaEvent_t> m_event_pool;
  cudaEvent_t get_event();
  template <typename F>
  void stage(cudaStream_t s, int id, F&& launch);

  WaveSet m_set[2];
  bool m_single_launch = false;
  DevBuf<uint32_t> m_first_hit;           // [n_pixels], single-launch mode only
  DevBuf<float4> m_pix_aov[3];            // [n_pixels] each

  uint32_t m_sort_mask, m_sort_bits;
  const uint32_t* sorted(cudaStream_t s, WaveSet& set, const SceneView& scene, const WaveBuffers& wb, int which, bool use_octant);
  void render_wave(cudaStream_t s, WaveSet& set, const WaveBuffers& wb, const WaveParams& wp, const SceneView& scene,
                   uint32_t class_mask);
  void start_wave(cudaStream_t s, const WaveBuffers& wb, const WaveParams& wp);
  // probe_alive: after the shade launches of the last bounce of the range, the number of paths that go on
  // (WaveControl::n[Q_NEXT]) is copied to m_alive_host on the second stream, under the visibility launches that
  // follow on `s`; wait for m_ev_alive before reading it
  void run_bounces(cudaStream_t s, WaveSet& set, const WaveBuffers& wb, const WaveParams& wp, const SceneView& scene,
                   uint32_t class_mask, uint32_t depth_begin, uint32_t depth_end, bool probe_alive = false);
  void ensure_aux_stream();
  void render_compacted(const SceneView& scene, WaveParams wp, const fredholm::RenderLayer& layers, uint32_t n_samples,
                        uint32_t per_wave, uint32_t per_pass, int film_mode, uint32_t class_mask);
};

}  // namespace frd
