// Host-callable launchers of the wavefront stages (shade.cu, trace.cu).
#pragma once
#include <cuda_runtime.h>

#include "wavefront.h"

namespace frd
{

enum FilmMode : int {
  FILM_MEAN = 0,  // reference behaviour: streaming mean into the layers (pt.cu:480-501)
  FILM_SUM = 1,   // accumulate sums (multi-GPU sample slices; divide after the reduce)
};

// shade.cu
void launch_wave_begin(cudaStream_t s, const WaveBuffers& wb, unsigned long long n_paths);
void launch_generate(cudaStream_t s, const WaveParams& wp, const WaveBuffers& wb);
void launch_shade(cudaStream_t s, const WaveParams& wp, const SceneView& sc, const WaveBuffers& wb, uint32_t depth,
                  int cls);
void launch_miss(cudaStream_t s, const WaveParams& wp, const SceneView& sc, const WaveBuffers& wb);
void launch_first_hit(cudaStream_t s, const WaveParams& wp, const WaveBuffers& wb);
void launch_advance(cudaStream_t s, const WaveBuffers& wb);
// wave compaction: move the live paths of `src` (about to trace bounce `depth`) to the end of the straggler set
// `dst` (two launches); origin_base = index of the wave * dst.origin_stride.  The caller has checked the capacity.
void launch_migrate(cudaStream_t s, const WaveBuffers& src, const WaveBuffers& dst, uint32_t depth, uint32_t origin_base);
// radiance of the first n straggler slots back into the waves' radiance arrays
void launch_migrate_back(cudaStream_t s, const WaveBuffers& late, uint32_t n, float4* L_all);
void launch_film(cudaStream_t s, const WaveParams& wp, const WaveBuffers& wb, const fredholm::RenderLayer& layers,
                 int film_mode);
void launch_scale_layers(cudaStream_t s, const fredholm::RenderLayer& layers, uint32_t n_pixels, float scale);

// trace.cu
// `order` (optional, device): the order in which the queue items are traced (coherence sort);
// for the radiance queue it holds path slots, for the record queues item indices
void launch_trace_closest(cudaStream_t s, const SceneView& sc, const WaveBuffers& wb, uint32_t depth,
                          const uint32_t* order = nullptr);
// coherent: the queue is in beam order (first-bounce sun rays), see FRD_REFILL_LANES_COHERENT
void launch_trace_shadow(cudaStream_t s, const SceneView& sc, const WaveBuffers& wb, int which,
                         const uint32_t* order = nullptr, bool coherent = false);
void launch_trace_light(cudaStream_t s, const SceneView& sc, const WaveBuffers& wb, const uint32_t* order = nullptr);
// counting instantiations of the three stages: nodes visited / triangles tested go to WaveControl::nodes / tris
// (measurement only; process-wide switch)
void set_traversal_counting(bool on);

// sort.cu: counting sort of queue `which` (SortQueue) by origin cell + direction octant.
// keys: [queue size] scratch, bins: [sort_bins(g)] scratch, out: [queue size] sorted order
void launch_coherence_sort(cudaStream_t s, const WaveBuffers& wb, const SortGrid& g, int which, uint32_t* keys,
                           uint32_t* bins, uint32_t* out);
// stand-alone batch query (tests, tools): rays are (o.xyz, d.xyz) per entry;
// out_id = (instance, primitive) or 0xffffffff, out_tuv = (t, u, v).  All HOST pointers.
// counters (optional, host): nodes visited, triangles tested
void trace_batch_closest(const SceneView& sc, const uint32_t* d_submesh_offsets, const float* rays_host, uint32_t n,
                         float tmin, float tmax, uint32_t* out_id_host, float* out_tuv_host,
                         unsigned long long* counters2_host);

// unit-test entry points (shade.cu)
void test_sampler(uint32_t width, uint32_t height, uint32_t seed, uint32_t image_idx, uint32_t n_spp,
                  const char* kinds, float* out_host, uint32_t n_out);
void test_bsdf(const float* in_host, uint32_t n, float* out_host);
void test_sky(const SceneView& sc, const float* dirs_host, uint32_t n, float* out_host);
void test_primary_rays(const WaveParams& wp, float* out_host);

}  // namespace frd
