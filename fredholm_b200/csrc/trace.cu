// Traversal stages of the wavefront integrator: persistent-thread kernels that pull
// rays from the HBM queues in warp-sized batches and walk the CWBVH (bvh.cuh).
//
//   k_trace_closest  radiance rays    (reference trace_radiance, pt.cu:82-94)
//   k_trace_shadow   visibility rays  (trace_shadow + __miss__/__closesthit__shadow,
//                                      pt.cu:96-109, 525-529, 946-950)
//   k_trace_light    MIS rays         (trace_light + __closesthit__light / __miss__light,
//                                      pt.cu:111-123, 531-543, 952-998) with the emitter
//                                      record and MIS weight resolved in the epilogue
// All three honour the reference's alpha cut-out (any-hit programs, pt.cu:545-678).
#include <algorithm>
#include <cstdlib>

#include "cuda_util.h"
#include "queue.cuh"
#include "surface.cuh"
#include "wavefront.h"
#include "wavefront_kernels.h"

namespace frd
{
namespace
{

constexpr int kBlock = 128;
#ifndef FRD_TRACE_BLOCKS
#define FRD_TRACE_BLOCKS 8  // resident CTAs per SM: 64 registers per thread
#endif
#ifndef FRD_TRACE_BLOCKS_TWO
#define FRD_TRACE_BLOCKS_TWO 6  // two-level instantiations (instance index, stack floor, record index): 80 registers
#endif

// alpha cut-out: candidate is ignored if base-colour alpha or the alpha map is < 0.5
struct AlphaTest {
  const SceneView* sc;
  FR_D bool operator()(uint32_t face, float u, float v) const
  {
    const fredholm::Material& m = sc->materials[sc->material_ids[face]];
    const uint3 idx = sc->indices[face];
    const float2 uv = bary2(sc->texcoords[idx.x], sc->texcoords[idx.y], sc->texcoords[idx.z], u, v);
    const SceneTex tex{sc->textures, sc->srgb_lut};
    bool accept = true;
    if (m.base_color_texture_id >= 0 && tex.fetch(m.base_color_texture_id, uv).w < 0.5f) accept = false;
    if (m.alpha_texture_id >= 0 && tex.fetch(m.alpha_texture_id, uv).x < 0.5f) accept = false;
    return accept;
  }
};

#define FR_DECLARE_STACK()                       \
  __shared__ uint2 s_stack[kSmemStack * kBlock]; \
  uint2* const stack_column = s_stack + threadIdx.x

// Counting builds of the stages (set_traversal_counting): nodes visited / triangles tested per ray type,
// summed over the retiring lanes (retire() runs with the warp converged) -- the measured N_node / N_tri of
// the issue-slot roofline.  The timed path uses the COUNT = false instantiations.
template <bool COUNT>
FR_D void count_retire(WaveControl* ctl, int type, bool has, const TraceCounters& c)
{
  if (!COUNT) return;
  const uint32_t nn = __reduce_add_sync(0xffffffffu, has ? c.nodes : 0u);
  const uint32_t nt = __reduce_add_sync(0xffffffffu, has ? c.tris : 0u);
  if (lane_id() == 0 && (nn | nt)) {
    atomicAdd(&ctl->nodes[type], (unsigned long long)nn);
    atomicAdd(&ctl->tris[type], (unsigned long long)nt);
  }
}

// ---- radiance rays: closest hit, then "sort by material" ---------------------------------
template <bool COUNT>
struct ClosestPolicy {
  const SceneView& sc;
  const WaveBuffers& wb;
  const uint32_t* q;
  uint32_t depth;
  uint32_t slot;
  FR_D AlphaTest anyhit() const { return AlphaTest{&sc}; }
  FR_D void load(uint32_t item, float3& o, float3& d, float& tmin, float& tmax)
  {
    slot = q[item];
    // path state and ray records are read once per bounce: streaming loads / stores (evict-first)
    // leave the L2 to the tree (-1.2 % any-hit time)
    const float4 ro = __ldcs(&wb.ray_o[slot]), rd = __ldcs(&wb.ray_d[slot]);
    o = f3(ro);
    d = f3(rd);
    tmin = 0.0f;
    tmax = 1e9f;
  }
  FR_D void world_ray(float3& o, float3& d) const
  {
    o = f3(wb.ray_o[slot]);
    d = f3(wb.ray_d[slot]);
  }
  FR_D void retire(bool has, const HitRecord& h, const TraceCounters& cnt)
  {
    count_retire<COUNT>(wb.ctl, 0, has, cnt);
    int cls = -1;
    if (has) {
      __stcs(&wb.hit[slot], make_float4(h.t, h.u, h.v, __uint_as_float(h.face)));
      // a miss only needs work for camera rays (sky seen directly, pt.cu:504-523)
      cls = h.face != kNoHit ? (int)sc.face_class[h.face] : (depth == 0 ? (int)CLS_MISS : -1);
    }
    // append the path to the shade queue of the class it hit: the retiring lanes group by class
    // (match_any), the first lane of every group reserves for its group, and all groups' atomics
    // are in flight together -- one round trip however many classes the warp retires
    const uint32_t appending = __ballot_sync(0xffffffffu, cls >= 0);
    if (cls >= 0) {
      const uint32_t peers = __match_any_sync(appending, cls);
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if ((int)lane_id() == leader) base = atomicAdd(&wb.ctl->n_class[cls], (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      wb.class_queue[cls][base + __popc(peers & ((1u << lane_id()) - 1u))] = slot;
    }
  }
};

template <bool COUNT, bool TWO>
__global__ void __launch_bounds__(kBlock, TWO ? FRD_TRACE_BLOCKS_TWO : FRD_TRACE_BLOCKS) k_trace_closest(SceneView sc, WaveBuffers wb, uint32_t depth, int refill, int tri_lanes,
                                                             const uint32_t* order)
{
  FR_DECLARE_STACK();
  ClosestPolicy<COUNT> pol{sc, wb, order ? order : wb.queue[depth & 1u], depth, 0u};
  trace_queue<false, COUNT, TWO>(sc.bvh, pol, &wb.ctl->cursor[0], wb.ctl->n[Q_CUR], stack_column, kBlock, refill, tri_lanes);
}

// ---- visibility rays: any hit; an unoccluded ray adds its contribution ----------------------
template <bool COUNT>
struct ShadowPolicy {
  const SceneView& sc;
  const WaveBuffers& wb;
  const float4* q;
  const uint32_t* order;
  uint32_t path;
  float3 c;
  uint32_t record;  // index of the ray record (read again by world_ray in two-level mode only)
  FR_D AlphaTest anyhit() const { return AlphaTest{&sc}; }
  FR_D void load(uint32_t item, float3& o, float3& d, float& tmin, float& tmax)
  {
    if (order) item = order[item];
    record = item;
    const float4 r0 = __ldcs(q + 3ull * item), r1 = __ldcs(q + 3ull * item + 1), r2 = __ldcs(q + 3ull * item + 2);
    o = f3(r0);
    d = f3(r1);
    tmin = 0.0f;
    tmax = r0.w;
    path = __float_as_uint(r1.w);
    c = f3(r2);
  }
  FR_D void world_ray(float3& o, float3& d) const
  {
    o = f3(q[3ull * record]);
    d = f3(q[3ull * record + 1]);
  }
  FR_D void retire(bool has, const HitRecord& h, const TraceCounters& cnt)
  {
    count_retire<COUNT>(wb.ctl, 1, has, cnt);
    if (has && h.face == kNoHit) {
      // one ray per path and kernel: plain read-modify-write, deterministic order
      float4 L = wb.L[path];
      L.x += c.x;
      L.y += c.y;
      L.z += c.z;
      wb.L[path] = L;
    }
  }
};

template <bool COUNT, bool TWO>
__global__ void __launch_bounds__(kBlock, TWO ? FRD_TRACE_BLOCKS_TWO : FRD_TRACE_BLOCKS) k_trace_shadow(SceneView sc, WaveBuffers wb, int which, int refill, int tri_lanes,
                                                            const uint32_t* order)
{
  FR_DECLARE_STACK();
  // which == 3: the MIS-ray queue holding visibility records (scenes without emitters, shade.cu)
  ShadowPolicy<COUNT> pol{sc, wb, reinterpret_cast<const float4*>(which < 3 ? wb.shadow[which] : (const ShadowRay*)wb.light), order,
                   0u, f3(0.f), 0u};
  trace_queue<true, COUNT, TWO>(sc.bvh, pol, &wb.ctl->cursor[2 + which], wb.ctl->n[Q_SHADOW0 + which], stack_column, kBlock,
                           refill, tri_lanes);
}

// ---- MIS rays: closest hit, emitter record + MIS weight in the epilogue -----------------------
template <bool COUNT>
struct LightPolicy {
  const SceneView& sc;
  const WaveBuffers& wb;
  const float4* q;
  const uint32_t* order;
  float3 o, d, w;
  float pdf_bsdf, cos_wi;
  uint32_t path;
  FR_D AlphaTest anyhit() const { return AlphaTest{&sc}; }
  FR_D void load(uint32_t item, float3& ro, float3& rd, float& tmin, float& tmax)
  {
    if (order) item = order[item];
    const float4 r0 = __ldcs(q + 3ull * item), r1 = __ldcs(q + 3ull * item + 1), r2 = __ldcs(q + 3ull * item + 2);
    ro = o = f3(r0);
    rd = d = f3(r1);
    tmin = 0.0f;
    tmax = 1e9f;
    pdf_bsdf = r0.w;
    path = __float_as_uint(r1.w);
    w = f3(r2);
    cos_wi = r2.w;
  }
  FR_D void world_ray(float3& ro, float3& rd) const
  {
    ro = o;
    rd = d;
  }
  FR_D void retire(bool has, const HitRecord& h, const TraceCounters& cnt)
  {
    count_retire<COUNT>(wb.ctl, 2, has, cnt);
    if (!has) return;
    float3 le = f3(0.0f);
    float pdf_light = cos_wi / kPi;
    bool contributes = true;
    if (h.face == kNoHit) {
      // __miss__light: environment radiance, cosine pdf of the sky NEE strategy
      le = sky_radiance(sc, d);
    } else {
      // __closesthit__light: an emitter record only for emissive, front-facing faces
      const fredholm::Material& m = sc.materials[sc.material_ids[h.face]];
      contributes = false;
      if (is_emissive(m)) {
        const FaceGeom g = load_face(sc, sc.indices[h.face], sc.face_submesh[h.face]);
        const float3 nl = bary3(g.n0, g.n1, g.n2, h.u, h.v);
        const float cos_l = dot(-d, nl);
        if (cos_l > 0.0f) {
          const float3 p = bary3(g.v0, g.v1, g.v2, h.u, h.v);
          const float2 uv = bary2(g.t0, g.t1, g.t2, h.u, h.v);
          const SceneTex tex{sc.textures, sc.srgb_lut};
          le = emission_of(m, tex, uv);
          const float area = 0.5f * length(cross(g.v1 - g.v0, g.v2 - g.v0));
          const float3 dp = p - o;
          const float r2d = dot(dp, dp);
          const float pdf_area = 1.0f / (sc.n_lights * area);
          pdf_light = r2d / fabsf(cos_l) * pdf_area;
          contributes = true;
        }
      }
    }
    if (contributes) {
      const float mis = pdf_bsdf / (pdf_bsdf + pdf_light);
      const float3 ww = saturate3(f3(w.x * mis, w.y * mis, w.z * mis));  // regularize_weight, pt.cu:373-376
      float4 L = wb.L[path];
      L.x += ww.x * le.x;
      L.y += ww.y * le.y;
      L.z += ww.z * le.z;
      wb.L[path] = L;
    }
  }
};

// (scenes without any emissive face never get here: their MIS rays are visibility rays with
// the sky contribution precomputed by the shade stage, traced by k_trace_shadow)
template <bool COUNT, bool TWO>
__global__ void __launch_bounds__(kBlock, TWO ? FRD_TRACE_BLOCKS_TWO : FRD_TRACE_BLOCKS) k_trace_light(SceneView sc, WaveBuffers wb, int refill, int tri_lanes,
                                                           const uint32_t* order)
{
  FR_DECLARE_STACK();
  LightPolicy<COUNT> pol{sc, wb, reinterpret_cast<const float4*>(wb.light), order, f3(0.f), f3(0.f), f3(0.f), 0.f, 0.f, 0u};
  trace_queue<false, COUNT, TWO>(sc.bvh, pol, &wb.ctl->cursor[5], wb.ctl->n[Q_LIGHT], stack_column, kBlock, refill, tri_lanes);
}

// stand-alone batch query for the parity tests: same driver and phases as the stages above
struct BatchPolicy {
  const SceneView& sc;
  const uint32_t* submesh_offsets;
  const float* rays;
  float tmin0, tmax0;
  uint32_t* out_id;
  float* out_tuv;
  unsigned long long* counters;
  uint32_t i;
  FR_D NoAnyHit anyhit() const { return NoAnyHit{}; }
  FR_D void load(uint32_t item, float3& o, float3& d, float& tmin, float& tmax)
  {
    i = item;
    o = f3(rays[6ull * i], rays[6ull * i + 1], rays[6ull * i + 2]);
    d = f3(rays[6ull * i + 3], rays[6ull * i + 4], rays[6ull * i + 5]);
    tmin = tmin0;
    tmax = tmax0;
  }
  FR_D void world_ray(float3& o, float3& d) const
  {
    o = f3(rays[6ull * i], rays[6ull * i + 1], rays[6ull * i + 2]);
    d = f3(rays[6ull * i + 3], rays[6ull * i + 4], rays[6ull * i + 5]);
  }
  FR_D void retire(bool has, const HitRecord& h, const TraceCounters& cnt)
  {
    if (!has) return;
    if (h.face != kNoHit) {
      const uint32_t sm = sc.face_submesh[h.face];
      out_id[2ull * i] = sm;
      out_id[2ull * i + 1] = h.face - submesh_offsets[sm];
      out_tuv[3ull * i] = h.t;
      out_tuv[3ull * i + 1] = h.u;
      out_tuv[3ull * i + 2] = h.v;
    } else {
      out_id[2ull * i] = out_id[2ull * i + 1] = kNoHit;
      out_tuv[3ull * i] = out_tuv[3ull * i + 1] = out_tuv[3ull * i + 2] = 0.0f;
    }
    if (counters) {
      atomicAdd(&counters[0], (unsigned long long)cnt.nodes);
      atomicAdd(&counters[1], (unsigned long long)cnt.tris);
    }
  }
};

template <bool TWO>
__global__ void __launch_bounds__(kBlock) k_trace_batch(SceneView sc, const uint32_t* __restrict__ submesh_offsets,
                                                        const float* __restrict__ rays, uint32_t n, float tmin,
                                                        float tmax, uint32_t* __restrict__ out_id,
                                                        float* __restrict__ out_tuv, unsigned long long* counters,
                                                        uint32_t* cursor, int refill, int tri_lanes)
{
  FR_DECLARE_STACK();
  BatchPolicy pol{sc, submesh_offsets, rays, tmin, tmax, out_id, out_tuv, counters, 0u};
  trace_queue<false, true, TWO>(sc.bvh, pol, cursor, n, stack_column, kBlock, refill, tri_lanes);
}

GridCache g_grid_closest, g_grid_shadow, g_grid_light;
GridCache g_grid_closest2, g_grid_shadow2, g_grid_light2;  // two-level instantiations
GridCache g_grid_batch, g_grid_batch2;

// idle lanes per warp that trigger a refill (tunable for experiments: FRD_REFILL_LANES)
int env_int(const char* name, int fallback)
{
  const char* e = getenv(name);
  const int v = e ? (int)strtol(e, nullptr, 0) : fallback;
  return v < 1 ? 1 : v;
}
int refill_lanes()
{
  static const int v = env_int("FRD_REFILL_LANES", 8);
  return v;
}
// the same threshold for launches whose queue is in beam order (camera rays, first-bounce sun rays)
int refill_lanes_coherent()
{
  static const int v = env_int("FRD_REFILL_LANES_COHERENT", refill_lanes());
  return v;
}
// lanes with a pending triangle that trigger a triangle phase (closest-hit / any-hit kernels)
int tri_lanes_closest()
{
  static const int v = env_int("FRD_TRI_LANES", 1);
  return v;
}
int tri_lanes_any()
{
  static const int v = env_int("FRD_TRI_LANES_ANY", 1);
  return v;
}


}  // namespace

bool g_count_traversal = false;
void set_traversal_counting(bool on) { g_count_traversal = on; }

void launch_trace_closest(cudaStream_t s, const SceneView& sc, const WaveBuffers& wb, uint32_t depth,
                          const uint32_t* order)
{
  const int refill = depth == 0 ? refill_lanes_coherent() : refill_lanes();
  if (sc.bvh.instances) {
    const int grid = g_grid_closest2.get(reinterpret_cast<const void*>(k_trace_closest<false, true>), kBlock);
    k_trace_closest<false, true><<<grid, kBlock, 0, s>>>(sc, wb, depth, refill, tri_lanes_closest(), order);
    FR_CUDA_LAUNCH_CHECK();
    return;
  }
  const int grid = g_grid_closest.get(reinterpret_cast<const void*>(k_trace_closest<false, false>), kBlock);
  if (g_count_traversal)
    k_trace_closest<true, false><<<grid, kBlock, 0, s>>>(sc, wb, depth, refill, tri_lanes_closest(), order);
  else
    k_trace_closest<false, false><<<grid, kBlock, 0, s>>>(sc, wb, depth, refill, tri_lanes_closest(), order);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_trace_shadow(cudaStream_t s, const SceneView& sc, const WaveBuffers& wb, int which, const uint32_t* order,
                         bool coherent)
{
  const int refill = coherent ? refill_lanes_coherent() : refill_lanes();
  if (sc.bvh.instances) {
    const int grid = g_grid_shadow2.get(reinterpret_cast<const void*>(k_trace_shadow<false, true>), kBlock);
    k_trace_shadow<false, true><<<grid, kBlock, 0, s>>>(sc, wb, which, refill, tri_lanes_any(), order);
    FR_CUDA_LAUNCH_CHECK();
    return;
  }
  const int grid = g_grid_shadow.get(reinterpret_cast<const void*>(k_trace_shadow<false, false>), kBlock);
  if (g_count_traversal)
    k_trace_shadow<true, false><<<grid, kBlock, 0, s>>>(sc, wb, which, refill, tri_lanes_any(), order);
  else
    k_trace_shadow<false, false><<<grid, kBlock, 0, s>>>(sc, wb, which, refill, tri_lanes_any(), order);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_trace_light(cudaStream_t s, const SceneView& sc, const WaveBuffers& wb, const uint32_t* order)
{
  if (sc.n_lights == 0) {
    launch_trace_shadow(s, sc, wb, 3, order, false);
    return;
  }
  if (sc.bvh.instances) {
    const int grid = g_grid_light2.get(reinterpret_cast<const void*>(k_trace_light<false, true>), kBlock);
    k_trace_light<false, true><<<grid, kBlock, 0, s>>>(sc, wb, refill_lanes(), tri_lanes_closest(), order);
    FR_CUDA_LAUNCH_CHECK();
    return;
  }
  const int grid = g_grid_light.get(reinterpret_cast<const void*>(k_trace_light<false, false>), kBlock);
  if (g_count_traversal)
    k_trace_light<true, false><<<grid, kBlock, 0, s>>>(sc, wb, refill_lanes(), tri_lanes_closest(), order);
  else
    k_trace_light<false, false><<<grid, kBlock, 0, s>>>(sc, wb, refill_lanes(), tri_lanes_closest(), order);
  FR_CUDA_LAUNCH_CHECK();
}

void trace_batch_closest(const SceneView& sc, const uint32_t* d_submesh_offsets, const float* rays_host, uint32_t n,
                         float tmin, float tmax, uint32_t* out_id_host, float* out_tuv_host,
                         unsigned long long* counters2_host)
{
  if (n == 0) return;
  DevBuf<float> d_rays(6ull * n), d_tuv(3ull * n);
  DevBuf<uint32_t> d_id(2ull * n);
  DevBuf<unsigned long long> d_cnt(2);
  DevBuf<uint32_t> d_cursor(1);
  d_cnt.zero();
  d_cursor.zero();
  FR_CUDA_CHECK(cudaMemcpy(d_rays.get(), rays_host, sizeof(float) * 6ull * n, cudaMemcpyHostToDevice));
  if (sc.bvh.instances) {
    const int grid = std::min<int>((n + kBlock - 1) / kBlock, g_grid_batch2.get(reinterpret_cast<const void*>(k_trace_batch<true>), kBlock));
    k_trace_batch<true><<<grid, kBlock>>>(sc, d_submesh_offsets, d_rays.get(), n, tmin, tmax, d_id.get(), d_tuv.get(),
                                          counters2_host ? d_cnt.get() : nullptr, d_cursor.get(), refill_lanes(),
                                          tri_lanes_closest());
  } else {
    const int grid = std::min<int>((n + kBlock - 1) / kBlock, g_grid_batch.get(reinterpret_cast<const void*>(k_trace_batch<false>), kBlock));
    k_trace_batch<false><<<grid, kBlock>>>(sc, d_submesh_offsets, d_rays.get(), n, tmin, tmax, d_id.get(), d_tuv.get(),
                                           counters2_host ? d_cnt.get() : nullptr, d_cursor.get(), refill_lanes(),
                                           tri_lanes_closest());
  }
  FR_CUDA_LAUNCH_CHECK();
  FR_CUDA_CHECK(cudaMemcpy(out_id_host, d_id.get(), sizeof(uint32_t) * 2ull * n, cudaMemcpyDeviceToHost));
  FR_CUDA_CHECK(cudaMemcpy(out_tuv_host, d_tuv.get(), sizeof(float) * 3ull * n, cudaMemcpyDeviceToHost));
  if (counters2_host)
    FR_CUDA_CHECK(cudaMemcpy(counters2_host, d_cnt.get(), sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost));
}

}  // namespace frd
