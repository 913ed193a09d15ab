// Wavefront stages that are not traversal: camera ray generation, the shade stage
// (surface + BSDF + next-event estimation + path continuation) and the film stage.
//
// Together with trace.cu these replace the reference's megakernel
//   __raygen__rg          pt.cu:418-502   -> k_generate, k_shade (tail), k_film
//   __miss__radiance      pt.cu:504-523   -> k_shade (miss branch)
//   __closesthit__radiance pt.cu:680-944  -> k_shade (hit branch)
// Sample semantics are those of the reference launched with n_samples = 1 per
// launch ("canonical mode"): every sample starts with a
// fresh payload, so `firsthit` is simply "bounce 0".
#include <cstring>
#include <stdexcept>

#include "tables.cuh"
//
#include "bsdf.cuh"
#include "cuda_util.h"
#include "queue.cuh"
#include "sampler.cuh"
#include "surface.cuh"
#include "wavefront.h"
#include "wavefront_kernels.h"

namespace frd
{
namespace
{

constexpr int kBlock = 128;
#ifndef FRD_SHADE_BLOCKS
#define FRD_SHADE_BLOCKS 4
#endif

FR_D float pack_draws(const PathSampler& s) { return __uint_as_float(s.cmj_draws | (s.sobol_dim << 16)); }

FR_D PathSampler restore_sampler(const WaveParams& wp, uint32_t sample, uint32_t x, uint32_t y, float packed)
{
  PathSampler s;
  const uint32_t n_pixels = wp.film.width * wp.film.height;
  s.init(x + wp.film.width * y, wp.sample_base + sample, n_pixels, wp.seed);
  const uint32_t bits = __float_as_uint(packed);
  s.cmj_draws = bits & 0xffffu;
  s.sobol_dim = bits >> 16;
  return s;
}

// ---- camera rays ----------------------------------------------------------------------
constexpr int kGenBlock = 256;

__global__ void __launch_bounds__(kGenBlock) k_generate(WaveParams wp, WaveBuffers wb)
{
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_slots = film_groups(wp.film, wp.n_samples) * wp.film.slots_per_group;
  bool alive = false;
  if (slot < n_slots) {
    uint32_t x, y, sample;
    const bool inside = slot_to_pixel(wp.film, slot, x, y, sample) && sample < wp.n_samples;
    wb.L[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wp.single_launch) wb.hit[slot] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kNoHit));  // read by k_first_hit
    if (wp.want_aov && !wp.single_launch) {  // the first-hit words are only kept when a layer will read them
      wb.aov0[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
      wb.aov1[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
      wb.aov2[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (inside && wp.max_depth > 0) {
      PathSampler s;
      s.init(x + wp.film.width * y, wp.sample_base + sample, wp.film.width * wp.film.height, wp.seed);
      // pixel jitter then lens sample (pt.cu:438-446); image x is flipped
      const float2 j = s.next2d();
      const float w = (float)wp.film.width, h = (float)wp.film.height;
      float2 uv = make_float2((2.0f * (x + j.x) - w) / h, (2.0f * (y + j.y) - h) / h);
      uv.x = -uv.x;
      const float2 lens = s.next2d();
      LensModel lm;
      lm.init(wp.camera);
      float3 o, d;
      thin_lens_ray(wp.camera, lm, uv, lens, o, d);
      // bounce-0 roulette: probability 1, but the draw is consumed and a draw of
      // exactly 1.0f still stops the path (pt.cu:457-462)
      const float u = s.next1d();
      alive = !(u >= 1.0f);
      wb.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.f);
      wb.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
      wb.thr[slot] = make_float4(1.f, 1.f, 1.f, pack_draws(s));
    }
  }
  // One reservation per 256-thread block: with one per warp the kernel ran at the rate a single address
  // takes atomics (33 M paths / 32 = 1 M atomics on ctl->n[Q_CUR] in 1.35 ms).
  __shared__ uint32_t s_count[kGenBlock / 32];
  __shared__ uint32_t s_base;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t mask = __ballot_sync(0xffffffffu, alive);
  if (lane == 0) s_count[warp] = __popc(mask);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int w = 0; w < kGenBlock / 32; ++w) total += s_count[w];
    s_base = total ? atomicAdd(&wb.ctl->n[Q_CUR], total) : 0u;
  }
  __syncthreads();
  if (alive) {
    uint32_t pos = s_base + __popc(mask & ((1u << lane) - 1u));
    for (uint32_t w = 0; w < warp; ++w) pos += s_count[w];
    wb.queue[0][pos] = slot;
  }
}

// ---- shade ----------------------------------------------------------------------------
template <bool TEX>
FR_D void load_surface_params(const fredholm::Material& m, const SceneTex& tex, const float2& uv,
                              SurfaceParams& p)
{
  // texture-or-constant resolution of the material inputs (pt.cu:181-280)
  p.diffuse = m.diffuse;
  p.diffuse_roughness = m.diffuse_roughness;
  p.base_color = m.base_color;
  p.specular = m.specular;
  p.specular_color = m.specular_color;
  float spec_rough = m.specular_roughness;
  p.metalness = m.metalness;
  float coat = m.coat, coat_rough = m.coat_roughness;
  if (TEX) {
    if (m.base_color_texture_id >= 0) p.base_color = f3(tex.fetch(m.base_color_texture_id, uv));
    if (m.specular_color_texture_id >= 0) p.specular_color = f3(tex.fetch(m.specular_color_texture_id, uv));
    if (m.specular_roughness_texture_id >= 0) spec_rough = tex.fetch(m.specular_roughness_texture_id, uv).x;
    if (m.metalness_texture_id >= 0) p.metalness = tex.fetch(m.metalness_texture_id, uv).x;
    if (m.coat_texture_id >= 0) coat = tex.fetch(m.coat_texture_id, uv).x;
    if (m.coat_roughness_texture_id >= 0) coat_rough = tex.fetch(m.coat_roughness_texture_id, uv).y;
  }
  p.specular_roughness = clampf(spec_rough, 0.01f, 1.0f);
  if (TEX && m.metallic_roughness_texture_id >= 0) {
    const float4 mr = tex.fetch(m.metallic_roughness_texture_id, uv);
    p.specular_roughness = clampf(mr.y, 0.01f, 1.0f);
    p.metalness = clampf(mr.z, 0.0f, 1.0f);
  }
  p.coat = clampf(coat, 0.0f, 1.0f);
  // quirk: the reference never copies Material::coat_color into its shading
  // parameters (pt.cu:238-255), so the coat is always colourless on the device
  p.coat_color = f3(1.0f);
  p.coat_roughness = clampf(coat_rough, 0.0f, 1.0f);
  p.transmission = m.transmission;
  p.transmission_color = m.transmission_color;
  p.sheen = m.sheen;
  p.sheen_color = m.sheen_color;
  p.sheen_roughness = m.sheen_roughness;
  p.subsurface = m.subsurface;
  p.subsurface_color = m.subsurface_color;
  p.thin_walled = m.thin_walled;
}

// firefly clamp of the reference (pt.cu:373-376): clamp(w, 0, 1) compiled by nvcc to .sat, i.e. NaN -> 0
// (a NaN weight -- every lobe weight 0 on the back face of an opaque surface -- contributes nothing)
FR_D float3 regularize(const float3& w) { return saturate3(w); }
FR_D bool nonzero3(const float3& v) { return v.x != 0.0f || v.y != 0.0f || v.z != 0.0f; }

constexpr float kShadowEps = 0.001f;  // SHADOW_RAY_EPS, pt.cu:11
constexpr float kRayMax = 1e9f;

FR_D void add_radiance(const WaveBuffers& wb, uint32_t slot, const float3& v)
{
  float4 L = wb.L[slot];
  L.x += v.x;
  L.y += v.y;
  L.z += v.z;
  wb.L[slot] = L;
}

// camera rays that left the scene: __miss__radiance with firsthit (pt.cu:504-523)
__global__ void __launch_bounds__(kBlock) k_miss(WaveParams wp, SceneView sc, WaveBuffers wb)
{
  WaveControl* ctl = wb.ctl;
  const uint32_t n = ctl->n_class[CLS_MISS];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = wb.class_queue[CLS_MISS][i];
    if (wp.single_launch) {
      // firsthit is already false once an earlier sample of the launch hit geometry (pt.cu:509)
      uint32_t px, py, sample;
      slot_to_pixel(wp.film, slot, px, py, sample);
      if (wp.sample_base + sample > wb.first_hit[px + wp.film.width * py]) continue;
    }
    const float3 d = f3(wb.ray_d[slot]);
    const float3 thr = f3(wb.thr[slot]);
    add_radiance(wb, slot, thr * sky_radiance(sc, d));
  }
}

// One material class at one bounce.  MASK = lobes the class can have, TEX = whether its
// materials read textures.  Every warp pulls 32 paths of the class queue; the next-event
// strategies and the two BSDF samples run as (non-unrolled) loops so that the BSDF code
// exists once per kernel, and queue appends happen with the whole warp converged.
template <uint32_t MASK, bool TEX>
__global__ void __launch_bounds__(kBlock, FRD_SHADE_BLOCKS) k_shade(WaveParams wp, SceneView sc, WaveBuffers wb, uint32_t depth, int cls)
{
  WaveControl* ctl = wb.ctl;
  const uint32_t n = ctl->n_class[cls];
  const uint32_t* q_in = wb.class_queue[cls];
  uint32_t* q_out = wb.queue[(depth & 1u) ^ 1u];
  const SceneTex tex{sc.textures, sc.srgb_lut};
  uint32_t item;
  uint32_t skipped = 0;  // visibility / MIS rays the reference would trace whose contribution is exactly zero
  while (fetch_batch(&ctl->cursor_class[cls], n, item)) {
    bool active = item < n;
    uint32_t slot = 0;
    float3 ray_d = f3(0.f), throughput = f3(0.f), x = f3(0.f), n_g = f3(0.f), shadow_o = f3(0.f);
    Frame fr;
    fr.t = fr.n = fr.b = f3(0.f);
    PathSampler smp;
    smp.pixel = smp.n_spp = smp.sobol_index = smp.seed_hash = smp.cmj_draws = smp.sobol_dim = 0;
    Closure<MASK> bsdf;

    if (active) {
      slot = q_in[item];
      const float4 rd = wb.ray_d[slot];
      const float4 hit = wb.hit[slot];
      const float4 thr4 = wb.thr[slot];
      ray_d = f3(rd);
      throughput = f3(thr4);
      const uint32_t face = __float_as_uint(hit.w);
      uint32_t px, py, sample;
      path_identity(wp.film, wb, slot, px, py, sample);
      smp = restore_sampler(wp, sample, px, py, thr4.w);

      // ---- surface (fill_surface_info, pt.cu:141-179) ----
      const uint3 idx = sc.indices[face];
      const FaceGeom g = load_face(sc, idx, sc.face_submesh[face]);
      const float bu = hit.y, bv = hit.z;
      x = bary3(g.v0, g.v1, g.v2, bu, bv);
      n_g = normalize(cross(g.v1 - g.v0, g.v2 - g.v0));
      float3 n_s = normalize(bary3(g.n0, g.n1, g.n2, bu, bv));
      const float2 uv = bary2(g.t0, g.t1, g.t2, bu, bv);
      const bool entering = dot(-ray_d, n_g) > 0.0f;
      if (!entering) {
        n_s = -n_s;
        n_g = -n_g;
      }
      fr.n = n_s;
      onb(n_s, fr.t, fr.b);

      const fredholm::Material& mat = sc.materials[sc.material_ids[face]];
      SurfaceParams sp;
      load_surface_params<TEX>(mat, tex, uv, sp);

      if (TEX) {
        // bump / normal mapping (pt.cu:709-742)
        if (mat.heightmap_texture_id >= 0) {
          const TexView& hm = sc.textures[mat.heightmap_texture_id];
          const float du = 1.0f / hm.width, dv = 1.0f / hm.height;
          const float v = tex.fetch(mat.heightmap_texture_id, uv).x;
          const float dfdu = tex.fetch(mat.heightmap_texture_id, make_float2(uv.x + du, uv.y)).x - v;
          const float dfdv = tex.fetch(mat.heightmap_texture_id, make_float2(uv.x, uv.y + dv)).x - v;
          const float3 t0 = fr.t, b0 = fr.b;
          fr.t = normalize(t0 + dfdu * n_s);
          fr.b = normalize(b0 + dfdv * n_s);
          fr.n = normalize(cross(fr.t, fr.b));
        }
        if (mat.normalmap_texture_id >= 0) {
          float3 value = f3(tex.fetch(mat.normalmap_texture_id, uv));
          value = 2.0f * value - 1.0f;
          // the reference maps through the UNBUMPED frame here (pt.cu:739-740)
          // (tangent-space map: x -> tangent, y -> bitangent, z -> normal)
          float3 t0, b0;
          onb(n_s, t0, b0);
          fr.n = normalize(value.x * t0 + value.y * b0 + value.z * n_s);
          onb(fr.n, fr.t, fr.b);
        }
      }

      // single-launch mode: only the first sample of the launch that hits geometry still has
      // payload.firsthit (pt.cu:744-759); later ones shade a directly visible emitter like any surface
      const uint32_t pixel = px + wp.film.width * py;
      const bool first_hit = depth == 0 && (!wp.single_launch || wb.first_hit[pixel] == smp.n_spp);
      if (first_hit && wp.single_launch) {
        wb.pix_aov0[pixel] = make_float4(x.x, x.y, x.z, hit.x);
        wb.pix_aov1[pixel] = make_float4(fr.n.x, fr.n.y, fr.n.z, uv.x);
        wb.pix_aov2[pixel] = make_float4(sp.base_color.x, sp.base_color.y, sp.base_color.z, uv.y);
      }
      if (first_hit) {
        // first-hit AOVs and directly visible emitters (pt.cu:745-760)
        if (wp.want_aov && !wp.single_launch) {
          wb.aov0[slot] = make_float4(x.x, x.y, x.z, hit.x);
          wb.aov1[slot] = make_float4(fr.n.x, fr.n.y, fr.n.z, uv.x);
          wb.aov2[slot] = make_float4(sp.base_color.x, sp.base_color.y, sp.base_color.z, uv.y);
        }
        if (is_emissive(mat)) {
          const float3 le = TEX ? emission_of(mat, tex, uv) : mat.emission_color;
          add_radiance(wb, slot, throughput * le);
          active = false;
        }
      }
      if (active) {
        bsdf.init(fr.to_local(-ray_d), sp, entering);
        shadow_o = offset_origin(x, n_g);
      }
    }

    // ---- next-event estimation (pt.cu:766-890): 0 = sun disk, 1 = sky, 2 = area light ----
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
      if (k == 0 && !sc.has_dir_light) continue;
      if (k == 2 && sc.n_lights == 0) continue;
      bool want = false;
      float3 dir = f3(0.f), c = f3(0.f);
      float tmax = kRayMax - kShadowEps;
      if (active) {
        float3 wi, le;
        float pdf;
        bool valid = true;
        if (k == 0) {
          const float2 pd = concentric_disk(smp.next2d());
          const float3 p = kRayMax * sc.dir_light.dir + sc.dir_disk_radius * (sc.dir_t * pd.x + sc.dir_b * pd.y);
          dir = normalize(p - shadow_o);
          wi = fr.to_local(dir);
          pdf = 1.0f;
          le = sc.dir_light.le;
        } else if (k == 1) {
          // cosine-hemisphere sample of the sky, always drawn (pt.cu:796-857)
          wi = cosine_hemisphere(smp.next2d());
          dir = fr.to_world(wi);
          pdf = abs_cos(wi) / kPi;
          le = sky_radiance(sc, dir);
        } else {
          // uniformly chosen emissive triangle, uniform point on it (pt.cu:282-322, 859-889)
          const float u1 = smp.next1d();
          const float2 u2 = smp.next2d();
          const uint32_t li = min((uint32_t)(u1 * sc.n_lights), sc.n_lights - 1u);
          const fredholm::AreaLight light = sc.lights[li];
          const float su = sqrtf(u2.x);
          const float b0 = 1.0f - su, b1 = u2.y * su;
          const FaceGeom lg = load_face(sc, light.indices, light.instance_idx);
          const float3 p = bary3(lg.v0, lg.v1, lg.v2, b0, b1);
          const float3 nl = bary3(lg.n0, lg.n1, lg.n2, b0, b1);
          const float2 luv = bary2(lg.t0, lg.t1, lg.t2, b0, b1);
          const float area = 0.5f * length(cross(lg.v1 - lg.v0, lg.v2 - lg.v0));
          const float pdf_area = 1.0f / (sc.n_lights * area);
          const float3 to_l = p - shadow_o;
          dir = normalize(to_l);
          const float r = length(to_l);
          const float cos_l = dot(-dir, nl);
          valid = cos_l > 0.0f;
          wi = fr.to_local(dir);
          pdf = r * r / fabsf(cos_l) * pdf_area;
          const fredholm::Material& lm = sc.materials[light.material_id];
          le = TEX ? emission_of(lm, tex, luv) : lm.emission_color;
          tmax = r - kShadowEps;
        }
        if (valid) {
          float3 f;
          float pdf_bsdf;
          bsdf.eval(wi, f, pdf_bsdf);
          const float mis = pdf / (pdf + pdf_bsdf);
          const float3 weight = regularize(throughput * mis * f * abs_cos(wi) / pdf);
          c = weight * le;
          want = nonzero3(c);  // a zero contribution needs no visibility test (NaN is kept)
          skipped += want ? 0u : 1u;
        }
      }
      const uint32_t pos = queue_reserve(&ctl->n[Q_SHADOW0 + k], want);
      if (want) {
        float4* dst = reinterpret_cast<float4*>(wb.shadow[k] + pos);
        dst[0] = make_float4(shadow_o.x, shadow_o.y, shadow_o.z, tmax);
        dst[1] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(slot));
        dst[2] = make_float4(c.x, c.y, c.z, 0.0f);
      }
    }

    // ---- two independent BSDF samples: 0 = MIS ray towards emitters / sky (pt.cu:892-925),
    //      1 = path continuation (pt.cu:927-943) ----
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
      bool want = false;
      float3 o = f3(0.f), dir = f3(0.f), w = f3(0.f);
      float pdf = 0.0f, cos_wi = 0.0f;
      if (active) {
        const float u1 = smp.next1d();
        const float2 u2 = smp.next2d();
        float3 f;
        const float3 wi = bsdf.sample(u1, u2, f, pdf);
        dir = fr.to_world(wi);
        const bool transmitted = dot(dir, n_g) < 0.0f;
        o = offset_origin(x, transmitted ? -n_g : n_g);
        cos_wi = abs_cos(wi);
        if (s == 0) {
          w = throughput * f * cos_wi / pdf;
          want = nonzero3(w);
          skipped += want ? 0u : 1u;
          if (sc.n_lights == 0) {
            // no emissive face anywhere: the MIS ray can only contribute by leaving the scene
            // (__miss__light, pt.cu:531-543), so its contribution is known here and the ray
            // becomes a plain visibility ray (cosine pdf of the sky NEE strategy)
            const float mis = pdf / (pdf + cos_wi / kPi);
            w = saturate3(w * mis) * sky_radiance(sc, dir);
          }
        } else {
          throughput *= f * cos_wi / pdf;
          // raygen loop tail + head of the next iteration (pt.cu:455-471)
          want = !bad3(throughput) && depth + 1 < wp.max_depth;
          if (want) {
            const float p = saturate1(luminance(throughput));  // pt.cu:460, .sat on the device
            const float u = smp.next1d();
            want = !(u >= p);
            throughput = throughput / p;
          }
        }
      }
      if (s == 0) {
        const uint32_t pos = queue_reserve(&ctl->n[Q_LIGHT], want);
        if (want) {
          float4* dst = reinterpret_cast<float4*>(wb.light + pos);
          dst[0] = make_float4(o.x, o.y, o.z, sc.n_lights == 0 ? kRayMax : pdf);
          dst[1] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(slot));
          dst[2] = make_float4(w.x, w.y, w.z, cos_wi);
        }
      } else {
        const uint32_t pos = queue_reserve(&ctl->n[Q_NEXT], want);
        if (want) {
          wb.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.f);
          wb.ray_d[slot] = make_float4(dir.x, dir.y, dir.z, 0.f);
          wb.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, pack_draws(smp));
          q_out[pos] = slot;
        }
      }
    }
  }
  // ray accounting only (the reference's trace-call count = traced + skipped): one atomic per warp and kernel
  skipped = __reduce_add_sync(0xffffffffu, skipped);
  if (lane_id() == 0 && skipped) atomicAdd(&ctl->rays_skipped, (unsigned long long)skipped);
}

// end of a bounce: rotate the radiance queue, clear the secondary queues and cursors
__global__ void k_advance(WaveControl* ctl)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ctl->rays_closest += ctl->n[Q_CUR];
    ctl->rays_shadow += (unsigned long long)ctl->n[Q_SHADOW0] + ctl->n[Q_SHADOW1] + ctl->n[Q_SHADOW2];
    ctl->rays_light += ctl->n[Q_LIGHT];
    ctl->n[Q_CUR] = ctl->n[Q_NEXT];
    ctl->n[Q_NEXT] = 0;
    ctl->n[Q_SHADOW0] = ctl->n[Q_SHADOW1] = ctl->n[Q_SHADOW2] = 0;
    ctl->n[Q_LIGHT] = 0;
    for (int i = 0; i < 8; ++i) ctl->cursor[i] = 0;
    for (int i = 0; i < CLS_COUNT; ++i) ctl->n_class[i] = ctl->cursor_class[i] = 0;
  }
}

// ---- wave compaction (integrator.cpp): the paths a wave still has alive move to the straggler set ---------
// Everything a path carries from one bounce to the next is its next ray, its throughput (with the sampler's
// draw counters in .w) and its radiance so far.
__global__ void __launch_bounds__(256) k_migrate(WaveBuffers src, WaveBuffers dst, uint32_t parity, uint32_t origin_base)
{
  const uint32_t n = src.ctl->n[Q_CUR];
  const uint32_t base = dst.ctl->n[Q_CUR];  // not written by this kernel (k_migrate_commit)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = src.queue[parity][i];
    const uint32_t j = base + i;
    dst.ray_o[j] = src.ray_o[slot];
    dst.ray_d[j] = src.ray_d[slot];
    dst.thr[j] = src.thr[slot];
    dst.L[j] = src.L[slot];
    dst.origin[j] = origin_base + slot;
    dst.queue[parity][j] = j;
  }
}

__global__ void k_migrate_commit(WaveControl* src, WaveControl* dst)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    dst->n[Q_CUR] += src->n[Q_CUR];
    src->n[Q_CUR] = 0;
  }
}

// the stragglers' finished radiance goes back to the slot of the wave they came from (L_all = the waves' radiance
// arrays, one after the other)
__global__ void __launch_bounds__(256) k_migrate_back(WaveBuffers late, uint32_t n, float4* __restrict__ L_all)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) L_all[late.origin[j]] = late.L[j];
}

__global__ void k_wave_begin(WaveControl* ctl, unsigned long long n_paths)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int i = 0; i < Q_COUNT; ++i) ctl->n[i] = 0;
    for (int i = 0; i < 8; ++i) ctl->cursor[i] = 0;
    for (int i = 0; i < CLS_COUNT; ++i) ctl->n_class[i] = ctl->cursor_class[i] = 0;
    ctl->paths += n_paths;
  }
}

// ---- single-launch mode: which sample of the launch consumed payload.firsthit ------------
__global__ void __launch_bounds__(256) k_first_hit(WaveParams wp, WaveBuffers wb)
{
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= wp.film.width || y >= wp.film.height) return;
  const uint32_t pixel = x + wp.film.width * y;
  if (wb.first_hit[pixel] != 0xffffffffu) return;  // an earlier wave of this launch
  for (uint32_t s = 0; s < wp.n_samples; ++s) {
    if (__float_as_uint(wb.hit[pixel_to_slot(wp.film, x, y, s)].w) != kNoHit) {
      wb.first_hit[pixel] = wp.sample_base + s;
      return;
    }
  }
}

// ---- film -----------------------------------------------------------------------------
// Streaming mean of the reference (pt.cu:480-501), applied sample by sample in sample
// order so that the result is independent of how many samples one wave carries.
__global__ void __launch_bounds__(256) k_film(WaveParams wp, WaveBuffers wb, fredholm::RenderLayer layers, int film_mode)
{
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= wp.film.width || y >= wp.film.height) return;
  const uint32_t pixel = x + wp.film.width * y;

  // (wave compaction applies the first-hit layers and the beauty layer in two passes: either may be unbound)
  float3 beauty = layers.beauty ? f3(layers.beauty[pixel]) : f3(0.f);
  float3 position = layers.position ? f3(layers.position[pixel]) : f3(0.f);
  float3 normal = layers.normal ? f3(layers.normal[pixel]) : f3(0.f);
  float depth = layers.depth ? layers.depth[pixel] : 0.f;
  float2 texcoord = make_float2(0.f, 0.f);
  if (layers.texcoord) {
    const float4 t = layers.texcoord[pixel];
    texcoord = make_float2(t.x, t.y);
  }
  float3 albedo = layers.albedo ? f3(layers.albedo[pixel]) : f3(0.f);

  uint32_t n_spp = wp.sample_base;
  for (uint32_t s = 0; s < wp.n_samples; ++s) {
    const uint32_t slot = pixel_to_slot(wp.film, x, y, s);
    float3 radiance = layers.beauty ? f3(wb.L[slot]) : f3(0.f);
    if (bad3(radiance)) radiance = f3(0.f);  // pt.cu:475-478
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
    if (wp.single_launch) {
      // the payload keeps the first hit's values for every later sample of the launch (pt.cu:482-487)
      if (n_spp >= wb.first_hit[pixel]) a0 = wb.pix_aov0[pixel], a1 = wb.pix_aov1[pixel], a2 = wb.pix_aov2[pixel];
    } else if (wp.want_aov) {
      a0 = wb.aov0[slot], a1 = wb.aov1[slot], a2 = wb.aov2[slot];
    }
    if (film_mode == FILM_MEAN) {
      const float nf = (float)n_spp;
      const float coef = 1.0f / (nf + 1.0f);
      beauty = coef * (nf * beauty + radiance);
      position = coef * (nf * position + f3(a0));
      normal = coef * (nf * normal + f3(a1));
      depth = coef * (nf * depth + a0.w);
      texcoord = make_float2(coef * (nf * texcoord.x + a1.w), coef * (nf * texcoord.y + a2.w));
      albedo = coef * (nf * albedo + f3(a2));
    } else {
      beauty += radiance;
      position += f3(a0);
      normal += f3(a1);
      depth += a0.w;
      texcoord = make_float2(texcoord.x + a1.w, texcoord.y + a2.w);
      albedo += f3(a2);
    }
    n_spp++;
  }
  if (layers.beauty) layers.beauty[pixel] = make_float4(beauty.x, beauty.y, beauty.z, 1.0f);
  if (layers.position) layers.position[pixel] = make_float4(position.x, position.y, position.z, 1.0f);
  if (layers.normal) layers.normal[pixel] = make_float4(normal.x, normal.y, normal.z, 1.0f);
  if (layers.depth) layers.depth[pixel] = depth;
  if (layers.texcoord) layers.texcoord[pixel] = make_float4(texcoord.x, texcoord.y, 0.0f, 1.0f);
  if (layers.albedo) layers.albedo[pixel] = make_float4(albedo.x, albedo.y, albedo.z, 1.0f);
}

// accumulated sums -> means (used after the multi-GPU reduce of FILM_SUM layers)
__global__ void k_scale_layers(fredholm::RenderLayer layers, uint32_t n_pixels, float scale)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pixels) return;
  auto sc4 = [&](float4* p) {
    if (!p) return;
    const float4 v = p[i];
    p[i] = make_float4(v.x * scale, v.y * scale, v.z * scale, 1.0f);
  };
  sc4(layers.beauty);
  sc4(layers.position);
  sc4(layers.normal);
  sc4(layers.albedo);
  if (layers.texcoord) {
    const float4 v = layers.texcoord[i];
    layers.texcoord[i] = make_float4(v.x * scale, v.y * scale, 0.0f, 1.0f);
  }
  if (layers.depth) layers.depth[i] *= scale;
}

// ---- unit-test kernels (sampler / BSDF / sky / camera known-answer vectors) -------------
__global__ void k_test_sampler(uint32_t width, uint32_t height, uint32_t seed, uint32_t image_idx,
                               uint32_t n_spp, const char* kinds, uint32_t n_kinds, float* out)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  PathSampler s;
  s.init(image_idx, n_spp, width * height, seed);
  for (uint32_t k = 0; k < n_kinds; ++k) {
    if (kinds[k] == '1') {
      *out++ = s.next1d();
    } else {
      const float2 v = s.next2d();
      *out++ = v.x;
      *out++ = v.y;
    }
  }
}

// params: n x 30 floats laid out like the reference ShadingParams (shared.h:173-199)
FR_D SurfaceParams unpack_params(const float* p)
{
  SurfaceParams s;
  s.diffuse = p[0];
  s.base_color = f3(p[1], p[2], p[3]);
  s.diffuse_roughness = p[4];
  s.specular = p[5];
  s.specular_color = f3(p[6], p[7], p[8]);
  s.specular_roughness = p[9];
  s.metalness = p[10];
  s.coat = p[11];
  s.coat_color = f3(p[12], p[13], p[14]);
  s.coat_roughness = p[15];
  s.transmission = p[16];
  s.transmission_color = f3(p[17], p[18], p[19]);
  s.sheen = p[20];
  s.sheen_color = f3(p[21], p[22], p[23]);
  s.sheen_roughness = p[24];
  s.subsurface = p[25];
  s.subsurface_color = f3(p[26], p[27], p[28]);
  s.thin_walled = p[29];
  return s;
}

// in: params[30], wo[3], entering, wi[3], u, v[2] per case (40 floats)
// out: eval f[3], pdf, sample wi[3], f[3], pdf (11 floats)
__global__ void k_test_bsdf(const float* in, uint32_t n, float* out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* c = in + 40ull * i;
  const SurfaceParams sp = unpack_params(c);
  const float3 wo = f3(c[30], c[31], c[32]);
  const bool entering = c[33] != 0.0f;
  const float3 wi = f3(c[34], c[35], c[36]);
  Closure<M_ALL> b;
  b.init(wo, sp, entering);
  float3 f;
  float pdf;
  b.eval(wi, f, pdf);
  float* o = out + 11ull * i;
  o[0] = f.x;
  o[1] = f.y;
  o[2] = f.z;
  o[3] = pdf;
  float3 fs;
  float pdfs;
  const float3 ws = b.sample(c[37], make_float2(c[38], c[39]), fs, pdfs);
  o[4] = ws.x;
  o[5] = ws.y;
  o[6] = ws.z;
  o[7] = fs.x;
  o[8] = fs.y;
  o[9] = fs.z;
  o[10] = pdfs;
}

__global__ void k_test_sky(SceneView sc, const float* dirs, uint32_t n, float* out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float3 r = sky_radiance(sc, f3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
  out[3 * i] = r.x;
  out[3 * i + 1] = r.y;
  out[3 * i + 2] = r.z;
}

// primary rays of sample `sample_base` for every pixel, row-major (o, d)
__global__ void k_test_primary_rays(WaveParams wp, float* out)
{
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= wp.film.width || y >= wp.film.height) return;
  PathSampler s;
  s.init(x + wp.film.width * y, wp.sample_base, wp.film.width * wp.film.height, wp.seed);
  const float2 j = s.next2d();
  const float w = (float)wp.film.width, h = (float)wp.film.height;
  float2 uv = make_float2((2.0f * (x + j.x) - w) / h, (2.0f * (y + j.y) - h) / h);
  uv.x = -uv.x;
  const float2 lens = s.next2d();
  LensModel lm;
  lm.init(wp.camera);
  float3 o, d;
  thin_lens_ray(wp.camera, lm, uv, lens, o, d);
  float* r = out + 6ull * (x + wp.film.width * y);
  r[0] = o.x;
  r[1] = o.y;
  r[2] = o.z;
  r[3] = d.x;
  r[4] = d.y;
  r[5] = d.z;
}


}  // namespace

// ---- launch wrappers (declared in wavefront_kernels.h) -------------------------------------
void launch_wave_begin(cudaStream_t s, const WaveBuffers& wb, unsigned long long n_paths)
{
  k_wave_begin<<<1, 32, 0, s>>>(wb.ctl, n_paths);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_migrate(cudaStream_t s, const WaveBuffers& src, const WaveBuffers& dst, uint32_t depth, uint32_t origin_base)
{
  static GridCache cache;
  const int grid = cache.get(reinterpret_cast<const void*>(k_migrate), 256);
  k_migrate<<<grid, 256, 0, s>>>(src, dst, depth & 1u, origin_base);
  FR_CUDA_LAUNCH_CHECK();
  k_migrate_commit<<<1, 32, 0, s>>>(src.ctl, dst.ctl);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_migrate_back(cudaStream_t s, const WaveBuffers& late, uint32_t n, float4* L_all)
{
  if (n == 0) return;
  k_migrate_back<<<(n + 255) / 256, 256, 0, s>>>(late, n, L_all);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_generate(cudaStream_t s, const WaveParams& wp, const WaveBuffers& wb)
{
  const uint32_t n_slots = film_groups(wp.film, wp.n_samples) * wp.film.slots_per_group;
  k_generate<<<(n_slots + kGenBlock - 1) / kGenBlock, kGenBlock, 0, s>>>(wp, wb);
  FR_CUDA_LAUNCH_CHECK();
}

template <uint32_t MASK, bool TEX>
void launch_shade_t(cudaStream_t s, const WaveParams& wp, const SceneView& sc, const WaveBuffers& wb, uint32_t depth,
                    int cls)
{
  static GridCache cache;
  const int grid = cache.get(reinterpret_cast<const void*>(k_shade<MASK, TEX>), kBlock);
  k_shade<MASK, TEX><<<grid, kBlock, 0, s>>>(wp, sc, wb, depth, cls);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_shade(cudaStream_t s, const WaveParams& wp, const SceneView& sc, const WaveBuffers& wb, uint32_t depth,
                  int cls)
{
  switch (cls) {
    case CLS_DIFFUSE: launch_shade_t<M_DIFFUSE_R, false>(s, wp, sc, wb, depth, cls); break;
    case CLS_PLASTIC: launch_shade_t<M_SPECULAR | M_DIFFUSE_R, false>(s, wp, sc, wb, depth, cls); break;
    case CLS_METAL: launch_shade_t<M_METAL | M_DIFFUSE_R, false>(s, wp, sc, wb, depth, cls); break;
    case CLS_COATED: launch_shade_t<M_COAT | M_SPECULAR | M_DIFFUSE_R, false>(s, wp, sc, wb, depth, cls); break;
    case CLS_GLASS: launch_shade_t<M_SPECULAR | M_TRANSMISSION | M_DIFFUSE_R, false>(s, wp, sc, wb, depth, cls); break;
    case CLS_SHEEN: launch_shade_t<M_SHEEN | M_DIFFUSE_R, false>(s, wp, sc, wb, depth, cls); break;
    case CLS_GENERIC: launch_shade_t<M_ALL, false>(s, wp, sc, wb, depth, cls); break;
    case CLS_GENERIC_TEX: launch_shade_t<M_ALL, true>(s, wp, sc, wb, depth, cls); break;
    default: throw std::runtime_error("launch_shade: bad class");
  }
}

void launch_first_hit(cudaStream_t s, const WaveParams& wp, const WaveBuffers& wb)
{
  const dim3 block(32, 8);
  const dim3 grid((wp.film.width + 31) / 32, (wp.film.height + 7) / 8);
  k_first_hit<<<grid, block, 0, s>>>(wp, wb);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_miss(cudaStream_t s, const WaveParams& wp, const SceneView& sc, const WaveBuffers& wb)
{
  static GridCache cache;
  const int grid = cache.get(reinterpret_cast<const void*>(k_miss), kBlock);
  k_miss<<<grid, kBlock, 0, s>>>(wp, sc, wb);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_advance(cudaStream_t s, const WaveBuffers& wb)
{
  k_advance<<<1, 32, 0, s>>>(wb.ctl);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_film(cudaStream_t s, const WaveParams& wp, const WaveBuffers& wb, const fredholm::RenderLayer& layers,
                 int film_mode)
{
  const dim3 block(32, 8);
  const dim3 grid((wp.film.width + 31) / 32, (wp.film.height + 7) / 8);
  k_film<<<grid, block, 0, s>>>(wp, wb, layers, film_mode);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_scale_layers(cudaStream_t s, const fredholm::RenderLayer& layers, uint32_t n_pixels, float scale)
{
  k_scale_layers<<<(n_pixels + 255) / 256, 256, 0, s>>>(layers, n_pixels, scale);
  FR_CUDA_LAUNCH_CHECK();
}

void test_sampler(uint32_t width, uint32_t height, uint32_t seed, uint32_t image_idx, uint32_t n_spp,
                  const char* kinds, float* out_host, uint32_t n_out)
{
  const uint32_t n_kinds = (uint32_t)strlen(kinds);
  DevBuf<char> d_kinds(n_kinds + 1);
  DevBuf<float> d_out(n_out);
  FR_CUDA_CHECK(cudaMemcpy(d_kinds.get(), kinds, n_kinds + 1, cudaMemcpyHostToDevice));
  k_test_sampler<<<1, 32>>>(width, height, seed, image_idx, n_spp, d_kinds.get(), n_kinds, d_out.get());
  FR_CUDA_LAUNCH_CHECK();
  FR_CUDA_CHECK(cudaMemcpy(out_host, d_out.get(), sizeof(float) * n_out, cudaMemcpyDeviceToHost));
}

void test_bsdf(const float* in_host, uint32_t n, float* out_host)
{
  DevBuf<float> d_in(40ull * n), d_out(11ull * n);
  FR_CUDA_CHECK(cudaMemcpy(d_in.get(), in_host, sizeof(float) * 40ull * n, cudaMemcpyHostToDevice));
  k_test_bsdf<<<(n + 63) / 64, 64>>>(d_in.get(), n, d_out.get());
  FR_CUDA_LAUNCH_CHECK();
  FR_CUDA_CHECK(cudaMemcpy(out_host, d_out.get(), sizeof(float) * 11ull * n, cudaMemcpyDeviceToHost));
}

void test_sky(const SceneView& sc, const float* dirs_host, uint32_t n, float* out_host)
{
  DevBuf<float> d_in(3ull * n), d_out(3ull * n);
  FR_CUDA_CHECK(cudaMemcpy(d_in.get(), dirs_host, sizeof(float) * 3ull * n, cudaMemcpyHostToDevice));
  k_test_sky<<<(n + 63) / 64, 64>>>(sc, d_in.get(), n, d_out.get());
  FR_CUDA_LAUNCH_CHECK();
  FR_CUDA_CHECK(cudaMemcpy(out_host, d_out.get(), sizeof(float) * 3ull * n, cudaMemcpyDeviceToHost));
}

void test_primary_rays(const WaveParams& wp, float* out_host)
{
  const size_t n = (size_t)wp.film.width * wp.film.height;
  DevBuf<float> d_out(6 * n);
  const dim3 block(32, 8);
  const dim3 grid((wp.film.width + 31) / 32, (wp.film.height + 7) / 8);
  k_test_primary_rays<<<grid, block>>>(wp, d_out.get());
  FR_CUDA_LAUNCH_CHECK();
  FR_CUDA_CHECK(cudaMemcpy(out_host, d_out.get(), sizeof(float) * 6 * n, cudaMemcpyDeviceToHost));
}

}  // namespace frd
