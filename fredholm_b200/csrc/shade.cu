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

FR_D float pack_draws(const PathSampler& s) { return __uint_as_float(s.cmj_draws | (s.sobol_dim << 16)); }

FR_D PathSampler restore_sampler(const WaveParams& wp, uint32_t slot, uint32_t x, uint32_t y, float packed)
{
  PathSampler s;
  const uint32_t n_pixels = wp.film.width * wp.film.height;
  s.init(x + wp.film.width * y, wp.sample_base + slot / wp.film.slots_per_sample, n_pixels, wp.seed);
  const uint32_t bits = __float_as_uint(packed);
  s.cmj_draws = bits & 0xffffu;
  s.sobol_dim = bits >> 16;
  return s;
}

// ---- camera rays ----------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_generate(WaveParams wp, WaveBuffers wb)
{
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_slots = wp.n_samples * wp.film.slots_per_sample;
  bool alive = false;
  if (slot < n_slots) {
    uint32_t x, y;
    const bool inside = slot_to_pixel(wp.film, slot % wp.film.slots_per_sample, x, y);
    wb.L[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    wb.aov0[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    wb.aov1[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    wb.aov2[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside && wp.max_depth > 0) {
      PathSampler s;
      s.init(x + wp.film.width * y, wp.sample_base + slot / wp.film.slots_per_sample,
             wp.film.width * wp.film.height, wp.seed);
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
  const uint32_t pos = queue_reserve(&wb.ctl->n[Q_CUR], alive);
  if (alive) wb.queue[0][pos] = slot;
}

// ---- shade ----------------------------------------------------------------------------
struct ShadeOut {
  bool has_shadow[3];
  ShadowRay shadow[3];
  bool has_light;
  LightRay light;
  bool continues;
};

FR_D void load_surface_params(const fredholm::Material& m, const SceneTex& tex, const float2& uv,
                              SurfaceParams& p)
{
  // texture-or-constant resolution of the material inputs (pt.cu:181-280)
  p.diffuse = m.diffuse;
  p.diffuse_roughness = m.diffuse_roughness;
  p.base_color = m.base_color_texture_id >= 0 ? f3(tex.fetch(m.base_color_texture_id, uv)) : m.base_color;
  p.specular = m.specular;
  p.specular_color =
      m.specular_color_texture_id >= 0 ? f3(tex.fetch(m.specular_color_texture_id, uv)) : m.specular_color;
  p.specular_roughness = clampf(
      m.specular_roughness_texture_id >= 0 ? tex.fetch(m.specular_roughness_texture_id, uv).x : m.specular_roughness,
      0.01f, 1.0f);
  p.metalness = m.metalness_texture_id >= 0 ? tex.fetch(m.metalness_texture_id, uv).x : m.metalness;
  if (m.metallic_roughness_texture_id >= 0) {
    const float4 mr = tex.fetch(m.metallic_roughness_texture_id, uv);
    p.specular_roughness = clampf(mr.y, 0.01f, 1.0f);
    p.metalness = clampf(mr.z, 0.0f, 1.0f);
  }
  p.coat = clampf(m.coat_texture_id >= 0 ? tex.fetch(m.coat_texture_id, uv).x : m.coat, 0.0f, 1.0f);
  // quirk: the reference never copies Material::coat_color into its shading
  // parameters (pt.cu:238-255), so the coat is always colourless on the device
  p.coat_color = f3(1.0f);
  p.coat_roughness =
      clampf(m.coat_roughness_texture_id >= 0 ? tex.fetch(m.coat_roughness_texture_id, uv).y : m.coat_roughness,
             0.0f, 1.0f);
  p.transmission = m.transmission;
  p.transmission_color = m.transmission_color;
  p.sheen = m.sheen;
  p.sheen_color = m.sheen_color;
  p.sheen_roughness = m.sheen_roughness;
  p.subsurface = m.subsurface;
  p.subsurface_color = m.subsurface_color;
  p.thin_walled = m.thin_walled;
}

// firefly clamp of the reference (pt.cu:373-376).  fmaxf(a, fminf(x, b)) maps NaN to b.
FR_D float3 regularize(const float3& w) { return clamp3(w, 0.0f, 1.0f); }
FR_D bool nonzero3(const float3& v) { return v.x != 0.0f || v.y != 0.0f || v.z != 0.0f; }

FR_D void set_shadow(ShadowRay& r, const float3& o, const float3& d, float tmax, uint32_t path, const float3& c)
{
  r.ox = o.x;
  r.oy = o.y;
  r.oz = o.z;
  r.tmax = tmax;
  r.dx = d.x;
  r.dy = d.y;
  r.dz = d.z;
  r.path = path;
  r.cr = c.x;
  r.cg = c.y;
  r.cb = c.z;
  r.pad_ = 0;
}

constexpr float kShadowEps = 0.001f;  // SHADOW_RAY_EPS, pt.cu:11
constexpr float kRayMax = 1e9f;

// One path at one bounce.  Returns what has to be enqueued.
FR_D void shade_path(const WaveParams& wp, const SceneView& sc, const WaveBuffers& wb, uint32_t slot,
                     uint32_t depth, ShadeOut& out)
{
  out.has_shadow[0] = out.has_shadow[1] = out.has_shadow[2] = false;
  out.has_light = false;
  out.continues = false;

  const float4 ro = wb.ray_o[slot], rd = wb.ray_d[slot];
  const float4 hit = wb.hit[slot];
  const float4 thr4 = wb.thr[slot];
  const float3 ray_o = f3(ro), ray_d = f3(rd);
  float3 throughput = f3(thr4);
  const uint32_t face = __float_as_uint(hit.w);

  if (face == kNoHit) {
    // __miss__radiance: sky is only added for camera rays; later bounces receive it
    // through next-event estimation and the MIS ray
    if (depth == 0) {
      const float3 le = sky_radiance(sc, ray_d);
      float4 L = wb.L[slot];
      L.x += throughput.x * le.x;
      L.y += throughput.y * le.y;
      L.z += throughput.z * le.z;
      wb.L[slot] = L;
    }
    return;
  }

  uint32_t px, py;
  slot_to_pixel(wp.film, slot % wp.film.slots_per_sample, px, py);
  PathSampler smp = restore_sampler(wp, slot, px, py, thr4.w);

  // ---- surface (fill_surface_info, pt.cu:141-179) ----
  const uint3 idx = sc.indices[face];
  const uint32_t xform = sc.face_submesh[face];
  const FaceGeom g = load_face(sc, idx, xform);
  const float bu = hit.y, bv = hit.z;
  const float3 x = bary3(g.v0, g.v1, g.v2, bu, bv);
  float3 n_g = normalize(cross(g.v1 - g.v0, g.v2 - g.v0));
  float3 n_s = normalize(bary3(g.n0, g.n1, g.n2, bu, bv));
  const float2 uv = bary2(g.t0, g.t1, g.t2, bu, bv);
  const bool entering = dot(-ray_d, n_g) > 0.0f;
  if (!entering) {
    n_s = -n_s;
    n_g = -n_g;
  }
  Frame fr;
  fr.n = n_s;
  onb(n_s, fr.t, fr.b);

  const fredholm::Material& mat = sc.materials[sc.material_ids[face]];
  const SceneTex tex{sc.textures, sc.srgb_lut};
  SurfaceParams sp;
  load_surface_params(mat, tex, uv, sp);

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

  if (depth == 0) {
    // first-hit AOVs and directly visible emitters (pt.cu:745-760)
    wb.aov0[slot] = make_float4(x.x, x.y, x.z, hit.x);
    wb.aov1[slot] = make_float4(fr.n.x, fr.n.y, fr.n.z, uv.x);
    wb.aov2[slot] = make_float4(sp.base_color.x, sp.base_color.y, sp.base_color.z, uv.y);
    if (is_emissive(mat)) {
      const float3 le = emission_of(mat, tex, uv);
      float4 L = wb.L[slot];
      L.x += throughput.x * le.x;
      L.y += throughput.y * le.y;
      L.z += throughput.z * le.z;
      wb.L[slot] = L;
      return;
    }
  }

  const float3 wo = fr.to_local(-ray_d);
  Closure bsdf;
  bsdf.init(wo, sp, entering);

  const float3 shadow_o = offset_origin(x, n_g);

  // ---- next-event estimation (pt.cu:766-890) ----
  if (sc.has_dir_light) {
    const float2 u = smp.next2d();
    const float2 pd = concentric_disk(u);
    const float3 p = kRayMax * sc.dir_light.dir + sc.dir_disk_radius * (sc.dir_t * pd.x + sc.dir_b * pd.y);
    const float3 dir = normalize(p - shadow_o);
    const float3 wi = fr.to_local(dir);
    float3 f;
    float pdf_bsdf;
    bsdf.eval(wi, f, pdf_bsdf);
    const float mis = 1.0f / (1.0f + pdf_bsdf);
    const float3 weight = regularize(throughput * mis * f * abs_cos(wi) / 1.0f);
    const float3 c = weight * sc.dir_light.le;
    if (nonzero3(c)) {
      out.has_shadow[0] = true;
      set_shadow(out.shadow[0], shadow_o, dir, kRayMax - kShadowEps, slot, c);
    }
  }
  {
    // sky: cosine-hemisphere sample, always drawn (pt.cu:796-857)
    const float2 u = smp.next2d();
    const float3 wi = cosine_hemisphere(u);
    const float3 dir = fr.to_world(wi);
    float3 f;
    float pdf_bsdf;
    bsdf.eval(wi, f, pdf_bsdf);
    const float pdf = abs_cos(wi) / kPi;
    const float mis = pdf / (pdf + pdf_bsdf);
    const float3 weight = regularize(throughput * mis * f * abs_cos(wi) / pdf);
    const float3 c = weight * sky_radiance(sc, dir);
    if (nonzero3(c)) {
      out.has_shadow[1] = true;
      set_shadow(out.shadow[1], shadow_o, dir, kRayMax - kShadowEps, slot, c);
    }
  }
  if (sc.n_lights > 0) {
    // uniformly chosen emissive triangle, uniform point on it (pt.cu:282-322, 859-889)
    const float u1 = smp.next1d();
    const float2 u2 = smp.next2d();
    const uint32_t li = min((uint32_t)(u1 * sc.n_lights), sc.n_lights - 1u);
    const fredholm::AreaLight light = sc.lights[li];
    const float su = sqrtf(u2.x);
    const float b0 = 1.0f - su, b1 = u2.y * su;
    const FaceGeom lg = load_face(sc, light.indices, light.instance_idx);
    const float3 p = bary3(lg.v0, lg.v1, lg.v2, b0, b1);
    const float3 n = bary3(lg.n0, lg.n1, lg.n2, b0, b1);
    const float2 luv = bary2(lg.t0, lg.t1, lg.t2, b0, b1);
    const float area = 0.5f * length(cross(lg.v1 - lg.v0, lg.v2 - lg.v0));
    const float pdf_area = 1.0f / (sc.n_lights * area);
    const float3 to_l = p - shadow_o;
    const float3 dir = normalize(to_l);
    const float r = length(to_l);
    const float cos_l = dot(-dir, n);
    if (cos_l > 0.0f) {
      const float3 le = emission_of(sc.materials[light.material_id], tex, luv);
      const float3 wi = fr.to_local(dir);
      float3 f;
      float pdf_bsdf;
      bsdf.eval(wi, f, pdf_bsdf);
      const float pdf = r * r / fabsf(cos_l) * pdf_area;
      const float mis = pdf / (pdf + pdf_bsdf);
      const float3 weight = regularize(throughput * mis * f * abs_cos(wi) / pdf);
      const float3 c = weight * le;
      if (nonzero3(c)) {
        out.has_shadow[2] = true;
        set_shadow(out.shadow[2], shadow_o, dir, r - kShadowEps, slot, c);
      }
    }
  }

  // ---- MIS ray: BSDF sample traced towards emitters / sky (pt.cu:892-925) ----
  {
    const float u1 = smp.next1d();
    const float2 u2 = smp.next2d();
    float3 f;
    float pdf;
    const float3 wi = bsdf.sample(u1, u2, f, pdf);
    const float3 dir = fr.to_world(wi);
    const bool transmitted = dot(dir, n_g) < 0.0f;
    const float3 o = offset_origin(x, transmitted ? -n_g : n_g);
    const float3 w = throughput * f * abs_cos(wi) / pdf;
    if (nonzero3(w)) {
      out.has_light = true;
      LightRay& r = out.light;
      r.ox = o.x;
      r.oy = o.y;
      r.oz = o.z;
      r.pdf_bsdf = pdf;
      r.dx = dir.x;
      r.dy = dir.y;
      r.dz = dir.z;
      r.path = slot;
      r.wr = w.x;
      r.wg = w.y;
      r.wb = w.z;
      r.cos_wi = abs_cos(wi);
    }
  }

  // ---- continuation: an independent second BSDF sample (pt.cu:927-943) ----
  {
    const float u1 = smp.next1d();
    const float2 u2 = smp.next2d();
    float3 f;
    float pdf;
    const float3 wi = bsdf.sample(u1, u2, f, pdf);
    const float3 dir = fr.to_world(wi);
    throughput *= f * abs_cos(wi) / pdf;
    const bool transmitted = dot(dir, n_g) < 0.0f;
    const float3 o = offset_origin(x, transmitted ? -n_g : n_g);

    // raygen loop tail + head of the next iteration (pt.cu:455-471)
    if (bad3(throughput)) return;
    if (depth + 1 >= wp.max_depth) return;
    const float p = clampf(luminance(throughput), 0.0f, 1.0f);
    const float u = smp.next1d();
    if (u >= p) return;
    throughput = throughput / p;
    wb.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.f);
    wb.ray_d[slot] = make_float4(dir.x, dir.y, dir.z, 0.f);
    wb.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, pack_draws(smp));
    out.continues = true;
  }
}

__global__ void __launch_bounds__(kBlock) k_shade(WaveParams wp, SceneView sc, WaveBuffers wb, uint32_t depth)
{
  WaveControl* ctl = wb.ctl;
  const uint32_t n = ctl->n[Q_CUR];
  const uint32_t* q_in = wb.queue[depth & 1u];
  uint32_t* q_out = wb.queue[(depth & 1u) ^ 1u];
  uint32_t item;
  while (fetch_batch(&ctl->cursor[1], n, item)) {
    ShadeOut out;
    out.has_shadow[0] = out.has_shadow[1] = out.has_shadow[2] = false;
    out.has_light = false;
    out.continues = false;
    uint32_t slot = 0;
    if (item < n) {
      slot = q_in[item];
      shade_path(wp, sc, wb, slot, depth, out);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint32_t pos = queue_reserve(&ctl->n[Q_SHADOW0 + k], out.has_shadow[k]);
      if (out.has_shadow[k]) {
        float4* dst = reinterpret_cast<float4*>(wb.shadow[k] + pos);
        const float4* src = reinterpret_cast<const float4*>(&out.shadow[k]);
        dst[0] = src[0];
        dst[1] = src[1];
        dst[2] = src[2];
      }
    }
    {
      const uint32_t pos = queue_reserve(&ctl->n[Q_LIGHT], out.has_light);
      if (out.has_light) {
        float4* dst = reinterpret_cast<float4*>(wb.light + pos);
        const float4* src = reinterpret_cast<const float4*>(&out.light);
        dst[0] = src[0];
        dst[1] = src[1];
        dst[2] = src[2];
      }
    }
    {
      const uint32_t pos = queue_reserve(&ctl->n[Q_NEXT], out.continues);
      if (out.continues) q_out[pos] = slot;
    }
  }
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
  }
}

__global__ void k_wave_begin(WaveControl* ctl, unsigned long long n_paths)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int i = 0; i < Q_COUNT; ++i) ctl->n[i] = 0;
    for (int i = 0; i < 8; ++i) ctl->cursor[i] = 0;
    ctl->paths += n_paths;
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
  const uint32_t slot0 = pixel_to_slot(wp.film, x, y);

  float3 beauty = f3(layers.beauty[pixel]);
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
    const uint32_t slot = s * wp.film.slots_per_sample + slot0;
    const float4 L = wb.L[slot];
    float3 radiance = f3(L);
    if (bad3(radiance)) radiance = f3(0.f);  // pt.cu:475-478
    const float4 a0 = wb.aov0[slot], a1 = wb.aov1[slot], a2 = wb.aov2[slot];
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
  layers.beauty[pixel] = make_float4(beauty.x, beauty.y, beauty.z, 1.0f);
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
  Closure b;
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

int g_shade_grid = 0;

int persistent_grid(const void* kernel, int block)
{
  int dev = 0, sms = 0, per_sm = 0;
  FR_CUDA_CHECK(cudaGetDevice(&dev));
  FR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  FR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0));
  return sms * (per_sm > 0 ? per_sm : 1);
}

}  // namespace

// ---- launch wrappers (declared in wavefront_kernels.h) -------------------------------------
void launch_wave_begin(cudaStream_t s, const WaveBuffers& wb, unsigned long long n_paths)
{
  k_wave_begin<<<1, 32, 0, s>>>(wb.ctl, n_paths);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_generate(cudaStream_t s, const WaveParams& wp, const WaveBuffers& wb)
{
  const uint32_t n_slots = wp.n_samples * wp.film.slots_per_sample;
  k_generate<<<(n_slots + kBlock - 1) / kBlock, kBlock, 0, s>>>(wp, wb);
  FR_CUDA_LAUNCH_CHECK();
}

void launch_shade(cudaStream_t s, const WaveParams& wp, const SceneView& sc, const WaveBuffers& wb, uint32_t depth)
{
  if (g_shade_grid == 0) g_shade_grid = persistent_grid(reinterpret_cast<const void*>(k_shade), kBlock);
  k_shade<<<g_shade_grid, kBlock, 0, s>>>(wp, sc, wb, depth);
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
