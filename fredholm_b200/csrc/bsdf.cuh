// Layered Standard-Surface BSDF of the shade stage.
//
// Behavioural spec: the reference's BSDF class and lobes
//   bsdf.cu:8-379 (layer weights, lobe CDF, eval / sample / eval_pdf)
//   bxdf.cu:119-822 (Oren-Nayar, diffuse transmission, GGX reflection with
//                    dielectric / conductor Fresnel, Walter GGX transmission,
//                    Estevez-Kulla sheen), sampling.cu:87-151 (VNDF, lobe CDF),
//   lut.cu:957-1081 (directional-albedo fetch)
// including the quirks that change pixels: coat absorption is
// formed with a zero coat albedo (bsdf.cu:26-39), sheen is sampled through a
// cosine-distributed half vector but reports a cosine pdf (bxdf.cu:759-778),
// sample() returns only the chosen lobe's value (bsdf.cu:214-293).
//
// It is NOT the reference's object layout: instead of seven lobe objects and a
// 17-float CDF (408 bytes, 1.1 KB stack frame on sm_100a) the closure keeps one
// flat set of scalars, shares the half vector / D / G terms between the lobes
// that use the same roughness, and evaluates value and pdf in one pass.
// Local shading frame: +y is the normal.
#pragma once
#include <cstdint>

#include "lobes.h"
#include "vecmath.cuh"

namespace frd
{

// c_lut_reflection / c_lut_sheen come from tables.cuh (include it first).

// Surface inputs after texture lookup (reference ShadingParams, shared.h:173-199).
struct SurfaceParams {
  float diffuse;
  float3 base_color;
  float diffuse_roughness;
  float specular;
  float3 specular_color;
  float specular_roughness;
  float metalness;
  float coat;
  float3 coat_color;
  float coat_roughness;
  float transmission;
  float3 transmission_color;
  float sheen;
  float3 sheen_color;
  float sheen_roughness;
  float subsurface;
  float3 subsurface_color;
  float thin_walled;
};

// ---- small lobe helpers ------------------------------------------------------
FR_D float abs_cos(const float3& w) { return fabsf(w.y); }

FR_D int lut_cell(float x)
{
  const int i = static_cast<int>(x * 16.0f);
  return min(max(i, 0), 15);
}

// bilinear fetch with clamped +1 neighbours (lut.cu:957-987)
FR_D float albedo_ggx(float cos_o, float roughness, float F0)
{
  const float u = fabsf(cos_o), v = clampf(roughness, 0.0f, 1.0f);
  const int i = lut_cell(u), j = lut_cell(v);
  const int i1 = min(i + 1, 15), j1 = min(j + 1, 15);
  const float hx = u * 16.0f - i, hy = v * 16.0f - j;
  const float2* lut = reinterpret_cast<const float2*>(c_lut_reflection);
  const float2 t0 = lut[i + 16 * j], t1 = lut[i1 + 16 * j];
  const float2 t2 = lut[i + 16 * j1], t3 = lut[i1 + 16 * j1];
  const float rx0 = (1.0f - hx) * t0.x + hx * t1.x, ry0 = (1.0f - hx) * t0.y + hx * t1.y;
  const float rx1 = (1.0f - hx) * t2.x + hx * t3.x, ry1 = (1.0f - hx) * t2.y + hx * t3.y;
  const float r = (1.0f - hy) * rx0 + hy * rx1;
  const float g = (1.0f - hy) * ry0 + hy * ry1;
  return F0 * r + (1.0f - F0) * g;
}

FR_D float albedo_sheen(float cos_o, float roughness)
{
  const float u = fabsf(cos_o), v = clampf(roughness, 0.0f, 1.0f);
  const int i = lut_cell(u), j = lut_cell(v);
  const int i1 = min(i + 1, 15), j1 = min(j + 1, 15);
  const float hx = u * 16.0f - i, hy = v * 16.0f - j;
  const float t0 = c_lut_sheen[i + 16 * j], t1 = c_lut_sheen[i1 + 16 * j];
  const float t2 = c_lut_sheen[i + 16 * j1], t3 = c_lut_sheen[i1 + 16 * j1];
  return (1.0f - hy) * ((1.0f - hx) * t0 + hx * t1) + hy * ((1.0f - hx) * t2 + hx * t3);
}

// exact dielectric Fresnel for relative ior (bxdf.cu:274-283)
FR_D float fresnel_dielectric(float c, float ior)
{
  const float g2 = ior * ior + c * c - 1.0f;
  if (g2 < 0.0f) return 1.0f;
  const float g = sqrtf(g2);
  const float t0 = (g - c) / (g + c);
  const float t1 = ((g + c) * c - 1.0f) / ((g - c) * c + 1.0f);
  return 0.5f * t0 * t0 * (1.0f + t1 * t1);
}

// conductor Fresnel (bxdf.cu:286-299)
FR_D float3 fresnel_conductor(float c, const float3& n, const float3& k)
{
  const float c2 = c * c;
  const float3 two_nc = 2.0f * n * c;
  const float3 t0 = n * n + k * k;
  const float3 t1 = t0 * c2;
  const float3 rs = (t0 - two_nc + c2) / (t0 + two_nc + c2);
  const float3 rp = (t1 - two_nc + 1.0f) / (t1 + two_nc + 1.0f);
  return 0.5f * (rp + rs);
}

// isotropic GGX terms in the y-up frame, alpha = roughness^2 (bxdf.cu:484-512)
FR_D float ggx_D(float a, const float3& h)
{
  const float a2 = a * a;
  const float t = h.x * h.x / a2 + h.z * h.z / a2 + h.y * h.y;
  return 1.0f / (kPi * a * a * t * t);
}
FR_D float ggx_lambda(float a, const float3& w)
{
  const float a2 = a * a;
  const float t = (a2 * w.x * w.x + a2 * w.z * w.z) / (w.y * w.y);
  return 0.5f * (-1.0f + sqrtf(1.0f + t));
}

// Heitz 2018 sampling of visible normals (sampling.cu:87-110), alpha_x = alpha_y
FR_D float3 sample_vndf(const float3& wo, float a, const float2& u)
{
  const float3 vh = normalize(f3(a * wo.x, wo.y, a * wo.z));
  const float lensq = vh.x * vh.x + vh.z * vh.z;
  const float3 t1 = lensq > 0.0f ? f3(vh.z, 0.0f, -vh.x) / sqrtf(lensq) : f3(0.0f, 0.0f, 1.0f);
  const float3 t2 = cross(vh, t1);
  const float r = sqrtf(u.x);
  const float phi = 2.0f * kPi * u.y;
  float s1, c1;
  sincosf(phi, &s1, &c1);
  const float p1 = r * c1;
  float p2 = r * s1;
  const float s = 0.5f * (1.0f + vh.y);
  p2 = (1.0f - s) * sqrtf(fmaxf(1.0f - p1 * p1, 0.0f)) + s * p2;
  const float3 nh = p1 * t1 + p2 * t2 + sqrtf(fmaxf(1.0f - p1 * p1 - p2 * p2, 0.0f)) * vh;
  return normalize(f3(a * nh.x, fmaxf(0.0f, nh.y), a * nh.z));
}

FR_D float3 mirror(const float3& w, const float3& n) { return normalize(-w + 2.0f * dot(w, n) * n); }

// Estevez-Kulla sheen fit (bxdf.cu:781-819)
struct SheenFit {
  float a, b, c, d, e, inv_r;
  FR_D void init(float roughness)
  {
    const float t = 1.0f - roughness;
    const float t2 = t * t;
    a = t2 * 25.3245f + (1.0f - t2) * 21.5473f;
    b = t2 * 3.32435f + (1.0f - t2) * 3.82987f;
    c = t2 * 0.16801f + (1.0f - t2) * 0.19823f;
    d = t2 * -1.27393f + (1.0f - t2) * -1.97760f;
    e = t2 * -4.85967f + (1.0f - t2) * -4.32054f;
    inv_r = 1.0f / roughness;
  }
  FR_D float L(float x) const { return a / (1.0f + b * powf(x, c)) + d * x + e; }
  FR_D float lambda(const float3& w) const
  {
    const float c0 = abs_cos(w);
    return (c0 < 0.5f) ? expf(L(c0)) : expf(2.0f * L(0.5f) - L(1.0f - c0));
  }
  FR_D float D(const float3& h) const
  {
    const float s = sqrtf(fmaxf(1.0f - h.y * h.y, 0.0f));
    return (2.0f + inv_r) * powf(s, inv_r) / (2.0f * kPi);
  }
};

// Oren-Nayar in the qualitative A/B form (bxdf.cu:151-205); also used, mirrored,
// as the diffuse-transmission lobe (bxdf.cu:209-264)
FR_D float oren_nayar_scale(float A, float B, const float3& wo, const float3& wi)
{
  const float so = sqrtf(fmaxf(1.0f - wo.y * wo.y, 0.0f));
  const float si = sqrtf(fmaxf(1.0f - wi.y * wi.y, 0.0f));
  float c_max = 0.0f;
  if (si > 1e-4f && so > 1e-4f) {
    const float c = (wi.x / si) * (wo.x / so) + (wi.z / si) * (wo.z / so);
    c_max = fmaxf(c, 0.0f);
  }
  const bool b = abs_cos(wi) > abs_cos(wo);
  const float s_alpha = b ? so : si;
  const float t_beta = b ? si / abs_cos(wi) : so / abs_cos(wo);
  return (A + B * c_max * s_alpha * t_beta) / kPi;
}

FR_D float3 guard(const float3& v) { return bad3(v) ? f3(0.0f) : v; }
FR_D float guard(float v) { return (isinf(v) || isnan(v)) ? 0.0f : v; }

// -----------------------------------------------------------------------------
template <uint32_t MASK>
struct Closure {
  static constexpr bool kCoat = MASK & M_COAT, kMetal = MASK & M_METAL, kSpec = MASK & M_SPECULAR,
                        kTrans = MASK & M_TRANSMISSION, kSheen = MASK & M_SHEEN, kDiffT = MASK & M_DIFFUSE_T;
  float3 wo;
  // layer scalars after the "seen from inside" masking (bsdf.cu:56-62)
  float coat, metalness, specular, transmission, sheen, subsurface, thin_walled, diffuse;
  float3 base_color, specular_color, transmission_color, sheen_color, subsurface_color;
  float3 coat_absorption;
  bool coat_on, spec_on, sheen_on;  // luminance gates (bsdf.cu:132,144,159)
  bool metal_on, trans_on, difft_on;
  float spec_albedo, sheen_albedo;
  float a_coat, a_spec;  // GGX alpha
  float ni, nt, eta;
  float3 metal_n, metal_k;
  float on_A, on_B;
  SheenFit sheen_fit;
  float pmf[LOBE_COUNT];  // differences of the reference's float CDF

  FR_D void init(const float3& wo_, const SurfaceParams& p, bool entering)
  {
    wo = wo_;
    ni = entering ? 1.0f : 1.5f;
    nt = entering ? 1.5f : 1.0f;
    eta = nt / ni;

    const float coat_lum = luminance(p.coat_color);
    const float spec_lum = luminance(p.specular_color);
    const float sheen_lum = luminance(p.sheen_color);

    // quirk: formed while the coat albedo member is still 0 (bsdf.cu:26-30)
    coat_absorption = f3(1.0f) + p.coat * (p.coat_color * (1.0f - 0.0f) - f3(1.0f));

    const float f0t = (nt - ni) / (nt + ni);
    const float F0 = f0t * f0t;
    float coat_albedo = 0.0f;
    if (kCoat && p.coat * coat_lum > 0.0f) coat_albedo = entering ? albedo_ggx(wo.y, p.coat_roughness, F0) : 0.0f;
    spec_albedo = 0.0f;
    if (kSpec && p.specular * spec_lum > 0.0f)
      spec_albedo = eta >= 1.0f ? albedo_ggx(wo.y, p.specular_roughness, F0) : 0.0f;
    sheen_albedo = 0.0f;
    if (kSheen && p.sheen * sheen_lum) sheen_albedo = entering ? albedo_sheen(wo.y, p.sheen_roughness) : 0.0f;

    coat = entering ? p.coat : 0.0f;
    metalness = entering ? p.metalness : 0.0f;
    specular = entering ? p.specular : 0.0f;
    sheen = entering ? p.sheen : 0.0f;
    diffuse = entering ? p.diffuse : 0.0f;
    transmission = p.transmission;
    subsurface = p.subsurface;
    thin_walled = p.thin_walled;
    base_color = p.base_color;
    specular_color = p.specular_color;
    transmission_color = p.transmission_color;
    sheen_color = p.sheen_color;
    subsurface_color = p.subsurface_color;
    coat_on = kCoat && coat * coat_lum > 0.0f;
    spec_on = kSpec && specular * spec_lum > 0.0f;
    sheen_on = kSheen && sheen * sheen_lum > 0.0f;
    metal_on = kMetal && metalness > 0.0f;
    trans_on = kTrans && transmission > 0.0f;
    difft_on = kDiffT && subsurface * thin_walled > 0.0f;

    // layer weights (bsdf.cu:67-93)
    float w[LOBE_COUNT];
    const float below_coat = 1.0f - coat * coat_albedo;
    w[LOBE_COAT] = coat * coat_albedo;
    w[LOBE_METAL] = below_coat * metalness;
    w[LOBE_SPECULAR] = below_coat * (1.0f - metalness) * specular * spec_albedo;
    w[LOBE_TRANSMISSION] =
        below_coat * (1.0f - metalness) * (1.0f - specular * spec_albedo) * transmission;
    w[LOBE_SHEEN] = below_coat * (1.0f - metalness) * (1.0f - specular * spec_albedo) * sheen *
                    sheen_albedo;
    w[LOBE_DIFFUSE_T] = below_coat * (1.0f - metalness) * (1.0f - specular * spec_albedo) *
                        (1.0f - transmission) * (1.0f - sheen * sheen_albedo) * subsurface *
                        thin_walled;
    w[LOBE_DIFFUSE_R] = below_coat * (1.0f - metalness) * (1.0f - specular * spec_albedo) *
                        (1.0f - transmission) * (1.0f - sheen * sheen_albedo) *
                        (1.0f - subsurface) * diffuse;

    // the reference keeps a running float CDF and later takes differences of it
    // (sampling.cu:115-151); reproduce those roundings so lobe choice agrees
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < LOBE_COUNT; ++i) sum += w[i];
    float cdf_prev = 0.0f;
#pragma unroll
    for (int i = 0; i < LOBE_COUNT; ++i) {
      const float cdf = cdf_prev + w[i] / sum;
      pmf[i] = cdf - cdf_prev;
      cdf_prev = cdf;
    }

    a_coat = p.coat_roughness * p.coat_roughness;
    a_spec = p.specular_roughness * p.specular_roughness;

    // artist-friendly metallic Fresnel, Gulbrandsen 2014 (bxdf.cu:107-116)
    const float3 r = clamp3(p.base_color, 0.0f, 0.99f);
    const float3 g = clamp3(p.specular_color, 0.0f, 0.99f);
    const float3 rs = sqrt3(r);
    metal_n = g * (1.0f - r) / (1.0f + r) + (1.0f - g) * (1.0f + rs) / (1.0f - rs);
    const float3 n1 = metal_n + 1.0f, n2 = metal_n - 1.0f;
    metal_k = sqrt3((r * (n1 * n1) - n2 * n2) / (1.0f - r));

    const float sigma2 = p.diffuse_roughness * p.diffuse_roughness;
    on_A = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.33f)));
    on_B = 0.45f * sigma2 / (sigma2 + 0.09f);
    if (kSheen) sheen_fit.init(p.sheen_roughness);
  }

  // ---- per-lobe value / pdf ---------------------------------------------------
  // microfacet reflection, shared by coat / specular / metal: value without Fresnel
  FR_D void ggx_reflect_terms(float a, const float3& wi, const float3& h, float& dg, float& pdf) const
  {
    const float d = ggx_D(a, h);
    const float lo = ggx_lambda(a, wo), li = ggx_lambda(a, wi);
    const float g2 = 1.0f / (1.0f + lo + li);
    const float g1 = 1.0f / (1.0f + lo);
    const float odh = fabsf(dot(wo, h));
    dg = 0.25f * (d * g2) / (abs_cos(wo) * abs_cos(wi));
    pdf = 0.25f * (g1 * odh * d / abs_cos(wo)) / odh;
  }

  FR_D float3 transmission_half(const float3& wi) const
  {
    float3 h = normalize(-(ni * wo + nt * wi));
    if (h.y < 0.0f) h = -h;
    return h;
  }
  FR_D float3 transmission_value(const float3& wi, const float3& h) const
  {
    const float f = fresnel_dielectric(fabsf(dot(wo, h)), nt / ni);
    const float d = ggx_D(a_spec, h);
    const float g2 = 1.0f / (1.0f + ggx_lambda(a_spec, wo) + ggx_lambda(a_spec, wi));
    const float odh = dot(wo, h), idh = dot(wi, h);
    const float t = ni * odh + nt * idh;
    const float v = fabsf(odh) * fabsf(idh) * nt * nt * fmaxf(1.0f - f, 0.0f) * g2 * d /
                    (abs_cos(wo) * abs_cos(wi) * t * t);
    return f3(v);
  }
  FR_D float transmission_pdf(const float3& wi, const float3& h) const
  {
    const float idh = dot(wi, h);
    const float t = ni * dot(wo, h) + nt * idh;
    const float g1 = 1.0f / (1.0f + ggx_lambda(a_spec, wo));
    const float dv = g1 * fabsf(dot(wo, h)) * ggx_D(a_spec, h) / abs_cos(wo);
    return dv * nt * nt * fabsf(idh) / (t * t);
  }
  FR_D float3 sheen_value(const float3& wi) const
  {
    const float3 h = normalize(wo + wi);
    const float d = sheen_fit.D(h);
    const float g = 1.0f / (1.0f + sheen_fit.lambda(wo) + sheen_fit.lambda(wi));
    return f3(0.25f * (1.0f * d * g) / (abs_cos(wo) * abs_cos(wi)));
  }

  // ---- full evaluation: value and pdf for a given direction (bsdf.cu:129-212,
  // 295-345) ---------------------------------------------------------------------
  FR_D void eval(const float3& wi, float3& f_out, float& pdf_out) const
  {
    const float cos_pdf = abs_cos(wi) / kPi;
    const bool need_refl = coat_on || spec_on || metal_on;
    float3 h = f3(0.0f);
    if (need_refl || sheen_on) h = normalize(wo + wi);

    float3 v_coat = f3(0.0f), v_metal = f3(0.0f), v_spec = f3(0.0f);
    float p_coat = 0.0f, p_metal = 0.0f, p_spec = 0.0f;
    if (coat_on) {
      float dg, pdf;
      ggx_reflect_terms(a_coat, wi, h, dg, pdf);
      v_coat = guard(f3(fresnel_dielectric(fabsf(dot(wo, h)), eta) * dg));
      p_coat = guard(pdf);
    }
    if (spec_on || metal_on) {
      float dg, pdf;
      ggx_reflect_terms(a_spec, wi, h, dg, pdf);
      if (metal_on) {
        v_metal = guard(fresnel_conductor(fabsf(dot(wo, h)), metal_n, metal_k) * dg);
        p_metal = guard(pdf);
      }
      if (spec_on) {
        v_spec = guard(f3(fresnel_dielectric(fabsf(dot(wo, h)), eta) * dg));
        p_spec = guard(pdf);
      }
    }
    float3 v_trans = f3(0.0f);
    float p_trans = 0.0f;
    if (trans_on) {
      const float3 ht = transmission_half(wi);
      v_trans = guard(transmission_value(wi, ht));
      p_trans = guard(transmission_pdf(wi, ht));
    }
    float3 v_sheen = f3(0.0f);
    float p_sheen = 0.0f;
    if (sheen_on) {
      v_sheen = guard(sheen_value(wi));
      p_sheen = guard(cos_pdf);
    }
    float3 v_dt = f3(0.0f), v_dr = f3(0.0f);
    float p_dt = 0.0f, p_dr = 0.0f;
    if (difft_on || diffuse > 0.0f) {
      const float3 on = guard(base_color * oren_nayar_scale(on_A, on_B, wo, wi));
      if (difft_on) {
        v_dt = on;
        p_dt = guard(cos_pdf);
      }
      if (diffuse > 0.0f) {
        v_dr = on;
        p_dr = guard(cos_pdf);
      }
    }

    // energy-compensated layering (bsdf.cu:178-211)
    float3 ret = coat * v_coat;
    float3 mult = coat_absorption;
    ret += mult * metalness * v_metal;
    mult *= (1.0f - metalness);
    ret += mult * specular * specular_color * v_spec;
    mult *= (1.0f - specular * specular_color * spec_albedo);
    ret += mult * transmission * transmission_color * v_trans;
    mult *= (1.0f - transmission);
    ret += mult * sheen * sheen_color * v_sheen;
    mult *= (1.0f - sheen * sheen_albedo);
    ret += mult * subsurface * subsurface_color * thin_walled * v_dt;
    mult *= (1.0f - subsurface);
    ret += mult * diffuse * v_dr;
    f_out = ret;

    pdf_out = pmf[LOBE_COAT] * p_coat + pmf[LOBE_METAL] * p_metal + pmf[LOBE_SPECULAR] * p_spec +
              pmf[LOBE_TRANSMISSION] * p_trans + pmf[LOBE_SHEEN] * p_sheen +
              pmf[LOBE_DIFFUSE_T] * p_dt + pmf[LOBE_DIFFUSE_R] * p_dr;
  }

  // ---- sampling: choose a lobe with the 1-D number, sample it with the 2-D number;
  // f is ONLY the chosen lobe's contribution scaled by its layer factor, pdf is
  // lobe pdf x lobe probability (bsdf.cu:214-293) --------------------------------
  FR_D int pick_lobe(float u, float& prob) const
  {
    float cdf = 0.0f;
#pragma unroll
    for (int i = 0; i < LOBE_COUNT; ++i) {
      cdf += pmf[i];
      if (u < cdf) {
        prob = pmf[i];
        return i;
      }
    }
    prob = pmf[LOBE_COUNT - 1];
    return LOBE_COUNT - 1;
  }

  FR_D float3 sample(float u, const float2& v, float3& f, float& pdf) const
  {
    float prob;
    const int lobe = pick_lobe(u, prob);
    float3 wi;
    // lobes outside MASK have zero probability; their cases compile away
    int lobe_c = lobe;
    if (!kCoat && lobe_c == LOBE_COAT) lobe_c = LOBE_DIFFUSE_R;
    if (!kMetal && lobe_c == LOBE_METAL) lobe_c = LOBE_DIFFUSE_R;
    if (!kSpec && lobe_c == LOBE_SPECULAR) lobe_c = LOBE_DIFFUSE_R;
    if (!kTrans && lobe_c == LOBE_TRANSMISSION) lobe_c = LOBE_DIFFUSE_R;
    if (!kSheen && lobe_c == LOBE_SHEEN) lobe_c = LOBE_DIFFUSE_R;
    if (!kDiffT && lobe_c == LOBE_DIFFUSE_T) lobe_c = LOBE_DIFFUSE_R;
    switch (lobe_c) {
      case LOBE_COAT:
      case LOBE_METAL:
      case LOBE_SPECULAR: {
        if (!(kCoat || kMetal || kSpec)) break;
        const float a = lobe == LOBE_COAT ? a_coat : a_spec;
        const float3 h = sample_vndf(wo, a, v);
        wi = mirror(wo, h);
        // the reference re-derives the half vector from (wo, wi) for value and pdf
        const float3 hh = normalize(wo + wi);
        float dg, p;
        ggx_reflect_terms(a, wi, hh, dg, p);
        const float c = fabsf(dot(wo, hh));
        if (lobe == LOBE_COAT) {
          f = f3(fresnel_dielectric(c, eta) * dg) * coat;
        } else if (lobe == LOBE_METAL) {
          f = fresnel_conductor(c, metal_n, metal_k) * dg * (coat_absorption * metalness);
        } else {
          f = f3(fresnel_dielectric(c, eta) * dg) *
              (coat_absorption * (1.0f - metalness) * specular * specular_color);
        }
        pdf = p;
      } break;
      case LOBE_TRANSMISSION: {
        if (!kTrans) break;
        const float3 h = sample_vndf(wo, a_spec, v);
        const float3 th = -ni / nt * (wo - dot(wo, h) * h);
        const float th2 = dot(th, th);
        if (th2 > 1.0f) {
          // total internal reflection (bxdf.cu:660-679)
          wi = mirror(wo, h);
          const float fr = fresnel_dielectric(fabsf(dot(wo, h)), nt / ni);
          const float d = ggx_D(a_spec, h);
          const float lo = ggx_lambda(a_spec, wo);
          const float g2 = 1.0f / (1.0f + lo + ggx_lambda(a_spec, wi));
          f = f3(0.25f * (fr * d * g2) / (abs_cos(wo) * abs_cos(wi)));
          const float dv = (1.0f / (1.0f + lo)) * fabsf(dot(wo, h)) * d / abs_cos(wo);
          pdf = 0.25f * dv / fabsf(dot(wi, h));
        } else {
          wi = th + (-sqrtf(fmaxf(1.0f - th2, 0.0f)) * h);
          const float3 ht = transmission_half(wi);
          f = transmission_value(wi, ht);
          pdf = transmission_pdf(wi, ht);
        }
        f *= coat_absorption * (1.0f - metalness) *
             (1.0f - specular * specular_color * spec_albedo) * transmission * transmission_color;
      } break;
      case LOBE_SHEEN: {
        if (!kSheen) break;
        const float3 h = cosine_hemisphere_local(v);
        wi = mirror(wo, h);
        f = sheen_value(wi) * (coat_absorption * (1.0f - metalness) *
                               (1.0f - specular * specular_color * spec_albedo) *
                               (1.0f - transmission) * sheen * sheen_color);
        pdf = abs_cos(wi) / kPi;
      } break;
      case LOBE_DIFFUSE_T: {
        if (!kDiffT) break;
        wi = -cosine_hemisphere_local(v);
        f = base_color * oren_nayar_scale(on_A, on_B, wo, wi) *
            (coat_absorption * (1.0f - metalness) *
             (1.0f - specular * specular_color * spec_albedo) * (1.0f - transmission) *
             (1.0f - sheen * sheen_albedo) * subsurface * subsurface_color * thin_walled);
        pdf = abs_cos(wi) / kPi;
      } break;
      default: {
        wi = cosine_hemisphere_local(v);
        f = base_color * oren_nayar_scale(on_A, on_B, wo, wi) *
            (coat_absorption * (1.0f - metalness) *
             (1.0f - specular * specular_color * spec_albedo) * (1.0f - transmission) *
             (1.0f - sheen * sheen_albedo) * (1.0f - subsurface) * diffuse);
        pdf = abs_cos(wi) / kPi;
      } break;
    }
    pdf *= prob;
    return wi;
  }

  // cosine-weighted hemisphere around +y via the concentric map
  FR_D static float3 cosine_hemisphere_local(const float2& u)
  {
    const float a = 2.0f * u.x - 1.0f, b = 2.0f * u.y - 1.0f;
    float dx = 0.0f, dz = 0.0f;
    if (!(a == 0.0f && b == 0.0f)) {
      float r, theta;
      if (fabsf(a) > fabsf(b)) {
        r = a;
        theta = 0.25f * kPi * b / a;
      } else {
        r = b;
        theta = 0.5f * kPi - 0.25f * kPi * a / b;
      }
      float s, c;
      sincosf(theta, &s, &c);
      dx = r * c;
      dz = r * s;
    }
    return f3(dx, sqrtf(fmaxf(0.0f, 1.0f - dx * dx - dz * dz)), dz);
  }
};

}  // namespace frd
