// Camera ray generation and the sky / miss stage.
//   thin-lens camera        reference camera.cu:24-53 (pinhole variant is unused)
//   concentric disk map     reference sampling.cu:54-64
//   Hosek-Wilkie RGB sky    reference arhosek.cu:103-127, pt.cu:352-363
#pragma once
#include "fredholm/shared.h"
#include "vecmath.cuh"

namespace frd
{

FR_HD float3 xform_point(const fredholm::Matrix3x4& m, const float3& p)
{
  return f3(m.m[0].x * p.x + m.m[0].y * p.y + m.m[0].z * p.z + m.m[0].w,
            m.m[1].x * p.x + m.m[1].y * p.y + m.m[1].z * p.z + m.m[1].w,
            m.m[2].x * p.x + m.m[2].y * p.y + m.m[2].z * p.z + m.m[2].w);
}
FR_HD float3 xform_vector(const fredholm::Matrix3x4& m, const float3& v)
{
  return f3(m.m[0].x * v.x + m.m[0].y * v.y + m.m[0].z * v.z,
            m.m[1].x * v.x + m.m[1].y * v.y + m.m[1].z * v.z,
            m.m[2].x * v.x + m.m[2].y * v.y + m.m[2].z * v.z);
}
// normals go through the transpose of world_to_object (shared.h:42-50)
FR_HD float3 xform_normal(const fredholm::Matrix3x4& w2o, const float3& n)
{
  return f3(w2o.m[0].x * n.x + w2o.m[1].x * n.y + w2o.m[2].x * n.z,
            w2o.m[0].y * n.x + w2o.m[1].y * n.y + w2o.m[2].y * n.z,
            w2o.m[0].z * n.x + w2o.m[1].z * n.y + w2o.m[2].z * n.z);
}

// Shirley-Chiu concentric map of [0,1)^2 onto the unit disk
FR_HD float2 concentric_disk(const float2& u)
{
  const float a = 2.0f * u.x - 1.0f;
  const float b = 2.0f * u.y - 1.0f;
  if (a == 0.0f && b == 0.0f) return make_float2(0.0f, 0.0f);
  float r, theta;
  if (fabsf(a) > fabsf(b)) {
    r = a;
    theta = 0.25f * kPi * b / a;
  } else {
    r = b;
    theta = 0.5f * kPi - 0.25f * kPi * a / b;
  }
  return make_float2(r * cosf(theta), r * sinf(theta));
}

// cosine-weighted hemisphere around +y (sampling.cu:66-78)
FR_HD float3 cosine_hemisphere(const float2& u)
{
  const float2 d = concentric_disk(u);
  return f3(d.x, sqrtf(fmaxf(0.0f, 1.0f - d.x * d.x - d.y * d.y)), d.y);
}

// Constants of the thin-lens model that do not depend on the pixel.
struct LensModel {
  float f;            // sensor-to-lens distance, 1/tan(fov/2)
  float lens_radius;  // 2 f / F
  float ab;           // a + b: sensor-to-focus-plane distance term
  FR_HD void init(const fredholm::CameraParams& c)
  {
    f = 1.0f / tanf(0.5f * c.fov);
    const float b = c.focus;
    const float a = 1.0f / (1.0f + f - 1.0f / b);
    lens_radius = 2.0f * f / c.F;
    ab = a + b;
  }
};

// uv: sensor position (already x-flipped, pt.cu:439-442); u: lens sample
FR_HD void thin_lens_ray(const fredholm::CameraParams& cam, const LensModel& lm,
                         const float2& uv, const float2& u, float3& origin,
                         float3& direction)
{
  const float3 sensor = f3(uv.x, uv.y, 0.0f);
  const float3 lens_c = f3(0.0f, 0.0f, lm.f);
  const float2 disk = concentric_disk(u);
  const float3 lens_p = f3(lm.lens_radius * disk.x, lm.lens_radius * disk.y, lm.f);
  const float3 to_c = normalize(lens_c - sensor);
  const float3 focus_p = sensor + (lm.ab / to_c.z) * to_c;
  origin = xform_point(cam.transform, lens_p);
  float3 dir = normalize(focus_p - lens_p);
  dir.z = -dir.z;  // the reference's "adhoc fix" (camera.cu:48)
  direction = xform_vector(cam.transform, dir);
}

// ---- sky -------------------------------------------------------------------
// Cooked Hosek state: 3 channels x 9 coefficients + 3 mean radiances (the only
// members of the reference's 544-byte ArHosekSkyModelState the device reads).
struct HosekSky {
  float cfg[3][9];
  float rad[3];
};

enum SkyMode : int { SKY_CONSTANT = 0, SKY_HOSEK = 1, SKY_IBL = 2 };

// Closed-form radiance for all three channels at once; the terms that do not
// depend on the channel (cos/exp of the angles) are evaluated once instead of
// three times.  sqrtf(cos theta) is NaN below the horizon
// exactly as in the reference -- callers rely on the NaN guard of the film stage.
FR_D float3 hosek_radiance(const HosekSky& s, const float3& v, const float3& sun_dir)
{
  // cos(theta) and cos(gamma) are the dot products themselves (the reference takes acos and then
  // cos again, arhosek.cu:103-127; equal to an ulp), and x^1.5 is x * sqrt(x): the three powf calls
  // were 7 % of the shade stage's instructions (profiles/r1i_shade_lines_1.txt)
  const float ct = clampf(v.y, -1.0f, 1.0f);
  const float cg = dot(sun_dir, v);
  const float gamma = acosf(cg);  // NaN for |cg| > 1, as in the reference
  const float ray_m = cg * cg;
  const float zenith = sqrtf(ct);
  float out[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* k = s.cfg[c];
    const float exp_m = expf(k[4] * gamma);
    const float x = 1.0f + k[8] * k[8] - 2.0f * k[8] * cg;
    const float mie_m = (1.0f + cg * cg) / (x * sqrtf(x));
    out[c] = (1.0f + k[0] * expf(k[1] / (ct + 0.01f))) *
             (k[2] + k[3] * exp_m + k[5] * ray_m + k[6] * mie_m + k[7] * zenith) * s.rad[c];
  }
  return f3(out[0], out[1], out[2]);
}

}  // namespace frd
