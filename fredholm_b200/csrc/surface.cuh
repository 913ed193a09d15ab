// Surface reconstruction at a hit: texture fetch, emission, interpolation of the
// vertex attributes, ray-origin offsetting.  Shared by the shade stage and by the
// epilogue of the MIS-ray traversal (emitter record) and the alpha test.
//   reference: pt.cu:125-179 (emission, fill_surface_info), pt.cu:402-416
//   (ray_origin_offset), pt.cu:545-588 (alpha test), cwl/texture.h:35-47 (sampler).
#pragma once
#include "wavefront.h"

namespace frd
{

FR_D int wrap_index(int i, int n)
{
  i %= n;
  return i < 0 ? i + n : i;
}

// Software version of the reference's texture sampler state: normalized
// coordinates, wrap addressing, bilinear filter, 8-bit -> [0,1], optional sRGB
// decode through a 256-entry table (fp32 weights; see DESIGN.md "textures").
FR_D float4 texel_rgba(const TexView& t, const float* __restrict__ srgb_lut, int i, int j)
{
  const uchar4 c = __ldg(t.texels + (size_t)j * t.width + i);
  if (t.srgb) return make_float4(__ldg(srgb_lut + c.x), __ldg(srgb_lut + c.y), __ldg(srgb_lut + c.z), c.w / 255.0f);
  return make_float4(c.x / 255.0f, c.y / 255.0f, c.z / 255.0f, c.w / 255.0f);
}

FR_D float4 tex_fetch(const TexView& t, const float* __restrict__ srgb_lut, float x, float y)
{
  const float xb = x * t.width - 0.5f, yb = y * t.height - 0.5f;
  const float fx = floorf(xb), fy = floorf(yb);
  const float a = xb - fx, b = yb - fy;
  const int i0 = wrap_index((int)fx, (int)t.width), i1 = wrap_index((int)fx + 1, (int)t.width);
  const int j0 = wrap_index((int)fy, (int)t.height), j1 = wrap_index((int)fy + 1, (int)t.height);
  const float4 t00 = texel_rgba(t, srgb_lut, i0, j0), t10 = texel_rgba(t, srgb_lut, i1, j0);
  const float4 t01 = texel_rgba(t, srgb_lut, i0, j1), t11 = texel_rgba(t, srgb_lut, i1, j1);
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  return make_float4(w00 * t00.x + w10 * t10.x + w01 * t01.x + w11 * t11.x,
                     w00 * t00.y + w10 * t10.y + w01 * t01.y + w11 * t11.y,
                     w00 * t00.z + w10 * t10.z + w01 * t01.z + w11 * t11.z,
                     w00 * t00.w + w10 * t10.w + w01 * t01.w + w11 * t11.w);
}

struct SceneTex {
  const TexView* textures;
  const float* srgb_lut;
  FR_D float4 fetch(int id, const float2& uv) const { return tex_fetch(textures[id], srgb_lut, uv.x, uv.y); }
};

FR_D bool is_emissive(const fredholm::Material& m)
{
  return m.emission_color.x > 0 || m.emission_color.y > 0 || m.emission_color.z > 0 ||
         m.emission_texture_id != -1;
}

FR_D float3 emission_of(const fredholm::Material& m, const SceneTex& tex, const float2& uv)
{
  return m.emission_texture_id >= 0 ? f3(tex.fetch(m.emission_texture_id, uv)) : m.emission_color;
}

// RT Gems ch. 6 self-intersection offset
FR_D float3 offset_origin(const float3& p, const float3& n)
{
  constexpr float origin = 1.0f / 32.0f;
  constexpr float float_scale = 1.0f / 65536.0f;
  constexpr float int_scale = 256.0f;
  const int ix = (int)(int_scale * n.x), iy = (int)(int_scale * n.y), iz = (int)(int_scale * n.z);
  const float px = __int_as_float(__float_as_int(p.x) + (p.x < 0 ? -ix : ix));
  const float py = __int_as_float(__float_as_int(p.y) + (p.y < 0 ? -iy : iy));
  const float pz = __int_as_float(__float_as_int(p.z) + (p.z < 0 ? -iz : iz));
  return f3(fabsf(p.x) < origin ? p.x + float_scale * n.x : px,
            fabsf(p.y) < origin ? p.y + float_scale * n.y : py,
            fabsf(p.z) < origin ? p.z + float_scale * n.z : pz);
}

// world-space vertices / attributes of one face
struct FaceGeom {
  float3 v0, v1, v2;
  float3 n0, n1, n2;  // transformed by transpose(world_to_object), not normalized
  float2 t0, t1, t2;
};

FR_D FaceGeom load_face(const SceneView& sc, const uint3& idx, uint32_t xform)
{
  FaceGeom g;
  const fredholm::Matrix3x4 o2w = sc.o2w[xform];
  const fredholm::Matrix3x4 w2o = sc.w2o[xform];
  g.v0 = xform_point(o2w, sc.vertices[idx.x]);
  g.v1 = xform_point(o2w, sc.vertices[idx.y]);
  g.v2 = xform_point(o2w, sc.vertices[idx.z]);
  g.n0 = xform_normal(w2o, sc.normals[idx.x]);
  g.n1 = xform_normal(w2o, sc.normals[idx.y]);
  g.n2 = xform_normal(w2o, sc.normals[idx.z]);
  g.t0 = sc.texcoords[idx.x];
  g.t1 = sc.texcoords[idx.y];
  g.t2 = sc.texcoords[idx.z];
  return g;
}

FR_D float3 bary3(const float3& a, const float3& b, const float3& c, float u, float v)
{
  return (1.0f - u - v) * a + u * b + v * c;
}
FR_D float2 bary2(const float2& a, const float2& b, const float2& c, float u, float v)
{
  const float w = 1.0f - u - v;
  return make_float2(w * a.x + u * b.x + v * c.x, w * a.y + u * b.y + v * c.y);
}

// sky radiance seen along `dir` (priority IBL > Hosek > constant, pt.cu:511-517)
FR_D float3 sky_radiance(const SceneView& sc, const float3& dir)
{
  if (sc.sky_mode == SKY_IBL) {
    float theta = acosf(clampf(dir.y, -1.0f, 1.0f));
    float phi = atan2f(dir.z, dir.x);
    if (phi < 0) phi += 2.0f * kPi;
    const float x = phi / (2.0f * kPi), y = theta / kPi;
    // float4 lat-long image, same sampler state as 8-bit textures
    const float xb = x * sc.ibl_width - 0.5f, yb = y * sc.ibl_height - 0.5f;
    const float fx = floorf(xb), fy = floorf(yb);
    const float a = xb - fx, b = yb - fy;
    const int w = (int)sc.ibl_width, h = (int)sc.ibl_height;
    const int i0 = wrap_index((int)fx, w), i1 = wrap_index((int)fx + 1, w);
    const int j0 = wrap_index((int)fy, h), j1 = wrap_index((int)fy + 1, h);
    const float4 t00 = __ldg(sc.ibl_texels + (size_t)j0 * w + i0), t10 = __ldg(sc.ibl_texels + (size_t)j0 * w + i1);
    const float4 t01 = __ldg(sc.ibl_texels + (size_t)j1 * w + i0), t11 = __ldg(sc.ibl_texels + (size_t)j1 * w + i1);
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    return sc.sky_intensity * f3(w00 * t00.x + w10 * t10.x + w01 * t01.x + w11 * t11.x,
                                 w00 * t00.y + w10 * t10.y + w01 * t01.y + w11 * t11.y,
                                 w00 * t00.z + w10 * t10.z + w01 * t01.z + w11 * t11.z);
  }
  if (sc.sky_mode == SKY_HOSEK) return sc.sky_intensity * hosek_radiance(sc.hosek, dir, sc.sun_dir);
  return sc.bg_color;
}

}  // namespace frd
