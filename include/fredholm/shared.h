// Host/device plain-data contract of the fredholm rendering core.
//
// These structs are the data interface between an application and the
// renderer; field order, types and defaults mirror the reference's
// fredholm/include/fredholm/shared.h so that applications written against the
// reference (scene arrays, AOV buffers, materials) keep working:
//   Matrix3x4        shared.h:11-14     CameraParams     shared.h:59-64
//   Material (180 B) shared.h:100-142   AreaLight        shared.h:149-153
//   DirectionalLight shared.h:155-159   RenderLayer      shared.h:201-208
// OptiX-only types of the reference (LaunchParams, SBT records, sampler state
// structs) have no equivalent here: the wavefront core keeps its own compact
// per-path state (fredholm_b200/csrc/wavefront.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace fredholm
{

// three rows of a row-major 3x4 affine transform
struct Matrix3x4 {
  float4 m[3];
};

inline Matrix3x4 make_mat3x4(const float4& r0, const float4& r1, const float4& r2)
{
  Matrix3x4 m;
  m.m[0] = r0;
  m.m[1] = r1;
  m.m[2] = r2;
  return m;
}

enum class RayType : unsigned int {
  RAY_TYPE_RADIANCE = 0,
  RAY_TYPE_SHADOW = 1,
  RAY_TYPE_LIGHT = 2,
  RAY_TYPE_COUNT
};

struct CameraParams {
  Matrix3x4 transform;  // camera to world
  float fov;            // vertical field of view [rad]
  float F;              // F number
  float focus;          // focus distance
};

// Arnold-Standard-Surface-like material (https://autodesk.github.io/standard-surface/)
struct Material {
  float diffuse = 1.0f;
  float3 base_color = make_float3(1, 1, 1);
  int base_color_texture_id = -1;
  float diffuse_roughness = 0.0f;

  float specular = 1.0f;
  float3 specular_color = make_float3(1, 1, 1);
  int specular_color_texture_id = -1;
  float specular_roughness = 0.2f;
  int specular_roughness_texture_id = -1;

  float metalness = 0;
  int metalness_texture_id = -1;

  int metallic_roughness_texture_id = -1;

  float coat = 0;
  int coat_texture_id = -1;
  float3 coat_color = make_float3(1, 1, 1);
  float coat_roughness = 0.1f;
  int coat_roughness_texture_id = -1;

  float transmission = 0;
  float3 transmission_color = make_float3(1, 1, 1);

  float sheen = 0.0f;
  float3 sheen_color = make_float3(1.0f, 1.0f, 1.0f);
  float sheen_roughness = 0.3f;

  float subsurface = 0;
  float3 subsurface_color = make_float3(1.0f, 1.0f, 1.0f);

  float thin_walled = 0.0f;

  float emission = 0;
  float3 emission_color = make_float3(0, 0, 0);
  int emission_texture_id = -1;

  int heightmap_texture_id = -1;
  int normalmap_texture_id = -1;
  int alpha_texture_id = -1;
};
static_assert(sizeof(Material) == 180, "Material must stay layout-compatible (180 bytes)");

struct AreaLight {
  uint3 indices;              // vertex indices of the emissive face
  unsigned int material_id;
  unsigned int instance_idx;  // transform used for the face
};

struct DirectionalLight {
  float3 le;        // emitted radiance
  float3 dir;       // direction TO the light, normalized
  float angle = 0;  // angular diameter [deg]
};

// Caller-owned device AOV buffers, one element per pixel, row 0 = image top.
struct RenderLayer {
  float4* beauty;
  float4* position;
  float* depth;
  float4* normal;
  float4* texcoord;
  float4* albedo;
};

}  // namespace fredholm
