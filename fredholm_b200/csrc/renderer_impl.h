// Private state of fredholm::Renderer (shared by renderer.cpp and the C ABI).
#pragma once
#include <memory>
#include <vector>

#include "bvh_build.h"
#include "cuda_util.h"
#include "fredholm/renderer.h"
#include "integrator.h"

namespace fredholm
{

struct Renderer::Impl {
  int device = 0;
  cudaStream_t stream = nullptr;

  uint32_t width = 0, height = 0;
  uint32_t sample_count = 0;  // uniform over the image (reference: per-pixel buffer, always uniform)
  FilmMode film_mode = FilmMode::MEAN;

  Scene scene;

  // scene on device
  frd::DevBuf<float3> d_vertices, d_normals;
  frd::DevBuf<float2> d_texcoords;
  frd::DevBuf<uint3> d_indices;
  frd::DevBuf<uint32_t> d_material_ids, d_face_submesh, d_face_flags, d_submesh_offsets;
  frd::DevBuf<uint8_t> d_face_class;
  uint32_t class_mask = 0;  // ShadeClass values present in the scene
  frd::DevBuf<Material> d_materials;
  std::vector<frd::DevBuf<uchar4>> d_texture_data;
  frd::DevBuf<frd::TexView> d_textures;
  frd::DevBuf<float> d_srgb_lut;
  frd::DevBuf<Matrix3x4> d_o2w, d_w2o;
  frd::DevBuf<AreaLight> d_lights;
  uint32_t n_lights = 0;
  frd::DevBuf<float4> d_ibl;
  uint32_t ibl_w = 0, ibl_h = 0;

  bool has_dir_light = false;
  DirectionalLight dir_light;
  float sky_intensity = 1.0f;
  float3 sun_direction = make_float3(0.0f, 1.0f, 0.0f);
  bool has_hosek = false;
  frd::HosekSky hosek;

  frd::DeviceBvh bvh;           // flat world-space tree
  frd::TwoLevelBvh bvh2;        // per-mesh trees + instance tree
  bool two_level = false;       // which of the two the last build produced
  AccelMode accel_mode = AccelMode::AUTO;
  bool accel_valid = false;
  AccelInfo accel_info;
  // distinct meshes of the scene (sub-meshes with identical triangle positions share one): filled by upload_scene
  std::vector<uint32_t> mesh_of_submesh, mesh_representative;
  void find_distinct_meshes();
  bool want_two_level() const;
  void refresh_accel_info(float build_ms);
  void update_accel_after_transform_change();

  std::unique_ptr<frd::Integrator> integrator;

  void upload_scene();
  void upload_transforms();
  void build_accel();
  frd::SceneView view(const float3& bg_color) const;
};

}  // namespace fredholm
