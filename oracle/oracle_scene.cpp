// =============================================================================
// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_host.cpp).
//
// File loading for the oracle goes through the reference's OWN loader,
// /root/reference/fredholm/src/scene.cpp (tinyobjloader / tinygltf / stb),
// compiled unchanged by the Makefile.  This file only copies the resulting
// fredholm::Scene members (scene.h:107-130) into the oracle state and exposes
// them to the tests, so that the product's own .obj/.gltf parsers can be
// checked array-for-array against the reference loader.
// =============================================================================
#include <cstring>
#include <string>
#include <vector>

#include "fredholm/camera.h"
#include "fredholm/scene.h"

extern "C" {
void orc_reset();
void orc_set_scene(const float* vertices, const float* normals,
                   const float* texcoords, uint n_vertices, const uint* indices,
                   const uint* material_ids, const uint* instance_ids,
                   uint n_faces, const void* materials, uint n_materials,
                   const uint* submesh_offsets, const uint* submesh_n_faces,
                   const float* transforms, uint n_submeshes);
void orc_set_transforms(const float* transforms, uint n_submeshes);
int orc_add_texture(const unsigned char* rgba8, int w, int h, int is_color);
}

namespace
{
fredholm::Scene g_scene;
}

extern "C" {

// Scene::load_model (scene.cpp:103-117) then the uploads of Renderer::load_scene.
// Returns 0 on success, -1 on failure (message in orc_last_error()).
static std::string g_err;
const char* orc_last_error() { return g_err.c_str(); }

int orc_load_scene(const char* path, int do_clear)
{
  try {
    g_scene.load_model(path, do_clear != 0);
    if (!g_scene.is_valid()) {
      g_err = "invalid scene";
      return -1;
    }
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
  orc_reset();
  orc_set_scene(reinterpret_cast<const float*>(g_scene.m_vertices.data()),
                reinterpret_cast<const float*>(g_scene.m_normals.data()),
                reinterpret_cast<const float*>(g_scene.m_texcoords.data()),
                g_scene.m_vertices.size(),
                reinterpret_cast<const uint*>(g_scene.m_indices.data()),
                g_scene.m_material_ids.data(), g_scene.m_instance_ids.data(),
                g_scene.m_indices.size(), g_scene.m_materials.data(),
                g_scene.m_materials.size(), g_scene.m_submesh_offsets.data(),
                g_scene.m_submesh_n_faces.data(),
                reinterpret_cast<const float*>(g_scene.m_transforms.data()),
                g_scene.m_submesh_offsets.size());
  for (const auto& t : g_scene.m_textures) {
    orc_add_texture(reinterpret_cast<const unsigned char*>(t.m_data.data()),
                    t.m_width, t.m_height,
                    t.m_texture_type == fredholm::TextureType::COLOR);
  }
  return 0;
}

// Renderer::set_time (renderer.h:614-640)
void orc_set_time(float time)
{
  g_scene.update_animation(time);
  orc_set_transforms(reinterpret_cast<const float*>(g_scene.m_transforms.data()),
                     g_scene.m_transforms.size());
}

// sizes: n_vertices, n_faces, n_materials, n_textures, n_submeshes,
// has_camera_transform
void orc_scene_sizes(uint* out6)
{
  out6[0] = g_scene.m_vertices.size();
  out6[1] = g_scene.m_indices.size();
  out6[2] = g_scene.m_materials.size();
  out6[3] = g_scene.m_textures.size();
  out6[4] = g_scene.m_submesh_offsets.size();
  out6[5] = g_scene.m_has_camera_transform ? 1 : 0;
}

void orc_scene_copy(float* vertices, float* normals, float* texcoords,
                    uint* indices, uint* material_ids, uint* instance_ids,
                    void* materials, uint* submesh_offsets,
                    uint* submesh_n_faces, float* transforms,
                    float* camera_transform)
{
  const auto& s = g_scene;
  std::memcpy(vertices, s.m_vertices.data(), sizeof(float3) * s.m_vertices.size());
  std::memcpy(normals, s.m_normals.data(), sizeof(float3) * s.m_normals.size());
  std::memcpy(texcoords, s.m_texcoords.data(), sizeof(float2) * s.m_texcoords.size());
  std::memcpy(indices, s.m_indices.data(), sizeof(uint3) * s.m_indices.size());
  std::memcpy(material_ids, s.m_material_ids.data(), 4 * s.m_material_ids.size());
  std::memcpy(instance_ids, s.m_instance_ids.data(), 4 * s.m_instance_ids.size());
  std::memcpy(materials, s.m_materials.data(),
              sizeof(fredholm::Material) * s.m_materials.size());
  std::memcpy(submesh_offsets, s.m_submesh_offsets.data(), 4 * s.m_submesh_offsets.size());
  std::memcpy(submesh_n_faces, s.m_submesh_n_faces.data(), 4 * s.m_submesh_n_faces.size());
  std::memcpy(transforms, s.m_transforms.data(), 64 * s.m_transforms.size());
  std::memcpy(camera_transform, &s.m_camera_transform, 64);
}

void orc_scene_texture_info(uint i, uint* w, uint* h, uint* is_color)
{
  *w = g_scene.m_textures[i].m_width;
  *h = g_scene.m_textures[i].m_height;
  *is_color = g_scene.m_textures[i].m_texture_type == fredholm::TextureType::COLOR;
}
void orc_scene_texture_copy(uint i, unsigned char* rgba8)
{
  const auto& t = g_scene.m_textures[i];
  std::memcpy(rgba8, t.m_data.data(), 4 * (size_t)t.m_width * t.m_height);
}

// The reference's image decoding on its own (Texture / FloatTexture constructors,
// scene.cpp:7-67 -> stb_image): 8-bit RGBA, bottom row first; float RGBA, top row first.
static std::vector<uchar4> g_tex;
static std::vector<float4> g_ftex;
int orc_image8_load(const char* path, uint* w, uint* h)
{
  try {
    const fredholm::Texture t(path, fredholm::TextureType::NONCOLOR);
    g_tex = t.m_data;
    *w = t.m_width;
    *h = t.m_height;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
  return 0;
}
void orc_image8_copy(unsigned char* rgba8) { std::memcpy(rgba8, g_tex.data(), 4 * g_tex.size()); }
int orc_imagef_load(const char* path, uint* w, uint* h)
{
  try {
    const fredholm::FloatTexture t(path);
    g_ftex = t.m_data;
    *w = t.m_width;
    *h = t.m_height;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
  return 0;
}
void orc_imagef_copy(float* rgba32f) { std::memcpy(rgba32f, g_ftex.data(), 16 * g_ftex.size()); }

// Camera (camera.h:51-69): camera-to-world 3x4 (row-major, as Renderer::render
// packs it, renderer.h:678-684) for a camera at `origin` looking down -z.
void orc_camera_transform(const float* origin, float* out12)
{
  const fredholm::Camera cam(make_float3(origin[0], origin[1], origin[2]));
  const glm::mat4& m = cam.m_transform;
  const float t[12] = {m[0][0], m[1][0], m[2][0], m[3][0], m[0][1], m[1][1],
                       m[2][1], m[3][1], m[0][2], m[1][2], m[2][2], m[3][2]};
  std::memcpy(out12, t, sizeof(t));
}

// Camera::lookAround + move (camera.h:85-135) for parity of the host mirror
void orc_camera_walk(const float* origin, float d_phi, float d_theta, int movement,
                     float dt, float* out12)
{
  fredholm::Camera cam(make_float3(origin[0], origin[1], origin[2]));
  cam.lookAround(d_phi, d_theta);
  cam.move(static_cast<fredholm::CameraMovement>(movement), dt);
  const glm::mat4& m = cam.m_transform;
  const float t[12] = {m[0][0], m[1][0], m[2][0], m[3][0], m[0][1], m[1][1],
                       m[2][1], m[3][1], m[0][2], m[1][2], m[2][2], m[3][2]};
  std::memcpy(out12, t, sizeof(t));
}

}  // extern "C"
