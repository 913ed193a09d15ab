// Scene description consumed by the renderer: flat vertex / index / material
// arrays plus per-submesh transforms.
//
// Member names and meaning follow the reference's fredholm::Scene
// (fredholm/include/fredholm/scene.h:103-179) -- these arrays ARE the interface
// between loaders and the renderer.  The reference fills them with tinyobjloader
// / tinygltf / stb; this implementation carries its own parsers (scene.cpp).
#pragma once
#include <cstdint>
#include <filesystem>
#include <string>
#include <vector>

#include "fredholm/camera.h"
#include "fredholm/shared.h"

namespace fredholm
{

enum class TextureType { COLOR, NONCOLOR };

// 8-bit RGBA texture; row 0 is the BOTTOM row of the image file (the reference
// loads with a vertical flip, scene.cpp:15).  COLOR textures are sRGB-decoded on
// fetch, NONCOLOR are linear.
struct Texture {
  uint32_t m_width = 0;
  uint32_t m_height = 0;
  std::vector<uchar4> m_data;
  TextureType m_texture_type = TextureType::NONCOLOR;

  Texture() = default;
  Texture(uint32_t width, uint32_t height, const uchar4* data, const TextureType& texture_type)
      : m_width(width), m_height(height), m_data(data, data + size_t(width) * height), m_texture_type(texture_type)
  {
  }
  Texture(const std::filesystem::path& filepath, const TextureType& texture_type);
};

// float RGBA lat-long environment image; row 0 is the TOP row (no flip, scene.cpp:44)
struct FloatTexture {
  uint32_t m_width = 0;
  uint32_t m_height = 0;
  std::vector<float4> m_data;

  FloatTexture() = default;
  FloatTexture(const std::filesystem::path& filepath);
};

struct Node {
  int idx = -1;  // node index in the source file
  std::vector<Node> children;
  mat4 transform;
  int camera_id = -1;
  int submesh_id = -1;
};

struct quat {
  float x = 0, y = 0, z = 0, w = 1;
};

struct Animation {
  int node_idx = -1;   // target node (index in the source file)
  int root_slot = -1;  // position of that node in Scene::m_nodes (only root nodes can be driven)

  std::vector<float> translation_input;
  std::vector<vec3> translation_output;
  std::vector<float> rotation_input;
  std::vector<quat> rotation_output;
  std::vector<float> scale_input;
  std::vector<vec3> scale_output;
};

struct Scene {
  bool m_has_camera_transform = false;
  mat4 m_camera_transform = {};

  // vertex data (one index addresses position, normal and texcoord)
  std::vector<float3> m_vertices = {};
  std::vector<uint3> m_indices = {};
  std::vector<float2> m_texcoords = {};
  std::vector<float3> m_normals = {};
  std::vector<float3> m_tangents = {};

  // per-face material id
  std::vector<unsigned int> m_material_ids = {};
  std::vector<Material> m_materials;
  std::vector<Texture> m_textures;

  // offset / face count of each sub-mesh in the index buffer; sub-mesh i is
  // instance i and uses m_transforms[i]
  std::vector<unsigned int> m_submesh_offsets = {};
  std::vector<unsigned int> m_submesh_n_faces = {};

  // per-face instance id (used by the area-light list only)
  std::vector<unsigned int> m_instance_ids = {};

  // per-instance object-to-world transform
  std::vector<mat4> m_transforms = {};

  std::vector<Node> m_nodes = {};
  std::vector<Animation> m_animations = {};

  Scene() = default;

  bool is_valid() const;
  // extension: full consistency check of the arrays the renderer indexes on the device (array lengths, vertex /
  // material / instance / texture indices, sub-meshes tiling the face range); throws std::runtime_error
  // "invalid scene: ..." -- Renderer::load_scene / set_scene call it before uploading
  void validate() const;
  void clear();

  // .obj or .gltf by extension; throws std::runtime_error otherwise
  void load_model(const std::filesystem::path& filepath, bool do_clear);
  void load_obj(const std::filesystem::path& filepath);
  void load_gltf(const std::filesystem::path& filepath);

  void update_transform();
  void update_animation(float time);
};

}  // namespace fredholm
