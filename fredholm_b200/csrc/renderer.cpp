// Host side of fredholm::Renderer: scene upload, light extraction, sky setup,
// acceleration-structure build and the render call.
//
// Upload logic follows Renderer::load_scene of the reference (renderer.h:354-432):
// flat arrays -> device buffers, one texture header per texture, the area-light
// list = every face whose material has emission (renderer.h:388-402), one
// object-to-world / world-to-object 3x4 pair per sub-mesh (renderer.h:404-421).
#include "fredholm/renderer.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <thread>
#include <vector>

#include "lobes.h"
#include "renderer_impl.h"

namespace fredholm
{

namespace
{

// ---- Hosek-Wilkie RGB datasets (data: tables/hosek_rgb_*.inc) ----
const float kHosekConfig[3][1080] = {{
#include "tables/hosek_rgb_config_r.inc"
                                     },
                                     {
#include "tables/hosek_rgb_config_g.inc"
                                     },
                                     {
#include "tables/hosek_rgb_config_b.inc"
                                     }};
const float kHosekRadiance[3][120] = {{
#include "tables/hosek_rgb_radiance_r.inc"
                                      },
                                      {
#include "tables/hosek_rgb_radiance_g.inc"
                                      },
                                      {
#include "tables/hosek_rgb_radiance_b.inc"
                                      }};

// quintic Bezier weights in the solar-elevation parameter
void elevation_weights(float elevation, float w[6])
{
  const float x = std::pow(elevation / (3.141592653589793f / 2.0f), (1.0f / 3.0f));
  w[0] = std::pow(1.0f - x, 5.0f);
  w[1] = 5.0f * std::pow(1.0f - x, 4.0f) * x;
  w[2] = 10.0f * std::pow(1.0f - x, 3.0f) * std::pow(x, 2.0f);
  w[3] = 10.0f * std::pow(1.0f - x, 2.0f) * std::pow(x, 3.0f);
  w[4] = 5.0f * (1.0f - x) * std::pow(x, 4.0f);
  w[5] = std::pow(x, 5.0f);
}

Matrix3x4 rows_of(const mat4& m)
{
  return make_mat3x4(make_float4(m[0][0], m[1][0], m[2][0], m[3][0]), make_float4(m[0][1], m[1][1], m[2][1], m[3][1]),
                     make_float4(m[0][2], m[1][2], m[2][2], m[3][2]));
}

float lum(const float3& c) { return 0.2126729f * c.x + 0.7151522f * c.y + 0.0721750f * c.z; }

// Smallest shade-kernel variant whose lobe set covers everything this material can
// ever evaluate.  A lobe may only be dropped when its run-time gate in the BSDF is
// provably false (or its layer factor exactly zero) for the material's constants.
frd::ShadeClass classify_material(const Material& m)
{
  using namespace frd;
  const bool textured = m.base_color_texture_id >= 0 || m.specular_color_texture_id >= 0 ||
                        m.specular_roughness_texture_id >= 0 || m.metalness_texture_id >= 0 ||
                        m.metallic_roughness_texture_id >= 0 || m.coat_texture_id >= 0 ||
                        m.coat_roughness_texture_id >= 0 || m.emission_texture_id >= 0 ||
                        m.heightmap_texture_id >= 0 || m.normalmap_texture_id >= 0 || m.alpha_texture_id >= 0;
  if (textured) return CLS_GENERIC_TEX;
  uint32_t need = M_DIFFUSE_R;
  if (m.coat > 0.0f) need |= M_COAT;  // coat colour is always white on the device
  if (m.metalness > 0.0f) need |= M_METAL;
  const bool full_metal = m.metalness == 1.0f;  // everything under the metal layer is scaled by exactly 0
  if (!full_metal) {
    if (m.specular * lum(m.specular_color) > 0.0f) need |= M_SPECULAR;
    if (m.transmission > 0.0f) need |= M_TRANSMISSION;
    if (m.sheen * lum(m.sheen_color) != 0.0f) need |= M_SHEEN;
    if (m.subsurface * m.thin_walled > 0.0f) need |= M_DIFFUSE_T;
  }
  const struct {
    ShadeClass cls;
    uint32_t mask;
  } variants[] = {{CLS_DIFFUSE, M_DIFFUSE_R},
                  {CLS_PLASTIC, M_SPECULAR | M_DIFFUSE_R},
                  {CLS_METAL, M_METAL | M_DIFFUSE_R},
                  {CLS_COATED, M_COAT | M_SPECULAR | M_DIFFUSE_R},
                  {CLS_GLASS, M_SPECULAR | M_TRANSMISSION | M_DIFFUSE_R},
                  {CLS_SHEEN, M_SHEEN | M_DIFFUSE_R}};
  for (const auto& v : variants)
    if ((need & ~v.mask) == 0) return v.cls;
  return CLS_GENERIC;
}

float3 normalized(const float3& v)
{
  const float inv = 1.0f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  return make_float3(v.x * inv, v.y * inv, v.z * inv);
}

}  // namespace

// Interpolation of the Hosek coefficient tables: bilinear in (albedo, turbidity),
// quintic Bezier in cbrt(elevation / 90deg)  (reference arhosek.h:145-323).
void arhosek_rgb_cook(float turbidity, float albedo, float elevation, float out30[30])
{
  const int it = static_cast<int>(turbidity);
  const float rem = turbidity - static_cast<float>(it);
  float w[6];
  elevation_weights(elevation, w);
  for (int ch = 0; ch < 3; ++ch) {
    // (albedo index, turbidity slot, blend factor) in the reference's summation order
    const struct {
      int alb, turb;
      float k;
    } corners[4] = {{0, it - 1, (1.0f - albedo) * (1.0f - rem)},
                    {1, it - 1, albedo * (1.0f - rem)},
                    {0, it, (1.0f - albedo) * rem},
                    {1, it, albedo * rem}};
    const int n_corners = it == 10 ? 2 : 4;
    for (int i = 0; i < 9; ++i) {
      float acc = 0.0f;
      for (int c = 0; c < n_corners; ++c) {
        const float* e = kHosekConfig[ch] + 9 * 6 * 10 * corners[c].alb + 9 * 6 * corners[c].turb;
        const float bez = w[0] * e[i] + w[1] * e[i + 9] + w[2] * e[i + 18] + w[3] * e[i + 27] + w[4] * e[i + 36] +
                          w[5] * e[i + 45];
        acc = c == 0 ? corners[c].k * bez : acc + corners[c].k * bez;
      }
      out30[9 * ch + i] = acc;
    }
    float acc = 0.0f;
    for (int c = 0; c < n_corners; ++c) {
      const float* e = kHosekRadiance[ch] + 6 * 10 * corners[c].alb + 6 * corners[c].turb;
      const float bez = w[0] * e[0] + w[1] * e[1] + w[2] * e[2] + w[3] * e[3] + w[4] * e[4] + w[5] * e[5];
      acc = c == 0 ? corners[c].k * bez : acc + corners[c].k * bez;
    }
    out30[27 + ch] = acc;
  }
}

// ---------------------------------------------------------------------------------------
void Renderer::Impl::upload_transforms()
{
  std::vector<Matrix3x4> o2w(scene.m_transforms.size()), w2o(scene.m_transforms.size());
  for (size_t i = 0; i < scene.m_transforms.size(); ++i) {
    o2w[i] = rows_of(scene.m_transforms[i]);
    w2o[i] = rows_of(inverse(scene.m_transforms[i]));
  }
  d_o2w.upload(o2w, stream);
  d_w2o.upload(w2o, stream);
  FR_CUDA_CHECK(cudaStreamSynchronize(stream));
  accel_valid = false;
}

void Renderer::Impl::upload_scene()
{
  const Scene& s = scene;
  s.validate();
  FR_CUDA_CHECK(cudaStreamSynchronize(stream));
  d_vertices.upload(s.m_vertices, stream);
  d_normals.upload(s.m_normals, stream);
  d_texcoords.upload(s.m_texcoords, stream);
  d_indices.upload(s.m_indices, stream);
  d_material_ids.upload(s.m_material_ids, stream);
  d_materials.upload(s.m_materials, stream);
  d_submesh_offsets.upload(s.m_submesh_offsets, stream);

  // face -> sub-mesh (== instance, == transform) and the alpha-test flag
  std::vector<uint32_t> face_submesh(s.m_indices.size(), 0), face_flags(s.m_indices.size(), 0);
  for (size_t sm = 0; sm < s.m_submesh_offsets.size(); ++sm)
    for (uint32_t f = 0; f < s.m_submesh_n_faces[sm]; ++f) face_submesh[s.m_submesh_offsets[sm] + f] = (uint32_t)sm;
  for (size_t f = 0; f < s.m_indices.size(); ++f) {
    const uint32_t mid = s.m_material_ids[f];
    if (mid >= s.m_materials.size()) throw std::runtime_error("invalid scene: face without a valid material");
    const Material& m = s.m_materials[mid];
    face_flags[f] = (m.base_color_texture_id >= 0 || m.alpha_texture_id >= 0) ? 1u : 0u;
  }
  d_face_submesh.upload(face_submesh, stream);
  d_face_flags.upload(face_flags, stream);
  // material class per face (shade-stage queue selection)
  std::vector<uint8_t> material_class(s.m_materials.size()), face_class(s.m_indices.size());
  class_mask = 0;
  for (size_t i = 0; i < s.m_materials.size(); ++i) material_class[i] = (uint8_t)classify_material(s.m_materials[i]);
  for (size_t f = 0; f < s.m_indices.size(); ++f) {
    face_class[f] = material_class[s.m_material_ids[f]];
    class_mask |= 1u << face_class[f];
  }
  d_face_class.upload(face_class, stream);

  // textures
  d_texture_data.clear();
  d_texture_data.resize(s.m_textures.size());
  std::vector<frd::TexView> views(s.m_textures.size());
  for (size_t i = 0; i < s.m_textures.size(); ++i) {
    const Texture& t = s.m_textures[i];
    d_texture_data[i].upload(t.m_data, stream);
    views[i].texels = d_texture_data[i].get();
    views[i].width = t.m_width;
    views[i].height = t.m_height;
    views[i].srgb = t.m_texture_type == TextureType::COLOR ? 1u : 0u;
    views[i].pad_ = 0;
  }
  if (!views.empty()) d_textures.upload(views, stream);
  if (d_srgb_lut.size() == 0) {
    std::vector<float> lut(256);
    for (int c = 0; c < 256; ++c) {
      const double v = c / 255.0;
      lut[c] = static_cast<float>(v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4));
    }
    d_srgb_lut.upload(lut, stream);
  }

  // emissive faces become area lights
  std::vector<AreaLight> lights;
  for (size_t f = 0; f < s.m_material_ids.size(); ++f) {
    const Material& m = s.m_materials[s.m_material_ids[f]];
    if (m.emission_color.x > 0 || m.emission_color.y > 0 || m.emission_color.z > 0 || m.emission_texture_id != -1) {
      AreaLight l;
      l.indices = s.m_indices[f];
      l.material_id = s.m_material_ids[f];
      l.instance_idx = s.m_instance_ids[f];
      lights.push_back(l);
    }
  }
  n_lights = (uint32_t)lights.size();
  if (n_lights) d_lights.upload(lights, stream);
  find_distinct_meshes();
  upload_transforms();
}

// Sub-meshes whose triangles have identical object-space positions, in the same order, are copies of one mesh
// (the reference's glTF path makes one sub-mesh with its own vertex copy per node, scene.cpp:730-822, and so does
// every instancing exporter): they can share one object-space tree.  128-bit content hash per sub-mesh, computed
// on all host threads; candidates are grouped by (face count, hash).
void Renderer::Impl::find_distinct_meshes()
{
  const Scene& s = scene;
  const size_t n_sm = s.m_submesh_offsets.size();
  struct Key {
    uint64_t a, b;
    uint32_t n;
    bool operator<(const Key& o) const { return n != o.n ? n < o.n : (a != o.a ? a < o.a : b < o.b); }
  };
  std::vector<Key> keys(n_sm);
  auto hash_range = [&](size_t begin, size_t end) {
    for (size_t sm = begin; sm < end; ++sm) {
      uint64_t a = 0x9e3779b97f4a7c15ull, b = 0xc2b2ae3d27d4eb4full;
      const uint32_t off = s.m_submesh_offsets[sm], nf = s.m_submesh_n_faces[sm];
      for (uint32_t f = 0; f < nf; ++f) {
        const uint3 idx = s.m_indices[off + f];
        const uint32_t vid[3] = {idx.x, idx.y, idx.z};
        for (int k = 0; k < 3; ++k) {
          uint32_t w[3];
          std::memcpy(w, &s.m_vertices[vid[k]], 12);
          for (int c = 0; c < 3; ++c) {
            a = (a ^ w[c]) * 0x100000001b3ull;
            a ^= a >> 29;
            b = (b + w[c]) * 0xff51afd7ed558ccdull;
            b ^= b >> 32;
          }
        }
      }
      keys[sm] = Key{a, b, nf};
    }
  };
  const unsigned n_threads = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)n_sm));
  if (n_threads <= 1 || s.m_indices.size() < (1u << 16)) {
    hash_range(0, n_sm);
  } else {
    // contiguous ranges with about the same number of faces each
    std::vector<std::thread> th;
    const size_t total = s.m_indices.size();
    size_t begin = 0;
    for (unsigned t = 0; t < n_threads && begin < n_sm; ++t) {
      const size_t target = total * (t + 1) / n_threads;
      size_t end = begin;
      while (end < n_sm && (end == begin || (size_t)s.m_submesh_offsets[end] + s.m_submesh_n_faces[end] <= target)) ++end;
      if (t + 1 == n_threads) end = n_sm;
      th.emplace_back(hash_range, begin, end);
      begin = end;
    }
    for (auto& t : th) t.join();
  }
  std::map<Key, uint32_t> first;
  mesh_of_submesh.assign(n_sm, 0);
  mesh_representative.clear();
  for (size_t sm = 0; sm < n_sm; ++sm) {
    auto it = first.find(keys[sm]);
    if (it == first.end()) {
      it = first.emplace(keys[sm], (uint32_t)mesh_representative.size()).first;
      mesh_representative.push_back((uint32_t)sm);
    }
    mesh_of_submesh[sm] = it->second;
  }
}

bool Renderer::Impl::want_two_level() const
{
  if (const char* e = getenv("FRD_ACCEL")) {  // experiments / tests: flat | two_level
    if (std::strcmp(e, "flat") == 0) return false;
    if (std::strcmp(e, "two_level") == 0) return true;
  }
  if (accel_mode == AccelMode::FLAT) return false;
  for (uint32_t nf : scene.m_submesh_n_faces)
    if (nf == 0) return false;  // an empty sub-mesh has no box to place
  if (accel_mode == AccelMode::TWO_LEVEL) return true;
  // AUTO: the flat world-space tree is the bit-exact and the faster one, so it stays the choice while it fits;
  // a scene whose flat tree (128 B per triangle incl. nodes, ~3x that while building) would take more than a
  // third of the device's memory goes two-level if at least half of its triangles are copies of another sub-mesh
  size_t stored = 0;
  for (uint32_t rep : mesh_representative) stored += scene.m_submesh_n_faces[rep];
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
  const size_t flat_build_bytes = scene.m_indices.size() * size_t(3 * 128);
  return flat_build_bytes > total_b / 3 && 2 * stored <= scene.m_indices.size();
}

void Renderer::Impl::refresh_accel_info(float build_ms)
{
  accel_info = AccelInfo();
  accel_info.build_ms = build_ms;
  accel_info.n_faces = (uint32_t)scene.m_indices.size();
  accel_info.two_level = two_level;
  if (two_level) {
    accel_info.n_nodes = bvh2.n_nodes;
    accel_info.depth = bvh2.depth;
    accel_info.bytes = bvh2.bytes();
    accel_info.n_instances = bvh2.n_instances;
    accel_info.n_meshes = bvh2.n_meshes;
    accel_info.n_stored_faces = bvh2.n_blas_faces;
    accel_info.tlas_update_ms = bvh2.tlas_ms;
    accel_info.tlas_refitted = bvh2.last_update_was_refit;
  } else {
    accel_info.n_nodes = bvh.n_nodes;
    accel_info.depth = bvh.depth;
    accel_info.bytes = size_t(bvh.n_nodes) * sizeof(frd::Node8) + size_t(bvh.n_faces) * 3 * sizeof(float4);
    accel_info.n_instances = (uint32_t)scene.m_submesh_offsets.size();
    accel_info.n_meshes = accel_info.n_instances;
    accel_info.n_stored_faces = bvh.n_faces;
  }
}

void Renderer::Impl::build_accel()
{
  cudaEvent_t e0, e1;
  FR_CUDA_CHECK(cudaEventCreate(&e0));
  FR_CUDA_CHECK(cudaEventCreate(&e1));
  FR_CUDA_CHECK(cudaEventRecord(e0, stream));
  two_level = want_two_level();
  if (two_level) {
    // alpha-test flag of a shared triangle: set if ANY instance of the mesh has an alpha-tested material there
    // (the any-hit functor looks at the real material of the face that was hit)
    std::vector<std::vector<uint32_t>> flags(mesh_representative.size());
    bool any_flag = false;
    for (size_t sm = 0; sm < scene.m_submesh_offsets.size(); ++sm) {
      const uint32_t off = scene.m_submesh_offsets[sm], nf = scene.m_submesh_n_faces[sm];
      for (uint32_t f = 0; f < nf; ++f) {
        const Material& m = scene.m_materials[scene.m_material_ids[off + f]];
        if (m.base_color_texture_id >= 0 || m.alpha_texture_id >= 0) {
          std::vector<uint32_t>& v = flags[mesh_of_submesh[sm]];
          if (v.empty()) v.assign(nf, 0u);
          v[f] = 1u;
          any_flag = true;
        }
      }
    }
    frd::build_two_level(stream, d_vertices.get(), d_indices.get(), scene.m_submesh_offsets, scene.m_submesh_n_faces,
                         mesh_of_submesh, mesh_representative, any_flag ? &flags : nullptr, d_o2w.get(), bvh2);
    bvh = frd::DeviceBvh();  // the flat tree of an earlier build is not kept
  } else {
    frd::build_bvh(stream, d_vertices.get(), d_indices.get(), d_face_submesh.get(), d_face_flags.get(), d_o2w.get(),
                   (uint32_t)scene.m_indices.size(), bvh);
    bvh2 = frd::TwoLevelBvh();
  }
  FR_CUDA_CHECK(cudaEventRecord(e1, stream));
  FR_CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0.0f;
  FR_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  refresh_accel_info(ms);
  accel_valid = true;
}

// transforms changed: the instance tree follows (two-level), or the whole world-space tree is rebuilt (flat)
void Renderer::Impl::update_accel_after_transform_change()
{
  upload_transforms();
  if (two_level && bvh2.n_instances == scene.m_transforms.size() && bvh2.n_instances > 0) {
    // refit the instance tree in place (one launch); rebuild it when the refit reports that the instances have
    // spread out so much that the old topology is no longer a good one (or FRD_TLAS_REBUILD=1 asks for it)
    const bool force_rebuild = getenv("FRD_TLAS_REBUILD") != nullptr;
    if (force_rebuild || !frd::refit_tlas(stream, d_o2w.get(), bvh2)) frd::update_tlas(stream, d_o2w.get(), bvh2);
    refresh_accel_info(accel_info.build_ms);
    accel_valid = true;
  } else {
    build_accel();
  }
}

frd::SceneView Renderer::Impl::view(const float3& bg_color) const
{
  frd::SceneView v{};
  v.vertices = d_vertices.get();
  v.normals = d_normals.get();
  v.texcoords = d_texcoords.get();
  v.indices = d_indices.get();
  v.material_ids = d_material_ids.get();
  v.face_submesh = d_face_submesh.get();
  v.face_class = d_face_class.get();
  v.materials = d_materials.get();
  v.textures = d_textures.get();
  v.srgb_lut = d_srgb_lut.get();
  v.o2w = d_o2w.get();
  v.w2o = d_w2o.get();
  v.lights = d_lights.get();
  v.n_lights = n_lights;
  if (two_level) {
    v.bvh = bvh2.view(d_w2o.get());
    v.bounds_lo = make_float3(bvh2.bounds_lo[0], bvh2.bounds_lo[1], bvh2.bounds_lo[2]);
    v.bounds_hi = make_float3(bvh2.bounds_hi[0], bvh2.bounds_hi[1], bvh2.bounds_hi[2]);
  } else {
    v.bvh = bvh.view();
    v.bounds_lo = make_float3(bvh.bounds_lo[0], bvh.bounds_lo[1], bvh.bounds_lo[2]);
    v.bounds_hi = make_float3(bvh.bounds_hi[0], bvh.bounds_hi[1], bvh.bounds_hi[2]);
  }
  v.has_dir_light = has_dir_light ? 1 : 0;
  v.dir_light = dir_light;
  if (has_dir_light) {
    // disk of directions at distance 1e9 (pt.cu:324-342)
    const float3 n = dir_light.dir;
    const float sign = std::copysign(1.0f, n.z);
    const float a = -1.0f / (sign + n.z);
    const float b = n.x * n.y * a;
    v.dir_t = make_float3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    v.dir_b = make_float3(b, sign + n.y * n.y * a, -n.y);
    const float half_angle_rad = (0.5f * dir_light.angle) * 3.14159265358979323846f / 180.0f;
    v.dir_disk_radius = 1e9f * std::tan(half_angle_rad);
  }
  v.sky_mode = d_ibl.size() ? frd::SKY_IBL : (has_hosek ? frd::SKY_HOSEK : frd::SKY_CONSTANT);
  v.sky_intensity = sky_intensity;
  v.bg_color = bg_color;
  v.sun_dir = sun_direction;
  v.hosek = hosek;
  v.ibl_texels = d_ibl.get();
  v.ibl_width = ibl_w;
  v.ibl_height = ibl_h;
  return v;
}

// ---------------------------------------------------------------------------------------
Renderer::Renderer(int cuda_device) : m_impl(new Impl())
{
  m_impl->device = cuda_device;
  FR_CUDA_CHECK(cudaSetDevice(cuda_device));
  FR_CUDA_CHECK(cudaFree(nullptr));
  FR_CUDA_CHECK(cudaStreamCreate(&m_impl->stream));
  m_impl->integrator = std::make_unique<frd::Integrator>(m_impl->stream);
}

Renderer::~Renderer() noexcept(false)
{
  if (m_impl && m_impl->stream) {
    cudaStreamSynchronize(m_impl->stream);
    m_impl->integrator.reset();
    cudaStreamDestroy(m_impl->stream);
    m_impl->stream = nullptr;
  }
}

void Renderer::create_module(const std::filesystem::path&) {}
void Renderer::create_program_group() {}
void Renderer::create_pipeline() {}
void Renderer::create_sbt() {}

void Renderer::load_scene(const std::filesystem::path& filepath, bool clear)
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  m_impl->scene.load_model(filepath, clear);
  if (!m_impl->scene.is_valid()) throw std::runtime_error("invalid scene");
  m_impl->upload_scene();
}

void Renderer::set_scene(const Scene& scene)
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  m_impl->scene = scene;
  if (!m_impl->scene.is_valid()) throw std::runtime_error("invalid scene");
  m_impl->upload_scene();
}

const Scene& Renderer::get_scene() const { return m_impl->scene; }

void Renderer::build_gas()
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  m_impl->build_accel();
}

void Renderer::build_ias()
{
  // instances are flattened into the one world-space tree; nothing to do unless
  // the tree is stale (transforms changed)
  if (!m_impl->accel_valid) build_gas();
}

void Renderer::set_time(float time)
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  // the reference re-uploads every transform and rebuilds its IAS on each call
  // (renderer.h:614-640); here the world-space tree only depends on the sub-mesh
  // transforms, so a time step that moves nothing but the camera keeps the tree
  const std::vector<mat4> before = m_impl->scene.m_transforms;
  m_impl->scene.update_animation(time);
  const std::vector<mat4>& after = m_impl->scene.m_transforms;
  const bool moved = before.size() != after.size() ||
                     (!after.empty() && std::memcmp(before.data(), after.data(), sizeof(mat4) * after.size()) != 0);
  if (!moved && m_impl->accel_valid) return;
  if (!m_impl->accel_valid) {
    m_impl->upload_transforms();
    m_impl->build_accel();
  } else {
    m_impl->update_accel_after_transform_change();
  }
}

void Renderer::set_accel_mode(AccelMode mode)
{
  if (mode != m_impl->accel_mode) m_impl->accel_valid = false;
  m_impl->accel_mode = mode;
}

void Renderer::set_transforms(const float* transforms16, uint32_t n_submeshes)
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  if (n_submeshes != m_impl->scene.m_transforms.size()) throw std::runtime_error("transform count mismatch");
  std::memcpy(m_impl->scene.m_transforms.data(), transforms16, sizeof(float) * 16 * n_submeshes);
  if (!m_impl->accel_valid) {
    m_impl->upload_transforms();
    m_impl->build_accel();
  } else {
    m_impl->update_accel_after_transform_change();
  }
}

void Renderer::set_directional_light(const float3& le, const float3& dir, float angle)
{
  m_impl->dir_light.le = le;
  m_impl->dir_light.dir = normalized(dir);
  m_impl->dir_light.angle = angle;
  m_impl->sun_direction = normalized(dir);
  m_impl->has_dir_light = true;
}
void Renderer::clear_directional_light() { m_impl->has_dir_light = false; }
void Renderer::set_sky_intensity(float sky_intensity) { m_impl->sky_intensity = sky_intensity; }

void Renderer::load_ibl(const std::filesystem::path& filepath)
{
  const FloatTexture ibl(filepath);
  set_ibl(ibl.m_data.data(), ibl.m_width, ibl.m_height);
}
void Renderer::set_ibl(const float4* texels, uint32_t width, uint32_t height)
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  FR_CUDA_CHECK(cudaStreamSynchronize(m_impl->stream));
  m_impl->d_ibl.alloc((size_t)width * height);
  FR_CUDA_CHECK(cudaMemcpy(m_impl->d_ibl.get(), texels, sizeof(float4) * width * height, cudaMemcpyHostToDevice));
  m_impl->ibl_w = width;
  m_impl->ibl_h = height;
}
void Renderer::clear_ibl()
{
  FR_CUDA_CHECK(cudaStreamSynchronize(m_impl->stream));
  m_impl->d_ibl.release();
  m_impl->ibl_w = m_impl->ibl_h = 0;
}

void Renderer::load_arhosek_sky(float turbidity, float albedo)
{
  // solar elevation from the current sun direction (renderer.h:592-603)
  const float cy = std::fmax(-1.0f, std::fmin(m_impl->sun_direction.y, 1.0f));
  float elevation = std::acos(cy);
  elevation = 0.5f * M_PI - elevation;
  float cooked[30];
  arhosek_rgb_cook(turbidity, albedo, elevation, cooked);
  for (int c = 0; c < 3; ++c) {
    for (int i = 0; i < 9; ++i) m_impl->hosek.cfg[c][i] = cooked[9 * c + i];
    m_impl->hosek.rad[c] = cooked[27 + c];
  }
  m_impl->has_hosek = true;
}
void Renderer::clear_arhosek_sky() { m_impl->has_hosek = false; }

void Renderer::set_resolution(uint32_t width, uint32_t height)
{
  m_impl->width = width;
  m_impl->height = height;
  init_render_states();
}
void Renderer::init_render_states() { m_impl->sample_count = 0; }

void Renderer::render(const Camera& camera, const float3& bg_color, const RenderLayer& render_layer,
                      uint32_t n_samples, uint32_t max_depth)
{
  render(camera_params(camera), bg_color, render_layer, n_samples, max_depth);
}

CameraParams Renderer::camera_params(const Camera& camera) const
{
  CameraParams cp;
  // a camera node of the scene overrides the application camera (renderer.h:671-685)
  cp.transform = rows_of(m_impl->scene.m_has_camera_transform ? m_impl->scene.m_camera_transform : camera.m_transform);
  cp.fov = camera.m_fov;
  cp.F = camera.m_F;
  cp.focus = camera.m_focus;
  return cp;
}

void Renderer::render(const CameraParams& camera, const float3& bg_color, const RenderLayer& render_layer,
                      uint32_t n_samples, uint32_t max_depth)
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  // A path draws up to four 1-D Sobol dimensions per bounce (light choice, two lobe choices, roulette) after
  // dimension 1 of the camera ray, and the table has 1024 dimensions (sobol.cu; the reference reads past it
  // beyond that, pt.cu:455-471 has no bound): 2 + 4 * max_depth <= 1024.
  if (max_depth > 255u) throw std::invalid_argument("render: max_depth must be <= 255 (1024 Sobol dimensions)");
  if (!m_impl->accel_valid) m_impl->build_accel();
  const frd::SceneView view = m_impl->view(bg_color);
  m_impl->integrator->render(view, camera, m_impl->width, m_impl->height, render_layer, m_impl->sample_count,
                             n_samples, max_depth, /*seed=*/1u,
                             m_impl->film_mode == FilmMode::MEAN ? frd::FILM_MEAN : frd::FILM_SUM,
                             m_impl->class_mask);
  m_impl->sample_count += n_samples;
}

void Renderer::wait_for_completion()
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  FR_CUDA_CHECK(cudaDeviceSynchronize());
  FR_CUDA_CHECK(cudaGetLastError());
}

void Renderer::set_sample_offset(uint32_t first_sample) { m_impl->sample_count = first_sample; }
uint32_t Renderer::get_sample_count() const { return m_impl->sample_count; }
void Renderer::set_film_mode(FilmMode mode) { m_impl->film_mode = mode; }
FilmMode Renderer::get_film_mode() const { return m_impl->film_mode; }
void Renderer::scale_layers(const RenderLayer& render_layer, float scale)
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  m_impl->integrator->scale_layers(render_layer, m_impl->width * m_impl->height, scale);
}
void Renderer::set_max_wave_paths(size_t n_paths) { m_impl->integrator->set_max_wave_paths(n_paths); }
size_t Renderer::get_wave_state_bytes() const { return m_impl->integrator->state_bytes(); }
void Renderer::set_wave_overlap(bool on) { m_impl->integrator->set_wave_overlap(on); }
void Renderer::set_wave_compaction(bool on, uint32_t depth) { m_impl->integrator->set_wave_compaction(on, depth); }
void Renderer::set_single_launch(bool on) { m_impl->integrator->set_single_launch(on); }
void Renderer::set_samples_per_warp(uint32_t spw) { m_impl->integrator->set_samples_per_warp(spw); }
void Renderer::set_traversal_counting(bool on) { frd::set_traversal_counting(on); }

void Renderer::set_stage_timing(bool on) { m_impl->integrator->set_stage_timing(on); }
void Renderer::get_stage_times(double ms[kStageCount], unsigned long long launches[kStageCount])
{
  static_assert(kStageCount == frd::STAGE_COUNT, "stage list out of sync");
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  const frd::StageTimes t = m_impl->integrator->stage_times();
  for (int i = 0; i < kStageCount; ++i) {
    ms[i] = t.ms[i];
    launches[i] = t.launches[i];
  }
}

RenderStatistics Renderer::get_statistics()
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  const frd::RenderStats s = m_impl->integrator->stats();
  RenderStatistics r;
  r.paths = s.paths;
  r.rays_radiance = s.rays_closest;
  r.rays_shadow = s.rays_shadow;
  r.rays_light = s.rays_light;
  r.kernel_launches = s.launches;
  r.rays_skipped = s.rays_skipped;
  for (int i = 0; i < 3; ++i) {
    r.nodes_visited[i] = s.nodes[i];
    r.tris_tested[i] = s.tris[i];
  }
  return r;
}
void Renderer::reset_statistics()
{
  FR_CUDA_CHECK(cudaSetDevice(m_impl->device));
  m_impl->integrator->reset_stats();
}
AccelInfo Renderer::get_accel_info() const { return m_impl->accel_info; }
cudaStream_t Renderer::get_stream() const { return m_impl->stream; }
int Renderer::get_device() const { return m_impl->device; }

}  // namespace fredholm
