// extern "C" facade over fredholm::Renderer (see include/fredholm_b200.h).
#include "fredholm_b200.h"
#include "image_codec.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "fredholm/batch.h"
#include "fredholm/denoiser.h"
#include "fredholm/multi_gpu.h"
#include "kernels/post-process.h"
#include "renderer_impl.h"

using namespace fredholm;

struct fr_renderer {
  std::unique_ptr<Renderer> owned;  // null for the per-rank views of an fr_multi
  Renderer& renderer;
  std::unique_ptr<ShardedRenderer> rank;  // fr_comm_init: this renderer is one rank of a multi-GPU world
  std::vector<Texture> staged_textures;
  // device layers owned by fr_render_frame_host
  frd::DevBuf<float4> h_beauty, h_position, h_normal, h_texcoord, h_albedo;
  frd::DevBuf<float> h_depth;
  explicit fr_renderer(int dev) : owned(new Renderer(dev)), renderer(*owned) {}
  explicit fr_renderer(Renderer& external) : renderer(external) {}
};

struct fr_scene {
  Scene scene;
};

namespace
{
thread_local std::string g_error;

template <typename F>
int guarded(F&& f)
{
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
  } catch (...) {
    g_error = "unknown error";
  }
  return -1;
}

CameraParams make_camera(const float* t12, float fov, float F, float focus)
{
  CameraParams c;
  std::memcpy(&c.transform, t12, sizeof(float) * 12);
  c.fov = fov;
  c.F = F;
  c.focus = focus;
  return c;
}

RenderLayer make_layers(const fr_layers* l)
{
  RenderLayer r;
  r.beauty = static_cast<float4*>(l->beauty);
  r.position = static_cast<float4*>(l->position);
  r.depth = static_cast<float*>(l->depth);
  r.normal = static_cast<float4*>(l->normal);
  r.texcoord = static_cast<float4*>(l->texcoord);
  r.albedo = static_cast<float4*>(l->albedo);
  return r;
}

void camera_rows(const Camera& cam, float* out12) { cam.to_rows(out12); }

void scene_sizes(const Scene& s, uint32_t* out6)
{
  out6[0] = (uint32_t)s.m_vertices.size();
  out6[1] = (uint32_t)s.m_indices.size();
  out6[2] = (uint32_t)s.m_materials.size();
  out6[3] = (uint32_t)s.m_textures.size();
  out6[4] = (uint32_t)s.m_submesh_offsets.size();
  out6[5] = s.m_has_camera_transform ? 1u : 0u;
}

void scene_arrays(const Scene& s, float* vertices, float* normals, float* texcoords, uint32_t* indices,
                  uint32_t* material_ids, uint32_t* instance_ids, void* materials, uint32_t* submesh_offsets,
                  uint32_t* submesh_n_faces, float* transforms, float* camera_transform16)
{
  std::memcpy(vertices, s.m_vertices.data(), sizeof(float3) * s.m_vertices.size());
  std::memcpy(normals, s.m_normals.data(), sizeof(float3) * s.m_normals.size());
  std::memcpy(texcoords, s.m_texcoords.data(), sizeof(float2) * s.m_texcoords.size());
  std::memcpy(indices, s.m_indices.data(), sizeof(uint3) * s.m_indices.size());
  std::memcpy(material_ids, s.m_material_ids.data(), 4 * s.m_material_ids.size());
  std::memcpy(instance_ids, s.m_instance_ids.data(), 4 * s.m_instance_ids.size());
  std::memcpy(materials, s.m_materials.data(), sizeof(Material) * s.m_materials.size());
  std::memcpy(submesh_offsets, s.m_submesh_offsets.data(), 4 * s.m_submesh_offsets.size());
  std::memcpy(submesh_n_faces, s.m_submesh_n_faces.data(), 4 * s.m_submesh_n_faces.size());
  std::memcpy(transforms, s.m_transforms.data(), 64 * s.m_transforms.size());
  std::memcpy(camera_transform16, &s.m_camera_transform, 64);
}

void texture_info(const Scene& s, uint32_t i, uint32_t* width, uint32_t* height, uint32_t* is_color)
{
  const Texture& t = s.m_textures.at(i);
  *width = t.m_width;
  *height = t.m_height;
  *is_color = t.m_texture_type == TextureType::COLOR;
}

void texture_data(const Scene& s, uint32_t i, uint8_t* rgba8)
{
  const Texture& t = s.m_textures.at(i);
  std::memcpy(rgba8, t.m_data.data(), 4 * (size_t)t.m_width * t.m_height);
}
}  // namespace

extern "C" {

const char* fr_last_error(void) { return g_error.c_str(); }

int fr_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* fr_version(void) { return "fredholm_b200 0.1.0 sm_100a"; }

fr_renderer* fr_renderer_create(int cuda_device)
{
  fr_renderer* r = nullptr;
  if (guarded([&] { r = new fr_renderer(cuda_device); }) != 0) return nullptr;
  return r;
}

void fr_renderer_destroy(fr_renderer* r)
{
  guarded([&] { delete r; });
}

int fr_load_scene(fr_renderer* r, const char* path, int clear)
{
  return guarded([&] { r->renderer.load_scene(path, clear != 0); });
}

int fr_stage_texture(fr_renderer* r, const uint8_t* rgba8, uint32_t width, uint32_t height, int is_color)
{
  int id = -1;
  const int rc = guarded([&] {
    r->staged_textures.emplace_back(width, height, reinterpret_cast<const uchar4*>(rgba8),
                                    is_color ? TextureType::COLOR : TextureType::NONCOLOR);
    id = (int)r->staged_textures.size() - 1;
  });
  return rc == 0 ? id : -1;
}

static void fill_scene(Scene& s, const float* vertices, const float* normals, const float* texcoords,
                       uint32_t n_vertices, const uint32_t* indices, const uint32_t* material_ids,
                       const uint32_t* instance_ids, uint32_t n_faces, const void* materials, uint32_t n_materials,
                       const uint32_t* submesh_offsets, const uint32_t* submesh_n_faces, const float* transforms,
                       uint32_t n_submeshes)
{
  if (!vertices || !normals || !texcoords || !indices || !material_ids || !instance_ids || !materials ||
      !submesh_offsets || !submesh_n_faces || !transforms)
    throw std::runtime_error("invalid scene: null array");
  s.m_vertices.resize(n_vertices);
  s.m_normals.resize(n_vertices);
  s.m_texcoords.resize(n_vertices);
  std::memcpy(s.m_vertices.data(), vertices, sizeof(float3) * n_vertices);
  std::memcpy(s.m_normals.data(), normals, sizeof(float3) * n_vertices);
  std::memcpy(s.m_texcoords.data(), texcoords, sizeof(float2) * n_vertices);
  s.m_indices.resize(n_faces);
  std::memcpy(s.m_indices.data(), indices, sizeof(uint3) * n_faces);
  s.m_material_ids.assign(material_ids, material_ids + n_faces);
  s.m_instance_ids.assign(instance_ids, instance_ids + n_faces);
  s.m_materials.resize(n_materials);
  std::memcpy(s.m_materials.data(), materials, sizeof(Material) * n_materials);
  s.m_submesh_offsets.assign(submesh_offsets, submesh_offsets + n_submeshes);
  s.m_submesh_n_faces.assign(submesh_n_faces, submesh_n_faces + n_submeshes);
  s.m_transforms.resize(n_submeshes);
  std::memcpy(s.m_transforms.data(), transforms, sizeof(float) * 16 * n_submeshes);
}

int fr_set_scene_arrays(fr_renderer* r, const float* vertices, const float* normals, const float* texcoords,
                        uint32_t n_vertices, const uint32_t* indices, const uint32_t* material_ids,
                        const uint32_t* instance_ids, uint32_t n_faces, const void* materials, uint32_t n_materials,
                        const uint32_t* submesh_offsets, const uint32_t* submesh_n_faces, const float* transforms,
                        uint32_t n_submeshes)
{
  return guarded([&] {
    Scene s;
    fill_scene(s, vertices, normals, texcoords, n_vertices, indices, material_ids, instance_ids, n_faces, materials,
               n_materials, submesh_offsets, submesh_n_faces, transforms, n_submeshes);
    s.m_textures = std::move(r->staged_textures);
    r->staged_textures.clear();
    r->renderer.set_scene(s);
  });
}

int fr_get_scene_sizes(fr_renderer* r, uint32_t* out6)
{
  return guarded([&] { scene_sizes(r->renderer.get_scene(), out6); });
}

int fr_get_scene_arrays(fr_renderer* r, float* vertices, float* normals, float* texcoords, uint32_t* indices,
                        uint32_t* material_ids, uint32_t* instance_ids, void* materials, uint32_t* submesh_offsets,
                        uint32_t* submesh_n_faces, float* transforms, float* camera_transform16)
{
  return guarded([&] {
    scene_arrays(r->renderer.get_scene(), vertices, normals, texcoords, indices, material_ids, instance_ids,
                 materials, submesh_offsets, submesh_n_faces, transforms, camera_transform16);
  });
}

int fr_get_texture_info(fr_renderer* r, uint32_t i, uint32_t* width, uint32_t* height, uint32_t* is_color)
{
  return guarded([&] { texture_info(r->renderer.get_scene(), i, width, height, is_color); });
}

int fr_get_texture_data(fr_renderer* r, uint32_t i, uint8_t* rgba8)
{
  return guarded([&] { texture_data(r->renderer.get_scene(), i, rgba8); });
}

fr_scene* fr_scene_create(void)
{
  fr_scene* s = nullptr;
  if (guarded([&] { s = new fr_scene(); }) != 0) return nullptr;
  return s;
}
void fr_scene_destroy(fr_scene* s) { delete s; }
int fr_scene_load(fr_scene* s, const char* path, int clear)
{
  return guarded([&] {
    s->scene.load_model(path, clear != 0);
    if (!s->scene.is_valid()) throw std::runtime_error("invalid scene");
  });
}
int fr_scene_set_arrays(fr_scene* s, const float* vertices, const float* normals, const float* texcoords,
                        uint32_t n_vertices, const uint32_t* indices, const uint32_t* material_ids,
                        const uint32_t* instance_ids, uint32_t n_faces, const void* materials, uint32_t n_materials,
                        const uint32_t* submesh_offsets, const uint32_t* submesh_n_faces, const float* transforms,
                        uint32_t n_submeshes)
{
  return guarded([&] {
    s->scene.clear();
    fill_scene(s->scene, vertices, normals, texcoords, n_vertices, indices, material_ids, instance_ids, n_faces,
               materials, n_materials, submesh_offsets, submesh_n_faces, transforms, n_submeshes);
  });
}
int fr_scene_validate(const fr_scene* s)
{
  return guarded([&] { s->scene.validate(); });
}
int fr_scene_get_sizes(fr_scene* s, uint32_t* out6)
{
  return guarded([&] { scene_sizes(s->scene, out6); });
}
int fr_scene_get_arrays(fr_scene* s, float* vertices, float* normals, float* texcoords, uint32_t* indices,
                        uint32_t* material_ids, uint32_t* instance_ids, void* materials, uint32_t* submesh_offsets,
                        uint32_t* submesh_n_faces, float* transforms, float* camera_transform16)
{
  return guarded([&] {
    scene_arrays(s->scene, vertices, normals, texcoords, indices, material_ids, instance_ids, materials,
                 submesh_offsets, submesh_n_faces, transforms, camera_transform16);
  });
}
int fr_scene_get_texture_info(fr_scene* s, uint32_t i, uint32_t* width, uint32_t* height, uint32_t* is_color)
{
  return guarded([&] { texture_info(s->scene, i, width, height, is_color); });
}
int fr_scene_get_texture_data(fr_scene* s, uint32_t i, uint8_t* rgba8)
{
  return guarded([&] { texture_data(s->scene, i, rgba8); });
}
int fr_scene_update_animation(fr_scene* s, float time)
{
  return guarded([&] { s->scene.update_animation(time); });
}
int fr_set_scene(fr_renderer* r, const fr_scene* s)
{
  return guarded([&] { r->renderer.set_scene(s->scene); });
}

int fr_build_accel(fr_renderer* r)
{
  return guarded([&] {
    r->renderer.build_gas();
    r->renderer.build_ias();
  });
}

int fr_get_accel_info(fr_renderer* r, uint32_t* out3, float* build_ms, uint64_t* bytes)
{
  return guarded([&] {
    const AccelInfo a = r->renderer.get_accel_info();
    out3[0] = a.n_faces;
    out3[1] = a.n_nodes;
    out3[2] = a.depth;
    if (build_ms) *build_ms = a.build_ms;
    if (bytes) *bytes = a.bytes;
  });
}

int fr_set_accel_mode(fr_renderer* r, int mode)
{
  return guarded([&] {
    if (mode < 0 || mode > 2) throw std::invalid_argument("fr_set_accel_mode: 0 = auto, 1 = flat, 2 = two-level");
    r->renderer.set_accel_mode(static_cast<AccelMode>(mode));
  });
}
int fr_get_accel_info2(fr_renderer* r, uint32_t* out5, float* tlas_update_ms)
{
  return guarded([&] {
    const AccelInfo a = r->renderer.get_accel_info();
    out5[0] = (a.two_level ? 1u : 0u) | (a.tlas_refitted ? 2u : 0u);
    out5[1] = a.n_instances;
    out5[2] = a.n_meshes;
    out5[3] = a.n_stored_faces;
    out5[4] = a.n_nodes;
    if (tlas_update_ms) *tlas_update_ms = a.tlas_update_ms;
  });
}

int fr_get_accel_data(fr_renderer* r, void* nodes80, void* tris48)
{
  return guarded([&] {
    if (r->renderer.impl()->two_level) {
      // two-level structure: the combined arrays -- n_nodes nodes, (instances + stored faces) 48-byte records
      const frd::TwoLevelBvh& t = r->renderer.impl()->bvh2;
      FR_CUDA_CHECK(cudaDeviceSynchronize());
      if (nodes80) FR_CUDA_CHECK(cudaMemcpy2D(nodes80, 80, t.nodes.get(), sizeof(frd::Node8), 80, t.n_nodes, cudaMemcpyDeviceToHost));
      if (tris48) FR_CUDA_CHECK(cudaMemcpy(tris48, t.tris.get(), 48ull * (t.n_instances + t.n_blas_faces), cudaMemcpyDeviceToHost));
      return;
    }
    const frd::DeviceBvh& b = r->renderer.impl()->bvh;
    FR_CUDA_CHECK(cudaDeviceSynchronize());
    // 80 bytes per node whatever the stride of the device array (Node8 may carry alignment padding)
    if (nodes80) FR_CUDA_CHECK(cudaMemcpy2D(nodes80, 80, b.nodes.get(), sizeof(frd::Node8), 80, b.n_nodes, cudaMemcpyDeviceToHost));
    if (tris48) FR_CUDA_CHECK(cudaMemcpy(tris48, b.tris.get(), 48ull * b.n_faces, cudaMemcpyDeviceToHost));
  });
}

int fr_set_time(fr_renderer* r, float time)
{
  return guarded([&] { r->renderer.set_time(time); });
}

int fr_set_transforms(fr_renderer* r, const float* transforms, uint32_t n_submeshes)
{
  return guarded([&] {
    r->renderer.set_transforms(transforms, n_submeshes);
  });
}

int fr_set_directional_light(fr_renderer* r, const float* le3, const float* dir3, float angle_deg)
{
  return guarded([&] {
    r->renderer.set_directional_light(make_float3(le3[0], le3[1], le3[2]), make_float3(dir3[0], dir3[1], dir3[2]),
                                      angle_deg);
  });
}
int fr_clear_directional_light(fr_renderer* r)
{
  return guarded([&] { r->renderer.clear_directional_light(); });
}
int fr_set_sky_intensity(fr_renderer* r, float v)
{
  return guarded([&] { r->renderer.set_sky_intensity(v); });
}
int fr_load_arhosek_sky(fr_renderer* r, float turbidity, float albedo)
{
  return guarded([&] { r->renderer.load_arhosek_sky(turbidity, albedo); });
}
int fr_clear_arhosek_sky(fr_renderer* r)
{
  return guarded([&] { r->renderer.clear_arhosek_sky(); });
}
int fr_set_ibl(fr_renderer* r, const float* rgba32f, uint32_t width, uint32_t height)
{
  return guarded([&] { r->renderer.set_ibl(reinterpret_cast<const float4*>(rgba32f), width, height); });
}
int fr_load_ibl(fr_renderer* r, const char* path)
{
  return guarded([&] { r->renderer.load_ibl(path); });
}
int fr_clear_ibl(fr_renderer* r)
{
  return guarded([&] { r->renderer.clear_ibl(); });
}

int fr_set_resolution(fr_renderer* r, uint32_t width, uint32_t height)
{
  return guarded([&] { r->renderer.set_resolution(width, height); });
}
int fr_init_render_states(fr_renderer* r)
{
  return guarded([&] { r->renderer.init_render_states(); });
}
int fr_set_sample_offset(fr_renderer* r, uint32_t first_sample)
{
  return guarded([&] { r->renderer.set_sample_offset(first_sample); });
}
uint32_t fr_get_sample_count(fr_renderer* r) { return r->renderer.get_sample_count(); }
int fr_set_film_mode(fr_renderer* r, int mode)
{
  return guarded([&] { r->renderer.set_film_mode(mode == 0 ? FilmMode::MEAN : FilmMode::SUM); });
}
int fr_set_max_wave_paths(fr_renderer* r, uint64_t n_paths)
{
  return guarded([&] { r->renderer.set_max_wave_paths((size_t)n_paths); });
}
int fr_set_single_launch(fr_renderer* r, int on)
{
  return guarded([&] { r->renderer.set_single_launch(on != 0); });
}

int fr_render(fr_renderer* r, const float* camera_transform12, float fov, float F, float focus, const float* bg,
              const fr_layers* layers_dev, uint32_t n_samples, uint32_t max_depth)
{
  return guarded([&] {
    if (!layers_dev || !layers_dev->beauty) throw std::runtime_error("fr_render: beauty layer is required");
    r->renderer.render(make_camera(camera_transform12, fov, F, focus), make_float3(bg[0], bg[1], bg[2]),
                       make_layers(layers_dev), n_samples, max_depth);
  });
}

int fr_wait(fr_renderer* r)
{
  return guarded([&] { r->renderer.wait_for_completion(); });
}

int fr_render_frame_host(fr_renderer* r, const float* camera_transform12, float fov, float F, float focus,
                         const float* bg, const fr_layers* layers_host, uint32_t n_samples, uint32_t max_depth)
{
  return guarded([&] {
    if (!layers_host || !layers_host->beauty) throw std::runtime_error("fr_render_frame_host: beauty is required");
    Renderer::Impl* im = r->renderer.impl();
    const size_t n = (size_t)im->width * im->height;
    cudaStream_t s = im->stream;
    auto prep4 = [&](frd::DevBuf<float4>& b, bool want) -> float4* {
      if (!want) return nullptr;
      b.reserve(n);
      FR_CUDA_CHECK(cudaMemsetAsync(b.get(), 0, sizeof(float4) * n, s));
      return b.get();
    };
    RenderLayer L;
    L.beauty = prep4(r->h_beauty, true);
    L.position = prep4(r->h_position, layers_host->position != nullptr);
    L.normal = prep4(r->h_normal, layers_host->normal != nullptr);
    L.texcoord = prep4(r->h_texcoord, layers_host->texcoord != nullptr);
    L.albedo = prep4(r->h_albedo, layers_host->albedo != nullptr);
    L.depth = nullptr;
    if (layers_host->depth) {
      r->h_depth.reserve(n);
      FR_CUDA_CHECK(cudaMemsetAsync(r->h_depth.get(), 0, sizeof(float) * n, s));
      L.depth = r->h_depth.get();
    }
    r->renderer.init_render_states();
    r->renderer.render(make_camera(camera_transform12, fov, F, focus), make_float3(bg[0], bg[1], bg[2]), L, n_samples,
                       max_depth);
    auto back = [&](void* dst, const void* src, size_t bytes) {
      if (dst) FR_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    };
    back(layers_host->beauty, L.beauty, sizeof(float4) * n);
    back(layers_host->position, L.position, sizeof(float4) * n);
    back(layers_host->normal, L.normal, sizeof(float4) * n);
    back(layers_host->texcoord, L.texcoord, sizeof(float4) * n);
    back(layers_host->albedo, L.albedo, sizeof(float4) * n);
    back(layers_host->depth, L.depth, sizeof(float) * n);
    FR_CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

// ---- multi-GPU (include/fredholm/multi_gpu.h) ----
int fr_comm_get_unique_id(uint8_t* out128)
{
  return guarded([&] {
    const CommId id = make_comm_id();
    std::memcpy(out128, id.bytes, sizeof(id.bytes));
  });
}
int fr_comm_init(fr_renderer* r, const uint8_t* id128, int rank, int world)
{
  return guarded([&] {
    CommId id;
    std::memcpy(id.bytes, id128, sizeof(id.bytes));
    r->rank.reset();
    r->rank = std::make_unique<ShardedRenderer>(r->renderer, id, rank, world);
  });
}
int fr_comm_destroy(fr_renderer* r)
{
  return guarded([&] { r->rank.reset(); });
}
int fr_sample_slice(uint32_t total_spp, int rank, int world, uint32_t* first, uint32_t* count)
{
  return guarded([&] { sample_slice(total_spp, rank, world, *first, *count); });
}
int fr_render_sharded(fr_renderer* r, const float* camera_transform12, float fov, float F, float focus,
                      const float* bg, const fr_layers* layers_dev, uint32_t total_spp, uint32_t max_depth, int root)
{
  return guarded([&] {
    if (!r->rank) throw std::runtime_error("fr_render_sharded: call fr_comm_init first");
    r->rank->render(make_camera(camera_transform12, fov, F, focus), make_float3(bg[0], bg[1], bg[2]),
                    make_layers(layers_dev), total_spp, max_depth, root);
  });
}
int fr_reduce_layers(fr_renderer* r, const fr_layers* layers_dev, uint32_t total_spp, int root)
{
  return guarded([&] {
    if (!r->rank) throw std::runtime_error("fr_reduce_layers: call fr_comm_init first");
    r->rank->reduce(make_layers(layers_dev), total_spp, root);
  });
}

struct fr_multi {
  MultiGpuRenderer multi;
  std::vector<std::unique_ptr<fr_renderer>> views;  // per-rank handles for the fr_* scene / light / film calls
  explicit fr_multi(const std::vector<int>& devs) : multi(devs)
  {
    for (int i = 0; i < multi.size(); ++i) views.emplace_back(new fr_renderer(multi.renderer(i)));
  }
};
fr_multi* fr_multi_create(const int* devices, int n_devices)
{
  fr_multi* m = nullptr;
  if (guarded([&] {
        std::vector<int> devs;
        if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
        m = new fr_multi(devs);
      }) != 0)
    return nullptr;
  return m;
}
void fr_multi_destroy(fr_multi* m)
{
  guarded([&] { delete m; });
}
int fr_multi_size(fr_multi* m) { return m ? m->multi.size() : 0; }
fr_renderer* fr_multi_renderer(fr_multi* m, int rank)
{
  if (!m || rank < 0 || rank >= (int)m->views.size()) {
    g_error = "fr_multi_renderer: rank outside the world";
    return nullptr;
  }
  return m->views[rank].get();
}
int fr_multi_render(fr_multi* m, const float* camera_transform12, float fov, float F, float focus, const float* bg,
                    const fr_layers* layers_dev_rank0, uint32_t total_spp, uint32_t max_depth)
{
  return guarded([&] {
    m->multi.render(make_camera(camera_transform12, fov, F, focus), make_float3(bg[0], bg[1], bg[2]),
                    make_layers(layers_dev_rank0), total_spp, max_depth);
  });
}
int fr_multi_wait(fr_multi* m)
{
  return guarded([&] { m->multi.wait_for_completion(); });
}

int fr_set_device(int device)
{
  return guarded([&] { FR_CUDA_CHECK(cudaSetDevice(device)); });
}

int fr_get_device_attributes(int device, uint32_t* out4, uint64_t* total_mem)
{
  return guarded([&] {
    cudaDeviceProp p;
    FR_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    int clock_khz = 0;
    FR_CUDA_CHECK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, device));
    out4[0] = (uint32_t)p.multiProcessorCount;
    out4[1] = (uint32_t)clock_khz;
    out4[2] = (uint32_t)p.major * 10u + (uint32_t)p.minor;
    out4[3] = (uint32_t)p.l2CacheSize;
    if (total_mem) *total_mem = (uint64_t)p.totalGlobalMem;
  });
}

int fr_scale_layers(fr_renderer* r, const fr_layers* layers_dev, float scale)
{
  return guarded([&] { r->renderer.scale_layers(make_layers(layers_dev), scale); });
}

int fr_get_statistics(fr_renderer* r, uint64_t* out6)
{
  return guarded([&] {
    const RenderStatistics s = r->renderer.get_statistics();
    out6[0] = s.paths;
    out6[1] = s.rays_radiance;
    out6[2] = s.rays_shadow;
    out6[3] = s.rays_light;
    out6[4] = s.kernel_launches;
    out6[5] = s.rays_skipped;
  });
}
int fr_set_wave_overlap(fr_renderer* r, int on)
{
  return guarded([&] { r->renderer.set_wave_overlap(on != 0); });
}
int fr_set_wave_compaction(fr_renderer* r, int on, uint32_t depth)
{
  return guarded([&] { r->renderer.set_wave_compaction(on != 0, depth ? depth : 3u); });
}
uint64_t fr_get_wave_state_bytes(fr_renderer* r) { return r ? (uint64_t)r->renderer.get_wave_state_bytes() : 0; }
int fr_set_traversal_counting(fr_renderer* r, int on)
{
  return guarded([&] { r->renderer.set_traversal_counting(on != 0); });
}
int fr_get_traversal_counters(fr_renderer* r, uint64_t* out6)
{
  return guarded([&] {
    const RenderStatistics s = r->renderer.get_statistics();
    for (int i = 0; i < 3; ++i) {
      out6[i] = s.nodes_visited[i];
      out6[3 + i] = s.tris_tested[i];
    }
  });
}
int fr_set_samples_per_warp(fr_renderer* r, uint32_t spw)
{
  return guarded([&] { r->renderer.set_samples_per_warp(spw); });
}
int fr_reset_statistics(fr_renderer* r)
{
  return guarded([&] { r->renderer.reset_statistics(); });
}
int fr_set_stage_timing(fr_renderer* r, int on)
{
  return guarded([&] { r->renderer.set_stage_timing(on != 0); });
}
int fr_get_stage_times(fr_renderer* r, double* ms7, uint64_t* launches7)
{
  return guarded([&] {
    unsigned long long l[Renderer::kStageCount];
    r->renderer.get_stage_times(ms7, l);
    for (int i = 0; i < Renderer::kStageCount; ++i) launches7[i] = l[i];
  });
}
void* fr_event_create(void)
{
  cudaEvent_t e = nullptr;
  if (guarded([&] { FR_CUDA_CHECK(cudaEventCreate(&e)); }) != 0) return nullptr;
  return e;
}
int fr_event_destroy(void* ev)
{
  return guarded([&] { FR_CUDA_CHECK(cudaEventDestroy(static_cast<cudaEvent_t>(ev))); });
}
int fr_event_record(fr_renderer* r, void* ev)
{
  return guarded([&] { FR_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(ev), r->renderer.get_stream())); });
}
int fr_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms)
{
  return guarded([&] {
    FR_CUDA_CHECK(cudaEventSynchronize(static_cast<cudaEvent_t>(ev_stop)));
    FR_CUDA_CHECK(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(ev_start), static_cast<cudaEvent_t>(ev_stop)));
  });
}
uint64_t fr_get_stream(fr_renderer* r) { return reinterpret_cast<uint64_t>(r->renderer.get_stream()); }

int fr_post_process(const void* beauty_in_dev, void* high_luminance_dev, void* temp_dev, int width, int height,
                    const fr_post_process_params* p, void* beauty_out_dev)
{
  return guarded([&] {
    PostProcessParams pp;
    pp.use_bloom = p->use_bloom != 0;
    pp.bloom_threshold = p->bloom_threshold;
    pp.bloom_sigma = p->bloom_sigma;
    pp.ISO = p->ISO;
    pp.chromatic_aberration = p->chromatic_aberration;
    post_process_kernel_launch(static_cast<const float4*>(beauty_in_dev), static_cast<float4*>(high_luminance_dev),
                               static_cast<float4*>(temp_dev), width, height, pp,
                               static_cast<float4*>(beauty_out_dev));
    FR_CUDA_CHECK(cudaDeviceSynchronize());
  });
}

int fr_tone_mapping(const void* beauty_in_dev, int width, int height, float ISO, float chromatic_aberration,
                    void* beauty_out_dev)
{
  return guarded([&] {
    tone_mapping_kernel_launch(static_cast<const float4*>(beauty_in_dev), width, height, ISO, chromatic_aberration,
                               static_cast<float4*>(beauty_out_dev));
    FR_CUDA_CHECK(cudaDeviceSynchronize());
  });
}

int fr_denoise(const void* beauty_dev, const void* normal_dev, const void* albedo_dev, void* denoised_dev,
               uint32_t width, uint32_t height, int upscale, int iterations, float sigma_color, float sigma_albedo,
               float albedo_floor, float firefly_k)
{
  return guarded([&] {
    Denoiser d(width, height, static_cast<const float4*>(beauty_dev), static_cast<const float4*>(normal_dev),
               static_cast<const float4*>(albedo_dev), static_cast<float4*>(denoised_dev), upscale != 0);
    DenoiserParams p;
    if (iterations > 0) p.iterations = iterations;
    if (sigma_color > 0.0f) p.sigma_color = sigma_color;
    if (sigma_albedo > 0.0f) p.sigma_albedo = sigma_albedo;
    if (albedo_floor > 0.0f) p.albedo_floor = albedo_floor;
    if (firefly_k >= 0.0f) p.firefly_k = firefly_k;
    d.set_params(p);
    d.denoise();
    d.wait_for_completion();
  });
}

int fr_batch_run(fr_renderer* r, const fr_batch_config* c, const float* camera_transforms12, uint32_t n_camera_frames,
                 float fov, float F, float focus, fr_frame_record* records, uint32_t max_records,
                 uint8_t* frames_rgba8, uint32_t* out5, double* wall_s)
{
  return guarded([&] {
    BatchConfig cfg;
    cfg.width = c->width;
    cfg.height = c->height;
    cfg.n_spp = c->n_spp;
    cfg.max_depth = c->max_depth;
    cfg.post.use_bloom = c->post.use_bloom != 0;
    cfg.post.bloom_threshold = c->post.bloom_threshold;
    cfg.post.bloom_sigma = c->post.bloom_sigma;
    cfg.post.ISO = c->post.ISO;
    cfg.post.chromatic_aberration = c->post.chromatic_aberration;
    cfg.denoise = c->denoise != 0;
    cfg.upscale = c->upscale != 0;
    if (c->dn_iterations > 0) cfg.denoiser.iterations = c->dn_iterations;
    if (c->dn_sigma_color > 0.0f) cfg.denoiser.sigma_color = c->dn_sigma_color;
    if (c->dn_sigma_albedo > 0.0f) cfg.denoiser.sigma_albedo = c->dn_sigma_albedo;
    if (c->dn_albedo_floor > 0.0f) cfg.denoiser.albedo_floor = c->dn_albedo_floor;
    if (c->dn_firefly_k >= 0.0f) cfg.denoiser.firefly_k = c->dn_firefly_k;
    cfg.fps = c->fps;
    cfg.start_time = c->start_time;
    cfg.max_time = c->max_time;
    cfg.kill_time_s = c->kill_time_s;
    cfg.first_frame = c->first_frame;
    cfg.frame_stride = c->frame_stride;
    cfg.max_frames = std::min(c->max_frames, max_records);
    for (int i = 0; i < 3; ++i) cfg.bg_color[i] = c->bg_color[i];
    cfg.animate = c->animate != 0;
    cfg.output_dir = c->output_dir ? c->output_dir : "";
    cfg.n_save_threads = c->n_save_threads;
    cfg.n_slots = c->n_slots;
    cfg.keep_frames = frames_rgba8 != nullptr;

    auto set_camera = [&](Camera& cam, uint32_t i) {
      const float* t = camera_transforms12 + 12ull * i;
      cam.m_transform = mat4();
      for (int row = 0; row < 3; ++row)
        for (int col = 0; col < 4; ++col) cam.m_transform[col][row] = t[4 * row + col];
    };
    Camera camera;
    camera.m_fov = fov;
    camera.m_F = F;
    camera.m_focus = focus;
    set_camera(camera, 0);
    FrameBatch batch(r->renderer, cfg);
    FrameBatch::FrameHook hook;
    if (n_camera_frames > 0)
      hook = [&](uint32_t frame_idx, float, Camera& cam) { set_camera(cam, std::min(frame_idx, n_camera_frames - 1)); };
    const BatchResult res = batch.run(camera, hook);
    const size_t frame_bytes = (size_t)res.out_width * res.out_height * 4;
    for (size_t i = 0; i < res.frames.size(); ++i) {
      const FrameRecord& f = res.frames[i];
      fr_frame_record& o = records[i];
      o.frame_idx = f.frame_idx;
      o.time = f.time;
      o.accel_ms = f.accel_ms;
      o.render_ms = f.render_ms;
      o.denoise_ms = f.denoise_ms;
      o.post_ms = f.post_ms;
      o.transfer_ms = f.transfer_ms;
      o.encode_ms = f.encode_ms;
      o.save_ms = f.save_ms;
      o.png_bytes = f.png_bytes;
      if (frames_rgba8) std::memcpy(frames_rgba8 + i * frame_bytes, f.rgba8.data(), frame_bytes);
    }
    out5[0] = (uint32_t)res.frames.size();
    out5[1] = res.out_width;
    out5[2] = res.out_height;
    out5[3] = res.killed ? 1u : 0u;
    out5[4] = 0u;
    if (wall_s) *wall_s = res.wall_s;
  });
}

void* fr_device_alloc(size_t bytes)
{
  void* p = nullptr;
  if (guarded([&] { FR_CUDA_CHECK(cudaMalloc(&p, bytes)); }) != 0) return nullptr;
  return p;
}
int fr_device_free(void* p)
{
  return guarded([&] { FR_CUDA_CHECK(cudaFree(p)); });
}
int fr_device_memset(void* p, int value, size_t bytes)
{
  return guarded([&] { FR_CUDA_CHECK(cudaMemset(p, value, bytes)); });
}
int fr_copy_to_device(void* dst_dev, const void* src_host, size_t bytes)
{
  return guarded([&] { FR_CUDA_CHECK(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice)); });
}
int fr_copy_to_host(void* dst_host, const void* src_dev, size_t bytes)
{
  return guarded([&] { FR_CUDA_CHECK(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost)); });
}
void* fr_host_alloc_pinned(size_t bytes)
{
  void* p = nullptr;
  if (guarded([&] { FR_CUDA_CHECK(cudaMallocHost(&p, bytes)); }) != 0) return nullptr;
  return p;
}
int fr_host_free_pinned(void* p)
{
  return guarded([&] { FR_CUDA_CHECK(cudaFreeHost(p)); });
}
int fr_device_synchronize(void)
{
  return guarded([&] { FR_CUDA_CHECK(cudaDeviceSynchronize()); });
}

namespace
{
thread_local fredholm::Texture t_image8;
thread_local fredholm::FloatTexture t_imagef;
}  // namespace

int fr_image8_load(const char* path, uint32_t* width, uint32_t* height)
{
  return guarded([&] {
    t_image8 = fredholm::Texture(std::filesystem::path(path), fredholm::TextureType::NONCOLOR);
    *width = t_image8.m_width;
    *height = t_image8.m_height;
  });
}
int fr_image8_copy(uint8_t* rgba8)
{
  return guarded([&] { std::memcpy(rgba8, t_image8.m_data.data(), 4 * t_image8.m_data.size()); });
}
int fr_imagef_load(const char* path, uint32_t* width, uint32_t* height)
{
  return guarded([&] {
    t_imagef = fredholm::FloatTexture(std::filesystem::path(path));
    *width = t_imagef.m_width;
    *height = t_imagef.m_height;
  });
}
int fr_imagef_copy(float* rgba32f)
{
  return guarded([&] { std::memcpy(rgba32f, t_imagef.m_data.data(), 16 * t_imagef.m_data.size()); });
}
int fr_write_png(const char* path, const uint8_t* pixels, uint32_t width, uint32_t height, uint32_t channels)
{
  return guarded([&] { fredholm::codec::write_png(path, pixels, (int)width, (int)height, (int)channels); });
}

int fr_trace_closest(fr_renderer* r, const float* rays, uint32_t n, float tmin, float tmax, uint32_t* out_id,
                     float* out_tuv, uint64_t* counters2)
{
  return guarded([&] {
    Renderer::Impl* im = r->renderer.impl();
    if (!im->accel_valid) im->build_accel();
    FR_CUDA_CHECK(cudaStreamSynchronize(im->stream));
    const frd::SceneView v = im->view(make_float3(0, 0, 0));
    frd::trace_batch_closest(v, im->d_submesh_offsets.get(), rays, n, tmin, tmax, out_id, out_tuv,
                             reinterpret_cast<unsigned long long*>(counters2));
  });
}

int fr_primary_rays(fr_renderer* r, const float* camera_transform12, float fov, float F, float focus, uint32_t n_spp,
                    float* out_rays)
{
  return guarded([&] {
    Renderer::Impl* im = r->renderer.impl();
    frd::WaveParams wp;
    wp.film = frd::make_film_geom(im->width, im->height);
    wp.n_samples = 1;
    wp.sample_base = n_spp;
    wp.max_depth = 1;
    wp.seed = 1;
    wp.camera = make_camera(camera_transform12, fov, F, focus);
    frd::test_primary_rays(wp, out_rays);
  });
}

int fr_sampler_sequence(uint32_t width, uint32_t height, uint32_t seed, uint32_t image_idx, uint32_t n_spp,
                        const char* kinds, float* out)
{
  return guarded([&] {
    uint32_t n_out = 0;
    for (const char* k = kinds; *k; ++k) n_out += (*k == '1') ? 1u : 2u;
    frd::test_sampler(width, height, seed, image_idx, n_spp, kinds, out, n_out);
  });
}

int fr_bsdf_eval_sample(const float* in, uint32_t n, float* out)
{
  return guarded([&] { frd::test_bsdf(in, n, out); });
}

int fr_sky_radiance(fr_renderer* r, const float* dirs, uint32_t n, float* out)
{
  return guarded([&] {
    const frd::SceneView v = r->renderer.impl()->view(make_float3(0, 0, 0));
    frd::test_sky(v, dirs, n, out);
  });
}

int fr_arhosek_cook(float turbidity, float albedo, float elevation, float* out30)
{
  return guarded([&] { arhosek_rgb_cook(turbidity, albedo, elevation, out30); });
}

int fr_camera_transform(const float* origin3, float* out12)
{
  return guarded([&] {
    const Camera cam(make_float3(origin3[0], origin3[1], origin3[2]));
    camera_rows(cam, out12);
  });
}

int fr_camera_walk(const float* origin3, float d_phi, float d_theta, int movement, float dt, float* out12)
{
  return guarded([&] {
    Camera cam(make_float3(origin3[0], origin3[1], origin3[2]));
    cam.lookAround(d_phi, d_theta);
    cam.move(static_cast<CameraMovement>(movement), dt);
    camera_rows(cam, out12);
  });
}

}  // extern "C"
