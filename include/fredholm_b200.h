/* C ABI of libfredholm_b200.so -- the drop-in boundary of the B200 rendering core.
 *
 * The reference exposes its hot path as a C++ class API (fredholm::Renderer,
 * fredholm/include/fredholm/renderer.h), not as an FFI; the C++ mirror of that
 * API is include/fredholm/renderer.h.  This header flattens the same calls to
 * plain C (pointers + sizes, no C++ or torch types) so that any host language
 * can bind them (ctypes / cgo / JNI).  Every entry names the reference interface
 * it stands for.  All functions return 0 on success and -1 on failure unless
 * stated otherwise; fr_last_error() then returns the message of the C++
 * exception (the reference reports errors as std::runtime_error, cwl/util.h:11-56).
 *
 * Pointer conventions: `const float*` / `const uint32_t*` arguments are HOST
 * pointers unless the name ends in `_dev`.  Matrices: `transforms` are
 * column-major 4x4 (glm::mat4 layout, 16 floats); camera transforms are the
 * row-major 3x4 camera-to-world block (12 floats) that Renderer::render packs
 * into CameraParams (renderer.h:678-684).
 */
#ifndef FREDHOLM_B200_H
#define FREDHOLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fr_renderer fr_renderer;

/* six AOV buffers of shared.h:201-208 (float4 per pixel, depth: float per pixel);
 * NULL entries (except beauty) are skipped */
typedef struct fr_layers {
  void* beauty;
  void* position;
  void* depth;
  void* normal;
  void* texcoord;
  void* albedo;
} fr_layers;

/* kernels/post-process.h:4-10 */
typedef struct fr_post_process_params {
  int use_bloom;
  float bloom_threshold;
  float bloom_sigma;
  float ISO;
  float chromatic_aberration;
} fr_post_process_params;

const char* fr_last_error(void);
int fr_device_count(void);
/* version string: "fredholm_b200 <semver> sm_100a" */
const char* fr_version(void);

/* ---- lifetime: Renderer(ctx) / ~Renderer (renderer.h:32-122) ---- */
fr_renderer* fr_renderer_create(int cuda_device);
void fr_renderer_destroy(fr_renderer* r);

/* ---- scene: Renderer::load_scene (renderer.h:354-432), Scene arrays (scene.h:107-130) ---- */
int fr_load_scene(fr_renderer* r, const char* path, int clear);
/* staged textures are attached by the next fr_set_scene_arrays; returns the texture id */
int fr_stage_texture(fr_renderer* r, const uint8_t* rgba8, uint32_t width, uint32_t height, int is_color);
int fr_set_scene_arrays(fr_renderer* r, const float* vertices, const float* normals, const float* texcoords,
                        uint32_t n_vertices, const uint32_t* indices, const uint32_t* material_ids,
                        const uint32_t* instance_ids, uint32_t n_faces, const void* materials /* 180 B each */,
                        uint32_t n_materials, const uint32_t* submesh_offsets, const uint32_t* submesh_n_faces,
                        const float* transforms, uint32_t n_submeshes);
/* loader introspection: out6 = n_vertices, n_faces, n_materials, n_textures, n_submeshes, has_camera */
int fr_get_scene_sizes(fr_renderer* r, uint32_t* out6);
int fr_get_scene_arrays(fr_renderer* r, float* vertices, float* normals, float* texcoords, uint32_t* indices,
                        uint32_t* material_ids, uint32_t* instance_ids, void* materials, uint32_t* submesh_offsets,
                        uint32_t* submesh_n_faces, float* transforms, float* camera_transform16);
int fr_get_texture_info(fr_renderer* r, uint32_t i, uint32_t* width, uint32_t* height, uint32_t* is_color);
int fr_get_texture_data(fr_renderer* r, uint32_t i, uint8_t* rgba8);
/* Renderer::build_gas + build_ias (renderer.h:434-552): GPU LBVH -> CWBVH */
int fr_build_accel(fr_renderer* r);
/* out3 = n_faces, n_nodes, depth */
int fr_get_accel_info(fr_renderer* r, uint32_t* out3, float* build_ms, uint64_t* bytes);
/* acceleration-structure layout (extension): 0 = auto, 1 = flat world-space tree, 2 = two-level (one object-space
 * tree per distinct mesh + an instance tree, like the reference's GAS + IAS, renderer.h:434-552); takes effect at the
 * next fr_build_accel.  fr_get_accel_info2: out5 = flags (bit 0 two-level, bit 1 the last instance-tree update was
 * a refit in place), instances, distinct meshes, triangles stored, nodes; tlas_update_ms = device time of the last
 * instance-tree update (fr_set_time / fr_set_transforms in two-level mode: a one-launch bottom-up refit, or a
 * rebuild when the instances have spread so far that the refit asks for one) */
int fr_set_accel_mode(fr_renderer* r, int mode);
int fr_get_accel_info2(fr_renderer* r, uint32_t* out5, float* tlas_update_ms);
/* structure inspection (tests / tools): copies the n_nodes 80-byte CWBVH nodes and the n_faces
 * 48-byte leaf triangles (layout: fredholm_b200/csrc/bvh.cuh) to host memory; NULL skips */
int fr_get_accel_data(fr_renderer* r, void* nodes80, void* tris48);
/* Renderer::set_time (renderer.h:614-640) and direct transform replacement */
int fr_set_time(fr_renderer* r, float time);
int fr_set_transforms(fr_renderer* r, const float* transforms, uint32_t n_submeshes);

/* ---- stand-alone Scene (host only, no GPU needed): Scene::load_model (scene.cpp:103-117) ---- */
typedef struct fr_scene fr_scene;
fr_scene* fr_scene_create(void);
void fr_scene_destroy(fr_scene* s);
int fr_scene_load(fr_scene* s, const char* path, int clear);
/* fills the flat arrays of Scene (scene.h:107-130) directly; layout as fr_set_scene_arrays */
int fr_scene_set_arrays(fr_scene* s, const float* vertices, const float* normals, const float* texcoords,
                        uint32_t n_vertices, const uint32_t* indices, const uint32_t* material_ids,
                        const uint32_t* instance_ids, uint32_t n_faces, const void* materials /* 180 B each */,
                        uint32_t n_materials, const uint32_t* submesh_offsets, const uint32_t* submesh_n_faces,
                        const float* transforms, uint32_t n_submeshes);
/* Scene::validate (extension): every index the device code follows is in range, sub-meshes tile the faces;
 * fails with "invalid scene: ..." -- the check every upload (fr_set_scene*, fr_load_scene) runs first */
int fr_scene_validate(const fr_scene* s);
int fr_scene_get_sizes(fr_scene* s, uint32_t* out6);
int fr_scene_get_arrays(fr_scene* s, float* vertices, float* normals, float* texcoords, uint32_t* indices,
                        uint32_t* material_ids, uint32_t* instance_ids, void* materials, uint32_t* submesh_offsets,
                        uint32_t* submesh_n_faces, float* transforms, float* camera_transform16);
int fr_scene_get_texture_info(fr_scene* s, uint32_t i, uint32_t* width, uint32_t* height, uint32_t* is_color);
int fr_scene_get_texture_data(fr_scene* s, uint32_t i, uint8_t* rgba8);
/* Scene::update_animation (scene.cpp:862-898) */
int fr_scene_update_animation(fr_scene* s, float time);
/* hands the scene to a renderer (Renderer::load_scene's upload part) */
int fr_set_scene(fr_renderer* r, const fr_scene* s);

/* ---- lights / sky (renderer.h:554-612) ---- */
int fr_set_directional_light(fr_renderer* r, const float* le3, const float* dir3, float angle_deg);
int fr_clear_directional_light(fr_renderer* r);
int fr_set_sky_intensity(fr_renderer* r, float sky_intensity);
int fr_load_arhosek_sky(fr_renderer* r, float turbidity, float albedo);
int fr_clear_arhosek_sky(fr_renderer* r);
int fr_set_ibl(fr_renderer* r, const float* rgba32f, uint32_t width, uint32_t height);
int fr_load_ibl(fr_renderer* r, const char* path);
int fr_clear_ibl(fr_renderer* r);

/* ---- film (renderer.h:642-655) ---- */
int fr_set_resolution(fr_renderer* r, uint32_t width, uint32_t height);
int fr_init_render_states(fr_renderer* r);
int fr_set_sample_offset(fr_renderer* r, uint32_t first_sample);
uint32_t fr_get_sample_count(fr_renderer* r);
/* 0 = streaming mean (reference), 1 = sums (multi-GPU slices) */
int fr_set_film_mode(fr_renderer* r, int mode);
int fr_set_max_wave_paths(fr_renderer* r, uint64_t n_paths);
/* two waves in flight on two streams: fr_set_max_wave_paths is then the total of both (extension) */
int fr_set_wave_overlap(fr_renderer* r, int on);
/* wave compaction (extension, on by default): the paths a wave still has alive after `depth` bounces (0 = default,
   3) move to a dense straggler set shared by up to eight waves; images are bit-identical either way */
int fr_set_wave_compaction(fr_renderer* r, int on, uint32_t depth);
/* device memory held for path state + ray queues (268 B / path for a beauty-only render with sun + sky) */
uint64_t fr_get_wave_state_bytes(fr_renderer* r);
/* 1: fr_render(n_samples) reproduces ONE reference launch of n_samples, where RadiancePayload is declared
   outside the sample loop and firsthit is never reset (pt.cu:432-433, 509, 744-759; app/rtcamp8.cpp:186-190);
   0 (default): n_samples launches of one sample (app/controller.cpp:221-224). */
int fr_set_single_launch(fr_renderer* r, int on);

/* ---- render: Renderer::render / wait_for_completion (renderer.h:657-736) ----
 * layers hold DEVICE pointers owned by the caller; asynchronous on the renderer's stream */
int fr_render(fr_renderer* r, const float* camera_transform12, float fov, float F, float focus, const float* bg_color3,
              const fr_layers* layers_dev, uint32_t n_samples, uint32_t max_depth);
int fr_wait(fr_renderer* r);
/* one whole frame through HOST buffers: clears internal device layers, renders n_samples,
 * copies the non-NULL layers back (cwl::CUDABuffer::copy_from_device_to_host, buffer.h:64-69)
 * and waits.  This is the call the end-to-end benchmark times. */
int fr_render_frame_host(fr_renderer* r, const float* camera_transform12, float fov, float F, float focus,
                         const float* bg_color3, const fr_layers* layers_host, uint32_t n_samples, uint32_t max_depth);
/* cudaSetDevice for the calling thread: which device fr_device_alloc allocates on.  Every fr_* call on a
 * renderer switches the calling thread to that renderer's device, so a process that drives several devices sets
 * the device again before allocating caller-owned layers. */
int fr_set_device(int device);
/* out4 = SM count, SM clock (kHz, the attribute's maximum), compute capability x 10, L2 bytes */
int fr_get_device_attributes(int device, uint32_t* out4, uint64_t* total_mem);
int fr_scale_layers(fr_renderer* r, const fr_layers* layers_dev, float scale);

/* ---- multi-GPU: sample-sharded render with ONE ncclReduce of the accumulation buffers, inside the C++ core
 * (include/fredholm/multi_gpu.h; SURVEY.md 8(e): the reference is single-GPU, its batch application
 * app/rtcamp8.cpp:159-246 is the caller shape).  NCCL is loaded at run time; without it these fail.
 * (a) one rank per renderer -- threads of one process or one process per GPU: rank 0 makes the 128-byte id,
 *     every rank calls fr_comm_init with the same bytes (collective), then fr_render_sharded (collective):
 *     renders this rank's slice (fr_sample_slice: whole 16-sample CMJ patterns) of a total_spp frame into the
 *     ZEROED layers as sums, reduces every bound layer onto `root` on the renderer's stream and divides by
 *     total_spp there.  Asynchronous; fr_wait.  fr_reduce_layers is the exchange step alone. */
int fr_comm_get_unique_id(uint8_t* out128);
int fr_comm_init(fr_renderer* r, const uint8_t* id128, int rank, int world);
int fr_comm_destroy(fr_renderer* r);
int fr_sample_slice(uint32_t total_spp, int rank, int world, uint32_t* first, uint32_t* count);
int fr_render_sharded(fr_renderer* r, const float* camera_transform12, float fov, float F, float focus,
                      const float* bg_color3, const fr_layers* layers_dev, uint32_t total_spp, uint32_t max_depth,
                      int root);
int fr_reduce_layers(fr_renderer* r, const fr_layers* layers_dev, uint32_t total_spp, int root);
/* (b) one process, n devices (NULL / 0: all): fredholm::MultiGpuRenderer, one renderer and one host thread per
 *     device.  fr_multi_renderer returns the BORROWED handle of a rank for the scene / light / film calls above
 *     (do not destroy it); fr_multi_render takes device pointers on the first device. */
typedef struct fr_multi fr_multi;
fr_multi* fr_multi_create(const int* devices, int n_devices);
void fr_multi_destroy(fr_multi* m);
int fr_multi_size(fr_multi* m);
fr_renderer* fr_multi_renderer(fr_multi* m, int rank);
int fr_multi_render(fr_multi* m, const float* camera_transform12, float fov, float F, float focus,
                    const float* bg_color3, const fr_layers* layers_dev_rank0, uint32_t total_spp, uint32_t max_depth);
int fr_multi_wait(fr_multi* m);
/* out6 = paths, radiance rays, shadow rays, light rays, kernel launches, zero-contribution rays not traced
 * (traced + not traced = the reference's trace-call count for the same samples) */
int fr_get_statistics(fr_renderer* r, uint64_t* out6);
int fr_reset_statistics(fr_renderer* r);
/* measurement: counting instantiations of the traversal kernels (process-wide switch); out6 = CWBVH nodes
 * visited by radiance / shadow / MIS rays, then triangles tested by the same three, since the last reset */
int fr_set_traversal_counting(fr_renderer* r, int on);
int fr_get_traversal_counters(fr_renderer* r, uint64_t* out6);
/* samples of one pixel block that share a warp (1, 2, 4, 8, 16, 32; csrc/wavefront.h FilmGeom) */
int fr_set_samples_per_warp(fr_renderer* r, uint32_t spw);
/* per-stage device time from CUDA events on the renderer's stream; 7 stages:
 * generate, trace_closest, shade, trace_shadow, trace_light, advance, film.
 * fr_get_stage_times synchronises, returns the accumulated ms / launch counts and clears them */
int fr_set_stage_timing(fr_renderer* r, int on);
int fr_get_stage_times(fr_renderer* r, double* ms7, uint64_t* launches7);
/* CUDA events on the renderer's stream (timing on the launching stream) */
void* fr_event_create(void);
int fr_event_destroy(void* ev);
int fr_event_record(fr_renderer* r, void* ev);
int fr_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms); /* synchronises on ev_stop */
/* CUDA stream of the renderer as an integer handle (cudaStream_t) */
uint64_t fr_get_stream(fr_renderer* r);

/* ---- post-process: kernels/post-process.cu:5-47 (device pointers, float4 per pixel) ---- */
int fr_post_process(const void* beauty_in_dev, void* high_luminance_dev, void* temp_dev, int width, int height,
                    const fr_post_process_params* params, void* beauty_out_dev);
int fr_tone_mapping(const void* beauty_in_dev, int width, int height, float ISO, float chromatic_aberration,
                    void* beauty_out_dev);

/* ---- denoise stage: fredholm::Denoiser (fredholm/include/fredholm/denoiser.h:14-145: ctor with the
 * beauty / normal / albedo / output device pointers, denoise(), wait_for_completion()).  The OptiX AI
 * denoiser is replaced by an albedo/normal-guided a-trous wavelet filter (include/fredholm/denoiser.h).
 * One-shot form: construct, denoise, wait.  Output is width x height float4 (2x each with upscale).
 * iterations <= 0 or sigmas / floor <= 0 select the defaults (5, 0.5, 0.1, 0.01); firefly_k < 0 selects
 * the default (2), 0 switches the firefly clamp off. */
int fr_denoise(const void* beauty_dev, const void* normal_dev, const void* albedo_dev, void* denoised_dev,
               uint32_t width, uint32_t height, int upscale, int iterations, float sigma_color, float sigma_albedo,
               float albedo_floor, float firefly_k);

/* ---- multi-frame batch: the render / save loop of app/rtcamp8.cpp:47-303 as a library call
 * (include/fredholm/batch.h).  Per frame: clear layers, init_render_states, set_time(t), render,
 * denoise, post-process, RGBA8 conversion, read-back, "<output_dir>/<frame>.png". */
typedef struct fr_batch_config {
  uint32_t width, height, n_spp, max_depth;   /* rtcamp8.cpp:47-54 */
  fr_post_process_params post;                /* rtcamp8.cpp:55-58,201-206 */
  int denoise, upscale;                       /* rtcamp8.cpp:48,113-118 */
  int dn_iterations;                          /* <= 0: default */
  float dn_sigma_color, dn_sigma_albedo, dn_albedo_floor, dn_firefly_k; /* as for fr_denoise */
  float fps, start_time, max_time, kill_time_s; /* rtcamp8.cpp:59-62 */
  uint32_t first_frame, frame_stride, max_frames; /* multi-GPU: rank, world size */
  float bg_color[3];
  int animate;                                /* call set_time every frame */
  const char* output_dir;                     /* NULL or "": no files */
  uint32_t n_save_threads, n_slots;
} fr_batch_config;

typedef struct fr_frame_record {
  uint32_t frame_idx;
  float time, accel_ms, render_ms, denoise_ms, post_ms, transfer_ms, encode_ms, save_ms;
  uint64_t png_bytes;
} fr_frame_record;

/* camera: one 3x4 block (n_camera_frames = 0) or a path of n_camera_frames blocks indexed by
 * min(frame_idx, n_camera_frames - 1).  records: room for max_records entries; frames_rgba8 (optional):
 * max_records x out_w x out_h x 4 bytes, filled in record order.  out5 = n_records, out_w, out_h, killed, 0. */
int fr_batch_run(fr_renderer* r, const fr_batch_config* config, const float* camera_transforms12,
                 uint32_t n_camera_frames, float fov, float F, float focus, fr_frame_record* records,
                 uint32_t max_records, uint8_t* frames_rgba8, uint32_t* out5, double* wall_s);

/* ---- raw device memory helpers (cwl::CUDABuffer, buffer.h:18-85) ---- */
void* fr_device_alloc(size_t bytes);
int fr_device_free(void* p);
int fr_device_memset(void* p, int value, size_t bytes);
int fr_copy_to_device(void* dst_dev, const void* src_host, size_t bytes);
int fr_copy_to_host(void* dst_host, const void* src_dev, size_t bytes);
int fr_device_synchronize(void);
/* page-locked host memory for the read-back path */
void* fr_host_alloc_pinned(size_t bytes);
int fr_host_free_pinned(void* p);

/* ---- image files on the boundary (host only, no device needed) ----
 * fr_image8_load  = what fredholm::Texture(path, type) holds (fredholm/src/scene.cpp:7-37: stbi_load,
 *                   4 channels, vertical flip -> row 0 is the BOTTOM row): PNG and JPEG
 * fr_imagef_load  = what fredholm::FloatTexture(path) holds (scene.cpp:39-67: stbi_loadf, no flip):
 *                   Radiance .hdr, or an 8-bit file through the gamma-2.2 rule
 * The decoded image stays in a per-thread slot until the matching *_copy call.
 * fr_write_png    = stbi_write_png(path, w, h, channels, data, w * channels) of the frame savers
 *                   (app/controller.cpp:291-308, app/rtcamp8.cpp:286-293); channels 3 or 4 */
int fr_image8_load(const char* path, uint32_t* width, uint32_t* height);
int fr_image8_copy(uint8_t* rgba8);
int fr_imagef_load(const char* path, uint32_t* width, uint32_t* height);
int fr_imagef_copy(float* rgba32f);
int fr_write_png(const char* path, const uint8_t* pixels, uint32_t width, uint32_t height, uint32_t channels);

/* ---- stage-level entry points used by the parity tests ---- */
/* closest hit for n rays (6 floats each: origin, direction): out_id = (instance, primitive)
 * or 0xffffffff, out_tuv = (t, u, v); counters2 (optional) = nodes visited, triangles tested */
int fr_trace_closest(fr_renderer* r, const float* rays, uint32_t n, float tmin, float tmax, uint32_t* out_id,
                     float* out_tuv, uint64_t* counters2);
int fr_primary_rays(fr_renderer* r, const float* camera_transform12, float fov, float F, float focus, uint32_t n_spp,
                    float* out_rays);
int fr_sampler_sequence(uint32_t width, uint32_t height, uint32_t seed, uint32_t image_idx, uint32_t n_spp,
                        const char* kinds, float* out);
/* in: n x 40 floats (ShadingParams[30], wo[3], entering, wi[3], u, v[2]); out: n x 11 floats */
int fr_bsdf_eval_sample(const float* in, uint32_t n, float* out);
int fr_sky_radiance(fr_renderer* r, const float* dirs, uint32_t n, float* out);
int fr_arhosek_cook(float turbidity, float albedo, float elevation, float* out30);
/* fredholm::Camera mirror (camera.h:51-135) */
int fr_camera_transform(const float* origin3, float* out12);
int fr_camera_walk(const float* origin3, float d_phi, float d_theta, int movement, float dt, float* out12);

#ifdef __cplusplus
}
#endif
#endif /* FREDHOLM_B200_H */
