// Data layout of the wavefront integrator: scene view, per-path state and the
// ray queues that live in HBM between stages.
//
// One "wave" processes n_samples x (tiled) pixels paths at once.  Path state is
// addressed by a fixed path slot (never moved); the queues hold slot indices or
// compact 48-byte ray records.  All records are 16-byte multiples so that even
// gathered accesses move whole 32-byte sectors.
#pragma once
#include <cstdint>

#include "bvh.cuh"
#include "camera_sky.cuh"
#include "fredholm/shared.h"

namespace frd
{

struct TexView {
  const uchar4* texels;  // row-major, row 0 first (as uploaded by the loader)
  uint32_t width, height;
  uint32_t srgb;  // decode sRGB -> linear on fetch (COLOR textures)
  uint32_t pad_;
};

struct SceneView {
  const float3* vertices;
  const float3* normals;
  const float2* texcoords;
  const uint3* indices;            // global vertex ids per face
  const uint32_t* material_ids;    // per face
  const uint32_t* face_submesh;    // per face: submesh == instance == transform index
  const uint8_t* face_class;       // per face: ShadeClass of its material
  const fredholm::Material* materials;
  const TexView* textures;
  const float* srgb_lut;  // 256-entry sRGB -> linear table
  const fredholm::Matrix3x4* o2w;
  const fredholm::Matrix3x4* w2o;
  const fredholm::AreaLight* lights;
  uint32_t n_lights;
  BvhView bvh;
  float3 bounds_lo, bounds_hi;  // world bounds of the geometry (coherence-sort grid)

  // lights & sky
  int has_dir_light;
  fredholm::DirectionalLight dir_light;
  float3 dir_t, dir_b;    // basis around dir_light.dir
  float dir_disk_radius;  // 1e9 * tan(angle/2)
  int sky_mode;           // SkyMode
  float sky_intensity;
  float3 bg_color;
  float3 sun_dir;
  HosekSky hosek;
  const float4* ibl_texels;  // lat-long float4 image
  uint32_t ibl_width, ibl_height;
};

// 48-byte shadow (visibility) ray: adds `contrib` to the path's radiance if nothing
// is hit in (0, tmax).
struct alignas(16) ShadowRay {
  float ox, oy, oz, tmax;
  float dx, dy, dz;
  uint32_t path;
  float cr, cg, cb;
  uint32_t pad_;
};
static_assert(sizeof(ShadowRay) == 48, "shadow ray record");

// 48-byte MIS ray (closest hit): radiance += clamp(w * mis) * Le where Le / the light
// pdf come from whatever the ray reaches (emitter, non-emitter or sky).
struct alignas(16) LightRay {
  float ox, oy, oz, pdf_bsdf;
  float dx, dy, dz;
  uint32_t path;
  float wr, wg, wb;  // throughput * f * |cos| / pdf_bsdf
  float cos_wi;      // |cos| of the sampled direction in the shading frame
};
static_assert(sizeof(LightRay) == 48, "light ray record");

enum QueueId : int { Q_CUR = 0, Q_NEXT, Q_SHADOW0, Q_SHADOW1, Q_SHADOW2, Q_LIGHT, Q_COUNT };

// Material classes of the shade stage.  After closest-hit traversal every path is
// appended to the queue of the class of the material it hit ("sorting by material"):
// each class has its own kernel instantiation carrying only the lobes the class needs.
enum ShadeClass : int {
  CLS_DIFFUSE = 0,   // Oren-Nayar / Lambert only
  CLS_PLASTIC,       // specular + diffuse
  CLS_METAL,         // metalness == 1
  CLS_COATED,        // coat + specular + diffuse
  CLS_GLASS,         // specular + transmission + diffuse
  CLS_SHEEN,         // sheen + diffuse
  CLS_GENERIC,       // every lobe, no textures
  CLS_GENERIC_TEX,   // every lobe, textured inputs / bump / normal map
  CLS_MISS,          // camera rays that left the scene (sky)
  CLS_COUNT
};

// device-resident control block: queue sizes, work cursors, statistics
struct WaveControl {
  uint32_t n[Q_COUNT];
  uint32_t cursor[8];
  uint32_t n_class[CLS_COUNT];
  uint32_t cursor_class[CLS_COUNT];
  unsigned long long rays_closest, rays_shadow, rays_light;  // traced rays
  unsigned long long nodes[3], tris[3];  // counting builds only: per ray type (closest, shadow, MIS)
  unsigned long long paths;
  unsigned long long rays_skipped;  // zero-contribution visibility / MIS rays that were not traced (shade.cu)
};

struct WaveBuffers {
  float4* ray_o;  // [n_slots] xyz origin
  float4* ray_d;  // [n_slots] xyz direction
  float4* hit;    // [n_slots] t, u, v, face bits
  float4* thr;    // [n_slots] throughput xyz, w = cmj_draws | sobol_dim << 16
  float4* L;      // [n_slots] radiance accumulator
  float4* aov0;   // [n_slots] position xyz, depth
  float4* aov1;   // [n_slots] normal xyz, texcoord.x
  float4* aov2;   // [n_slots] albedo xyz, texcoord.y
  uint32_t* queue[2];       // [n_slots] path slots of live radiance rays (ping-pong)
  uint32_t* class_queue[CLS_COUNT];  // [n_slots] each: paths to shade, by material class
  ShadowRay* shadow[3];     // [n_slots] directional / sky / area-light NEE rays
  LightRay* light;          // [n_slots] MIS rays
  WaveControl* ctl;
  // single-launch mode (reference quirk: RadiancePayload outlives the sample loop, pt.cu:432-433)
  uint32_t* first_hit;  // [n_pixels] index of the first sample of this launch whose camera ray hit geometry
  float4* pix_aov0;     // [n_pixels] that sample's first-hit words (position|depth, normal|u, albedo|v)
  float4* pix_aov1;
  float4* pix_aov2;
  // Straggler set (integrator.cpp, wave compaction): its slots are handed out densely to the paths that several
  // waves still had alive after a few bounces, so a slot no longer says which pixel / sample it is.  origin[slot] =
  // wave index * origin_stride + the slot the path had in its own wave; that wave started origin_samples samples
  // per wave index after the first one.  nullptr in an ordinary wave.
  uint32_t* origin;
  uint32_t origin_stride;
  uint32_t origin_samples;
};


// image <-> path slot mapping.  A warp owns 32 consecutive path slots = a small pixel block times
// `spw` consecutive samples of it (spw = 2^spw_log2 samples per warp, block = 32 / spw pixels):
//   spw  1: 8x4 pixels x 1 sample     spw  8: 2x2 pixels x  8 samples
//   spw  2: 4x4 pixels x 2 samples    spw 16: 2x1 pixels x 16 samples
//   spw  4: 4x2 pixels x 4 samples    spw 32: 1   pixel  x 32 samples
// Camera rays (and the first-bounce sun rays) of one warp then form a beam about one pixel wide:
// their traversals take the same decisions, so the per-lane state machines of trace_queue() stay in
// lock step (node and triangle phases run with most lanes active).  The samples of a wave are
// handed out in groups of spw; the slots of a group beyond n_samples stay empty.
struct FilmGeom {
  uint32_t width, height;
  uint32_t tiles_x, tiles_y;
  uint32_t slots_per_group;  // tiles_x * tiles_y * 32: one group = spw samples of every pixel
  uint32_t spw_log2;         // log2(samples per warp)
  uint32_t bw_log2, bh_log2;  // log2 of the pixel block a warp covers
};

__host__ __device__ inline FilmGeom make_film_geom(uint32_t w, uint32_t h, uint32_t spw_log2 = 0)
{
  FilmGeom g;
  g.width = w;
  g.height = h;
  g.spw_log2 = spw_log2 > 5u ? 5u : spw_log2;
  const uint32_t pix_log2 = 5u - g.spw_log2;  // pixels per warp: 8x4, 4x4, 4x2, 2x2, 2x1, 1x1
  g.bw_log2 = (pix_log2 + 1u) >> 1;
  g.bh_log2 = pix_log2 >> 1;
  g.tiles_x = (w + (1u << g.bw_log2) - 1u) >> g.bw_log2;
  g.tiles_y = (h + (1u << g.bh_log2) - 1u) >> g.bh_log2;
  g.slots_per_group = g.tiles_x * g.tiles_y * 32u;
  return g;
}
// sample groups (of spw samples) needed for n samples, and the slots they occupy
__host__ __device__ inline uint32_t film_groups(const FilmGeom& g, uint32_t n_samples)
{
  return (n_samples + (1u << g.spw_log2) - 1u) >> g.spw_log2;
}
// slot -> pixel and sample index within the wave; false if the slot lies outside the image
__host__ __device__ inline bool slot_to_pixel(const FilmGeom& g, uint32_t slot, uint32_t& x, uint32_t& y, uint32_t& s)
{
  const uint32_t group = slot / g.slots_per_group, rem = slot - group * g.slots_per_group;
  const uint32_t tile = rem >> 5, lane = rem & 31u;
  const uint32_t pix_log2 = 5u - g.spw_log2;
  const uint32_t p = lane & ((1u << pix_log2) - 1u);
  s = (group << g.spw_log2) + (lane >> pix_log2);
  const uint32_t ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
  x = (tx << g.bw_log2) + (p & ((1u << g.bw_log2) - 1u));
  y = (ty << g.bh_log2) + (p >> g.bw_log2);
  return x < g.width && y < g.height;
}
__host__ __device__ inline uint32_t pixel_to_slot(const FilmGeom& g, uint32_t x, uint32_t y, uint32_t s)
{
  const uint32_t pix_log2 = 5u - g.spw_log2;
  const uint32_t group = s >> g.spw_log2, s_in = s & ((1u << g.spw_log2) - 1u);
  const uint32_t tile = (y >> g.bh_log2) * g.tiles_x + (x >> g.bw_log2);
  const uint32_t p = ((y & ((1u << g.bh_log2) - 1u)) << g.bw_log2) | (x & ((1u << g.bw_log2) - 1u));
  return group * g.slots_per_group + tile * 32u + (s_in << pix_log2) + p;
}

// Coherence sort (sort.cu): which queue, and the grid the ray origins are binned on.
enum SortQueue : int { SORT_RADIANCE0 = 0, SORT_RADIANCE1, SORT_SHADOW0, SORT_SHADOW1, SORT_SHADOW2, SORT_LIGHT };
struct SortGrid {
  float3 lo;          // scene bounds, lower corner
  float3 inv_cell;    // cells per world unit, per axis
  uint32_t cell_bits;  // 2^cell_bits cells per axis (<= 10)
  uint32_t use_octant;  // append the direction octant to the key
};
__host__ __device__ inline uint32_t sort_bins(const SortGrid& g) { return 1u << (3u * g.cell_bits + (g.use_octant ? 3u : 0u)); }

// which pixel and sample (relative to WaveParams::sample_base) a path slot of `wb` belongs to
__host__ __device__ inline void path_identity(const FilmGeom& g, const WaveBuffers& wb, uint32_t slot, uint32_t& x, uint32_t& y,
                                     uint32_t& sample)
{
  uint32_t first = 0;
  if (wb.origin) {
    const uint32_t o = wb.origin[slot];
    const uint32_t wave = o / wb.origin_stride;
    slot = o - wave * wb.origin_stride;
    first = wave * wb.origin_samples;
  }
  slot_to_pixel(g, slot, x, y, sample);
  sample += first;
}

struct WaveParams {
  FilmGeom film;
  uint32_t n_samples;    // samples in this wave
  uint32_t sample_base;  // sample index of the first one (reference: sample_count)
  uint32_t max_depth;
  uint32_t seed;
  uint32_t want_aov;     // any first-hit layer (position / normal / depth / texcoord / albedo) is bound
  uint32_t single_launch;  // 1: all samples of this render() call share one payload like ONE reference launch
  fredholm::CameraParams camera;
};

}  // namespace frd
