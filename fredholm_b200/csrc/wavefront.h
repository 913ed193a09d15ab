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
  unsigned long long nodes_visited, tris_tested;             // only in stats builds
  unsigned long long paths;
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
};

// image <-> path slot mapping: pixels are walked in 8x4 tiles so that one warp owns
// a compact screen tile (coherent primary rays).
struct FilmGeom {
  uint32_t width, height;
  uint32_t tiles_x, tiles_y;
  uint32_t slots_per_sample;  // tiles_x * tiles_y * 32
};

__host__ __device__ inline FilmGeom make_film_geom(uint32_t w, uint32_t h)
{
  FilmGeom g;
  g.width = w;
  g.height = h;
  g.tiles_x = (w + 7) / 8;
  g.tiles_y = (h + 3) / 4;
  g.slots_per_sample = g.tiles_x * g.tiles_y * 32u;
  return g;
}
__host__ __device__ inline bool slot_to_pixel(const FilmGeom& g, uint32_t slot_in_sample, uint32_t& x, uint32_t& y)
{
  const uint32_t tile = slot_in_sample >> 5, lane = slot_in_sample & 31u;
  x = (tile % g.tiles_x) * 8u + (lane & 7u);
  y = (tile / g.tiles_x) * 4u + (lane >> 3);
  return x < g.width && y < g.height;
}
__host__ __device__ inline uint32_t pixel_to_slot(const FilmGeom& g, uint32_t x, uint32_t y)
{
  return ((y >> 2) * g.tiles_x + (x >> 3)) * 32u + ((y & 3u) << 3) + (x & 7u);
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
