// =============================================================================
// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// This translation unit is the CPU oracle for the fredholm path-tracing core.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the library built from it.  The product
// (fredholm_b200/) never includes, links or calls anything under oracle/.
//
// What it is: the reference's own integrator sources
//   /root/reference/fredholm/modules/pt.cu  (+ bsdf.cu, bxdf.cu, lut.cu,
//   sampling.cu, cmj.cu, sobol.cu, camera.cu, arhosek.cu, math.cu, shared.h)
// compiled VERBATIM for the host; the only edit is the
// mandatory one-token patch of pt.cu:181 applied by the Makefile to a build-dir
// copy.  This file supplies ONLY what OptiX / CUDA supplied to that code:
//   (i)   optixTrace: a CPU BVH + watertight ray/triangle test (Woop et al.
//         2013) + any-hit / closest-hit / miss program dispatch by ray type
//         (SBT layout of renderer.h:306-327, 519-521),
//   (ii)  tex2D emulation (cwl/texture.h:35-47: wrap, bilinear, normalized,
//         uchar4 -> [0,1], sRGB decode for COLOR textures),
//   (iii) the optixLaunch loop over pixels with `params` as a host global
//         (pt.cu:15-17),
//   (iv)  the scene-upload logic of Renderer::load_scene (light list, 3x4
//         matrices; renderer.h:388-421) and Renderer::render (renderer.h:657-734).
//
// Parity status: the reference ships no tests or golden vectors,
// so parity is pinned to the reference SOURCE (this build), not to reference
// test fixtures.  At the OptiX boundary (traversal) the spec is the OptiX
// programming-guide semantics restated in (i): "parity unpinned" there.
// =============================================================================

// ---- host stand-ins for the CUDA qualifiers (SURVEY.md 8(c) recipe) ----
#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline
#define __align__(x) alignas(x)

#include <sys/types.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
using std::isinf;
using std::isnan;

#include <optix.h>  // oracle/shim/optix.h

// Device semantics of clamp(x, 0.0f, 1.0f): nvcc lowers the reference's fmaxf(0, fminf(x, 1)) to the
// .sat modifier, which maps NaN to 0 (the host expression maps NaN to 1).  Used by the patched copies
// of pt.cu:375 (regularize_weight) and pt.cu:460 (roulette probability); see oracle/Makefile.
static inline float orc_sat(float x) { return x != x ? 0.0f : fmaxf(0.0f, fminf(x, 1.0f)); }
static inline float3 orc_sat3(const float3& v) { return make_float3(orc_sat(v.x), orc_sat(v.y), orc_sat(v.z)); }

// the reference integrator, verbatim (build-dir copy with the patches listed in oracle/Makefile)
#include "patched/pt.cu"

// glm only for glm::inverse, to mirror renderer.h:404-421 bit-for-bit
#include "glm/glm.hpp"

namespace
{

// -----------------------------------------------------------------------------
// host textures (stand-in for cwl::CUDATexture, cwl/texture.h:13-74)
// -----------------------------------------------------------------------------
struct HostTexture {
  int w = 0, h = 0;
  bool srgb = false;
  bool is_float = false;
  std::vector<uchar4> data8;
  std::vector<float4> dataf;
};

float srgb8_to_linear(unsigned char c)
{
  // IEC 61966-2-1 decode, evaluated in double then rounded (matches the
  // product's 256-entry table, which is built with the same expression).
  const double v = c / 255.0;
  const double l = v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4);
  return static_cast<float>(l);
}

inline float4 texel(const HostTexture* t, int i, int j)
{
  if (t->is_float) return t->dataf[i + t->w * j];
  const uchar4 c = t->data8[i + t->w * j];
  if (t->srgb) {
    return make_float4(srgb8_to_linear(c.x), srgb8_to_linear(c.y),
                       srgb8_to_linear(c.z), c.w / 255.0f);
  }
  return make_float4(c.x / 255.0f, c.y / 255.0f, c.z / 255.0f, c.w / 255.0f);
}

inline int wrapi(int i, int n)
{
  i %= n;
  return i < 0 ? i + n : i;
}

// -----------------------------------------------------------------------------
// ray / triangle: watertight test (Woop, Benthin, Wald 2013).  THIS DEFINES THE
// ARITHMETIC the product's device code must reproduce bit-for-bit:
//   shear + T use fused multiply-add (std::fmaf / __fmaf_rn), the edge functions
//   U,V,W use separately rounded products (watertightness needs antisymmetry),
//   exact-zero edge functions fall back to double precision.
// This file is compiled with -ffp-contract=off so nothing else gets fused.
// -----------------------------------------------------------------------------
struct RaySetup {
  float ox, oy, oz;
  float cx[3], cy[3], cz[3];  // shear rows: Ax = dot(A, cx) etc.
};

inline RaySetup make_ray_setup(const float3& o, const float3& d)
{
  const float dv[3] = {d.x, d.y, d.z};
  int kz = 0;
  if (std::fabs(dv[1]) > std::fabs(dv[kz])) kz = 1;
  if (std::fabs(dv[2]) > std::fabs(dv[kz])) kz = 2;
  int kx = (kz + 1) % 3;
  int ky = (kx + 1) % 3;
  if (dv[kz] < 0.0f) std::swap(kx, ky);
  const float Sx = dv[kx] / dv[kz];
  const float Sy = dv[ky] / dv[kz];
  const float Sz = 1.0f / dv[kz];
  RaySetup r;
  r.ox = o.x;
  r.oy = o.y;
  r.oz = o.z;
  for (int i = 0; i < 3; ++i) {
    r.cx[i] = (i == kx) ? 1.0f : (i == kz) ? -Sx : 0.0f;
    r.cy[i] = (i == ky) ? 1.0f : (i == kz) ? -Sy : 0.0f;
    r.cz[i] = (i == kz) ? Sz : 0.0f;
  }
  return r;
}

inline float shear(const float a[3], const float c[3])
{
  return std::fmaf(a[2], c[2], std::fmaf(a[1], c[1], a[0] * c[0]));
}

// returns true and (t,u,v) if tmin < t < tmax
inline bool intersect_tri(const RaySetup& r, const float3& v0, const float3& v1,
                          const float3& v2, float tmin, float tmax, float& t,
                          float& bu, float& bv)
{
  const float A[3] = {v0.x - r.ox, v0.y - r.oy, v0.z - r.oz};
  const float B[3] = {v1.x - r.ox, v1.y - r.oy, v1.z - r.oz};
  const float C[3] = {v2.x - r.ox, v2.y - r.oy, v2.z - r.oz};
  const float Ax = shear(A, r.cx), Ay = shear(A, r.cy);
  const float Bx = shear(B, r.cx), By = shear(B, r.cy);
  const float Cx = shear(C, r.cx), Cy = shear(C, r.cy);
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    const double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
    U = (float)(CxBy - CyBx);
    const double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
    V = (float)(AxCy - AyCx);
    const double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
    W = (float)(BxAy - ByAx);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
    return false;
  const float det = (U + V) + W;
  if (det == 0.0f) return false;
  const float Az = shear(A, r.cz), Bz = shear(B, r.cz), Cz = shear(C, r.cz);
  const float T = std::fmaf(W, Cz, std::fmaf(V, Bz, U * Az));
  const float rcp = 1.0f / det;
  const float tt = T * rcp;
  if (!(tt > tmin && tt < tmax)) return false;
  t = tt;
  bu = V * rcp;
  bv = W * rcp;
  return true;
}

// -----------------------------------------------------------------------------
// CPU BVH2 (binned SAH) over world-space triangles.  Correctness only; the box
// test is padded so that it can never cull a triangle the watertight test
// would accept.
// -----------------------------------------------------------------------------
struct Box {
  float lo[3], hi[3];
  void reset()
  {
    for (int i = 0; i < 3; ++i) {
      lo[i] = 3.0e38f;
      hi[i] = -3.0e38f;
    }
  }
  void grow(const float3& p)
  {
    const float v[3] = {p.x, p.y, p.z};
    for (int i = 0; i < 3; ++i) {
      lo[i] = std::min(lo[i], v[i]);
      hi[i] = std::max(hi[i], v[i]);
    }
  }
  void grow(const Box& b)
  {
    for (int i = 0; i < 3; ++i) {
      lo[i] = std::min(lo[i], b.lo[i]);
      hi[i] = std::max(hi[i], b.hi[i]);
    }
  }
  float half_area() const
  {
    const float e[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    if (e[0] < 0) return 0.0f;
    return e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
  }
};

struct BvhNode {
  Box box;
  uint left;   // internal: left child (right = left + 1); leaf: first prim slot
  uint count;  // 0 => internal
};

struct CpuBvh {
  std::vector<BvhNode> nodes;
  std::vector<uint> prim;  // leaf order -> global face index

  void build(const std::vector<float3>& wv)
  {
    const uint n = wv.size() / 3;
    prim.resize(n);
    std::vector<Box> pb(n);
    std::vector<float> cen(3 * (size_t)n);
    for (uint i = 0; i < n; ++i) {
      prim[i] = i;
      pb[i].reset();
      for (int k = 0; k < 3; ++k) pb[i].grow(wv[3 * (size_t)i + k]);
      for (int a = 0; a < 3; ++a)
        cen[3 * (size_t)i + a] = 0.5f * (pb[i].lo[a] + pb[i].hi[a]);
    }
    nodes.clear();
    nodes.reserve(2 * (size_t)n + 1);
    nodes.push_back(BvhNode{});
    if (n == 0) {
      nodes[0].box.reset();
      nodes[0].left = 0;
      nodes[0].count = 0;
      return;
    }
    struct Task {
      uint node, first, count;
    };
    std::vector<Task> stack;
    stack.push_back({0, 0, n});
    constexpr int NB = 16;
    while (!stack.empty()) {
      const Task tk = stack.back();
      stack.pop_back();
      Box nb, cb;
      nb.reset();
      cb.reset();
      for (uint i = tk.first; i < tk.first + tk.count; ++i) {
        const uint p = prim[i];
        nb.grow(pb[p]);
        cb.grow(make_float3(cen[3 * (size_t)p], cen[3 * (size_t)p + 1],
                            cen[3 * (size_t)p + 2]));
      }
      nodes[tk.node].box = nb;
      if (tk.count <= 4) {
        nodes[tk.node].left = tk.first;
        nodes[tk.node].count = tk.count;
        continue;
      }
      // binned SAH over the three axes
      int best_axis = -1, best_bin = -1;
      float best_cost = 3.0e38f;
      for (int a = 0; a < 3; ++a) {
        const float ext = cb.hi[a] - cb.lo[a];
        if (!(ext > 0.0f)) continue;
        Box bb[NB];
        uint bc[NB];
        for (int b = 0; b < NB; ++b) {
          bb[b].reset();
          bc[b] = 0;
        }
        const float scale = NB / ext;
        for (uint i = tk.first; i < tk.first + tk.count; ++i) {
          const uint p = prim[i];
          int b = (int)((cen[3 * (size_t)p + a] - cb.lo[a]) * scale);
          b = std::min(std::max(b, 0), NB - 1);
          bb[b].grow(pb[p]);
          bc[b]++;
        }
        float ra[NB];
        uint rc[NB];
        Box acc;
        acc.reset();
        uint cnt = 0;
        for (int b = NB - 1; b > 0; --b) {
          acc.grow(bb[b]);
          cnt += bc[b];
          ra[b] = acc.half_area();
          rc[b] = cnt;
        }
        acc.reset();
        cnt = 0;
        for (int b = 0; b < NB - 1; ++b) {
          acc.grow(bb[b]);
          cnt += bc[b];
          if (cnt == 0 || rc[b + 1] == 0) continue;
          const float cost = acc.half_area() * cnt + ra[b + 1] * rc[b + 1];
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = a;
            best_bin = b;
          }
        }
      }
      uint mid;
      if (best_axis < 0) {
        mid = tk.first + tk.count / 2;  // all centroids coincide: split in half
      } else {
        const int a = best_axis;
        const float scale = NB / (cb.hi[a] - cb.lo[a]);
        uint* b0 = prim.data() + tk.first;
        uint* b1 = b0 + tk.count;
        uint* m = std::partition(b0, b1, [&](uint p) {
          int b = (int)((cen[3 * (size_t)p + a] - cb.lo[a]) * scale);
          b = std::min(std::max(b, 0), NB - 1);
          return b <= best_bin;
        });
        mid = tk.first + (uint)(m - b0);
        if (mid == tk.first || mid == tk.first + tk.count)
          mid = tk.first + tk.count / 2;
      }
      const uint l = nodes.size();
      nodes.push_back(BvhNode{});
      nodes.push_back(BvhNode{});
      nodes[tk.node].left = l;
      nodes[tk.node].count = 0;
      stack.push_back({l, tk.first, mid - tk.first});
      stack.push_back({l + 1, mid, tk.first + tk.count - mid});
    }
  }
};

inline bool hit_box(const Box& b, const float3& o, const float3& inv,
                    float tmin, float tmax)
{
  float t0 = tmin, t1 = tmax;
  const float ov[3] = {o.x, o.y, o.z};
  const float iv[3] = {inv.x, inv.y, inv.z};
  for (int a = 0; a < 3; ++a) {
    float n = (b.lo[a] - ov[a]) * iv[a];
    float f = (b.hi[a] - ov[a]) * iv[a];
    if (n > f) std::swap(n, f);
    // pad generously: the box test must be conservative (NaN => keep)
    n -= 1e-5f * std::fabs(n) + 1e-6f;
    f += 1e-5f * std::fabs(f) + 1e-6f;
    if (n > t0) t0 = n;
    if (f < t1) t1 = f;
    if (t0 > t1) return false;
  }
  return true;
}

// -----------------------------------------------------------------------------
// oracle state
// -----------------------------------------------------------------------------
struct OracleState {
  // scene (host copies of what Renderer::load_scene uploads, renderer.h:361-421)
  std::vector<float3> vertices, normals;
  std::vector<float2> texcoords;
  std::vector<uint3> indices;
  std::vector<uint> material_ids, submesh_offsets, submesh_n_faces, instance_ids;
  std::vector<Material> materials;
  std::vector<std::unique_ptr<HostTexture>> textures;
  std::vector<TextureHeader> texture_headers;
  std::vector<glm::mat4> transforms;
  std::vector<Matrix3x4> o2w, w2o;
  std::vector<AreaLight> lights;
  std::vector<HitGroupSbtRecordData> sbt;  // one per submesh (x3 ray types share data)

  DirectionalLight dir_light;
  bool has_dir_light = false;
  float sky_intensity = 1.0f;
  float3 sun_direction = make_float3(0.0f, 1.0f, 0.0f);
  ArHosekSkyModelState arhosek;
  bool has_arhosek = false;
  std::unique_ptr<HostTexture> ibl;

  // accel
  std::vector<float3> wv;         // world-space triangle vertices, 3 per face
  std::vector<uint> face_submesh;  // global face -> submesh (= OptiX instance)
  CpuBvh bvh;
  bool accel_valid = false;

  // film state (renderer.h:642-655)
  uint width = 0, height = 0;
  std::vector<uint> sample_count;

  // counters
  std::atomic<unsigned long long> n_rays[3];
};

OracleState* S = nullptr;

// per-thread "OptiX" state
struct TraceCtx {
  uint3 launch_index;
  uint p0, p1;
  uint prim, inst;
  float2 bary;
  float3 o, d;
  float tmax;
  const HitGroupSbtRecordData* sbt;
  bool ignore;
};
thread_local TraceCtx g_ctx;
thread_local unsigned long long g_rays[3];

struct Hit {
  float t, u, v;
  uint face;
  bool valid;
};

typedef void (*Program)();
Program g_anyhit[3] = {__anyhit__radiance, __anyhit__shadow, __anyhit__light};
Program g_closesthit[3] = {__closesthit__radiance, __closesthit__shadow,
                           __closesthit__light};
Program g_miss[3] = {__miss__radiance, __miss__shadow, __miss__light};

inline bool face_needs_anyhit(uint face)
{
  const Material& m = S->materials[S->material_ids[face]];
  return m.base_color_texture_id >= 0 || m.alpha_texture_id >= 0;
}

// runs the reference's any-hit program for a candidate; returns true if the
// intersection is ACCEPTED
inline bool run_anyhit(uint raytype, uint face, float t, float u, float v)
{
  if (!face_needs_anyhit(face)) return true;  // program body is a no-op then
  const uint sm = S->face_submesh[face];
  g_ctx.prim = face - S->submesh_offsets[sm];
  g_ctx.inst = sm;
  g_ctx.bary = make_float2(u, v);
  g_ctx.tmax = t;
  g_ctx.sbt = &S->sbt[sm];
  g_ctx.ignore = false;
  g_anyhit[raytype]();
  return !g_ctx.ignore;
}

// closest (or first, if `any`) accepted hit.  Tie rule: on exactly equal t the
// lower global face index wins.
Hit traverse(const float3& o, const float3& d, float tmin, float tmax,
             bool any, uint raytype, bool with_anyhit)
{
  Hit best;
  best.valid = false;
  best.t = tmax;
  best.face = 0xffffffffu;
  best.u = best.v = 0.0f;
  if (S->bvh.prim.empty()) return best;
  const RaySetup rs = make_ray_setup(o, d);
  const float3 inv = make_float3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  uint stack[256];
  int sp = 0;
  stack[sp++] = 0;
  const std::vector<BvhNode>& nodes = S->bvh.nodes;
  while (sp > 0) {
    const BvhNode& nd = nodes[stack[--sp]];
    // "<= best.t" semantics: hit_box is padded, so ties are never culled
    if (!hit_box(nd.box, o, inv, tmin, best.t)) continue;
    if (nd.count == 0) {
      if (sp + 2 > 256) {
        std::fprintf(stderr, "oracle: traversal stack overflow\n");
        std::abort();
      }
      stack[sp++] = nd.left;
      stack[sp++] = nd.left + 1;
      continue;
    }
    for (uint i = nd.left; i < nd.left + nd.count; ++i) {
      const uint face = S->bvh.prim[i];
      float t, u, v;
      // allow t == best.t so the tie rule can be applied
      const float tlim = best.valid ? std::nextafter(best.t, 3.0e38f) : tmax;
      if (!intersect_tri(rs, S->wv[3 * (size_t)face], S->wv[3 * (size_t)face + 1],
                         S->wv[3 * (size_t)face + 2], tmin, tlim, t, u, v))
        continue;
      if (best.valid && t == best.t && face > best.face) continue;
      if (with_anyhit && !run_anyhit(raytype, face, t, u, v)) continue;
      best.valid = true;
      best.t = t;
      best.u = u;
      best.v = v;
      best.face = face;
      if (any) return best;
    }
  }
  return best;
}

}  // namespace

// =============================================================================
// shim implementation (what OptiX provided to pt.cu)
// =============================================================================
void optixTrace(OptixTraversableHandle, float3 o, float3 d, float tmin,
                float tmax, float, OptixVisibilityMask, unsigned int rayFlags,
                unsigned int SBToffset, unsigned int, unsigned int missSBTIndex,
                unsigned int& p0, unsigned int& p1)
{
  const TraceCtx saved = g_ctx;
  const uint raytype = SBToffset;
  g_rays[raytype]++;
  g_ctx.p0 = p0;
  g_ctx.p1 = p1;
  g_ctx.o = o;
  g_ctx.d = d;
  const bool any = (rayFlags & OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT) != 0;
  const Hit h = traverse(o, d, tmin, tmax, any, raytype, true);
  if (h.valid) {
    const uint sm = S->face_submesh[h.face];
    g_ctx.prim = h.face - S->submesh_offsets[sm];
    g_ctx.inst = sm;
    g_ctx.bary = make_float2(h.u, h.v);
    g_ctx.tmax = h.t;
    g_ctx.sbt = &S->sbt[sm];
    g_closesthit[raytype]();
  } else {
    g_ctx.tmax = tmax;
    g_miss[missSBTIndex]();
  }
  g_ctx = saved;
}

uint3 optixGetLaunchIndex() { return g_ctx.launch_index; }
uint3 optixGetLaunchDimensions() { return make_uint3(S->width, S->height, 1); }
unsigned int optixGetPayload_0() { return g_ctx.p0; }
unsigned int optixGetPayload_1() { return g_ctx.p1; }
unsigned long long optixGetSbtDataPointer()
{
  return reinterpret_cast<unsigned long long>(g_ctx.sbt);
}
unsigned int optixGetPrimitiveIndex() { return g_ctx.prim; }
unsigned int optixGetInstanceIndex() { return g_ctx.inst; }
float2 optixGetTriangleBarycentrics() { return g_ctx.bary; }
float3 optixGetWorldRayOrigin() { return g_ctx.o; }
float3 optixGetWorldRayDirection() { return g_ctx.d; }
float optixGetRayTmax() { return g_ctx.tmax; }
void optixIgnoreIntersection() { g_ctx.ignore = true; }

// CUDA texture fetch, linear filtering, normalized coordinates, wrap addressing
// (CUDA C++ Programming Guide, "Texture Fetching": xB = N*x - 0.5, i = floor(xB),
// alpha = frac(xB)).  Weights are kept in full fp32 (the hardware quantises them
// to 8 fractional bits; the product filters in software with fp32 weights too --
// see DESIGN.md "textures").
template <>
float4 tex2D<float4>(cudaTextureObject_t tex, float x, float y)
{
  const HostTexture* t = reinterpret_cast<const HostTexture*>(tex);
  const float xb = x * t->w - 0.5f;
  const float yb = y * t->h - 0.5f;
  const float fx = std::floor(xb), fy = std::floor(yb);
  const float a = xb - fx, b = yb - fy;
  const int i0 = wrapi((int)fx, t->w), i1 = wrapi((int)fx + 1, t->w);
  const int j0 = wrapi((int)fy, t->h), j1 = wrapi((int)fy + 1, t->h);
  const float4 t00 = texel(t, i0, j0), t10 = texel(t, i1, j0);
  const float4 t01 = texel(t, i0, j1), t11 = texel(t, i1, j1);
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b);
  const float w01 = (1.0f - a) * b, w11 = a * b;
  return make_float4(w00 * t00.x + w10 * t10.x + w01 * t01.x + w11 * t11.x,
                     w00 * t00.y + w10 * t10.y + w01 * t01.y + w11 * t11.y,
                     w00 * t00.z + w10 * t10.z + w01 * t01.z + w11 * t11.z,
                     w00 * t00.w + w10 * t10.w + w01 * t01.w + w11 * t11.w);
}

// =============================================================================
// C API (ctypes) -- mirrors the Renderer call sequence of the reference apps
// =============================================================================
namespace
{
Matrix3x4 mat3x4_from_glm(const glm::mat4& m)
{
  // renderer.h:409-412
  return make_mat3x4(make_float4(m[0][0], m[1][0], m[2][0], m[3][0]),
                     make_float4(m[0][1], m[1][1], m[2][1], m[3][1]),
                     make_float4(m[0][2], m[1][2], m[2][2], m[3][2]));
}

void rebuild_derived()
{
  // light list: every face whose material has emission (renderer.h:388-402)
  S->lights.clear();
  for (size_t face = 0; face < S->material_ids.size(); ++face) {
    const uint material_id = S->material_ids[face];
    const Material& m = S->materials[material_id];
    if (m.emission_color.x > 0 || m.emission_color.y > 0 ||
        m.emission_color.z > 0 || m.emission_texture_id != -1) {
      AreaLight light;
      light.indices = S->indices[face];
      light.material_id = material_id;
      light.instance_idx = S->instance_ids[face];
      S->lights.push_back(light);
    }
  }
  // transforms (renderer.h:404-421)
  S->o2w.resize(S->transforms.size());
  S->w2o.resize(S->transforms.size());
  for (size_t i = 0; i < S->transforms.size(); ++i) {
    S->o2w[i] = mat3x4_from_glm(S->transforms[i]);
    S->w2o[i] = mat3x4_from_glm(glm::inverse(S->transforms[i]));
  }
  // SBT hit-group data (renderer.h:306-327)
  S->sbt.resize(S->submesh_offsets.size());
  S->face_submesh.assign(S->indices.size(), 0);
  for (size_t sm = 0; sm < S->submesh_offsets.size(); ++sm) {
    S->sbt[sm].indices = S->indices.data() + S->submesh_offsets[sm];
    S->sbt[sm].material_ids = S->material_ids.data() + S->submesh_offsets[sm];
    for (uint f = 0; f < S->submesh_n_faces[sm]; ++f)
      S->face_submesh[S->submesh_offsets[sm] + f] = sm;
  }
  S->texture_headers.resize(S->textures.size());
  for (size_t i = 0; i < S->textures.size(); ++i) {
    S->texture_headers[i].size = make_uint2(S->textures[i]->w, S->textures[i]->h);
    S->texture_headers[i].texture_object =
        reinterpret_cast<cudaTextureObject_t>(S->textures[i].get());
  }
  S->accel_valid = false;
}

void build_accel()
{
  // instance i = submesh i with transform i (renderer.h:509-527); triangles are
  // moved to world space with the reference's own transform_position (shared.h:28)
  const size_t nf = S->indices.size();
  S->wv.resize(3 * nf);
  for (size_t f = 0; f < nf; ++f) {
    const Matrix3x4& m = S->o2w[S->face_submesh[f]];
    const uint3 idx = S->indices[f];
    S->wv[3 * f + 0] = transform_position(m, S->vertices[idx.x]);
    S->wv[3 * f + 1] = transform_position(m, S->vertices[idx.y]);
    S->wv[3 * f + 2] = transform_position(m, S->vertices[idx.z]);
  }
  S->bvh.build(S->wv);
  S->accel_valid = true;
}

void fill_params(const float* cam_transform12, float fov, float F, float focus,
                 const float* bg_color, const RenderLayer& layers, uint n_samples,
                 uint max_depth)
{
  // Renderer::render, renderer.h:661-724
  params.render_layer = layers;
  params.sample_count = S->sample_count.data();
  params.seed = 1;
  params.width = S->width;
  params.height = S->height;
  params.n_samples = n_samples;
  params.max_depth = max_depth;
  std::memcpy(&params.camera.transform, cam_transform12, sizeof(float) * 12);
  params.camera.fov = fov;
  params.camera.F = F;
  params.camera.focus = focus;
  params.object_to_world = S->o2w.data();
  params.world_to_object = S->w2o.data();
  params.vertices = S->vertices.data();
  params.normals = S->normals.data();
  params.texcoords = S->texcoords.data();
  params.materials = S->materials.data();
  params.textures = S->texture_headers.data();
  params.lights = S->lights.data();
  params.n_lights = S->lights.size();
  params.directional_light = S->has_dir_light ? &S->dir_light : nullptr;
  params.bg_color = make_float3(bg_color[0], bg_color[1], bg_color[2]);
  params.sky_intensity = S->sky_intensity;
  params.ibl = S->ibl ? reinterpret_cast<cudaTextureObject_t>(S->ibl.get()) : 0;
  params.sun_direction = S->sun_direction;
  params.arhosek = S->has_arhosek ? &S->arhosek : nullptr;
  params.ias_handle = 1;
}
}  // namespace

extern "C" {

void orc_reset()
{
  delete S;
  S = new OracleState();
  for (auto& c : S->n_rays) c = 0;
}

// Flat scene arrays == the Scene members the renderer consumes (scene.h:107-130).
// `materials` is an array of the reference's 180-byte Material (shared.h:100-142).
// `transforms` holds one column-major glm::mat4 (16 floats) per submesh.
void orc_set_scene(const float* vertices, const float* normals,
                   const float* texcoords, uint n_vertices, const uint* indices,
                   const uint* material_ids, const uint* instance_ids,
                   uint n_faces, const void* materials, uint n_materials,
                   const uint* submesh_offsets, const uint* submesh_n_faces,
                   const float* transforms, uint n_submeshes)
{
  if (!S) orc_reset();
  S->vertices.resize(n_vertices);
  S->normals.resize(n_vertices);
  S->texcoords.resize(n_vertices);
  std::memcpy(S->vertices.data(), vertices, sizeof(float3) * n_vertices);
  std::memcpy(S->normals.data(), normals, sizeof(float3) * n_vertices);
  std::memcpy(S->texcoords.data(), texcoords, sizeof(float2) * n_vertices);
  S->indices.resize(n_faces);
  std::memcpy(S->indices.data(), indices, sizeof(uint3) * n_faces);
  S->material_ids.assign(material_ids, material_ids + n_faces);
  S->instance_ids.assign(instance_ids, instance_ids + n_faces);
  static_assert(sizeof(Material) == 180, "Material layout (shared.h:100)");
  S->materials.resize(n_materials);
  std::memcpy(S->materials.data(), materials, sizeof(Material) * n_materials);
  S->submesh_offsets.assign(submesh_offsets, submesh_offsets + n_submeshes);
  S->submesh_n_faces.assign(submesh_n_faces, submesh_n_faces + n_submeshes);
  S->transforms.resize(n_submeshes);
  std::memcpy(S->transforms.data(), transforms, sizeof(float) * 16 * n_submeshes);
  rebuild_derived();
}

// animation (Renderer::set_time re-uploads transforms, renderer.h:614-640)
void orc_set_transforms(const float* transforms, uint n_submeshes)
{
  S->transforms.resize(n_submeshes);
  std::memcpy(S->transforms.data(), transforms, sizeof(float) * 16 * n_submeshes);
  rebuild_derived();
}

// returns texture id; rgba8 rows are in the order Texture::m_data holds them
int orc_add_texture(const unsigned char* rgba8, int w, int h, int is_color)
{
  auto t = std::make_unique<HostTexture>();
  t->w = w;
  t->h = h;
  t->srgb = is_color != 0;
  t->data8.resize((size_t)w * h);
  std::memcpy(t->data8.data(), rgba8, (size_t)w * h * 4);
  S->textures.push_back(std::move(t));
  rebuild_derived();
  return (int)S->textures.size() - 1;
}

void orc_set_ibl(const float* rgba32f, int w, int h)
{
  if (!rgba32f) {
    S->ibl.reset();
    return;
  }
  S->ibl = std::make_unique<HostTexture>();
  S->ibl->w = w;
  S->ibl->h = h;
  S->ibl->is_float = true;
  S->ibl->dataf.resize((size_t)w * h);
  std::memcpy(S->ibl->dataf.data(), rgba32f, (size_t)w * h * 16);
}

// Renderer::set_directional_light, renderer.h:554-567
void orc_set_directional_light(const float* le, const float* dir, float angle)
{
  S->dir_light.le = make_float3(le[0], le[1], le[2]);
  S->dir_light.dir = normalize(make_float3(dir[0], dir[1], dir[2]));
  S->dir_light.angle = angle;
  S->sun_direction = normalize(make_float3(dir[0], dir[1], dir[2]));
  S->has_dir_light = true;
}
void orc_clear_directional_light() { S->has_dir_light = false; }
void orc_set_sky_intensity(float v) { S->sky_intensity = v; }

// Renderer::load_arhosek_sky, renderer.h:588-607 (reference's own cook functions)
void orc_load_arhosek_sky(float turbidity, float albedo)
{
  const auto c2s = [](const float3& w) {
    float2 ret;
    ret.x = acosf(clamp(w.y, -1.0f, 1.0f));
    ret.y = atan2f(w.z, w.x);
    if (ret.y < 0) ret.y += 2.0f * M_PIf;
    return ret;
  };
  float elevation = c2s(S->sun_direction).x;
  elevation = 0.5f * M_PI - elevation;
  S->arhosek = arhosek_rgb_skymodelstate_alloc_init(turbidity, albedo, elevation);
  S->has_arhosek = true;
}
void orc_clear_arhosek_sky() { S->has_arhosek = false; }
// raw cook for unit tests: out = 3*9 configs then 3 radiances
void orc_arhosek_cook(float turbidity, float albedo, float elevation, float* out)
{
  const ArHosekSkyModelState st =
      arhosek_rgb_skymodelstate_alloc_init(turbidity, albedo, elevation);
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < 9; ++i) out[9 * c + i] = st.configs[c][i];
  for (int c = 0; c < 3; ++c) out[27 + c] = st.radiances[c];
}

// Renderer::set_resolution / init_render_states, renderer.h:642-655
void orc_set_resolution(uint w, uint h)
{
  S->width = w;
  S->height = h;
  S->sample_count.assign((size_t)w * h, 0);
}
void orc_init_render_states() { S->sample_count.assign((size_t)S->width * S->height, 0); }
void orc_set_sample_count(uint v) { S->sample_count.assign((size_t)S->width * S->height, v); }

void orc_build_accel() { build_accel(); }
uint orc_n_lights() { return S->lights.size(); }

// Renderer::render (renderer.h:657-734) + the optixLaunch loop, restricted to the
// pixel window [x0,x1) x [y0,y1).  Layers are HOST arrays laid out exactly like the
// device AOV buffers (float4 per pixel; depth float per pixel).  Returns seconds
// spent in the launch loop.
double orc_render(const float* cam_transform12, float fov, float F, float focus,
                  const float* bg_color, float* beauty, float* position,
                  float* depth, float* normal, float* texcoord, float* albedo,
                  uint n_samples, uint max_depth, uint x0, uint y0, uint x1,
                  uint y1, int n_threads)
{
  if (!S->accel_valid) build_accel();
  RenderLayer layers;
  layers.beauty = reinterpret_cast<float4*>(beauty);
  layers.position = reinterpret_cast<float4*>(position);
  layers.depth = depth;
  layers.normal = reinterpret_cast<float4*>(normal);
  layers.texcoord = reinterpret_cast<float4*>(texcoord);
  layers.albedo = reinterpret_cast<float4*>(albedo);
  fill_params(cam_transform12, fov, F, focus, bg_color, layers, n_samples,
              max_depth);
  n_threads = std::max(n_threads, 1);
  std::atomic<uint> next_row{y0};
  const auto worker = [&]() {
    for (int k = 0; k < 3; ++k) g_rays[k] = 0;
    for (;;) {
      const uint y = next_row.fetch_add(1);
      if (y >= y1) break;
      for (uint x = x0; x < x1; ++x) {
        g_ctx = TraceCtx{};
        g_ctx.launch_index = make_uint3(x, y, 0);
        __raygen__rg();
      }
    }
    for (int k = 0; k < 3; ++k) S->n_rays[k] += g_rays[k];
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (n_threads == 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Same launch loop over a LIST of pixel windows (tiles4 = n_tiles x (x0, y0, x1, y1)): bench.py's reference arm
// samples the frame in tiles spread over the whole image instead of one centre crop.  The work items handed
// to the threads are tile rows.
double orc_render_tiles(const float* cam_transform12, float fov, float F, float focus,
                        const float* bg_color, float* beauty, float* position,
                        float* depth, float* normal, float* texcoord, float* albedo,
                        uint n_samples, uint max_depth, const uint* tiles4, uint n_tiles, int n_threads)
{
  if (!S->accel_valid) build_accel();
  RenderLayer layers;
  layers.beauty = reinterpret_cast<float4*>(beauty);
  layers.position = reinterpret_cast<float4*>(position);
  layers.depth = depth;
  layers.normal = reinterpret_cast<float4*>(normal);
  layers.texcoord = reinterpret_cast<float4*>(texcoord);
  layers.albedo = reinterpret_cast<float4*>(albedo);
  fill_params(cam_transform12, fov, F, focus, bg_color, layers, n_samples, max_depth);
  struct Row {
    uint y, x0, x1;
  };
  std::vector<Row> rows;
  for (uint t = 0; t < n_tiles; ++t)
    for (uint y = tiles4[4 * t + 1]; y < tiles4[4 * t + 3]; ++y) rows.push_back(Row{y, tiles4[4 * t], tiles4[4 * t + 2]});
  n_threads = std::max(n_threads, 1);
  std::atomic<size_t> next_row{0};
  const auto worker = [&]() {
    for (int k = 0; k < 3; ++k) g_rays[k] = 0;
    for (;;) {
      const size_t i = next_row.fetch_add(1);
      if (i >= rows.size()) break;
      for (uint x = rows[i].x0; x < rows[i].x1; ++x) {
        g_ctx = TraceCtx{};
        g_ctx.launch_index = make_uint3(x, rows[i].y, 0);
        __raygen__rg();
      }
    }
    for (int k = 0; k < 3; ++k) S->n_rays[k] += g_rays[k];
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (n_threads == 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

void orc_get_ray_counts(unsigned long long* out3)
{
  for (int k = 0; k < 3; ++k) out3[k] = S->n_rays[k];
}
void orc_reset_ray_counts()
{
  for (auto& c : S->n_rays) c = 0;
}

// Batch closest-hit queries with the oracle's traversal (no any-hit programs):
// out_id[2*i] = instance (submesh), out_id[2*i+1] = primitive, 0xffffffff on miss;
// out_tuv[3*i..] = t, u, v.
void orc_trace_closest(const float* origins, const float* dirs, uint n, float tmin,
                       float tmax, uint* out_id, float* out_tuv)
{
  if (!S->accel_valid) build_accel();
  for (uint i = 0; i < n; ++i) {
    const float3 o = make_float3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
    const float3 d = make_float3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    const Hit h = traverse(o, d, tmin, tmax, false, 0, false);
    if (h.valid) {
      const uint sm = S->face_submesh[h.face];
      out_id[2 * i] = sm;
      out_id[2 * i + 1] = h.face - S->submesh_offsets[sm];
      out_tuv[3 * i] = h.t;
      out_tuv[3 * i + 1] = h.u;
      out_tuv[3 * i + 2] = h.v;
    } else {
      out_id[2 * i] = out_id[2 * i + 1] = 0xffffffffu;
      out_tuv[3 * i] = out_tuv[3 * i + 1] = out_tuv[3 * i + 2] = 0.0f;
    }
  }
}

// brute-force variant (no BVH) to validate the oracle's own BVH
void orc_trace_closest_bruteforce(const float* origins, const float* dirs, uint n,
                                  float tmin, float tmax, uint* out_id,
                                  float* out_tuv)
{
  if (!S->accel_valid) build_accel();
  const uint nf = S->indices.size();
  for (uint i = 0; i < n; ++i) {
    const float3 o = make_float3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
    const float3 d = make_float3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    const RaySetup rs = make_ray_setup(o, d);
    float bt = tmax, bu = 0, bv = 0;
    uint bf = 0xffffffffu;
    for (uint f = 0; f < nf; ++f) {
      float t, u, v;
      const float tlim = bf != 0xffffffffu ? std::nextafter(bt, 3.0e38f) : tmax;
      if (!intersect_tri(rs, S->wv[3 * (size_t)f], S->wv[3 * (size_t)f + 1],
                         S->wv[3 * (size_t)f + 2], tmin, tlim, t, u, v))
        continue;
      if (bf != 0xffffffffu && t == bt && f > bf) continue;
      bt = t;
      bu = u;
      bv = v;
      bf = f;
    }
    if (bf != 0xffffffffu) {
      const uint sm = S->face_submesh[bf];
      out_id[2 * i] = sm;
      out_id[2 * i + 1] = bf - S->submesh_offsets[sm];
      out_tuv[3 * i] = bt;
      out_tuv[3 * i + 1] = bu;
      out_tuv[3 * i + 2] = bv;
    } else {
      out_id[2 * i] = out_id[2 * i + 1] = 0xffffffffu;
      out_tuv[3 * i] = out_tuv[3 * i + 1] = out_tuv[3 * i + 2] = 0.0f;
    }
  }
}

// Primary rays exactly as __raygen__rg generates them for sample index `n_spp`
// (pt.cu:435-446): out_ray[6*i..] = origin, direction for pixel i (row-major).
void orc_primary_rays(const float* cam_transform12, float fov, float F,
                      float focus, uint n_spp, float* out_ray)
{
  RenderLayer layers{};
  const float bg[3] = {0, 0, 0};
  fill_params(cam_transform12, fov, F, focus, bg, layers, 1, 1);
  for (uint y = 0; y < S->height; ++y) {
    for (uint x = 0; x < S->width; ++x) {
      const uint3 idx = make_uint3(x, y, 0);
      const uint3 dim = make_uint3(S->width, S->height, 1);
      const uint image_idx = x + S->width * y;
      SamplerState st;
      init_sampler_state(idx, image_idx, n_spp, st);
      float2 u = sample_2d(st);
      float2 uv = make_float2((2.0f * (idx.x + u.x) - dim.x) / dim.y,
                              (2.0f * (idx.y + u.y) - dim.y) / dim.y);
      uv.x = -uv.x;
      u = sample_2d(st);
      float pdf;
      float3 o, d;
      sample_ray_thinlens_camera(params.camera, uv, u, o, d, pdf);
      float* r = out_ray + 6 * (size_t)image_idx;
      r[0] = o.x;
      r[1] = o.y;
      r[2] = o.z;
      r[3] = d.x;
      r[4] = d.y;
      r[5] = d.z;
    }
  }
}

// Sampler known-answer vectors: draws the sequence described by `kinds`
// ('1' = sample_1d, '2' = sample_2d) from a freshly initialised SamplerState
// (pt.cu:378-399) and writes 1 or 2 floats per draw.
void orc_sampler_sequence(uint width, uint height, uint seed, uint image_idx,
                          uint n_spp, const char* kinds, float* out)
{
  params.width = width;
  params.height = height;
  params.seed = seed;
  SamplerState st;
  init_sampler_state(make_uint3(image_idx % width, image_idx / width, 0),
                     image_idx, n_spp, st);
  for (const char* k = kinds; *k; ++k) {
    if (*k == '1') {
      *out++ = sample_1d(st);
    } else {
      const float2 v = sample_2d(st);
      *out++ = v.x;
      *out++ = v.y;
    }
  }
}

// integer-level sampler primitives
uint orc_xxhash32_1(uint p) { return xxhash32(p); }
uint orc_xxhash32_4(uint x, uint y, uint z, uint w)
{
  return xxhash32(make_uint4(x, y, z, w));
}
uint orc_cmj_permute(uint i, uint l, uint p) { return cmj_permute(i, l, p); }
uint orc_sobol(unsigned long long index, uint dimension, uint scramble)
{
  return sobol(index, dimension, scramble);
}
uint orc_owen(uint x, uint seed) { return nested_uniform_scramble_base2(x, seed); }

// BSDF known-answer vectors.  `sp` is the reference's 120-byte ShadingParams.
void orc_bsdf_eval(const void* sp, const float* wo, int is_entering,
                   const float* wi, float* out_f3_pdf)
{
  static_assert(sizeof(ShadingParams) == 120, "ShadingParams layout");
  ShadingParams p;
  std::memcpy(&p, sp, sizeof(p));
  const float3 o = make_float3(wo[0], wo[1], wo[2]);
  const float3 i = make_float3(wi[0], wi[1], wi[2]);
  const BSDF bsdf(o, p, is_entering != 0);
  const float3 f = bsdf.eval(o, i);
  out_f3_pdf[0] = f.x;
  out_f3_pdf[1] = f.y;
  out_f3_pdf[2] = f.z;
  out_f3_pdf[3] = bsdf.eval_pdf(o, i);
}
void orc_bsdf_sample(const void* sp, const float* wo, int is_entering, float u,
                     const float* v2, float* out_wi3_f3_pdf)
{
  ShadingParams p;
  std::memcpy(&p, sp, sizeof(p));
  const float3 o = make_float3(wo[0], wo[1], wo[2]);
  const BSDF bsdf(o, p, is_entering != 0);
  float3 f;
  float pdf;
  const float3 wi = bsdf.sample(o, u, make_float2(v2[0], v2[1]), f, pdf);
  out_wi3_f3_pdf[0] = wi.x;
  out_wi3_f3_pdf[1] = wi.y;
  out_wi3_f3_pdf[2] = wi.z;
  out_wi3_f3_pdf[3] = f.x;
  out_wi3_f3_pdf[4] = f.y;
  out_wi3_f3_pdf[5] = f.z;
  out_wi3_f3_pdf[6] = pdf;
}

// Hosek sky radiance for a world direction (pt.cu:352-363)
void orc_sky_radiance(const float* dir, float* out3)
{
  params.sky_intensity = S->sky_intensity;
  params.sun_direction = S->sun_direction;
  params.arhosek = &S->arhosek;
  const float3 r = evaluate_arhosek_sky(make_float3(dir[0], dir[1], dir[2]));
  out3[0] = r.x;
  out3[1] = r.y;
  out3[2] = r.z;
}

uint orc_sizeof_material() { return sizeof(Material); }
uint orc_sizeof_shading_params() { return sizeof(ShadingParams); }
uint orc_sizeof_launch_params() { return sizeof(LaunchParams); }

}  // extern "C"
