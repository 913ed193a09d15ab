// Compressed 8-wide BVH (CWBVH, Ylitie/Karras/Laine 2017) and its traversal.
//
// Replaces the reference's OptiX acceleration structures + optixTrace
// (renderer.h:434-552, pt.cu:82-123): B200 has no RT cores, so closest-hit /
// any-hit queries are answered by this software traversal.
//
// Geometry is flattened to WORLD space at build time (instance i = submesh i with
// transform i, renderer.h:509-527), so one single-level tree serves the whole
// scene and a hit reports the global face index; (instance, primitive) are
// recovered from the submesh table.
//
// The ray/triangle test is the watertight test of Woop, Benthin, Wald (2013).  Its
// arithmetic is pinned instruction by instruction (explicit _rn intrinsics, no
// compiler contraction) because primary-hit parity with the host oracle is
// required to be bit-exact: see oracle/oracle_host.cpp (intersect_tri).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "vecmath.cuh"

namespace frd
{

// 80-byte node in a 96-byte, 32-byte-aligned record: the traversal reads it as two 256-bit words (LDG.E.256, new
// on sm_100) and one 128-bit word -- three load instructions per node visit instead of five.  The traversal
// kernels keep the L1 at 40-56 % of its peak throughput with every lane on a different node
// (profiles/r2_kernels_ncu_full.txt), and for such loads the tag stage works per instruction and line, so fewer,
// wider loads pay: +1.1 % on the bench frame (profiles/r2l_node_loads.txt) for 20 % more node bytes.
struct alignas(32) Node8 {
  float px, py, pz;          // quantisation origin (node box lower corner)
  uint8_t ex, ey, ez;        // biased exponents of the per-axis grid step
  uint8_t imask;             // bit s set: slot s holds an internal child
  uint32_t child_base;       // index of first internal child node
  uint32_t tri_base;         // index of first leaf triangle
  uint8_t meta[8];           // per slot: inner -> 0b001xxxxx (24+slot); leaf -> unary count<<5 | offset
  uint8_t qlox[8], qloy[8];  // quantised child boxes
  uint8_t qloz[8], qhix[8];
  uint8_t qhiy[8], qhiz[8];
  uint32_t pad_[4];          // alignment only
};
static_assert(sizeof(Node8) == 96 && offsetof(Node8, pad_) == 80, "CWBVH node: 80 bytes of payload in a 96-byte record");
constexpr uint32_t kNodeWords = sizeof(Node8) / 16;  // float4 words per node

// 256-bit read-only load (sm_100: LDG.E.256); p must be 32-byte aligned
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b)
{
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

// 48-byte leaf triangle: world-space vertices; v0.w = global face index (bits),
// v1.w = flags (bit 0: alpha-tested material, needs the any-hit path).
struct alignas(16) LeafTri {
  float4 v0, v1, v2;
};

// Two-level mode (instanced scenes, set_time without a rebuild): the node / triangle arrays hold an instance tree
// (TLAS) at index 0 followed by one object-space tree (BLAS) per DISTINCT mesh.  A TLAS "triangle" is a
// placeholder whose three vertices span the instance's world box; its v0.w is the instance index.  All child /
// triangle indices are absolute.  instances == nullptr: one flat world-space tree (the default).
struct InstanceRecord {
  uint32_t blas_root;    // node index of the mesh's root
  uint32_t face_offset;  // global face index of the instance's first face (BLAS triangles carry mesh-local ids)
};

struct BvhView {
  const float4* nodes;  // Node8 as float4[kNodeWords]
  const float4* tris;   // LeafTri as float4[3]
  const InstanceRecord* instances;  // two-level mode only
  const float4* w2o;                // world-to-object rows, 3 float4 per instance (two-level mode only)
};

struct HitRecord {
  float t;        // hit distance (ray tmax if miss)
  float u, v;     // barycentrics: p = (1-u-v) v0 + u v1 + v v2
  uint32_t face;  // global face index, 0xffffffff = miss
};

constexpr uint32_t kNoHit = 0xffffffffu;

// ---- watertight ray setup / test (must mirror oracle_host.cpp exactly) ---------
struct RayShear {
  float cx[3], cy[3], cz[3];
};

FR_D RayShear make_shear(const float3& d)
{
  const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  int kz = 0;
  float am = ax;
  if (ay > am) {
    kz = 1;
    am = ay;
  }
  if (az > am) kz = 2;
  int kx = kz == 2 ? 0 : kz + 1;
  int ky = kx == 2 ? 0 : kx + 1;
  const float dz = kz == 0 ? d.x : (kz == 1 ? d.y : d.z);
  if (dz < 0.0f) {
    const int t = kx;
    kx = ky;
    ky = t;
  }
  const float dx = kx == 0 ? d.x : (kx == 1 ? d.y : d.z);
  const float dy = ky == 0 ? d.x : (ky == 1 ? d.y : d.z);
  const float Sx = __fdiv_rn(dx, dz);
  const float Sy = __fdiv_rn(dy, dz);
  const float Sz = __fdiv_rn(1.0f, dz);
  RayShear r;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    r.cx[i] = (i == kx) ? 1.0f : (i == kz) ? -Sx : 0.0f;
    r.cy[i] = (i == ky) ? 1.0f : (i == kz) ? -Sy : 0.0f;
    r.cz[i] = (i == kz) ? Sz : 0.0f;
  }
  return r;
}

FR_D float shear_dot(float ax, float ay, float az, const float c[3])
{
  return __fmaf_rn(az, c[2], __fmaf_rn(ay, c[1], __fmul_rn(ax, c[0])));
}

// true if tmin < t < tlim (or t == tlim when allow_equal), outputs t,u,v
// OCCLUSION = true (any-hit rays): only WHETHER the triangle is hit inside (tmin, tlim) matters, so the division
// is replaced by comparing the scaled distance T against tmin |det| and tlim |det| (same edge functions, same
// sign rules; the outcome can differ from the divided form only for a distance within an ulp of an end of the
// interval).  t, bu, bv are then only computed when want_uv is set (alpha-tested triangles).
template <bool OCCLUSION>
FR_D bool watertight_hit(const RayShear& s, const float3& o, const float4& v0, const float4& v1,
                         const float4& v2, float tmin, float tlim, bool allow_equal, float& t,
                         float& bu, float& bv, bool want_uv = true)
{
  const float Ax0 = __fsub_rn(v0.x, o.x), Ay0 = __fsub_rn(v0.y, o.y), Az0 = __fsub_rn(v0.z, o.z);
  const float Bx0 = __fsub_rn(v1.x, o.x), By0 = __fsub_rn(v1.y, o.y), Bz0 = __fsub_rn(v1.z, o.z);
  const float Cx0 = __fsub_rn(v2.x, o.x), Cy0 = __fsub_rn(v2.y, o.y), Cz0 = __fsub_rn(v2.z, o.z);
  const float Ax = shear_dot(Ax0, Ay0, Az0, s.cx), Ay = shear_dot(Ax0, Ay0, Az0, s.cy);
  const float Bx = shear_dot(Bx0, By0, Bz0, s.cx), By = shear_dot(Bx0, By0, Bz0, s.cy);
  const float Cx = shear_dot(Cx0, Cy0, Cz0, s.cx), Cy = shear_dot(Cx0, Cy0, Cz0, s.cy);
  float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
  float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
  float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = __double2float_rn(__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx)));
    V = __double2float_rn(__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx)));
    W = __double2float_rn(__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax)));
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  const float det = __fadd_rn(__fadd_rn(U, V), W);
  if (det == 0.0f) return false;
  const float Az = shear_dot(Ax0, Ay0, Az0, s.cz);
  const float Bz = shear_dot(Bx0, By0, Bz0, s.cz);
  const float Cz = shear_dot(Cx0, Cy0, Cz0, s.cz);
  const float T = __fmaf_rn(W, Cz, __fmaf_rn(V, Bz, __fmul_rn(U, Az)));
  if (OCCLUSION) {
    const float ad = fabsf(det);
    const float Ts = __uint_as_float(__float_as_uint(T) ^ (__float_as_uint(det) & 0x80000000u));
    if (!(Ts > tmin * ad) || !(Ts < tlim * ad)) return false;
    if (!want_uv) return true;
  }
  const float rcp = __frcp_rn(det);  // the correctly rounded 1 / det, as the host oracle's 1.0f / det (shorter than __fdiv_rn)
  const float tt = __fmul_rn(T, rcp);
  if (!(tt > tmin)) return false;
  if (!(tt < tlim || (allow_equal && tt == tlim))) return false;
  t = tt;
  bu = __fmul_rn(V, rcp);
  bv = __fmul_rn(W, rcp);
  return true;
}

// ---- traversal -------------------------------------------------------------------
FR_D uint32_t sign_extend_s8x4(uint32_t x)
{
  uint32_t r;
  asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
  return r;
}
FR_D float byte_f(uint32_t w, int j) { return (float)((w >> (8 * j)) & 0xffu); }

// Short stack kept in shared memory (one column per thread, 8-byte entries, so a
// warp's accesses are conflict free); deep paths overflow to thread-local memory.
#ifndef FRD_SMEM_STACK
#define FRD_SMEM_STACK 12
#endif
constexpr int kSmemStack = FRD_SMEM_STACK;
constexpr int kLocalStack = 36;

struct TravStack {
  uint2* smem;  // &smem_stack[0][thread]
  int stride;   // threads per block
  uint2* local;  // [kLocalStack] thread-local overflow, declared by the caller so that the rest of
                 // the traversal state stays in registers
  int sp;
  FR_D void push(const uint2& e)
  {
    if (sp < kSmemStack)
      smem[sp * stride] = e;
    else
      local[sp - kSmemStack] = e;
    sp++;
  }
  FR_D uint2 pop()
  {
    sp--;
    return sp < kSmemStack ? smem[sp * stride] : local[sp - kSmemStack];
  }
};

struct TraceCounters {
  uint32_t nodes, tris;
};

// Alpha-test hook: returns true if the candidate hit is accepted.  Opaque scenes
// compile this away.
struct NoAnyHit {
  FR_D bool operator()(uint32_t, float, float) const { return true; }
};

// Per-lane traversal state machine, driven warp-synchronously by trace_queue().
//
// A lane owns one ray.  Its work is split into two phases that the WARP schedules:
//   node phase      pop one child of the current node group, fetch the 80-byte node, test
//                   its 8 quantised boxes (~280 SASS instructions, ~90 % of lanes busy)
//   triangle phase  test ONE pending leaf triangle (watertight, ~100 instructions)
// Testing a node's leaf triangles right away (the textbook while-while loop) ran the
// triangle code with ~3 of 32 lanes active and cost as many issue slots as all node tests
// together (profiles/r1c_trace_closest_source.txt).  Instead a lane parks the triangle
// group it found (two slots) and keeps descending; the warp runs a triangle phase only when
// enough lanes have a triangle pending (or nothing else can make progress), so the
// triangle code runs with most lanes active.  The order of triangle tests does not change
// the result: closest hit is a minimum with a fixed tie rule, any-hit only reports
// occlusion.
//
// Closest hit (ANY = false) or first accepted hit (ANY = true).
// Tie rule on exactly equal t: lower global face index wins (as in the oracle).
template <bool TWO>
struct Traverser {
  float3 o;
  RayShear sh;
  float3 idir;  // 1 / d (components of |d| < tiny replaced, box tests only)
  uint32_t octinv;
  float tmin;
  HitRecord best;
  uint2 ngroup;   // current node group: x = child base, y = hit bits (31..24) | imask; y == 0: none
  uint2 tgroup;   // pending triangles: x = triangle base, y = 24-bit mask
  uint2 tgroup2;  // second parking slot
  TravStack st;
  // two-level mode: the instance being traversed (kNoHit: in the TLAS) and the stack height its tree started at
  uint32_t inst;
  int sp_base;

  FR_D void begin(const float3& org, const float3& d, float t_min, float t_max)
  {
    tmin = t_min;
    best.t = t_max;
    best.u = best.v = 0.0f;
    best.face = kNoHit;
    set_ray(org, d);
    st.sp = 0;
    ngroup = make_uint2(0u, 0x80000000u);
    tgroup = make_uint2(0u, 0u);
    tgroup2 = make_uint2(0u, 0u);
    inst = kNoHit;
    sp_base = 0;
  }

  // direction-dependent constants of the ray (shear, reciprocal direction, octant)
  FR_D void set_ray(const float3& org, const float3& d)
  {
    o = org;
    sh = make_shear(d);
    // box tests are conservative: guard against 0 * inf and widen by a few ulps
    const float tiny = 1e-20f;
    const float3 ds = f3(fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x),
                         fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y),
                         fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
    idir = f3(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
    octinv = (d.x >= 0.0f ? 1u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 4u : 0u);
  }

  // ---- two-level mode -------------------------------------------------------------------------
  // The instance's tree is walked with the ray taken to object space, o' = W2O o, d' = W2O d (not
  // normalised, so t is the same parameter in both spaces and `best` carries over).  What is left of the
  // TLAS walk -- node group, parked placeholders -- goes on the stack; sp_base marks where the instance's own
  // stack starts.  The world ray is not kept: the policy reloads it from the queue record on the way back.
  FR_D bool in_instance() const { return TWO && inst != kNoHit; }
  FR_D bool instance_done() const { return TWO && inst != kNoHit && ngroup.y == 0u && tgroup.y == 0u; }
  FR_D int stack_floor() const { return TWO ? sp_base : 0; }

  FR_D void enter_instance(const BvhView& bvh, uint32_t instance, const float3& wo, const float3& wd)
  {
    st.push(ngroup);
    st.push(tgroup);
    st.push(tgroup2);
    const float4 r0 = __ldg(bvh.w2o + 3ull * instance), r1 = __ldg(bvh.w2o + 3ull * instance + 1),
                 r2 = __ldg(bvh.w2o + 3ull * instance + 2);
    const float3 oo = f3(r0.x * wo.x + r0.y * wo.y + r0.z * wo.z + r0.w, r1.x * wo.x + r1.y * wo.y + r1.z * wo.z + r1.w,
                         r2.x * wo.x + r2.y * wo.y + r2.z * wo.z + r2.w);
    const float3 od = f3(r0.x * wd.x + r0.y * wd.y + r0.z * wd.z, r1.x * wd.x + r1.y * wd.y + r1.z * wd.z,
                         r2.x * wd.x + r2.y * wd.y + r2.z * wd.z);
    set_ray(oo, od);
    inst = instance;
    sp_base = st.sp;
    ngroup = make_uint2(bvh.instances[instance].blas_root, 0x80000000u);
    tgroup = make_uint2(0u, 0u);
    tgroup2 = make_uint2(0u, 0u);
  }

  // back to the TLAS walk with the world ray (wo, wd)
  FR_D void leave_instance(const float3& wo, const float3& wd)
  {
    tgroup2 = st.pop();
    tgroup = st.pop();
    ngroup = st.pop();
    inst = kNoHit;
    sp_base = 0;
    set_ray(wo, wd);
  }

  FR_D bool has_triangles() const { return tgroup.y != 0u; }
  // node work available and a free slot to park the triangles it may produce
  FR_D bool can_descend() const { return ngroup.y != 0u && tgroup2.y == 0u; }
  FR_D bool finished() const { return ngroup.y == 0u && tgroup.y == 0u && !in_instance(); }

  template <bool COUNT>
  FR_D void node_phase(const BvhView& bvh, TraceCounters* cnt)
  {
    constexpr float kFar = 1.0000005f, kNear = 0.9999995f;
    const uint32_t hits_imask = ngroup.y;
    const uint32_t bit = 31u - __clz(hits_imask);
    ngroup.y &= ~(1u << bit);  // what is left of this group (pushed below if the child has inner hits)
    const uint32_t slot = (bit ^ octinv) & 7u;
    const uint32_t rel = __popc(hits_imask & ~(0xffffffffu << slot));
    const float4* np = bvh.nodes + (unsigned long long)kNodeWords * (ngroup.x + rel);
    float4 n0, n1, n2, n3;
    ldg256(np, n0, n1);
    ldg256(np + 2, n2, n3);
    const float4 n4 = __ldg(np + 4);
    if (COUNT) cnt->nodes++;
    const uint32_t ew = __float_as_uint(n0.w);
    const float sx = __uint_as_float((ew & 0xffu) << 23);
    const float sy = __uint_as_float(((ew >> 8) & 0xffu) << 23);
    const float sz = __uint_as_float(((ew >> 16) & 0xffu) << 23);
    const float ax = sx * idir.x, ay = sy * idir.y, az = sz * idir.z;
    const float bx = (n0.x - o.x) * idir.x, by = (n0.y - o.y) * idir.y, bz = (n0.z - o.z) * idir.z;
    const bool negx = !(octinv & 1u), negy = !(octinv & 2u), negz = !(octinv & 4u);
    const uint32_t octinv4 = octinv * 0x01010101u;
    const float tnear0 = tmin, tfar0 = best.t;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const uint32_t meta4 = __float_as_uint(half == 0 ? n1.z : n1.w);
      // no child in slots 4-7: the builder packs nodes of <= 4 children into the lower half
      // (-1.7 % closest-hit time; the branch is per lane, it only pays when the warp agrees)
      if (half == 1 && meta4 == 0u) continue;
      const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
      const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
      const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
      const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
      const uint32_t qlox = __float_as_uint(half == 0 ? n2.x : n2.y);
      const uint32_t qloy = __float_as_uint(half == 0 ? n2.z : n2.w);
      const uint32_t qloz = __float_as_uint(half == 0 ? n3.x : n3.y);
      const uint32_t qhix = __float_as_uint(half == 0 ? n3.z : n3.w);
      const uint32_t qhiy = __float_as_uint(half == 0 ? n4.x : n4.y);
      const uint32_t qhiz = __float_as_uint(half == 0 ? n4.z : n4.w);
      const uint32_t nx = negx ? qhix : qlox, fx = negx ? qlox : qhix;
      const uint32_t ny = negy ? qhiy : qloy, fy = negy ? qloy : qhiy;
      const uint32_t nz = negz ? qhiz : qloz, fz = negz ? qloz : qhiz;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float t0x = fmaf(byte_f(nx, j), ax, bx), t1x = fmaf(byte_f(fx, j), ax, bx);
        const float t0y = fmaf(byte_f(ny, j), ay, by), t1y = fmaf(byte_f(fy, j), ay, by);
        const float t0z = fmaf(byte_f(nz, j), az, bz), t1z = fmaf(byte_f(fz, j), az, bz);
        const float cnear = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tnear0)) * kNear;
        const float cfar = fminf(fminf(t1x, t1y), fminf(t1z, tfar0)) * kFar;
        if (cnear <= cfar) {
          const uint32_t cb = (child_bits4 >> (8 * j)) & 0xffu;
          const uint32_t bi = (bit_index4 >> (8 * j)) & 0xffu;
          hitmask |= cb << bi;
        }
      }
    }
    // inner children hit: descend into this node's group, the rest of the parent's group
    // goes to the stack; none: carry on with the rest of the group, or resume from the stack
    if (hitmask & 0xff000000u) {
      if (ngroup.y > 0x00ffffffu) st.push(ngroup);
      ngroup.x = __float_as_uint(n1.x);
      ngroup.y = (hitmask & 0xff000000u) | (ew >> 24);
    } else if (ngroup.y <= 0x00ffffffu) {
      ngroup = st.sp > stack_floor() ? st.pop() : make_uint2(0u, 0u);
    }
    // park the leaf triangles this node produced
    const uint32_t tmask = hitmask & 0x00ffffffu;
    if (tmask) {
      const uint2 tg = make_uint2(__float_as_uint(n1.y), tmask);
      if (tgroup.y == 0u)
        tgroup = tg;
      else
        tgroup2 = tg;
    }
  }

  // tests ONE pending triangle; returns true if an ANY ray found its hit
  // Policy: world_ray(o, d) reloads the lane's ray (two-level mode only)
  template <bool ANY, bool COUNT, typename AnyHit, typename Policy>
  FR_D bool triangle_phase(const BvhView& bvh, const AnyHit& anyhit, TraceCounters* cnt, Policy& pol)
  {
    const uint32_t bit = __ffs(tgroup.y) - 1u;
    tgroup.y &= tgroup.y - 1u;
    const float4* tp = bvh.tris + 3ull * (tgroup.x + bit);
    if (TWO && inst == kNoHit) {
      // a TLAS leaf entry: the placeholder of an instance (its box was tested with the node); walk its tree
      const float4 lo = __ldg(tp), hi = __ldg(tp + 1);  // the placeholder's v0 / v1 are the corners of the world box
      const uint32_t instance = __float_as_uint(lo.w);
      if (tgroup.y == 0u) {
        tgroup = tgroup2;
        tgroup2 = make_uint2(0u, 0u);
      }
      // Entering costs a ray transform and a walk from the mesh's root, and the placeholder may have been parked
      // before a closer hit was found: test its exact box against the CURRENT best.t first (conservative slabs).
      {
        const float ax = (lo.x - o.x) * idir.x, bx = (hi.x - o.x) * idir.x;
        const float ay = (lo.y - o.y) * idir.y, by = (hi.y - o.y) * idir.y;
        const float az = (lo.z - o.z) * idir.z, bz = (hi.z - o.z) * idir.z;
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin)) * 0.9999995f;
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), best.t)) * 1.0000005f;
        if (!(tn <= tf)) return false;
      }
      float3 wo, wd;
      pol.world_ray(wo, wd);
      enter_instance(bvh, instance, wo, wd);
      return false;
    }
    const float4 v0 = __ldg(tp + 0), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    if (tgroup.y == 0u) {
      tgroup = tgroup2;
      tgroup2 = make_uint2(0u, 0u);
    }
    if (COUNT) cnt->tris++;
    float t = 0.0f, u = 0.0f, v = 0.0f;
    const bool alpha_tested = (__float_as_uint(v1.w) & 1u) != 0u;
    if (!watertight_hit<ANY>(sh, o, v0, v1, v2, tmin, best.t, best.face != kNoHit, t, u, v, alpha_tested)) return false;
    const uint32_t face = __float_as_uint(v0.w) + (TWO ? bvh.instances[inst].face_offset : 0u);
    if (!ANY && t == best.t && best.face != kNoHit && face > best.face) return false;
    if (alpha_tested && !anyhit(face, u, v)) return false;
    best.t = t;
    best.u = u;
    best.v = v;
    best.face = face;
    return ANY;
  }
};

// Persistent-warp queue driver (trace.cu kernels).  Every lane owns one ray; the warp
// iterates { vote, triangle phase, node phase } fully converged at the top of each
// iteration:
//  * lane refill (Aila & Laine 2009 "dynamic fetch"): when at least `refill_lanes` lanes are
//    idle the finished rays are retired (Policy::retire, whole warp converged so it can use
//    warp-aggregated queue appends) and the idle lanes fetch new items with ONE atomic on the
//    queue cursor; once the queue is exhausted the warp drains;
//  * triangle phase when at least `tri_lanes` lanes have a triangle pending, or when no lane
//    can descend.
//
// Policy interface:
//   void load(uint32_t item, float3& o, float3& d, float& tmin, float& tmax)  (lane-local payload)
//   void retire(bool has_result, const HitRecord& h, const TraceCounters&)    all 32 lanes together
//   anyhit()                                                                  alpha-test functor
//   void world_ray(float3& o, float3& d)                                       two-level mode: reload the ray
template <bool ANY, bool COUNT, bool TWO, typename Policy>
FR_D void trace_queue(const BvhView& bvh, Policy& pol, uint32_t* cursor, uint32_t n, uint2* smem_column,
                      int stride, int refill_lanes, int tri_lanes)
{
  uint2 overflow[kLocalStack];
  Traverser<TWO> tr;
  tr.st.smem = smem_column;
  tr.st.stride = stride;
  tr.st.local = overflow;
  tr.st.sp = 0;
  tr.ngroup = tr.tgroup = tr.tgroup2 = make_uint2(0u, 0u);
  tr.inst = kNoHit;
  tr.sp_base = 0;
  TraceCounters cnt{0u, 0u};
  const uint32_t lane = threadIdx.x & 31u;
  bool have = false;      // lane is traversing a ray
  bool finished = false;  // lane holds a finished ray that has not been retired yet
  bool exhausted = false;
  // Work items are taken from the global cursor 32 at a time and handed out from a warp-local reserve
  // [res_next, res_end): two refills out of three need no atomic at all, and the cursor -- one address
  // for the whole grid -- sees a third of the traffic.
  uint32_t res_next = 0u, res_end = 0u;  // warp-uniform
  bool global_done = false;              // the cursor has passed n
  for (;;) {
    const uint32_t idle = __ballot_sync(0xffffffffu, !have);
    if (idle == 0xffffffffu || (!exhausted && __popc(idle) >= refill_lanes)) {
      const uint32_t need = (uint32_t)__popc(idle);
      const uint32_t left = res_end - res_next;
      // the cursor atomic is issued first and consumed after the retire step, so that its round
      // trip overlaps the retire step's own memory traffic (queue appends, radiance updates)
      const bool fetch = !global_done && left < need;
      uint32_t chunk = 0;
      if (fetch && lane == 0) chunk = atomicAdd(cursor, 32u);
      pol.retire(finished, tr.best, cnt);
      finished = false;
      if (!exhausted) {
        const uint32_t k = __popc(idle & ((1u << lane) - 1u));  // this lane's rank among the idle lanes
        uint32_t item = res_next + k;                           // from what is left of the reserve ...
        if (fetch) {
          chunk = __shfl_sync(0xffffffffu, chunk, 0);
          if (k >= left) item = chunk + (k - left);             // ... then from the new chunk
          res_next = chunk + (need - left);
          res_end = chunk + 32u;
          global_done = res_end >= n;
        } else {
          res_next += need < left ? need : left;
          if (k >= left) item = 0xffffffffu;                    // reserve ran dry and the queue is finished
        }
        if (!have && item < n) {
          float3 o, d;
          float tmin, tmax;
          pol.load(item, o, d, tmin, tmax);
          tr.begin(o, d, tmin, tmax);
          if (COUNT) cnt.nodes = cnt.tris = 0u;
          have = true;
        }
        exhausted = global_done && (res_next >= res_end || res_next >= n);
      }
      if (__ballot_sync(0xffffffffu, have) == 0u) break;
    }
    // ---- vote ----
    const bool want_tri = have && tr.has_triangles();
    const uint32_t tri_votes = __ballot_sync(0xffffffffu, want_tri);
    const uint32_t node_votes = __ballot_sync(0xffffffffu, have && tr.can_descend());
    bool done = false;
    if (tri_votes != 0u && (__popc(tri_votes) >= tri_lanes || node_votes == 0u)) {
      if (want_tri) done = tr.template triangle_phase<ANY, COUNT>(bvh, pol.anyhit(), &cnt, pol);
      __syncwarp();
    }
    if (have && !done && tr.can_descend()) tr.template node_phase<COUNT>(bvh, &cnt);
    if (TWO && have && !done && tr.instance_done()) {
      // the instance's tree is exhausted: back to the instance tree with the world ray
      float3 wo, wd;
      pol.world_ray(wo, wd);
      tr.leave_instance(wo, wd);
    }
    if (have && (done || tr.finished())) {
      have = false;
      finished = true;
    }
    __syncwarp();
  }
}

}  // namespace frd
