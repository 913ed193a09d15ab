// GPU BVH construction: world-space flattening -> 63-bit Morton codes -> radix sort -> binary tree
// (default: PLOC, parallel locally-ordered clustering, Meister & Bittner 2018, search radius 4;
// FRD_BVH_BUILDER=lbvh: Karras 2012 radix tree + bottom-up box fit) -> greedy collapse to an 8-wide
// tree -> CWBVH quantisation (Ylitie et al. 2017).
//
// Stands in for the reference's optixAccelBuild calls (per-submesh GAS + one IAS,
// renderer.h:434-552).  Like the reference on set_time (renderer.h:614-619) the
// tree is rebuilt, not refitted, when transforms change -- a rebuild of 1 M
// triangles is 3.6 ms on a B200 (52 M triangles: 53 ms).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>

#include "bvh_build.h"
#include "bvh_quant.cuh"

namespace frd
{
namespace
{

#ifndef FR_LEAF_MAX
#define FR_LEAF_MAX 3
#endif
constexpr int kLeafMax = FR_LEAF_MAX;  // triangles per leaf slot (unary count fits 3 bits)

// ---- stage 1: world-space triangles + scene bounds --------------------------------
__device__ __forceinline__ float xf_row(const float4& r, const float3& p)
{
  // same rounding sequence as the oracle's transform_position (no fused ops)
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.x, p.x), __fmul_rn(r.y, p.y)), __fmul_rn(r.z, p.z)),
                   __fmul_rn(r.w, 1.0f));
}

__device__ __forceinline__ void atomic_min_f(float* addr, float v)
{
  // valid for any sign: compare as ordered ints
  if (v >= 0.0f)
    atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v)
{
  if (v >= 0.0f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void k_world_tris(const float3* __restrict__ vertices, const uint3* __restrict__ indices,
                             const uint32_t* __restrict__ face_submesh,
                             const uint32_t* __restrict__ face_flags,
                             const fredholm::Matrix3x4* __restrict__ o2w, uint32_t n,
                             float4* __restrict__ wtri, float* __restrict__ bounds6)
{
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  if (f < n) {
    const uint3 idx = indices[f];
    const fredholm::Matrix3x4 m = o2w[face_submesh[f]];
    const uint32_t vid[3] = {idx.x, idx.y, idx.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float3 p = vertices[vid[k]];
      const float x = xf_row(m.m[0], p), y = xf_row(m.m[1], p), z = xf_row(m.m[2], p);
      const float w = k == 0 ? __uint_as_float(f) : (k == 1 ? __uint_as_float(face_flags ? face_flags[f] : 0u) : 0.0f);
      wtri[3ull * f + k] = make_float4(x, y, z, w);
      lo[0] = fminf(lo[0], x);
      lo[1] = fminf(lo[1], y);
      lo[2] = fminf(lo[2], z);
      hi[0] = fmaxf(hi[0], x);
      hi[1] = fmaxf(hi[1], y);
      hi[2] = fmaxf(hi[2], z);
    }
  }
  // warp reduce, one atomic set per warp
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
    }
  }
  if ((threadIdx.x & 31) == 0 && lo[0] <= hi[0]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomic_min_f(bounds6 + a, lo[a]);
      atomic_max_f(bounds6 + 3 + a, hi[a]);
    }
  }
}

// ---- stage 2: Morton codes ------------------------------------------------------------
__device__ __forceinline__ uint64_t spread21(uint64_t x)
{
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void k_morton(const float4* __restrict__ wtri, const float* __restrict__ bounds6, uint32_t n,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const float4 a = wtri[3ull * f], b = wtri[3ull * f + 1], c = wtri[3ull * f + 2];
  const float cx = 0.5f * (fminf(fminf(a.x, b.x), c.x) + fmaxf(fmaxf(a.x, b.x), c.x));
  const float cy = 0.5f * (fminf(fminf(a.y, b.y), c.y) + fmaxf(fmaxf(a.y, b.y), c.y));
  const float cz = 0.5f * (fminf(fminf(a.z, b.z), c.z) + fmaxf(fmaxf(a.z, b.z), c.z));
  const float ex = fmaxf(bounds6[3] - bounds6[0], 1e-30f);
  const float ey = fmaxf(bounds6[4] - bounds6[1], 1e-30f);
  const float ez = fmaxf(bounds6[5] - bounds6[2], 1e-30f);
  // one grid step for all axes keeps cells cubic (better trees for flat scenes)
  const float e = fmaxf(ex, fmaxf(ey, ez));
  const float s = 2097151.0f / e;
  const uint64_t qx = (uint64_t)fminf(fmaxf((cx - bounds6[0]) * s, 0.0f), 2097151.0f);
  const uint64_t qy = (uint64_t)fminf(fmaxf((cy - bounds6[1]) * s, 0.0f), 2097151.0f);
  const uint64_t qz = (uint64_t)fminf(fmaxf((cz - bounds6[2]) * s, 0.0f), 2097151.0f);
  keys[f] = spread21(qx) | (spread21(qy) << 1) | (spread21(qz) << 2);
  vals[f] = f;
}

// ---- stage 3: Karras binary radix tree -------------------------------------------------
// node ids: internal i -> i (0..n-2), leaf j -> (n-1)+j
struct Lbvh {
  uint2* child;        // [n-1] left/right node ids
  uint32_t* parent;    // [2n-1]
  uint2* range;        // [n-1] first,last sorted leaf covered
  float4* lo;          // [2n-1]
  float4* hi;          // [2n-1]
  uint32_t* visit;     // [n-1]
  const uint32_t* leaf_pos;  // [n] position of sorted leaf j in the triangle order the collapse reads
                             // (null: the Morton order itself, LBVH)
  // SAH-cost collapse (optional, null = greedy largest-area-first): per internal node, the split of j slots
  // between its two children (j = 2..8, 3 bits each, bits 0..20) and, for i = 2..7, whether i slots are no better
  // than i - 1 (bits 21..26)
  const uint32_t* sah_decision;
};

__device__ __forceinline__ int delta(const uint64_t* keys, int n, int i, int j)
{
  if (j < 0 || j >= n) return -1;
  const uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll(a ^ b);
}

__global__ void k_karras(const uint64_t* __restrict__ keys, int n, Lbvh t)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int s = lmax / 2; s >= 1; s /= 2)
    if (delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0;
  int tt = l;
  do {
    tt = (tt + 1) / 2;
    if (delta(keys, n, i, i + (s + tt) * d) > dnode) s += tt;
  } while (tt > 1);
  const int gamma = i + s * d + min(d, 0);
  const int first = min(i, j), last = max(i, j);
  const uint32_t left = (first == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
  const uint32_t right = (last == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
  t.child[i] = make_uint2(left, right);
  t.range[i] = make_uint2((uint32_t)first, (uint32_t)last);
  t.parent[left] = i;
  t.parent[right] = i;
  if (i == 0) t.parent[0] = 0xffffffffu;
}

// ---- stage 4: bottom-up fit ----------------------------------------------------------------
__global__ void k_fit(const float4* __restrict__ wtri, const uint32_t* __restrict__ sorted, int n, Lbvh t)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t f = sorted[j];
  const float4 a = wtri[3ull * f], b = wtri[3ull * f + 1], c = wtri[3ull * f + 2];
  float4 lo = make_float4(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z), 0.0f);
  float4 hi = make_float4(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z), 0.0f);
  const uint32_t leaf = n - 1 + j;
  t.lo[leaf] = lo;
  t.hi[leaf] = hi;
  if (n == 1) return;
  uint32_t cur = t.parent[leaf];
  for (;;) {
    __threadfence();
    if (atomicAdd(&t.visit[cur], 1u) == 0u) return;  // sibling not done yet
    const uint2 ch = t.child[cur];
    const volatile float4* vlo = t.lo;
    const volatile float4* vhi = t.hi;
    const float4 l0 = make_float4(vlo[ch.x].x, vlo[ch.x].y, vlo[ch.x].z, 0.f);
    const float4 l1 = make_float4(vlo[ch.y].x, vlo[ch.y].y, vlo[ch.y].z, 0.f);
    const float4 h0 = make_float4(vhi[ch.x].x, vhi[ch.x].y, vhi[ch.x].z, 0.f);
    const float4 h1 = make_float4(vhi[ch.y].x, vhi[ch.y].y, vhi[ch.y].z, 0.f);
    lo = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.0f);
    hi = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.0f);
    t.lo[cur] = lo;
    t.hi[cur] = hi;
    if (cur == 0) return;
    cur = t.parent[cur];
  }
}


// ---- stage 3b/4b (alternative): PLOC, parallel locally-ordered clustering ------------------
// Meister & Bittner 2018.  Bottom-up agglomerative build over the Morton-ordered clusters: every
// cluster looks `radius` positions to either side for the neighbour that gives the smallest
// merged box, mutual nearest neighbours merge, the cluster array is compacted (order kept) and
// the round repeats until one cluster is left.  Tree quality is close to a full SAH sweep (the
// host-SAH bound in profiles/r1_bvh_collapse_experiment.txt) at a few ms per million triangles.
// Node ids follow the LBVH convention (internal 0..n-2 with the root at 0, leaf j at n-1+j), so
// the 8-wide collapse below serves both builders; internal ids are handed out downwards from
// n-2 so that the last merge -- the root -- gets id 0.
constexpr uint32_t kInvalidCluster = 0xffffffffu;
constexpr int kPlocBlock = 256;
constexpr int kPlocMaxRadius = 32;

__global__ void k_leaf_boxes(const float4* __restrict__ wtri, const uint32_t* __restrict__ sorted, int n, Lbvh t,
                             uint32_t* __restrict__ clusters)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t f = sorted[j];
  const float4 a = wtri[3ull * f], b = wtri[3ull * f + 1], c = wtri[3ull * f + 2];
  const uint32_t leaf = n - 1 + j;
  t.lo[leaf] = make_float4(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z), 0.0f);
  t.hi[leaf] = make_float4(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z), 0.0f);
  clusters[j] = leaf;
}

__global__ void __launch_bounds__(kPlocBlock) k_ploc_nearest(const uint32_t* __restrict__ clusters, uint32_t c, int radius, Lbvh t,
                                                             uint32_t* __restrict__ nearest)
{
  __shared__ float s_lo[kPlocBlock + 2 * kPlocMaxRadius][3];
  __shared__ float s_hi[kPlocBlock + 2 * kPlocMaxRadius][3];
  const int base = (int)(blockIdx.x * kPlocBlock) - radius;
  for (int k = threadIdx.x; k < kPlocBlock + 2 * radius; k += kPlocBlock) {
    const int g = base + k;
    if (g >= 0 && g < (int)c) {
      const uint32_t id = clusters[g];
      const float4 l = t.lo[id], h = t.hi[id];
      s_lo[k][0] = l.x, s_lo[k][1] = l.y, s_lo[k][2] = l.z;
      s_hi[k][0] = h.x, s_hi[k][1] = h.y, s_hi[k][2] = h.z;
    }
  }
  __syncthreads();
  const int i = blockIdx.x * kPlocBlock + threadIdx.x;
  if (i >= (int)c) return;
  const int me = threadIdx.x + radius;
  const float lx = s_lo[me][0], ly = s_lo[me][1], lz = s_lo[me][2];
  const float hx = s_hi[me][0], hy = s_hi[me][1], hz = s_hi[me][2];
  // The key (merged area, pair hash) is symmetric in (i, j) bit for bit, so the pair with the globally
  // smallest key is always mutual (progress).  The hash breaks the area ties of regular meshes: with
  // "lower index wins" every cluster of a regular grid would point at its left neighbour and only
  // one pair per round would merge.
  unsigned long long best = ~0ull;
  int best_j = -1;
  const int j0 = max(i - radius, 0), j1 = min(i + radius, (int)c - 1);
  for (int j = j0; j <= j1; ++j) {
    if (j == i) continue;
    const int k = j - base;
    const float ex = fmaxf(hx, s_hi[k][0]) - fminf(lx, s_lo[k][0]);
    const float ey = fmaxf(hy, s_hi[k][1]) - fminf(ly, s_lo[k][1]);
    const float ez = fmaxf(hz, s_hi[k][2]) - fminf(lz, s_lo[k][2]);
    const float a = ex * ey + ey * ez + ez * ex;
    uint32_t h = (uint32_t)min(i, j) * 0x9E3779B1u ^ (uint32_t)max(i, j) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    const unsigned long long key = ((unsigned long long)__float_as_uint(a) << 32) | h;  // a >= 0
    if (key < best) {
      best = key;
      best_j = j;
    }
  }
  nearest[i] = (uint32_t)best_j;
}

__global__ void k_ploc_merge(const uint32_t* __restrict__ clusters, uint32_t c, const uint32_t* __restrict__ nearest,
                             int n, Lbvh t, uint32_t* __restrict__ n_merged, uint32_t* __restrict__ out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const uint32_t j = nearest[i];
  const uint32_t a = clusters[i];
  if (j < c && nearest[j] == i) {
    if (i < j) {
      const uint32_t b = clusters[j];
      const uint32_t id = (uint32_t)(n - 2) - atomicAdd(n_merged, 1u);
      t.child[id] = make_uint2(a, b);
      t.parent[a] = id;
      t.parent[b] = id;
      const float4 la = t.lo[a], lb = t.lo[b], ha = t.hi[a], hb = t.hi[b];
      t.lo[id] = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 0.0f);
      t.hi[id] = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.0f);
      const uint32_t ca = a >= (uint32_t)(n - 1) ? 1u : t.visit[a];
      const uint32_t cb = b >= (uint32_t)(n - 1) ? 1u : t.visit[b];
      t.visit[id] = ca + cb;  // triangles below the node (the LBVH fit counter is free here)
      out[i] = id;
    } else {
      out[i] = kInvalidCluster;
    }
  } else {
    out[i] = a;
  }
}

struct ValidCluster {
  __device__ __forceinline__ bool operator()(const uint32_t& v) const { return v != kInvalidCluster; }
};

// Left-to-right position of the first triangle under `id`: walk to the root adding the sizes of
// the left siblings passed on the way.  Gives every sub-tree a contiguous triangle range.
constexpr uint32_t kPlocMaxDepth = 2048;  // binary levels; deeper trees fall back to the radix tree

__device__ __forceinline__ uint32_t ploc_first(const Lbvh& t, int n, uint32_t id, uint32_t* too_deep)
{
  uint32_t first = 0;
  uint32_t cur = id;
  for (uint32_t steps = 0;; ++steps) {
    const uint32_t p = t.parent[cur];
    if (p == 0xffffffffu) break;
    if (steps >= kPlocMaxDepth) {  // a degenerate chain: give up, the host rebuilds with the radix tree
      *too_deep = 1u;
      break;
    }
    const uint2 ch = t.child[p];
    if (ch.y == cur) first += ch.x >= (uint32_t)(n - 1) ? 1u : t.visit[ch.x];
    cur = p;
  }
  return first;
}

__global__ void k_ploc_ranges(int n, Lbvh t, const uint32_t* __restrict__ sorted, uint32_t* __restrict__ leaf_pos,
                              uint32_t* __restrict__ order, uint32_t* __restrict__ too_deep)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n - 1) return;
  const uint32_t first = ploc_first(t, n, (uint32_t)i, too_deep);
  if (first >= (uint32_t)n) return;  // only possible after a depth bail-out
  if (i < n - 1) {
    t.range[i] = make_uint2(first, first + t.visit[i] - 1u);
  } else {
    leaf_pos[i - (n - 1)] = first;
    order[first] = sorted[i - (n - 1)];
  }
}

// ---- stage 5: collapse to 8-wide + quantise --------------------------------------------------
__device__ __forceinline__ uint32_t tri_count(const Lbvh& t, int n, uint32_t id)
{
  if (id >= (uint32_t)(n - 1)) return 1u;
  const uint2 r = t.range[id];
  return r.y - r.x + 1u;
}
__device__ __forceinline__ uint32_t first_leaf(const Lbvh& t, int n, uint32_t id)
{
  if (id >= (uint32_t)(n - 1)) return t.leaf_pos ? t.leaf_pos[id - (uint32_t)(n - 1)] : id - (uint32_t)(n - 1);
  return t.range[id].x;
}
__device__ __forceinline__ float half_area(const float4& lo, const float4& hi)
{
  const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
  return ex * ey + ey * ez + ez * ex;
}


// ---- stage 5a (optional): which children an 8-wide node should get, by SAH cost ---------------------------
// Ylitie, Karras, Laine 2017, section 4.1: c(n, i) = cheapest way to represent the binary sub-tree n as a
// forest of at most i 8-wide sub-trees,
//   c(n, 1) = A(n) P(n) c_tri                       if n fits a leaf slot (P <= kLeafMax)
//           = A(n) c_node + dist(n, 8)              otherwise (n becomes an 8-wide node)
//   c(n, i) = min(dist(n, i), c(n, i - 1)),         dist(n, j) = min_k c(left, k) + c(right, j - k)
// computed bottom-up (second thread to arrive at a node does it); the collapse then follows the recorded splits
// instead of opening the child with the largest area.  c_tri / c_node is the measured cost ratio of a triangle
// test and a node visit in issue slots (13.3 vs 10.2 warp-instructions on the bench frame: the triangle phase runs
// with few lanes); FRD_SAH_CT overrides it.
struct SahDp {
  float* cost;         // [8 per internal node], entries 1..7
  uint32_t* decision;  // [internal nodes]
  uint32_t* arrive;    // [internal nodes], zeroed
  float c_tri;
};

__global__ void k_sah_dp(int n, Lbvh t, SahDp dp)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || n < 2) return;
  uint32_t cur = t.parent[(uint32_t)(n - 1 + j)];
  for (;;) {
    __threadfence();
    if (atomicAdd(&dp.arrive[cur], 1u) == 0u) return;  // sibling not done yet
    const uint2 ch = t.child[cur];
    float cl[8], cr[8];
    const volatile float* vc = dp.cost;
    for (int side = 0; side < 2; ++side) {
      const uint32_t c = side == 0 ? ch.x : ch.y;
      float* dst = side == 0 ? cl : cr;
      if (c >= (uint32_t)(n - 1)) {
        const float a = half_area(t.lo[c], t.hi[c]) * dp.c_tri;
        for (int i = 1; i < 8; ++i) dst[i] = a;
      } else {
        for (int i = 1; i < 8; ++i) dst[i] = vc[8ull * c + i];
      }
    }
    const float area = half_area(t.lo[cur], t.hi[cur]);
    const uint32_t P = tri_count(t, n, cur);
    float dist[9];
    uint32_t decision = 0;
    for (int jj = 2; jj <= 8; ++jj) {
      float best = 3.0e38f;
      int best_k = 1;
      for (int k = 1; k < jj; ++k) {
        if (k > 7 || jj - k > 7) continue;
        const float v = cl[k] + cr[jj - k];
        if (v < best) {
          best = v;
          best_k = k;
        }
      }
      dist[jj] = best;
      decision |= (uint32_t)best_k << (3 * (jj - 2));
    }
    float c[8];
    c[1] = P <= (uint32_t)kLeafMax ? area * (float)P * dp.c_tri : area + dist[8];
    for (int i = 2; i <= 7; ++i) {
      if (dist[i] < c[i - 1]) {
        c[i] = dist[i];
      } else {
        c[i] = c[i - 1];
        decision |= 1u << (21 + (i - 2));
      }
    }
    for (int i = 1; i < 8; ++i) dp.cost[8ull * cur + i] = c[i];
    dp.decision[cur] = decision;
    if (cur == 0) return;
    cur = t.parent[cur];
  }
}

struct CollapseCounters {
  uint32_t n_nodes;
  uint32_t n_tris;
};

// one 8-wide node: n8 is built from the binary sub-tree rooted at work[n8]
__device__ void collapse_node(uint32_t n8, uint32_t* __restrict__ work, int n, const Lbvh& t,
                              const float4* __restrict__ wtri, const uint32_t* __restrict__ sorted,
                              CollapseCounters* __restrict__ counters, Node8* __restrict__ nodes,
                              float4* __restrict__ tris)
{
  const uint32_t b2 = work[n8];

  uint32_t c[8];
  int nc;
  const bool root_is_cluster = tri_count(t, n, b2) <= (uint32_t)kLeafMax;  // only for tiny scenes
  if (root_is_cluster) {
    c[0] = b2;
    nc = 1;
  } else {
    const uint2 ch = t.child[b2];
    c[0] = ch.x;
    c[1] = ch.y;
    nc = 2;
    if (t.sah_decision) {
      // follow the splits recorded by k_sah_dp: (node, slots) pairs on a small stack
      nc = 0;
      uint32_t st_node[16];
      int st_slots[16];
      int sp = 0;
      st_node[sp] = b2;
      st_slots[sp++] = 8;
      bool root = true;
      while (sp > 0) {
        const uint32_t m = st_node[--sp];
        int i = st_slots[sp];
        const bool is_binary_leaf = m >= (uint32_t)(n - 1);
        uint32_t d = 0;
        if (!is_binary_leaf) {
          d = t.sah_decision[m];
          if (!root)
            while (i > 1 && (d >> (21 + (i - 2))) & 1u) i--;
        }
        if (is_binary_leaf || (i == 1 && !root)) {
          c[nc++] = m;
          continue;
        }
        root = false;
        const int k = (int)((d >> (3 * (i - 2))) & 7u);
        const uint2 mc = t.child[m];
        // right first so that the left sub-tree is emitted first
        st_node[sp] = mc.y;
        st_slots[sp++] = i - k;
        st_node[sp] = mc.x;
        st_slots[sp++] = k;
      }
    }
    while (!t.sah_decision && nc < 8) {
      int best = -1;
      float best_area = -1.0f;
      for (int k = 0; k < nc; ++k) {
        if (tri_count(t, n, c[k]) > (uint32_t)kLeafMax) {
          const float a = half_area(t.lo[c[k]], t.hi[c[k]]);
          if (a > best_area) {
            best_area = a;
            best = k;
          }
        }
      }
      if (best < 0) break;
      const uint2 ch2 = t.child[c[best]];
      c[best] = ch2.x;
      c[nc++] = ch2.y;
    }
#ifndef FR_NO_LEAF_SPLIT
    // slots left over: split multi-triangle leaves (largest first) so that their triangles get
    // boxes of their own -- no extra node, fewer triangle tests
    while (nc < 8) {
      int best = -1;
      float best_area = -1.0f;
      for (int k = 0; k < nc; ++k) {
        const uint32_t cnt = tri_count(t, n, c[k]);
        if (cnt > 1u && cnt <= (uint32_t)kLeafMax) {
          const float a = half_area(t.lo[c[k]], t.hi[c[k]]);
          if (a > best_area) {
            best_area = a;
            best = k;
          }
        }
      }
      if (best < 0) break;
      const uint2 ch2 = t.child[c[best]];
      c[best] = ch2.x;
      c[nc++] = ch2.y;
    }
#endif
  }

  const float4 nlo = t.lo[b2], nhi = t.hi[b2];
  const float3 ncen = f3(0.5f * (nlo.x + nhi.x), 0.5f * (nlo.y + nhi.y), 0.5f * (nlo.z + nhi.z));

  // greedy octant-order slot assignment (Ylitie et al. sec. 4.2)
  float cost[8][8];
  for (int k = 0; k < nc; ++k) {
    const float4 l = t.lo[c[k]], h = t.hi[c[k]];
    const float3 dc = f3(0.5f * (l.x + h.x) - ncen.x, 0.5f * (l.y + h.y) - ncen.y, 0.5f * (l.z + h.z) - ncen.z);
    for (int s = 0; s < 8; ++s) {
      cost[k][s] = ((s & 1) ? dc.x : -dc.x) + ((s & 2) ? dc.y : -dc.y) + ((s & 4) ? dc.z : -dc.z);
    }
  }
  // nodes with at most four children keep them in slots 0-3 (ordered along x and y only), so that
  // the traversal can skip the empty upper half of the node test (bvh.cuh, node_phase)
  const int n_slots = nc <= 4 ? 4 : 8;
  int slot_of[8];
  int child_at[8];
  for (int s = 0; s < 8; ++s) child_at[s] = -1;
  for (int k = 0; k < 8; ++k) slot_of[k] = -1;
  for (int it = 0; it < nc; ++it) {
    int bk = -1, bs = -1;
    float bc = -3.0e38f;
    for (int k = 0; k < nc; ++k) {
      if (slot_of[k] >= 0) continue;
      for (int s = 0; s < n_slots; ++s) {
        if (child_at[s] >= 0) continue;
        if (cost[k][s] > bc) {
          bc = cost[k][s];
          bk = k;
          bs = s;
        }
      }
    }
    slot_of[bk] = bs;
    child_at[bs] = bk;
  }

  uint32_t n_inner = 0, n_leaf_tris = 0;
  for (int k = 0; k < nc; ++k) {
    const uint32_t cnt = tri_count(t, n, c[k]);
    if (cnt > (uint32_t)kLeafMax)
      n_inner++;
    else
      n_leaf_tris += cnt;
  }
  const uint32_t cbase = n_inner ? atomicAdd(&counters->n_nodes, n_inner) : 0u;
  const uint32_t tbase = n_leaf_tris ? atomicAdd(&counters->n_tris, n_leaf_tris) : 0u;

  Node8 out;
  out.pad_[0] = out.pad_[1] = out.pad_[2] = out.pad_[3] = 0u;
  out.px = nlo.x;
  out.py = nlo.y;
  out.pz = nlo.z;
  const uint32_t ex = grid_exponent(__fsub_ru(nhi.x, nlo.x)), ey = grid_exponent(__fsub_ru(nhi.y, nlo.y)),
                 ez = grid_exponent(__fsub_ru(nhi.z, nlo.z));
  out.ex = (uint8_t)ex;
  out.ey = (uint8_t)ey;
  out.ez = (uint8_t)ez;
  out.imask = 0;
  out.child_base = cbase;
  out.tri_base = tbase;
  // 1 / 2^(e-127) = 2^(127-e) -> biased exponent 254 - e
  const float isx = grid_inverse_step(ex), isy = grid_inverse_step(ey), isz = grid_inverse_step(ez);
  uint32_t rank = 0, toff = 0;
  for (int s = 0; s < 8; ++s) {
    const int k = child_at[s];
    if (k < 0) {
      out.meta[s] = 0;
      out.qlox[s] = out.qloy[s] = out.qloz[s] = 255;
      out.qhix[s] = out.qhiy[s] = out.qhiz[s] = 0;
      continue;
    }
    const uint32_t id = c[k];
    const uint32_t cnt = tri_count(t, n, id);
    const float4 l = t.lo[id], h = t.hi[id];
    // conservative: lower bounds round down, upper bounds round up
    out.qlox[s] = quantize_lo(l.x, nlo.x, isx);
    out.qloy[s] = quantize_lo(l.y, nlo.y, isy);
    out.qloz[s] = quantize_lo(l.z, nlo.z, isz);
    out.qhix[s] = quantize_hi(h.x, nlo.x, isx);
    out.qhiy[s] = quantize_hi(h.y, nlo.y, isy);
    out.qhiz[s] = quantize_hi(h.z, nlo.z, isz);
    if (cnt > (uint32_t)kLeafMax) {
      out.imask |= (uint8_t)(1u << s);
      out.meta[s] = (uint8_t)(0x20u | (24u + s));
      work[cbase + rank] = id;
      rank++;
    } else {
      const uint32_t unary = cnt == 1 ? 1u : (cnt == 2 ? 3u : 7u);
      out.meta[s] = (uint8_t)((unary << 5) | toff);
      const uint32_t first = first_leaf(t, n, id);
      for (uint32_t q = 0; q < cnt; ++q) {
        const uint32_t f = sorted[first + q];
        float4* dst = tris + 3ull * (tbase + toff + q);
        dst[0] = wtri[3ull * f];
        dst[1] = wtri[3ull * f + 1];
        dst[2] = wtri[3ull * f + 2];
      }
      toff += cnt;
    }
  }
  nodes[n8] = out;
}

// one level of the 8-wide tree per launch (the host reads the node counter between levels)
__global__ void k_collapse(uint32_t level_begin, uint32_t level_end, uint32_t* __restrict__ work, int n,
                           Lbvh t, const float4* __restrict__ wtri, const uint32_t* __restrict__ sorted,
                           CollapseCounters* __restrict__ counters, Node8* __restrict__ nodes,
                           float4* __restrict__ tris)
{
  const uint32_t n8 = level_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (n8 >= level_end) return;
  collapse_node(n8, work, n, t, wtri, sorted, counters, nodes, tris);
}

// small trees (an instance tree of a few thousand boxes): all levels in ONE launch of one block, which walks the
// levels itself -- no host round trip per level.  depth_out receives the number of levels.
constexpr int kCollapseSmallThreads = 1024;
constexpr int kSmallTree = 16384;  // primitives
constexpr int kMaxLevels = kSmemStack + kLocalStack;
__global__ void __launch_bounds__(kCollapseSmallThreads) k_collapse_small(uint32_t* __restrict__ work, int n, Lbvh t,
                                                                         const float4* __restrict__ wtri,
                                                                         const uint32_t* __restrict__ sorted,
                                                                         CollapseCounters* __restrict__ counters,
                                                                         Node8* __restrict__ nodes, float4* __restrict__ tris,
                                                                         uint32_t max_nodes, uint32_t* __restrict__ depth_out)
{
  // depth_out[0] = number of levels, depth_out[1 + k] = first node of level k, depth_out[1 + levels] = node count
  __shared__ uint32_t s_begin, s_end;
  if (threadIdx.x == 0) {
    s_begin = 0;
    s_end = 1;
  }
  __syncthreads();
  uint32_t depth = 0;
  while (s_begin < s_end && s_end <= max_nodes) {
    if (threadIdx.x == 0 && depth < (uint32_t)kMaxLevels) depth_out[1 + depth] = s_begin;
    for (uint32_t n8 = s_begin + threadIdx.x; n8 < s_end; n8 += blockDim.x)
      collapse_node(n8, work, n, t, wtri, sorted, counters, nodes, tris);
    __threadfence_block();
    __syncthreads();
    if (threadIdx.x == 0) {
      s_begin = s_end;
      s_end = counters->n_nodes;
    }
    depth++;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    depth_out[0] = depth;
    if (depth <= (uint32_t)kMaxLevels) depth_out[1 + depth] = s_begin;
  }
}

__global__ void k_empty_root(Node8* nodes)
{
  Node8 out = {};
  out.ex = out.ey = out.ez = 127;
  for (int s = 0; s < 8; ++s) {
    out.qlox[s] = out.qloy[s] = out.qloz[s] = 255;
  }
  nodes[0] = out;
}

struct TreeTooDeep : std::runtime_error {
  TreeTooDeep() : std::runtime_error("bvh: tree too deep for the traversal stack") {}
};
// a PLOC round that merged nothing (cannot happen with the symmetric tie break, kept as a guard): the
// caller rebuilds with the radix tree, like TreeTooDeep
struct PlocStalled : std::runtime_error {
  PlocStalled() : std::runtime_error("bvh ploc: no progress") {}
};

// builder selection (experiments / fallback): FRD_BVH_BUILDER=lbvh|ploc, FRD_PLOC_RADIUS=1..32
bool builder_is_ploc()
{
  const char* e = getenv("FRD_BVH_BUILDER");
  return !(e && strcmp(e, "lbvh") == 0);
}
// k_sah_dp chooses the children of the 8-wide nodes (FRD_COLLAPSE=greedy: largest area first, the round-1 rule);
// FRD_SAH_CT = cost of a triangle test relative to a node visit.  Measured on the bench frame
// (profiles/r2j_collapse_sah.txt): 1.0 is the flat optimum, 0.3 ... 1.5 all beat the greedy rule.
bool collapse_by_sah()
{
  const char* e = getenv("FRD_COLLAPSE");
  return !(e && strcmp(e, "greedy") == 0);
}
float sah_tri_cost()
{
  const char* e = getenv("FRD_SAH_CT");
  const float v = e ? (float)atof(e) : 1.0f;
  return v > 0.0f ? v : 1.0f;
}
// Search radius 4: with the SAH-cost collapse a smaller window gives the better 8-wide tree (bench scene: 12.30 / 10.62
// node visits per radiance / visibility ray against 12.60 / 11.13 at radius 8, frame +2.1 %; the 52 M-triangle scene
// +1.2 %; profiles/r2s_ploc_radius.txt) and fewer, cheaper rounds.  With the round-1 largest-area-first collapse
// the sweep was flat.
int ploc_radius()
{
  const char* e = getenv("FRD_PLOC_RADIUS");
  const int r = e ? atoi(e) : 4;
  return r < 1 ? 1 : (r > kPlocMaxRadius ? kPlocMaxRadius : r);
}

}  // namespace

namespace
{
// FRD_BVH_VERBOSE=1: host wall time per build phase (synchronises at every mark)
struct PhaseClock {
  bool on = getenv("FRD_BVH_VERBOSE") != nullptr;
  cudaStream_t s;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseClock(cudaStream_t st) : s(st), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* what)
  {
    if (!on) return;
    cudaStreamSynchronize(s);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[bvh] %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};
}  // namespace

namespace
{
// Build temporaries come from a PRIVATE stream-ordered memory pool per device (cudaMemPoolCreate +
// cudaMallocFromPoolAsync) with the release threshold lifted, so a rebuild -- set_time on an animated scene --
// reuses the pool's memory instead of paying cudaMalloc / cudaFree for every buffer (these calls were 20-90 ms
// of a 52 M-triangle build).  The device's default pool, which the host application may share (e.g. PyTorch's
// cudaMallocAsync backend), is never touched.  What the pool keeps between builds: up to FRD_BVH_POOL_KEEP_GB
// (default 16, 0 = give everything back), never more than a tenth of the device's memory -- mapping 10 GB of
// fresh pool memory costs more than the 52 M-triangle build that uses it (1.7 s against 85 ms).
thread_local cudaStream_t g_scratch_stream = nullptr;
thread_local cudaMemPool_t g_scratch_pool = nullptr;

constexpr int kMaxDevices = 64;
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[kMaxDevices] = {};
unsigned g_builds[kMaxDevices] = {};

int current_device()
{
  int dev = 0;
  FR_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) throw std::runtime_error("bvh: device index out of range");
  return dev;
}

void init_scratch_pool()
{
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (!g_pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    FR_CUDA_CHECK(cudaMemPoolCreate(&g_pools[dev], &props));
    uint64_t keep = UINT64_MAX;
    FR_CUDA_CHECK(cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
  }
  g_scratch_pool = g_pools[dev];
}

// after a build: give the temporaries back (first build) or keep them for the next rebuild
void trim_scratch_pool()
{
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (!g_pools[dev]) return;
  const char* e = getenv("FRD_BVH_POOL_KEEP_GB");
  const size_t keep_gb = e ? (size_t)strtoull(e, nullptr, 0) : 16;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = 0;
  g_builds[dev]++;
  cudaMemPoolTrimTo(g_pools[dev], std::min(keep_gb << 30, total_b / 10));
}

template <typename T>
class ScratchBuf
{
 public:
  ScratchBuf() = default;
  explicit ScratchBuf(size_t n) { alloc(n); }
  ScratchBuf(const ScratchBuf&) = delete;
  ScratchBuf& operator=(const ScratchBuf&) = delete;
  ~ScratchBuf() { release(); }
  void alloc(size_t n)
  {
    release();
    if (n) FR_CUDA_CHECK(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&p_), n * sizeof(T), g_scratch_pool, g_scratch_stream));
    n_ = n;
  }
  void release()
  {
    if (p_) cudaFreeAsync(p_, g_scratch_stream);
    p_ = nullptr;
    n_ = 0;
  }
  void zero(cudaStream_t s)
  {
    if (n_) FR_CUDA_CHECK(cudaMemsetAsync(p_, 0, n_ * sizeof(T), s));
  }
  T* get() const { return p_; }
  size_t size() const { return n_; }

 private:
  T* p_ = nullptr;
  size_t n_ = 0;
};

void build_bvh_with(bool use_ploc, cudaStream_t stream, const float3* d_vertices, const uint3* d_indices,
                    const uint32_t* d_face_submesh, const uint32_t* d_face_flags,
                    const fredholm::Matrix3x4* d_o2w, uint32_t n_faces, DeviceBvh& out)
{
  out.n_faces = n_faces;
  out.depth = 1;
  if (n_faces == 0) {
    out.nodes.alloc(1);
    out.tris.alloc(3);
    out.n_nodes = 1;
    k_empty_root<<<1, 1, 0, stream>>>(out.nodes.get());
    FR_CUDA_LAUNCH_CHECK();
    FR_CUDA_CHECK(cudaStreamSynchronize(stream));
    return;
  }
  const int n = (int)n_faces;
  const int B = 256;
  const int G = (n + B - 1) / B;

  g_scratch_stream = stream;
  init_scratch_pool();
  PhaseClock clk(stream);
  ScratchBuf<float4> wtri(3ull * n);
  ScratchBuf<float> bounds(6);
  const float init_b[6] = {3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
  FR_CUDA_CHECK(cudaMemcpyAsync(bounds.get(), init_b, sizeof(init_b), cudaMemcpyHostToDevice, stream));
  k_world_tris<<<G, B, 0, stream>>>(d_vertices, d_indices, d_face_submesh, d_face_flags, d_o2w, n_faces,
                                    wtri.get(), bounds.get());
  FR_CUDA_LAUNCH_CHECK();

  ScratchBuf<uint64_t> keys(n), keys_sorted(n);
  ScratchBuf<uint32_t> vals(n), sorted(n);
  k_morton<<<G, B, 0, stream>>>(wtri.get(), bounds.get(), n_faces, keys.get(), vals.get());
  FR_CUDA_LAUNCH_CHECK();
  size_t tmp_bytes = 0;
  FR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.get(), keys_sorted.get(), vals.get(),
                                                sorted.get(), n, 0, 63, stream));
  ScratchBuf<unsigned char> tmp(tmp_bytes);
  FR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.get(), tmp_bytes, keys.get(), keys_sorted.get(), vals.get(),
                                                sorted.get(), n, 0, 63, stream));

  clk.mark("sort");
  const int n_int = n > 1 ? n - 1 : 1;
  ScratchBuf<uint2> child(n_int), range(n_int);
  ScratchBuf<uint32_t> parent(2ull * n), visit(n_int);
  ScratchBuf<float4> lo(2ull * n), hi(2ull * n);
  visit.zero(stream);
  Lbvh t{child.get(), parent.get(), range.get(), lo.get(), hi.get(), visit.get(), nullptr};
  ScratchBuf<uint32_t> leaf_pos, order;
  const uint32_t* tri_order = sorted.get();
  out.ploc_rounds = 0;
  if (use_ploc && n > 2) {
    // ---- PLOC ----
    const int radius = ploc_radius();
    ScratchBuf<uint32_t> cl_a(n), cl_b(n), nearest(n), counters2(3);
    counters2.zero(stream);
    uint32_t* n_merged = counters2.get();
    uint32_t* n_selected = counters2.get() + 1;
    clk.mark("ploc alloc");
    k_leaf_boxes<<<G, B, 0, stream>>>(wtri.get(), sorted.get(), n, t, cl_a.get());
    FR_CUDA_LAUNCH_CHECK();
    size_t sel_bytes = 0;
    FR_CUDA_CHECK(cub::DeviceSelect::If(nullptr, sel_bytes, cl_b.get(), cl_a.get(), n_selected, n, ValidCluster{}, stream));
    ScratchBuf<unsigned char> sel_tmp(sel_bytes);
    uint32_t c = (uint32_t)n;
    uint32_t* cur = cl_a.get();
    uint32_t* nxt = cl_b.get();
    while (c > 1) {
      const int g = (int)((c + kPlocBlock - 1) / kPlocBlock);
      k_ploc_nearest<<<g, kPlocBlock, 0, stream>>>(cur, c, radius, t, nearest.get());
      k_ploc_merge<<<g, kPlocBlock, 0, stream>>>(cur, c, nearest.get(), n, t, n_merged, nxt);
      FR_CUDA_LAUNCH_CHECK();
      // compact (order kept) back into `cur`
      FR_CUDA_CHECK(cub::DeviceSelect::If(sel_tmp.get(), sel_bytes, nxt, cur, n_selected, (int)c, ValidCluster{}, stream));
      uint32_t c_next = 0;
      FR_CUDA_CHECK(cudaMemcpyAsync(&c_next, n_selected, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      FR_CUDA_CHECK(cudaStreamSynchronize(stream));
      if (c_next >= c) throw PlocStalled();
      c = c_next;
      out.ploc_rounds++;
    }
    clk.mark("ploc loop");
    if (getenv("FRD_BVH_VERBOSE")) fprintf(stderr, "[bvh] ploc radius %d: %u rounds\n", radius, out.ploc_rounds);
    const uint32_t none = 0xffffffffu;
    FR_CUDA_CHECK(cudaMemcpyAsync(parent.get(), &none, sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    leaf_pos.alloc(n);
    order.alloc(n);
    k_ploc_ranges<<<(2 * n - 1 + B - 1) / B, B, 0, stream>>>(n, t, sorted.get(), leaf_pos.get(), order.get(),
                                                             counters2.get() + 2);
    FR_CUDA_LAUNCH_CHECK();
    uint32_t too_deep = 0;
    FR_CUDA_CHECK(cudaMemcpyAsync(&too_deep, counters2.get() + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    FR_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (too_deep) throw TreeTooDeep();
    t.leaf_pos = leaf_pos.get();
    tri_order = order.get();
  } else {
    if (n > 1) {
      k_karras<<<(n - 1 + B - 1) / B, B, 0, stream>>>(keys_sorted.get(), n, t);
      FR_CUDA_LAUNCH_CHECK();
    }
    k_fit<<<G, B, 0, stream>>>(wtri.get(), sorted.get(), n, t);
    FR_CUDA_LAUNCH_CHECK();
  }

  clk.mark("binary");
  // optional: choose the children of every 8-wide node by SAH cost instead of largest-area-first
  ScratchBuf<float> sah_cost;
  ScratchBuf<uint32_t> sah_decision, sah_arrive;
  if (n > 2 && collapse_by_sah()) {
    sah_cost.alloc(8ull * n_int);
    sah_decision.alloc(n_int);
    sah_arrive.alloc(n_int);
    sah_arrive.zero(stream);
    k_sah_dp<<<G, B, 0, stream>>>(n, t, SahDp{sah_cost.get(), sah_decision.get(), sah_arrive.get(), sah_tri_cost()});
    FR_CUDA_LAUNCH_CHECK();
    t.sah_decision = sah_decision.get();
    clk.mark("sah dp");
  }
  // collapse, level by level; node n8 is built from binary node work[n8]
  const size_t max_nodes = (size_t)n / 2 + 2;
  ScratchBuf<Node8> nodes(max_nodes);
  ScratchBuf<uint32_t> work(max_nodes);
  out.tris.reserve(3ull * n);  // grow-only: a rebuild of the same scene keeps its buffers
  ScratchBuf<CollapseCounters> counters(1);
  const CollapseCounters init_c{1u, 0u};
  const uint32_t root_id = n > 1 ? 0u : 0u;  // n == 1: leaf 0 has node id (n-1)+0 = 0
  FR_CUDA_CHECK(cudaMemcpyAsync(counters.get(), &init_c, sizeof(init_c), cudaMemcpyHostToDevice, stream));
  FR_CUDA_CHECK(cudaMemcpyAsync(work.get(), &root_id, sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  uint32_t begin = 0, end = 1, depth = 0;
  if (n <= kSmallTree) {
    // Instance trees and other small inputs: ONE launch for all levels and ONE host round trip for the whole
    // build -- the node pool is copied at its capacity (a few hundred KB), so nothing has to be known on the host
    // before the final synchronisation.  This is the per-frame cost of an animated two-level scene.
    ScratchBuf<uint32_t> d_depth(kMaxLevels + 2);
    k_collapse_small<<<1, kCollapseSmallThreads, 0, stream>>>(work.get(), n, t, wtri.get(), tri_order, counters.get(), nodes.get(),
                                                               out.tris.get(), (uint32_t)max_nodes, d_depth.get());
    FR_CUDA_LAUNCH_CHECK();
    out.nodes.reserve(max_nodes);
    FR_CUDA_CHECK(cudaMemcpyAsync(out.nodes.get(), nodes.get(), sizeof(Node8) * max_nodes, cudaMemcpyDeviceToDevice, stream));
    struct {
      CollapseCounters c;
      uint32_t depth[kMaxLevels + 2];
      float bounds[6];
    } h;
    FR_CUDA_CHECK(cudaMemcpyAsync(&h.c, counters.get(), sizeof(h.c), cudaMemcpyDeviceToHost, stream));
    FR_CUDA_CHECK(cudaMemcpyAsync(h.depth, d_depth.get(), sizeof(h.depth), cudaMemcpyDeviceToHost, stream));
    FR_CUDA_CHECK(cudaMemcpyAsync(h.bounds, bounds.get(), sizeof(h.bounds), cudaMemcpyDeviceToHost, stream));
    FR_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (h.c.n_nodes > max_nodes) throw std::runtime_error("bvh collapse: node pool overflow");
    out.depth = h.depth[0];
    out.n_nodes = h.c.n_nodes;
    out.level_begin.clear();
    if (out.depth <= (uint32_t)kMaxLevels) out.level_begin.assign(h.depth + 1, h.depth + 2 + out.depth);
    for (int a = 0; a < 3; ++a) {
      out.bounds_lo[a] = h.bounds[a];
      out.bounds_hi[a] = h.bounds[3 + a];
    }
    clk.mark("collapse+finish");
    if (out.depth + 2 > (uint32_t)(kSmemStack + kLocalStack)) throw TreeTooDeep();
    if (use_ploc && getenv("FRD_PLOC_FORCE_FALLBACK")) throw TreeTooDeep();  // test hook for the fallback path
    return;
  }
  while (begin < end) {
    const uint32_t cnt = end - begin;
    k_collapse<<<(cnt + 63) / 64, 64, 0, stream>>>(begin, end, work.get(), n, t, wtri.get(), tri_order,
                                                    counters.get(), nodes.get(), out.tris.get());
    FR_CUDA_LAUNCH_CHECK();
    CollapseCounters h;
    FR_CUDA_CHECK(cudaMemcpyAsync(&h, counters.get(), sizeof(h), cudaMemcpyDeviceToHost, stream));
    FR_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (h.n_nodes > max_nodes) throw std::runtime_error("bvh collapse: node pool overflow");
    begin = end;
    end = h.n_nodes;
    depth++;
  }
  clk.mark("collapse");
  out.depth = depth;
  out.n_nodes = end;
  out.level_begin.clear();
  // shrink the node pool to its final size
  out.nodes.reserve(end);
  FR_CUDA_CHECK(cudaMemcpyAsync(out.nodes.get(), nodes.get(), sizeof(Node8) * end, cudaMemcpyDeviceToDevice, stream));
  float hb[6];
  FR_CUDA_CHECK(cudaMemcpyAsync(hb, bounds.get(), sizeof(hb), cudaMemcpyDeviceToHost, stream));
  FR_CUDA_CHECK(cudaStreamSynchronize(stream));
  for (int a = 0; a < 3; ++a) {
    out.bounds_lo[a] = hb[a];
    out.bounds_hi[a] = hb[3 + a];
  }
  clk.mark("finish");
  if (depth + 2 > (uint32_t)(kSmemStack + kLocalStack)) throw TreeTooDeep();
  if (use_ploc && getenv("FRD_PLOC_FORCE_FALLBACK")) throw TreeTooDeep();  // test hook for the fallback path
}
}  // namespace

void build_bvh(cudaStream_t stream, const float3* d_vertices, const uint3* d_indices,
               const uint32_t* d_face_submesh, const uint32_t* d_face_flags,
               const fredholm::Matrix3x4* d_o2w, uint32_t n_faces, DeviceBvh& out, int builder)
{
  const bool ploc = builder < 0 ? builder_is_ploc() : builder == 1;
  // a 52 M-triangle build uses ~10 GB of temporaries (trim_scratch_pool: what stays cached afterwards)
  struct TrimPool {
    ~TrimPool() { trim_scratch_pool(); }
  } trim_on_exit;
  auto rebuild_with_radix_tree = [&](const char* why) {
    // an agglomerative tree has no depth bound; the radix tree's depth is bounded by the 63 key bits
    if (!ploc) throw;
    if (getenv("FRD_BVH_VERBOSE")) fprintf(stderr, "[bvh] ploc %s: rebuilding with the radix tree\n", why);
    build_bvh_with(false, stream, d_vertices, d_indices, d_face_submesh, d_face_flags, d_o2w, n_faces, out);
  };
  try {
    build_bvh_with(ploc, stream, d_vertices, d_indices, d_face_submesh, d_face_flags, d_o2w, n_faces, out);
  } catch (const TreeTooDeep&) {
    rebuild_with_radix_tree("tree too deep");
  } catch (const PlocStalled&) {
    rebuild_with_radix_tree("made no progress");
  }
}

}  // namespace frd
