// Vertex expansion, face-normal generation and de-duplication on the GPU (mesh_prep.h).
//
//  k_expand    one thread per triangle: gathers position / normal / texcoord of its three corners from the file's
//              attribute pools; a triangle without three normals gets the face normal of its normalised edges, one
//              without three texcoords gets (0,0),(1,0),(0,1) (scene.cpp:334-378).  Every operation of the normal
//              is a single IEEE op (_rn intrinsics, no FMA) so that the floats equal the host loop's.
//  k_insert    one thread per corner: open-addressing hash table of corner indices; a slot's value is the SMALLEST
//              corner index whose 8-float record compares equal (float ==, so -0 == +0 and NaN equals nothing --
//              the semantics of the reference's Vertex::operator==).
//  k_resolve   representative of every corner = that smallest index = its first occurrence in file order.
//  scan + k_emit  representatives are numbered in corner order (exclusive sum), which is exactly the order in
//              which the serial loop would have appended them; indices follow.
#include <cub/device/device_scan.cuh>

#include <stdexcept>

#include "cuda_util.h"
#include "mesh_prep.h"

namespace frd
{
namespace
{

struct alignas(16) CornerRecord {
  float v[8];  // position, normal, texcoord
};

constexpr uint32_t kEmpty = 0xffffffffu;

__device__ __forceinline__ float3 normalize_rn(float x, float y, float z)
{
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float inv = __fdiv_rn(1.0f, __fsqrt_rn(d));
  return make_float3(__fmul_rn(x, inv), __fmul_rn(y, inv), __fmul_rn(z, inv));
}

__global__ void k_expand(const float* __restrict__ pos, uint32_t n_pos, const float* __restrict__ nrm, uint32_t n_nrm,
                         const float* __restrict__ tex, uint32_t n_tex, const ObjCorner* __restrict__ corners,
                         uint32_t n_faces, CornerRecord* __restrict__ rec, uint32_t* __restrict__ error)
{
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_faces) return;
  float P[3][3], N[3][3], T[3][2];
  int n_normals = 0, n_texcoords = 0;
  for (int k = 0; k < 3; ++k) {
    const ObjCorner c = corners[3ull * f + k];
    if (c.v < 0 || 3ull * (uint32_t)c.v + 2 >= n_pos) {
      atomicExch(error, 1u);
      return;
    }
    for (int a = 0; a < 3; ++a) P[k][a] = pos[3ull * c.v + a];
    if (c.vn >= 0 && 3ull * (uint32_t)c.vn + 2 < n_nrm) {
      for (int a = 0; a < 3; ++a) N[n_normals][a] = nrm[3ull * c.vn + a];
      n_normals++;
    }
    if (c.vt >= 0 && 2ull * (uint32_t)c.vt + 1 < n_tex) {
      for (int a = 0; a < 2; ++a) T[n_texcoords][a] = tex[2ull * c.vt + a];
      n_texcoords++;
    }
  }
  if (n_normals < 3) {
    const float3 e1 = normalize_rn(__fsub_rn(P[1][0], P[0][0]), __fsub_rn(P[1][1], P[0][1]), __fsub_rn(P[1][2], P[0][2]));
    const float3 e2 = normalize_rn(__fsub_rn(P[2][0], P[0][0]), __fsub_rn(P[2][1], P[0][1]), __fsub_rn(P[2][2], P[0][2]));
    const float cx = __fsub_rn(__fmul_rn(e1.y, e2.z), __fmul_rn(e1.z, e2.y));
    const float cy = __fsub_rn(__fmul_rn(e1.z, e2.x), __fmul_rn(e1.x, e2.z));
    const float cz = __fsub_rn(__fmul_rn(e1.x, e2.y), __fmul_rn(e1.y, e2.x));
    float3 n = normalize_rn(cx, cy, cz);
    // a degenerate triangle gives 0 * inf = NaN: the x86 host loop stores the SSE default NaN (0xffc00000), the GPU
    // would store 0x7fffffff -- keep the arrays bit-identical
    if (n.x != n.x) n.x = __uint_as_float(0xffc00000u);
    if (n.y != n.y) n.y = __uint_as_float(0xffc00000u);
    if (n.z != n.z) n.z = __uint_as_float(0xffc00000u);
    for (int k = 0; k < 3; ++k) N[k][0] = n.x, N[k][1] = n.y, N[k][2] = n.z;
  }
  if (n_texcoords < 3) {
    T[0][0] = 0.f, T[0][1] = 0.f;
    T[1][0] = 1.f, T[1][1] = 0.f;
    T[2][0] = 0.f, T[2][1] = 1.f;
  }
  for (int k = 0; k < 3; ++k) {
    CornerRecord r;
    r.v[0] = P[k][0], r.v[1] = P[k][1], r.v[2] = P[k][2];
    r.v[3] = N[k][0], r.v[4] = N[k][1], r.v[5] = N[k][2];
    r.v[6] = T[k][0], r.v[7] = T[k][1];
    rec[3ull * f + k] = r;
  }
}

__device__ __forceinline__ bool same_vertex(const CornerRecord& a, const CornerRecord& b)
{
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (!(a.v[i] == b.v[i])) return false;
  return true;
}

__device__ __forceinline__ uint32_t hash_vertex(const CornerRecord& r)
{
  uint64_t h = 1469598103934665603ull;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float f = r.v[i] == 0.0f ? 0.0f : r.v[i];  // -0 and +0 compare equal, so they must hash alike
    h = (h ^ __float_as_uint(f)) * 1099511628211ull;
  }
  return (uint32_t)(h ^ (h >> 32));
}

__global__ void k_insert(const CornerRecord* __restrict__ rec, uint32_t n, uint32_t* table, uint32_t mask,
                         uint32_t* __restrict__ slot_of)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const CornerRecord mine = rec[i];
  uint32_t h = hash_vertex(mine) & mask;
  for (;;) {
    uint32_t cur = table[h];
    if (cur == kEmpty) {
      cur = atomicCAS(&table[h], kEmpty, i);
      if (cur == kEmpty) break;  // this corner opened the slot
    }
    // every corner that ever owns a slot carries the same vertex, so comparing with whoever owns it now is enough
    if (same_vertex(rec[cur], mine)) {
      atomicMin(&table[h], i);
      break;
    }
    h = (h + 1u) & mask;
  }
  slot_of[i] = h;
}

__global__ void k_resolve(const uint32_t* __restrict__ table, const uint32_t* __restrict__ slot_of, uint32_t n,
                          uint32_t* __restrict__ rep, uint32_t* __restrict__ is_rep)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = table[slot_of[i]];
  rep[i] = r;
  is_rep[i] = r == i ? 1u : 0u;
}

__global__ void k_emit(const CornerRecord* __restrict__ rec, const uint32_t* __restrict__ rep,
                       const uint32_t* __restrict__ new_id, uint32_t n, float3* __restrict__ vertices,
                       float3* __restrict__ normals, float2* __restrict__ texcoords, uint32_t* __restrict__ indices)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = rep[i];
  const uint32_t id = new_id[r];
  const CornerRecord c = rec[i];
  bool has_nan = false;
#pragma unroll
  for (int a = 0; a < 8; ++a) has_nan = has_nan || c.v[a] != c.v[a];
  // quirk of the reference's map lookup (scene.cpp:380-387, see csrc/scene.cpp): a NaN vertex is appended but its
  // corner refers to vertex 0
  indices[i] = has_nan ? 0u : id;
  if (r == i) {
    vertices[id] = make_float3(c.v[0], c.v[1], c.v[2]);
    normals[id] = make_float3(c.v[3], c.v[4], c.v[5]);
    texcoords[id] = make_float2(c.v[6], c.v[7]);
  }
}

}  // namespace

bool mesh_prep_available()
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return n > 0;
}

void prepare_mesh_gpu(const std::vector<float>& pos, const std::vector<float>& nrm, const std::vector<float>& tex,
                      const std::vector<ObjCorner>& corners, PreparedMesh& out)
{
  out = PreparedMesh();
  const uint32_t n = (uint32_t)corners.size(), n_faces = n / 3;
  if (n == 0) return;
  if (corners.size() % 3 != 0 || corners.size() > 0xfffffff0ull) throw std::runtime_error("mesh prep: bad corner count");
  cudaEvent_t e0, e1;
  FR_CUDA_CHECK(cudaEventCreate(&e0));
  FR_CUDA_CHECK(cudaEventCreate(&e1));
  DevBuf<float> d_pos, d_nrm, d_tex;
  DevBuf<ObjCorner> d_corners;
  d_pos.upload(pos);
  if (!nrm.empty()) d_nrm.upload(nrm);
  if (!tex.empty()) d_tex.upload(tex);
  d_corners.upload(corners);
  FR_CUDA_CHECK(cudaEventRecord(e0));
  DevBuf<CornerRecord> rec(n);
  DevBuf<uint32_t> error(1), slot_of(n), rep(n), is_rep(n), new_id(n);
  error.zero();
  const int B = 256;
  k_expand<<<(n_faces + B - 1) / B, B>>>(d_pos.get(), (uint32_t)pos.size(), d_nrm.get(), (uint32_t)nrm.size(), d_tex.get(),
                                         (uint32_t)tex.size(), d_corners.get(), n_faces, rec.get(), error.get());
  FR_CUDA_LAUNCH_CHECK();
  uint32_t table_size = 1024;
  while (table_size < 2ull * n) table_size <<= 1;
  DevBuf<uint32_t> table(table_size);
  FR_CUDA_CHECK(cudaMemset(table.get(), 0xff, sizeof(uint32_t) * table_size));
  k_insert<<<(n + B - 1) / B, B>>>(rec.get(), n, table.get(), table_size - 1, slot_of.get());
  FR_CUDA_LAUNCH_CHECK();
  k_resolve<<<(n + B - 1) / B, B>>>(table.get(), slot_of.get(), n, rep.get(), is_rep.get());
  FR_CUDA_LAUNCH_CHECK();
  size_t tmp_bytes = 0;
  FR_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, is_rep.get(), new_id.get(), (int)n));
  DevBuf<unsigned char> tmp(tmp_bytes);
  FR_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.get(), tmp_bytes, is_rep.get(), new_id.get(), (int)n));
  uint32_t h_err = 0, last_id = 0, last_flag = 0;
  FR_CUDA_CHECK(cudaMemcpy(&h_err, error.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (h_err) throw std::runtime_error("vertex index out of range");
  FR_CUDA_CHECK(cudaMemcpy(&last_id, new_id.get() + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
  FR_CUDA_CHECK(cudaMemcpy(&last_flag, is_rep.get() + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
  const uint32_t n_unique = last_id + last_flag;
  DevBuf<float3> d_v(n_unique), d_n(n_unique);
  DevBuf<float2> d_t(n_unique);
  DevBuf<uint32_t> d_idx(n);
  k_emit<<<(n + B - 1) / B, B>>>(rec.get(), rep.get(), new_id.get(), n, d_v.get(), d_n.get(), d_t.get(), d_idx.get());
  FR_CUDA_LAUNCH_CHECK();
  FR_CUDA_CHECK(cudaEventRecord(e1));
  out.vertices.resize(n_unique);
  out.normals.resize(n_unique);
  out.texcoords.resize(n_unique);
  out.indices.resize(n_faces);
  FR_CUDA_CHECK(cudaMemcpy(out.vertices.data(), d_v.get(), sizeof(float3) * n_unique, cudaMemcpyDeviceToHost));
  FR_CUDA_CHECK(cudaMemcpy(out.normals.data(), d_n.get(), sizeof(float3) * n_unique, cudaMemcpyDeviceToHost));
  FR_CUDA_CHECK(cudaMemcpy(out.texcoords.data(), d_t.get(), sizeof(float2) * n_unique, cudaMemcpyDeviceToHost));
  FR_CUDA_CHECK(cudaMemcpy(out.indices.data(), d_idx.get(), sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
  FR_CUDA_CHECK(cudaEventElapsedTime(&out.gpu_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

}  // namespace frd
