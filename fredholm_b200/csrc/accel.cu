// Two-level acceleration structure: per-mesh object-space trees + an instance tree (bvh_build.h TwoLevelBvh).
// Built from the same GPU builder as the flat tree: a BLAS is build_bvh() over one mesh's faces with the identity
// transform; the TLAS is build_bvh() over one PLACEHOLDER triangle per instance whose vertices span the
// instance's world box (so the builder's leaf boxes are exactly the instance boxes and its leaf "triangles" carry
// the instance index).  The traversal kernel (bvh.cuh, TWO = true) enters an instance where the flat kernel would
// test a triangle.
#include <algorithm>
#include <stdexcept>

#include "bvh_build.h"

namespace frd
{
namespace
{

// copies a BLAS into the combined array, making its child / triangle indices absolute
__global__ void k_relocate_nodes(const Node8* __restrict__ src, uint32_t n, uint32_t node_offset, uint32_t tri_offset,
                                 Node8* __restrict__ dst)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Node8 nd = src[i];
  nd.child_base += node_offset;
  nd.tri_base += tri_offset;
  dst[node_offset + i] = nd;
}

// world box of every instance (the 8 corners of its mesh's object-space box through the instance transform)
// as a placeholder triangle: v0 = lo, v1 = hi, v2 = a third corner -- its bounding box is the instance box
__global__ void k_instance_placeholders(const float* __restrict__ mesh_bounds, const fredholm::Matrix3x4* __restrict__ o2w,
                                        uint32_t n, float3* __restrict__ vertices, uint3* __restrict__ indices)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* b = mesh_bounds + 6ull * i;
  const fredholm::Matrix3x4 m = o2w[i];
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float x = (c & 1) ? b[3] : b[0], y = (c & 2) ? b[4] : b[1], z = (c & 4) ? b[5] : b[2];
    const float4 r[3] = {m.m[0], m.m[1], m.m[2]};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = r[a].x * x + r[a].y * y + r[a].z * z + r[a].w;
      lo[a] = fminf(lo[a], v);
      hi[a] = fmaxf(hi[a], v);
    }
  }
  // a few ulps of slack for the rounding of the corner transform and of the ray transform at traversal time
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pad = 4e-7f * fmaxf(fabsf(lo[a]), fabsf(hi[a])) + 1e-30f;
    lo[a] -= pad;
    hi[a] += pad;
  }
  vertices[3ull * i] = make_float3(lo[0], lo[1], lo[2]);
  vertices[3ull * i + 1] = make_float3(hi[0], hi[1], hi[2]);
  vertices[3ull * i + 2] = make_float3(lo[0], hi[1], lo[2]);
  indices[i] = make_uint3(3u * i, 3u * i + 1u, 3u * i + 2u);
}

}  // namespace

void update_tlas(cudaStream_t stream, const fredholm::Matrix3x4* d_o2w, TwoLevelBvh& out)
{
  const uint32_t n = out.n_instances;
  if (n == 0) throw std::runtime_error("two-level bvh: no instances");
  cudaEvent_t e0, e1;
  FR_CUDA_CHECK(cudaEventCreate(&e0));
  FR_CUDA_CHECK(cudaEventCreate(&e1));
  FR_CUDA_CHECK(cudaEventRecord(e0, stream));
  k_instance_placeholders<<<(n + 127) / 128, 128, 0, stream>>>(out.mesh_bounds.get(), d_o2w, n, out.placeholder_vertices.get(),
                                                               out.placeholder_indices.get());
  FR_CUDA_LAUNCH_CHECK();
  // radix tree: no merge rounds, a handful of launches -- this is the per-frame cost of an animated scene
  build_bvh(stream, out.placeholder_vertices.get(), out.placeholder_indices.get(), out.zeros.get(), nullptr, out.identity.get(), n,
            out.tlas, /*builder=*/0);
  if (out.tlas.n_nodes > out.tlas_node_capacity) throw std::runtime_error("two-level bvh: TLAS exceeds its reserved node range");
  FR_CUDA_CHECK(cudaMemcpyAsync(out.nodes.get(), out.tlas.nodes.get(), sizeof(Node8) * out.tlas.n_nodes, cudaMemcpyDeviceToDevice, stream));
  FR_CUDA_CHECK(cudaMemcpyAsync(out.tris.get(), out.tlas.tris.get(), sizeof(float4) * 3ull * n, cudaMemcpyDeviceToDevice, stream));
  FR_CUDA_CHECK(cudaEventRecord(e1, stream));
  FR_CUDA_CHECK(cudaEventSynchronize(e1));
  FR_CUDA_CHECK(cudaEventElapsedTime(&out.tlas_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  for (int a = 0; a < 3; ++a) {
    out.bounds_lo[a] = out.tlas.bounds_lo[a];
    out.bounds_hi[a] = out.tlas.bounds_hi[a];
  }
  out.depth = out.tlas.depth + out.blas_depth;
  // TLAS levels + what enter_instance parks (3 entries) + BLAS levels must fit the traversal stack
  if (out.depth + 3 + 2 > (uint32_t)(kSmemStack + kLocalStack)) throw std::runtime_error("two-level bvh: tree too deep for the traversal stack");
}

void build_two_level(cudaStream_t stream, const float3* d_vertices, const uint3* d_indices,
                     const std::vector<uint32_t>& submesh_offsets, const std::vector<uint32_t>& submesh_n_faces,
                     const std::vector<uint32_t>& mesh_of_submesh, const std::vector<uint32_t>& representative,
                     const std::vector<std::vector<uint32_t>>* face_flags_of_mesh,
                     const fredholm::Matrix3x4* d_o2w, TwoLevelBvh& out)
{
  const uint32_t n_inst = (uint32_t)submesh_offsets.size(), n_mesh = (uint32_t)representative.size();
  if (n_inst == 0 || n_mesh == 0) throw std::runtime_error("two-level bvh: empty scene");
  out.n_instances = n_inst;
  out.n_meshes = n_mesh;
  uint32_t max_faces = n_inst;
  for (uint32_t m = 0; m < n_mesh; ++m) max_faces = std::max(max_faces, submesh_n_faces[representative[m]]);
  out.zeros.reserve(max_faces);
  FR_CUDA_CHECK(cudaMemsetAsync(out.zeros.get(), 0, sizeof(uint32_t) * max_faces, stream));
  const fredholm::Matrix3x4 ident = fredholm::make_mat3x4(make_float4(1, 0, 0, 0), make_float4(0, 1, 0, 0), make_float4(0, 0, 1, 0));
  out.identity.upload(&ident, 1, stream);

  // ---- one object-space tree per distinct mesh ----
  std::vector<DeviceBvh> blas(n_mesh);
  DevBuf<uint32_t> d_flags;
  for (uint32_t m = 0; m < n_mesh; ++m) {
    const uint32_t s = representative[m], nf = submesh_n_faces[s];
    if (nf == 0) throw std::runtime_error("two-level bvh: sub-mesh without faces");
    const uint32_t* flags = nullptr;
    if (face_flags_of_mesh && !(*face_flags_of_mesh)[m].empty()) {
      d_flags.upload((*face_flags_of_mesh)[m], stream);
      flags = d_flags.get();
    }
    build_bvh(stream, d_vertices, d_indices + submesh_offsets[s], out.zeros.get(), flags, out.identity.get(), nf, blas[m]);
  }

  // ---- layout: [TLAS capacity][BLAS 0][BLAS 1] ... ----
  out.tlas_node_capacity = n_inst / 2 + 2;
  std::vector<uint32_t> node_off(n_mesh), tri_off(n_mesh);
  uint32_t nodes_total = out.tlas_node_capacity, tris_total = n_inst;
  out.blas_depth = 0;
  for (uint32_t m = 0; m < n_mesh; ++m) {
    node_off[m] = nodes_total;
    tri_off[m] = tris_total;
    nodes_total += blas[m].n_nodes;
    tris_total += blas[m].n_faces;
    out.blas_depth = std::max(out.blas_depth, blas[m].depth);
  }
  out.n_nodes = nodes_total;
  out.n_blas_faces = tris_total - n_inst;
  out.nodes.alloc(nodes_total);
  out.tris.alloc(3ull * tris_total);
  FR_CUDA_CHECK(cudaMemsetAsync(out.nodes.get(), 0, sizeof(Node8) * out.tlas_node_capacity, stream));
  for (uint32_t m = 0; m < n_mesh; ++m) {
    const uint32_t n = blas[m].n_nodes;
    k_relocate_nodes<<<(n + 255) / 256, 256, 0, stream>>>(blas[m].nodes.get(), n, node_off[m], tri_off[m], out.nodes.get());
    FR_CUDA_LAUNCH_CHECK();
    FR_CUDA_CHECK(cudaMemcpyAsync(out.tris.get() + 3ull * tri_off[m], blas[m].tris.get(), sizeof(float4) * 3ull * blas[m].n_faces,
                                  cudaMemcpyDeviceToDevice, stream));
  }

  // ---- instances ----
  std::vector<InstanceRecord> rec(n_inst);
  std::vector<float> bounds(6ull * n_inst);
  for (uint32_t i = 0; i < n_inst; ++i) {
    const uint32_t m = mesh_of_submesh[i];
    if (submesh_n_faces[i] != submesh_n_faces[representative[m]]) throw std::runtime_error("two-level bvh: instance / mesh face count mismatch");
    rec[i].blas_root = node_off[m];
    rec[i].face_offset = submesh_offsets[i];
    for (int a = 0; a < 3; ++a) {
      bounds[6ull * i + a] = blas[m].bounds_lo[a];
      bounds[6ull * i + 3 + a] = blas[m].bounds_hi[a];
    }
  }
  out.instances.upload(rec, stream);
  out.mesh_bounds.upload(bounds, stream);
  out.placeholder_vertices.reserve(3ull * n_inst);
  out.placeholder_indices.reserve(n_inst);
  FR_CUDA_CHECK(cudaStreamSynchronize(stream));  // blas[] and the host vectors go out of scope
  update_tlas(stream, d_o2w, out);
}

}  // namespace frd
